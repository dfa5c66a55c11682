"""ctypes binding of libfs2b200.so -- mirrors include/fs2_b200.h one to one."""
import ctypes as C
import os

from .build import LIB_PATH

FS2_OK = 0
MATH_TF32, MATH_BF16, MATH_TF32X3 = 0, 1, 2

c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_f32p = C.POINTER(C.c_float)
c_u8p = C.POINTER(C.c_uint8)


class Config(C.Structure):
    _fields_ = [("n_src_vocab", C.c_int32), ("n_speaker", C.c_int32), ("n_emotion", C.c_int32),
                ("n_arousal", C.c_int32), ("n_valence", C.c_int32), ("max_seq_len", C.c_int32),
                ("math_mode", C.c_int32),
                ("pitch_frame_level", C.c_int32), ("energy_frame_level", C.c_int32)]


class Inputs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("max_src_len", C.c_int32),
                ("speakers", C.c_void_p), ("emotions", C.c_void_p), ("arousals", C.c_void_p),
                ("valences", C.c_void_p), ("texts", C.c_void_p), ("src_lens", C.c_void_p),
                ("p_targets", C.c_void_p), ("e_targets", C.c_void_p), ("d_targets", C.c_void_p),
                ("p_control", C.c_float), ("e_control", C.c_float), ("d_control", C.c_float),
                ("max_mel_len", C.c_int32)]


class Stage1Out(C.Structure):
    _fields_ = [("pitch", C.c_void_p), ("energy", C.c_void_p), ("log_d", C.c_void_p),
                ("d_rounded", C.c_void_p), ("src_mask", C.c_void_p), ("mel_lens", C.c_void_p),
                ("total_frames", C.c_int64), ("max_mel_len", C.c_int32)]


class Stage2IO(C.Structure):
    _fields_ = [("mel", C.c_void_p), ("postnet", C.c_void_p), ("mel_mask", C.c_void_p),
                ("pitch_frames", C.c_void_p), ("energy_frames", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/fs2_b200.h declares
SIGNATURES = {
    "fs2_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]),
    "fs2_destroy": (None, [C.c_void_p]),
    "fs2_last_error": (C.c_char_p, [C.c_void_p]),
    "fs2_version": (C.c_int, []),
    "fs2_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, c_i64p, C.c_int]),
    "fs2_prepare": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fs2_forward_stage1": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Inputs), C.POINTER(Stage1Out)]),
    "fs2_forward_stage2": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Stage2IO)]),
    "fs2_export_stage1": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fs2_import_stage1": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, c_i64p, C.POINTER(C.c_int32)]),
    "fs2_last_launch_count": (C.c_int, [C.c_void_p]),
    "fs2_set_eager_stage2": (C.c_int, [C.c_void_p, C.c_int]),
    "fs2_set_stage2_wait_event": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fs2_read_packed_postnet": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, c_i64p]),
    "fs2_debug_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "fs2_debug_fetch": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, c_i64p, c_i64p]),
    "fs2_debug_set_flag": (C.c_int, [C.c_int, C.c_int]),
    "fs2_debug_read_trace": (C.c_int, [c_i64p, C.c_int]),
    "fs2_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "fs2_profile_read": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "fs2_op_conv_gemm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "fs2_op_conv_gemm_ln": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fs2_op_ffn_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fs2_op_conv_gemm_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "fs2_op_conv_gemm_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "fs2_op_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p]),
    "fs2_op_attention_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "fs2_op_durations": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "fs2_op_bucketize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    "fs2_op_frame_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fs2_voc_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "fs2_voc_destroy": (None, [C.c_void_p]),
    "fs2_voc_last_error": (C.c_char_p, [C.c_void_p]),
    "fs2_voc_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, c_i64p, C.c_int]),
    "fs2_voc_prepare": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fs2_voc_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p]),
    "fs2_voc_last_launch_count": (C.c_int, [C.c_void_p]),
}

_LIB = None


def load_library(path=None):
    """dlopen libfs2b200.so and bind every symbol.  Raises if the library is missing:
    the product has no other implementation to fall back to."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc); fs2_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def check(lib, ctx, code):
    if code != FS2_OK:
        msg = lib.fs2_last_error(ctx)
        raise RuntimeError(f"libfs2b200 error {code}: {msg.decode() if msg else '?'}")
