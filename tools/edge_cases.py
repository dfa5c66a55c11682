import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np, fs2_b200
from gpu_util import model_for, run, DEV
syn=fs2_b200.synthetic
sd=syn.synthetic_state_dict(0)
for mode in ("tf32","bf16"):
    m=model_for(sd, math_mode=mode)
    for lens in ([1],[1,1,1],[2,120,1],[5,0,7]):
        b=syn.make_batch([max(l,1) for l in lens], seed=3)
        b["src_lens"]=torch.tensor(lens)
        out=run(m,b)
        print(mode, lens, "mel", tuple(out[0].shape), "mel_lens", out[9].tolist(), "finite", bool(torch.isfinite(out[1]).all()))
    # all-zero durations: d_control tiny
    b=syn.make_batch([6,9], seed=4)
    out=run(m,b,d_control=0.01)
    print(mode, "d_control 0.01 -> mel", tuple(out[0].shape), out[9].tolist())
    # long single utterance T>2000
    b=syn.make_batch([400], seed=5)
    out=run(m,b,d_control=2.0)
    print(mode, "long", tuple(out[0].shape), bool(torch.isfinite(out[1]).all()))
voc=fs2_b200.HiFiGANGeneratorB200(); voc.load_state_dict(syn.synthetic_vocoder_state_dict(0)); voc=voc.to(DEV)
for B,T,lens in ((1,1,None),(2,3,[3,0]),(1,2000,None),(5,7,[1,7,2,7,3])):
    mel=torch.randn(B,80,T,device=DEV)
    w=voc(mel, mel_lens=None if lens is None else torch.tensor(lens))
    print("voc",B,T,lens,tuple(w.shape),bool(torch.isfinite(w).all()))
