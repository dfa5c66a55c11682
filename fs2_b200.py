"""Import alias: `import fs2_b200` loads the package that lives in the
directory `expressive-fastspeech2-mandarin_b200/` (a hyphenated name is not a
Python identifier, so it is registered under this importable name)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "expressive-fastspeech2-mandarin_b200")
_spec = importlib.util.spec_from_file_location(
    "fs2_b200", os.path.join(_PKG_DIR, "__init__.py"),
    submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fs2_b200"] = _mod
_spec.loader.exec_module(_mod)
