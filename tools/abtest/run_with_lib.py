#!/usr/bin/env python3
"""A/B helper: run tools/profile_run.py against another build of the library (older builds may lack newer exports).
usage: run_with_lib.py <lib.so> [profile_run args...]"""
import ctypes as C, os, runpy, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from r2p2_b200 import _lib
L = C.CDLL(os.path.abspath(sys.argv[1]))
for name, (res, args) in _lib._SIGS.items():
    try:
        fn = getattr(L, name)
    except AttributeError:
        continue
    fn.restype, fn.argtypes = res, args
_lib._lib = L
sys.argv = [os.path.join(ROOT, "tools", "profile_run.py")] + sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
