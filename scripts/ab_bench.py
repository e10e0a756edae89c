"""A/B experiments: run bench.py against an alternative build of the CUDA library.  usage: ab_bench.py LIB.so [bench args]"""
import os, runpy, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from openballbot_rl_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv.pop(1))
_lib.needs_build = lambda: False           # the alternative build is used as it is
_lib.torch_ops = lambda: (_ for _ in ()).throw(_lib.EngineError('A/B run: ctypes binding only (the ops library links against the in-tree build)'))
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bench.py"), run_name="__main__")
