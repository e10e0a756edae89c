"""A/B experiments: run bench.py against an alternative build of the CUDA library.  usage: ab_bench.py LIB.so [bench args]"""
import os, runpy, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from openballbot_rl_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv.pop(1))
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bench.py"), run_name="__main__")
