"""Build experiment variants of libfwsim.so side by side (build_ab/lib_<name>.so) for A/B measurements on one GPU box:
    python scripts/ab_build.py name1:DEF1,DEF2 name2:DEF3 ...   then   FWSIM_LIB=build_ab/lib_name1.so python bench.py ..."""
import os, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyflyt_drone_b200 import build
os.makedirs("build_ab", exist_ok=True)
def one(spec):
    name, _, defs = spec.partition(":")
    out = os.path.abspath(f"build_ab/lib_{name}.so")
    build.build(force=True, out=out, defines=tuple(d for d in defs.split(",") if d))
    return out
with ThreadPoolExecutor(4) as ex:
    for o in ex.map(one, sys.argv[1:]):
        print(o)
