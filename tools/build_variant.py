"""Builds a variant of the library with extra nvcc flags into build/variants/<name>/lib.so (A/B experiments, trace builds):
    SOCCDPT_NVCC_FLAGS="-DSOCCDPT_..." python tools/build_variant.py <name>
and select it in-process with SOCCDPT_LIB=build/variants/<name>/lib.so."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccdpt_b200 import build as B

name = sys.argv[1]
out = os.path.join(ROOT, "build", "variants", name)
os.makedirs(out, exist_ok=True)
B.OBJ = os.path.join(out, "obj")
B.LIB_DIR = out
B.LIB = os.path.join(out, "lib.so")
print(B.build(force=True))
