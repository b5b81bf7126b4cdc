"""Build tuning variants of the library into build/variants/ (occupancy knobs of the tile kernels)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coivo_b200 import _lib

VARIANTS = {
    "base": [],
    "sm32x16": ["COLVO_SM_BW=32"],
    "sm64x8": ["COLVO_SM_BH=8"],
    "sm128x8": ["COLVO_SM_BW=128", "COLVO_SM_BH=8"],
    "sm32x32": ["COLVO_SM_BW=32", "COLVO_SM_BH=32"],
}
out = os.path.join(os.path.dirname(_lib.PKG_DIR), "build", "variants")
os.makedirs(out, exist_ok=True)
procs = []
for name, defs in VARIANTS.items():
    cmd = _lib.nvcc_command(os.path.join(out, f"lib_{name}.so"), defs)
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    o, _ = p.communicate()
    print(name, "ok" if p.returncode == 0 else "FAILED\n" + o)
