"""Build tuning variants of the library into build/variants/ (occupancy knobs of the tile kernels)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coivo_b200 import _lib

VARIANTS = {
    "base": [],
    # backward ablations (timing only: these builds give WRONG results and are never loaded by the package)
    "abl_nored": ["COLVO_EXP_NORED=1"],
    "abl_nogather": ["COLVO_EXP_NOGATHER=1"],
    "abl_nobar": ["COLVO_EXP_NOBAR=1"],
}
if len(sys.argv) > 1:            # python scripts/build_variants.py name=DEF1,DEF2 ...
    VARIANTS = {"base": []}
    for a in sys.argv[1:]:
        nm, _, defs = a.partition("=")
        VARIANTS[nm] = [d for d in defs.split(",") if d]
out = os.path.join(os.path.dirname(_lib.PKG_DIR), "build", "variants")
os.makedirs(out, exist_ok=True)
procs = []
for name, defs in VARIANTS.items():
    cmd = _lib.nvcc_command(os.path.join(out, f"lib_{name}.so"), defs)
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    o, _ = p.communicate()
    print(name, "ok" if p.returncode == 0 else "FAILED\n" + o)
