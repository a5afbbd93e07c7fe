#!/usr/bin/env python
"""Builds tuning variants of libvoxgraph_b200.so into voxgraph_b200/variants/ (git-ignored, they
travel to the GPU box); scripts/tune_reg.sh times each one there."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voxgraph_b200 import build as b  # noqa: E402

VARIANTS = {
    "all_on_b5": ["-DVGX_REG_THREADS=128", "-DVGX_REG_MIN_BLOCKS=5"],
    "all_on_b6": ["-DVGX_REG_THREADS=128", "-DVGX_REG_MIN_BLOCKS=6"],
    "all_on_b4": ["-DVGX_REG_THREADS=128", "-DVGX_REG_MIN_BLOCKS=4"],
    "no_ldg256": ["-DVGX_REG_THREADS=128", "-DVGX_REG_MIN_BLOCKS=5", "-DVGX_REG_LDG256=0"],
    "no_earlyout": ["-DVGX_REG_THREADS=128", "-DVGX_REG_MIN_BLOCKS=5", "-DVGX_REG_EARLYOUT=0"],
    "no_skipgram": ["-DVGX_REG_THREADS=128", "-DVGX_REG_MIN_BLOCKS=5", "-DVGX_REG_SKIPGRAM=0"],
    "all_off": ["-DVGX_REG_THREADS=128", "-DVGX_REG_MIN_BLOCKS=5", "-DVGX_REG_LDG256=0", "-DVGX_REG_EARLYOUT=0", "-DVGX_REG_SKIPGRAM=0"],
    "all_on_t256_b2": ["-DVGX_REG_THREADS=256", "-DVGX_REG_MIN_BLOCKS=2"],
    "all_on_t256_b3": ["-DVGX_REG_THREADS=256", "-DVGX_REG_MIN_BLOCKS=3"],
}

if __name__ == "__main__":
    out_dir = os.path.join(ROOT, "voxgraph_b200", "variants")
    os.makedirs(out_dir, exist_ok=True)
    names = sys.argv[1:] or list(VARIANTS)
    for name in names:
        out = os.path.join(out_dir, "libvgx_%s.so" % name)
        b.build(force=True, defines=VARIANTS[name], out=out)
        print("built", out)
