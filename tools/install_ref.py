"""Installs the UNMODIFIED reference hot-path files into baseline/_ref/ (git-ignored; it ships to the GPU box with the
gpurun snapshot, like the built .so files) so that `bench.py --impl reference` and the `cpu_baseline` leg time the
reference itself and not a restatement.

The reference is a pure-Python mmcv project (mmcv-full / mmdet / mmpose are absent and there is no network), so
`pip install --target baseline/_ref /root/reference` cannot resolve; the files the DS-GCN path executes import unmodified
once oracle/ref_loader.py stubs the handful of mmcv names they touch.  This script copies exactly those
files, byte for byte, keeping their relative paths; nothing under baseline/_ref is tracked by git or edited.

    python tools/install_ref.py [--src /root/reference] [--dst baseline/_ref]
"""
import argparse
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FILES = [
    "pyskl/utils/graph.py",                      # Graph tables
    "pyskl/models/gcns/utils/init_func.py",
    "pyskl/models/gcns/utils/gcn.py",            # unit_gcn, dgphgcn1 (a7-a9, a14)
    "pyskl/models/gcns/utils/tcn.py",            # unit_tcn, mstcn, dgmstcn (a10-a13)
    "pyskl/models/gcns/dgstgcn.py",              # DGBlock, DGSTGCN (a4-a6)
    "pyskl/models/gcns/stgcn.py",                # STGCN / STGCNBlock (config 5)
    "pyskl/models/gcns/utils/msg3d_utils.py",    # MSTCN (config 5, CTR-GCN's temporal unit)
    "pyskl/models/gcns/ctrgcn.py",               # CTRGCN / CTRGCNBlock (config 5)
    "LICENSE",
]


def install(src="/root/reference", dst=None, quiet=False):
    dst = dst or os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(src, FILES[4])):
        if not quiet:
            print(f"install_ref: {src} has no reference tree; nothing installed", file=sys.stderr)
        return None
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.isfile(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
    with open(os.path.join(dst, "README"), "w") as fh:
        fh.write("Unmodified copies of the reference files the DS-GCN path executes (tools/install_ref.py).\n"
                 "Not tracked by git; loaded by oracle/ref_loader.py for bench.py's reference arm.\n")
    return dst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=None)
    a = ap.parse_args()
    out = install(a.src, a.dst)
    print(out or "not installed")
