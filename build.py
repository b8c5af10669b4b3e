"""Builds libdsgcn_b200.so (sm_100a) in-tree.  `python build.py` or __graft_entry__.build()."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "ds-gcn_b200", "csrc")
OUT = os.path.join(ROOT, "ds-gcn_b200", "libdsgcn_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    return srcs + [os.path.join(ROOT, "include", "dsgcn_b200.h")]


def build(force=False, verbose=False):
    srcs = sources()
    if not force and not _stale(OUT, srcs):
        return OUT
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
           "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           os.path.join(CSRC, "api.cu"), "-o", OUT]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
