"""TEST INFRASTRUCTURE — builds tests/emu/libdsgcn_emu.so: the library's kernel sources compiled by g++
against the host-side SIMT simulator (emu_runtime.h).  Loaded only by the CPU test-suite."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "ds-gcn_b200", "csrc")
OUT = os.path.join(HERE, "libdsgcn_emu.so")


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("emu_runtime.h", "emu_runtime.cc")]
    srcs.append(os.path.join(ROOT, "include", "dsgcn_b200.h"))
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in srcs):
        return OUT
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-shared", "-fPIC", "-DDSG_EMU", "-Wno-unknown-pragmas", "-Wno-attributes",
           "-I", HERE, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-x", "c++", os.path.join(CSRC, "api.cu"), os.path.join(HERE, "emu_runtime.cc"), "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
