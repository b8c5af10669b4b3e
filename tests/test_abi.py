"""The C-ABI library builds for sm_100a, loads on a GPU-less host, and exports every symbol include/dsgcn_b200.h declares
(no compute calls here).  Also: the ctypes mirror structs have the sizes the C compiler gives the header's structs."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dsgcn_b200.h")


def _declared():
    src = open(HEADER).read()
    names = re.findall(r"^(?:int|long long|const char\*)\s+(dsg_\w+)\s*\(", src, flags=re.M)
    assert len(names) >= 18
    return names


def _cuda_lib():
    sys.path.insert(0, ROOT)
    import build
    if not os.path.exists(build.OUT):
        if not os.path.exists(build.NVCC):
            pytest.skip("nvcc not available and the library is not built")
        build.build()
    return build.OUT


def test_cuda_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_cuda_lib())
    for name in _declared():
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    lib.dsg_abi_version.restype = ctypes.c_int
    assert lib.dsg_abi_version() == 1
    assert lib.dsg_is_device_build() == 1


def test_binding_matches_header():
    import dsgcn_b200
    from dsgcn_b200 import _lib as L
    assert set(L.EXPORTS) == set(_declared())
    # struct sizes as the C compiler sees them
    prog = r'''
#include <stdio.h>
#include "dsgcn_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(dsg_ctr_topology_args), sizeof(dsg_act_src), sizeof(dsg_conv_gemm_args), sizeof(dsg_conv_wgrad_args),
         sizeof(dsg_bn_job), sizeof(dsg_topology_args), sizeof(dsg_graph_agg_args), sizeof(dsg_graph_agg_dadj_args),
         sizeof(dsg_ms_combine_args), sizeof(dsg_pointwise_args), sizeof(dsg_ms_branch), sizeof(dsg_ms_temporal_args));
  return 0;
}'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    mirror = [L.CtrTopologyArgs, L.ActSrc, L.ConvGemmArgs, L.ConvWgradArgs, L.BnJob, L.TopologyArgs, L.GraphAggArgs, L.GraphAggDadjArgs, L.MsCombineArgs,
              L.PointwiseArgs, L.MsBranch, L.MsTemporalArgs]
    assert sizes == [ctypes.sizeof(m) for m in mirror]


def test_no_cpu_fallback_and_oracle_isolation():
    """The product package never imports oracle/ and refuses CPU tensors when bound to the CUDA library."""
    pkg = os.path.join(ROOT, "ds-gcn_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f
