"""ctypes binding of libdsgcn_b200.so (include/dsgcn_b200.h).

The product path loads exactly one library: the in-tree CUDA build
`ds-gcn_b200/libdsgcn_b200.so` (sm_100a).  There is no CPU fallback: if the library
is missing, or a tensor is not on a CUDA device, the call raises.  The CPU
test-suite can point the binding at the host-side simulator build of the same
sources (tests/emu) through `_testing_use_library`; nothing in the package does.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdsgcn_b200.so")

F32, BF16 = 0, 1
_DT = {torch.float32: F32, torch.bfloat16: BF16}

c_ll, c_int, c_f, c_d, vp = C.c_longlong, C.c_int, C.c_float, C.c_double, C.c_void_p


class ActSrc(C.Structure):
    _fields_ = [("x1", vp), ("x2", vp), ("a1", vp), ("b1", vp), ("a2", vp), ("b2", vp),
                ("ld1", c_ll), ("ld2", c_ll), ("relu", c_int), ("pad_", c_int)]


class ConvGemmArgs(C.Structure):
    _fields_ = [("src", ActSrc), ("dtype", c_int), ("K", c_int), ("N", c_int), ("W", vp),
                ("ws_n", c_ll), ("ws_k", c_ll), ("ws_tap", c_ll), ("bias", vp),
                ("taps", c_int), ("tap_step", c_int), ("tap_off", c_int), ("t_mul", c_int), ("t_div", c_int),
                ("n_samples", c_int), ("T_in", c_int), ("T_out", c_int), ("Vin", c_int),
                ("ext_in", c_int), ("contract_ext", c_int),
                ("out", vp), ("ld_out", c_ll), ("add", vp), ("ld_add", c_ll), ("add2", vp), ("ld_add2", c_ll),
                ("bcast", vp), ("bcast_scale", c_f), ("has_mask", c_int), ("mask", ActSrc),
                ("stat_sum", vp), ("stat_sq", vp), ("partner", vp), ("ld_partner", c_ll), ("wpack", vp),
                ("out_f32", c_int), ("pad2_", c_int), ("adyn", vp), ("y_out", vp), ("ld_y", c_ll)]


class ConvWgradArgs(C.Structure):
    _fields_ = [("A", ActSrc), ("B", ActSrc), ("dtype", c_int), ("K", c_int), ("N", c_int), ("dW", vp),
                ("ws_n", c_ll), ("ws_k", c_ll), ("ws_tap", c_ll), ("db", vp),
                ("taps", c_int), ("tap_step", c_int), ("tap_off", c_int), ("t_mul", c_int), ("t_div", c_int),
                ("n_samples", c_int), ("T_in", c_int), ("T_out", c_int), ("Vin", c_int), ("ext_in", c_int)]


class BnJob(C.Structure):
    _fields_ = [("mode", c_int), ("C", c_int), ("sum", vp), ("sq", vp), ("count", c_d),
                ("gamma", vp), ("beta", vp), ("running_mean", vp), ("running_var", vp),
                ("save_mean", vp), ("save_invstd", vp), ("a", vp), ("b", vp), ("c", vp),
                ("dgamma", vp), ("dbeta", vp), ("momentum", c_f), ("eps", c_f)]


class TopologyArgs(C.Structure):
    _fields_ = [("H", vp), ("ld_h", c_ll), ("n_samples", c_int), ("V", c_int), ("R", c_int),
                ("node_type", vp), ("edge_type", vp), ("A", vp), ("alpha", vp), ("beta", vp),
                ("We", vp), ("be", vp), ("adyn", vp), ("adyn_dtype", c_int), ("S", vp),
                ("dadyn", vp), ("dH", vp), ("dA", vp), ("dalpha", vp), ("dbeta", vp), ("dWe", vp), ("dbe", vp),
                ("dH_bf16", vp), ("variant", c_int), ("subset_wise", c_int)]


class CtrTopologyArgs(C.Structure):
    _fields_ = [("H", vp), ("ld_h", c_ll), ("n_samples", c_int), ("V", c_int), ("R", c_int), ("C", c_int),
                ("A", vp), ("alpha", vp), ("W4", vp), ("b4", vp), ("adyn", vp), ("adyn_dtype", c_int),
                ("dadyn", vp), ("dH", vp), ("dH_bf16", vp), ("dA", vp), ("dalpha", vp), ("dW4", vp), ("db4", vp)]


class GraphAggArgs(C.Structure):
    _fields_ = [("src", ActSrc), ("dtype", c_int), ("mode", c_int),
                ("n_samples", c_int), ("T", c_int), ("V", c_int), ("KC", c_int), ("Ksub", c_int),
                ("adyn", vp), ("A", vp), ("out", vp), ("ld_out", c_ll), ("has_mask", c_int), ("mask", ActSrc),
                ("stat_sum", vp), ("stat_sq", vp), ("partner", vp), ("ld_partner", c_ll)]


class GraphAggDadjArgs(C.Structure):
    _fields_ = [("p", ActSrc), ("dy", ActSrc), ("dtype", c_int), ("is_static", c_int),
                ("n_samples", c_int), ("T", c_int), ("V", c_int), ("KC", c_int), ("Ksub", c_int), ("dadj", vp)]


class MsCombineArgs(C.Structure):
    _fields_ = [("dtype", c_int), ("n_samples", c_int), ("T_in", c_int), ("T_out", c_int), ("stride", c_int),
                ("V", c_int), ("has_ext", c_int), ("C", c_int),
                ("conv_lo", c_int), ("conv_hi", c_int), ("max_lo", c_int), ("max_hi", c_int),
                ("pass_lo", c_int), ("pass_hi", c_int),
                ("b", ActSrc), ("o", vp), ("ld_o", c_ll), ("add_coeff", vp), ("feat", vp), ("ld_feat", c_ll),
                ("oglob", vp), ("stat_sum", vp), ("stat_sq", vp),
                ("dfeat", ActSrc), ("d_o", vp), ("ld_do", c_ll), ("e", vp), ("ld_e", c_ll),
                ("b_raw", vp), ("ld_b", c_ll), ("e_sum", vp), ("e_sq", vp), ("dadd_coeff", vp),
                ("d_o_full", c_int), ("pad_", c_int)]


class MsBranch(C.Structure):
    _fields_ = [("kind", c_int), ("lo", c_int), ("hi", c_int), ("dilation", c_int), ("W", vp), ("bias", vp), ("dW", vp), ("db", vp)]


class MsTemporalArgs(C.Structure):
    _fields_ = [("n_samples", c_int), ("T_in", c_int), ("T_out", c_int), ("stride", c_int), ("V", c_int), ("has_ext", c_int),
                ("C", c_int), ("n_branches", c_int), ("br", MsBranch * 8), ("b", ActSrc), ("add_coeff", vp),
                ("feat", vp), ("ld_feat", c_ll), ("oglob", vp), ("stat_sum", vp), ("stat_sq", vp),
                ("dfeat", ActSrc), ("e", vp), ("ld_e", c_ll), ("e_sum", vp), ("e_sq", vp), ("dadd_coeff", vp), ("wpack", vp)]


class MsConvArgs(C.Structure):
    _fields_ = [("n_samples", c_int), ("T_in", c_int), ("T_out", c_int), ("stride", c_int), ("Vr", c_int), ("transposed", c_int),
                ("n_branches", c_int), ("br", MsBranch * 8), ("src", vp), ("ld_src", c_ll), ("out", vp), ("ld_out", c_ll),
                ("has_mask", c_int), ("mask", ActSrc), ("partner", vp), ("ld_partner", c_ll), ("stat_sum", vp), ("stat_sq", vp),
                ("wpack", vp)]


class PointwiseArgs(C.Structure):
    _fields_ = [("src", ActSrc), ("dtype", c_int), ("C", c_int), ("rows", c_ll), ("out", vp), ("ld_out", c_ll),
                ("out_dtype", c_int), ("has_mask", c_int), ("mask", ActSrc),
                ("stat_sum", vp), ("stat_sq", vp), ("partner", vp), ("ld_partner", c_ll),
                ("partner_dtype", c_int), ("pad_", c_int)]


EXPORTS = {
    # name: (restype, argtypes)
    "dsg_conv_gemm": (c_int, [C.POINTER(ConvGemmArgs), vp]),
    "dsg_conv_gemm_wpack_bytes": (c_ll, [c_int, c_int]),
    "dsg_conv_wgrad": (c_int, [C.POINTER(ConvWgradArgs), vp]),
    "dsg_bn_finalize": (c_int, [C.POINTER(BnJob), c_int, vp]),
    "dsg_tmean": (c_int, [vp, c_int, c_ll, c_int, c_int, c_int, c_int, vp, vp]),
    "dsg_tmean2": (c_int, [vp, c_int, c_ll, c_int, c_int, c_int, c_int, vp, vp, vp]),
    "dsg_topology_fwd": (c_int, [C.POINTER(TopologyArgs), vp]),
    "dsg_topology_bwd": (c_int, [C.POINTER(TopologyArgs), vp]),
    "dsg_head_ce_fwd": (c_int, [vp, vp, vp, vp, c_int, c_int, c_int, vp, vp, vp]),
    "dsg_head_ce_bwd": (c_int, [vp, vp, vp, vp, vp, c_int, c_int, c_int, vp, vp, vp, vp, vp]),
    "dsg_ctr_topology_fwd": (c_int, [C.POINTER(CtrTopologyArgs), vp]),
    "dsg_ctr_topology_bwd": (c_int, [C.POINTER(CtrTopologyArgs), vp]),
    "dsg_graph_agg": (c_int, [C.POINTER(GraphAggArgs), vp]),
    "dsg_graph_agg_dadj": (c_int, [C.POINTER(GraphAggDadjArgs), vp]),
    "dsg_ms_combine_fwd": (c_int, [C.POINTER(MsCombineArgs), vp]),
    "dsg_ms_combine_bwd": (c_int, [C.POINTER(MsCombineArgs), vp]),
    "dsg_ms_combine_bwd_part": (c_int, [C.POINTER(MsCombineArgs), c_int, vp]),
    "dsg_pointwise": (c_int, [C.POINTER(PointwiseArgs), vp]),
    "dsg_ms_temporal_supported": (c_int, [C.POINTER(MsTemporalArgs)]),
    "dsg_ms_temporal_wpack_bytes": (c_ll, [C.POINTER(MsTemporalArgs)]),
    "dsg_ms_temporal_fwd": (c_int, [C.POINTER(MsTemporalArgs), vp]),
    "dsg_ms_temporal_bwd_data": (c_int, [C.POINTER(MsTemporalArgs), vp]),
    "dsg_ms_temporal_bwd_weight": (c_int, [C.POINTER(MsTemporalArgs), vp]),
    "dsg_ms_conv_wpack_bytes": (c_ll, [C.POINTER(MsConvArgs)]),
    "dsg_ms_conv": (c_int, [C.POINTER(MsConvArgs), C.POINTER(c_int), vp]),
    "dsg_ms_conv_wgrad": (c_int, [C.POINTER(MsConvArgs), C.POINTER(c_int), vp]),
    "dsg_sgd_step": (c_int, [vp, vp, vp, c_ll, c_f, c_f, c_f, c_int, c_f, vp]),
    "dsg_sgd_step_dev": (c_int, [vp, vp, vp, c_ll, vp, c_f, c_f, c_int, c_f, vp]),
    "dsg_debug_counter": (c_ll, [c_int]),
    "dsg_last_error": (C.c_char_p, []),
    "dsg_abi_version": (c_int, []),
    "dsg_is_device_build": (c_int, []),
}

ABI_VERSION = 1
_lib = None
_is_device = True
launch_count = 0      # kernels-launching ABI calls made through this binding (bench.py reports it)


class DsgError(RuntimeError):
    pass


def _bind(path):
    lib = C.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.dsg_abi_version() != ABI_VERSION:
        raise DsgError(f"{path}: ABI version {lib.dsg_abi_version()} != {ABI_VERSION}")
    return lib


def lib():
    """The CUDA library, loaded on first use.  Raises if it has not been built."""
    global _lib, _is_device
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DsgError(f"{LIB_PATH} not found: run `python build.py` (nvcc, sm_100a). "
                           "There is no CPU fallback for the DS-GCN kernels.")
        _lib = _bind(LIB_PATH)
        _is_device = bool(_lib.dsg_is_device_build())
    return _lib


def _testing_use_library(path):
    """CPU test-suite hook: bind the host-side simulator build of the kernel sources (tests/emu)."""
    global _lib, _is_device
    _lib = _bind(path) if path else None
    _is_device = bool(_lib.dsg_is_device_build()) if _lib else True
    return _lib


def is_device_build():
    lib()
    return _is_device


def check_tensor(t):
    lib()
    if _is_device and not t.is_cuda:
        raise DsgError("DS-GCN kernels need CUDA tensors (no CPU fallback); got a CPU tensor")
    if not _is_device and t.is_cuda:
        raise DsgError("simulator build bound but a CUDA tensor was passed")


def ptr(t):
    if t is None:
        return None
    check_tensor(t)
    return t.data_ptr()


def stream():
    if not _is_device:
        return None
    return torch.cuda.current_stream().cuda_stream


def dt(t_or_dtype):
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    try:
        return _DT[d]
    except KeyError:
        raise DsgError(f"unsupported activation dtype {d} (fp32 or bf16)")


# ---- side stream: independent work (weight gradients) overlaps the data-gradient chain -------------------------
side_enabled = os.environ.get("DSG_SIDE_STREAM", "1") != "0"
_side_streams = {}
_side_used = False
keepalive = []        # tensors read by side-stream kernels stay referenced until join_side()


class side_stream:
    """`with side_stream():` launches the enclosed ABI calls on a second CUDA stream that first waits for everything
    already queued on the current stream.  join_side() makes the current stream wait for it again.  Works under CUDA
    graph capture (fork/join become graph edges)."""

    def __enter__(self):
        global _side_used
        self.active = side_enabled and _is_device and torch.cuda.is_available()
        if self.active:
            main = torch.cuda.current_stream()
            key = main.device.index
            if key not in _side_streams:
                _side_streams[key] = torch.cuda.Stream(device=main.device)
            s = _side_streams[key]
            s.wait_stream(main)
            self.cm = torch.cuda.stream(s)
            self.cm.__enter__()
            _side_used = True
        return self

    def __exit__(self, *exc):
        if self.active:
            self.cm.__exit__(*exc)
        return False


def join_side():
    global _side_used
    if _side_used:
        main = torch.cuda.current_stream()
        s = _side_streams.get(main.device.index)
        if s is not None:
            main.wait_stream(s)
        _side_used = False
    keepalive.clear()


profile = None        # when a list: (name, algorithmic_bytes, start_event, end_event) per ABI call (bench.py roofline leg)


def call(name, *args, nbytes=0, tag=None):
    global launch_count
    if profile is not None and _is_device:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib(), name)(*args)
        e1.record()
        profile.append((name, nbytes, e0, e1, tag))
    else:
        rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise DsgError(lib().dsg_last_error().decode())
    launch_count += 1
