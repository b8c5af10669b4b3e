"""TEST INFRASTRUCTURE — loads the *unmodified* reference hot-path files from
/root/reference so the oracle restatement (oracle/dsgcn_oracle.py) can be pinned
against them and golden vectors can be generated (tests/golden/make_golden.py).

Nothing in the product package imports this file.  /root/reference does not
exist on the GPU box; there the loader falls back to baseline/_ref/ — byte-identical
copies of the same files made by tools/install_ref.py (git-ignored, shipped with the
gpurun snapshot) — which is what bench.py's reference arm times.  `available()` says
whether either tree is present.

The reference imports mmcv (absent) and three junk modules (tkinter, turtle,
matplotlib: pyskl/models/gcns/utils/gcn.py:5-9).  We register tiny stand-ins in
sys.modules for exactly the names the five hot-path files touch
and load each file by path under a fake package `refpyskl`.
"""
import importlib.util
import os
import sys
import types

import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
_PROBE = "pyskl/models/gcns/dgstgcn.py"


def _find_root():
    cands = [os.environ.get("DSGCN_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, _PROBE)):
            return c
    return cands[0] or "/root/reference"


REF_ROOT = _find_root()
_PKG = "refpyskl"
_loaded = {}


def available():
    return os.path.isfile(os.path.join(REF_ROOT, _PROBE))


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave like a package so sub-imports resolve
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


class _Registry:
    def __init__(self, name="models", parent=None, **kw):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        return deco

    def build(self, cfg):
        cfg = dict(cfg)
        return self.module_dict[cfg.pop("type")](**cfg)


def _build_norm_layer(cfg, num_features, postfix=""):
    # mmcv-full 1.5.0 build_norm_layer(dict(type='BN'), C) -> ('bn', nn.BatchNorm2d(C, eps=1e-5))
    assert cfg["type"] in ("BN", "BN2d"), cfg
    return "bn" + str(postfix), nn.BatchNorm2d(num_features, eps=1e-5)


def _build_activation_layer(cfg):
    return {"ReLU": nn.ReLU, "Tanh": nn.Tanh, "Sigmoid": nn.Sigmoid}[cfg["type"]]()


def _install_stubs():
    _mod("tkinter", N=None)
    _mod("turtle", screensize=None)
    _mod("matplotlib")
    _mod("matplotlib.pyplot", axes=None, axis=None)
    reg = _Registry()
    _mod("mmcv")
    _mod("mmcv.cnn", build_norm_layer=_build_norm_layer, build_activation_layer=_build_activation_layer,
         MODELS=reg, normal_init=lambda m, mean=0, std=1, bias=0: (nn.init.normal_(m.weight, mean, std),
                                                                    nn.init.constant_(m.bias, bias)))
    _mod("mmcv.runner", load_checkpoint=lambda *a, **k: None)
    _mod("mmcv.utils", Registry=_Registry, _BatchNorm=nn.modules.batchnorm._BatchNorm)


def _load(modname, relpath):
    full = f"{_PKG}.{modname}"
    if full in sys.modules and getattr(sys.modules[full], "__file__", None):
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(REF_ROOT, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[full] = m
    spec.loader.exec_module(m)
    return m


def load():
    """Returns a namespace with the reference classes: Graph, unit_gcn, dgphgcn1,
    unit_tcn, mstcn, dgmstcn, DGBlock, DGSTGCN, STGCN (if importable)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    _install_stubs()
    for p in (_PKG, f"{_PKG}.models", f"{_PKG}.models.gcns", f"{_PKG}.models.gcns.utils", f"{_PKG}.utils"):
        _mod(p)
    graph = _load("utils.graph", "pyskl/utils/graph.py")
    _mod(f"{_PKG}.utils", Graph=graph.Graph, cache_checkpoint=lambda x: x, graph=graph)
    sys.modules[f"{_PKG}.utils.graph"] = graph
    _mod(f"{_PKG}.models.builder", BACKBONES=_Registry("backbones"))
    _load("models.gcns.utils.init_func", "pyskl/models/gcns/utils/init_func.py")
    gcn = _load("models.gcns.utils.gcn", "pyskl/models/gcns/utils/gcn.py")
    tcn = _load("models.gcns.utils.tcn", "pyskl/models/gcns/utils/tcn.py")
    names = {}
    for src in (gcn, tcn):
        for k in dir(src):
            v = getattr(src, k)
            if isinstance(v, type) and issubclass(v, nn.Module) and v.__module__ == src.__name__:
                names[k] = v
    _mod(f"{_PKG}.models.gcns.utils", **names)
    dg = _load("models.gcns.dgstgcn", "pyskl/models/gcns/dgstgcn.py")
    _loaded.update(Graph=graph.Graph, graph_module=graph, unit_gcn=gcn.unit_gcn, dgphgcn1=gcn.dgphgcn1,
                   dggcn=gcn.dggcn, unit_tcn=tcn.unit_tcn, mstcn=tcn.mstcn, dgmstcn=tcn.dgmstcn,
                   DGBlock=dg.DGBlock, DGSTGCN=dg.DGSTGCN)
    _loaded.update(dghgcn=gcn.dghgcn, dgphgcn=gcn.dgphgcn, unit_ctrgcn=gcn.unit_ctrgcn, CTRGC=gcn.CTRGC)
    try:   # config-5 extras: MSTCN (msg3d_utils.py:64-150) and the CTR-GCN backbone (ctrgcn.py)
        ms = _load("models.gcns.utils.msg3d_utils", "pyskl/models/gcns/utils/msg3d_utils.py")
        _mod(f"{_PKG}.models.gcns.utils", MSTCN=ms.MSTCN)
        _loaded.update(MSTCN=ms.MSTCN)
        ct = _load("models.gcns.ctrgcn", "pyskl/models/gcns/ctrgcn.py")
        _loaded.update(CTRGCN=ct.CTRGCN, CTRGCNBlock=ct.CTRGCNBlock)
    except Exception as e:  # pragma: no cover
        _loaded.update(MSTCN=None, CTRGCN=None, CTRGCNBlock=None, ctrgcn_error=repr(e))
    try:
        st = _load("models.gcns.stgcn", "pyskl/models/gcns/stgcn.py")
        _loaded.update(STGCN=st.STGCN, STGCNBlock=st.STGCNBlock)
    except Exception as e:  # pragma: no cover - config-5 extras are optional
        _loaded.update(STGCN=None, STGCNBlock=None, stgcn_error=repr(e))
    return types.SimpleNamespace(**_loaded)


NORTH_STAR_BACKBONE = dict(
    # configs/dsstgcn/DSSTGCN_model.py:4-33
    gcn_type="dgphgcn1", gcn_ratio=0.125, gcn_node_attention=True, gcn_edge_attention=True,
    gcn_decompose=True, gcn_subset_wise=True, gcn_ctr="T", gcn_ada="T", tcn_type="dgmstcn",
    graph_cfg=dict(layout="nturgb+d", mode="random", num_filter=3, init_off=.04, init_std=.02),
    tcn_ms_cfg=[(3, 1), (3, 2), (3, 3), (3, 4), ("max", 3), "1x1"],
)
