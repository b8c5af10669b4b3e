"""B200-native DS-GCN backbone (drop-in for the DS-STGCN path of davelailai/DS-GCN).

Import as `import dsgcn_b200` (repo-root alias; the directory name `ds-gcn_b200` is not a
Python identifier).  Host side: Python/PyTorch modules mirroring the reference API.
Device side: hand-written sm_100a CUDA kernels behind a C ABI (include/dsgcn_b200.h).
"""
from . import _lib, ops  # noqa: F401
from . import functional, graph, modules  # noqa: F401,E402
from .graph import Graph  # noqa: F401,E402
from .modules import (CTRGC, CTRGCN, CTRGCNBlock, DGBlock, DGSTGCN, MSTCN, STGCN, STGCNBlock, dggcn, dghgcn, dgmstcn,  # noqa: F401,E402
                      dgphgcn, dgphgcn1, get_compute_dtype, mstcn, set_compute_dtype, unit_ctrgcn, unit_gcn, unit_tcn)
from . import recognizer  # noqa: F401,E402
from .recognizer import (BACKBONES, HEADS, LOSSES, MODELS, RECOGNIZERS, CrossEntropyLoss, GCNHead, RecognizerGCN,  # noqa: F401,E402
                         build_backbone, build_head, build_loss, build_model, build_recognizer)
from . import parallel  # noqa: F401,E402
from . import pipeline, train  # noqa: F401,E402
