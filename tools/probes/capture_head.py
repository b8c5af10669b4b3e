"""Which variant of the head path survives CUDA-graph capture of a whole training iteration (train.GraphedTrainStep)."""
import sys, os, itertools
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dsgcn_b200
from dsgcn_b200 import parallel, train as TR, recognizer as R
NS = dict(gcn_type="dgphgcn1", gcn_ratio=0.5, gcn_node_attention=True, gcn_edge_attention=True, gcn_decompose=True, gcn_subset_wise=True,
          gcn_ctr="T", gcn_ada="T", tcn_type="dgmstcn", base_channels=16,
          graph_cfg=dict(layout="nturgb+d", mode="random", num_filter=3, init_off=.04, init_std=.02),
          tcn_ms_cfg=[(3, 1), (3, 2), (3, 3), (3, 4), ("max", 3), "1x1"])
dev = torch.device("cuda:0")
for fused, direct in ((False, True), (True, False), (True, True)):
    torch.manual_seed(0); np.random.seed(0)
    m = dsgcn_b200.RecognizerGCN(backbone=dict(type="DGSTGCN", **NS), cls_head=dict(type="GCNHead", num_classes=10, in_channels=64)).to(dev).train()
    m.cls_head.use_fused = fused
    R._HeadCEFn.direct_sink = direct
    opt = parallel.FlatSGD(parallel.GradBuckets(m, n_buckets=3), lr=0.01, momentum=0.9, weight_decay=5e-4, nesterov=True)
    x = torch.randn(4, 1, 2, 16, 25, 3, device=dev); y = torch.randint(0, 10, (4, 1), device=dev)
    try:
        st = TR.GraphedTrainStep(m, opt, x, y, warmup=2)
        out = st(x, y)
        print(f"fused={fused} direct={direct}: OK loss {out['log_vars']['loss']:.4f}")
        st.release()
    except Exception as e:
        print(f"fused={fused} direct={direct}: FAILED {type(e).__name__}: {str(e)[:120]}")
        torch.cuda.synchronize()
