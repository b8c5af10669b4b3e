"""Benchmark of the DS-GCN hot path (BASELINE.json): DS-GCN clips/sec, M=2, T=100, V=25, C=3.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mode train|fwd] [--batch B] [--impl ours|reference]

One "step" = one pass of the path over one synthetic batch: `train` = forward + loss + backward + SGD step of
RecognizerGCN(DGSTGCN north-star config + GCNHead) on `--batch` clips per GPU (default 128 = videos_per_gpu of
configs/_init_/lr_schedual.py:1-8, BASELINE configs[1]); `fwd` = eval forward.  N>1: one process per GPU under
torchrun, batch sharded (weak scaling), gradient all-reduce over NCCL, per-rank BatchNorm statistics like the
reference's DDP (broadcast_buffers=False, no SyncBN).

Printed JSON line: see DESIGN.md "Measurement".  `value` = device-resident inputs, CUDA-graph replay;
`e2e` = the public API (RecognizerGCN.train_step / forward) with host buffers: H2D of the batch from pinned memory and
D2H of the loss inside the timed region.  `--impl reference` times the UNMODIFIED reference (imported by path from
/root/reference, or from baseline/_ref = tools/install_ref.py copies on the GPU box) on the host cores at N=16 clips per
step, all host threads; `cpu_baseline` in the product line is the same thing on a shorter sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NORTH_STAR = dict(gcn_type="dgphgcn1", gcn_ratio=0.125, gcn_node_attention=True, gcn_edge_attention=True, gcn_decompose=True,
                  gcn_subset_wise=True, gcn_ctr="T", gcn_ada="T", tcn_type="dgmstcn",
                  graph_cfg=dict(layout="nturgb+d", mode="random", num_filter=3, init_off=.04, init_std=.02),
                  tcn_ms_cfg=[(3, 1), (3, 2), (3, 3), (3, 4), ("max", 3), "1x1"])
M_, T_, V_, C_ = 2, 100, 25, 3
NUM_CLASSES = 60
SGD = dict(lr=0.1, momentum=0.9, weight_decay=5e-4, nesterov=True)     # configs/_init_/lr_schedual.py:11
# stage-granular algorithmic elements per clip, forward; train = 3x (read x, read dy, write dx)
ELEMS_PER_CLIP_FWD = 16.655e6
LAUNCH_LIST = "r02_launch_list_train_b128.json"      # ncu launch list of one eager training step of this build (tools/ncu_launch_list.py)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--mode", default="train", choices=["train", "fwd"])
    p.add_argument("--batch", type=int, default=128, help="clips per GPU per step")
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--e2e-eager", action="store_true", help="time the e2e leg launch by launch (train_step + OptimizerHook sequence) instead of GraphedTrainStep")
    p.add_argument("--buckets", type=int, default=3, help="gradient buckets (all-reduce overlapped with backward)")
    p.add_argument("--sweep", default=None, choices=["coco"],
                   help="BASELINE config 4: large-batch inference sweep on 2-D HRNet keypoints (COCO layout, V=17, pixel-unit inputs), "
                        "batch 1..4096 clips on one GPU; prints one JSON line with the whole sweep (and writes --sweep-out)")
    p.add_argument("--sweep-out", default=None)
    p.add_argument("--sweep-max", type=int, default=4096)
    p.add_argument("--profile-step", action="store_true",
                   help="profiling aid: after the eager warm-up run ONE eager step between cudaProfilerStart/Stop and exit "
                        "(ncu --profile-from-start off ...); prints no bench line")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--ref-clips", type=int, default=REF_CLIPS, help="clips per step of the CPU reference arm (N=16: SURVEY 8d)")
    p.add_argument("--ref-budget", type=float, default=240.0, help="seconds the whole reference run may take (the batch shrinks to fit)")
    p.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-eager-on-this-GPU context number")
    p.add_argument("--no-fwd", action="store_true", help="skip the forward-metric leg reported in extra.fwd")
    return p.parse_args()


# ----------------------------------------------------------------------------------------------------------------
# Reference arm: the UNMODIFIED reference (pyskl DGSTGCN, imported by path from /root/reference or, on the GPU box, from
# baseline/_ref = byte-identical copies made by tools/install_ref.py) on the box's host cores.  oracle/ref_loader.py only
# stubs the mmcv names those files import; none of this repo's models or kernels are on this path.  If no reference tree
# is present the oracle port is timed instead and the line says kind "port".
# ----------------------------------------------------------------------------------------------------------------

REF_CLIPS = 16          # SURVEY.md §8(d) / BASELINE.md §4: N=16 for the CPU point


def _reference_model(device):
    """(backbone, head, params) of the unmodified reference, north-star config, reference default init + live dynamic branches."""
    from oracle import ref_loader as RL
    if not RL.available():
        return None
    ns = RL.load()
    torch.manual_seed(0)
    np.random.seed(0)
    backbone = ns.DGSTGCN(**RL.NORTH_STAR_BACKBONE)
    head = torch.nn.Linear(256, NUM_CLASSES)           # GCNHead.fc_cls (heads/simple_head.py:83-97): pool(T,V), mean(M), Linear
    torch.nn.init.normal_(head.weight, 0, 0.01)
    torch.nn.init.constant_(head.bias, 0)
    with torch.no_grad():
        for n_, p in backbone.named_parameters():
            if n_.rsplit(".", 1)[-1] in ("alpha", "beta", "add_coeff"):
                p.normal_(0, 0.1)
    return backbone.to(device), head.to(device), RL.REF_ROOT


def reference_arm(mode, clips, steps, warmup, device="cpu", budget_s=None):
    """Times `steps` steps of the reference itself after `warmup`; returns (clips/s, s/step, clips, description) or None."""
    built = _reference_model(device)
    if built is None:
        return None
    backbone, head, root = built
    if device == "cpu":
        torch.set_num_threads(os.cpu_count())
    train = mode == "train"
    backbone.train(train)
    params = [p for n_, p in backbone.named_parameters() if "conv2_se" not in n_] + list(head.parameters())
    opt = torch.optim.SGD(params, **SGD) if train else None

    def make(n):
        g = torch.Generator().manual_seed(0)
        return (torch.randn(n, M_, T_, V_, C_, generator=g).to(device), torch.randint(0, NUM_CLASSES, (n,), generator=g).to(device))

    def step(x, y):
        if not train:
            with torch.no_grad():
                return head(backbone(x).mean((3, 4)).mean(1))
        feat = backbone(x)
        loss = torch.nn.functional.cross_entropy(head(feat.mean((3, 4)).mean(1)), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def sync():
        if device != "cpu":
            torch.cuda.synchronize()

    x, y = make(clips)
    t0 = time.perf_counter()
    step(x, y)
    sync()
    t1 = time.perf_counter() - t0
    if budget_s is not None and t1 * (steps + warmup) > budget_s and clips > 2:      # bounded sample: shrink the batch, not the step count
        clips = max(2, int(clips * budget_s / (t1 * (steps + warmup))))
        x, y = make(clips)
    for _ in range(max(warmup - 1, 0)):
        step(x, y)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        step(x, y)
    sync()
    dt = (time.perf_counter() - t0) / steps
    what = (f"unmodified reference DGSTGCN ({root}) + Linear head, {mode}, {clips} clips/step x {steps} steps after {warmup} warm-up, "
            f"fp32, torch {torch.__version__} {device}")
    return clips / dt, dt, clips, what


def cpu_reference(mode, clips, steps, warmup):
    """Oracle port (only used when no reference tree is present)."""
    from oracle import dsgcn_oracle as O
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    np.random.seed(0)
    V = V_
    # parameters with the reference's names/shapes and PyTorch default init, via the oracle's own table helpers
    sd = _oracle_state(O)
    O.randomize_state(sd, 1)
    head_w = (torch.randn(NUM_CLASSES, 256) * 0.01).requires_grad_()
    head_b = torch.zeros(NUM_CLASSES, requires_grad=True)
    params = {k: v for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    for v in params.values():
        v.requires_grad_()
    opt = torch.optim.SGD([v for k, v in params.items() if "conv2_se" not in k] + [head_w, head_b], **SGD)
    x = torch.randn(clips, M_, T_, V, C_)
    label = torch.randint(0, NUM_CLASSES, (clips,))

    def step():
        if mode == "fwd":
            with torch.no_grad():
                return O.dgstgcn_forward(x, sd, training=False)
        feat = O.dgstgcn_forward(x, sd, training=True)
        logits = O.gcn_head_forward(feat, {"fc_cls.weight": head_w, "fc_cls.bias": head_b})
        loss = torch.nn.functional.cross_entropy(logits, label)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return clips / dt, dt


def cpu_baseline_leg(mode, clips, steps, warmup, budget_s=None):
    """dict for the `cpu_baseline` key: the reference itself when a reference tree is present, else the oracle port."""
    unit = "clips/s"
    try:
        r = reference_arm(mode, clips, steps, warmup, "cpu", budget_s)
    except Exception as e:                       # never lose the GPU line to the baseline leg
        print(f"[bench] reference arm failed ({type(e).__name__}: {str(e)[:200]}); timing the oracle port", file=sys.stderr)
        r = None
    if r is not None:
        val, dt, n, what = r
        return dict(value=val, unit=unit, cores=os.cpu_count(), kind="reference", sample=what, clips_per_step=n, ms_per_step=dt * 1e3,
                    dtype="f32", note="reference = fp32 on host cores at N=16 (SURVEY 8d); product arm = bf16 at 128 clips/GPU")
    n = min(clips, 4)
    val, dt = cpu_reference(mode, n, steps, warmup)
    return dict(value=val, unit=unit, cores=os.cpu_count(), kind="port", clips_per_step=n, ms_per_step=dt * 1e3, dtype="f32",
                sample=f"no reference tree present: oracle port of the reference path ({mode}), {n} clips/step x {steps} steps, torch {torch.__version__} CPU")


def _oracle_state(O):
    """Random-init state dict with the reference layout (shapes from the oracle's plan; PyTorch conv default init scale)."""
    import math
    sd = {}
    g = torch.Generator().manual_seed(0)

    def conv(name, cout, cin, k=1):
        bound = 1 / math.sqrt(cin * k)
        sd[name + ".weight"] = (torch.rand(cout, cin, k, 1, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def bn(name, c):
        sd[name + ".weight"], sd[name + ".bias"] = torch.ones(c), torch.zeros(c)
        sd[name + ".running_mean"], sd[name + ".running_var"] = torch.zeros(c), torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    bn("data_bn", V_ * C_)
    for i, (cin, cout, stride, res) in enumerate(O.dgstgcn_plan()):
        p = f"gcn.{i}.gcn."
        R = int(0.125 * cout)
        sd[p + "A"] = torch.randn(3, V_, V_, generator=g) * 0.02 + 0.04
        sd[p + "alpha"], sd[p + "beta"] = torch.zeros(3), torch.zeros(3)
        conv(p + "pre.0", 3 * R, cin); bn(p + "pre.1", 3 * R); conv(p + "post", cout, 3 * R)
        conv(p + "conv1_se", 5 * R, cin); conv(p + "conv2_se", 5 * R, cin); conv(p + "conv1", 2 * R, cin); conv(p + "conv2", 2 * R, cin)
        conv(p + "edge_linears", 15 * R, R)
        if cin != cout:
            conv(p + "down.0", cout, cin); bn(p + "down.1", cout)
        bn(p + "bn", cout)
        p = f"gcn.{i}.tcn."
        sd[p + "add_coeff"] = torch.zeros(25)
        mid = cout // 6
        widths = [cout - 5 * mid] + [mid] * 5
        for j, w in enumerate(widths):
            if j == 5:
                conv(p + "branches.5", w, cout)
            else:
                conv(p + f"branches.{j}.0", w, cout); bn(p + f"branches.{j}.1", w)
                if j < 4:
                    conv(p + f"branches.{j}.3.conv", w, w, 3)
        bn(p + "transform.0", cout); conv(p + "transform.2", cout, cout); bn(p + "bn", cout)
        if res == "conv":
            conv(f"gcn.{i}.residual.conv", cout, cin); bn(f"gcn.{i}.residual.bn", cout)
    return sd


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------

class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


def run_sweep(args):
    """Forward (eval, no_grad) clips/s of the COCO-layout DS-GCN over batch 1 .. 4096 on one GPU: device-resident inputs,
    CUDA events, >= 3 warm-up and >= 5 timed steps per size (sized to ~0.5 s), inputs re-used (sizes above ~16 clips exceed L2)."""
    import dsgcn_b200
    from dsgcn_b200 import modules as Mod
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    Mod.set_compute_dtype(dtype)
    torch.manual_seed(1234)
    np.random.seed(0)
    cfg = {**NORTH_STAR, "graph_cfg": dict(layout="coco", mode="random", num_filter=3, init_off=.04, init_std=.02)}
    model = dsgcn_b200.RecognizerGCN(backbone=dict(type="DGSTGCN", **cfg),
                                    cls_head=dict(type="GCNHead", num_classes=400, in_channels=256)).to(dev).eval()
    with torch.no_grad():
        for n_, p_ in model.named_parameters():
            if n_.rsplit(".", 1)[-1] in ("alpha", "beta", "add_coeff"):
                p_.normal_(0, 0.1)
        model.backbone.data_bn.running_mean.copy_(torch.tensor([128.0, 128.0, 0.5]).repeat(17))
        model.backbone.data_bn.running_var.copy_(torch.tensor([74.0 ** 2, 74.0 ** 2, 0.083]).repeat(17))
    T, V = 100, 17
    rows = []
    n = 1
    while n <= args.sweep_max:
        g = torch.Generator().manual_seed(n)
        x = torch.cat([torch.rand(n, 2, T, V, 2, generator=g) * 256, torch.rand(n, 2, T, V, 1, generator=g)], -1).to(dev)   # (x, y) pixels, score

        def eager():
            with torch.no_grad():
                return model.cls_head(model.extract_feat(x))
        for _ in range(3):
            eager()
        torch.cuda.synchronize()
        step, graphed = eager, False
        if not args.no_graph:                       # serving path: one captured forward per batch size, replayed
            try:
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    eager()
                gr.replay()
                torch.cuda.synchronize()
                step, graphed = gr.replay, True
            except Exception as e:
                print(f"[sweep] graph capture failed at batch {n} ({type(e).__name__}); eager", file=sys.stderr)
                torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        k = int(max(5, min(50, 500.0 / max(e0.elapsed_time(e1), 1e-3))))
        e0.record()
        for _ in range(k):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / k
        rows.append(dict(batch=n, ms_per_step=round(ms, 4), clips_per_s=round(n / (ms * 1e-3), 1), steps=k, cuda_graph=graphed,
                         peak_mem_gb=round(torch.cuda.max_memory_allocated() / 1e9, 2)))
        if graphed:
            del gr
        print(f"[sweep] batch {n:5d}: {ms:9.3f} ms  {n / (ms * 1e-3):10.1f} clips/s", file=sys.stderr, flush=True)
        del x
        n *= 2
    best = max(rows, key=lambda r: r["clips_per_s"])
    line = dict(metric="DS-GCN fwd clips/sec (COCO layout, M=2,T=100,V=17,C=3), batch sweep", value=best["clips_per_s"], unit="clips/s", n_gpus=1,
                higher_is_better=True, dtype=args.dtype, data="synthetic (pixel-unit x,y ~ U(0,256), score ~ U(0,1))",
                config=dict(workload="DS-GCN kinetics400_hrnet joint eval forward, batch sweep", layout="coco", M=2, T=T, V=V, C=3,
                            best_batch=best["batch"]), sweep=rows)
    print(json.dumps(line))
    if args.sweep_out:
        with open(args.sweep_out, "w") as fh:
            json.dump(line, fh, indent=1)


def _trace(msg):
    if os.environ.get("DSG_BENCH_TRACE"):
        print(f"[bench r{os.environ.get('RANK', 0)} {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    unit = "clips/s"
    metric = f"DS-GCN {args.mode} clips/sec (M=2,T=100,V=25,C=3)"
    workload = ("DS-GCN ntu60_xsub_3dkp joint training step (fwd+loss+bwd+SGD)" if args.mode == "train"
                else "DS-GCN ntu60_xsub_3dkp joint eval forward")

    if args.sweep:
        if rank == 0:
            run_sweep(args)
        return
    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_baseline_leg(args.mode, args.ref_clips, args.steps, args.warmup, budget_s=args.ref_budget)
        val, dt = cb["value"], cb["ms_per_step"] * 1e-3
        line = dict(impl="reference", metric=metric, value=val, unit=unit, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=workload, clips_per_step=cb["clips_per_step"], M=M_, T=T_, V=V_, C=C_, device="cpu",
                                note="reference's own CPU implementation: fp32, N=16 clips/step (SURVEY 8d); the product arm runs bf16 at 128 clips/GPU"),
                    cpu_baseline=cb, e2e=dict(value=val, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import dsgcn_b200
    from dsgcn_b200 import _lib as L
    from dsgcn_b200 import modules as Mod
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    Mod.set_compute_dtype(dtype)
    torch.manual_seed(1234 + rank)
    np.random.seed(0)
    model = dsgcn_b200.RecognizerGCN(backbone=dict(type="DGSTGCN", **NORTH_STAR),
                                    cls_head=dict(type="GCNHead", num_classes=NUM_CLASSES, in_channels=256)).to(dev)
    with torch.no_grad():       # make the dynamic branches live (zero at init: SURVEY.md §0 item 6)
        for n_, p in model.named_parameters():
            if n_.rsplit(".", 1)[-1] in ("alpha", "beta", "add_coeff"):
                p.normal_(0, 0.1)
    dsgcn_b200.parallel.broadcast_parameters(model)
    B = args.batch
    train = args.mode == "train"
    model.train(train)
    # flat parameter / gradient buckets in gradient-ready order (conv2_se never receives gradients, gcn.py:2253-2254, and stays
    # outside); per-bucket all-reduce launched from autograd hooks while backward is still running; one fused SGD launch per bucket
    gb = dsgcn_b200.parallel.GradBuckets(model, n_buckets=args.buckets) if train else None
    opt = dsgcn_b200.parallel.FlatSGD(gb, **SGD) if train else None

    # synthetic NTU-shaped data: a pool of pinned host batches (e2e) and device-resident batches (value)
    g = torch.Generator().manual_seed(rank)
    host_x = [torch.randn(B, 1, M_, T_, V_, C_, generator=g).pin_memory() for _ in range(2)]
    host_y = [torch.randint(0, NUM_CLASSES, (B, 1), generator=g).pin_memory() for _ in range(2)]
    dev_x = [h.to(dev) for h in host_x]
    dev_y = [h.to(dev) for h in host_y]
    sx, sy = dev_x[0].clone(), dev_y[0].clone()

    def device_step(x, y):
        if not train:
            with torch.no_grad():
                feat = model.extract_feat(x[:, 0])
                return model.cls_head(feat)
        losses = model(x, y, return_loss=True)
        loss = losses["loss_cls"]
        opt.zero_grad()
        loss.backward()                  # bucket all-reduces start inside (hooks), overlapped with the remaining backward
        opt.step()                       # waits for the collectives, then one fused update per bucket
        return loss

    # ---- warm-up (eager, on a side stream so no autograd node is tied to the default stream), then CUDA-graph
    #      capture of one whole step on static buffers
    L.launch_count = 0
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(max(args.warmup, 3)):
            device_step(dev_x[i % 2], dev_y[i % 2])
        torch.cuda.synchronize()
        l0 = L.launch_count
        device_step(sx, sy)
        launches_per_step = L.launch_count - l0
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    _trace("eager warm-up done")
    if args.profile_step:
        L.side_enabled = False                      # serialised launches: one stream, the order of the step
        device_step(dev_x[0], dev_y[0])
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        device_step(dev_x[1], dev_y[1])
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    graph = None
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):            # at N > 1 the NCCL collectives of the step are captured too
                device_step(sx, sy)
            _trace("captured")
            graph.replay()
            torch.cuda.synchronize()
            _trace("first replay done")
        except Exception as e:   # report, fall back to eager launches (still our kernels)
            print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {str(e)[:300]}); timing eager launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    used_graph = graph is not None

    def run_step(i):
        if graph is not None:
            sx.copy_(dev_x[i % 2], non_blocking=True)       # device-to-device refresh of the static input (inputs differ per step)
            sy.copy_(dev_y[i % 2], non_blocking=True)
            graph.replay()
        else:
            device_step(dev_x[i % 2], dev_y[i % 2])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(3):
        run_step(i)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        run_step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if sampler else None
    _trace(f"timed region done: {ms:.2f} ms/step")
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B / (ms * 1e-3)

    # ---- e2e: public API with host buffers (H2D of the batch from pinned memory, D2H of the logged scalars / scores).
    #      Training goes through dsgcn_b200.train.GraphedTrainStep — the whole iteration (forward, loss, backward with the bucketed
    #      all-reduce, update) as one captured graph replayed on static inputs; `--e2e-eager` times RecognizerGCN.train_step + the
    #      OptimizerHook sequence launch by launch instead.
    gstep = None
    if train and not args.e2e_eager and not args.no_graph:
        try:
            gstep = dsgcn_b200.train.GraphedTrainStep(model, opt, dev_x[0], dev_y[0])
        except Exception as e:
            print(f"[bench] GraphedTrainStep failed ({type(e).__name__}: {str(e)[:300]}); eager e2e", file=sys.stderr)
            gstep = None
            torch.cuda.synchronize()

    def e2e_step(i):
        if gstep is not None:
            return gstep(host_x[i % 2], host_y[i % 2])["log_vars"]["loss"]      # pinned host batch in, python floats out
        x = host_x[i % 2].to(dev, non_blocking=True)
        y = host_y[i % 2].to(dev, non_blocking=True)
        if train:
            opt.zero_grad()
            out = model.train_step(dict(keypoint=x, label=y), opt)       # the logged scalars come back to the host inside = D2H
            out["loss"].backward()
            opt.step()
            return out["log_vars"]["loss"]
        with torch.no_grad():
            return model(x, return_loss=False)                           # numpy scores = D2H
    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for i in range(n_e2e):
        e2e_step(i)
    barrier()
    e2e_ms = (time.perf_counter() - t0) / n_e2e * 1e3
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_graph = gstep is not None
    if gstep is not None:
        gstep.release()
        gstep = None
    _trace("e2e done")
    h2d = host_x[0].numel() * 4 + host_y[0].numel() * 8
    d2h = 16 if train else B * NUM_CLASSES * 4      # train: top1, top5, loss_cls, loss (one packed copy)

    # ---- roofline leg: per-ABI-call CUDA events over extra eager steps (same shapes), dominant kernel.
    #      Every rank runs the steps (they contain the gradient all-reduce); only rank 0 reports.
    L.profile = []
    side_was, L.side_enabled = L.side_enabled, False     # per-call events must not time kernels that overlap a side-stream kernel
    for i in range(2):
        device_step(dev_x[i % 2], dev_y[i % 2])
    torch.cuda.synchronize()
    L.side_enabled = side_was
    prof, L.profile = L.profile, None
    _trace("roofline leg done")
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
        if graph is not None:            # the captured NCCL kernels hold the communicator: release the graph before tearing it down
            graph.reset()
            graph = None
        _trace("graph released")
        torch.distributed.destroy_process_group()
        _trace("process group destroyed")
    if rank != 0:
        return
    agg = {}
    for name, nbytes, a, b, _tag in prof:
        d = agg.setdefault(name, [0.0, 0, 0])
        d[0] += a.elapsed_time(b)
        d[1] += nbytes
        d[2] += 1
    if os.environ.get("DSG_BENCH_DUMP"):       # per-call dump for kernel work (name, ms, GB/s, shape tag)
        with open(os.environ["DSG_BENCH_DUMP"], "w") as fh:
            for name, nbytes, a_, b_, tag in prof[len(prof) // 2:]:
                t_ = a_.elapsed_time(b_)
                fh.write(json.dumps(dict(name=name, ms=round(t_, 4), gbs=round(nbytes / (t_ * 1e-3) / 1e9, 1) if t_ > 0 else 0, tag=tag)) + "\n")
    total_ms = sum(v[0] for v in agg.values())
    top = max(agg, key=lambda k: agg[k][0])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    achieved = agg[top][1] / (agg[top][0] * 1e-3) / 1e9
    # DRAM traffic of the dominant entry point, from the committed ncu launch list of this exact workload (bytes per launch,
    # to set against the algorithmic bytes per launch); null for any other workload
    traffic, traffic_src, own_kernels, step_traffic = None, None, None, None
    if train and B == 128 and dtype == torch.bfloat16:
        try:
            ll = json.load(open(os.path.join(ROOT, "profiles", LAUNCH_LIST)))
            stem = top.replace("dsg_", "")
            # kernels behind the entry point: dsg_conv_gemm = the TMA engine tc4_gemm_kernel<...> (which also serves the 20
            # dsg_ms_conv launches of the step: same kernel name, they are in the mean) + the older conv_gemm_* engines
            stems = (stem, "tc4_gemm_kernel") if stem == "conv_gemm" else (stem,)
            ks = [k for k in ll["kernels"] if any(s_ in k["kernel"] for s_ in stems) and "wpack" not in k["kernel"]
                  and not k["kernel"].startswith("cutlass") and ("wgrad" in stem) == ("wgrad" in k["kernel"])]
            if ks:
                traffic = sum(k["dram_read_mb"] + k["dram_write_mb"] for k in ks) * 1e6 / sum(k["launches"] for k in ks)
                traffic_src = (f"profiles/{LAUNCH_LIST} (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per launch over "
                               f"{sum(k['launches'] for k in ks)} launches of {', '.join(stems)})")
            step_traffic = sum(k["dram_read_mb"] + k["dram_write_mb"] for k in ll["kernels"]) * 1e6
            own_kernels = sum(k["launches"] for k in ll["kernels"]
                              if not k["kernel"].startswith(("void at::", "void at_cuda", "cutlass", "void sbtopk", "void cub", "void cutlass")))
        except Exception:
            pass
    roofline = dict(bound="hbm", kernel=top, achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    traffic_unit="bytes per launch", algorithmic_bytes_per_launch=agg[top][1] / agg[top][2], traffic_source=traffic_src,
                    peak_source="MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                    share_of_step=agg[top][0] / total_ms, launches_per_step=agg[top][2] // 2,
                    per_kernel_ms_per_step={k: round(v[0] / 2, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])},
                    per_kernel_gbs={k: round(v[1] / (v[0] * 1e-3) / 1e9, 1) for k, v in agg.items() if v[1] > 0 and v[0] > 0},
                    step_algorithmic_gbs=ELEMS_PER_CLIP_FWD * (3 if train else 1) * (2 if dtype == torch.bfloat16 else 4) * B / (ms * 1e-3) / 1e9)
    # whole-step roofline: SURVEY 8(d) stage-granular algorithmic bytes of the step / step time / measured HBM peak, and the DRAM
    # bytes the step really moves (committed ncu launch list of this workload)
    roofline["step_algorithmic_bytes"] = ELEMS_PER_CLIP_FWD * (3 if train else 1) * (2 if dtype == torch.bfloat16 else 4) * B
    roofline["step_frac"] = roofline["step_algorithmic_gbs"] / peak
    roofline["step_traffic_bytes"] = step_traffic
    roofline["step_traffic_over_algorithmic"] = (step_traffic / roofline["step_algorithmic_bytes"]) if step_traffic else None
    # context for `frac`: what a plain device copy moving the same bytes per launch reaches on this GPU (the measured HBM peak is a
    # 2 GiB copy; at the network's tensor sizes a launch has a fixed cost of ~8 us) — tools/probes/copy_small.py, DESIGN.md 4.7
    if world == 1:
        try:
            nel = max(int(roofline["algorithmic_bytes_per_launch"]) // 4, 1 << 20)          # bf16 elements read (= written)
            k = 4
            srcs = [torch.empty(nel, dtype=torch.bfloat16, device=dev).normal_() for _ in range(k)]
            dsts = [torch.empty(nel, dtype=torch.bfloat16, device=dev) for _ in range(k)]
            for i in range(k):
                dsts[i].copy_(srcs[i])
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for i in range(5 * k):
                dsts[i % k].copy_(srcs[i % k])
            c1.record()
            torch.cuda.synchronize()
            cg = 4.0 * nel / (c0.elapsed_time(c1) * 1e-3 / (5 * k)) / 1e9
            roofline["copy_same_bytes_gbs"] = round(cg, 1)
            roofline["frac_of_copy_same_bytes"] = round(achieved / cg, 4)
            del srcs, dsts
        except Exception:
            pass
    cpu = None
    extra = {}
    if train and world == 1 and not args.no_fwd:
        # the forward metric of BASELINE.json in the same run: eval forward (no_grad) of the same model, CUDA-graph replay
        try:
            model.eval()
            def fwd_step(x):
                with torch.no_grad():
                    return model.cls_head(model.extract_feat(x[:, 0]))
            side2 = torch.cuda.Stream()
            side2.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side2):
                for i in range(3):
                    fwd_step(dev_x[i % 2])
            torch.cuda.current_stream().wait_stream(side2)
            torch.cuda.synchronize()
            gf = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gf):
                fwd_step(sx)
            for i in range(3):
                sx.copy_(dev_x[i % 2]); gf.replay()
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for i in range(args.steps):
                sx.copy_(dev_x[i % 2], non_blocking=True); gf.replay()
            f1.record()
            torch.cuda.synchronize()
            fms = f0.elapsed_time(f1) / args.steps
            fgbs = ELEMS_PER_CLIP_FWD * (2 if dtype == torch.bfloat16 else 4) * B / (fms * 1e-3) / 1e9
            extra["fwd"] = dict(metric="DS-GCN fwd clips/sec (M=2,T=100,V=25,C=3)", value=B / (fms * 1e-3), unit=unit, ms_per_step=fms, clips_per_gpu=B,
                                cuda_graph=True, step_algorithmic_gbs=fgbs, step_frac=fgbs / peak)
            gf.reset()
            model.train(True)
        except Exception as e:
            extra["fwd"] = dict(unavailable=f"{type(e).__name__}: {str(e)[:200]}")
    if not args.no_cpu_baseline and world == 1:      # the CPU baseline is reported at N=1 only (rank 0's host cores)
        cpu = cpu_baseline_leg(args.mode, args.ref_clips, 2, 1, budget_s=40.0)
    if not args.no_ref_gpu and world == 1:
        # context only (never a denominator): the unmodified reference run eagerly on this GPU, same batch as the product arm
        try:
            del model, opt
            torch.cuda.empty_cache()
            r = reference_arm(args.mode, B, 3, 2, device=f"cuda:{local}")
            if r is not None:
                extra["reference_gpu_eager"] = dict(value=r[0], unit=unit, ms_per_step=r[1] * 1e3, clips_per_step=r[2], sample=r[3])
        except Exception as e:
            extra["reference_gpu_eager"] = dict(unavailable=f"{type(e).__name__}: {str(e)[:200]}")
    line = dict(metric=metric, value=value, unit=unit, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.dtype, data="synthetic",
                config=dict(workload=workload, clips_per_gpu=B, M=M_, T=T_, V=V_, C=C_, parallelism=f"dp{world}", cuda_graph=used_graph,
                            allreduce=(f"{len(gb.buckets)} flat buckets ({gb.grad_bytes()} B), NCCL AVG launched from autograd hooks during backward"
                                       if (train and world > 1) else None),
                            cache="inputs + activations per step (~GBs) exceed the 126 MB L2; two input batches alternate"),
                e2e=dict(value=world * B / (e2e_ms * 1e-3), unit=unit, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=e2e_ms,
                         api="dsgcn_b200.train.GraphedTrainStep (captured iteration, host batch in, logged scalars out)" if e2e_graph
                         else ("RecognizerGCN.train_step + zero_grad/backward/step" if train else "RecognizerGCN.forward(return_loss=False)")),
                gpu_launches=launches_per_step * args.steps, abi_calls_per_step=launches_per_step, kernels_per_step_ncu=own_kernels, clocks=clocks, roofline=roofline, cpu_baseline=cpu, extra=extra)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
