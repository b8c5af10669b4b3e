"""What a plain device copy reaches at the network's tensor sizes (the 6544 GB/s peak is a 2 GiB copy): b.copy_(a) over rotating
buffers (working set > L2), CUDA events, per-copy time and read+write GB/s.  Context for the per-kernel roofline numbers."""
import torch
dev = torch.device("cuda")
for mb in (20, 41, 82, 164, 328, 1024):
    n = mb * 1000 * 1000 // 2
    k = max(4, (600 * 1000 * 1000) // (mb * 1000 * 1000))       # rotate over > 600 MB of sources
    src = [torch.randn(n, device=dev).to(torch.bfloat16) for _ in range(k)]
    dst = [torch.empty(n, dtype=torch.bfloat16, device=dev) for _ in range(k)]
    for i in range(k):
        dst[i].copy_(src[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5 * k
    e0.record()
    for r in range(reps):
        dst[r % k].copy_(src[r % k])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"copy {mb:5d} MB (read) + same (write): {us:8.1f} us  {2 * n * 2 / us / 1e3:7.0f} GB/s", flush=True)
