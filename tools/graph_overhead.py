"""Micro-benchmark: per-node latency of dependent tiny kernels, in a CUDA graph and as plain stream launches."""
import torch, time
dev = torch.device("cuda:0")
x = torch.zeros(1024, device=dev)
N = 200
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(10):
        x.add_(1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(N):
            x.add_(1)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(10):
        g.replay()
    b.record(s)
    torch.cuda.synchronize()
    print("graph: us per tiny dependent kernel node:", a.elapsed_time(b) * 1e3 / (10 * N))
    a.record(s)
    for _ in range(10 * N):
        x.add_(1)
    b.record(s)
    torch.cuda.synchronize()
    print("stream launches: us per tiny kernel:", a.elapsed_time(b) * 1e3 / (10 * N))
    # bigger kernel: 64 MB elementwise (should be ~20 us of work)
    y = torch.zeros(16 << 20, device=dev)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2, stream=s):
        for _ in range(50):
            y.add_(1)
    g2.replay(); torch.cuda.synchronize()
    a.record(s)
    for _ in range(10):
        g2.replay()
    b.record(s)
    torch.cuda.synchronize()
    print("graph: us per 64MB add_ node:", a.elapsed_time(b) * 1e3 / 500, "(ideal ~", 2 * 64e6 / 6.5e12 * 1e6, "us )")
