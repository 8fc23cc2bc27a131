"""Stream launches (with programmatic dependent launch) vs CUDA-graph replay of one DDS data-consistency step.

    python tools/graph_time.py
Measured on B200: B=1 146 -> 139 us, B=8 278 -> 267 us, B=32 766 -> 735 us (PDL already hides most of the launch latency).
"""
import sys, torch, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusion_models_dev_project_b200 as pkg
dev = torch.device('cuda')
for B in (1, 8, 32):
    rt = pkg.B200RayTrafo((256, 256), 60)
    abar = pkg.DDPM().alpha_bar_table(dev)
    x = torch.rand(B, 1, 256, 256, device=dev); s = torch.randn_like(x); eps = torch.randn_like(x)
    atb = rt.trafo_adjoint(rt(torch.rand_like(x)))
    t = torch.ones(B, device=dev) * 500.; tp = torch.ones(B, device=dev) * 490.
    f = lambda: rt.dds_step(x, s, atb, eps, t, tp, abar, 0.01, 0.15, 5)
    for _ in range(5): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = f()
    def timeit(fn, n=50):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3
    print('B=%d  stream %.1f us   graph replay %.1f us' % (B, timeit(f), timeit(g.replay)))
