"""Developer tool: launch latency-bound small GEMMs of the hot path (for ncu captures / timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from otter_b200 import functional as F
from otter_b200.graph import GraphedStep

dev = "cuda:0"
torch.manual_seed(0)
shapes = {"perceiver to_kv(latents) 512x1024x1024": (512, 1024, 1024), "clip out_proj 2056x1024x1024": (2056, 1024, 1024),
          "clip fc2 2056x1024x4096": (2056, 1024, 4096), "perceiver ff2 512x1024x4096": (512, 1024, 4096)}
for name, (M, N, K) in shapes.items():
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        F.linear_fwd(x, w, out=out)
    torch.cuda.synchronize()
    def ten():
        for _ in range(10):
            F.linear_fwd(x, w, out=out)
    g = GraphedStep(ten)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(f"{name}: {us:.1f} us per launch back-to-back in a graph ({2 * M * N * K / us / 1e6:.0f} TFLOP/s)")
