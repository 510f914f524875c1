"""Developer tool: the single fused kernel (attention + to_out + gate + residual) vs the two-kernel path at the c2 shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from otter_b200 import functional as F

dev = "cuda:0"
torch.manual_seed(0)
BF = torch.bfloat16
B, L, D, n, H, inner = 8, 256, 4096, 64, 8, 512
sets = []
for _ in range(6):          # rotating operand sets (each ~55 MB): cold-ish operands
    q = torch.randn(B * L, inner, device=dev).to(BF)
    kv = torch.randn(B * n, 2 * inner, device=dev).to(BF)
    wo = (torch.randn(D, inner, device=dev) * inner ** -0.5).to(BF)
    x = torch.randn(B * L, D, device=dev).to(BF)
    loc = torch.zeros(B, L, dtype=torch.bool, device=dev)
    loc[:, 0] = True
    spec = F.AttnSpec(q, 0, kv, 0, inner, B, H, L, n, 0.125, text_time=F.text_time(loc), n_per_media=n, T_img=1)
    sets.append((spec, wo, x))
gate = torch.tensor([0.5], device=dev)


def fused(s):
    return F.xattn_out_fused(s[0], s[1], gate, s[2])


def two(s):
    o, lse = F.attn_fwd(s[0])
    a = torch.empty(B * L, D, device=dev, dtype=BF)
    return F.linear_fwd(o, s[1], aux_out=a, scale_ptr=gate, scale_tanh=True, residual=s[2])


fused(sets[0]); fused(sets[0]); two(sets[0])
torch.cuda.synchronize()
if "--time" in sys.argv:
    print(f"fused kernel      : {bench.graph_time_us([(lambda s=s: fused(s)) for s in sets] * 3):.2f} us per block")
    print(f"two-kernel path   : {bench.graph_time_us([(lambda s=s: two(s)) for s in sets] * 3):.2f} us per block (attention + to_out GEMM)")
