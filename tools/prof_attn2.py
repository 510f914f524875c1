"""Developer tool: the fused attention kernels at the step's shapes — A (perceiver), B (gated x-attn), CLIP, and the
backward of A and B — two launches each (for `ncu -k regex:attn`), then CUDA-graph timings (us per launch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from otter_b200 import functional as F

dev = "cuda:0"
torch.manual_seed(0)
BF = torch.bfloat16


def rn(*s):
    return torch.randn(*s, device=dev).to(BF)


def specA(P=8, n1=256, n2=64, H=8):
    q, kv = rn(P * n2, 512), rn(P * (n1 + n2), 1024)
    return F.AttnSpec(q, 0, kv[:P * n1], 0, 512, P, H, n2, n1, 0.125, kv2=kv[P * n1:], k2_col0=0, v2_col0=512, Sk2=n2), kv


def specB(B=8, L=256, H=8):
    q, kv = rn(B * L, 512), rn(B * 64, 1024)
    loc = torch.zeros(B, L, dtype=torch.bool, device=dev)
    loc[:, 0] = True
    return F.AttnSpec(q, 0, kv, 0, 512, B, H, L, 64, 0.125, text_time=F.text_time(loc), n_per_media=64, T_img=1), kv


def specC(N=8, S=257, H=16):
    qkv = rn(N * S, 3 * H * 64)
    return F.AttnSpec(qkv, 0, qkv, H * 64, 2 * H * 64, N, H, S, S, 0.125), qkv


def bwd_call(spec, kv, P, Sq, Sk1):
    o, lse = F.attn_fwd(spec)
    do, dq, dkv = rn(P * Sq, 512), torch.empty(P * Sq, 512, device=dev, dtype=BF), torch.empty_like(kv)
    if spec.kv2 is not None:
        return lambda: F.attn_bwd(spec, o, 0, lse, do, 0, dq, 0, dkv[:P * Sk1], 0, 512, dkv[P * Sk1:], 0, 512)
    return lambda: F.attn_bwd(spec, o, 0, lse, do, 0, dq, 0, dkv, 0, 512)


sa, kva = specA()
sb, kvb = specB()
sc, _ = specC()
calls = {"A fwd": lambda: F.attn_fwd(sa), "B fwd": lambda: F.attn_fwd(sb), "CLIP fwd": lambda: F.attn_fwd(sc, want_lse=False),
         "A bwd": bwd_call(sa, kva, 8, 64, 256), "B bwd": bwd_call(sb, kvb, 8, 256, 64)}
for name, c in calls.items():
    c(); c()
torch.cuda.synchronize()
if "--time" in sys.argv:
    for name, c in calls.items():
        print(f"{name}: {bench.graph_time_us([c] * 20):.2f} us per launch (graph replay, warm L2)")
