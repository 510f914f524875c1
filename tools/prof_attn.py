"""Developer tool: launch the fused attention forward at the three hot-path shapes (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from otter_b200 import functional as F

dev = "cuda:0"
torch.manual_seed(0)
def clip(N=8, S=257, H=16):
    qkv = torch.randn(N * S, 3 * H * 64, device=dev).to(torch.bfloat16)
    return F.AttnSpec(qkv, 0, qkv, H * 64, 2 * H * 64, N, H, S, S, 0.125)
def perceiver(BT=8, n1=256, n2=64, H=8):
    q = torch.randn(BT * n2, 512, device=dev).to(torch.bfloat16)
    kvx = torch.randn(BT * n1, 1024, device=dev).to(torch.bfloat16)
    kvl = torch.randn(BT * n2, 1024, device=dev).to(torch.bfloat16)
    return F.AttnSpec(q, 0, kvx, 0, 512, BT, H, n2, n1, 0.125, kv2=kvl, k2_col0=0, v2_col0=512, Sk2=n2)
def xattn(B=8, L=256, H=8):
    q = torch.randn(B * L, 512, device=dev).to(torch.bfloat16)
    kv = torch.randn(B * 64, 1024, device=dev).to(torch.bfloat16)
    loc = torch.zeros(B, L, dtype=torch.bool, device=dev); loc[:, 0] = True
    return F.AttnSpec(q, 0, kv, 0, 512, B, H, L, 64, 0.125, text_time=F.text_time(loc), n_per_media=64, T_img=1)
specs = {"clip": clip(), "perceiver": perceiver(), "xattn": xattn()}
for name, sp in specs.items():
    for _ in range(3):
        F.attn_fwd(sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        F.attn_fwd(sp)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
