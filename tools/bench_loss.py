"""Developer tool: HBM roofline of the §8f row-2 kernels (fused shifted cross-entropy, label masking)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from otter_b200 import losses

dev = "cuda:0"
B, L, V = 8, 256, 50432            # MPT-7B vocab (Otter-MPT7B-config.json), per-GPU batch 8, L = 256
torch.manual_seed(0)
logits = torch.randn(B, L, V, device=dev).to(torch.bfloat16).requires_grad_(True)
labels = torch.randint(0, V, (B, L), device=dev)
ids = torch.randint(0, 50, (B, L), device=dev)
from otter_b200.graph import GraphedStep


def graph_time(fn, iters=20):
    """GPU time per call with the host launch path taken out (captured once, replayed)."""
    g = GraphedStep(fn)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = graph_time(lambda: losses.shifted_cross_entropy(logits, labels))
alg = B * L * V * 2 * 2            # algorithmic bytes: logits read once + dlogits written once (bf16)
peak = 6569.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"kernel": "otb_shifted_cross_entropy (count + row + finalize)", "shape": [B, L, V], "dtype": "bf16",
                  "ms": round(ms, 4), "algorithmic_GB": round(alg / 1e9, 3), "achieved_GBps": round(alg / ms / 1e6, 1),
                  "peak_GBps": peak, "frac": round(alg / ms / 1e6 / peak, 3),
                  "note": "row staged in shared memory: logits read once, gradient written once; graph-replayed (no host launch cost)"}))
us = graph_time(lambda: losses.label_mask(ids, 2, 7, 8)) * 1e3
print(json.dumps({"kernel": "otb_label_mask", "shape": [B, L], "us": round(us, 2), "note": "graph-replayed"}))
