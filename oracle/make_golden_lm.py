"""TEST INFRASTRUCTURE — generates tests/golden/mpt_block_tiny.pt from the UNMODIFIED reference MPTBlock
(/root/reference/src/otter_ai/models/mpt/blocks.py:23-88, attn_impl "torch", alibi, no bias) on seeded inputs.
Run in the build container only:  python oracle/make_golden_lm.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims  # noqa: E402


def reference_block(d_model, n_heads, no_bias=True, seed=0):
    ref_shims._install_stubs()
    sys.path.insert(0, os.path.join(ref_shims.REF_ROOT, "src"))
    from otter_ai.models.mpt.attention import build_attn_bias
    from otter_ai.models.mpt.blocks import MPTBlock
    torch.manual_seed(seed)
    attn_config = {"attn_type": "multihead_attention", "attn_pdrop": 0.0, "attn_impl": "torch", "qk_ln": False,
                   "clip_qkv": None, "softmax_scale": None, "prefix_lm": False, "attn_uses_sequence_id": False,
                   "alibi": True, "alibi_bias_max": 8}
    blk = MPTBlock(d_model=d_model, n_heads=n_heads, expansion_ratio=4, attn_config=attn_config).eval()
    if no_bias:                                   # modeling_mpt.py: `no_bias` drops every bias (and LN bias) parameter
        for m in blk.modules():
            if hasattr(m, "bias") and isinstance(m.bias, torch.nn.Parameter):
                m.register_parameter("bias", None)
    with torch.no_grad():
        for prm in blk.parameters():
            prm.copy_(torch.randn_like(prm) * (0.5 if prm.ndim == 1 else prm.shape[-1] ** -0.5))
            if prm.ndim == 1:
                prm.add_(1.0)

    def run(x):
        S = x.shape[1]
        bias = build_attn_bias("torch", torch.zeros(1, n_heads, 1, S), n_heads, S, causal=True, alibi=True,
                               alibi_bias_max=8)
        y, _, _ = blk(x, attn_bias=bias, is_causal=True)
        return y

    return blk, run


def main():
    out = {}
    for name, (B, S, D, H) in {"a": (2, 19, 64, 4), "b": (1, 33, 96, 6)}.items():     # H=6: non-power-of-two slopes
        blk, run = reference_block(D, H, no_bias=True, seed=3)
        x = torch.randn(B, S, D, generator=torch.Generator().manual_seed(11)).requires_grad_(True)
        y = run(x)
        gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(12))
        (gx,) = torch.autograd.grad(y, x, gy)
        out[name] = {"shape": (B, S, D, H), "params": {k: v.detach().clone() for k, v in blk.state_dict().items()},
                     "x": x.detach().clone(), "y": y.detach().clone(), "gy": gy, "gx": gx.detach().clone()}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "mpt_block_tiny.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
