"""GPU parity of §8f rank 1: causal / ALiBi head-dim-128 attention (otb_lm_attn_fwd / _bwd) and the
frozen MPT block built on it, against fp32 torch math and the reference-pinned oracle (oracle/restatement_lm.py).
Tolerances are bf16-storage tolerances (inputs, P and outputs are rounded to bf16 in the kernels)."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def _ref_attn(qkv, B, S, H, slopes, causal):
    D = H * 128
    q, k, v = qkv.float().view(B, S, 3 * D).split(D, dim=2)
    q = q.view(B, S, H, 128).permute(0, 2, 1, 3)
    k = k.view(B, S, H, 128).permute(0, 2, 1, 3)
    v = v.view(B, S, H, 128).permute(0, 2, 1, 3)
    w = q @ k.transpose(-1, -2) / math.sqrt(128)
    if slopes is not None:
        pos = torch.arange(1 - S, 1, device=qkv.device, dtype=torch.float32)
        w = w + slopes.view(1, H, 1, 1) * pos.view(1, 1, 1, S)
    if causal:
        w = w.masked_fill(~torch.ones(S, S, dtype=torch.bool, device=qkv.device).tril(), float("-inf"))
    p = torch.softmax(w, dim=-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * S, D)


@pytest.mark.parametrize("B,S,H,alibi,causal", [(2, 200, 2, True, True), (1, 256, 3, True, True), (1, 77, 1, False, True),
                                                 (1, 384, 2, True, True), (2, 130, 1, True, False)])
def test_lm_attention_fwd_bwd(B, S, H, alibi, causal):
    from otter_b200 import functional as F
    from otter_b200.lm_mpt import alibi_slopes
    torch.manual_seed(S + H)
    D = H * 128
    qkv = (torch.randn(B * S, 3 * D, device=dev()) * 0.8).to(torch.bfloat16)
    slopes = alibi_slopes(H, device=dev()) if alibi else None
    out, lse = F.lm_attn_fwd(qkv, B, S, H, slopes=slopes, causal=causal)
    ref_in = qkv.float().requires_grad_(True)
    want = _ref_attn(ref_in, B, S, H, slopes, causal)
    err = (out.float() - want).abs().max().item()
    assert err < 3e-2, err
    dout = (torch.randn(B * S, D, device=dev()) * 0.5).to(torch.bfloat16)
    want.backward(dout.float())
    dqkv = F.lm_attn_bwd(dout, qkv, out, lse, B, S, H, slopes=slopes, causal=causal)
    g = ref_in.grad
    rel = (dqkv.float() - g).abs().max().item() / g.abs().max().item()
    assert rel < 3e-2, rel


def test_frozen_mpt_block_matches_oracle():
    from oracle import restatement_lm as R
    from otter_b200.lm_mpt import FrozenMPTBlock
    torch.manual_seed(0)
    B, S, D, H = 2, 150, 256, 2
    blk = FrozenMPTBlock(D, H).to(dev())
    with torch.no_grad():
        for prm in blk.parameters():
            prm.copy_(torch.randn_like(prm) * (0.3 if prm.ndim == 1 else prm.shape[-1] ** -0.5))
            if prm.ndim == 1:
                prm.add_(1.0)
    x = torch.randn(B, S, D, device=dev()).to(torch.bfloat16).requires_grad_(True)
    y, _, _ = blk(x)
    gy = torch.randn(B, S, D, device=dev()).to(torch.bfloat16)
    (gx,) = torch.autograd.grad(y, x, gy)
    params = {k: v.detach().float().cpu() for k, v in blk.state_dict().items()}
    xr = x.detach().float().cpu().requires_grad_(True)
    yr = R.mpt_block(xr, params, H)
    (gxr,) = torch.autograd.grad(yr, xr, gy.float().cpu())
    rel_y = (y.float().cpu() - yr).abs().max().item() / yr.abs().max().item()
    rel_g = (gx.float().cpu() - gxr).abs().max().item() / gxr.abs().max().item()
    assert rel_y < 3e-2, rel_y
    assert rel_g < 5e-2, rel_g
    assert all(p.grad is None for p in blk.parameters())          # frozen: no weight gradients are produced


def test_golden_key_names_match_reference_block():
    from otter_b200.lm_mpt import FrozenMPTBlock
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "mpt_block_tiny.pt"))
    assert set(FrozenMPTBlock(256, 2).state_dict().keys()) == set(gold["a"]["params"].keys())
