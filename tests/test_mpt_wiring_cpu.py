"""CPU: host logic of the frozen MPT block (otter_b200/lm_mpt.py) — the autograd wiring of _FrozenMPTBlockFn with the
CUDA entry points replaced by plain torch formulas of what each computes (include/otter_b200.h), against the
reference-pinned oracle (oracle/restatement_lm.py, pinned to the reference's MPTBlock in tests/test_oracle_lm_cpu.py).
The kernels themselves are checked on the GPU (tests/test_lm_gpu.py)."""
import math

import pytest
import torch

from oracle import restatement_lm as RL


def _gelu(v):
    return 0.5 * v * (1 + torch.erf(v / math.sqrt(2)))


@pytest.fixture
def torch_kernels(monkeypatch):
    from otter_b200 import functional as F
    from otter_b200 import lm_mpt
    from otter_b200 import params as P

    def layernorm_fwd(x, g, b, eps=1e-5):
        mean = x.mean(-1)
        rstd = torch.rsqrt(x.var(-1, unbiased=False) + eps)
        return (x - mean[:, None]) * rstd[:, None] * g + b, mean, rstd

    def layernorm_bwd(dy, x, mean, rstd, g, add=None, want_param_grads=True):
        assert not want_param_grads                                   # frozen layer
        dy, add = dy.float(), (add.float() if add is not None else None)
        xh = (x - mean[:, None]) * rstd[:, None]
        gy = dy * g
        dx = rstd[:, None] * (gy - gy.mean(-1, keepdim=True) - xh * (gy * xh).mean(-1, keepdim=True))
        return (dx if add is None else dx + add), None, None

    def linear_fwd(x, w, *, bias=None, act=0, aux_out=None, residual=None):
        z = x @ w.t() if bias is None else x @ w.t() + bias
        if aux_out is not None:
            aux_out.copy_(z)                                          # pre-activation kept for the dGELU epilogue
        y = _gelu(z) if act == 1 else z
        return y if residual is None else y + residual

    def linear_dgrad(dy, w, *, aux_in=None):
        dx = dy.float() @ w
        if aux_in is not None:                                        # epilogue: times gelu'(z)
            z = aux_in.detach().requires_grad_(True)
            with torch.enable_grad():
                dx = torch.autograd.grad(_gelu(z), z, dx)[0]
        return dx

    def _attn(qkv, B, S, H, slopes, scale):
        D = H * 128
        sp = lambda t: t.reshape(B, S, H, 128).transpose(1, 2)
        q, k, v = sp(qkv[:, :D]), sp(qkv[:, D:2 * D]), sp(qkv[:, 2 * D:])
        s = (q @ k.transpose(-1, -2)) * scale
        if slopes is not None:
            s = s + slopes.view(1, H, 1, 1) * torch.arange(1 - S, 1, dtype=torch.float32).view(1, 1, 1, S)
        s = s.masked_fill(~torch.ones(S, S, dtype=torch.bool).tril(), float("-inf"))
        return (s.softmax(-1) @ v).transpose(1, 2).reshape(B * S, D)

    def lm_attn_fwd(qkv, B, S, H, *, slopes=None, causal=True, scale=None):
        assert causal
        return _attn(qkv, B, S, H, slopes, scale), torch.zeros(B, H, S)

    def lm_attn_bwd(dout, qkv, out, lse, B, S, H, *, slopes=None, causal=True, scale=None):
        q = qkv.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            return torch.autograd.grad(_attn(q, B, S, H, slopes, scale), q, dout.float())[0]

    for name, fn in dict(layernorm_fwd=layernorm_fwd, layernorm_bwd=layernorm_bwd, linear_fwd=linear_fwd,
                         linear_dgrad=linear_dgrad, lm_attn_fwd=lm_attn_fwd, lm_attn_bwd=lm_attn_bwd).items():
        monkeypatch.setattr(F, name, fn)
    monkeypatch.setattr(P, "bf16_of", lambda p: p.detach().float())
    monkeypatch.setattr(P, "f32_of", lambda p: p.detach().float())
    real_empty = torch.empty
    monkeypatch.setattr(lm_mpt.torch, "empty", lambda *a, **k: real_empty(*a, **{**k, "dtype": torch.float32}))
    real_fn = lm_mpt._FrozenMPTBlockFn.apply
    monkeypatch.setattr(lm_mpt._FrozenMPTBlockFn, "apply", staticmethod(lambda x, *a: real_fn(x.float(), *a)))


@pytest.mark.parametrize("no_bias", [True, False])
def test_frozen_mpt_block_wiring_matches_the_oracle(torch_kernels, no_bias):
    from otter_b200.lm_mpt import FrozenMPTBlock
    B, S, D, H = 2, 19, 256, 2
    torch.manual_seed(1)
    blk = FrozenMPTBlock(D, H, no_bias=no_bias)
    with torch.no_grad():
        for p in blk.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    p_ref = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    # bf16-representable inputs: the module's casts to the kernels' storage type are then lossless, and the stand-in
    # arithmetic (fp32) can be compared with the fp32 oracle at fp32 tolerances
    x = torch.randn(B, S, D).bfloat16().float()
    w = torch.randn(B, S, D).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    ref = RL.mpt_block(xr, p_ref, H)
    (ref * w).sum().backward()
    x2 = x.clone().requires_grad_(True)
    mask = torch.ones(B, S, dtype=torch.long)
    mask[1, 15:] = 0                                                   # right padding is accepted
    y, attn_w, past = blk(x2, attention_mask=mask)
    assert attn_w is None and past is None                             # the reference block's return tuple
    (y * w).sum().backward()
    assert torch.allclose(y, ref, rtol=1e-4, atol=1e-5)
    # the module hands bf16 to the kernels, so autograd rounds the returned gradient to bf16 at that cast (8 mantissa bits)
    assert torch.allclose(x2.grad, xr.grad, rtol=8e-3, atol=2e-3)
    assert all(p.grad is None for p in blk.parameters())
    with pytest.raises(NotImplementedError):
        blk(x2, attention_mask=mask.flip(1))                           # left padding is refused
    with pytest.raises(NotImplementedError):
        blk(x2, past_key_value=(x, x))
