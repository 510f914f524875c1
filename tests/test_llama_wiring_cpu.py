"""CPU: host logic of the frozen LLaMA decoder layer (otter_b200/lm_llama.py) — the autograd wiring of
_FrozenLlamaLayerFn with the CUDA entry points replaced by plain torch formulas of what each computes
(include/otter_b200.h), against HF's LlamaDecoderLayer (what the reference instantiates, modeling_otter.py:52-54).
The kernels themselves are checked on the GPU (tests/test_llama_gpu.py)."""
import math

import pytest
import torch


def _rot(x):
    return torch.cat((-x[..., 64:], x[..., :64]), -1)


@pytest.fixture
def torch_kernels(monkeypatch):
    from otter_b200 import functional as F
    from otter_b200 import params as P

    def rmsnorm_fwd(x, w, eps, want_rstd=True):
        rstd = torch.rsqrt(x.float().pow(2).mean(-1) + eps)
        return x.float() * rstd[:, None] * w, rstd

    def rmsnorm_bwd(dy, x, rstd, w, add=None):
        xh = x.float() * rstd[:, None]
        g = dy.float() * w
        dx = rstd[:, None] * (g - xh * (g * xh).mean(-1, keepdim=True))
        return dx if add is None else dx + add

    def linear_fwd(x, w, *, out=None, residual=None, **kw):
        assert not kw
        y = x @ w.t()
        if residual is not None:
            y = y + residual
        if out is not None:
            out.copy_(y)
            return out
        return y

    def linear_dgrad(dy, w, *, residual=None, **kw):
        assert not kw
        dx = dy @ w
        return dx if residual is None else dx + residual

    def rope128_(buf, H, S, nblk, theta, backward=False):
        rows = buf.shape[0]
        inv = 1.0 / (theta ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))
        ang = (torch.arange(rows) % S).float()[:, None] * inv[None]
        cos, sin = torch.cat((ang, ang), -1).cos()[:, None, None], torch.cat((ang, ang), -1).sin()[:, None, None]
        if backward:
            sin = -sin
        v = buf[:, :nblk * H * 128].reshape(rows, nblk, H, 128)
        buf[:, :nblk * H * 128] = (v * cos + _rot(v) * sin).reshape(rows, -1)
        return buf

    def _attn(qkv, B, S, H):
        D = H * 128
        sp = lambda t: t.reshape(B, S, H, 128).transpose(1, 2)
        q, k, v = sp(qkv[:, :D]), sp(qkv[:, D:2 * D]), sp(qkv[:, 2 * D:])
        s = (q @ k.transpose(-1, -2)) / math.sqrt(128)
        s = s.masked_fill(~torch.ones(S, S, dtype=torch.bool).tril(), float("-inf"))
        return (s.softmax(-1) @ v).transpose(1, 2).reshape(B * S, D)

    def lm_attn_fwd(qkv, B, S, H, *, slopes=None, causal=True, scale=None):
        assert slopes is None and causal
        return _attn(qkv, B, S, H), torch.zeros(B, H, S)

    def lm_attn_bwd(dout, qkv, out, lse, B, S, H, *, slopes=None, causal=True, scale=None):
        q = qkv.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            return torch.autograd.grad(_attn(q, B, S, H), q, dout)[0]

    def swiglu_bwd(dh, g, u):
        gr, ur = g.detach().requires_grad_(True), u.detach().requires_grad_(True)
        with torch.enable_grad():
            return torch.autograd.grad(torch.nn.functional.silu(gr) * ur, (gr, ur), dh)

    for name, fn in dict(rmsnorm_fwd=rmsnorm_fwd, rmsnorm_bwd=rmsnorm_bwd, linear_fwd=linear_fwd, linear_dgrad=linear_dgrad,
                         rope128_=rope128_, lm_attn_fwd=lm_attn_fwd, lm_attn_bwd=lm_attn_bwd,
                         swiglu_fwd=lambda g, u: torch.nn.functional.silu(g) * u, swiglu_bwd=swiglu_bwd).items():
        monkeypatch.setattr(F, name, fn)
    monkeypatch.setattr(P, "bf16_of", lambda p: p.detach().float())
    monkeypatch.setattr(P, "f32_of", lambda p: p.detach().float())
    from otter_b200 import lm_llama
    monkeypatch.setattr(lm_llama, "BF16", torch.float32)          # keep the stand-in arithmetic in fp32


def test_frozen_llama_layer_wiring_matches_hf(torch_kernels):
    from transformers import LlamaConfig
    from transformers.models.llama import modeling_llama as M
    from otter_b200.lm_llama import FrozenLlamaDecoderLayer
    B, S, D, H = 2, 21, 256, 2
    cfg = LlamaConfig(hidden_size=D, num_attention_heads=H, num_key_value_heads=H, intermediate_size=3 * D, num_hidden_layers=1,
                      vocab_size=32, max_position_embeddings=128, rms_norm_eps=1e-6)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    hf = M.LlamaDecoderLayer(cfg, layer_idx=0)
    with torch.no_grad():
        hf.input_layernorm.weight.add_(0.1 * torch.randn(D))
        hf.post_attention_layernorm.weight.add_(0.1 * torch.randn(D))
    x = torch.randn(B, S, D, requires_grad=True)
    pos = torch.arange(S)[None].expand(B, S)
    mask = torch.full((S, S), float("-inf")).triu(1)[None, None].expand(B, 1, S, S)
    out = hf(x, attention_mask=mask, position_ids=pos, position_embeddings=M.LlamaRotaryEmbedding(cfg)(x, pos))
    out = out[0] if isinstance(out, tuple) else out
    w = torch.randn_like(out)
    (out * w).sum().backward()

    mine = FrozenLlamaDecoderLayer(D, H, 3 * D)
    mine.load_state_dict(hf.state_dict(), strict=True)
    x2 = x.detach().clone().requires_grad_(True)
    y = mine(x2, position_ids=pos[:1])
    (y * w).sum().backward()
    assert torch.allclose(y, out, rtol=1e-4, atol=1e-5)
    assert torch.allclose(x2.grad, x.grad, rtol=1e-3, atol=1e-5)
    assert all(p.grad is None for p in mine.parameters())          # frozen: activation gradient only
    with pytest.raises(NotImplementedError):
        mine(x2, position_ids=pos[:1] + 1)
