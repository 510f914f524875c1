"""Frozen MPT decoder layer on the otter_b200 kernels (candidate, see R2_PREP.md).

Mirrors `MPTBlock` (/root/reference/src/otter_ai/models/mpt/blocks.py:23-88) — same sub-module and parameter names
(`norm_1`, `attn.Wqkv`, `attn.out_proj`, `norm_2`, `ffn.up_proj`, `ffn.down_proj`), same `forward()` signature and
return tuple — for the configuration OTTER-Image-MPT7B ships: multi-head attention, `attn_impl: torch` semantics
(mpt/attention.py:22-84), ALiBi key bias, causal mask, erf-GELU MLP, dropouts 0.  The layer is FROZEN in Otter
(modeling_otter.py:905 freezes the LM): backward produces the input gradient only — five dgrad GEMMs, the causal
attention backward and two LayerNorm backwards; no wgrad, and the bf16 weight copies are cast once.

head_dim must be 128 (MPT-7B / LLaMA-7B); there is no CPU path.
"""
import math

import torch
from torch import nn

from . import functional as F
from . import params as P


def alibi_slopes(n_heads, alibi_bias_max=8, device=None):
    """mpt/attention.py:447-454 gen_slopes -> fp32 [H]."""
    n2 = 2 ** math.ceil(math.log2(n_heads))
    m = torch.arange(1, n2 + 1, dtype=torch.float32, device=device) * (alibi_bias_max / n2)
    slopes = 1.0 / torch.pow(2, m)
    if n2 != n_heads:
        slopes = torch.cat([slopes[1::2], slopes[::2]])[:n_heads]
    return slopes.contiguous()


class _FrozenMPTBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, blk, B, S):
        D, H = blk.d_model, blk.n_heads
        x2 = x.reshape(B * S, D)
        g1, b1 = P.f32_of(blk.norm_1.weight), blk._beta(blk.norm_1)
        g2, b2 = P.f32_of(blk.norm_2.weight), blk._beta(blk.norm_2)
        w_qkv, w_out = P.bf16_of(blk.attn.Wqkv.weight), P.bf16_of(blk.attn.out_proj.weight)
        w_up, w_down = P.bf16_of(blk.ffn.up_proj.weight), P.bf16_of(blk.ffn.down_proj.weight)

        a, mean1, rstd1 = F.layernorm_fwd(x2, g1, b1)
        qkv = F.linear_fwd(a, w_qkv, bias=blk._bias(blk.attn.Wqkv))
        o, lse = F.lm_attn_fwd(qkv, B, S, H, slopes=blk.slopes, causal=True, scale=blk.softmax_scale)
        x1 = F.linear_fwd(o, w_out, bias=blk._bias(blk.attn.out_proj), residual=x2)
        m, mean2, rstd2 = F.layernorm_fwd(x1, g2, b2)
        z = torch.empty((B * S, w_up.shape[0]), device=x.device, dtype=torch.bfloat16)
        h = F.linear_fwd(m, w_up, bias=blk._bias(blk.ffn.up_proj), act=1, aux_out=z)
        y = F.linear_fwd(h, w_down, bias=blk._bias(blk.ffn.down_proj), residual=x1)
        ctx.blk, ctx.B, ctx.S = blk, B, S
        ctx.save_for_backward(x2, mean1, rstd1, qkv, o, lse, x1, mean2, rstd2, z)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        blk, B, S = ctx.blk, ctx.B, ctx.S
        x2, mean1, rstd1, qkv, o, lse, x1, mean2, rstd2, z = ctx.saved_tensors
        D, H = blk.d_model, blk.n_heads
        dy2 = dy.reshape(B * S, D)
        if dy2.dtype != torch.bfloat16 or dy2.stride(-1) != 1:
            dy2 = dy2.to(torch.bfloat16).contiguous()
        g1, g2 = P.f32_of(blk.norm_1.weight), P.f32_of(blk.norm_2.weight)
        w_qkv, w_out = P.bf16_of(blk.attn.Wqkv.weight), P.bf16_of(blk.attn.out_proj.weight)
        w_up, w_down = P.bf16_of(blk.ffn.up_proj.weight), P.bf16_of(blk.ffn.down_proj.weight)

        dz = F.linear_dgrad(dy2, w_down, aux_in=z)                       # (dy W_down) * gelu'(z)
        dm = F.linear_dgrad(dz, w_up)
        dx1, _, _ = F.layernorm_bwd(dm, x1, mean2, rstd2, g2, add=dy2, want_param_grads=False)   # + residual branch
        do = F.linear_dgrad(dx1, w_out)
        dqkv = F.lm_attn_bwd(do, qkv, o, lse, B, S, H, slopes=blk.slopes, causal=True, scale=blk.softmax_scale)
        da = F.linear_dgrad(dqkv, w_qkv)
        dx, _, _ = F.layernorm_bwd(da, x2, mean1, rstd1, g1, add=dx1, want_param_grads=False)
        return dx.view(dy.shape), None, None, None


class _Attn(nn.Module):
    def __init__(self, d_model, bias):
        super().__init__()
        self.Wqkv = nn.Linear(d_model, 3 * d_model, bias=bias)
        self.out_proj = nn.Linear(d_model, d_model, bias=bias)


class _MLP(nn.Module):
    def __init__(self, d_model, expansion_ratio, bias):
        super().__init__()
        self.up_proj = nn.Linear(d_model, expansion_ratio * d_model, bias=bias)
        self.down_proj = nn.Linear(expansion_ratio * d_model, d_model, bias=bias)


class FrozenMPTBlock(nn.Module):
    """Drop-in for a frozen `MPTBlock` (state-dict keys identical); trains nothing, propagates input gradients."""

    def __init__(self, d_model, n_heads, expansion_ratio=4, no_bias=True, alibi=True, alibi_bias_max=8,
                 softmax_scale=None):
        super().__init__()
        if d_model != n_heads * 128:
            raise ValueError("FrozenMPTBlock: head_dim must be 128 (MPT-7B: 4096 = 32 x 128)")
        self.d_model, self.n_heads = d_model, n_heads
        self.norm_1 = nn.LayerNorm(d_model, eps=1e-5, bias=not no_bias)
        self.attn = _Attn(d_model, bias=not no_bias)
        self.norm_2 = nn.LayerNorm(d_model, eps=1e-5, bias=not no_bias)
        self.ffn = _MLP(d_model, expansion_ratio, bias=not no_bias)
        self.alibi, self.alibi_bias_max = alibi, alibi_bias_max
        self.softmax_scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(128)
        self.slopes = None
        self._zeros = None
        self.requires_grad_(False)

    def _beta(self, ln):
        if ln.bias is not None:
            return P.f32_of(ln.bias)
        if self._zeros is None or self._zeros.device != ln.weight.device:
            self._zeros = torch.zeros(self.d_model, device=ln.weight.device, dtype=torch.float32)
        return self._zeros

    @staticmethod
    def _bias(lin):
        return P.f32_of(lin.bias) if lin.bias is not None else None

    def forward(self, x, past_key_value=None, attn_bias=None, attention_mask=None, is_causal=True):
        if past_key_value is not None or not is_causal:
            raise NotImplementedError("FrozenMPTBlock: training path only (causal, no KV cache)")
        if attention_mask is not None:
            # Right padding needs no key mask under a causal mask (a real token never sees a later pad key, and the
            # ALiBi key bias is a per-row constant shift away from the relative form); anything else is refused.
            am = attention_mask.to(torch.bool)
            if not bool((am[:, 1:] <= am[:, :-1]).all()):
                raise NotImplementedError("FrozenMPTBlock: only right-padded batches (no interior / left padding)")
        if self.alibi and (self.slopes is None or self.slopes.device != x.device):
            self.slopes = alibi_slopes(self.n_heads, self.alibi_bias_max, device=x.device)
        B, S, _ = x.shape
        xb = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
        y = _FrozenMPTBlockFn.apply(xb.contiguous(), self, B, S)
        return (y, None, past_key_value)
