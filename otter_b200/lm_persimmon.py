"""Persimmon decoder layer (the LM of OtterHD / Fuyu-8B) on the otter_b200 kernels — SURVEY.md §8f rank 3.

Mirrors `PersimmonDecoderLayer` / `PersimmonAttention` / `PersimmonMLP`
(/root/reference/src/otter_ai/models/fuyu/modeling_persimmon.py:170-193,196-319,322-400): same sub-module and
parameter names (`input_layernorm`, `self_attn.{query_key_value, dense, q_layernorm, k_layernorm}`,
`post_attention_layernorm`, `mlp.{dense_h_to_4h, dense_4h_to_h}`), same forward() arguments and return tuple.  The
reference swaps HF's ops for flash-attn's fused CUDA ops (fused_layer_norm, fused_apply_rotary_emb, flash_attn_func,
fused_mlp_func "sqrelu"); the math is HF Persimmon's.  OtterHD fine-tunes the WHOLE model, so unlike the frozen MPT
layer this one returns weight gradients too.

  LN -> GEMM(+bias) -> [split + qk-LayerNorm + partial RoPE: one kernel] -> fused causal attention (head dim 64)
     -> GEMM(+bias, +residual) -> LN -> GEMM(+bias, relu^2, pre-activation kept) -> GEMM(+bias, +residual)
Backward: the same GEMM kernels for dgrad / wgrad (relu^2' in the dgrad epilogue), bias gradients as deterministic
column sums.  head_dim must be 64; dropouts 0; training path (no KV cache).
"""
import math

import torch
from torch import nn

from . import functional as F
from .params import GradSink, bf16_of, f32_of

BF16 = torch.bfloat16


class _PersimmonLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, layer, B, S, *params):          # params: layer.parameters(), so autograd tracks them
        a, m = layer.self_attn, layer.mlp
        D, H = layer.hidden_size, a.num_heads
        x2 = x.reshape(B * S, D)
        h1, mean1, rstd1 = F.layernorm_fwd(x2, f32_of(layer.input_layernorm.weight), f32_of(layer.input_layernorm.bias),
                                           layer.input_layernorm.eps)                                   # :372
        fused = F.linear_fwd(h1, bf16_of(a.query_key_value.weight), bias=f32_of(a.query_key_value.bias))  # :278
        qkv, stats = F.qkln_rope_fwd(fused, H, S, f32_of(a.q_layernorm.weight), f32_of(a.q_layernorm.bias),
                                     f32_of(a.k_layernorm.weight), f32_of(a.k_layernorm.bias), a.rotary_ndims,
                                     a.rope_theta, a.q_layernorm.eps)                                   # :281-303
        spec = F.AttnSpec(qkv, 0, qkv, D, 2 * D, B, H, S, S, 1.0 / math.sqrt(64), causal=True)
        o, lse = F.attn_fwd(spec)                                                                        # :304 causal
        x1 = F.linear_fwd(o, bf16_of(a.dense.weight), bias=f32_of(a.dense.bias), residual=x2)            # :308,383
        h2, mean2, rstd2 = F.layernorm_fwd(x1, f32_of(layer.post_attention_layernorm.weight),
                                           f32_of(layer.post_attention_layernorm.bias),
                                           layer.post_attention_layernorm.eps)                           # :387
        z = torch.empty((B * S, m.dense_h_to_4h.weight.shape[0]), device=x.device, dtype=BF16)
        act = F.linear_fwd(h2, bf16_of(m.dense_h_to_4h.weight), bias=f32_of(m.dense_h_to_4h.bias), act=3, aux_out=z)
        y = F.linear_fwd(act, bf16_of(m.dense_4h_to_h.weight), bias=f32_of(m.dense_4h_to_h.bias), residual=x1)  # :388-391
        ctx.layer, ctx.B, ctx.S = layer, B, S
        ctx.save_for_backward(x2, mean1, rstd1, h1, fused, stats, qkv, o, lse, x1, mean2, rstd2, h2, z, act)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        layer, B, S = ctx.layer, ctx.B, ctx.S
        x2, mean1, rstd1, h1, fused, stats, qkv, o, lse, x1, mean2, rstd2, h2, z, act = ctx.saved_tensors
        a, m = layer.self_attn, layer.mlp
        D, H = layer.hidden_size, a.num_heads
        sink = GradSink()
        dy2 = dy.reshape(B * S, D)
        if dy2.dtype != BF16 or not dy2.is_contiguous():
            dy2 = dy2.to(BF16).contiguous()

        def lin_bwd(lin, dout, inp, **dgrad_epi):
            """wgrad + bias grad into the sinks, returns the input gradient."""
            g, acc = sink.target(lin.weight)
            F.linear_wgrad(dout, inp, out=g, accumulate=acc)
            gb, accb = sink.target(lin.bias)
            F.grouped_colsum(dout, 1, 1, out=gb.view(1, -1), accumulate=accb)
            return F.linear_dgrad(dout, bf16_of(lin.weight), **dgrad_epi)

        def ln_bwd(ln, dout, inp, mean, rstd, add):
            gw, acc = sink.target(ln.weight)
            gb, _ = sink.target(ln.bias)
            dx, _, _ = F.layernorm_bwd(dout, inp, mean, rstd, f32_of(ln.weight), add=add, dgamma=gw, dbeta=gb, accumulate=acc)
            return dx

        dz = lin_bwd(m.dense_4h_to_h, dy2, act, aux_in=z, act=3)            # (dy W2) * 2 relu(z)
        dh2 = lin_bwd(m.dense_h_to_4h, dz, h2)
        dx1 = ln_bwd(layer.post_attention_layernorm, dh2, x1, mean2, rstd2, dy2)
        do = lin_bwd(a.dense, dx1, o)
        spec = F.AttnSpec(qkv, 0, qkv, D, 2 * D, B, H, S, S, 1.0 / math.sqrt(64), causal=True)
        dqkv = torch.empty_like(qkv)
        F.attn_bwd(spec, o, 0, lse, do, 0, dqkv, 0, dqkv, D, 2 * D)
        gqw, acc = sink.target(a.q_layernorm.weight)
        gqb, _ = sink.target(a.q_layernorm.bias)
        gkw, _ = sink.target(a.k_layernorm.weight)
        gkb, _ = sink.target(a.k_layernorm.bias)
        dfused = F.qkln_rope_bwd(dqkv, fused, stats, H, S, f32_of(a.q_layernorm.weight), f32_of(a.k_layernorm.weight),
                                 a.rotary_ndims, a.rope_theta, gqw, gqb, gkw, gkb, accumulate=acc)
        dh1 = lin_bwd(a.query_key_value, dfused, h1)
        dx = ln_bwd(layer.input_layernorm, dh1, x2, mean1, rstd1, dx1)
        return (dx.view(dy.shape), None, None, None) + tuple(sink.result(p) for p in layer.parameters())


class PersimmonMLP(nn.Module):
    """modeling_persimmon.py:170-193 (relu^2 between the two Linears, both with bias)."""

    def __init__(self, hidden_size, intermediate_size):
        super().__init__()
        self.dense_h_to_4h = nn.Linear(hidden_size, intermediate_size)
        self.dense_4h_to_h = nn.Linear(intermediate_size, hidden_size)


class PersimmonAttention(nn.Module):
    """modeling_persimmon.py:196-319 (parameters only; the arithmetic is in _PersimmonLayerFn)."""

    def __init__(self, hidden_size, num_heads, layer_norm_eps=1e-5, partial_rotary_factor=0.5, rope_theta=25000.0,
                 qk_layernorm=True):
        super().__init__()
        self.hidden_size, self.num_heads = hidden_size, num_heads
        self.head_dim = hidden_size // num_heads
        if self.head_dim * num_heads != hidden_size:
            raise ValueError(f"hidden_size must be divisible by num_heads (got `hidden_size`: {hidden_size} and "
                             f"`num_heads`: {num_heads}).")                   # :215-216
        if self.head_dim != 64:
            raise ValueError("otter_b200 Persimmon attention is built for head_dim == 64 (Persimmon-8B: 4096 / 64)")
        if not qk_layernorm:
            raise NotImplementedError("qk_layernorm=False is not built (Persimmon-8B / Fuyu-8B ship with it on)")
        self.rope_theta, self.partial_rotary_factor = float(rope_theta), partial_rotary_factor
        self.rotary_ndims = int(partial_rotary_factor * self.head_dim)
        self.query_key_value = nn.Linear(hidden_size, 3 * hidden_size, bias=True)
        self.dense = nn.Linear(hidden_size, hidden_size, bias=True)
        self.qk_layernorm = True
        self.q_layernorm = nn.LayerNorm(self.head_dim, eps=layer_norm_eps, elementwise_affine=True)
        self.k_layernorm = nn.LayerNorm(self.head_dim, eps=layer_norm_eps, elementwise_affine=True)


class PersimmonDecoderLayer(nn.Module):
    """modeling_persimmon.py:322-400.  Construct from a HF PersimmonConfig (`config=`) or explicit sizes."""

    def __init__(self, config=None, *, hidden_size=None, num_attention_heads=None, intermediate_size=None,
                 layer_norm_eps=1e-5, partial_rotary_factor=0.5, rope_theta=25000.0, qk_layernorm=True):
        super().__init__()
        if config is not None:
            hidden_size, num_attention_heads = config.hidden_size, config.num_attention_heads
            intermediate_size, layer_norm_eps = config.intermediate_size, config.layer_norm_eps
            qk_layernorm = getattr(config, "qk_layernorm", True)
            rp = getattr(config, "rope_parameters", None) or {}
            partial_rotary_factor = getattr(config, "partial_rotary_factor", rp.get("partial_rotary_factor", 0.5))
            rope_theta = getattr(config, "rope_theta", rp.get("rope_theta", 25000.0))
            if getattr(config, "hidden_dropout", 0.0) or getattr(config, "attention_dropout", 0.0):
                raise NotImplementedError("dropout probabilities must be 0")
            if getattr(config, "hidden_act", "relu2") != "relu2":
                raise NotImplementedError("hidden_act must be relu2")
        self.hidden_size = hidden_size
        self.self_attn = PersimmonAttention(hidden_size, num_attention_heads, layer_norm_eps, partial_rotary_factor,
                                            rope_theta, qk_layernorm)
        self.mlp = PersimmonMLP(hidden_size, intermediate_size)
        self.input_layernorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.post_attention_layernorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None,
                output_attentions=False, use_cache=False):
        assert past_key_value is None                                           # :275 (the reference's own assert)
        if output_attentions:
            raise NotImplementedError("output_attentions is not available from the fused attention kernel")
        if position_ids is not None:
            S_ = hidden_states.shape[1]
            want = torch.arange(S_, device=position_ids.device)
            if position_ids.shape[-1] != S_ or not bool((position_ids.reshape(-1, S_) == want).all()):
                raise NotImplementedError("only the default position_ids (arange) are supported")
        B, S, _ = hidden_states.shape
        xb = hidden_states if hidden_states.dtype == BF16 else hidden_states.to(BF16)
        y = _PersimmonLayerFn.apply(xb.contiguous(), self, B, S, *self.parameters())
        outputs = (y.to(hidden_states.dtype),)
        if use_cache:
            outputs += (None,)
        return outputs
