"""Frozen LLaMA decoder layer on the otter_b200 kernels — SURVEY.md §8f rank 1 (the LM of OTTER-Video-LLaMA7B, config c3).

Mirrors `LlamaDecoderLayer` (/root/reference/xformers_model/llama.py:262-320; the reference falls back to HF's identical
`transformers.models.llama` classes when xformers is absent, modeling_otter.py:52-54): same sub-module and parameter names
(`input_layernorm`, `self_attn.{q,k,v,o}_proj`, `post_attention_layernorm`, `mlp.{gate,up,down}_proj`), no biases.
The layer is FROZEN in Otter (modeling_otter.py:897-905): backward produces the input gradient only.

  RMSNorm -> q / k / v GEMMs into one [rows, 3D] buffer -> rotary embedding in place (rotate_half, head dim 128)
     -> causal attention (otb_lm_attn_*, no key bias) -> o_proj GEMM (+residual)
     -> RMSNorm -> gate / up GEMMs -> silu(g) * u -> down GEMM (+residual)
head_dim must be 128 (LLaMA-7B: 4096 = 32 x 128); training path (no KV cache); default position ids.
`swap_llama_layers(model)` replaces the decoder layers of an HF `LlamaForCausalLM` by this class, sharing the Parameter
objects.  Attention masks: the kernel applies the causal mask only, which equals HF's causal + padding mask at every
non-pad position when padding is on the RIGHT (the training collate); a forward pre-hook on the LlamaModel
checks the 2-D mask once per forward and raises on left padding.  Incremental decoding with a KV cache (generation — out of
the hot path, SURVEY.md §8 "out of scope") is handed to the original HF layer, which is what the reference itself runs.
"""
import math

import torch
from torch import nn

from . import functional as F
from . import params as P

BF16 = torch.bfloat16


class _FrozenLlamaLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, layer, B, S):
        a, m = layer.self_attn, layer.mlp
        D, H = layer.hidden_size, layer.num_heads
        x2 = x.reshape(B * S, D)
        w_in, w_post = P.f32_of(layer.input_layernorm.weight), P.f32_of(layer.post_attention_layernorm.weight)
        h1, r1 = F.rmsnorm_fwd(x2, w_in, layer.eps)                                        # llama.py:296
        qkv = torch.empty((B * S, 3 * D), device=x.device, dtype=BF16)
        F.linear_fwd(h1, P.bf16_of(a.q_proj.weight), out=qkv[:, :D])                       # :222-224
        F.linear_fwd(h1, P.bf16_of(a.k_proj.weight), out=qkv[:, D:2 * D])
        F.linear_fwd(h1, P.bf16_of(a.v_proj.weight), out=qkv[:, 2 * D:])
        F.rope128_(qkv, H, S, 2, layer.rope_theta)                                         # :229-230
        o, lse = F.lm_attn_fwd(qkv, B, S, H, slopes=None, causal=True, scale=1.0 / math.sqrt(128))   # :241-246
        x1 = F.linear_fwd(o, P.bf16_of(a.o_proj.weight), residual=x2)                      # :254,304
        h2, r2 = F.rmsnorm_fwd(x1, w_post, layer.eps)                                      # :308
        g = F.linear_fwd(h2, P.bf16_of(m.gate_proj.weight))
        u = F.linear_fwd(h2, P.bf16_of(m.up_proj.weight))
        hh = F.swiglu_fwd(g, u)                                                            # :184
        y = F.linear_fwd(hh, P.bf16_of(m.down_proj.weight), residual=x1)                   # :309-310
        ctx.layer, ctx.B, ctx.S = layer, B, S
        ctx.save_for_backward(x2, r1, qkv, o, lse, x1, r2, g, u)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        layer, B, S = ctx.layer, ctx.B, ctx.S
        x2, r1, qkv, o, lse, x1, r2, g, u = ctx.saved_tensors
        a, m = layer.self_attn, layer.mlp
        D, H = layer.hidden_size, layer.num_heads
        dy2 = dy.reshape(B * S, D)
        if dy2.dtype != BF16 or not dy2.is_contiguous():
            dy2 = dy2.to(BF16).contiguous()
        dh = F.linear_dgrad(dy2, P.bf16_of(m.down_proj.weight))
        dg, du = F.swiglu_bwd(dh, g, u)
        dh2 = F.linear_dgrad(dg, P.bf16_of(m.gate_proj.weight))
        dh2 = F.linear_dgrad(du, P.bf16_of(m.up_proj.weight), residual=dh2)
        dx1 = F.rmsnorm_bwd(dh2, x1, r2, P.f32_of(layer.post_attention_layernorm.weight), add=dy2)
        do = F.linear_dgrad(dx1, P.bf16_of(a.o_proj.weight))
        dqkv = F.lm_attn_bwd(do, qkv, o, lse, B, S, H, slopes=None, causal=True, scale=1.0 / math.sqrt(128))
        F.rope128_(dqkv, H, S, 2, layer.rope_theta, backward=True)
        dh1 = F.linear_dgrad(dqkv[:, :D], P.bf16_of(a.q_proj.weight))
        dh1 = F.linear_dgrad(dqkv[:, D:2 * D], P.bf16_of(a.k_proj.weight), residual=dh1)
        dh1 = F.linear_dgrad(dqkv[:, 2 * D:], P.bf16_of(a.v_proj.weight), residual=dh1)
        dx = F.rmsnorm_bwd(dh1, x2, r1, P.f32_of(layer.input_layernorm.weight), add=dx1)
        return dx.view(dy.shape), None, None, None


class _RMSNormParams(nn.Module):
    def __init__(self, d, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.variance_epsilon = eps


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q_proj, self.k_proj = nn.Linear(d, d, bias=False), nn.Linear(d, d, bias=False)
        self.v_proj, self.o_proj = nn.Linear(d, d, bias=False), nn.Linear(d, d, bias=False)


class _MLP(nn.Module):
    def __init__(self, d, i):
        super().__init__()
        self.gate_proj, self.up_proj = nn.Linear(d, i, bias=False), nn.Linear(d, i, bias=False)
        self.down_proj = nn.Linear(i, d, bias=False)


class FrozenLlamaDecoderLayer(nn.Module):
    def __init__(self, hidden_size, num_attention_heads, intermediate_size, rms_norm_eps=1e-6, rope_theta=10000.0):
        super().__init__()
        if hidden_size != num_attention_heads * 128:
            raise ValueError("FrozenLlamaDecoderLayer: head_dim must be 128 (LLaMA-7B: 4096 = 32 x 128)")
        self.hidden_size, self.num_heads, self.eps, self.rope_theta = hidden_size, num_attention_heads, rms_norm_eps, rope_theta
        self.self_attn = _Attn(hidden_size)
        self.mlp = _MLP(hidden_size, intermediate_size)
        self.input_layernorm = _RMSNormParams(hidden_size, rms_norm_eps)
        self.post_attention_layernorm = _RMSNormParams(hidden_size, rms_norm_eps)
        self.requires_grad_(False)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None, past_key_values=None,
                output_attentions=False, use_cache=False, **kwargs):
        if past_key_value is not None or past_key_values is not None:
            hf = self.__dict__.get("_hf_layer")
            if hf is None:
                raise NotImplementedError("FrozenLlamaDecoderLayer: training path only (no KV cache)")
            kw = dict(kwargs, attention_mask=attention_mask, position_ids=position_ids, use_cache=use_cache)
            kw["past_key_values" if past_key_values is not None else "past_key_value"] = \
                past_key_values if past_key_values is not None else past_key_value
            return hf(hidden_states, **kw)
        if output_attentions:
            raise NotImplementedError("output_attentions is not available from the fused attention kernel")
        B, S, _ = hidden_states.shape
        if position_ids is not None:
            want = torch.arange(S, device=position_ids.device)
            if position_ids.shape[-1] != S or not bool((position_ids.reshape(-1, S) == want).all()):
                raise NotImplementedError("only the default position_ids (arange) are supported")
        xb = hidden_states if hidden_states.dtype == BF16 else hidden_states.to(BF16)
        y = _FrozenLlamaLayerFn.apply(xb.contiguous(), self, B, S)
        return y.to(hidden_states.dtype)          # HF 5.x decoder layers return the tensor


def _share_parameters(new, old):
    for name, _ in list(new.named_parameters()):
        mod_path, _, leaf = name.rpartition(".")
        src, dst = old.get_submodule(mod_path), new.get_submodule(mod_path)
        p = getattr(src, leaf)
        if tuple(p.shape) != tuple(getattr(dst, leaf).shape):
            raise ValueError(f"{name}: shape {tuple(p.shape)} does not fit FrozenLlamaDecoderLayer")
        dst._parameters[leaf] = p          # the SAME Parameter object: .to() / optimizer / checkpoint see one tensor


def _right_padding_check(module, args, kwargs):
    mask = kwargs.get("attention_mask")
    if mask is not None and mask.dim() == 2 and mask.shape[1] > 1:
        m = mask.to(torch.bool)
        if bool((m[:, 1:] & ~m[:, :-1]).any()):
            raise NotImplementedError("FrozenLlamaDecoderLayer applies the causal mask only: attention_mask must be "
                                      "right-padded (ones then zeros per row)")
    elif mask is not None and not isinstance(mask, dict) and mask.dim() != 2:
        raise NotImplementedError("FrozenLlamaDecoderLayer: pass the 2-D padding mask, not a prepared 4-D mask")
    return None


def swappable(cfg):
    """True when the HF LlamaConfig describes layers this module covers (MHA, head_dim 128, default RoPE, no biases)."""
    rp = getattr(cfg, "rope_parameters", None) or {}
    rope_type = rp.get("rope_type", "default") if isinstance(rp, dict) else "default"
    hd = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
    return (hd == 128 and cfg.hidden_size == cfg.num_attention_heads * 128
            and getattr(cfg, "num_key_value_heads", cfg.num_attention_heads) == cfg.num_attention_heads
            and rope_type == "default" and not getattr(cfg, "attention_bias", False) and not getattr(cfg, "mlp_bias", False)
            and getattr(cfg, "hidden_act", "silu") == "silu" and cfg.hidden_size % 64 == 0 and cfg.intermediate_size % 64 == 0)


def swap_llama_layers(llama_model, keep_hf_for_cache=True):
    """Replace the decoder layers of an HF LlamaForCausalLM / LlamaModel by FrozenLlamaDecoderLayer (Parameters shared)."""
    model = llama_model.model if hasattr(llama_model, "model") else llama_model
    cfg = model.config
    if not swappable(cfg):
        raise ValueError("swap_llama_layers: this LlamaConfig is outside FrozenLlamaDecoderLayer's coverage (see swappable())")
    rp = getattr(cfg, "rope_parameters", None) or {}
    theta = rp.get("rope_theta", getattr(cfg, "rope_theta", 10000.0))
    for i, old in enumerate(model.layers):
        new = FrozenLlamaDecoderLayer(cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size, cfg.rms_norm_eps, theta)
        _share_parameters(new, old)
        new.requires_grad_(False)
        if keep_hf_for_cache:
            new.__dict__["_hf_layer"] = old            # not a registered sub-module: no duplicate state-dict keys
        model.layers[i] = new
    if not getattr(model, "_otb_mask_hook", False):
        model.register_forward_pre_hook(_right_padding_check, with_kwargs=True)
        model._otb_mask_hook = True
    return llama_model
