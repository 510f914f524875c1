"""TEST INFRASTRUCTURE — CPU restatement (oracle) of one Persimmon decoder layer, the LM layer of OtterHD / Fuyu
(/root/reference/src/otter_ai/models/fuyu/modeling_persimmon.py:170-193,266-319,322-400).

The reference file is HF's `modeling_persimmon.py` with four ops swapped for flash-attn CUDA extensions
(fused_layer_norm :283-285,372,387; fused_apply_rotary_emb :300-301; flash_attn_func(causal=True) :304;
fused_mlp_func "sqrelu" :187-193) — it cannot run on a CPU, and flash-attn's extensions are the third-party code:
flash_attn 2.8.3 is in this image, the reference pins none.  The algorithm is therefore restated here from HF's
published form and PINNED against `transformers.models.persimmon.modeling_persimmon.PersimmonDecoderLayer`
(transformers 5.5.0 in this image; the reference vendors 4.35's) in tests/test_oracle_persimmon_cpu.py.
Parameters: dict keyed by the layer's state-dict names.
"""
import math

import torch


def layer_norm(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def persimmon_layer(x, p, num_heads, rotary_ndims=32, rope_theta=25000.0, eps=1e-5, q=None):
    """x [B, S, D] -> [B, S, D].  `q`: optional rounding hook at the points the CUDA pipeline stores bf16."""
    q_ = q or (lambda t: t)
    B, S, D = x.shape
    hd = D // num_heads
    h = q_(layer_norm(x, p["input_layernorm.weight"], p["input_layernorm.bias"], eps))                      # :372
    fused = q_(h @ p["self_attn.query_key_value.weight"].t() + p["self_attn.query_key_value.bias"])         # :278
    fused = fused.view(B, S, num_heads, 3, hd)                                                              # :263-264
    qs, ks, vs = fused[..., 0, :], fused[..., 1, :], fused[..., 2, :]
    qs = layer_norm(qs, p["self_attn.q_layernorm.weight"], p["self_attn.q_layernorm.bias"], eps)            # :283-285
    ks = layer_norm(ks, p["self_attn.k_layernorm.weight"], p["self_attn.k_layernorm.bias"], eps)
    inv_freq = 1.0 / (rope_theta ** (torch.arange(0, rotary_ndims, 2, dtype=torch.float32) / rotary_ndims))
    freqs = torch.arange(S, dtype=torch.float32)[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos()[None, :, None, :], emb.sin()[None, :, None, :]                                     # :288-289
    q_rot, q_pass = qs[..., :rotary_ndims], qs[..., rotary_ndims:]                                          # :292-299
    k_rot, k_pass = ks[..., :rotary_ndims], ks[..., rotary_ndims:]
    qs = q_(torch.cat((q_rot * cos + rotate_half(q_rot) * sin, q_pass), dim=-1))                            # :300-303
    ks = q_(torch.cat((k_rot * cos + rotate_half(k_rot) * sin, k_pass), dim=-1))
    qh, kh, vh = (t.permute(0, 2, 1, 3) for t in (qs, ks, q_(vs)))
    sim = (qh @ kh.transpose(-1, -2)) / math.sqrt(hd)                                                       # :304 scale
    mask = torch.ones(S, S, dtype=torch.bool).tril()
    sim = sim.masked_fill(~mask, float("-inf"))                                                             # causal=True
    o = q_((sim.softmax(-1) @ vh).permute(0, 2, 1, 3).reshape(B, S, D))
    x1 = q_(o @ p["self_attn.dense.weight"].t() + p["self_attn.dense.bias"] + x)                            # :308,383
    h2 = q_(layer_norm(x1, p["post_attention_layernorm.weight"], p["post_attention_layernorm.bias"], eps))  # :387
    z = q_(h2 @ p["mlp.dense_h_to_4h.weight"].t() + p["mlp.dense_h_to_4h.bias"])
    a = q_(torch.relu(z) ** 2)                                                                              # "sqrelu"
    return q_(a @ p["mlp.dense_4h_to_h.weight"].t() + p["mlp.dense_4h_to_h.bias"] + x1)                     # :388-391
