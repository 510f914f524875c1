"""TEST INFRASTRUCTURE — CPU restatement (oracle) of one LLaMA decoder layer
(/root/reference/xformers_model/llama.py:74-89,150-185,186-320; the reference instantiates HF's identical classes when
xformers is absent, modeling_otter.py:52-54).  Pinned against `transformers.models.llama.modeling_llama.LlamaDecoderLayer`
(transformers 5.5.0) in tests/test_oracle_llama_cpu.py.  Parameters: dict keyed by the layer's state-dict names."""
import math

import torch


def rms_norm(x, w, eps):
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w               # llama.py:83-89


def rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)                                             # :150-154


def llama_layer(x, p, num_heads, eps=1e-6, rope_theta=10000.0):
    B, S, D = x.shape
    hd = D // num_heads
    h = rms_norm(x, p["input_layernorm.weight"], eps)                               # :296
    sp = lambda t: t.view(B, S, num_heads, hd).transpose(1, 2)
    q, k, v = (sp(h @ p[f"self_attn.{n}_proj.weight"].t()) for n in ("q", "k", "v"))   # :222-224
    inv_freq = 1.0 / (rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    fr = torch.arange(S, dtype=torch.float32)[:, None] * inv_freq[None]
    emb = torch.cat((fr, fr), -1)
    cos, sin = emb.cos()[None, None], emb.sin()[None, None]
    q, k = q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin           # :157-166
    sim = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    sim = sim.masked_fill(~torch.ones(S, S, dtype=torch.bool).tril(), float("-inf"))   # LowerTriangularMask :241-246
    o = (sim.softmax(-1) @ v).transpose(1, 2).reshape(B, S, D)
    x1 = x + o @ p["self_attn.o_proj.weight"].t()                                   # :254,304
    h2 = rms_norm(x1, p["post_attention_layernorm.weight"], eps)                    # :308
    g, u = h2 @ p["mlp.gate_proj.weight"].t(), h2 @ p["mlp.up_proj.weight"].t()
    return x1 + (torch.nn.functional.silu(g) * u) @ p["mlp.down_proj.weight"].t()   # :184,309-310
