"""CPU: the Persimmon-layer oracle (oracle/restatement_persimmon.py) is pinned against HF transformers' own
PersimmonDecoderLayer — the published form of the algorithm the reference file wraps in flash-attn ops
(fuyu/modeling_persimmon.py:266-319,322-400) — forward, input gradient and parameter gradients, fp32."""
import pytest
import torch

from oracle import restatement_persimmon as RP


def _hf_layer(D, H, I):
    from transformers import PersimmonConfig
    from transformers.models.persimmon import modeling_persimmon as M
    cfg = PersimmonConfig(hidden_size=D, num_attention_heads=H, intermediate_size=I, num_hidden_layers=1, vocab_size=32,
                          qk_layernorm=True, hidden_dropout=0.0, attention_dropout=0.0, layer_norm_eps=1e-5,
                          max_position_embeddings=512)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    layer = M.PersimmonDecoderLayer(cfg, layer_idx=0)
    with torch.no_grad():
        for n, p in layer.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    return cfg, layer, M.PersimmonRotaryEmbedding(cfg)


@pytest.mark.parametrize("B,S,D,H", [(2, 37, 128, 2), (1, 130, 256, 4)])
def test_oracle_matches_hf_persimmon_layer(B, S, D, H):
    cfg, layer, rope = _hf_layer(D, H, 4 * D)
    x = torch.randn(B, S, D, requires_grad=True)
    pos = torch.arange(S)[None].expand(B, S)
    mask = torch.full((S, S), float("-inf")).triu(1)[None, None].expand(B, 1, S, S)
    out = layer(x, attention_mask=mask, position_ids=pos, position_embeddings=rope(x, pos))
    out = out[0] if isinstance(out, tuple) else out
    w = torch.randn_like(out)
    (out * w).sum().backward()
    rp = getattr(cfg, "rope_parameters", None) or {}
    theta = rp.get("rope_theta", getattr(cfg, "rope_theta", 25000.0))
    prf = rp.get("partial_rotary_factor", getattr(cfg, "partial_rotary_factor", 0.5))
    p = {k: v.detach().clone().requires_grad_(True) for k, v in layer.state_dict().items()}
    xr = x.detach().clone().requires_grad_(True)
    ref = RP.persimmon_layer(xr, p, H, rotary_ndims=int(prf * (D // H)), rope_theta=theta, eps=cfg.layer_norm_eps)
    (ref * w).sum().backward()
    assert torch.allclose(ref, out, rtol=1e-4, atol=1e-5)
    assert torch.allclose(xr.grad, x.grad, rtol=1e-3, atol=1e-5)
    for k, v in layer.named_parameters():
        assert torch.allclose(p[k].grad, v.grad, rtol=1e-3, atol=2e-5), k
