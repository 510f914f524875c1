"""CPU: the LLaMA-layer oracle is pinned against HF transformers' LlamaDecoderLayer (what the reference instantiates when
xformers is absent, modeling_otter.py:52-54) — forward and input gradient, fp32."""
import pytest
import torch

from oracle import restatement_llama as RL


@pytest.mark.parametrize("B,S,D,H", [(2, 37, 256, 2), (1, 130, 512, 4)])
def test_oracle_matches_hf_llama_layer(B, S, D, H):
    from transformers import LlamaConfig
    from transformers.models.llama import modeling_llama as M
    cfg = LlamaConfig(hidden_size=D, num_attention_heads=H, num_key_value_heads=H, intermediate_size=3 * D, num_hidden_layers=1,
                      vocab_size=32, max_position_embeddings=512, rms_norm_eps=1e-6)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    layer = M.LlamaDecoderLayer(cfg, layer_idx=0)
    with torch.no_grad():
        for n, p_ in layer.named_parameters():
            if p_.dim() == 1:
                p_.add_(0.1 * torch.randn_like(p_))
    rope = M.LlamaRotaryEmbedding(cfg)
    x = torch.randn(B, S, D, requires_grad=True)
    pos = torch.arange(S)[None].expand(B, S)
    mask = torch.full((S, S), float("-inf")).triu(1)[None, None].expand(B, 1, S, S)
    out = layer(x, attention_mask=mask, position_ids=pos, position_embeddings=rope(x, pos))
    out = out[0] if isinstance(out, tuple) else out
    w = torch.randn_like(out)
    (out * w).sum().backward()
    rp = getattr(cfg, "rope_parameters", None) or {}
    theta = rp.get("rope_theta", getattr(cfg, "rope_theta", 10000.0))
    p = {k: v.detach().clone() for k, v in layer.state_dict().items()}
    xr = x.detach().clone().requires_grad_(True)
    ref = RL.llama_layer(xr, p, H, eps=cfg.rms_norm_eps, rope_theta=theta)
    (ref * w).sum().backward()
    assert torch.allclose(ref, out, rtol=1e-4, atol=1e-5)
    assert torch.allclose(xr.grad, x.grad, rtol=1e-3, atol=1e-5)
