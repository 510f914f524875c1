"""GPU: frozen LLaMA decoder layer (SURVEY.md §8f rank 1, config c3's LM) — RMSNorm / rotary / SwiGLU kernels against plain
torch and the whole layer (forward + input gradient) against the CPU oracle pinned to HF's LlamaDecoderLayer."""
import pytest
import torch

from oracle import restatement_llama as RL

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF16 = torch.bfloat16


def _rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-20)).item()


def test_rmsnorm_rope_swiglu_kernels():
    from otter_b200 import functional as F
    g = torch.Generator().manual_seed(0)
    rows, D = 75, 512
    x = (torch.randn(rows, D, generator=g) * 2).to(BF16)
    w = 1 + 0.2 * torch.randn(D, generator=g)
    y, rstd = F.rmsnorm_fwd(x.to(DEV), w.to(DEV), 1e-6)
    xr = x.float().requires_grad_(True)
    yr = RL.rms_norm(xr, w, 1e-6)
    assert _rel(y, yr) <= 5e-3
    dy, add = torch.randn(rows, D, generator=g).to(BF16), torch.randn(rows, D, generator=g).to(BF16)
    yr.backward(dy.float())
    dx = F.rmsnorm_bwd(dy.to(DEV), x.to(DEV), rstd, w.to(DEV), add=add.to(DEV))
    assert _rel(dx, xr.grad + add.float()) <= 5e-3
    # rotary embedding, 2 of 3 blocks, H = 2, S = 25 (rows = 3 sequences)
    H, S = 2, 25
    buf = torch.randn(rows, 3 * H * 128, generator=g).to(BF16)
    out = F.rope128_(buf.clone().to(DEV), H, S, 2, 10000.0)
    t = buf.float().view(3, S, 3, H, 128)
    inv = 1.0 / (10000.0 ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))
    fr = torch.arange(S, dtype=torch.float32)[:, None] * inv[None]
    emb = torch.cat((fr, fr), -1)
    cos, sin = emb.cos()[None, :, None, None], emb.sin()[None, :, None, None]
    ref = t.clone()
    ref[:, :, :2] = t[:, :, :2] * cos + RL.rotate_half(t[:, :, :2]) * sin
    assert _rel(out, ref.reshape(rows, -1)) <= 5e-3
    assert torch.equal(out[:, 2 * H * 128:].cpu(), buf[:, 2 * H * 128:])                 # v untouched
    back = F.rope128_(out.clone(), H, S, 2, 10000.0, backward=True)                      # R^T R = I
    assert _rel(back, buf) <= 1e-2
    # SwiGLU
    I = 384
    gg, uu, dh = (torch.randn(rows, I, generator=g).to(BF16) for _ in range(3))
    h = F.swiglu_fwd(gg.to(DEV), uu.to(DEV))
    gr, ur = gg.float().requires_grad_(True), uu.float().requires_grad_(True)
    hr = torch.nn.functional.silu(gr) * ur
    assert _rel(h, hr) <= 5e-3
    hr.backward(dh.float())
    dg, du = F.swiglu_bwd(dh.to(DEV), gg.to(DEV), uu.to(DEV))
    assert _rel(dg, gr.grad) <= 5e-3 and _rel(du, ur.grad) <= 5e-3


@pytest.mark.parametrize("B,S,D,H", [(2, 100, 256, 2), (1, 300, 512, 4)])
def test_frozen_llama_layer_vs_oracle(B, S, D, H):
    from otter_b200.lm_llama import FrozenLlamaDecoderLayer
    torch.manual_seed(S)
    layer = FrozenLlamaDecoderLayer(D, H, 3 * D)
    with torch.no_grad():
        layer.input_layernorm.weight.add_(0.1 * torch.randn(D))
        layer.post_attention_layernorm.weight.add_(0.1 * torch.randn(D))
    p_ref = {k: v.detach().clone() for k, v in layer.state_dict().items()}
    layer.to(DEV)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, S, D, generator=g).to(BF16).float()
    w = torch.randn(B, S, D, generator=g).to(BF16).float()
    xg = x.to(DEV).requires_grad_(True)
    y = layer(xg, position_ids=torch.arange(S, device=DEV)[None])
    (y.float() * w.to(DEV)).sum().backward()
    xr = x.clone().requires_grad_(True)
    ref = RL.llama_layer(xr, p_ref, H)
    (ref * w).sum().backward()
    assert _rel(y, ref) <= 1.5e-2, _rel(y, ref)
    assert _rel(xg.grad, xr.grad) <= 3e-2, _rel(xg.grad, xr.grad)
    assert all(p.grad is None for p in layer.parameters())                                # frozen: activation grads only


def test_swap_llama_layers_in_an_hf_model():
    from transformers import LlamaConfig, LlamaForCausalLM
    from otter_b200.lm_llama import FrozenLlamaDecoderLayer, swap_llama_layers
    cfg = LlamaConfig(vocab_size=64, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=2, max_position_embeddings=128)
    torch.manual_seed(0)
    model = LlamaForCausalLM(cfg).to(DEV)
    ids = torch.randint(0, 64, (2, 40), device=DEV)
    mask = torch.ones_like(ids)
    mask[1, 30:] = 0                                                                      # right padding
    with torch.no_grad():
        ref = model(input_ids=ids, attention_mask=mask, use_cache=False).logits.float()
    keys = set(model.state_dict().keys())
    swap_llama_layers(model)
    assert set(model.state_dict().keys()) == keys                                         # checkpoint-compatible
    assert all(isinstance(l, FrozenLlamaDecoderLayer) for l in model.model.layers)
    with torch.no_grad():
        got = model(input_ids=ids, attention_mask=mask, use_cache=False).logits.float()
    valid = mask.bool()
    assert _rel(got[valid], ref[valid]) <= 2e-2, _rel(got[valid], ref[valid])
    # incremental decoding with a KV cache is handed to the HF layer the reference itself runs
    with torch.no_grad():
        out = model(input_ids=ids[:1], use_cache=True)
    assert out.past_key_values is not None and _rel(out.logits, ref[:1]) <= 2e-2
    left = mask.flip(1)
    with pytest.raises(NotImplementedError):
        model(input_ids=ids, attention_mask=left, use_cache=False)
