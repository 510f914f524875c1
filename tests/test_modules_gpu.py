"""GPU parity of the drop-in modules against (1) golden outputs of the UNMODIFIED reference
(tests/golden/*.pt, made by oracle/make_golden.py) and (2) the CPU oracle (oracle/restatement.py).

Production numerics are bf16 operands / fp32 accumulate (the reference's autocast(bf16) recipe), compared
with the reference's fp32 results.  Tolerance for this mode, stated once here:
    relative Frobenius error <= 1.5e-2   and   max |err| <= 4e-2 * max |ref|   (outputs)
    gradient pins: relative error of the norm <= 3e-2, sampled entries within 6e-2 * max |sample| (+1e-6)
Index / mask tensors are compared bit-exact.
"""
import os

import pytest
import torch

from oracle import restatement as R
from oracle.ref_shims import FakeTokenizer
from oracle.seeded import load_seeded_, sample_flat, seeded_tensor

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def gold(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=False)


def check_out(got, ref, what, fro=1.5e-2, mx=4e-2):
    got, ref = got.detach().float().cpu(), ref.float()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    rel = (got - ref).norm().item() / max(ref.norm().item(), 1e-12)
    m = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
    assert rel <= fro and m <= mx, f"{what}: rel_fro={rel:.3e} max_rel={m:.3e}"


def check_grads(named_params, pins, what, norm_tol=3e-2, samp_tol=6e-2, gate_tol=None):
    """gate_tol: the two tanh-gate gradients are single scalars = sum(dy * a) over the whole activation, a sum
    with heavy sign cancellation, so bf16 rounding of dy / a shows up amplified by |sum|terms|| / |sum|."""
    named = dict(named_params)
    assert set(pins) <= set(named), (what, set(pins) - set(named))
    for k, pin in pins.items():
        g = named[k].grad
        assert g is not None, f"{what}: no grad for {k}"
        n = g.float().norm().item()
        nt = gate_tol if (gate_tol is not None and k.endswith("_gate")) else norm_tol
        assert abs(n - pin["norm"]) <= nt * pin["norm"] + 1e-7, f"{what}: |grad {k}| {n:.4e} vs {pin['norm']:.4e}"
        if gate_tol is not None and k.endswith("_gate"):
            continue
        s = sample_flat(g).cpu()
        tol = samp_tol * pin["sample"].abs().max().item() + 1e-6
        err = (s - pin["sample"]).abs().max().item()
        assert err <= tol, f"{what}: grad sample {k} err {err:.3e} tol {tol:.3e}"


def loss_of(out):
    return out.float().pow(2).mean()


def test_perceiver_block_golden():
    from otter_b200.modeling_otter import OtterPerceiverBlock
    g = gold("perceiver_block.pt")
    c = g["cfg"]
    blk = OtterPerceiverBlock(dim=c["dim"])
    load_seeded_(blk, g["seed"])
    blk.to(DEV)
    x = seeded_tensor("in.x", (c["b"], c["T"], c["n1"], c["dim"]), g["seed"], "randn").to(DEV)
    lat = seeded_tensor("in.latents", (c["b"], c["T"], c["n2"], c["dim"]), g["seed"], "randn").to(DEV).requires_grad_(True)
    out = blk(x, lat)
    assert out.dtype == torch.float32
    check_out(out, g["out"], "perceiver block out")
    loss_of(out).backward()
    check_out(lat.grad, g["dlat"], "perceiver block dlatents", fro=3e-2, mx=6e-2)
    check_grads(blk.named_parameters(), g["grads"], "perceiver block")


@pytest.mark.parametrize("tag", ["small", "image", "video"])
def test_resampler_golden(tag):
    from otter_b200.modeling_otter import OtterPerceiverResampler
    g = gold(f"resampler_{tag}.pt")
    rs = OtterPerceiverResampler(**g["cfg"])
    load_seeded_(rs, g["seed"], kinds={"latents": "randn", "frame_embs": "randn"})
    rs.to(DEV)
    x = seeded_tensor(f"in.resampler.{tag}", g["in_shape"], g["seed"], "randn").to(DEV)
    out = rs(x)
    check_out(out, g["out"], f"resampler {tag}")
    loss_of(out).backward()
    check_grads(rs.named_parameters(), g["grads"], f"resampler {tag}")


MASK_CASES = ["no_image", "leading_image", "two_images", "more_tokens_than_media", "attend_previous_false", "none"]


@pytest.mark.parametrize("name", MASK_CASES)
def test_masked_cross_attention_golden(name):
    from otter_b200 import functional as F
    from otter_b200.modeling_otter import OtterMaskedCrossAttention
    g = gold(f"xattn_{name}.pt")
    c = g["cfg"]
    att = OtterMaskedCrossAttention(dim=c["D"], dim_visual=c["Dv"])
    load_seeded_(att, g["seed"])
    att.to(DEV)
    x = seeded_tensor("in.xattn.x", (c["B"], c["L"], c["D"]), g["seed"], "randn").to(DEV).requires_grad_(True)
    media = seeded_tensor("in.xattn.media", (c["B"], c["T"], c["n"], c["Dv"]), g["seed"], "randn").to(DEV).requires_grad_(True)
    loc = g["media_locations"].to(DEV) if g["media_locations"] is not None else None
    if loc is not None:   # index tensor: bit-exact against the reference's own text_time
        tt = F.text_time(loc, c["attend_previous"])
        assert torch.equal(tt.cpu().long(), g["text_time"])
    out = att(x, media, media_locations=loc, attend_previous=c["attend_previous"])
    check_out(out, g["out"], f"xattn {name}")
    loss_of(out).backward()
    check_out(x.grad, g["dx"], f"xattn {name} dx", fro=3e-2, mx=6e-2)
    check_out(media.grad, g["dmedia"], f"xattn {name} dmedia", fro=3e-2, mx=6e-2)
    check_grads(att.named_parameters(), g["grads"], f"xattn {name}")


@pytest.mark.parametrize("name", ["two_images", "more_tokens_than_media"])
def test_gated_block_golden(name):
    from otter_b200.modeling_otter import OtterGatedCrossAttentionBlock
    g = gold(f"gated_{name}.pt")
    c = g["cfg"]
    gb = OtterGatedCrossAttentionBlock(dim=c["D"], dim_visual=c["Dv"])
    load_seeded_(gb, g["seed"])
    gb.to(DEV)
    x = seeded_tensor("in.gated.x", (c["B"], c["L"], c["D"]), g["seed"], "randn").to(DEV).requires_grad_(True)
    media = seeded_tensor("in.gated.media", (c["B"], c["T"], c["n"], c["Dv"]), g["seed"], "randn").to(DEV).requires_grad_(True)
    loc = torch.zeros(c["B"], c["L"], dtype=torch.bool)
    for b, ps in enumerate(c["pos"]):
        loc[b, ps] = True
    out = gb(x, media, media_locations=loc.to(DEV), attend_previous=c["attend_previous"])
    check_out(out, g["out"], f"gated {name}")
    loss_of(out).backward()
    check_out(x.grad, g["dx"], f"gated {name} dx", fro=3e-2, mx=6e-2)
    check_out(media.grad, g["dmedia"], f"gated {name} dmedia", fro=3e-2, mx=6e-2)
    check_grads(gb.named_parameters(), g["grads"], f"gated {name}")     # incl. attn_gate / ff_gate


@pytest.mark.parametrize("tag,img", [("small", 56), ("vitl_2layer", 224)])
def test_clip_golden(tag, img):
    from transformers import CLIPVisionConfig
    from otter_b200.modeling_clip import CLIPVisionModel
    g = gold(f"clip_{tag}.pt")
    clip = CLIPVisionModel(CLIPVisionConfig(hidden_act="quick_gelu", **g["cfg"]))
    sd = load_seeded_(clip, g["seed"], kinds={"vision_model.embeddings.class_embedding": "emb",
                                              "vision_model.embeddings.position_embedding.weight": "emb"})
    clip.to(DEV).requires_grad_(False)
    px = seeded_tensor(f"in.clip.{tag}", (2, 3, img, img), g["seed"], "randn")
    out = clip(px.to(DEV))[0]
    check_out(out, g["out"], f"clip {tag}", fro=2e-2, mx=6e-2)
    # and against the CPU oracle with bf16 rounding at the same materialisation points (tighter)
    ref_q = R.clip_vision_last_hidden(px, {k: v.float() for k, v in sd.items()}, heads=g["cfg"]["num_attention_heads"],
                                      q=R.bf16_round)
    check_out(out, ref_q, f"clip {tag} vs bf16-rounded oracle", fro=1e-2, mx=4e-2)


def test_gated_block_vs_oracle_identity_and_errors():
    """Gates at 0 => block is the identity on x (tanh(0)=0, reference :362,371); unconditioned layer errors."""
    from otter_b200.modeling_otter import OtterGatedCrossAttentionBlock, OtterLayer
    gb = OtterGatedCrossAttentionBlock(dim=256, dim_visual=1024).to(DEV)
    x = torch.randn(2, 16, 256, device=DEV).to(torch.bfloat16)
    media = torch.randn(2, 1, 64, 1024, device=DEV).to(torch.bfloat16)
    loc = torch.zeros(2, 16, dtype=torch.bool, device=DEV)
    loc[:, 0] = True
    out = gb(x, media, media_locations=loc)
    assert torch.equal(out, x)
    layer = OtterLayer(gb, torch.nn.Identity())
    with pytest.raises(ValueError, match="vis_x must be conditioned"):
        layer(x)
    layer.condition_vis_x(media)
    with pytest.raises(ValueError, match="media_locations must be conditioned"):
        layer(x)


def test_tiny_full_model_golden(monkeypatch):
    """Config 1: OpenFlamingo-tiny shape (2-layer LM, 1 gated block, 1-layer CLIP-L-width tower, 6-block perceiver)."""
    from transformers import CLIPVisionConfig, LlamaConfig
    from otter_b200 import otter_hf
    monkeypatch.setattr(otter_hf, "AutoTokenizer", FakeTokenizer)
    g = gold("tiny_full_model.pt")
    tc = LlamaConfig(**{k: v for k, v in g["text_config"].items() if k in (
        "vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
        "num_key_value_heads", "max_position_embeddings")})
    td = tc.to_dict()
    td["_name_or_path"] = "llama-tiny"
    td["architectures"] = ["LlamaForCausalLM"]
    cfg = otter_hf.OtterConfig(vision_config={k: v for k, v in g["vision_config"].items() if k in (
        "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "image_size", "patch_size",
        "hidden_act")}, text_config=td, cross_attn_every_n_layers=2)
    cfg.text_config._name_or_path = "llama-tiny"
    cfg.text_config.architectures = ["LlamaForCausalLM"]
    model = otter_hf.OtterForConditionalGeneration(cfg)
    assert model.media_token_id == g["media_token_id"]
    load_seeded_(model, g["seed"], kinds={"perceiver.latents": "randn",
                                          "vision_encoder.vision_model.embeddings.class_embedding": "emb",
                                          "vision_encoder.vision_model.embeddings.position_embedding.weight": "emb",
                                          "lang_encoder.model.embed_tokens.weight": "emb"})
    model.to(DEV)
    vision_x = seeded_tensor("in.full.vision_x", (2, 1, 1, 3, 224, 224), g["seed"], "randn").to(DEV)
    lang_x, labels = g["lang_x"].to(DEV), g["labels"].to(DEV)
    out = model(vision_x=vision_x, lang_x=lang_x, attention_mask=torch.ones_like(lang_x), labels=labels)
    assert abs(out.loss.item() - g["loss"].item()) <= 2e-2 * abs(g["loss"].item()), (out.loss.item(), g["loss"].item())
    check_out(out.logits, g["logits"], "tiny model logits", fro=2e-2, mx=6e-2)
    out.loss.backward()
    check_grads(model.named_parameters(), g["grads"], "tiny model", norm_tol=5e-2, samp_tol=1e-1, gate_tol=0.12)
    assert not model.lang_encoder.is_conditioned()      # cleared after forward (reference :970-971)
    with pytest.raises(AssertionError):
        model(vision_x=torch.zeros(2, 3, 224, 224, device=DEV), lang_x=lang_x)


# =================================================================================================
# fp32-grade forward mode: the north star's own tolerance, on the CUDA path
#   "outputs match the reference PyTorch forward on identical random inputs within 1e-3 rel / 1e-5 abs fp32"
# =================================================================================================
NS_RTOL, NS_ATOL = 1e-3, 1e-5


def ns_close(got, ref, what):
    got, ref = got.detach().float().cpu(), ref.float()
    assert got.shape == ref.shape
    err = (got - ref).abs()
    bad = (err > NS_ATOL + NS_RTOL * ref.abs()).sum().item()
    assert bad == 0, f"{what}: {bad}/{err.numel()} outside 1e-3 rel / 1e-5 abs (max err {err.max().item():.3e})"


@torch.no_grad()
def test_fp32_mode_north_star_tolerance():
    import otter_b200
    from transformers import CLIPVisionConfig
    from otter_b200.modeling_clip import CLIPVisionModel
    from otter_b200.modeling_otter import (OtterGatedCrossAttentionBlock, OtterMaskedCrossAttention,
                                           OtterPerceiverBlock, OtterPerceiverResampler)
    with otter_b200.precision("fp32"):
        g = gold("perceiver_block.pt")
        c = g["cfg"]
        blk = OtterPerceiverBlock(dim=c["dim"])
        load_seeded_(blk, g["seed"])
        blk.to(DEV)
        x = seeded_tensor("in.x", (c["b"], c["T"], c["n1"], c["dim"]), g["seed"], "randn").to(DEV)
        lat = seeded_tensor("in.latents", (c["b"], c["T"], c["n2"], c["dim"]), g["seed"], "randn").to(DEV)
        ns_close(blk(x, lat), g["out"], "fp32 perceiver block")
        for tag in ("small", "image", "video"):
            g = gold(f"resampler_{tag}.pt")
            rs = OtterPerceiverResampler(**g["cfg"])
            load_seeded_(rs, g["seed"], kinds={"latents": "randn", "frame_embs": "randn"})
            rs.to(DEV)
            xin = seeded_tensor(f"in.resampler.{tag}", g["in_shape"], g["seed"], "randn").to(DEV)
            ns_close(rs(xin), g["out"], f"fp32 resampler {tag}")
        for name in MASK_CASES:
            g = gold(f"xattn_{name}.pt")
            c = g["cfg"]
            att = OtterMaskedCrossAttention(dim=c["D"], dim_visual=c["Dv"])
            load_seeded_(att, g["seed"])
            att.to(DEV)
            xin = seeded_tensor("in.xattn.x", (c["B"], c["L"], c["D"]), g["seed"], "randn").to(DEV)
            media = seeded_tensor("in.xattn.media", (c["B"], c["T"], c["n"], c["Dv"]), g["seed"], "randn").to(DEV)
            loc = g["media_locations"].to(DEV) if g["media_locations"] is not None else None
            ns_close(att(xin, media, media_locations=loc, attend_previous=c["attend_previous"]), g["out"], f"fp32 xattn {name}")
        for name in ("two_images", "more_tokens_than_media"):
            g = gold(f"gated_{name}.pt")
            c = g["cfg"]
            gb = OtterGatedCrossAttentionBlock(dim=c["D"], dim_visual=c["Dv"])
            load_seeded_(gb, g["seed"])
            gb.to(DEV)
            xin = seeded_tensor("in.gated.x", (c["B"], c["L"], c["D"]), g["seed"], "randn").to(DEV)
            media = seeded_tensor("in.gated.media", (c["B"], c["T"], c["n"], c["Dv"]), g["seed"], "randn").to(DEV)
            loc = torch.zeros(c["B"], c["L"], dtype=torch.bool)
            for b, ps in enumerate(c["pos"]):
                loc[b, ps] = True
            ns_close(gb(xin, media, media_locations=loc.to(DEV), attend_previous=c["attend_previous"]), g["out"],
                     f"fp32 gated {name}")
        for tag, img in (("small", 56), ("vitl_2layer", 224)):
            g = gold(f"clip_{tag}.pt")
            clip = CLIPVisionModel(CLIPVisionConfig(hidden_act="quick_gelu", **g["cfg"]))
            load_seeded_(clip, g["seed"], kinds={"vision_model.embeddings.class_embedding": "emb",
                                                 "vision_model.embeddings.position_embedding.weight": "emb"})
            clip.to(DEV).requires_grad_(False)
            px = seeded_tensor(f"in.clip.{tag}", (2, 3, img, img), g["seed"], "randn").to(DEV)
            got, ref = clip(px)[0].float().cpu(), g["out"]
            err = (got - ref).abs()
            # residual stream is O(10) after two ViT-L layers: 1e-3 rel with the abs floor scaled to the tensor's size
            assert (err > 1e-5 * ref.abs().max() + NS_RTOL * ref.abs()).sum().item() == 0, (tag, err.max().item())


def test_fp32_mode_clip_is_forward_only():
    """The trainable blocks have an fp32-grade backward (tests/test_fp32_backward_gpu.py); the frozen CLIP tower does not."""
    import otter_b200
    from transformers import CLIPVisionConfig
    from otter_b200.modeling_clip import CLIPVisionModel
    clip = CLIPVisionModel(CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2,
                                            image_size=28, patch_size=14, hidden_act="quick_gelu")).to(DEV)
    with otter_b200.precision("fp32"), pytest.raises(RuntimeError, match="forward-only"):
        clip(torch.randn(1, 3, 28, 28, device=DEV))


@torch.no_grad()
def test_fp32_mode_tiny_full_model_logits(monkeypatch):
    """Config 1 (OpenFlamingo-tiny shape) end to end in the fp32-grade mode: logits/loss vs the reference's fp32
    forward at the north-star tolerance (the frozen LM runs stock torch fp32)."""
    import otter_b200
    from transformers import LlamaConfig
    from otter_b200 import otter_hf
    monkeypatch.setattr(otter_hf, "AutoTokenizer", FakeTokenizer)
    torch.backends.cuda.matmul.allow_tf32 = False
    g = gold("tiny_full_model.pt")
    tc = LlamaConfig(**{k: v for k, v in g["text_config"].items() if k in (
        "vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
        "num_key_value_heads", "max_position_embeddings")})
    td = tc.to_dict()
    td["_name_or_path"] = "llama-tiny"
    td["architectures"] = ["LlamaForCausalLM"]
    cfg = otter_hf.OtterConfig(vision_config={k: v for k, v in g["vision_config"].items() if k in (
        "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "image_size", "patch_size",
        "hidden_act")}, text_config=td, cross_attn_every_n_layers=2)
    cfg.text_config._name_or_path = "llama-tiny"
    cfg.text_config.architectures = ["LlamaForCausalLM"]
    model = otter_hf.OtterForConditionalGeneration(cfg)
    load_seeded_(model, g["seed"], kinds={"perceiver.latents": "randn",
                                          "vision_encoder.vision_model.embeddings.class_embedding": "emb",
                                          "vision_encoder.vision_model.embeddings.position_embedding.weight": "emb",
                                          "lang_encoder.model.embed_tokens.weight": "emb"})
    model.to(DEV)
    vision_x = seeded_tensor("in.full.vision_x", (2, 1, 1, 3, 224, 224), g["seed"], "randn").to(DEV)
    lang_x, labels = g["lang_x"].to(DEV), g["labels"].to(DEV)
    with otter_b200.precision("fp32"):
        out = model(vision_x=vision_x, lang_x=lang_x, attention_mask=torch.ones_like(lang_x), labels=labels)
    ns_close(out.logits, g["logits"], "fp32 tiny model logits")
    assert abs(out.loss.item() - g["loss"].item()) <= 1e-3 * abs(g["loss"].item()) + 1e-5


def _tiny_model(monkeypatch, g):
    from transformers import LlamaConfig
    from otter_b200 import otter_hf
    monkeypatch.setattr(otter_hf, "AutoTokenizer", FakeTokenizer)
    tc = LlamaConfig(**{k: v for k, v in g["text_config"].items() if k in (
        "vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
        "num_key_value_heads", "max_position_embeddings")})
    td = tc.to_dict()
    td["_name_or_path"] = "llama-tiny"
    td["architectures"] = ["LlamaForCausalLM"]
    cfg = otter_hf.OtterConfig(vision_config={k: v for k, v in g["vision_config"].items() if k in (
        "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "image_size", "patch_size",
        "hidden_act")}, text_config=td, cross_attn_every_n_layers=2)
    cfg.text_config._name_or_path = "llama-tiny"
    cfg.text_config.architectures = ["LlamaForCausalLM"]
    model = otter_hf.OtterForConditionalGeneration(cfg)
    load_seeded_(model, g["seed"], kinds={"perceiver.latents": "randn",
                                          "vision_encoder.vision_model.embeddings.class_embedding": "emb",
                                          "vision_encoder.vision_model.embeddings.position_embedding.weight": "emb",
                                          "lang_encoder.model.embed_tokens.weight": "emb"})
    return model.to(DEV).eval()


def test_generate_matches_reference_greedy(monkeypatch):
    """generate() (reference :1000-1042; pipeline/demos/demo_models.py:64 call pattern): greedy decoding with the HF
    KV cache — cached steps see a single token with no <image>, i.e. the zero-attention rows.
    fp32-grade mode must reproduce the reference's token ids exactly; production mode must run the same path."""
    import otter_b200
    torch.backends.cuda.matmul.allow_tf32 = False
    g = gold("tiny_generate.pt")
    model = _tiny_model(monkeypatch, g)
    lang_x = g["lang_x"].to(DEV)
    vision_x = seeded_tensor("in.gen.vision_x", (2, 1, 1, 3, 224, 224), g["seed"], "randn").to(DEV)
    kw = dict(vision_x=vision_x, lang_x=lang_x, attention_mask=torch.ones_like(lang_x), max_new_tokens=5,
              do_sample=False, num_beams=1)
    with otter_b200.precision("fp32"):
        out32 = model.generate(**kw)
    assert torch.equal(out32.cpu(), g["generated"]), (out32.cpu(), g["generated"])
    assert not model.lang_encoder.is_conditioned()
    out = model.generate(**kw)                                   # production (bf16) numerics: same plumbing
    assert out.shape == g["generated"].shape and torch.equal(out[:, :9].cpu(), g["lang_x"])


def test_backward_is_deterministic():
    """No atomics anywhere: two identical forward/backward passes give bit-identical outputs and gradients."""
    from otter_b200.modeling_otter import OtterGatedCrossAttentionBlock, OtterPerceiverResampler
    torch.manual_seed(0)
    rs = OtterPerceiverResampler(dim=256, depth=2, max_num_frames=4).to(DEV)
    gb = OtterGatedCrossAttentionBlock(dim=512, dim_visual=256).to(DEV)
    with torch.no_grad():
        gb.attn_gate.fill_(0.5), gb.ff_gate.fill_(0.5)
    feats = torch.randn(3, 2, 2, 100, 256, device=DEV)
    x = torch.randn(3, 200, 512, device=DEV)
    loc = torch.zeros(3, 200, dtype=torch.bool, device=DEV)
    loc[:, 0] = True
    loc[:, 90] = True
    runs = []
    for _ in range(2):
        for p in list(rs.parameters()) + list(gb.parameters()):
            p.grad = None
        out = gb(x, rs(feats), media_locations=loc)
        out.float().pow(2).mean().backward()
        runs.append([out.detach().clone()] + [p.grad.clone() for p in list(rs.parameters()) + list(gb.parameters())])
    assert all(torch.equal(a, b) for a, b in zip(*runs))


@pytest.mark.parametrize("pos,T,attend_previous", [([[0, 12], [3, 20]], 2, True), ([[5], []], 1, True),
                                                    ([[0, 8, 16], [2, 9]], 2, False)])
def test_masked_cross_attention_attend_all_previous_media(pos, T, attend_previous):
    """only_attend_immediate_media=False (reference :246,317: mask_op = torch.ge; no zeroing of rows, :326 guard):
    fwd + bwd against the oracle in production mode, forward in fp32-grade mode at the north-star tolerance."""
    import otter_b200
    from otter_b200.modeling_otter import OtterMaskedCrossAttention
    torch.manual_seed(7)
    B, L, D, Dv, n = 2, 32, 256, 128, 64
    att = OtterMaskedCrossAttention(dim=D, dim_visual=Dv, only_attend_immediate_media=False).to(DEV)
    x = torch.randn(B, L, D, device=DEV).requires_grad_(True)
    media = torch.randn(B, T, n, Dv, device=DEV).requires_grad_(True)
    loc = torch.zeros(B, L, dtype=torch.bool)
    for b, ps in enumerate(pos):
        loc[b, ps] = True
    p = {k: v.detach().float().cpu().requires_grad_(True) for k, v in att.state_dict().items()}
    xr, mr = x.detach().cpu().requires_grad_(True), media.detach().cpu().requires_grad_(True)
    ref = R.masked_cross_attention(xr, mr, loc, p, attend_previous=attend_previous, only_attend_immediate_media=False)
    wgt = torch.randn(B, L, D)
    (ref * wgt).sum().backward()
    out = att(x, media, media_locations=loc.to(DEV), attend_previous=attend_previous)
    (out.float() * wgt.to(DEV)).sum().backward()
    check_out(out, ref.detach(), "xattn ge")
    check_out(x.grad, xr.grad, "xattn ge dx", fro=3e-2, mx=6e-2)
    check_out(media.grad, mr.grad, "xattn ge dmedia", fro=3e-2, mx=6e-2)
    for k, v in att.named_parameters():
        check_out(v.grad, p[k].grad, f"xattn ge grad {k}", fro=4e-2, mx=1e-1)
    with torch.no_grad(), otter_b200.precision("fp32"):
        out32 = att(x.detach(), media.detach(), media_locations=loc.to(DEV), attend_previous=attend_previous)
    ns_close(out32, ref.detach(), "xattn ge fp32-grade")
