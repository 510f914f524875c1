"""OpenFlamingo twin of the Otter modules (reference: src/otter_ai/models/flamingo/modeling_flamingo.py).

Same kernels, `Flamingo*` names.  Differences kept from the reference: no `<answer>`
token (:521,734), `use_media_placement_augmentation` honoured (:534,747), perceiver built without
frame_embs (:540,753), `assert F == 1` in FlamingoModel only (:683 vs :899); the vision tower is frozen
and forward-only in both (the reference wraps it in torch.no_grad(), :686,902).
"""
from .modeling_otter import (OtterGatedCrossAttentionBlock, OtterLayer, OtterLMMixin, OtterMaskedCrossAttention,
                             OtterPerceiverBlock, OtterPerceiverResampler)
from .otter_hf import OtterConfig, _OtterBase


class FlamingoPerceiverBlock(OtterPerceiverBlock):
    pass


class FlamingoPerceiverResampler(OtterPerceiverResampler):
    _block_cls = FlamingoPerceiverBlock


class FlamingoMaskedCrossAttention(OtterMaskedCrossAttention):
    pass


class FlamingoGatedCrossAttentionBlock(OtterGatedCrossAttentionBlock):
    _attn_cls = FlamingoMaskedCrossAttention


class FlamingoLayer(OtterLayer):
    pass


class FlamingoLMMixin(OtterLMMixin):
    """modeling_flamingo.py:399-474 — `init_flamingo` is the reference's name for the same wiring."""
    _layer_cls = FlamingoLayer
    _gated_cls = FlamingoGatedCrossAttentionBlock

    def init_flamingo(self, media_token_id, vis_hidden_size, cross_attn_every_n_layers,
                      use_media_placement_augmentation):
        return self.init_otter(media_token_id=media_token_id, vis_hidden_size=vis_hidden_size,
                               cross_attn_every_n_layers=cross_attn_every_n_layers,
                               use_media_placement_augmentation=use_media_placement_augmentation)


class FlamingoConfig(OtterConfig):
    """configuration_flamingo.py — same fields as OtterConfig under model_type "flamingo"."""
    model_type = "flamingo"

    def __init__(self, vision_config=None, text_config=None, cross_attn_every_n_layers: int = 4,
                 use_media_placement_augmentation: bool = True, **kwargs):
        # explicit: transformers 5.x synthesises an __init__ for config subclasses that do not define one, which would
        # skip OtterConfig's dispatch of text_config / vision_config into config objects
        super().__init__(vision_config=vision_config, text_config=text_config,
                         cross_attn_every_n_layers=cross_attn_every_n_layers,
                         use_media_placement_augmentation=use_media_placement_augmentation, **kwargs)


class _FlamingoBase(_OtterBase):
    config_class = FlamingoConfig
    base_model_prefix = "flamingo"
    _no_split_modules = ["FlamingoPerceiverBlock", "CLIPEncoderLayer", "FlamingoLayer"]
    _special_tokens = ["<|endofchunk|>", "<image>"]
    _honour_media_placement_augmentation = True
    _use_frame_embs = False
    _perceiver_cls = FlamingoPerceiverResampler
    _lm_mixin_cls = FlamingoLMMixin


class FlamingoPreTrainedModel(_FlamingoBase):
    pass


class FlamingoModel(_FlamingoBase):
    _assert_single_frame = True


class FlamingoForConditionalGeneration(_FlamingoBase):
    generate = None  # replaced below


from .otter_hf import OtterForConditionalGeneration as _OFCG  # noqa: E402

FlamingoForConditionalGeneration.generate = _OFCG.generate
