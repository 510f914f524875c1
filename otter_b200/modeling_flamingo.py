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
    pass


class FlamingoMaskedCrossAttention(OtterMaskedCrossAttention):
    pass


class FlamingoGatedCrossAttentionBlock(OtterGatedCrossAttentionBlock):
    pass


class FlamingoLayer(OtterLayer):
    pass


class FlamingoLMMixin(OtterLMMixin):
    pass


class FlamingoConfig(OtterConfig):
    model_type = "flamingo"


class _FlamingoBase(_OtterBase):
    config_class = FlamingoConfig
    base_model_prefix = "flamingo"
    _no_split_modules = ["FlamingoPerceiverBlock", "CLIPEncoderLayer", "FlamingoLayer"]
    _special_tokens = ["<|endofchunk|>", "<image>"]
    _honour_media_placement_augmentation = True
    _use_frame_embs = False
    _perceiver_cls = FlamingoPerceiverResampler


class FlamingoPreTrainedModel(_FlamingoBase):
    pass


class FlamingoModel(_FlamingoBase):
    _assert_single_frame = True


class FlamingoForConditionalGeneration(_FlamingoBase):
    generate = None  # replaced below


from .otter_hf import OtterForConditionalGeneration as _OFCG  # noqa: E402

FlamingoForConditionalGeneration.generate = _OFCG.generate
