"""CLIP ViT vision tower on the otter_b200 kernels — drop-in for the `CLIPVisionModel` the reference
instantiates (modeling_otter.py:52-54,768; xformers_model/clip.py:453-504).

Same constructor (`CLIPVisionModel(config)` with an HF `CLIPVisionConfig`), same state-dict keys
(`vision_model.embeddings.{class_embedding,patch_embedding.weight,position_embedding.weight}`,
`vision_model.pre_layrnorm.*` (sic), `vision_model.encoder.layers.{i}.{layer_norm1,layer_norm2,
self_attn.{q,k,v,out}_proj,mlp.{fc1,fc2}}.*`, `vision_model.post_layernorm.*`), same call convention:
`model(pixel_values)[0]` is `last_hidden_state` [N, 1+patches, D].  Forward only (the tower is frozen on
the hot path, modeling_otter.py:851-858); all arithmetic runs in otter_b200 kernels.
"""
import torch
import torch.nn as nn

from . import functional as F
from .blocks import clip_vision_forward
from .params import bf16_of, f32_of


class _Embeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        D = cfg.hidden_size
        self.class_embedding = nn.Parameter(torch.randn(D))
        self.patch_embedding = nn.Conv2d(cfg.num_channels, D, kernel_size=cfg.patch_size, stride=cfg.patch_size,
                                         bias=False)
        self.num_patches = (cfg.image_size // cfg.patch_size) ** 2
        self.position_embedding = nn.Embedding(self.num_patches + 1, D)
        self.register_buffer("position_ids", torch.arange(self.num_patches + 1).expand((1, -1)), persistent=False)


class _Attention(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.k_proj, self.v_proj = nn.Linear(D, D), nn.Linear(D, D)
        self.q_proj, self.out_proj = nn.Linear(D, D), nn.Linear(D, D)


class _MLP(nn.Module):
    def __init__(self, D, I):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(D, I), nn.Linear(I, D)


class _EncoderLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        D = cfg.hidden_size
        self.self_attn = _Attention(D)
        self.layer_norm1 = nn.LayerNorm(D, eps=cfg.layer_norm_eps)
        self.mlp = _MLP(D, cfg.intermediate_size)
        self.layer_norm2 = nn.LayerNorm(D, eps=cfg.layer_norm_eps)


class _Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayer(cfg) for _ in range(cfg.num_hidden_layers)])


class _VisionTransformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        D = cfg.hidden_size
        self.embeddings = _Embeddings(cfg)
        self.pre_layrnorm = nn.LayerNorm(D, eps=cfg.layer_norm_eps)   # (sic) key name is load-bearing
        self.encoder = _Encoder(cfg)
        self.post_layernorm = nn.LayerNorm(D, eps=cfg.layer_norm_eps)


class CLIPVisionOutput(tuple):
    """Indexable like the HF output: [0] last_hidden_state, [1] pooler_output."""

    @property
    def last_hidden_state(self):
        return self[0]

    @property
    def pooler_output(self):
        return self[1]


class CLIPVisionModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        if config.hidden_size % 64 != 0 or config.hidden_size // config.num_attention_heads != 64:
            raise ValueError("otter_b200 CLIP requires head_dim == 64 (CLIP ViT-L/14: 1024 / 16)")
        if getattr(config, "hidden_act", "quick_gelu") != "quick_gelu":
            raise ValueError("otter_b200 CLIP implements hidden_act='quick_gelu' (the CLIP ViT-L/14 setting)")
        self.vision_model = _VisionTransformer(config)
        self._prepared = None

    def get_input_embeddings(self):
        return self.vision_model.embeddings.patch_embedding

    # ---- weight preparation: bf16 shadows, fused q|k|v, padded patch-embedding GEMM weight ----
    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def prepare(self):
        sig = self._signature()
        if self._prepared is not None and self._prepared[0] == sig:
            return self._prepared[1]
        cfg, vm = self.config, self.vision_model
        D, P = cfg.hidden_size, cfg.patch_size
        K = cfg.num_channels * P * P
        Kpad = (K + 63) // 64 * 64
        dev = vm.embeddings.class_embedding.device
        pw = torch.zeros((D, Kpad), device=dev, dtype=torch.float32)
        pw[:, :K] = vm.embeddings.patch_embedding.weight.detach().float().reshape(D, K)
        W = dict(D=D, heads=cfg.num_attention_heads, patch=P, Kpad=Kpad, eps=cfg.layer_norm_eps,
                 patch_w=F.cast_bf16(pw), cls=f32_of(vm.embeddings.class_embedding),
                 pos=f32_of(vm.embeddings.position_embedding.weight),
                 pre_ln_w=f32_of(vm.pre_layrnorm.weight), pre_ln_b=f32_of(vm.pre_layrnorm.bias), layers=[])
        for l in vm.encoder.layers:
            a = l.self_attn
            qkv_w = torch.cat([a.q_proj.weight.detach(), a.k_proj.weight.detach(), a.v_proj.weight.detach()], 0).float()
            qkv_b = torch.cat([a.q_proj.bias.detach(), a.k_proj.bias.detach(), a.v_proj.bias.detach()], 0).float()
            W["layers"].append(dict(
                ln1_w=f32_of(l.layer_norm1.weight), ln1_b=f32_of(l.layer_norm1.bias),
                ln2_w=f32_of(l.layer_norm2.weight), ln2_b=f32_of(l.layer_norm2.bias),
                qkv_w=F.cast_bf16(qkv_w.contiguous()), qkv_b=qkv_b.contiguous(),
                out_w=bf16_of(a.out_proj.weight), out_b=f32_of(a.out_proj.bias),
                fc1_w=bf16_of(l.mlp.fc1.weight), fc1_b=f32_of(l.mlp.fc1.bias),
                fc2_w=bf16_of(l.mlp.fc2.weight), fc2_b=f32_of(l.mlp.fc2.bias)))
        self._prepared = (sig, W)
        return W

    def last_hidden_bf16(self, pixel_values):
        """bf16 [N, 1+patches, D] — what the hot path consumes (no dtype round trip)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("otter_b200 CLIP tower is forward-only (frozen on the hot path, "
                                      "modeling_otter.py:851-858); train_vision_encoder is not supported")
        with torch.no_grad():
            return clip_vision_forward(pixel_values, self.prepare())

    def forward(self, pixel_values=None, output_attentions=None, output_hidden_states=None, return_dict=None, **kw):
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")
        from . import fp32_path
        if fp32_path.is_fp32():                                        # fp32-grade parity mode (forward only)
            fp32_path.require_no_grad(pixel_values, *self.parameters())
            with torch.no_grad():
                h32 = fp32_path.clip_last_hidden(self, pixel_values)
                vm = self.vision_model
                pooled = F.layernorm_fwd_f32(h32[:, 0, :].contiguous(), f32_of(vm.post_layernorm.weight),
                                             f32_of(vm.post_layernorm.bias), self.config.layer_norm_eps)
            return CLIPVisionOutput((h32, pooled))
        h = self.last_hidden_bf16(pixel_values)                        # bf16 [N, S, D]
        vm = self.vision_model
        with torch.no_grad():
            pooled, _, _ = F.layernorm_fwd(h[:, 0, :].contiguous(), f32_of(vm.post_layernorm.weight),
                                           f32_of(vm.post_layernorm.bias), self.config.layer_norm_eps,
                                           want_stats=False)
        if pixel_values.dtype in (torch.float32, torch.float16):
            return CLIPVisionOutput((h.to(pixel_values.dtype), pooled.to(pixel_values.dtype)))
        return CLIPVisionOutput((h, pooled))
