"""fp32-grade FORWARD path ("parity mode"): reproduces the reference's fp32 forward within the north star's
1e-3 rel / 1e-5 abs on the CUDA path itself.  Selected with `with otter_b200.precision("fp32"):` (no-grad only).

Dense contractions still run on the tcgen05 GEMM (three-term bf16 split, six cross products over a 6x longer
reduction — functional.linear_f32); LayerNorm, softmax/attention and the few element-wise adds are fp32 CUDA-core
kernels (csrc/otb_fp32.cu).  Pure data movement (drop CLS, concat CLS token, im2col, broadcast of the latents) is
torch indexing — no arithmetic.  Production numerics (bf16 operands, otter_b200.blocks) are untouched.
"""
import contextlib
import weakref

import torch

from . import functional as F
from .params import f32_of

_state = {"fp32": False}


@contextlib.contextmanager
def precision(mode):
    assert mode in ("bf16", "fp32")
    old = _state["fp32"]
    _state["fp32"] = (mode == "fp32")
    try:
        yield
    finally:
        _state["fp32"] = old


def is_fp32():
    return _state["fp32"]


def require_no_grad(*tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise RuntimeError("otter_b200 precision('fp32') is a forward-only parity mode: call it under torch.no_grad()")


_w6 = {}


def w6_of(p, pad_to=None):
    """Cached split3_concat(W, pattern B) of a weight [N, K] (optionally zero-padded along K)."""
    key = (id(p), pad_to)
    hit = _w6.get(key)
    if hit is not None and hit[0]() is p and hit[1] == p._version and hit[2] == p.data_ptr():
        return hit[3]
    w = f32_of(p).reshape(p.shape[0], -1)
    if pad_to is not None and w.shape[1] < pad_to:
        w = torch.nn.functional.pad(w, (0, pad_to - w.shape[1]))
    out = F.split3_concat(w.contiguous(), 1)

    def _drop(r, key=key):
        hit = _w6.get(key)
        if hit is not None and hit[0] is r:
            del _w6[key]

    _w6[key] = (weakref.ref(p, _drop), p._version, p.data_ptr(), out)
    return out


def lin(x, weight, **kw):
    return F.linear_f32(x, w6_of(weight), weight.shape[0], **kw)


def _f32_2d(t, cols):
    t = t.reshape(-1, cols)
    return t.float().contiguous() if (t.dtype != torch.float32 or not t.is_contiguous()) else t


def perceiver_block(blk, x2d, lat2d, BT):
    n1, n2, inner = x2d.shape[0] // BT, lat2d.shape[0] // BT, blk.heads * 64
    ff = blk.feed_forward
    xn = F.layernorm_fwd_f32(x2d, f32_of(blk.norm_media.weight), f32_of(blk.norm_media.bias))
    ln = F.layernorm_fwd_f32(lat2d, f32_of(blk.norm_latents.weight), f32_of(blk.norm_latents.bias))
    q = lin(ln, blk.to_q.weight)
    kv_x, kv_l = lin(xn, blk.to_kv.weight), lin(ln, blk.to_kv.weight)
    spec = F.AttnSpec(q, 0, kv_x, 0, inner, BT, blk.heads, n2, n1, 0.125, kv2=kv_l, k2_col0=0, v2_col0=inner, Sk2=n2,
                      dtype=torch.float32)
    o = F.attn_fwd_f32(spec)
    lat1 = lin(o, blk.to_out.weight, residual=lat2d)
    h = F.layernorm_fwd_f32(lat1, f32_of(ff[0].weight), f32_of(ff[0].bias))
    h = lin(h, ff[1].weight, act=1)
    return lin(h, ff[3].weight, residual=lat1)


def resample_media(rs, media2d, BT):
    n, D = rs.latents.shape
    lat = f32_of(rs.latents).unsqueeze(0).expand(BT, n, D).reshape(BT * n, D).contiguous()
    for blk in rs.layers:
        lat = perceiver_block(blk, media2d, lat, BT)
    return F.layernorm_fwd_f32(lat, f32_of(rs.norm.weight), f32_of(rs.norm.bias), rs.norm.eps)


def resampler_forward(rs, x):
    b, T, Fr, v, D = x.shape
    media = _f32_2d(x, D)
    if rs.frame_embs is not None:
        media = F.add_rowbias_f32(media, f32_of(rs.frame_embs)[:Fr].contiguous(), v, Fr)
    if rs.media_time_embs is not None:
        media = F.add_rowbias_f32(media, f32_of(rs.media_time_embs).reshape(-1, D)[:T].contiguous(), Fr * v, T)
    return resample_media(rs, media, b * T).view(b, T, rs.latents.shape[0], D)


def masked_cross_attention(att, x2d, media2d, tt, B, L, T_img, n, gate=None, residual=None):
    inner = att.heads * 64
    xn = F.layernorm_fwd_f32(x2d, f32_of(att.norm.weight), f32_of(att.norm.bias))
    q = lin(xn, att.to_q.weight)
    kv = lin(media2d, att.to_kv.weight)
    spec = F.AttnSpec(q, 0, kv, 0, inner, B, att.heads, L, T_img * n, 0.125, text_time=tt, n_per_media=n, T_img=T_img,
                      dtype=torch.float32, mask_ge=not att.only_attend_immediate_media)
    o = F.attn_fwd_f32(spec)
    if gate is None:
        return lin(o, att.to_out.weight)
    return lin(o, att.to_out.weight, scale_ptr=f32_of(gate), scale_tanh=True, residual=residual)


def gated_block(gb, x2d, media2d, tt, B, L, T_img, n):
    ff = gb.feed_forward
    x1 = masked_cross_attention(gb.attn, x2d, media2d, tt, B, L, T_img, n, gate=gb.attn_gate, residual=x2d)
    h = F.layernorm_fwd_f32(x1, f32_of(ff[0].weight), f32_of(ff[0].bias))
    h = lin(h, ff[1].weight, act=1)
    return lin(h, ff[3].weight, scale_ptr=f32_of(gb.ff_gate), scale_tanh=True, residual=x1)


def clip_last_hidden(clip, pixel_values):
    cfg, vm = clip.config, clip.vision_model
    D, P, heads, eps = cfg.hidden_size, cfg.patch_size, cfg.num_attention_heads, cfg.layer_norm_eps
    N = pixel_values.shape[0]
    cols = torch.nn.functional.unfold(pixel_values.float(), P, stride=P).transpose(1, 2)      # data movement only
    np_ = cols.shape[1]
    K = cols.shape[2]
    Kp = (K + 7) // 8 * 8
    cols = torch.nn.functional.pad(cols.reshape(N * np_, K), (0, Kp - K)).contiguous()
    pe = F.linear_f32(cols, w6_of(vm.embeddings.patch_embedding.weight, pad_to=Kp), D)
    cls = f32_of(vm.embeddings.class_embedding).reshape(1, 1, D).expand(N, 1, D)
    h = torch.cat([cls, pe.view(N, np_, D)], dim=1).reshape(N * (np_ + 1), D).contiguous()   # data movement only
    S = np_ + 1
    h = F.add_rowbias_f32(h, f32_of(vm.embeddings.position_embedding.weight), 1, S)
    h = F.layernorm_fwd_f32(h, f32_of(vm.pre_layrnorm.weight), f32_of(vm.pre_layrnorm.bias), eps)
    for l in vm.encoder.layers:
        a = l.self_attn
        x = F.layernorm_fwd_f32(h, f32_of(l.layer_norm1.weight), f32_of(l.layer_norm1.bias), eps)
        q = lin(x, a.q_proj.weight, bias=f32_of(a.q_proj.bias))
        k = lin(x, a.k_proj.weight, bias=f32_of(a.k_proj.bias))
        v = lin(x, a.v_proj.weight, bias=f32_of(a.v_proj.bias))
        kv = torch.cat([k, v], dim=1)                                                         # data movement only
        o = F.attn_fwd_f32(F.AttnSpec(q, 0, kv, 0, D, N, heads, S, S, 0.125, dtype=torch.float32))
        h = lin(o, a.out_proj.weight, bias=f32_of(a.out_proj.bias), residual=h)
        x = F.layernorm_fwd_f32(h, f32_of(l.layer_norm2.weight), f32_of(l.layer_norm2.bias), eps)
        x = lin(x, l.mlp.fc1.weight, bias=f32_of(l.mlp.fc1.bias), act=2)
        h = lin(x, l.mlp.fc2.weight, bias=f32_of(l.mlp.fc2.bias), residual=h)
    return h.view(N, S, D)
