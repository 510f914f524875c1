"""TEST INFRASTRUCTURE — CPU restatement (the oracle) of Otter's vision-fusion hot path.

Plain functional torch fp32 on CPU (integer/mask parts in numpy), written from the reference's
math, NOT imported by anything under otter_b200/.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.

Pinned against the real reference: oracle/make_golden.py runs the UNMODIFIED reference modules
(/root/reference, via oracle/ref_shims.py) on seeded inputs and stores their outputs in
tests/golden/*.pt; tests/test_oracle_cpu.py checks every function below against those fixtures
(and live against the reference when /root/reference exists).

All file:line citations are relative to /root/reference/src/otter_ai/models/otter/modeling_otter.py
unless another file is named.  Parameters are passed as dicts keyed by the reference's state-dict
names, so a reference `module.state_dict()` can be fed in directly.

`q` (optional) is a rounding hook applied wherever the CUDA pipeline materialises a bf16 tensor in
HBM; `q=None` is the exact fp32 reference math.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _id(t):
    return t


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def layer_norm(x, w, b, eps=1e-5):
    # nn.LayerNorm(dim) — :136-137,144,251,364 ; eps default 1e-5
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x):
    # nn.GELU() exact erf form — :146,367
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def quick_gelu(x):
    # CLIP hidden_act "quick_gelu": x * sigmoid(1.702 x) — xformers_model/clip.py:137-149 (ACT2FN)
    return x * torch.sigmoid(1.702 * x)


# --------------------------------------------------------------------------------------------
# a-4  OtterPerceiverBlock.forward  :151-184
# --------------------------------------------------------------------------------------------
def perceiver_block(x, latents, p, prefix="", heads=8, dim_head=64, q=None):
    """x [b,T,n1,D] media features, latents [b,T,n2,D]."""
    q_ = q or _id
    g = lambda k: p[prefix + k]
    xn = q_(layer_norm(x, g("norm_media.weight"), g("norm_media.bias")))            # :159
    ln = q_(layer_norm(latents, g("norm_latents.weight"), g("norm_latents.bias")))  # :161
    qq = q_(ln @ g("to_q.weight").t())                                              # :165
    kv_in = torch.cat((xn, ln), dim=-2)                                             # :166
    kv = q_(kv_in @ g("to_kv.weight").t())                                          # :167
    k, v = kv.chunk(2, dim=-1)
    b, T, n2, _ = qq.shape
    nk = k.shape[2]

    def split(t, n):  # "b t n (h d) -> b h t n d"  :168-170
        return t.reshape(b, T, n, heads, dim_head).permute(0, 3, 1, 2, 4)

    qh, kh, vh = split(qq, n2) * dim_head ** -0.5, split(k, nk), split(v, nk)       # :171
    sim = qh @ kh.transpose(-1, -2)                                                 # :174
    sim = sim - sim.amax(dim=-1, keepdim=True).detach()                             # :175
    attn = sim.softmax(dim=-1)                                                      # :176
    out = attn @ vh                                                                 # :178
    out = q_(out.permute(0, 2, 3, 1, 4).reshape(b, T, n2, heads * dim_head))        # :179
    out = q_(out @ g("to_out.weight").t() + latents)                                # :180
    res = out
    h = q_(layer_norm(out, g("feed_forward.0.weight"), g("feed_forward.0.bias")))   # :182-183
    z = q_(h @ g("feed_forward.1.weight").t())
    h = q_(gelu_erf(z))
    return q_(h @ g("feed_forward.3.weight").t() + res)                             # :184


# --------------------------------------------------------------------------------------------
# a-3  OtterPerceiverResampler.forward  :213-235
# --------------------------------------------------------------------------------------------
def perceiver_resampler(x, p, prefix="", depth=None, heads=8, dim_head=64, q=None):
    """x [b,T,F,v,D] -> [b,T,num_latents,D]."""
    q_ = q or _id
    b, T, Fr, v, D = x.shape
    if prefix + "frame_embs" in p:                                                  # :224-226
        x = x + p[prefix + "frame_embs"][:Fr].reshape(1, 1, Fr, 1, D)
    x = q_(x.reshape(b, T, Fr * v, D))                                              # :227
    if prefix + "media_time_embs" in p:                                             # :228-229
        x = x + p[prefix + "media_time_embs"][:T]
    lat = p[prefix + "latents"]
    lat = q_(lat.reshape(1, 1, *lat.shape).expand(b, T, -1, -1))                    # :232
    if depth is None:
        depth = 1 + max(int(k[len(prefix) + 7:].split(".")[0]) for k in p if k.startswith(prefix + "layers."))
    for i in range(depth):                                                          # :233-234
        lat = perceiver_block(x, lat, p, prefix=f"{prefix}layers.{i}.", heads=heads, dim_head=dim_head, q=q)
    return q_(layer_norm(lat, p[prefix + "norm.weight"], p[prefix + "norm.bias"]))  # :235


# --------------------------------------------------------------------------------------------
# a-5  mask construction (integer, bit-exact)  :296-311
# --------------------------------------------------------------------------------------------
def text_time_np(media_locations, attend_previous=True):
    """media_locations bool [B,L] -> int64 text_time [B,L]  (numpy; bit-exact contract)."""
    loc = np.asarray(media_locations).astype(bool)
    tt = np.cumsum(loc.astype(np.int64), axis=-1)                                   # :298
    if not attend_previous:                                                         # :301-311
        tt = tt.copy()
        tt[~loc] += 1
        cnt = np.count_nonzero(loc, axis=1)[:, None]
        tt[tt > cnt] = 0
    return tt


def keep_mask_np(text_time, T_img, n, only_attend_immediate_media=True):
    """bool [B,L,T_img*n]: True where the key may be attended  :313-320"""
    media_time = np.repeat(np.arange(T_img) + 1, n)[None, None, :]
    tt = np.asarray(text_time)[:, :, None]
    return (tt == media_time) if only_attend_immediate_media else (tt >= media_time)


# --------------------------------------------------------------------------------------------
# a-5  OtterMaskedCrossAttention.forward  :262-340
# --------------------------------------------------------------------------------------------
def masked_cross_attention(x, media, media_locations, p, prefix="", attend_previous=True, heads=8, dim_head=64,
                           only_attend_immediate_media=True, q=None, sim_dtype=torch.float32):
    """x [B,L,D], media [B,T,n,Dv], media_locations bool [B,L] or None -> [B,L,D] (after to_out)."""
    q_ = q or _id
    g = lambda k: p[prefix + k]
    B, T_img, n = media.shape[:3]
    L = x.shape[1]
    xn = q_(layer_norm(x, g("norm.weight"), g("norm.bias")))                        # :283
    qq = q_(xn @ g("to_q.weight").t())                                              # :285
    med = media.reshape(B, T_img * n, -1)                                           # :286
    kv = q_(med @ g("to_kv.weight").t())                                            # :288
    k, v = kv.chunk(2, dim=-1)

    def split(t):  # "b n (h d) -> b h n d"  :290-292
        return t.reshape(B, -1, heads, dim_head).permute(0, 2, 1, 3)

    qh, kh, vh = split(qq) * dim_head ** -0.5, split(k), split(v)                   # :293
    sim = qh @ kh.transpose(-1, -2)                                                 # :295
    tt = None
    if media_locations is not None:
        tt = torch.from_numpy(text_time_np(media_locations.cpu().numpy(), attend_previous))
        keep = torch.from_numpy(keep_mask_np(tt.numpy(), T_img, n, only_attend_immediate_media))
        sim = sim.masked_fill(~keep[:, None], -torch.finfo(sim_dtype).max)          # :321
    sim = sim - sim.amax(dim=-1, keepdim=True).detach()                             # :323
    attn = sim.softmax(dim=-1)                                                      # :324
    if media_locations is not None and only_attend_immediate_media:                 # :326-330
        attn = attn.masked_fill((tt == 0)[:, None, :, None], 0.0)
    out = attn @ vh                                                                 # :332
    out = q_(out.permute(0, 2, 1, 3).reshape(B, L, heads * dim_head))               # :333
    return out @ g("to_out.weight").t()                                             # :340


# --------------------------------------------------------------------------------------------
# a-6  OtterGatedCrossAttentionBlock.forward  :373-395
# --------------------------------------------------------------------------------------------
def gated_cross_attention_block(x, media, media_locations, p, prefix="", attend_previous=True, q=None):
    q_ = q or _id
    g = lambda k: p[prefix + k]
    a = masked_cross_attention(x, media, media_locations, p, prefix=prefix + "attn.",
                               attend_previous=attend_previous, q=q)
    x = q_(a * torch.tanh(g("attn_gate")) + x)                                      # :380-389
    res = x
    h = q_(layer_norm(x, g("feed_forward.0.weight"), g("feed_forward.0.bias")))     # :391-392
    z = q_(h @ g("feed_forward.1.weight").t())
    h = q_(gelu_erf(z))
    f = h @ g("feed_forward.3.weight").t()
    return q_(f * torch.tanh(g("ff_gate")) + res)                                   # :393


# --------------------------------------------------------------------------------------------
# a-2  CLIPVisionModel.forward -> last_hidden_state   xformers_model/clip.py:50-81,84-134,137-199,393-446
# --------------------------------------------------------------------------------------------
def clip_vision_last_hidden(pixel_values, p, prefix="vision_model.", num_layers=None, heads=16, patch=14,
                            eps=1e-5, q=None):
    """pixel_values [N,3,H,W] -> last_hidden_state [N,1+(H/patch)*(W/patch),D] (no post_layernorm)."""
    q_ = q or _id
    g = lambda k: p[prefix + k]
    N = pixel_values.shape[0]
    W = g("embeddings.patch_embedding.weight")                                      # clip.py:61-63 conv, no bias
    D = W.shape[0]
    pe = F.conv2d(pixel_values, W, stride=patch)                                    # clip.py:74
    pe = pe.flatten(2).transpose(1, 2)                                              # clip.py:75
    cls = g("embeddings.class_embedding").reshape(1, 1, D).expand(N, 1, D)          # clip.py:77
    h = torch.cat([cls, pe], dim=1) + g("embeddings.position_embedding.weight")[None]   # clip.py:78-80
    h = q_(layer_norm(h, g("pre_layrnorm.weight"), g("pre_layrnorm.bias"), eps))    # clip.py:425 (sic)
    if num_layers is None:
        num_layers = 1 + max(int(k.split("encoder.layers.")[1].split(".")[0]) for k in p if "encoder.layers." in k)
    S = h.shape[1]
    dh = D // heads
    for i in range(num_layers):                                                     # clip.py:347
        lp = f"{prefix}encoder.layers.{i}."
        r = h
        x = q_(layer_norm(h, p[lp + "layer_norm1.weight"], p[lp + "layer_norm1.bias"], eps))   # clip.py:178
        qh = q_(x @ p[lp + "self_attn.q_proj.weight"].t() + p[lp + "self_attn.q_proj.bias"])   # clip.py:106-110
        kh = q_(x @ p[lp + "self_attn.k_proj.weight"].t() + p[lp + "self_attn.k_proj.bias"])
        vh = q_(x @ p[lp + "self_attn.v_proj.weight"].t() + p[lp + "self_attn.v_proj.bias"])
        sp = lambda t: t.reshape(N, S, heads, dh).permute(0, 2, 1, 3)
        att = (sp(qh) * dh ** -0.5) @ sp(kh).transpose(-1, -2)                      # clip.py:123 (scale, no mask)
        o = att.softmax(-1) @ sp(vh)
        o = q_(o.permute(0, 2, 1, 3).reshape(N, S, D))
        h = q_(o @ p[lp + "self_attn.out_proj.weight"].t() + p[lp + "self_attn.out_proj.bias"] + r)  # clip.py:131,185
        r = h
        x = q_(layer_norm(h, p[lp + "layer_norm2.weight"], p[lp + "layer_norm2.bias"], eps))   # clip.py:188
        x = q_(quick_gelu(x @ p[lp + "mlp.fc1.weight"].t() + p[lp + "mlp.fc1.bias"]))          # clip.py:145-147
        h = q_(x @ p[lp + "mlp.fc2.weight"].t() + p[lp + "mlp.fc2.bias"] + r)                  # clip.py:148,190
    return h                                                                        # clip.py:430-434 [0]


# --------------------------------------------------------------------------------------------
# a-1  _encode_vision_x  :975-997
# --------------------------------------------------------------------------------------------
def encode_vision_x(vision_x, clip_p, perc_p, clip_prefix="vision_model.", perc_prefix="", q=None):
    assert vision_x.ndim == 6, "vision_x should be of shape (b, T_img, F, C, H, W)"   # :987
    b, T, Fr = vision_x.shape[:3]
    feats = clip_vision_last_hidden(vision_x.reshape(b * T * Fr, *vision_x.shape[3:]), clip_p, clip_prefix, q=q)
    feats = feats[:, 1:, :]                                                         # :991 drop CLS
    feats = feats.reshape(b, T, Fr, feats.shape[1], feats.shape[2])                 # :992
    return perceiver_resampler(feats, perc_p, perc_prefix, q=q)                     # :994


# --------------------------------------------------------------------------------------------
# a-9  Fuyu patch-linear + scatter   fuyu/modeling_fuyu.py:44-77,122-131
# --------------------------------------------------------------------------------------------
def fuyu_gather_continuous_embeddings(word_embeddings, continuous_embeddings, image_patch_input_indices):
    """word_embeddings [b,s,D]; continuous_embeddings: list of [n_i,D]; indices int [b,s] (-1 = keep word)."""
    out = word_embeddings.clone()
    for bi in range(word_embeddings.shape[0]):                                      # modeling_fuyu.py:65
        dst = torch.nonzero(image_patch_input_indices[bi] >= 0, as_tuple=True)[0]   # :68
        src = image_patch_input_indices[bi][dst]                                    # :71
        if src.shape[0] > continuous_embeddings[bi].shape[0]:                       # :73-76
            raise ValueError("Number of continuous embeddings does not match number of continuous token ids")
        out[bi, dst] = continuous_embeddings[bi][src]                               # :77
    return out


def fuyu_patch_embed(image_patches, w, bias, word_embeddings, image_patches_indices, q=None):
    """image_patches: list of [n_i, 2700]; Linear(2700->D, bias) then scatter — modeling_fuyu.py:126-131."""
    q_ = q or _id
    emb = [q_(pt @ w.t() + bias) for pt in image_patches]
    return fuyu_gather_continuous_embeddings(word_embeddings, emb, image_patches_indices)


# --------------------------------------------------------------------------------------------
# M1 harness composition: CLIP fwd (no grad) -> perceiver -> n gated blocks chained on a
# synthetic hidden state; loss = out.float().pow(2).mean()
# --------------------------------------------------------------------------------------------
def m1_forward(vision_x, hidden, media_locations, clip_p, perc_p, gated_ps, q=None):
    with torch.no_grad():
        b, T, Fr = vision_x.shape[:3]
        feats = clip_vision_last_hidden(vision_x.reshape(b * T * Fr, *vision_x.shape[3:]), clip_p, q=q)[:, 1:, :]
        feats = feats.reshape(b, T, Fr, feats.shape[1], feats.shape[2])
    media = perceiver_resampler(feats, perc_p, q=q)
    x = hidden
    for gp in gated_ps:
        x = gated_cross_attention_block(x, media, media_locations, gp, q=q)
    return x, media


# --------------------------------------------------------------------------------------------
# label masking (pipeline/train/instruction_following.py:163-190) and the shifted LM loss
# (src/otter_ai/models/mpt/modeling_mpt.py:430-436)
# --------------------------------------------------------------------------------------------
def label_mask_np(input_ids, eos_token_id, answer_token_id, endofchunk_token_id, masking_number=-100):
    """numpy restatement of `masking()` — integer, bit-exact contract."""
    ids = np.asarray(input_ids).astype(np.int64)
    labels = np.where(ids == eos_token_id, eos_token_id, masking_number).astype(np.int64)   # :166
    for i in range(ids.shape[0]):
        A = np.nonzero(ids[i] == answer_token_id)[0]                                         # :167
        E = np.nonzero(ids[i] == endofchunk_token_id)[0]                                     # :168
        j = 0
        for a in A:                                                                          # :171-183
            while j < len(E) and E[j] < a:
                j += 1
            if j < len(E):
                labels[i, a + 1:E[j] + 1] = ids[i, a + 1:E[j] + 1]
                j += 1
        for a, e in zip(A, E):                                                               # :185-186
            labels[i, a + 1:e + 1] = ids[i, a + 1:e + 1]
    labels[:, 0] = masking_number                                                            # :188
    return labels


def shifted_cross_entropy(logits, labels):
    """F.cross_entropy(logits.view(-1, V), roll(labels, -1) with last column -100) — modeling_mpt.py:430-436."""
    _labels = torch.roll(labels, shifts=-1)
    _labels[:, -1] = -100
    return F.cross_entropy(logits.reshape(-1, logits.size(-1)).float(), _labels.reshape(-1))
