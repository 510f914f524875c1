"""GPU: fp32-grade BACKWARD (verdict r1 #10) — the gradients of the trainable blocks on the CUDA path, held to the north
star's tolerance (1e-3 rel, abs floor 1e-5 x max|ref|) against the gradients the UNMODIFIED reference produced in fp32
(tests/golden/*.pt: full dx / dmedia / dlatents tensors, norm + sampled entries of every parameter gradient), plus the
fp32 backward kernels one by one against torch fp32."""
import math

import pytest
import torch

from oracle.seeded import load_seeded_, sample_flat, seeded_tensor
from test_fp32_autograd_cpu import _attn_ref, gold, loss_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL, AFLOOR = 1e-3, 1e-5


def ns_grad_close(got, ref, what):
    got, ref = got.detach().float().cpu(), ref.float()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs()
    tol = RTOL * ref.abs() + AFLOOR * ref.abs().max()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{what}: {bad}/{err.numel()} outside 1e-3 rel (+1e-5 max|ref|); max err {err.max().item():.3e}, max|ref| {ref.abs().max().item():.3e}"


def check_pins(named, pins, what):
    named = dict(named)
    assert set(pins) <= set(named)
    for k, pin in pins.items():
        g = named[k].grad
        assert g is not None, f"{what}: no grad for {k}"
        n = g.float().norm().item()
        assert abs(n - pin["norm"]) <= RTOL * pin["norm"] + 1e-12, f"{what}: |grad {k}| {n:.6e} vs {pin['norm']:.6e}"
        ns_grad_close(sample_flat(g), pin["sample"], f"{what}: grad sample {k}")


def test_fp32_perceiver_block_backward():
    import otter_b200
    from otter_b200.modeling_otter import OtterPerceiverBlock
    g = gold("perceiver_block.pt")
    c = g["cfg"]
    blk = OtterPerceiverBlock(dim=c["dim"])
    load_seeded_(blk, g["seed"])
    blk.to(DEV)
    x = seeded_tensor("in.x", (c["b"], c["T"], c["n1"], c["dim"]), g["seed"], "randn").to(DEV)
    lat = seeded_tensor("in.latents", (c["b"], c["T"], c["n2"], c["dim"]), g["seed"], "randn").to(DEV).requires_grad_(True)
    with otter_b200.precision("fp32"):
        loss_of(blk(x, lat)).backward()
    ns_grad_close(lat.grad, g["dlat"], "fp32 perceiver block dlatents")
    check_pins(blk.named_parameters(), g["grads"], "fp32 perceiver block")


@pytest.mark.parametrize("tag", ["small", "image", "video"])
def test_fp32_resampler_backward(tag):
    import otter_b200
    from otter_b200.modeling_otter import OtterPerceiverResampler
    g = gold(f"resampler_{tag}.pt")
    rs = OtterPerceiverResampler(**g["cfg"])
    load_seeded_(rs, g["seed"], kinds={"latents": "randn", "frame_embs": "randn"})
    rs.to(DEV)
    x = seeded_tensor(f"in.resampler.{tag}", g["in_shape"], g["seed"], "randn").to(DEV)
    with otter_b200.precision("fp32"):
        loss_of(rs(x)).backward()
    check_pins(rs.named_parameters(), g["grads"], f"fp32 resampler {tag}")


@pytest.mark.parametrize("name", ["no_image", "leading_image", "two_images", "more_tokens_than_media", "attend_previous_false",
                                  "none"])
def test_fp32_masked_cross_attention_backward(name):
    import otter_b200
    from otter_b200.modeling_otter import OtterMaskedCrossAttention
    g = gold(f"xattn_{name}.pt")
    c = g["cfg"]
    att = OtterMaskedCrossAttention(dim=c["D"], dim_visual=c["Dv"])
    load_seeded_(att, g["seed"])
    att.to(DEV)
    x = seeded_tensor("in.xattn.x", (c["B"], c["L"], c["D"]), g["seed"], "randn").to(DEV).requires_grad_(True)
    media = seeded_tensor("in.xattn.media", (c["B"], c["T"], c["n"], c["Dv"]), g["seed"], "randn").to(DEV).requires_grad_(True)
    loc = g["media_locations"].to(DEV) if g["media_locations"] is not None else None
    with otter_b200.precision("fp32"):
        loss_of(att(x, media, media_locations=loc, attend_previous=c["attend_previous"])).backward()
    ns_grad_close(x.grad, g["dx"], f"fp32 xattn {name} dx")
    ns_grad_close(media.grad, g["dmedia"], f"fp32 xattn {name} dmedia")
    check_pins(att.named_parameters(), g["grads"], f"fp32 xattn {name}")


@pytest.mark.parametrize("name", ["two_images", "more_tokens_than_media"])
def test_fp32_gated_block_backward(name):
    import otter_b200
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
    with otter_b200.precision("fp32"):
        loss_of(gb(x, media, media_locations=loc.to(DEV), attend_previous=c["attend_previous"])).backward()
    ns_grad_close(x.grad, g["dx"], f"fp32 gated {name} dx")
    ns_grad_close(media.grad, g["dmedia"], f"fp32 gated {name} dmedia")
    check_pins(gb.named_parameters(), g["grads"], f"fp32 gated {name}")      # incl. attn_gate / ff_gate at 1e-3


# ------------------------------------------------------------------------------------------------
# the kernels one by one
# ------------------------------------------------------------------------------------------------
def _close(got, ref, what, rtol=2e-5):
    got, ref = got.detach().float().cpu(), ref.float()
    err = (got - ref).abs().max().item()
    assert err <= rtol * ref.abs().max().item() + 1e-12, f"{what}: max err {err:.3e} vs max|ref| {ref.abs().max().item():.3e}"


def test_fp32_layernorm_act_gate_rowbias_kernels():
    from otter_b200 import functional as F
    g = torch.Generator().manual_seed(0)
    rows, D = 77, 200
    x, dy = torch.randn(rows, D, generator=g) * 2 + 0.5, torch.randn(rows, D, generator=g)
    gamma = 1 + 0.3 * torch.randn(D, generator=g)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), torch.zeros(D, requires_grad=True)
    torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5).backward(dy)
    dx, dg, db = F.layernorm_bwd_f32(dy.to(DEV), x.to(DEV), gamma.to(DEV), 1e-5)
    _close(dx, xr.grad, "ln dx")
    _close(dg, gr.grad, "ln dgamma")
    _close(db, br.grad, "ln dbeta")
    dx2, dg2, _ = F.layernorm_bwd_f32(dy.to(DEV), x.to(DEV), gamma.to(DEV), 1e-5, need_dx=False)
    assert dx2 is None
    _close(dg2, gr.grad, "ln dgamma (no dx)")
    for act, fn in ((1, lambda v: torch.nn.functional.gelu(v)), (2, lambda v: v * torch.sigmoid(1.702 * v))):
        pr = (x * 1.5).clone().requires_grad_(True)
        fn(pr).backward(dy)
        _close(F.act_bwd_f32(dy.to(DEV), (x * 1.5).to(DEV), act), pr.grad, f"act {act} derivative")
    gate = torch.tensor([0.37])
    ref = (dy.double() * x.double()).sum() * (1 - math.tanh(0.37) ** 2)
    _close(F.gate_grad_f32(dy.to(DEV), x.to(DEV), gate.to(DEV)), ref.float().reshape(1), "gate grad", rtol=1e-5)
    # rows = 3 sequences x 5 frames x 4 tokens; frame table of 8 rows, 5 used
    dyb = torch.randn(60, D, generator=g)
    ref = torch.zeros(8, D).index_add_(0, (torch.arange(60) // 4) % 5, dyb)
    _close(F.rowbias_grad_f32(dyb.to(DEV), 4, 5, 8), ref, "rowbias grad")


@pytest.mark.parametrize("case", ["two_sources", "eq_mask", "ge_mask"])
def test_fp32_attention_backward_kernel(case):
    from otter_b200 import functional as F
    g = torch.Generator().manual_seed(3)
    P, H, inner = 2, 3, 192
    if case == "two_sources":
        Sq, Sk1, Sk2, n, T, tt, ge = 20, 70, 20, 0, 0, None, False
    else:
        Sq, Sk1, Sk2, n, T, ge = 37, 24, 0, 8, 3, case == "ge_mask"
        tt = torch.randint(0, 5, (P, Sq), generator=g).to(torch.int32)       # 0: no media yet, 4 > T_img: the uniform class
    q = torch.randn(P * Sq, inner, generator=g)
    kv1 = torch.randn(P * Sk1, 2 * inner, generator=g)
    kv2 = torch.randn(P * Sk2, 2 * inner, generator=g) if Sk2 else None
    dout = torch.randn(P * Sq, inner, generator=g)

    def spec_of(dev, *ts):
        q_, kv1_, kv2_ = (t.to(dev) if t is not None else None for t in ts)
        s = F.AttnSpec.__new__(F.AttnSpec)
        s.q, s.q_col0, s.kv1, s.k1_col0, s.v1_col0, s.kv2, s.k2_col0, s.v2_col0 = q_, 0, kv1_, 0, inner, kv2_, 0, inner
        s.P, s.H, s.Sq, s.Sk1, s.Sk2, s.scale = P, H, Sq, Sk1, Sk2, 0.125
        s.text_time, s.n_per_media, s.T_img = (tt.to(dev) if tt is not None else None), n, T
        s.mask_ge, s.causal = ge, False
        return s

    qr, k1r = q.clone().requires_grad_(True), kv1.clone().requires_grad_(True)
    k2r = kv2.clone().requires_grad_(True) if kv2 is not None else None
    ref = _attn_ref(spec_of("cpu", qr, k1r, k2r), qr, k1r, k2r)
    ref.backward(dout)
    sp = spec_of(DEV, q, kv1, kv2)
    out = F.attn_fwd_f32(sp)
    _close(out, ref, f"{case} forward")
    dq, dkv1, dkv2 = F.attn_bwd_f32(sp, out, dout.to(DEV))
    _close(dq, qr.grad, f"{case} dq")
    _close(dkv1, k1r.grad, f"{case} dkv1")
    if kv2 is not None:
        _close(dkv2, k2r.grad, f"{case} dkv2")
