"""CPU: host logic of the fp32-grade autograd composition (otter_b200/fp32_path.py).  The CUDA kernels are replaced by
plain torch formulas of what each C entry point computes (include/otter_b200.h), so this checks the WIRING — which gradient
goes where, transposes, fan-in, saved tensors, parameter shapes — against the gradients the unmodified reference produced
(tests/golden/*.pt).  The kernels themselves are checked on the GPU (tests/test_fp32_backward_gpu.py)."""
import math
import os

import pytest
import torch

from oracle.seeded import load_seeded_, sample_flat, seeded_tensor

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=False)


class _Split:          # stands for the [rows, 6K] bf16 operand of the split GEMM: carries the fp32 matrix
    def __init__(self, m):
        self.m, self.shape = m, (m.shape[0], 6 * m.shape[1])


def _epilogue(acc, bias=None, act=0, scale_ptr=None, scale_tanh=False, residual=None):
    v = acc if bias is None else acc + bias
    if act == 1:
        v = 0.5 * v * (1 + torch.erf(v / math.sqrt(2)))
    elif act == 2:
        v = v * torch.sigmoid(1.702 * v)
    if scale_ptr is not None:
        v = v * (torch.tanh(scale_ptr) if scale_tanh else scale_ptr)
    return v if residual is None else v + residual


def _attn_ref(spec, q, kv1, kv2):
    P, H, Sq, Sk1, Sk2, inner = spec.P, spec.H, spec.Sq, spec.Sk1, spec.Sk2, spec.H * 64
    qh = q.view(P, Sq, H, 64).permute(0, 2, 1, 3) * spec.scale
    ks, vs = [kv1[:, :inner].view(P, Sk1, H, 64)], [kv1[:, inner:].view(P, Sk1, H, 64)]
    if Sk2:
        ks.append(kv2[:, :inner].view(P, Sk2, H, 64))
        vs.append(kv2[:, inner:].view(P, Sk2, H, 64))
    k, v = torch.cat(ks, 1).permute(0, 2, 1, 3), torch.cat(vs, 1).permute(0, 2, 1, 3)
    s = qh @ k.transpose(-1, -2)                                    # [P, H, Sq, nk]
    nk = Sk1 + Sk2
    if spec.text_time is not None:
        tt = spec.text_time.view(P, 1, Sq, 1).long()
        slot = (torch.arange(nk) // spec.n_per_media + 1).view(1, 1, 1, nk)
        allowed = (slot <= tt) if spec.mask_ge else (slot == tt)
        if spec.mask_ge:
            zero, uni = torch.zeros_like(tt, dtype=torch.bool), tt == 0
        else:
            zero, uni = tt == 0, tt > spec.T_img
        s = s.masked_fill(~allowed, float("-inf"))
        s = torch.where(uni | zero, torch.zeros_like(s), s)         # uniform rows; zero rows are blanked below
        p = s.softmax(-1)
        p = torch.where(zero, torch.zeros_like(p), p)
    else:
        p = s.softmax(-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(P * Sq, inner)


@pytest.fixture
def torch_kernels(monkeypatch):
    from otter_b200 import functional as F

    def linear_f32(x, w6, N, *, bias=None, act=0, scale_ptr=None, scale_tanh=False, residual=None):
        assert isinstance(w6, _Split) and w6.m.shape == (N, x.shape[1]), (w6.m.shape, N, x.shape)
        return _epilogue(x @ w6.m.t(), bias, act, scale_ptr, scale_tanh, residual)

    def layernorm_bwd_f32(dy, x, gamma, eps=1e-5, need_dx=True, need_params=True):
        xr = x.detach().requires_grad_(True)
        g, b = gamma.detach().requires_grad_(True), torch.zeros_like(gamma).requires_grad_(True)
        with torch.enable_grad():
            y = torch.nn.functional.layer_norm(xr, (x.shape[1],), g, b, eps)
            dx, dg, db = torch.autograd.grad(y, (xr, g, b), dy)
        return (dx if need_dx else None), (dg if need_params else None), (db if need_params else None)

    def attn_bwd_f32(spec, out, dout):
        q, kv1 = spec.q.detach().requires_grad_(True), spec.kv1.detach().requires_grad_(True)
        kv2 = spec.kv2.detach().requires_grad_(True) if spec.kv2 is not None else None
        with torch.enable_grad():
            o = _attn_ref(spec, q, kv1, kv2)
            assert torch.allclose(o, out, atol=1e-5)
            gs = torch.autograd.grad(o, [t for t in (q, kv1, kv2) if t is not None], dout, allow_unused=True)
        gs = [g if g is not None else torch.zeros_like(t) for g, t in zip(gs, (q, kv1, kv2))]
        return gs[0], gs[1], (gs[2] if kv2 is not None else None)

    def act_bwd_f32(dy, pre, act):
        pr = pre.detach().requires_grad_(True)
        with torch.enable_grad():
            return torch.autograd.grad(_epilogue(pr, act=act), pr, dy)[0]

    def rowbias_grad_f32(dy, div, mod, out_rows):
        out = torch.zeros(out_rows, dy.shape[1])
        idx = (torch.arange(dy.shape[0]) // div) % mod
        return out.index_add_(0, idx, dy)

    monkeypatch.setattr(F, "_mat", lambda t, name="m", dtype=None: t if t.dim() == 2 else t.reshape(-1, t.shape[-1]))
    monkeypatch.setattr(F, "_req", lambda t, dtype=None, name="t": t)
    monkeypatch.setattr(F, "split3_concat", lambda src, pattern: _Split(src))
    monkeypatch.setattr(F, "linear_f32", linear_f32)
    monkeypatch.setattr(F, "epilogue_f32", lambda acc, **kw: _epilogue(acc, **kw))
    monkeypatch.setattr(F, "layernorm_fwd_f32", lambda x, g, b, eps=1e-5: torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps))
    monkeypatch.setattr(F, "layernorm_bwd_f32", layernorm_bwd_f32)
    monkeypatch.setattr(F, "add_rowbias_f32", lambda x, bias, div, mod: x + bias[(torch.arange(x.shape[0]) // div) % mod])
    monkeypatch.setattr(F, "attn_fwd_f32", lambda spec: _attn_ref(spec, spec.q, spec.kv1, spec.kv2))
    monkeypatch.setattr(F, "attn_bwd_f32", attn_bwd_f32)
    monkeypatch.setattr(F, "act_bwd_f32", act_bwd_f32)
    monkeypatch.setattr(F, "gate_grad_f32", lambda dy, f, gate: ((dy * f).sum() * (1 - torch.tanh(gate) ** 2)).reshape(1))
    monkeypatch.setattr(F, "rowbias_grad_f32", rowbias_grad_f32)
    monkeypatch.setattr(F, "text_time", lambda loc, attend_previous=True: _text_time(loc, attend_previous))
    from otter_b200 import fp32_path
    fp32_path._w6.clear()
    yield
    fp32_path._w6.clear()


def _text_time(loc, attend_previous):
    from oracle.restatement import text_time_np
    return torch.from_numpy(text_time_np(loc.numpy(), attend_previous)).to(torch.int32)


TOL = 2e-4


def close(got, ref, what):
    got, ref = got.detach().float(), ref.float()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs().max().item()
    assert err <= TOL * ref.abs().max().item() + 1e-9, f"{what}: max err {err:.3e} vs max |ref| {ref.abs().max().item():.3e}"


def check_pins(named, pins, what):
    named = dict(named)
    for k, pin in pins.items():
        g = named[k].grad
        assert g is not None, f"{what}: no grad for {k}"
        assert g.shape == named[k].shape
        assert abs(g.float().norm().item() - pin["norm"]) <= TOL * pin["norm"] + 1e-9, (what, k, g.norm().item(), pin["norm"])
        err = (sample_flat(g) - pin["sample"]).abs().max().item()
        assert err <= TOL * pin["sample"].abs().max().item() + 1e-9, (what, k, err)


def loss_of(out):
    return out.float().pow(2).mean()


def test_perceiver_block_wiring(torch_kernels):
    import otter_b200
    from otter_b200.modeling_otter import OtterPerceiverBlock
    g = gold("perceiver_block.pt")
    c = g["cfg"]
    blk = OtterPerceiverBlock(dim=c["dim"])
    load_seeded_(blk, g["seed"])
    x = seeded_tensor("in.x", (c["b"], c["T"], c["n1"], c["dim"]), g["seed"], "randn")
    lat = seeded_tensor("in.latents", (c["b"], c["T"], c["n2"], c["dim"]), g["seed"], "randn").requires_grad_(True)
    with otter_b200.precision("fp32"):
        out = blk(x, lat)
        close(out, g["out"], "out")
        loss_of(out).backward()
    close(lat.grad, g["dlat"], "dlat")
    check_pins(blk.named_parameters(), g["grads"], "perceiver block")


@pytest.mark.parametrize("tag", ["small", "video"])
def test_resampler_wiring(torch_kernels, tag):
    import otter_b200
    from otter_b200.modeling_otter import OtterPerceiverResampler
    g = gold(f"resampler_{tag}.pt")
    rs = OtterPerceiverResampler(**g["cfg"])
    load_seeded_(rs, g["seed"], kinds={"latents": "randn", "frame_embs": "randn"})
    x = seeded_tensor(f"in.resampler.{tag}", g["in_shape"], g["seed"], "randn")
    with otter_b200.precision("fp32"):
        out = rs(x)
        close(out, g["out"], "out")
        loss_of(out).backward()
    check_pins(rs.named_parameters(), g["grads"], f"resampler {tag}")


@pytest.mark.parametrize("name", ["two_images", "more_tokens_than_media"])
def test_gated_block_wiring(torch_kernels, name):
    import otter_b200
    from otter_b200.modeling_otter import OtterGatedCrossAttentionBlock
    g = gold(f"gated_{name}.pt")
    c = g["cfg"]
    gb = OtterGatedCrossAttentionBlock(dim=c["D"], dim_visual=c["Dv"])
    load_seeded_(gb, g["seed"])
    x = seeded_tensor("in.gated.x", (c["B"], c["L"], c["D"]), g["seed"], "randn").requires_grad_(True)
    media = seeded_tensor("in.gated.media", (c["B"], c["T"], c["n"], c["Dv"]), g["seed"], "randn").requires_grad_(True)
    loc = torch.zeros(c["B"], c["L"], dtype=torch.bool)
    for b, ps in enumerate(c["pos"]):
        loc[b, ps] = True
    with otter_b200.precision("fp32"):
        out = gb(x, media, media_locations=loc, attend_previous=c["attend_previous"])
        close(out, g["out"], "out")
        loss_of(out).backward()
    close(x.grad, g["dx"], "dx")
    close(media.grad, g["dmedia"], "dmedia")
    check_pins(gb.named_parameters(), g["grads"], f"gated {name}")
