"""GPU: the north star's single kernel — masked cross-attention + to_out + tanh gate + residual (otb_xattn_out_fused,
reference modeling_otter.py:290-340,380-389) — against the two-kernel path it replaces (same bf16 rounding points, so
agreement is tight) and, through the gated block, against the CPU oracle (forward and backward)."""
import pytest
import torch

from oracle import restatement as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF16 = torch.bfloat16


def _rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-20)).item()


@pytest.mark.parametrize("B,L,D,n,locs", [
    (8, 256, 4096, 64, None),                       # the c2 shape, <image> at position 0
    (2, 77, 512, 64, [[5], []]),                    # ragged row tile, a sample without <image> (all rows zeroed)
    (3, 130, 1024, 32, [[0, 64], [3], [129]]),      # 32 latents, more <image> tokens than media (uniform rows), tile edge
])
def test_fused_kernel_matches_two_kernel_path(B, L, D, n, locs):
    from otter_b200 import functional as F
    g = torch.Generator().manual_seed(L + D)
    H, inner = 8, 512
    q = torch.randn(B * L, inner, generator=g).to(BF16).to(DEV)
    kv = torch.randn(B * n, 2 * inner, generator=g).to(BF16).to(DEV)
    wo = (torch.randn(D, inner, generator=g) * inner ** -0.5).to(BF16).to(DEV)
    x = torch.randn(B * L, D, generator=g).to(BF16).to(DEV)
    gate = torch.tensor([0.6], device=DEV)
    loc = torch.zeros(B, L, dtype=torch.bool)
    if locs is None:
        loc[:, 0] = True
    else:
        for b, ps in enumerate(locs):
            loc[b, ps] = True
    tt = F.text_time(loc.to(DEV), True)
    spec = F.AttnSpec(q, 0, kv, 0, inner, B, H, L, n, 0.125, text_time=tt, n_per_media=n, T_img=1)
    o_ref, lse_ref = F.attn_fwd(spec)
    a_ref = torch.empty(B * L, D, device=DEV, dtype=BF16)
    y_ref = F.linear_fwd(o_ref, wo, aux_out=a_ref, scale_ptr=gate, scale_tanh=True, residual=x)
    y, a, o, lse = F.xattn_out_fused(spec, wo, gate, x)
    # same arithmetic up to the summation order of the softmax denominator: at most one bf16 ulp apart
    assert (o.float() - o_ref.float()).abs().max().item() <= 2.0 ** -7 * o_ref.float().abs().max().item()
    assert _rel(o, o_ref) <= 2e-3
    assert torch.allclose(lse, lse_ref, rtol=1e-5, atol=1e-5)
    assert (a.float() - a_ref.float()).abs().max().item() <= 2e-2 * a_ref.float().abs().max().item() + 1e-3
    assert _rel(a, a_ref) <= 2e-3 and _rel(y, y_ref) <= 2e-3


def test_gated_block_with_the_fused_kernel_vs_oracle(monkeypatch):
    from otter_b200 import blocks
    from otter_b200.modeling_otter import OtterGatedCrossAttentionBlock
    monkeypatch.setattr(blocks, "XATTN_OUT_FUSED", True)
    B, L, D, Dv, n = 2, 200, 1024, 256, 64
    torch.manual_seed(2)
    gb = OtterGatedCrossAttentionBlock(dim=D, dim_visual=Dv)
    with torch.no_grad():
        gb.attn_gate.fill_(0.5), gb.ff_gate.fill_(-0.4)
    p_ref = {k: v.detach().clone().float().requires_grad_(True) for k, v in gb.state_dict().items()}
    gb.to(DEV)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, L, D, generator=g).to(BF16).float()
    media = torch.randn(B, 1, n, Dv, generator=g).to(BF16).float()
    w = torch.randn(B, L, D, generator=g).to(BF16).float()
    loc = torch.zeros(B, L, dtype=torch.bool)
    loc[0, 0], loc[1, 4] = True, True
    xg, mg = x.to(DEV).requires_grad_(True), media.to(DEV).requires_grad_(True)
    with torch.no_grad():
        gb(xg, mg, media_locations=loc.to(DEV))               # warm-up: builds the bf16 weight copies (cast launches)
    n0 = blocks.F.launch_count()
    y = gb(xg, mg, media_locations=loc.to(DEV))
    launches_fwd = blocks.F.launch_count() - n0
    (y.float() * w.to(DEV)).sum().backward()
    xr, mr = x.clone().requires_grad_(True), media.clone().requires_grad_(True)
    yr = R.gated_cross_attention_block(xr, mr, loc, p_ref)
    (yr * w).sum().backward()
    assert _rel(y, yr) <= 1.5e-2 and _rel(xg.grad, xr.grad) <= 3e-2 and _rel(mg.grad, mr.grad) <= 3e-2
    named = dict(gb.named_parameters())
    for k, pr in p_ref.items():
        if k.endswith("_gate"):
            assert abs(named[k].grad.item() - pr.grad.item()) <= 0.12 * abs(pr.grad.item()) + 1e-6, k
        else:
            assert _rel(named[k].grad, pr.grad) <= 4e-2, (k, _rel(named[k].grad, pr.grad))
    monkeypatch.setattr(blocks, "XATTN_OUT_FUSED", False)
    n0 = blocks.F.launch_count()
    gb(xg.detach(), mg.detach(), media_locations=loc.to(DEV))
    assert launches_fwd == blocks.F.launch_count() - n0 - 1        # one launch fewer than the two-kernel path
