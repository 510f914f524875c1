"""CPU: the frozen-LM oracle (oracle/restatement_lm.py) against the golden fixture generated from the
unmodified reference MPTBlock, and — when /root/reference is present — against the live reference classes."""
import os

import pytest
import torch

from oracle import ref_shims
from oracle import restatement_lm as R

GOLD = os.path.join(os.path.dirname(__file__), "golden", "mpt_block_tiny.pt")


def test_alibi_slopes_power_of_two_and_not():
    s8 = R.alibi_slopes(8)
    assert torch.allclose(s8, torch.tensor([2.0 ** -(i + 1) for i in range(8)]))
    s6 = R.alibi_slopes(6)                      # 8 slopes computed, odd-indexed first, then even, truncated
    full = torch.tensor([2.0 ** -(i + 1) for i in range(8)])
    assert torch.allclose(s6, torch.cat([full[1::2], full[::2]])[:6])
    b = R.alibi_key_bias(4, 5)
    assert b.shape == (4, 5) and torch.all(b[:, -1] == 0) and torch.all(b[:, 0] < 0)


def test_block_forward_and_input_grad_match_reference_golden():
    gold = torch.load(GOLD)
    for name, g in gold.items():
        B, S, D, H = g["shape"]
        x = g["x"].clone().requires_grad_(True)
        y = R.mpt_block(x, g["params"], H)
        assert torch.allclose(y, g["y"], rtol=1e-5, atol=1e-5), name
        (gx,) = torch.autograd.grad(y, x, g["gy"])
        assert torch.allclose(gx, g["gx"], rtol=1e-4, atol=1e-5), name


def test_causality_and_shift_property():
    """Outputs at position i do not depend on later tokens; the key-only ALiBi bias equals the relative-distance form."""
    g = torch.load(GOLD)["a"]
    B, S, D, H = g["shape"]
    y = R.mpt_block(g["x"], g["params"], H)
    x2 = g["x"].clone()
    x2[:, S // 2:] += 1.0
    y2 = R.mpt_block(x2, g["params"], H)
    assert torch.equal(y[:, :S // 2], y2[:, :S // 2])
    assert not torch.allclose(y[:, S // 2:], y2[:, S // 2:])


@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference (build container only)")
def test_against_live_reference_block():
    from oracle.make_golden_lm import reference_block
    blk, run = reference_block(128, 8, no_bias=True, seed=5)
    x = torch.randn(2, 40, 128, generator=torch.Generator().manual_seed(2))
    want = run(x)
    got = R.mpt_block(x, dict(blk.state_dict()), 8)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    blk_b, run_b = reference_block(64, 4, no_bias=False, seed=6)        # with biases (generic MPT configs)
    xb = torch.randn(1, 9, 64, generator=torch.Generator().manual_seed(3))
    assert torch.allclose(R.mpt_block(xb, dict(blk_b.state_dict()), 4), run_b(xb), rtol=1e-5, atol=1e-5)
