"""TEST INFRASTRUCTURE — deterministic, reference-independent parameter/input generation.

Golden fixtures store only seeds + outputs (weights at vis_dim=1024 would be tens of MB); both
oracle/make_golden.py (running the real reference) and the tests regenerate identical weights from
(name, shape, seed) with the rules below.  CPU torch.Generator streams are stable for a given torch
build, and the GPU box runs the same image.
"""
import zlib

import torch


def _gen(seed, name):
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def seeded_tensor(name, shape, seed, kind=None):
    g = _gen(seed, name)
    shape = tuple(shape)
    leaf = name.split(".")[-1]
    if kind is None:
        if name.endswith("attn_gate") or name.endswith("ff_gate"):
            kind = "gate"
        elif len(shape) == 1 and leaf == "weight":
            kind = "ln_w"
        elif len(shape) == 1:
            kind = "bias"
        elif leaf == "weight" and len(shape) >= 2:
            kind = "linear"
        else:
            kind = "randn"
    if kind == "gate":
        return torch.full(shape, 0.5 if name.endswith("attn_gate") else -0.75)
    if kind == "ln_w":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if kind == "bias":
        return 0.1 * torch.randn(shape, generator=g)
    if kind == "linear":
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) / fan_in ** 0.5
    if kind == "emb":
        return 0.02 * torch.randn(shape, generator=g)
    return torch.randn(shape, generator=g)


def seeded_state(shapes, seed, kinds=None):
    """shapes: {name: shape} -> {name: fp32 tensor}; deterministic in (name, shape, seed)."""
    kinds = kinds or {}
    return {k: seeded_tensor(k, s, seed, kinds.get(k)) for k, s in sorted(shapes.items())}


def load_seeded_(module, seed, kinds=None):
    """Overwrite every parameter/buffer of `module` that appears in its state_dict with seeded values."""
    sd = module.state_dict()
    new = seeded_state({k: v.shape for k, v in sd.items() if v.is_floating_point()}, seed, kinds)
    for k, v in sd.items():
        if not v.is_floating_point():
            new[k] = v
    module.load_state_dict(new, strict=True)
    return new


def sample_flat(t, n=64):
    """Deterministic strided sample used to pin large gradient tensors in small fixtures."""
    f = t.detach().reshape(-1).float()
    if f.numel() <= n:
        return f.clone()
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return f[idx].clone()
