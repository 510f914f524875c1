"""Parameter plumbing: bf16 compute shadows of fp32 master weights, and gradient sinks.

The reference trains fp32 weights under `autocast(bf16)` (pipeline/train/instruction_following.py:311-314):
autocast re-casts each weight to bf16 for the matmul.  Here the cast is an explicit kernel
(otb_cast_f32_bf16) whose result is cached until the parameter changes (`_version` bump by the
optimizer / load_state_dict), so a step pays for it exactly once per weight.

Gradient sinks: by default backward returns fresh fp32 gradients to autograd.  When a parameter
carries `_otb_grad` (an fp32 view into the flat data-parallel gradient buffer, see otter_b200.dp),
the wgrad GEMM epilogue writes/accumulates straight into it and autograd receives None — no gather
copy before the single NCCL all-reduce.
"""
import torch

from . import functional as F

_shadow = {}   # id(param) -> (version, data_ptr, bf16 tensor)
_f32 = {}


def bf16_of(p):
    """bf16 compute copy of a weight (identity for bf16 parameters)."""
    t = p.detach() if isinstance(p, torch.nn.Parameter) or p.requires_grad else p
    if t.dtype == torch.bfloat16:
        return t if t.is_contiguous() else t.contiguous()
    key = id(p)
    ver, ptr = p._version, t.data_ptr()
    hit = _shadow.get(key)
    if hit is not None and hit[0] == ver and hit[1] == ptr:
        return hit[2]
    out = hit[2] if (hit is not None and hit[2].shape == t.shape and hit[2].device == t.device) else None
    out = F.cast_bf16(t.float() if t.dtype != torch.float32 else t, out)
    _shadow[key] = (ver, ptr, out)
    return out


def f32_of(p):
    """fp32 view/copy of a small parameter (LayerNorm affine, bias, gate)."""
    t = p.detach()
    if t.dtype == torch.float32:
        return t if t.is_contiguous() else t.contiguous()
    key = id(p)
    ver, ptr = p._version, t.data_ptr()
    hit = _f32.get(key)
    if hit is not None and hit[0] == ver and hit[1] == ptr:
        return hit[2]
    out = t.float().contiguous()
    _f32[key] = (ver, ptr, out)
    return out


def clear_caches():
    _shadow.clear()
    _f32.clear()


class GradSink:
    """Resolves where a parameter's gradient is written during backward."""

    def __init__(self):
        self.returned = {}

    def target(self, p):
        """-> (fp32 tensor to write, accumulate flag)."""
        buf = getattr(p, "_otb_grad", None)
        if buf is not None:
            acc = bool(getattr(p, "_otb_grad_live", False))
            p._otb_grad_live = True
            return buf, acc
        if id(p) in self.returned:          # second contribution within the same backward
            return self.returned[id(p)], True
        t = torch.empty(p.shape, device=p.device, dtype=torch.float32)
        self.returned[id(p)] = t
        return t, False

    def result(self, p):
        """What backward() returns to autograd for this parameter."""
        t = self.returned.get(id(p))
        if t is None:
            return None
        return t if p.dtype == torch.float32 else t.to(p.dtype)
