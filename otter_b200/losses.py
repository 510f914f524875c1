"""SURVEY.md §8f row 2 — the ops immediately either side of the hot path in the reference's training step, on device.

* `label_mask`  — pipeline/train/instruction_following.py:163-190 (`masking()`): the reference builds the labels with
  a python loop per sequence and `torch.where` (one device sync per row); here one kernel, bit-exact.
* `shifted_cross_entropy` — the LM-head loss of the frozen LM (src/otter_ai/models/mpt/modeling_mpt.py:430-436;
  HF LLaMA shifts the same way): fused log-softmax + NLL + gradient over the [B, L, V] logits (V = 50 k for MPT).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check
from .functional import _on_device, _p, _stream


@_on_device
def label_mask(input_ids, eos_token_id, answer_token_id, endofchunk_token_id, masking_number=-100):
    """int64 [B, L] -> int64 labels [B, L] exactly as the reference's masking()."""
    if not input_ids.is_cuda or input_ids.dtype != torch.int64:
        raise _lib.OtbError("label_mask: input_ids must be a CUDA int64 tensor")
    ids = input_ids.contiguous()
    B, L = ids.shape
    labels = torch.empty_like(ids)
    check(_lib.load().otb_label_mask(_p(ids), B, L, int(eos_token_id), int(answer_token_id), int(endofchunk_token_id),
                                     int(masking_number), _p(labels), _stream()), "otb_label_mask")
    return labels


class _ShiftedCE(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, logits, labels):
        lib = _lib.load()
        B, L, V = logits.shape
        lg = logits if logits.is_contiguous() else logits.contiguous()
        if lg.dtype not in (torch.float32, torch.bfloat16):
            raise _lib.OtbError("shifted_cross_entropy: logits must be fp32 or bf16")
        lab = labels.contiguous()
        loss = torch.empty((), device=lg.device, dtype=torch.float32)
        need_grad = logits.requires_grad
        dlogits = torch.empty_like(lg) if need_grad else None
        ws = torch.empty(1 + B * L, device=lg.device, dtype=torch.float32)
        check(lib.otb_shifted_cross_entropy(_p(lg), int(lg.dtype == torch.float32), V, _p(lab), B, L, V, _p(loss),
                                            _p(dlogits), V, _p(ws), _stream()), "otb_shifted_cross_entropy")
        ctx.dlogits = dlogits
        return loss

    @staticmethod
    @_on_device
    def backward(ctx, grad_out):
        d = ctx.dlogits
        if d is None:
            return None, None
        g = grad_out.to(torch.float32).contiguous()
        check(_lib.load().otb_scale_by_scalar(_p(d), int(d.dtype == torch.float32), d.numel(), _p(g), _stream()),
              "otb_scale_by_scalar")
        return d, None


def shifted_cross_entropy(logits, labels):
    """logits [B, L, V] (fp32 / bf16, CUDA), labels int64 [B, L] with -100 = ignore.  Equals
    F.cross_entropy(logits.view(-1, V), roll(labels, -1) with the last column set to -100)."""
    if not logits.is_cuda:
        raise _lib.OtbError("shifted_cross_entropy: CUDA tensors only (otter_b200 has no CPU path)")
    return _ShiftedCE.apply(logits, labels.to(torch.int64))
