"""Input pipeline -> device: what sits between the decoded sample and `forward_pass`.

Mirrors, with the reference's names and argument meaning:
  * `patch_resize_transform`  pipeline/mimicit_utils/mimicit_dataset.py:132-143 — Resize((S,S), BICUBIC) -> ToTensor ->
    Normalize(FLAMINGO_MEAN, FLAMINGO_STD), here `ImagePreprocessor`: decoded uint8 HWC images are staged once in pinned
    memory, copied to the GPU and resampled + normalised + cast there by two kernels (csrc/otb_data.cu), bit-exact
    w.r.t. Pillow / torchvision;
  * `process_images`          :329-350 — images of one sample -> [n, 3, S, S] (image) / [1, n, 3, S, S] (video);
  * `collate_tokens` / `collate_fn`  :510-620 — right-padded token / mask tensors and the stacked `patch_images`;
  * `resample_frames_fn`      :307-311.
`DeviceBatcher` strings them together for a whole batch: ONE staging buffer, one H2D copy on a copy stream, one
kernel pair, labels built on the device by `losses.label_mask` (instruction_following.py:163-190).

Host side of the kernel: Pillow's coefficient tables (Resample.c precompute_coeffs + normalize_coeffs_8bpc) are
computed here in float64 — the same doubles Pillow computes — and cached per (input size, output size).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check
from .functional import _on_device, _p, _stream

FLAMINGO_MEAN = [0.481, 0.458, 0.408]      # mimicit_dataset.py:28-29
FLAMINGO_STD = [0.269, 0.261, 0.276]
_PREC = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0,
                    np.where(x < 2.0, (((x - 5.0) * x + 8.0) * x - 4.0) * a, 0.0))


def resample_coeffs(in_size, out_size):
    """-> (bounds int32 [out, 2], weights int32 [out, ksize]); vectorised over the output pixels."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)          # C (int) cast of a positive-or-clamped double
    xmin = np.where(center - support + 0.5 < 0, 0, xmin)
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size)
    n = (xmax - xmin).astype(np.int64)
    taps = np.arange(ksize, dtype=np.float64)[None, :]
    w = _bicubic((taps + xmin[:, None] - center[:, None] + 0.5) * (1.0 / filterscale))
    w = np.where(taps < n[:, None], w, 0.0)
    # Pillow sums the taps left to right in double: reproduce the summation order (np.sum pairs terms differently)
    ww = np.zeros(out_size, dtype=np.float64)
    for i in range(ksize):
        ww = ww + w[:, i]
    k = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    ki = np.where(k < 0, -0.5 + k * (1 << _PREC), 0.5 + k * (1 << _PREC)).astype(np.int64).astype(np.int32)
    ki = np.where(taps < n[:, None], ki, 0).astype(np.int32)
    bounds = np.stack([xmin, n], axis=1).astype(np.int32)
    return bounds, ki


class ImagePreprocessor:
    """`patch_resize_transform` for a LIST of decoded images, on the device.

    __call__(images) -> bf16 [N, 3, S, S] (CUDA).  images: uint8 [H, W, 3] numpy arrays / CPU tensors / PIL images."""

    def __init__(self, patch_image_size=224, mean=FLAMINGO_MEAN, std=FLAMINGO_STD, device="cuda", out_dtype=torch.bfloat16):
        self.S, self.device, self.out_dtype = int(patch_image_size), torch.device(device), out_dtype
        self.mean = [float(np.float32(m)) for m in mean]
        self.std = [float(np.float32(s)) for s in std]
        self._coef_host = {}           # (in, out) -> (offset of bounds, offset of weights, ksize) in the pool
        self._pool = np.zeros(0, dtype=np.int32)
        self._pool_dev = None
        self._stage = None             # pinned uint8 staging buffer
        self.copy_stream = None

    def _coef(self, in_size):
        key = (in_size, self.S)
        hit = self._coef_host.get(key)
        if hit is None:
            bounds, ki = resample_coeffs(in_size, self.S)
            off_b = self._pool.size
            self._pool = np.concatenate([self._pool, bounds.reshape(-1), ki.reshape(-1)])
            hit = (off_b, off_b + bounds.size, ki.shape[1])
            self._coef_host[key] = hit
            self._pool_dev = None
        return hit

    @staticmethod
    def _as_u8(img):
        if isinstance(img, torch.Tensor):
            img = img.numpy()
        elif not isinstance(img, np.ndarray):                   # PIL image: .convert("RGB") as mimicit_dataset.py:338
            img = np.asarray(img.convert("RGB"))
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
            raise ValueError(f"expected a uint8 [H, W, 3] image, got {img.dtype} {img.shape}")
        return np.ascontiguousarray(img)

    @_on_device
    def _launch(self, table, pool, N, max_h, tmp, out):
        m, s = self.mean, self.std
        check(_lib.load().otb_preprocess_images(_p(table), _p(pool), N, max_h, self.S, _p(tmp), m[0], m[1], m[2], s[0],
                                                s[1], s[2], _p(out), int(out.dtype == torch.float32), _stream()),
              "otb_preprocess_images")

    def __call__(self, images, out=None):
        imgs = [self._as_u8(i) for i in images]
        N, S = len(imgs), self.S
        if N == 0:
            return torch.empty((0, 3, S, S), device=self.device, dtype=self.out_dtype)
        sizes = [im.shape[:2] for im in imgs]
        src_off, tmp_off, rows = [], [], []
        so = to = 0
        for (H, W), im in zip(sizes, imgs):
            src_off.append(so)
            tmp_off.append(to)
            so += (H * W * 3 + 15) // 16 * 16
            to += (H * S * 3 + 15) // 16 * 16
        if self._stage is None or self._stage.numel() < so:
            self._stage = torch.empty(max(so, 1 << 22), dtype=torch.uint8).pin_memory()
        stage_np = self._stage.numpy()
        for im, o in zip(imgs, src_off):
            stage_np[o:o + im.size] = im.reshape(-1)
        coefs = [(self._coef(W), self._coef(H)) for H, W in sizes]
        with torch.cuda.device(self.device):
            dev_src = self._stage[:so].to(self.device, non_blocking=True)
            if self._pool_dev is None:
                self._pool_dev = torch.from_numpy(self._pool).to(self.device)
            base = dev_src.data_ptr()
            for (H, W), o, t, (ch, cv) in zip(sizes, src_off, tmp_off, coefs):
                rows.append([base + o, H, W, ch[0], ch[1], ch[2], cv[0], cv[1], cv[2], t])
            table = torch.tensor(rows, dtype=torch.int64).to(self.device, non_blocking=True)
            tmp = torch.empty(to, device=self.device, dtype=torch.uint8)
            if out is None:
                out = torch.empty((N, 3, S, S), device=self.device, dtype=self.out_dtype)
            self._launch(table, self._pool_dev, N, max(h for h, _ in sizes), tmp, out)
        return out


def resample_frames_fn(image_ids, resample_frames):
    """mimicit_dataset.py:307-311."""
    indices = np.linspace(0, len(image_ids) - 1, resample_frames, dtype=int)
    image_ids = [image_ids[i] for i in indices]
    assert len(image_ids) == resample_frames
    return image_ids


def collate_tokens(values, pad_idx, eos_idx=None, left_pad=False, move_eos_to_beginning=False, pad_to_length=None,
                   pad_to_multiple=1, pad_to_bsz=None):
    """mimicit_dataset.py:552-592 — a list of 1-D (or 2-D) tensors -> one padded tensor (same argument meaning)."""
    size = max(v.size(0) for v in values)
    size = size if pad_to_length is None else max(size, pad_to_length)
    if pad_to_multiple != 1 and size % pad_to_multiple != 0:
        size = int(((size - 0.1) // pad_to_multiple + 1) * pad_to_multiple)
    if pad_idx is None:
        pad_idx = eos_idx
    if values[0].dim() == 1:
        res = values[0].new_full((len(values), size), pad_idx)
    elif values[0].dim() == 2:
        assert move_eos_to_beginning is False
        res = values[0].new_full((len(values), size, values[0].size(1)), pad_idx)
    else:
        raise NotImplementedError
    for i, v in enumerate(values):
        dst = res[i][size - len(v):] if left_pad else res[i][:len(v)]
        assert dst.numel() == v.numel()
        if move_eos_to_beginning:
            dst[0] = v[-1] if eos_idx is None else eos_idx
            dst[1:] = v[:-1]
        else:
            dst.copy_(v)
    return res


class DeviceBatcher:
    """collate_fn (mimicit_dataset.py:510-549) + the H2D step + label masking, for one training batch.

    samples: dicts with "source" (1-D token ids), "text_mask" (1-D), "images" (list of decoded uint8 HWC images; the
    reference's `patch_images` after the transform) and optionally "is_video".  Returns what `forward_pass` consumes:
    images bf16 [B, T, F, 3, S, S] in the reference's layout ([B, 1, n, ...] for image samples — mimicit_dataset.py:
    380-382 `unsqueeze(0)` — and for videos), input_ids / attention_mask int64 [B, L] and labels, all on the device."""

    def __init__(self, pad_idx, eos_idx, answer_token_id=None, endofchunk_token_id=None, patch_image_size=224,
                 device="cuda"):
        self.pad_idx, self.eos_idx = pad_idx, eos_idx
        self.answer_token_id, self.endofchunk_token_id = answer_token_id, endofchunk_token_id
        self.device = torch.device(device)
        self.pre = ImagePreprocessor(patch_image_size, device=device)

    def __call__(self, samples):
        if len(samples) == 0:
            return {}
        larger_size = max(s["source"].size(0) for s in samples)
        ids = collate_tokens([s["source"] for s in samples], self.pad_idx, eos_idx=self.eos_idx, pad_to_length=larger_size)
        mask = collate_tokens([s["text_mask"] for s in samples], 0, eos_idx=self.eos_idx, pad_to_length=larger_size)
        counts = [len(s["images"]) for s in samples]
        if len(set(counts)) != 1:
            raise RuntimeError("stack expects each tensor to be equal size (samples with different image counts)")
        flat = [im for s in samples for im in s["images"]]
        S = self.pre.S
        px = self.pre(flat).view(len(samples), 1, counts[0], 3, S, S)
        input_ids = ids.pin_memory().to(self.device, non_blocking=True)
        attention_mask = mask.pin_memory().to(self.device, non_blocking=True)
        batch = {"net_input": {"input_ids": input_ids, "attention_masks": attention_mask, "patch_images": px}}
        if self.answer_token_id is not None:
            from .losses import label_mask
            batch["labels"] = label_mask(input_ids, self.eos_idx, self.answer_token_id, self.endofchunk_token_id)
        return batch
