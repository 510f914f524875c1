"""TEST INFRASTRUCTURE — CPU restatement of the image transform the reference's dataset applies before the hot path
(pipeline/mimicit_utils/mimicit_dataset.py:132-143): torchvision `Resize((S, S), BICUBIC)` on a PIL image (= Pillow's
two-pass antialiased resampler with 8-bit intermediates, libImaging/Resample.c) -> `ToTensor()` -> `Normalize(mean, std)`.
Integer / byte work: the contract is bit-exact uint8 after the resize and bit-exact fp32 after the normalisation.
Pinned against Pillow + torchvision themselves in tests/test_data_cpu.py (both are in the image; third-party arithmetic
on the path, versions Pillow 12.2 / torchvision 0.26).
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c
FLAMINGO_MEAN = [0.481, 0.458, 0.408]     # mimicit_dataset.py:28-29
FLAMINGO_STD = [0.269, 0.261, 0.276]


def _bicubic(x):
    a = -0.5                          # Resample.c bicubic_filter
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0,
                    np.where(x < 2.0, (((x - 5.0) * x + 8.0) * x - 4.0) * a, 0.0))


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc -> (bounds int32 [out, 2] = (xmin, xsize), kk int32
    [out, ksize]) in float64 exactly as Pillow computes them."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = _bicubic((np.arange(xmax) + xmin - center + 0.5) * ss)
        ww = w.sum()
        k = w / ww if ww != 0.0 else w
        ki = np.where(k < 0, -0.5 + k * (1 << PRECISION_BITS), 0.5 + k * (1 << PRECISION_BITS)).astype(np.int64)   # C (int) cast truncates
        kk[xx, :xmax] = ki.astype(np.int32)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img, out_h, out_w):
    """img uint8 [H, W, C] -> uint8 [out_h, out_w, C]; horizontal pass first, 8-bit intermediate (Resample.c
    ImagingResampleInner + ImagingResampleHorizontal/Vertical_8bpc)."""
    H, W, C = img.shape
    src = img.astype(np.int64)
    if W != out_w:
        bx, kx = precompute_coeffs(W, out_w)
        tmp = np.empty((H, out_w, C), dtype=np.uint8)
        for xx in range(out_w):
            x0, n = bx[xx]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, x0:x0 + n, :], kx[xx, :n].astype(np.int64), axes=([1], [0]))
            tmp[:, xx, :] = _clip8(acc)
        src = tmp.astype(np.int64)
    if H != out_h:
        by, ky = precompute_coeffs(H, out_h)
        out = np.empty((out_h, src.shape[1], C), dtype=np.uint8)
        for yy in range(out_h):
            y0, n = by[yy]
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(ky[yy, :n].astype(np.int64), src[y0:y0 + n], axes=([0], [0]))
            out[yy] = _clip8(acc)
        return out
    return src.astype(np.uint8)


def to_tensor_normalize(img_u8, mean=FLAMINGO_MEAN, std=FLAMINGO_STD):
    """uint8 [H, W, 3] -> fp32 [3, H, W]: ToTensor (x / 255) then Normalize ((x - mean) / std), fp32 arithmetic in
    torchvision's operation order."""
    x = img_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32)[:, None, None]
    s = np.asarray(std, dtype=np.float32)[:, None, None]
    return (x - m) / s


def patch_resize_transform(img_u8, size=224, mean=FLAMINGO_MEAN, std=FLAMINGO_STD):
    return to_tensor_normalize(resize_bicubic_u8(img_u8, size, size), mean, std)
