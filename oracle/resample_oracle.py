"""CPU oracle of the image-size transforms either side of the hot path.  TEST INFRASTRUCTURE ONLY: imported by
tests/ (and nothing in the product path).

Reference call sites: src/transformers.py:73-77 ``downward_img_quality`` = ``transforms.Resize((clip_height // 4,
clip_width // 4))`` (torchvision's default BILINEAR, antialiased, on a PIL image) -> ``ToTensor`` -> ``x + randn_like(x) *
uniform(0, 0.03)``; src/transformers.py:79-82 ``normalize_img_size`` = ``transforms.Resize((clip_height, clip_width),
Image.BICUBIC)`` -> ``ToTensor``.  On PIL images torchvision delegates to ``PIL.Image.resize`` -- the arithmetic lives in
Pillow (requirements.txt:3 pins Pillow==11.1.0; installed here: 12.2.0, same algorithm), ``src/libImaging/Resample.c``:
``ImagingResample`` for 8-bit channels = separable two-pass convolution (horizontal, then vertical), per-output-pixel
windows ``[center - support, center + support]`` with ``support = filter_support * max(scale, 1)`` (antialiasing),
coefficients normalised in double precision, converted to fixed point with 22 fractional bits, accumulated in int32 from
``1 << 21`` and clipped to uint8 after each pass.  This file restates that algorithm in numpy; **pin:**
tests/test_resample_cpu.py checks it bit for bit against the installed Pillow on random images.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
BILINEAR, BICUBIC = 0, 1


def _bilinear(x: float) -> float:
    if x < 0.0:
        x = -x
    return 1.0 - x if x < 1.0 else 0.0


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


_FILTERS = {BILINEAR: (_bilinear, 1.0), BICUBIC: (_bicubic, 2.0)}


def precompute_coeffs(in_size: int, out_size: int, filt: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box (in0 = 0, in1 = in_size).
    Returns (ksize, bounds int32 [out_size][2] = (xmin, count), coeffs int32 [out_size][ksize])."""
    fn, fsupport = _FILTERS[filt]
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeffs = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ww = 0.0
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        for x in range(xmax):
            w = fn((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x]
            coeffs[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, coeffs


def _pass(img: np.ndarray, axis: int, out_size: int, filt: int) -> np.ndarray:
    """One separable pass over ``axis`` of a uint8 [H][W][C] image."""
    in_size = img.shape[axis]
    _, bounds, coeffs = precompute_coeffs(in_size, out_size, filt)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(n):
            acc += src[xmin + x] * int(coeffs[xx, x])
        # int32 arithmetic in C; the sums stay far below 2^31, so int64 here is the same value
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_u8(img: np.ndarray, out_h: int, out_w: int, filt: int) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), resample) for an 8-bit [H][W][C] image: horizontal pass (if the width changes),
    then vertical pass (if the height changes), each clipped to uint8 -- ImagingResample, Resample.c."""
    H, W = img.shape[:2]
    out = img
    if out_w != W:
        out = _pass(out, 1, out_w, filt)
    if out_h != H:
        out = _pass(out, 0, out_h, filt)
    return out.copy() if out is img else out


def to_tensor(img_u8: np.ndarray) -> np.ndarray:
    """torchvision ToTensor on an 8-bit image: [H][W][C] uint8 -> [C][H][W] float32 / 255."""
    return (np.moveaxis(img_u8, 2, 0).astype(np.float32) / np.float32(255)).astype(np.float32)


def downward_img_quality(img_u8: np.ndarray, out_h: int, out_w: int, noise: np.ndarray, sigma: float) -> np.ndarray:
    """src/transformers.py:73-77 with the random draws supplied: x + noise * sigma (fp32, two roundings like torch)."""
    x = to_tensor(resize_u8(img_u8, out_h, out_w, BILINEAR))
    return (x + (noise.astype(np.float32) * np.float32(sigma)).astype(np.float32)).astype(np.float32)
