"""Pins oracle/resample_oracle.py (numpy restatement of Pillow's ImagingResample, the arithmetic behind the reference's
``transforms.Resize`` on PIL images, src/transformers.py:73-82) to the installed Pillow, bit for bit.  CPU only."""
import numpy as np
import pytest

from oracle import resample_oracle as R

PIL = pytest.importorskip("PIL.Image")

CASES = [
    # in (H, W) -> out (H, W)
    ((64, 96), (16, 24)),        # the reference's / 4 (src/transformers.py:74)
    ((50, 37), (13, 9)),         # ragged, non-integer scale
    ((20, 30), (45, 64)),        # upscale
    ((33, 47), (33, 20)),        # width only
    ((33, 47), (8, 47)),         # height only
    ((9, 9), (9, 9)),            # identity
    ((5, 3), (1, 1)),            # everything into one pixel
]


@pytest.mark.parametrize("filt,pil_filter", [(R.BILINEAR, 2), (R.BICUBIC, 3)])     # Image.BILINEAR = 2, Image.BICUBIC = 3
@pytest.mark.parametrize("src,dst", CASES)
def test_resize_matches_pillow(src, dst, filt, pil_filter):
    rng = np.random.default_rng(src[0] * 131 + dst[1] + filt)
    img = rng.integers(0, 256, size=(src[0], src[1], 3), dtype=np.uint8)
    img[: src[0] // 2, : src[1] // 2] = rng.choice([0, 255], size=(src[0] // 2, src[1] // 2, 3))     # saturating edges
    ref = np.asarray(PIL.fromarray(img, "RGB").resize((dst[1], dst[0]), resample=pil_filter))
    out = R.resize_u8(img, dst[0], dst[1], filt)
    assert out.dtype == np.uint8 and out.shape == ref.shape
    assert np.array_equal(out, ref)


def test_downward_img_quality_matches_torchvision_pipeline():
    """src/transformers.py:73-77 with the noise drawn up front: Resize (PIL, antialiased bilinear) -> ToTensor -> + noise."""
    torch = pytest.importorskip("torch")
    T = pytest.importorskip("torchvision.transforms")
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, size=(64, 128, 3), dtype=np.uint8)
    pil = PIL.fromarray(img, "RGB")
    x = T.ToTensor()(T.Resize((16, 32))(pil))
    noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(3))
    sigma = 0.0173
    ref = x + noise * sigma
    out = R.downward_img_quality(img, 16, 32, noise.numpy(), sigma)
    assert np.array_equal(out, ref.numpy())
    big = T.ToTensor()(T.Resize((40, 72), PIL.BICUBIC)(pil))                      # normalize_img_size, :79-82
    assert np.array_equal(R.to_tensor(R.resize_u8(img, 40, 72, R.BICUBIC)), big.numpy())


def test_coefficient_tables():
    k, b, c = R.precompute_coeffs(64, 16, R.BILINEAR)
    assert k == 9 and b.shape == (16, 2) and c.shape == (16, 9)
    assert (c.sum(axis=1) - (1 << R.PRECISION_BITS)).__abs__().max() <= 8        # rows sum to one in fixed point
    k2, _, _ = R.precompute_coeffs(16, 64, R.BICUBIC)
    assert k2 == 5                                                              # upscaling: the filter's own support
