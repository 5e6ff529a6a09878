"""GPU parity tests of the on-device image-size transforms (src/transformers.py:73-82): bit-exact against
the CPU oracle (oracle/resample_oracle.py, itself pinned to Pillow in tests/test_resample_cpu.py) and against Pillow /
torchvision directly where they are installed."""
import numpy as np
import pytest
import torch

from oracle import resample_oracle as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import srgan_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return srgan_b200.transformers


CASES = [((64, 96), (16, 24)), ((50, 37), (13, 9)), ((20, 30), (45, 64)), ((33, 47), (33, 20)), ((33, 47), (8, 47)),
         ((9, 9), (9, 9)), ((5, 3), (1, 1)), ((512, 1024), (128, 256))]


@pytest.mark.parametrize("filt", [R.BILINEAR, R.BICUBIC])
@pytest.mark.parametrize("src,dst", CASES)
def test_resize_u8_bit_exact(T, src, dst, filt):
    rng = np.random.default_rng(src[0] * 17 + dst[0] + filt)
    N = 1 if src[0] > 100 else 3
    img = rng.integers(0, 256, size=(N, src[0], src[1], 3), dtype=np.uint8)
    img[:, : src[0] // 2, : src[1] // 2] = rng.choice([0, 255], size=(N, src[0] // 2, src[1] // 2, 3))
    out = T.resize_u8(torch.from_numpy(img).cuda(), dst, filt).cpu().numpy()
    for n in range(N):
        assert np.array_equal(out[n], R.resize_u8(img[n], dst[0], dst[1], filt)), n
    PIL = pytest.importorskip("PIL.Image")
    ref = np.asarray(PIL.fromarray(img[0], "RGB").resize((dst[1], dst[0]), resample=2 if filt == R.BILINEAR else 3))
    assert np.array_equal(out[0], ref)


def test_transform_pipelines_bit_exact(T):
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, size=(2, 64, 128, 3), dtype=np.uint8)
    d = torch.from_numpy(img).cuda()
    # ToTensor
    assert np.array_equal(T.to_tensor(d).cpu().numpy(), np.stack([R.to_tensor(i) for i in img]))
    # normalize_img_size: bicubic to a given clip size, then ToTensor
    hr = T.normalize_img_size(d, 40, 72).cpu().numpy()
    assert np.array_equal(hr, np.stack([R.to_tensor(R.resize_u8(i, 40, 72, R.BICUBIC)) for i in img]))
    # downward_img_quality with the random draws supplied
    g = torch.Generator(device="cuda").manual_seed(5)
    noise = torch.randn(2, 3, 16, 32, device="cuda", generator=g)
    sigma = torch.tensor([0.0173, 0.0291], device="cuda")
    lr = T.downward_img_quality(d, 16, 32, noise=noise, sigma=sigma).cpu().numpy()
    ref = np.stack([R.downward_img_quality(img[n], 16, 32, noise[n].cpu().numpy(), float(sigma[n])) for n in range(2)])
    assert np.array_equal(lr, ref)
    # add_noise: no resize
    nz = torch.randn(2, 3, 64, 128, device="cuda", generator=g)
    an = T.add_noise(d, noise=nz, sigma=sigma).cpu().numpy()
    ref = np.stack([(R.to_tensor(img[n]) + (nz[n].cpu().numpy() * np.float32(float(sigma[n]))).astype(np.float32)) for n in range(2)])
    assert np.array_equal(an, ref.astype(np.float32))


def test_random_degradation_statistics_and_defaults(T):
    """Default sizes follow src/variables.py (512 x 1024 clips, / 4 for LR); the noise has zero mean and a per-image
    standard deviation inside U(0, 0.03)."""
    img = torch.full((4, 96, 160, 3), 128, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    hr, lr = T.synthesize_pair(img, generator=g)
    assert tuple(hr.shape) == (4, 3, 512, 1024) and tuple(lr.shape) == (4, 3, 128, 256)
    assert float((hr - 128 / 255).abs().max()) < 1e-6            # constant image stays constant under either filter
    resid = lr - 128 / 255
    assert abs(float(resid.mean())) < 2e-4
    sd = resid.flatten(1).std(dim=1)
    assert float(sd.max()) < 0.0305 and float(sd.min()) >= 0.0 and float(sd.max() - sd.min()) > 1e-4
    with pytest.raises(RuntimeError):
        T.resize_u8(img.cpu(), (8, 8))
    with pytest.raises(RuntimeError):
        T.resize_u8(img.float(), (8, 8))
