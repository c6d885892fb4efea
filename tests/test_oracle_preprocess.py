"""Pre-processing: scale selection, cv2-faithful bilinear resize, padding (lib/utils/test_utils.py,
lib/utils/blob.py, lib/test.py:30-38,131-137)."""
import numpy as np
import pytest

from oracle import preprocess as P


def test_pyramid_scales_1024():
    s = P.pyramid_scales((1024, 1024, 3))
    assert np.allclose(s, [0.09765625, 0.29296875, 0.5859375, 0.9765625, 1.3671875])
    # long side capped: 768x1024 -> base = 800/768 but round(1066.7)=1067 <= 1200 so no cap
    assert np.isclose(P.compute_scaling_factor((768, 1024, 3), 800, 1200), 800 / 768)
    assert np.isclose(P.compute_scaling_factor((500, 1024, 3), 800, 1200), 1200 / 1024)


@pytest.mark.parametrize("hw", [(224, 224), (100, 137), (480, 640)])
def test_resize_restatement_matches_cv2(hw):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    im = rng.randint(0, 256, hw + (3,)).astype(np.uint8)
    imc = im.astype(np.float32) - P.PIXEL_MEANS
    assert imc.dtype == np.float64
    for s in P.pyramid_scales(im.shape) + [0.5, 2.0, 1.3333]:
        a = P.resize_linear_cv2(imc, s, s)
        b = P.resize_linear(imc, s, s)
        assert a.shape == b.shape == (int(np.rint(hw[0] * s)), int(np.rint(hw[1] * s)), 3)
        assert np.abs(a - b).max() < 1e-7
        a32, b32 = a.astype(np.float32), b.astype(np.float32)
        assert (a32 != b32).mean() < 1e-3
        assert np.abs(a32 - b32).max() <= 2e-5           # <= 1 float32 ulp at |x| < 256


def test_blobs_and_padding():
    im = np.random.RandomState(3).randint(0, 256, (50, 70, 3)).astype(np.uint8)
    blobs = P.get_image_blobs(im, [1.0, 0.5])
    assert blobs[0].shape == (1, 3, 50, 70) and blobs[0].dtype == np.float32
    assert np.allclose(blobs[0][0, 1], im[:, :, 1].astype(np.float64) - 115.9465, atol=1e-5)
    assert blobs[1].shape == (1, 3, 25, 35)
    p = P.pad_to_multiple(blobs[0])
    assert p.shape == (1, 3, 64, 80) and np.all(p[:, :, 50:] == 0) and np.all(p[:, :, :, 70:] == 0)
