"""CPU: the numpy oracle of the depth-aware effects against the golden vectors produced by the reference's own
PostProcessor (oracle/make_golden_effects.py), and its two OpenCV restatements against cv2 when it is importable."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import effects_oracle as eo  # noqa: E402


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(ROOT, "tests", "golden", "effects.npz"))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_fog_and_hologram_bit_exact(fx, tag):
    img, depth = fx[f"img_{tag}"], fx[f"depth_{tag}"]
    assert np.array_equal(eo.fog(img, depth, fog_start=0.1), fx[f"fog_{tag}"])
    lines = [tuple(int(v) for v in r) for r in fx[f"lines_{tag}"]]
    assert np.array_equal(eo.hologram(img, depth, 50, fx[f"noise_{tag}"], lines), fx[f"holo_{tag}"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_toon(fx, tag):
    got, mag = eo.toon(fx[f"img_{tag}"], fx[f"depth_{tag}"], 5, 1.0)
    assert float((got != fx[f"toon_{tag}"]).any(axis=-1).mean()) <= 2e-3
    assert np.array_equal(mag, fx[f"mag_toon_{tag}"])
    assert np.abs(eo.bilateral(eo.normalise_depth(fx[f"depth_{tag}"])) - fx[f"bilateral_{tag}"]).max() <= 2e-6


def test_opencv_restatements_match_cv2(fx):
    cv2 = pytest.importorskip("cv2")
    dn = eo.normalise_depth(fx["depth_a"])
    gx, gy = cv2.Sobel(dn, cv2.CV_32F, 1, 0, ksize=3), cv2.Sobel(dn, cv2.CV_32F, 0, 1, ksize=3)
    assert np.abs(np.sqrt(gx ** 2 + gy ** 2) - eo.sobel_magnitude(dn)).max() <= 1e-6
    assert np.abs(cv2.bilateralFilter(dn, 9, 75, 75) - eo.bilateral(dn)).max() <= 2e-6


def test_edge_cases():
    img = np.full((5, 7, 3), 200, np.uint8)
    flat = np.full((5, 7), 0.5, np.float32)              # max <= 1: no normalisation; no gradient anywhere
    assert eo.sobel_magnitude(flat).max() == 0.0
    t, _ = eo.toon(img, flat)
    assert np.array_equal(t, np.full_like(img, int(np.floor(200 / 255.0 * 5) / 5 * 255.0)))
    h = eo.hologram(img, None)
    assert h.shape == img.shape and h[..., 2].max() <= int(200 / 255 * 0.2 * 255) + 1
