import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def rt():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def scene_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("scenes"))


@pytest.fixture(scope="session")
def have_ref():
    """True when the unmodified reference was compiled into oracle/_ref (always here; prebuilt on the GPU box)."""
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_raytracer"))


def block_means(img, b=8):
    h, w, _ = img.shape
    return img[: h // b * b, : w // b * b].reshape(h // b, b, w // b, b, 3).mean(axis=(1, 3))


def assert_statistically_equal(ref_a, ref_b, img, b=8, k=1.6):
    """Statistical parity against an irreproducible reference: `ref_a`, `ref_b` are two runs of the
    reference (own, unseeded RNG), `img` the render under test.  Per-block deviations from the reference mean must be
    of the size of the reference's own run-to-run differences (medians: robust against the heavy-tailed 1/d^2 and sun
    fireflies), and the global means must agree to 2 % + 3x the reference's own spread."""
    import numpy as np
    a, bb, g = block_means(ref_a, b), block_means(ref_b, b), block_means(img, b)
    noise = np.median(np.abs(a - bb))
    dev = np.median(np.abs(g - 0.5 * (a + bb)))
    # unbiased g: dev/noise ~ sqrt(1.5)/sqrt(2) = 0.87
    assert dev <= k * noise + 2e-3 * np.abs(a).mean(), (dev, noise)
    ma, mb, mg = ref_a.mean(), ref_b.mean(), img.mean()
    assert abs(mg - 0.5 * (ma + mb)) <= 0.02 * ma + 3 * abs(ma - mb), (mg, ma, mb)
