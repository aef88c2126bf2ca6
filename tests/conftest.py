import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "css_golden_small.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def small_weights():
    from oracle import css_oracle as O
    return O.random_weights(seed=1, d_model=128, n_heads=2, d_ff=256, n_blocks=2)


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


@pytest.fixture(scope="session")
def golden_t186():
    """Reference-generated fixture at the production segment shape (tests/golden/make_golden_t186.py)."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "css_golden_t186.npz")))


def t186_inputs(g):
    """Regenerates the seeded inputs of css_golden_t186.npz and checks them against the stored checksums."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import inputs
    x = inputs.conditioned_mixture(int(g["n_samples"]))
    assert inputs.checksum(x) == int(g["mixture_crc"]), "numpy regenerated a different mixture than the fixture was made from"
    masks = inputs.synthetic_segment_masks(4, 186, 93)
    assert inputs.checksum(masks) == int(g["masks_crc"]), "numpy regenerated different plug-in masks than the fixture was made from"
    return x, masks


def ipd_flip_report(f: np.ndarray, ref: np.ndarray):
    """Features [T, 257 * (1 + pairs)] against the reference's: an entry is a *flip* when the two differ by ~2 pi, which can
    only happen where the IPD atan2(yi - mean, yr - mean) (feature.py:217-222) sits on the +-pi cut: always in the real-valued
    DC and Nyquist bins with a negative real part (their imaginary parts are 1e-7-level residues whose sign hangs on the last
    bit of the STFT), and by chance (~1e-7 per entry) anywhere else.  Returns (flip mask, bins in which flips occur, max
    |difference| over the non-flipped entries); asserts that every flipped entry is within 1e-3 of +-pi on both sides."""
    d = np.abs(f - ref)
    flips = d > 6.0
    assert np.all(np.abs(np.abs(f[flips]) - np.pi) < 1e-3) and np.all(np.abs(np.abs(ref[flips]) - np.pi) < 1e-3), \
        "a feature differs by more than 6 away from the +-pi cut: not an IPD sign flip"
    col = np.nonzero(flips.any(axis=0))[0]
    return flips, sorted(set((col % 257).tolist())), float(d[~flips].max())
