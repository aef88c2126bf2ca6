"""Seeded inputs shared by the fixture generators (build container) and the tests that replay them (GPU box).

numpy PCG64 streams only, so the arrays regenerate bit-identically wherever the same numpy is installed; every fixture
that depends on one of them stores a checksum of the int16 / float32 bytes and the tests assert it before comparing.
"""
from __future__ import annotations

import zlib

import numpy as np


def checksum(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def conditioned_mixture(n_samples: int = 100_000, seed: int = 17, noise_db: float = -10.0) -> np.ndarray:
    """[n_samples, 7] float32 mixture on which the reference's *complex64* MVDR is trustworthy: three intermittent white
    talkers seen through different integer inter-microphone delays plus independent white sensor noise only ``-noise_db``
    below a talker.  The sensor noise makes every spatial covariance full rank with a condition number of a few tens in
    every bin (a real array is coherent at low frequencies and reaches 1e5..1e7), so the reference's
    fp32 einsum / LAPACK solve agrees with its own fp64 evaluation to ~1e-5 and its *actual* output can be the ground truth."""
    rng = np.random.default_rng(seed)
    x = np.zeros((n_samples, 7), np.float64)
    t = np.arange(n_samples)
    for s in range(3):
        sig = rng.standard_normal(n_samples + 16)
        period = 16000 * (0.7 + 0.45 * s)
        gate = (np.sin(2 * np.pi * (t + 3000 * s) / period) > -0.2).astype(np.float64)
        delays = rng.integers(0, 7, size=7)
        for c in range(7):
            x[:, c] += sig[8 - delays[c]: 8 - delays[c] + n_samples] * gate
    x += rng.standard_normal((n_samples, 7)) * 10.0 ** (noise_db / 20.0)
    x *= 0.01 / np.sqrt(np.mean(x * x))
    return x.astype(np.float32)


CHANNEL_SHUFFLES = [(0, 1, 2), (2, 0, 1), (1, 0, 2), (2, 1, 0)]


def synthetic_segment_masks(num_segments: int, seg_frames: int, hop_frames: int, seed: int = 23, sharpness: float = 2.5,
                            num_bins: int = 257) -> np.ndarray:
    """[num_segments, 4, F, T] float32 masks of a stand-in separator plug-in (README.md:229-232: separators are plug-ins):
    one long-form softmax field over {3 talkers, noise}, cut into segments, perturbed by 2 % per segment and with the
    talker channels shuffled per segment (CHANNEL_SHUFFLES) so that the permutation chain has work to do.  Winners are
    independent per (bin, frame): every mask wins ~46 of a segment's 186 frames in every bin, which keeps every masked
    covariance of mvdr_util.py:58-66 full rank (a mask that wins fewer than 7 frames leaves a rank-deficient matrix plus
    1e-10 x total, condition 1e10, and the reference's complex64 solve is then noise)."""
    rng = np.random.default_rng(seed)
    t_long = (num_segments - 1) * hop_frames + seg_frames
    logits = rng.standard_normal((4, num_bins, t_long)) * sharpness
    # slow talker activity on top, so that the stitched activity crosses its threshold in runs, not frame by frame
    env = np.sin(2 * np.pi * np.arange(t_long)[None, :] / np.array([140.0, 205.0, 320.0, 1e9])[:, None] + np.arange(4)[:, None])
    logits += 1.5 * env[:, None, :]
    e = np.exp(logits - logits.max(axis=0, keepdims=True))
    field = e / e.sum(axis=0, keepdims=True)
    out = np.empty((num_segments, 4, num_bins, seg_frames), np.float32)
    for i in range(num_segments):
        m = field[:, :, i * hop_frames:i * hop_frames + seg_frames] * (1.0 + 0.02 * rng.standard_normal((4, num_bins, seg_frames)))
        m = np.clip(m, 0.0, 1.0).astype(np.float32)
        sh = list(CHANNEL_SHUFFLES[i % len(CHANNEL_SHUFFLES)])
        out[i, :3] = m[:3][sh]
        out[i, 3] = m[3]
    return out
