"""CPU oracle for the CSS hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-numpy restatement of the algorithm the reference implements in
``css/css.py::separate_and_stitch`` and everything below it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module, and only as the *checker* (or as the
timed CPU baseline).  The product path (``notsofar_b200``) never imports it.

Parity pinning: every function below is checked against the reference's own
code, executed in the build container through ``oracle/reference_shim.py``, by
``tests/golden/make_golden.py`` (fixtures committed under ``tests/golden/``) and
by ``tests/test_oracle_pinned.py``; the reference's two known-answer tests on
this path (``utils/numpy_utils.py:16-22`` morphology, ``css/training/losses.py:
109-123`` PIT) are re-run against this restatement as well.

Every function takes ``dtype`` = np.float32 (the reference's arithmetic) or
np.float64 (the "fp64-lifted" reference of SURVEY.md section 8c, used as the
ground truth for the ill-conditioned MVDR solve).

All citations are relative to the reference repository root.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

FRAME_LEN = 512
FRAME_HOP = 256
NUM_BINS = 257
EPS32 = float(np.finfo(np.float32).eps)          # feature.py:15  EPSILON = th.finfo(th.float32).eps


def _cdtype(dtype):
    return np.complex64 if np.dtype(dtype) == np.float32 else np.complex128


# --------------------------------------------------------------------------- STFT / iSTFT

def hann_periodic(n: int = FRAME_LEN, dtype=np.float64) -> np.ndarray:
    """th.hann_window(frame_len) is the *periodic* hann (feature.py:29)."""
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(dtype)


def _exact_trig() -> Tuple[np.ndarray, np.ndarray]:
    """cos/sin(2 pi k n / 512) for k < 257, n < 512 in float64, with the integer product
    reduced mod 512 and the exact zeros of the quarter-circle kept exact (an FFT of the
    identity, which is how the reference builds its kernel, has exact zeros there; the
    DC and Nyquist rows of the imaginary kernel are identically zero)."""
    n = np.arange(FRAME_LEN)
    k = np.arange(NUM_BINS)
    m = (k[:, None] * n[None, :]) % FRAME_LEN
    ang = 2.0 * np.pi * m.astype(np.float64) / FRAME_LEN
    c, s = np.cos(ang), np.sin(ang)
    c[(m == FRAME_LEN // 4) | (m == 3 * FRAME_LEN // 4)] = 0.0
    s[(m == 0) | (m == FRAME_LEN // 2)] = 0.0
    return c, s


def stft_kernel(dtype=np.float32) -> Tuple[np.ndarray, np.ndarray]:
    """Real/imag analysis kernels [257, 512]  (feature.py:19-45 with window='hann', S=1).

    K = view_as_real(rfft(eye(N)))[:frame_len] * W: row k of the real part is
    cos(2 pi k n / N) w[n], of the imaginary part -sin(2 pi k n / N) w[n].
    The integer product k*n is reduced mod N before the trigonometric call so the
    table is exact to the last bit of ``dtype``.
    """
    c, s = _exact_trig()
    w = hann_periodic(FRAME_LEN, np.float64)
    return (c * w).astype(dtype), (-s * w).astype(dtype) + 0.0


def istft_kernel(dtype=np.float32) -> Tuple[np.ndarray, np.ndarray]:
    """Synthesis kernels [257, 512] (feature.py:19-45 with the *default* window
    'sqrt_hann' and normalize=True, because FeatureExtractor does not forward
    ``window`` to iSTFT, feature.py:422-425): S = 0.5*sqrt(N*N/hop) = 16."""
    c, sn = _exact_trig()
    w = np.sqrt(hann_periodic(FRAME_LEN, np.float64))
    s = 0.5 * math.sqrt(FRAME_LEN * FRAME_LEN / FRAME_HOP)
    return (c / s * w).astype(dtype), (-sn / s * w).astype(dtype) + 0.0


def num_frames(n_samples: int) -> int:
    """conv1d with stride 256, kernel 512, no padding (feature.py:105/116)."""
    return 0 if n_samples < FRAME_LEN else (n_samples - FRAME_LEN) // FRAME_HOP + 1


def stft(x: np.ndarray, dtype=np.float32) -> np.ndarray:
    """ConformerCssWrapper.stft (conformer_wrapper.py:106-129) -> STFT.forward (feature.py:88-128).

    x: [N, C] real.  Returns [F=257, T, C] complex.  The reference goes through
    (mag, atan2) and th.polar; that round trip is kept.
    """
    x = np.asarray(x, dtype=dtype)
    n, c = x.shape
    t = num_frames(n)
    kr, ki = stft_kernel(dtype)
    idx = np.arange(t)[:, None] * FRAME_HOP + np.arange(FRAME_LEN)[None, :]
    out = np.empty((NUM_BINS, t, c), dtype=_cdtype(dtype))
    for ch in range(c):
        frames = x[:, ch][idx]                       # [T, 512]
        r = frames @ kr.T                            # [T, 257]
        i = frames @ ki.T
        # DC / Nyquist rows of the imaginary kernel are identically zero: the reference's conv
        # yields +0 there [probed: atan2 -> +pi for negative real parts], so pin the sign.
        i[:, 0] = 0.0
        i[:, NUM_BINS - 1] = 0.0
        mag = np.sqrt(r * r + i * i)                 # feature.py:126
        pha = np.arctan2(i, r)                       # feature.py:127
        out[:, :, ch] = (mag * np.cos(pha) + 1j * (mag * np.sin(pha))).T   # th.polar, conformer_wrapper.py:124
    return out


def istft(s: np.ndarray, dtype=np.float32) -> np.ndarray:
    """ConformerCssWrapper.istft (conformer_wrapper.py:131-146) -> iSTFT.forward (feature.py:138-167).

    s: [B, F, T] complex -> [B, (T-1)*256+512] real.  conv_transpose1d == overlap-add
    of K^T-weighted frames; no window-sum normalisation, no one-sided doubling.
    """
    s = np.asarray(s)
    b, f, t = s.shape
    kr, ki = istft_kernel(dtype)
    mag = np.abs(s).astype(dtype)
    pha = np.angle(s).astype(dtype)
    r = mag * np.cos(pha)                            # feature.py:157
    i = mag * np.sin(pha)                            # feature.py:158
    n_out = (t - 1) * FRAME_HOP + FRAME_LEN if t > 0 else 0
    out = np.zeros((b, n_out), dtype=dtype)
    for bi in range(b):
        frames = r[bi].T @ kr + i[bi].T @ ki         # [T, 512]
        for off in (0, 1):                           # frames 2j+off do not overlap each other
            sel = frames[off::2]
            if len(sel) == 0:
                continue
            seg = sel.reshape(-1)
            start = off * FRAME_HOP
            out[bi, start:start + seg.size] += seg
    return out


# --------------------------------------------------------------------------- features

def css_features(stft_seg: np.ndarray, dtype=np.float32) -> np.ndarray:
    """ConformerCssWrapper.separate front half (conformer_wrapper.py:91-94) +
    FeatureExtractor.forward (feature.py:543-569): MVN magnitude of mic 0
    (compute_spectra, feature.py:478-508) and mean-normalised IPD, version 1, of
    mics 1..6 against mic 0 (IPDFeature.forward, feature.py:198-249).

    stft_seg: [F, T, C] complex.  Returns [T, 257*(1 + (C-1))] (time-major rows, the
    layout the mask network consumes after its transpose, conformer.py:294-295).
    """
    f_, t_, c_ = stft_seg.shape
    mag = np.abs(stft_seg).astype(dtype)             # stft.abs()
    # stft.angle().  The DC and Nyquist bins are real up to the sin(pi_f32) residue th.polar
    # leaves, so their phases sit one ulp from -pi and the *sign* of every IPD there hangs on
    # the last bit of atan2/sin.  torch's float32 kernels are correctly rounded at those
    # arguments [probed: angle -> -3.1415925, sin(3.1415925) -> 1.509958e-07]; numpy's float32
    # arctan2 is not (-3.1415927), so evaluate in float64 and round once.
    pha = np.angle(stft_seg.astype(np.complex128)).astype(dtype)
    eps = dtype(EPS32)
    f0 = np.maximum(mag[:, :, 0], eps)               # th.clamp(mag[:,0], min=EPSILON)
    mean = f0.mean(axis=1, keepdims=True, dtype=dtype)
    std = f0.std(axis=1, keepdims=True, ddof=1, dtype=dtype)   # torch.std is unbiased
    feat = [(f0 - mean) / (std + eps)]
    for m in range(1, c_):
        d = pha[:, :, m] - pha[:, :, 0]
        yr, yi = np.cos(d.astype(np.float64)).astype(dtype), np.sin(d.astype(np.float64)).astype(dtype)
        yrm = yr.mean(axis=1, keepdims=True, dtype=dtype)
        yim = yi.mean(axis=1, keepdims=True, dtype=dtype)
        feat.append(np.arctan2(yi - yim, yr - yrm))  # ipd_mean_normalize_version == 1
    return np.concatenate(feat, axis=0).T.astype(dtype)        # [T, 1799]


# --------------------------------------------------------------------------- mask network

@dataclass
class NetDims:
    d_model: int
    n_heads: int
    d_ff: int
    n_blocks: int
    kernel_size: int
    in_features: int
    n_out: int              # num_bins * (num_spks + num_nois)

    @property
    def d_k(self) -> int:
        return self.d_model // self.n_heads


_P = "executor.nnet."


def net_dims(w: Dict[str, np.ndarray]) -> NetDims:
    d_model, in_features = w[_P + "conformer.embed.0.weight"].shape
    d_k = w[_P + "conformer.pos_emb.pe_k.weight"].shape[1]
    d_ff = w[_P + "conformer.encoders.0.feed_forward_in.net.0.weight"].shape[0]
    ks = w[_P + "conformer.encoders.0.conv.dw_conv_1d.weight"].shape[2]
    n_blocks = 0
    while (_P + f"conformer.encoders.{n_blocks}.layer_norm.weight") in w:
        n_blocks += 1
    return NetDims(d_model, d_model // d_k, d_ff, n_blocks, ks, in_features, w[_P + "linear.weight"].shape[0])


def _ln(x, g, b, dtype):
    mu = x.mean(axis=-1, keepdims=True, dtype=dtype)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True, dtype=dtype)     # biased, nn.LayerNorm
    return (x - mu) / np.sqrt(var + dtype(1e-5)) * g + b


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def conformer_masks(w: Dict[str, np.ndarray], feat: np.ndarray, dtype=np.float32) -> np.ndarray:
    """ConformerCSS.forward (conformer.py:287-310) in eval mode (dropout identity, BN running stats).

    feat: [B, T, in_features].  Returns masks [B, n_out/257, 257, T]: rows 0..num_spks-1
    are the speaker masks, the rest the noise masks (torch.chunk along F, conformer.py:308-309).
    """
    W = {k: np.asarray(v, dtype=dtype) for k, v in w.items() if np.asarray(v).dtype.kind == "f"}
    d = net_dims(w)
    x = np.asarray(feat, dtype=dtype)
    b_, t_, _ = x.shape
    x = (x + W[_P + "input_bias"].reshape(1, 1, -1)) * W[_P + "input_scale"].reshape(1, 1, -1)   # conformer.py:297-299
    c = _P + "conformer."
    x = x @ W[c + "embed.0.weight"].T + W[c + "embed.0.bias"]                        # conformer.py:205-210
    x = np.maximum(_ln(x, W[c + "embed.1.weight"], W[c + "embed.1.bias"], dtype), 0)  # Linear -> LN -> (Dropout) -> ReLU
    # relative positions (conformer.py:229-233, 23-29): pos_k[t1, t2] = pe_k[clamp(t1 - t2, -1000, 999) + 1000]
    pe = W[c + "pos_emb.pe_k.weight"]
    maxlen = pe.shape[0] // 2
    rel = np.clip(np.arange(t_)[:, None] - np.arange(t_)[None, :], -maxlen, maxlen - 1) + maxlen
    pos_k = pe[rel]                                                                   # [T, T, d_k]
    pos_kT = np.ascontiguousarray(pos_k.transpose(0, 2, 1))
    inv_sqrt_dk = dtype(1.0 / math.sqrt(d.d_k))
    pad = (d.kernel_size - 1) // 2
    for l in range(d.n_blocks):
        p = c + f"encoders.{l}."

        def ff(x, q):                                                                 # FeedForward.forward conformer.py:146-150
            h = _ln(x, W[q + "layer_norm.weight"], W[q + "layer_norm.bias"], dtype)
            h = np.maximum(h @ W[q + "net.0.weight"].T + W[q + "net.0.bias"], 0)
            return h @ W[q + "net.3.weight"].T + W[q + "net.3.bias"]

        x = x + dtype(0.5) * ff(x, p + "feed_forward_in.")                            # conformer.py:179
        # MultiHeadedAttention.forward conformer.py:57-92
        q_ = p + "self_attn."
        h = _ln(x, W[q_ + "layer_norm.weight"], W[q_ + "layer_norm.bias"], dtype)
        qh = (h @ W[q_ + "linear_q.weight"].T + W[q_ + "linear_q.bias"]).reshape(b_, t_, d.n_heads, d.d_k).transpose(0, 2, 1, 3)
        kh = (h @ W[q_ + "linear_k.weight"].T + W[q_ + "linear_k.bias"]).reshape(b_, t_, d.n_heads, d.d_k).transpose(0, 2, 1, 3)
        vh = (h @ W[q_ + "linear_v.weight"].T + W[q_ + "linear_v.bias"]).reshape(b_, t_, d.n_heads, d.d_k).transpose(0, 2, 1, 3)
        A = qh @ kh.transpose(0, 1, 3, 2)                                             # [B, H, T, T]
        # conformer.py:74-77: reshape_q [T, B*H, d_k] @ pos_k^T [T, d_k, T] -> [T, B*H, T] -> [B, H, T, T]
        rq = qh.reshape(b_ * d.n_heads, t_, d.d_k).transpose(1, 0, 2)
        Bm = np.matmul(rq, pos_kT).transpose(1, 0, 2).reshape(b_, d.n_heads, t_, t_)
        s = (A + Bm) * inv_sqrt_dk
        s = s - s.max(axis=-1, keepdims=True)
        e = np.exp(s)
        pr = e / e.sum(axis=-1, keepdims=True, dtype=dtype)
        o = (pr @ vh).transpose(0, 2, 1, 3).reshape(b_, t_, d.d_model)
        x = x + (o @ W[q_ + "linear_out.weight"].T + W[q_ + "linear_out.bias"])       # conformer.py:180
        # ConvModule.forward conformer.py:113-127
        q_ = p + "conv."
        h = _ln(x, W[q_ + "layer_norm.weight"], W[q_ + "layer_norm.bias"], dtype)
        w1 = W[q_ + "pw_conv_1.weight"].reshape(2)
        b1 = W[q_ + "pw_conv_1.bias"].reshape(2)
        u = (w1[0] * h + b1[0]) * _sigmoid(w1[1] * h + b1[1])                          # scalar 1x1 Conv2d(1->2) + GLU
        up = np.pad(u, ((0, 0), (pad, pad), (0, 0)))
        dw = W[q_ + "dw_conv_1d.weight"].reshape(d.d_model, d.kernel_size)
        acc = np.zeros_like(u)
        for j in range(d.kernel_size):                                                # depthwise cross-correlation over time
            acc += up[:, j:j + t_, :] * dw[:, j]
        acc += W[q_ + "dw_conv_1d.bias"]
        acc = (acc - W[q_ + "BN.running_mean"]) / np.sqrt(W[q_ + "BN.running_var"] + dtype(1e-5)) * W[q_ + "BN.weight"] + W[q_ + "BN.bias"]
        acc = np.maximum(acc, 0)
        acc = W[q_ + "pw_conv_2.weight"].reshape(()) * acc + W[q_ + "pw_conv_2.bias"].reshape(())
        x = x + acc                                                                   # conformer.py:181
        x = x + dtype(0.5) * ff(x, p + "feed_forward_out.")                           # conformer.py:182
        x = _ln(x, W[p + "layer_norm.weight"], W[p + "layer_norm.bias"], dtype)       # conformer.py:184
    m = _sigmoid(x @ W[_P + "linear.weight"].T + W[_P + "linear.bias"])               # conformer.py:302-304
    n_masks = d.n_out // NUM_BINS
    return m.reshape(b_, t_, n_masks, NUM_BINS).transpose(0, 2, 3, 1).astype(dtype)   # [B, 4, F, T]


# --------------------------------------------------------------------------- MVDR

def make_wta(spk_masks: np.ndarray, noise_masks: np.ndarray) -> np.ndarray:
    """mvdr_util.py:50-55 -- winner-take-all over {speakers, sum of noise masks}."""
    noise = noise_masks.sum(axis=0, keepdims=True)
    m = np.concatenate([spk_masks, noise], axis=0)
    mx = m.max(axis=0, keepdims=True)
    return np.where(m == mx, m, m.dtype.type(1e-10))


def mask_scm(mix_ftc: np.ndarray, mask_ft: np.ndarray) -> np.ndarray:
    """mvdr_util.py:58-66: R[f] = sum_t m[f,t] x[f,t] x[f,t]^H + 1e-15 I  (the in-place
    += keeps the dtype of the einsum result, so complex64 stays complex64)."""
    r = np.einsum("ft,ftm,ftn->fmn", mask_ft, mix_ftc, mix_ftc.conj())
    r += (1e-15 * np.eye(mix_ftc.shape[2]))[None].astype(r.dtype)
    return r


def bf_coeffs(noi_scm: np.ndarray, tgt_scm: np.ndarray) -> np.ndarray:
    """mvdr_util.py:69-75: W = (solve(N, R) / trace(solve(N, R)))[:, :, 0]; the 1e-15 only
    reaches bin 0 (den[0] += 1e-15)."""
    num = np.linalg.solve(noi_scm, tgt_scm)
    den = np.trace(num, axis1=-2, axis2=-1)[..., None, None]
    den[0] += 1e-15
    return (num / den)[..., 0]


def make_mvdr(spk_masks: np.ndarray, noise_masks: np.ndarray, mix_stft: np.ndarray, dtype=np.float32) -> np.ndarray:
    """mvdr_util.py:5-47 with return_stft=True.

    spk_masks [S, F, T], noise_masks [Nn, F, T] real; mix_stft [C, F, T] complex.
    Returns [S, F, T] complex.  dtype=np.float64 is the fp64-lifted reference.
    """
    cd = _cdtype(dtype)
    spk = np.asarray(spk_masks, dtype=dtype)
    noi = np.asarray(noise_masks, dtype=dtype)
    mix = np.asarray(mix_stft, dtype=cd)
    allm = make_wta(spk, noi)
    L = min(allm.shape[-1], mix.shape[-1])
    mix, allm = mix[:, :, :L], allm[:, :, :L]
    mix_ftc = np.ascontiguousarray(mix.transpose(1, 2, 0))
    scms = [mask_scm(mix_ftc, m) for m in allm]
    spk_scms = np.stack(scms[:-1])
    noise_scm = scms[-1]
    out = []
    n_spk = spk_scms.shape[0]
    for i in range(n_spk):
        other = spk_scms[np.arange(n_spk) != i].sum(axis=0)
        wcoef = bf_coeffs(noise_scm + other, spk_scms[i])           # [F, C]
        out.append(np.einsum("fc,ftc->ft", wcoef.conj(), mix_ftc))  # get_bf, mvdr_util.py:78-80
    return np.stack(out).astype(cd)


# --------------------------------------------------------------------------- stitching

def calc_segment_weight(seg_frames: int, m0: int, m1: int, is_first: bool = False, is_last: bool = False) -> np.ndarray:
    """css.py:341-390 -- trapezoid weights; edges of the first/last segment are 0.1."""
    assert seg_frames > 2 * m1
    w = np.ones(seg_frames, dtype=np.float32)
    w[:m0] = 0
    w[seg_frames - m0:] = 0
    n = m1 - m0
    # torch.linspace(0.1, 1, n) in float32: symmetric evaluation start + i*step / end - (n-1-i)*step
    lin = linspace_f32(0.1, 1.0, n)
    w[m0:m1] = lin
    w[seg_frames - m1:seg_frames - m0] = lin[::-1]
    if is_first:
        w[:m0] = 0.1
    if is_last:
        w[seg_frames - m0:] = 0.1
    return w


def linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace for float32 on CPU: step = (end-start)/(steps-1) in float32;
    the first half counts up from start, the second half counts down from end."""
    if steps == 1:
        return np.array([start], dtype=np.float32)
    s, e = np.float32(start), np.float32(end)
    step = np.float32((e - s) / np.float32(steps - 1))
    half = steps // 2
    out = np.empty(steps, dtype=np.float32)
    for i in range(steps):
        if i < half:
            out[i] = np.float32(s + np.float32(step * np.float32(i)))
        else:
            out[i] = np.float32(e - np.float32(step * np.float32(steps - i - 1)))
    return out


def pit_cost_l1(left: np.ndarray, right: np.ndarray, dtype=np.float32) -> np.ndarray:
    """PitWrapper._opt_perm_loss with l1_loss (losses.py:50-71, 104-106):
    C[i, j] = mean_{f,t} |left[..., i] - right[..., j]|;  left/right: [F, T_ov, S]."""
    s = left.shape[-1]
    c = np.empty((s, s), dtype=dtype)
    for i in range(s):
        for j in range(s):
            c[i, j] = np.abs(left[..., i].astype(dtype) - right[..., j].astype(dtype)).mean(dtype=dtype)
    return c


def pit_cost_mse(left: np.ndarray, right: np.ndarray, dtype=np.float32) -> np.ndarray:
    """losses.py:100-102 variant."""
    s = left.shape[-1]
    c = np.empty((s, s), dtype=dtype)
    for i in range(s):
        for j in range(s):
            c[i, j] = ((left[..., i].astype(dtype) - right[..., j].astype(dtype)) ** 2).mean(dtype=dtype)
    return c


def assign(cost: np.ndarray) -> np.ndarray:
    """PitWrapper._fast_pit (losses.py:32-48): scipy Hungarian; returns right_perm with
    new channel k <- old channel perm[k]."""
    from scipy.optimize import linear_sum_assignment
    li, ri = linear_sum_assignment(np.asarray(cost))
    assert (li == np.arange(len(li))).all()
    return ri


def assign_bruteforce(cost: np.ndarray) -> np.ndarray:
    """arg-min over all permutations (first minimum in lexicographic order)."""
    s = cost.shape[0]
    best, best_p = None, None
    for p in itertools.permutations(range(s)):
        v = sum(float(cost[i, p[i]]) for i in range(s))
        if best is None or v < best:
            best, best_p = v, p
    return np.array(best_p)


def dilate(arr: np.ndarray, iters: int) -> np.ndarray:
    """numpy_utils.py:10-13: sliding max over a window of 2*iters+1, zero padding."""
    a = np.pad(arr, iters, mode="constant", constant_values=0)
    return np.lib.stride_tricks.sliding_window_view(a, 2 * iters + 1).max(1)


def erode(arr: np.ndarray, iters: int) -> np.ndarray:
    """numpy_utils.py:4-7: sliding min over a window of 2*iters+1, one padding."""
    a = np.pad(arr, iters, mode="constant", constant_values=1)
    return np.lib.stride_tricks.sliding_window_view(a, 2 * iters + 1).min(1)


# --------------------------------------------------------------------------- the whole path

@dataclass
class OracleCfg:
    """The CssCfg fields (css.py:24-48) the numeric path reads."""
    segment_size_sec: float = 3.0
    hop_size_sec: float = 1.5
    normalize_segment_power: bool = False
    stitching_loss: str = "l1"
    stitching_input: str = "mask"
    seg_weight_m0_sec: float = 0.15
    seg_weight_m1_sec: float = 0.3
    activity_th: float = 0.4
    activity_dilation_sec: float = 0.4
    activity_erosion_sec: float = 0.2
    num_spks: int = 3
    mc_mvdr: bool = True
    mc_mask_floor_db: float = 0.0
    sc_mask_floor_db: float = -np.inf


@dataclass
class SegmentPlan:
    """Integer frame bookkeeping of css.py:141-169."""
    segment_frames: int
    hop_frames: int
    m0_frames: int
    m1_frames: int
    dilation_frames: int
    erosion_frames: int
    mix_frames: int          # after the short-input zero padding (css.py:159-164)
    raw_frames: int          # before it
    num_segments: int

    @property
    def overlap_frames(self) -> int:
        return self.segment_frames - self.hop_frames


def plan_segments(n_samples: int, fs: int, cfg) -> SegmentPlan:
    seg = num_frames(int(cfg.segment_size_sec * fs))
    hop = int(seg * cfg.hop_size_sec / cfg.segment_size_sec)
    m0 = int(seg * cfg.seg_weight_m0_sec / cfg.segment_size_sec)
    m1 = int(seg * cfg.seg_weight_m1_sec / cfg.segment_size_sec)
    dil = int(seg * cfg.activity_dilation_sec / cfg.segment_size_sec)
    ero = int(seg * cfg.activity_erosion_sec / cfg.segment_size_sec)
    raw = num_frames(n_samples)
    mix = max(raw, seg)
    nseg = int(np.ceil((mix - (seg - hop)) / hop))
    return SegmentPlan(seg, hop, m0, m1, dil, ero, mix, raw, nseg)


def separate_and_stitch(speech_mix: np.ndarray, weights: Dict[str, np.ndarray], fs: int = 16000,
                        cfg: Optional[OracleCfg] = None, dtype=np.float32, mvdr_dtype=None,
                        masks_override: Optional[np.ndarray] = None, batch: int = 8,
                        return_stages: bool = False, stft_override: Optional[np.ndarray] = None):
    """css.py:110-338 restated.  speech_mix [1, N, C] float.  Returns (list of num_spks
    waveforms, side_info) like the reference; with return_stages also the per-segment tensors.

    mvdr_dtype: arithmetic of the MVDR stage (default = dtype).  np.float64 with
    dtype=np.float32 is "fp32 masks, fp64-lifted beamformer" (SURVEY 8c protocol).
    masks_override: [num_segments, 4, F, T] masks to use instead of the network's.
    stft_override: [F, T_long, C] long-form STFT to use instead of stft(x) (stage isolation in the
    parity tests: the ill-conditioned MVDR amplifies even 1e-7 differences of its input).
    """
    cfg = cfg or OracleCfg()
    mvdr_dtype = mvdr_dtype or dtype
    cd = _cdtype(dtype)
    assert speech_mix.ndim == 3 and speech_mix.shape[0] == 1
    x = np.asarray(speech_mix[0], dtype=dtype)
    n, c = x.shape
    plan = plan_segments(n, fs, cfg)
    T, hop = plan.segment_frames, plan.hop_frames
    stft_mix = stft(x, dtype) if stft_override is None else np.asarray(stft_override)[:, :plan.raw_frames].astype(cd)
    if plan.raw_frames < T:                                          # css.py:159-164
        stft_mix = np.pad(stft_mix, ((0, 0), (0, T - plan.raw_frames), (0, 0)))
    mix_frames = plan.mix_frames
    S = cfg.num_spks

    segs = np.zeros((plan.num_segments, NUM_BINS, T, c), dtype=cd)
    for i in range(plan.num_segments):                               # css.py:182-193
        st = i * hop
        en = min(st + T, mix_frames)
        segs[i, :, :en - st] = stft_mix[:, st:en]

    if masks_override is None:
        masks = np.empty((plan.num_segments, net_dims(weights).n_out // NUM_BINS, NUM_BINS, T), dtype=dtype)
        for b0 in range(0, plan.num_segments, batch):
            feats = np.stack([css_features(s, dtype) for s in segs[b0:b0 + batch]])
            masks[b0:b0 + batch] = conformer_masks(weights, feats, dtype)
    else:
        masks = np.asarray(masks_override, dtype=dtype)

    spk_masks = masks[:, :S]                                         # [n, S, F, T]
    noise_masks = masks[:, S:]
    sep = np.empty((plan.num_segments, S, NUM_BINS, T), dtype=cd)
    floor_db = cfg.mc_mask_floor_db if c > 1 else cfg.sc_mask_floor_db
    mask_floor = dtype(10.0 ** (floor_db / 20.0))
    for i in range(plan.num_segments):
        if c > 1 and cfg.mc_mvdr:                                    # css.py:210-217
            base = make_mvdr(spk_masks[i], noise_masks[i], segs[i].transpose(2, 0, 1), mvdr_dtype).astype(cd)
        else:
            base = np.repeat(segs[i][None, :, :, 0], S, axis=0)
        sep[i] = base * np.maximum(spk_masks[i], mask_floor)         # css.py:223-227
        if cfg.normalize_segment_power:                              # css.py:233-247
            t = min(i * hop + T, mix_frames) - i * hop
            mix_e = np.sqrt(np.mean(np.abs(segs[i][:, :t, 0]) ** 2, dtype=dtype))
            sep_e = np.sqrt(np.mean(np.abs(sep[i][:, :, :t].sum(axis=0)) ** 2, dtype=dtype))
            sep[i] = (mix_e / sep_e) * sep[i]

    # II. permutation chain + weighted overlap-add (css.py:254-299)
    stft_st = np.zeros((NUM_BINS, mix_frames, S), dtype=cd)
    mask_st = np.zeros((NUM_BINS, mix_frames, S), dtype=dtype)
    wg_st = np.zeros(mix_frames, dtype=np.float32)
    perms = np.tile(np.arange(S), (plan.num_segments, 1))
    costs = np.zeros((plan.num_segments, S, S), dtype=dtype)
    spk_masks = spk_masks.copy()
    ov = plan.overlap_frames
    cost_fn = {"l1": pit_cost_l1, "mse": pit_cost_mse}[cfg.stitching_loss]
    for i in range(plan.num_segments):
        if i > 0:
            if cfg.stitching_input == "mask":
                left, right = spk_masks[i - 1], spk_masks[i]
            else:
                left, right = np.abs(sep[i - 1]), np.abs(sep[i])
            costs[i] = cost_fn(left[:, :, T - ov:].transpose(1, 2, 0), right[:, :, :ov].transpose(1, 2, 0), dtype)
            perm = assign(costs[i])
            perms[i] = perm
            spk_masks[i] = spk_masks[i][perm]
            sep[i] = sep[i][perm]
        st = i * hop
        en = min(st + T, mix_frames)
        w = calc_segment_weight(T, plan.m0_frames, plan.m1_frames, is_first=(i == 0),
                                is_last=(i == plan.num_segments - 1 and i > 0))[:en - st]
        wg_st[st:en] += w
        stft_st[:, st:en] += (w[None, :, None] * sep[i][:, :, :en - st].transpose(1, 2, 0)).astype(cd)
        mask_st[:, st:en] += (w[None, :, None] * spk_masks[i][:, :, :en - st].transpose(1, 2, 0)).astype(dtype)
    assert (wg_st > 1e-5).all(), "zero weights found. check hop_size, segment_size or m0, m1"
    stft_st = (stft_st / wg_st[None, :, None]).astype(cd)
    mask_st = (mask_st / wg_st[None, :, None]).astype(dtype)

    # III. activity gating (css.py:303-312)
    activity = mask_st.mean(axis=0, dtype=dtype)                     # [T_long, S]
    activity_b = activity >= dtype(cfg.activity_th)
    activity_final = np.stack([erode(dilate(activity_b[:, k], plan.dilation_frames), plan.erosion_frames)
                               for k in range(S)], axis=1)
    stft_st = stft_st * activity_final[None]
    wavs = istft(np.ascontiguousarray(stft_st.transpose(2, 0, 1)), dtype)      # [S, N']
    side = {
        "mask_stitched": mask_st[None],                 # [1, F, T_long, S]
        "activity_b": activity_b,                       # [T_long, S]
        "activity_final": activity_final[None],         # [1, T_long, S]
        "segment_frames": T,
    }
    if return_stages:
        side.update(stft_mix=stft_mix, segs=segs, masks=masks, sep=sep, perms=perms, costs=costs,
                    stft_stitched=stft_st, wg=wg_st, plan=plan, activity=activity)
    return [wavs[k] for k in range(S)], side


# --------------------------------------------------------------------------- file boundary

def peaknorm(samps: np.ndarray) -> np.ndarray:
    """write_wav max_norm (audio_utils.py:44-45); float64 because np.max(np.abs(float32)) + 1e-7
    promotes nothing in numpy 2 (python float is weak) -- stays float32."""
    samps = np.asarray(samps)
    return samps * 0.99 / (np.max(np.abs(samps)) + 1e-7)


def pcm16(samps: np.ndarray) -> np.ndarray:
    """libsndfile float -> PCM_16 [upstream, not verifiable here]: lrintf(x * 0x7FFF) with
    clipping off by default (values are inside [-0.99, 0.99] after peaknorm)."""
    return np.rint(np.asarray(samps, dtype=np.float32) * np.float32(32767.0)).astype(np.int16)


# --------------------------------------------------------------------------- deterministic weights

def random_weights(seed: int, d_model: int = 512, n_heads: int = 8, d_ff: int = 1024, n_blocks: int = 18,
                   kernel_size: int = 33, in_features: int = 1799, num_spks: int = 3, num_nois: int = 1,
                   gain: float = 1.0) -> Dict[str, np.ndarray]:
    """Seeded weights of the reference architecture, keyed by the reference's state_dict
    names (conformer.py; names probed from ConformerCssWrapper.state_dict()).  numpy PCG64
    so the same arrays can be regenerated on the GPU box without torch RNG parity.
    Distributions are chosen to give non-degenerate masks (not all ~0.5) and BN / LN
    parameters away from identity so every term of the forward pass is exercised."""
    rng = np.random.default_rng(seed)
    d_k = d_model // n_heads

    def lin(o, i, g=1.0):
        return (rng.standard_normal((o, i)) * (g / math.sqrt(i))).astype(np.float32), \
               (rng.standard_normal(o) * 0.1).astype(np.float32)

    def ln(n):
        return (1.0 + 0.1 * rng.standard_normal(n)).astype(np.float32), (0.1 * rng.standard_normal(n)).astype(np.float32)

    w: Dict[str, np.ndarray] = {}
    w[_P + "input_bias"] = (0.05 * rng.standard_normal((1, 1, in_features))).astype(np.float32)
    w[_P + "input_scale"] = (1.0 + 0.05 * rng.standard_normal((1, 1, in_features))).astype(np.float32)
    c = _P + "conformer."
    w[c + "embed.0.weight"], w[c + "embed.0.bias"] = lin(d_model, in_features)
    w[c + "embed.1.weight"], w[c + "embed.1.bias"] = ln(d_model)
    w[c + "pos_emb.pe_k.weight"] = (rng.standard_normal((2000, d_k)) * 0.5).astype(np.float32)
    for l in range(n_blocks):
        p = c + f"encoders.{l}."
        for ffn in ("feed_forward_in.", "feed_forward_out."):
            w[p + ffn + "layer_norm.weight"], w[p + ffn + "layer_norm.bias"] = ln(d_model)
            w[p + ffn + "net.0.weight"], w[p + ffn + "net.0.bias"] = lin(d_ff, d_model, gain)
            w[p + ffn + "net.3.weight"], w[p + ffn + "net.3.bias"] = lin(d_model, d_ff, gain)
        a = p + "self_attn."
        w[a + "layer_norm.weight"], w[a + "layer_norm.bias"] = ln(d_model)
        for nm in ("linear_q", "linear_k", "linear_v", "linear_out"):
            w[a + nm + ".weight"], w[a + nm + ".bias"] = lin(d_model, d_model, gain)
        cv = p + "conv."
        w[cv + "layer_norm.weight"], w[cv + "layer_norm.bias"] = ln(d_model)
        w[cv + "pw_conv_1.weight"] = (1.0 + 0.2 * rng.standard_normal((2, 1, 1, 1))).astype(np.float32)
        w[cv + "pw_conv_1.bias"] = (0.1 * rng.standard_normal(2)).astype(np.float32)
        w[cv + "dw_conv_1d.weight"] = (rng.standard_normal((d_model, 1, kernel_size)) / math.sqrt(kernel_size)).astype(np.float32)
        w[cv + "dw_conv_1d.bias"] = (0.1 * rng.standard_normal(d_model)).astype(np.float32)
        w[cv + "BN.weight"], w[cv + "BN.bias"] = ln(d_model)
        w[cv + "BN.running_mean"] = (0.1 * rng.standard_normal(d_model)).astype(np.float32)
        w[cv + "BN.running_var"] = (0.5 + rng.random(d_model)).astype(np.float32)
        w[cv + "BN.num_batches_tracked"] = np.array(0, dtype=np.int64)
        w[cv + "pw_conv_2.weight"] = (1.0 + 0.2 * rng.standard_normal((1, 1, 1, 1))).astype(np.float32)
        w[cv + "pw_conv_2.bias"] = (0.1 * rng.standard_normal(1)).astype(np.float32)
        w[p + "layer_norm.weight"], w[p + "layer_norm.bias"] = ln(d_model)
    w[_P + "linear.weight"], w[_P + "linear.bias"] = lin(NUM_BINS * (num_spks + num_nois), d_model, 4.0 * gain)
    kr, ki = stft_kernel(np.float32)
    w["executor.extractor.forward_stft.K"] = np.concatenate([kr, ki]).reshape(2 * NUM_BINS, 1, FRAME_LEN)
    kr, ki = istft_kernel(np.float32)
    w["executor.extractor.inverse_stft.K"] = np.concatenate([kr, ki]).reshape(2 * NUM_BINS, 1, FRAME_LEN)
    return w
