"""Seeded synthetic inputs for benchmarks and smoke tests (there is no network for the real checkpoints or
meetings): random weights of the reference's mask-network architecture keyed by its state_dict names, and a
7-channel 16 kHz "meeting" of intermittent band-limited talkers seen by the NOTSOFAR circular array."""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

_P = "executor.nnet."
NUM_BINS = 257


def random_state_dict(seed: int, d_model: int = 512, n_heads: int = 8, d_ff: int = 1024, n_blocks: int = 18,
                      kernel_size: int = 33, in_features: int = 1799, num_spks: int = 3, num_nois: int = 1,
                      gain: float = 1.0) -> Dict[str, np.ndarray]:
    """Defaults = the shipped v1.0 multi-channel model (configs/train_css/local/conformer_v1.0_mc.yaml:36-42:
    attention_dim 512, 8 heads, 18 blocks, linear_units 1024, kernel 33; 59.25 M parameters).  numpy PCG64 so the
    stream is identical on every box."""
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
    return w


def mic_positions_m() -> np.ndarray:
    """utils/mic_array_model.py:4-27: centre mic + 6 on a 4.25 cm circle."""
    pos = np.zeros((7, 3))
    for i in range(1, 7):
        ang = np.deg2rad(60.0 * (i - 1))
        pos[i, 0], pos[i, 1] = 0.0425 * np.cos(ang), 0.0425 * np.sin(ang)
    return pos


def synthetic_meeting(seconds: float, seed: int = 0, fs: int = 16000, n_spk: int = 4, out: np.ndarray = None) -> np.ndarray:
    """[n_samples, 7] float32: talkers = band-limited (100-7000 Hz) noise with 1-10 s talk spurts arriving as plane
    waves (fractional delays by linear interpolation), plus diffuse noise 25 dB down; rms ~ 0.0065 like the
    reference's sample_data mixture."""
    from scipy.signal import butter, lfilter
    rng = np.random.default_rng(seed)
    n = int(round(seconds * fs))
    x = out if out is not None else np.zeros((n, 7), np.float32)
    x[:] = 0
    pos = mic_positions_m()
    b, a = butter(2, [100.0 / (fs / 2), 7000.0 / (fs / 2)], btype="band")
    block = 1 << 22
    for s in range(n_spk):
        az = rng.uniform(0, 2 * np.pi)
        direction = np.array([np.cos(az), np.sin(az), 0.0])
        delays = -(pos @ direction) / 343.0 * fs                      # samples, |d| <= 2
        delays -= delays.min()
        # talk-spurt gate
        gate = np.zeros(n, np.float32)
        t = 0
        on = rng.random() < 0.5
        while t < n:
            dur = int(rng.uniform(1.0, 10.0) * fs)
            if on:
                gate[t:t + dur] = 1.0
            on = not on
            t += dur
        for st in range(0, n, block):
            en = min(n, st + block)
            sig = lfilter(b, a, rng.standard_normal(en - st + 8)).astype(np.float32)
            am = (1.0 + 0.5 * np.sin(2 * np.pi * 4.0 * (np.arange(en - st + 8) + st) / fs)).astype(np.float32)
            sig *= am
            for c in range(7):
                d = delays[c]
                i0 = int(np.floor(d))
                fr = np.float32(d - i0)
                seg = (1 - fr) * sig[6 - i0:6 - i0 + (en - st)] + fr * sig[5 - i0:5 - i0 + (en - st)]
                x[st:en, c] += seg * gate[st:en]
    for st in range(0, n, block):
        en = min(n, st + block)
        x[st:en] += (rng.standard_normal((en - st, 7)) * 10 ** (-25 / 20)).astype(np.float32)
    rms = float(np.sqrt(np.mean(np.square(x[: min(n, 1 << 22)], dtype=np.float64))))
    x *= np.float32(0.0065 / max(rms, 1e-12))
    return x
