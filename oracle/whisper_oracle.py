"""CPU restatement of the Whisper log-mel front end and audio encoder  --  TEST INFRASTRUCTURE (numpy).

Only tests/ may import this module; it is the checker for csrc/whisper.cu, never the thing shipped.

The algorithm lives in a third-party dependency that is absent from /root/reference: openai-whisper, installed by the
reference from ``git+https://github.com/openai/whisper.git`` at an unpinned HEAD (requirements.txt:2; call site
asr/asr.py:69-74).  What is restated here is its published algorithm [upstream]:
  whisper/audio.py  log_mel_spectrogram: torch.stft(n_fft 400, hop 160, hann(400), centred, reflect), |.|^2 of frames [:-1],
                    mel filterbank (librosa slaney), log10(clamp(1e-10)), max(., max - 8), (. + 4) / 4
  whisper/model.py  AudioEncoder / ResidualAttentionBlock / MultiHeadAttention (see csrc/whisper.cu header)
**Parity unpinned** with respect to openai-whisper itself: the pin available offline is the transformers implementation of
the same model (WhisperFeatureExtractor, WhisperModel.encoder), checked in tests/test_whisper.py.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np


def log_mel(audio: np.ndarray, filters: np.ndarray) -> np.ndarray:
    """audio [480000] -> [n_mels, 3000] float32."""
    x = np.asarray(audio, np.float64)
    n_fft, hop = 400, 160
    xp = np.pad(x, (n_fft // 2, n_fft // 2), mode="reflect")
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)
    n_frames = 1 + (len(xp) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    spec = np.fft.rfft(xp[idx] * win, axis=1)[:-1]                       # drop the last frame (stft[..., :-1])
    power = (np.abs(spec) ** 2).T                                        # [201, 3000]
    mel = filters.astype(np.float64) @ power
    log_spec = np.log10(np.maximum(mel, 1e-10))
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return ((log_spec + 4.0) / 4.0).astype(np.float32)


def _ln(x, g, b):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + 1e-5) * g + b


def _gelu(x):
    from scipy.special import erf
    return 0.5 * x * (1.0 + erf(x / math.sqrt(2.0)))


def encoder(w: Dict[str, np.ndarray], mel: np.ndarray, dtype=np.float64) -> np.ndarray:
    """w: openai-whisper encoder names (see notsofar_b200.whisper._canon); mel [n_mels, 3000] -> [1500, d]."""
    W = {k: np.asarray(v, dtype) for k, v in w.items()}
    x = np.asarray(mel, dtype)
    d = W["conv1.weight"].shape[0]

    def conv1d(x, wt, b, stride):
        c_out, c_in, k = wt.shape
        xp = np.pad(x, ((0, 0), (1, 1)))
        t_out = (x.shape[1] + 2 - k) // stride + 1
        cols = np.stack([xp[:, kk:kk + stride * t_out:stride] for kk in range(k)], axis=1)     # [c_in, k, t_out]
        return np.einsum("ock,ckt->ot", wt, cols) + b[:, None]

    x = _gelu(conv1d(x, W["conv1.weight"], W["conv1.bias"], 1))
    x = _gelu(conv1d(x, W["conv2.weight"], W["conv2.bias"], 2))
    x = x.T + W["positional_embedding"]
    n_heads = d // 64
    l = 0
    while f"blocks.{l}.attn.query.weight" in W:
        p = f"blocks.{l}."
        h = _ln(x, W[p + "attn_ln.weight"], W[p + "attn_ln.bias"])
        q = h @ W[p + "attn.query.weight"].T + W[p + "attn.query.bias"]
        k = h @ W[p + "attn.key.weight"].T
        v = h @ W[p + "attn.value.weight"].T + W[p + "attn.value.bias"]
        T = x.shape[0]
        q = q.reshape(T, n_heads, 64).transpose(1, 0, 2) * 64 ** -0.25
        k = k.reshape(T, n_heads, 64).transpose(1, 0, 2) * 64 ** -0.25
        v = v.reshape(T, n_heads, 64).transpose(1, 0, 2)
        s = q @ k.transpose(0, 2, 1)
        s = s - s.max(-1, keepdims=True)
        pr = np.exp(s)
        pr /= pr.sum(-1, keepdims=True)
        o = (pr @ v).transpose(1, 0, 2).reshape(T, d)
        x = x + o @ W[p + "attn.out.weight"].T + W[p + "attn.out.bias"]
        h = _ln(x, W[p + "mlp_ln.weight"], W[p + "mlp_ln.bias"])
        x = x + _gelu(h @ W[p + "mlp.0.weight"].T + W[p + "mlp.0.bias"]) @ W[p + "mlp.2.weight"].T + W[p + "mlp.2.bias"]
        l += 1
    return _ln(x, W["ln_post.weight"], W["ln_post.bias"])


def apply_logit_rules(logits: np.ndarray, tokens: np.ndarray, sample_begin: int, timestamp_begin: int, no_timestamps: int, eot: int,
                      max_initial_timestamp_index=None, suppress=(), suppress_first=()) -> np.ndarray:
    """The logit filters of openai-whisper's decoding loop [upstream whisper/decoding.py: SuppressBlank, SuppressTokens,
    ApplyTimestampRules, applied in this order by DecodingTask] on logits [B, vocab] given the tokens so far [B, n]
    (n >= sample_begin).  timestamp_begin < 0 switches the timestamp rules off.  Pinned in tests/test_whisper.py against
    transformers' WhisperTimeStampLogitsProcessor, which implements the same timestamp rules."""
    lg = np.array(logits, np.float32, copy=True)
    B, n = tokens.shape
    if n == sample_begin and len(suppress_first):
        lg[:, list(suppress_first)] = -np.inf                                        # SuppressBlank
    if len(suppress):
        lg[:, list(suppress)] = -np.inf                                              # SuppressTokens
    if timestamp_begin < 0:
        return lg
    tb = timestamp_begin
    lg[:, no_timestamps] = -np.inf
    for k in range(B):
        seq = tokens[k, sample_begin:].tolist()
        last = len(seq) >= 1 and seq[-1] >= tb
        penult = len(seq) < 2 or seq[-2] >= tb
        if last:
            if penult:
                lg[k, tb:] = -np.inf                                                  # has to be non-timestamp
            else:
                lg[k, :eot] = -np.inf                                                 # cannot be normal text tokens
        ts = [t for t in seq if t >= tb]
        if ts:                                                                        # timestamps shouldn't decrease; segments have non-zero length
            t_last = ts[-1] if (last and not penult) else ts[-1] + 1
            lg[k, tb:t_last] = -np.inf
    if n == sample_begin:
        lg[:, :tb] = -np.inf                                                          # the first sampled token is a timestamp
        if max_initial_timestamp_index is not None:
            lg[:, tb + max_initial_timestamp_index + 1:] = -np.inf
    for k in range(B):                                                               # timestamp mass above every text token -> timestamp
        row = lg[k].astype(np.float64)
        m = row.max()
        logprobs = row - (m + np.log(np.exp(row - m).sum())) if np.isfinite(m) else row
        t = logprobs[tb:]
        tm = t.max()
        ts_lp = tm + np.log(np.exp(t - tm).sum()) if np.isfinite(tm) else -np.inf
        if ts_lp > logprobs[:tb].max():
            lg[k, :tb] = -np.inf
    return lg


def median_filter(x: np.ndarray, width: int = 7) -> np.ndarray:
    """whisper/timing.py::median_filter [upstream]: median of `width` along the last axis, reflect padding; inputs not longer
    than width // 2 are returned unchanged."""
    pad = width // 2
    if x.shape[-1] <= pad:
        return x
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad, pad)], mode="reflect")
    win = np.lib.stride_tricks.sliding_window_view(xp, width, axis=-1)
    return np.sort(win, axis=-1)[..., pad]


def dtw(cost: np.ndarray):
    """whisper/timing.py::dtw_cpu + backtrace [upstream]: fp32 accumulated cost, ties resolved diagonal < up < left as there.
    cost [N, M] -> (text_indices, time_indices) of the monotone path."""
    N, M = cost.shape
    D = np.full((N + 1, M + 1), np.inf, np.float32)
    T = -np.ones((N + 1, M + 1), np.int8)
    D[0, 0] = 0
    x = cost.astype(np.float32)
    for j in range(1, M + 1):
        for i in range(1, N + 1):
            c0, c1, c2 = D[i - 1, j - 1], D[i - 1, j], D[i, j - 1]
            if c0 < c1 and c0 < c2:
                c, t = c0, 0
            elif c1 < c0 and c1 < c2:
                c, t = c1, 1
            else:
                c, t = c2, 2
            D[i, j] = x[i - 1, j - 1] + c
            T[i, j] = t
    i, j = N, M
    T[0, :] = 2
    T[:, 0] = 1
    ti, tj = [], []
    while i > 0 or j > 0:
        ti.append(i - 1)
        tj.append(j - 1)
        if T[i, j] == 0:
            i, j = i - 1, j - 1
        elif T[i, j] == 1:
            i -= 1
        else:
            j -= 1
    return np.asarray(ti)[::-1], np.asarray(tj)[::-1]


def alignment(weights: np.ndarray, m_valid: int):
    """whisper/timing.py::find_alignment, numeric core [upstream]: weights [A, N, M] (softmax cross-attention rows of the
    alignment heads for the N tokens to align) -> (cost [N, m_valid] float32, start_frame [N]): crop to m_valid audio positions,
    normalise over the tokens (population std), median filter 7, mean over the heads, dtw(-matrix), first audio position of
    every token on the path (jump_times = start_frame / 50 s)."""
    w = np.asarray(weights, np.float32)[..., :m_valid]
    w = (w - w.mean(-2, keepdims=True)) / w.std(-2, keepdims=True)
    cost = -median_filter(w, 7).mean(0)
    ti, tj = dtw(cost)
    jumps = np.pad(np.diff(ti), (1, 0), constant_values=1).astype(bool)
    return cost.astype(np.float32), tj[jumps]
