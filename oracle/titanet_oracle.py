"""CPU restatement of the TitaNet speaker-embedding forward  --  TEST INFRASTRUCTURE (numpy).

Only tests/ may import this module; it is the checker for csrc/titanet.cu, never the thing shipped.

The algorithm lives in a third-party dependency that is absent from /root/reference: NeMo (``nemo_toolkit[all]``, unpinned,
requirements.txt:18), model ``titanet_large`` (configs/inference/*.yaml: embedding_model_name), called from
diarization/word_based_diarization.py:26 (EncDecSpeakerLabelModel.from_pretrained) and :105 (spk_model.forward(input_signal,
input_signal_length) -> (logits, embeddings)).  What is restated here is its published architecture [upstream, from the
titanet-large.yaml recipe and the modules it instantiates]:
  AudioToMelSpectrogramPreprocessor / FilterbankFeatures   pre-emphasis 0.97, torch.stft(n_fft 512, hop 160, win 400 symmetric
        hann, centred, reflect), power spectrum, 80 slaney mel bands (0..8 kHz), log(. + 2^-24), per-feature mean / unbiased-std
        normalisation over the valid frames (+1e-5), frames beyond the length zeroed, time padded to a multiple of 16
  ConvASREncoder (JasperBlock x5, conv_mask)                separable 1-D convolutions (depthwise k, pointwise 1x1, BatchNorm
        eps 1e-3) x repeat with ReLU in between, squeeze-excite (global masked mean, 1/8 bottleneck, no biases) after the last
        BatchNorm, residual 1x1 conv + BatchNorm branch on blocks 1-3, ReLU; every convolution sees its input masked beyond
        the sequence length.  filters / repeat / kernel: 1024/1/3, 1024/3/7, 1024/3/11, 1024/3/15, 3072/1/1
  SpeakerDecoder (pool_mode attention, emb_sizes 192)       attentive statistics pooling with global context (TDNN 9216->128
        1x1 conv, ReLU, BatchNorm; tanh; 1x1 conv 128->3072; masked softmax over time; weighted mean and std, clamp 1e-10),
        BatchNorm(6144) + 1x1 conv -> 192-d embedding
**Parity unpinned** with respect to NeMo itself: NeMo and the checkpoint are absent offline and the reference holds no test or
golden vector for this path.  The pins available offline (tests/test_titanet.py): the front end against the
torch.stft / torch.hann_window / torch.std calls FilterbankFeatures is made of (1e-9), the network against the same recipe
assembled from torch.nn Conv1d / BatchNorm1d modules loading these weights by name (1e-10), and the published parameter count.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

SR, N_FFT, WIN, HOP, N_MELS = 16000, 512, 400, 160, 80
BLOCKS = [  # (filters, repeat, kernel, residual)
    (1024, 1, 3, False), (1024, 3, 7, True), (1024, 3, 11, True), (1024, 3, 15, True), (3072, 1, 1, False)]
ATT_CH, EMB = 128, 192
BN_EPS_ENC, BN_EPS_DEC = 1e-3, 1e-5


def block_plan(blocks=BLOCKS, feat_in=N_MELS) -> List[Tuple[int, int, int, int, bool]]:
    """-> [(c_in, c_out, repeat, kernel, residual)] per block."""
    out, c = [], feat_in
    for f, r, k, res in blocks:
        out.append((c, f, r, k, res))
        c = f
    return out


def random_weights(seed: int = 0, blocks=BLOCKS, feat_in=N_MELS, att_ch=ATT_CH, emb=EMB) -> Dict[str, np.ndarray]:
    """Random weights under NeMo's state_dict names (encoder.encoder.{b}.mconv.{i}.conv.weight ...), with non-trivial
    BatchNorm statistics so that the folding is exercised."""
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}

    def bn(name, c):
        w[name + ".weight"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
        w[name + ".bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        w[name + ".running_mean"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        w[name + ".running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)

    for b, (ci, co, rep, k, res) in enumerate(block_plan(blocks, feat_in)):
        p = f"encoder.encoder.{b}."
        c, i = ci, 0
        for r in range(rep):
            w[p + f"mconv.{i}.conv.weight"] = (rng.standard_normal((c, 1, k)) / np.sqrt(k)).astype(np.float32)       # depthwise
            w[p + f"mconv.{i + 1}.conv.weight"] = (rng.standard_normal((co, c, 1)) / np.sqrt(c)).astype(np.float32)  # pointwise
            bn(p + f"mconv.{i + 2}", co)
            i += 3 if r == rep - 1 else 5            # ReLU + dropout between repeats
            c = co
        w[p + f"mconv.{i}.fc.0.weight"] = (rng.standard_normal((co // 8, co)) / np.sqrt(co)).astype(np.float32)
        w[p + f"mconv.{i}.fc.2.weight"] = (rng.standard_normal((co, co // 8)) / np.sqrt(co // 8)).astype(np.float32)
        if res:
            w[p + "res.0.0.conv.weight"] = (rng.standard_normal((co, ci, 1)) / np.sqrt(ci)).astype(np.float32)
            bn(p + "res.0.1", co)
    c = block_plan(blocks, feat_in)[-1][1]
    p = "decoder._pooling.attention_layer."
    w[p + "0.conv_layer.weight"] = (rng.standard_normal((att_ch, 3 * c, 1)) / np.sqrt(3 * c)).astype(np.float32)
    w[p + "0.conv_layer.bias"] = (0.1 * rng.standard_normal(att_ch)).astype(np.float32)
    bn(p + "0.bn", att_ch)
    w[p + "2.weight"] = (rng.standard_normal((c, att_ch, 1)) / np.sqrt(att_ch)).astype(np.float32)
    w[p + "2.bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
    bn("decoder.emb_layers.0.0", 2 * c)
    w["decoder.emb_layers.0.1.weight"] = (rng.standard_normal((emb, 2 * c, 1)) / np.sqrt(2 * c)).astype(np.float32)
    w["decoder.emb_layers.0.1.bias"] = (0.1 * rng.standard_normal(emb)).astype(np.float32)
    return w


def mel_filterbank(n_mels=N_MELS, sr=SR, n_fft=N_FFT) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin 0, fmax sr/2): htk=False, norm='slaney' -> [n_mels, n_fft/2+1] float64."""
    def hz_to_mel(f):
        f = np.asarray(f, np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0), f / (200.0 / 3))

    def mel_to_hz(m):
        m = np.asarray(m, np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), m * (200.0 / 3))

    freqs = np.linspace(0, sr / 2, n_fft // 2 + 1)
    pts = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2), n_mels + 2))
    ramps = pts[:, None] - freqs[None, :]
    lower, upper = -ramps[:-2] / np.diff(pts)[:-1, None], ramps[2:] / np.diff(pts)[1:, None]
    fb = np.maximum(0, np.minimum(lower, upper))
    return fb * (2.0 / (pts[2:] - pts[:-2]))[:, None]


def seq_len(n_samples: int) -> int:
    """FilterbankFeatures.get_seq_len for centred frames: floor(n / hop) + 1."""
    return n_samples // HOP + 1


def features(audio: np.ndarray, dtype=np.float64) -> np.ndarray:
    """One crop [n] -> normalised log-mel [seq_len(n), 80] (time major; the zero frames NeMo pads to a multiple of 16 are
    left to the caller)."""
    x = np.asarray(audio, dtype)
    n = len(x)
    x = np.concatenate([x[:1], x[1:] - 0.97 * x[:-1]])                           # pre-emphasis
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    win = np.zeros(N_FFT, dtype)
    off = (N_FFT - WIN) // 2
    win[off:off + WIN] = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(WIN) / (WIN - 1))    # hann_window(400, periodic=False), centred in n_fft
    T = seq_len(n)
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(T)[:, None]
    power = np.abs(np.fft.rfft(xp[idx] * win, axis=1)) ** 2                        # [T, 257]
    mel = power @ mel_filterbank().T.astype(dtype)
    lm = np.log(mel + 2.0 ** -24)
    mean = lm.mean(0)
    std = np.sqrt(((lm - mean) ** 2).sum(0) / (T - 1)) + 1e-5
    return ((lm - mean) / std).astype(dtype)


def _bn(w, name, eps):
    g, b, m, v = (np.asarray(w[name + s], np.float64) for s in (".weight", ".bias", ".running_mean", ".running_var"))
    a = g / np.sqrt(v + eps)
    return a, b - m * a


def encoder(w: Dict[str, np.ndarray], feats: List[np.ndarray], blocks=BLOCKS) -> List[np.ndarray]:
    """feats: per crop [T_i, c_in] (valid frames only; masking beyond the length == zero padding of the valid part).
    -> per crop [T_i, c_last]."""
    outs = []
    for x in feats:
        x = np.asarray(x, np.float64)
        for b, (ci, co, rep, k, res) in enumerate(block_plan(blocks, x.shape[1])):
            p = f"encoder.encoder.{b}."
            x_in, i = x, 0
            for r in range(rep):
                dw = np.asarray(w[p + f"mconv.{i}.conv.weight"], np.float64)[:, 0, :]          # [c, k]
                pw = np.asarray(w[p + f"mconv.{i + 1}.conv.weight"], np.float64)[:, :, 0]      # [co, c]
                a, c0 = _bn(w, p + f"mconv.{i + 2}", BN_EPS_ENC)
                T = x.shape[0]
                xp = np.pad(x, ((k // 2, k // 2), (0, 0)))
                y = sum(xp[j:j + T] * dw[:, j] for j in range(k))                               # cross-correlation, 'same' padding
                x = (y @ pw.T) * a + c0
                last = r == rep - 1
                if not last:
                    x = np.maximum(x, 0.0)
                i += 3 if last else 5
            f0 = np.asarray(w[p + f"mconv.{i}.fc.0.weight"], np.float64)
            f2 = np.asarray(w[p + f"mconv.{i}.fc.2.weight"], np.float64)
            gate = 1.0 / (1.0 + np.exp(-(np.maximum(x.mean(0) @ f0.T, 0.0) @ f2.T)))           # squeeze-excite over the valid frames
            x = x * gate
            if res:
                a, c0 = _bn(w, p + "res.0.1", BN_EPS_ENC)
                x = x + (x_in @ np.asarray(w[p + "res.0.0.conv.weight"], np.float64)[:, :, 0].T) * a + c0
            x = np.maximum(x, 0.0)
        outs.append(x)
    return outs


def decoder(w: Dict[str, np.ndarray], enc: List[np.ndarray]) -> np.ndarray:
    """Attentive statistics pooling + embedding layer -> [n, 192]."""
    p = "decoder._pooling.attention_layer."
    W1 = np.asarray(w[p + "0.conv_layer.weight"], np.float64)[:, :, 0]
    b1 = np.asarray(w[p + "0.conv_layer.bias"], np.float64)
    a1, c1 = _bn(w, p + "0.bn", BN_EPS_DEC)
    W2 = np.asarray(w[p + "2.weight"], np.float64)[:, :, 0]
    b2 = np.asarray(w[p + "2.bias"], np.float64)
    ae, ce = _bn(w, "decoder.emb_layers.0.0", BN_EPS_DEC)
    We = np.asarray(w["decoder.emb_layers.0.1.weight"], np.float64)[:, :, 0]
    be = np.asarray(w["decoder.emb_layers.0.1.bias"], np.float64)
    embs = []
    for x in enc:
        T, c = x.shape
        mean = x.mean(0)
        std = np.sqrt(np.maximum(((x - mean) ** 2).mean(0), 1e-10))
        ctx = np.concatenate([x, np.broadcast_to(mean, (T, c)), np.broadcast_to(std, (T, c))], axis=1)
        h = np.tanh(np.maximum(ctx @ W1.T + b1, 0.0) * a1 + c1)
        logit = h @ W2.T + b2
        logit = logit - logit.max(0)
        alpha = np.exp(logit)
        alpha /= alpha.sum(0)
        mu = (alpha * x).sum(0)
        sg = np.sqrt(np.maximum((alpha * (x - mu) ** 2).sum(0), 1e-10))
        pool = np.concatenate([mu, sg]) * ae + ce
        embs.append(pool @ We.T + be)
    return np.stack(embs)


def embed(w: Dict[str, np.ndarray], crops: List[np.ndarray], blocks=BLOCKS) -> np.ndarray:
    """spk_model.forward(input_signal, input_signal_length)[1] for a list of crops -> [n, 192] float64."""
    return decoder(w, encoder(w, [features(c) for c in crops], blocks))


def cos_affinity(emb: np.ndarray) -> np.ndarray:
    """getCosAffinityMatrix [upstream offline_clustering.py]: cosine similarity (norm + 3.5e-4), unit diagonal, global min-max
    scaling.  emb [n, d] -> [n, n]."""
    e = np.asarray(emb, np.float64)
    if e.shape[0] == 1:
        return np.ones((1, 1))
    en = e / (np.linalg.norm(e, axis=1, keepdims=True) + 3.5e-4)
    sim = en @ en.T
    np.fill_diagonal(sim, 1.0)
    return (sim - sim.min()) / (sim.max() - sim.min())
