"""Whisper audio encoder on the B200 (libnsf_b200.so: nsf_whisper_logmel / nsf_whisper_encoder_forward).

The reference transcribes every separated stream with openai-whisper (asr/asr.py:69-74); that package, its weights and
its tokenizer are absent offline.  This module hosts the part of it that is built so far -- the log-mel
front end and the audio encoder, bf16 tensor cores with an fp32 residual stream -- from a state dict in either naming
(openai-whisper ``encoder.blocks.N.attn.query.weight`` ... or transformers ``model.encoder.layers.N.self_attn.q_proj.weight``
...).  The decoder (greedy / beam search, word timestamps) is not built yet: ``notsofar_b200.asr.set_transcriber`` remains the
plug-in point for a complete transcriber.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _cabi

N_FFT, HOP, N_SAMPLES, N_FRAMES, PAD_ROWS = 400, 160, 480000, 3000, 3002


class WhisperDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n_mels", "n_ctx", "d_model", "n_heads", "n_layers", "d_ff")]


def mel_filterbank(n_mels: int, sr: int = 16000, n_fft: int = N_FFT) -> np.ndarray:
    """Slaney-style mel filterbank [n_mels, n_fft/2+1] (librosa.filters.mel(sr, n_fft, n_mels): htk=False, norm='slaney'),
    the matrix openai-whisper ships as assets/mel_filters.npz."""
    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        mel = f / (200.0 / 3)
        log_t = f >= 1000.0
        mel = np.where(log_t, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0), mel)
        return mel

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        f = m * (200.0 / 3)
        log_t = m >= 15.0
        return np.where(log_t, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), f)

    fft_freqs = np.linspace(0, sr / 2, n_fft // 2 + 1)
    mel_pts = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2), n_mels + 2))
    fdiff = np.diff(mel_pts)
    ramps = mel_pts[:, None] - fft_freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0, np.minimum(lower, upper))
    w *= (2.0 / (mel_pts[2:n_mels + 2] - mel_pts[:n_mels]))[:, None]
    return w.astype(np.float32)


def _bf16_bits(a: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).view(torch.int16).numpy()


def _canon(sd: Dict[str, object]) -> Dict[str, np.ndarray]:
    """-> openai-whisper encoder names (conv1.weight, blocks.N.attn.query.weight, ..., ln_post.weight)."""
    out = {}
    hf = {"self_attn.q_proj": "attn.query", "self_attn.k_proj": "attn.key", "self_attn.v_proj": "attn.value",
          "self_attn.out_proj": "attn.out", "self_attn_layer_norm": "attn_ln", "fc1": "mlp.0", "fc2": "mlp.2",
          "final_layer_norm": "mlp_ln"}
    for k, v in sd.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().float().cpu().numpy()
        v = np.asarray(v)
        for pre in ("model.encoder.", "encoder."):
            if k.startswith(pre):
                k = k[len(pre):]
                break
        else:
            if k.startswith(("model.decoder.", "decoder.", "proj_out.")):
                continue
        if k.startswith("layers."):
            _, n, rest = k.split(".", 2)
            for a, b in hf.items():
                if rest.startswith(a + "."):
                    rest = b + rest[len(a):]
                    break
            k = f"blocks.{n}.{rest}"
        elif k.startswith("layer_norm."):
            k = "ln_post." + k[len("layer_norm."):]
        elif k == "embed_positions.weight":
            k = "positional_embedding"
        out[k] = v
    return out


def pack_whisper_encoder(state_dict: Dict[str, object]):
    """-> (WhisperDims, blob float32[...], offsets int64[...], mel filters).  Blob entries (csrc/whisper.cu WhGlobal/WhLayer):
    conv1.W as bf16 head + remainder planes [d][3*n_mels] (K index = tap * n_mels + channel), conv1.b, conv2.W bf16
    [d][3*d] (K index = tap * d + channel), conv2.b, positional embedding [1500][d], ln_post g, b; per layer: ln1 g, b,
    Wqkv bf16 [3d][d] with q and k rows scaled by d_k^-0.25, bqkv (zero for k), Wo bf16, bo, ln2 g, b, W1 bf16 [4d][d], b1,
    W2 bf16 [d][4d], b2."""
    w = _canon(state_dict)
    d, n_mels, _ = w["conv1.weight"].shape
    n_ctx = w["positional_embedding"].shape[0]
    n_layers = 0
    while f"blocks.{n_layers}.attn.query.weight" in w:
        n_layers += 1
    d_ff = w["blocks.0.mlp.0.weight"].shape[0]
    n_heads = d // 64
    chunks, offsets, cursor = [], [], 0

    def add(a32: np.ndarray):
        nonlocal cursor
        a32 = np.ascontiguousarray(a32).reshape(-1)
        assert a32.dtype == np.float32
        offsets.append(cursor)
        chunks.append(a32)
        pad = (-a32.size) % 64
        if pad:
            chunks.append(np.zeros(pad, np.float32))
        cursor += a32.size + pad

    def add_bf16(a: np.ndarray):
        bits = _bf16_bits(a).reshape(-1)
        assert bits.size % 2 == 0
        add(bits.view(np.float32))

    c1 = np.ascontiguousarray(w["conv1.weight"].transpose(0, 2, 1)).reshape(d, 3 * n_mels).astype(np.float32)    # [d][tap][c]
    hi = torch.from_numpy(c1).to(torch.bfloat16)
    lo = (torch.from_numpy(c1) - hi.float()).to(torch.bfloat16)
    add(hi.view(torch.int16).numpy().reshape(-1).view(np.float32)); add(lo.view(torch.int16).numpy().reshape(-1).view(np.float32))
    add(w["conv1.bias"].astype(np.float32))
    add_bf16(np.ascontiguousarray(w["conv2.weight"].transpose(0, 2, 1)).reshape(d, 3 * d))
    add(w["conv2.bias"].astype(np.float32))
    add(w["positional_embedding"].astype(np.float32))
    add(w["ln_post.weight"].astype(np.float32)); add(w["ln_post.bias"].astype(np.float32))
    sc = np.float32(64 ** -0.25)
    for l in range(n_layers):
        p = f"blocks.{l}."
        add(w[p + "attn_ln.weight"].astype(np.float32)); add(w[p + "attn_ln.bias"].astype(np.float32))
        wq, wk, wv = w[p + "attn.query.weight"] * sc, w[p + "attn.key.weight"] * sc, w[p + "attn.value.weight"]
        add_bf16(np.concatenate([wq, wk, wv], 0))
        add(np.concatenate([w[p + "attn.query.bias"] * sc, np.zeros(d, np.float32), w[p + "attn.value.bias"]]).astype(np.float32))
        add_bf16(w[p + "attn.out.weight"]); add(w[p + "attn.out.bias"].astype(np.float32))
        add(w[p + "mlp_ln.weight"].astype(np.float32)); add(w[p + "mlp_ln.bias"].astype(np.float32))
        add_bf16(w[p + "mlp.0.weight"]); add(w[p + "mlp.0.bias"].astype(np.float32))
        add_bf16(w[p + "mlp.2.weight"]); add(w[p + "mlp.2.bias"].astype(np.float32))
    dims = WhisperDims(n_mels=n_mels, n_ctx=n_ctx, d_model=d, n_heads=n_heads, n_layers=n_layers, d_ff=d_ff)
    assert len(offsets) == 8 + 12 * n_layers
    return dims, np.concatenate(chunks), np.asarray(offsets, np.int64), mel_filterbank(n_mels)


class WhisperEncoderB200:
    """Log-mel front end + audio encoder of one Whisper model on one B200."""

    def __init__(self, state_dict: Dict[str, object], device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise _cabi.NsfError("WhisperEncoderB200 needs a CUDA device (sm_100a); there is no CPU path")
        self._lib = _cabi.load()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        dims, blob, offsets, filters = pack_whisper_encoder(state_dict)
        self.dims = dims
        self._blob = torch.from_numpy(blob).to(self.device)
        self._filters = torch.from_numpy(filters).to(self.device)
        self._handle = C.c_void_p()
        offs = (C.c_int64 * len(offsets))(*offsets.tolist())
        _cabi.check(self._lib.nsf_whisper_encoder_create(C.byref(dims), _cabi.ptr(self._blob), self._blob.numel(), offs, len(offsets),
                                                         C.byref(self._handle)), "nsf_whisper_encoder_create")
        self._ws = None

    def __del__(self):
        try:
            self._lib.nsf_whisper_encoder_destroy(self._handle)
        except Exception:
            pass

    def log_mel(self, audio: torch.Tensor):
        """audio [n_batch, 480000] float32 on the device -> (mel_hi, mel_lo) int16 [n_batch, 3002, n_mels] (bf16 head and
        remainder planes, time-major, zero rows in front and behind) and log_spec [n_batch, n_mels, 3000] before the clamp."""
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 2 and audio.shape[1] == N_SAMPLES and audio.is_contiguous()
        B, nm = audio.shape[0], self.dims.n_mels
        log_spec = torch.empty((B, nm, N_FRAMES), dtype=torch.float32, device=audio.device)
        gmax = torch.empty((B,), dtype=torch.int32, device=audio.device)
        hi = torch.empty((B, PAD_ROWS, nm), dtype=torch.int16, device=audio.device)
        lo = torch.empty_like(hi)
        with torch.cuda.device(audio.device):
            _cabi.check(self._lib.nsf_whisper_logmel(_cabi.ptr(audio), B, N_SAMPLES, _cabi.ptr(self._filters), nm, _cabi.ptr(log_spec),
                                                     _cabi.ptr(gmax), _cabi.ptr(hi), _cabi.ptr(lo), _cabi.stream_ptr()), "nsf_whisper_logmel")
        return hi, lo, log_spec

    def encode_mel(self, mel_hi: torch.Tensor, mel_lo: torch.Tensor) -> torch.Tensor:
        """-> [n_batch, 1500, d_model] float32 (the encoder output after ln_post)."""
        B = mel_hi.shape[0]
        need = int(self._lib.nsf_whisper_encoder_workspace_bytes(C.byref(self.dims), B))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty((B, self.dims.n_ctx, self.dims.d_model), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.nsf_whisper_encoder_forward(self._handle, _cabi.ptr(mel_hi), _cabi.ptr(mel_lo), B, _cabi.ptr(out), None,
                                                              _cabi.ptr(self._ws), need, _cabi.stream_ptr()), "nsf_whisper_encoder_forward")
        return out

    def encode_mel_f32(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [n_batch, n_mels, 3000] float32 (whisper's input_features) -> encoder output; splits into bf16 planes on the host side
        of the API (tests / external front ends)."""
        B, nm, _ = mel.shape
        t = torch.zeros((B, PAD_ROWS, nm), dtype=torch.float32, device=mel.device)
        t[:, 1:N_FRAMES + 1] = mel.transpose(1, 2)
        hi = t.to(torch.bfloat16)
        lo = (t - hi.float()).to(torch.bfloat16)
        return self.encode_mel(hi.view(torch.int16).contiguous(), lo.view(torch.int16).contiguous())

    def encode_audio(self, audio: torch.Tensor) -> torch.Tensor:
        hi, lo, _ = self.log_mel(audio)
        return self.encode_mel(hi, lo)

    def log_mel_recording(self, audio: torch.Tensor):
        """The front end of transcribe() [upstream whisper/transcribe.py: log_mel_spectrogram(audio, n_mels, padding=N_SAMPLES)]:
        audio [n] float32 on the device -> (log_spec [n_mels, n_frames] before the clamp, gmax [1] order-encoded recording
        maximum, content_frames = n_frames - 3000).  The 30 s of zero padding are appended here."""
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.dim() == 1
        n = audio.shape[0]
        padded = torch.zeros((n + N_SAMPLES,), dtype=torch.float32, device=audio.device)
        padded[:n] = audio
        n_frames = (n + N_SAMPLES) // HOP
        nm = self.dims.n_mels
        log_spec = torch.empty((nm, n_frames), dtype=torch.float32, device=audio.device)
        gmax = torch.empty((1,), dtype=torch.int32, device=audio.device)
        with torch.cuda.device(audio.device):
            _cabi.check(self._lib.nsf_whisper_logmel_recording(_cabi.ptr(padded), n + N_SAMPLES, _cabi.ptr(self._filters), nm, n_frames,
                                                               _cabi.ptr(log_spec), _cabi.ptr(gmax), _cabi.stream_ptr()), "nsf_whisper_logmel_recording")
        return log_spec, gmax, n_frames - N_FRAMES

    def mel_windows(self, log_spec: torch.Tensor, gmax: torch.Tensor, seeks, sizes=None):
        """30-s windows of a recording's log-mel starting at mel frames ``seeks`` with ``sizes`` content frames each (the rest is
        zero: pad_or_trim), normalised with the recording's maximum -> (mel_hi, mel_lo) int16 [n, 3002, n_mels] for ``encode_mel``."""
        nm, n_frames = log_spec.shape
        B = len(seeks)
        dev = log_spec.device
        sk = torch.tensor([int(v) for v in seeks], dtype=torch.int32, device=dev)
        sz = None if sizes is None else torch.tensor([int(v) for v in sizes], dtype=torch.int32, device=dev)
        hi = torch.empty((B, PAD_ROWS, nm), dtype=torch.int16, device=dev)
        lo = torch.empty_like(hi)
        with torch.cuda.device(dev):
            _cabi.check(self._lib.nsf_whisper_mel_windows(_cabi.ptr(log_spec), n_frames, _cabi.ptr(gmax), nm, _cabi.ptr(sk), _cabi.ptr(sz), B,
                                                          _cabi.ptr(hi), _cabi.ptr(lo), _cabi.stream_ptr()), "nsf_whisper_mel_windows")
        return hi, lo


# ------------------------------------------------------------------------------------------- text decoder (greedy)
class WhisperDecDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("vocab", "n_text_ctx", "d_model", "n_heads", "n_layers", "d_ff", "n_audio_ctx")]


def _canon_decoder(sd: Dict[str, object]) -> Dict[str, np.ndarray]:
    """-> openai-whisper decoder names (token_embedding.weight, positional_embedding, blocks.N.attn.query.weight, ...,
    blocks.N.cross_attn.*, ln.weight) from either naming."""
    hf = {"self_attn.q_proj": "attn.query", "self_attn.k_proj": "attn.key", "self_attn.v_proj": "attn.value",
          "self_attn.out_proj": "attn.out", "self_attn_layer_norm": "attn_ln",
          "encoder_attn.q_proj": "cross_attn.query", "encoder_attn.k_proj": "cross_attn.key", "encoder_attn.v_proj": "cross_attn.value",
          "encoder_attn.out_proj": "cross_attn.out", "encoder_attn_layer_norm": "cross_attn_ln",
          "fc1": "mlp.0", "fc2": "mlp.2", "final_layer_norm": "mlp_ln"}
    out = {}
    for k, v in sd.items():
        for pre in ("model.decoder.", "decoder."):
            if k.startswith(pre):
                k = k[len(pre):]
                break
        else:
            continue
        if isinstance(v, torch.Tensor):
            v = v.detach().float().cpu().numpy()
        v = np.asarray(v)
        if k.startswith("layers."):
            _, n, rest = k.split(".", 2)
            for a, b in hf.items():
                if rest.startswith(a + "."):
                    rest = b + rest[len(a):]
                    break
            k = f"blocks.{n}.{rest}"
        elif k.startswith("layer_norm."):
            k = "ln." + k[len("layer_norm."):]
        elif k == "embed_tokens.weight":
            k = "token_embedding.weight"
        elif k == "embed_positions.weight":
            k = "positional_embedding"
        out[k] = v
    return out


def pack_whisper_decoder(state_dict: Dict[str, object], n_audio_ctx: int = 1500):
    """-> (WhisperDecDims, blob, offsets); entry order: csrc/whisper_dec.cu WdGlobal / WdLayer."""
    w = _canon_decoder(state_dict)
    vocab, d = w["token_embedding.weight"].shape
    n_text_ctx = w["positional_embedding"].shape[0]
    n_layers = 0
    while f"blocks.{n_layers}.attn.query.weight" in w:
        n_layers += 1
    d_ff = w["blocks.0.mlp.0.weight"].shape[0]
    chunks, offsets, cursor = [], [], 0

    def add(a32):
        nonlocal cursor
        a32 = np.ascontiguousarray(a32).reshape(-1)
        assert a32.dtype == np.float32
        offsets.append(cursor)
        chunks.append(a32)
        pad = (-a32.size) % 64
        if pad:
            chunks.append(np.zeros(pad, np.float32))
        cursor += a32.size + pad

    def add_bf16(a):
        bits = _bf16_bits(a).reshape(-1)
        if bits.size % 2:
            bits = np.concatenate([bits, np.zeros(1, bits.dtype)])
        add(bits.view(np.float32))

    f32 = lambda a: np.asarray(a, np.float32)
    add_bf16(w["token_embedding.weight"]); add(f32(w["positional_embedding"])); add(f32(w["ln.weight"])); add(f32(w["ln.bias"]))
    sc = np.float32(64 ** -0.25)
    z = np.zeros(d, np.float32)
    for l in range(n_layers):
        p = f"blocks.{l}."
        add(f32(w[p + "attn_ln.weight"])); add(f32(w[p + "attn_ln.bias"]))
        add_bf16(np.concatenate([w[p + "attn.query.weight"] * sc, w[p + "attn.key.weight"] * sc, w[p + "attn.value.weight"]], 0))
        add(f32(np.concatenate([w[p + "attn.query.bias"] * sc, z, w[p + "attn.value.bias"]])))
        add_bf16(w[p + "attn.out.weight"]); add(f32(w[p + "attn.out.bias"]))
        add(f32(w[p + "cross_attn_ln.weight"])); add(f32(w[p + "cross_attn_ln.bias"]))
        add_bf16(w[p + "cross_attn.query.weight"] * sc); add(f32(w[p + "cross_attn.query.bias"] * sc))
        add_bf16(np.concatenate([w[p + "cross_attn.key.weight"] * sc, w[p + "cross_attn.value.weight"]], 0))
        add(f32(np.concatenate([z, w[p + "cross_attn.value.bias"]])))
        add_bf16(w[p + "cross_attn.out.weight"]); add(f32(w[p + "cross_attn.out.bias"]))
        add(f32(w[p + "mlp_ln.weight"])); add(f32(w[p + "mlp_ln.bias"]))
        add_bf16(w[p + "mlp.0.weight"]); add(f32(w[p + "mlp.0.bias"]))
        add_bf16(w[p + "mlp.2.weight"]); add(f32(w[p + "mlp.2.bias"]))
    dims = WhisperDecDims(vocab=vocab, n_text_ctx=n_text_ctx, d_model=d, n_heads=d // 64, n_layers=n_layers, d_ff=d_ff, n_audio_ctx=n_audio_ctx)
    assert len(offsets) == 4 + 20 * n_layers
    return dims, np.concatenate(chunks), np.asarray(offsets, np.int64)


class _RulesStruct(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("sample_begin", "timestamp_begin", "no_timestamps", "eot", "max_initial_timestamp_index",
                                        "n_suppress", "n_suppress_first")]


class WhisperRules:
    """Logit filters of the decoding loop (SuppressBlank, SuppressTokens, ApplyTimestampRules [upstream whisper/decoding.py]); the
    token ids come from the tokenizer of the checkpoint in use (multilingual large-v3: eot 50257, <|notimestamps|> 50364,
    <|0.00|> 50365).  ``timestamp_begin=None``: no timestamp rules (the without_timestamps mode)."""

    def __init__(self, eot: int, timestamp_begin: Optional[int] = None, no_timestamps: int = -1,
                 max_initial_timestamp_index: Optional[int] = 50, suppress=(), suppress_first=()):
        self.eot, self.timestamp_begin, self.no_timestamps = int(eot), timestamp_begin, int(no_timestamps)
        self.max_initial_timestamp_index = max_initial_timestamp_index
        self.suppress, self.suppress_first = [int(t) for t in suppress], [int(t) for t in suppress_first]

    def to_device(self, sample_begin: int, device):
        st = _RulesStruct(int(sample_begin), -1 if self.timestamp_begin is None else int(self.timestamp_begin), self.no_timestamps, self.eot,
                          -1 if self.max_initial_timestamp_index is None else int(self.max_initial_timestamp_index),
                          len(self.suppress), len(self.suppress_first))
        sup = torch.tensor(self.suppress or [0], dtype=torch.int32, device=device)
        sup1 = torch.tensor(self.suppress_first or [0], dtype=torch.int32, device=device)
        return st, sup, sup1


def apply_logit_rules(logits: torch.Tensor, tokens: torch.Tensor, sample_begin: int, rules: WhisperRules) -> torch.Tensor:
    """logits [B, vocab] f32 cuda, tokens [B, n] int32 (the tokens so far, n >= sample_begin) -> filtered copy (nsf_whisper_logit_rules)."""
    if not (logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2):
        raise _cabi.NsfError("apply_logit_rules needs float32 CUDA logits [B, vocab]; there is no CPU path")
    lib = _cabi.load()
    out = logits.clone().contiguous()
    tok = tokens.to(device=logits.device, dtype=torch.int32).contiguous()
    pos = torch.tensor([tok.shape[1] - 1], dtype=torch.int32, device=logits.device)
    st, sup, sup1 = rules.to_device(sample_begin, logits.device)
    with torch.cuda.device(logits.device):
        _cabi.check(lib.nsf_whisper_logit_rules(_cabi.ptr(out), out.shape[0], out.shape[1], _cabi.ptr(tok), tok.shape[1], _cabi.ptr(pos),
                                                C.byref(st), _cabi.ptr(sup), _cabi.ptr(sup1), _cabi.stream_ptr()), "nsf_whisper_logit_rules")
    return out


def split_segments(tokens, timestamp_begin: int, time_offset: float, segment_size: int, time_precision: float = 0.02,
                   input_stride: int = 2, hop_seconds: float = 0.01):
    """One decoded 30-s window -> (segments, frames to advance the seek by): the segmentation rule of whisper/transcribe.py's
    main loop [upstream, restated; openai-whisper is absent offline, so this host logic is checked by hand-built cases only].
    ``tokens``: the sampled ids of the window (sot sequence and eot stripped); ``segment_size``: mel frames of the window
    (min(3000, frames left)).  Consecutive timestamp pairs <|t1|><|t2|> cut the window into segments [t_start, t_end]; a window
    ending on a single timestamp is consumed whole, otherwise the seek moves to the last closed timestamp and the unfinished
    tail is decoded again; a window without pairs is one segment ending at its last timestamp (or at the window end)."""
    tok = [int(t) for t in tokens]
    is_ts = [t >= timestamp_begin for t in tok]
    single_timestamp_ending = is_ts[-2:] == [False, True]
    consecutive = [i + 1 for i in range(len(tok) - 1) if is_ts[i] and is_ts[i + 1]]
    segments = []
    if consecutive:
        slices = list(consecutive)
        if single_timestamp_ending:
            slices.append(len(tok))
        last = 0
        for cur in slices:
            sl = tok[last:cur]
            segments.append(dict(start=time_offset + (sl[0] - timestamp_begin) * time_precision,
                                 end=time_offset + (sl[-1] - timestamp_begin) * time_precision, tokens=sl))
            last = cur
        if single_timestamp_ending:
            advance = segment_size                         # no speech after the last timestamp
        else:
            advance = (tok[last - 1] - timestamp_begin) * input_stride
    else:
        duration = segment_size * hop_seconds
        ts = [t for t in tok if t >= timestamp_begin]
        if ts and ts[-1] != timestamp_begin:
            duration = (ts[-1] - timestamp_begin) * time_precision
        segments.append(dict(start=time_offset, end=time_offset + duration, tokens=tok))
        advance = segment_size
    return segments, advance


def transcribe_windows(content_frames: int, decode_window, timestamp_begin: int, eot: int, n_frames: int = N_FRAMES,
                       hop_seconds: float = HOP / 16000.0):
    """The seek loop of whisper/transcribe.py [upstream, restated] over one recording of ``content_frames`` mel frames:
    ``decode_window(seek, segment_size)`` returns the sampled token ids of the 30-s window that starts at mel frame ``seek``
    (e.g. ``WhisperB200.decode_greedy`` with ``WhisperRules``); the window is cut into segments by ``split_segments`` and the
    seek advances to the last closed timestamp (a whole window when it ends on a single timestamp or holds no pair).
    Returns the segments (dicts with start / end seconds, tokens, seek).  Windows depend on each other through the seek, so
    the loop is sequential per stream; streams and sessions batch across it."""
    seek, out = 0, []
    while seek < content_frames:
        segment_size = min(n_frames, content_frames - seek)
        tokens = []
        for t in decode_window(seek, segment_size):
            if int(t) == eot:
                break
            tokens.append(int(t))
        if not tokens:
            seek += segment_size
            continue
        segs, advance = split_segments(tokens, timestamp_begin, seek * hop_seconds, segment_size, hop_seconds=hop_seconds)
        for sg in segs:
            if any(t < timestamp_begin for t in sg["tokens"]):            # segments without text carry nothing
                out.append(dict(sg, seek=seek))
        seek += advance if advance > 0 else segment_size                   # a window closed at <|0.00|> must not stall the loop
    return out


def token_alignment(weights: torch.Tensor, m_valid: Optional[int] = None, n_tokens: Optional[torch.Tensor] = None, return_cost: bool = False):
    """weights [B, A, N, M] f32 cuda: cross-attention softmax rows of the A alignment heads for the N token positions to align
    (``decode_greedy(..., align_heads=...)`` captures them) -> start_frame int32 [B, N]: the audio position (20 ms units) at which
    the DTW path enters each token (whisper/timing.py::find_alignment [upstream]; nsf_whisper_alignment)."""
    if not (weights.is_cuda and weights.dtype == torch.float32 and weights.dim() == 4):
        raise _cabi.NsfError("token_alignment needs a float32 CUDA tensor [B, A, N, M]; there is no CPU path")
    lib = _cabi.load()
    w = weights.contiguous()
    B, A, N, M = w.shape
    mv = M if m_valid is None else int(m_valid)
    need = int(lib.nsf_whisper_alignment_workspace_bytes(B, A, N, M))
    ws = torch.empty(need, dtype=torch.uint8, device=w.device)
    out = torch.zeros((B, N), dtype=torch.int32, device=w.device)
    cost = torch.empty((B, N, mv), dtype=torch.float32, device=w.device) if return_cost else None
    nt = None if n_tokens is None else n_tokens.to(device=w.device, dtype=torch.int32).contiguous()
    with torch.cuda.device(w.device):
        _cabi.check(lib.nsf_whisper_alignment(_cabi.ptr(w), B, A, N, M, mv, _cabi.ptr(nt), _cabi.ptr(out), _cabi.ptr(cost), _cabi.ptr(ws), need,
                                              _cabi.stream_ptr()), "nsf_whisper_alignment")
    return (out, cost) if return_cost else out


class WhisperB200:
    """Encoder + greedy decoder of one Whisper model on one B200: audio chunks in, token ids out.

    Plain greedy arg-max decoding (whisper/decoding.py GreedyDecoder with temperature 0) from a caller-given prompt
    (e.g. <|startoftranscript|><|en|><|transcribe|><|notimestamps|>); the logit filters, beam search, temperature
    fallback and word timestamps of openai-whisper's DecodingTask are not built."""

    def __init__(self, state_dict: Dict[str, object], device: Optional[torch.device] = None):
        self.encoder = WhisperEncoderB200(state_dict, device=device)
        self.device = self.encoder.device
        self._lib = self.encoder._lib
        dims, blob, offsets = pack_whisper_decoder(state_dict, self.encoder.dims.n_ctx)
        self.dec_dims = dims
        self._dblob = torch.from_numpy(blob).to(self.device)
        self._dh = C.c_void_p()
        offs = (C.c_int64 * len(offsets))(*offsets.tolist())
        _cabi.check(self._lib.nsf_whisper_decoder_create(C.byref(dims), _cabi.ptr(self._dblob), self._dblob.numel(), offs, len(offsets),
                                                         C.byref(self._dh)), "nsf_whisper_decoder_create")
        self._state = None

    def __del__(self):
        try:
            self._lib.nsf_whisper_decoder_destroy(self._dh)
        except Exception:
            pass

    def encode(self, mel_hi: torch.Tensor, mel_lo: torch.Tensor):
        """-> (encoder output fp32 [B, 1500, d], the same as a bf16 plane int16 [B, 1500, d])"""
        enc = self.encoder
        B = mel_hi.shape[0]
        need = int(self._lib.nsf_whisper_encoder_workspace_bytes(C.byref(enc.dims), B))
        if enc._ws is None or enc._ws.numel() < need:
            enc._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty((B, enc.dims.n_ctx, enc.dims.d_model), dtype=torch.float32, device=self.device)
        out16 = torch.empty((B, enc.dims.n_ctx, enc.dims.d_model), dtype=torch.int16, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.nsf_whisper_encoder_forward(enc._handle, _cabi.ptr(mel_hi), _cabi.ptr(mel_lo), B, _cabi.ptr(out),
                                                              _cabi.ptr(out16), _cabi.ptr(enc._ws), need, _cabi.stream_ptr()),
                        "nsf_whisper_encoder_forward")
        return out, out16

    def log_mel_recording(self, audio: torch.Tensor):
        return self.encoder.log_mel_recording(audio)

    def mel_windows(self, log_spec, gmax, seeks, sizes=None):
        return self.encoder.mel_windows(log_spec, gmax, seeks, sizes)

    # ---- one decoder position at a time with the logits handed back: the engine under beam search / temperature sampling --------
    def begin_sequences(self, enc_bf16: torch.Tensor):
        """Projects the cross-attention keys / values of enc_bf16 int16 [B, 1500, d] and resets the position: ``step_logits`` then
        feeds one token per sequence and position."""
        B = enc_bf16.shape[0]
        need = self._ensure_state(B)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.nsf_whisper_decoder_prefill_cross(self._dh, _cabi.ptr(enc_bf16.contiguous()), B, _cabi.ptr(self._state), need,
                                                                    _cabi.stream_ptr()), "nsf_whisper_decoder_prefill_cross")
        key = (B, self._state.data_ptr())
        if getattr(self, "_fwd", None) is None:
            self._fwd = {}
        if key not in self._fwd:
            bufs = dict(cur=torch.zeros((B,), dtype=torch.int32, device=self.device), pos=torch.zeros((1,), dtype=torch.int32, device=self.device),
                        logits=torch.empty((B, self.dec_dims.vocab), dtype=torch.float32, device=self.device))

            def launch():
                _cabi.check(self._lib.nsf_whisper_decoder_forward(self._dh, _cabi.ptr(bufs["cur"]), _cabi.ptr(bufs["pos"]), B, _cabi.ptr(self._state),
                                                                  need, _cabi.ptr(bufs["logits"]), _cabi.stream_ptr()), "nsf_whisper_decoder_forward")
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                launch()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                launch()
            self._fwd[key] = (bufs, g)
        self._cur_fwd = self._fwd[key]
        self._cur_fwd[0]["pos"].zero_()
        return B

    def step_logits(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens int32 [B] at the current position -> logits [B, vocab] (a view that the next call overwrites); the position
        advances by one."""
        bufs, g = self._cur_fwd
        bufs["cur"].copy_(tokens.to(torch.int32))
        g.replay()
        bufs["pos"] += 1
        return bufs["logits"]

    def reorder_sequences(self, src: torch.Tensor):
        """Sequence b continues the hypothesis of slot src[b] (beam search): permutes the self-attention caches up to the current
        position (rearrange_kv_cache [upstream])."""
        bufs, _ = self._cur_fwd
        B = bufs["cur"].shape[0]
        need = self._ensure_state(B)
        sb = int(self._lib.nsf_whisper_decoder_reorder_scratch_bytes(C.byref(self.dec_dims), B))
        if getattr(self, "_scratch", None) is None or self._scratch.numel() < sb:
            self._scratch = torch.empty(sb, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.nsf_whisper_decoder_reorder(self._dh, _cabi.ptr(src.to(device=self.device, dtype=torch.int32).contiguous()), _cabi.ptr(bufs["pos"]),
                                                              B, _cabi.ptr(self._state), need, _cabi.ptr(self._scratch), sb, _cabi.stream_ptr()),
                        "nsf_whisper_decoder_reorder")

    def _ensure_state(self, B: int):
        need = int(self._lib.nsf_whisper_decoder_state_bytes(C.byref(self.dec_dims), B))
        if self._state is None or self._state.numel() < need:
            self._state = torch.empty(need, dtype=torch.uint8, device=self.device)
        return need

    @torch.no_grad()
    def decode_greedy(self, enc_bf16: torch.Tensor, prompt, max_new_tokens: int = 224, eot: Optional[int] = None,
                      forced_tokens: Optional[torch.Tensor] = None, return_logits: bool = False, rules: Optional["WhisperRules"] = None,
                      align_heads=None):
        """enc_bf16 int16 [B, 1500, d] (from ``encode``).  prompt: list of token ids fed first.  Returns tokens int32
        [B, len(prompt) + n_new] (and the fp32 logits of every step if return_logits).  forced_tokens [B, n]: teacher forcing
        (the arg-max is still computed and returned, the forced token is fed) -- used by the parity tests.  rules: logit filters
        (timestamp rules, suppressed tokens) applied on the device between the logits and the arg-max of every sampled position
        (graph path only).  align_heads: [(layer, head), ...] -- the cross-attention softmax rows of these heads are captured
        for every position; returns (tokens, probs [B, n_heads, total, 1500]) for ``token_alignment`` (word timestamps)."""
        B = enc_bf16.shape[0]
        D = self.dec_dims
        need = self._ensure_state(B)
        lib, st = self._lib, self._state
        with torch.cuda.device(self.device):
            _cabi.check(lib.nsf_whisper_decoder_prefill_cross(self._dh, _cabi.ptr(enc_bf16), B, _cabi.ptr(st), need, _cabi.stream_ptr()),
                        "nsf_whisper_decoder_prefill_cross")
            n_prompt = len(prompt)
            total = min(D.n_text_ctx, n_prompt + max_new_tokens)
            if rules is not None and return_logits:
                raise _cabi.NsfError("decode_greedy: rules are applied inside the graph-replayed step; return_logits is the unfiltered test path")
            if not return_logits:
                return self._decode_graph(B, need, prompt, total, eot, forced_tokens, rules, align_heads)
            if align_heads is not None:
                raise _cabi.NsfError("decode_greedy: align_heads is captured by the graph-replayed step, not with return_logits")
            tokens = torch.zeros((B, total), dtype=torch.int32, device=self.device)
            tokens[:, :n_prompt] = torch.tensor(prompt, dtype=torch.int32, device=self.device)
            nxt = torch.empty((B,), dtype=torch.int32, device=self.device)
            argmaxes = torch.zeros((B, total), dtype=torch.int32, device=self.device)
            logits_all = torch.empty((total, B, D.vocab), dtype=torch.float32, device=self.device) if return_logits else None
            done = torch.zeros((B,), dtype=torch.bool, device=self.device)
            n_done_check = 16
            for pos in range(total - 1):
                cur = tokens[:, pos].contiguous()
                lg = logits_all[pos] if return_logits else None
                _cabi.check(lib.nsf_whisper_decoder_step(self._dh, _cabi.ptr(cur), pos, B, _cabi.ptr(st), need, _cabi.ptr(lg), _cabi.ptr(nxt),
                                                         _cabi.stream_ptr()), "nsf_whisper_decoder_step")
                argmaxes[:, pos + 1] = nxt
                if pos + 1 >= n_prompt:
                    if forced_tokens is not None:
                        tokens[:, pos + 1] = forced_tokens[:, pos + 1 - n_prompt]
                    else:
                        tokens[:, pos + 1] = torch.where(done, torch.full_like(nxt, eot if eot is not None else 0), nxt)
                        if eot is not None:
                            done |= nxt == eot
                            if (pos + 1) % n_done_check == 0 and bool(done.all()):
                                tokens = tokens[:, :pos + 2]
                                argmaxes = argmaxes[:, :pos + 2]
                                break
        if return_logits:
            return tokens, argmaxes, logits_all
        return tokens

    def _decode_graph(self, B: int, need: int, prompt, total: int, eot: Optional[int], forced_tokens: Optional[torch.Tensor],
                      rules: Optional["WhisperRules"] = None, align_heads=None):
        """The decode loop as replays of ONE captured CUDA graph: position, current tokens, done flags and the token record
        live on the device (nsf_whisper_decoder_step_dev), so every step launches the same ~450 kernels with the same
        arguments."""
        dev = self.device
        n_prompt = len(prompt)
        rkey = None if rules is None else (n_prompt, rules.eot, rules.timestamp_begin, rules.no_timestamps, rules.max_initial_timestamp_index,
                                           tuple(rules.suppress), tuple(rules.suppress_first))
        akey = None if align_heads is None else tuple((int(l), int(h)) for l, h in align_heads)
        key = (B, total, self._state.data_ptr(), rkey, akey)
        if getattr(self, "_graphs", None) is None:
            self._graphs = {}
        if key not in self._graphs:
            bufs = dict(cur=torch.zeros((B,), dtype=torch.int32, device=dev), pos=torch.zeros((1,), dtype=torch.int32, device=dev),
                        forced=torch.full((B, total), -1, dtype=torch.int32, device=dev),
                        out=torch.zeros((B, total), dtype=torch.int32, device=dev), arg=torch.zeros((B, total), dtype=torch.int32, device=dev),
                        done=torch.zeros((B,), dtype=torch.uint8, device=dev), eot=torch.zeros((), dtype=torch.int32))

            rdev = rules.to_device(n_prompt, dev) if rules is not None else None
            bufs["rules"] = rdev                                  # keeps the device lists alive as long as the graph

            if akey is not None:
                D = self.dec_dims
                amap = torch.full((D.n_layers, D.n_heads), -1, dtype=torch.int32)
                for slot, (l, h) in enumerate(akey):
                    amap[l, h] = slot
                bufs["amap"] = amap.to(dev)
                bufs["probs"] = torch.zeros((B, len(akey), total, D.n_audio_ctx), dtype=torch.float32, device=dev)

            def launch(eot_val):
                if rdev is not None or akey is not None:
                    _cabi.check(self._lib.nsf_whisper_decoder_step_rules(
                        self._dh, _cabi.ptr(bufs["cur"]), _cabi.ptr(bufs["pos"]), B, _cabi.ptr(self._state), need, _cabi.ptr(bufs["forced"]), total,
                        eot_val, _cabi.ptr(bufs["out"]), _cabi.ptr(bufs["arg"]), _cabi.ptr(bufs["done"]),
                        C.byref(rdev[0]) if rdev is not None else None, _cabi.ptr(rdev[1]) if rdev is not None else None,
                        _cabi.ptr(rdev[2]) if rdev is not None else None, _cabi.ptr(bufs.get("probs")), _cabi.ptr(bufs.get("amap")),
                        len(akey) if akey is not None else 0, _cabi.stream_ptr()), "nsf_whisper_decoder_step_rules")
                    return
                _cabi.check(self._lib.nsf_whisper_decoder_step_dev(
                    self._dh, _cabi.ptr(bufs["cur"]), _cabi.ptr(bufs["pos"]), B, _cabi.ptr(self._state), need, _cabi.ptr(bufs["forced"]), total,
                    eot_val, _cabi.ptr(bufs["out"]), _cabi.ptr(bufs["arg"]), _cabi.ptr(bufs["done"]), _cabi.stream_ptr()),
                    "nsf_whisper_decoder_step_dev")
            self._graphs[key] = (bufs, {}, launch)
        bufs, graphs, launch = self._graphs[key]
        eot_val = -1 if eot is None else int(eot)
        if eot_val not in graphs:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                launch(eot_val)                      # warm-up outside the capture (lazy attribute / table initialisation)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                launch(eot_val)
            graphs[eot_val] = g
        g = graphs[eot_val]
        bufs["forced"].fill_(-1)
        bufs["forced"][:, :n_prompt] = torch.tensor(prompt, dtype=torch.int32, device=dev)
        if forced_tokens is not None:
            n = min(forced_tokens.shape[1], total - n_prompt)
            bufs["forced"][:, n_prompt:n_prompt + n] = forced_tokens[:, :n]
        bufs["out"].zero_(); bufs["arg"].zero_(); bufs["done"].zero_(); bufs["pos"].zero_()
        if "probs" in bufs:
            bufs["probs"].zero_()
        bufs["out"][:, 0] = prompt[0]
        bufs["cur"].fill_(prompt[0])
        steps = total - 1
        for i in range(steps):
            g.replay()
            if eot is not None and forced_tokens is None and (i + 1) % 16 == 0 and i + 1 >= n_prompt and bool(bufs["done"].all()):
                steps = i + 1
                break
        if akey is not None:
            return bufs["out"][:, :steps + 1].clone(), bufs["probs"][:, :, :steps + 1].clone()
        return bufs["out"][:, :steps + 1].clone()
