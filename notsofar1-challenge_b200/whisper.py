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
import math
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
            _cabi.check(self._lib.nsf_whisper_encoder_forward(self._handle, _cabi.ptr(mel_hi), _cabi.ptr(mel_lo), B, _cabi.ptr(out),
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
