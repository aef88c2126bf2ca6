"""TitaNet speaker-embedding model on B200 (csrc/titanet.cu).

Host side of ``spk_model.forward(input_signal, input_signal_length)[1]`` as diarization/word_based_diarization.py:105 calls
it: weights under NeMo's state_dict names are folded (BatchNorm into the 1x1 convolutions) and packed once into a device
blob; ``embed`` turns zero-padded crops + lengths into [n, 192] embeddings; ``multiscale_affinity`` is the per-scale
getCosAffinityMatrix + mean of :171-177 on the GPU.  ``as_embedding_backend`` plugs the model into
``notsofar_b200.diarization.set_embedding_backend``.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from .separator import _align, _split16
from .whisper import mel_filterbank

SR, N_FFT, HOP, N_MELS = 16000, 512, 160, 80
TITANET_LARGE = ((1024, 1, 3, False), (1024, 3, 7, True), (1024, 3, 11, True), (1024, 3, 15, True), (3072, 1, 1, False))
BN_EPS_ENC, BN_EPS_DEC = 1e-3, 1e-5


class TitanetDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("feat_in", "n_blocks", "att_ch", "emb")] + \
               [(n, C.c_int * 8) for n in ("filters", "repeat", "kernel", "residual")] + [("precision", C.c_int)]


PRECISION_FP32, PRECISION_FP16 = 0, 1      # nsf_titanet_dims.precision: bf16 head + remainder pairs (fp32-grade) / one fp16 plane


def _np(v) -> np.ndarray:
    if isinstance(v, torch.Tensor):
        v = v.detach().float().cpu().numpy()
    return np.asarray(v, np.float64)


def _bn_fold(w: Dict[str, object], name: str, eps: float) -> Tuple[np.ndarray, np.ndarray]:
    """BatchNorm in eval mode as y = a x + c."""
    g, b, m, v = (_np(w[name + s]) for s in (".weight", ".bias", ".running_mean", ".running_var"))
    a = g / np.sqrt(v + eps)
    return a, b - m * a


def infer_blocks(w: Dict[str, object]) -> Tuple[Tuple[int, int, int, bool], ...]:
    """(filters, repeat, kernel, residual) per Jasper block from the state_dict (encoder.encoder.{b}.mconv.{i}...)."""
    blocks = []
    b = 0
    while f"encoder.encoder.{b}.mconv.0.conv.weight" in w:
        p = f"encoder.encoder.{b}."
        rep, i = 0, 0
        while p + f"mconv.{i}.conv.weight" in w:
            rep += 1
            i += 5
        k = int(_np(w[p + "mconv.0.conv.weight"]).shape[-1])
        f = int(_np(w[p + "mconv.1.conv.weight"]).shape[0])
        blocks.append((f, rep, k, p + "res.0.0.conv.weight" in w))
        b += 1
    return tuple(blocks)


def pack_titanet(w: Dict[str, object], blocks: Sequence[Tuple[int, int, int, bool]], precision: int = PRECISION_FP32):
    """-> (dims, blob float32 [n], offsets int64) in the order csrc/titanet.cu::tn_plan expects.  GEMM weights are stored as
    bf16 head / remainder planes (two 16-bit values per float32 word) with the following BatchNorm folded in; with
    ``precision = PRECISION_FP16`` the head entry is one fp16 plane (saturating) and the remainder entry a placeholder."""
    chunks, offsets, cursor = [], [], 0

    def add(a: np.ndarray):
        nonlocal cursor
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
        offsets.append(cursor)
        chunks.append(a)
        pad = _align(a.size, 64) - a.size
        if pad:
            chunks.append(np.zeros(pad, np.float32))
        cursor += a.size + pad

    def add_split(a: np.ndarray):
        if precision == PRECISION_FP16:
            h = np.clip(np.asarray(a, np.float32), -65504.0, 65504.0).astype(np.float16)
            assert h.size % 2 == 0
            add(h.reshape(-1).view(np.float32))
            add(np.zeros(64, np.float32))
            return
        hi, lo = _split16(np.asarray(a, np.float32), _cabi.SPLIT_BF16)
        add(hi)
        add(lo)

    feat_in = int(_np(w["encoder.encoder.0.mconv.0.conv.weight"]).shape[0])
    c_in = feat_in
    for b, (co, rep, k, res) in enumerate(blocks):
        p = f"encoder.encoder.{b}."
        c, i = c_in, 0
        for r in range(rep):
            add(_np(w[p + f"mconv.{i}.conv.weight"])[:, 0, :].T)                   # depthwise [k][c]
            a, c0 = _bn_fold(w, p + f"mconv.{i + 2}", BN_EPS_ENC)
            add_split(_np(w[p + f"mconv.{i + 1}.conv.weight"])[:, :, 0] * a[:, None])    # pointwise [co][c], BatchNorm folded
            add(c0)
            i += 3 if r == rep - 1 else 5
            c = co
        add(_np(w[p + f"mconv.{i}.fc.0.weight"]))
        add(_np(w[p + f"mconv.{i}.fc.2.weight"]).T)                                # transposed [co / 8][co]
        if res:
            a, c0 = _bn_fold(w, p + "res.0.1", BN_EPS_ENC)
            add_split(_np(w[p + "res.0.0.conv.weight"])[:, :, 0] * a[:, None])
            add(c0)
        c_in = co
    C_ = c_in
    p = "decoder._pooling.attention_layer."
    W1 = _np(w[p + "0.conv_layer.weight"])[:, :, 0]                                   # [att][3C]: x | mean | std
    att = W1.shape[0]
    add_split(W1[:, :C_])
    add(W1[:, C_:])                                                                   # [att][2C] against [mean | std]
    add(_np(w[p + "0.conv_layer.bias"]))
    a1, c1 = _bn_fold(w, p + "0.bn", BN_EPS_DEC)
    add(a1); add(c1)
    add_split(_np(w[p + "2.weight"])[:, :, 0])
    add(_np(w[p + "2.bias"]))
    ae, ce = _bn_fold(w, "decoder.emb_layers.0.0", BN_EPS_DEC)
    We = _np(w["decoder.emb_layers.0.1.weight"])[:, :, 0]
    add(We * ae[None, :])
    add(_np(w["decoder.emb_layers.0.1.bias"]) + We @ ce)
    dims = TitanetDims()
    dims.feat_in, dims.n_blocks, dims.att_ch, dims.emb = feat_in, len(blocks), att, We.shape[0]
    dims.precision = int(precision)
    for b, (co, rep, k, res) in enumerate(blocks):
        dims.filters[b], dims.repeat[b], dims.kernel[b], dims.residual[b] = co, rep, k, int(res)
    return dims, np.concatenate(chunks), np.asarray(offsets, np.int64)


class TitaNetB200:
    """``state_dict``: NeMo EncDecSpeakerLabelModel names (encoder.encoder.*, decoder._pooling.*, decoder.emb_layers.*)."""

    def __init__(self, state_dict: Dict[str, object], device: Optional[torch.device] = None,
                 blocks: Optional[Sequence[Tuple[int, int, int, bool]]] = None, precision: str = "fp16"):
        """``precision``: "fp16" (default) runs the 1x1 convolutions / linear layers the way the reference does under
        ``torch.cuda.amp.autocast()`` (word_based_diarization.py:102-105): fp16 operands, fp32 accumulation, one tensor-core pass;
        "fp32" keeps bf16 head + remainder planes (three passes, fp32-grade: 2e-6 against the fp64 restatement)."""
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _cabi.NsfError("TitaNetB200 needs a CUDA device; there is no CPU path")
        self._lib = _cabi.load()
        self.blocks = tuple(blocks) if blocks is not None else infer_blocks(state_dict)
        if precision not in ("fp16", "fp32"):
            raise _cabi.NsfError(f"TitaNetB200: precision must be 'fp16' or 'fp32', got {precision!r}")
        self.precision = precision
        self.dims, blob, offsets = pack_titanet(state_dict, self.blocks, PRECISION_FP16 if precision == "fp16" else PRECISION_FP32)
        self._blob = torch.from_numpy(blob).to(self.device)
        self._offsets = offsets
        self._filters = torch.from_numpy(np.ascontiguousarray(mel_filterbank(self.dims.feat_in, SR, N_FFT).T)).to(self.device)   # [257][n_mels]
        self._handle = C.c_void_p()
        _cabi.check(self._lib.nsf_titanet_create(C.byref(self.dims), _cabi.ptr(self._blob), self._blob.numel(),
                                                 offsets.ctypes.data_as(C.POINTER(C.c_int64)), len(offsets), C.byref(self._handle)),
                    "nsf_titanet_create")
        self._ws = None

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            self._lib.nsf_titanet_destroy(h)
            self._handle = None

    @property
    def emb_dim(self) -> int:
        return int(self.dims.emb)

    def features(self, crops: torch.Tensor, lengths: torch.Tensor):
        """crops [n, max_len] f32 cuda (zero padded), lengths [n] int32 cuda -> (feat_hi, feat_lo bf16 planes [n, t_pad, 80],
        n_frames [n] int32, t_pad)."""
        if not (crops.is_cuda and crops.dtype == torch.float32 and crops.dim() == 2 and crops.is_contiguous()):
            raise _cabi.NsfError("TitaNetB200.features needs a contiguous float32 CUDA tensor [n, max_len]")
        lengths = lengths.to(device=crops.device, dtype=torch.int32).contiguous()
        n, max_len = crops.shape
        t_pad = _align(max_len // HOP + 1, 16)                                       # pad_to 16 of the NeMo preprocessor
        nm = int(self.dims.feat_in)
        lm = torch.empty((n, t_pad, nm), dtype=torch.float32, device=crops.device)
        hi = torch.empty((n, t_pad, nm), dtype=torch.bfloat16, device=crops.device)
        lo = torch.empty_like(hi)
        nf = torch.empty(n, dtype=torch.int32, device=crops.device)
        with torch.cuda.device(crops.device):
            _cabi.check(self._lib.nsf_titanet_features(_cabi.ptr(crops), _cabi.ptr(lengths), n, max_len, t_pad, _cabi.ptr(self._filters),
                                                       nm, _cabi.ptr(lm), _cabi.ptr(hi), _cabi.ptr(lo), _cabi.ptr(nf), _cabi.stream_ptr()),
                        "nsf_titanet_features")
        return hi, lo, nf, t_pad

    def forward_features(self, hi: torch.Tensor, lo: torch.Tensor, nf: torch.Tensor, t_pad: int) -> torch.Tensor:
        n = hi.shape[0]
        need = int(self._lib.nsf_titanet_workspace_bytes(C.byref(self.dims), n, t_pad))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=hi.device)
        emb = torch.empty((n, self.emb_dim), dtype=torch.float32, device=hi.device)
        with torch.cuda.device(hi.device):
            _cabi.check(self._lib.nsf_titanet_forward(self._handle, _cabi.ptr(hi), _cabi.ptr(lo), _cabi.ptr(nf), n, t_pad, _cabi.ptr(emb),
                                                      _cabi.ptr(self._ws), self._ws.numel(), _cabi.stream_ptr()), "nsf_titanet_forward")
        return emb

    def embed(self, crops: torch.Tensor, lengths: torch.Tensor, bucket: bool = True) -> torch.Tensor:
        """-> [n, emb] float32 embeddings (spk_model.forward(input_signal=crops, input_signal_length=lengths)[1]).

        Padding is masked everywhere, so a crop's embedding does not depend on what it is batched with: with ``bucket`` the
        crops are grouped by padded frame count (the six window scales of a word batch give six groups) and every group runs
        at its own length instead of the longest crop's."""
        if not bucket or crops.shape[0] < 2:
            hi, lo, nf, t_pad = self.features(crops, lengths)
            return self.forward_features(hi, lo, nf, t_pad)
        lens_h = lengths.detach().cpu().numpy().astype(np.int64)
        t_pads = (lens_h // HOP + 1 + 15) // 16 * 16
        out = torch.empty((crops.shape[0], self.emb_dim), dtype=torch.float32, device=crops.device)
        for tp in np.unique(t_pads):
            idx = np.nonzero(t_pads == tp)[0]
            idx_t = torch.from_numpy(idx).to(crops.device)
            L = max(int(lens_h[idx].max()), 1)
            sub = crops.index_select(0, idx_t)[:, :L].contiguous()
            hi, lo, nf, t_pad = self.features(sub, lengths.index_select(0, idx_t))
            out.index_copy_(0, idx_t, self.forward_features(hi, lo, nf, t_pad))
        return out

    def as_embedding_backend(self):
        return lambda crops, lens, cfg: self.embed(crops, lens)


def multiscale_affinity(emb: torch.Tensor) -> torch.Tensor:
    """emb [n_words, n_scales, D] f32 cuda -> mean over the scales of getCosAffinityMatrix(emb[:, scale])
    (word_based_diarization.py:171-177), [n_words, n_words] f32."""
    if not (emb.is_cuda and emb.dtype == torch.float32 and emb.dim() == 3):
        raise _cabi.NsfError("multiscale_affinity needs a float32 CUDA tensor [n_words, n_scales, D]; there is no CPU path")
    emb = emb.contiguous()
    n, ns, d = emb.shape
    if n == 1:
        return torch.ones((1, 1), dtype=torch.float32, device=emb.device)
    lib = _cabi.load()
    acc = torch.zeros((n, n), dtype=torch.float32, device=emb.device)
    en = torch.empty((n, d), dtype=torch.float32, device=emb.device)
    sim = torch.empty((n, n), dtype=torch.float32, device=emb.device)
    mm = torch.empty(2, dtype=torch.int32, device=emb.device)
    with torch.cuda.device(emb.device):
        for s in range(ns):
            _cabi.check(lib.nsf_cos_affinity_accum(C.c_void_p(emb.data_ptr() + 4 * s * d), ns * d, d, n, 1.0 / ns, _cabi.ptr(en),
                                                   _cabi.ptr(sim), _cabi.ptr(mm), _cabi.ptr(acc), _cabi.stream_ptr()),
                        "nsf_cos_affinity_accum")
    return acc


def load_titanet_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """State dict of a NeMo speaker model from a ``.nemo`` archive (a tar file holding ``model_weights.ckpt``) or from a file
    written by ``torch.save(model.state_dict())`` -- what EncDecSpeakerLabelModel.from_pretrained would have downloaded
    (word_based_diarization.py:26)."""
    import io
    import tarfile
    if tarfile.is_tarfile(path):
        with tarfile.open(path, "r:*") as tar:
            names = [m for m in tar.getmembers() if m.name.endswith(".ckpt")]
            if not names:
                raise _cabi.NsfError(f"{path}: no *.ckpt member in the archive")
            sd = torch.load(io.BytesIO(tar.extractfile(names[0]).read()), map_location="cpu", weights_only=True)
    else:
        sd = torch.load(path, map_location="cpu", weights_only=True)
    sd = sd.get("state_dict", sd)
    return {k: v for k, v in sd.items() if k.startswith(("encoder.", "decoder."))}


def load_titanet(path: str, device) -> TitaNetB200:
    return TitaNetB200(load_titanet_state_dict(path), device)
