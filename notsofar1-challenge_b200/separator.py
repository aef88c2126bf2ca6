"""B200 separator: the reference's segment-wise CSS model (ConformerCssWrapper,
css/training/conformer_wrapper.py:51-146) re-hosted on libnsf_b200.so.

It keeps the reference's separator contract -- ``.stft(s)``, ``.separate(stft)``, ``.istft(stft)``,
``.eval()``, ``.cpu()``, ``.to(device)``, ``.training`` (README.md:229-232) -- so code written against
the reference's plug-in interface keeps working, and adds the batched entry points
(``features`` / ``masks`` / ``mvdr``) that ``css.separate_and_stitch`` drives.

Weights come from a reference ``state_dict`` (677 tensors for the v1.0 MC model, names
``executor.nnet...``, optional DDP ``module.`` prefix as in css/helpers.py:30-36) and are repacked
once into a single device blob: GEMM weights split into head + remainder in the engine's format (two fp32
arrays for the TF32 engines, two 16-bit arrays for 2xBF16 / 2xF16), K padded to a multiple of 32, BatchNorm
folded into scale/shift.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _cabi

NUM_BINS = 257
FRAME_LEN = 512
FRAME_HOP = 256
_P = "executor.nnet."


def _align(n: int, a: int) -> int:
    return (n + a - 1) // a * a


def _split_tf32(w: np.ndarray):
    """fp32 -> (TF32-representable head with the 13 low mantissa bits cleared, exact remainder)."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    hi = (w.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = w - hi
    return hi, lo


def _split16(w: np.ndarray, fmt: int):
    """fp32 -> (head, remainder) as 16-bit patterns packed two per float32 word (csrc/common.cuh SplitFmt):
    SPLIT_BF16: hi = bf16(w), lo = bf16(w - hi); SPLIT_F16: the same in fp16 on 2^8 w (saturating at 65504)."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    if fmt == _cabi.SPLIT_BF16:
        t = torch.from_numpy(w)
        hi = t.to(torch.bfloat16)
        lo = (t - hi.to(torch.float32)).to(torch.bfloat16)
        hi, lo = hi.view(torch.int16).numpy(), lo.view(torch.int16).numpy()
    else:
        ws = np.clip(w * np.float32(_cabi.F16_WEIGHT_SCALE), -65504.0, 65504.0).astype(np.float32)
        hi = ws.astype(np.float16)
        lo = (ws - hi.astype(np.float32)).astype(np.float16)
        hi, lo = hi.view(np.int16), lo.view(np.int16)
    assert hi.size % 2 == 0
    return hi.reshape(-1).view(np.float32), lo.reshape(-1).view(np.float32)


def split_activations(a: np.ndarray, gemm_engine: int):
    """Host-side twin of the device's split store (csrc/common.cuh split_store): fp32 activations [rows, ld] ->
    (head, remainder) in the layout nsf_conformer_forward expects for ``gemm_engine`` (tests, external feature paths)."""
    fmt = _cabi.split_fmt_of_engine(gemm_engine)
    a = np.ascontiguousarray(a, dtype=np.float32)
    if fmt == _cabi.SPLIT_TF32:
        return _split_tf32(a)
    if fmt == _cabi.SPLIT_BF16:
        t = torch.from_numpy(a)
        hi = t.to(torch.bfloat16)
        lo = (t - hi.to(torch.float32)).to(torch.bfloat16)
        return hi.view(torch.int16).numpy(), lo.view(torch.int16).numpy()
    s = np.clip(a * np.float32(_cabi.F16_ACT_SCALE), -65504.0, 65504.0).astype(np.float32)
    hi = s.astype(np.float16)
    lo = (s - hi.astype(np.float32)).astype(np.float16)
    return hi.view(np.int16), lo.view(np.int16)


def _strip_prefix(sd: Dict[str, object]) -> Dict[str, np.ndarray]:
    out = {}
    for k, v in sd.items():
        if k.startswith("module."):
            k = k[len("module."):]
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def pack_weights(state_dict: Dict[str, object], T: int, gemm_engine: int):
    """Returns (dims: _cabi.ConformerDims, blob: np.float32[...], offsets: np.int64[...]).

    Blob order (offsets index): 12 global entries
      0 embed.W hi [d][Kf] | 1 embed.W lo | 2 embed.b | 3 embed LN g | 4 embed LN b | 5 pe_k hi [2*maxlen][d_k] | 6 pe_k lo (TF32 pairs)
      7 head.W hi [n_out][d] | 8 head.W lo | 9 head.b | 10 pe_k hi | 11 pe_k lo (bf16 pairs)
    then 34 per encoder block
      0-7   feed_forward_in : LN g, LN b, W1 hi, W1 lo, b1, W2 hi, W2 lo, b2
      8-15  self_attn       : LN g, LN b, Wqkv hi [3d][d], Wqkv lo, bqkv, Wo hi, Wo lo, bo
      16-21 conv            : LN g, LN b, scalars[8] = (w1a, b1a, w1g, b1g, w2, b2, 0, 0), dw W [d][ks], BN scale, BN shift
      22-29 feed_forward_out: as feed_forward_in
      30-31 layer_norm      : g, b
      32-33 column sums of the stored Wqkv / feed_forward_out W1 (folded LayerNorms, see below; zeros otherwise)

    When the library folds the attention and feed_forward_out LayerNorms into the following GEMM (nsf_conformer_ln_fold:
    2xBF16 engine, d_model = 128 * {1, 2, 4}), LN(x) W^T + b = rstd (x (gamma W)^T - mean colsum(gamma W)) + (b + W beta):
    those two weights are stored gamma-scaled, their biases as b + W beta (float64 sums), and entries 32-33 hold the column
    sums of the weights *as stored* (head + remainder), so that the mean term cancels what the tensor cores accumulate.
    """
    w = _strip_prefix(state_dict)
    d_model, in_features = w[_P + "conformer.embed.0.weight"].shape
    two_maxlen, d_k = w[_P + "conformer.pos_emb.pe_k.weight"].shape
    d_ff = w[_P + "conformer.encoders.0.feed_forward_in.net.0.weight"].shape[0]
    ks = w[_P + "conformer.encoders.0.conv.dw_conv_1d.weight"].shape[2]
    n_out = w[_P + "linear.weight"].shape[0]
    n_blocks = 0
    while (_P + f"conformer.encoders.{n_blocks}.layer_norm.weight") in w:
        n_blocks += 1
    n_heads = d_model // d_k
    Kf = _align(in_features, 32)

    chunks, offsets = [], []
    cursor = 0

    def add(a: np.ndarray):
        nonlocal cursor
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
        offsets.append(cursor)
        chunks.append(a)
        pad = _align(a.size, 64) - a.size
        if pad:
            chunks.append(np.zeros(pad, np.float32))
        cursor += a.size + pad

    fmt = _cabi.split_fmt_of_engine(gemm_engine)
    dims = _cabi.ConformerDims(d_model=d_model, n_heads=n_heads, d_ff=d_ff, n_blocks=n_blocks, kernel_size=ks,
                               in_features=in_features, n_out=n_out, maxlen=two_maxlen // 2, T=T, gemm_engine=gemm_engine)
    ln_fold = bool(_cabi.load().nsf_conformer_ln_fold(C.byref(dims)))
    csums = []

    def add_split(a: np.ndarray, tf32: bool = False):
        """GEMM weight in the engine's split format (pe_k feeds the attention kernel, which stays SPLIT_TF32)."""
        hi, lo = _split_tf32(a) if (tf32 or fmt == _cabi.SPLIT_TF32) else _split16(a, fmt)
        add(hi)
        add(lo)
        return hi, lo

    def add_linear_after_ln(W: np.ndarray, b: np.ndarray, g: np.ndarray, beta: np.ndarray):
        """W [N, K], b [N] of a Linear that follows LayerNorm(gamma=g, beta); returns the column sums entry."""
        if not ln_fold:
            add_split(W); add(b)
            return np.zeros(W.shape[0], np.float32)
        Wg = (W.astype(np.float32) * g.astype(np.float32)[None, :]).astype(np.float32)
        hi, lo = add_split(Wg)
        add((b.astype(np.float64) + W.astype(np.float64) @ beta.astype(np.float64)).astype(np.float32))
        assert fmt == _cabi.SPLIT_BF16
        stored = (torch.from_numpy(hi.view(np.int16).copy()).view(torch.bfloat16).to(torch.float64)
                  + torch.from_numpy(lo.view(np.int16).copy()).view(torch.bfloat16).to(torch.float64)).reshape(W.shape)
        return stored.sum(dim=1).to(torch.float32).numpy()

    def add_ffn(q: str, after_ln: bool = False):
        add(w[q + "layer_norm.weight"]); add(w[q + "layer_norm.bias"])
        if after_ln:
            csums.append(add_linear_after_ln(w[q + "net.0.weight"], w[q + "net.0.bias"], w[q + "layer_norm.weight"], w[q + "layer_norm.bias"]))
        else:
            add_split(w[q + "net.0.weight"]); add(w[q + "net.0.bias"])
        add_split(w[q + "net.3.weight"]); add(w[q + "net.3.bias"])

    c = _P + "conformer."
    emb = np.zeros((d_model, Kf), np.float32)
    emb[:, :in_features] = w[c + "embed.0.weight"]
    add_split(emb)
    add(w[c + "embed.0.bias"])
    add(w[c + "embed.1.weight"])
    add(w[c + "embed.1.bias"])
    # pe_k feeds the attention kernels: TF32 pairs here (attention.cu and the unfused path), bf16 pairs at the end
    # (attention16.cu); which kernel runs depends on the engine and on the segment length
    pe = w[c + "pos_emb.pe_k.weight"]
    add_split(pe, tf32=True)
    add_split(w[_P + "linear.weight"])
    add(w[_P + "linear.bias"])
    pe16_hi, pe16_lo = _split16(pe, _cabi.SPLIT_BF16)
    add(pe16_hi); add(pe16_lo)
    for l in range(n_blocks):
        p = c + f"encoders.{l}."
        add_ffn(p + "feed_forward_in.")
        a = p + "self_attn."
        add(w[a + "layer_norm.weight"]); add(w[a + "layer_norm.bias"])
        csums.clear()
        csums.append(add_linear_after_ln(
            np.concatenate([w[a + "linear_q.weight"], w[a + "linear_k.weight"], w[a + "linear_v.weight"]], axis=0),
            np.concatenate([w[a + "linear_q.bias"], w[a + "linear_k.bias"], w[a + "linear_v.bias"]]),
            w[a + "layer_norm.weight"], w[a + "layer_norm.bias"]))
        add_split(w[a + "linear_out.weight"]); add(w[a + "linear_out.bias"])
        cv = p + "conv."
        add(w[cv + "layer_norm.weight"]); add(w[cv + "layer_norm.bias"])
        w1 = w[cv + "pw_conv_1.weight"].reshape(2)
        b1 = w[cv + "pw_conv_1.bias"].reshape(2)
        add(np.array([w1[0], b1[0], w1[1], b1[1], w[cv + "pw_conv_2.weight"].reshape(()), w[cv + "pw_conv_2.bias"].reshape(()),
                      0.0, 0.0], np.float32))
        add(w[cv + "dw_conv_1d.weight"].reshape(d_model, ks))
        # BatchNorm1d in eval mode folded with the depthwise-conv bias (conformer.py:119-121)
        scale = w[cv + "BN.weight"].astype(np.float64) / np.sqrt(w[cv + "BN.running_var"].astype(np.float64) + 1e-5)
        shift = (w[cv + "dw_conv_1d.bias"].astype(np.float64) - w[cv + "BN.running_mean"].astype(np.float64)) * scale \
            + w[cv + "BN.bias"].astype(np.float64)
        add(scale.astype(np.float32)); add(shift.astype(np.float32))
        add_ffn(p + "feed_forward_out.", after_ln=True)
        add(w[p + "layer_norm.weight"]); add(w[p + "layer_norm.bias"])
        add(csums[0]); add(csums[1])

    assert len(offsets) == 12 + 34 * n_blocks, len(offsets)
    return dims, np.concatenate(chunks), np.asarray(offsets, dtype=np.int64), \
        dict(input_bias=w[_P + "input_bias"].reshape(-1).astype(np.float32),
             input_scale=w[_P + "input_scale"].reshape(-1).astype(np.float32))


class ConformerCssB200:
    """Segment-wise CSS model on one B200.  Not an nn.Module: all arithmetic is in libnsf_b200.so."""

    def __init__(self, state_dict: Dict[str, object], num_spks: int = 3, device: Optional[torch.device] = None,
                 gemm_engine: int = _cabi.GEMM_TC_2XBF16, segments_per_batch: int = 1280):
        self._lib = _cabi.load()
        self._sd = _strip_prefix(state_dict)
        self.training = False
        self.num_spks = num_spks
        self.gemm_engine = gemm_engine
        self.segments_per_batch = segments_per_batch
        self.device = torch.device(device) if device is not None else None
        n_out = self._sd[_P + "linear.weight"].shape[0]
        self.num_masks = n_out // NUM_BINS
        self.num_nois = self.num_masks - num_spks
        self.in_features = self._sd[_P + "conformer.embed.0.weight"].shape[1]
        self.num_mics = self.in_features // NUM_BINS
        self.ldf = _align(self.in_features, 32)
        self._handles = {}          # T -> (handle, dims)
        self._blob = None
        self._ws = None
        self._feat = None
        if self.device is not None and self.device.type == "cuda":
            self._upload()

    # ---- reference-compatible module surface ------------------------------------------------
    def eval(self):
        self.training = False
        return self

    def cpu(self):
        """The reference moves the model to the CPU for the long-form STFT (css.py:141); this path
        never computes on the host, so the call only drops the device copy of the weights."""
        return self

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            return self
        if self.device != device or self._blob is None:
            self.device = device
            self._handles.clear()
            self._upload()
        return self

    def state_dict(self):
        return dict(self._sd)

    # ---- plumbing ---------------------------------------------------------------------------
    def _require_cuda(self):
        if self.device is None or self.device.type != "cuda" or self._blob is None:
            if not torch.cuda.is_available():
                raise _cabi.NsfError("ConformerCssB200 needs a CUDA device (sm_100a); there is no CPU path")
            self.to(torch.device("cuda", torch.cuda.current_device()))

    def _upload(self):
        dims, blob, offsets, extra = pack_weights(self._sd, T=186, gemm_engine=self.gemm_engine)
        self._offsets = offsets
        self._blob = torch.from_numpy(blob).to(self.device)
        self._in_bias = torch.from_numpy(extra["input_bias"]).to(self.device)
        self._in_scale = torch.from_numpy(extra["input_scale"]).to(self.device)
        self._dims_proto = dims

    def _handle(self, T: int):
        if T not in self._handles:
            d = self._dims_proto
            dims = _cabi.ConformerDims(d.d_model, d.n_heads, d.d_ff, d.n_blocks, d.kernel_size, d.in_features, d.n_out,
                                       d.maxlen, T, self.gemm_engine)
            h = C.c_void_p()
            offs = (C.c_int64 * len(self._offsets))(*self._offsets.tolist())
            _cabi.check(self._lib.nsf_conformer_create(C.byref(dims), _cabi.ptr(self._blob), self._blob.numel(), offs,
                                                       len(self._offsets), C.byref(h)), "nsf_conformer_create")
            self._handles[T] = (h, dims)
        return self._handles[T]

    def __del__(self):
        try:
            for h, _ in self._handles.values():
                self._lib.nsf_conformer_destroy(h)
        except Exception:
            pass

    def _workspace(self, dims, n_seg: int):
        need = int(self._lib.nsf_conformer_workspace_bytes(C.byref(dims), n_seg))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws, need

    # ---- batched stages (device tensors in, device tensors out) -------------------------------
    def stft_device(self, x: torch.Tensor, T_alloc: Optional[int] = None) -> torch.Tensor:
        """x [N, C] float32 on the device -> X [F, T_alloc, C] complex64 (frames beyond the signal are zero)."""
        self._require_cuda()
        assert x.dim() == 2 and x.dtype == torch.float32 and x.is_cuda and x.is_contiguous()
        n, c = x.shape
        nf = int(self._lib.nsf_num_frames(n))
        T_alloc = max(nf, T_alloc or 0)
        X = torch.zeros((NUM_BINS, T_alloc, c), dtype=torch.complex64, device=x.device)
        _cabi.check(self._lib.nsf_stft_mc(_cabi.ptr(x), n, c, _cabi.ptr(X), T_alloc, nf, _cabi.stream_ptr()), "nsf_stft_mc")
        return X

    def stft_alloc(self, n_ch: int, T_alloc: int, T_valid: int, device) -> torch.Tensor:
        """X [F, T_alloc, C] complex64 with the frames >= T_valid (zero padding of short inputs, css.py:159-164) cleared."""
        X = torch.empty((NUM_BINS, T_alloc, n_ch), dtype=torch.complex64, device=device)
        if T_alloc > T_valid:
            X[:, T_valid:, :].zero_()
        return X

    def stft_frames(self, x: torch.Tensor, X: torch.Tensor, f0: int, f1: int):
        """Frames [f0, f1) of x [N, C] into X[:, f0:f1, :] (the rest of X is untouched)."""
        self._require_cuda()
        n, c = x.shape
        assert x.is_contiguous() and X.is_contiguous() and X.shape[2] == c and 0 <= f0 <= f1 <= int(self._lib.nsf_num_frames(n))
        if f1 == f0:
            return
        xo = x[f0 * FRAME_HOP:]
        Xo = X[:, f0:, :]
        _cabi.check(self._lib.nsf_stft_mc(_cabi.ptr(xo), n - f0 * FRAME_HOP, c, _cabi.ptr(Xo), X.shape[1], f1 - f0,
                                          _cabi.stream_ptr()), "nsf_stft_mc")

    def features(self, X: torch.Tensor, T_valid: int, seg_first: int, n_seg: int, T: int, hop: int,
                 normalize_input: bool = False, split: bool = False):
        """feat [n_seg*T, ldf] (and its TF32 remainder when split) for segments seg_first.. of X."""
        self._require_cuda()
        F_, T_long, c = X.shape
        assert F_ == NUM_BINS and X.dtype == torch.complex64 and X.is_contiguous()
        rows = n_seg * T
        fmt = _cabi.split_fmt_of_engine(self.gemm_engine) if split else _cabi.SPLIT_TF32
        # 16-bit split formats: bf16 / scaled-fp16 bit patterns (the kernel zero-fills the K padding columns)
        feat = torch.empty((rows, self.ldf), dtype=torch.float32 if fmt == _cabi.SPLIT_TF32 else torch.int16, device=X.device)
        feat_lo = torch.empty_like(feat) if split else None
        _cabi.check(self._lib.nsf_css_features(
            _cabi.ptr(X), T_long, T_valid, c, seg_first, n_seg, T, hop,
            _cabi.ptr(self._in_bias) if normalize_input else None, _cabi.ptr(self._in_scale) if normalize_input else None,
            _cabi.ptr(feat), _cabi.ptr(feat_lo), self.ldf, fmt, _cabi.stream_ptr()), "nsf_css_features")
        return feat, feat_lo

    def masks_from_features(self, feat: torch.Tensor, feat_lo: Optional[torch.Tensor], n_seg: int, T: int,
                            out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Mask network on already input-normalised features -> [n_seg, num_masks, F, T]."""
        self._require_cuda()
        h, dims = self._handle(T)
        ws, need = self._workspace(dims, n_seg)
        if out is None:
            out = torch.empty((n_seg, self.num_masks, NUM_BINS, T), dtype=torch.float32, device=feat.device)
        _cabi.check(self._lib.nsf_conformer_forward(h, _cabi.ptr(feat), _cabi.ptr(feat_lo), feat.shape[1], n_seg, _cabi.ptr(out),
                                                    _cabi.ptr(ws), need, _cabi.stream_ptr()), "nsf_conformer_forward")
        return out

    def masks(self, X: torch.Tensor, T_valid: int, seg_first: int, n_seg: int, T: int, hop: int,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        split = self.gemm_engine != _cabi.GEMM_SIMT_FP32
        feat, feat_lo = self.features(X, T_valid, seg_first, n_seg, T, hop, normalize_input=True, split=split)
        return self.masks_from_features(feat, feat_lo, n_seg, T, out=out)

    def mvdr(self, masks: torch.Tensor, X: torch.Tensor, T_valid: int, seg_first: int, hop: int, mask_floor: float,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """masks [n_seg, S+Nn, F, T], X [F, T_long, C] -> Y [n_seg, S, F, T] complex64."""
        self._require_cuda()
        n_seg, n_m, n_bins, T = masks.shape
        assert masks.dtype == torch.float32 and masks.is_contiguous() and X.is_contiguous()
        if out is None:
            out = torch.empty((n_seg, self.num_spks, n_bins, T), dtype=torch.complex64, device=masks.device)
        _cabi.check(self._lib.nsf_mvdr(_cabi.ptr(masks), self.num_spks, n_m - self.num_spks, _cabi.ptr(X), X.shape[1], T_valid,
                                       X.shape[2], seg_first, n_seg, T, hop, n_bins, float(mask_floor), _cabi.ptr(out),
                                       _cabi.stream_ptr()), "nsf_mvdr")
        return out

    def mvdr_utterance(self, spk_masks: torch.Tensor, noise_masks: torch.Tensor, X: torch.Tensor, mask_floor: float = 1.0) -> torch.Tensor:
        """make_mvdr(spk_masks, noise_masks, mix_stft=..., return_stft=True) on ONE utterance of any length (mvdr_util.py:5-47):
        spk_masks [S, F, T], noise_masks [Nn, F, T] float32, X [F, T, C] complex64 -> [S, F, T] complex64 (split-T kernels)."""
        self._require_cuda()
        S, F_, T = spk_masks.shape
        masks = torch.cat([spk_masks, noise_masks], 0).contiguous()
        assert masks.dtype == torch.float32 and X.dtype == torch.complex64 and X.is_contiguous() and tuple(X.shape[:2]) == (F_, T)
        need = int(self._lib.nsf_mvdr_utterance_workspace_bytes(S, T, F_))
        ws = torch.empty(need, dtype=torch.uint8, device=masks.device)
        out = torch.empty((S, F_, T), dtype=torch.complex64, device=masks.device)
        _cabi.check(self._lib.nsf_mvdr_utterance(_cabi.ptr(masks), S, masks.shape[0] - S, _cabi.ptr(X), T, X.shape[2], F_, float(mask_floor),
                                                 _cabi.ptr(out), _cabi.ptr(ws), need, _cabi.stream_ptr()), "nsf_mvdr_utterance")
        return out

    def mask_apply(self, masks: torch.Tensor, X: torch.Tensor, T_valid: int, seg_first: int, hop: int, mask_floor: float,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Separation without the beamformer (css.py:218-227): reference channel x floored mask -> Y [n_seg, S, F, T]."""
        self._require_cuda()
        n_seg, n_m, n_bins, T = masks.shape
        assert masks.dtype == torch.float32 and masks.is_contiguous() and X.is_contiguous()
        if out is None:
            out = torch.empty((n_seg, self.num_spks, n_bins, T), dtype=torch.complex64, device=masks.device)
        _cabi.check(self._lib.nsf_mask_apply(_cabi.ptr(masks), self.num_spks, n_m - self.num_spks, _cabi.ptr(X), X.shape[1], T_valid,
                                             X.shape[2], seg_first, n_seg, T, hop, n_bins, float(mask_floor), _cabi.ptr(out),
                                             _cabi.stream_ptr()), "nsf_mask_apply")
        return out

    def power_norm(self, Y: torch.Tensor, X: torch.Tensor, T_valid: int, seg_first: int, hop: int, mix_frames: int) -> torch.Tensor:
        """CssCfg.normalize_segment_power (css.py:233-247), in place on Y [n_seg, S, F, T]; returns the per-segment ratios."""
        self._require_cuda()
        n_seg, S, n_bins, T = Y.shape
        assert Y.dtype == torch.complex64 and Y.is_contiguous() and X.is_contiguous()
        ratio = torch.empty((n_seg,), dtype=torch.float32, device=Y.device)
        _cabi.check(self._lib.nsf_segment_power_norm(_cabi.ptr(Y), S, _cabi.ptr(X), X.shape[1], T_valid, X.shape[2], seg_first, n_seg,
                                                     T, hop, n_bins, mix_frames, _cabi.ptr(ratio), _cabi.stream_ptr()),
                    "nsf_segment_power_norm")
        return ratio

    def istft_device(self, S_st: torch.Tensor) -> torch.Tensor:
        """S_st [n_streams, T_long, F] complex64 (frame-major) -> wav [n_streams, (T_long-1)*256+512]."""
        self._require_cuda()
        n_streams, T_long, F_ = S_st.shape
        assert F_ == NUM_BINS and S_st.dtype == torch.complex64 and S_st.is_contiguous()
        wav = torch.empty((n_streams, (T_long - 1) * FRAME_HOP + FRAME_LEN), dtype=torch.float32, device=S_st.device)
        _cabi.check(self._lib.nsf_istft(_cabi.ptr(S_st), n_streams, T_long, _cabi.ptr(wav), _cabi.stream_ptr()), "nsf_istft")
        return wav

    # ---- the reference's separator protocol (conformer_wrapper.py:79-146) ---------------------
    def stft(self, s: torch.Tensor) -> torch.Tensor:
        """[Batch, T, Mics] (or [Batch, T]) float -> [Batch, F, T, Mics] (or [Batch, F, T]) complex64."""
        self._require_cuda()
        squeeze = s.dim() == 2
        if squeeze:
            s = s.unsqueeze(-1)
        outs = []
        for b in range(s.shape[0]):
            x = s[b].to(self.device, torch.float32).contiguous()
            outs.append(self.stft_device(x))
        X = torch.stack(outs)
        return X[..., 0] if squeeze else X

    def separate(self, stft: torch.Tensor):
        """[Batch, F, T, Mics] complex -> {'spk_masks': [Batch, F, T, S], 'noise_masks': [Batch, F, T, Nn]}."""
        self._require_cuda()
        assert torch.is_complex(stft)
        if stft.dim() == 3:                                 # single-channel model: [Batch, F, T]
            stft = stft.unsqueeze(-1)
        stft = stft.to(self.device)
        B, F_, T, c = stft.shape
        outs = []
        for b in range(B):
            X = stft[b].contiguous()
            outs.append(self.masks(X, T_valid=T, seg_first=0, n_seg=1, T=T, hop=T)[0])
        m = torch.stack(outs).permute(0, 2, 3, 1)          # [B, F, T, masks]
        return {"spk_masks": m[..., :self.num_spks], "noise_masks": m[..., self.num_spks:]}

    def istft(self, stft: torch.Tensor) -> torch.Tensor:
        """[Batch, F, T] complex -> [Batch, NSamples]."""
        self._require_cuda()
        assert torch.is_complex(stft) and stft.dim() == 3
        S_st = stft.to(self.device).permute(0, 2, 1).contiguous()
        return self.istft_device(S_st)

    def forward(self, mix: torch.Tensor):
        return self.separate(self.stft(mix))

    __call__ = forward
