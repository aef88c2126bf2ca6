"""Drop-in replacement for the reference's CSS plug-in (css/css.py): same names, same signatures,
same outputs -- ``CssCfg``, ``css_inference``, ``separate_and_stitch``, ``calc_segment_weight`` --
with every stage running on a B200 through libnsf_b200.so.

Reference flow (css/css.py:110-338) vs this module
  * STFT of the whole meeting on the CPU (:155)             -> one STFT kernel, X stays in HBM
  * per 3-s segment: H2D, Conformer, D2H, NumPy MVDR, H2D, D2H (:182-250, 4 PCIe crossings/segment)
                                                             -> segments are the batch dimension; features,
                                                                mask network and MVDR run per chunk of
                                                                ``segments_per_batch`` segments, nothing leaves HBM
  * sequential PIT alignment + overlap-add on the CPU (:266-299)
                                                             -> all 3x3 costs in one kernel, the 6-permutation
                                                                chain on the host (3 ints per segment), then one
                                                                gather-style overlap-add kernel (no atomics)
  * activity gate with NumPy morphology (:303-312)          -> threshold / dilate / erode kernels
  * iSTFT on the CPU (:316-319)                              -> iSTFT kernel; only the 3 waveforms cross PCIe
"""
from __future__ import annotations

import ctypes
import os
import itertools
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _cabi
from .separator import ConformerCssB200, NUM_BINS, FRAME_HOP, FRAME_LEN


# CSS inference configuration -- field for field the reference's CssCfg (css/css.py:24-48)
@dataclass
class CssCfg:
    segment_size_sec: float = 3.  # in seconds
    hop_size_sec: float = 1.5     # in seconds
    normalize_segment_power: bool = False
    stitching_loss: str = 'l1'  # loss function for stitching adjacent segments ('l1' or 'mse')
    stitching_input: str = 'mask'  # type of input for stitching loss ('mask' or 'separation_result')
    seg_weight_m0_sec: float = 0.15  # see calc_segment_weight
    seg_weight_m1_sec: float = 0.3
    activity_th: float = 0.4  # threshold for segmentation mask
    activity_dilation_sec: float = 0.4  # dilation and erosion for segmentation mask
    activity_erosion_sec: float = 0.2
    device: Optional[str] = None
    show_progressbar: bool = True
    # segment-wise single-channel model
    checkpoint_sc: str = 'notsofar/conformer1.0/sc'
    # segment-wise multi-channel model
    checkpoint_mc: str = 'notsofar/conformer1.0/mc'
    device_id: int = 0
    num_spks: int = 3  # the number of streams the separation models outputs
    mc_mvdr: bool = True  # if True, applies MVDR to the multi-channel input
    mc_mask_floor_db: float = 0.  # mask floor in db. -inf means no floor. 0 means mask has no effect
    sc_mask_floor_db: float = -np.inf
    pass_through_ch0: bool = False  # if True, simply returns the first channel of the input and skips CSS
    slice_audio_for_debug: bool = False  # if True, only processes 10 seconds of the input audio


@dataclass
class SegmentPlan:
    """Integer frame bookkeeping of css.py:141-169 (bit-exact by construction: same int() truncations)."""
    segment_frames: int
    hop_frames: int
    m0_frames: int
    m1_frames: int
    dilation_frames: int
    erosion_frames: int
    raw_frames: int      # STFT frames of the signal
    mix_frames: int      # after zero-padding inputs shorter than one segment (css.py:159-164)
    num_segments: int

    @property
    def overlap_frames(self) -> int:
        return self.segment_frames - self.hop_frames


def _num_frames(n_samples: int) -> int:
    return 0 if n_samples < FRAME_LEN else (n_samples - FRAME_LEN) // FRAME_HOP + 1


def plan_segments(n_samples: int, fs: int, cfg: CssCfg) -> SegmentPlan:
    seg = _num_frames(int(cfg.segment_size_sec * fs))        # dummy 3-s STFT of css.py:142-144
    hop = int(seg * cfg.hop_size_sec / cfg.segment_size_sec)
    m0 = int(seg * cfg.seg_weight_m0_sec / cfg.segment_size_sec)
    m1 = int(seg * cfg.seg_weight_m1_sec / cfg.segment_size_sec)
    dil = int(seg * cfg.activity_dilation_sec / cfg.segment_size_sec)
    ero = int(seg * cfg.activity_erosion_sec / cfg.segment_size_sec)
    raw = _num_frames(n_samples)
    mix = max(raw, seg)
    nseg = int(np.ceil((mix - (seg - hop)) / hop))
    return SegmentPlan(seg, hop, m0, m1, dil, ero, raw, mix, nseg)


def calc_segment_weight(seg_frames: int, m0_frames: int, m1_frames: int,
                        is_first_seg: bool = False, is_last_seg: bool = False):
    """Trapezoid weights of css.py:341-390 (0 on [0, m0), linear 0.1 -> 1 on [m0, m1), 1 in the middle,
    mirrored on the right; the outer edge of the first / last segment is 0.1).  Returns a float32
    torch tensor like the reference."""
    assert seg_frames > 2 * m1_frames, \
        'not enough frames to fit weighting window. try modifying hop_size, segment_size or m0, m1'
    wg_win = torch.ones(seg_frames, dtype=torch.float32)
    wg_win[:m0_frames] = 0
    wg_win[len(wg_win) - m0_frames:] = 0
    linear = torch.linspace(0.1, 1, m1_frames - m0_frames)
    wg_win[m0_frames:m1_frames] = linear
    wg_win[-m1_frames:-m0_frames] = torch.flip(linear, (0,))
    if is_first_seg:
        wg_win[:m0_frames] = 0.1
    if is_last_seg:
        wg_win[len(wg_win) - m0_frames:] = 0.1
    return wg_win


_PERM_TABLES: Dict[int, tuple] = {}


def _perm_tables(S: int):
    """All S! channel orders, and next-state bookkeeping for the chain: cand [P, S]."""
    if S not in _PERM_TABLES:
        cand = np.asarray(list(itertools.permutations(range(S))), dtype=np.int64)
        _PERM_TABLES[S] = (cand,)
    return _PERM_TABLES[S]


def permutation_chain(costs: np.ndarray, prev_state: Optional[int] = None, return_state: bool = False):
    """Sequential alignment of css.py:266-285 from the pairwise costs of the *unpermuted* segments.

    costs[i][a][b] = loss(left channel a of segment i-1, right channel b of segment i) on the original
    channel order.  The reference aligns segment i against the already permuted segment i-1, i.e. it sees
    rows perm[i-1] of costs[i]; the optimal assignment of a 3x3 (<= 4x4) matrix is found by enumerating
    the permutations (== scipy's Hungarian answer away from exact ties).  Returns perms [n_seg, S] with
    new channel k of segment i <- old channel perms[i][k].

    The enumeration is vectorised over segments: total[i][q][p] = sum_a costs[i][cand[q][a]][cand[p][a]] (float64,
    summed in channel order) for every possible previous order q, best[i][q] = first arg-min over p; the chain
    itself is then a table walk (0.3 ms instead of 14 ms of Python for a 30-minute meeting, during which the GPU
    would sit idle behind the cost read-back).

    prev_state / return_state continue a chain: with prev_state = the state returned for segments [0, s), costs[s:] yields
    the rows s.. of the one-shot call (the progressive tail of css_device walks the chain chunk by chunk).
    """
    n_seg, S, _ = costs.shape
    (cand,) = _perm_tables(S)
    P = len(cand)
    c = np.ascontiguousarray(np.asarray(costs, dtype=np.float64).transpose(1, 2, 0))      # [S, S, n_seg]: contiguous per entry
    total = np.empty((P, P, n_seg), np.float64)
    for qi in range(P):
        for pi in range(P):
            acc = total[qi, pi]
            np.copyto(acc, c[cand[qi, 0], cand[pi, 0]])   # ((0 + c_0) + c_1) + c_2: the order of the scalar loop it replaces
            for a in range(1, S):
                acc += c[cand[qi, a], cand[pi, a]]
    bval = total[:, 0].copy()                              # [P (previous order), n_seg]
    bidx = np.zeros((P, n_seg), np.int64)
    for pi in range(1, P):                                 # strict '<': the first minimum wins, like the scalar loop
        better = total[:, pi] < bval
        bidx[better] = pi
        np.minimum(bval, total[:, pi], out=bval)
    best = np.ascontiguousarray(bidx.T)                    # [n_seg, P]
    if prev_state is None:
        best[0] = np.arange(P)                           # segment 0 keeps its order
    # state_i = best[i][state_{i-1}], state_{-1} = 0 (identity = cand[0]).  The maps compose associatively: blocked prefix
    # composition -- inside blocks of ~sqrt(n) segments vectorised across blocks, then one short carry walk over the
    # blocks (0.6 ms for the 9 677 segments every rank of a sharded 4-hour meeting replays with its GPU waiting).
    L = max(1, int(np.sqrt(n_seg)))
    nb = -(-n_seg // L)
    G = np.tile(np.arange(P, dtype=np.int64), (nb * L, 1))
    G[:n_seg] = best
    G = G.reshape(nb, L, P)
    rows = np.arange(nb)[:, None]
    for j in range(1, L):
        G[:, j] = G[:, j][rows, G[:, j - 1]]             # G_j <- best_j o G_{j-1} inside every block
    carry = np.zeros(nb, np.int64)
    last = G[:, L - 1].tolist()
    q = 0 if prev_state is None else int(prev_state)
    for b in range(nb):
        carry[b] = q
        q = last[b][q]
    states = np.take_along_axis(G, np.broadcast_to(carry[:, None, None], (nb, L, 1)), axis=2).reshape(-1)[:n_seg]
    perms = cand[states].astype(np.int32)
    return (perms, int(states[-1])) if return_state else perms


def plan_batches(n_seg: int, max_batch: int, streaming: bool = False, first_batch: int = 176, progressive: int = 0,
                 last_batch: int = 0):
    """Chunks of segments the mask network / MVDR run on: [(first segment, count), ...].

    Large, equal chunks keep the persistent GEMMs' last wave full (a 128 x 256 tile grid over 148 SMs quantises badly for
    small M); when the audio is still streaming in from the host the first chunk is kept short so that compute starts
    after the first few MB have landed -- but long enough (176 segments = 15 ms of network time) that the rest of a
    30-minute recording (0.8 GB, ~15 ms over PCIe) has landed when the second chunk wants it.

    progressive > 0 (the separated waveforms go back to the host piece by piece, css_device(host_out=...)): what follows the
    first chunk is cut into chunks of at most that many segments, so that all but the last chunk's share of the output has
    crossed PCIe when the network finishes.  last_batch > 0 cuts a short chunk off the end of the last one: what is copied
    after the last kernel shrinks with it (worth a fourth chunk where several ranks share the host's memory bandwidth)."""
    max_batch = max(1, max_batch)
    if streaming and progressive > 0 and n_seg >= 2 * first_batch:
        max_batch = min(max_batch, max(progressive, first_batch + 1))
        out, s0 = [(0, first_batch)], first_batch
        while n_seg - s0 > max_batch + max_batch // 4:          # full chunks, then the remainder (a short one is merged)
            out.append((s0, max_batch))
            s0 += max_batch
        rest = n_seg - s0
        if last_batch > 0 and rest >= 2 * last_batch:
            out.append((s0, rest - last_batch))
            s0 += rest - last_batch
        out.append((s0, n_seg - s0))
        return out
    out, s0 = [], 0
    if streaming and n_seg >= 2 * first_batch and max_batch > first_batch:      # short sessions: one chunk (a small second chunk wastes GEMM waves)
        out.append((0, first_batch))
        s0 = first_batch
    rest = n_seg - s0
    if rest > 0:
        parts = -(-rest // max_batch)
        base, extra = divmod(rest, parts)
        for i in range(parts):
            nb = base + (1 if i < extra else 0)
            out.append((s0, nb))
            s0 += nb
    return out


_SEG_W_CACHE: Dict[tuple, tuple] = {}


def _segment_weights(plan: SegmentPlan):
    """seg_w [n_seg, T] and its overlap-added sum wg_stitched [mix_frames] exactly as css.py:258-259,288-291
    accumulate them (float32, ascending segment order).  Cached per plan: callers must not modify the arrays."""
    key = (plan.segment_frames, plan.hop_frames, plan.m0_frames, plan.m1_frames, plan.num_segments, plan.mix_frames)
    if key in _SEG_W_CACHE:
        return _SEG_W_CACHE[key]
    T, n, hop = plan.segment_frames, plan.num_segments, plan.hop_frames
    first = calc_segment_weight(T, plan.m0_frames, plan.m1_frames, is_first_seg=True).numpy()
    mid = calc_segment_weight(T, plan.m0_frames, plan.m1_frames).numpy()
    last = calc_segment_weight(T, plan.m0_frames, plan.m1_frames, is_last_seg=True).numpy()
    seg_w = np.empty((n, T), np.float32)
    seg_w[:] = mid
    seg_w[n - 1] = last
    seg_w[0] = first                                     # a lone segment is a "first" one (css.py:257-259 before :283-284)
    wsum = np.zeros(plan.mix_frames, np.float32)
    if T <= 2 * hop:
        # at most two segments meet in a frame: even and odd segments tile the axis without overlap, and a + b is
        # the same float32 whichever comes first
        for parity in (0, 1):
            part = np.zeros(plan.mix_frames + T, np.float32)
            idx = np.arange(parity, n, 2)
            if len(idx):
                pos = (idx[:, None] * hop + np.arange(T)[None, :]).reshape(-1)
                part[pos] = seg_w[idx].reshape(-1)
            wsum += part[:plan.mix_frames]
    else:
        for i in range(n):
            st = i * hop
            en = min(st + T, plan.mix_frames)
            wsum[st:en] += seg_w[i][:en - st]
    _SEG_W_CACHE[key] = (seg_w, wsum)
    if len(_SEG_W_CACHE) > 16:
        _SEG_W_CACHE.pop(next(iter(_SEG_W_CACHE)))
    return seg_w, wsum


_SEG_W_DEV: Dict[tuple, tuple] = {}


def _segment_weights_device(plan: SegmentPlan, device: torch.device):
    """_segment_weights on the device, kept per plan: the two small uploads would otherwise queue behind the recording on the
    host -> device copy engine (15 ms for a 30-min meeting) when the tail is set up before the network runs."""
    key = (plan.segment_frames, plan.hop_frames, plan.m0_frames, plan.m1_frames, plan.num_segments, plan.mix_frames, str(device))
    hit = _SEG_W_DEV.get(key)
    if hit is None:
        seg_w_np, wsum_np = _segment_weights(plan)
        assert (wsum_np > 1e-5).all(), 'zero weights found. check hop_size, segment_size or m0, m1'
        hit = (torch.from_numpy(seg_w_np).to(device), torch.from_numpy(wsum_np).to(device))
        _SEG_W_DEV[key] = hit
        if len(_SEG_W_DEV) > 8:
            _SEG_W_DEV.pop(next(iter(_SEG_W_DEV)))
    return hit


class HostFeeder:
    """Streams a pinned (or pageable) host recording [Nsamples, Channels] into HBM in chunks on a side stream, so the
    STFT / mask network of the first segments run while the rest of the meeting is still crossing PCIe.
    ``ready(n)`` makes the *current* stream wait until samples [0, n) have landed."""

    def __init__(self, x_host: torch.Tensor, device: torch.device, chunk_samples: int):
        assert x_host.dim() == 2 and x_host.dtype == torch.float32 and not x_host.is_cuda
        n = x_host.shape[0]
        self.x_dev = torch.empty(x_host.shape, dtype=torch.float32, device=device)
        self.stream = _side_stream(device)
        self.stream.wait_stream(torch.cuda.current_stream(device))     # x_dev may recycle memory still in use upstream
        self.bounds, self.events = [], []
        chunk_samples = max(int(chunk_samples), FRAME_LEN)
        with torch.cuda.stream(self.stream):
            for s0 in range(0, n, chunk_samples):
                s1 = min(n, s0 + chunk_samples)
                self.x_dev[s0:s1].copy_(x_host[s0:s1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                self.bounds.append(s1)
                self.events.append(ev)
        self._next = 0

    def ready(self, n_samples: int):
        cur = torch.cuda.current_stream(self.x_dev.device)
        while self._next < len(self.bounds) and (self._next == 0 or self.bounds[self._next - 1] < n_samples):
            cur.wait_event(self.events[self._next])
            self._next += 1


_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}
_TAIL_STREAMS: Dict[int, "torch.cuda.Stream"] = {}
_SMALL_PINNED: Dict[tuple, "torch.Tensor"] = {}
# Segments per chunk behind the first one when the output streams back piece by piece; 0: one-shot tail.  356 segments of
# 186 frames are 259 = 7 x 37 row tiles of 256: every GEMM of the block (2, 6 or 8 column tiles) fills its last wave of
# 74 CTA pairs (345-segment chunks measured 4 % slower per segment).
PROGRESSIVE_CHUNK = int(os.environ.get("NSF_PROGRESSIVE_CHUNK", "356"))


def _tail_stream(device: torch.device):
    """High-priority stream of the progressive tail (stitching + iSTFT + D2H of what is final) next to the mask network."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _TAIL_STREAMS:
        _TAIL_STREAMS[idx] = torch.cuda.Stream(device, priority=-1)      # 0.4-0.7 ms per chunk; 1.7 ms when it has to share fairly
    return _TAIL_STREAMS[idx]


def _small_pinned(tag: str, shape: tuple, dtype) -> "torch.Tensor":
    """Page-locked scratch for the few KB of costs / permutations that cross PCIe per chunk (kept: pinning is a syscall)."""
    key = (tag, dtype)
    n = int(np.prod(shape))
    buf = _SMALL_PINNED.get(key)
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 4096), dtype=dtype, pin_memory=True)
        _SMALL_PINNED[key] = buf
    return buf[:n].view(shape)


def _side_stream(device: torch.device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[idx]


@torch.no_grad()
def css_device(x, separator: ConformerCssB200, fs: int, cfg: CssCfg, want_side_info: bool = True,
               host_out: Optional[torch.Tensor] = None) -> Dict:
    """The whole CSS path on tensors that already live in HBM (or are on their way there: x may be a HostFeeder).

    x: [Nsamples, Channels] float32 on the CUDA device.  Returns a dict of device tensors:
    'wav' [S, N'], 'mask_stitched' [F, T_long, S], 'activity' [T_long, S] float, 'activity_b' / 'activity_final'
    [T_long, S] uint8, plus the per-segment 'masks' [n_seg, S+Nn, F, T], 'Y' [n_seg, S, F, T], 'X' [F, T_long, C],
    'perms' (numpy [n_seg, S]) and 'plan'.  The only host round trip inside is the [n_seg, S, S] cost matrix of
    the permutation chain (css.py:266-285), 36 bytes per segment.

    host_out: optional page-locked [S, N'] float32 tensor that receives 'wav' (asynchronously: synchronise the current
    stream before reading it).  With it, and more than one chunk of segments, the tail runs *progressively*: after every
    chunk the costs of its segments go to the host, the chain advances, and nsf_stitch_progress stitches / gates /
    inverse-transforms the frames that have become final on a side stream whose D2H copies overlap the next chunk's
    mask network -- only the last chunk's share of the output is copied after the network has finished.  Every tail
    stage is local in time, so the result is bit for bit that of the one-shot tail (tests/test_gpu_parity.py).
    """
    feeder = x if isinstance(x, HostFeeder) else None
    if feeder is not None:
        x = feeder.x_dev
    assert x.dim() == 2 and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    n_samples, num_channels = x.shape
    device = x.device
    lib = _cabi.load()
    assert num_channels == separator.num_mics, f'the model expects {separator.num_mics} channel(s), got {num_channels}'
    use_mvdr = num_channels > 1 and cfg.mc_mvdr              # css.py:209
    assert cfg.stitching_loss in ('l1', 'mse'), f'unexpected stitching_loss: {cfg.stitching_loss}'
    assert cfg.stitching_input in ('mask', 'separation_result'), f'unexpected stitching_input: {cfg.stitching_input}'

    plan = plan_segments(n_samples, fs, cfg)
    T, hop, S = plan.segment_frames, plan.hop_frames, cfg.num_spks
    assert S == separator.num_spks
    mix_frames, n_seg = plan.mix_frames, plan.num_segments
    mask_floor_db = cfg.mc_mask_floor_db if num_channels > 1 else cfg.sc_mask_floor_db
    assert mask_floor_db <= 0
    mask_floor = 10. ** (mask_floor_db / 20.)

    with torch.cuda.device(device):
        sp = _cabi.stream_ptr
        T_valid = plan.raw_frames
        X = separator.stft_alloc(num_channels, mix_frames, T_valid, device)   # [F, mix_frames, C], padding frames zeroed
        frames_done = 0

        # I. masks + MVDR per chunk of segments (css.py:182-250; segments are the batch dimension here)
        n_masks = separator.num_masks
        masks = torch.empty((n_seg, n_masks, NUM_BINS, T), dtype=torch.float32, device=device)
        Y = torch.empty((n_seg, S, NUM_BINS, T), dtype=torch.complex64, device=device)
        chunks = plan_batches(n_seg, int(separator.segments_per_batch), streaming=feeder is not None,
                              progressive=PROGRESSIVE_CHUNK if host_out is not None else 0)
        progressive = host_out is not None and len(chunks) > 1 and PROGRESSIVE_CHUNK > 0
        costs = torch.empty((n_seg, S, S), dtype=torch.float32, device=device)
        in_kind = 0 if cfg.stitching_input == 'mask' else 1
        loss_kind = 0 if cfg.stitching_loss == 'l1' else 1
        src = masks if in_kind == 0 else Y
        n_out = (mix_frames - 1) * FRAME_HOP + FRAME_LEN
        if host_out is not None:
            assert tuple(host_out.shape) == (S, n_out) and host_out.dtype == torch.float32 and not host_out.is_cuda
        if progressive:
            # II'/III'. everything behind the chain, advanced chunk by chunk on a side stream (nsf_stitch_progress)
            seg_w, wsum = _segment_weights_device(plan, device)
            mask_st = torch.empty((NUM_BINS, mix_frames, S), dtype=torch.float32, device=device)
            activity = torch.empty((mix_frames, S), dtype=torch.float32, device=device)
            act_b = torch.empty((mix_frames, S), dtype=torch.uint8, device=device)
            act_tmp = torch.empty_like(act_b)
            act_final = torch.empty_like(act_b)
            S_st = torch.empty((S, mix_frames, NUM_BINS), dtype=torch.complex64, device=device)
            wav = torch.empty((S, n_out), dtype=torch.float32, device=device)
            perms = torch.empty((n_seg, S), dtype=torch.int32, device=device)
            costs_host = _small_pinned("costs", (n_seg, S, S), torch.float32)
            perms_host = _small_pinned("perms", (n_seg, S), torch.int32)
            tail = _tail_stream(device)
            main = torch.cuda.current_stream(device)
            tail.wait_stream(main)                     # the buffers above may recycle memory still in use upstream
            cost_events = []
            hops = (ctypes.c_int64 * 2)()
            chain_state = [None]
            th = float(np.float32(cfg.activity_th))

            def advance(ci):
                c0, cn = chunks[ci]
                cost_events[ci].synchronize()          # the GPU is busy with chunk ci + 1 meanwhile
                p_np, chain_state[0] = permutation_chain(costs_host[c0:c0 + cn].numpy(), prev_state=chain_state[0], return_state=True)
                perms_host[c0:c0 + cn] = torch.from_numpy(p_np)
                with torch.cuda.stream(tail):
                    tail.wait_event(cost_events[ci])   # the tail of this chunk reads its masks / Y (complete: formal ordering)
                    perms[c0:c0 + cn].copy_(perms_host[c0:c0 + cn], non_blocking=True)
                    _cabi.check(lib.nsf_stitch_progress(
                        _cabi.ptr(masks), n_masks, _cabi.ptr(Y), _cabi.ptr(perms), _cabi.ptr(seg_w), _cabi.ptr(wsum), n_seg, c0, c0 + cn,
                        S, NUM_BINS, T, hop, mix_frames, th, plan.dilation_frames, plan.erosion_frames, _cabi.ptr(mask_st),
                        _cabi.ptr(activity), _cabi.ptr(act_b), _cabi.ptr(act_tmp), _cabi.ptr(act_final), _cabi.ptr(S_st), _cabi.ptr(wav),
                        hops, sp()), "nsf_stitch_progress")
                    a, b = int(hops[0]) * FRAME_HOP, min(int(hops[1]) * FRAME_HOP, n_out)
                    if b > a:
                        for k in range(S):             # contiguous runs: plain cudaMemcpyAsync (a strided copy_ is staged through pageable memory)
                            host_out[k, a:b].copy_(wav[k, a:b], non_blocking=True)

        for ci, (s0, nb) in enumerate(chunks):
            # STFT of the frames this chunk of segments needs (and, when streaming from the host, only once they landed)
            f_need = min(T_valid, (s0 + nb - 1) * hop + T)
            if f_need > frames_done:
                if feeder is not None:
                    feeder.ready((f_need - 1) * FRAME_HOP + FRAME_LEN)
                separator.stft_frames(x, X, frames_done, f_need)
                frames_done = f_need
            separator.masks(X, T_valid, s0, nb, T, hop, out=masks[s0:s0 + nb])
            if use_mvdr:
                separator.mvdr(masks[s0:s0 + nb], X, T_valid, s0, hop, mask_floor, out=Y[s0:s0 + nb])
            else:
                separator.mask_apply(masks[s0:s0 + nb], X, T_valid, s0, hop, mask_floor, out=Y[s0:s0 + nb])
            if cfg.normalize_segment_power:
                separator.power_norm(Y[s0:s0 + nb], X, T_valid, s0, hop, mix_frames)
            if progressive:
                _cabi.check(lib.nsf_pit_cost_range(_cabi.ptr(src), in_kind, loss_kind, s0, s0 + nb, n_masks if in_kind == 0 else S, S,
                                                   NUM_BINS, T, plan.overlap_frames, _cabi.ptr(costs), sp()), "nsf_pit_cost_range")
                costs_host[s0:s0 + nb].copy_(costs[s0:s0 + nb], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(main)
                cost_events.append(ev)
                if ci >= 1:
                    advance(ci - 1)                    # one chunk behind: the queue of the main stream never runs dry

        if feeder is not None:
            feeder.ready(n_samples)       # every copy has been ordered before the buffers can be recycled

        if progressive:
            advance(len(chunks) - 1)
            main.wait_stream(tail)                     # downstream readers (and the allocator) see the tail's work
            perms_np = perms_host.numpy().copy()
        else:
            # II. permutation chain + weighted overlap-add (css.py:254-299)
            _cabi.check(lib.nsf_pit_cost(_cabi.ptr(src), in_kind, loss_kind, n_seg, n_masks if in_kind == 0 else S, S, NUM_BINS, T,
                                         plan.overlap_frames, _cabi.ptr(costs), sp()), "nsf_pit_cost")
            # host work that does not depend on the costs runs while the GPU is still busy with the segments
            seg_w_np, wsum_np = _segment_weights(plan)
            assert (wsum_np > 1e-5).all(), 'zero weights found. check hop_size, segment_size or m0, m1'
            seg_w = torch.from_numpy(seg_w_np).to(device)
            wsum = torch.from_numpy(wsum_np).to(device)
            mask_st = torch.empty((NUM_BINS, mix_frames, S), dtype=torch.float32, device=device)
            activity = torch.empty((mix_frames, S), dtype=torch.float32, device=device)
            perms_np = permutation_chain(costs.cpu().numpy())            # the only device -> host sync of the path
            perms = torch.from_numpy(perms_np).to(device)
            _cabi.check(lib.nsf_stitch_masks(_cabi.ptr(masks), n_masks, _cabi.ptr(perms), _cabi.ptr(seg_w), _cabi.ptr(wsum), n_seg, S,
                                             NUM_BINS, T, hop, mix_frames, _cabi.ptr(mask_st), _cabi.ptr(activity), sp()),
                        "nsf_stitch_masks")
            # III. activity gate (css.py:303-312)
            act_b = torch.empty((mix_frames, S), dtype=torch.uint8, device=device)
            act_tmp = torch.empty_like(act_b)
            act_final = torch.empty_like(act_b)
            _cabi.check(lib.nsf_activity(_cabi.ptr(activity), mix_frames, S, float(np.float32(cfg.activity_th)), plan.dilation_frames,
                                         plan.erosion_frames, _cabi.ptr(act_b), _cabi.ptr(act_tmp), _cabi.ptr(act_final), sp()),
                        "nsf_activity")
            S_st = torch.empty((S, mix_frames, NUM_BINS), dtype=torch.complex64, device=device)
            _cabi.check(lib.nsf_stitch_stft(_cabi.ptr(Y), _cabi.ptr(perms), _cabi.ptr(seg_w), _cabi.ptr(wsum), _cabi.ptr(act_final),
                                            n_seg, S, NUM_BINS, T, hop, mix_frames, _cabi.ptr(S_st), sp()), "nsf_stitch_stft")
            wav = separator.istft_device(S_st)                                      # [S, N']
            if host_out is not None:
                host_out.copy_(wav, non_blocking=True)
    return dict(wav=wav, mask_stitched=mask_st, activity=activity, activity_b=act_b, activity_final=act_final, masks=masks,
                Y=Y, X=X, S_st=S_st, costs=costs, perms=perms_np, plan=plan)


@torch.no_grad()
def separate_and_stitch(speech_mix, separator: ConformerCssB200, fs: int, device: torch.device, cfg: CssCfg,
                        return_side_info: bool = True, _stages: Optional[dict] = None) -> (List[np.ndarray], Dict):
    """Block-online CSS of a long-form recording: same contract as the reference's
    separate_and_stitch (css/css.py:110-338).

    Args:
        speech_mix: [Batch=1, Nsamples, Channels] float32 numpy array (or a CPU / pinned / CUDA torch tensor).
        separator: ConformerCssB200 (the B200 re-hosting of the reference's ConformerCssWrapper).
        fs: sample rate.  device: CUDA device.  cfg: CssCfg.
    Returns:
        separated_wavs: list of num_spks float32 arrays [(T_long-1)*256+512];
        side_info: {'mask_stitched' [1,F,T_long,S] float32, 'activity_b' [T_long,S] bool,
                    'activity_final' [1,T_long,S] bool, 'segment_frames' int} (CPU torch tensors, like the reference).
    """
    assert speech_mix.ndim == 3, f'expecting 3 dimensions, got {speech_mix.shape}'
    batch_size = speech_mix.shape[0]
    assert batch_size == 1, 'assuming 1 example in batch. easy to support more.'
    if not isinstance(separator, ConformerCssB200):
        raise TypeError("notsofar_b200.separate_and_stitch drives the fused B200 path and needs a ConformerCssB200 "
                        "separator (build one with ConformerCssB200(reference_state_dict) or load_css_model)")
    assert not separator.training
    device = torch.device(device)
    if device.type != "cuda":
        raise _cabi.NsfError("notsofar_b200 has no CPU path: pass a CUDA device")
    separator.to(device)
    # H2D of the raw audio -- the only input that crosses PCIe -- streamed in chunks behind the compute
    if isinstance(speech_mix, torch.Tensor) and speech_mix.is_cuda:
        x = speech_mix[0].to(device=device, dtype=torch.float32).contiguous()
    else:
        x_host = speech_mix[0] if isinstance(speech_mix, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(speech_mix[0], dtype=np.float32))
        x_host = x_host.to(torch.float32).contiguous()
        plan0 = plan_segments(x_host.shape[0], fs, cfg)
        # copy granularity: the first (short) chunk of segments, so that the STFT of chunk 0 can start early
        chunk = min(32, max(1, int(separator.segments_per_batch))) * plan0.hop_frames * FRAME_HOP   # 21 MB per copy: 0.4 ms
        with torch.cuda.device(device):
            x = HostFeeder(x_host, device, chunk)
    # D2H of the separated streams into a (pooled) pinned buffer, piece by piece behind the mask network (css_device)
    n_samples = x.x_dev.shape[0] if isinstance(x, HostFeeder) else x.shape[0]
    plan0 = plan_segments(n_samples, fs, cfg)
    with torch.cuda.device(device):
        host = _pinned_out((cfg.num_spks, (plan0.mix_frames - 1) * FRAME_HOP + FRAME_LEN))
        out = css_device(x, separator, fs, cfg, host_out=host)
        torch.cuda.current_stream(device).synchronize()
    separated = _lend(host)                 # owned by the caller: the buffer is not reused while these arrays are alive
    separated_wavs = [separated[k] for k in range(cfg.num_spks)]
    del separated
    side_info = {'segment_frames': out["plan"].segment_frames}
    if return_side_info:
        side_info.update({
            'mask_stitched': out["mask_stitched"].cpu().unsqueeze(0),           # [1, F, T_long, S]
            'activity_b': out["activity_b"].cpu().bool(),                       # [T_long, S]
            'activity_final': out["activity_final"].cpu().bool().unsqueeze(0),  # [1, T_long, S]
        })
    if _stages is not None:
        _stages.update(out)
    return separated_wavs, side_info


_PINNED_POOL: Dict[int, list] = {}         # capacity in bytes -> [[pinned uint8 tensor, weakref to the numpy array handed out or None], ...]
_PINNED_KEEP_FREE = 2


def _pinned_out(shape: tuple, dtype=torch.float32) -> torch.Tensor:
    """Page-locked host buffer of the given shape (the separated waveforms; the PCM16 staging of css_inference).  Pinning
    345 MB costs more than the whole separation, so buffers are pooled -- by capacity, rounded up to 64 MB (1 MB for small
    ones), because every recording has its own length.  A buffer is handed out again only after every numpy array that was
    returned on top of it has been garbage collected (``_lend`` keeps a weak reference to it): a caller that keeps the streams
    of three sessions holds three distinct buffers -- the reference returns freshly owned arrays (css.py:316-319), so must this."""
    numel = 1
    for d in shape:
        numel *= int(d)
    nbytes = max(1, numel * torch.empty((), dtype=dtype).element_size())
    gran = (1 << 26) if nbytes > (1 << 24) else (1 << 20)
    cap = -(-nbytes // gran) * gran
    entries = _PINNED_POOL.setdefault(cap, [])
    base = None
    for e in entries:
        if e[1] is None or e[1]() is None:
            e[1] = None
            base = e[0]
            break
    if base is None:
        base = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        entries.append([base, None])
    return base[:nbytes].view(dtype).view(tuple(int(d) for d in shape))


def _lend(buf: torch.Tensor) -> np.ndarray:
    """numpy view of a pooled pinned buffer; the buffer stays out of circulation while the view (or any slice of it) lives."""
    import weakref
    arr = buf.numpy()
    ptr = buf.untyped_storage().data_ptr()
    for entries in _PINNED_POOL.values():
        for e in entries:
            if e[0].untyped_storage().data_ptr() == ptr:
                e[1] = weakref.ref(arr)
        # buffers nobody holds any more beyond a small reserve go back to the OS
        free = [e for e in entries if e[1] is None or e[1]() is None]
        drop = {id(e) for e in free[_PINNED_KEEP_FREE:]}
        if drop:
            entries[:] = [e for e in entries if id(e) not in drop]    # by identity: comparing entries would compare tensors
    return arr


class CfgNode(dict):
    """The yaml of a checkpoint with attribute access (``cfg.single_channel``, ``cfg.conformer_css_cfg.nnet_conf.num_spks``),
    standing in for the reference's TrainCfg dataclass tree (css/training/train.py:48-91), which css/helpers.py:26 builds from
    the same file; only the keys present in the yaml exist (the training defaults are not part of the inference path)."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None


def load_css_model(model_dir: Path, device: Optional[torch.device] = None, **kw) -> (ConformerCssB200, "CfgNode"):
    """Counterpart of css/helpers.py:14-37: one ``*.pt`` (``checkpoint['model']`` with the DDP ``module.``
    prefix) and one ``*.yaml`` (the TrainCfg, returned with attribute access like the reference's dataclass) in
    ``model_dir``.  The network shape is taken from the checkpoint tensors themselves."""
    import yaml

    def fetch_one_file(path: Path, suffix: str):
        files = list(Path(path).glob(suffix))
        if len(files) == 0:
            raise FileNotFoundError(f'expecting at least one {suffix} file in {path}')
        assert len(files) == 1, f'expecting exactly one {suffix} file in {path}'
        return str(files[0])

    yaml_path = fetch_one_file(model_dir, '*.yaml')
    checkpoint_path = fetch_one_file(model_dir, '*.pt')
    with open(yaml_path) as f:
        train_cfg = CfgNode(yaml.safe_load(f) or {})
    checkpoint = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    state = {k[len("module."):]: v for k, v in checkpoint["model"].items() if k.startswith("module.")}
    return ConformerCssB200(state, device=device, **kw), train_cfg


def load_audio(wav_file_names: List, is_mc: bool) -> (np.ndarray, int):
    """css/helpers.py:40-65: 7 mono wav files (MC) or one (SC) -> [1, n_samples, n_channels] float32."""
    import scipy.io.wavfile as wf

    def read(fname):
        sr, data = wf.read(str(fname))
        if data.dtype == np.int16:
            data = data.astype(np.float32) / np.float32(32768.0)       # libsndfile's int16 -> float32 scaling
        elif data.dtype == np.int32:
            data = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            data = data.astype(np.float32)
        return data, sr

    if is_mc:
        assert len(wav_file_names) == 7, 'expecting 7 microphones'
        audio_data, srs = zip(*[read(f) for f in wav_file_names])
        mix_wav = np.stack(audio_data, axis=-1)[np.newaxis, ...]
        assert mix_wav.ndim == 3 and mix_wav.shape[2] in (1, 7)
        sr = srs[0]
    else:
        assert len(wav_file_names) == 1
        mix_wav, sr = read(wav_file_names[0])
        assert mix_wav.ndim == 1
        mix_wav = mix_wav[np.newaxis, :, np.newaxis]
    return mix_wav, sr


def write_wav(fname, samps: np.ndarray, sr: int = 16000, max_norm: bool = True):
    """utils/audio_utils.py:37-49: peak-normalise to 0.99 and write PCM_16 (libsndfile's float -> short
    conversion is lrintf(x * 0x7FFF))."""
    import os
    import scipy.io.wavfile as wf
    assert samps.ndim == 1
    if max_norm:
        samps = samps * 0.99 / (np.max(np.abs(samps)) + 1e-7)
    os.makedirs(os.path.dirname(str(fname)), exist_ok=True)
    pcm = np.clip(np.rint(samps.astype(np.float32) * np.float32(32767.0)), -32768, 32767).astype(np.int16)
    wf.write(str(fname), sr, pcm)


_MODEL_CACHE: Dict[str, ConformerCssB200] = {}
# CSS -> ASR / diarization hand-off without the disk round trip: the separated streams of the most recent sessions as the
# PCM16 the WAV files hold (peak-normalised to 0.99, utils/audio_utils.py:44-45), still in HBM.  Keyed by the absolute
# paths of the WAV files the same call wrote -- a different out_dir, or a later run that only hits the disk cache, can
# never pick up another run's samples -- with the sample rate beside the tensor.
DEVICE_STREAMS: Dict[tuple, tuple] = {}
_DEVICE_STREAMS_KEEP = 2


def _streams_key(wav_file_names) -> tuple:
    import os
    return tuple(os.path.abspath(str(f)) for f in wav_file_names)


def device_streams_for(wav_file_names):
    """(int16 CUDA tensor [n_streams, n], sample rate) of exactly these separated WAV files if this process produced them
    and they are still resident, else None.  The order follows ``wav_file_names``."""
    want = _streams_key(wav_file_names)
    for key, (pcm, sr) in DEVICE_STREAMS.items():
        if set(want) <= set(key) and len(want) > 0:
            idx = [key.index(w) for w in want]
            return (pcm if idx == list(range(len(key))) else pcm[idx].contiguous()), sr
    return None


def streams_to_pcm16(wav: "torch.Tensor") -> "torch.Tensor":
    """wav [S, N] float32 on the device -> int16 [S, N]: write_wav's 0.99 peak normalisation + libsndfile's PCM_16 rounding
    (nsf_peaknorm_pcm16), i.e. exactly the samples read_wav would read back from the files css_inference writes."""
    lib = _cabi.load()
    assert wav.is_cuda and wav.dtype == torch.float32 and wav.dim() == 2 and wav.is_contiguous()
    peak = torch.empty((wav.shape[0],), dtype=torch.float32, device=wav.device)
    pcm = torch.empty(wav.shape, dtype=torch.int16, device=wav.device)
    with torch.cuda.device(wav.device):
        _cabi.check(lib.nsf_peaknorm_pcm16(_cabi.ptr(wav), wav.shape[0], wav.shape[1], _cabi.ptr(peak), _cabi.ptr(pcm), _cabi.stream_ptr()),
                    "nsf_peaknorm_pcm16")
    return pcm


# ---- file boundary helpers of css_inference: threaded WAV I/O, PCM16 in and out of the device ------------------------------------
ASYNC_WAV_WRITES = False        # True: css_inference returns while its WAV files are still being written (flush_wav_writes waits)
_WAV_POOL = None
_WAV_PENDING: Dict[str, object] = {}


def _wav_pool():
    global _WAV_POOL
    if _WAV_POOL is None:
        import atexit
        from concurrent.futures import ThreadPoolExecutor
        _WAV_POOL = ThreadPoolExecutor(max_workers=8, thread_name_prefix="nsf-wav")
        atexit.register(flush_wav_writes)
    return _WAV_POOL


def _write_pcm16_file(fname: str, pcm: np.ndarray, sr: int):
    import os
    import scipy.io.wavfile as wf
    os.makedirs(os.path.dirname(fname), exist_ok=True)
    wf.write(fname, sr, pcm)


def flush_wav_writes(paths=None):
    """Waits for the WAV files css_inference handed to its writer threads (all of them, or the given paths); re-raises a
    writer's exception.  asr_inference / diarization_inference call it before they touch files; so does interpreter exit."""
    import os
    keys = list(_WAV_PENDING) if paths is None else [os.path.abspath(str(p)) for p in paths]
    for k in keys:
        fut = _WAV_PENDING.pop(k, None)
        if fut is not None:
            fut.result()


def _read_pcm16_channels(wav_file_names):
    """The channel files as one int16 array [C, n] (+ sample rate) when every file is mono PCM_16 of the same length and rate
    -- the NOTSOFAR recordings are -- else None (the caller then goes through load_audio).  Files are read concurrently."""
    import scipy.io.wavfile as wf

    def read(f):
        return wf.read(str(f), mmap=True)
    try:
        res = list(_wav_pool().map(read, wav_file_names))
    except Exception:
        return None
    srs = {r[0] for r in res}
    datas = [r[1] for r in res]
    if len(srs) != 1 or any(d.dtype != np.int16 or d.ndim != 1 or d.size != datas[0].size for d in datas):
        return None
    return datas, int(res[0][0])


def pcm16_to_device(datas, device) -> "torch.Tensor":
    """List of C int16 arrays [n] -> x [n, C] float32 on the device = pcm / 32768 (nsf_pcm16_to_float_interleaved)."""
    lib = _cabi.load()
    n, c = datas[0].size, len(datas)
    with torch.cuda.device(device):
        pcm = torch.empty((c, n), dtype=torch.int16, device=device)
        # the channel files (memory-mapped) are copied into one pooled page-locked buffer by the reader threads, every channel
        # goes up as soon as its copy is done: a pageable upload of a 6-minute session took 8 ms, this takes ~2
        stage = _pinned_out((c, n), torch.int16)
        view = _lend(stage)                                     # out of circulation until this call is over

        def fill(k):
            np.copyto(view[k], datas[k])
            return k
        for k in _wav_pool().map(fill, range(c)):
            pcm[k].copy_(stage[k], non_blocking=True)
        x = torch.empty((n, c), dtype=torch.float32, device=device)
        _cabi.check(lib.nsf_pcm16_to_float_interleaved(_cabi.ptr(pcm), c, n, _cabi.ptr(x), _cabi.stream_ptr()), "nsf_pcm16_to_float_interleaved")
        torch.cuda.current_stream(device).synchronize()        # the staging buffer goes back to the pool when `view` dies
        del view
    return x


def css_inference(out_dir: str, models_dir: str, session, cfg: CssCfg, fetch_from_cache: bool):
    """Applies CSS to one session -- same signature, file layout and cache semantics as the reference's
    css_inference (css/css.py:51-107).  Returns a copy of ``session`` with 'sep_wav_file_names'.

    The file boundary is kept off the critical path: 16-bit channel files are read concurrently and uploaded as int16 (the
    float conversion and the [N, C] interleave happen on the device), the separated streams are peak-normalised and quantised
    to PCM_16 on the device (nsf_peaknorm_pcm16: the very samples write_wav would produce), so that 2 bytes per sample cross
    PCIe instead of 4 and no host pass over the waveforms remains, and the four WAV files are written by parallel threads
    (by default the call still returns only once they are on disk, like the reference)."""
    session_css = session.copy()

    assert isinstance(session.wav_file_names, list)
    if cfg.pass_through_ch0:
        session_css['sep_wav_file_names'] = session.wav_file_names[0:1]
        return session_css

    css_out_dir = Path(out_dir) / "css_inference" / session.session_id
    if fetch_from_cache and css_out_dir.exists():
        flush_wav_writes()
        sep_wav_file_names = sorted(css_out_dir.glob('sep*.wav'))
        session_css['sep_wav_file_names'] = sep_wav_file_names
        # the files on disk are the truth for this run: drop any in-HBM copy an earlier call left for the same paths
        DEVICE_STREAMS.pop(_streams_key(sep_wav_file_names), None)
        return session_css

    if not torch.cuda.is_available():
        raise _cabi.NsfError("notsofar_b200.css_inference needs a CUDA device (sm_100a); there is no CPU path")
    device = torch.device('cuda', torch.cuda.current_device())
    model_dir = str(Path(models_dir) / (cfg.checkpoint_mc if session.is_mc else cfg.checkpoint_sc))
    # the reference re-loads the checkpoint for every session (css.py:85); the weights are kept resident here
    if model_dir not in _MODEL_CACHE:
        _MODEL_CACHE[model_dir] = load_css_model(Path(model_dir), device=device)[0]
    separator = _MODEL_CACHE[model_dir]
    separator.eval()
    separator.to(device)

    n_expected = 7 if session.is_mc else 1
    assert len(session.wav_file_names) == n_expected, 'expecting 7 microphones' if session.is_mc else 'expecting one file'
    import os, time
    timing = os.environ.get("NSF_TIMING") == "1"          # per-phase wall clock of the call (forces a device sync per phase)
    marks = [("start", time.perf_counter())]

    def tick(name):
        if timing:
            torch.cuda.synchronize(device)
            marks.append((name, time.perf_counter()))
    fast = _read_pcm16_channels(session.wav_file_names)
    tick("read")
    if fast is not None:
        datas, sr = fast
        x = pcm16_to_device(datas, device)
    else:                                                       # float / 24-bit / ragged files: the reference's host path
        mixwav, sr = load_audio(session.wav_file_names, is_mc=session.is_mc)
        x = torch.from_numpy(np.ascontiguousarray(mixwav[0], dtype=np.float32)).to(device)
    if cfg.slice_audio_for_debug:
        x = x[sr * 20:sr * 30].contiguous()

    tick("upload")
    out = css_device(x, separator, sr, cfg)
    tick("css_device")
    # write_wav (utils/audio_utils.py:37-49) on the device: 0.99 peak normalisation + PCM_16 rounding of the three streams and
    # of channel 0 of the input (input_mixture.wav); the int16 streams also stay in HBM for the stages downstream
    pcm_dev = streams_to_pcm16(out["wav"])
    mix_pcm = streams_to_pcm16(x[:, 0].contiguous()[None])
    del out
    # one pooled page-locked buffer for the four files' samples; the writer threads hold views of it (it is not handed out again
    # before they are done)
    with torch.cuda.device(device):
        host = _pinned_out(tuple(pcm_dev.shape), torch.int16)
        pcm_host = _lend(host)                                        # lent before the next request: never the same buffer twice
        host_mix = _pinned_out((mix_pcm.shape[1],), torch.int16)      # the streams are a few samples shorter than the input
        mix_host = _lend(host_mix)
        host.copy_(pcm_dev, non_blocking=True)
        host_mix.copy_(mix_pcm[0], non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
    tick("pcm16+d2h")

    names = [str(css_out_dir / 'input_mixture.wav')] + [str(css_out_dir / f"sep_stream{i}.wav") for i in range(pcm_host.shape[0])]
    for fname, samps in zip(names, [mix_host] + [pcm_host[i] for i in range(pcm_host.shape[0])]):
        flush_wav_writes([fname])                               # an earlier write of the same path must not overtake this one
        _WAV_PENDING[os.path.abspath(fname)] = _wav_pool().submit(_write_pcm16_file, fname, samps, int(sr))
    if not ASYNC_WAV_WRITES:
        flush_wav_writes(names)

    tick("wav_submit/flush")
    if timing:
        print("css_inference phases [ms]: " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.2f}" for a, b in zip(marks, marks[1:])), flush=True)
    sep_wav_file_names = names[1:]
    DEVICE_STREAMS.pop(_streams_key(sep_wav_file_names), None)
    DEVICE_STREAMS[_streams_key(sep_wav_file_names)] = (pcm_dev, int(sr))
    while len(DEVICE_STREAMS) > _DEVICE_STREAMS_KEEP:
        DEVICE_STREAMS.pop(next(iter(DEVICE_STREAMS)))
    session_css['sep_wav_file_names'] = sep_wav_file_names
    return session_css
