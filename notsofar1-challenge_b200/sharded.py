"""One meeting across the GPUs of a box: contiguous blocks of segments per rank, three small exchanges.

The reference separates a session on one device, segment after segment (css/css.py:182-250).  Loop I has no
cross-segment dependency; only the permutation chain (css.py:266-285), the 50 %-overlap weighted overlap-add
(:287-299) and the activity morphology (:303-312) couple neighbours.  Rank r therefore owns the segments
[lo_r, hi_r) plus a one-segment halo on the left (recomputed, not communicated), and the ranks meet three times:

  1. all-gather of the 3x3 stitching costs of the owned segments (36 B / segment) -> every rank replays the
     6-permutation chain on the host and knows the global channel order of its block;
  2. all-gather of the per-frame mask means (12 B / frame) -> every rank runs the (global) dilate / erode gate;
  3. the separated waveforms of the owned frames go to the rank that hands the streams to ASR / diarization
     (3 x 4 B / sample; 345 MB per 30 min): point-to-point, straight into their final place; only the 256-sample
     seams between ranks are added there (gather_waveforms).

Everything else (STFT of the rank's sample range, features, mask network, MVDR, local WOLA, iSTFT) is the
single-GPU path of css.py on the rank's slice.  When a rank also reads its own samples back to the host
(phase1(host_piece) / finish_host), the interior of its piece leaves while its segments are still in the mask network:
a local permutation chain from the identity, the progressive tail of css_device on the local arrays, and a relabelling
of the three streams once the global chain is known (the assignment is equivariant under a relabelling of the previous
segment; checked, with a fall-back to copying the whole piece).  torch.distributed (NCCL over NVLink on the GPUs, gloo in the
CPU tests of the exchange logic) carries the three exchanges; there is no CPU compute path.
"""
from __future__ import annotations

import sys
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _cabi
import ctypes
import os

from . import css as _css
from .css import (CssCfg, SegmentPlan, plan_segments, plan_batches, permutation_chain, _segment_weights, HostFeeder)
from .separator import NUM_BINS, FRAME_HOP, FRAME_LEN


@dataclass
class Shard:
    """Segment / frame / sample ranges of one rank (all global indices; *_hi exclusive)."""
    rank: int
    world: int
    seg_lo: int
    seg_hi: int
    halo: int            # segments left of seg_lo recomputed locally: every segment that reaches into the owned frames,
                         # min(seg_lo, ceil(T / hop) - 1) -- one for the default 50 % overlap, two for hop = T / 3, ...
    frame0: int          # global index of local frame 0
    n_frames: int        # local frames (pitch of the local X / stitched arrays)
    valid_frames: int    # local frames that exist in the signal (the rest is the zero padding of css.py:159-164,185-190)
    own_lo: int          # owned output frames [own_lo, own_hi)
    own_hi: int
    sample_lo: int       # samples of the recording this rank needs: [sample_lo, sample_hi)
    sample_hi: int

    @property
    def n_own_seg(self) -> int:
        return self.seg_hi - self.seg_lo

    @property
    def n_loc_seg(self) -> int:
        return self.seg_hi - self.seg_lo + self.halo if self.seg_hi > self.seg_lo else 0

    @property
    def n_own_frames(self) -> int:
        return self.own_hi - self.own_lo


def shard_bounds(n_seg: int, world: int) -> List[int]:
    """Balanced contiguous blocks: rank r owns segments [b[r], b[r+1])."""
    base, extra = divmod(n_seg, world)
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < extra else 0))
    return b


def make_shard(plan: SegmentPlan, rank: int, world: int) -> Shard:
    T, hop = plan.segment_frames, plan.hop_frames
    b = shard_bounds(plan.num_segments, world)
    lo, hi = b[rank], b[rank + 1]
    if hi == lo:           # more ranks than segments: nothing to do, nothing owned
        return Shard(rank, world, lo, hi, 0, 0, 0, 0, 0, 0, 0, 0)
    # frame lo*hop is covered by the segments s with s*hop <= lo*hop < s*hop + T, i.e. s > lo - T/hop: all of them must be
    # local, or the owned frames next to the seam miss a term of the overlap-add while still being divided by the global
    # weight sum (hop_size_sec=1.0 has three segments per frame).  At least one, for the stitching cost of seg_lo.
    halo = min(lo, max(1, -(-T // hop) - 1))
    frame0 = (lo - halo) * hop
    n_frames = min(plan.mix_frames, (hi - 1) * hop + T) - frame0
    valid = max(0, min(plan.raw_frames - frame0, n_frames))
    own_lo = lo * hop
    own_hi = hi * hop if hi < plan.num_segments else plan.mix_frames
    s_lo = frame0 * FRAME_HOP
    s_hi = (frame0 + valid - 1) * FRAME_HOP + FRAME_LEN if valid > 0 else s_lo
    return Shard(rank, world, lo, hi, halo, frame0, n_frames, valid, own_lo, own_hi, s_lo, s_hi)


# ------------------------------------------------------------------------------------------- exchanges
def allgather_varlen(own: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """All-gather along dim 0 of per-rank tensors with counts[r] rows (known to every rank from the plan)."""
    import torch.distributed as dist
    world = len(counts)
    assert own.shape[0] == counts[dist.get_rank(group)]
    mx = max(max(counts), 1)
    pad = torch.zeros((mx,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
    pad[:own.shape[0]] = own
    out = torch.empty((world * mx,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.view((world, mx) + tuple(own.shape[1:]))
    return torch.cat([out[r, :counts[r]] for r in range(world)], dim=0)


def gather_varlen(own: torch.Tensor, counts: Sequence[int], dst: int = 0, group=None, dim: int = 0) -> Optional[List[torch.Tensor]]:
    """Gather per-rank tensors whose size along ``dim`` is counts[r] to rank ``dst``; returns the list there, None elsewhere."""
    import torch.distributed as dist
    world = len(counts)
    rank = dist.get_rank(group)
    assert own.shape[dim] == counts[rank]
    mx = max(max(counts), 1)
    shape = list(own.shape)
    shape[dim] = mx
    pad = torch.zeros(shape, dtype=own.dtype, device=own.device)
    pad.narrow(dim, 0, own.shape[dim]).copy_(own)
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return [bufs[r].narrow(dim, 0, counts[r]) for r in range(world)]


def assemble_waveforms(pieces: Sequence[torch.Tensor], shards: Sequence[Shard], mix_frames: int) -> torch.Tensor:
    """Overlap-adds the per-rank waveform pieces [S, n_own_frames*256 + 256] (piece r starts at sample own_lo*256;
    its last 256 samples are the tail of its last frame, which lands on the head of rank r+1's first frame)."""
    n_out = (mix_frames - 1) * FRAME_HOP + FRAME_LEN
    S = pieces[0].shape[0]
    out = torch.zeros((S, n_out), dtype=pieces[0].dtype, device=pieces[0].device)
    for p, sh in zip(pieces, shards):
        if sh.n_own_frames == 0:
            continue
        n = sh.n_own_frames * FRAME_HOP + FRAME_HOP
        assert p.shape[1] == n
        out[:, sh.own_lo * FRAME_HOP: sh.own_lo * FRAME_HOP + n] += p
    return out


def gather_waveforms(piece: torch.Tensor, shards: Sequence[Shard], mix_frames: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """The waveform hand-off without padding, zero-fill or full-size adds: rank ``dst`` receives the body of every piece
    (its first n_own_frames*256 samples) straight into its final place in the [S, N'] output -- the bodies tile the time axis
    exactly -- and only the 256-sample tails (the second half of each rank's last frame, which overlaps the next rank's first
    256 samples) are added on top.  One point-to-point message per (rank, stream) plus one small one per rank."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    world = len(shards)
    S = piece.shape[0]
    sh = shards[rank]
    n_out = (mix_frames - 1) * FRAME_HOP + FRAME_LEN
    body = sh.n_own_frames * FRAME_HOP
    assert piece.shape[1] == (body + FRAME_HOP if sh.n_own_frames else 0)
    ops, tails = [], {}
    if rank == dst:
        out = torch.empty((S, n_out), dtype=piece.dtype, device=piece.device)
        for r, s_r in enumerate(shards):
            if s_r.n_own_frames == 0:
                continue
            lo, n = s_r.own_lo * FRAME_HOP, s_r.n_own_frames * FRAME_HOP
            if r == rank:
                out[:, lo:lo + n].copy_(piece[:, :n])
                tails[r] = piece[:, n:]
            else:
                for k in range(S):
                    ops.append(dist.P2POp(dist.irecv, out[k, lo:lo + n], _global_rank(r, group), group))
                tails[r] = torch.empty((S, FRAME_HOP), dtype=piece.dtype, device=piece.device)
                ops.append(dist.P2POp(dist.irecv, tails[r], _global_rank(r, group), group))
    elif sh.n_own_frames:
        tail = piece[:, body:].contiguous()
        for k in range(S):
            ops.append(dist.P2POp(dist.isend, piece[k, :body], _global_rank(dst, group), group))
        ops.append(dist.P2POp(dist.isend, tail, _global_rank(dst, group), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank != dst:
        return None
    last = max((r for r, s_r in enumerate(shards) if s_r.n_own_frames), default=None)
    for r, t in tails.items():
        at = shards[r].own_hi * FRAME_HOP
        if r == last:
            out[:, at:at + FRAME_HOP].copy_(t)                        # the very end of the recording: nobody else writes it
        else:
            out[:, at:at + FRAME_HOP] += t
    return out


def _global_rank(r: int, group) -> int:
    import torch.distributed as dist
    return r if group is None else dist.get_global_rank(group, r)


# ------------------------------------------------------------------------------------------- per-rank work
_SHARD_W_DEV: Dict[tuple, tuple] = {}


class ShardWorker:
    """The work of one rank, split at the three exchange points so that the same code runs under torch.distributed
    (css_device_sharded) and, rank after rank on one device, in the single-GPU parity test of the sharding logic."""

    def __init__(self, separator, fs: int, cfg: CssCfg, n_samples_total: int, rank: int, world: int):
        self.sep, self.fs, self.cfg = separator, fs, cfg
        self.plan = plan_segments(n_samples_total, fs, cfg)
        self.world = world
        self.shards = [make_shard(self.plan, r, world) for r in range(world)]
        self.sh = self.shards[rank]
        self.lib = _cabi.load()
        assert cfg.stitching_loss in ('l1', 'mse') and cfg.stitching_input in ('mask', 'separation_result')

    # ---- phase 1: STFT, mask network, MVDR and stitching costs of the local block --------------------------------
    @torch.no_grad()
    def phase1(self, x_local, host_piece: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x_local: samples [sample_lo, sample_hi) of the recording, [n, C] float32 on the device (or a HostFeeder
        over that slice).  Returns the costs of the owned segments [n_own_seg, S, S] (device).

        host_piece: optional page-locked [S, n_own_frames*256 + 256] float32 tensor that is to receive this rank's
        waveform piece (finish_host).  With it the *interior* of the piece crosses PCIe while the mask network is still
        running: a rank cannot know the global channel order of its block before every earlier rank has finished, but that
        order is only a relabelling of its three streams -- the optimal assignment is equivariant, p*(q) = sigma o q --, so
        the rank walks a local chain from the identity, runs the progressive tail of css_device on its local arrays
        (nsf_stitch_progress) and copies out, still under the local labels, the hops that depend neither on a
        neighbour's activity (dilation + erosion frames next to a seam) nor on the last chunk of segments."""
        sh, plan, cfg, sep = self.sh, self.plan, self.cfg, self.sep
        feeder = x_local if isinstance(x_local, HostFeeder) else None
        x = feeder.x_dev if feeder is not None else x_local
        device = x.device
        self.device = device
        T, hop, S = plan.segment_frames, plan.hop_frames, cfg.num_spks
        n_masks = sep.num_masks
        n_loc = sh.n_loc_seg
        if n_loc == 0:
            self.masks = self.Y = None
            return torch.zeros((0, S, S), dtype=torch.float32, device=device)
        assert x.shape[0] == sh.sample_hi - sh.sample_lo, (x.shape, sh)
        num_channels = x.shape[1]
        use_mvdr = num_channels > 1 and cfg.mc_mvdr
        mask_floor = 10. ** ((cfg.mc_mask_floor_db if num_channels > 1 else cfg.sc_mask_floor_db) / 20.)
        with torch.cuda.device(device):
            X = sep.stft_alloc(num_channels, sh.n_frames, sh.valid_frames, device)
            self.masks = torch.empty((n_loc, n_masks, NUM_BINS, T), dtype=torch.float32, device=device)
            self.Y = torch.empty((n_loc, S, NUM_BINS, T), dtype=torch.complex64, device=device)
            frames_done = 0
            # NSF_SHARD_LAST_CHUNK=111 cuts a short chunk off the end (less to copy after the last kernel where several ranks
            # share the host's memory bandwidth); measured at N = 4: 125.6 vs 124.8 ms without -- off by default
            chunks = plan_batches(n_loc, int(sep.segments_per_batch), streaming=feeder is not None,
                                  progressive=_css.PROGRESSIVE_CHUNK if host_piece is not None else 0,
                                  last_batch=int(os.environ.get("NSF_SHARD_LAST_CHUNK", "0")) if self.world > 1 else 0)
            costs = torch.empty((n_loc, S, S), dtype=torch.float32, device=device)
            in_kind = 0 if cfg.stitching_input == 'mask' else 1
            loss_kind = 0 if cfg.stitching_loss == 'l1' else 1
            src = self.masks if in_kind == 0 else self.Y
            self.prog = None
            if host_piece is not None and len(chunks) > 1 and _css.PROGRESSIVE_CHUNK > 0:
                self._progressive_setup(host_piece, chunks, device)
            for ci, (s0, nb) in enumerate(chunks):
                f_need = min(sh.valid_frames, (s0 + nb - 1) * hop + T)
                if f_need > frames_done:
                    if feeder is not None:
                        feeder.ready((f_need - 1) * FRAME_HOP + FRAME_LEN)
                    sep.stft_frames(x, X, frames_done, f_need)
                    frames_done = f_need
                sep.masks(X, sh.valid_frames, s0, nb, T, hop, out=self.masks[s0:s0 + nb])
                if use_mvdr:
                    sep.mvdr(self.masks[s0:s0 + nb], X, sh.valid_frames, s0, hop, mask_floor, out=self.Y[s0:s0 + nb])
                else:
                    sep.mask_apply(self.masks[s0:s0 + nb], X, sh.valid_frames, s0, hop, mask_floor, out=self.Y[s0:s0 + nb])
                if cfg.normalize_segment_power:
                    # the reference's t = en - st counts frames up to mix_frames (global): local pitch ends there too
                    sep.power_norm(self.Y[s0:s0 + nb], X, sh.valid_frames, s0, hop, plan.mix_frames - sh.frame0)
                if self.prog is not None:
                    _cabi.check(self.lib.nsf_pit_cost_range(_cabi.ptr(src), in_kind, loss_kind, s0, s0 + nb, n_masks if in_kind == 0 else S,
                                                            S, NUM_BINS, T, plan.overlap_frames, _cabi.ptr(costs), _cabi.stream_ptr()),
                                "nsf_pit_cost_range")
                    self._progressive_chunk_done(ci, costs)
            if feeder is not None:
                feeder.ready(x.shape[0])
            if self.prog is None:
                _cabi.check(self.lib.nsf_pit_cost(_cabi.ptr(src), in_kind, loss_kind, n_loc, n_masks if in_kind == 0 else S, S, NUM_BINS,
                                                  T, plan.overlap_frames, _cabi.ptr(costs), _cabi.stream_ptr()), "nsf_pit_cost")
            elif len(chunks) >= 2:
                self._progressive_advance(len(chunks) - 2)      # the last chunk's frames wait for the global tail (phase 3)
        self.X = X
        return costs[sh.halo:]

    # ---- progressive read-back of the piece's interior (see phase1) ------------------------------------------------
    def _progressive_setup(self, host_piece: torch.Tensor, chunks, device):
        sh, plan, cfg = self.sh, self.plan, self.cfg
        S, T = cfg.num_spks, plan.segment_frames
        assert tuple(host_piece.shape) == (S, sh.n_own_frames * FRAME_HOP + FRAME_HOP) and host_piece.dtype == torch.float32
        loc0 = sh.seg_lo - sh.halo
        key = (plan.segment_frames, plan.hop_frames, plan.m0_frames, plan.m1_frames, plan.num_segments, plan.mix_frames, loc0, sh.seg_hi,
               sh.frame0, sh.n_frames, str(device))
        hit = _SHARD_W_DEV.get(key)
        if hit is None:            # small uploads, kept: next to the recording they would queue behind it on the copy engine
            seg_w_np, wsum_np = _segment_weights(plan)
            hit = (torch.from_numpy(np.ascontiguousarray(seg_w_np[loc0:sh.seg_hi])).to(device),
                   torch.from_numpy(np.ascontiguousarray(wsum_np[sh.frame0:sh.frame0 + sh.n_frames])).to(device))
            _SHARD_W_DEV[key] = hit
            if len(_SHARD_W_DEV) > 8:
                _SHARD_W_DEV.pop(next(iter(_SHARD_W_DEV)))
        nf = sh.n_frames
        p = dict(chunks=chunks, host=host_piece, seg_w=hit[0], wsum=hit[1], events=[], state=None, copied=[],
                 mask_st=torch.empty((NUM_BINS, nf, S), dtype=torch.float32, device=device),
                 activity=torch.empty((nf, S), dtype=torch.float32, device=device),
                 act_b=torch.empty((nf, S), dtype=torch.uint8, device=device),
                 act_tmp=torch.empty((nf, S), dtype=torch.uint8, device=device),
                 act_final=torch.empty((nf, S), dtype=torch.uint8, device=device),
                 S_st=torch.empty((S, nf, NUM_BINS), dtype=torch.complex64, device=device),
                 wav=torch.empty((S, (nf - 1) * FRAME_HOP + FRAME_LEN), dtype=torch.float32, device=device),
                 perms=torch.empty((sh.n_loc_seg, S), dtype=torch.int32, device=device),
                 costs_host=_css._small_pinned(f"shard_costs{sh.rank}", (sh.n_loc_seg, S, S), torch.float32),
                 perms_host=_css._small_pinned(f"shard_perms{sh.rank}", (sh.n_loc_seg, S), torch.int32), perms_np=[],
                 tail=_css._tail_stream(device), main=torch.cuda.current_stream(device), hops=(ctypes.c_int64 * 2)())
        # hops of the local waveform that may leave early: complete frames on both sides (hop j reads frames j-1 and j),
        # gate decided without a neighbour's activity
        R = plan.dilation_frames + plan.erosion_frames
        a, b = sh.own_lo - sh.frame0, sh.own_hi - sh.frame0
        p["a"] = a
        p["hop_lo"] = a + R + 1 if sh.seg_lo > 0 else 0
        p["hop_hi"] = b - R if sh.seg_hi < plan.num_segments else b + 1
        p["tail"].wait_stream(p["main"])
        self.prog = p

    def _progressive_chunk_done(self, ci: int, costs: torch.Tensor):
        p = self.prog
        c0, cn = p["chunks"][ci]
        p["costs_host"][c0:c0 + cn].copy_(costs[c0:c0 + cn], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(p["main"])
        p["events"].append(ev)
        if ci >= 1:
            self._progressive_advance(ci - 1)          # one chunk behind: the main stream's queue never runs dry

    def _progressive_advance(self, ci: int):
        p, sh, plan, cfg = self.prog, self.sh, self.plan, self.cfg
        if ci < 0 or ci >= len(p["chunks"]) - 1 or ci < len(p["copied"]):
            return
        S = cfg.num_spks
        c0, cn = p["chunks"][ci]
        p["events"][ci].synchronize()
        p_np, p["state"] = permutation_chain(p["costs_host"][c0:c0 + cn].numpy(), prev_state=p["state"], return_state=True)
        p["perms_host"][c0:c0 + cn] = torch.from_numpy(p_np)
        p["perms_np"].append(p_np)
        with torch.cuda.stream(p["tail"]):
            p["tail"].wait_event(p["events"][ci])
            p["perms"][c0:c0 + cn].copy_(p["perms_host"][c0:c0 + cn], non_blocking=True)
            _cabi.check(self.lib.nsf_stitch_progress(
                _cabi.ptr(self.masks), self.sep.num_masks, _cabi.ptr(self.Y), _cabi.ptr(p["perms"]), _cabi.ptr(p["seg_w"]), _cabi.ptr(p["wsum"]),
                sh.n_loc_seg, c0, c0 + cn, S, NUM_BINS, plan.segment_frames, plan.hop_frames, sh.n_frames, float(np.float32(cfg.activity_th)),
                plan.dilation_frames, plan.erosion_frames, _cabi.ptr(p["mask_st"]), _cabi.ptr(p["activity"]), _cabi.ptr(p["act_b"]),
                _cabi.ptr(p["act_tmp"]), _cabi.ptr(p["act_final"]), _cabi.ptr(p["S_st"]), _cabi.ptr(p["wav"]), p["hops"], _cabi.stream_ptr()),
                "nsf_stitch_progress")
            h0, h1 = max(int(p["hops"][0]), p["hop_lo"]), min(int(p["hops"][1]), p["hop_hi"])
            if h1 > h0:
                lo, hi = (h0 - p["a"]) * FRAME_HOP, (h1 - p["a"]) * FRAME_HOP          # samples of the piece
                for k in range(S):
                    p["host"][k, lo:hi].copy_(p["wav"][k, h0 * FRAME_HOP:h1 * FRAME_HOP], non_blocking=True)
                p["copied"].append((lo, hi))
            else:
                p["copied"].append((0, 0))

    @torch.no_grad()
    def finish_host(self, wav_piece: torch.Tensor, host_piece: torch.Tensor) -> List[torch.Tensor]:
        """The rank's waveform piece on the host: rows in the global stream order (views of host_piece; synchronise the
        current stream before reading).  What left early under the local labels is kept if the local chain, relabelled by
        the channel order at the rank's first local segment, is the global chain on those segments (it is, away from exact
        ties between assignments); the rest -- seams, last chunk -- is copied now.  Otherwise the whole piece is copied."""
        S = self.cfg.num_spks
        p = getattr(self, "prog", None)
        n = wav_piece.shape[1]
        if p is None or n == 0:
            host_piece.copy_(wav_piece, non_blocking=True)
            return [host_piece[k] for k in range(S)]
        p["main"].wait_stream(p["tail"])
        sh = self.sh
        loc0 = sh.seg_lo - sh.halo
        local = np.concatenate(p["perms_np"], axis=0)                          # local segments the local chain has walked
        n_adv = local.shape[0]
        tau = np.asarray(self.perms[loc0], dtype=np.int64)                       # global slot k == local slot tau[k]
        self.relabel = tau
        self.progressive_ok = bool(np.array_equal(self.perms[loc0:loc0 + n_adv], local[:, tau]))
        if not self.progressive_ok:
            host_piece.copy_(wav_piece, non_blocking=True)
            return [host_piece[k] for k in range(S)]
        done = sorted((lo, hi) for lo, hi in p["copied"] if hi > lo)
        at = 0
        for lo, hi in done + [(n, n)]:
            if lo > at:
                for k in range(S):
                    host_piece[int(tau[k]), at:lo].copy_(wav_piece[k, at:lo], non_blocking=True)
            at = max(at, hi)
        return [host_piece[int(tau[k])] for k in range(S)]

    # ---- phase 2: permutation chain (host, replicated), local mask WOLA -> owned rows of the activity mean --------
    @torch.no_grad()
    def phase2(self, costs_all: np.ndarray) -> torch.Tensor:
        sh, plan, cfg = self.sh, self.plan, self.cfg
        S = cfg.num_spks
        assert costs_all.shape == (plan.num_segments, S, S)
        self.perms = permutation_chain(costs_all)
        device = self.device
        if sh.n_loc_seg == 0:
            return torch.zeros((0, S), dtype=torch.float32, device=device)
        seg_w_np, wsum_np = _segment_weights(plan)
        assert (wsum_np > 1e-5).all(), 'zero weights found. check hop_size, segment_size or m0, m1'
        loc0 = sh.seg_lo - sh.halo
        with torch.cuda.device(device):
            self.perms_loc = torch.from_numpy(np.ascontiguousarray(self.perms[loc0:sh.seg_hi])).to(device)
            self.seg_w = torch.from_numpy(np.ascontiguousarray(seg_w_np[loc0:sh.seg_hi])).to(device)
            self.wsum = torch.from_numpy(np.ascontiguousarray(wsum_np[sh.frame0:sh.frame0 + sh.n_frames])).to(device)
            self.mask_st = torch.empty((NUM_BINS, sh.n_frames, S), dtype=torch.float32, device=device)
            activity = torch.empty((sh.n_frames, S), dtype=torch.float32, device=device)
            _cabi.check(self.lib.nsf_stitch_masks(_cabi.ptr(self.masks), self.sep.num_masks, _cabi.ptr(self.perms_loc), _cabi.ptr(self.seg_w),
                                                  _cabi.ptr(self.wsum), sh.n_loc_seg, S, NUM_BINS, plan.segment_frames, plan.hop_frames,
                                                  sh.n_frames, _cabi.ptr(self.mask_st), _cabi.ptr(activity), _cabi.stream_ptr()),
                        "nsf_stitch_masks")
        return activity[sh.own_lo - sh.frame0: sh.own_hi - sh.frame0]

    # ---- phase 3: global activity gate (replicated, tiny), local STFT WOLA + iSTFT -> waveform of the owned frames --
    @torch.no_grad()
    def phase3(self, activity_all: torch.Tensor) -> Dict[str, torch.Tensor]:
        sh, plan, cfg = self.sh, self.plan, self.cfg
        S = cfg.num_spks
        device = self.device
        mix = plan.mix_frames
        assert tuple(activity_all.shape) == (mix, S)
        with torch.cuda.device(device):
            activity_all = activity_all.contiguous()
            act_b = torch.empty((mix, S), dtype=torch.uint8, device=device)
            act_tmp = torch.empty_like(act_b)
            act_final = torch.empty_like(act_b)
            _cabi.check(self.lib.nsf_activity(_cabi.ptr(activity_all), mix, S, float(np.float32(cfg.activity_th)), plan.dilation_frames,
                                              plan.erosion_frames, _cabi.ptr(act_b), _cabi.ptr(act_tmp), _cabi.ptr(act_final),
                                              _cabi.stream_ptr()), "nsf_activity")
            out = dict(activity=activity_all, activity_b=act_b, activity_final=act_final)
            if sh.n_loc_seg == 0:
                out["wav_piece"] = torch.zeros((S, 0), dtype=torch.float32, device=device)
                out["mask_piece"] = torch.zeros((NUM_BINS, 0, S), dtype=torch.float32, device=device)
                return out
            act_loc = act_final[sh.frame0: sh.frame0 + sh.n_frames]            # contiguous rows
            S_st = torch.empty((S, sh.n_frames, NUM_BINS), dtype=torch.complex64, device=device)
            _cabi.check(self.lib.nsf_stitch_stft(_cabi.ptr(self.Y), _cabi.ptr(self.perms_loc), _cabi.ptr(self.seg_w), _cabi.ptr(self.wsum),
                                                 _cabi.ptr(act_loc), sh.n_loc_seg, S, NUM_BINS, plan.segment_frames, plan.hop_frames,
                                                 sh.n_frames, _cabi.ptr(S_st), _cabi.stream_ptr()), "nsf_stitch_stft")
            # frames outside the owned range are incomplete here (their other segment lives on a neighbour): they
            # contribute nothing to this rank's piece
            a, b = sh.own_lo - sh.frame0, sh.own_hi - sh.frame0
            if a > 0:
                S_st[:, :a].zero_()
            if b < sh.n_frames:
                S_st[:, b:].zero_()
            wav = self.sep.istft_device(S_st)                                   # [S, (n_frames-1)*256+512]
            out["wav_piece"] = wav[:, a * FRAME_HOP: b * FRAME_HOP + FRAME_HOP].contiguous()
            out["mask_piece"] = self.mask_st[:, a:b]
        return out


@torch.no_grad()
def css_device_sharded(x_local, separator, fs: int, cfg: CssCfg, n_samples_total: int, group=None, dst: int = 0,
                       want_side_info: bool = False, host_piece: Optional[torch.Tensor] = None) -> Dict:
    """One meeting sharded over the ranks of ``group`` (default: the world).  x_local is this rank's sample range
    (``make_shard(plan, rank, world).sample_lo/hi``), on the device or behind a HostFeeder.  Returns on rank ``dst``
    {'wav' [S, N'], 'activity_b', 'activity_final', 'perms', 'plan' (+ 'mask_stitched' if want_side_info)}; on the
    other ranks the dict has no 'wav'.  host_piece (page-locked [S, n_own_frames*256 + 256]): this rank's own samples are
    also read back to the host, their interior while the mask network is still running (ShardWorker.phase1); the dict then
    holds 'wav_host', the rows in the global stream order (synchronise the current stream before reading them)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    wk = ShardWorker(separator, fs, cfg, n_samples_total, rank, world)
    S = cfg.num_spks
    import os
    import time
    timing = os.environ.get("NSF_SHARD_TIMING") == "1"

    def mark(label, _t=[None]):
        if timing:
            torch.cuda.synchronize()
            now = time.perf_counter()
            if _t[0] is not None and rank == dst:
                print(f"[shard timing] {label}: {(now - _t[0]) * 1e3:.2f} ms", file=sys.stderr, flush=True)
            _t[0] = now
    mark("start")
    own_costs = wk.phase1(x_local, host_piece)
    mark("phase1 (segments)")
    costs_all = allgather_varlen(own_costs, [s.n_own_seg for s in wk.shards], group)
    costs_np = costs_all.cpu().numpy()
    mark("all-gather costs + read-back")
    own_act = wk.phase2(costs_np)
    mark("phase2 (chain + mask WOLA)")
    activity_all = allgather_varlen(own_act, [s.n_own_frames for s in wk.shards], group)
    mark("all-gather activity")
    out = wk.phase3(activity_all)
    mark("phase3 (gate + STFT WOLA + iSTFT)")
    wav = gather_waveforms(out["wav_piece"], wk.shards, wk.plan.mix_frames, dst, group)
    mark("gather waveforms")
    res = dict(activity_b=out["activity_b"], activity_final=out["activity_final"], perms=wk.perms, plan=wk.plan, shard=wk.sh,
               wav_piece=out["wav_piece"])        # this rank's own samples [S, n_own_frames*256 + 256], first sample own_lo*256
    if host_piece is not None:
        res["wav_host"] = wk.finish_host(out["wav_piece"], host_piece)
    mask_pieces = None
    if want_side_info:
        mask_pieces = gather_varlen(out["mask_piece"].contiguous(), [s.n_own_frames for s in wk.shards], dst, group, dim=1)
    if rank == dst:
        res["wav"] = wav
        if mask_pieces is not None:
            res["mask_stitched"] = torch.cat(mask_pieces, dim=1)
    return res


@torch.no_grad()
def css_sharded_on_one_device(x: torch.Tensor, separator, fs: int, cfg: CssCfg, world: int) -> Dict:
    """The sharded algorithm with all ``world`` ranks played by one device, one after the other (exchanges become
    concatenations).  Used to check the sharding logic against css_device on a single GPU."""
    n = x.shape[0]
    wks = [ShardWorker(separator, fs, cfg, n, r, world) for r in range(world)]
    costs = [w.phase1(x[w.sh.sample_lo:w.sh.sample_hi].contiguous()) for w in wks]
    costs_all = torch.cat(costs, 0).cpu().numpy()
    acts = [w.phase2(costs_all) for w in wks]
    activity_all = torch.cat(acts, 0)
    outs = [w.phase3(activity_all) for w in wks]
    wav = assemble_waveforms([o["wav_piece"] for o in outs], [w.sh for w in wks], wks[0].plan.mix_frames)
    return dict(wav=wav, mask_stitched=torch.cat([o["mask_piece"] for o in outs], 1), activity=activity_all,
                activity_b=outs[0]["activity_b"], activity_final=outs[0]["activity_final"], perms=wks[0].perms,
                plan=wks[0].plan, masks=[w.masks for w in wks], Y=[w.Y for w in wks], shards=[w.sh for w in wks])
