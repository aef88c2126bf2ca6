#!/usr/bin/env python
"""bench.py -- audio-seconds per second (xRT) of the CSS + MVDR hot path on a synthetic 7-channel 16 kHz meeting.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--seconds 1800]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): continuous speech separation of a 30-minute 7-channel meeting -- multichannel
STFT, 1 209 overlapping 3-s segments through the v1.0-MC Conformer mask network (d=512, 8 heads, 18 blocks, seeded
random weights: no checkpoints offline) with fp32-grade GEMMs (bf16 head + remainder pairs, three tcgen05 kind::f16 MMAs
per product, fp32 accumulate), fp64 mask-weighted MVDR, permutation-aligned overlap-add, activity gate, iSTFT.  One "step"
= one pass of that path over the whole meeting.

  value : audio-seconds / wall-second with the raw audio already resident in HBM (device-timed, CUDA events,
          max over ranks).  N > 1 (default --multi shard): ONE meeting of N x --seconds, its segments sharded over the
          ranks (notsofar_b200.sharded: two small NCCL all-gathers + the waveform hand-off to rank 0) -> weak scaling;
          --scaling strong keeps the meeting at --seconds in total; --multi replicas gives every rank its own meeting.
  e2e   : the same through the public API notsofar_b200.separate_and_stitch with HOST buffers -- H2D of the pinned
          raw audio and D2H of the three separated waveforms inside the timed region (the audio streams in behind the
          first chunk of segments, the waveforms stream out behind the mask network: progressive tail, DESIGN.md 4;
          N > 1: every rank reads its own samples back the same way, css_device_sharded(host_piece=...)).
  roofline     : the dominant kernel class of the step (the tcgen05 GEMM), algorithmic flops / event-timed duration
                 vs MEASURED_PEAKS.json; the other classes are listed in "kernels".
  cpu_baseline : the numpy port of the reference's path (oracle/, same algorithm, all host cores) on a bounded slice.

--impl reference times that CPU port alone (rank 0 only), same metric / unit / config.
--workload session: a 6-minute session (the NOTSOFAR session length) per GPU through the plug-in call css_inference, 7 WAV
  files in (pageable numpy from the file reader), 4 WAV files out, inside the timed region.
--check: correctness of the sharded path under the real process group: rank 0's sharded waveforms / permutations /
  activity against the single-device path on a 2-minute meeting (one JSON line, non-zero exit on mismatch).
tcpWER is not evaluated anywhere here: dataset, checkpoints and meeteval are unavailable offline;
the stage-wise parity suite under tests/ stands in.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 16000
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel class, from the committed
# `ncu --set full` capture under profiles/ (None until such a capture exists for the class)
NCU_TRAFFIC = {
    # mean over the six GEMMs of one Conformer block at the bench size (1 209 segments in one chunk), CTA-pair kernels with folded
    # LayerNorms (profiles/r02_ncu_full_gemm16_cta_pairs_summary.csv: 2.28 / 1.81 / 1.34 / 1.34 / 1.82 / 1.34 GB read + written):
    # activations in + out dominate (e.g. FFN W2: 920 MB of bf16 pairs read, 460 MB of x read and written, 460 MB of operand planes)
    ("gemm_tc", "2xbf16"): 1657.0e6,
}
METRIC = "audio-sec/sec (xRT) CSS+MVDR 7-ch 16kHz"
UNIT = "audio-s/s"
QUALITY = "tcpWER not evaluated (dataset, checkpoints and meeteval unavailable offline); stage-wise parity vs the reference: tests/ -m gpu"


def restore_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arms are meant to use every host core (the BLAS pools were
    sized at import time, so the limit is lifted at run time)."""
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(cores)
    except Exception:
        pass
    return cores


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    which="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, which="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi SM clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_meeting(seconds: float, seed: int):
    """5-minute seeded pattern (notsofar_b200.synth.synthetic_meeting) tiled to the requested length."""
    from notsofar_b200 import synth
    n = int(round(seconds * FS))
    base_s = min(seconds, 300.0)
    base = synth.synthetic_meeting(base_s, seed=seed)
    reps = -(-n // len(base))
    return np.tile(base, (reps, 1))[:n] if reps > 1 else base[:n]


def cpu_port_xrt(x_slice: np.ndarray, weights, repeats: int = 1):
    """Times the numpy port of the reference path on x_slice [n, 7]; returns (xRT, seconds of CPU work)."""
    from oracle import css_oracle as O
    restore_host_threads()
    t0 = time.perf_counter()
    for _ in range(repeats):
        O.separate_and_stitch(x_slice[None], weights, FS, O.OracleCfg(activity_th=0.3), dtype=np.float32)
    dt = (time.perf_counter() - t0) / repeats
    return len(x_slice) / FS / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (numpy port under oracle/) on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from notsofar_b200 import synth
    cores = restore_host_threads()
    sample_s = args.ref_seconds
    x = make_meeting(sample_s, seed=0)
    w = synth.random_state_dict(0)
    from oracle import css_oracle as O
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.separate_and_stitch(x[None], w, FS, O.OracleCfg(activity_th=0.3), dtype=np.float32)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = sample_s / (ms / 1e3)
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            # the arm's own workload; every step runs a bounded slice of it (cost is linear in segments)
            "config": {"workload": f"CSS Conformer v1.0-MC + MVDR, 7-ch 16 kHz, {args.seconds / 60:.0f}-min synthetic meeting per GPU "
                                   f"({int(O.plan_segments(int(args.seconds * FS), FS, O.OracleCfg()).num_segments)} segments of 186 frames)",
                       "sample": f"{sample_s:.0f}-s slice per step ({int(O.plan_segments(len(x), FS, O.OracleCfg()).num_segments)} segments)",
                       "gemm_engine": "numpy fp32 (BLAS)", "parallelism": "host threads"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample_s:.0f} s of the synthetic 7-ch meeting per step, numpy/BLAS on all host threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "quality": QUALITY}
    print(json.dumps(line), flush=True)


def roofline_of(prof, prof_ms, steps, engine):
    """Per-kernel-class durations (library CUDA-event brackets on the launching stream) -> (kernels, roofline of the
    class that takes the largest share of the step)."""
    peaks = _peaks()
    kernels = {}
    for name, (ms, work, cnt) in prof.items():
        if cnt == 0:
            continue
        per = {"ms_per_step": ms / steps, "brackets_per_step": cnt / steps}
        if name.startswith("gemm") or name == "attention":
            per.update(bound="tensor", achieved=work / (ms * 1e-3) / 1e12 if ms > 0 else None, unit="TFLOP/s")
        elif name != "net_other":
            per.update(bound="hbm", achieved=work / (ms * 1e-3) / 1e9 if ms > 0 else None, unit="GB/s")
        kernels[name] = per
    dom = max((k for k in kernels if "achieved" in kernels[k]), key=lambda k: kernels[k]["ms_per_step"])
    d = kernels[dom]
    if d["bound"] == "tensor":
        # the fraction is reported against the measured sustained bf16 GEMM peak (the only measured tensor number)
        peak = peaks["bf16_sustained"]
        note = f"algorithmic fp32 flops (the 3 tensor-core passes of a split product count once) vs sustained bf16 cuBLAS peak, {peaks['which']}"
    else:
        peak = peaks["hbm_gbs"]
        note = f"algorithmic bytes vs copy bandwidth, {peaks['which']}"
    extra = {}
    if d["bound"] == "tensor" and engine == "3xtf32":
        # every algorithmic product costs three kind::tf32 MMAs, and the dense tf32 rate is half the bf16 rate:
        # the ceiling of this arithmetic is peak / 6
        extra = {"issued_tf32_tflops": 3 * d["achieved"], "frac_of_3xtf32_ceiling": 6 * d["achieved"] / peak}
    elif d["bound"] == "tensor" and engine in ("2xbf16", "2xf16"):
        # three kind::f16 MMAs per algorithmic product: the ceiling of this arithmetic is peak / 3
        extra = {"issued_f16_tflops": 3 * d["achieved"], "frac_of_split16_ceiling": 3 * d["achieved"] / peak}
    roofline = {**extra, "kernel": dom, "bound": d["bound"], "achieved": d["achieved"], "peak": peak, "unit": d["unit"],
                "frac": d["achieved"] / peak, "traffic": NCU_TRAFFIC.get((dom, engine)), "share_of_step": d["ms_per_step"] / prof_ms, "note": note}
    return kernels, roofline


def run_b200_sharded(args, world, rank, local_rank, dev, lib):
    """N > 1: ONE synthetic meeting of N x --seconds, segments sharded over the ranks in contiguous blocks with a
    one-segment halo; exchanges: all-gather of the 3x3 stitching costs and of the per-frame mask means, gather of the
    separated waveforms to rank 0 (NCCL over NVLink).  Per-rank work is fixed as N grows -> weak scaling."""
    import torch
    import torch.distributed as dist
    import notsofar_b200 as N
    from notsofar_b200 import synth, _cabi
    from notsofar_b200.css import HostFeeder
    from notsofar_b200.sharded import css_device_sharded, make_shard

    from notsofar_b200.scheduler import bind_host_to_gpu
    numa = bind_host_to_gpu(local_rank)          # before the first page-locked allocation: host buffers on the GPU's NUMA node
    seconds = args.seconds if args.scaling == "weak" else args.seconds / world
    total_s = seconds * world
    n_total = int(round(total_s * FS))
    cfg = N.CssCfg(activity_th=0.3, show_progressbar=False)
    plan = N.plan_segments(n_total, FS, cfg)
    sh = make_shard(plan, rank, world)
    # the 5-min seeded pattern tiles the long meeting; every rank materialises only its own sample range
    base = synth.synthetic_meeting(min(total_s, 300.0), seed=0)
    idx0, idx1 = sh.sample_lo, sh.sample_hi
    reps = -(-idx1 // len(base))
    x_np = np.ascontiguousarray(np.tile(base, (reps - idx0 // len(base), 1))[idx0 % len(base): idx0 % len(base) + (idx1 - idx0)])
    x_pinned = torch.from_numpy(x_np).pin_memory()
    weights = synth.random_state_dict(0)
    engine = {"3xtf32": N.GEMM_TC_3XTF32, "tf32": N.GEMM_TC_TF32, "simt": N.GEMM_SIMT_FP32, "2xbf16": N.GEMM_TC_2XBF16,
              "2xf16": N.GEMM_TC_2XF16}[args.engine]
    sep = N.ConformerCssB200(weights, device=dev, gemm_engine=engine, segments_per_batch=args.segments_per_batch)
    x_dev = x_pinned.to(dev)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return css_device_sharded(x_dev, sep, FS, cfg, n_total)

    def step_e2e():
        chunk = 32 * plan.hop_frames * 256
        feeder = HostFeeder(x_pinned, dev, chunk)
        # the assembled streams stay on rank 0's GPU (the ASR / diarization hand-off); the host copy of the result is read
        # by every rank for its own samples in parallel (each over its own PCIe link), seams included in the pieces -- the
        # interior of a piece while the mask network is still running (ShardWorker.phase1 / finish_host)
        from notsofar_b200.css import _pinned_out
        host = _pinned_out((3, sh.n_own_frames * 256 + 256))
        out = css_device_sharded(feeder, sep, FS, cfg, n_total, host_piece=host)
        torch.cuda.current_stream(dev).synchronize()
        return out

    for _ in range(args.warmup):
        out = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.nsf_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    launches = (lib.nsf_launch_count() - launches0) // args.steps
    clocks = sampler.stop()
    # rank 0's per-kernel-class brackets (same second pass as the single-GPU arm)
    lib.nsf_prof_enable(1)
    _cabi.prof_collect()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
    ev1.record()
    barrier()
    prof_ms = ev0.elapsed_time(ev1) / args.steps
    prof = _cabi.prof_collect()
    lib.nsf_prof_enable(0)
    del out
    e2e_ms = float("nan")
    if not args.skip_e2e:
        for _ in range(min(2, args.warmup)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(args.steps):
            step_e2e()
        ev1.record()
        barrier()
        e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3) / args.steps
    t = torch.tensor([dev_ms, e2e_ms, float(launches)], device=dev, dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms = tmax[0].item(), tmax[1].item()
    if rank == 0:
        n_out = (plan.mix_frames - 1) * 256 + 512
        kernels, roofline = roofline_of(prof, prof_ms, args.steps, args.engine)
        line = {"metric": METRIC, "value": total_s / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": DTYPES[args.engine], "quality": QUALITY,
                "data": "synthetic (seeded 5-min 7-ch pattern tiled; random-init v1.0-MC weights)",
                "config": {"workload": f"CSS Conformer v1.0-MC + MVDR, 7-ch 16 kHz, ONE {total_s / 60:.0f}-min synthetic meeting "
                                       f"({plan.num_segments} segments of 186 frames) sharded over {world} GPUs = {seconds / 60:.0f} min per GPU",
                           "segments_per_batch": args.segments_per_batch, "gemm_engine": args.engine,
                           "parallelism": f"segment-sharded x{world}: contiguous blocks + 1-segment halo; NCCL all-gather of stitching costs "
                                          f"(36 B/segment) and mask means (12 B/frame), point-to-point hand-off of the separated streams to rank 0",
                           "l2": "inputs and intermediates exceed L2 (>5 GB touched per step per GPU)"},
                "clocks": clocks,
                "e2e": {"value": total_s / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(n_total * 7 * 4 + (world - 1) * (plan.segment_frames + 1) * 256 * 7 * 4),
                        "d2h_bytes_per_step": int(3 * (n_out + 256 * (world - 1)) * 4 + plan.num_segments * 36 * world),
                        "d2h": "every rank reads its own samples (pieces overlap by one 256-sample seam); the NVLink gather to rank 0 is inside the step",
                        "host_numa": numa},
                "gpu_launches": int(t[2].item()) * args.steps, "gpu_launches_per_step": int(t[2].item()),
                "roofline": roofline, "kernels": kernels}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


DTYPES = {"3xtf32": "f32 (3xTF32 tensor-core GEMMs, fp64 MVDR)", "tf32": "tf32", "simt": "f32",
          "2xbf16": "f32 (bf16 head+remainder pairs, 3 kind::f16 MMAs per product, fp32 accumulate; fp64 MVDR)",
          "2xf16": "f32 (scaled fp16 head+remainder pairs = 22 mantissa bits, 3 kind::f16 MMAs per product, "
                   "fp32 accumulate; fp64 MVDR)"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import notsofar_b200 as N
    from notsofar_b200 import synth, _cabi
    from notsofar_b200.css import css_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug output (NCCL_DEBUG=VERSION|INFO) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()

    seconds = args.seconds
    sharded = world > 1 and args.multi == "shard"
    if sharded:
        # one meeting of world x `seconds`, its segments sharded over the ranks (notsofar_b200.sharded): every rank
        # separates `seconds` of audio (weak scaling) and the separated streams are gathered on rank 0
        return run_b200_sharded(args, world, rank, local_rank, dev, lib)
    x_np = make_meeting(seconds, seed=rank)                   # every rank its own meeting (sessions are independent)
    n = len(x_np)
    x_pinned = torch.from_numpy(x_np).pin_memory()
    weights = synth.random_state_dict(0)
    engine = {"3xtf32": N.GEMM_TC_3XTF32, "tf32": N.GEMM_TC_TF32, "simt": N.GEMM_SIMT_FP32, "2xbf16": N.GEMM_TC_2XBF16,
              "2xf16": N.GEMM_TC_2XF16}[args.engine]
    sep = N.ConformerCssB200(weights, device=dev, gemm_engine=engine, segments_per_batch=args.segments_per_batch)
    cfg = N.CssCfg(activity_th=0.3, show_progressbar=False)   # inference_v1.yaml:17
    plan = N.plan_segments(n, FS, cfg)
    x_dev = x_pinned.to(dev)
    h2d_bytes = x_pinned.numel() * 4
    n_out = (plan.mix_frames - 1) * 256 + 512
    d2h_bytes = 3 * n_out * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return css_device(x_dev, sep, FS, cfg)

    def step_e2e():
        return N.separate_and_stitch(x_pinned[None], sep, FS, dev, cfg, return_side_info=False)

    # ---- warm-up
    for _ in range(args.warmup):
        out = step_device()
    barrier()

    # ---- timed: device-resident input.  Inputs + intermediates (> 5 GB per step) far exceed the 126 MB L2.
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.nsf_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    launches = (lib.nsf_launch_count() - launches0) // args.steps
    clocks = sampler.stop()

    # ---- the same steps again with the library's per-kernel-class CUDA-event brackets switched on (they cost a few
    #      percent, so they stay out of the timed loop above): per-class durations and algorithmic work for the roofline
    lib.nsf_prof_enable(1)
    _cabi.prof_collect()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
    ev1.record()
    barrier()
    prof_ms = ev0.elapsed_time(ev1) / args.steps
    prof = _cabi.prof_collect()
    lib.nsf_prof_enable(0)
    del out

    # ---- timed: end to end through the public API with host buffers
    e2e_ms = float("nan")
    if not args.skip_e2e:
        # the caller keeps the previous result while the next call runs (results are owned arrays: a pinned buffer is only
        # recycled once its arrays were dropped), so two page-locked buffers alternate; both exist after the warm-up
        wavs = None
        for _ in range(max(3, min(3, args.warmup))):
            wavs, _ = step_e2e()
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(args.steps):
            wavs, _ = step_e2e()
        ev1.record()
        barrier()
        e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3) / args.steps
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = t.tolist()

    if rank == 0:
        value = world * seconds / (dev_ms / 1e3)
        e2e_val = world * seconds / (e2e_ms / 1e3)
        kernels, roofline = roofline_of(prof, prof_ms, args.steps, args.engine)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": DTYPES[args.engine], "quality": QUALITY,
                "data": "synthetic (seeded 5-min 7-ch pattern tiled; random-init v1.0-MC weights)",
                "config": {"workload": f"CSS Conformer v1.0-MC + MVDR, 7-ch 16 kHz, {seconds / 60:.0f}-min synthetic meeting per GPU "
                                       f"({plan.num_segments} segments of 186 frames)",
                           "segments_per_batch": args.segments_per_batch, "gemm_engine": args.engine,
                           "parallelism": f"{world} independent meetings (one per GPU)",
                           "l2": "inputs and intermediates exceed L2 (>5 GB touched per step)"},
                "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes + plan.num_segments * 36},
                "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
                "roofline": roofline, "kernels": kernels}
        if world == 1 and not args.no_cpu_baseline:
            sl = x_np[: int(args.cpu_seconds * FS)]
            xrt, dt = cpu_port_xrt(sl, weights)
            line["cpu_baseline"] = {"value": xrt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"first {args.cpu_seconds:.0f} s of the same meeting ({dt:.1f} s of CPU work), numpy port of the "
                                              f"reference path on all host threads; cost is linear in segments"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_check(args):
    """--check: the sharded path under the real process group (NCCL on the GPUs of this box) against the single-device path
    on the same 2-minute meeting: permutations and activity masks bit-exact, waveforms to float32 round-off."""
    import torch
    import torch.distributed as dist
    import notsofar_b200 as N
    from notsofar_b200 import synth
    from notsofar_b200.css import css_device
    from notsofar_b200.sharded import css_device_sharded, make_shard
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    seconds = args.check_seconds
    x_np = synth.synthetic_meeting(seconds, seed=1)
    n_total = len(x_np)
    sep = N.ConformerCssB200(synth.random_state_dict(0), device=dev, segments_per_batch=args.segments_per_batch)
    probe_cfg = N.CssCfg(show_progressbar=False)
    plan = N.plan_segments(n_total, FS, probe_cfg)
    x_full = torch.from_numpy(x_np).to(dev)
    # a threshold inside the activity spread so that the gate (dilate / erode across the shard seams) is exercised; every
    # rank derives the same one from the same single-device probe
    probe = css_device(x_full, sep, FS, probe_cfg)
    cfg = N.CssCfg(activity_th=float(torch.quantile(probe["activity"].flatten(), 0.9)), show_progressbar=False)
    one = css_device(x_full, sep, FS, cfg)
    sh = make_shard(plan, rank, world)
    x_loc = x_full[sh.sample_lo:sh.sample_hi].contiguous()
    out = css_device_sharded(x_loc, sep, FS, cfg, n_total)
    torch.cuda.synchronize()
    ok = torch.tensor([1], device=dev)
    line = None
    if rank == 0:
        a, b = out["wav"], one["wav"]
        rel = float((a - b).norm() / b.norm())
        perms_equal = bool(np.array_equal(out["perms"], one["perms"]))
        act_equal = bool(torch.equal(out["activity_b"], one["activity_b"]) and torch.equal(out["activity_final"], one["activity_final"]))
        gate_on = float(one["activity_final"].float().mean())
        good = perms_equal and act_equal and rel < 1e-5 and a.shape == b.shape
        line = {"check": "sharded (NCCL) vs single-device css_device", "n_gpus": world, "seconds": seconds, "segments": plan.num_segments,
                "wav_rel_l2": rel, "wav_max_abs": float((a - b).abs().max()), "perms_equal": perms_equal, "activity_equal": act_equal,
                "activity_final_on_fraction": gate_on, "tolerance": 1e-5, "ok": bool(good)}
        ok[0] = 1 if good else 0
    dist.broadcast(ok, 0)
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
    if int(ok.item()) != 1:
        sys.exit(1)


def run_session(args):
    """--workload session: one NOTSOFAR-sized session (default 6 min) per GPU through the reference-facing plug-in call
    css_inference -- 7 mono WAVs read from disk into pageable numpy, model resident, separated streams peak-normalised and
    written as 16-bit WAVs -- everything inside the timed region (wall clock, max over ranks)."""
    import tempfile
    import pandas as pd
    import scipy.io.wavfile as wf
    import torch
    import torch.distributed as dist
    import notsofar_b200 as N
    from notsofar_b200 import synth, _cabi
    from notsofar_b200 import css as css_mod
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()
    seconds = args.session_seconds
    tmp = tempfile.mkdtemp(prefix=f"nsf_bench_r{rank}_")
    model_dir = os.path.join(tmp, "models", "notsofar", "conformer1.0", "mc")
    os.makedirs(model_dir)
    torch.save({"model": {"module." + k: torch.from_numpy(np.asarray(v)) for k, v in synth.random_state_dict(0).items()}},
               os.path.join(model_dir, "model.pt"))
    open(os.path.join(model_dir, "cfg.yaml"), "w").write("single_channel: false\n")
    cfg = N.CssCfg(show_progressbar=False, activity_th=0.3)
    css_mod.ASYNC_WAV_WRITES = bool(args.async_wav)

    def make_session(idx, secs):
        x = synth.synthetic_meeting(secs, seed=idx)
        names = []
        for c in range(7):
            f = os.path.join(tmp, f"s{idx}_ch{c}.wav")
            wf.write(f, FS, np.clip(np.rint(x[:, c] * 32768.0 * 8), -32768, 32767).astype(np.int16))
            names.append(f)
        return pd.Series(dict(session_id=f"multichannel/MTG_bench_{idx}", meeting_id=f"MTG_bench_{idx}", is_mc=True, wav_file_names=names)), len(x)

    if world == 1:
        session, n_samp = make_session(0, seconds)
        total_audio_s, total_samples = seconds, n_samp

        def step(i):
            return N.css_inference(os.path.join(tmp, f"out{i % 2}"), os.path.join(tmp, "models"), session, cfg, fetch_from_cache=False)
    else:
        # 2 sessions per GPU of 2/3, 1 and 4/3 of --session-seconds, handed out longest-first by the session scheduler
        # (scheduler.css_inference_distributed: one all_gather_object of the finished rows over the process group)
        from notsofar_b200.scheduler import assign_sessions, css_inference_distributed
        durations = [seconds * (2 + (i % 3)) / 3.0 for i in range(2 * world)]
        mine = set(assign_sessions(durations, world)[rank])
        sessions, total_samples = [], 0
        for i, d in enumerate(durations):
            if i in mine:
                sess, n_samp = make_session(i, d)
            else:                                                   # another rank's session: the row alone (its files live in that rank's directory)
                sess, n_samp = pd.Series(dict(session_id=f"multichannel/MTG_bench_{i}", meeting_id=f"MTG_bench_{i}", is_mc=True,
                                              wav_file_names=[f"/nonexistent/s{i}_ch{c}.wav" for c in range(7)])), int(round(d * FS))
            sessions.append(sess)
            total_samples += n_samp
        total_audio_s = float(sum(durations))

        def step(i):
            return css_inference_distributed(os.path.join(tmp, f"out{i % 2}"), os.path.join(tmp, "models"), sessions, cfg, False, durations=durations)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 1)):
        step(i)
    barrier()
    launches0 = lib.nsf_launch_count()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    css_mod.flush_wav_writes()
    barrier()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    launches = (lib.nsf_launch_count() - launches0) // args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if rank == 0:
        val = total_audio_s / (ms / 1e3)
        print(json.dumps({"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPES["2xbf16"],
                          "quality": QUALITY, "data": "synthetic (seeded 7-ch session; random-init v1.0-MC weights)",
                          "config": {"workload": (f"css_inference plug-in call on a {seconds / 60:.0f}-min 7-ch 16 kHz session" if world == 1 else
                                                  f"{2 * world} sessions of {seconds * 2 / 180:.0f}-{seconds * 4 / 180:.0f} min through scheduler.css_inference_distributed") +
                                                 ": 7 WAV files in (pageable host memory), model resident, 1 + 3 peak-normalised PCM16 WAV files out per session",
                                     "parallelism": "one session at a time on one GPU" if world == 1 else f"sessions assigned longest-first to {world} GPUs (one process each)",
                                     "timing": "wall clock around the call, max over ranks",
                                     "wav_writes": "asynchronous (flushed before the clock stops)" if args.async_wav else "synchronous (parallel writer threads)"},
                          "e2e": {"value": val, "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": int(total_samples * 7 * 2),
                                  "d2h_bytes_per_step": int(4 * total_samples * 2), "file_bytes_read": int(total_samples * 7 * 2),
                                  "file_bytes_written": int(4 * total_samples * 2), "note": "int16 crosses PCIe in both directions (device-side conversion / quantisation)"},
                          "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches)}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=float, default=1800.0, help="meeting length per GPU")
    ap.add_argument("--engine", default="2xbf16", choices=["3xtf32", "tf32", "simt", "2xbf16", "2xf16"])
    ap.add_argument("--segments-per-batch", type=int, default=1280)
    ap.add_argument("--cpu-seconds", type=float, default=120.0, help="slice of the meeting the CPU baseline leg runs")
    ap.add_argument("--ref-seconds", type=float, default=15.0, help="--impl reference: audio seconds per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--multi", default="shard", choices=["shard", "replicas"],
                    help="N > 1: shard ONE N x --seconds meeting by segments over the ranks (default), or give every rank its own meeting")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: device-resident loop alone")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1, --multi shard: weak = one meeting of N x --seconds, strong = one meeting of --seconds split over the ranks")
    ap.add_argument("--workload", default="meeting", choices=["meeting", "session"],
                    help="meeting: BASELINE config 2 (default); session: a --session-seconds session per GPU through css_inference with file I/O")
    ap.add_argument("--session-seconds", type=float, default=360.0)
    ap.add_argument("--async-wav", action="store_true", help="--workload session: css_inference returns while the WAV files are still being written")
    ap.add_argument("--check", action="store_true", help="compare the sharded path under the real process group with the single-device path")
    ap.add_argument("--check-seconds", type=float, default=120.0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1 and args.impl == "b200":
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    elif args.check:
        run_check(args)
    elif args.workload == "session":
        run_session(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
