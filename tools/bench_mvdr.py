#!/usr/bin/env python
"""BASELINE.json config 5: MVDR kernel sweep -- 257 bins x {1k ... 100k} frames x 7 mics as batches of 186-frame segments
(the pipeline shape), achieved algorithmic GB/s (96 B per bin-frame) against the measured HBM copy bandwidth.
Inputs follow SURVEY 8d: mix ~ CN(0,1) with a per-bin rank-1 + identity spatial structure, masks = softmax(3 N(0,1)).
Run on a B200: python tools/bench_mvdr.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from notsofar_b200 import _cabi

def main():
    """Under torchrun every rank sweeps its own data (bins x frames are independent: no collective on this path) and rank 0
    prints the aggregate of the per-rank rates."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    lib = _cabi.load()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
        if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
    T, F, C, S = 186, 257, 7, 3
    g = torch.Generator(device=dev).manual_seed(0)
    sizes = [int(a) for a in sys.argv[1:]] or [1000, 3000, 10000, 30000, 100000]      # python tools/bench_mvdr.py [frames ...]
    # form "segments": the pipeline shape -- 186-frame segments every 93 frames (css.py:141-152); NSF_MVDR_HOP=186 gives
    # disjoint segments (the round-1 sweep; one warp per (segment, bin) kernel)
    hop = int(os.environ.get("NSF_MVDR_HOP", "93"))
    for frames in sizes:
        n_seg = max(1, -(-(frames - (T - hop)) // hop))
        T_long = (n_seg - 1) * hop + T
        a = torch.randn(F, 1, C, 2, device=dev, generator=g)
        s = torch.randn(F, T_long, 1, 2, device=dev, generator=g)
        steer = torch.view_as_complex(a.contiguous()) * torch.view_as_complex(s.contiguous())
        X = (steer + torch.view_as_complex(torch.randn(F, T_long, C, 2, device=dev, generator=g))).contiguous()
        masks = torch.softmax(3 * torch.randn(n_seg, S + 1, F, T, device=dev, generator=g), dim=1).contiguous()
        Y = torch.empty(n_seg, S, F, T, dtype=torch.complex64, device=dev)
        for i in range(8):
            if i == 3:
                lib.nsf_prof_enable(1); _cabi.prof_collect()
            _cabi.check(lib.nsf_mvdr(_cabi.ptr(masks), S, 1, _cabi.ptr(X), T_long, T_long, C, 0, n_seg, T, hop, F, 1.0, _cabi.ptr(Y),
                                     _cabi.stream_ptr()), "nsf_mvdr")
        prof = _cabi.prof_collect(); lib.nsf_prof_enable(0)
        ms, work, cnt = prof["mvdr"]
        gbs = work / (ms * 1e-3) / 1e9
        if world > 1:
            t = torch.tensor([gbs], device=dev, dtype=torch.float64)
            dist.all_reduce(t)
            gbs = t.item()
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"form": f"batched {T}-frame segments, hop {hop}", "impl": os.environ.get("NSF_MVDR_IMPL", "default"), "frames_per_gpu": frames, "segments_per_gpu": n_seg, "n_gpus": world, "us_per_launch": ms / cnt * 1e3,
                              "GB/s": gbs, "frac_of_measured_hbm": gbs / (world * peaks["hbm_gbs"]),
                              "note": "96 algorithmic bytes per (bin, segment frame): 56 mix + 16 masks + 24 out"}), flush=True)
    # form "one long utterance" (make_mvdr accepts any T, mvdr_util.py:5-47): split-T kernels
    if os.environ.get("NSF_MVDR_FORMS", "both") == "both":
        for frames in sizes:
            a = torch.randn(F, 1, 7, 2, device=dev, generator=g)
            s_ = torch.randn(F, frames, 1, 2, device=dev, generator=g)
            X = (torch.view_as_complex(a.contiguous()) * torch.view_as_complex(s_.contiguous()) +
                 torch.view_as_complex(torch.randn(F, frames, 7, 2, device=dev, generator=g))).contiguous()
            masks = torch.softmax(3 * torch.randn(S + 1, F, frames, device=dev, generator=g), dim=0).contiguous()
            Y = torch.empty(S, F, frames, dtype=torch.complex64, device=dev)
            need = int(lib.nsf_mvdr_utterance_workspace_bytes(S, frames, F))
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            for i in range(8):
                if i == 3:
                    lib.nsf_prof_enable(1); _cabi.prof_collect()
                _cabi.check(lib.nsf_mvdr_utterance(_cabi.ptr(masks), S, 1, _cabi.ptr(X), frames, 7, F, 1.0, _cabi.ptr(Y), _cabi.ptr(ws), need,
                                                   _cabi.stream_ptr()), "nsf_mvdr_utterance")
            prof = _cabi.prof_collect(); lib.nsf_prof_enable(0)
            ms, work, cnt = prof["mvdr"]
            gbs = work / (ms * 1e-3) / 1e9
            if world > 1:
                t = torch.tensor([gbs], device=dev, dtype=torch.float64)
                dist.all_reduce(t)
                gbs = t.item()
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"form": "one long utterance", "frames_per_gpu": frames, "n_gpus": world, "us_per_launch": ms / cnt * 1e3,
                                  "GB/s": gbs, "frac_of_measured_hbm": gbs / (world * peaks["hbm_gbs"]),
                                  "note": "96 algorithmic bytes per (bin, frame); three kernels: partial covariances, per-bin solve, apply"}), flush=True)
    # the reference's CPU path beside it: the numpy port of make_mvdr (oracle/) on a bounded sample, all host threads
    if int(os.environ.get("RANK", "0")) == 0 and os.environ.get("NSF_MVDR_CPU", "1") == "1":
        import time
        import numpy as np
        from oracle import css_oracle as O
        try:
            from threadpoolctl import threadpool_limits
            threadpool_limits(limits=os.cpu_count())
        except Exception:
            pass
        rng = np.random.default_rng(0)
        for frames, form in ((1000, "one long utterance"), (3000, "one long utterance"), (186 * 8, "batched 186-frame segments, hop 186")):
            Xc = (rng.standard_normal((7, F, frames)) + 1j * rng.standard_normal((7, F, frames))).astype(np.complex64)
            m = rng.random((4, F, frames)).astype(np.float32)
            t0 = time.perf_counter()
            if form.startswith("one"):
                O.make_mvdr(m[:3], m[3:], Xc, np.float32)
            else:
                for i in range(frames // 186):
                    O.make_mvdr(m[:3, :, i * 186:(i + 1) * 186], m[3:, :, i * 186:(i + 1) * 186], Xc[:, :, i * 186:(i + 1) * 186], np.float32)
            dt = time.perf_counter() - t0
            print(json.dumps({"form": form, "impl": "cpu reference (numpy port of make_mvdr, complex64, all host threads)", "frames": frames,
                              "cores": os.cpu_count(), "ms": dt * 1e3, "GB/s": F * frames * 96 / dt / 1e9}), flush=True)
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
