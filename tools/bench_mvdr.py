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
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
