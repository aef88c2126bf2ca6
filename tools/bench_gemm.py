#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM engine on the mask network's shapes (uses the C-ABI test hook nsf_gemm_test and the
library's event profiler, so only the GEMM launch itself is timed).  Run on a B200: python tools/bench_gemm.py"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from notsofar_b200 import _cabi

def main():
    lib = _cabi.load()
    dev = torch.device("cuda", 0)
    n_seg = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    M = n_seg * 186
    shapes = [("embed", M, 512, 1824), ("ffn1", M, 1024, 512), ("ffn2", M, 512, 1024), ("qkv", M, 1536, 512),
              ("out", M, 512, 512), ("head", M, 1028, 512)]
    for eng_name, eng in (("3xtf32", 1), ("tf32", 2)):
        for name, m, n, k in shapes:
            A = torch.randn(m, k, device=dev); W = torch.randn(n, k, device=dev); b = torch.randn(n, device=dev)
            out = torch.empty(m, n, device=dev)
            ws = torch.empty(((m * k + 63) // 64 * 64 * 2 + (n * k + 63) // 64 * 64 * 2) * 4, dtype=torch.uint8, device=dev)
            for i in range(6):
                if i == 3:
                    lib.nsf_prof_enable(1); _cabi.prof_collect()
                _cabi.check(lib.nsf_gemm_test(eng, _cabi.ptr(A), _cabi.ptr(W), _cabi.ptr(b), _cabi.ptr(out), m, n, k, _cabi.ptr(ws),
                                              ws.numel(), _cabi.stream_ptr()), "gemm")
            prof = _cabi.prof_collect(); lib.nsf_prof_enable(0)
            ms, work, cnt = prof["gemm_tc"]
            ref = (A[:256].double() @ W.double().T + b.double())
            err = ((out[:256].double() - ref).norm() / ref.norm()).item()
            print(f"{eng_name:7s} {name:6s} M={m} N={n} K={k}: {ms / cnt * 1e3:8.1f} us  {work / (ms * 1e-3) / 1e12:7.1f} TFLOP/s  rel_err {err:.2e}", flush=True)

if __name__ == "__main__":
    main()
