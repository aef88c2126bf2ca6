#!/usr/bin/env python
"""BASELINE.json config 3 (encoder part): Whisper-large-v3-shaped audio encoder (32 layers, d = 1280, 20 heads, ffn 5120,
128 mels), random weights, batches of 30-s chunks on one B200.  Reports ms per batch, algorithmic TFLOP/s against the
measured bf16 peak, audio-seconds per second, and the per-kernel-class split from the library's event profiler.
    python tools/bench_whisper_encoder.py [n_chunks=64] [--audio]   (--audio: include the log-mel front end)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from notsofar_b200 import _cabi
from notsofar_b200.whisper import WhisperEncoderB200

def random_large_v3(seed=0, d=1280, layers=32, ffn=5120, n_mels=128):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g) * 0.02
    sd = {"encoder.conv1.weight": r(d, n_mels, 3), "encoder.conv1.bias": r(d), "encoder.conv2.weight": r(d, d, 3), "encoder.conv2.bias": r(d),
          "encoder.positional_embedding": r(1500, d), "encoder.ln_post.weight": torch.ones(d), "encoder.ln_post.bias": torch.zeros(d)}
    for l in range(layers):
        p = f"encoder.blocks.{l}."
        for nm in ("query", "key", "value", "out"):
            sd[p + f"attn.{nm}.weight"] = r(d, d)
            if nm != "key":
                sd[p + f"attn.{nm}.bias"] = r(d)
        sd[p + "attn_ln.weight"] = torch.ones(d); sd[p + "attn_ln.bias"] = torch.zeros(d)
        sd[p + "mlp_ln.weight"] = torch.ones(d); sd[p + "mlp_ln.bias"] = torch.zeros(d)
        sd[p + "mlp.0.weight"] = r(ffn, d); sd[p + "mlp.0.bias"] = r(ffn)
        sd[p + "mlp.2.weight"] = r(d, ffn); sd[p + "mlp.2.bias"] = r(d)
    return sd

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
    with_audio = "--audio" in sys.argv
    dev = torch.device("cuda", 0)
    lib = _cabi.load()
    enc = WhisperEncoderB200(random_large_v3(), device=dev)
    D = enc.dims
    T, d, L, ff, nm = D.n_ctx, D.d_model, D.n_layers, D.d_ff, D.n_mels
    flop = 2.0 * 3000 * d * 3 * nm + 2.0 * T * d * 3 * d + L * (2.0 * T * d * (4 * d + 2 * ff) + 4.0 * T * T * d)
    audio = torch.randn(B, 480000, device=dev) * 0.05
    hi, lo, _ = enc.log_mel(audio)
    step = (lambda: enc.encode_audio(audio)) if with_audio else (lambda: enc.encode_mel(hi, lo))
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        out = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    lib.nsf_prof_enable(1); _cabi.prof_collect()
    step(); torch.cuda.synchronize()
    prof = _cabi.prof_collect(); lib.nsf_prof_enable(0)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0}
    tf = flop * B / (ms * 1e-3) / 1e12
    print(json.dumps({"workload": f"Whisper-large-v3-shaped encoder, {B} x 30-s chunks, bf16 tensor cores" + (" + log-mel front end" if with_audio else ""),
                      "ms_per_batch": ms, "audio_s_per_s": 30.0 * B / (ms * 1e-3), "TFLOP/s": tf, "TFLOP_per_chunk": flop / 1e12,
                      "frac_of_bf16_sustained_peak": tf / peaks["bf16_tflops_sustained"],
                      "classes_ms": {k: round(v[0], 2) for k, v in prof.items() if v[2]}, "finite": bool(torch.isfinite(out).all())}))

if __name__ == "__main__":
    main()
