#!/usr/bin/env python
"""Row a16: titanet-large-shaped speaker-embedding forward (random weights) on word crops of the six window scales of
configs/inference (3.0 ... 0.5 s), one B200.  Reports crops/s, audio-seconds of crops per second, algorithmic TFLOP/s of the
1x1-convolution GEMMs (valid frames only) against the measured bf16 peak and its third (the 2xBF16 split arithmetic issues
three MMAs per product), and the per-kernel-class split from the library's event profiler.
    python tools/bench_titanet.py [n_words=256] [--padded]   (--padded: one batch padded to the longest crop, as the reference batches)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from notsofar_b200 import _cabi
from notsofar_b200.titanet import TitaNetB200, TITANET_LARGE, HOP


def random_titanet_large(seed=0):
    """NeMo state_dict names with random values (no checkpoint offline); shapes of titanet-large."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    w, c_in = {}, 80

    def bn(name, c):
        w[name + ".weight"] = torch.rand(c, generator=g) + 0.5; w[name + ".bias"] = 0.1 * r(c)
        w[name + ".running_mean"] = 0.1 * r(c); w[name + ".running_var"] = torch.rand(c, generator=g) + 0.5

    for b, (co, rep, k, res) in enumerate(TITANET_LARGE):
        p, c, i = f"encoder.encoder.{b}.", c_in, 0
        for j in range(rep):
            w[p + f"mconv.{i}.conv.weight"] = r(c, 1, k) / k ** 0.5
            w[p + f"mconv.{i + 1}.conv.weight"] = r(co, c, 1) / c ** 0.5
            bn(p + f"mconv.{i + 2}", co)
            i += 3 if j == rep - 1 else 5
            c = co
        w[p + f"mconv.{i}.fc.0.weight"] = r(co // 8, co) / co ** 0.5
        w[p + f"mconv.{i}.fc.2.weight"] = r(co, co // 8) / (co // 8) ** 0.5
        if res:
            w[p + "res.0.0.conv.weight"] = r(co, c_in, 1) / c_in ** 0.5
            bn(p + "res.0.1", co)
        c_in = co
    p = "decoder._pooling.attention_layer."
    w[p + "0.conv_layer.weight"] = r(128, 3 * c_in, 1) / (3 * c_in) ** 0.5; w[p + "0.conv_layer.bias"] = 0.1 * r(128)
    bn(p + "0.bn", 128)
    w[p + "2.weight"] = r(c_in, 128, 1) / 128 ** 0.5; w[p + "2.bias"] = 0.1 * r(c_in)
    bn("decoder.emb_layers.0.0", 2 * c_in)
    w["decoder.emb_layers.0.1.weight"] = r(192, 2 * c_in, 1) / (2 * c_in) ** 0.5; w["decoder.emb_layers.0.1.bias"] = 0.1 * r(192)
    return w


def flop_per_frame():
    f, c_in = 0.0, 80
    for co, rep, k, res in TITANET_LARGE:
        c = c_in
        for j in range(rep):
            f += 2.0 * c * co
            c = co
        if res:
            f += 2.0 * c_in * co
        c_in = co
    return f + 2.0 * c_in * 128 * 2


def main():
    n_words = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
    padded = "--padded" in sys.argv
    dev = torch.device("cuda", 0)
    lib = _cabi.load()
    model = TitaNetB200(random_titanet_large(), dev)
    scales = [3.0, 2.5, 2.0, 1.5, 1.0, 0.5]
    lens = np.tile(np.asarray([int(s * 16000) for s in scales], np.int32), n_words)          # word-major like the crop plan
    n = len(lens)
    crops = torch.randn(n, int(lens.max()), device=dev) * 0.05
    lens_t = torch.from_numpy(lens).to(dev)
    crops *= (torch.arange(crops.shape[1], device=dev)[None, :] < lens_t[:, None])
    step = lambda: model.embed(crops, lens_t, bucket=not padded)
    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        out = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    lib.nsf_prof_enable(1); _cabi.prof_collect()
    step(); torch.cuda.synchronize()
    prof = _cabi.prof_collect(); lib.nsf_prof_enable(0)
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["bf16_tflops_sustained"] if os.path.exists(pk) else 1400.0
    frames = float((lens // HOP + 1).sum())
    tf = flop_per_frame() * frames / (ms * 1e-3) / 1e12
    print(json.dumps({"workload": f"titanet-large-shaped embedding forward, {n_words} words x 6 scales ({n} crops, {lens.sum() / 16000:.0f} s of audio), "
                                  + ("one batch padded to 3 s" if padded else "bucketed by length"),
                      "ms_per_batch": ms, "crops_per_s": n / (ms * 1e-3), "words_per_s": n_words / (ms * 1e-3),
                      "crop_audio_s_per_s": float(lens.sum()) / 16000 / (ms * 1e-3), "GEMM_TFLOP/s_algorithmic": tf,
                      "frac_of_bf16_sustained_peak": tf / peak, "frac_of_split16_ceiling": 3 * tf / peak,
                      "MFLOP_per_frame": flop_per_frame() / 1e6, "classes_ms": {k: round(v[0], 3) for k, v in prof.items() if v[2]},
                      "finite": bool(torch.isfinite(out).all())}))


if __name__ == "__main__":
    main()
