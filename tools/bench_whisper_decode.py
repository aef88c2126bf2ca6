#!/usr/bin/env python
"""BASELINE.json config 3: Whisper-large-v3-shaped encoder + greedy decode on one B200 (random weights: the token stream is
noise, the work per step is the real model's).  python tools/bench_whisper_decode.py [n_chunks=64] [n_steps=64]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from notsofar_b200 import _cabi
from notsofar_b200.whisper import WhisperB200
from tools.bench_whisper_encoder import random_large_v3

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    dev = torch.device("cuda", 0)
    d, L, ffn, vocab = 1280, 32, 5120, 51866
    sd = random_large_v3()
    g = torch.Generator().manual_seed(1)
    r = lambda *s: torch.randn(*s, generator=g) * 0.02
    sd["decoder.token_embedding.weight"] = r(vocab, d); sd["decoder.positional_embedding"] = r(448, d)
    sd["decoder.ln.weight"] = torch.ones(d); sd["decoder.ln.bias"] = torch.zeros(d)
    for l in range(L):
        p = f"decoder.blocks.{l}."
        for a in ("attn", "cross_attn"):
            for nm in ("query", "key", "value", "out"):
                sd[p + f"{a}.{nm}.weight"] = r(d, d)
                if nm != "key":
                    sd[p + f"{a}.{nm}.bias"] = r(d)
            sd[p + f"{a}_ln.weight"] = torch.ones(d); sd[p + f"{a}_ln.bias"] = torch.zeros(d)
        sd[p + "mlp_ln.weight"] = torch.ones(d); sd[p + "mlp_ln.bias"] = torch.zeros(d)
        sd[p + "mlp.0.weight"] = r(ffn, d); sd[p + "mlp.0.bias"] = r(ffn)
        sd[p + "mlp.2.weight"] = r(d, ffn); sd[p + "mlp.2.bias"] = r(d)
    wb = WhisperB200(sd, device=dev)
    del sd
    audio = torch.randn(B, 480000, device=dev) * 0.05
    hi, lo, _ = wb.encoder.log_mel(audio)
    _, enc16 = wb.encode(hi, lo)
    prompt = [50258, 50259, 50360, 50364]
    wb.decode_greedy(enc16, prompt, max_new_tokens=steps)          # warm-up: captures the step graph
    torch.cuda.synchronize()
    # one un-captured, teacher-forced pass of a few steps with the library's event brackets: where a step goes
    lib = _cabi.load()
    lib.nsf_prof_enable(1); _cabi.prof_collect()
    forced = torch.zeros((B, 8), dtype=torch.int32, device=dev)
    wb.decode_greedy(enc16, prompt, max_new_tokens=8, forced_tokens=forced, return_logits=True)
    torch.cuda.synchronize()
    prof = _cabi.prof_collect(); lib.nsf_prof_enable(0)
    n_prof = len(prompt) + 8 - 1
    classes = {("cross_attn_cache" if k == "mvdr" else ("self_attn_cache" if k == "attention" else k)): round(v[0] / n_prof, 3) for k, v in prof.items() if v[2]}
    t0 = time.perf_counter()
    _, enc16 = wb.encode(hi, lo)
    torch.cuda.synchronize()
    t_enc = time.perf_counter() - t0
    t0 = time.perf_counter()
    toks = wb.decode_greedy(enc16, prompt, max_new_tokens=steps)
    torch.cuda.synchronize()
    t_dec = time.perf_counter() - t0
    n_steps = toks.shape[1] - 1
    print(json.dumps({"workload": f"Whisper-large-v3-shaped encoder + greedy decode, {B} x 30-s chunks, bf16", "encoder_ms": t_enc * 1e3,
                      "decode_ms_total": t_dec * 1e3, "decode_steps": n_steps, "ms_per_step": t_dec * 1e3 / n_steps,
                      "tokens_per_s": B * n_steps / t_dec,
                      "audio_s_per_s_at_224_tokens": 30.0 * B / (t_enc + t_dec / n_steps * 228),
                      "ms_per_step_by_class (incl. prefill GEMMs / steps)": classes}))

if __name__ == "__main__":
    main()
