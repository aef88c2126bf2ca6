"""Host / device timeline of one separate_and_stitch call on a 30-min recording (progressive tail on or off: NSF_PROGRESSIVE_CHUNK)."""
import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import notsofar_b200 as N
from notsofar_b200 import css as C
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
from oracle import css_oracle as O
w = O.random_weights(seed=0, gain=0.5)
sep = N.ConformerCssB200(w, device=dev)
n = 16000 * 1800
rng = np.random.default_rng(0)
base = (rng.standard_normal((16000 * 60, 7)) * 0.05).astype(np.float32)
x = torch.from_numpy(np.tile(base, (30, 1))[:n]).pin_memory()
cfg = N.CssCfg(activity_th=0.3, show_progressbar=False)
# instrument
marks = []
orig_check = C._cabi.check
def mark(name, stream=None):
    ev = torch.cuda.Event(enable_timing=True); ev.record(stream or torch.cuda.current_stream()); marks.append((name, time.perf_counter(), ev))
orig_masks = sep.masks
def masks(*a, **k):
    mark("net_begin"); r = orig_masks(*a, **k); mark("net_end"); return r
sep.masks = masks
lib = C._cabi.load()
orig_prog = lib.nsf_stitch_progress
def prog(*a):
    mark("tail_begin"); r = orig_prog(*a); mark("tail_end"); return r
lib.nsf_stitch_progress = prog
orig_chain = C.permutation_chain
def chain(*a, **k):
    t = time.perf_counter(); r = orig_chain(*a, **k); marks.append(("chain_done", time.perf_counter(), None)); return r
C.permutation_chain = chain
if os.environ.get("NSF_TRACE_PLAN"):            # e.g. "176,356,356,321": chunk sizes to try instead of plan_batches' own
    sizes = [int(v) for v in os.environ["NSF_TRACE_PLAN"].split(",")]
    def fixed_plan(n_seg, *a, **k):
        assert sum(sizes) == n_seg, (sum(sizes), n_seg)
        out, s0 = [], 0
        for c in sizes:
            out.append((s0, c)); s0 += c
        return out
    C.plan_batches = fixed_plan
walls = []
for it in range(9):
    marks.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0 = torch.cuda.Event(enable_timing=True); e0.record()
    wavs, _ = N.separate_and_stitch(x[None], sep, 16000, dev, cfg, return_side_info=False)
    t1 = time.perf_counter()
    if it >= 3:
        walls.append((t1 - t0) * 1e3)
    if it == 8:
        print("mean wall of 6 calls %.2f ms (min %.2f)" % (sum(walls) / len(walls), min(walls)))
        print("total wall %.1f ms" % ((t1 - t0) * 1e3))
        for name, th, ev in marks:
            print(f"{name:12s} host {1e3*(th-t0):7.2f} ms   gpu " + ("   -" if ev is None else f"{e0.elapsed_time(ev):7.2f} ms"))
