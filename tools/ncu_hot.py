#!/usr/bin/env python
"""Hot SASS instructions of one kernel in an `ncu --set full --import-source on` report:
    python tools/ncu_hot.py gpurun_out/X.ncu-rep [min_pct]
prints the per-instruction warp-stall samples (all samples / long-scoreboard share) above min_pct of the total."""
import csv, subprocess, sys

def main(rep, min_pct=0.5, ctx=0):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[1]
    c, s, ie = H.index("# Samples"), H.index("Source"), H.index("Instructions Executed")
    stalls = [(i, h) for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
    data = rows[2:]
    tot = sum(float(r[c] or 0) for r in data)
    tot_inst = sum(float(r[ie] or 0) for r in data)
    print(f"{rows[0][1][:80]}: {tot:.0f} samples, {tot_inst/1e6:.2f} M warp instructions")
    agg = {h: sum(float(r[i] or 0) for r in data) for i, h in stalls}
    print("stall totals:", ", ".join(f"{h[6:]} {100*v/tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot))
    for k, r in enumerate(data):
        v = float(r[c] or 0)
        if v >= min_pct / 100 * tot:
            top = max(stalls, key=lambda ih: float(r[ih[0]] or 0))
            print(f"{100*v/tot:5.1f}% i{k:5d} exec={r[ie]:>8s} {top[1][6:]:>12s} {r[s].strip()[:90]}")

if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)
