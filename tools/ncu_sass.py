"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` of one kernel into the SASS evidence kept
under profiles/: opcode mix weighted by executions, the bulk-copy (TMA) / mbarrier instructions with their execution
counts, and a window of the hot loop with per-instruction executions and stall samples.
    ncu -i gpurun_out/X.ncu-rep --page source --csv --print-source sass > /tmp/x_sass.csv
    python tools/ncu_sass.py /tmp/x_sass.csv [--window-at MUFU.EX2 --window 170] > profiles/rNN_x_sass.txt"""
import argparse
import csv
from collections import Counter

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--title", default="")
ap.add_argument("--window-at", default="MUFU.EX2", help="the window starts a few instructions before the N-th match")
ap.add_argument("--nth", type=int, default=17)
ap.add_argument("--before", type=int, default=62)
ap.add_argument("--window", type=int, default=170)
a = ap.parse_args()
hdr, ins = None, []
for r in csv.reader(open(a.csv)):
    if r and r[0] == "Address":
        if hdr is not None:   # the page lists the kernel once per source view: keep the first listing
            break
        hdr = r
        continue
    if hdr is None:
        continue
    try:
        ins.append((r[1].strip(), int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])))
    except (ValueError, IndexError):
        pass
tot, smp = sum(i[1] for i in ins), sum(i[2] for i in ins)
if a.title:
    print("# " + a.title)
print(f"# total warp instructions {tot:,}, stall samples {smp:,}\n#\n# (1) opcode mix weighted by executions")
mix = Counter()
for s, ex, _ in ins:
    t = s.split()
    op = t[1] if t[0].startswith("@") else t[0]
    mix[op.split(".")[0]] += ex
for op, ex in mix.most_common(24):
    print(f"#   {op:10s} {100 * ex / tot:5.1f} %")
fmt = lambda i: f"  [{i:4d}] {ins[i][0]:94s} exec {ins[i][1]:10d}  samples {ins[i][2]}"
print("#\n# (2) the bulk-copy (TMA) / mbarrier instructions: cp.async.bulk -> UBLKCP.S.G, mbarrier.init -> SYNCS.EXCH.64,\n"
      "#     arrive.expect_tx -> SYNCS.ARRIVE.TRANS64, try_wait.parity -> SYNCS.PHASECHK.TRANS64.TRYWAIT, fence.proxy.async -> FENCE.VIEW.ASYNC")
for i, (s, _, _) in enumerate(ins):
    if any(k in s for k in ("UBLKCP", "SYNCS", "FENCE.VIEW", "UTMA")):
        print(fmt(i))
hits = [i for i, (s, _, _) in enumerate(ins) if a.window_at in s]
if hits:
    i0 = max(0, hits[min(a.nth, len(hits)) - 1] - a.before)
    print(f"#\n# (3) the hot loop around the {a.nth}-th {a.window_at}: {a.window} instructions with executions and stall samples")
    for i in range(i0, min(len(ins), i0 + a.window)):
        print(fmt(i))
