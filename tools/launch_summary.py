"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) of bench.py:
per kernel the number of launches, the mean device time and its share of one step.
    python tools/launch_summary.py gpurun_out/launches.csv --steps 3 > profiles/rNN_launches.txt
`--steps` = warm-up + timed steps the profiled command ran (kernels launched once per step are divided by it)."""
import argparse
import csv
import re
from collections import defaultdict

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--steps", type=int, required=True)
ap.add_argument("--title", default="")
a = ap.parse_args()
rows = [r for r in csv.reader(l for l in open(a.csv) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
t = defaultdict(list)
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
    ns = float(r[iv].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[iu], 1.0)
    t[name].append(ns)
per_step = {k: sum(v) / a.steps for k, v in t.items() if len(v) >= a.steps}
step_ns = sum(per_step.values())
if a.title:
    print("# " + a.title)
print(f"# {len(rows) - 1} launches; kernels launched at least once per step sum to {step_ns / 1e6:.3f} ms per step "
      f"(cold-cache, serialised: shares are meaningful, absolute times are not bench values)")
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    share = f"{100 * per_step[k] / step_ns:5.1f} %" if k in per_step else "  (setup)"
    print(f"{k:44s} launches {len(v):3d}  mean {sum(v) / len(v) / 1e6:8.4f} ms  share of one step {share}")
