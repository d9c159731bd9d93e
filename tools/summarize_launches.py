"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a per-kernel table for ONE training
step (markdown).  usage: python tools/summarize_launches.py gpurun_out/launches.csv [step_index [marker_kernel]] > profiles/xxx.md"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hdr, data = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
marker = sys.argv[3] if len(sys.argv) > 3 else "counts_to_bf16_kernel"  # launched exactly twice per step (once per group)
start = [i for i, d in enumerate(data) if marker in d["Kernel Name"]]
i0, i1 = start[2 * step], start[2 * step + 2]
agg = collections.OrderedDict()
tot = 0.0
for d in data[i0:i1]:
    n = re.sub(r"\(.*", "", d["Kernel Name"])
    n = re.sub(r"^void ", "", n).replace("<unnamed>::", "")
    v = float(d["Metric Value"].replace(",", "")) / 1000
    tot += v
    key = (n, d["Grid Size"], d["Block Size"])
    agg.setdefault(key, [0, 0.0])
    agg[key][0] += 1
    agg[key][1] += v
print(f"one training step (both groups): {i1 - i0} launches, {tot:.1f} us summed kernel time (ncu: serialised, cold caches)\n")
print("| us / step | share | launches | kernel | grid | block |")
print("|---:|---:|---:|---|---|---|")
for (n, g, b), (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| {t:.1f} | {100 * t / tot:.1f}% | {c} | `{n}` | {g} | {b} |")
