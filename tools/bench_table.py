"""bench.py JSON line(s) -> markdown summary (profiles/r2_bench_final.md): python tools/bench_table.py ours.json [reference.json]"""
import json
import sys


def last_line(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


d = last_line(sys.argv[1])
ref = last_line(sys.argv[2]) if len(sys.argv) > 2 else None
print(f"# r2 — `python bench.py --steps {d['steps']} --warmup {d['warmup']}` on one B200 (clocks {d['clocks']['sm_mhz']:.0f} / {d['clocks']['sm_max_mhz']:.0f} MHz, "
      f"reasons {d['clocks']['reasons'] or 'none'}; wall {d['wall_s']} s)\n")
print("cells/s = cells of both groups per step / step time; `value` device-resident (one CUDA-graph replay per step), `e2e` through the plugin "
      "call with pinned host minibatches in scvi's float32 layout (H2D of the step's inputs and D2H of the loss inside the timed region).\n")
print("| workload | ms / step | value (cells/s) | e2e (cells/s) | e2e ms | launches / step | forward-only sweep | training sweep |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")


def row(name, ms, value, e2e, launches, roof):
    trn = [o for o in roof["other"] if "train" in o["kernel"] and o["bound"] == "hbm"]
    t = f"{trn[0]['avg_launch_ms'] * 1e3:.0f} us ({trn[0]['frac'] * 100:.0f} % HBM)" if trn else "-"
    print(f"| {name} | {ms:.4f} | {value:,.0f} | {e2e['value']:,.0f} | {e2e['ms_per_step']:.3f} | {launches} | "
          f"{roof['avg_launch_ms'] * 1e3:.0f} us ({roof['frac'] * 100:.1f} % HBM, {roof['other'][0]['frac'] * 100:.0f} % SFU) | {t} |")


row(d["config"]["workload"].split(":")[0] + " (headline)", d["ms_per_step"], d["value"], d["e2e"], d["launches_per_step"], d["roofline"])
for c in d["configs"]:
    if "value" in c:
        row(c["config"]["workload"].split(":")[0], c["ms_per_step"], c["value"], c["e2e"], c["launches_per_step"], c["roofline"])
    else:
        print(f"| {c.get('workload')} | {c.get('skipped') or c.get('error')} |")
print()
for k in ("e2e_uint16_input", "e2e_trainloop", "e2e_torch_adam"):
    if d.get(k):
        print(f"* `{k}`: {d[k]['value']:,.0f} cells/s ({d[k]['ms_per_step']:.3f} ms / step, {d[k]['h2d_bytes_per_step'] / 1e6:.0f} MB H2D per step)")
cb = d.get("cpu_baseline")
if cb:
    print(f"* `cpu_baseline` ({cb['kind']}, {cb['cores']} host threads): {cb['value']:,.0f} cells/s - {cb['sample']}")
if ref:
    print(f"* `--impl reference` (same box): {ref['value']:,.0f} cells/s ({ref['ms_per_step']:.0f} ms / step, {ref['cpu_baseline']['cores']} threads) "
          f"-> e2e ratio {d['e2e']['value'] / ref['value']:.0f}x, device-resident ratio {d['value'] / ref['value']:.0f}x")
r = d["roofline"]
print(f"\nRoofline of the headline workload (peak {r['peak']:.1f} GB/s, {r['peak_source']}): forward-only sweep {r['achieved']:.0f} GB/s algorithmic "
      f"({r['algorithmic_bytes_per_launch'] / 1e6:.1f} MB by SURVEY 8(d); {r['traffic'] / 1e6 if r.get('traffic') else float('nan'):.1f} MB measured DRAM traffic) = {r['frac']:.3f}.")
for o in r["other"]:
    print(f"* {o['kernel'][:60]}... [{o['bound']}]: {o['achieved']:.3g} of {o['peak']:.4g} {o['unit']} = {o['frac']:.3f} ({o['avg_launch_ms'] * 1e3:.0f} us)")
