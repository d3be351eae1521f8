"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, ms, share.

usage: python tools/summarise_launches.py launches.csv "<command that was profiled>" > summary.json"""
import csv, json, re, sys

rows = []
with open(sys.argv[1], newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    rows.append((r["Kernel Name"], ms))

def short(name):
    m = re.search(r"gemm_nt(?:_tma)?_kernel<.*?TileCfg<(\d+), (\d+).*?(Epi\w+(?:<[^>]*>)?)", name)
    if m:
        return "gemm_nt_kernel<%sx%s,%s>" % (m.group(1), m.group(2), m.group(3).replace("stpyb::", ""))
    return re.sub(r"\(.*", "", name).replace("stpyb::", "").replace("(anonymous namespace)::", "")

agg = {}
for name, ms in rows:
    k = short(name)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
total = sum(ms for _, ms in rows)
out = {"command": sys.argv[2] if len(sys.argv) > 2 else "", "launches": len(rows), "total_ms": total,
       "kernels": [{"kernel": k, "launches": c, "ms": ms, "share": ms / total}
                   for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
json.dump(out, sys.stdout, indent=1)
