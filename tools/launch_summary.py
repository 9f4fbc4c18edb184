"""Aggregates an ncu launch list (ncu --metrics gpu__time_duration.sum[,smsp__inst_executed.sum] --csv --log-file X.csv) per kernel:
launches, mean duration, share of the summed kernel time, mean warp instructions.  python tools/launch_summary.py X.csv [slices]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
iname, imetric, ivalue, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
slices = float(sys.argv[2]) if len(sys.argv) > 2 else None
agg = OrderedDict()
launch_ids = set()
for r in rows[1:]:
    name = r[iname].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, {"n": 0, "ns": 0.0, "inst": 0.0})
    v = float(r[ivalue].replace(",", ""))
    if r[imetric] == "gpu__time_duration.sum":
        a["n"] += 1
        a["ns"] += v
        launch_ids.add(r[iid])
    elif r[imetric] == "smsp__inst_executed.sum":
        a["inst"] += v
total = sum(a["ns"] for a in agg.values())
print(f"{len(launch_ids)} launches, summed kernel time {total / 1e3:.1f} us" + (f" = {total / 1e3 / slices:.1f} us per slice (ncu: serialised, cold caches)" if slices else ""))
print(f"{'kernel':44s} {'launches':>8s} {'mean us':>9s} {'share':>7s} {'warp inst / launch':>19s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    print(f"{name:44s} {a['n']:8d} {a['ns'] / a['n'] / 1e3:9.2f} {a['ns'] / total:7.1%} {a['inst'] / max(a['n'], 1):19.0f}")
