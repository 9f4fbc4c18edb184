"""Builds profiles/ncu_dram_bytes.json -- dram__bytes_read.sum + dram__bytes_write.sum per launch, by the kernel names bench.py's per-kernel table
uses -- from an ncu launch list made by tools/ncu_capture.sh (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum).
  python tools/ncu_dram_table.py gpurun_out/r04_launches_ncu.csv profiles/r04_launches_ncu.csv
The second argument is the path the launch list is committed under (recorded as the provenance of every entry)."""
import csv
import json
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bench_name(kernel, grid, half_grid):
    base = kernel.split("(")[0].replace("void ", "").replace("qpic::", "")
    t = re.sub(r"\s", "", base)
    if t.startswith("k_particles_predict"):
        return "k_particles_predict"
    if t.startswith("k_tile_cell_deposit") or t.startswith("k_tile_list_deposit"):
        return "k_tile_deposit_jy_jz"
    if t.startswith("k_tile_assemble") or t.startswith("k_tile_stay_deposit"):
        return "k_tile_deposit_charge_mid"
    if t.startswith("k_tile_fold_rho_chi"):
        return "k_tile_fold_rho_chi"
    if t.startswith("k_tile_deposit<2"):
        return "k_tile_deposit_rho_chi"
    if t.startswith("k_tile_deposit<1"):
        return "k_tile_deposit_charge"
    if t.startswith("k_rank_to_slots"):
        return "k_particles_to_slots"
    if t.startswith("k_p2_cols_solve") or t.startswith("k_fft_cols_solve"):
        return "k_fft_cols_solve" if grid == half_grid else "k_fft_cols_solve_b"
    m = re.match(r"k_(p2|fft)_rows_(r2c|c2r|b_forward|b_backward|c2c)", t)
    if m:
        return "k_fft_rows_" + m.group(2)
    return t.split("<")[0]


def main():
    src, committed = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    iid, iname, igrid, imetric, ivalue, iunit = (hdr.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Value", "Metric Unit"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    launches = {}
    for r in rows[1:]:
        L = launches.setdefault(r[iid], {"name": r[iname], "grid": r[igrid], "bytes": 0.0})
        if r[imetric] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            L["bytes"] += float(r[ivalue].replace(",", "")) * scale.get(r[iunit], 1.0)
    # the column kernel serves the half spectrum of the real solves and the full spectrum of the By + i Bz solve: two launch shapes, told apart
    # by their mean traffic (the full spectrum moves twice the bytes)
    by_grid = defaultdict(list)
    for L in launches.values():
        if "cols_solve" in L["name"]:
            by_grid[(L["grid"], L.get("block"))].append(L["bytes"])
    order = sorted(by_grid, key=lambda g: sum(by_grid[g]) / len(by_grid[g]))
    half = order[0][0] if order else None
    agg = defaultdict(list)
    for L in launches.values():
        agg[bench_name(L["name"], L["grid"], half)].append(L["bytes"])
    out = {}
    for k, v in sorted(agg.items()):
        # the first launch of a kernel in the process pays its cold start; use the median
        v = sorted(v)
        out[k] = {"bytes": v[len(v) // 2], "launches": len(v), "source": committed}
    path = os.path.join(ROOT, "profiles", "ncu_dram_bytes.json")
    json.dump(out, open(path, "w"), indent=1)
    for k, e in out.items():
        print(f"{k:32s} {e['bytes'] / 1e6:9.2f} MB  x{e['launches']}")


if __name__ == "__main__":
    main()
