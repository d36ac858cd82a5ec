#!/usr/bin/env python
"""Summarise the ncu launch list of tools/hbm_bench.py: per memory-bound kernel the algorithmic bytes of one launch,
its duration, the achieved GB/s (algorithmic and DRAM-measured) and the fraction of the measured HBM peak.
usage: tools/hbm_summary.py gpurun_out/hbm_kernels.csv N C K [hbm_peak_gbs] > profiles/rNx_hbm_kernels.md"""
import csv, json, os, re, sys
from collections import defaultdict

path, N, C, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = float(sys.argv[5]) if len(sys.argv) > 5 else json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
launches = defaultdict(dict)
order = []
for r in rows[1:]:
    d = dict(zip(hdr, r))
    lid = int(d["ID"])
    if lid not in launches:
        order.append(lid)
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    launches[lid]["name"] = name
    launches[lid]["grid"] = d["Grid Size"]
    val = float(d["Metric Value"].replace(",", ""))
    unit = d["Metric Unit"].lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(unit, 1)
    launches[lid][d["Metric Name"]] = val * mult
E = 8  # float64
ntiles = (C + 3) // 4
npairs = ntiles * (ntiles + 1) // 2
# algorithmic bytes per launch, keyed on kernel name and grid
def algorithmic(name, grid):
    gy = int(re.findall(r"\d+", grid)[1]) if grid else 1
    if name.startswith("colsum_kernel"):
        return None  # depends on d: resolved below from the launch order
    if name.startswith("colsum_all_kernel"):
        return E * N * C
    if name.startswith("gather_rows_kernel"):
        return (2 * E + 4) * N * C          # read value + int32 index, write value, per column
    if name.startswith("fold_gram_kernel"):
        return E * N * 8 * npairs            # 8 columns per 4 x 4 tile pair
    if name.startswith("lg_logl_kernel"):
        return E * N * 4 + 8 * N             # 4 columns in, one double out
    if name.startswith("sum_partial_kernel"):
        return 8 * N
    return None
seen = defaultdict(int)
out = []
for lid in order:
    L = launches[lid]
    name = L["name"]
    seen[name] += 1
    t = L.get("gpu__time_duration.sum")
    if not t or t < 20e-6:
        continue
    alg = algorithmic(name, L["grid"])
    gy = [int(x) for x in re.findall(r"\d+", L["grid"])]
    if name.startswith("cov_tile_kernel"):
        side = int(round(gy[1] ** 0.5))
        alg = E * N * 8 * (side * (side + 1) // 2)
        d = min(C, side * 4)
    if name.startswith("colsum_kernel"):
        # one pass over d columns (d = 8 for the moments call, 4 for the bandwidth / centring passes of the KDE fit):
        # d is read off the DRAM bytes of the launch
        d = max(1, int(round(L.get("dram__bytes_read.sum", 0) / (E * N))))
        alg = E * N * d
    if name.startswith("whiten_kernel"):
        alg = E * N * 4 * 2 + 8 * N         # 4 columns in, 4 whitened coordinates + the row norm out
    if alg is None:
        continue
    dram = L.get("dram__bytes_read.sum", 0) + L.get("dram__bytes_write.sum", 0)
    out.append((name, seen[name], L["grid"], t, alg, dram))
print("# Memory-bound helper kernels on B200 (tools/hbm_bench.py, N = %d rows, %d float64 columns, k = %d folds)" % (N, C, K))
print()
print("ncu `gpu__time_duration.sum` / `dram__bytes_*` per launch (`--clock-control none`; ncu serialises launches and")
print("starts each with a cold L2, so these are the conservative numbers).  Peak = %.0f GB/s (MEASURED_PEAKS.json)." % peak)
print()
print("| kernel | launch | grid | ms | algorithmic MB | DRAM MB (ncu) | GB/s (algorithmic) | frac of HBM peak | GB/s (DRAM) |")
print("|---|---|---|---|---|---|---|---|---|")
for name, k, grid, t, alg, dram in out:
    print("| `%s` | %d | %s | %.3f | %.1f | %.1f | %.0f | %.2f | %.0f |" % (name, k, grid, t * 1e3, alg / 1e6, dram / 1e6, alg / t / 1e9, alg / t / 1e9 / peak, dram / t / 1e9))
