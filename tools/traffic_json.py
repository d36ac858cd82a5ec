#!/usr/bin/env python
"""ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv log of one pair-kernel launch -> a record in
profiles/pair_kernel_traffic.json (read by bench.py for roofline.traffic).
usage: tools/traffic_json.py gpurun_out/traffic.csv n_train n_test [label]"""
import csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, n_train, n_test = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
label = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(path)
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
rec = {}
for r in rows[1:]:
    d = dict(zip(hdr, r))
    m = re.search(r"pair_kernel<[^>]*>", d.get("Kernel Name", ""))
    if not m:
        continue
    val = float(d["Metric Value"].replace(",", ""))
    unit = d["Metric Unit"].lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
    rec["kernel"] = m.group(0)
    rec["dram_bytes_read" if "read" in d["Metric Name"] else "dram_bytes_write"] = val * mult
assert "dram_bytes_read" in rec and "dram_bytes_write" in rec, rec
rec.update({"n_train": n_train, "n_test": n_test, "source": label})
out = os.path.join(ROOT, "profiles", "pair_kernel_traffic.json")
recs = json.load(open(out)) if os.path.exists(out) else []
recs = [r for r in recs if not (r["n_train"] == n_train and r["n_test"] == n_test and r["kernel"] == rec["kernel"])] + [rec]
json.dump(recs, open(out, "w"), indent=1)
print(rec)
