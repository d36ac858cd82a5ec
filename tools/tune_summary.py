"""Prints gpurun_out/tune.log (tools/tune_all.sh) as one throughput row and one slogl row per build."""
import json, os, sys
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join("gpurun_out", "tune.log")
for line in open(path):
    name, _, js = line.partition(" ")
    try:
        d = json.loads(js)
    except Exception:
        print(line.rstrip())
        continue
    print("%-24s" % name, " ".join("%s %s" % (k[:-3], v) for k, v in d.items() if not k.endswith("slogl")))
    print(" " * 24, " ".join(v for k, v in d.items() if k.endswith("slogl")))
