"""Workloads for the ncu captures of the round-2 kernels that bench.py does not launch:
   WHICH=gskip : CKDE d=4 float64, 1M training x 262144 test rows, tile skipping on -> pair_kernel<double, 4, 1, 0, 0, 1> (pass B)
   WHICH=soft  : KDE d=1 float32, 1M x 262144, all pairs -> pair_kernel<float, 1, 0, 0, 0, 0> with the MUFU offload
Each runs the call twice (the first loads the modules).  Not part of the product or the tests."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util_data
import pybnesian_b200 as pbn

which = os.environ.get("WHICH", "gskip")
ctx = pbn.default_context()
if which == "gskip":
    tr = util_data.generate_normal_data(1_000_000, 0); te = util_data.generate_normal_data(262_144, 1)
    f = pbn.CKDE("d", ["a", "b", "c"]); ctx.set_skipping(True)
else:
    tr = util_data.iid_normal(1_000_000, 1, 0, "float32"); te = util_data.iid_normal(262_144, 1, 1, "float32")
    f = pbn.KDE(list(tr.columns)); ctx.set_skipping(False)
ftr, fte = pbn.DataFrame(tr), pbn.DataFrame(te)
f.fit(ftr)
for _ in range(2):
    print(which, f.slogl(fte), ctx.skip_stats())
