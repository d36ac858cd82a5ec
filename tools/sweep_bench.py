"""BASELINE.json configs[4]: KDE slogl sweep, n_train = n_test = N, d = 1..8,
float32 and float64, i.i.d. N(0, I_d), NormalReferenceRule; pair-evals/s of the pair kernel (CUDA events
inside the library) and its fraction of the FP64 / FP32+MUFU roofline of SURVEY §8(d).
usage: python tools/sweep_bench.py [--n 1000000] [--dims 1,2,4,8] [--dtypes float64,float32] [--modes off,on] [--json out.json]
`--modes`: tile skipping off (every pair evaluated: the roofline figures) and / or on (the library default).
Under torchrun the test rows of every point are sharded over the ranks (API-level sharding, parallel.py); as a plain
process with several GPUs visible (and PBN_CUDA_DEVICE unset) the in-process multi-device context shards them.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", default="1000000")
    ap.add_argument("--dims", default="1,2,3,4,5,6,7,8")
    ap.add_argument("--dtypes", default="float64,float32")
    ap.add_argument("--json", default="")
    ap.add_argument("--modes", default="off")
    ap.add_argument("--reps", type=int, default=1, help="timed calls per point; the fastest is reported")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        lr = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(lr)
        os.environ["PBN_CUDA_DEVICE"] = str(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import util_data
    import pybnesian_b200 as pbn
    ctx = pbn.default_context()
    f_hz, sms = 1.965e9, ctx.sm_count
    rows = []
    for n in [int(x) for x in a.n.split(",")]:
        for dt in a.dtypes.split(","):
            for d in [int(x) for x in a.dims.split(",")]:
                tr = pbn.DataFrame(util_data.iid_normal(n, d, 0, dt))
                te = pbn.DataFrame(util_data.iid_normal(n, d, 1, dt))
                k = pbn.KDE(list(tr.columns))
                k.fit(tr)
                warm = pbn.DataFrame(util_data.iid_normal(min(n, 65536), d, 2, dt))
                for mode in a.modes.split(","):
                    ctx.set_skipping(mode == "on")
                    k.slogl(warm)  # loads the kernels; the timed call below is the only full-size one
                    ctx.set_timing(True)
                    ctx.pair_kernel_time(reset=True)
                    ctx.skip_stats(reset=True)
                    wall, ms, pe = float("inf"), 0.0, 0
                    for _ in range(max(1, a.reps)):
                        ctx.pair_kernel_time(reset=True)
                        ctx.skip_stats(reset=True)
                        t0 = time.perf_counter()
                        s = k.slogl(te)
                        w = time.perf_counter() - t0
                        if w < wall:
                            wall = w
                            ms, nl, pe = ctx.pair_kernel_time(reset=False)
                            st = ctx.skip_stats(reset=False)
                    ctx.pair_kernel_time(reset=True)
                    ctx.skip_stats(reset=True)
                    ctx.set_timing(False)
                    if dt == "float64":
                        peak_survey = sms * 64 * f_hz / (2 * d + 18)
                        # dot-product form with the hoisted test-row norm (pair_kernel.cuh: tile_f64_dot): d FMA + 6
                        # (table exp2 with a degree-2 polynomial: 3 DADD + 2 DFMA, + 1 DFMA to accumulate)
                        peak_own = sms * 64 * f_hz / (d + 6)
                    else:
                        # FP32 pipe: d sub + d FMA + 1 add lane-ops per pair at 128 lanes/clk/SM; MUFU.EX2 16/clk/SM
                        peak_survey = min(sms * 128 * f_hz / (2 * d + 2), sms * 16 * f_hz)
                        peak_own = min(sms * 128 * f_hz / (2 * d + 1), sms * 16 * f_hz)
                    ndev = max(world, getattr(ctx, "num_devices", 1))
                    rate = pe / (ms * 1e-3) if ms > 0 else float("nan")
                    row = {"n": n, "d": d, "dtype": dt, "n_gpus": ndev, "tile_skipping": mode, "wall_s": wall,
                           "pair_evals_per_s_job_wall": float(n) * n / wall, "slogl": s, "fallback_rows": ctx.last_fallback_rows(),
                           "units_evaluated_fraction": st["timed_evaluated"] / max(st["timed_total"], 1)}
                    if mode == "off":   # roofline figures only make sense when every pair is evaluated
                        row.update({"pair_evals_per_s_per_gpu_kernel": rate, "frac_survey_roofline": rate / peak_survey,
                                    "frac_own_roofline": rate / peak_own})
                    rows.append(row)
                    if int(os.environ.get("RANK", "0")) == 0:
                        print(json.dumps(rows[-1]), flush=True)
                ctx.set_skipping(True)
                del k, tr, te
    if a.json and int(os.environ.get("RANK", "0")) == 0:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
