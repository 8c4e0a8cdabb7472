#!/usr/bin/env python
"""The whole box from ONE host thread through qsb_group (the shape the single-threaded Terra host uses): BASELINE workloads
sharded over every visible GPU by the library itself, wall-clock timed around group.advance / group.sync (no torchrun).

    python tools/group_bench.py [c3|c4|c5] [--devices N] [--steps K]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as Bn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="c3")
    ap.add_argument("--devices", type=int, default=None)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    import torch
    import quicksand_b200 as qs
    from quicksand_b200 import diagnostics as D
    nd = a.devices or torch.cuda.device_count()
    wl = Bn.workload(a.workload)
    total = wl["chains"] if a.workload == "c3" else wl["chains"] * nd      # c3: BASELINE's 8 192 chains in total; c4/c5: their 8-GPU totals / 8 per device
    burn = {"c3": 150, "c2": 300, "c4": 300, "c5": 100}[a.workload]
    g = qs.Group(wl["prog"], wl["kernel"], list(range(nd)), numchains=total, seed=Bn.SEED, sample_chains=min(total // nd, 2048) * nd if a.workload != "c3" else 0)
    g.init()
    Kd = 200
    g.begin(Kd, burnin=burn + 3 + a.steps, lag=1, numiters=burn + 3 + a.steps + Kd)
    g.advance(burn + 3); g.sync()
    s0 = g.stats()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        g.advance(1)
    g.sync()
    t = time.perf_counter() - t0
    s1 = g.stats()
    key = "props_made" if a.workload == "c4" else "leapfrog_steps"
    g.advance(Kd); g.sync()
    raw = g.diagnostics_raw(maxlag=50)
    d = D.summarize(raw, wl["prog"].retdim, 50)
    print(json.dumps({"workload": wl["desc"], "api": "qsb_group (one host thread, %d devices)" % nd, "chains_total": total, "steps": a.steps,
                      "value": (s1[key] - s0[key]) / t, "unit": wl["unit"], "ms_per_step_wall": 1e3 * t / a.steps,
                      "max_split_rhat_all_devices": float(np.nanmax(d["rhat"])), "timing": "host wall clock around advance x steps + sync"}))
    g.close()


if __name__ == "__main__":
    main()
