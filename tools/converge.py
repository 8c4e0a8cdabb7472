#!/usr/bin/env python
"""Convergence exploration for the c4 / c5 workloads: split-R-hat over successive windows of a long run,
for a list of kernel settings, at a reduced chain count (a chain's convergence does not depend on how many others run).

    python tools/converge.py c4 [--chains 2048] [--chunk 5000] [--chunks 10]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def settings(qs, name):
    if name == "c4":
        y, sigma = qs.programs.c4_eight_schools_data(1000)
        prog = qs.programs.eight_schools(y, sigma)
        return prog, [("drift scale=0.02 adapt", qs.DriftKernel(scale=0.02)),
                      ("drift scale=0.1 adapt", qs.DriftKernel(scale=0.1)),
                      ("drift scale=0.1 fixed", qs.DriftKernel(scale=0.1, doScaleAdapt=False)),
                      ("drift scale=0.05 fixed", qs.DriftKernel(scale=0.05, doScaleAdapt=False)),
                      ("drift scale=0.1 lexical adapt", qs.DriftKernel(scale=0.1, lexicalScaleSharing=True))]
    prog = qs.programs.funnel(1000)
    mk = lambda L, **kw: qs.MixtureKernel([qs.HARMKernel(), qs.HMCKernel(numSteps=L, **kw)], [0.5, 0.5])
    return prog, [("HARM+HMC L=16 adaptive", mk(16)), ("HARM+HMC L=16 eps=0.05 fixed", mk(16, stepSize=0.05, doStepSizeAdapt=False)),
                  ("HARM+HMC L=16 eps=0.02 fixed", mk(16, stepSize=0.02, doStepSizeAdapt=False)),
                  ("HARM+HMC L=64 adaptive", mk(64))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--chains", type=int, default=2048)
    ap.add_argument("--chunk", type=int, default=5000)
    ap.add_argument("--chunks", type=int, default=10)
    ap.add_argument("--draws", type=int, default=100)
    ap.add_argument("--only", type=int, default=None)
    a = ap.parse_args()
    import quicksand_b200 as qs
    from quicksand_b200 import diagnostics as D
    prog, sets = settings(qs, a.workload)
    for si, (name, kernel) in enumerate(sets):
        if a.only is not None and si != a.only:
            continue
        s = qs.Sampler(prog, kernel, numchains=a.chains, seed=20261019)
        s.init()
        lag = max(a.chunk // a.draws, 1)
        it = 0
        prev = s.stats()
        for c in range(a.chunks):
            s.begin(a.draws, burnin=0, lag=lag, numiters=a.draws * lag)
            s.advance(a.draws * lag)
            it += a.draws * lag
            raw = s.diagnostics_raw(maxlag=20)
            d = D.summarize(raw, prog.retdim, 20)
            st = s.stats()
            acc = (st["props_accepted"] - prev["props_accepted"]) / max(st["props_made"] - prev["props_made"], 1)
            prev = st
            rh = d["rhat"]
            print(json.dumps({"workload": a.workload, "kernel": name, "iters": it, "max_rhat": float(np.nanmax(rh)), "median_rhat": float(np.nanmedian(rh)),
                              "rhat_dim0": float(rh[0]), "rhat_dim1": float(rh[1]), "n_rhat_gt_1.05": int((rh > 1.05).sum()),
                              "accept": acc, "min_ess_per_draw": float(np.nanmin(d["ess"]) / (a.chains * a.draws)),
                              "mean0": float(d["mean"][0]), "var0": float(d["var"][0]), "mean1": float(d["mean"][1]), "var1": float(d["var"][1])}), flush=True)
        s.close()


if __name__ == "__main__":
    main()
