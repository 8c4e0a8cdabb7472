#!/usr/bin/env python
"""Is the posterior-variance gap between c3's adaptive HMC (the reference's never-ending dual averaging) and fixed-step HMC a
property of the algorithm?  Three arms on the device at the full shape: adaptive, fixed eps = 0.04, fixed eps = 0.025 (longer)."""
import json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quicksand_b200 as qs

X, y, theta = qs.programs.c3_logistic_data(10000, 100)
prog = qs.programs.logistic(X, y, 2.5)

def arm(kernel, chains, nsamps, burnin, seed):
    s = qs.Sampler(prog, kernel, numchains=chains, seed=seed); s.init(); s.run(nsamps, burnin, 1)
    v = s.fetch_values(); st = s.stats(); s.close()
    return v, st["props_accepted"] / st["props_made"]

res = {}
for name, k, C, n, b, seed in [("adaptive", qs.HMCKernel(numSteps=20), 8192, 60, 150, 20261019),
                               ("adaptive_long", qs.HMCKernel(numSteps=20), 4096, 60, 600, 5),
                               ("fixed0.04", qs.HMCKernel(numSteps=20, stepSize=0.04, doStepSizeAdapt=False), 4096, 60, 400, 7),
                               ("fixed0.025", qs.HMCKernel(numSteps=20, stepSize=0.025, doStepSizeAdapt=False), 4096, 60, 800, 9)]:
    v, acc = arm(k, C, n, b, seed)
    flat = v.reshape(-1, 100)
    res[name] = dict(mean=flat.mean(axis=0), var=flat.var(axis=0), acc=acc)
    print(name, "accept %.3f" % acc, "mean var over dims %.5g" % flat.var(axis=0).mean(), flush=True)
base = res["fixed0.025"]
for name in res:
    r = res[name]["var"] / base["var"]
    dm = (res[name]["mean"] - base["mean"]) / np.sqrt(base["var"])
    print(json.dumps({"arm": name, "var_ratio_vs_fixed0.025": [float(r.min()), float(np.median(r)), float(r.max())],
                      "mean_shift_in_posterior_sd": [float(dm.min()), float(np.median(dm)), float(dm.max())], "accept": res[name]["acc"]}))
