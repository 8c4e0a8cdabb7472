"""Shared helpers of the GPU parity tests: run the same program / kernel / Philox stream through the
CUDA engine (C ABI) and through the CPU oracle, and compare per iteration."""
import numpy as np

import oracle as O
import quicksand_b200 as qs
from quicksand_b200 import _lib as L


def oracle_specs(kernel):
    return [O.spec(s.kind, stepSize=s.stepSize, numSteps=int(s.numSteps), doAdapt=bool(s.doAdapt), scale=s.scale, weight=s.weight)
            for s in kernel.specs]


def oracle_program(p):
    a = p.arrays
    if p.family == L.FAM_INDEP:
        return O.Program.indep(a["erp_kind"], a["erp_p0"], a["erp_p1"], p.ret_mode, p.ret_div)
    if p.family == L.FAM_GAUSS_PREC:
        return O.Program.gauss_prec(a["A"])
    if p.family == L.FAM_LOGISTIC:
        return O.Program.logistic(a["X"], a["ybin"], p.prior_sigma)
    if p.family == L.FAM_EIGHT_SCHOOLS:
        return O.Program.eight_schools(a["y"], a["sigma"])
    return O.Program.funnel(p.dim)


def run_both(program, kernel, chains, iters, seed=1234, chain_begin=0, leap=False, mapping=0, burnin=0, lag=1):
    """Returns (gpu, cpu): dicts with state [C][T][n], accept, dH, step, kidx, leapfrog, samples, logprobs, init."""
    nsamps = (iters - burnin) // lag
    assert burnin + nsamps * lag == iters
    flags = L.FLAG_RECORD | (L.FLAG_RECORD_LEAP if leap else 0)
    s = qs.Sampler(program, kernel, numchains=chains, seed=seed, chain_begin=chain_begin, flags=flags, mapping=mapping)
    s.init()
    init = s.get_state()
    s.run(nsamps, burnin, lag)
    gpu = s.fetch_record(leap=leap)
    gpu["init"] = init
    smp = s.fetch_samples()
    gpu["samples"] = smp["value"]
    gpu["logprobs"] = smp["logprob"]
    gpu["stats"] = s.stats()
    s.close()
    cpu = O.run(oracle_program(program), oracle_specs(kernel), seed, chains, nsamps, burnin, lag, chain_begin=chain_begin, record=True)
    return gpu, cpu


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1.0)
    return np.abs(a - b) / scale


def compare(gpu, cpu, tol=1e-10, min_agree=0.999, leap=False, what=""):
    """north_star gates: positions within `tol` relative; accept decisions identical on >= min_agree of steps.
    A chain is compared up to (and including) its first differing accept decision -- after a flip the two
    chains are different Markov chains and no longer comparable."""
    acc_g, acc_c = gpu["accept"], cpu["accept"]
    C, T = acc_g.shape
    differ = acc_g != acc_c
    first = np.where(differ.any(axis=1), differ.argmax(axis=1), T)           # first flipped iteration per chain
    steps = int(sum(min(f + 1, T) for f in first))
    agree = 1.0 - float((first < T).sum()) / max(steps, 1)
    valid = np.arange(T)[None, :] < first[:, None]                            # iterations strictly before the flip
    e_init = relerr(gpu["init"], cpu["init"]).max()
    e_state = relerr(gpu["state"], cpu["state"])[valid].max() if valid.any() else 0.0
    e_dh = 0.0
    finite = valid & np.isfinite(cpu["dH"]) & np.isfinite(gpu["dH"])
    if finite.any():
        e_dh = (np.abs(gpu["dH"] - cpu["dH"])[finite] / np.maximum(1.0, np.abs(cpu["dH"][finite]))).max()
    e_leap = 0.0
    if leap:
        e_leap = relerr(gpu["leapfrog"], cpu["leapfrog"])[valid].max()
    msg = "%s init %.2e state %.2e dH %.2e leap %.2e agree %.5f (%d/%d chain-steps compared)" % (
        what, e_init, e_state, e_dh, e_leap, agree, steps, C * T)
    print(msg)
    assert e_init <= tol, msg
    assert e_state <= tol, msg
    assert e_leap <= tol, msg
    assert agree >= min_agree, msg
    assert steps >= 0.9 * C * T, msg
    return dict(init=e_init, state=e_state, dH=e_dh, leap=e_leap, agree=agree)
