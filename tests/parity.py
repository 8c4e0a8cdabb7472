"""Shared helpers of the GPU parity tests: run the same program / kernel / Philox stream through the
CUDA engine (C ABI) and through the CPU oracle, and compare per iteration."""
import numpy as np

import oracle as O
import quicksand_b200 as qs
from quicksand_b200 import _lib as L


def oracle_specs(kernel):
    return [O.spec(s.kind, stepSize=s.stepSize, numSteps=int(s.numSteps), doAdapt=bool(s.doAdapt), scale=s.scale, weight=s.weight,
                   lexicalScaleSharing=bool(s.lexicalScaleSharing)) for s in kernel.specs]


def oracle_program(p):
    a = p.arrays
    if p.family == L.FAM_INDEP:
        return O.Program.indep(a["erp_kind"], a["erp_p0"], a["erp_p1"], p.ret_mode, p.ret_div, lexids=a.get("erp_lexid"),
                               fmu=a.get("erp_fmu"), fsigma=a.get("erp_fsigma"),
                               pref=a.get("erp_pref"), pcoef=a.get("erp_pcoef"), ptf=a.get("erp_ptf"))
    if p.family == L.FAM_GAUSS_PREC:
        return O.Program.gauss_prec(a["A"])
    if p.family == L.FAM_LOGISTIC:
        return O.Program.logistic(a["X"], a["ybin"], p.prior_sigma)
    if p.family == L.FAM_EIGHT_SCHOOLS:
        return O.Program.eight_schools(a["y"], a["sigma"])
    return O.Program.funnel(p.dim)


def run_oracle(program, kernel, chains, iters, seed=1234, chain_begin=0, burnin=0, lag=1, variant="strict"):
    nsamps = (iters - burnin) // lag
    O.use(variant)
    try:
        return O.run(oracle_program(program), oracle_specs(kernel), seed, chains, nsamps, burnin, lag, chain_begin=chain_begin, record=True)
    finally:
        O.use("strict")


def run_both(program, kernel, chains, iters, seed=1234, chain_begin=0, leap=False, mapping=0, burnin=0, lag=1, temps=None):
    """Returns (gpu, cpu): dicts with state [C][T][n], accept, dH, step, kidx, leapfrog, samples, logprobs, init."""
    nsamps = (iters - burnin) // lag
    assert burnin + nsamps * lag == iters
    flags = L.FLAG_RECORD | (L.FLAG_RECORD_LEAP if leap else 0)
    s = qs.Sampler(program, kernel, numchains=chains, seed=seed, chain_begin=chain_begin, flags=flags, mapping=mapping)
    s.init()
    init = s.get_state()
    if temps is not None:
        s.set_temperatures(temps)
    s.run(nsamps, burnin, lag)
    gpu = s.fetch_record(leap=leap)
    gpu["init"] = init
    smp = s.fetch_samples()
    gpu["samples"] = smp["value"]
    gpu["logprobs"] = smp["logprob"]
    gpu["stats"] = s.stats()
    s.close()
    if temps is None and kernel.anneal is not None:
        temps = [float(kernel.anneal(i, iters)) for i in range(iters)]
    cpu = O.run(oracle_program(program), oracle_specs(kernel), seed, chains, nsamps, burnin, lag, chain_begin=chain_begin, record=True,
                temps=temps)
    return gpu, cpu


def relerr(a, b):
    """Relative error against max(|a|, |b|, 1); entries where both sides are non-finite in the same way count as equal."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        scale = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1.0)
        e = np.abs(a - b) / scale
    both_bad = ~np.isfinite(a) & ~np.isfinite(b)
    e = np.where(both_bad, 0.0, e)
    return np.where(np.isnan(e), np.inf, e)


def compare(gpu, cpu, tol=1e-10, min_agree=0.999, leap=False, what="", min_compared=1.0, check=True):
    """north_star gates: per-step positions within `tol` relative, accept decisions identical on >= min_agree
    of the compared steps.

    A chain is compared up to its first *divergence event*: an accept decision that differs, or a state
    whose error exceeds `tol`.  After either, the two runs are different Markov chains and later steps are
    not comparable.  With min_compared == 1 (stable dynamics: fixed step sizes, drift, HARM) no divergence
    by error is tolerated at all.  With min_compared < 1 (the reference's step-size search and dual
    averaging deliberately drive HMC through numerically unstable step sizes: an unstable leapfrog
    trajectory amplifies a 1-ulp difference -- FMA contraction, libm -- by (eps/sigma)^(2L), so two correct
    implementations separate) the test requires that at least that fraction of chain-steps was compared
    and matched, and that accept flips stay below 1 - min_agree of them."""
    acc_g, acc_c = gpu["accept"], cpu["accept"]
    C, T = acc_g.shape
    e_state_all = relerr(gpu["state"], cpu["state"]).max(axis=2)             # [C][T]
    if leap:
        e_leap_all = relerr(gpu["leapfrog"], cpu["leapfrog"]).reshape(C, T, -1).max(axis=2)
    else:
        e_leap_all = np.zeros((C, T))
    flip = acc_g != acc_c
    bad = flip | (e_state_all > tol) | (e_leap_all > tol)
    first = np.where(bad.any(axis=1), bad.argmax(axis=1), T)                  # first divergence event per chain
    valid = np.arange(T)[None, :] < first[:, None]
    compared = int(valid.sum())
    flips = int(sum(1 for c in range(C) if first[c] < T and flip[c, first[c]]))
    agree = 1.0 - flips / max(compared + flips, 1)
    e_init = relerr(gpu["init"], cpu["init"]).max()
    e_state = e_state_all[valid].max() if compared else 0.0
    e_leap = e_leap_all[valid].max() if compared else 0.0
    with np.errstate(invalid="ignore"):
        e_dh_all = np.abs(gpu["dH"] - cpu["dH"]) / np.maximum(1.0, np.abs(cpu["dH"]))
    fin = valid & np.isfinite(e_dh_all)
    e_dh = e_dh_all[fin].max() if fin.any() else 0.0
    frac = compared / float(C * T)
    msg = "%s init %.2e state %.2e dH %.2e leap %.2e agree %.5f compared %.4f (%d/%d chain-steps, %d flips)" % (
        what, e_init, e_state, e_dh, e_leap, agree, frac, compared, C * T, flips)
    print(msg)
    if check:
        assert e_init <= tol, msg
        assert agree >= min_agree, msg
        assert frac >= min_compared - 1e-12, msg
    return dict(init=e_init, state=e_state, dH=e_dh, leap=e_leap, agree=agree, compared=frac, valid=valid)
