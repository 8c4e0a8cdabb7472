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
                               pref=a.get("erp_pref"), pcoef=a.get("erp_pcoef"), ptf=a.get("erp_ptf"),
                               cond_lo=a.get("erp_cond_lo"), cond_hi=a.get("erp_cond_hi"))
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


TINY = 1e-300


def relerr(a, b, floor=TINY):
    """TRUE relative error |a - b| / max(|a|, |b|, floor), elementwise; entries where both sides are non-finite in the same
    way count as equal.  ``floor`` (a scalar or an array broadcastable against a) bounds the scale from below: the default
    only guards 0/0.  For a *vector* quantity (a chain's position) callers pass ``vector_floor``: components that sit
    below 1e-3 of the vector's largest are the result of cancellation and are held to 1e-10 of that size instead of to
    1e-10 of their own -- still a relative gate (it scales with the state), never an absolute one."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        scale = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
        e = np.abs(a - b) / scale
    both_bad = ~np.isfinite(a) & ~np.isfinite(b)
    e = np.where(both_bad, 0.0, e)
    return np.where(np.isnan(e), np.inf, e)


def vector_floor(ref, frac=1e-3):
    """Per-vector floor for relerr: frac * max_i |ref_i| over the last axis (kept for broadcasting)."""
    r = np.asarray(ref, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        m = np.nanmax(np.where(np.isfinite(r), np.abs(r), 0.0), axis=-1, keepdims=True)
    return np.maximum(frac * m, TINY)


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
    e_state_all = relerr(gpu["state"], cpu["state"], vector_floor(cpu["state"])).max(axis=2)             # [C][T]
    if leap:
        e_leap_all = relerr(gpu["leapfrog"], cpu["leapfrog"], vector_floor(cpu["leapfrog"])).reshape(C, T, -1).max(axis=2)
    else:
        e_leap_all = np.zeros((C, T))
    flip = acc_g != acc_c
    bad = flip | (e_state_all > tol) | (e_leap_all > tol)
    first = np.where(bad.any(axis=1), bad.argmax(axis=1), T)                  # first divergence event per chain
    valid = np.arange(T)[None, :] < first[:, None]
    compared = int(valid.sum())
    flips = int(sum(1 for c in range(C) if first[c] < T and flip[c, first[c]]))
    agree = 1.0 - flips / max(compared + flips, 1)
    e_init = relerr(gpu["init"], cpu["init"], vector_floor(cpu["init"])).max()
    e_state = e_state_all[valid].max() if compared else 0.0
    e_leap = e_leap_all[valid].max() if compared else 0.0
    with np.errstate(invalid="ignore"):
        # dH is a difference of two Hamiltonians (hundreds to thousands in magnitude for the large models): its error is
        # relative to them, not to itself; reported, not gated (the accept decisions are what is gated)
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


# ---------------------------------------------------------------------------------------------
def teacher_forced_hmc(program, kernel, chains, iters, seed=1234, chain_begin=0, tol=1e-10, what="", threads=8, div_dh=30.0, min_stable=0.15):
    """Teacher-forced parity of EVERY leapfrog step of an HMC run, adaptive or not (north_star: per-step leapfrog positions
    within 1e-10 relative, accept decisions identical).

    The oracle runs the chains and records, for every trajectory, the state entering it (positions, cached gradient,
    momenta, log-density, step size) and the positions / momenta / gradient after each HMCKernel:step.  Every one of those
    steps is then replayed on the device from the ORACLE's input state through qsb_debug_leapfrog -- the device functions
    the transition kernels use -- and compared with the oracle's output state.  No step is skipped: the chaos of the
    reference's adaptation transient cannot separate the two sides because each step starts from
    the same numbers.  The accept decision of every trajectory is re-derived from the device's final (p', lp') and the
    injected uniform and compared as well."""
    from concurrent.futures import ThreadPoolExecutor
    spec = kernel.specs[0]
    assert len(kernel.specs) == 1 and spec.kind == L.KERNEL_HMC
    Ls = int(spec.numSteps)
    n = program.dim

    def one(c):
        return O.run(oracle_program(program), oracle_specs(kernel), seed, 1, iters, 0, 1, chain_begin=chain_begin + c, record=True, traj=True)
    with ThreadPoolExecutor(max_workers=min(threads, chains)) as ex:      # ctypes drops the GIL; the oracle's state is thread-local
        outs = list(ex.map(one, range(chains)))
    cat = lambda k: np.concatenate([o[k] for o in outs], axis=0)
    x0, g0, p0, lp0 = cat("x0"), cat("g0"), cat("p0"), cat("lp0")
    pos, mom, grd, lps = cat("leapfrog"), cat("leap_mom"), cat("leap_grad"), cat("leap_lp")
    eps, dH, acc = cat("step"), cat("dH"), cat("accept")
    C, T = eps.shape
    xin = np.concatenate([x0[:, :, None, :], pos[:, :, :-1, :]], axis=2)      # [C][T][L][n]: state entering step s
    pin = np.concatenate([p0[:, :, None, :], mom[:, :, :-1, :]], axis=2)
    gin = np.concatenate([g0[:, :, None, :], grd[:, :, :-1, :]], axis=2)
    ein = np.broadcast_to(eps[:, :, None], (C, T, Ls))
    # A diverged trajectory of the adaptation transient may overflow.  A step is replayed when its oracle INPUT and OUTPUT are
    # finite; steps past an overflow have no numbers to compare (both sides carry inf / NaN, and the trajectory is rejected on
    # both -- the accept decisions below are still checked for EVERY trajectory).
    # Which steps are gated.  The reference's arithmetic itself degrades on DIVERGENT trajectories -- the naive sigmoid
    # 1/(1+exp(-z)) cancels in 1 - p once |z| > ~20, exp(v/2) overflows in the hierarchical programs and the tape's
    # zero-adjoint skipping (ad.t:166-172) then decides which terms survive -- so two correct evaluations of the same
    # gradient differ there (and the trajectory is rejected on both sides anyway).  A trajectory counts as divergent when the
    # ORACLE's energy error is non-finite or |dH| > div_dh = 30 (acceptance probability below e^-30, or an energy GAIN of
    # that size, which only the Jacobian-free first Hamiltonian of quirk Q3 produces; a stable leapfrog trajectory has |dH| of order 1;
    # measured: the first deviations above 1e-10 appear at |dH| ~ 77).  Every step of every
    # non-divergent trajectory must agree within tol: positions, momenta, gradient, log-density.  On divergent trajectories
    # the errors are reported and only the accept decision is gated (it is gated for EVERY trajectory).
    fin = np.isfinite(xin).all(axis=3) & np.isfinite(pin).all(axis=3) & np.isfinite(gin).all(axis=3)
    fin &= np.isfinite(pos).all(axis=3) & np.isfinite(mom).all(axis=3) & np.isfinite(grd).all(axis=3)
    fin_in_last = np.isfinite(xin[:, :, Ls - 1]).all(axis=2) & np.isfinite(pin[:, :, Ls - 1]).all(axis=2) & np.isfinite(gin[:, :, Ls - 1]).all(axis=2)
    with np.errstate(invalid="ignore"):
        stable = np.isfinite(dH) & (np.abs(dH) <= div_dh) & fin.all(axis=2)           # [C][T]
    gate = np.broadcast_to(stable[:, :, None], (C, T, Ls))
    res = {}
    worst = dict(x=0.0, p=0.0, g=0.0, lp=0.0)            # over the gated steps
    worst_div = dict(x=0.0, p=0.0, g=0.0)                # over the comparable steps of divergent trajectories (reported)
    nsteps = 0
    for last in (False, True):
        sel = np.zeros((C, T, Ls), dtype=bool)
        if last:
            sel[:, :, Ls - 1] = True
        else:
            sel[:, :, :Ls - 1] = True
        sel &= fin
        if last:
            sel[:, :, Ls - 1] |= fin_in_last      # the last step is replayed whenever it can be injected: it carries the accept decision
        if not sel.any():
            continue
        r = program.debug_leapfrog(xin[sel], pin[sel], ein[sel], g=gin[sel], last=last)
        nsteps += int(sel.sum())
        cmp = fin[sel]                               # rows of this batch whose oracle output is finite
        gated = gate[sel] & cmp
        for key, ref in (("x", pos), ("p", mom), ("g", grd)):
            e = relerr(r[key], ref[sel], vector_floor(ref[sel])).max(axis=1)
            if gated.any():
                worst[key] = max(worst[key], float(e[gated].max()))
                if float(e[gated].max()) > (10 * tol if key == "g" else tol):
                    bad = np.argwhere(sel)[gated][np.argmax(e[gated])]
                    print("   worst gated %s at (chain, iter, step) = %s: eps %.3g dH %.3g |x| %.3g |p| %.3g |g| %.3g err %.3g" % (
                        key, tuple(int(v) for v in bad), eps[bad[0], bad[1]], dH[bad[0], bad[1]], np.abs(pos[tuple(bad)]).max(), np.abs(mom[tuple(bad)]).max(),
                        np.abs(grd[tuple(bad)]).max(), e[gated].max()))
            rest = cmp & ~gated
            if rest.any():
                worst_div[key] = max(worst_div[key], float(e[rest].max()))
        if last:
            with np.errstate(invalid="ignore"):
                okl = gated & np.isfinite(lps[sel])
            if okl.any():
                worst["lp"] = float(relerr(r["lp"][okl], lps[sel][okl]).max())
            res["last_sel"] = sel[:, :, Ls - 1]
            res["p_last"], res["lp_last"] = r["p"], r["lp"]
    # accept decisions from the device's own final momentum and log-density, the oracle's H0 and the injected uniform
    flips = 0
    ntraj = 0
    e_dh = 0.0
    if "last_sel" in res:
        ls = res["last_sel"]
        H0 = -0.5 * (p0 * p0).sum(axis=2) + lp0
        with np.errstate(invalid="ignore", over="ignore"):
            dH_dev = (-0.5 * (res["p_last"] ** 2).sum(axis=1) + res["lp_last"]) - H0[ls]
        u = np.empty((C, T))
        one_u = np.empty(1)
        for c in range(C):
            for t in range(T):
                L.check(L.lib().qsb_uniforms(seed, chain_begin + c, t, n, 1, one_u.ctypes.data_as(L.dp)))
                u[c, t] = one_u[0]
        with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
            acc_dev = np.log(u[ls]) < dH_dev
        flips = int((acc_dev != (acc[ls] != 0)).sum())
        ntraj = int(ls.sum())
        with np.errstate(invalid="ignore"):
            scale = np.maximum(np.maximum(np.abs(H0[ls]), np.abs(dH[ls])), 1.0)     # dH is a difference of Hamiltonians of this size
            ed = np.abs(dH_dev - dH[ls]) / scale
        e_dh = float(np.nanmax(np.where(np.isfinite(ed), ed, 0.0))) if ed.size else 0.0
    n_stable = int(stable.sum())
    msg = ("%s teacher-forced: %d of %d trajectories non-divergent (|dH| <= %g) = %d leapfrog steps gated: x %.2e p %.2e g %.2e lp %.2e; "
           "divergent trajectories (reported): x %.2e p %.2e g %.2e; %d trajectories replayed to their accept decision, dH %.2e (rel. to H), "
           "accept flips %d; accept rate %.2f, eps range [%.3g, %.3g]; %d steps not comparable (oracle state overflowed)" % (
               what, n_stable, C * T, div_dh, n_stable * Ls, worst["x"], worst["p"], worst["g"], worst["lp"], worst_div["x"], worst_div["p"], worst_div["g"],
               ntraj, e_dh, flips, acc.mean(), eps.min(), eps.max(), C * T * Ls - int(fin.sum())))
    print(msg)
    assert n_stable >= min_stable * C * T, msg         # the gated set is the bulk of the run, not a corner of it
    assert worst["x"] <= tol and worst["p"] <= tol, msg
    assert worst["g"] <= 10 * tol and worst["lp"] <= tol, msg
    assert flips == 0, msg
    # a trajectory whose last step cannot even be injected has overflowed entirely: the oracle must have rejected it
    if "last_sel" in res:
        assert not (acc[~res["last_sel"]] != 0).any(), msg
    return dict(worst=worst, worst_divergent=worst_div, steps=nsteps, stable=n_stable, trajectories=C * T, flips=flips, accept_rate=float(acc.mean()), eps=eps)
