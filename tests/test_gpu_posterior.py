"""Statistical parity on the BASELINE configurations c3, c4 and c5 (north_star: posterior means and variances agree within
3 Monte-Carlo standard errors).

How a standard error is obtained here.  Chains are independent, so a per-chain statistic (the chain's time average of a
coordinate over the kept draws, or of its squared deviation) is an i.i.d. sample over chains and its ensemble mean has the
EXACT standard error sd/sqrt(chains) -- no autocorrelation model, no ESS estimate, valid whether or not a chain has mixed.
Two ensembles (device vs CPU oracle, or two device runs) of the SAME algorithm started from the SAME initial distribution
have the same law at every iteration, converged or not; so the comparison is sharp for c4 and c5 as well, whose reference
kernels do not reach their posterior at these shapes in any affordable run (profiles/r02_convergence_c{4,5}.jsonl).  The
oracle ensembles use chain ids disjoint from the device's (independent samples, not replicas of the same chains).

Gate over the 2 x dim simultaneous comparisons of a workload ("within 3 standard errors" read for a family of tests): the
number of |z| above 3 stays within the 99.9 % quantile of its binomial law under the null (Student t with the smaller
ensemble's degrees of freedom), and no |z| exceeds the Bonferroni bound at a family-wise 1e-4."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import parity as P

pytestmark = pytest.mark.gpu


def _chain_stats(values, centre):
    """values [C][S][dim] -> per-chain time means a [C][dim] and mean squared deviations from `centre` q [C][dim]."""
    a = values.mean(axis=1)
    q = ((values - centre[None, None, :]) ** 2).mean(axis=1)
    return a, q


def _z(a, b):
    """Two-sample z per column of the ensembles a [Ca][dim], b [Cb][dim]."""
    se = np.sqrt(a.var(axis=0, ddof=1) / a.shape[0] + b.var(axis=0, ddof=1) / b.shape[0])
    return (a.mean(axis=0) - b.mean(axis=0)) / np.maximum(se, 1e-300)


def _gate(za, zq, dof, what):
    from scipy import stats
    z = np.abs(np.concatenate([za, zq]))
    n3 = int((z > 3.0).sum())
    p3 = 2.0 * stats.t.sf(3.0, dof)
    allowed = int(stats.binom.ppf(0.999, z.size, p3))
    zmax = float(stats.t.isf(1e-4 / (2.0 * z.size), dof))
    msg = "%s: %d comparisons, max |z| mean %.2f variance %.2f (bound %.2f), %d above 3 (expected %.1f, allowed %d)" % (
        what, z.size, np.abs(za).max(), np.abs(zq).max(), zmax, n3, p3 * z.size, allowed)
    print(msg)
    assert np.isfinite(z).all(), msg
    assert z.max() < zmax and n3 <= allowed, msg


def _device_values(qs, prog, kernel, chains, nsamps, burnin, lag, seed, chain_begin=0):
    s = qs.Sampler(prog, kernel, numchains=chains, seed=seed, chain_begin=chain_begin)
    s.init(); s.run(nsamps, burnin, lag)
    v = s.fetch_values()
    st = s.stats()
    s.close()
    return v, st


def _oracle_values(oracle, prog, kernel, chains, nsamps, burnin, lag, seed, chain_begin, threads=16):
    op, specs = P.oracle_program(prog), P.oracle_specs(kernel)

    def one(c):        # ctypes releases the GIL; the oracle's tape is thread-local
        return oracle.run(op, specs, seed, 1, nsamps, burnin, lag, chain_begin=chain_begin + c)["samples"]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return np.concatenate(list(ex.map(one, range(chains))), axis=0)


def _compare(dev, ora, what):
    centre = np.concatenate([dev.reshape(-1, dev.shape[2]), ora.reshape(-1, ora.shape[2])]).mean(axis=0)
    ad, qd = _chain_stats(dev, centre)
    ao, qo = _chain_stats(ora, centre)
    _gate(_z(ad, ao), _z(qd, qo), min(dev.shape[0], ora.shape[0]) - 1, what)


def test_c3_posterior_moments_device_vs_oracle(qs, oracle):
    """config c3's program and kernel (Bayesian logistic regression d = 100, HMCKernel{numSteps=20}, step-size search + dual
    averaging) at N = 1 000 rows, where the CPU oracle affords the run: 32 oracle chains x (100 burn-in + 60 kept transitions)
    against 4 096 device chains, all 100 coefficients, mean and variance."""
    X, y, _ = qs.programs.c3_logistic_data(1000, 100)
    prog = qs.programs.logistic(X, y, 2.5)
    k = qs.HMCKernel(numSteps=20)
    dev, st = _device_values(qs, prog, k, 4096, 60, 100, 1, seed=20261019)
    ora = _oracle_values(oracle, prog, k, 32, 60, 100, 1, seed=20261019, chain_begin=1 << 20)
    assert 0.4 < st["props_accepted"] / st["props_made"] < 0.95
    _compare(dev, ora, "c3 (N=1000) adaptive HMC L=20, device vs oracle")


def test_c3_full_shape_adaptive_vs_fixed_step_posterior(qs):
    """config c3 at its full shape (N = 10 000, 8 192 chains, the run bench.py times) against an exactly invariant kernel on
    the same posterior: HMC with a FIXED step size (no adaptation: a valid Metropolis kernel), other seed, longer burn-in.
    The posterior means and variances of all 100 coefficients agree, and both agree with the data-generating coefficients to
    the posterior's own width."""
    X, y, theta = qs.programs.c3_logistic_data(10000, 100)
    prog = qs.programs.logistic(X, y, 2.5)
    a, st = _device_values(qs, prog, qs.HMCKernel(numSteps=20), 8192, 60, 150, 1, seed=20261019)
    # (step 0.025: L * eps = 0.5 stays well below the half period pi * sd ~ 0.7 of this nearly isotropic posterior; at 0.04 the
    # fixed-length trajectories sit on the resonance and the variance estimate is still 4-15 % off after 400 transitions --
    # tools/c3_variance_probe.py, gpurun_out r02k)
    b, _ = _device_values(qs, prog, qs.HMCKernel(numSteps=20, stepSize=0.025, doStepSizeAdapt=False), 4096, 60, 800, 1, seed=9)
    _compare(a, b, "c3 full shape: adaptive (bench configuration) vs fixed-step HMC")
    post_sd = a.reshape(-1, 100).std(axis=0)
    assert (np.abs(a.reshape(-1, 100).mean(axis=0) - theta) < 5 * post_sd).all()


def test_c4_full_dimension_ensemble_device_vs_oracle(qs, oracle):
    """config c4's model at full dimension (J = 1 000, d = 1 002) under its kernel (DriftKernel{scale=0.02}, scale adaptation
    on): 1 500 proposals per chain, the last 500 kept with lag 20; 128 oracle chains against 1 024 device chains, all 1 002
    coordinates."""
    y, sigma = qs.programs.c4_eight_schools_data(1000)
    prog = qs.programs.eight_schools(y, sigma)
    k = qs.DriftKernel(scale=0.02)
    dev, _ = _device_values(qs, prog, k, 1024, 25, 1000, 20, seed=20261019)
    ora = _oracle_values(oracle, prog, k, 128, 25, 1000, 20, seed=20261019, chain_begin=1 << 20)
    _compare(dev, ora, "c4 (J=1000) drift, device vs oracle")


def test_c4_small_posterior_device_vs_oracle(qs, oracle):
    """The same program at J = 30 groups, where the drift kernel converges: posterior moments after 20 000 burn-in proposals."""
    rng = np.random.default_rng(5)
    sigma = rng.uniform(8.0, 18.0, 30)
    yv = 4.0 + 3.0 * rng.standard_normal(30) + sigma * rng.standard_normal(30)
    prog = qs.programs.eight_schools(yv, sigma)
    k = qs.DriftKernel(scale=0.3)
    dev, _ = _device_values(qs, prog, k, 2048, 100, 20000, 50, seed=3)
    ora = _oracle_values(oracle, prog, k, 128, 100, 20000, 50, seed=3, chain_begin=1 << 20)
    _compare(dev, ora, "eight schools J=30 drift posterior, device vs oracle")


def test_c5_full_dimension_ensemble_device_vs_oracle(qs, oracle):
    """config c5 at full dimension (1000-D funnel) under its kernel (MixtureKernel{HARM, HMC L=16} 0.5/0.5, adaptive): 120
    iterations, the last 60 kept with lag 2; 192 oracle chains against 1 024 device chains, all 1 000 coordinates."""
    prog = qs.programs.funnel(1000)
    k = qs.MixtureKernel([qs.HARMKernel(), qs.HMCKernel(numSteps=16)], [0.5, 0.5])
    dev, _ = _device_values(qs, prog, k, 1024, 30, 60, 2, seed=20261019)
    ora = _oracle_values(oracle, prog, k, 192, 30, 60, 2, seed=20261019, chain_begin=1 << 20)
    _compare(dev, ora, "c5 (d=1000) HARM+HMC mixture, device vs oracle")


def test_c5_analytic_truth_under_a_fixed_step_kernel(qs):
    """The funnel is its own prior (v ~ N(0, 3), x_i | v ~ N(0, e^{v/2})), so the forward-sampling initialisation
    (trace.t:592-624) starts every chain IN the target and a valid Metropolis kernel (HMC with a fixed step size) keeps the
    ensemble there: after 40 transitions v has mean 0 and variance 9 and x_i e^{-v/2} has unit variance, within 3.5 exact
    standard errors over 16 384 chains."""
    prog = qs.programs.funnel(1000)
    k = qs.HMCKernel(numSteps=16, stepSize=0.03, doStepSizeAdapt=False)
    v, st = _device_values(qs, prog, k, 16384, 1, 39, 1, seed=11)
    v = v[:, 0, :]
    C = v.shape[0]
    assert st["props_accepted"] > 0.2 * st["props_made"]
    assert abs(v[:, 0].mean()) <= 3.5 * 3.0 / np.sqrt(C), v[:, 0].mean()
    assert abs(v[:, 0].var() - 9.0) <= 3.5 * 9.0 * np.sqrt(2.0 / C), v[:, 0].var()
    zz = (v[:, 1:] ** 2 * np.exp(-v[:, :1])).mean(axis=1)          # per chain: mean_i x_i^2 e^{-v}, expectation 1
    assert abs(zz.mean() - 1.0) <= 3.5 * zz.std(ddof=1) / np.sqrt(C), zz.mean()
