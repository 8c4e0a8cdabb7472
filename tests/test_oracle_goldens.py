"""CPU suite, part 1: pin the oracle (the CPU restatement of the reference algorithm) against every
fixture the reference's own tests hold for this path: the 26 log-density known-answer values of
test/testsuite.t:159-187 (committed as tests/golden/testsuite_logprob_goldens.json) and the posterior
expectation truths of testsuite.t:621-691, 719-730, 776-787, 812-823 under the reference's own
protocol (150 samples x lag 20 x 5 runs, mean abs error <= 0.07: testsuite.t:84-89, 126-140)."""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "testsuite_logprob_goldens.json")))


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_density_goldens(oracle, case):
    got = oracle.logprob(case["which"], case["args"])
    assert abs(got - case["value"]) <= GOLD["tolerance"], (case, got)


def test_lanczos_uses_five_coefficients(oracle):
    """gamma lp(4;2,2) is -2.000000048134726 (not -2): pinned by testsuite.t:166."""
    assert abs(oracle.logprob(3, [4.0, 2.0, 2.0]) - (-2.000000048134726)) < 1e-14
    assert abs(oracle.logprob(9, [2.0]) - 4.8134726e-08) < 1e-12


def test_tape_ad_against_finite_differences(oracle):
    for which, args in [(2, [0.3, 0.1, 0.5]), (3, [1.7, 2.0, 2.0]), (4, [0.3, 2.0, 5.0])]:
        for wrt in range(3):
            lp, g = oracle.logprob_grad(which, args, wrt)
            h = 1e-6
            a1 = list(args); a2 = list(args); a1[wrt] += h; a2[wrt] -= h
            fd = (oracle.logprob(which, a1) - oracle.logprob(which, a2)) / (2 * h)
            assert abs(g - fd) < 1e-6 * max(1, abs(fd)), (which, wrt, g, fd)


def _expected_value(oracle, program, specs, truth, runs=5, numsamps=150, lag=20, errtol=0.07, seed0=100):
    errs = []
    for r in range(runs):
        out = oracle.run(program, specs, seed0 + r, 1, numsamps, 0, lag)
        est = oracle.expectation(out["samples"][0])[0]
        errs.append(abs(est - truth))
    return float(np.mean(errs))


HMC_TRUTHS = [("gaussian", (0, 0.1, 0.5), 1.0, 0.1), ("uniform", (3, 0.1, 0.4), 1.0, 0.25),
              ("gamma", (1, 2.0, 2.0), 10.0, 0.4), ("beta", (2, 2.0, 5.0), 1.0, 2.0 / 7.0)]


@pytest.mark.parametrize("name,erp,div,truth", HMC_TRUTHS, ids=[t[0] for t in HMC_TRUTHS])
@pytest.mark.parametrize("steps", [1, 10])
def test_hmc_expectation_truths(oracle, name, erp, div, truth, steps):
    """testsuite.t:621-691: Langevin (numSteps=1) and HMC (numSteps=10) on the four continuous ERPs."""
    prog = oracle.Program.indep([erp[0]], [erp[1]], [erp[2]], 0, div)
    err = _expected_value(oracle, prog, oracle.spec(oracle.HMC, numSteps=steps), truth)
    assert err <= 0.07, err


@pytest.mark.parametrize("kind", ["hmc", "harm", "drift"])
def test_ten_gaussians_truth(oracle, kind):
    """testsuite.t:719-730 (HMC), 776-787 (HARM), 812-823 (Drift): mean of 10 iid gaussian(0.1, 0.5) -> 0.1."""
    prog = oracle.Program.indep([0] * 10, [0.1] * 10, [0.5] * 10, 0, 10.0)
    spec = {"hmc": oracle.spec(oracle.HMC, numSteps=10), "harm": oracle.spec(oracle.HARM), "drift": oracle.spec(oracle.DRIFT)}[kind]
    assert _expected_value(oracle, prog, spec, 0.1) <= 0.07


def test_expectation_quirk(oracle):
    """SURVEY Q9 (infer.t:154-180): mean = sum(v)/(N-1), variance / N."""
    v = np.array([1.0, 2.0, 3.0, 4.0])
    m, var = oracle.expectation(v, do_variance=True)
    assert abs(m[0] - 10.0 / 3.0) < 1e-15
    assert abs(var - np.mean((v - 10.0 / 3.0) ** 2)) < 1e-15


def test_stream_is_partition_invariant(oracle):
    prog = oracle.Program.funnel(6)
    a = oracle.run(prog, oracle.spec(oracle.HMC, numSteps=4), 7, 4, 10)
    b = oracle.run(prog, oracle.spec(oracle.HMC, numSteps=4), 7, 2, 10, chain_begin=2)
    assert np.array_equal(a["samples"][2:], b["samples"])


def test_dirichlet_unit_simplex_restatement(oracle):
    """erp.t:200-254 + distrib.t:336-362 in the oracle: the dual-trace density is lp(fwd(x)) + logJ(x); its tape-AD gradient
    equals the closed form alpha_k (1 - z_k) - z_k * sum_{m>k} alpha_m the device uses; component K-1 is inert; a
    Dirichlet(2,3,5) chain under HMC has mean alpha / sum(alpha)."""
    alpha = np.array([2.0, 3.0, 5.0])
    prog = oracle.Program.indep([4, 4, 4], alpha, [0, 1, 2], 1, 1.0)
    assert prog.dim == 3 and prog.retdim == 3
    x = np.array([0.3, -0.2, 1.7])
    lpd, g, lpf = prog.logp_grad(x)
    z = [1.0 / (1.0 + np.exp(-(x[k] - np.log(2 - k)))) for k in range(2)]
    want = [alpha[k] * (1 - z[k]) - z[k] * alpha[k + 1:].sum() for k in range(2)] + [0.0]
    assert np.allclose(g, want, rtol=0, atol=1e-13)
    y = [z[0], (1 - z[0]) * z[1], (1 - z[0]) * (1 - z[1])]
    assert abs(lpf - oracle.logprob(8, list(y) + list(alpha))) < 1e-13
    lj = sum(np.log(s) + np.log(zz) + np.log(1 - zz) for s, zz in zip([1.0, 1 - z[0]], z))
    assert abs(lpd - (lpf + lj)) < 1e-12
    out = oracle.run(prog, oracle.spec(oracle.HMC, numSteps=5), 11, 16, 300, 100)
    s = out["samples"]
    assert np.abs(s.sum(axis=2) - 1.0).max() < 1e-12
    assert np.abs(s.mean(axis=(0, 1)) - alpha / alpha.sum()).max() < 0.03


def test_annealing_kernel_restatement(oracle):
    """mcmc.t:297-319: temperature 1 everywhere is the plain kernel; a hot schedule changes the chain and kept samples store
    logprob * temperature (mcmc.t:69), i.e. the untempered log-density of the kept state when it was scored this iteration."""
    prog = oracle.Program.funnel(6)
    sp = oracle.spec(oracle.DRIFT, scale=0.5)
    a = oracle.run(prog, sp, 5, 3, 40, record=True)
    b = oracle.run(prog, sp, 5, 3, 40, record=True, temps=[1.0] * 40)
    c = oracle.run(prog, sp, 5, 3, 40, record=True, temps=[5.0] * 40)
    assert np.array_equal(a["state"], b["state"]) and np.array_equal(a["logprobs"], b["logprobs"])
    assert not np.array_equal(a["state"], c["state"])
    assert c["accept"].mean() > a["accept"].mean()          # a flatter target accepts more
    # constant temperature: logprob/T * T restores the log-density of the state
    for ch in range(3):
        lp = [prog.logp_grad(c["state"][ch, i])[2] for i in range(40)]
        acc = c["accept"][ch].astype(bool)
        assert np.allclose(np.array(lp)[acc], c["logprobs"][ch][acc], rtol=1e-12)


def test_drift_lexical_scale_sharing_truth(oracle):
    """testsuite.t:846-860: 10 gaussians drawn by one call site, DriftKernel{lexicalScaleSharing=true} -> 0.1."""
    prog = oracle.Program.indep([0] * 10, [0.1] * 10, [0.5] * 10, 0, 10.0, lexids=[0] * 10)
    assert _expected_value(oracle, prog, oracle.spec(oracle.DRIFT, lexicalScaleSharing=True), 0.1) <= 0.07


@pytest.mark.parametrize("name,erp,div,truth", HMC_TRUTHS, ids=[t[0] for t in HMC_TRUTHS])
def test_tracemh_expectation_truths(oracle, name, erp, div, truth):
    """testsuite.t:192-266 (multiMethodExpectedValueTest, the MCMC leg with TraceMHKernel) on the continuous ERPs."""
    prog = oracle.Program.indep([erp[0]], [erp[1]], [erp[2]], 0, div)
    assert _expected_value(oracle, prog, oracle.spec(oracle.TRACEMH), truth) <= 0.07


@pytest.mark.parametrize("kind", ["tracemh", "hmc"])
def test_factor_and_observe_expectation(oracle, kind):
    """testsuite.t:394-442 "factor expectation (mcmc)" / "observe expectation (mcmc)": x ~ uniform(-1, 1) with
    qs.factor(gaussian.logprob(x, 0.3, 0.1)) == qs.gaussian.observe(x, 0.3, 0.1) -> 0.3 (the reference runs it under TraceMH)."""
    prog = oracle.Program.indep([3], [-1.0], [1.0], 0, 1.0, fmu=[0.3], fsigma=[0.1])
    spec = oracle.spec(oracle.TRACEMH) if kind == "tracemh" else oracle.spec(oracle.HMC, numSteps=10)
    assert _expected_value(oracle, prog, spec, 0.3) <= 0.07


# ---- latent ERP parameters: parameters of a choice computed from earlier choices, differentiated by the tape
def _hier10(oracle):
    """mu, logs ~ gaussian; x ~ gaussian(mu, exp(logs/2)); z ~ gaussian(0.5 + 2x, 1); k, th ~ gamma; y ~ gamma(k, th);
    a ~ gamma; lb ~ gaussian; w ~ beta(a, exp(lb))."""
    kinds = [0, 0, 0, 0, 1, 1, 1, 1, 0, 2]
    p0 = [1.0, 0.0, 0.0, 0.5, 3.0, 2.0, 0.0, 2.5, 0.7, 0.0]
    p1 = [0.5, 0.5, 0.0, 1.0, 1.0, 0.5, 0.0, 1.0, 0.3, 0.0]
    pref = [-1, -1, -1, -1, 0, 1, 2, -1, -1, -1, -1, -1, 4, 5, -1, -1, -1, -1, 7, 8]
    pcoef = [0, 0, 0, 0, 1.0, 0.5, 2.0, 0, 0, 0, 0, 0, 1.0, 1.0, 0, 0, 0, 0, 1.0, 1.0]
    ptf = [0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1]
    return oracle.Program.indep(kinds, p0, p1, 1, 1.0, pref=pref, pcoef=pcoef, ptf=ptf)


def test_latent_parameter_gradient_matches_finite_differences(oracle):
    prog = _hier10(oracle)
    rng = np.random.default_rng(0)
    for _ in range(4):
        x = rng.normal(0, 0.7, 10)
        _, g, _ = prog.logp_grad(x)
        h = 1e-6
        gn = np.array([(prog.logp_grad(x + h * np.eye(10)[i])[0] - prog.logp_grad(x - h * np.eye(10)[i])[0]) / (2 * h) for i in range(10)])
        assert np.abs(g - gn).max() <= 1e-7 * max(1.0, np.abs(g).max())


def _lg5_digamma(z):
    """d/dz of the reference's 5-coefficient log-gamma (distrib.t:84-106), SURVEY Appendix A."""
    cof = [76.18009172947146, -86.50532032941677, 24.01409824083091, -1.231739572450155, 0.1208650973866179e-2]
    t = z + 4.5
    ser = 1.000000000190015 + sum(c / (z + j) for j, c in enumerate(cof))
    dser = -sum(c / (z + j) ** 2 for j, c in enumerate(cof))
    return np.log(t) + (z - 0.5) / t - 1.0 + dser / ser


def test_latent_shape_gradient_is_the_derivative_of_the_5_coefficient_log_gamma(oracle):
    """k ~ gamma(3, 1); y ~ gamma(k, 2): d/dx0 of the dual-trace log-density with k = exp(x0) is
    [ (3-1)/k - 1 + ln y - psi5(k) - ln 2 ] k + 1, psi5 the derivative of the reference's own log-gamma -- not the true
    digamma, which differs from it by ~1e-8 (a 1e-10 parity gate sees that)."""
    from scipy.special import digamma
    prog = oracle.Program.indep([1, 1], [3.0, 0.0], [1.0, 2.0], 1, 1.0, pref=[-1, -1, 0, -1], pcoef=[0, 0, 1.0, 0], ptf=[0, 0, 0, 0])
    worst_true = 0.0
    for x0, x1 in [(0.3, 0.2), (-0.5, 1.0), (1.1, -0.4)]:
        _, g, _ = prog.logp_grad(np.array([x0, x1]))
        k, y = np.exp(x0), np.exp(x1)
        want = (2.0 / k - 1.0 + np.log(y) - _lg5_digamma(k) - np.log(2.0)) * k + 1.0
        assert abs(g[0] - want) <= 1e-12 * max(1.0, abs(want)), (g[0], want)
        worst_true = max(worst_true, abs(g[0] - ((2.0 / k - 1.0 + np.log(y) - digamma(k) - np.log(2.0)) * k + 1.0)))
    assert worst_true > 1e-10


@pytest.mark.parametrize("kind", ["hmc", "drift", "tracemh"])
def test_latent_parameter_expectations(oracle, kind):
    """mu ~ gaussian(1, 0.5); x ~ gaussian(0.5 + 2 mu, 1): E[x] = 2.5.   k ~ gamma(3, 1); y ~ gamma(k, 2)/10: E = 0.6."""
    spec = {"hmc": oracle.spec(oracle.HMC, numSteps=10), "drift": oracle.spec(oracle.DRIFT), "tracemh": oracle.spec(oracle.TRACEMH)}[kind]
    g = oracle.Program.indep([0, 0], [1.0, 0.5], [0.5, 1.0], 1, 1.0, pref=[-1, -1, 0, -1], pcoef=[0, 0, 2.0, 0], ptf=[0, 0, 0, 0])
    errs = []
    for r in range(5):
        out = oracle.run(g, spec, 300 + r, 1, 150, 0, 20)
        errs.append(abs(out["samples"][0][:, 1].mean() - 2.5))
    assert np.mean(errs) <= 0.2, errs
    if kind == "drift":   # Drift scores bounded choices without the log-Jacobian: biased on a gamma by design of the reference
        return
    y = oracle.Program.indep([1, 1], [3.0, 0.0], [1.0, 2.0], 1, 1.0, pref=[-1, -1, 0, -1], pcoef=[0, 0, 1.0, 0], ptf=[0, 0, 0, 0])
    errs = []
    for r in range(5):
        out = oracle.run(y, spec, 400 + r, 1, 150, 0, 20)
        errs.append(abs(out["samples"][0][:, 1].mean() / 10.0 - 0.6))
    assert np.mean(errs) <= 0.12, errs
