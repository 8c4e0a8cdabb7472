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
