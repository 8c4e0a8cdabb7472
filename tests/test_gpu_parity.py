"""GPU parity tests proper: the CUDA engine, called through the C ABI, against the CPU oracle on the
same seeded inputs and the same injected Philox stream (north_star: per-step leapfrog positions
within 1e-10 relative in fp64, accept/reject decisions identical on >= 99.9 % of steps)."""
import numpy as np
import pytest

import parity as P

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_rng_stream_bit_exact(qs, oracle):
    from quicksand_b200 import _lib as L
    import ctypes as C
    for seed, chain, it, slot in [(0, 0, 0, 0), (20261019, 65535, 17, 9), (2**63 + 5, 2**31, oracle.ITER_INIT, 3),
                                  (7, 1, oracle.ITER_SEARCH + 2, oracle.SLOT_MIXTURE)]:
        out = np.empty(64)
        L.check(L.lib().qsb_uniforms(seed, chain, it, slot, 64, out.ctypes.data_as(L.dp)))
        ref = oracle.uniforms(seed, chain, it, slot, 64)
        assert np.array_equal(out, ref)
        g = np.empty(4096)
        L.check(L.lib().qsb_gaussians(seed, chain, it, 0, 4096, 0.3, 1.7, g.ctypes.data_as(L.dp)))
        gref = oracle.gaussians(seed, chain, it, 0, 4096, 0.3, 1.7)
        assert np.array_equal(g, gref), np.abs(g - gref).max()   # Leva sampler: un-contracted arithmetic => bit-exact


def test_device_densities_against_reference_goldens(qs):
    """The known-answer values of test/testsuite.t:159-170 (tolerance 1e-8, testsuite.t:34) on the device functions."""
    from quicksand_b200 import _lib as L
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "testsuite_logprob_goldens.json")))
    for g in gold["cases"]:
        if g["which"] == 8:      # dirichlet (testsuite.t:184-187): K tuples (theta_k, alpha_k, -)
            K = len(g["args"]) // 2
            a = np.zeros((K, 3)); a[:, 0] = g["args"][:K]; a[:, 1] = g["args"][K:]
            out = np.empty(K)
            L.check(L.lib().qsb_logprob(8, a.ctypes.data_as(L.dp), K, out.ctypes.data_as(L.dp)))
            assert abs(out[0] - g["value"]) <= 1e-8, (g, out[0])
            continue
        if g["which"] not in (0, 1, 2, 3, 4):
            continue
        a = np.array((g["args"] + [0.0, 0.0, 0.0])[:3], dtype=np.float64)
        out = np.empty(1)
        L.check(L.lib().qsb_logprob(g["which"], a.ctypes.data_as(L.dp), 1, out.ctypes.data_as(L.dp)))
        assert abs(out[0] - g["value"]) <= 1e-8, (g, out[0])


def _programs(qs):
    rng = np.random.default_rng(5)
    progs = {}
    progs["gaussian1"] = qs.programs.indep([("gaussian", 0.1, 0.5)])
    progs["gamma1"] = qs.programs.indep([("gamma", 2.0, 2.0)], ret_div=10.0)
    progs["beta1"] = qs.programs.indep([("beta", 2.0, 5.0)])
    progs["uniform1"] = qs.programs.indep([("uniform", 0.1, 0.4)])
    progs["gauss10"] = qs.programs.indep([("gaussian", 0.1, 0.5)] * 10, ret_div=10.0)
    progs["mixed10"] = qs.programs.indep([("gaussian", 0.1, 0.5), ("gamma", 2.0, 2.0), ("beta", 2.0, 5.0), ("uniform", -1.0, 3.0),
                                           ("gamma", 0.7, 1.5)] * 2, ret="vector")
    # vector-valued choices (testsuite.t:693-717, 886-897) and the mean/variance parametrisations (erpdefs.t:72-125)
    progs["dirichlet4"] = qs.programs.indep([("dirichlet", [1.0, 1.0, 1.0, 1.0])], ret="vector")
    progs["dirmix9"] = qs.programs.indep([("gaussian", 0.1, 0.5), ("dirichlet", [2.0, 3.0, 5.0]), ("gammamv", 4.0, 8.0),
                                           ("dirichlet", [0.7, 1.5, 1.0]), ("betamv", 0.3, 0.02)], ret="vector")
    # qs.factor / erp.observe statements on the value of a choice (testsuite.t:394-442: uniform(-1,1) with a gaussian(0.3, 0.1) factor)
    progs["factor1"] = qs.programs.indep([("uniform", -1.0, 1.0)], factors=[(0.3, 0.1)])
    progs["factor6"] = qs.programs.indep([("uniform", -1.0, 1.0), ("gaussian", 0.0, 2.0), ("gamma", 2.0, 2.0), ("beta", 2.0, 5.0), ("gaussian", 1.0, 1.0),
                                           ("gammamv", 3.0, 2.0)], ret="vector",
                                          factors=[(0.3, 0.1), (1.0, 0.5), (3.0, 1.0), None, (0.5, 0.7), (2.5, 0.4)])
    progs["wide24"] = qs.programs.indep([("dirichlet", [1.5] * 8), ("gaussian", 0.0, 1.0), ("gamma", 3.0, 0.5), ("dirichlet", [0.8, 1.0, 2.0, 3.0, 1.2, 0.9]),
                                          ("beta", 2.0, 2.0), ("uniform", -2.0, 2.0)] + [("gaussian", 1.0, 0.2)] * 6, ret="vector")
    # call sites shared by several choices (a loop body), for DriftKernel{lexicalScaleSharing=true} (testsuite.t:846-860)
    progs["gauss10loop"] = qs.programs.indep([("gaussian", 0.1, 0.5)] * 10, ret_div=10.0, lexids=[0] * 10)
    progs["sites9"] = qs.programs.indep([("gaussian", 0.1, 0.5), ("dirichlet", [2.0, 3.0, 5.0]), ("gaussian", 1.0, 2.0),
                                          ("dirichlet", [0.7, 1.5, 1.0]), ("gaussian", -1.0, 0.3)], ret="vector", lexids=[0, 1, 0, 1, 2])
    # latent parameters: parameters of a choice computed from earlier choices (hierarchical programs written with plain ERPs)
    lat = qs.programs.latent
    progs["hier4"] = qs.programs.indep([("gaussian", 1.0, 0.5), ("gaussian", 0.0, 0.5), ("gaussian", lat(0), lat(1, b=0.5, exp=True)),
                                         ("gaussian", lat(2, a=0.5, b=2.0), 1.0)], ret="vector")
    progs["hier12"] = qs.programs.indep([("gaussian", 1.0, 0.5), ("gaussian", 0.0, 0.5), ("gaussian", lat(0), lat(1, b=0.5, exp=True)),
                                          ("gaussian", lat(2, a=0.5, b=2.0), 1.0), ("gamma", 3.0, 1.0), ("gamma", 2.0, 0.5), ("gamma", lat(4), lat(5)),
                                          ("gamma", 2.5, 1.0), ("gaussian", 0.7, 0.3), ("beta", lat(7), lat(8, exp=True)), ("uniform", -1.0, 1.0),
                                          ("gaussian", lat(10, b=3.0), lat(6, a=0.2))], ret="vector",
                                         factors=[None] * 9 + [(0.4, 0.2), None, (0.5, 1.0)])
    progs["hierdir9"] = qs.programs.indep([("gamma", 3.0, 1.0), ("dirichlet", [2.0, 3.0, 5.0]), ("gamma", lat(0), 2.0), ("gaussian", 0.0, 1.0),
                                            ("dirichlet", [0.7, 1.5, 1.0]), ("gaussian", lat(3), lat(2, a=0.1))], ret="vector")
    # Bounds.Upper (erp.t:158-173) through a user-defined ERP, and qs.condition statements (trace.t:794-796): hard constraints with
    # rejection initialisation and `conditionsSatisfied and ...` accept tests
    progs["neggamma1"] = qs.programs.indep([("neggamma", 2.0, 2.0)], ret_div=10.0)
    progs["upper5"] = qs.programs.indep([("neggamma", 2.0, 2.0), ("gaussian", 0.1, 0.5), ("neggamma", 0.7, 1.5), ("gamma", 2.0, 2.0), ("neggamma", 3.0, 0.5)], ret="vector")
    progs["cond5"] = qs.programs.indep([("gaussian", 0.0, 1.0), ("neggamma", 2.0, 1.5), ("gamma", 2.0, 2.0), ("uniform", -1.0, 1.0), ("gaussian", 1.0, 2.0)], ret="vector",
                                        conditions=[(-0.5, None), (None, -0.1), (0.2, 9.0), None, (-1.0, 3.0)])
    progs["cond5w"] = qs.programs.indep([("gaussian", 0.0, 1.0), ("neggamma", 2.0, 1.5), ("gamma", 2.0, 2.0), ("uniform", -1.0, 1.0), ("gaussian", 1.0, 2.0)] * 8, ret="vector",
                                         conditions=[(-1.5, None), (None, -0.01), (0.02, 30.0), None, (-5.0, 7.0)] * 8)          # 40 components: warp per chain
    progs["condhier4"] = qs.programs.indep([("gaussian", 1.0, 0.5), ("gaussian", 0.0, 0.5), ("gaussian", lat(0), lat(1, b=0.5, exp=True)),
                                             ("gaussian", lat(2, a=0.5, b=2.0), 1.0)], ret="vector", conditions=[(0.0, None), None, (-1.0, 4.0), None])
    # vector choices and latent parameters beyond 32 components: warp per chain, the general evaluator on scratch memory
    progs["hier70"] = qs.programs.indep([("gaussian", 1.0, 0.5), ("gaussian", 0.0, 0.5)] + [("gaussian", lat(0), lat(1, b=0.5, exp=True))] * 60
                                         + [("gamma", 3.0, 1.0), ("gamma", lat(62), 2.0), ("dirichlet", [2.0, 3.0, 5.0, 1.5, 2.5, 4.0])], ret="vector",
                                         factors=[None] * 2 + [(1.2, 0.8)] * 30 + [None] * 32 + [None])
    progs["hier64"] = qs.programs.indep([("gaussian", 1.0, 0.5), ("gaussian", 0.0, 0.5)] + [("gaussian", lat(0), lat(1, b=0.5, exp=True))] * 60
                                         + [("gamma", 3.0, 1.0), ("gamma", lat(62), 2.0)], ret="vector", factors=[None] * 2 + [(1.2, 0.8)] * 30 + [None] * 32)
    progs["dirwide40"] = qs.programs.indep([("dirichlet", [1.5, 0.8, 2.0, 3.0, 1.2, 0.9, 1.0, 2.5])] * 5, ret="vector")
    progs["c2"] = qs.programs.c2_correlated_gaussian()[0]
    J = 11
    progs["schools13"] = qs.programs.eight_schools(rng.normal(4, 10, J), rng.uniform(8, 18, J))
    J = 60
    progs["schools62"] = qs.programs.eight_schools(rng.normal(4, 10, J), rng.uniform(8, 18, J))
    progs["funnel12"] = qs.programs.funnel(12)
    progs["funnel70"] = qs.programs.funnel(70)
    X = rng.standard_normal((333, 20)) / np.sqrt(20)
    y = (rng.random(333) < 0.5).astype(np.uint8)
    progs["logistic20"] = qs.programs.logistic(X, y, 2.5)
    return progs


@pytest.mark.parametrize("name", ["gaussian1", "gamma1", "beta1", "uniform1", "gauss10", "mixed10", "dirichlet4", "dirmix9", "wide24", "factor1", "factor6", "hier4", "hier12", "hierdir9", "c2",
                                  "schools13", "schools62", "funnel12", "funnel70", "logistic20", "neggamma1", "upper5", "cond5", "cond5w", "condhier4"])
def test_logp_and_gradient_match_tape_ad(qs, oracle, name):
    """Closed-form device gradients vs the oracle's tape AD (qs/lib/ad.t) on random unconstrained points."""
    prog = _programs(qs)[name]
    op = P.oracle_program(prog)
    rng = np.random.default_rng(11)
    x = rng.normal(0, 1.0, (37, prog.dim))
    lpd, g, lpf = prog.logp_grad(x)
    for c in range(x.shape[0]):
        lpd_o, g_o, lpf_o = op.logp_grad(x[c])
        assert P.relerr(lpd[c], lpd_o) <= 1e-12, (name, c, lpd[c], lpd_o)
        assert P.relerr(lpf[c], lpf_o) <= 1e-12, (name, c, lpf[c], lpf_o)
        assert P.relerr(g[c], g_o, P.vector_floor(g_o)).max() <= 1e-11, (name, c, np.abs(g[c] - g_o).max())


# ---- HMC with fixed step sizes in the numerically stable regime: strict per-leapfrog-step parity over whole runs
HMC_FIXED = [
    ("gaussian1", dict(numSteps=1, stepSize=0.3), 0, 1.0), ("gaussian1", dict(numSteps=10, stepSize=0.2), 0, 1.0),
    ("gamma1", dict(numSteps=10, stepSize=0.2), 0, 1.0), ("beta1", dict(numSteps=10, stepSize=0.2), 0, 1.0),
    ("uniform1", dict(numSteps=10, stepSize=0.5), 0, 1.0),
    ("gauss10", dict(numSteps=10, stepSize=0.2), 0, 1.0), ("mixed10", dict(numSteps=5, stepSize=0.1), 0, 1.0),
    ("dirichlet4", dict(numSteps=10, stepSize=0.2), 0, 1.0), ("dirmix9", dict(numSteps=5, stepSize=0.1), 0, 1.0),
    ("wide24", dict(numSteps=5, stepSize=0.05), 0, 1.0),
    ("factor1", dict(numSteps=10, stepSize=0.05), 0, 1.0), ("factor6", dict(numSteps=5, stepSize=0.05), 0, 1.0), ("factor6", dict(numSteps=5, stepSize=0.05), 32, 1.0),
    ("mixed10", dict(numSteps=5, stepSize=0.1), 32, 1.0),
    ("hier4", dict(numSteps=10, stepSize=0.05), 0, 1.0), ("hier12", dict(numSteps=5, stepSize=0.02), 0, 0.9), ("hierdir9", dict(numSteps=5, stepSize=0.03), 0, 0.9),
    ("c2", dict(numSteps=16, stepSize=0.15), 0, 1.0), ("c2", dict(numSteps=16, stepSize=0.15), 4, 1.0),   # 4: four lanes per chain
    ("neggamma1", dict(numSteps=10, stepSize=0.2), 0, 1.0), ("upper5", dict(numSteps=5, stepSize=0.1), 0, 1.0), ("upper5", dict(numSteps=5, stepSize=0.1), 32, 1.0),
    ("cond5", dict(numSteps=5, stepSize=0.15), 0, 1.0), ("cond5", dict(numSteps=5, stepSize=0.15), 32, 1.0), ("cond5w", dict(numSteps=4, stepSize=0.05), 0, 1.0),
    ("condhier4", dict(numSteps=6, stepSize=0.05), 0, 1.0),
    ("schools13", dict(numSteps=8, stepSize=0.02), 1, 0.95), ("schools13", dict(numSteps=8, stepSize=0.02), 32, 0.95),
    ("schools62", dict(numSteps=8, stepSize=0.02), 0, 0.95),
    ("funnel12", dict(numSteps=16, stepSize=0.01), 1, 0.9), ("funnel12", dict(numSteps=16, stepSize=0.01), 32, 0.9),
    ("funnel70", dict(numSteps=16, stepSize=0.01), 0, 0.9),
    ("logistic20", dict(numSteps=20, stepSize=0.05), 0, 1.0), ("logistic20", dict(numSteps=7, stepSize=0.1), 0, 1.0),
]


@pytest.mark.parametrize("name,kp,mapping,min_compared", HMC_FIXED)
def test_hmc_fixed_step_per_leapfrog_parity(qs, oracle, name, kp, mapping, min_compared):
    prog = _programs(qs)[name]
    kp = dict(kp, doStepSizeAdapt=False)
    gpu, cpu = P.run_both(prog, qs.HMCKernel(kp), chains=96, iters=40, seed=20261019, leap=True, mapping=mapping)
    r = P.compare(gpu, cpu, TOL, leap=True, what="HMC fixed %s %s map%d" % (name, kp, mapping), min_compared=min_compared)
    v = r["valid"]
    assert P.relerr(gpu["samples"], cpu["samples"], P.vector_floor(cpu["samples"]))[v].max() <= 1e-9
    assert P.relerr(gpu["logprobs"], cpu["logprobs"])[v].max() <= 1e-9


# ---- HMC with the reference's step-size search + dual averaging (hmc.t:345-402).  Iteration 0 starts from identical
# states on both sides, so the search result and the first transition are compared strictly; later iterations in
# prefix mode (see parity.compare): the adaptation transient runs through unstable step sizes.
# The last column is the fraction of chain-steps that must stay within 1e-10 of the oracle.  It is high wherever
# the dynamics are stable.  For long trajectories on the hierarchical / GLM targets it is low on purpose: the
# searched step size sits at the edge of stability and the dual averaging (x0 = eps, eps = exp(x): hmc.t:261,
# 394-402) then raises it further, so a single leapfrog step has a Lipschitz constant of eps^2 |Hessian| ~ 1e4
# and L of them amplify the few-ulp difference between a sequential sum over observations (oracle) and a tiled
# tensor-core sum (GPU) past 1e-10 within the first transitions.  The printed control shows the same effect
# between two CPU builds of the oracle that differ only in FMA contraction.
HMC_ADAPT = [
    ("gaussian1", 1, 0, 0.9), ("gaussian1", 10, 0, 0.5), ("gamma1", 10, 0, 0.5), ("beta1", 1, 0, 0.9), ("uniform1", 10, 0, 0.8),
    ("gauss10", 10, 0, 0.4), ("mixed10", 5, 32, 0.8), ("c2", 16, 0, 0.4), ("c2", 16, 4, 0.4), ("dirichlet4", 10, 0, 0.4), ("dirmix9", 1, 0, 0.8),
    ("schools13", 1, 32, 0.9), ("schools13", 2, 1, 0.8), ("schools62", 1, 0, 0.9), ("funnel12", 1, 1, 0.9), ("funnel12", 2, 32, 0.8),
    ("funnel70", 1, 0, 0.9), ("logistic20", 1, 0, 0.9), ("logistic20", 2, 0, 0.05), ("hier4", 1, 0, 0.8), ("hier12", 2, 0, 0.3),
    ("schools13", 8, 32, 0.05), ("funnel12", 16, 1, 0.03), ("logistic20", 20, 0, 0.01),
]


@pytest.mark.parametrize("name,L,mapping,min_compared", HMC_ADAPT)
def test_hmc_adaptive_search_and_dual_averaging(qs, oracle, name, L, mapping, min_compared):
    prog = _programs(qs)[name]
    gpu, cpu = P.run_both(prog, qs.HMCKernel(numSteps=L), chains=96, iters=25, seed=7, leap=True, mapping=mapping)
    # the searched initial step size (a chain of single leapfrog steps from the identical start) must agree exactly
    assert np.array_equal(gpu["step"][:, 0], cpu["step"][:, 0]), np.abs(gpu["step"][:, 0] - cpu["step"][:, 0]).max()
    ctl = P.run_oracle(prog, qs.HMCKernel(numSteps=L), chains=96, iters=25, seed=7, variant="fma")
    rc = P.compare(ctl, cpu, TOL, leap=True, what="   control (oracle_fma vs oracle)", check=False)
    r = P.compare(gpu, cpu, TOL, leap=True, what="HMC adaptive %s L=%d map%d" % (name, L, mapping),
                  min_compared=min(min_compared, 0.8 * rc["compared"]))
    v = r["valid"]
    assert P.relerr(gpu["step"], cpu["step"])[v].max() <= 1e-9       # dual-averaging sequence on the compared prefix
    assert gpu["stats"]["grad_evals"] > 0


@pytest.mark.parametrize("name,mapping", [("gauss10", 0), ("c2", 0), ("c2", 4), ("schools13", 1), ("schools13", 32), ("schools62", 0),
                                          ("funnel12", 1), ("funnel70", 0), ("mixed10", 0), ("dirichlet4", 0), ("dirmix9", 0), ("wide24", 0), ("logistic20", 0), ("factor6", 0), ("factor6", 32),
                                          ("hier4", 0), ("hier12", 0), ("hierdir9", 0), ("upper5", 0), ("upper5", 32), ("cond5", 0), ("cond5", 32), ("cond5w", 0), ("condhier4", 0)])
def test_drift_per_step_parity(qs, oracle, name, mapping):
    prog = _programs(qs)[name]
    gpu, cpu = P.run_both(prog, qs.DriftKernel(scale=0.3), chains=64, iters=400, seed=99, mapping=mapping)
    P.compare(gpu, cpu, TOL, what="Drift %s map%d" % (name, mapping))


LEXICAL = [("gauss10loop", 1), ("gauss10loop", 32), ("gauss10", 0), ("sites9", 0), ("hier12", 0), ("c2", 0), ("schools13", 1), ("schools13", 32),
           ("schools62", 0), ("funnel12", 1), ("funnel12", 32), ("funnel70", 0)]


@pytest.mark.parametrize("name,mapping", LEXICAL)
def test_drift_lexical_scale_sharing_parity(qs, oracle, name, mapping):
    """DriftKernel{lexicalScaleSharing=true} (drift.t:101-115, 208-235): choices of one call site share a ScaleAdapter.
    Thread per chain feeds the shared RunningVar in the reference's component order; warp per chain folds a site's
    components in as one batch (same value up to rounding)."""
    prog = _programs(qs)[name]
    k = qs.DriftKernel(scale=0.3, lexicalScaleSharing=True)
    gpu, cpu = P.run_both(prog, k, chains=64, iters=500, seed=98, mapping=mapping)
    P.compare(gpu, cpu, TOL, what="Drift lexical %s map%d" % (name, mapping))
    if name in ("gauss10loop", "schools13", "funnel12"):   # sharing matters: the per-component adapters give another chain
        plain, _ = P.run_both(prog, qs.DriftKernel(scale=0.3), chains=8, iters=500, seed=98, mapping=mapping)
        assert not np.array_equal(plain["state"], gpu["state"][:8])


@pytest.mark.parametrize("name,mapping", [("gauss10", 0), ("gaussian1", 0), ("c2", 0), ("c2", 4), ("schools13", 32), ("schools62", 0),
                                          ("funnel12", 1), ("funnel70", 0), ("dirichlet4", 0), ("logistic20", 0), ("hier4", 0), ("hier12", 0),
                                          ("upper5", 0), ("cond5", 0), ("cond5", 32), ("cond5w", 0), ("condhier4", 0)])
def test_harm_per_step_parity(qs, oracle, name, mapping):
    prog = _programs(qs)[name]
    gpu, cpu = P.run_both(prog, qs.HARMKernel(), chains=64, iters=300, seed=5, mapping=mapping)
    P.compare(gpu, cpu, TOL, what="HARM %s map%d" % (name, mapping))
    assert P.relerr(gpu["step"], cpu["step"]).max() <= 1e-9


TRACEMH = [("hier4", 0), ("hier12", 0), ("hierdir9", 0), ("factor1", 0), ("factor6", 0), ("factor6", 32), ("gaussian1", 0), ("gamma1", 0), ("beta1", 0), ("uniform1", 0), ("gauss10", 0), ("mixed10", 0), ("mixed10", 32), ("dirichlet4", 0),
           ("dirmix9", 0), ("wide24", 0), ("c2", 0), ("c2", 4), ("schools13", 1), ("schools13", 32), ("schools62", 0), ("funnel12", 1), ("funnel12", 32),
           ("funnel70", 0), ("neggamma1", 0), ("upper5", 0), ("cond5", 0), ("cond5", 32), ("condhier4", 0)]


@pytest.mark.parametrize("name,mapping", TRACEMH)
def test_tracemh_nonstructural_parity(qs, oracle, name, mapping):
    """TraceMHKernel{doStruct=false} (mcmc.t:134-193): which choice, the proposal (gaussian drift with the ERP's own sigma,
    prior resampling for the other ERPs incl. vector choices), the forward/reverse proposal terms and the accept decision."""
    prog = _programs(qs)[name]
    gpu, cpu = P.run_both(prog, qs.TraceMHKernel(doStruct=False), chains=64, iters=300, seed=41, mapping=mapping)
    assert np.array_equal(gpu["step"], cpu["step"])            # the chosen choice index of every iteration
    P.compare(gpu, cpu, TOL, what="TraceMH %s map%d" % (name, mapping))
    assert gpu["accept"].mean() > 0          # (resampling a lone ERP from its prior is always accepted)


def test_tracemh_in_mixtures_and_truth(qs, oracle):
    """The reference's most common kernel shape: MixtureKernel({TraceMHKernel, HMCKernel}) (testsuite.t:732-751 on a
    structural program; here its fixed-structure counterpart), per step against the oracle, and the expectation truths of
    the continuous ERPs under TraceMH alone (testsuite.t:192-266 protocol: 150 samples x lag 20, tolerance 0.07)."""
    prog = _programs(qs)["mixed10"]
    k = qs.MixtureKernel([qs.TraceMHKernel(doStruct=False), qs.HMCKernel(numSteps=3, stepSize=0.05, doStepSizeAdapt=False),
                          qs.DriftKernel(scale=0.2)], [0.5, 0.25, 0.25])
    gpu, cpu = P.run_both(prog, k, chains=64, iters=200, seed=42)
    assert np.array_equal(gpu["kidx"], cpu["kidx"])
    P.compare(gpu, cpu, TOL, what="Mixture TraceMH+HMC+Drift mixed10", min_compared=0.9)
    k = qs.AnnealingKernel(qs.TraceMHKernel(doStruct=False), _cooling)
    gpu, cpu = P.run_both(_programs(qs)["funnel12"], k, chains=32, iters=100, seed=43, mapping=32)
    P.compare(gpu, cpu, TOL, what="Annealed TraceMH funnel12")
    for name, truth in (("gaussian1", 0.1), ("gamma1", 0.4), ("beta1", 2.0 / 7.0), ("uniform1", 0.25)):
        m = qs.MCMC(qs.TraceMHKernel(doStruct=False), dict(numsamps=150, lag=20, numchains=512, seed=9))
        est = qs.infer(_programs(qs)[name], qs.Expectation(), m)()
        assert abs(np.mean(est) - truth) < 0.02, (name, np.mean(est))
        assert np.mean(np.abs(est - truth) <= 0.07) > 0.5
    # testsuite.t:394-442 "factor / observe expectation (mcmc)": uniform(-1, 1) with a gaussian(0.3, 0.1) factor -> 0.3
    for kern in (qs.TraceMHKernel(doStruct=False), qs.HMCKernel(numSteps=10)):
        m = qs.MCMC(kern, dict(numsamps=150, lag=20, numchains=512, seed=10))
        est = qs.infer(_programs(qs)["factor1"], qs.Expectation(), m)()
        assert abs(np.mean(est) - 0.3) < 0.02 and np.mean(np.abs(est - 0.3) <= 0.07) > 0.9, np.mean(est)
    with pytest.raises(NotImplementedError):
        qs.TraceMHKernel()                  # the default (structural) form is not a device kernel
    from quicksand_b200 import _lib as L
    with pytest.raises(L.QsbError):
        qs.Sampler(_programs(qs)["logistic20"], qs.TraceMHKernel(doStruct=False), numchains=4)


@pytest.mark.parametrize("name,mapping,eps", [("funnel12", 1, 0.01), ("funnel70", 0, 0.01), ("gauss10", 0, 0.2), ("schools13", 32, 0.02), ("c2", 4, 0.15)])
def test_mixture_harm_hmc_parity(qs, oracle, name, mapping, eps):
    """config c5's kernel: MixtureKernel({HARMKernel{}, HMCKernel{numSteps=L}}, {0.5, 0.5}) (mcmc.t:199-291)."""
    prog = _programs(qs)[name]
    k = qs.MixtureKernel([qs.HARMKernel(), qs.HMCKernel(numSteps=16, stepSize=eps, doStepSizeAdapt=False)], [0.5, 0.5])
    gpu, cpu = P.run_both(prog, k, chains=64, iters=60, seed=31, mapping=mapping)
    assert np.array_equal(gpu["kidx"], cpu["kidx"])
    P.compare(gpu, cpu, TOL, what="Mixture %s map%d" % (name, mapping), min_compared=0.9)
    assert gpu["stats"]["sub_made"] == [int((cpu["kidx"] == 0).sum()), int((cpu["kidx"] == 1).sum())]


def test_mixture_with_adaptive_hmc(qs, oracle):
    prog = _programs(qs)["funnel12"]
    k = qs.MixtureKernel([qs.HARMKernel(), qs.HMCKernel(numSteps=16)], [0.5, 0.5])
    gpu, cpu = P.run_both(prog, k, chains=64, iters=40, seed=32)
    assert np.array_equal(gpu["kidx"], cpu["kidx"])
    P.compare(gpu, cpu, TOL, what="Mixture adaptive funnel12", min_compared=0.25)


def test_three_way_mixture_and_lag(qs, oracle):
    prog = _programs(qs)["gauss10"]
    k = qs.MixtureKernel([qs.DriftKernel(), qs.HARMKernel(), qs.HMCKernel(numSteps=3, stepSize=0.2, doStepSizeAdapt=False)], [0.2, 0.3, 0.5])
    gpu, cpu = P.run_both(prog, k, chains=32, iters=10 + 20 * 3, seed=8, burnin=10, lag=3)
    assert np.array_equal(gpu["kidx"], cpu["kidx"])
    P.compare(gpu, cpu, TOL, what="3-way mixture")
    assert gpu["samples"].shape == cpu["samples"].shape == (32, 20, 1)
    assert P.relerr(gpu["samples"], cpu["samples"], P.vector_floor(cpu["samples"])).max() <= 1e-9
    assert P.relerr(gpu["logprobs"], cpu["logprobs"]).max() <= 1e-9


def _cooling(i, n):
    """annealSched(iter, numiters): geometric cooling from 4 to 1 over the first 3/4 of the run."""
    return max(1.0, 4.0 * (0.25 ** (i / (0.75 * n))))


ANNEAL = [
    ("logistic20", "drift", 0), ("logistic20", "harm", 0),
    ("c2", "hmc", 0), ("c2", "hmc", 4), ("c2", "mix3", 4), ("gauss10", "mix3", 0), ("schools13", "drift", 32), ("schools13", "hmc", 32), ("schools13", "hmc", 1),
    ("funnel12", "mix", 32), ("funnel12", "mix", 1), ("funnel70", "harm", 0), ("logistic20", "hmc", 0), ("mixed10", "hmc", 32),
]


@pytest.mark.parametrize("name,kind,mapping", ANNEAL)
def test_annealing_kernel_parity(qs, oracle, name, kind, mapping):
    """AnnealingKernel(kernel, annealSched) (mcmc.t:297-319): the trace temperature changes every iteration; cached
    log-densities and gradients keep the temperature they were computed under (hmc.t:144, 254, 266; drift.t:155;
    harm.t:90), kept samples store logprob * temperature (mcmc.t:69)."""
    prog = _programs(qs)[name]
    eps = {"c2": 0.15, "schools13": 0.02, "funnel12": 0.01, "logistic20": 0.05, "mixed10": 0.1, "gauss10": 0.2}.get(name, 0.1)
    hmc = qs.HMCKernel(numSteps=6, stepSize=eps, doStepSizeAdapt=False)
    inner = {"hmc": hmc, "drift": qs.DriftKernel(scale=0.3), "harm": qs.HARMKernel(),
             "mix": qs.MixtureKernel([qs.HARMKernel(), hmc], [0.5, 0.5]),
             "mix3": qs.MixtureKernel([qs.DriftKernel(), qs.HARMKernel(), hmc], [0.3, 0.3, 0.4])}[kind]
    k = qs.AnnealingKernel(inner, _cooling)
    gpu, cpu = P.run_both(prog, k, chains=64, iters=80, seed=77, mapping=mapping, leap=(kind == "hmc"))
    r = P.compare(gpu, cpu, TOL, leap=(kind == "hmc"), what="Annealing %s %s map%d" % (name, kind, mapping), min_compared=0.9)
    v = r["valid"]
    assert P.relerr(gpu["logprobs"], cpu["logprobs"])[v].max() <= 1e-9
    # the schedule really was applied: an un-annealed run differs
    plain, _ = P.run_both(prog, inner, chains=8, iters=80, seed=77, mapping=mapping)
    assert not np.array_equal(plain["state"], gpu["state"][:8])


def test_annealing_with_adaptive_hmc_first_iteration(qs, oracle):
    """The step-size search and the first gradient run under the first iteration's temperature (hmc.t:218, 254-262)."""
    prog = _programs(qs)["funnel12"]
    k = qs.AnnealingKernel(qs.HMCKernel(numSteps=2), lambda i, n: 3.0 if i < 5 else 1.0)
    for mapping in (1, 32):
        gpu, cpu = P.run_both(prog, k, chains=64, iters=12, seed=78, mapping=mapping, leap=True)
        assert np.array_equal(gpu["step"][:, 0], cpu["step"][:, 0])
        P.compare(gpu, cpu, TOL, leap=True, what="Annealing adaptive HMC map%d" % mapping, min_compared=0.5)
    prog = _programs(qs)["logistic20"]
    gpu, cpu = P.run_both(prog, k, chains=64, iters=12, seed=79, leap=True)
    assert np.array_equal(gpu["step"][:, 0], cpu["step"][:, 0])
    # adaptive HMC on the GLM target leaves the stable region within a few transitions (see HMC_ADAPT): prefix mode
    P.compare(gpu, cpu, TOL, leap=True, what="Annealing adaptive HMC logistic", min_compared=0.05)


def test_full_dimension_configs_against_oracle(qs, oracle):
    """BASELINE configs c4 and c5 at their full dimensions (d = 1002, d = 1000: 32 component rows per lane, the largest
    the warp-per-chain kernels take) on a handful of chains, per step against the oracle; odd chain counts exercise the
    partially filled tail block."""
    y, sigma = qs.programs.c4_eight_schools_data(1000)
    prog = qs.programs.eight_schools(y, sigma)
    gpu, cpu = P.run_both(prog, qs.DriftKernel(scale=0.02), chains=5, iters=260, seed=20261019)
    P.compare(gpu, cpu, TOL, what="c4 full dimension drift")
    assert gpu["accept"].sum() > 5 * 6                      # the adaptive branch (n > 100 needs > 5 acceptances) was reached
    gpu, cpu = P.run_both(prog, qs.DriftKernel(scale=0.02, lexicalScaleSharing=True), chains=3, iters=120, seed=20261019)
    P.compare(gpu, cpu, TOL, what="c4 full dimension drift, lexical scale sharing")
    prog = qs.programs.funnel(1000)
    k = qs.MixtureKernel([qs.HARMKernel(), qs.HMCKernel(numSteps=16, stepSize=0.01, doStepSizeAdapt=False)], [0.5, 0.5])
    gpu, cpu = P.run_both(prog, k, chains=5, iters=24, seed=20261019)
    assert np.array_equal(gpu["kidx"], cpu["kidx"])
    P.compare(gpu, cpu, TOL, what="c5 full dimension HARM+HMC mixture", min_compared=0.9)
    gpu, cpu = P.run_both(prog, qs.HMCKernel(numSteps=3), chains=3, iters=6, seed=20261019, leap=True)
    assert np.array_equal(gpu["step"][:, 0], cpu["step"][:, 0])     # step-size search at d = 1000: identical result
    # the searched step size sits at the edge of stability of a 1000-dimensional funnel; only the first trajectory is
    # comparable before two correct implementations separate (see HMC_ADAPT)
    assert P.relerr(gpu["leapfrog"][:, 0, 0], cpu["leapfrog"][:, 0, 0], P.vector_floor(cpu["leapfrog"][:, 0, 0])).max() <= TOL
    P.compare(gpu, cpu, TOL, leap=True, what="c5 full dimension adaptive HMC", min_compared=0.0)
    with pytest.raises(Exception):
        qs.Sampler(qs.programs.funnel(1025), qs.HARMKernel(), numchains=2)   # beyond the largest variant: loud failure


def test_c3_full_shape_gradient_kernel_against_oracle(qs, oracle):
    """BASELINE config c3 at its full data shape (N = 10 000 rows, d = 100: 13 dimension tiles with the half-empty last
    tile, 313 row blocks split 12 ways) -- the variant of the tensor-core gradient kernel the benchmark runs -- on a
    handful of chains against the oracle's tape AD, plus a partial chain tile (chain counts that are not a multiple of 96)."""
    X, y, _ = qs.programs.c3_logistic_data(10000, 100)
    prog = qs.programs.logistic(X, y, 2.5)
    rng = np.random.default_rng(3)
    x = rng.normal(0, 0.3, (5, 100))
    lpd, g, lpf = prog.logp_grad(x)
    op = P.oracle_program(prog)
    for c in range(5):
        lpd_o, g_o, lpf_o = op.logp_grad(x[c])
        assert P.relerr(lpd[c], lpd_o) <= 1e-12 and P.relerr(g[c], g_o, P.vector_floor(g_o)).max() <= 1e-11, (c, lpd[c], lpd_o)
    k = qs.HMCKernel(numSteps=4, stepSize=0.015, doStepSizeAdapt=False)
    gpu, cpu = P.run_both(prog, k, chains=6, iters=4, seed=20261019, leap=True)
    P.compare(gpu, cpu, TOL, leap=True, what="c3 full shape HMC")
    # QSB_FLAG_INVARIANT: a chain's result depends neither on the tile it sits in nor on how many chains its sampler holds --
    # 200 chains (2 full tiles + a partial one) vs the same ids in shards of 10 and of 150 chains, bit for bit
    from quicksand_b200 import _lib as L
    full = qs.Sampler(prog, k, numchains=200, seed=5, flags=L.FLAG_INVARIANT); full.init(); full.run(3)
    tail = qs.Sampler(prog, k, numchains=10, seed=5, chain_begin=190, flags=L.FLAG_INVARIANT); tail.init(); tail.run(3)
    head = qs.Sampler(prog, k, numchains=150, seed=5, chain_begin=0, flags=L.FLAG_INVARIANT); head.init(); head.run(3)
    fv = full.fetch_values()
    assert np.array_equal(fv[190:], tail.fetch_values()) and np.array_equal(fv[:150], head.fetch_values())
    # the default (balanced) decomposition groups the row blocks by the sampler's chain count: equal to rounding only
    dflt = qs.Sampler(prog, k, numchains=200, seed=5); dflt.init(); dflt.run(3)
    dv = dflt.fetch_values()
    assert P.relerr(dv, fv, P.vector_floor(fv)).max() <= 1e-10
    full.close(); tail.close(); head.close(); dflt.close()


def test_shared_and_per_lane_gaussian_fill_agree(qs, oracle):
    """Thread-per-chain kernels draw the Gaussians of a full warp's 32 chains cooperatively and those of a partial tail warp
    (or of a mixture, where lanes sit in different sub-kernels) lane by lane; a chain's draws must not depend on which:
    40 chains = one full warp + a tail of 8, against the oracle and against the same chain ids run as a lone partial warp."""
    prog = _programs(qs)["c2"]
    k = qs.HMCKernel(numSteps=4, stepSize=0.15, doStepSizeAdapt=False)
    gpu, cpu = P.run_both(prog, k, chains=40, iters=30, seed=12, leap=True)
    P.compare(gpu, cpu, TOL, leap=True, what="HMC c2, 40 chains")
    a = qs.Sampler(prog, k, numchains=40, seed=12); a.init(); a.run(30)
    b = qs.Sampler(prog, k, numchains=8, seed=12, chain_begin=5); b.init(); b.run(30)      # chains 5..12 of the full warp, as a partial warp
    assert np.array_equal(a.fetch_values()[5:13], b.fetch_values())
    a.close(); b.close()
    gpu, cpu = P.run_both(_programs(qs)["gauss10"], qs.HARMKernel(), chains=40, iters=100, seed=13)
    P.compare(gpu, cpu, TOL, what="HARM gauss10, 40 chains")


@pytest.mark.parametrize("d,N", [(128, 700), (104, 333), (97, 257), (57, 100), (8, 40), (3, 33)])
def test_glm_kernel_dimension_edges_against_oracle(qs, oracle, d, N):
    """The tensor-core gradient kernel at the edges of its shapes: the largest dimension (d = 128: the Theta tile leaves room for
    a 2-deep ring only), full and partly filled dimension tiles with and without the trimmed last tile, row counts that are
    not a multiple of the 32-row block."""
    rng = np.random.default_rng(d)
    X = rng.standard_normal((N, d)) / np.sqrt(d)
    y = (rng.random(N) < 0.5).astype(np.uint8)
    prog = qs.programs.logistic(X, y, 2.5)
    op = P.oracle_program(prog)
    x = rng.normal(0, 0.5, (3, d))
    lpd, g, lpf = prog.logp_grad(x)
    for c in range(3):
        lpd_o, g_o, lpf_o = op.logp_grad(x[c])
        assert P.relerr(lpd[c], lpd_o) <= 1e-12 and P.relerr(g[c], g_o, P.vector_floor(g_o)).max() <= 1e-11, (d, c)
    gpu, cpu = P.run_both(prog, qs.HMCKernel(numSteps=3, stepSize=0.05, doStepSizeAdapt=False), chains=5, iters=4, seed=d, leap=True)
    P.compare(gpu, cpu, TOL, leap=True, what="GLM d=%d N=%d" % (d, N))


def test_chain_ids_are_partition_invariant(qs):
    """Philox key = global chain id: chains [32,64) run as a shard equal chains 32..63 of the full run, bit for bit
    (the multi-GPU invariance the chain sharding relies on)."""
    prog = _programs(qs)["funnel12"]
    k = qs.HMCKernel(numSteps=4)
    full = qs.Sampler(prog, k, numchains=64, seed=3); full.init(); full.run(20)
    part = qs.Sampler(prog, k, numchains=32, seed=3, chain_begin=32); part.init(); part.run(20)
    a = full.fetch_values(); b = part.fetch_values()
    assert np.array_equal(a[32:], b)


def test_device_diagnostics_match_numpy(qs):
    """qsb_sampler_diagnostics (device reduction to plain sums) against a numpy evaluation of the fetched draws,
    for both sample layouts (thread-per-chain and warp-per-chain)."""
    from quicksand_b200 import diagnostics as D
    from test_multirank_gloo import raw_from_draws
    for prog, mapping in [(_programs(qs)["funnel12"], 1), (_programs(qs)["funnel70"], 32)]:
        s = qs.Sampler(prog, qs.HMCKernel(numSteps=4, stepSize=0.05, doStepSizeAdapt=False), numchains=48, seed=5, mapping=mapping)
        s.init(); s.run(60, burnin=10, lag=2)
        draws = s.fetch_values()
        assert draws.shape == (48, 60, prog.dim)
        raw = s.diagnostics_raw(maxlag=20)
        ref = raw_from_draws(draws, 20)
        assert np.allclose(raw, ref, rtol=1e-9, atol=1e-9), np.abs(raw - ref).max()
        summ = D.summarize(raw, prog.dim, 20)
        assert np.all(np.isfinite(summ["ess"])) and np.all(summ["ess"] > 0)
        st = s.stats()
        assert st["iterations"] == 130 and st["props_made"] == 48 * 130
        s.close()


def test_sample_struct_layout_and_lag(qs):
    """fetch_samples fills S.Vector(Sample(T)) elements {value[retdim], logprob, loglikelihood} (infer.t:13-18);
    iteration i is kept iff i >= burnin and i % lag == 0 (mcmc.t:64)."""
    prog = _programs(qs)["gauss10"]
    s = qs.Sampler(prog, qs.DriftKernel(scale=0.2), numchains=8, seed=1, flags=2)
    s.init(); s.run(7, burnin=5, lag=3)     # iters = 26; kept i = 6, 9, ..., 24 -> 7 samples
    rec = s.fetch_record()
    smp = s.fetch_samples()
    assert smp.shape == (8, 7) and smp.dtype.itemsize == 8 * 3
    kept = [i for i in range(26) if i >= 5 and i % 3 == 0]
    assert len(kept) == 7
    vals = rec["state"][:, kept, :].sum(axis=2) / 10.0
    assert np.allclose(smp["value"][:, :, 0], vals, rtol=1e-13)
    assert np.all(smp["loglikelihood"] == 0.0)
    s.close()


def test_queries_on_device_match_reference_arithmetic(qs, oracle):
    """Expectation / Autocorrelation / MAP (infer.t:146-264) run as CUDA kernels: on a hand-made S.Vector(Sample) handed
    over from the host, and on the device-resident samples of a sampler; both against the oracle's restatement
    (mean from value[0] with weights summed from i = 1, variance / N)."""
    rng = np.random.default_rng(0)
    dt = np.dtype([("value", np.float64, (3,)), ("logprob", np.float64), ("loglikelihood", np.float64)])
    data = np.zeros((2, 17), dtype=dt)
    data["value"] = rng.normal(size=(2, 17, 3))
    data["logprob"] = rng.normal(size=(2, 17))
    data["logprob"][1, 5] = data["logprob"][1, 9] = 10.0          # tie: MAP keeps the first
    vec = qs.SampleVector(data)
    m, v = qs.Expectation(True)(None)(vec)
    for c in range(2):
        mo, vo = oracle.expectation(data["value"][c], do_variance=True)
        assert np.array_equal(m[c], mo) and abs(v[c] - vo) < 1e-14
    ac = qs.Autocorrelation()(None)(vec)
    for c in range(2):
        assert np.allclose(ac[c], oracle.autocorrelation(data["value"][c]), rtol=0, atol=1e-13)
    ac2 = qs.Autocorrelation([0.0, 0.0, 0.0], 3.0)(None)(vec)
    v0 = data["value"][0]
    want = [(v0[:17 - t] * v0[t:]).sum() / ((17 - t) * 3.0) for t in range(17)]
    assert np.allclose(ac2[0], want, rtol=0, atol=1e-13)
    mp = qs.MAP(None)(vec)
    assert np.array_equal(mp[0], data["value"][0, np.argmax(data["logprob"][0])]) and np.array_equal(mp[1], data["value"][1, 5])
    with pytest.raises(AssertionError):
        qs.Autocorrelation(mean=0.0)
    # device-resident: both sample layouts (thread per chain / warp per chain) and the GLM path
    for name, mapping in (("gauss10", 0), ("mixed10", 32), ("schools13", 32), ("logistic20", 0)):
        prog = _programs(qs)[name]
        s = qs.Sampler(prog, qs.HMCKernel(numSteps=3, stepSize=0.05, doStepSizeAdapt=False), numchains=5, seed=4, mapping=mapping)
        s.init(); s.run(23, burnin=3, lag=2)
        vec = qs.SampleVector(s.fetch_samples(), sampler=s)
        host = qs.SampleVector(vec.data)
        md, vd = qs.Expectation(True)(prog)(vec)
        mh, vh = qs.Expectation(True)(prog)(host)
        assert np.array_equal(md, mh) and np.array_equal(vd, vh)
        vals = vec.value
        for c in range(5):
            mo, vo = oracle.expectation(vals[c], do_variance=True)
            assert np.array_equal(np.reshape(md[c], -1), mo) and abs(vd[c] - vo) <= 1e-13 * max(1.0, abs(vo))
        assert np.array_equal(qs.Autocorrelation()(prog)(vec), qs.Autocorrelation()(prog)(host))
        assert np.allclose(qs.Autocorrelation()(prog)(vec)[2], oracle.autocorrelation(vals[2]), rtol=1e-12, atol=1e-13)
        assert np.array_equal(qs.MAP(prog)(vec), qs.MAP(prog)(host))
        best = np.argmax(vec.logprob, axis=1)
        assert np.array_equal(np.reshape(qs.MAP(prog)(vec), (5, -1)), vals[np.arange(5), best])
        s.close()


def test_dirichlet_samples_live_on_the_simplex(qs):
    """testsuite.t:693-717 'dirichlet array compile and run (HMC)' with many chains: values are a point of the simplex and
    the pooled mean of Dirichlet(2,3,5) is alpha / sum(alpha).  Long burn-in on purpose: until a chain's first acceptance
    the reference's Hamiltonian uses the float trace's logprob, which lacks the log-Jacobian (hmc.t:144, 179-180),
    so chains initialised near the boundary of the simplex stay put for a while and the transient decays slowly."""
    from quicksand_b200 import _lib as L
    prog = qs.programs.indep([("dirichlet", [2.0, 3.0, 5.0])], ret="vector")
    s = qs.Sampler(prog, qs.HMCKernel(numSteps=5, stepSize=0.3, doStepSizeAdapt=False), numchains=2048, seed=3)
    s.init(); s.run(40, burnin=400)
    v = s.fetch_values()
    s.close()
    assert np.abs(v.sum(axis=2) - 1.0).max() < 1e-12 and v.min() > 0.0
    assert np.abs(v.mean(axis=(0, 1)) - np.array([0.2, 0.3, 0.5])).max() < 0.01
    # the warp-per-chain mapping takes vector choices too (tests/test_gpu_vector_warp.py): same simplex, same chains
    s = qs.Sampler(prog, qs.HMCKernel(), numchains=4, mapping=32)
    s.init(); s.run(20, 5, 1)
    v = s.fetch_values(); s.close()
    assert np.abs(v.sum(axis=2) - 1.0).max() < 1e-12 and (v >= 0).all()


def test_reference_compile_and_run_programs(qs, oracle):
    """The reference's compile-and-run tests on the hot path, through the mirrored interface with their own parameters
    (numsamps = 150): "dirichlet array compile and run (HMC)" (testsuite.t:693-702, HMCKernel() defaults) and "drift vector
    compile and run (lexical state sharing)" (testsuite.t:886-897: two dirichlet choices from two call sites)."""
    prog = qs.programs.indep([("dirichlet", [1.0, 1.0, 1.0, 1.0])], ret="vector")
    samples = qs.infer(prog, qs.Samples(), qs.MCMC(qs.HMCKernel(), dict(numsamps=150, numchains=4, seed=1)))()
    v = samples.value
    assert v.shape == (4, 150, 4) and np.all(np.isfinite(v)) and np.abs(v.sum(axis=2) - 1.0).max() < 1e-12 and v.min() > 0.0
    prog = qs.programs.indep([("dirichlet", [1.0, 1.0, 1.0, 1.0]), ("dirichlet", [1.0, 1.0, 1.0, 1.0])], ret="vector", lexids=[0, 1])
    k = qs.DriftKernel(lexicalScaleSharing=True)
    samples = qs.infer(prog, qs.Samples(), qs.MCMC(k, dict(numsamps=150, numchains=4, seed=1)))()
    v = samples.value
    assert v.shape == (4, 150, 8) and np.all(np.isfinite(v)) and np.abs(v[:, :, :4].sum(axis=2) - 1.0).max() < 1e-12
    gpu, cpu = P.run_both(prog, k, chains=16, iters=150, seed=1)
    P.compare(gpu, cpu, TOL, what="two dirichlets, drift with lexical scale sharing")


def test_error_behaviour(qs):
    """Unsupported requests fail loudly with a message (the shim turns them into the reference's print + abort)."""
    from quicksand_b200 import _lib as L
    with pytest.raises(L.QsbError):   # the tensor-core GLM path runs one kernel, not a mixture
        qs.Sampler(_programs(qs)["logistic20"], qs.MixtureKernel([qs.DriftKernel(), qs.HARMKernel()], [0.5, 0.5]), numchains=4)
    with pytest.raises(L.QsbError):
        qs.Sampler(qs.programs.funnel(5000), qs.HMCKernel(), numchains=4)
    m = qs.MCMC(qs.HMCKernel(numSteps=3), dict(numsamps=20, numchains=3, seed=2))
    mean = qs.infer(_programs(qs)["gaussian1"], qs.Expectation(), m)()
    assert mean.shape == (3,) and np.all(np.isfinite(mean))


# ---- statistical parity (north_star: posterior means and variances agree within 3 Monte-Carlo standard errors)
def _pooled_moments(qs, prog, kernel, chains, nsamps, burnin, lag=1, seed=77):
    from quicksand_b200 import diagnostics as D
    s = qs.Sampler(prog, kernel, numchains=chains, seed=seed)
    s.init(); s.run(nsamps, burnin, lag)
    raw = s.diagnostics_raw(maxlag=min(nsamps - 1, 60))
    d = D.summarize(raw, prog.retdim, min(nsamps - 1, 60))
    vals = s.fetch_values()
    s.close()
    return d, vals


REF_TRUTHS = [("gaussian1", 0.1, 0.25), ("uniform1", 0.25, 0.3 ** 2 / 12.0), ("gamma1", 0.4, 8.0 / 100.0), ("beta1", 2.0 / 7.0, 10.0 / (49.0 * 8.0))]


def _oracle_moments(oracle, prog, kernel, chains, nsamps, burnin, lag, seed=177):
    from quicksand_b200 import diagnostics as D
    out = oracle.run(P.oracle_program(prog), P.oracle_specs(kernel), seed, chains, nsamps, burnin, lag)
    v = out["samples"][:, :, 0]
    ess = max(D.ess_numpy(v, min(nsamps - 1, 60)), 50.0)
    return v.mean(), v.var(), ess, v


@pytest.mark.parametrize("name,mean,var", REF_TRUTHS, ids=[t[0] for t in REF_TRUTHS])
@pytest.mark.parametrize("L", [1, 10])
def test_reference_expectation_truths_on_device(qs, oracle, name, mean, var, L):
    """The reference's own HMC expectation tests (testsuite.t:621-691: gaussian 0.1, uniform 0.25, gamma(2,2)/10 -> 0.4,
    beta(2,5) -> 2/7; Langevin L=1 and HMC L=10) with 2048 chains on the device.

    (i) the truth within the reference's own tolerance 0.07 (testsuite.t:84-87);
    (ii) posterior mean AND variance within 3 Monte-Carlo standard errors of the CPU oracle's (1024 independent chains of
    the same algorithm, different seed).  The comparison is against the oracle, not the analytic truth, because the
    reference's dual averaging never stops (hmc.t:159-161): the step size keeps reacting to the chain, the kernel is not
    exactly invariant, and the reference algorithm itself is biased at the 1e-2 level (e.g. variance 0.227 instead of 0.25
    for gaussian(0.1, 0.5) with L=10) -- both implementations show the same bias."""
    prog = _programs(qs)[name]
    k = qs.HMCKernel(numSteps=L)
    d, vals = _pooled_moments(qs, prog, k, chains=2048, nsamps=150, burnin=300, lag=2)
    assert abs(d["mean"][0] - mean) < 0.07
    om, ov, oess, ovals = _oracle_moments(oracle, prog, k, 1024, 150, 300, 2)
    ess = max(d["ess"][0], 100.0)
    se_mean = np.sqrt(ov / ess + ov / oess)
    m4 = np.var((ovals - om) ** 2)
    se_var = np.sqrt(m4 / ess + m4 / oess)
    assert abs(d["mean"][0] - om) <= 3 * se_mean, (d["mean"][0], om, se_mean)
    assert abs(d["var"][0] - ov) <= 3 * se_var, (d["var"][0], ov, se_var)
    assert d["rhat"][0] < 1.1


@pytest.mark.parametrize("name,mean,var", REF_TRUTHS, ids=[t[0] for t in REF_TRUTHS])
def test_fixed_step_hmc_hits_the_truth(qs, name, mean, var):
    """Sanity of the transition itself: with adaptation off the kernel is a valid MCMC kernel and reproduces the analytic
    mean (4 MCSE) and variance (4 MCSE or 3 %) of the reference's test programs."""
    prog = _programs(qs)[name]
    eps = {"gaussian1": 0.3, "uniform1": 0.6, "gamma1": 0.35, "beta1": 0.35}[name]
    d, vals = _pooled_moments(qs, prog, qs.HMCKernel(numSteps=5, stepSize=eps, doStepSizeAdapt=False), chains=4096, nsamps=200, burnin=600, lag=3)
    ess = max(min(d["ess"][0], 4096 * 200), 100.0)
    assert abs(d["mean"][0] - mean) <= 4 * np.sqrt(var / ess) + 1e-12, (d["mean"][0], mean, ess)
    se_var = np.sqrt(np.var((vals[:, :, 0] - mean) ** 2) / ess)
    assert abs(d["var"][0] - var) <= max(4 * se_var, 0.03 * var), (d["var"][0], var, se_var)
    assert d["rhat"][0] < 1.1


@pytest.mark.parametrize("kind", ["hmc", "harm", "drift", "drift_lexical"])
def test_ten_gaussians_truth_on_device(qs, oracle, kind):
    """testsuite.t:719-730, 776-787, 812-823, 846-860 with many chains: mean of 10 iid gaussian(0.1, 0.5) -> 0.1 within the
    reference's tolerance, and mean/variance within 3 MCSE of the oracle's."""
    prog = _programs(qs)["gauss10loop" if kind == "drift_lexical" else "gauss10"]
    kernel = {"hmc": qs.HMCKernel(numSteps=10), "harm": qs.HARMKernel(), "drift": qs.DriftKernel(),
              "drift_lexical": qs.DriftKernel(lexicalScaleSharing=True)}[kind]
    burnin, lag = (300, 2) if kind == "hmc" else (1500, 20)
    d, vals = _pooled_moments(qs, prog, kernel, chains=2048, nsamps=100, burnin=burnin, lag=lag)
    assert abs(d["mean"][0] - 0.1) < 0.07
    om, ov, oess, ovals = _oracle_moments(oracle, prog, kernel, 256, 100, burnin, lag)
    ess = max(d["ess"][0], 100.0)
    assert abs(d["mean"][0] - om) <= 3 * np.sqrt(ov / ess + ov / oess), (d["mean"][0], om)
    m4 = np.var((ovals - om) ** 2)
    assert abs(d["var"][0] - ov) <= 3 * np.sqrt(m4 / ess + m4 / oess), (d["var"][0], ov)


def test_correlated_gaussian_moments_and_oracle_agreement(qs, oracle):
    """config c2 under the adaptive kernel: pooled means and variances of all 10 coordinates within 3 combined MCSE of the
    CPU oracle's (independent chains of the same algorithm), and the empirical covariance close to rho^|i-j|."""
    prog, Sigma = qs.programs.c2_correlated_gaussian(10, 0.9)
    k = qs.HMCKernel(numSteps=16)
    d, vals = _pooled_moments(qs, prog, k, chains=4096, nsamps=100, burnin=300)
    out = oracle.run(P.oracle_program(prog), P.oracle_specs(k), 5, 512, 100, 300, 1)
    o = out["samples"]
    from quicksand_b200 import diagnostics as D
    for j in range(10):
        oess = max(D.ess_numpy(o[:, :, j], 60), 50.0)
        ov = o[:, :, j].var()
        se = np.sqrt(ov / max(d["ess"][j], 100.0) + ov / oess)
        assert abs(d["mean"][j] - o[:, :, j].mean()) <= 3 * se, (j, d["mean"][j], o[:, :, j].mean(), se)
        m4 = np.var((o[:, :, j] - o[:, :, j].mean()) ** 2)
        # 20 simultaneous comparisons (10 coordinates x {mean, variance}): 4 SE for the variances, whose effective
        # sample size is smaller than the mean's
        assert abs(d["var"][j] - ov) <= 4 * np.sqrt(m4 / max(d["ess"][j], 100.0) + m4 / oess), (j, d["var"][j], ov)
    emp = np.cov(vals.reshape(-1, 10).T)
    assert np.abs(emp - Sigma).max() < 0.15      # the never-ending adaptation biases the reference algorithm itself


def test_cxx_mirror_runs_the_reference_test(qs, tmp_path):
    """The C++ mirror (include/qsb200.hpp) end to end on the device: the reference's 10-gaussian HMC expectation test."""
    import os, subprocess
    from quicksand_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "example"
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cxx", "example.cpp"),
                    "-o", str(exe), "-L", libdir, "-lqsb200", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("kind", ["hmc", "tracemh"])
def test_latent_parameter_truths_on_device(qs, kind):
    """Hierarchical programs written with plain ERPs whose parameters are earlier choices:
    mu ~ gaussian(1, 0.5); x ~ gaussian(0.5 + 2 mu, 1) -> E[x] = 2.5, Var[x] = 1 + 4 * 0.25 = 2;
    k ~ gamma(3, 1); y ~ gamma(k, 2) -> E[y] = 6 (latent shape: the gradient runs through the 5-coefficient log-gamma)."""
    lat = qs.programs.latent
    k = qs.HMCKernel(numSteps=10) if kind == "hmc" else qs.TraceMHKernel(doStruct=False)
    g = qs.programs.indep([("gaussian", 1.0, 0.5), ("gaussian", lat(0, a=0.5, b=2.0), 1.0)], ret="vector")
    d, vals = _pooled_moments(qs, g, k, chains=2048, nsamps=150, burnin=300, lag=2)
    assert abs(d["mean"][1] - 2.5) < 0.07 and abs(d["mean"][0] - 1.0) < 0.07, d["mean"]
    assert abs(d["var"][1] - 2.0) < 0.2, d["var"]
    y = qs.programs.indep([("gamma", 3.0, 1.0), ("gamma", lat(0), 2.0)], ret="vector")
    d, vals = _pooled_moments(qs, y, k, chains=2048, nsamps=150, burnin=300, lag=2)
    assert abs(d["mean"][0] - 3.0) < 0.15 and abs(d["mean"][1] - 6.0) < 0.35, d["mean"]


def test_latent_parameter_descriptor_errors(qs):
    from quicksand_b200 import _lib as L
    lat = qs.programs.latent
    p = qs.programs.indep([("gaussian", 0.0, 1.0), ("gaussian", lat(0), 1.0)], ret="vector")
    p.arrays["erp_pref"][0] = 1                       # choice 0 depending on the later choice 1
    with pytest.raises(L.QsbError):
        p.model()
    p = qs.programs.indep([("gaussian", 0.0, 1.0), ("gaussian", lat(0), 1.0)], ret="vector")
    p.arrays["erp_kind"][1] = L.ERP_UNIFORM          # a uniform's parameters are its bounds: not supported as latent
    with pytest.raises(L.QsbError):
        p.model()
    p = qs.programs.indep([("gaussian", 0.0, 1.0)] + [("gaussian", lat(0), 1.0)] * 40, ret="vector")   # 41 components: warp per chain
    qs.Sampler(p, qs.HMCKernel(), numchains=4).close()
    with pytest.raises(L.QsbError):                   # ... but thread per chain holds at most 32
        qs.Sampler(p, qs.HMCKernel(), numchains=4, mapping=1)
    p = qs.programs.indep([("gaussian", 0.0, 1.0)] + [("gaussian", lat(0), 1.0)] * 140, ret="vector")  # 141 components: beyond the INDEP family
    with pytest.raises(L.QsbError):
        qs.Sampler(p, qs.HMCKernel(), numchains=4)


def test_glm_stream_k_decomposition_parity(qs, oracle, monkeypatch):
    """The gradient kernel's balanced decomposition (the default: equal-cost runs of row-block units per SM, runs crossing
    chain-tile boundaries, a varying number of partial sums per tile, idle tail warps skipped) against the oracle and against
    the (chain tile x row split) grid of QSB_FLAG_INVARIANT.  The two group a chain's row blocks differently: equal to
    rounding, not bit for bit."""
    from quicksand_b200 import _lib as L
    rng = np.random.default_rng(7)
    N, d = 10000, 100
    X = rng.standard_normal((N, d)) / np.sqrt(d)
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(-X @ rng.normal(0, 1, d)))).astype(np.uint8)
    big = qs.programs.logistic(X, y, 2.5)
    k = qs.HMCKernel(numSteps=4, stepSize=0.015, doStepSizeAdapt=False)
    ref = qs.Sampler(big, k, numchains=300, seed=5, flags=L.FLAG_INVARIANT); ref.init(); ref.run(3)
    ref_vals = ref.fetch_values(); ref.close()
    monkeypatch.setenv("QSB_GLM_STREAMK", "1")
    gpu, cpu = P.run_both(big, k, chains=6, iters=4, seed=20261019, leap=True)
    P.compare(gpu, cpu, TOL, leap=True, what="c3 full shape HMC, stream-K")
    sk = qs.Sampler(big, k, numchains=300, seed=5); sk.init(); sk.run(3)     # 4 chain tiles: runs cross tile boundaries
    sk_vals = sk.fetch_values(); sk.close()
    assert P.relerr(sk_vals, ref_vals, P.vector_floor(ref_vals)).max() <= 1e-10
    monkeypatch.delenv("QSB_GLM_STREAMK")
    fl = qs.Sampler(big, k, numchains=300, seed=5); fl.init(); fl.run(3)       # no flag: the balanced decomposition is the default
    assert np.array_equal(fl.fetch_values(), sk_vals) and not np.array_equal(sk_vals, ref_vals)
    fl.close()
    monkeypatch.setenv("QSB_GLM_STREAMK", "0")      # the tuning override selects the grid as well
    g0 = qs.Sampler(big, k, numchains=300, seed=5); g0.init(); g0.run(3)
    assert np.array_equal(g0.fetch_values(), ref_vals)
    g0.close()
    monkeypatch.setenv("QSB_GLM_STREAMK", "1")
    small = _programs(qs)["logistic20"]
    for kern, iters, minc in [(qs.HMCKernel(numSteps=7, stepSize=0.1, doStepSizeAdapt=False), 20, 1.0), (qs.DriftKernel(scale=0.3), 100, 1.0),
                              (qs.HARMKernel(), 60, 1.0), (qs.HMCKernel(numSteps=1), 15, 0.8)]:
        gpu, cpu = P.run_both(small, kern, chains=200, iters=iters, seed=3)      # 200 chains of 333 rows: 3 tiles x 11 row blocks in runs of 12
        P.compare(gpu, cpu, TOL, what="logistic20 stream-K %d" % kern.specs[0].kind, min_compared=minc)


def test_conditions_hold_and_reject(qs, oracle):
    """qs.condition (trace.t:794-796): every chain starts inside the constraints (rejection initialisation, trace.t:612-623,
    bit-equal to the oracle's retries), never leaves them under any kernel, and proposals that violate them are rejected
    whatever their density (the accept decisions equal the oracle's, which evaluates `conditionsSatisfied and ...`)."""
    prog = _programs(qs)["cond5"]
    lo = np.array([-0.5, -np.inf, 0.2, -np.inf, -1.0]); hi = np.array([np.inf, -0.1, 9.0, np.inf, 3.0])
    for kern, mapping in [(qs.HMCKernel(numSteps=5, stepSize=0.4, doStepSizeAdapt=False), 0), (qs.DriftKernel(scale=1.5), 0), (qs.DriftKernel(scale=1.5), 32),
                          (qs.HARMKernel(scale=3.0, doScaleAdapt=False), 32), (qs.TraceMHKernel(doStruct=False), 0),
                          (qs.MixtureKernel([qs.HARMKernel(), qs.HMCKernel(numSteps=4, stepSize=0.3, doStepSizeAdapt=False), qs.DriftKernel()], [0.3, 0.4, 0.3]), 0)]:
        gpu, cpu = P.run_both(prog, kern, chains=96, iters=150, seed=8, mapping=mapping)
        P.compare(gpu, cpu, TOL, what="conditions %s map%d" % ([s.kind for s in kern.specs], mapping))
        v = gpu["samples"]
        assert (v > lo).all() and (v < hi).all()
        assert 0.05 < gpu["accept"].mean() < 0.98
    # without the conditions the same kernel wanders outside: the constraint is what keeps it in
    free = qs.programs.indep([("gaussian", 0.0, 1.0), ("neggamma", 2.0, 1.5), ("gamma", 2.0, 2.0), ("uniform", -1.0, 1.0), ("gaussian", 1.0, 2.0)], ret="vector")
    s = qs.Sampler(free, qs.DriftKernel(scale=1.5), numchains=96, seed=8); s.init(); s.run(150)
    fv = s.fetch_values(); s.close()
    assert not ((fv > lo).all() and (fv < hi).all())


def test_chain_health_counters(qs):
    """qsb_stats.nonfinite / divergent: a stable run reports none; a funnel driven with a step size far beyond
    its stability limit reports divergent (|dH| > 1000) and non-finite transitions, on the engine and on the tensor-core path."""
    s = qs.Sampler(_programs(qs)["gauss10"], qs.HMCKernel(numSteps=10, stepSize=0.2, doStepSizeAdapt=False), numchains=256, seed=1)
    s.init(); s.run(50)
    st = s.stats(); s.close()
    assert st["nonfinite"] == 0 and st["divergent"] == 0
    for mapping in (1, 32):
        s = qs.Sampler(_programs(qs)["funnel12"], qs.HMCKernel(numSteps=16, stepSize=3.0, doStepSizeAdapt=False), numchains=256, seed=1, mapping=mapping)
        s.init(); s.run(50)
        st = s.stats(); s.close()
        assert st["divergent"] + st["nonfinite"] > 1000 and st["props_accepted"] < st["props_made"]
    s = qs.Sampler(_programs(qs)["logistic20"], qs.HMCKernel(numSteps=20, stepSize=5.0, doStepSizeAdapt=False), numchains=192, seed=1)
    s.init(); s.run(30)
    st = s.stats(); s.close()
    assert st["divergent"] + st["nonfinite"] > 100


def test_single_chain_reference_protocol(qs, oracle):
    """Config c1: with numchains = 1 (the default) qs.MCMC is the reference's single chain.  The reference's own protocol --
    150 samples x lag 20, 5 runs, mean absolute error <= 0.07 (testsuite.t:84-141) -- on its HMC programs (testsuite.t:621-691,
    719-730), and that one chain is the oracle's chain step for step."""
    for name, truth in [("gaussian1", 0.1), ("gauss10", 0.1), ("gamma1", 0.4), ("beta1", 2.0 / 7.0)]:
        prog = _programs(qs)[name]
        errs = []
        for r in range(5):
            m = qs.MCMC(qs.HMCKernel(numSteps=10), dict(numsamps=150, lag=20, seed=100 + r))
            est = qs.infer(prog, qs.Expectation(), m)()
            assert np.size(est) == 1                       # one chain: the reference's scalar, not a per-chain vector
            errs.append(abs(float(np.ravel(est)[0]) - truth))
        assert np.mean(errs) <= 0.07, (name, errs)
    k = qs.HMCKernel(numSteps=10, stepSize=0.2, doStepSizeAdapt=False)
    gpu, cpu = P.run_both(_programs(qs)["gauss10"], k, chains=1, iters=200, seed=100, leap=True)
    P.compare(gpu, cpu, TOL, leap=True, what="single chain HMC gauss10")
    gpu, cpu = P.run_both(_programs(qs)["funnel12"], qs.MixtureKernel([qs.HARMKernel(), qs.DriftKernel(scale=0.3)], [0.5, 0.5]), chains=1, iters=300, seed=101)
    P.compare(gpu, cpu, TOL, what="single chain HARM+Drift funnel12")
