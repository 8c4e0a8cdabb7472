"""Teacher-forced parity of the HMC leapfrog on the headline configurations: the oracle's recorded
states of ADAPTIVE runs -- BASELINE config c3 at its full shape with L = 20, config c5's and c4's models at full
dimension, config c2 -- are replayed step by step through ``qsb_debug_leapfrog``.  Gates (tests/parity.py: teacher_forced_hmc): every leapfrog
step of every non-divergent trajectory within 1e-10 relative (positions, momenta, log-density; gradient 1e-9), and the accept
decision of EVERY trajectory identical.  No prefix: a step is compared from the oracle's own input state, wherever it sits in
the run.  Divergent trajectories (oracle |dH| > 30 or overflow: the adaptation transient's rejected excursions, where the
reference's own naive sigmoid / exp arithmetic cancels or overflows) are replayed as well, their errors reported, their
accept decisions gated."""
import numpy as np
import pytest

import parity as P

pytestmark = pytest.mark.gpu


def test_c3_full_shape_adaptive_L20_every_leapfrog_step(qs, oracle):
    """Bayesian logistic regression d = 100, N = 10 000, HMCKernel{numSteps=20} with the reference's step-size search and
    dual averaging (hmc.t:345-402) -- the configuration bench.py times.  8 chains x 16 transitions x 20 steps on the oracle
    (tape AD over 2 M nodes per gradient), every step replayed through the tensor-core gradient kernel."""
    X, y, _ = qs.programs.c3_logistic_data(10000, 100)
    prog = qs.programs.logistic(X, y, 2.5)
    r = P.teacher_forced_hmc(prog, qs.HMCKernel(numSteps=20), chains=8, iters=16, seed=20261019, what="c3 full shape adaptive L=20")
    assert r["stable"] >= 24


@pytest.mark.parametrize("mapping", [32])
def test_c5_funnel_1000_adaptive_every_leapfrog_step(qs, oracle, mapping):
    """config c5's model at full dimension (1000-D funnel) under its HMC sub-kernel, adaptive, L = 16: the scalar-gradient
    warp-per-chain trajectory code."""
    prog = qs.programs.funnel(1000)
    P.teacher_forced_hmc(prog, qs.HMCKernel(numSteps=16), chains=12, iters=12, seed=20261019, what="c5 funnel-1000 adaptive L=16")


def test_c4_eight_schools_1002_adaptive_every_leapfrog_step(qs, oracle):
    """config c4's model at full dimension (J = 1000, d = 1002) under adaptive HMC (c4's own kernel, drift, has no leapfrog;
    its per-step parity is test_c4_full_dimension_drift)."""
    y, sigma = qs.programs.c4_eight_schools_data(1000)
    prog = qs.programs.eight_schools(y, sigma)
    P.teacher_forced_hmc(prog, qs.HMCKernel(numSteps=8), chains=12, iters=10, seed=20261019, what="c4 schools-1002 adaptive L=8")


def test_c2_adaptive_L16_every_leapfrog_step(qs, oracle):
    prog, _ = qs.programs.c2_correlated_gaussian(10, 0.9)
    P.teacher_forced_hmc(prog, qs.HMCKernel(numSteps=16), chains=64, iters=40, seed=20261019, what="c2 adaptive L=16")


SMALL = [("gaussian1", 10, 0), ("gamma1", 10, 0), ("beta1", 10, 0), ("uniform1", 10, 0), ("gauss10", 10, 0), ("mixed10", 5, 32), ("mixed10", 5, 1),
         ("dirichlet4", 10, 0), ("schools13", 8, 32), ("schools13", 8, 1), ("funnel12", 16, 1), ("funnel12", 16, 32), ("logistic20", 20, 0),
         ("hier12", 2, 0), ("c2", 16, 4), ("upper5", 5, 0), ("upper5", 5, 32)]


@pytest.mark.parametrize("name,L,mapping", SMALL)
def test_small_programs_adaptive_every_leapfrog_step(qs, oracle, monkeypatch, name, L, mapping):
    """The adaptive configurations test_hmc_adaptive_search_and_dual_averaging can only compare on a prefix (down to 1 % of
    the chain-steps for logistic20 L=20), here on 100 % of their leapfrog steps and accept decisions, in both mappings."""
    from test_gpu_parity import _programs
    if mapping:
        monkeypatch.setenv("QSB_TEST_MAPPING", str(mapping))
    prog = _programs(qs)[name]
    P.teacher_forced_hmc(prog, qs.HMCKernel(numSteps=L), chains=48, iters=25, seed=7, what="%s adaptive L=%d map%d" % (name, L, mapping))
