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


def test_c3_full_shape_working_regime_ties_the_device_run_to_the_oracle(qs, oracle):
    """The headline configuration where it spends its time: AFTER the adaptation transient.  A free-running device run of c3
    (full shape, adaptive, L = 20, the bench's 150 burn-in transitions) is recorded; for each of its next transitions the
    oracle side re-computes the whole trajectory from the DEVICE's own state and step size -- momenta from the shared Philox
    stream, `HMCKernel:step` (hmc.t:307-323) in the reference's operation order on the oracle's tape-AD gradient -- and

    * every one of the 20 leapfrog steps is replayed on the device from the oracle's input (qsb_debug_leapfrog: the tensor-core
      gradient kernel) and agrees within 1e-10 (positions, momenta, log-density; gradient 1e-9): no trajectory is exempt;
    * the accept decision the device run actually took, its energy error and the state it moved to equal the oracle's.

    Nothing here depends on a prefix or on a divergence filter: in the working regime |dH| is of order one."""
    from concurrent.futures import ThreadPoolExecutor
    from quicksand_b200 import _lib as L
    X, y, _ = qs.programs.c3_logistic_data(10000, 100)
    prog = qs.programs.logistic(X, y, 2.5)
    C, B, T, Ls, seed, d = 16, 150, 4, 20, 20261019, 100
    s = qs.Sampler(prog, qs.HMCKernel(numSteps=Ls), numchains=C, seed=seed, flags=L.FLAG_RECORD)
    s.init(); s.run(T, B, 1)
    rec = s.fetch_record(); s.close()
    state, step, acc, dH = rec["state"], rec["step"], rec["accept"], rec["dH"]

    def seqsum(v):
        t = 0.0
        for a in v:
            t = t + a
        return t

    def one(c):
        op = P.oracle_program(prog)          # (the oracle's tape is thread-local)
        out = []
        for t in range(B, B + T):
            x, eps = state[c, t - 1].copy(), float(step[c, t])
            p = oracle.gaussians(seed, c, t, 0, d)
            lp, g, _ = op.logp_grad(x)
            H0 = -0.5 * seqsum(p * p * 1.0) + lp
            xin, pin, gin, xo, po, go = [], [], [], [], [], []
            for _ in range(Ls):
                xin.append(x.copy()); pin.append(p.copy()); gin.append(g.copy())
                p = p + (0.5 * eps) * g
                x = x + (eps * p) * 1.0
                lp, g, _ = op.logp_grad(x)
                p = p + (0.5 * eps) * g
                xo.append(x.copy()); po.append(p.copy()); go.append(g.copy())
            dh = (-0.5 * seqsum(p * p * 1.0) + lp) - H0
            u = oracle.uniforms(seed, c, t, d, 1)[0]
            out.append(dict(c=c, t=t, eps=eps, xin=np.array(xin), pin=np.array(pin), gin=np.array(gin), xo=np.array(xo), po=np.array(po), go=np.array(go),
                            lp=lp, dH=dh, accept=bool(np.log(u) < dh), x_start=state[c, t - 1]))
        return out
    with ThreadPoolExecutor(max_workers=16) as ex:
        trajs = [tr for chain in ex.map(one, range(C)) for tr in chain]
    worst = dict(x=0.0, p=0.0, g=0.0, lp=0.0, dH=0.0, state=0.0)
    flips = 0
    for last in (False, True):
        sl = slice(Ls - 1, Ls) if last else slice(0, Ls - 1)
        xin = np.concatenate([tr["xin"][sl] for tr in trajs]); pin = np.concatenate([tr["pin"][sl] for tr in trajs]); gin = np.concatenate([tr["gin"][sl] for tr in trajs])
        eps = np.concatenate([np.full(xin.shape[0] // len(trajs), tr["eps"]) for tr in trajs])
        r = prog.debug_leapfrog(xin, pin, eps, g=gin, last=last)
        for key, ref in (("x", "xo"), ("p", "po"), ("g", "go")):
            want = np.concatenate([tr[ref][sl] for tr in trajs])
            worst[key] = max(worst[key], float(P.relerr(r[key], want, P.vector_floor(want)).max()))
        if last:
            worst["lp"] = float(P.relerr(r["lp"], np.array([tr["lp"] for tr in trajs])).max())
    for tr in trajs:
        c, t = tr["c"], tr["t"]
        flips += int(bool(acc[c, t]) != tr["accept"])
        worst["dH"] = max(worst["dH"], abs(dH[c, t] - tr["dH"]) / max(1.0, abs(tr["lp"])))        # a difference of two Hamiltonians of size |lp|
        want = tr["xo"][-1] if tr["accept"] else tr["x_start"]
        worst["state"] = max(worst["state"], float(P.relerr(state[c, t], want, P.vector_floor(want)).max()))
    dhs = np.array([tr["dH"] for tr in trajs])
    msg = ("c3 full shape, working regime: %d trajectories x %d steps, x %.2e p %.2e g %.2e lp %.2e; device run vs oracle: dH %.2e (rel. to H) state %.2e "
           "accept flips %d; accept rate %.2f, |dH| max %.2f, eps range [%.3g, %.3g]" % (
               len(trajs), Ls, worst["x"], worst["p"], worst["g"], worst["lp"], worst["dH"], worst["state"], flips, float(np.mean([tr["accept"] for tr in trajs])),
               float(np.abs(dhs).max()), min(tr["eps"] for tr in trajs), max(tr["eps"] for tr in trajs)))
    print(msg)
    assert np.abs(dhs).max() < 30.0, msg                      # the working regime: no divergent trajectory among them
    assert worst["x"] <= 1e-10 and worst["p"] <= 1e-10 and worst["lp"] <= 1e-10 and worst["g"] <= 1e-9, msg
    assert flips == 0 and worst["dH"] <= 1e-10 and worst["state"] <= 1e-10, msg
