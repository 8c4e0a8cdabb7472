"""Multi-GPU behind the C ABI (SURVEY.md 8b Threading / 8e): ``qsb_group`` drives several shards of one chain population
from one host thread; ``qsb_comm`` is the one-process-per-device shape.  Chains are keyed by their global id, so a sharded
run must reproduce the unsharded one chain for chain; the diagnostics' sufficient statistics are all-reduced inside the
library (NCCL when the shards sit on different devices, a device kernel when a test box has a single GPU)."""
import numpy as np
import pytest

import parity as P

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _cases(qs):
    rng = np.random.default_rng(11)
    X = rng.standard_normal((200, 12)) / np.sqrt(12)
    y = (rng.random(200) < 0.5).astype(np.uint8)
    return [
        ("funnel mixture", qs.programs.funnel(40), qs.MixtureKernel([qs.HARMKernel(), qs.HMCKernel(numSteps=6)], [0.5, 0.5]), 203),
        ("c2 HMC", qs.programs.c2_correlated_gaussian(10, 0.9)[0], qs.HMCKernel(numSteps=8), 260),
        ("schools drift", qs.programs.eight_schools(np.linspace(-3, 9, 30), np.linspace(8, 18, 30)), qs.DriftKernel(scale=0.3), 131),
        ("logistic HMC (tensor-core path)", qs.programs.logistic(X, y, 2.5), qs.HMCKernel(numSteps=5), 192),
    ]


def _check_group_equals_single(qs, devices):
    from quicksand_b200 import diagnostics as D
    for name, prog, kernel, chains in _cases(qs):
        single = qs.Sampler(prog, kernel, numchains=chains, seed=31, device=devices[0])
        single.init(); single.run(30, 10, 2)
        ref = single.fetch_samples()
        ref_raw = single.diagnostics_raw(maxlag=10)
        ref_st = single.stats()
        single.close()
        g = qs.Group(prog, kernel, devices, numchains=chains, seed=31)
        assert g.size() == len(devices)
        g.init(); g.run(30, 10, 2)
        got = g.fetch_samples()
        raw = g.diagnostics_raw(maxlag=10)
        st = g.stats()
        g.close()
        assert np.array_equal(got["value"], ref["value"]), name          # chain for chain, bit for bit
        assert np.array_equal(got["logprob"], ref["logprob"]), name
        for k in ("props_made", "props_accepted", "leapfrog_steps", "grad_evals"):
            assert st[k] == ref_st[k], (name, k, st[k], ref_st[k])
        # sufficient statistics: plain sums accumulated in a different order (atomics, shards)
        assert P.relerr(raw, ref_raw, 1e-9 * np.abs(ref_raw).max()).max() <= 1e-9, name
        a, b = D.summarize(raw, prog.retdim, 10), D.summarize(ref_raw, prog.retdim, 10)
        assert np.allclose(a["rhat"], b["rhat"], rtol=1e-9, equal_nan=True) and np.allclose(a["mean"], b["mean"], rtol=1e-9, atol=1e-12)


def test_group_of_two_shards_on_one_device_equals_the_unsharded_run(qs):
    _check_group_equals_single(qs, [0, 0])
    _check_group_equals_single(qs, [0, 0, 0])       # uneven block partition (203 = 68 + 68 + 67)


@pytest.mark.skipif("_ndev() < 2")
def test_group_of_two_devices_equals_the_unsharded_run(qs):
    """One host thread, two B200s, the library's own NCCL communicator."""
    _check_group_equals_single(qs, [0, 1])


def test_group_through_the_mcmc_interface(qs):
    """qs.MCMC(kernel, {numchains=, devices=}) (qs/mcmc.t:36-43 parameter table + devices) returns the same sample vector."""
    prog = qs.programs.indep([("gaussian", 0.1, 0.5)] * 10, ret_div=10.0)
    k = qs.HMCKernel(numSteps=10)
    a = qs.MCMC(k, numsamps=40, lag=2, numchains=64, seed=3)(prog)()
    b = qs.MCMC(k, numsamps=40, lag=2, numchains=64, seed=3, devices=[0, 0])(prog)()
    assert np.array_equal(a.value, b.value) and np.array_equal(a.logprob, b.logprob)
    m = qs.Expectation()(prog)(b)           # queries run on the host copy of a group's samples
    assert m.shape == (64,)


def test_streaming_moments_match_stored_draws(qs):
    """QSB_FLAG_MOMENTS: running statistics in the transition kernels instead of stored draws.  The same run with and without
    the flag (same seed = same draws): pooled mean / variance and split-R-hat from the 56 bytes per (chain, dimension) equal
    those from the stored draws; the batch-means ESS agrees with Geyer's within a factor of two; in every sample layout
    (thread per chain, warp per chain, tensor-core path) and through a sharded group."""
    from quicksand_b200 import _lib as L, diagnostics as D
    rng = np.random.default_rng(11)
    X = rng.standard_normal((200, 12)) / np.sqrt(12)
    y = (rng.random(200) < 0.5).astype(np.uint8)
    cases = [(qs.programs.c2_correlated_gaussian(10, 0.9)[0], qs.HMCKernel(numSteps=8, stepSize=0.15, doStepSizeAdapt=False), 0),
             (qs.programs.funnel(40), qs.MixtureKernel([qs.HARMKernel(), qs.HMCKernel(numSteps=6, stepSize=0.05, doStepSizeAdapt=False)], [0.5, 0.5]), 32),
             (qs.programs.indep([("gaussian", 0.1, 0.5)] * 10, ret_div=10.0), qs.DriftKernel(), 0),
             (qs.programs.logistic(X, y, 2.5), qs.HMCKernel(numSteps=5, stepSize=0.1, doStepSizeAdapt=False), 0)]
    for prog, kern, mapping in cases:
        a = qs.Sampler(prog, kern, numchains=256, seed=9, mapping=mapping); a.init(); a.run(400, 50, 1)
        ref = D.summarize(a.diagnostics_raw(maxlag=60), prog.retdim, 60); a.close()
        b = qs.Sampler(prog, kern, numchains=256, seed=9, mapping=mapping, flags=L.FLAG_MOMENTS); b.init(); b.run(400, 50, 1)
        got = D.summarize_moments(b.diagnostics_raw(), prog.retdim)
        with pytest.raises(L.QsbError):
            b.fetch_samples()                     # nothing is stored
        b.close()
        assert np.allclose(got["mean"], ref["mean"], rtol=1e-9, atol=1e-12) and np.allclose(got["var"], ref["var"], rtol=1e-9)
        assert np.allclose(got["rhat"], ref["rhat"], rtol=1e-9)
        # batch means (batch = 20 draws here) estimate the asymptotic variance when the autocorrelation time is well below the
        # batch length: there the two ESS estimates agree within a factor of two; on slowly mixing dimensions (the funnel
        # under a small fixed step) batch means are biased low, the ESS is an upper bound and is flagged as such
        ratio = got["ess"] / ref["ess"]
        fast = ref["tau"] <= got["batch"] / 4.0
        assert (ratio[fast] > 0.4).all() and (ratio[fast] < 2.5).all(), ratio
        assert (ratio[~fast] > 0.4).all(), (ratio, ref["tau"])
        assert got["batch_too_short"].dtype == bool and not got["batch_too_short"][ref["tau"] < 1.5].any()
        g = qs.Group(prog, kern, [0, 0], numchains=256, seed=9, flags=L.FLAG_MOMENTS); g.init(); g.run(400, 50, 1)
        gg = D.summarize_moments(g.diagnostics_raw(), prog.retdim); g.close()
        if mapping == 0:      # (the group picks its own mapping)
            assert np.allclose(gg["rhat"], got["rhat"], rtol=1e-9) and np.allclose(gg["ess"], got["ess"], rtol=1e-9)


def test_comm_single_rank_allreduce_is_identity(qs):
    """qsb_comm with one rank: the NCCL communicator is created and the all-reduced diagnostics equal the local ones."""
    prog = qs.programs.funnel(12)
    s = qs.Sampler(prog, qs.HARMKernel(), numchains=64, seed=2)
    s.init(); s.run(20)
    c = qs.Comm(qs.Comm.unique_id(), 0, 1)
    a = c.diagnostics_raw(s, maxlag=5)
    b = s.diagnostics_raw(maxlag=5)
    assert P.relerr(a, b, 1e-9 * np.abs(b).max()).max() <= 1e-9
    c.close(); s.close()


def test_measured_fp64_peak_is_plausible(qs):
    dmma, dfma = qs.measure_fp64_peak("dmma", 0.1), qs.measure_fp64_peak("dfma", 0.1)
    assert 20.0 < dmma < 60.0 and 20.0 < dfma < 60.0, (dmma, dfma)


def test_library_is_built_from_this_tree(qs):
    from quicksand_b200 import build
    assert qs.source_hash() == build.source_hash()
