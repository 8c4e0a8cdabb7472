"""CPU suite, part 3: the N > 1 host logic under torch.distributed (gloo, world_size 2): chains are block-partitioned
(rank r owns global ids [r*C, (r+1)*C)), each rank reduces its shard to plain-sum sufficient statistics, ONE
all-reduce(sum) combines them, and the diagnostics of the combined statistics equal those of the unsharded run.
(The per-shard statistics here are computed by a numpy restatement of diagnostics_kernel -- test code, not product.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quicksand_b200 import diagnostics as D


def raw_from_draws(draws, maxlag):
    """numpy restatement of diagnostics_kernel (csrc/capi.cu): draws [chains][N][retdim] -> raw statistics."""
    M, N, R = draws.shape
    h = N // 2
    out = np.zeros(R * D.HDR + R * (maxlag + 1))
    for j in range(R):
        x = draws[:, :, j]
        o = out[j * D.HDR:(j + 1) * D.HDR]
        o[0], o[1] = 2 * M, h
        for half in range(2):
            seg = x[:, half * h:(half + 1) * h]
            m = seg.mean(axis=1)
            o[2] += m.sum(); o[3] += (m * m).sum(); o[4] += seg.var(axis=1, ddof=1).sum()
        o[5], o[6] = x.sum(), (x * x).sum()
        o[7], o[8] = M, N
        m = x.mean(axis=1)
        o[9], o[10], o[11] = m.sum(), (m * m).sum(), x.var(axis=1, ddof=1).sum()
        dlt = x - m[:, None]
        ac = out[R * D.HDR + j * (maxlag + 1):R * D.HDR + (j + 1) * (maxlag + 1)]
        for t in range(maxlag + 1):
            ac[t] = (dlt[:, :N - t] * dlt[:, t:]).sum()
    return out


def make_draws(chain_begin, chains, N=200, R=3):
    """AR(1) draws whose content depends only on the GLOBAL chain id (like the Philox stream)."""
    out = np.empty((chains, N, R))
    for c in range(chains):
        rng = np.random.default_rng(1000 + chain_begin + c)
        e = rng.standard_normal((N, R))
        x = np.zeros(R)
        for i in range(N):
            x = np.array([0.5, 0.9, -0.3]) * x + e[i]
            out[c, i] = x
    return out


def _worker(rank, world, port, C, maxlag, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    draws = make_draws(rank * C, C)                      # this rank's shard: global chains [rank*C, (rank+1)*C)
    raw = torch.from_numpy(raw_from_draws(draws, maxlag))
    dist.all_reduce(raw)                                 # the only collective of the path
    if rank == 0:
        ret["raw"] = raw.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_sharded_diagnostics_equal_unsharded():
    world, C, maxlag, R = 2, 24, 30, 3
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), C, maxlag, ret), nprocs=world, join=True)
    sharded = D.summarize(ret["raw"], R, maxlag, nshards=world)
    full = D.summarize(raw_from_draws(make_draws(0, world * C), maxlag), R, maxlag)
    for k in ("mean", "var", "rhat", "ess", "tau"):
        assert np.allclose(sharded[k], full[k], rtol=1e-10, atol=1e-12), (k, sharded[k], full[k])
    assert np.all(sharded["chains"] == world * C) and np.all(sharded["draws"] == 200)
    # sanity of the estimator itself: AR(1) with rho has tau = (1+rho)/(1-rho)
    for j, rho in enumerate([0.5, 0.9, -0.3]):
        assert abs(full["tau"][j] / ((1 + rho) / (1 - rho)) - 1) < 0.25
    assert np.all(np.abs(full["rhat"] - 1) < 0.25)


def test_ess_matches_direct_formula():
    draws = make_draws(0, 16, N=300)
    raw = raw_from_draws(draws, 40)
    s = D.summarize(raw, 3, 40)
    for j in range(3):
        assert abs(s["ess"][j] / D.ess_numpy(draws[:, :, j], 40) - 1) < 1e-9
