"""Host-side index arithmetic of the CUDA library, exercised through the C ABI without a GPU (no oracle)."""
import ctypes as C
import random

import numpy as np


def _layout(numchains, nobs, nsm):
    from quicksand_b200 import _lib as L
    nt, nb, nc = C.c_int32(), C.c_int32(), C.c_int32()
    L.check(L.lib().qsb_debug_glm_layout(numchains, nobs, nsm, 0, None, None, None, C.byref(nt), C.byref(nb), C.byref(nc)))
    cap = max(nt.value, nc.value + 1)
    start, first, np_ = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    L.check(L.lib().qsb_debug_glm_layout(numchains, nobs, nsm, cap, start.ctypes.data_as(L.ip), first.ctypes.data_as(L.ip), np_.ctypes.data_as(L.ip),
                                         C.byref(nt), C.byref(nb), C.byref(nc)))
    return nt.value, nb.value, nc.value, start[:nc.value + 1].tolist(), first[:nt.value].tolist(), np_[:nt.value].tolist()


def _segments(nt, nb, nc, start):
    """What glm_grad_kernel does with the table: CTA c works off units [start[c], start[c+1]) tile by tile."""
    segs = {}
    for c in range(nc):
        su, end = start[c], start[c + 1]
        while su < end:
            tile = su // nb
            b0 = su - tile * nb
            b1 = min(nb, b0 + (end - su))
            segs.setdefault(tile, []).append((c, b0, b1))
            su += b1 - b0
    return segs


def test_balanced_layout_covers_every_tile_exactly_once():
    """The balanced decomposition of the logistic-regression gradient kernel (glm_hmc.cu: glm_layout): every chain tile's row
    blocks are covered once, by contiguous segments of consecutive CTAs, landing in partial-sum slots 0..np-1 (what the vector
    kernels sum); no CTA is empty; for the benchmark shapes and a few thousand random ones."""
    rnd = random.Random(1)
    shapes = [(c, n, 148) for c in (1, 7, 8, 96, 97, 200, 1024, 2048, 4096, 8192, 65536) for n in (1, 31, 32, 33, 333, 384, 10000, 10016, 32000)]
    shapes += [(rnd.randint(1, 200000), rnd.randint(1, 120000), rnd.choice([148, 132, 108, 16])) for _ in range(1500)]
    for chains, nobs, nsm in shapes:
        nt, nb, nc, start, first, np_ = _layout(chains, nobs, nsm)
        assert 1 <= nc <= nsm and start[0] == 0 and start[-1] == nt * nb, (chains, nobs, nsm)
        assert all(start[i + 1] > start[i] for i in range(nc)), (chains, nobs, nsm)
        segs = _segments(nt, nb, nc, start)
        for t in range(nt):
            s = segs[t]
            assert [c - first[t] for c, _, _ in s] == list(range(np_[t])), (chains, nobs, nsm, t)
            assert s[0][1] == 0 and s[-1][2] == nb and all(s[i][2] == s[i + 1][1] and s[i][2] > s[i][1] for i in range(len(s) - 1))


def test_balanced_layout_equalises_cost():
    """BASELINE config c3 (N = 10 000 -> 313 row blocks) at its strong-scaling shard sizes: one run per SM, and the costliest
    run is within 2 % of the mean (a unit of the partially filled last tile costs 78 % / 57 % of a full one with 2 / 1 live warps
    per SM sub-partition: the measured table of glm_unit_cost)."""
    for chains in (8192, 4096, 2048, 1024):
        nt, nb, nc, start, first, np_ = _layout(chains, 10000, 148)
        assert nc == 148 and nb == 313
        live_last = min(12, (chains - (nt - 1) * 96 + 7) // 8)
        w_last = {3: 100, 2: 78, 1: 57}[(live_last + 3) // 4]
        cost = []
        for c in range(nc):
            full_units = max(0, min(start[c + 1], (nt - 1) * nb) - start[c])
            last_units = (start[c + 1] - start[c]) - full_units
            cost.append(100 * full_units + w_last * last_units)
        assert max(cost) <= 1.02 * (sum(cost) / nc) + 100, (chains, max(cost), sum(cost) / nc)
        assert max(np_) <= 16
