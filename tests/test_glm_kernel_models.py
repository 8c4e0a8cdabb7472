"""CPU checks of two design claims of the GLM gradient kernel (quicksand_b200/csrc/glm_hmc.cu), from the source itself:

* the table-driven sigmoid (256-entry 2^(j/256) table parsed out of the .cu file, degree-4 series, one cubic reciprocal step on
  a 24-bit seed) is accurate to ~1.5 ulp -- a numpy model of the device arithmetic (FMAs emulated in extended precision)
  against 200-bit arithmetic;
* the column-permuted X layout (logical dimension 8J + n in column 14 n + J, row stride 114, identity row map) makes the
  LDS.128 fragment loads of BOTH contractions conflict-free per quarter warp (8 lanes x 16 bytes must hit 8 distinct 16-byte
  bank groups), and no even stride with the old row map does."""
import itertools
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "quicksand_b200", "csrc", "glm_hmc.cu")).read()


def _table():
    body = SRC[SRC.index("c_exp2_tab[TAB_N] = {"):]
    body = body[:body.index("};")]
    vals = [float.fromhex(h) for h in re.findall(r"0x1\.[0-9a-f]+p[+-]\d+", body)]
    assert len(vals) == 256
    return np.array(vals)


def test_exp2_table_is_correctly_rounded():
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    tab = _table()
    for j in range(256):
        exact = mp.mpf(2) ** (mp.mpf(j) / 256)
        assert tab[j] == float(exact), j            # float(mpf) rounds to nearest


def test_table_driven_sigmoid_model_is_within_two_ulp():
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    tab = _table()
    # the constants the kernel uses, read from the source so that the model cannot drift from it
    c_scale = float(re.search(r"fma\(t, ([0-9.]+), SHIFT\)", SRC).group(1))
    c_hi = float.fromhex(re.search(r"fma\(nd, -(0x1\.[0-9a-f]+p-\d+), t\)", SRC).group(1))
    c_lo = float(re.search(r"fma\(nd, -([0-9.e-]+), r\)", SRC).group(1))
    assert abs(c_scale - 256 / math.log(2)) < 1e-12 and abs((c_hi + c_lo) - math.log(2) / 256) < 1e-18
    ld = np.longdouble

    def fma(a, b, c):
        return (ld(a) * ld(b) + ld(c)).astype(np.float64)
    rng = np.random.default_rng(0)
    z = np.concatenate([rng.normal(0, 3, 4000), rng.uniform(-40, 40, 4000), rng.uniform(-1e-3, 1e-3, 500), rng.uniform(-80, 80, 500)])
    t = np.clip(-z, -708.0, 708.0)
    n = np.rint(fma(t, c_scale, 0.0)).astype(np.int64)
    nd = n.astype(np.float64)
    r = fma(nd, -c_hi, t)
    r = fma(nd, -c_lo, r)
    assert np.abs(r).max() <= math.log(2) / 512 * (1 + 1e-9)
    q = np.full_like(r, 4.1666666666666664e-02)
    for c in (1.6666666666666666e-01, 0.5, 1.0, 1.0):
        q = fma(q, r, c)
    ts = np.ldexp(tab[n & 255], (n >> 8).astype(int))
    d = fma(q, ts, 1.0)
    y = (1.0 / d).astype(np.float32).astype(np.float64)          # ~24-bit seed (MUFU.RCP64H)
    err = fma(-d, y, 1.0)
    err = fma(err, err, err)
    y = fma(y, err, y)
    worst = 0.0
    for zi, pi in zip(z, y):
        exact = 1 / (1 + mp.exp(-mp.mpf(float(zi))))
        worst = max(worst, float(abs((mp.mpf(float(pi)) - exact) / exact)))
    assert worst < 4.5e-16, worst            # 2 ulp of 2.2e-16


def _ok128(addrs):
    return len({(a // 2) % 8 for a in addrs}) == 8


def _conflict_free(S, rho, PJ=14):
    for qw in range(4):
        lanes = [(g, t) for g in (2 * qw, 2 * qw + 1) for t in range(4)]
        for s in (0, 1):                 # first contraction: X[rho(gid) + 8T][(2 tig + s) PJ + J], J even
            if not _ok128([rho[g] * S + (2 * t + s) * PJ for g, t in lanes]):
                return False
        for e in (0, 1):                 # second contraction: X[rho(2 tig + e) + 8T][gid PJ + J], J even
            if not _ok128([rho[2 * t + e] * S + g * PJ for g, t in lanes]):
                return False
    return True


def test_permuted_layout_is_bank_conflict_free():
    assert re.search(r"PERM_PJ = 14, PERM_S = 114", SRC)
    assert re.search(r"RHO = PERM \? 0x76543210u", SRC)          # identity row map with the permuted layout
    assert _conflict_free(114, tuple(range(8)))
    old_rho = (0, 2, 1, 3, 6, 4, 7, 5)                           # the natural layout's row map does not carry over
    assert not any(_conflict_free(S, old_rho) for S in range(112, 132, 2))
    # the stride matters: 112 and 116 (the nearest multiples of 4) conflict for EVERY row permutation
    for S in (112, 116):
        assert not any(_conflict_free(S, rho) for rho in itertools.permutations(range(8)))
