"""ORACLE (test infrastructure, never shipped, never on the product path).

ctypes front end over ``oracle/libqsoracle.so``, the CPU restatement of the reference's
algorithm (qs/distrib.t, qs/erp.t, qs/trace.t, qs/lib/ad.t, qs/hmc.t, qs/drift.t,
qs/harm.t, qs/mcmc.t, qs/infer.t -- each C++ header cites the lines it follows).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``quicksand_b200/``
does: the product path has no CPU fallback.

Parity status: densities are pinned by the 26 known-answer values of
``test/testsuite.t:159-187``; kernels are pinned only statistically (expectation truths
of ``testsuite.t:621-691, 719-730, 776-787, 812-823``) because the reference cannot run
here (no Terra compiler) and holds no deterministic per-step fixture.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}
_ACTIVE = "strict"
_SO = {"strict": "libqsoracle.so", "fma": "libqsoracle_fma.so"}


def use(variant):
    """Select the oracle build new Programs are created in: "strict" (no FMA contraction, the oracle
    proper) or "fma" (contracted; a control for chaotic trajectories, see the Makefile)."""
    global _ACTIVE
    assert variant in _SO
    _ACTIVE = variant

HMC, DRIFT, HARM, TRACEMH = 1, 2, 3, 4
ERP_GAUSSIAN, ERP_GAMMA, ERP_BETA, ERP_UNIFORM, ERP_DIRICHLET, ERP_NEGGAMMA = 0, 1, 2, 3, 4, 5
ITER_INIT = 0xFFFFFFFF
ITER_SEARCH = 0xF0000000
SLOT_MIXTURE = 0xFFFFFFFF


class KernelSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("doAdapt", C.c_int32), ("stepSize", C.c_double),
                ("numSteps", C.c_uint64), ("scale", C.c_double), ("weight", C.c_double),
                ("lexicalScaleSharing", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("grad_evals", C.c_uint64), ("ad_nodes", C.c_uint64), ("rng_draws", C.c_uint64),
                ("props_made", C.c_uint64), ("props_accepted", C.c_uint64), ("seconds", C.c_double)]


def build(force=False):
    """Compile the restatement with g++ (Makefile in this directory)."""
    sos = [os.path.join(_HERE, f) for f in _SO.values()]
    if force or any(not os.path.exists(so) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(so)
            for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp"))) for so in sos):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return sos[0]


def lib(variant=None):
    variant = variant or _ACTIVE
    if variant not in _LIBS:
        so = os.path.join(_HERE, _SO[variant])
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        dp = C.POINTER(C.c_double)
        L.qso_logprob.restype = C.c_double
        L.qso_logprob.argtypes = [C.c_int, dp, C.c_int]
        L.qso_logprob_grad.restype = C.c_double
        L.qso_logprob_grad.argtypes = [C.c_int, dp, C.c_int, dp]
        L.qso_uniforms.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, dp]
        L.qso_gaussians.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_double, C.c_double, dp]
        L.qso_philox.argtypes = [C.POINTER(C.c_uint32)] * 3
        for f in ("qso_program_indep", "qso_program_gauss_prec", "qso_program_logistic",
                  "qso_program_eight_schools", "qso_program_funnel"):
            getattr(L, f).restype = C.c_void_p
        L.qso_program_indep.argtypes = [C.c_int, C.POINTER(C.c_int), dp, dp, C.c_int, C.c_double]
        L.qso_program_indep_lexids.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.qso_program_indep_factors.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.qso_program_indep_latent.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), dp, C.POINTER(C.c_int)]
        L.qso_program_indep_conditions.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.qso_program_gauss_prec.argtypes = [C.c_int, dp]
        L.qso_program_logistic.argtypes = [C.c_int, C.c_int, dp, C.POINTER(C.c_ubyte), C.c_double]
        L.qso_program_eight_schools.argtypes = [C.c_int, dp, dp]
        L.qso_program_funnel.argtypes = [C.c_int]
        L.qso_program_free.argtypes = [C.c_void_p]
        L.qso_program_retdim.argtypes = [C.c_void_p]
        L.qso_program_dim.argtypes = [C.c_void_p]
        ip = C.POINTER(C.c_int32)
        L.qso_run.argtypes = [C.c_void_p, C.POINTER(KernelSpec), C.c_int, C.c_uint64, C.c_uint32, C.c_uint32,
                              C.c_uint64, C.c_uint64, C.c_uint64, dp, dp, dp, ip, dp, dp, ip, dp, dp, C.POINTER(Stats)]
        L.qso_set_temperatures.argtypes = [dp, C.c_uint64]
        L.qso_set_traj_record.argtypes = [dp] * 7
        L.qso_logp_grad.argtypes = [C.c_void_p, dp, C.c_int, dp, dp, dp]
        L.qso_expectation.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, dp]
        L.qso_autocorrelation.argtypes = [dp, C.c_int, C.c_int, dp]
        _LIBS[variant] = L
    return _LIBS[variant]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def logprob(which, args):
    a = np.ascontiguousarray(args, dtype=np.float64)
    return lib().qso_logprob(which, _dp(a), len(a))


def logprob_grad(which, args, wrt):
    a = np.ascontiguousarray(args, dtype=np.float64)
    g = C.c_double()
    lp = lib().qso_logprob_grad(which, _dp(a), wrt, C.byref(g))
    return lp, g.value


def uniforms(seed, chain, it, slot, n):
    out = np.empty(n)
    lib().qso_uniforms(seed, chain, it, slot, n, _dp(out))
    return out


def gaussians(seed, chain, it, slot0, n, mu=0.0, sigma=1.0):
    out = np.empty(n)
    lib().qso_gaussians(seed, chain, it, slot0, n, mu, sigma, _dp(out))
    return out


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().qso_philox(c, k, o)
    return list(o)


class Program:
    """A fixed-structure program of one of the five families (oracle/programs.hpp)."""

    def __init__(self, handle, keep=()):
        self.h = C.c_void_p(handle)
        self._keep = keep
        self._lib = lib()          # the build this program's code lives in
        self.retdim = self._lib.qso_program_retdim(self.h)
        self.dim = self._lib.qso_program_dim(self.h)

    def __del__(self):
        try:
            self._lib.qso_program_free(self.h)
        except Exception:
            pass

    @staticmethod
    def indep(kinds, p0, p1, ret_mode=0, ret_div=1.0, lexids=None, fmu=None, fsigma=None, pref=None, pcoef=None, ptf=None, cond_lo=None, cond_hi=None):
        k = np.ascontiguousarray(kinds, dtype=np.int32)
        a = np.ascontiguousarray(p0, dtype=np.float64)
        b = np.ascontiguousarray(p1, dtype=np.float64)
        h = lib().qso_program_indep(len(k), k.ctypes.data_as(C.POINTER(C.c_int)), _dp(a), _dp(b), ret_mode, ret_div)
        if lexids is not None:
            lx = np.ascontiguousarray(lexids, dtype=np.int32)
            lib().qso_program_indep_lexids(C.c_void_p(h), len(lx), lx.ctypes.data_as(C.POINTER(C.c_int)))
        if fsigma is not None:
            fm = np.ascontiguousarray(fmu, dtype=np.float64); fs = np.ascontiguousarray(fsigma, dtype=np.float64)
            lib().qso_program_indep_factors(C.c_void_p(h), len(fs), _dp(fm), _dp(fs))
        if cond_lo is not None:
            lo = np.ascontiguousarray(cond_lo, dtype=np.float64); hi = np.ascontiguousarray(cond_hi, dtype=np.float64)
            lib().qso_program_indep_conditions(C.c_void_p(h), len(lo), _dp(lo), _dp(hi))
        if pref is not None:
            r = np.ascontiguousarray(pref, dtype=np.int32).ravel(); cb = np.ascontiguousarray(pcoef, dtype=np.float64).ravel()
            tf = np.ascontiguousarray(ptf if ptf is not None else np.zeros(len(r)), dtype=np.int32).ravel()
            ip = C.POINTER(C.c_int)
            lib().qso_program_indep_latent(C.c_void_p(h), len(k), r.ctypes.data_as(ip), _dp(cb), tf.ctypes.data_as(ip))
        return Program(h)

    @staticmethod
    def gauss_prec(A):
        A = np.ascontiguousarray(A, dtype=np.float64)
        return Program(lib().qso_program_gauss_prec(A.shape[0], _dp(A)))

    @staticmethod
    def logistic(X, y, prior_sigma=2.5):
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.uint8)
        return Program(lib().qso_program_logistic(X.shape[0], X.shape[1], _dp(X), y.ctypes.data_as(C.POINTER(C.c_ubyte)), prior_sigma))

    @staticmethod
    def eight_schools(y, sigma):
        y = np.ascontiguousarray(y, dtype=np.float64)
        s = np.ascontiguousarray(sigma, dtype=np.float64)
        return Program(lib().qso_program_eight_schools(len(y), _dp(y), _dp(s)))

    @staticmethod
    def funnel(d):
        return Program(lib().qso_program_funnel(d))

    def logp_grad(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.empty_like(x)
        lpd, lpf = C.c_double(), C.c_double()
        rc = self._lib.qso_logp_grad(self.h, _dp(x), len(x), C.byref(lpd), _dp(g), C.byref(lpf))
        if rc:
            raise ValueError("dimension mismatch")
        return lpd.value, g, lpf.value


def spec(kind, stepSize=1.0, numSteps=1, doAdapt=True, scale=1.0, weight=1.0, lexicalScaleSharing=False):
    return KernelSpec(kind, int(bool(doAdapt)), stepSize, numSteps, scale, weight, int(bool(lexicalScaleSharing)), 0)


def run(program, specs, seed, nchains, nsamps, burnin=0, lag=1, chain_begin=0, record=False, temps=None, traj=False):
    """doMCMC for each chain.  Returns a dict with samples [C][S][retdim], logprobs and, if
    ``record``, the per-iteration trace (state, accept, dH, step, kernel index, leapfrog positions).
    ``temps``: AnnealingKernel schedule, temps[i] = annealSched(i, iters) (qs/mcmc.t:297-319).
    ``traj`` (single HMC kernel, with ``record``): additionally the state entering every trajectory (x0, g0, p0, lp0)
    and the momenta / gradient / log-density after every leapfrog step -- what a teacher-forced replay needs."""
    if isinstance(specs, KernelSpec):
        specs = [specs]
    arr = (KernelSpec * len(specs))(*specs)
    iters = burnin + nsamps * lag
    n = program.dim
    samples = np.empty((nchains, nsamps, program.retdim))
    logprobs = np.empty((nchains, nsamps))
    out = {"samples": samples, "logprobs": logprobs}
    st = ac = dh = ss = ki = lf = init = None
    if record:
        st = np.zeros((nchains, iters, n)); ac = np.zeros((nchains, iters), dtype=np.int32)
        dh = np.zeros((nchains, iters)); ss = np.zeros((nchains, iters)); init = np.zeros((nchains, n))
        if len(specs) > 1:
            ki = np.zeros((nchains, iters), dtype=np.int32)
        elif specs[0].kind == HMC:
            lf = np.zeros((nchains, iters, specs[0].numSteps, n))
        out.update(state=st, accept=ac, dH=dh, step=ss, kidx=ki, leapfrog=lf, init=init)
    tr = None
    if traj:
        assert record and len(specs) == 1 and specs[0].kind == HMC
        Ls = int(specs[0].numSteps)
        tr = dict(x0=np.zeros((nchains, iters, n)), g0=np.zeros((nchains, iters, n)), p0=np.zeros((nchains, iters, n)),
                  lp0=np.zeros((nchains, iters)), leap_mom=np.zeros((nchains, iters, Ls, n)), leap_grad=np.zeros((nchains, iters, Ls, n)),
                  leap_lp=np.zeros((nchains, iters, Ls)))
        program._lib.qso_set_traj_record(_dp(tr["x0"]), _dp(tr["g0"]), _dp(tr["p0"]), _dp(tr["lp0"]), _dp(tr["leap_mom"]),
                                         _dp(tr["leap_grad"]), _dp(tr["leap_lp"]))
        out.update(tr)
    stats = Stats()
    tarr = np.ascontiguousarray(temps, dtype=np.float64) if temps is not None else None
    program._lib.qso_set_temperatures(_dp(tarr), 0 if tarr is None else len(tarr))
    rc = program._lib.qso_run(program.h, arr, len(specs), seed, chain_begin, nchains, nsamps, burnin, lag,
                       _dp(samples), _dp(logprobs), _dp(st), _ip(ac), _dp(dh), _dp(ss), _ip(ki), _dp(lf), _dp(init),
                       C.byref(stats))
    program._lib.qso_set_temperatures(None, 0)
    if traj:
        program._lib.qso_set_traj_record(None, None, None, None, None, None, None)
    if rc:
        raise RuntimeError("qso_run failed: %d" % rc)
    out["stats"] = {f: getattr(stats, f) for f, _ in Stats._fields_}
    return out


def expectation(values, do_variance=False):
    v = np.ascontiguousarray(values, dtype=np.float64)
    if v.ndim == 1:
        v = v[:, None]
    m = np.empty(v.shape[1]); var = C.c_double()
    lib().qso_expectation(_dp(v), v.shape[0], v.shape[1], int(do_variance), _dp(m), C.byref(var))
    return (m, var.value) if do_variance else m


def autocorrelation(values):
    v = np.ascontiguousarray(values, dtype=np.float64)
    if v.ndim == 1:
        v = v[:, None]
    ac = np.empty(v.shape[0])
    lib().qso_autocorrelation(_dp(v), v.shape[0], v.shape[1], _dp(ac))
    return ac
