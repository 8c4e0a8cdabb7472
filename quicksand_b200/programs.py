"""Fixed-structure continuous programs: the model-family descriptors that stand for a
``qs.program`` on the device.  Each function mirrors how a quicksand user
writes the program (ERP calls with {struc=false}, qs.factor, erp.observe) and returns a
:class:`Program` holding the POD descriptor the C ABI reads."""
import ctypes as C

import numpy as np

from . import _lib as L


def _dp(a):
    return a.ctypes.data_as(L.dp)


class Program:
    """Host-side stand-in for a compiled ``qs.program`` (qs/progmodule.t:58-70)."""

    def __init__(self, family, dim, **arrays):
        self.family, self.dim = family, dim
        self.nobs = arrays.pop("nobs", 0)
        self.ret_mode = arrays.pop("ret_mode", 1)
        self.ret_div = arrays.pop("ret_div", 1.0)
        self.prior_sigma = arrays.pop("prior_sigma", 0.0)
        self.arrays = arrays          # keeps the numpy buffers alive
        self._models = {}

    @property
    def retdim(self):
        return 1 if (self.family == L.FAM_INDEP and self.ret_mode == 0) else self.dim

    def desc(self):
        d = L.ModelDesc()
        d.family, d.dim, d.nobs, d.ret_mode = self.family, self.dim, self.nobs, self.ret_mode
        d.ret_div, d.prior_sigma = self.ret_div, self.prior_sigma
        a = self.arrays
        if "erp_kind" in a:
            d.erp_kind = a["erp_kind"].ctypes.data_as(L.ip); d.erp_p0 = _dp(a["erp_p0"]); d.erp_p1 = _dp(a["erp_p1"])
        if "A" in a:
            d.A = _dp(a["A"])
        if "X" in a:
            d.X = _dp(a["X"]); d.ybin = a["ybin"].ctypes.data_as(C.POINTER(C.c_uint8))
        if "y" in a:
            d.y = _dp(a["y"]); d.sigma = _dp(a["sigma"])
        if "erp_lexid" in a:
            d.erp_lexid = a["erp_lexid"].ctypes.data_as(L.ip)
        if "erp_fsigma" in a:
            d.erp_fmu = _dp(a["erp_fmu"]); d.erp_fsigma = _dp(a["erp_fsigma"])
        if "erp_cond_lo" in a:
            d.erp_cond_lo = _dp(a["erp_cond_lo"]); d.erp_cond_hi = _dp(a["erp_cond_hi"])
        if "erp_pref" in a:
            d.erp_pref = a["erp_pref"].ctypes.data_as(L.ip); d.erp_pcoef = _dp(a["erp_pcoef"]); d.erp_ptf = a["erp_ptf"].ctypes.data_as(L.ip)
        return d

    def model(self, device=-1):
        """Device-resident model handle (uploads the model data once per device)."""
        if device not in self._models:
            h = C.c_void_p()
            L.check(L.lib().qsb_model_create(C.byref(self.desc()), device, C.byref(h)))
            self._models[device] = h
        return self._models[device]

    def update_data(self, device=-1, stream=None, **arrays):
        """Refresh the observation data of the device-resident model from host arrays (same shapes)."""
        for k, v in arrays.items():
            assert k in self.arrays and v.shape == self.arrays[k].shape and v.dtype == self.arrays[k].dtype
            self.arrays[k] = v
        L.check(L.lib().qsb_model_update_data(self.model(device), C.byref(self.desc()), C.c_void_p(stream) if stream else None))

    def upload(self, device=-1):
        """Force a fresh host->device upload of the model data (used by the end-to-end benchmark)."""
        self.release(device)
        return self.model(device)

    def release(self, device=None):
        for dev in list(self._models) if device is None else [device]:
            h = self._models.pop(dev, None)
            if h is not None:
                L.lib().qsb_model_destroy(h)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def logp_grad(self, x, device=-1):
        """Test hook: (lp_dual, grad, lp_float) at unconstrained points x [n][dim]."""
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        n = x.shape[0]
        lpd, lpf, g = np.empty(n), np.empty(n), np.empty_like(x)
        L.check(L.lib().qsb_logp_grad(self.model(device), _dp(x), n, _dp(lpd), _dp(g), _dp(lpf)))
        return lpd, g, lpf


    def debug_leapfrog(self, x, p, eps, g=None, last=False, device=-1):
        """Test hook (qsb_debug_leapfrog): one HMCKernel:step of n injected states.  Returns dict(x, p, g, lp)."""
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
        n = x.shape[0]
        eps = np.ascontiguousarray(np.broadcast_to(np.asarray(eps, dtype=np.float64), (n,)))
        gp = None
        if g is not None:
            g = np.ascontiguousarray(np.atleast_2d(g), dtype=np.float64)
            gp = _dp(g)
        xo, po, go, lpo = np.empty_like(x), np.empty_like(x), np.empty_like(x), np.empty(n)
        L.check(L.lib().qsb_debug_leapfrog(self.model(device), n, _dp(x), _dp(p), gp, _dp(eps), int(bool(last)),
                                           _dp(xo), _dp(po), _dp(go), _dp(lpo)))
        return dict(x=xo, p=po, g=go, lp=lpo)


_KIND = {"gaussian": L.ERP_GAUSSIAN, "gamma": L.ERP_GAMMA, "beta": L.ERP_BETA, "uniform": L.ERP_UNIFORM, "neggamma": L.ERP_NEGGAMMA}


class latent:
    """A parameter of an ``indep`` entry that depends on the value of an earlier scalar entry: ``f(a + b * value[entry])``
    with f = identity or exp -- e.g. ``[("gaussian", 0, 1), ("gaussian", latent(0), 1.0)]`` for
    ``mu = qs.gaussian(0,1); x = qs.gaussian(mu, 1)`` and ``latent(1, b=0.5, exp=True)`` for ``exp(v/2)``."""

    def __init__(self, entry, a=0.0, b=1.0, exp=False):
        self.entry, self.a, self.b, self.exp = int(entry), float(a), float(b), bool(exp)


def indep(erps, ret="sum", ret_div=1.0, lexids=None, factors=None, conditions=None):
    """ERPs in program order with constant parameters (or ``latent`` ones, see above), e.g. ``[("gamma", 2.0, 2.0)]`` with ret_div=10 for
    ``return qs.gamma(2.0, 2.0, {struc=false})/10.0`` (test/testsuite.t:657-673).  Entries:

      ("gaussian", mu, sigma) ("gamma", shape, scale) ("beta", a, b) ("uniform", lo, hi)     erpdefs.t:28-36, 46-69, 92-100
      ("gammamv", m, v)  -> gamma(m*m/v, v/m)                                                  erpdefs.t:72-80
      ("betamv", m, v)   -> beta(-m*tmp/v, (m-1)*tmp/v), tmp = m*m - m + v (must be < 0)       erpdefs.t:103-117
      ("dirichlet", [alpha_0 .. alpha_{K-1}])  one vector-valued choice of K real components   erpdefs.t:179-212

    ``factors[i]``: ``None`` or ``(mu, sigma)``: the statement ``qs.factor(gaussian.logprob(value_i, mu, sigma))`` (equally
    ``qs.gaussian.observe(value_i, mu, sigma)``) right after scalar choice i (test/testsuite.t:394-442).

    ``conditions[i]``: ``None`` or ``(lo, hi)``: the statement ``qs.condition(value_i > lo and value_i < hi)`` right after scalar
    choice i (``None`` / +-inf on a side = no bound): hard constraint, rejection initialisation, accept tests that require it
    (trace.t:612-623, 794-796).  ("neggamma", shape, scale): value = -gamma(shape, scale) under Bounds.Upper(0) (erp.t:158-173).

    ``lexids[i]``: call-site id of entry i (entries produced by one ERP call inside a loop share an id); only
    DriftKernel{lexicalScaleSharing=true} looks at it (drift.t:208-235).  Default: every entry is its own call site."""
    kinds, p0, p1, lex, fmu, fsig, clo, chi = [], [], [], [], [], [], [], []
    pref, pcoef, ptf, comp_of_entry = [], [], [], []
    for n_e, e in enumerate(erps):
        k = e[0]
        comp_of_entry.append(len(kinds))
        site = n_e if lexids is None else int(lexids[n_e])
        fac = None if factors is None else factors[n_e]
        if k == "dirichlet":
            assert fac is None, "factors on a dirichlet choice are not supported"
            alphas = [float(a) for a in e[1]]
            assert conditions is None or conditions[n_e] is None, "conditions on a dirichlet choice are not supported"
            for j, a in enumerate(alphas):
                kinds.append(L.ERP_DIRICHLET); p0.append(a); p1.append(float(j)); lex.append(site); fmu.append(0.0); fsig.append(0.0)
                clo.append(-np.inf); chi.append(np.inf)
                pref += [-1, -1]; pcoef += [0.0, 0.0]; ptf += [0, 0]
            continue
        lex.append(site)
        fmu.append(0.0 if fac is None else float(fac[0])); fsig.append(0.0 if fac is None else float(fac[1]))
        cnd = None if conditions is None else conditions[n_e]
        clo.append(-np.inf if cnd is None or cnd[0] is None else float(cnd[0])); chi.append(np.inf if cnd is None or cnd[1] is None else float(cnd[1]))
        for q in (1, 2):
            if isinstance(e[q], latent):
                assert k in ("gaussian", "gamma", "beta", "neggamma"), "latent parameters: gaussian, gamma, beta and neggamma choices"
                assert 0 <= e[q].entry < n_e and erps[e[q].entry][0] != "dirichlet", "a latent parameter refers to an earlier scalar entry"
                pref.append(comp_of_entry[e[q].entry]); pcoef.append(e[q].b); ptf.append(1 if e[q].exp else 0)
            else:
                pref.append(-1); pcoef.append(0.0); ptf.append(0)
        a, b = (x.a if isinstance(x, latent) else float(x) for x in (e[1], e[2]))
        if k == "gammamv":
            k, a, b = "gamma", a * a / b, b / a
        elif k == "betamv":
            tmp = a * a - a + b
            assert tmp < 0.0, "betamv: v must be less than m - m^2"          # S.assert(tmp < 0.0) (erpdefs.t:110)
            k, a, b = "beta", -a * tmp / b, (a - 1) * tmp / b
        kinds.append(_KIND[k] if isinstance(k, str) else int(k)); p0.append(a); p1.append(b)
    extra = {} if lexids is None else {"erp_lexid": np.array(lex, dtype=np.int32)}
    if factors is not None:
        extra.update(erp_fmu=np.array(fmu, dtype=np.float64), erp_fsigma=np.array(fsig, dtype=np.float64))
    if conditions is not None:
        extra.update(erp_cond_lo=np.array(clo, dtype=np.float64), erp_cond_hi=np.array(chi, dtype=np.float64))
    if any(r >= 0 for r in pref):
        extra.update(erp_pref=np.array(pref, dtype=np.int32), erp_pcoef=np.array(pcoef, dtype=np.float64), erp_ptf=np.array(ptf, dtype=np.int32))
    return Program(L.FAM_INDEP, len(kinds), erp_kind=np.array(kinds, dtype=np.int32), erp_p0=np.array(p0, dtype=np.float64),
                   erp_p1=np.array(p1, dtype=np.float64), ret_mode=0 if ret == "sum" else 1, ret_div=float(ret_div), **extra)


def gauss_prec(A):
    """x_i ~ qs.gaussian(0,1,{struc=false}), i<d; qs.factor(-0.5 x'Ax).  log p = -0.5 x'(I+A)x + c."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    return Program(L.FAM_GAUSS_PREC, A.shape[0], A=A)


def logistic(X, y, prior_sigma=2.5):
    """theta_j ~ qs.gaussian(0, prior_sigma); qs.flip.observe(y_i, 1/(1+exp(-x_i . theta)))."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.uint8)
    return Program(L.FAM_LOGISTIC, X.shape[1], X=X, ybin=y, nobs=X.shape[0], prior_sigma=float(prior_sigma))


def eight_schools(y, sigma):
    """mu ~ gaussian(0,5); logtau ~ gaussian(0,1); theta_j ~ gaussian(mu, exp(logtau));
    qs.gaussian.observe(y_j, theta_j, sigma_j).  Components [mu, logtau, theta_0..theta_{J-1}]."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    s = np.ascontiguousarray(sigma, dtype=np.float64)
    return Program(L.FAM_EIGHT_SCHOOLS, len(y) + 2, y=y, sigma=s, nobs=len(y))


def funnel(d):
    """v ~ gaussian(0,3); x_i ~ gaussian(0, exp(v/2)), i<d-1.  Components [v, x_0..x_{d-2}]."""
    return Program(L.FAM_FUNNEL, int(d))


# ---- synthetic data of BASELINE.json's configs; numpy Philox streams, fixed seeds
def _rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def c2_correlated_gaussian(d=10, rho=0.9):
    i = np.arange(d)
    Sigma = rho ** np.abs(i[:, None] - i[None, :])
    P = np.linalg.inv(Sigma)
    return gauss_prec(P - np.eye(d)), Sigma


def c3_logistic_data(N=10000, d=100):
    X = _rng(1).standard_normal((N, d)) / np.sqrt(d)
    theta = _rng(2).standard_normal(d)
    p = 1.0 / (1.0 + np.exp(-(X @ theta)))
    y = (_rng(3).random(N) < p).astype(np.uint8)
    return X, y, theta


def c4_eight_schools_data(J=1000):
    r = _rng(4)
    sigma = r.uniform(8.0, 18.0, J)
    theta = 4.0 + 3.0 * r.standard_normal(J)
    y = theta + sigma * r.standard_normal(J)
    return y, sigma
