"""Cross-chain convergence diagnostics from the plain-sum sufficient statistics of
``qsb_sampler_diagnostics`` (include/qsb200.h): pooled moments, split-R-hat, and ESS with Geyer's
initial-monotone-sequence truncation on the chain-averaged autocovariance (Stan-style, not
rank-normalised).  The reference has none of these (its only related code is Autocorrelation,
qs/infer.t:213-264); they are the new statistics north_star asks for.  Because the inputs are sums
over chains, shards on different GPUs combine with ONE all-reduce(sum) -- the only collective of
the whole path."""
import numpy as np

HDR = 12


def unpack(raw, retdim, maxlag, nshards=1):
    raw = np.asarray(raw, dtype=np.float64)
    hdr = raw[:retdim * HDR].reshape(retdim, HDR).copy()
    hdr[:, 1] /= nshards      # draws per half-chain and per chain are constants, not sums
    hdr[:, 8] /= nshards
    acov = raw[retdim * HDR:].reshape(retdim, maxlag + 1) if maxlag > 0 else None
    return hdr, acov


def summarize(raw, retdim, maxlag=0, nshards=1):
    """dict(mean, var, rhat, ess, tau) per return-value dimension; `raw` may be the all-reduced sum of `nshards` shards."""
    hdr, acov = unpack(raw, retdim, maxlag, nshards)
    mh, nh = hdr[:, 0], hdr[:, 1]
    M, N = hdr[:, 7], hdr[:, 8]
    tot = M * N
    mean = hdr[:, 5] / tot
    var = hdr[:, 6] / tot - mean * mean
    out = dict(mean=mean, var=var, chains=M, draws=N)
    with np.errstate(divide="ignore", invalid="ignore"):
        # split-R-hat over the 2M half-chains
        Bn = (hdr[:, 3] - hdr[:, 2] ** 2 / mh) / (mh - 1.0)       # variance of the half-chain means = B/n
        W = hdr[:, 4] / mh
        varplus = (nh - 1.0) / nh * W + Bn
        out["rhat"] = np.sqrt(varplus / W)
        if acov is not None:
            Bn_f = (hdr[:, 10] - hdr[:, 9] ** 2 / M) / np.maximum(M - 1.0, 1.0)
            W_f = hdr[:, 11] / M
            vp = (N - 1.0) / N * W_f + Bn_f
            acov_mean = acov / (M[:, None] * N[:, None])          # chain-averaged, biased (1/N) autocovariance
            rho = 1.0 - (W_f[:, None] - acov_mean) / vp[:, None]
            rho[:, 0] = 1.0
            tau = np.empty(retdim)
            hit = np.zeros(retdim, dtype=bool)
            for j in range(retdim):
                tau[j], hit[j] = _geyer_tau(rho[j], with_flag=True)
            out["tau"] = tau
            out["truncated_at_maxlag"] = hit     # the pair sums were still positive at maxlag: tau (and the ESS) is a bound, not an estimate
            out["ess"] = tot / tau
    return out


def summarize_moments(raw, retdim, nshards=1):
    """The same summary from the QSB_FLAG_MOMENTS layout (15 doubles per dimension: the 12 sums + [sum over chains of the
    batch-mean variance, batch size, batches per chain]); the ESS is the batch-means estimate
    M N s^2 / (b * mean batch-mean variance), s^2 the mean within-chain variance."""
    raw = np.asarray(raw, dtype=np.float64)
    out = summarize(raw[:retdim * HDR], retdim, 0, nshards)
    bm = raw[retdim * HDR:].reshape(retdim, 3).copy()
    bm[:, 1] /= nshards; bm[:, 2] /= nshards
    hdr, _ = unpack(raw[:retdim * HDR], retdim, 0, nshards)
    M, N = hdr[:, 7], hdr[:, 8]
    with np.errstate(divide="ignore", invalid="ignore"):
        W = hdr[:, 11] / M
        sigma2 = bm[:, 1] * bm[:, 0] / M              # asymptotic variance of the chain mean times N
        tau = np.where(bm[:, 2] >= 2, sigma2 / W, np.nan)
        out["tau"] = tau
        out["ess"] = M * N / np.maximum(tau, 1e-3)
    out["batch"] = bm[:, 1]; out["batches"] = bm[:, 2]
    # batch means need batches much longer than the autocorrelation time; when the estimate itself says they are not
    # (tau > batch / 4) it is biased low and the ESS is an upper bound, not an estimate (the counterpart of Geyer's
    # `truncated_at_maxlag`)
    with np.errstate(invalid="ignore"):
        out["batch_too_short"] = ~(tau <= bm[:, 1] / 4.0)
    return out


def _geyer_tau(rho, with_flag=False):
    """tau = -1 + 2 * sum of the initial positive, monotone sequence of pair sums P_k = rho_2k + rho_2k+1.
    with_flag: also return whether the sequence ran into the end of rho (maxlag) before a non-positive pair sum."""
    T = len(rho)
    s = 0.0
    prev = np.inf
    k = 0
    stopped = False
    while 2 * k + 1 < T:
        P = rho[2 * k] + rho[2 * k + 1]
        if not (P > 0.0):
            stopped = True
            break
        P = min(P, prev)
        s += P
        prev = P
        k += 1
    tau = -1.0 + 2.0 * s
    tau = tau if tau > 1e-3 else 1e-3
    return (tau, not stopped) if with_flag else tau


def ess_numpy(draws, maxlag):
    """Reference implementation on raw draws [chains][N] for one dimension (used by the tests)."""
    M, N = draws.shape
    means = draws.mean(axis=1)
    W = draws.var(axis=1, ddof=1).mean()
    Bn = means.var(ddof=1) if M > 1 else 0.0
    vp = (N - 1.0) / N * W + Bn
    d = draws - means[:, None]
    rho = np.ones(maxlag + 1)
    for t in range(1, maxlag + 1):
        ac = (d[:, :N - t] * d[:, t:]).sum(axis=1).mean() / N
        rho[t] = 1.0 - (W - ac) / vp
    return M * N / _geyer_tau(rho)
