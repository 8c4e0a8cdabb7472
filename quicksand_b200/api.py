"""Host-side mirror of the reference's inference interface for the hot path:

    qs.infer(program, query, method)                               qs/infer.t:59-70
    qs.MCMC(kernel, {numsamps, burnin, lag, verbose})              qs/mcmc.t:36-105
    qs.HMCKernel{stepSize, numSteps, doStepSizeAdapt}              qs/hmc.t:57-66
    qs.DriftKernel{scale, doScaleAdapt, lexicalScaleSharing}       qs/drift.t:63-73
    qs.HARMKernel{scale, doScaleAdapt}                             qs/harm.t:16-22
    qs.TraceMHKernel{doStruct=false}                               qs/mcmc.t:134-193 (the non-structural half)
    qs.MixtureKernel(kernels, weights)                             qs/mcmc.t:199-291
    qs.AnnealingKernel(kernel, annealSched)                        qs/mcmc.t:297-319
    qs.Samples / qs.Expectation / qs.Autocorrelation               qs/infer.t:128-136, 146-187, 213-264

Same names, same parameter meaning, same defaults, same error behaviour (the reference aborts with a
message; Python raises with the same message).  New optional MCMC params: ``numchains`` (default 1 =
the reference's single chain, identical sample shape), ``seed``, ``device``, ``chain_begin``.
Everything numeric happens in libqsb200.so (CUDA); this module only marshals descriptors.
"""
import ctypes as C
import sys
import time

import numpy as np

from . import _lib as L
from .programs import Program


# ---------------------------------------------------------------------------------------------
class KernelSpecs:
    """What a kernel constructor yields: the reference returns ``function(TraceType) -> struct``;
    here the spec list the B200 driver reads."""

    def __init__(self, specs, names, anneal=None):
        self.specs, self.names = specs, names
        self.anneal = anneal      # annealSched(iter, numiters) -> temperature, or None


def HMCKernel(params=None, **kw):
    p = dict(params or {}); p.update(kw)
    unknown = set(p) - {"stepSize", "numSteps", "doStepSizeAdapt"}
    if unknown:
        raise TypeError("HMCKernel: unknown params %s" % sorted(unknown))
    adapt = p.get("doStepSizeAdapt", True)                     # hmc.t:61-62
    s = L.KernelSpec(L.KERNEL_HMC, int(bool(adapt)), float(p.get("stepSize", 1.0)), int(p.get("numSteps", 1)), 1.0, 1.0, 0, 0)
    return KernelSpecs([s], ["HMCKernel"])


def DriftKernel(params=None, **kw):
    p = dict(params or {}); p.update(kw)
    unknown = set(p) - {"scale", "doScaleAdapt", "lexicalScaleSharing"}
    if unknown:
        raise TypeError("DriftKernel: unknown params %s" % sorted(unknown))
    adapt = p.get("doScaleAdapt", True)                        # drift.t:66-67
    lexical = p.get("lexicalScaleSharing", False)              # drift.t:71-73
    s = L.KernelSpec(L.KERNEL_DRIFT, int(bool(adapt)), 1.0, 1, float(p.get("scale", 1.0)), 1.0, int(bool(lexical)), 0)
    return KernelSpecs([s], ["DriftKernel"])


def HARMKernel(params=None, **kw):
    p = dict(params or {}); p.update(kw)
    unknown = set(p) - {"scale", "doScaleAdapt"}
    if unknown:
        raise TypeError("HARMKernel: unknown params %s" % sorted(unknown))
    adapt = p.get("doScaleAdapt", True)                        # harm.t:19-20
    s = L.KernelSpec(L.KERNEL_HARM, int(bool(adapt)), 1.0, 1, float(p.get("scale", 1.0)), 1.0, 0, 0)
    return KernelSpecs([s], ["HARMKernel"])


def TraceMHKernel(params=None, **kw):
    """qs/mcmc.t:134-193.  Fixed-structure programs have no structural choices, so only ``doStruct=false`` (re-propose one
    random non-structural choice per iteration) exists on the device; the default ``doStruct=true`` is refused."""
    p = dict(params or {}); p.update(kw)
    unknown = set(p) - {"doStruct", "doNonstruct"}
    if unknown:
        raise TypeError("TraceMHKernel: unknown params %s" % sorted(unknown))
    doStruct, doNonstruct = p.get("doStruct", True), p.get("doNonstruct", True)
    assert doStruct or doNonstruct, "TraceMHKernel: cannot set both 'doStruct' and 'doNonstruct' to false"   # mcmc.t:139
    if doStruct:
        raise NotImplementedError("TraceMHKernel: structural choices are out of scope of the device path; pass doStruct=False")
    s = L.KernelSpec(L.KERNEL_TRACEMH, 0, 1.0, 1, 1.0, 1.0, 0, 0)
    return KernelSpecs([s], ["TraceMHKernel"])


def MixtureKernel(kernels, weights):
    assert len(kernels) == len(weights), "MixtureKernel: number of kernels must equal number of weights"   # mcmc.t:200-201
    specs, names = [], []
    for k, w in zip(kernels, weights):
        if len(k.specs) != 1:
            raise NotImplementedError("MixtureKernel: nested mixtures are not supported")
        if k.anneal is not None:
            raise NotImplementedError("MixtureKernel: anneal the mixture (AnnealingKernel(MixtureKernel(...))), not its parts")
        s = k.specs[0]
        specs.append(L.KernelSpec(s.kind, s.doAdapt, s.stepSize, s.numSteps, s.scale, float(w), s.lexicalScaleSharing, 0))
        names.append(k.names[0])
    return KernelSpecs(specs, names)


def AnnealingKernel(kernel, annealSched):
    """Simulated annealing on top of another kernel (qs/mcmc.t:297-319): before every iteration the trace
    temperature is set to ``annealSched(iter, numiters)`` (any callable of two ints returning a float).  The
    schedule is host code, as in the reference (a Terra function); it is tabulated once per run and handed to
    the device as an array (``qsb_sampler_set_temperatures``)."""
    if kernel.anneal is not None:
        raise NotImplementedError("AnnealingKernel: nested annealing schedules are not supported")
    return KernelSpecs(kernel.specs, kernel.names, anneal=annealSched)


# ---------------------------------------------------------------------------------------------
class Sampler:
    """One ``qsb_sampler``: the SoA state of ``numchains`` chains of one program under one kernel,
    resident in HBM.  Thin wrapper over the C ABI; used by MCMC(), the tests and bench.py."""

    def __init__(self, program, kernel, numchains=1, seed=0, device=-1, chain_begin=0, sample_chains=0,
                 flags=0, stream=None, mapping=0):
        assert isinstance(program, Program), "MCMC: expected a qs.program"      # progmod.assertIsProgram (mcmc.t:46)
        self.program, self.kernel = program, kernel
        self.numchains = int(numchains)
        self.sample_chains = int(sample_chains) if sample_chains else self.numchains
        self.retdim, self.dim = program.retdim, program.dim
        arr = (L.KernelSpec * len(kernel.specs))(*kernel.specs)
        cfg = L.RunConfig(int(seed), int(chain_begin), self.numchains, int(sample_chains), int(flags), int(device), int(mapping),
                          C.c_void_p(stream) if stream else None)
        self.h = C.c_void_p()
        self.flags = flags
        self._numiters = 0
        L.check(L.lib().qsb_sampler_create(program.model(device), arr, len(kernel.specs), C.byref(cfg), C.byref(self.h)))

    def close(self):
        if self.h:
            L.lib().qsb_sampler_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self):
        L.check(L.lib().qsb_sampler_init(self.h))

    def set_state(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.numchains, self.dim)
        L.check(L.lib().qsb_sampler_set_state(self.h, x.ctypes.data_as(L.dp)))

    def get_state(self):
        x = np.empty((self.numchains, self.dim))
        L.check(L.lib().qsb_sampler_get_state(self.h, x.ctypes.data_as(L.dp)))
        return x

    def set_temperatures(self, temps):
        """temps[i] = temperature of iteration i (None removes the schedule)."""
        if temps is None:
            L.check(L.lib().qsb_sampler_set_temperatures(self.h, None, 0))
            return
        t = np.ascontiguousarray(temps, dtype=np.float64)
        L.check(L.lib().qsb_sampler_set_temperatures(self.h, t.ctypes.data_as(L.dp), len(t)))

    def _tabulate_schedule(self, numiters):
        if self.kernel.anneal is not None:
            self.set_temperatures([float(self.kernel.anneal(i, numiters)) for i in range(numiters)])

    def run(self, nsamps, burnin=0, lag=1):
        self._numiters = burnin + nsamps * lag
        self._tabulate_schedule(self._numiters)
        L.check(L.lib().qsb_sampler_run(self.h, nsamps, burnin, lag))

    def begin(self, max_samples, burnin=0, lag=1, numiters=0):
        self._numiters = numiters
        self._tabulate_schedule(numiters)
        L.check(L.lib().qsb_sampler_begin(self.h, max_samples, burnin, lag, numiters))

    def advance(self, iters):
        L.check(L.lib().qsb_sampler_advance(self.h, iters))

    def num_samples(self):
        return int(L.lib().qsb_sampler_num_samples(self.h))

    def sample_dtype(self):
        """C layout of Sample(T) = { T value; double logprob; double loglikelihood } (infer.t:13-18)."""
        return np.dtype([("value", np.float64, (self.retdim,)), ("logprob", np.float64), ("loglikelihood", np.float64)])

    def fetch_samples(self, chain_begin=0, chain_end=None, sample_begin=0, sample_end=None, out=None):
        chain_end = self.sample_chains if chain_end is None else chain_end
        sample_end = self.num_samples() if sample_end is None else sample_end
        dt = self.sample_dtype()
        if out is None:
            out = np.empty((chain_end - chain_begin, sample_end - sample_begin), dtype=dt)
        L.check(L.lib().qsb_sampler_fetch_samples(self.h, chain_begin, chain_end, sample_begin, sample_end,
                                                  out.ctypes.data_as(C.c_void_p), dt.itemsize))
        return out

    def fetch_samples_async(self, chain_begin, chain_end, sample_begin, sample_end, out):
        """Queue the copy of kept samples into the (pinned) structured array ``out`` without waiting; ``fetch_sync`` waits."""
        dt = self.sample_dtype()
        L.check(L.lib().qsb_sampler_fetch_samples_async(self.h, chain_begin, chain_end, sample_begin, sample_end,
                                                        out.ctypes.data_as(C.c_void_p), dt.itemsize))

    def fetch_sync(self):
        L.check(L.lib().qsb_sampler_fetch_sync(self.h))

    def fetch_values(self, chain_begin=0, chain_end=None, sample_begin=0, sample_end=None, out=None):
        chain_end = self.sample_chains if chain_end is None else chain_end
        sample_end = self.num_samples() if sample_end is None else sample_end
        if out is None:
            out = np.empty((chain_end - chain_begin, sample_end - sample_begin, self.retdim))
        L.check(L.lib().qsb_sampler_fetch_values(self.h, chain_begin, chain_end, sample_begin, sample_end, out.ctypes.data_as(L.dp)))
        return out

    def stats(self):
        st = L.Stats()
        L.check(L.lib().qsb_sampler_stats(self.h, C.byref(st)))
        n = len(self.kernel.specs)
        return dict(props_made=st.props_made, props_accepted=st.props_accepted, sub_made=list(st.sub_made)[:n],
                    sub_accepted=list(st.sub_accepted)[:n], grad_evals=st.grad_evals, leapfrog_steps=st.leapfrog_steps,
                    iterations=st.iterations, kernel_launches=st.kernel_launches, seconds=st.seconds, nonfinite=st.nonfinite, divergent=st.divergent)

    def set_kernel_timing(self, on):
        L.check(L.lib().qsb_sampler_set_kernel_timing(self.h, int(bool(on))))

    def kernel_time(self):
        """(seconds, launches) of the dominant kernel since the last call (needs FLAG_TIME_KERNEL)."""
        sec, n = C.c_double(), C.c_uint64()
        L.check(L.lib().qsb_sampler_kernel_time(self.h, C.byref(sec), C.byref(n)))
        return sec.value, n.value

    def diagnostics_raw(self, maxlag=0, out_ptr=None):
        """Sufficient statistics (see qsb_sampler_diagnostics).  ``out_ptr``: device pointer (e.g. a torch
        tensor's data_ptr()) to write into for a following NCCL all-reduce; otherwise a numpy array."""
        if self.flags & L.FLAG_MOMENTS:     # running statistics instead of stored draws: 12 + 3 doubles per dimension
            maxlag, n = -1, self.retdim * 15
        else:
            n = self.retdim * 12 + (self.retdim * (maxlag + 1) if maxlag > 0 else 0)
        if out_ptr is not None:
            L.check(L.lib().qsb_sampler_diagnostics(self.h, maxlag, C.c_void_p(out_ptr), 1))
            return None
        out = np.empty(n)
        L.check(L.lib().qsb_sampler_diagnostics(self.h, maxlag, out.ctypes.data_as(C.c_void_p), 0))
        return out

    def fetch_record(self, leap=False):
        T, Cn, d = self._numiters, self.numchains, self.dim
        state = np.zeros((Cn, T, d)); acc = np.zeros((Cn, T), dtype=np.int32); dh = np.zeros((Cn, T)); step = np.zeros((Cn, T))
        nsub = len(self.kernel.specs)
        kidx = np.zeros((Cn, T), dtype=np.int32) if nsub > 1 else None
        lf = np.zeros((Cn, T, int(self.kernel.specs[0].numSteps), d)) if leap else None
        L.check(L.lib().qsb_sampler_fetch_record(
            self.h, state.ctypes.data_as(L.dp), acc.ctypes.data_as(L.ip), dh.ctypes.data_as(L.dp), step.ctypes.data_as(L.dp),
            kidx.ctypes.data_as(L.ip) if kidx is not None else None, lf.ctypes.data_as(L.dp) if lf is not None else None))
        return dict(state=state, accept=acc, dH=dh, step=step, kidx=kidx, leapfrog=lf)


# ---------------------------------------------------------------------------------------------
class Group:
    """One ``qsb_group``: the chains of one program block-partitioned over several devices of one box, driven from this
    (single) host thread; the library owns one stream per device and the NCCL communicator its diagnostics all-reduce
    runs on.  Same surface as :class:`Sampler`; chain indices are global within the group."""

    def __init__(self, program, kernel, devices, numchains=1, seed=0, chain_begin=0, sample_chains=0, flags=0):
        assert isinstance(program, Program), "MCMC: expected a qs.program"
        self.program, self.kernel = program, kernel
        self.devices = [int(d) for d in devices]
        self.numchains = int(numchains)
        self.retdim, self.dim = program.retdim, program.dim
        arr = (L.KernelSpec * len(kernel.specs))(*kernel.specs)
        cfg = L.RunConfig(int(seed), int(chain_begin), self.numchains, int(sample_chains), int(flags), -1, 0, None)
        dev = (C.c_int32 * len(self.devices))(*self.devices)
        self.h = C.c_void_p()
        self.flags = flags
        desc = program.desc()
        L.check(L.lib().qsb_group_create(C.byref(desc), arr, len(kernel.specs), C.byref(cfg), dev, len(self.devices), C.byref(self.h)))
        self._numiters = 0
        self._nsamps = 0

    def close(self):
        if self.h:
            L.lib().qsb_group_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        return int(L.lib().qsb_group_size(self.h))

    def init(self):
        L.check(L.lib().qsb_group_init(self.h))

    def _tabulate_schedule(self, numiters):
        if self.kernel.anneal is not None:
            t = np.ascontiguousarray([float(self.kernel.anneal(i, numiters)) for i in range(numiters)], dtype=np.float64)
            for i in range(self.size()):
                L.check(L.lib().qsb_sampler_set_temperatures(L.lib().qsb_group_sampler(self.h, i), t.ctypes.data_as(L.dp), len(t)))

    def run(self, nsamps, burnin=0, lag=1):
        self._numiters = burnin + nsamps * lag
        self._nsamps = nsamps
        self._tabulate_schedule(self._numiters)
        L.check(L.lib().qsb_group_run(self.h, nsamps, burnin, lag))

    def begin(self, max_samples, burnin=0, lag=1, numiters=0):
        self._numiters = numiters
        self._tabulate_schedule(numiters)
        L.check(L.lib().qsb_group_begin(self.h, max_samples, burnin, lag, numiters))

    def advance(self, iters):
        L.check(L.lib().qsb_group_advance(self.h, iters))

    def sync(self):
        L.check(L.lib().qsb_group_sync(self.h))

    def num_samples(self):
        return int(L.lib().qsb_sampler_num_samples(L.lib().qsb_group_sampler(self.h, 0)))

    def sample_dtype(self):
        return np.dtype([("value", np.float64, (self.retdim,)), ("logprob", np.float64), ("loglikelihood", np.float64)])

    def fetch_samples(self, chain_begin=0, chain_end=None, sample_begin=0, sample_end=None, out=None):
        chain_end = self.numchains if chain_end is None else chain_end
        sample_end = self.num_samples() if sample_end is None else sample_end
        dt = self.sample_dtype()
        if out is None:
            out = np.empty((chain_end - chain_begin, sample_end - sample_begin), dtype=dt)
        L.check(L.lib().qsb_group_fetch_samples(self.h, chain_begin, chain_end, sample_begin, sample_end,
                                                out.ctypes.data_as(C.c_void_p), dt.itemsize))
        return out

    def stats(self):
        st = L.Stats()
        L.check(L.lib().qsb_group_stats(self.h, C.byref(st)))
        n = len(self.kernel.specs)
        return dict(props_made=st.props_made, props_accepted=st.props_accepted, sub_made=list(st.sub_made)[:n],
                    sub_accepted=list(st.sub_accepted)[:n], grad_evals=st.grad_evals, leapfrog_steps=st.leapfrog_steps,
                    iterations=st.iterations, kernel_launches=st.kernel_launches, seconds=st.seconds, nonfinite=st.nonfinite, divergent=st.divergent)

    def diagnostics_raw(self, maxlag=0):
        """The sufficient statistics of qsb_sampler_diagnostics summed over the shards (NCCL all-reduce inside the library)."""
        if self.flags & L.FLAG_MOMENTS:
            maxlag = -1
        out = np.empty(self.retdim * 15 if maxlag < 0 else self.retdim * 12 + (self.retdim * (maxlag + 1) if maxlag > 0 else 0))
        L.check(L.lib().qsb_group_diagnostics(self.h, maxlag, out.ctypes.data_as(L.dp)))
        return out


class Comm:
    """One rank of a multi-process job (one process per device): the library's own NCCL communicator.  ``uid`` is the
    128-byte id rank 0 obtains from :meth:`unique_id` and the launcher distributes to the other ranks."""

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * L.COMM_ID_BYTES)()
        L.check(L.lib().qsb_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, uid, rank, nranks, device=-1):
        buf = (C.c_uint8 * L.COMM_ID_BYTES).from_buffer_copy(uid)
        self.h = C.c_void_p()
        self.rank, self.nranks = rank, nranks
        L.check(L.lib().qsb_comm_create(buf, rank, nranks, device, C.byref(self.h)))

    def close(self):
        if self.h:
            L.lib().qsb_comm_destroy(self.h)
            self.h = C.c_void_p()

    def allreduce_sum(self, dev_ptr, n, stream=None):
        L.check(L.lib().qsb_comm_allreduce_sum(self.h, C.c_void_p(dev_ptr), n, C.c_void_p(stream) if stream else None))

    def diagnostics_raw(self, sampler, maxlag=0):
        if sampler.flags & L.FLAG_MOMENTS:
            maxlag = -1
        out = np.empty(sampler.retdim * 15 if maxlag < 0 else sampler.retdim * 12 + (sampler.retdim * (maxlag + 1) if maxlag > 0 else 0))
        L.check(L.lib().qsb_sampler_diagnostics_allreduce(sampler.h, self.h, maxlag, out.ctypes.data_as(L.dp)))
        return out


def measure_fp64_peak(which="dmma", seconds=0.3, device=-1):
    """TFLOP/s of dependent DFMA chains ("dfma") or mma.sync.m8n8k4.f64 ("dmma") on this device, measured now."""
    v = C.c_double()
    L.check(L.lib().qsb_measure_fp64_peak(1 if which == "dmma" else 0, float(seconds), int(device), C.byref(v)))
    return v.value


def source_hash():
    return L.lib().qsb_source_hash().decode()


# ---------------------------------------------------------------------------------------------
class SampleVector:
    """The filled ``S.Vector(Sample(T))`` (qs/infer.t:12-36; qs/lib/std.t:219-351): ``data`` is a
    structured array [numchains][nsamps] of {value[retdim], logprob, loglikelihood}.  With
    numchains == 1 it is indexable like the reference's single vector of samples.  When it was filled by
    ``MCMC`` it also keeps the sampler whose kept samples it mirrors, so that queries run on the
    device-resident copy without re-uploading."""

    def __init__(self, data, sampler=None):
        self.data = data
        self._sampler = sampler

    def size(self):
        return self.data.shape[1]

    @property
    def value(self):
        return self.data["value"]

    @property
    def logprob(self):
        return self.data["logprob"]

    def __len__(self):
        return self.data.shape[1]

    def __getitem__(self, i):
        return self.data[0, i]

    # -- plumbing of the device queries
    def _shape(self):
        Cn, N = self.data.shape
        return Cn, N, self.data.dtype["value"].shape[0]

    def _host_args(self):
        d = np.ascontiguousarray(self.data)
        Cn, N, rd = self._shape()
        return d, (d.ctypes.data_as(C.c_void_p), d.dtype.itemsize, Cn, N, rd)


def MCMC(kernel, params=None, **kw):
    """``MCMC(kernel, params)`` -> ``function(program)`` -> ``run(samples)`` (qs/mcmc.t:36-105)."""
    p = dict(params or {}); p.update(kw)
    numsamps = int(p.get("numsamps", 1000))       # mcmc.t:38
    burnin = int(p.get("burnin", 0))              # mcmc.t:39
    lag = int(p.get("lag", 1))                    # mcmc.t:40 (the code's default, not the comment's)
    verbose = p.get("verbose", False)
    numchains = int(p.get("numchains", 1))
    seed = int(p.get("seed", 0))
    device = int(p.get("device", -1))
    chain_begin = int(p.get("chain_begin", 0))
    devices = p.get("devices", None)              # new: chains block-partitioned over these devices (one host thread, qsb_group)

    def method(program):
        assert isinstance(program, Program), "MCMC: argument is not a qs.program"

        def doMCMC(samples=None):
            out = sys.stdout if verbose is True else (verbose or None)
            if devices is not None and len(devices) > 0:
                s = Group(program, kernel, devices, numchains=numchains, seed=seed, chain_begin=chain_begin)
            else:
                s = Sampler(program, kernel, numchains=numchains, seed=seed, device=device, chain_begin=chain_begin)
            try:
                s.init()
                t0 = time.time()
                s.run(numsamps, burnin, lag)
                data = s.fetch_samples()
                if out:   # mcmc.t:73-89, same line formats
                    st = s.stats()
                    pm, pa = st["props_made"], st["props_accepted"]
                    out.write("Acceptance Ratio: %u/%u (%g%%)\n" % (pa, pm, 100.0 * pa / max(pm, 1)))
                    if len(kernel.specs) > 1:   # MixtureKernel:printStats (mcmc.t:256-273)
                        for name, m_, a_ in zip(kernel.names, st["sub_made"], st["sub_accepted"]):
                            out.write("   %s Acceptance Ratio: %u/%u (%g%%)\n" % (name, a_, m_, 100.0 * a_ / max(m_, 1)))
                    out.write("Time: %g\n" % (time.time() - t0))
            except Exception:
                s.close()
                raise
            dev_resident = s if isinstance(s, Sampler) else None    # queries on a group's samples go through the host copy
            if dev_resident is None:
                s.close()
            vec = SampleVector(data, sampler=dev_resident)     # the sampler (and its device-resident samples) lives as long as the vector
            if samples is not None:
                # the reference appends to the caller's vector (samples:insert(), mcmc.t:65); so do we
                if samples.data is not None and samples.data.size and samples.data.shape[0] == data.shape[0] \
                        and samples.data.dtype == data.dtype:
                    samples.data = np.concatenate([samples.data, data], axis=1)
                    samples._sampler = None            # the device copy holds this run's samples only
                else:
                    samples.data = data
                    samples._sampler = dev_resident
            return vec
        return doMCMC
    return method


# ---- queries (qs/infer.t:128-294, the ones the hot path feeds), evaluated on the device through the C ABI
def Samples():
    return lambda program: (lambda samples: samples)


def _expectation(samples, doVariance):
    Cn, N, rd = samples._shape()
    mean = np.empty((Cn, rd)); var = np.empty(Cn) if doVariance else None
    vp = var.ctypes.data_as(L.dp) if doVariance else None
    if samples._sampler is not None and samples._sampler.h:
        L.check(L.lib().qsb_query_expectation(samples._sampler.h, 0, Cn, mean.ctypes.data_as(L.dp), vp))
    else:
        keep, args = samples._host_args()
        L.check(L.lib().qsb_query_expectation_host(*args, mean.ctypes.data_as(L.dp), vp))
    return mean, var


def Expectation(doVariance=False):
    """infer.t:146-187, per chain: the mean starts at value[0] and the weights are summed from i=1,
    so for MCMC samples (loglikelihood = 0) mean = sum(v)/(N-1); variance divides by N.  Returns
    [numchains][retdim] (and [numchains]); squeezed to the reference's shape when numchains == 1."""
    def query(program):
        def run(samples):
            assert samples.size() > 0                          # S.assert(samples:size() > 0) (infer.t:153)
            m, var = _expectation(samples, doVariance)
            m_out = m[:, 0] if m.shape[1] == 1 else m
            if not doVariance:
                return m_out[0] if m.shape[0] == 1 else m_out
            return (m_out[0], var[0]) if m.shape[0] == 1 else (m_out, var)
        return run
    return query


def MAP(program):
    """infer.t:191-206, per chain: the value of the first sample with the largest logprob."""
    def run(samples):
        assert samples.size() > 0
        Cn, N, rd = samples._shape()
        val = np.empty((Cn, rd))
        if samples._sampler is not None and samples._sampler.h:
            L.check(L.lib().qsb_query_map(samples._sampler.h, 0, Cn, val.ctypes.data_as(L.dp), None))
        else:
            keep, args = samples._host_args()
            L.check(L.lib().qsb_query_map_host(*args, val.ctypes.data_as(L.dp), None))
        v = val[:, 0] if rd == 1 else val
        return v[0] if Cn == 1 else v
    return run


def Autocorrelation(mean=None, variance=None):
    """infer.t:213-264, per chain."""
    if mean is not None or variance is not None:
        assert mean is not None and variance is not None, \
            "Autocorrelation: both mean and variance must be provided if one is provided."     # infer.t:214-217

    def query(program):
        def run(samples):
            Cn, N, rd = samples._shape()
            ac = np.empty((Cn, N))
            mp = None
            if mean is not None:
                marr = np.ascontiguousarray(np.broadcast_to(np.reshape(np.asarray(mean, dtype=np.float64), (-1,)), (rd,)))
                mp = marr.ctypes.data_as(L.dp)
            v = float(variance) if variance is not None else 0.0
            if samples._sampler is not None and samples._sampler.h:
                L.check(L.lib().qsb_query_autocorrelation(samples._sampler.h, 0, Cn, mp, v, ac.ctypes.data_as(L.dp)))
            else:
                keep, args = samples._host_args()
                L.check(L.lib().qsb_query_autocorrelation_host(*args, mp, v, ac.ctypes.data_as(L.dp)))
            return ac[0] if Cn == 1 else ac
        return run
    return query


def infer(program, query, method):
    """qs/infer.t:59-70: returns a function; calling it runs ``method`` and applies ``query``."""
    methodfn = method(program)
    queryfn = query(program)

    def run():
        samples = methodfn()
        return queryfn(samples)
    return run
