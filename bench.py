#!/usr/bin/env python
"""bench.py -- the reference's headline metric on the many-chain MCMC hot path.

    python bench.py --gpus N --steps K --warmup W [--workload c3|c2|c4|c5] [--impl reference]

Metric (BASELINE.json): leapfrog steps/sec (all chains) and min-ESS/sec.  `value` is the leapfrog
(c2/c3/c5) or proposal (c4) throughput of the whole job with every input resident in HBM; the
secondary min-ESS/sec is computed from the draws of the timed steps.  A *step* is one MCMC transition
of every chain (c3: L = 20 leapfrog steps each).  Default workload c3 = Bayesian logistic regression
d=100, N=10 000, HMC L=20, 8 192 chains per GPU -- the configuration north_star quotes its target on.

N > 1: launched by torchrun, one rank per GPU; chains are sharded (rank r owns global chain ids
[r*C, (r+1)*C)), no data-path collective; the only collective is one NCCL all-reduce(sum) of the
diagnostics' sufficient statistics, outside the timed region, on the LIBRARY's own communicator (qsb_comm:
torch.distributed only distributes the 128-byte id and provides the barrier / max-over-ranks of the timing contract).
Scaling: c3 is BASELINE's "8 192 chains at 1/2/4/8 GPUs" = STRONG scaling (8 192 chains in total, 1 024 per GPU at N = 8);
the weak figure (8 192 per GPU) is reported in `secondary_weak`.  c2 / c4 / c5 default to weak scaling with the per-GPU
count that gives BASELINE's totals at 8 GPUs (65 536 per GPU; 2^20 / 8; 262 144 / 8).

--impl reference: the CPU restatement of the reference algorithm (oracle/: tape AD + whole-program
re-execution, like qs/hmc.t:325-342) on the host cores, one chain per core, same program, same
kernel parameters, same Philox stream.  The Terra reference itself cannot run (no Terra compiler).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261019
# FP64 peaks are MEASURED IN THIS RUN on this box (qsb_measure_fp64_peak, ~0.3 s each): MEASURED_PEAKS.json has no fp64 entry
DEFAULT_SCALING = {"c3": "strong", "c2": "weak", "c4": "weak", "c5": "weak"}


def workload(name, chains=None):
    """(program builder args, kernel spec dict, chains per GPU, description)."""
    import quicksand_b200 as qs
    if name == "c3":
        X, y, _ = qs.programs.c3_logistic_data(10000, 100)
        return dict(name="c3", prog=qs.programs.logistic(X, y, 2.5), kernel=qs.HMCKernel(numSteps=20), L=20, chains=chains or 8192,
                    desc="Bayesian logistic regression d=100 N=10000 synthetic, HMC L=20 (adaptive step size), fp64",
                    unit_per_iter=20, unit="leapfrog steps/s", arrays=dict(X=X, ybin=y),
                    flops_per_unit=4.0 * 10000 * 100, bound="tensor", peak_kind="dmma")
    if name == "c2":
        prog, _ = qs.programs.c2_correlated_gaussian(10, 0.9)
        return dict(name="c2", prog=prog, kernel=qs.HMCKernel(numSteps=16), L=16, chains=chains or 65536,
                    desc="10-D correlated Gaussian (rho=0.9), HMC L=16 (adaptive step size), fp64",
                    unit_per_iter=16, unit="leapfrog steps/s", arrays={}, flops_per_unit=290.0, bound="fp64", peak_kind="dfma")
    if name == "c4":
        y, sigma = qs.programs.c4_eight_schools_data(1000)
        return dict(name="c4", prog=qs.programs.eight_schools(y, sigma), kernel=qs.DriftKernel(scale=0.02), L=1, chains=chains or 131072,
                    desc="hierarchical 8-schools-style J=1000 (d=1002), drift RW-Metropolis scale=0.02 with scale adaptation, fp64",
                    unit_per_iter=1, unit="proposals/s", arrays=dict(y=y, sigma=sigma), bytes_per_unit=40.0 * 1002, bound="hbm")   # x, mean, m2 read; mean, m2 written
    if name == "c5":
        return dict(name="c5", prog=qs.programs.funnel(1000), kernel=qs.MixtureKernel([qs.HARMKernel(), qs.HMCKernel(numSteps=16)], [0.5, 0.5]),
                    L=16, chains=chains or 32768, desc="1000-D Neal's funnel, MixtureKernel{HARM, HMC L=16} 0.5/0.5, fp64",
                    unit_per_iter=None, unit="leapfrog steps/s", arrays={}, flops_per_unit=15.0 * 1000, bound="fp64", peak_kind="dfma")
    raise SystemExit("unknown workload %s" % name)


# ---------------------------------------------------------------------------------------------
def _oracle_worker(args):
    """One host core: one chain of the CPU restatement.  Returns (grad_evals, seconds, proposals).  The run keeps a
    single sample (lag = transitions) so that long samples of the 1000-dimensional workloads do not materialise every draw."""
    wl_name, chain, transitions, L = args
    import oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    wl = workload(wl_name)
    prog = parity.oracle_program(wl["prog"])
    specs = parity.oracle_specs(wl["kernel"])
    if L is not None:
        for s in specs:
            if s.kind == O.HMC:
                s.numSteps = L
    out = O.run(prog, specs, SEED, 1, 1, 0, transitions, chain_begin=chain)
    st = out["stats"]
    return st["grad_evals"], st["seconds"], st["props_made"]


def cpu_sample(wl_name, cores, transitions, L=None):
    """`cores` chains of the oracle in parallel processes (chains are independent; the reference itself is
    single-threaded: qs/trace.t:452-454, qs/lib/ad.t:9,19).  Returns units/s over all cores and details."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.time()
    with ctx.Pool(cores) as pool:
        res = pool.map(_oracle_worker, [(wl_name, c, transitions, L) for c in range(cores)])
    wall = time.time() - t0
    evals = sum(r[0] for r in res)
    props = sum(r[2] for r in res)
    busy = max(r[1] for r in res)
    return dict(grad_evals=evals, proposals=props, seconds=busy, wall=wall)


def cpu_units_per_sec(wl, s):
    """Throughput in the workload's unit: gradient evaluations (= leapfrog steps, incl. the step-size search probes)
    for HMC workloads, proposals for drift."""
    if wl["name"] == "c4":
        return s["proposals"] / s["seconds"]
    return s["grad_evals"] / s["seconds"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.workload)
    cores = os.cpu_count() or 1
    # warm-up: W tiny runs (page in the library, build the program); also calibrates the cost of one gradient
    t0 = time.time()
    cal = cpu_sample(args.workload, cores, 1, L=1)
    for _ in range(max(args.warmup - 1, 0)):
        cpu_sample(args.workload, cores, 1, L=1)
    per_eval = cal["seconds"] / max(cal["grad_evals"] / cores, 1)
    # each step = one transition of `cores` chains with L_s leapfrog steps, L_s bounded so K steps end within ~200 s
    L_full = wl["L"]
    L_s = L_full
    if wl["name"] != "c4":
        L_s = int(max(1, min(L_full, 200.0 / (max(args.steps, 1) * per_eval) - 0)))
    s = cpu_sample(args.workload, cores, args.steps, L=L_s if wl["name"] != "c4" else None)
    value = cpu_units_per_sec(wl, s)
    line = {
        "impl": "reference", "metric": "leapfrog steps/sec (all chains)" if wl["name"] != "c4" else "proposals/sec (all chains)",
        "value": value, "unit": wl["unit"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * s["seconds"] / max(args.steps, 1), "higher_is_better": True, "scaling": DEFAULT_SCALING.get(wl["name"], "weak"), "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (wl["name"], wl["desc"]), "chains": cores, "leapfrog_per_step": L_s},
        "cpu_baseline": {"value": value, "unit": wl["unit"], "cores": cores, "kind": "port",
                         "sample": "%d chains (one per core) x %d transitions x %d leapfrog steps of the CPU restatement of the reference "
                                   "algorithm (tape AD, whole-program re-execution); gradient evaluations / busy seconds" % (cores, args.steps, L_s)},
        "e2e": {"value": value, "unit": wl["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU DURING the timed region, sampled in-process through NVML by a thread (every
    ~10 ms; starts before the region so that even a 3 ms region -- c2 -- has samples on both sides of it)."""

    REASONS = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]

    def __init__(self, index):
        import threading
        self.samples = []        # (time, sm_mhz, reasons bitmask, power W)
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch / CUDA_VISIBLE_DEVICES index -> NVML handle by UUID-free best effort: visible index into the visible list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                toks = [t.strip() for t in vis.split(",") if t.strip()]
                if index < len(toks) and toks[index].isdigit():
                    phys = int(toks[index])
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:       # no NVML: report that, never a made-up clock
            self.err = repr(e)
        self.t0 = self.t1 = None
        if self.ok:
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()

    def _once(self):
        nv, h = self.nv, self.h
        try:
            sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            try:
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            try:
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = None
            self.samples.append((time.time(), sm, rs, pw))
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._once()
            time.sleep(0.01)

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.ok:
            out["error"] = "NVML unavailable: " + getattr(self, "err", "?")
            return out
        self._once()
        self._stop.set()
        self.th.join(timeout=1.0)
        inside = [x for x in self.samples if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or 1e300)]
        note = None
        if not inside:      # region shorter than one polling period: the samples bracketing it
            before = [x for x in self.samples if x[0] < self.t0][-1:]
            after = [x for x in self.samples if x[0] > self.t1][:1]
            inside = before + after
            note = "timed region shorter than the 10 ms polling period: the samples bracketing it"
        if not inside:
            return out
        sm = [x[1] for x in inside]
        mask = 0
        for x in inside:
            mask |= x[2]
        pw = [x[3] for x in inside if x[3] is not None]
        out = {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": self.max_mhz,
               "reasons": [n for n, b in self.REASONS if mask & b], "samples": len(inside), "source": "NVML in-process",
               "power_w_max": max(pw) if pw else None}
        if note:
            out["when"] = note
        return out


def tree_hash():
    from quicksand_b200 import build as B
    return B.source_hash()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=os.environ.get("QSB_WORKLOAD", "c3"))
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU (weak) / in total (strong)")
    ap.add_argument("--balanced", action="store_true", help="(no-op: the balanced decomposition of the c3 gradient kernel is the default since round 2)")
    ap.add_argument("--invariant", action="store_true",
                    help="c3: QSB_FLAG_INVARIANT (the fixed chain-tile x 12-row-split grid: draws bit-identical for any chain partition; slower on small shards)")
    ap.add_argument("--roofline-steps", type=int, default=None, help="steps of the separate pass that times every launch of the dominant kernel")
    ap.add_argument("--scaling", default=os.environ.get("QSB_SCALING"), choices=["weak", "strong"],
                    help="strong: the configuration's chain count in total, sharded over the GPUs (default for c3 = BASELINE's 8 192 chains "
                         "at 1/2/4/8 GPUs); weak: that count per GPU (default for c2, c4, c5)")
    ap.add_argument("--burnin", type=int, default=None, help="untimed adaptation transitions before the warm-up steps")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--ess-steps", type=int, default=None, help="untimed transitions after the timed region whose draws estimate ESS / split-Rhat")
    ap.add_argument("--ess-lag", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the secondary weak-scaling measurement of a strong-scaling run")
    ap.add_argument("--maxlag", type=int, default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(args.warmup, 3)
    scaling = args.scaling or DEFAULT_SCALING.get(args.workload, "weak")

    # CPU baseline first (fork-based pool before CUDA is initialised); rank 0 at N=1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        wl0 = workload(args.workload)
        trans = {"c3": 16, "c2": 100000, "c4": 60000, "c5": 6000}[wl0["name"]]     # ~10-15 s of CPU work per core
        s = cpu_sample(args.workload, cores, trans)
        cpu = {"value": cpu_units_per_sec(wl0, s), "unit": wl0["unit"], "cores": cores, "kind": "port",
               "sample": "%d chains (one per core) x %d transitions from initialisation of the CPU restatement of the reference algorithm "
                         "(tape AD + whole-program re-execution); %.1f s busy, %.1f s wall; units = gradient evaluations (HMC) or proposals"
                         % (cores, trans, s["seconds"], s["wall"])}

    import torch
    import torch.distributed as dist
    import quicksand_b200 as qs
    from quicksand_b200 import _lib as L, diagnostics as D
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # the data-plane collective (diagnostics all-reduce) runs on the library's own NCCL communicator; torch.distributed
        # only hands the 128-byte id around
        box = [qs.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = qs.Comm(box[0], rank, world, device=local)
    wl = workload(args.workload, args.chains)
    prog, kernel, Ctot = wl["prog"], wl["kernel"], wl["chains"]
    C = max(Ctot // world, 1) if scaling == "strong" else Ctot   # block partition of the global chain range over the ranks
    B = args.burnin if args.burnin is not None else {"c3": 150, "c2": 300, "c4": 300, "c5": 100}[wl["name"]]
    Ke = args.e2e_steps if args.e2e_steps is not None else min(K, 10)
    # draws for ESS / split-Rhat come from an untimed segment after the timed region (the timed K alone -- 20 at the driver's
    # setting -- is too short for an autocovariance); same kernel, same adaptation (which, as in the reference, never stops)
    Kd = args.ess_steps if args.ess_steps is not None else {"c3": 200, "c2": 400, "c4": 400, "c5": 200}[wl["name"]]
    dlag = args.ess_lag if args.ess_lag is not None else 1
    SCn = min(C, {"c3": 8192, "c2": 16384, "c4": 2048, "c5": 2048}[wl["name"]])     # chains whose draws are kept
    stream = torch.cuda.current_stream().cuda_stream

    def make_sampler(chains, begin):
        return qs.Sampler(prog, kernel, numchains=chains, seed=SEED, device=local, chain_begin=begin,
                          flags=(L.FLAG_INVARIANT if args.invariant else 0), stream=stream, sample_chains=min(chains, SCn))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_run(sampler, Kt, clocks=None):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if clocks:
            clocks.mark_begin()
        ev0.record()
        for _ in range(Kt):
            sampler.advance(1)
        ev1.record()
        barrier()
        if clocks:
            clocks.mark_end()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() * 1e-3

    def units_of(st1, st0):
        key = "props_made" if wl["name"] == "c4" else "leapfrog_steps"
        return st1[key] - st0[key]

    sampler = make_sampler(C, rank * C)
    sampler.init()
    n_keep = K + Ke + (Kd + dlag - 1) // dlag
    # kept draws: the K timed ones, the Ke end-to-end ones (lag 1), then every dlag-th of the Kd diagnostic ones
    Kr_plan = args.roofline_steps if args.roofline_steps is not None else min(K, 10)
    sampler.begin(K + Ke, burnin=B + W, lag=1, numiters=B + W + K + Ke + Kr_plan)
    sampler.advance(B)                      # adaptation / burn-in (untimed set-up; the reference adapts forever, so do we)
    for _ in range(W):
        sampler.advance(1)                  # warm-up steps
    torch.cuda.synchronize()
    st0 = sampler.stats()
    clocks = ClockSampler(local) if rank == 0 else None
    time.sleep(0.03)
    t = timed_run(sampler, K, clocks)
    clk = clocks.stop() if clocks else None
    st1 = sampler.stats()

    units_local = units_of(st1, st0)
    cnt = torch.tensor([float(units_local), float(st1["props_accepted"] - st0["props_accepted"]), float(st1["props_made"] - st0["props_made"])],
                       device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(cnt)
    units = cnt[0].item()
    value = units / t
    launches = st1["kernel_launches"] - st0["kernel_launches"]

    # ---- same-box FP64 peaks (roofline denominators), measured now
    peaks = {}
    if rank == 0 and wl["bound"] != "hbm":
        peaks = {"dmma": qs.measure_fp64_peak("dmma", 0.3, local), "dfma": qs.measure_fp64_peak("dfma", 0.3, local)}

    # ---- end to end through the host-facing API with HOST buffers: every step uploads the observation data from pinned
    # host memory, runs one transition of all chains and reads that step's draws back into pinned host memory.  The
    # read-back of step i (gather on the compute stream, D2H on the library's copy stream, two staging buffers and two
    # pinned destinations alternating) overlaps the transition of step i+1.
    h2d = d2h = 0
    e2e_val = None
    if Ke > 0:
        pinned = {}
        for k, v in wl["arrays"].items():
            tns = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
            pinned[k] = tns
            h2d += tns.numel() * tns.element_size()
        nsc = sampler.sample_chains
        dt = sampler.sample_dtype()
        outs = []
        for _ in range(2):
            ot = torch.empty(nsc * dt.itemsize, dtype=torch.uint8).pin_memory()
            outs.append((ot, ot.numpy().view(dt).reshape(nsc, 1)))
        d2h = nsc * dt.itemsize
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # warm-up of the read-back path (untimed, like the W warm-up transitions): the first two asynchronous fetches create the
        # copy stream and allocate the two device staging buffers -- at c2's 41 us per step that one-off cost was 90 % of the
        # timed region
        for b in range(min(2, K)):
            sampler.fetch_samples_async(0, nsc, b, b + 1, outs[b][1])
        sampler.fetch_sync()
        if pinned:
            prog.update_data(device=local, stream=stream, **{k: v.numpy() for k, v in pinned.items()})
        ue0 = sampler.stats()
        barrier()
        e0.record()
        for i in range(Ke):
            if pinned:
                prog.update_data(device=local, stream=stream, **{k: v.numpy() for k, v in pinned.items()})
            sampler.advance(1)
            sampler.fetch_samples_async(0, nsc, K + i, K + i + 1, outs[i % 2][1])
        sampler.fetch_sync()                 # the last copies have landed in host memory
        e1.record()
        barrier()
        ue1 = sampler.stats()
        mse = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        ucnt = torch.tensor([float(units_of(ue1, ue0))], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(mse, op=dist.ReduceOp.MAX); dist.all_reduce(ucnt)
        e2e_val = ucnt.item() / (mse.item() * 1e-3)

    # ---- roofline pass: the same transitions with a CUDA-event pair around every launch of the dominant kernel (the event
    # records sit between the launches of a transition and would suppress their programmatic-dependent-launch overlap, so
    # the throughput above is measured without them); these steps are not kept and not part of `value`
    Kr = args.roofline_steps if args.roofline_steps is not None else min(K, 10)
    ktime, klaunch, t_roof, units_roof = 0.0, 0, 0.0, 0
    if Kr > 0:
        sampler.set_kernel_timing(True)
        sampler.kernel_time()               # reset the dominant-kernel timer
        r0 = sampler.stats()
        t_roof = timed_run(sampler, Kr)
        ktime, klaunch = sampler.kernel_time()
        units_roof = units_of(sampler.stats(), r0)
        sampler.set_kernel_timing(False)

    # ---- diagnostics: an untimed segment of Kd transitions (every dlag-th kept); sufficient statistics on the device,
    # ONE all-reduce(sum) inside the library (qsb_comm), Geyer ESS on the host
    diag = {}
    min_ess = rhat_max = None
    nd = 0
    if Kd >= 4 * dlag:
        nd = Kd // dlag
        sampler.begin(nd, burnin=0, lag=dlag, numiters=nd * dlag)
        sampler.advance(nd * dlag)
        maxlag = args.maxlag if args.maxlag is not None else max(1, min(nd // 4, 50))
        raw = comm.diagnostics_raw(sampler, maxlag) if comm else sampler.diagnostics_raw(maxlag)
        diag = D.summarize(raw, sampler.retdim, maxlag)          # (the library already undid the sum of the per-shard constants)
        if np.isfinite(diag["rhat"]).any():
            rhat_max = float(np.nanmax(diag["rhat"]))
        if "ess" in diag and np.isfinite(diag["ess"]).any():
            min_ess = float(np.nanmin(diag["ess"]))
    converged = rhat_max is not None and rhat_max <= 1.05
    chains_diag = float(sampler.sample_chains) * world

    # ---- strong-scaling runs also report the weak figure (the configuration's chain count on EVERY GPU)
    weak = None
    if scaling == "strong" and world > 1 and not args.no_weak:
        sampler.close()
        sampler = make_sampler(Ctot, rank * Ctot)
        sampler.init()
        Bw, Kw = min(B, 40), min(K, 10)
        sampler.begin(0, burnin=0, lag=1, numiters=Bw + W + Kw)
        sampler.advance(Bw + W)
        torch.cuda.synchronize()
        w0 = sampler.stats()
        tw = timed_run(sampler, Kw)
        w1 = sampler.stats()
        wu = torch.tensor([float(units_of(w1, w0))], device="cuda", dtype=torch.float64)
        dist.all_reduce(wu)
        weak = {"scaling": "weak", "value": wu.item() / tw, "unit": wl["unit"], "chains_per_gpu": Ctot, "chains_total": Ctot * world,
                "steps": Kw, "ms_per_step": 1e3 * tw / Kw, "burnin_transitions": Bw}

    if rank == 0:
        if cpu is not None:
            cpu["note"] = ("throughput of the CPU restatement in the metric's unit; a min-ESS/s of the CPU arm is NOT measured (its sample is far "
                           "too short for an autocovariance) -- with the same algorithm and random stream the ESS per transition is the "
                           "same on both sides, so the min-ESS/s ratio equals the throughput ratio")
        per_launch_units = units_roof / max(klaunch, 1)
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel: from the ncu --set full captures under
        # profiles/ (a constant of the kernel at that chain count, not measurable outside a profiler); null at other chain counts
        traffic_tab = {"c3": {8192: 40.0e6}, "c2": {65536: 14.6e6}, "c4": {131072: 5.46e9}, "c5": {32768: 0.533e9}}[wl["name"]]
        tfile = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tfile):
            try:
                for kk, vv in json.load(open(tfile)).get(wl["name"], {}).items():
                    traffic_tab[int(kk)] = float(vv)
            except Exception:
                pass
        common = {"traffic": traffic_tab.get(C), "traffic_source": "ncu --set full capture under profiles/ (dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                  "kernel": "glm_grad_kernel" if wl["name"] == "c3" else "advance_kernel",
                  "kernel_ms_per_launch": 1e3 * ktime / max(klaunch, 1), "kernel_launches_timed": klaunch,
                  "kernel_share_of_step": (ktime / t_roof) if t_roof else None,
                  "timing_pass": "%d extra steps after the timed region with a CUDA-event pair around every launch of the kernel "
                                 "(%.3f ms per step in that pass vs %.3f ms in the timed region)" % (Kr, 1e3 * t_roof / max(Kr, 1), 1e3 * t / K)}
        if wl["bound"] == "hbm":
            achieved = wl["bytes_per_unit"] * per_launch_units / (ktime / max(klaunch, 1)) / 1e9 if klaunch else float("nan")
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            peak = float(mp.get("hbm_gbs", 6650.0))
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if mp else "fallback 6.65 TB/s"}
        else:
            achieved = wl["flops_per_unit"] * per_launch_units / (ktime / max(klaunch, 1)) / 1e12 if klaunch else float("nan")
            peak = peaks[wl["peak_kind"]]
            roof = {"bound": wl["bound"], "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "peak_source": "measured in this run on this GPU by qsb_measure_fp64_peak (%s: %s), best launch over ~0.3 s; "
                                   "MEASURED_PEAKS.json has no fp64 entry" % (wl["peak_kind"], "mma.sync.m8n8k4.f64" if wl["peak_kind"] == "dmma" else "8 independent DFMA chains per thread"),
                    "fp64_peaks_tflops": peaks}
        roof.update(common)
        line = {
            "metric": "leapfrog steps/sec (all chains)" if wl["name"] != "c4" else "proposals/sec (all chains)",
            "value": value, "unit": wl["unit"], "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * t / K,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (wl["name"], wl["desc"]), "chains_per_gpu": C, "chains_total": C * world,
                       "burnin_transitions": B, "l2": "working set (chain state + partial sums) larger than L2; inputs change every step",
                       "seed": SEED, **({"glm_decomposition": "invariant grid (QSB_FLAG_INVARIANT)"} if args.invariant else {})},
            "secondary": {"metric": "min-ESS/sec", "unit": "effective samples/s",
                          # ESS per kept draw x kept-draw rate of the timed region; only reported for converged chains
                          "value": (min_ess / (nd * dlag) * K / t * (C * world / chains_diag)) if (min_ess and converged) else None,
                          "min_ess": min_ess if converged else None, "converged": bool(converged), "max_split_rhat": rhat_max,
                          "draws_per_chain": nd, "draw_lag": dlag, "chains_in_estimate": int(chains_diag), "maxlag": (maxlag if nd else None),
                          "ess_truncated_at_maxlag": bool(np.any(diag.get("truncated_at_maxlag", False))) if nd else None,
                          "note": "draws of an untimed segment of %d transitions after the timed region; step-size / scale adaptation keeps running "
                                  "(as in the reference: hmc.t:159-161); min_ess is null unless max split-Rhat <= 1.05" % (nd * dlag),
                          "accept_rate": cnt[1].item() / max(cnt[2].item(), 1.0)},
            "roofline": roof, "cpu_baseline": cpu, "clocks": clk,
            "e2e": {"value": e2e_val, "unit": wl["unit"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "overlap": "D2H of step i on a copy stream, overlapping the transition of step i+1 (double-buffered)"},
            "gpu_launches": int(launches),
            "build": {"source_hash": qs.source_hash(), "tree_hash": tree_hash(), "matches_tree": qs.source_hash() == tree_hash()},
        }
        if weak:
            line["secondary_weak"] = weak
        print(json.dumps(line))
    sampler.close()
    if comm:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
