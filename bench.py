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
diagnostics' sufficient statistics, outside the timed region.

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
FP64_DMMA_PEAK_TFLOPS = 36.9      # own micro-benchmark on this pool's B200 (profiles/r01_microbench_b200.json: dmma_m8n8k4)
FP64_DFMA_PEAK_TFLOPS = 33.8      # same file: dfma


def workload(name, chains=None):
    """(program builder args, kernel spec dict, chains per GPU, description)."""
    import quicksand_b200 as qs
    if name == "c3":
        X, y, _ = qs.programs.c3_logistic_data(10000, 100)
        return dict(name="c3", prog=qs.programs.logistic(X, y, 2.5), kernel=qs.HMCKernel(numSteps=20), L=20, chains=chains or 8192,
                    desc="Bayesian logistic regression d=100 N=10000 synthetic, HMC L=20 (adaptive step size), fp64",
                    unit_per_iter=20, unit="leapfrog steps/s", arrays=dict(X=X, ybin=y),
                    flops_per_unit=4.0 * 10000 * 100, bound="tensor", peak=FP64_DMMA_PEAK_TFLOPS)
    if name == "c2":
        prog, _ = qs.programs.c2_correlated_gaussian(10, 0.9)
        return dict(name="c2", prog=prog, kernel=qs.HMCKernel(numSteps=16), L=16, chains=chains or 65536,
                    desc="10-D correlated Gaussian (rho=0.9), HMC L=16 (adaptive step size), fp64",
                    unit_per_iter=16, unit="leapfrog steps/s", arrays={}, flops_per_unit=290.0, bound="fp64", peak=FP64_DFMA_PEAK_TFLOPS)
    if name == "c4":
        y, sigma = qs.programs.c4_eight_schools_data(1000)
        return dict(name="c4", prog=qs.programs.eight_schools(y, sigma), kernel=qs.DriftKernel(scale=0.02), L=1, chains=chains or 131072,
                    desc="hierarchical 8-schools-style J=1000 (d=1002), drift RW-Metropolis scale=0.02 with scale adaptation, fp64",
                    unit_per_iter=1, unit="proposals/s", arrays=dict(y=y, sigma=sigma), bytes_per_unit=40.0 * 1002, bound="hbm")   # x, mean, m2 read; mean, m2 written
    if name == "c5":
        return dict(name="c5", prog=qs.programs.funnel(1000), kernel=qs.MixtureKernel([qs.HARMKernel(), qs.HMCKernel(numSteps=16)], [0.5, 0.5]),
                    L=16, chains=chains or 32768, desc="1000-D Neal's funnel, MixtureKernel{HARM, HMC L=16} 0.5/0.5, fp64",
                    unit_per_iter=None, unit="leapfrog steps/s", arrays={}, flops_per_unit=15.0 * 1000, bound="fp64", peak=FP64_DFMA_PEAK_TFLOPS)
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
        "ms_per_step": 1e3 * s["seconds"] / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
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
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        # a timed region shorter than nvidia-smi's start-up (c2: 3 ms) leaves no sample: take the first one right after it
        late = False
        t_end = time.time() + 2.0
        while os.path.getsize(self.f.name) == 0 and time.time() < t_end:
            late = True
            time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            parts = [t.strip() for t in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
            if late:
                out["when"] = "first sample after the timed region (region shorter than the sampler's start-up)"
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=os.environ.get("QSB_WORKLOAD", "c3"))
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU")
    ap.add_argument("--balanced", action="store_true",
                    help="c3: QSB_FLAG_BALANCED (one equal run of row blocks per SM in the gradient kernel; draws then depend on the chain tile in the last bits)")
    ap.add_argument("--scaling", default=os.environ.get("QSB_SCALING", "weak"), choices=["weak", "strong"],
                    help="weak (default): the configuration's chain count per GPU; strong: that chain count in total, sharded over the GPUs")
    ap.add_argument("--burnin", type=int, default=None, help="untimed adaptation transitions before the warm-up steps")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--maxlag", type=int, default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(args.warmup, 3)

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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = workload(args.workload, args.chains)
    prog, kernel, C = wl["prog"], wl["kernel"], wl["chains"]
    if args.scaling == "strong":
        C = max(C // world, 1)            # BASELINE's chain count in total (c3: 8 192 = 8 x 1 024), block partition over the ranks
    B = args.burnin if args.burnin is not None else {"c3": 150, "c2": 300, "c4": 300, "c5": 100}[wl["name"]]
    Ke = args.e2e_steps if args.e2e_steps is not None else min(K, 10)
    stream = torch.cuda.current_stream().cuda_stream
    sampler = qs.Sampler(prog, kernel, numchains=C, seed=SEED, device=local, chain_begin=rank * C, flags=L.FLAG_TIME_KERNEL | (L.FLAG_BALANCED if args.balanced else 0),
                         stream=stream, sample_chains=min(C, 16384))
    sampler.init()
    sampler.begin(K + Ke, burnin=B + W, lag=1, numiters=B + W + K + Ke)
    sampler.advance(B)                      # adaptation / burn-in (untimed set-up; the reference adapts forever, so do we)
    for _ in range(W):
        sampler.advance(1)                  # warm-up steps
    torch.cuda.synchronize()
    st0 = sampler.stats()
    sampler.kernel_time()                   # reset the dominant-kernel timer

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    ev0.record()
    for _ in range(K):
        sampler.advance(1)
    ev1.record()
    barrier()
    clk = clocks.stop() if clocks else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    t = ms.item() * 1e-3
    ktime, klaunch = sampler.kernel_time()
    st1 = sampler.stats()

    # units of work done in the timed region (all ranks)
    if wl["name"] == "c4":
        units_local = st1["props_made"] - st0["props_made"]
    else:
        units_local = st1["leapfrog_steps"] - st0["leapfrog_steps"]
    cnt = torch.tensor([float(units_local), float(st1["props_accepted"] - st0["props_accepted"]), float(st1["props_made"] - st0["props_made"])],
                       device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(cnt)
    units = cnt[0].item()
    value = units / t
    launches = st1["kernel_launches"] - st0["kernel_launches"]

    # ---- diagnostics over the K timed draws: sufficient statistics on the device, ONE all-reduce(sum), Geyer ESS on the host
    maxlag = args.maxlag if args.maxlag is not None else min(K - 1, 40)
    retdim = sampler.retdim
    raw = torch.zeros(retdim * 12 + retdim * (maxlag + 1), device="cuda", dtype=torch.float64)
    # statistics over the first K kept draws only: sampler currently holds exactly K draws
    sampler.diagnostics_raw(maxlag, out_ptr=raw.data_ptr())
    if world > 1:
        dist.all_reduce(raw)
    diag = D.summarize(raw.cpu().numpy(), retdim, maxlag, nshards=world)
    # too few timed draws for an autocovariance (K < 3): no ESS rather than a made-up one
    min_ess = float(np.nanmin(diag["ess"])) if "ess" in diag and np.isfinite(diag["ess"]).any() else None
    rhat_max = float(np.nanmax(diag["rhat"])) if np.isfinite(diag["rhat"]).any() else None

    # ---- end to end through the host-facing API with HOST buffers: every step uploads the observation data from pinned
    # host memory, runs one transition of all chains and reads that step's draws back into pinned host memory
    h2d = d2h = 0
    e2e_val = None
    if Ke > 0:
        pinned = {}
        for k, v in wl["arrays"].items():
            tns = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
            pinned[k] = tns
            h2d += tns.numel() * tns.element_size()
        SCn = sampler.sample_chains
        dt = sampler.sample_dtype()
        out_t = torch.empty(SCn * dt.itemsize, dtype=torch.uint8).pin_memory()
        out_np = out_t.numpy().view(dt).reshape(SCn, 1)
        d2h = SCn * dt.itemsize
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ue0 = sampler.stats()
        barrier()
        e0.record()
        for i in range(Ke):
            if pinned:
                prog.update_data(device=local, stream=stream, **{k: v.numpy() for k, v in pinned.items()})
            sampler.advance(1)
            sampler.fetch_samples(0, SCn, K + i, K + i + 1, out=out_np)
        e1.record()
        barrier()
        ue1 = sampler.stats()
        mse = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        ucnt = torch.tensor([float((ue1["props_made"] - ue0["props_made"]) if wl["name"] == "c4" else (ue1["leapfrog_steps"] - ue0["leapfrog_steps"]))],
                            device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(mse, op=dist.ReduceOp.MAX); dist.all_reduce(ucnt)
        e2e_val = ucnt.item() / (mse.item() * 1e-3)

    if rank == 0:
        if cpu is not None and wl["unit_per_iter"]:
            # Same algorithm, same random stream: the effective sample size gained per chain-transition is the same on both
            # sides, so the CPU arm's min-ESS/s is its chain-transitions/s times the ESS per chain-transition measured above
            # (the CPU sample itself is far too short to estimate an ESS from).
            ess_per_transition = (min_ess or 0.0) / float(C * world * K)
            cpu["min_ess_per_sec"] = ess_per_transition * cpu["value"] / float(wl["unit_per_iter"]) if min_ess else None
            cpu["min_ess_note"] = "CPU chain-transitions/s x min-ESS per chain-transition of the GPU draws (same algorithm and stream)"
        per_launch_units = units_local / max(klaunch, 1)
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at the default chain counts, from the
        # ncu --set full captures summarised in profiles/r01j_c3_grad_ncu_metrics.csv and profiles/r01i_c{2,4,5}_adv_ncu_metrics.csv
        traffic = {"c3": (8192, 40.0e6), "c2": (65536, 14.6e6), "c4": (131072, 5.46e9), "c5": (32768, 0.533e9)}[wl["name"]]   # (chains per GPU, bytes)
        common = {"traffic": traffic[1] if C == traffic[0] else None,
                  "kernel": "glm_grad_kernel" if wl["name"] == "c3" else "advance_kernel",
                  "kernel_ms_per_launch": 1e3 * ktime / max(klaunch, 1), "kernel_launches_timed": klaunch,
                  "kernel_share_of_step": ktime / t}
        if wl["bound"] == "hbm":
            achieved = wl["bytes_per_unit"] * per_launch_units / (ktime / max(klaunch, 1)) / 1e9
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            peak = float(peaks.get("hbm_gbs", 6650.0))
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"}
        else:
            achieved = wl["flops_per_unit"] * per_launch_units / (ktime / max(klaunch, 1)) / 1e12
            roof = {"bound": wl["bound"], "achieved": achieved, "peak": wl["peak"], "unit": "TFLOP/s", "frac": achieved / wl["peak"],
                    "peak_source": "fp64 peak measured by tools/microbench.cu on this pool (profiles/r01_microbench_b200.json); "
                                   "MEASURED_PEAKS.json has no fp64 entry"}
        roof.update(common)
        line = {
            "metric": "leapfrog steps/sec (all chains)" if wl["name"] != "c4" else "proposals/sec (all chains)",
            "value": value, "unit": wl["unit"], "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * t / K,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (wl["name"], wl["desc"]), "chains_per_gpu": C, "chains_total": C * world,
                       "burnin_transitions": B, "l2": "working set (chain state + partial sums) larger than L2; inputs change every step",
                       "seed": SEED, **({"glm_decomposition": "balanced (QSB_FLAG_BALANCED)"} if args.balanced else {})},
            "secondary": {"metric": "min-ESS/sec", "value": (min_ess / t) if min_ess else None, "unit": "effective samples/s", "min_ess": min_ess,
                          "draws_per_chain": K, "max_split_rhat": rhat_max, "maxlag": maxlag,
                          "accept_rate": cnt[1].item() / max(cnt[2].item(), 1.0)},
            "roofline": roof, "cpu_baseline": cpu, "clocks": clk,
            "e2e": {"value": e2e_val, "unit": wl["unit"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke},
            "gpu_launches": int(launches),
        }
        print(json.dumps(line))
    sampler.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
