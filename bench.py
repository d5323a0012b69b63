#!/usr/bin/env python
"""bench.py -- headline benchmark of the Monte-Carlo master-equation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload r1|r1b|z1] [--trajectories N] [--impl ours|reference]

A "step" is one pass of the hot path over one ensemble of synthetic trajectories.  Default workload `r1` is
BASELINE.json configs[1]: single-atom two-photon Rydberg Rabi oscillation with Doppler/position noise, laser phase
noise and intermediate decay, 10^6 trajectories (`get_default_configs()[1]` of the reference's default.jl with
n_samples=10^6, laser_noise=true).  With N>1 (torchrun, one rank per GPU) every rank integrates its own 10^6
trajectories of a N*10^6 ensemble (weak scaling) and the sums are combined by one NCCL all-reduce.

Prints ONE JSON line (rank 0).  `value` = trajectories/s with everything resident on the device (results stay in
HBM); `e2e` = the same metric through the public `simulation(cfg)` call with host buffers in and out.
`--impl reference` times the reference's CPU path (the C oracle port: Julia is not installable here) on the host
cores with the same metric/config on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

F_RHS_1 = 340.0    # canonical flop per one-atom RHS (frozen)
F_RHS_2 = 7500.0   # canonical flop per two-atom RHS
F_NOISE_PER = 3.0  # flop per (node, frequency) actually executed by the Reinsch recurrence (SURVEY counts 8)
# DRAM bytes per trajectory of the dominant kernel from the committed `ncu --set full` captures (dram__bytes_read.sum +
# dram__bytes_write.sum divided by the trajectories of the captured launch; profiles/r1/k_lindblad{1,2}_ncu_summary.txt).
# K3: the per-warp phase-noise tables (2 x 1001 nodes x 8 B written and read once) are the only real traffic.
NCU_DRAM_BYTES_PER_TRAJ = {1: (760.0e6 + 772.5e6) / 37888, 2: (12.3e6 + 13.5e6) / 2368}


def workload(pkg, name, n):
    cfg, cz = pkg.get_default_configs()
    if name in ("r1", "r1b"):
        cfg.laser_noise = True
        cfg.n_samples = n
        if name == "r1b":
            T0 = cfg.tspan[20]
            cfg.tspan = np.array([0.0, T0 / 2])
        desc = {"r1": "R1: 1-atom two-photon Rydberg Rabi, default.jl cfg, laser_noise=true, tspan=0:T0/20:5T0 (nt=101)",
                "r1b": "R1b: same model, pi pulse tspan=[0,T0/2]"}[name]
        return cfg, 1, desc
    cz.laser_noise = True
    cz.n_samples = n
    return cz, 2, "Z1: 2-atom CZ (Levine-Pichler), default.jl cfg_czlp, laser_noise=true, tspan=[0,2tau]"


def noise_coarse(cfg):
    """The coarsening factor the one-atom kernel picks for its phase-noise tables (choose_noise_coarse, ca_device.cuh)."""
    dtn = cfg.tspan[-1] / 1000.0
    for c in (32, 16, 8, 4, 2):
        if ((1000 + c - 1) // c + 8) * c > 2 * 1001:
            continue
        b = max(float(np.sum(np.abs(a) * (2 * np.pi * np.abs(cfg.f) * c * dtn) ** 8))
                for a in (cfg.red_laser_phase_amplitudes, cfg.blue_laser_phase_amplitudes))
        if b * 43.07 / 40320.0 <= 1e-14:
            return c
    return 1


def algorithmic_flops(cfg, natoms, stats, n):
    d = 5 if natoms == 1 else 25
    nt = len(cfg.tspan)
    f_rhs = F_RHS_1 if natoms == 1 else F_RHS_2
    traces = 2 * natoms if cfg.laser_noise else 0
    # flops the table generation really executes: the one-atom kernel sums only every c-th node and interpolates the rest
    c = noise_coarse(cfg) if natoms == 1 else 1
    nodes_summed = 1001 if c == 1 else (1000 + c - 1) // c + 8
    f_noise = traces * (nodes_summed * len(cfg.f) * F_NOISE_PER + (0 if c == 1 else 1001 * 16)) * n
    f_obs = nt * (8 * d ** 3 + 4 * d ** 2) * n
    return stats["n_rhs"] * f_rhs + f_noise + f_obs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference(pkg, oracle, cfg, natoms, budget_s, threads):
    """Reference CPU path (oracle port, OpenMP over trajectories) on a bounded sample of the same workload."""
    model = pkg.model_from_config(cfg, two_atom=(natoms == 2))
    kw = dict(f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes, nthreads=threads)
    if natoms == 2:
        kw["ode"] = pkg._abi.ode_opts(maxiters=10 ** 7)
    run = oracle.rydberg1_run if natoms == 1 else oracle.rydberg2_run
    n0 = max(threads, 8)
    t = time.perf_counter()
    run(model, cfg.tspan, cfg.ψ0, n0, **kw)
    dt0 = time.perf_counter() - t
    n = int(max(n0, min(20000, budget_s / max(dt0, 1e-3) * n0)))
    n -= n % threads
    t = time.perf_counter()
    _, _, st = run(model, cfg.tspan, cfg.ψ0, n, **kw)
    dt = time.perf_counter() - t
    return n / dt, n, st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="r1", choices=["r1", "r1b", "z1"])
    ap.add_argument("--trajectories", type=int, default=0, help="trajectories per GPU per step (default: BASELINE size)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pkg = graft.load_package()
    n_per = args.trajectories or {"r1": 1_000_000, "r1b": 1_000_000, "z1": 100_000}[args.workload]
    cfg, natoms, desc = workload(pkg, args.workload, n_per)
    metric = "master-eq trajectories/sec"
    unit = "trajectories/s"

    if args.impl == "reference":
        if rank != 0:
            return
        oracle = graft.load_oracle()
        oracle.build()
        threads = host_threads()   # torchrun exports OMP_NUM_THREADS=1: the reference arm uses every core the process may run on
        vals = []
        for i in range(args.warmup + args.steps):
            v, n_s, _ = cpu_reference(pkg, oracle, cfg, natoms, max(2.0, args.cpu_seconds / 2), threads)
            if i >= args.warmup:
                vals.append((v, n_s))
        value = float(np.mean([v for v, _ in vals]))
        n_s = vals[-1][1]
        line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * n_s / value, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "trajectories_per_step": n_s},
                "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port",
                                 "sample": f"{n_s} trajectories of the same workload per step (C oracle, OpenMP, DP5 rtol 1e-6/atol 1e-8)"},
                "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    abi = pkg._abi
    ctx = pkg._lib.default_context(local)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    model = pkg.model_from_config(cfg, two_atom=(natoms == 2))
    d = 5 if natoms == 1 else 25
    nt = len(cfg.tspan)
    buf = torch.zeros((2, nt, d, d, 2), dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    n_total = n_per * world
    run = ctx.rydberg1_run if natoms == 1 else ctx.rydberg2_run
    kw = dict(f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes)
    ode_kwargs = {}
    if natoms == 2:
        # OrdinaryDiffEq's default maxiters=1e5 is hit by the closest atom pairs of a 1e5-trajectory ensemble (V ~ R^-6 sets the step):
        # the caller raises it through ode_kwargs exactly as with the reference (cz_model.jl:163 forwards them)
        ode_kwargs = dict(maxiters=10 ** 7)
        kw["ode"] = abi.ode_opts(maxiters=10 ** 7)

    def step(i):
        flush.zero_()
        st = run(model, cfg.tspan, cfg.ψ0, n_per, traj_offset=rank * n_per, n_total=n_total, seed=1000 + i,
                 out_flags=abi.CA_OUT_SUMS | abi.CA_OUT_DEVICE, dev_out=(buf[0].data_ptr(), buf[1].data_ptr()), **kw)[2]
        if world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    peak_tf, peak_mhz = ctx.measure_fp64_peak()
    for i in range(args.warmup):
        step(i)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        st = step(args.warmup + i)
        launches += st["n_launches"] + (1 if world > 1 else 0)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clk = clocks.stop() if rank == 0 else None
    value = n_total * args.steps / (ms * 1e-3)

    # roofline of the dominant kernel: one extra step through the host-output path to read the device counters
    ctx.reset_stream()
    torch.cuda.synchronize(dev)
    _, _, st = run(model, cfg.tspan, cfg.ψ0, n_per, traj_offset=rank * n_per, n_total=n_total, seed=999, **kw)
    flops = algorithmic_flops(cfg, natoms, st, n_per)
    ach = flops / (st["kernel_ms"] * 1e-3) / 1e12
    peak_tf2, _ = ctx.measure_fp64_peak()

    # e2e through the public API (host buffers in/out, sampling on device, D2H of ρ and ρ2 inside the timed region)
    sim = pkg.simulation if natoms == 1 else pkg.simulation_czlp
    cfg_e = cfg.copy()
    cfg_e.n_samples = n_total
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    for i in range(e2e_steps):
        rho, rho2 = sim(cfg_e, seed=2000 + i, device=local, distributed=(world > 1), **ode_kwargs)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = n_total * e2e_steps / float(e2e_t.item())
    h2d = 8 * (nt + 10 + 3 * len(cfg.f)) + 2048  # tspan, ψ0, f/amplitudes, model POD
    d2h = 2 * nt * d * d * 16 + 64

    if rank == 0:
        cpu = None
        if not args.no_cpu:
            oracle = graft.load_oracle()
            oracle.build()
            threads = host_threads()
            v, n_s, _ = cpu_reference(pkg, oracle, cfg, natoms, args.cpu_seconds, threads)
            cpu = {"value": v, "unit": unit, "cores": threads, "kind": "port",
                   "sample": f"{n_s} trajectories of the same workload (C oracle, OpenMP over trajectories, DP5 rtol 1e-6/atol 1e-8)"}
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "trajectories_per_gpu_per_step": n_per, "nt": nt, "ode": "DP5 adaptive rtol 1e-6 atol 1e-8 dense output" + (", maxiters=1e7" if natoms == 2 else ""),
                       "l2": "flushed between steps (256 MiB write); the path is FP64-compute-bound with KB-sized inputs",
                       "parallelism": f"trajectory shards x{world}, one NCCL all-reduce of the rho sums"},
            "roofline": {"bound": "fp64", "achieved": ach, "peak": max(peak_tf, peak_tf2), "unit": "TFLOP/s",
                         "frac": ach / max(peak_tf, peak_tf2),
                         "traffic": NCU_DRAM_BYTES_PER_TRAJ[natoms] * n_per,
                         "traffic_note": "bytes per launch = ncu dram bytes per trajectory (captured launch, profiles/r1) x trajectories of this launch",
                         "kernel": "k_lindblad1" if natoms == 1 else "k_lindblad2", "kernel_ms": st["kernel_ms"],
                         "flop_model": f"n_rhs*{F_RHS_1 if natoms == 1 else F_RHS_2:.0f} + noise {F_NOISE_PER}/(summed node,freq) + obs (SURVEY 8d canonical, DESIGN.md)",
                         "n_rhs_per_traj": st["n_rhs"] / max(1, n_per),
                         "peak_source": "measured DFMA chain (ca_measure_fp64_peak) in this run; MEASURED_PEAKS.json has no FP64 figure"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": clk,
            "stats": {k: st[k] for k in ("n_rhs", "n_accept", "n_reject", "n_sample_attempts", "kernel_ms", "total_ms")},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
