#!/usr/bin/env python
"""bench.py -- benchmark of the Monte-Carlo master-equation hot path (BASELINE.json metric) on every BASELINE config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload r1|r1b|r2|z1|c1|s1] [--scaling weak|strong]
                    [--trajectories N] [--impl ours|reference]

A "step" is one pass of the hot path over one ensemble of synthetic trajectories.  Workloads:
  r1  (default) BASELINE configs[1]: single-atom two-photon Rydberg Rabi oscillation with Doppler/position noise, laser phase
      noise and intermediate decay, 10^6 trajectories (`get_default_configs()[1]`, n_samples = 10^6, laser_noise = true)
  r1b the same model on a pi pulse `tspan = [0, T0/2]` (what the fidelity callers run)
  r2  configs[2]: `simulation_blue_intens` semantics, uniform red + flat-top HG (n = m = 4; `--beam lg6` for LG l = 6) blue, 10^6 trajectories
  z1  configs[3]: two-atom CZ (Levine-Pichler), 10^5 trajectories, maxiters = 1e7
  c1  configs[0]: release-and-recapture curve, 10^4 atoms, t = 0..50 us (metric: atoms/s)
  s1  configs[4]: detuning x temperature x pulse-time sweep, 10^4 grid points x 10^4 trajectories = 10^8, sharded by grid point
With N > 1 (torchrun, one rank per GPU): `--scaling weak` (default) gives every rank the full BASELINE ensemble of a N-times larger
job, `--scaling strong` splits ONE BASELINE ensemble over the ranks; the un-normalised sums are combined by one NCCL all-reduce (s1:
no collective on the data path, the per-point results are gathered).

Prints ONE JSON line (rank 0).  `value` = trajectories/s with everything resident on the device (results stay in HBM); `e2e` = the
same metric through the public API call with host buffers in and out.  `--impl reference` times the reference's CPU path (the C
oracle port with OpenMP over trajectories: Julia is not installable here) on the host cores, same metric and `config`, on a bounded
sample.
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

F_RHS_1 = 340.0    # canonical flop per one-atom RHS (frozen): two Gaussian beams (86 of the 340)
F_RHS_2 = 7500.0   # canonical flop per two-atom RHS
F_NOISE_PER = 3.0  # flop per (node, frequency) actually executed by the Reinsch recurrence (SURVEY counts 8)
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e-3   # 148 SMs x 64 FMA/clk x 2 flop x 1.965 GHz = 37.2


def f_rhs_r2(kind, order):
    """SURVEY 8d mode-term adders: the 340 flop of the Gaussian pair minus its two beam evaluations (86), plus a uniform red beam
    (10: one sincos-free complex constant times the noise phase), the Gaussian envelope of the blue beam (43) and the flat-top mode sum:
    HG order n: (n/2+1)^2 terms x 30 flop + two Hermite recurrences (4n each); LG order l: (l/2+1) terms x 25 flop."""
    base = F_RHS_1 - 86.0 + 10.0 + 43.0
    return base + ((order // 2 + 1) ** 2 * 30.0 + 8.0 * order if kind == "HG" else (order // 2 + 1) * 25.0)


class Workload:
    """One BASELINE config: how to run a step on the device, through the public API, and on the CPU oracle."""

    def __init__(self, pkg, name, args, world):
        self.pkg, self.name, self.world = pkg, name, world
        abi = pkg._abi
        cfg, cz = pkg.get_default_configs()
        self.natoms, self.ode, self.ode_kwargs, self.metric, self.unit = 1, abi.ode_opts(), {}, "master-eq trajectories/sec", "trajectories/s"
        self.noise = True
        default_n = {"r1": 10 ** 6, "r1b": 10 ** 6, "r2": 10 ** 6, "z1": 10 ** 5, "c1": 10 ** 4, "s1": 10 ** 8}[name]
        self.n_config = args.trajectories or default_n   # the ensemble BASELINE.json names (per job under strong, per GPU under weak scaling)
        if name in ("r1", "r1b", "s1"):
            cfg.laser_noise = True
            if name == "r1b":
                cfg.tspan = np.array([0.0, cfg.tspan[20] / 2])
            self.cfg, self.model, self.tspan, self.psi0 = cfg, pkg.model_from_config(cfg), cfg.tspan, cfg.ψ0
            self.desc = {"r1": "R1: 1-atom two-photon Rydberg Rabi, default.jl cfg, laser_noise=true, tspan=0:T0/20:5T0 (nt=101)",
                         "r1b": "R1b: same model, pi pulse tspan=[0,T0/2]",
                         "s1": "S1: delta x T x pulse-time sweep of the 1-atom Rabi model (R1b points), 100 x 10 x 10 grid points"}[name]
            self.f_rhs = F_RHS_1
            if name == "s1":
                self.n_pp = args.traj_per_point
                self.n_points = max(1, self.n_config // self.n_pp)
                T0 = cfg.tspan[20]
                nd = max(1, self.n_points // 100)
                δs = cfg.detuning_params[1] + 2 * np.pi * np.linspace(-2, 2, nd)
                self.grid = [(δ, T, te) for δ in δs for T in np.linspace(10.0, 100.0, 10) for te in np.linspace(2 * T0 / 10, 2 * T0, 10)]
                self.n_points = len(self.grid)
                self.n_config = self.n_points * self.n_pp
        elif name == "r2":
            kind, order = ("HG", 4) if args.beam == "hg4" else ("LG", int(args.beam[2:])) if args.beam.startswith("lg") else ("HG", int(args.beam[2:]))
            Γ, Δ0, Ω = 2 * np.pi * 6.0, 2 * np.pi * 1500.0, 2 * np.pi * 96.0
            T0 = pkg.T_twophoton(Ω, Ω, Δ0)
            s = order // 2 + 1
            wb, zb = 2.0, pkg.w0_to_z0(2.0, 0.475)
            blue = [Ω, wb / np.sqrt(s), zb / s, "simp_flattop_" + kind, order, order if kind == "HG" else 0] + ([1.0] if kind == "HG" else [])
            self.r2 = dict(tspan=(T0 / 50) * np.arange(0, 301), ψ0=pkg.ket_1, atom_params=[86.9091835, 70.0],
                           trap_params=[1000.0, 1.1, pkg.w0_to_z0(1.1, 0.852, 1.3)], red_laser_params=[Ω, 100.0, pkg.w0_to_z0(100.0, 0.795)],
                           blue_laser_params=blue, detuning_params=[Δ0, pkg.δ_twophoton(Ω, Ω, Δ0)], decay_params=[Γ / 4, 3 * Γ / 4])
            a = self.r2
            self.model = pkg.blue_intens_model(a["atom_params"], a["trap_params"], a["red_laser_params"], a["blue_laser_params"],
                                               a["detuning_params"], a["decay_params"])
            self.tspan, self.psi0, self.noise, self.cfg = a["tspan"], a["ψ0"], False, None
            self.desc = f"R2: simulation_blue_intens, uniform red + flat-top {kind} order {order} blue, params/default.jl, tspan=0:T0/50:6T0 (nt=301)"
            # roofline: the Gaussian-pair canonical count (340) is used for `achieved` so that the fraction is never inflated by mode sums
            # the anchored kernel does not evaluate per RHS; the SURVEY 8d count with the mode-term adders is reported next to it
            self.f_rhs, self.f_rhs_modes = F_RHS_1, f_rhs_r2(kind, order)
        elif name == "z1":
            cz.laser_noise = True
            self.cfg, self.model, self.tspan, self.psi0, self.natoms = cz, pkg.model_from_config(cz, two_atom=True), cz.tspan, cz.ψ0, 2
            # OrdinaryDiffEq's default maxiters=1e5 is hit by the closest atom pairs of a 1e5-trajectory ensemble (V ~ R^-6 sets the
            # step): the caller raises it through ode_kwargs exactly as with the reference (cz_model.jl:163 forwards them)
            self.ode, self.ode_kwargs = abi.ode_opts(maxiters=10 ** 7), dict(maxiters=10 ** 7)
            self.desc = "Z1: 2-atom CZ (Levine-Pichler), default.jl cfg_czlp, laser_noise=true, tspan=[0,2tau]"
            self.f_rhs = F_RHS_2
        elif name == "c1":
            self.trap, self.atom, self.tspan = [1000.0, 1.0, pkg.w0_to_z0(1.0, 0.852)], [86.9091835, 50.0], np.arange(0.0, 51.0)
            self.desc = "C1: release-and-recapture, Rb-87 T=50 uK, U0=1 mK, w0=1 um, tspan=0:1:50 us (nt=51), harmonic sampler, eps=1e-3"
            self.metric, self.unit, self.noise, self.cfg = "release-recapture atoms/sec", "atoms/s", False, None
        self.nt = 2 if name == "s1" else len(self.tspan)   # (a sweep point is tspan = [0, t_end]; only rho(t_end) is kept)
        self.d = 5 if self.natoms == 1 else 25

    # ---- ensemble sizes -------------------------------------------------------------------------------------------------------
    def shard(self, rank, scaling):
        """(offset, count, n_total) of this rank's trajectories (s1: grid points)."""
        n = self.n_points if self.name == "s1" else self.n_config
        if self.name == "s1" or scaling == "strong":
            lo, hi = n * rank // self.world, n * (rank + 1) // self.world
            return lo, hi - lo, n
        return rank * n, n, n * self.world

    def config(self, scaling):
        """Workload-defining keys only: identical for the GPU arm and the reference arm."""
        c = {"workload": self.desc, "trajectories": self.n_config, "nt": self.nt, "scaling": scaling,
             "ode": "DP5 adaptive rtol 1e-6 atol 1e-8 dense output" + (", maxiters=1e7" if self.natoms == 2 else "")}
        if self.name == "c1":
            c["ode"] = "none (ballistic flight + energy test)"
        if self.name == "s1":
            c["grid_points"], c["trajectories_per_point"] = self.n_points, self.n_pp
        return c

    def noise_kw(self):
        c = self.cfg
        return dict(f=c.f, amp_red=c.red_laser_phase_amplitudes, amp_blue=c.blue_laser_phase_amplitudes) if self.noise else {}

    # ---- device-resident step ---------------------------------------------------------------------------------------------
    def s1_points(self, lo, cnt):
        models, tends = [], []
        for δ, T, te in self.grid[lo:lo + cnt]:
            c = self.cfg.copy()
            c.detuning_params[1] = float(δ)
            c.atom_params[1] = float(T)
            models.append(self.pkg.model_from_config(c))
            tends.append(float(te))
        return models, tends

    def algorithmic_flops(self, st, n):
        if self.name == "c1":
            return 25.0 * self.nt * n
        traces = 2 * self.natoms if self.noise else 0
        c, pw = noise_plan(self.cfg) if (self.natoms == 1 and self.noise) else (1, 8)
        nodes_summed = 1001 if c == 1 else (1000 + c - 1) // c + pw
        nf = len(self.cfg.f) if self.noise else 0
        f_noise = traces * (nodes_summed * nf * F_NOISE_PER + (0 if c == 1 else 1001 * 2 * pw)) * n   # (one FMA per window node and table entry)
        nt_out = 1 if self.name == "s1" else self.nt
        f_obs = nt_out * (8 * self.d ** 3 + 4 * self.d ** 2) * n
        return st["n_rhs"] * self.f_rhs + f_noise + f_obs


def noise_plan(cfg, n_nodes=1001, chunk=64):
    """(c, p): the interpolation plan the one-atom kernel picks for its phase-noise tables (choose_noise_coarse, ca_device.cuh):
    every c-th node is summed over the frequencies, the others follow from p-point Lagrange interpolation."""
    dtn = cfg.tspan[-1] / (n_nodes - 1.0)
    nf = len(cfg.f)
    best, best_cost = (1, 8), float("inf")
    for p in (8, 16):
        for c in ((32, 16, 8, 4, 2) if p == 8 else (32, 16, 8)):
            n_c = (n_nodes - 1 + c - 1) // c + p
            if p == 8 and n_c * c > 2 * n_nodes:
                continue
            b = max(float(np.sum(np.abs(a) * (2 * np.pi * np.abs(cfg.f) * c * dtn) ** p))
                    for a in (cfg.red_laser_phase_amplitudes, cfg.blue_laser_phase_amplitudes))
            if not b * (2.9966e-6 if p == 16 else 43.07 / 40320.0) <= 1e-14:
                continue
            cost = ((n_c + chunk - 1) // chunk) * nf * 45.0 + n_c * nf * 3.0 + n_nodes * p
            if cost < best_cost:
                best, best_cost = (c, p), cost
    return best


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


# ---- CPU arms (the only place bench.py executes oracle/) -----------------------------------------------------------------------
def cpu_reference(wl, oracle, budget_s, threads, dense=False, rhs_per_traj=None):
    """Reference CPU path on a bounded sample of the same workload: the C oracle port, OpenMP over trajectories (`dense=False`), or
    its single-thread dense-`mul!` emulation of the reference's own arithmetic (`dense=True`, SURVEY 8d: 32 dense 5x5 / 66 dense
    25x25 complex products per RHS; timed on a shortened window as RHS/s and converted with the RHS count of a full trajectory)."""
    if wl.name == "c1":
        n = 10 ** 4
        t = time.perf_counter()
        reps = 0
        while time.perf_counter() - t < min(budget_s, 5.0):
            oracle.release_recapture(wl.trap, wl.atom, wl.tspan, n, seed=5 + reps)
            reps += 1
        return n * reps / (time.perf_counter() - t), n * reps, f"{reps} x 10^4 atoms, single thread (C oracle)"
    run = oracle.rydberg1_run if wl.natoms == 1 else oracle.rydberg2_run
    model, tspan, psi0, kw = wl.model, wl.tspan, wl.psi0, dict(wl.noise_kw(), ode=wl.ode)
    if wl.name == "s1":   # one grid point of the sweep: the middle detuning and temperature at the mean pulse time
        models, _ = wl.s1_points(len(wl.grid) // 2, 1)
        model, tspan = models[0], np.array([0.0, float(np.mean([g[2] for g in wl.grid]))])
    oracle.set_dense_emulation(dense)
    try:
        nthreads = 1 if dense else threads
        if dense:
            tspan = np.array([tspan[0], tspan[0] + (tspan[-1] - tspan[0]) / (40 if wl.natoms == 2 else 4)])
        n0 = max(nthreads, 1 if dense else 8)
        t = time.perf_counter()
        run(model, tspan, psi0, n0, nthreads=nthreads, **kw)
        dt0 = time.perf_counter() - t
        n = int(max(n0, min(20000, budget_s / max(dt0, 1e-3) * n0)))
        n -= n % nthreads
        t = time.perf_counter()
        _, _, st = run(model, tspan, psi0, n, nthreads=nthreads, **kw)
        dt = time.perf_counter() - t
        if dense:
            rhs_s = st["n_rhs"] / dt
            return rhs_s / rhs_per_traj, n, (f"{n} trajectories on a 1/{40 if wl.natoms == 2 else 4} window: {rhs_s:.3g} RHS/s with 32 dense 5x5 / 66 dense 25x25 "
                                            f"products per RHS like QuantumOptics' mul!, single thread like the reference loop, / {rhs_per_traj:.0f} RHS per full trajectory")
        return n / dt, n, f"{n} trajectories of the same workload (C oracle, OpenMP over trajectories, DP5 rtol 1e-6/atol 1e-8)"
    finally:
        oracle.set_dense_emulation(False)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="r1", choices=["r1", "r1b", "r2", "z1", "c1", "s1"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--beam", default="hg4", help="r2: hgN or lgN")
    ap.add_argument("--trajectories", type=int, default=0, help="ensemble size (default: the BASELINE.json size of the workload)")
    ap.add_argument("--traj-per-point", type=int, default=10 ** 4, help="s1: trajectories per grid point")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pkg = graft.load_package()
    wl = Workload(pkg, args.workload, args, world if args.impl == "ours" else max(1, args.gpus))
    scaling = "strong" if args.workload == "s1" else args.scaling
    metric, unit = wl.metric, wl.unit

    if args.impl == "reference":
        if rank != 0:
            return
        oracle = graft.load_oracle()
        oracle.build()
        threads = host_threads()   # torchrun exports OMP_NUM_THREADS=1: the reference arm uses every core the process may run on
        vals = []
        for i in range(args.warmup + args.steps):
            v, n_s, sample = cpu_reference(wl, oracle, max(2.0, args.cpu_seconds / 2), threads)
            if i >= args.warmup:
                vals.append((v, n_s))
        value = float(np.mean([v for v, _ in vals]))
        n_s = vals[-1][1]
        line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * n_s / value, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": wl.config(scaling),
                "cpu_baseline": {"value": value, "unit": unit, "cores": 1 if args.workload == "c1" else threads, "kind": "port", "sample": sample + " per step"},
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
    nt, d = wl.nt, wl.d
    lo, cnt, n_total = wl.shard(rank, scaling)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    DEV = abi.CA_OUT_SUMS | abi.CA_OUT_DEVICE
    n_out = 2 * nt * d * d * 2

    if wl.name == "c1":
        buf = torch.zeros(nt + 8, dtype=torch.float64, device=dev)

        def step(i):
            flush.zero_()
            st = ctx.release_recapture(wl.trap, wl.atom, wl.tspan, cnt, traj_offset=lo, n_total=n_total, seed=1000 + i, out_flags=DEV, dev_out=buf.data_ptr())[2]
            if world > 1:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            return st
    elif wl.name == "s1":
        models, tends = wl.s1_points(lo, cnt)
        marr = (abi.ca_model * cnt)(*models)   # marshalled once: the sweep grid is the job's input, like tspan
        psis = [pkg.ket_1] * cnt

        sbuf = torch.zeros(2 * cnt * 50 + 8, dtype=torch.float64, device=dev)

        def step(i):
            flush.zero_()
            return ctx.rydberg1_sweep(marr, psis, tends, wl.n_pp, traj_offset=lo * wl.n_pp, n_total=wl.n_pp, seed=1000 + i, out_flags=abi.CA_OUT_DEVICE,
                                      dev_out=(sbuf.data_ptr(), sbuf.data_ptr() + 8 * 50 * cnt), **wl.noise_kw())[2]
    else:
        buf = torch.zeros(n_out + 8, dtype=torch.float64, device=dev)
        run = ctx.rydberg1_run if wl.natoms == 1 else ctx.rydberg2_run
        half = nt * d * d * 2

        def step(i):
            flush.zero_()
            st = run(wl.model, wl.tspan, wl.psi0, cnt, traj_offset=lo, n_total=n_total, seed=1000 + i, out_flags=DEV,
                     dev_out=(buf.data_ptr(), buf.data_ptr() + 8 * half), ode=wl.ode, **wl.noise_kw())[2]
            ctx.stats_to_device(buf.data_ptr() + 8 * n_out)   # the status counters ride in the same all-reduce
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
        launches += st["n_launches"] + (1 if wl.name not in ("c1", "s1") else 0)   # + k_export_counters (NCCL's kernel is not ours)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clk = clocks.stop() if rank == 0 else None
    n_job = wl.n_config if (scaling == "strong") else wl.n_config * world   # units all ranks processed per step
    value = n_job * args.steps / (ms * 1e-3)
    status = 0
    if wl.name not in ("c1", "s1"):
        tail = buf[n_out:].cpu().numpy()
        status = abi.CA_ERR_NONFINITE if tail[5] > 0 else (abi.CA_ERR_MAXITERS if tail[4] > 0 else 0)

    # roofline of the dominant kernel: one extra step through the host-output path of THIS rank's shard to read the device counters
    ctx.reset_stream()
    torch.cuda.synchronize(dev)
    if wl.name == "c1":
        _, _, st = ctx.release_recapture(wl.trap, wl.atom, wl.tspan, cnt, traj_offset=lo, n_total=n_total, seed=999)
        kernel = "k_release_recapture"
    elif wl.name == "s1":
        _, _, st = ctx.rydberg1_sweep(marr, psis, tends, wl.n_pp, traj_offset=lo * wl.n_pp, n_total=wl.n_pp, seed=999, **wl.noise_kw())
        kernel = "k_lindblad1"
    else:
        run = ctx.rydberg1_run if wl.natoms == 1 else ctx.rydberg2_run
        _, _, st = run(wl.model, wl.tspan, wl.psi0, cnt, traj_offset=lo, n_total=n_total, seed=999, ode=wl.ode, **wl.noise_kw())
        kernel = "k_lindblad1" if wl.natoms == 1 else "k_lindblad2"
    n_here = cnt * wl.n_pp if wl.name == "s1" else cnt
    flops = wl.algorithmic_flops(st, n_here)
    ach = flops / (max(st["kernel_ms"], 1e-6) * 1e-3) / 1e12
    peak_tf2, _ = ctx.measure_fp64_peak()
    peak = max(peak_tf, peak_tf2)

    # e2e through the public API (host buffers in/out, sampling on device, D2H of the results inside the timed region)
    e2e_steps = max(1, min(args.steps, 2))
    dist_flag = world > 1
    h2d = d2h = 0
    if wl.name == "r2":   # the reference API takes the thermal samples as a host array: drawn once, copied in on every call
        smp_host, _ = pkg.samples_generate(wl.r2["trap_params"], wl.r2["atom_params"], cnt, seed=77 + rank, device=local)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        if wl.name == "c1":
            probs, _ = pkg.release_recapture(wl.tspan, wl.trap, wl.atom, n_total if dist_flag else cnt, seed=2000 + i, device=local, distributed=dist_flag)
            h2d, d2h = 8 * (nt + 5), 8 * nt
        elif wl.name == "s1":
            srho, _, _ = ctx.rydberg1_sweep(models, psis, tends, wl.n_pp, traj_offset=lo * wl.n_pp, n_total=wl.n_pp, seed=2000 + i, **wl.noise_kw())   # marshals the grid
            h2d, d2h = cnt * (1248 + 10 * 8 + 16) + 8 * 3 * len(wl.cfg.f), cnt * 2 * 25 * 16
        elif wl.name == "r2":
            a = wl.r2
            rho, rho2 = pkg.simulation_blue_intens(a["tspan"], a["ψ0"], a["atom_params"], a["trap_params"], smp_host, a["red_laser_params"],
                                                   a["blue_laser_params"], a["detuning_params"], a["decay_params"], device=local)
            h2d, d2h = 8 * (nt + 10) + 2048 + 48 * cnt, 2 * nt * d * d * 16 + 64
        else:
            sim = pkg.simulation if wl.natoms == 1 else pkg.simulation_czlp
            cfg_e = wl.cfg.copy()
            cfg_e.n_samples = n_total if dist_flag else cnt
            rho, rho2 = sim(cfg_e, seed=2000 + i, device=local, distributed=dist_flag, **wl.ode_kwargs)
            h2d = 8 * (nt + 2 * d + 3 * len(wl.cfg.f)) + 2048   # tspan, ψ0, f/amplitudes, model POD
            d2h = 2 * nt * d * d * 16 + 64
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    n_e2e = n_total if (dist_flag and wl.name not in ("s1", "r2")) else (n_job if wl.name in ("s1",) else cnt * world)
    e2e_val = n_e2e * e2e_steps / float(e2e_t.item())

    # parity of the multi-rank route: N-rank result against rank 0 alone on 4096 trajectories of the same ensemble
    parity_multi = None
    if world > 1 and wl.name in ("r1", "r1b", "z1"):
        sim = pkg.simulation if wl.natoms == 1 else pkg.simulation_czlp
        c = wl.cfg.copy()
        c.n_samples = 4096 if wl.natoms == 1 else 256
        if wl.natoms == 1:
            c.tspan = np.linspace(0.0, 0.25, 6)
        else:
            c.tspan = np.array([0.0, 0.02])
        a, a2 = sim(c, seed=31, device=local, distributed=True, **wl.ode_kwargs)
        if rank == 0:
            b, b2 = sim(c, seed=31, device=local, **wl.ode_kwargs)
            parity_multi = {"max_abs_diff": float(max(np.abs(a - b).max(), np.abs(a2 - b2).max())), "trajectories": int(c.n_samples), "ranks": world}
        barrier()

    if rank == 0:
        cpu = cpu_dense = None
        if not args.no_cpu:
            oracle = graft.load_oracle()
            oracle.build()
            threads = host_threads()
            v, n_s, sample = cpu_reference(wl, oracle, args.cpu_seconds, threads)
            cpu = {"value": v, "unit": unit, "cores": 1 if wl.name == "c1" else threads, "kind": "port", "sample": sample}
            if wl.name != "c1":
                v, n_s, sample = cpu_reference(wl, oracle, args.cpu_seconds / 2, threads, dense=True, rhs_per_traj=st["n_rhs"] / max(1, n_here))
                cpu_dense = {"value": v, "unit": unit, "cores": 1, "kind": "port", "sample": sample}
        traffic, traffic_src = None, None
        tfile = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # written from the committed ncu --set full captures (profiles/ncu_summary.py)
        if os.path.exists(tfile):
            tj = json.load(open(tfile)).get(kernel if wl.name != "r2" else "k_lindblad1_r2")
            if tj:
                traffic, traffic_src = tj["dram_bytes_per_trajectory"] * n_here, tj["source"]
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": wl.config(scaling),
            "run": {"trajectories_this_rank_per_step": n_here, "trajectories_all_ranks_per_step": n_job,
                    "l2": "flushed between steps (256 MiB write); the path is FP64-compute-bound with KB-sized inputs",
                    "parallelism": f"trajectory shards x{world}, one NCCL all-reduce of [sum rho | sum rho.rho | status counters]" if wl.name != "s1"
                    else f"grid-point shards x{world}, no data-path collective", "device_status": status},
            "roofline": {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "frac_of_nominal": ach / FP64_NOMINAL_TFLOPS, "peak_nominal": FP64_NOMINAL_TFLOPS,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": kernel, "kernel_ms": st["kernel_ms"],
                         "flop_model": f"n_rhs*{wl.f_rhs:.0f} + noise {F_NOISE_PER}/(summed node,freq) + obs (SURVEY 8d canonical, DESIGN.md)" if wl.name != "c1"
                         else "25 flop per (atom, time point) (SURVEY 8d): latency-bound kernel, the fraction is not a target",
                         "n_rhs_per_traj": st["n_rhs"] / max(1, n_here), "exact_coefficient_steps": st.get("n_exact_steps", 0),
                         **({"achieved_with_mode_terms": ach * wl.f_rhs_modes / wl.f_rhs, "flop_per_rhs_with_mode_terms": wl.f_rhs_modes} if wl.name == "r2" else {}),
                         "peak_source": "measured DFMA chain (ca_measure_fp64_peak: 8 chains x 8 warps/SM) in this run; MEASURED_PEAKS.json has no FP64 figure; "
                                        "nominal = 148 SM x 64 FMA/clk x 1.965 GHz"},
            "cpu_baseline": cpu, "cpu_dense_emulation": cpu_dense,
            "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": launches, "clocks": clk, "parity_multi": parity_multi,
            "stats": {k: st[k] for k in ("n_rhs", "n_accept", "n_reject", "n_sample_attempts", "kernel_ms", "total_ms")},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
