"""S1 workload (BASELINE config 5) at 1/100 scale on one GPU: detuning x temperature x pulse-time grid of the 1-atom Rabi model
through ca_rydberg1_sweep (one launch for all points).   python profiles/prof_sweep.py [n_points] [traj_per_point]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import numpy as np
pkg = g.load_package()
n_points = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_pp = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cfg, _ = pkg.get_default_configs()
cfg.laser_noise = True
T0 = cfg.tspan[20]
rng = np.random.default_rng(0)
nd, nT, nt = 10, 10, max(1, n_points // 100)
δs = cfg.detuning_params[1] + 2 * np.pi * np.linspace(-2, 2, nd)
Ts = np.linspace(10.0, 100.0, nT)
ts = np.linspace(2 * T0 / nt, 2 * T0, nt)
models, psis, tends = [], [], []
for δ in δs:
    for T in Ts:
        for te in ts:
            c = cfg.copy()
            c.detuning_params[1] = float(δ)
            c.atom_params[1] = float(T)
            models.append(pkg.model_from_config(c))
            psis.append(pkg.ket_1)
            tends.append(float(te))
models, psis, tends = models[:n_points], psis[:n_points], tends[:n_points]
ctx = pkg._lib.default_context(0)
kw = dict(f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes)
for r in range(2):
    rho, rho2, st = ctx.rydberg1_sweep(models, psis, tends, n_pp, seed=r, **kw)
    n = len(models) * n_pp
    print("sweep points", len(models), "traj/point", n_pp, "kernel_ms %.1f" % st["kernel_ms"], "traj/s %.0f" % (n / st["kernel_ms"] * 1e3),
          "rhs/traj %.0f" % (st["n_rhs"] / n), "P_r range %.3f..%.3f" % (rho[:, 2, 2].real.min(), rho[:, 2, 2].real.max()), flush=True)
