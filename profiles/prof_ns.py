"""One-atom kernel with initial states outside the {1,r,p} block (spectator rows: NS = 15 / 21 variants of k_lindblad1), R1 model.
    python profiles/prof_ns.py [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import numpy as np
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256
cfg, _ = pkg.get_default_configs()
cfg.laser_noise = True
ctx = pkg._lib.default_context(0)
model = pkg.model_from_config(cfg)
states = {"ket_1 (NS=9)": pkg.ket_1, "(0+1)/sqrt2 (NS=15)": (pkg.ket_0 + pkg.ket_1) / np.sqrt(2),
          "general (NS=21)": (pkg.ket_0 + pkg.ket_1 + pkg.ket_l + 0.5j * pkg.ket_r) / np.sqrt(3.25)}
for name, psi in states.items():
    for r in range(2):
        _, _, st = ctx.rydberg1_run(model, cfg.tspan, psi, n, seed=r, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes)
    print(name, "n", n, "kernel_ms %.2f" % st["kernel_ms"], "traj/s %.0f" % (n / st["kernel_ms"] * 1e3), "rhs/traj %.0f" % (st["n_rhs"] / n), flush=True)
