"""Profiling driver: one call of the one-atom ensemble (R1 workload) with n trajectories; used under ncu.
    python profiles/prof_r1.py [n] [reps] [workload]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import numpy as np
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = sys.argv[3] if len(sys.argv) > 3 else "r1"
cfg, cz = pkg.get_default_configs()
ctx = pkg._lib.default_context(0)
if wl == "z1":
    cz.laser_noise = True
    model = pkg.model_from_config(cz, two_atom=True)
    run, c = ctx.rydberg2_run, cz
else:
    cfg.laser_noise = True
    model = pkg.model_from_config(cfg)
    run, c = ctx.rydberg1_run, cfg
for r in range(reps):
    _, _, st = run(model, c.tspan, c.ψ0, n, seed=r, f=c.f, amp_red=c.red_laser_phase_amplitudes, amp_blue=c.blue_laser_phase_amplitudes)
    print(wl, "n", n, "kernel_ms %.2f" % st["kernel_ms"], "traj/s %.0f" % (n / st["kernel_ms"] * 1e3), "rhs/traj %.0f" % (st["n_rhs"] / n), flush=True)
