"""Times the R1 kernel with and without laser noise (cost split of the fused phases)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 56832
cfg, _ = pkg.get_default_configs()
ctx = pkg._lib.default_context(0)
for noise in (True, False, True, False):
    cfg.laser_noise = noise
    model = pkg.model_from_config(cfg)
    _, _, st = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, seed=1, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes)
    print("noise", noise, "kernel_ms %.2f" % st["kernel_ms"], "noise_ms %.2f" % st["noise_ms"], "traj/s %.0f" % (n / st["kernel_ms"] * 1e3), flush=True)
# short window: noise generation dominates (R1b)
import numpy as np
cfg.laser_noise = True
model = pkg.model_from_config(cfg)
ts = np.array([0.0, cfg.tspan[10]])
_, _, st = ctx.rydberg1_run(model, ts, cfg.ψ0, n, seed=1, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes)
print("R1b pi pulse kernel_ms %.2f" % st["kernel_ms"], "traj/s %.0f" % (n / st["kernel_ms"] * 1e3), "rhs/traj", st["n_rhs"] / n)
