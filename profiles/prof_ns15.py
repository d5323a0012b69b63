import sys, os
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g
import numpy as np
pkg = g.load_package()
n = int(sys.argv[1])
cfg, _ = pkg.get_default_configs(); cfg.laser_noise = True
ctx = pkg._lib.default_context(0)
model = pkg.model_from_config(cfg)
psi = (pkg.ket_0 + pkg.ket_1) / np.sqrt(2)
_, _, st = ctx.rydberg1_run(model, cfg.tspan, psi, n, seed=1, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes)
print("ns15 n", n, "kernel_ms %.2f" % st["kernel_ms"], "traj/s %.0f" % (n / st["kernel_ms"] * 1e3))
