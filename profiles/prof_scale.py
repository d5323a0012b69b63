"""Kernel time vs ensemble size (multiples of one persistent wave)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
cfg, _ = pkg.get_default_configs()
cfg.laser_noise = True
ctx = pkg._lib.default_context(0)
model = pkg.model_from_config(cfg)
wave = 148 * 2 * 128
for mult in (1, 3, 6, 12, 24, 27):
    n = wave * mult
    _, _, st = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, seed=1, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes)
    print("waves", mult, "n", n, "kernel_ms %.1f" % st["kernel_ms"], "ms/wave %.2f" % (st["kernel_ms"] / mult), "traj/s %.0f" % (n / st["kernel_ms"] * 1e3), flush=True)
