import sys, os
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g
pkg = g.load_package()
cfg, cz = pkg.get_default_configs(); cfg.laser_noise = True
ctx = pkg._lib.default_context(0); model = pkg.model_from_config(cfg)
for n in (37888, 37888*4):
    for r in range(2):
        _, _, st = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, seed=r, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes)
        print("n", n, "kernel_ms %.2f noise_ms %.3f share %.4f traj/s %.0f" % (st["kernel_ms"], st["noise_ms"], st["noise_ms"]/st["kernel_ms"], n/st["kernel_ms"]*1e3), flush=True)
