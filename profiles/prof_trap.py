"""One-atom kernel with the atom oscillating in the trap (free_motion=false: the 'Laser noise' configuration of src/fidelity.jl:33-41)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 113664
cfg, _ = pkg.get_default_configs()
cfg.laser_noise = True
cfg.free_motion = False
ctx = pkg._lib.default_context(0)
model = pkg.model_from_config(cfg)
for r in range(2):
    _, _, st = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, seed=r, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes)
    print("trapped r1 n", n, "kernel_ms %.1f" % st["kernel_ms"], "traj/s %.0f" % (n / st["kernel_ms"] * 1e3), "rhs/traj %.0f" % (st["n_rhs"] / n), flush=True)
