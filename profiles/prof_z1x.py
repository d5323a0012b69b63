"""K4 cost-breakdown experiments: fixed-step runs (no controller) of n trajectories."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import numpy as np
pkg = g.load_package()
n = int(sys.argv[1]); tend = float(sys.argv[2]); fixed = int(sys.argv[3])
_, cz = pkg.get_default_configs()
cz.laser_noise = True
ctx = pkg._lib.default_context(0)
model = pkg.model_from_config(cz, two_atom=True)
ts = np.array([0.0, tend])
ode = pkg._abi.ode_opts(dt=5e-5, adaptive=False) if fixed else pkg._abi.ode_opts(maxiters=10**7)
for r in range(2):
    try:
        _, _, st = ctx.rydberg2_run(model, ts, cz.ψ0, n, seed=r, f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes, ode=ode)
    except Exception as e:
        st = pkg.last_stats() if hasattr(pkg, "last_stats") else {}
        print("error", str(e)[:60]); continue
    print("n", n, "fixed" if fixed else "adaptive", "kernel_ms %.1f" % st["kernel_ms"], "RHS/s %.3g" % (st["n_rhs"] / st["kernel_ms"] * 1e3), "cycles/RHS/warp %.0f" % (1184 * 1.965e9 / (st["n_rhs"] / st["kernel_ms"] * 1e3)), flush=True)
