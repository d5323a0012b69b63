"""Two-atom kernel timing: n trajectories of the Z1 workload on a shortened window (same per-RHS cost)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import numpy as np
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
tend = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
_, cz = pkg.get_default_configs()
cz.laser_noise = True
ctx = pkg._lib.default_context(0)
model = pkg.model_from_config(cz, two_atom=True)
ts = np.array([0.0, tend])
for r in range(2):
    _, _, st = ctx.rydberg2_run(model, ts, cz.ψ0, n, seed=r, f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes,
                                ode=pkg._abi.ode_opts(maxiters=10**7))
    rhs = st["n_rhs"]
    print("z1 n", n, "tend", tend, "kernel_ms %.1f" % st["kernel_ms"], "RHS/s %.3g" % (rhs / st["kernel_ms"] * 1e3), "full-window traj/s %.0f" % (rhs / st["kernel_ms"] * 1e3 / 67182), flush=True)
