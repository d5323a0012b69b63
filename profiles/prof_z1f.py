"""Full-form two-atom kernel (k_lindblad2_full) timing: the Z1 workload with a psi0 that has |l> components (forces the fallback)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import numpy as np
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 444
tend = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
_, cz = pkg.get_default_configs()
cz.laser_noise = True
ctx = pkg._lib.default_context(0)
model = pkg.model_from_config(cz, two_atom=True)
psi = np.array(cz.ψ0, dtype=complex)
psi[24] = 1e-3   # a tiny |ll> admixture
psi /= np.linalg.norm(psi)
ts = np.array([0.0, tend])
for r in range(2):
    _, _, st = ctx.rydberg2_run(model, ts, psi, n, seed=r, f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes,
                                ode=pkg._abi.ode_opts(maxiters=10**7))
    print("z1-full n", n, "tend", tend, "kernel_ms %.1f" % st["kernel_ms"], "RHS/s %.3g" % (st["n_rhs"] / st["kernel_ms"] * 1e3), flush=True)
