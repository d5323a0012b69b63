"""R2 workload timing (BASELINE config 3): one-atom ensemble with uniform red and flat-top HG / LG blue beams
(simulation_blue_intens semantics), on-device sampling.   python profiles/prof_r2.py [n] [HG|LG] [order]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import numpy as np
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256
kind = sys.argv[2] if len(sys.argv) > 2 else "HG"
order = int(sys.argv[3]) if len(sys.argv) > 3 else 4
Γ = 2 * np.pi * 6.0
Δ0, Ω = 2 * np.pi * 1500.0, 2 * np.pi * 96.0
wb, zb = 2.0, pkg.w0_to_z0(2.0, 0.475)
T0 = pkg.T_twophoton(Ω, Ω, Δ0)
tspan = (T0 / 50) * np.arange(0, 301)
s = order // 2 + 1
blue = [Ω, wb / np.sqrt(s), zb / s, "simp_flattop_HG", order, order, 1.0] if kind == "HG" else [Ω, wb / np.sqrt(s), zb / s, "simp_flattop_LG", order, 0]
m = pkg.blue_intens_model([86.9091835, 70.0], [1000.0, 1.1, pkg.w0_to_z0(1.1, 0.852, 1.3)], [Ω, 100.0, pkg.w0_to_z0(100.0, 0.795)], blue,
                          [Δ0, pkg.δ_twophoton(Ω, Ω, Δ0)], [Γ / 4, 3 * Γ / 4])
ctx = pkg._lib.default_context(0)
for r in range(2):
    _, _, st = ctx.rydberg1_run(m, tspan, pkg.ket_1, n, seed=r)
    print("r2", kind, order, "n", n, "kernel_ms %.2f" % st["kernel_ms"], "traj/s %.0f" % (n / st["kernel_ms"] * 1e3), "rhs/traj %.0f" % (st["n_rhs"] / n), flush=True)
