import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import numpy as np
pkg = g.load_package(); oracle = g.load_oracle()
_, cz = pkg.get_default_configs()
cz.laser_noise = True
ctx = pkg._lib.default_context(0)
model = pkg.model_from_config(cz, two_atom=True)
kw = dict(f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes)
for n, tend, mi in [(8, cz.tspan[-1], 10**6), (8, 0.1, 10**6), (1200, 0.024, 10**6), (2400, 0.024, 10**6), (8, cz.tspan[-1], 10**5)]:
    ts = np.array([0.0, tend])
    try:
        rho, rho2, st = ctx.rydberg2_run(model, ts, cz.ψ0, n, seed=0, ode=pkg._abi.ode_opts(maxiters=mi), **kw)
        o, o2, ost = oracle.rydberg2_run(model, ts, cz.ψ0, min(n, 8), seed=0, **kw)
        print(n, tend, "rhs/traj", st["n_rhs"] / n, "acc", st["n_accept"] / n, "rej", st["n_reject"] / n, "oracle rhs/traj", ost["n_rhs"] / min(n, 8), "kernel_ms", st["kernel_ms"],
              "err(n<=8)", np.abs(rho - o).max() if n <= 8 else None, flush=True)
    except Exception as e:
        print(n, tend, "EXC", e, flush=True)
