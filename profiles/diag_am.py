"""Diagnostic: the 'Atom motion' point of get_cz_infidelity (T = 100 uK, no decay, no noise), GPU vs oracle trajectory by trajectory."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package(); oracle = g.load_oracle()
ctx = pkg._lib.default_context(0)
_, cz = pkg.get_default_configs()
pts = pkg.cz_infidelity_points(cz, n_samples=5)
name, c = pts[5]
m = pkg.model_from_config(c, two_atom=True)
ode = pkg._abi.ode_opts(maxiters=10**6)
for i in range(24):
    gi = 109 + i
    r, r2, st = ctx.rydberg2_run(m, [0.0, c.tspan[-1]], c.ψ0, 1, traj_offset=gi, seed=11, ode=ode)
    o, o2, ost = oracle.rydberg2_run(m, [0.0, c.tspan[-1]], c.ψ0, 1, traj_offset=gi, seed=11, ode=ode)
    d = np.abs(r - o)[-1]
    print(i, "gpu", st["n_accept"], st["n_reject"], "oracle", ost["n_accept"], ost["n_reject"], "max|d| %.2e diag %.2e" % (d.max(), np.abs(np.diag(d)).max()), flush=True)
