"""Profiling driver for the phase-noise generator alone (k_phase_noise)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import numpy as np
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
cfg, _ = pkg.get_default_configs()
ctx = pkg._lib.default_context(0)
tn = np.arange(1001) * (2.5 / 1000)
for r in range(2):
    t = time.perf_counter()
    out = ctx.phase_noise(tn, cfg.f, cfg.red_laser_phase_amplitudes, n, seed=r)
    print("n", n, "wall %.1f ms" % ((time.perf_counter() - t) * 1e3), out.std())
