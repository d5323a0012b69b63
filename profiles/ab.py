"""A/B driver for developer builds (profiles/dev_fast.sh): runs the R1 workload once per library given on the command line, each in
its own process, and prints kernel time, trajectories/s, RHS count and the largest deviation of rho(t) from the FIRST library's.
    python profiles/ab.py n lib1.so lib2.so ..."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    n = int(sys.argv[2])
    cfg, _ = pkg.get_default_configs()
    cfg.laser_noise = True
    ctx = pkg._lib.default_context(0)
    best = None
    if os.environ.get("AB_WORKLOAD", "r1") == "r2":   # flat-top HG n = m = 4 blue, uniform red (profiles/prof_r2.py)
        Γ, Δ0, Ω = 2 * np.pi * 6.0, 2 * np.pi * 1500.0, 2 * np.pi * 96.0
        T0 = pkg.T_twophoton(Ω, Ω, Δ0)
        tspan = (T0 / 50) * np.arange(0, 301)
        blue = [Ω, 2.0 / np.sqrt(3), pkg.w0_to_z0(2.0, 0.475) / 3, "simp_flattop_HG", 4, 4, 1.0]
        model = pkg.blue_intens_model([86.9091835, 70.0], [1000.0, 1.1, pkg.w0_to_z0(1.1, 0.852, 1.3)], [Ω, 100.0, pkg.w0_to_z0(100.0, 0.795)], blue,
                                      [Δ0, pkg.δ_twophoton(Ω, Ω, Δ0)], [Γ / 4, 3 * Γ / 4])
        run = lambda: ctx.rydberg1_run(model, tspan, pkg.ket_1, n, seed=7)   # noqa: E731
    elif os.environ.get("AB_WORKLOAD") == "z1":   # two-atom CZ (BASELINE config 4); AB_TEND shortens the window
        _, cz = pkg.get_default_configs()
        cz.laser_noise = True
        m2 = pkg.model_from_config(cz, two_atom=True)
        ts = np.array([0.0, float(os.environ.get("AB_TEND", cz.tspan[-1]))])
        run = lambda: ctx.rydberg2_run(m2, ts, cz.ψ0, n, seed=7, f=cz.f, amp_red=cz.red_laser_phase_amplitudes,   # noqa: E731
                                       amp_blue=cz.blue_laser_phase_amplitudes, ode=pkg._abi.ode_opts(maxiters=10 ** 7))
    elif os.environ.get("AB_WORKLOAD") in ("trap", "ns15t"):   # atoms oscillating in the trap (free_motion = false): |1> / the superposition state
        cfg.free_motion = False
        model = pkg.model_from_config(cfg)
        psi = cfg.ψ0 if os.environ.get("AB_WORKLOAD") == "trap" else (pkg.ket_0 + 1j * pkg.ket_1) / np.sqrt(2)
        run = lambda: ctx.rydberg1_run(model, cfg.tspan, psi, n, seed=7, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes,   # noqa: E731
                                       amp_blue=cfg.blue_laser_phase_amplitudes)
    elif os.environ.get("AB_WORKLOAD") in ("ns15", "ns21"):   # R1 with a superposition initial state (fidelity callers) / a general one
        model = pkg.model_from_config(cfg)
        psi = (pkg.ket_0 + 1j * pkg.ket_1) / np.sqrt(2)
        if os.environ.get("AB_WORKLOAD") == "ns21":
            psi = (pkg.ket_0 + pkg.ket_1 + pkg.ket_l + 0.5j * pkg.ket_r) / np.sqrt(3.25)
        run = lambda: ctx.rydberg1_run(model, cfg.tspan, psi, n, seed=7, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes,   # noqa: E731
                                       amp_blue=cfg.blue_laser_phase_amplitudes)
    else:
        model = pkg.model_from_config(cfg)
        run = lambda: ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, seed=7, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes,   # noqa: E731
                                       amp_blue=cfg.blue_laser_phase_amplitudes)
    for r in range(3):
        rho, rho2, st = run()
        best = st["kernel_ms"] if best is None else min(best, st["kernel_ms"])
    np.save(sys.argv[3], rho)
    print("kernel_ms %.2f traj/s %.0f rhs/traj %.1f rhs/s %.3g rej %d exact %d" % (best, n / best * 1e3, st["n_rhs"] / n, st["n_rhs"] / best * 1e3, st["n_reject"],
                                                                                 st.get("n_exact_steps", 0)), flush=True)
    sys.exit(0)

n = int(sys.argv[1])
ref = None
for lib in sys.argv[2:]:
    out = "/tmp/ab_%s.npy" % os.path.basename(lib)
    env = dict(os.environ, COLDATOMS_B200_LIB=os.path.join(ROOT, lib))
    r = subprocess.run([sys.executable, __file__, "--child", str(n), out], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-300:]
    dev = ""
    if os.path.exists(out):
        rho = np.load(out)
        if ref is None:
            ref = rho
        dev = " max|drho| vs first %.2e" % np.abs(rho - ref).max()
    print("%-28s %s%s" % (lib, line, dev), flush=True)
