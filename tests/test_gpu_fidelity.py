"""GPU tests of the fidelity callers (src/fidelity.jl) built on the batched sweep and the CZ ensemble."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_rydberg_infidelity_batched_equals_serial(pkg):
    cfg, _ = pkg.get_default_configs()
    T0 = cfg.tspan[20]
    cfg.tspan = np.array([0.0, T0])   # 2π pulse: returns every qubit state to itself up to the known phase
    U = pkg.RZ(0.0)
    a = pkg.get_rydberg_infidelity(cfg, U=U, n_samples=6, seed=5, batched=True)
    b = pkg.get_rydberg_infidelity(cfg, U=U, n_samples=6, seed=5, batched=False)
    assert list(a) == list(b) == ["Intermdeiate state decay", "Rydberg state decay", "Laser noise", "Atom motion", "Total"]
    for k in a:
        assert abs(a[k] - b[k]) < 1e-9, (k, a[k], b[k])
        assert -1e-9 < a[k] < 1.0
    # each error source alone costs less than all of them together (up to sampling noise of 6 trajectories)
    assert a["Rydberg state decay"] < a["Total"] + 1e-3 and a["Intermdeiate state decay"] < a["Total"] + 1e-3


def test_cz_parity_pipeline(pkg, oracle):
    _, cz = pkg.get_default_configs()
    c = cz.copy()
    c.spontaneous_decay_intermediate = c.spontaneous_decay_rydberg = c.laser_noise = False
    c.atom_params[1] = 1.0
    c.n_samples = 1
    ϕs, F, ϕmax = pkg.get_parity_fidelity(c, seed=3)
    assert len(ϕs) == 62832 and F.shape == ϕs.shape and 0 <= ϕmax < 2 * np.pi
    c.ψ0 = pkg.tensor((pkg.ket_0 + pkg.ket_1) / np.sqrt(2), (pkg.ket_0 + pkg.ket_1) / np.sqrt(2))
    ρ = pkg.simulation_czlp(c, seed=3)[0][-1]
    Fmax, _ = pkg.get_parity_fidelity_temp(ρ, ϕmax)
    assert abs(Fmax - F.max()) < 1e-12
    # the same single trajectory (Philox seed 3, index 0) through the CPU oracle gives the same scan
    o, _, _ = oracle.rydberg2_run(pkg.model_from_config(c, two_atom=True), c.tspan, c.ψ0, 1, seed=3)
    assert np.abs(pkg.parity_scan(o[-1], ϕs[::97]) - F[::97]).max() < 1e-6
    inf, cal = pkg.get_cz_infidelity(cz, n_samples=2, seed=3)
    assert abs(cal - (1.0 - F.max())) < 0.05   # calibration run: ANOTHER 1 µK sample pair of the scan's configuration, at its own best phase
    assert set(inf) == {"Intermdeiate state decay", "Rydberg state decay", "Laser noise", "Atom motion", "Total"}
    assert all(0 <= v < 0.5 for v in inf.values())


def test_cz_infidelity_sweep_matches_oracle_and_serial(pkg, oracle, ctx):
    """get_cz_infidelity's seven ensembles as ONE ca_rydberg2_sweep launch: every point against the ORACLE on the same global
    trajectory indices, and the batched pipeline against the serial one."""
    _, cz = pkg.get_default_configs()
    pts = pkg.cz_infidelity_points(cz, n_samples=5)
    assert [n for n, _ in pts][:2] == ["parity scan", "calibration"] and len(pts) == 7
    models = [pkg.model_from_config(c, two_atom=True) for _, c in pts]
    counts = [int(c.n_samples) for _, c in pts]
    kw = dict(f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes)
    ode = pkg._abi.ode_opts(maxiters=10 ** 6)
    rho, rho2, st = ctx.rydberg2_sweep(models, [c.ψ0 for _, c in pts], [c.tspan[-1] for _, c in pts], counts, seed=11, traj_offset=100, ode=ode, **kw)
    assert st["n_traj"] == sum(counts) and st["status"] == 0
    off = 100
    for k, ((name, c), m, n) in enumerate(zip(pts, models, counts)):
        noise = kw if c.laser_noise else {}
        o, o2, _ = oracle.rydberg2_run(m, [0.0, c.tspan[-1]], c.ψ0, n, traj_offset=off, seed=11, ode=ode, **noise)
        assert np.abs(rho[k] - o[-1]).max() < 1e-6 and np.abs(rho2[k] - o2[-1]).max() < 1e-6, name
        off += n
    a, ca = pkg.get_cz_infidelity(cz, n_samples=3, seed=5, batched=True)
    b, cb = pkg.get_cz_infidelity(cz, n_samples=3, seed=5, batched=False)
    assert abs(ca - cb) < 1e-9 and all(abs(a[k] - b[k]) < 1e-8 for k in a)
