"""GPU parity tests of the one-atom path (K1+K2+K3+K6) through the C ABI against the CPU oracle.

Tolerances: fixed-step runs take identical step sequences on both sides -> 1e-11 absolute; adaptive runs may flip
an accept/reject decision at rounding level, after which both are equally valid solutions of the same tolerance
-> the north-star bound 1e-6 absolute on every ρ element (populations and fidelities included).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_FIXED = 1e-11
TOL_ADAPTIVE = 1e-6


def _inputs(oracle, cfg, n, seed=11):
    rng = np.random.default_rng(seed)
    samples, _ = oracle.sample_atoms(cfg.trap_params, cfg.atom_params, n, seed=seed)
    pr, pb = rng.uniform(0, 2 * np.pi, (2, n, len(cfg.f)))
    return dict(samples=samples, phases_red=pr, phases_blue=pb, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes,
                amp_blue=cfg.blue_laser_phase_amplitudes)


@pytest.mark.parametrize("psi", ["1", "0+i1", "general"])
@pytest.mark.parametrize("free", [True, False])
def test_fixed_step_parity(pkg, oracle, ctx, r1cfg, psi, free):
    cfg = r1cfg.copy()
    cfg.free_motion = free
    psi0 = {"1": pkg.ket_1, "0+i1": (pkg.ket_0 + 1j * pkg.ket_1) / np.sqrt(2),
            "general": (pkg.ket_0 + pkg.ket_1 + pkg.ket_l + 0.5j * pkg.ket_r) / np.sqrt(3.25)}[psi]
    n = 70  # ragged: two full warps + a partial one
    kw = _inputs(oracle, cfg, n)
    model = pkg.model_from_config(cfg)
    ode = pkg._abi.ode_opts(dt=2e-4, adaptive=False)
    rho, rho2, st = ctx.rydberg1_run(model, cfg.tspan, psi0, n, ode=ode, **kw)
    o, o2, ost = oracle.rydberg1_run(model, cfg.tspan, psi0, n, ode=ode, **kw)
    assert st["n_rhs"] == ost["n_rhs"] and st["n_accept"] == ost["n_accept"]
    assert np.abs(rho - o).max() < TOL_FIXED
    assert np.abs(rho2 - o2).max() < TOL_FIXED
    assert np.allclose(np.trace(rho, axis1=1, axis2=2), 1.0, atol=1e-10)


@pytest.mark.parametrize("dense", [True, False])
def test_adaptive_parity(pkg, oracle, ctx, r1cfg, dense):
    cfg = r1cfg.copy()
    n = 96
    kw = _inputs(oracle, cfg, n, seed=5)
    model = pkg.model_from_config(cfg)
    ode = pkg._abi.ode_opts(dense=dense)
    rho, rho2, st = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, ode=ode, **kw)
    o, o2, ost = oracle.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, ode=ode, **kw)
    assert np.abs(rho - o).max() < TOL_ADAPTIVE
    assert np.abs(rho2 - o2).max() < TOL_ADAPTIVE
    # same algorithm -> same work to within a fraction of a percent
    assert abs(st["n_rhs"] - ost["n_rhs"]) <= 0.005 * ost["n_rhs"]


def test_device_sampler_and_noise_match_oracle(pkg, oracle, ctx, r1cfg):
    """On-device Philox sampling + noise generation = oracle's Philox restatement (same seed, same global indices)."""
    cfg = r1cfg.copy()
    n = 40
    model = pkg.model_from_config(cfg)
    ode = pkg._abi.ode_opts(dt=2e-4, adaptive=False)
    kw = dict(f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes, ode=ode, seed=1234,
              traj_offset=1000)
    rho, rho2, st = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, **kw)
    o, o2, ost = oracle.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, **kw)
    assert st["n_sample_attempts"] == ost["n_sample_attempts"]
    assert np.abs(rho - o).max() < 1e-10
    assert np.abs(rho2 - o2).max() < 1e-10


def test_sharding_is_partition_independent(pkg, ctx, r1cfg):
    """Sums over [0,n) equal sums over [0,k) + [k,n): RNG is keyed by the global trajectory index."""
    cfg = r1cfg.copy()
    model = pkg.model_from_config(cfg)
    kw = dict(f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes, seed=99,
              out_flags=pkg._abi.CA_OUT_SUMS)
    full, full2, _ = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, 100, **kw)
    a, a2, _ = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, 37, traj_offset=0, **kw)
    b, b2, _ = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, 63, traj_offset=37, **kw)
    assert np.abs(full - (a + b)).max() < 1e-10
    assert np.abs(full2 - (a2 + b2)).max() < 1e-10


def test_edge_cases(pkg, ctx, r1cfg):
    cfg = r1cfg.copy()
    model = pkg.model_from_config(cfg)
    # empty ensemble
    rho, rho2, st = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, 0, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes,
                                     amp_blue=cfg.blue_laser_phase_amplitudes)
    assert np.all(rho == 0) and st["n_rhs"] == 0
    # single time point: returns ρ0
    rho, _, _ = ctx.rydberg1_run(model, [0.0], cfg.ψ0, 5, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes,
                                 amp_blue=cfg.blue_laser_phase_amplitudes)
    assert abs(rho[0, 1, 1] - 1.0) < 1e-15
    # maxiters is reported, not silently ignored
    with pytest.raises(pkg._lib.ColdAtomsError) as e:
        ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, 4, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes,
                         amp_blue=cfg.blue_laser_phase_amplitudes, ode=pkg._abi.ode_opts(maxiters=50))
    assert e.value.code == pkg._abi.CA_ERR_MAXITERS
    # bad arguments
    with pytest.raises(pkg._lib.ColdAtomsError):
        ctx.rydberg1_run(model, [0.0, 0.0], cfg.ψ0, 4)


def test_analytic_no_coupling_decay(pkg, ctx, r1cfg):
    """Ω=0: |p> decays with Γ0+Γ1+Γl into 0,1,l with the branching ratios; |r> decays into l with Γr."""
    cfg = r1cfg.copy()
    cfg.laser_noise = False
    cfg.red_laser_params[0] = 0.0
    cfg.blue_laser_params[0] = 0.0
    cfg.tspan = np.linspace(0, 0.2, 5)
    model = pkg.model_from_config(cfg)
    G0, G1, Gl, Gr = cfg.decay_params
    Gp = G0 + G1 + Gl
    rho, _, _ = ctx.rydberg1_run(model, cfg.tspan, pkg.ket_p, 8, ode=pkg._abi.ode_opts(1e-10, 1e-12))
    t = cfg.tspan
    assert np.allclose(rho[:, 3, 3].real, np.exp(-Gp * t), atol=1e-9)
    assert np.allclose(rho[:, 0, 0].real, G0 / Gp * (1 - np.exp(-Gp * t)), atol=1e-9)
    assert np.allclose(rho[:, 1, 1].real, G1 / Gp * (1 - np.exp(-Gp * t)), atol=1e-9)
    assert np.allclose(rho[:, 4, 4].real, Gl / Gp * (1 - np.exp(-Gp * t)), atol=1e-9)
    rho, _, _ = ctx.rydberg1_run(model, cfg.tspan, pkg.ket_r, 8, ode=pkg._abi.ode_opts(1e-10, 1e-12))
    assert np.allclose(rho[:, 2, 2].real, np.exp(-Gr * t), atol=1e-9)


def test_simulation_api_and_sweep(pkg, oracle, ctx, r1cfg):
    cfg = r1cfg.copy()
    cfg.n_samples = 64
    rho, rho2 = pkg.simulation(cfg, seed=42)
    assert rho.shape == (len(cfg.tspan), 5, 5)
    model = pkg.model_from_config(cfg)
    o, o2, _ = oracle.rydberg1_run(model, cfg.tspan, cfg.ψ0, 64, seed=42, f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes,
                                   amp_blue=cfg.blue_laser_phase_amplitudes)
    assert np.abs(rho - o).max() < TOL_ADAPTIVE and np.abs(rho2 - o2).max() < TOL_ADAPTIVE
    # sweep: 3 points with different δ0 / ψ0 / t_end; each must equal its own single run at [0, t_end]
    models, psis, tends = [], [], [0.1, 0.2, 0.25]
    for k, d in enumerate([-3.0, 0.0, 4.0]):
        c = cfg.copy()
        c.detuning_params[1] += d
        models.append(pkg.model_from_config(c))
        psis.append([pkg.ket_1, (pkg.ket_0 + pkg.ket_1) / np.sqrt(2), pkg.ket_1][k])
    n_pp = 48
    kw = dict(f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes, seed=7)
    srho, srho2, st = ctx.rydberg1_sweep(models, psis, tends, n_pp, **kw)
    for k in range(3):
        o, o2, _ = oracle.rydberg1_run(models[k], [0.0, tends[k]], psis[k], n_pp, traj_offset=k * n_pp, **kw)
        assert np.abs(srho[k] - o[-1]).max() < TOL_ADAPTIVE
        assert np.abs(srho2[k] - o2[-1]).max() < TOL_ADAPTIVE


def _r2_args(pkg, kind):
    """R2 workload: params/default.jl:19-54 geometry, uniform red, flat-top HG (n=m=4) or LG (l=6) blue scaled
    as in notebooks/modeling.ipynb cells 7-8 (wb -> wb/sqrt(s), zb -> zb/s)."""
    Γ = 2 * np.pi * 6.0
    Δ0, Ω = 2 * np.pi * 1500.0, 2 * np.pi * 96.0
    wb, zb = 2.0, pkg.w0_to_z0(2.0, 0.475)
    T0 = pkg.T_twophoton(Ω, Ω, Δ0)
    tspan = (T0 / 50) * np.arange(0, 61, 6)
    if kind == "HG":
        s = 4 // 2 + 1
        blue = [Ω, wb / np.sqrt(s), zb / s, "simp_flattop_HG", 4, 4, 1.0]
    else:
        s = 6 // 2 + 1
        blue = [Ω, wb / np.sqrt(s), zb / s, "simp_flattop_LG", 6, 0]
    return dict(tspan=tspan, ψ0=pkg.ket_1, atom_params=[86.9091835, 70.0], trap_params=[1000.0, 1.1, pkg.w0_to_z0(1.1, 0.852, 1.3)],
                red_laser_params=[Ω, 100.0, pkg.w0_to_z0(100.0, 0.795)], blue_laser_params=blue,
                detuning_params=[Δ0, pkg.δ_twophoton(Ω, Ω, Δ0)], decay_params=[Γ / 4, 3 * Γ / 4])


@pytest.mark.parametrize("kind", ["HG", "LG"])
@pytest.mark.parametrize("parallel", [False, True])
def test_arbitrary_beams_blue_intens(pkg, oracle, ctx, kind, parallel):
    """simulation_blue_intens (src/rydberg_model_arb_bms.jl:15-86) through the public entry vs the oracle on the same samples."""
    a = _r2_args(pkg, kind)
    n = 40
    samples, _ = oracle.sample_atoms(a["trap_params"], a["atom_params"], n, seed=11)
    rho, rho2 = pkg.simulation_blue_intens(a["tspan"], a["ψ0"], a["atom_params"], a["trap_params"], samples, a["red_laser_params"],
                                           a["blue_laser_params"], a["detuning_params"], a["decay_params"], parallel=parallel)
    m = pkg.blue_intens_model(a["atom_params"], a["trap_params"], a["red_laser_params"], a["blue_laser_params"], a["detuning_params"],
                              a["decay_params"], parallel=parallel)
    o, o2, ost = oracle.rydberg1_run(m, a["tspan"], a["ψ0"], n, samples=samples)
    assert rho.shape == (len(a["tspan"]), 5, 5)
    assert np.abs(rho - o).max() < TOL_ADAPTIVE and np.abs(rho2 - o2).max() < TOL_ADAPTIVE
    assert abs(pkg.last_stats()["n_rhs"] - ost["n_rhs"]) <= 0.01 * ost["n_rhs"]
    # fixed step: no accept/reject decisions, roundoff-level agreement
    ode = pkg._abi.ode_opts(dt=1e-4, adaptive=False)
    r, r2, _ = ctx.rydberg1_run(m, a["tspan"][:3], a["ψ0"], n, samples=samples, ode=ode)
    o, o2, _ = oracle.rydberg1_run(m, a["tspan"][:3], a["ψ0"], n, samples=samples, ode=ode)
    assert np.abs(r - o).max() < 1e-10 and np.abs(r2 - o2).max() < 1e-10


def test_sweep_over_beam_classes(pkg, oracle, ctx, r1cfg):
    """The anchored-coefficient code is compiled per beam class; a sweep may mix them (Gaussian pair / flat-top points in one launch)
    or consist of flat-top points only.  Every point must equal its own single-model oracle run on the same trajectory indices."""
    a = _r2_args(pkg, "HG")
    hg = pkg.blue_intens_model(a["atom_params"], a["trap_params"], a["red_laser_params"], a["blue_laser_params"], a["detuning_params"],
                               a["decay_params"])
    b = _r2_args(pkg, "LG")
    lg = pkg.blue_intens_model(b["atom_params"], b["trap_params"], b["red_laser_params"], b["blue_laser_params"], b["detuning_params"],
                               b["decay_params"])
    cfg = r1cfg.copy()
    cfg.laser_noise = False
    gauss = pkg.model_from_config(cfg)
    n_pp = 40
    for models in ([gauss, hg, lg], [hg, lg]):
        tends = [0.12, 0.2, 0.17][:len(models)]
        psis = [pkg.ket_1] * len(models)
        srho, srho2, _ = ctx.rydberg1_sweep(models, psis, tends, n_pp, seed=13)
        for k, m in enumerate(models):
            o, o2, _ = oracle.rydberg1_run(m, [0.0, tends[k]], psis[k], n_pp, traj_offset=k * n_pp, seed=13)
            assert np.abs(srho[k] - o[-1]).max() < TOL_ADAPTIVE, k
            assert np.abs(srho2[k] - o2[-1]).max() < TOL_ADAPTIVE, k


def test_rejected_steps_parity(pkg, oracle, ctx, r1cfg):
    """A too large initial step (ode dt = 0.05) forces every trajectory through a series of rejections before the controller
    settles: the accept/reject sequence and the solution must still be the oracle's."""
    cfg = r1cfg.copy()
    n = 70
    kw = _inputs(oracle, cfg, n, seed=9)
    model = pkg.model_from_config(cfg)
    ode = pkg._abi.ode_opts(dt=0.05)
    rho, rho2, st = ctx.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, ode=ode, **kw)
    o, o2, ost = oracle.rydberg1_run(model, cfg.tspan, cfg.ψ0, n, ode=ode, **kw)
    assert ost["n_reject"] >= 3 * n
    assert (st["n_rhs"], st["n_accept"], st["n_reject"]) == (ost["n_rhs"], ost["n_accept"], ost["n_reject"])
    assert np.abs(rho - o).max() < 1e-9 and np.abs(rho2 - o2).max() < 1e-9
