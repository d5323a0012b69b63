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
