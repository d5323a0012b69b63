"""GPU parity tests of the two-atom CZ path (K4 + finalize) through the C ABI against the CPU oracle
(full 25x25 sparse-operator restatement of src/cz_model.jl)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_FIXED = 1e-11
TOL_ADAPTIVE = 1e-6


@pytest.fixture(scope="module")
def czcfg(pkg):
    _, cz = pkg.get_default_configs()
    cz.laser_noise = True
    cz.tspan = np.array([0.0, 0.012, 0.024])
    cz.ξ = 1.3  # phase jump at tspan[end]/2 = 0.012
    return cz


def _inputs(oracle, cz, n, seed=3):
    rng = np.random.default_rng(seed)
    s1, _ = oracle.sample_atoms(cz.trap_params, cz.atom_params, n, seed=seed, stream=0)
    s2, _ = oracle.sample_atoms(cz.trap_params, cz.atom_params, n, seed=seed, stream=1)
    ph = list(rng.uniform(0, 2 * np.pi, (4, n, len(cz.f))))
    return dict(samples1=s1, samples2=s2, phases4=ph, f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes)


@pytest.mark.parametrize("free", [True, False])
def test_fixed_step_parity(pkg, oracle, ctx, czcfg, free):
    cz = czcfg.copy()
    cz.free_motion = free
    n = 11
    kw = _inputs(oracle, cz, n)
    model = pkg.model_from_config(cz, two_atom=True)
    ode = pkg._abi.ode_opts(dt=6e-5, adaptive=False)
    rho, rho2, st = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, n, ode=ode, **kw)
    o, o2, ost = oracle.rydberg2_run(model, cz.tspan, cz.ψ0, n, ode=ode, **kw)
    assert st["n_rhs"] == ost["n_rhs"]
    assert np.abs(rho - o).max() < TOL_FIXED
    assert np.abs(rho2 - o2).max() < TOL_FIXED
    assert np.allclose(np.trace(rho, axis1=1, axis2=2), 1.0, atol=1e-10)


def test_adaptive_parity_and_device_sampling(pkg, oracle, ctx, czcfg):
    cz = czcfg.copy()
    n = 12
    model = pkg.model_from_config(cz, two_atom=True)
    kw = dict(f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes, seed=17, traj_offset=5)
    rho, rho2, st = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, n, **kw)   # on-device Philox samples + phases
    o, o2, ost = oracle.rydberg2_run(model, cz.tspan, cz.ψ0, n, **kw)
    assert st["n_sample_attempts"] == ost["n_sample_attempts"]
    assert np.abs(rho - o).max() < TOL_ADAPTIVE and np.abs(rho2 - o2).max() < TOL_ADAPTIVE
    assert abs(st["n_rhs"] - ost["n_rhs"]) <= 0.01 * ost["n_rhs"]
    # tstops mode
    ode = pkg._abi.ode_opts(dense=False)
    rho, rho2, _ = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, n, ode=ode, **kw)
    o, o2, _ = oracle.rydberg2_run(model, cz.tspan, cz.ψ0, n, ode=ode, **kw)
    assert np.abs(rho - o).max() < TOL_ADAPTIVE and np.abs(rho2 - o2).max() < TOL_ADAPTIVE


def test_other_initial_states_and_flags(pkg, oracle, ctx, czcfg):
    cz = czcfg.copy()
    cz.laser_noise = False
    cz.spontaneous_decay_rydberg = False
    n = 6
    ode = pkg._abi.ode_opts(dt=6e-5, adaptive=False)
    for psi in (pkg.tensor(pkg.ket_1, pkg.ket_1), pkg.tensor(pkg.ket_0, pkg.ket_1), pkg.tensor((pkg.ket_0 + 1j * pkg.ket_1) / np.sqrt(2), pkg.ket_r),
                pkg.tensor((pkg.ket_p + pkg.ket_1) / np.sqrt(2), (pkg.ket_r - 1j * pkg.ket_0) / np.sqrt(2))):
        model = pkg.model_from_config(cz, two_atom=True, rho2_mode=pkg._abi.CA_RHO2_ABS2)
        kw = dict(seed=4, ode=ode)
        rho, rho2, _ = ctx.rydberg2_run(model, cz.tspan, psi, n, **kw)
        o, o2, _ = oracle.rydberg2_run(model, cz.tspan, psi, n, **kw)
        assert np.abs(rho - o).max() < TOL_FIXED and np.abs(rho2 - o2).max() < TOL_FIXED
    # a ψ0 with |l> components lies outside the 16x16 + 4x4 + 4x4 + 1 block structure: ca_rydberg2_run falls back to the full-form
    # kernel and still reproduces the oracle
    for psi_l in (pkg.tensor((pkg.ket_l + pkg.ket_1) / np.sqrt(2), pkg.ket_1), pkg.tensor(pkg.ket_1, pkg.ket_l)):
        model = pkg.model_from_config(cz, two_atom=True)
        rho, rho2, _ = ctx.rydberg2_run(model, cz.tspan, psi_l, 2, seed=4, ode=ode)
        o, o2, _ = oracle.rydberg2_run(model, cz.tspan, psi_l, 2, seed=4, ode=ode)
        assert np.abs(rho - o).max() < TOL_FIXED and np.abs(rho2 - o2).max() < TOL_FIXED
    # maxiters is reported (OrdinaryDiffEq default 1e5; close atom pairs need more: V ~ R^-6 sets the step)
    with pytest.raises(pkg._lib.ColdAtomsError) as e:
        ctx.rydberg2_run(pkg.model_from_config(cz, two_atom=True), cz.tspan, cz.ψ0, 2, ode=pkg._abi.ode_opts(maxiters=20))
    assert e.value.code == pkg._abi.CA_ERR_MAXITERS


def test_partition_independence_and_api(pkg, oracle, ctx, czcfg):
    cz = czcfg.copy()
    model = pkg.model_from_config(cz, two_atom=True)
    kw = dict(f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes, seed=9, out_flags=pkg._abi.CA_OUT_SUMS)
    full, full2, _ = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, 20, **kw)
    a, a2, _ = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, 7, traj_offset=0, **kw)
    b, b2, _ = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, 13, traj_offset=7, **kw)
    assert np.abs(full - (a + b)).max() < 1e-10 and np.abs(full2 - (a2 + b2)).max() < 1e-10
    cz.n_samples = 10
    rho, rho2 = pkg.simulation_czlp(cz, seed=9)
    assert rho.shape == (3, 25, 25)
    assert np.abs(rho - full[:, :, :] * 0 - ctx.rydberg2_run(model, cz.tspan, cz.ψ0, 10, seed=9, f=cz.f, amp_red=cz.red_laser_phase_amplitudes,
                                                             amp_blue=cz.blue_laser_phase_amplitudes)[0]).max() < 1e-12
    probs = pkg.get_two_qubit_probs(rho, rho2)
    assert abs(sum(probs[k][0] for k in ("P00", "P01", "P10", "P11")) - 1.0) < 1e-12
    # empty ensemble
    z, _, st = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, 0, f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes)
    assert np.all(z == 0)


def test_longest_first_handout_keeps_the_ensemble(pkg, oracle, ctx, czcfg):
    """More than two rounds of trajectories switch on the longest-first ordering (bucket sort by the blockade shift of the sampled
    pair): the ensemble sums must not depend on the hand-out order, with host-supplied and with device-drawn samples."""
    cz = czcfg.copy()
    cz.tspan = np.array([0.0, 0.006])
    n = 2600   # > 2 x 148 SMs x 8 warps
    model = pkg.model_from_config(cz, two_atom=True)
    ode = pkg._abi.ode_opts(maxiters=10 ** 7)   # adaptive: a fixed step is unstable for the closest pairs (V ~ R^-6) of 2600 samples
    kw = _inputs(oracle, cz, n)
    rho, rho2, st = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, n, ode=ode, **kw)
    assert st["n_launches"] == 5   # three ordering kernels + integration + finalize
    o, o2, _ = oracle.rydberg2_run(model, cz.tspan, cz.ψ0, n, ode=ode, **kw)
    assert np.abs(rho - o).max() < TOL_ADAPTIVE and np.abs(rho2 - o2).max() < TOL_ADAPTIVE
    # device-drawn samples: same ensemble as two unordered shards
    kw = dict(f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes, seed=3, ode=ode, out_flags=pkg._abi.CA_OUT_SUMS)
    full, _, _ = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, n, **kw)
    a, _, sa = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, 1300, traj_offset=0, **kw)
    b, _, _ = ctx.rydberg2_run(model, cz.tspan, cz.ψ0, 1300, traj_offset=1300, **kw)
    assert sa["n_launches"] == 2
    assert np.abs(full - (a + b)).max() < 1e-9


def test_unstable_fixed_step_is_reported(pkg, ctx, czcfg):
    """A fixed step beyond the stability limit of the closest pairs blows up: CA_ERR_NONFINITE, not silent NaNs."""
    cz = czcfg.copy()
    cz.tspan = np.array([0.0, 0.05])
    cz.laser_noise = False
    cz.atom_centers = [[-0.35, 0.0, 0.0], [0.35, 0.0, 0.0]]   # V ~ 2π·135298/0.7^6 ≈ 7e6 rad/µs: dt·V >> 3
    with pytest.raises(pkg._lib.ColdAtomsError) as e:
        ctx.rydberg2_run(pkg.model_from_config(cz, two_atom=True), cz.tspan, cz.ψ0, 8, seed=2, ode=pkg._abi.ode_opts(dt=1e-4, adaptive=False))
    assert e.value.code == pkg._abi.CA_ERR_NONFINITE
