"""CPU tests of the DEVICE SOURCE (coldatoms.jl_b200/csrc/ca_device.cuh is __host__ __device__): the per-lane
sampler, phase-noise recurrence, beam code, reduced Lindblad RHS and DP5 are compiled for the host by the
test-only harness tests/hostemu and compared with the oracle.  The GPU tests repeat this through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
dp = C.POINTER(C.c_double)


def P(a):
    return None if a is None else a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def he():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "hostemu"), "-s"])
    return C.CDLL(os.path.join(HERE, "hostemu", "libhostemu.so"))


def unpack(v):
    nt = v.shape[0]
    M = np.zeros((nt, 5, 5), complex)
    for a in range(5):
        M[:, a, a] = v[:, a]
    q = 0
    for a in range(5):
        for b in range(a + 1, 5):
            M[:, a, b] = v[:, 5 + 2 * q] + 1j * v[:, 6 + 2 * q]
            M[:, b, a] = np.conj(M[:, a, b])
            q += 1
    return M


def he_traj(he, model, tspan, psi0, sample, pr, pb, f, ar, ab, ode, seed=1, gi=0):
    tspan = np.ascontiguousarray(tspan, float)
    nt = len(tspan)
    out = np.zeros((nt, 50))
    st = (C.c_longlong * 4)()
    psi = np.ascontiguousarray(np.asarray(psi0, complex)).view(float)
    arrs = [None if a is None else np.ascontiguousarray(a, float) for a in (sample, pr, pb, f, ar, ab)]
    nf = 0 if f is None else len(f)
    rc = he.he_rydberg1_traj(C.byref(model), P(tspan), nt, P(psi), P(arrs[0]), P(arrs[1]), P(arrs[2]), P(arrs[3]), P(arrs[4]), P(arrs[5]), nf,
                             C.byref(ode), C.c_ulonglong(seed), C.c_ulonglong(gi), P(out), st)
    assert rc == 0, rc
    he_traj.n_exact = st[3]   # steps whose coefficients took the exact path
    return unpack(out[:, :25]), unpack(out[:, 25:]), list(st)[:3]


@pytest.mark.parametrize("psi", ["1", "0+i1", "l", "general"])
@pytest.mark.parametrize("free", [True, False])
def test_lane_matches_oracle(pkg, oracle, he, r1cfg, psi, free):
    cfg = r1cfg.copy()
    cfg.free_motion = free
    psi0 = {"1": pkg.ket_1, "0+i1": (pkg.ket_0 + 1j * pkg.ket_1) / np.sqrt(2), "l": (pkg.ket_l + pkg.ket_p) / np.sqrt(2),
            "general": (pkg.ket_0 + pkg.ket_1 + pkg.ket_l + 0.5j * pkg.ket_r) / np.sqrt(3.25)}[psi]
    rng = np.random.default_rng(1)
    samples, _ = oracle.sample_atoms(cfg.trap_params, cfg.atom_params, 2)
    ph = rng.uniform(0, 2 * np.pi, (2, 397))
    model = pkg.model_from_config(cfg)
    for ode, tol in [(pkg._abi.ode_opts(), 1e-11), (pkg._abi.ode_opts(dt=2e-4, adaptive=False), 1e-12),
                     (pkg._abi.ode_opts(dense=False), 1e-11)]:
        r1, r2, st = he_traj(he, model, cfg.tspan, psi0, samples[1], ph[0], ph[1], cfg.f, cfg.red_laser_phase_amplitudes,
                             cfg.blue_laser_phase_amplitudes, ode)
        o1, o2, ost = oracle.rydberg1_run(model, cfg.tspan, psi0, 1, samples=samples[1:2], phases_red=ph[0:1], phases_blue=ph[1:2], f=cfg.f,
                                          amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes, ode=ode)
        assert st == [ost["n_rhs"], ost["n_accept"], ost["n_reject"]]
        assert np.abs(r1 - o1).max() < tol and np.abs(r2 - o2).max() < tol


def test_lane_abs2_mode_and_cz_sign(pkg, oracle, he, r1cfg):
    cfg = r1cfg.copy()
    model = pkg.model_from_config(cfg, rho2_mode=pkg._abi.CA_RHO2_ABS2, compat=0)
    model.detuning_sign = pkg._abi.CA_SIGN_CZ
    samples, _ = oracle.sample_atoms(cfg.trap_params, cfg.atom_params, 1, seed=5)
    ode = pkg._abi.ode_opts(dt=2e-4, adaptive=False)
    psi0 = (pkg.ket_0 + pkg.ket_1) / np.sqrt(2)
    r1, r2, _ = he_traj(he, model, cfg.tspan, psi0, samples[0], None, None, cfg.f, cfg.red_laser_phase_amplitudes,
                        cfg.blue_laser_phase_amplitudes, ode, seed=77, gi=12)
    o1, o2, _ = oracle.rydberg1_run(model, cfg.tspan, psi0, 1, traj_offset=12, seed=77, samples=samples, f=cfg.f,
                                    amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes, ode=ode)
    assert np.abs(r1 - o1).max() < 1e-12 and np.abs(r2 - o2).max() < 1e-12


def test_lane_arbitrary_beams(pkg, oracle, he):
    """simulation_blue_intens semantics (uniform red, flat-top HG / LG blue, Doppler from vz, two decay channels)."""
    abi = pkg._abi
    for blue in ([2 * np.pi * 96, 2 / np.sqrt(4), 26.4 / 4, "simp_flattop_HG", 4, 4, 1.0], [2 * np.pi * 96, 2.0, 26.4, "simp_flattop_LG", 6, 0]):
        m = abi.ca_model()
        m.atom_m, m.atom_T, m.trap_U0, m.trap_w0, m.trap_z0 = 86.9091835, 70.0, 1000.0, 1.1, 4.6
        m.red = pkg.beam_from_params([2 * np.pi * 96, 100.0, 39517.0], arb_bms=True)
        m.blue = pkg.beam_from_params(blue, arb_bms=True)
        m.Delta0, m.delta0 = 2 * np.pi * 1500, 0.0
        m.Gamma1, m.Gamma0 = 2 * np.pi * 1.5, 2 * np.pi * 4.5
        m.atom_motion = m.free_motion = m.decay_intermediate = 1
        m.doppler_kind, m.parallel, m.n_noise_nodes = abi.CA_DOPPLER_Z, 0, 1001
        tspan = np.linspace(0, 0.3, 7)
        samples, _ = oracle.sample_atoms([1000.0, 1.1, 4.6], [86.9091835, 70.0], 1, seed=3)
        ode = abi.ode_opts()
        r1, r2, st = he_traj(he, m, tspan, pkg.ket_1, samples[0], None, None, None, None, None, ode)
        o1, o2, ost = oracle.rydberg1_run(m, tspan, pkg.ket_1, 1, samples=samples, ode=ode)
        assert st[0] == ost["n_rhs"]
        assert he_traj.n_exact == 0   # flat-top HG / LG beams ride the cubic anchor model, no exact fallback
        assert np.abs(r1 - o1).max() < 1e-10 and np.abs(r2 - o2).max() < 1e-10


def test_noise_recurrence_and_sampler(pkg, oracle, he):
    cfg, _ = pkg.get_default_configs()
    f, amp = np.ascontiguousarray(cfg.f), np.ascontiguousarray(cfg.red_laser_phase_amplitudes)
    n_nodes, dtn = 1001, 2.5 / 1000
    out = np.zeros(n_nodes)
    he.he_noise_trace(P(f), P(amp), len(f), n_nodes, C.c_double(dtn), None, C.c_ulonglong(9), C.c_ulonglong(3), C.c_uint(2), P(out))
    ref = oracle.phi(np.arange(n_nodes) * dtn, f, amp, oracle.phase_draw(9, 3, 2, len(f)))
    assert np.abs(out - ref).max() < 1e-13
    trap, atom = np.array([1000.0, 1.0, 3.7]), np.array([86.9091835, 100.0])
    o, _ = oracle.sample_atoms(trap, atom, 50, offset=4, seed=2, stream=1)
    for i in range(50):
        s = np.zeros(6)
        he.he_sample_atom(P(trap), P(atom), C.c_ulonglong(2), C.c_ulonglong(4 + i), C.c_uint(1), C.c_double(1e-2), P(s))
        assert np.array_equal(s, o[i])  # same operation order on the host: bit-identical


def test_device_beam_code(pkg, oracle, he):
    rng = np.random.default_rng(3)
    xyz = np.ascontiguousarray(rng.normal(size=(300, 3)) * [1.5, 1.5, 30.0])
    for p in ([2.0, 1.3, 7.0, "simp_flattop_HG", 6, 4, 1.3], [2.0, 1.3, 7.0, "simp_flattop_LG", 8, 0], [3.0, 10.0, 661.0, 0.3, 1],
              [3.0, 50.0, 9879.0, np.pi / 2, 1], [3.0, 10.0, 661.0, 0.0, 1]):
        b = pkg.beam_from_params(p, arb_bms=len(p) > 5)
        out = np.zeros((len(xyz), 2))
        assert he.he_eval_beam(C.byref(b), P(xyz), C.c_longlong(len(xyz)), P(out)) == 0
        o = oracle.eval_beam(b, xyz)
        assert np.abs(out[:, 0] + 1j * out[:, 1] - o).max() < 1e-12 * max(1.0, np.abs(o).max())


# ---- two-atom kernel: the per-lane template of ca_lindblad2.cuh executed by 32 host threads in lock-step -------------
def he_traj2(he, model, tspan, psi0, s1, s2, ph4, f, ar, ab, ode, seed=1, gi=0):
    tspan = np.ascontiguousarray(tspan, float)
    nt = len(tspan)
    rho, rho2 = np.zeros((nt, 25, 25), complex), np.zeros((nt, 25, 25), complex)
    st = (C.c_longlong * 4)()
    psi = np.ascontiguousarray(np.asarray(psi0, complex)).view(float)
    arrs = [None if a is None else np.ascontiguousarray(a, float) for a in (s1, s2, f, ar, ab)]
    keep = None if ph4 is None else [np.ascontiguousarray(p, float) for p in ph4]
    ph = None if keep is None else (dp * 4)(*[P(p) for p in keep])
    nf = 0 if f is None else len(f)
    rc = he.he_rydberg2_traj(C.byref(model), P(tspan), nt, P(psi), P(arrs[0]), P(arrs[1]), ph, P(arrs[2]), P(arrs[3]), P(arrs[4]), nf,
                             C.byref(ode), C.c_ulonglong(seed), C.c_ulonglong(gi), rho.ctypes.data_as(dp), rho2.ctypes.data_as(dp), st)
    he_traj2.n_exact = st[3]   # (lane, step) beam evaluations on the exact path
    return rc, rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), list(st)[:3]


@pytest.fixture(scope="module")
def czcfg(pkg):
    _, cz = pkg.get_default_configs()
    cz.laser_noise = True
    cz.tspan = np.array([0.0, 0.004, 0.008])
    cz.ξ = 1.3  # phase jump at tspan[end]/2
    return cz


@pytest.mark.parametrize("free", [True, False])
def test_warp2_fixed_step_matches_oracle(pkg, oracle, he, czcfg, free):
    cz = czcfg.copy()
    cz.free_motion = free
    rng = np.random.default_rng(3)
    s1, _ = oracle.sample_atoms(cz.trap_params, cz.atom_params, 1, seed=3, stream=0)
    s2, _ = oracle.sample_atoms(cz.trap_params, cz.atom_params, 1, seed=3, stream=1)
    ph = list(rng.uniform(0, 2 * np.pi, (4, 1, len(cz.f))))
    model = pkg.model_from_config(cz, two_atom=True)
    ode = pkg._abi.ode_opts(dt=6e-5, adaptive=False)
    kw = dict(f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes)
    rc, r, r2, st = he_traj2(he, model, cz.tspan, cz.ψ0, s1[0], s2[0], [p[0] for p in ph], cz.f, kw["amp_red"], kw["amp_blue"], ode)
    assert rc == 0
    o, o2, ost = oracle.rydberg2_run(model, cz.tspan, cz.ψ0, 1, samples1=s1, samples2=s2, phases4=ph, ode=ode, **kw)
    assert st[0] == ost["n_rhs"]
    assert np.abs(r - o).max() < 1e-12 and np.abs(r2 - o2).max() < 1e-12


def test_warp2_adaptive_states_and_modes(pkg, oracle, he, czcfg):
    cz = czcfg.copy()
    kw = dict(f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes)
    model = pkg.model_from_config(cz, two_atom=True)
    # adaptive with on-"device" Philox samples and phases, dense output and tstops mode
    for ode in (pkg._abi.ode_opts(), pkg._abi.ode_opts(dense=False)):
        rc, r, r2, st = he_traj2(he, model, cz.tspan, cz.ψ0, None, None, None, cz.f, kw["amp_red"], kw["amp_blue"], ode, seed=17, gi=5)
        assert rc == 0
        o, o2, ost = oracle.rydberg2_run(model, cz.tspan, cz.ψ0, 1, seed=17, traj_offset=5, ode=ode, **kw)
        # same accept/reject decisions (taken in double); the controller's fractional powers are single precision on the device
        # path, so step SIZES agree to ~1e-7 and the solutions to ~1e-9 (requirement: 1e-6)
        assert st == [ost["n_rhs"], ost["n_accept"], ost["n_reject"]]
        assert np.abs(r - o).max() < 1e-8 and np.abs(r2 - o2).max() < 1e-8
    # other initial states, |ρ_ij|² second moment, a decay channel switched off, no noise
    cz.laser_noise = False
    cz.spontaneous_decay_rydberg = False
    ode = pkg._abi.ode_opts(dt=6e-5, adaptive=False)
    model = pkg.model_from_config(cz, two_atom=True, rho2_mode=pkg._abi.CA_RHO2_ABS2)
    for psi in (pkg.tensor(pkg.ket_1, pkg.ket_1), pkg.tensor(pkg.ket_0, pkg.ket_1), pkg.tensor((pkg.ket_0 + 1j * pkg.ket_1) / np.sqrt(2), pkg.ket_r),
                pkg.tensor((pkg.ket_p + pkg.ket_1) / np.sqrt(2), (pkg.ket_r - 1j * pkg.ket_0) / np.sqrt(2))):
        rc, r, r2, _ = he_traj2(he, model, cz.tspan, psi, None, None, None, None, None, None, ode, seed=4, gi=0)
        assert rc == 0
        o, o2, _ = oracle.rydberg2_run(model, cz.tspan, psi, 1, seed=4, ode=ode)
        assert np.abs(r - o).max() < 1e-12 and np.abs(r2 - o2).max() < 1e-12
    # ψ0 with |l> components lies outside the block structure: refused, not silently wrong
    rc, *_ = he_traj2(he, model, cz.tspan, pkg.tensor((pkg.ket_l + pkg.ket_1) / np.sqrt(2), pkg.ket_1), None, None, None, None, None, None, ode)
    assert rc == pkg._abi.CA_ERR_UNSUPPORTED


def test_metropolis_chains_device_source_vs_oracle(pkg, oracle, he):
    """Metropolis branch (src/atom_sampler.jl:97-120) as independent chains: device source == oracle restatement on the host;
    one chain reproduces the reference's sequential chain; many chains keep the thermal widths."""
    he.he_metropolis.restype = C.c_double
    trap, atom = np.array([1000.0, 1.0, np.pi / 0.852]), np.array([86.9091835, 50.0])
    for K, harm in ((1, False), (7, False), (64, True)):
        N = 300
        out = np.zeros((N, 6))
        acc = he.he_metropolis(P(trap), P(atom), C.c_longlong(N), C.c_longlong(10), C.c_longlong(200), int(harm), C.c_double(1e-2),
                               C.c_longlong(K), C.c_ulonglong(11), C.c_uint(0), P(out))
        o, oacc = oracle.sample_atoms_metropolis_chains(trap, atom, N, freq=10, skip=200, harmonic=harm, n_chains=K, seed=11)
        assert np.abs(out - o).max() < 1e-12 and abs(acc - oacc) < 1e-12
        assert 0.05 < acc < 0.95
    # n_chains = 1 visits the states of the reference's single chain (legacy oracle routine; operation order differs by rounding)
    one, _ = oracle.sample_atoms_metropolis_chains(trap, atom, 200, freq=10, skip=100, n_chains=1, seed=5)
    ref, _ = oracle.sample_atoms_metropolis(trap, atom, 200, freq=10, skip=100, harmonic=False, seed=5)
    assert np.abs(one - ref).max() < 1e-9
    # many short chains sample the same Boltzmann distribution as the one long reference chain (the true Gaussian trap is
    # softer than its harmonic approximation: x-width ~10 % above sqrt(T w0²/4U0)); velocities are exactly thermal
    big, _ = oracle.sample_atoms_metropolis_chains(trap, atom, 20000, freq=10, skip=1000, n_chains=2000, seed=2)
    long1, _ = oracle.sample_atoms_metropolis(trap, atom, 20000, freq=10, skip=1000, harmonic=False, seed=3)
    for q in (0, 2, 3):
        assert abs(big[:, q].std() / long1[:, q].std() - 1) < 0.06, q
    assert abs(big[:, 3].std() / (0.09118367518929328 * np.sqrt(50 / 86.9091835)) - 1) < 0.04


def test_thermal_average_functionals_device_source_vs_reference_loops(pkg, oracle, he):
    """thermal_terms (ca_device.cuh) against the numpy restatement of calibrate_two_photon's and get_blockade_stark_shift_factor's
    sample loops (src/rydberg_model.jl:293,300; src/cz_model.jl:98-101)."""
    cfg, cz = pkg.get_default_configs()
    m = pkg.config.model_from_config(cfg)
    s1, _ = oracle.sample_atoms(cfg.trap_params, cfg.atom_params, 400, seed=5, stream=0)
    out = np.zeros((400, 2))
    assert he.he_thermal_terms(0, C.byref(m), P(s1), None, C.c_longlong(400), P(out)) == 0
    o2, d2 = oracle.twophoton_terms(m.red, m.blue, s1)
    # δ_twophoton is a cancelling difference (Ωr = Ωb on axis): its error is measured on the scale of each term, |Ωr Ωb / 2Δ|
    assert np.abs(out[:, 0] / o2 - 1).max() < 1e-13 and (np.abs(out[:, 1] - d2) / o2).max() < 1e-13
    m2 = pkg.config.model_from_config(cz, two_atom=True)
    s2, _ = oracle.sample_atoms(cz.trap_params, cz.atom_params, 400, seed=5, stream=1)
    assert he.he_thermal_terms(1, C.byref(m2), P(s1), P(s2), C.c_longlong(400), P(out)) == 0
    ref = oracle.rm6_terms(s1, s2, cz.atom_centers)
    assert np.abs(out[:, 0] / ref - 1).max() < 1e-14 and not out[:, 1].any()


def test_lane_rejected_steps_restore_the_state(pkg, oracle, he, r1cfg):
    """A too large initial step forces rejections: the lane advances y in place and re-forms the old state from the stored stage
    derivatives on a rejection, so the retried steps must reproduce the oracle's accept/reject sequence and its solution."""
    cfg = r1cfg
    m = pkg.config.model_from_config(cfg)
    rng = np.random.default_rng(5)
    sample, _ = oracle.sample_atoms(cfg.trap_params, cfg.atom_params, 1, seed=3)
    pr, pb = rng.uniform(0, 2 * np.pi, len(cfg.f)), rng.uniform(0, 2 * np.pi, len(cfg.f))
    for psi in (pkg.ket_1, (pkg.ket_0 + 1j * pkg.ket_1) / np.sqrt(2)):
        ode = pkg._abi.ode_opts(dt=0.05)
        r1, r2, st = he_traj(he, m, cfg.tspan, psi, sample[0], pr, pb, cfg.f, cfg.red_laser_phase_amplitudes, cfg.blue_laser_phase_amplitudes, ode)
        o1, o2, ost = oracle.rydberg1_run(m, cfg.tspan, psi, 1, samples=sample, phases_red=pr[None], phases_blue=pb[None], f=cfg.f,
                                          amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes, ode=ode)
        assert ost["n_reject"] >= 3 and st == [ost["n_rhs"], ost["n_accept"], ost["n_reject"]]
        assert np.abs(r1 - o1).max() < 1e-9 and np.abs(r2 - o2).max() < 1e-9
