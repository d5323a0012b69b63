"""GPU tests: Philox samplers, phase-noise tables, beams (golden vector), release-recapture (bit-exact indicators)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TRAP = [1000.0, 1.0, np.pi * 1.0 ** 2 / 0.852]
ATOM = [86.9091835, 50.0]


def test_sample_atoms_matches_oracle_philox(ctx, oracle):
    s, att = ctx.sample_atoms(TRAP, ATOM, 5000, traj_offset=17, seed=3, stream=1)
    o, oatt = oracle.sample_atoms(TRAP, ATOM, 5000, offset=17, seed=3, stream=1)
    assert att == oatt
    assert np.abs(s - o).max() < 1e-13 * max(1.0, np.abs(o).max())


def test_sample_atoms_moments(ctx, pkg):
    """Sampler moments: variances (T w0²/4U0, same, T z0²/2U0, vconst² T/m x3), zero means."""
    n = 200000
    s, att = ctx.sample_atoms(TRAP, ATOM, n, seed=1)
    U0, w0, z0 = TRAP
    m, T = ATOM
    sig = np.array([np.sqrt(T * w0 ** 2 / (4 * U0))] * 2 + [np.sqrt(T * z0 ** 2 / (2 * U0))] + [pkg.vconst * np.sqrt(T / m)] * 3)
    assert np.all(np.abs(s.mean(0)) < 5 * sig / np.sqrt(n))
    assert np.allclose(s.std(0), sig, rtol=0.01)  # the energy cut removes ~nothing at T/U0 = 0.05
    assert att >= n and att < 1.001 * n


def test_phase_noise_tables(ctx, oracle, pkg):
    cfg, _ = pkg.get_default_configs()
    tn = np.arange(1001) * (2.5 / 1000)
    n = 9
    rng = np.random.default_rng(0)
    ph = rng.uniform(0, 2 * np.pi, (n, len(cfg.f)))
    tab = ctx.phase_noise(tn, cfg.f, cfg.red_laser_phase_amplitudes, n, phases=ph)
    for i in range(n):
        ref = oracle.phi(tn, cfg.f, cfg.red_laser_phase_amplitudes, ph[i])
        assert np.abs(tab[i] - ref).max() < 1e-12
    # device-drawn phases = oracle's Philox phases
    tab = ctx.phase_noise(tn, cfg.f, cfg.red_laser_phase_amplitudes, 3, traj_offset=5, seed=77, stream=3)
    for i in range(3):
        ref = oracle.phi(tn, cfg.f, cfg.red_laser_phase_amplitudes, oracle.phase_draw(77, 5 + i, 3, len(cfg.f)))
        assert np.abs(tab[i] - ref).max() < 1e-12
    # rms ~ sqrt(sum a²/2)
    big = ctx.phase_noise(tn, cfg.f, cfg.red_laser_phase_amplitudes, 2000, seed=5)
    assert abs(big.std() - np.sqrt(np.sum(np.asarray(cfg.red_laser_phase_amplitudes) ** 2) / 2)) < 0.02 * big.std()


def test_beam_golden_vector_on_device(ctx, pkg):
    """The reference's only deterministic known answers (notebooks/modeling.ipynb cell 13), on the device beam code."""
    g = json.load(open(os.path.join(HERE, "golden", "beam_profiles.json")))
    x = np.array(g["x"])
    xyz = np.stack([x, 0 * x, 0 * x], 1)
    for i in range(0, 10, 2):
        s = max(i, 1)
        b = pkg.beam_from_params([1, 1 / s ** 0.5, 5, "simp_flattop_HG", i, i, 1.0], arb_bms=True)
        assert np.abs(np.abs(ctx.eval_beam(b, xyz)) ** 2 - np.array(g["traces"][f"HG_{i}"])).max() < 1e-13
        b = pkg.beam_from_params([1, 1 / s ** 0.5, 5, "simp_flattop_LG", i, i], arb_bms=True)
        assert np.abs(np.abs(ctx.eval_beam(b, xyz)) ** 2 + 0.001 - np.array(g["traces"][f"LG_{i}"])).max() < 1e-13


def test_beams_off_axis_match_oracle(ctx, oracle, pkg):
    rng = np.random.default_rng(3)
    xyz = rng.normal(size=(500, 3)) * [1.5, 1.5, 30.0]
    for p in ([2.0, 1.3, 7.0, "simp_flattop_HG", 6, 4, 1.3], [2.0, 1.3, 7.0, "simp_flattop_LG", 8, 0],
              [3.0, 10.0, 661.0, 0.3, 1], [3.0, 50.0, 9879.0, np.pi / 2, 1]):
        b = pkg.beam_from_params(p, arb_bms=len(p) > 5)
        d, o = ctx.eval_beam(b, xyz), oracle.eval_beam(b, xyz)
        assert np.abs(d - o).max() < 1e-12 * max(1.0, np.abs(o).max())


def test_release_recapture_bit_exact(ctx, oracle):
    tspan = np.arange(0.0, 51.0)
    n = 10000
    samples, _ = oracle.sample_atoms(TRAP, ATOM, n, seed=21)
    probs, ind, _ = ctx.release_recapture(TRAP, ATOM, tspan, n, samples=samples, want_indicators=True)
    oprobs, oind = oracle.release_recapture(TRAP, ATOM, tspan, n, samples=samples, want_indicators=True)
    assert np.array_equal(ind, oind)          # recapture indicators bit-exact (north star)
    assert np.array_equal(probs, oprobs)
    assert probs[0] == 1.0 and np.all(np.diff(probs) <= 0)  # monotone mask (basic_experiments.jl:39-44)
    # on-device sampling path
    p2, _, st = ctx.release_recapture(TRAP, ATOM, tspan, n, seed=8)
    o2, _ = oracle.release_recapture(TRAP, ATOM, tspan, n, seed=8)
    assert np.array_equal(p2, o2)
    # anchor curve of SURVEY App. C (N = 1e5: P(5,10,20,30,50) = .961 .721 .356 .205 .097), statistical tolerance
    p3, _, _ = ctx.release_recapture(TRAP, ATOM, tspan, 100000, seed=1)
    assert np.allclose(p3[[5, 10, 20, 30, 50]], [0.961, 0.721, 0.356, 0.205, 0.097], atol=0.006)
    # empty input and ragged sizes
    p0, _, _ = ctx.release_recapture(TRAP, ATOM, tspan, 0)
    assert np.all(np.isnan(p0))  # zeros ./ 0 in the reference (basic_experiments.jl:80)
    p1, i1, _ = ctx.release_recapture(TRAP, ATOM, tspan, 33, samples=samples[:33], want_indicators=True)
    assert np.array_equal(i1, oind[:33])


def test_metropolis_sampler_on_device(ctx, oracle, pkg):
    """samples_generate(harmonic=false) (src/atom_sampler.jl:97-120) as independent chains, one per thread."""
    for K in (1, 37, 0):
        n = 500
        s, acc = ctx.sample_atoms_metropolis(TRAP, ATOM, n, freq=10, skip=300, n_chains=K, seed=21)
        Keff = K if K > 0 else n
        o, oacc = oracle.sample_atoms_metropolis_chains(TRAP, ATOM, n, freq=10, skip=300, n_chains=Keff, seed=21)
        # same Philox keys and operation order; libdevice vs glibc log/sincos/exp differ by an ulp at most
        assert np.abs(s - o).max() < 1e-9 and abs(acc - oacc) < 1e-6
    # public entry points
    s, acc = pkg.samples_generate(TRAP, ATOM, 20000, harmonic=False, seed=4)
    assert s.shape == (20000, 6) and 0.05 < acc < 0.95
    h, _ = pkg.samples_generate(TRAP, ATOM, 20000, harmonic=True, seed=4)
    assert abs(s[:, 3].std() / h[:, 3].std() - 1) < 0.04          # thermal velocities
    assert 1.0 < s[:, 0].std() / h[:, 0].std() < 1.25             # the true trap is softer than its harmonic approximation
    tspan = np.arange(0.0, 51.0)
    p, acc = pkg.release_recapture(tspan, TRAP, ATOM, 20000, harmonic=False, seed=6)
    ph, _ = pkg.release_recapture(tspan, TRAP, ATOM, 20000, harmonic=True, seed=6)
    assert p[0] == 1.0 and np.all(np.diff(p) <= 0) and 0.05 < acc < 0.95
    assert np.abs(p - ph).max() < 0.1                              # same physics, slightly different initial ensemble
    # empty input
    s0, _ = ctx.sample_atoms_metropolis(TRAP, ATOM, 0)
    assert s0.shape == (0, 6)


def test_thermal_average_device_reduction_vs_reference_loops(ctx, pkg, oracle):
    """ca_thermal_average against the numpy restatement of the reference's sample loops over the SAME Philox ensemble
    (device-drawn) and over host-supplied samples; shards add up; the API mirrors reproduce calibrate_two_photon /
    get_blockade_stark_shift_factor computed on the host from the oracle's samples (rydberg_model.jl:285-306, cz_model.jl:77-104)."""
    cfg, cz = pkg.get_default_configs()
    abi = pkg._abi
    m = pkg.config.model_from_config(cfg)
    n = 20000
    s1, att1 = oracle.sample_atoms(cfg.trap_params, cfg.atom_params, n, seed=11, stream=0)
    o2, d2 = oracle.twophoton_terms(m.red, m.blue, s1)
    sums, att = ctx.thermal_average(abi.CA_AVG_TWOPHOTON, m, n, seed=11)
    assert att == att1
    # Σ 1/|Δ| is heavy-tailed: compare on the scale of Σ|terms|
    assert abs(sums[0] - o2.sum()) < 1e-12 * o2.sum() and abs(sums[1] - d2.sum()) < 1e-12 * o2.sum()
    sums_h, att_h = ctx.thermal_average(abi.CA_AVG_TWOPHOTON, m, n, samples1=s1)
    assert att_h == 0 and abs(sums_h[0] - o2.sum()) < 1e-12 * o2.sum() and abs(sums_h[1] - d2.sum()) < 1e-12 * o2.sum()
    a, _ = ctx.thermal_average(abi.CA_AVG_TWOPHOTON, m, 7000, traj_offset=0, seed=11)
    b, _ = ctx.thermal_average(abi.CA_AVG_TWOPHOTON, m, n - 7000, traj_offset=7000, seed=11)
    assert np.all(np.abs(a + b - sums) < 1e-12 * o2.sum())
    again, _ = ctx.thermal_average(abi.CA_AVG_TWOPHOTON, m, n, seed=11)
    assert np.array_equal(again, sums)   # fixed-order reduction: bit-reproducible

    m2 = pkg.config.model_from_config(cz, two_atom=True)
    t1, a1 = oracle.sample_atoms(cz.trap_params, cz.atom_params, n, seed=12, stream=0)
    t2, a2 = oracle.sample_atoms(cz.trap_params, cz.atom_params, n, seed=12, stream=1)
    ref = oracle.rm6_terms(t1, t2, cz.atom_centers)
    sums, att = ctx.thermal_average(abi.CA_AVG_RM6, m2, n, seed=12)
    assert att == a1 + a2 and abs(sums[0] / ref.sum() - 1) < 1e-12 and sums[1] == 0.0
    sums_h, _ = ctx.thermal_average(abi.CA_AVG_RM6, m2, n, samples1=t1, samples2=t2)
    assert abs(sums_h[0] / ref.sum() - 1) < 1e-12
    with pytest.raises(pkg._lib.ColdAtomsError):
        ctx.thermal_average(abi.CA_AVG_RM6, m2, n, samples1=t1)
    assert np.array_equal(ctx.thermal_average(abi.CA_AVG_RM6, m2, 0)[0], [0.0, 0.0])

    # API mirrors
    Ω2 = pkg.Ω_twophoton(cfg.red_laser_params[0], cfg.blue_laser_params[0], cfg.detuning_params[0])
    got = pkg.get_blockade_stark_shift_factor(cz.trap_params, cz.atom_params, cz.atom_centers, Ω2, cz.c6, n_samples=n, seed=12)
    assert abs(got / (-Ω2 / (2.0 * cz.c6 * ref.mean())) - 1) < 1e-12
    cal = pkg.calibrate_two_photon(cfg, n_samples=n, seed=11)
    Ωr, Ωb = cfg.red_laser_params[0], cfg.blue_laser_params[0]
    factor = np.sqrt(o2.mean() / Ω2)
    assert abs(cal.red_laser_params[0] / (Ωr / factor) - 1) < 1e-12 and abs(cal.blue_laser_params[0] / (Ωb / factor) - 1) < 1e-12
    mc = pkg.config.model_from_config(cal)
    _, d2c = oracle.twophoton_terms(mc.red, mc.blue, s1)
    want = cfg.detuning_params[1] + d2c.mean() - pkg.δ_twophoton(Ωr, Ωb, cfg.detuning_params[0])
    assert abs(cal.detuning_params[1] - want) < 1e-12 * np.abs(o2).mean()
    # the calibrated config runs through the ensemble entry point (temperature_calibrate=true, rydberg_model.jl:201-203)
    cfg.tspan = np.linspace(0.0, 0.25, 3)
    cfg.n_samples = 64
    ρ, _ = pkg.simulation(cfg, temperature_calibrate=True, seed=11)
    assert abs(np.trace(ρ[-1]).real - 1) < 1e-9
