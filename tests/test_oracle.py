"""CPU tests that PIN THE ORACLE: reference golden vector, independent scipy restatement (DOP853 rtol 1e-12 as
converged truth), analytic limits and invariants."""
import json
import os

import numpy as np
import pytest

import pyref

HERE = os.path.dirname(os.path.abspath(__file__))


def test_golden_beam_profiles(pkg, oracle):
    """notebooks/modeling.ipynb cell 13: 10 traces x 401 points, the reference's only deterministic known answers."""
    g = json.load(open(os.path.join(HERE, "golden", "beam_profiles.json")))
    x = np.array(g["x"])
    xyz = np.stack([x, 0 * x, 0 * x], 1)
    for i in range(0, 10, 2):
        s = max(i, 1)
        b = pkg.beam_from_params([1, 1 / s ** 0.5, 5, "simp_flattop_HG", i, i, 1.0], arb_bms=True)
        assert np.abs(np.abs(oracle.eval_beam(b, xyz)) ** 2 - np.array(g["traces"][f"HG_{i}"])).max() < 5e-15
        b = pkg.beam_from_params([1, 1 / s ** 0.5, 5, "simp_flattop_LG", i, i], arb_bms=True)
        assert np.abs(np.abs(oracle.eval_beam(b, xyz)) ** 2 + 0.001 - np.array(g["traces"][f"LG_{i}"])).max() < 5e-15
    # identity visible in the notebook output: n=m=0 flat-top HG == l=0 flat-top LG (both TEM00)
    b0 = pkg.beam_from_params([1, 1.0, 5, "simp_flattop_HG", 0, 0, 1.0], arb_bms=True)
    b1 = pkg.beam_from_params([1, 1.0, 5, "simp_flattop_LG", 0, 0], arb_bms=True)
    assert np.abs(oracle.eval_beam(b0, xyz) - oracle.eval_beam(b1, xyz)).max() < 1e-15


def test_coefficients_match_formulas(pkg, oracle):
    assert oracle.HG_coeff(0, 0) == 1.0 and oracle.LG_coeff(0, 0) == 1.0
    for n in range(5):
        for k in range(n + 1):
            assert abs(oracle.HG_coeff(k, n) - pkg.HG_coeff(k, n)) < 1e-15
            assert abs(oracle.LG_coeff(k, n) - pkg.LG_coeff(k, n)) < 1e-15


def test_philox_known_answer(oracle):
    """Random123 known-answer vectors for philox4x32-10."""
    assert oracle.philox(0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    # counter = ff.., key = ff..
    assert oracle.philox(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]


@pytest.mark.parametrize("noise", [True, False])
def test_one_atom_vs_scipy(pkg, oracle, r1cfg, noise):
    cfg = r1cfg.copy()
    cfg.laser_noise = noise
    cfg.tspan = np.linspace(0, 0.1, 6)
    rng = np.random.default_rng(1)
    samples, _ = oracle.sample_atoms(cfg.trap_params, cfg.atom_params, 1)
    ph = rng.uniform(0, 2 * np.pi, (2, 397))
    ref, _ = pyref.one_atom_traj(cfg, samples[0], ph[0], ph[1])
    model = pkg.model_from_config(cfg)
    kw = dict(samples=samples, phases_red=ph[0:1], phases_blue=ph[1:2], f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes,
              amp_blue=cfg.blue_laser_phase_amplitudes)
    tight, t2, _ = oracle.rydberg1_run(model, cfg.tspan, cfg.ψ0, 1, ode=pkg._abi.ode_opts(1e-11, 1e-13, maxiters=10 ** 7), **kw)
    assert np.abs(tight - ref).max() < 1e-9
    assert np.abs(t2 - tight @ tight).max() < 1e-14  # second output is the matrix square
    default, _, st = oracle.rydberg1_run(model, cfg.tspan, cfg.ψ0, 1, **kw)
    # the reference's default tolerance: ~1e-5 on elements, < 1e-6 on populations
    assert np.abs(default - ref).max() < 5e-5
    assert np.abs(np.einsum("tii->ti", default - ref)).max() < 1e-6
    # tstops mode and fixed-step mode converge to the same answer
    ts, _, _ = oracle.rydberg1_run(model, cfg.tspan, cfg.ψ0, 1, ode=pkg._abi.ode_opts(1e-11, 1e-13, maxiters=10 ** 7, dense=False), **kw)
    assert np.abs(ts - ref).max() < 1e-9
    fx, _, _ = oracle.rydberg1_run(model, cfg.tspan, cfg.ψ0, 1, ode=pkg._abi.ode_opts(dt=2e-5, adaptive=False, maxiters=10 ** 7), **kw)
    assert np.abs(fx - ref).max() < 2e-7


def test_one_atom_trapped_superposition_vs_scipy(pkg, oracle, r1cfg):
    cfg = r1cfg.copy()
    cfg.free_motion = False
    cfg.tspan = np.linspace(0, 0.06, 4)
    cfg.ψ0 = (pkg.ket_0 + 1j * pkg.ket_1) / np.sqrt(2)
    rng = np.random.default_rng(2)
    samples, _ = oracle.sample_atoms(cfg.trap_params, cfg.atom_params, 1, seed=9)
    ph = rng.uniform(0, 2 * np.pi, (2, 397))
    ref, _ = pyref.one_atom_traj(cfg, samples[0], ph[0], ph[1])
    model = pkg.model_from_config(cfg)
    out, _, _ = oracle.rydberg1_run(model, cfg.tspan, cfg.ψ0, 1, samples=samples, phases_red=ph[0:1], phases_blue=ph[1:2], f=cfg.f,
                                    amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes,
                                    ode=pkg._abi.ode_opts(1e-11, 1e-13, maxiters=10 ** 7))
    assert np.abs(out - ref).max() < 1e-9


def test_two_atom_vs_scipy(pkg, oracle):
    _, cz = pkg.get_default_configs()
    cz.laser_noise = True
    cz.tspan = np.array([0.0, 0.004, 0.008])
    cz.ξ = 1.3  # phase jump falls inside the window at t = 0.004
    rng = np.random.default_rng(3)
    s1, _ = oracle.sample_atoms(cz.trap_params, cz.atom_params, 1, seed=1, stream=0)
    s2, _ = oracle.sample_atoms(cz.trap_params, cz.atom_params, 1, seed=1, stream=1)
    ph = rng.uniform(0, 2 * np.pi, (4, 1, 397))
    ref, _ = pyref.two_atom_traj(cz, s1[0], s2[0], [p[0] for p in ph], rtol=1e-10, atol=1e-12)
    model = pkg.model_from_config(cz, two_atom=True)
    out, out2, st = oracle.rydberg2_run(model, cz.tspan, cz.ψ0, 1, samples1=s1, samples2=s2, phases4=list(ph), f=cz.f,
                                        amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes,
                                        ode=pkg._abi.ode_opts(1e-10, 1e-12, maxiters=10 ** 7))
    assert np.abs(out - ref).max() < 2e-8
    assert np.abs(out2 - out @ out).max() < 1e-14
    assert abs(np.trace(out[-1]) - 1) < 1e-9
    # block structure: coherences between an l state of one atom and a non-l state of the same atom vanish
    r = out[-1].reshape(5, 5, 5, 5)  # [a2, a1, b2, b1]
    for a in range(4):
        assert np.abs(r[:, 4, :, a]).max() == 0 and np.abs(r[4, :, a, :]).max() == 0


def test_far_detuned_two_photon_rabi(pkg, oracle):
    """Analytic limit: no motion/noise/decay, δ0 = δ_twophoton -> P_r(t) = sin²(Ω₂ t / 2) with Ω₂ = ΩrΩb/2Δ
    (src/rydberg_model.jl:267-277) up to O(Ω/Δ) corrections."""
    cfg, _ = pkg.get_default_configs()
    cfg.atom_motion = False
    cfg.spontaneous_decay_intermediate = cfg.spontaneous_decay_rydberg = False
    Δ0 = 2 * np.pi * 5000.0
    Ωr, Ωb = 2 * np.pi * 20.0, 2 * np.pi * 20.0
    cfg.red_laser_params[0], cfg.blue_laser_params[0] = Ωr, Ωb
    cfg.red_laser_params[3] = 0.0
    cfg.detuning_params = [Δ0, pkg.δ_twophoton(Ωr, Ωb, Δ0)]
    T0 = pkg.T_twophoton(Ωr, Ωb, Δ0)
    cfg.tspan = np.linspace(0, T0, 9)
    model = pkg.model_from_config(cfg)
    rho, _, _ = oracle.rydberg1_run(model, cfg.tspan, cfg.ψ0, 1, ode=pkg._abi.ode_opts(1e-8, 1e-10, maxiters=10 ** 8))
    Ω2 = pkg.Ω_twophoton(Ωr, Ωb, Δ0)
    assert np.abs(rho[:, 2, 2].real - np.sin(Ω2 * cfg.tspan / 2) ** 2).max() < 5e-3


def test_release_recapture_oracle(oracle):
    trap, atom = [1000.0, 1.0, np.pi / 0.852], [86.9091835, 50.0]
    tspan = np.arange(0.0, 51.0)
    p, ind = oracle.release_recapture(trap, atom, tspan, 20000, seed=4, want_indicators=True)
    assert p[0] == 1.0 and np.all(np.diff(p) <= 0)
    assert np.all(np.diff(ind.astype(int), axis=1) <= 0)  # cumulative AND (basic_experiments.jl:39-44)
    assert np.allclose(p[[5, 10, 20, 30, 50]], [0.961, 0.721, 0.356, 0.205, 0.097], atol=0.012)  # SURVEY App. C anchors
    # Metropolis branch (CPU-only): same thermal widths as the harmonic sampler
    s, acc = oracle.sample_atoms_metropolis(trap, atom, 4000, freq=10, skip=1000, harmonic=False)
    assert 0.05 < acc < 0.95
    assert abs(s[:, 0].std() - np.sqrt(50.0 / 4000.0)) < 0.03
    # empty input
    p0, _ = oracle.release_recapture(trap, atom, tspan, 0)
    assert np.all(np.isnan(p0))  # zeros ./ 0 in the reference (basic_experiments.jl:80)
