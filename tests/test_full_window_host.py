"""CPU mirror of tests/test_gpu_full_window.py: the DEVICE SOURCE (ca_device.cuh / ca_lindblad2.cuh compiled for the host by
tests/hostemu) against the oracle on the full benchmarked windows, trajectory by trajectory -- R1 0:T0/20:5T0 with noise, R2
0:T0/50:6T0 with flat-top beams, Z1 [0, 2τ] -- and on the stress cases of the anchored-coefficient guards.  Same step sequences
are required (host builds evaluate the error norm in full precision) and the solutions must agree far inside the 1e-6 bound."""
import ctypes as C
import os

import numpy as np
import pytest

from test_gpu_full_window import _noise_kw, _r2_model
from test_hostemu import he, he_traj, he_traj2  # noqa: F401  (fixture + helpers)

TOL = 1e-9


def _cmp1(pkg, oracle, he, model, tspan, psi0, n, seed, cfg=None, samples=None, tol=TOL):
    nk = _noise_kw(cfg) if cfg is not None else {}
    f, ar, ab = (nk.get("f"), nk.get("amp_red"), nk.get("amp_blue"))
    worst, steps, exact = 0.0, 0, 0
    ode = pkg._abi.ode_opts()
    for i in range(n):
        smp = None if samples is None else samples[i]
        r1, r2, st = he_traj(he, model, tspan, psi0, smp, None, None, f, ar, ab, ode, seed=seed, gi=i)
        kw = dict(traj_offset=i, seed=seed, ode=ode, **nk)
        if samples is not None:
            kw["samples"] = samples[i:i + 1]
        o1, o2, ost = oracle.rydberg1_run(model, tspan, psi0, 1, **kw)
        assert st == [ost["n_rhs"], ost["n_accept"], ost["n_reject"]], i
        err = max(np.abs(r1 - o1).max(), np.abs(r2 - o2).max())
        assert err < tol, (i, err)
        worst = max(worst, err)
        steps += st[1] + st[2]
        exact += he_traj.n_exact
    return worst, steps, exact


def test_r1_full_window_device_source(pkg, oracle, he):
    cfg, _ = pkg.get_default_configs()
    cfg.laser_noise = True
    worst, steps, exact = _cmp1(pkg, oracle, he, pkg.model_from_config(cfg), cfg.tspan, cfg.ψ0, 24, 20261020, cfg)
    assert exact <= 1e-3 * steps
    cfg.free_motion = False
    worst, steps, exact = _cmp1(pkg, oracle, he, pkg.model_from_config(cfg), cfg.tspan, cfg.ψ0, 6, 7, cfg)
    assert exact <= 1e-3 * steps
    cfg.free_motion = True
    worst, steps, exact = _cmp1(pkg, oracle, he, pkg.model_from_config(cfg), cfg.tspan, (pkg.ket_0 + 1j * pkg.ket_1) / np.sqrt(2), 6, 7, cfg)
    assert exact <= 1e-3 * steps
    _cmp1(pkg, oracle, he, pkg.model_from_config(cfg), cfg.tspan, (pkg.ket_0 + pkg.ket_1 + pkg.ket_l + 0.5j * pkg.ket_r) / np.sqrt(3.25), 3, 7, cfg)


@pytest.mark.parametrize("kind,order", [("HG", 4), ("LG", 6)])
def test_r2_full_window_device_source(pkg, oracle, he, kind, order):
    m, tspan, atom, trap = _r2_model(pkg, kind, order)
    samples, _ = oracle.sample_atoms(trap, atom, 6, seed=100 + order)
    worst, steps, exact = _cmp1(pkg, oracle, he, m, tspan, pkg.ket_1, 6, 1, samples=samples)
    assert exact <= 2e-2 * steps


@pytest.mark.parametrize("case", ["hot_500uK", "narrow_blue_1um", "HG16", "HG_squeezed", "trapped_hot", "late_window"])
def test_stress_cases_device_source(pkg, oracle, he, case):
    cfg, _ = pkg.get_default_configs()
    cfg.laser_noise = True
    if case in ("HG16", "HG_squeezed"):
        order, sqz = (16, 1.0) if case == "HG16" else (6, 1.6)
        m, tspan, atom, trap = _r2_model(pkg, "HG", order, T=150.0, sqz=sqz)
        n = 2 if case == "HG16" else 6
        samples, _ = oracle.sample_atoms(trap, atom, n, seed=3)
        _cmp1(pkg, oracle, he, m, tspan[:151 if case == "HG_squeezed" else 41], pkg.ket_1, n, 1, samples=samples)
        return
    if case == "hot_500uK":
        cfg.atom_params[1] = 500.0
    elif case == "narrow_blue_1um":
        cfg.blue_laser_params[1] = 1.0
        cfg.blue_laser_params[2] = pkg.w0_to_z0(1.0, 0.475)
    elif case == "trapped_hot":
        cfg.atom_params[1] = 300.0
        cfg.free_motion = False
    elif case == "late_window":
        cfg.tspan = np.linspace(0.4, 1.9, 31)
    _cmp1(pkg, oracle, he, pkg.model_from_config(cfg), cfg.tspan, cfg.ψ0, 8, 99, cfg)


def test_z1_full_window_device_source(pkg, oracle, he):
    _, cz = pkg.get_default_configs()
    cz.laser_noise = True
    model = pkg.model_from_config(cz, two_atom=True)
    ode = pkg._abi.ode_opts(maxiters=10 ** 7)
    kw = dict(f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes)
    for gi in (0, 1):
        rc, r, r2, st = he_traj2(he, model, cz.tspan, cz.ψ0, None, None, None, cz.f, kw["amp_red"], kw["amp_blue"], ode, seed=11, gi=gi)
        assert rc == 0
        o, o2, ost = oracle.rydberg2_run(model, cz.tspan, cz.ψ0, 1, seed=11, traj_offset=gi, ode=ode, **kw)
        assert st == [ost["n_rhs"], ost["n_accept"], ost["n_reject"]]
        assert max(np.abs(r - o).max(), np.abs(r2 - o2).max()) < 1e-7


def test_device_source_reproduces_full_window_golden(pkg, he):
    """The host-compiled DEVICE source against the committed scipy DOP853 trajectories over the benchmarked windows
    (tests/golden/pyref_full_window.npz): the CPU-suite twin of test_cuda_path_reproduces_full_window_golden."""
    from test_golden_trajectories import FW, fw_r1_case, fw_z1_case
    for i in range(2):
        cfg, model, kw = fw_r1_case(pkg, i)
        for ode, tol, ptol in [(pkg._abi.ode_opts(1e-11, 1e-13, maxiters=10 ** 7), 1e-9, 1e-9), (pkg._abi.ode_opts(), 3e-5, 1e-6)]:
            r1, r2, st = he_traj(he, model, cfg.tspan, cfg.ψ0, kw["samples"][0], kw["phases_red"][0], kw["phases_blue"][0], cfg.f, kw["amp_red"], kw["amp_blue"], ode)
            assert np.abs(r1 - FW["r1_rho"][i]).max() < tol, i
            assert np.abs(np.einsum("tii->ti", r1 - FW["r1_rho"][i])).max() < ptol, i
    cz, model, kw = fw_z1_case(pkg)
    # (the tight-tolerance Z1 leg, 48 k steps of 32 host threads at a barrier, is left to the oracle and to the GPU twin of this test)
    for ode, tol, ptol in [(pkg._abi.ode_opts(maxiters=10 ** 7), 5e-5, 1e-6)]:
        rc, r, r2, st = he_traj2(he, model, cz.tspan, cz.ψ0, kw["samples1"][0], kw["samples2"][0], [p[0] for p in kw["phases4"]], cz.f, kw["amp_red"], kw["amp_blue"], ode)
        assert rc == 0
        assert np.abs(r - FW["z1_rho"]).max() < tol
        assert np.abs(np.einsum("tii->ti", r - FW["z1_rho"])).max() < ptol
