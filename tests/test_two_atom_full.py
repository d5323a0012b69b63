"""The full-form two-atom path (k_lindblad2_full / ca_rydberg2_run_ex) and the legacy drivers of src/two_atom_model.jl.

CPU part: (1) the oracle's legacy parameterisations against the INDEPENDENT old-basis scipy restatement tests/pyref.py (dense operators
in the reference's (g, p, r, gt, zero) order, DOP853), through the host wrappers' level map; (2) the device source
(ca_lindblad2f.cuh run by 32 host threads, tests/hostemu) against the oracle: ψ0 with |l> components, a density matrix as the initial
state, two_atom_simulation and direct_CZ_simulation parameterisations, two time segments.  GPU part: the same through the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

from test_hostemu import he, P  # noqa: F401

dp = C.POINTER(C.c_double)


def he_traj2f(he, model, tspan, psi0=None, rho0=None, s1=None, s2=None, ph4=None, f=None, ar=None, ab=None, ode=None, nt_seg1=0, seed=1, gi=0):
    tspan = np.ascontiguousarray(tspan, float)
    nt = len(tspan)
    rho, rho2 = np.zeros((nt, 25, 25), complex), np.zeros((nt, 25, 25), complex)
    st = (C.c_longlong * 4)()
    psi = None if psi0 is None else np.ascontiguousarray(np.asarray(psi0, complex)).view(float)
    r0 = None if rho0 is None else np.ascontiguousarray(np.asarray(rho0, complex).reshape(25, 25).T).view(float)
    arrs = [None if a is None else np.ascontiguousarray(a, float) for a in (s1, s2, f, ar, ab)]
    keep = None if ph4 is None else [np.ascontiguousarray(p, float) for p in ph4]
    ph = None if keep is None else (dp * 4)(*[P(p) for p in keep])
    nf = 0 if f is None else len(f)
    rc = he.he_rydberg2f_traj(C.byref(model), P(tspan), nt, nt_seg1, P(psi), P(r0), P(arrs[0]), P(arrs[1]), ph, P(arrs[2]), P(arrs[3]), P(arrs[4]), nf,
                              C.byref(ode), C.c_ulonglong(seed), C.c_ulonglong(gi), rho.ctypes.data_as(dp), rho2.ctypes.data_as(dp), st)
    return rc, rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), list(st)[:3]


def _legacy_setup(pkg, oracle):
    Ω, Δ0, Γ = 2 * np.pi * 96.0, 2 * np.pi * 1500.0, 2 * np.pi * 6.0
    a = dict(atom=[86.9091835, 70.0], trap=[1000.0, 1.1, pkg.w0_to_z0(1.1, 0.852, 1.3)],
             red1=[Ω, 100.0, pkg.w0_to_z0(100.0, 0.795)], red2=[0.9 * Ω, 100.0, pkg.w0_to_z0(100.0, 0.795)],
             blue1=[Ω, 2.0 / np.sqrt(3), pkg.w0_to_z0(2.0, 0.475) / 3, "simp_flattop_HG", 4, 4, 1.0],
             blue2=[1.1 * Ω, 2.0 / np.sqrt(4), pkg.w0_to_z0(2.0, 0.475) / 4, "simp_flattop_LG", 6, 0],
             det1=[Δ0, 1.0], det2=[Δ0 * 1.01, -2.0], dec=[Γ / 4, 3 * Γ / 4], V_rr=2 * np.pi * 30.0, bc=(0.3, -0.2))
    a["tspan"] = np.linspace(0, pkg.T_twophoton(Ω, Ω, Δ0) / 2, 5)
    a["s1"], _ = oracle.sample_atoms(a["trap"], a["atom"], 3, seed=3, stream=0)
    a["s2"], _ = oracle.sample_atoms(a["trap"], a["atom"], 3, seed=3, stream=1)
    k = np.eye(5, dtype=complex)
    a1 = (k[0] + k[3]) / np.sqrt(2)
    a2 = k[0] + 1j * k[2] + 0.3 * k[1]
    a["psi_old"] = np.kron(a2 / np.linalg.norm(a2), a1)   # old basis (g, p, r, gt, zero), index a1 + 5 a2
    a["b1"] = [2 * np.pi * 2.0, 3.0, pkg.w0_to_z0(3.0, 0.475), 0.2, 1]
    a["b2"] = [2 * np.pi * 2.2, 3.5, pkg.w0_to_z0(3.5, 0.475), 0.0, 1]
    a["rho1"] = 0.8 * np.outer(a["psi_old"], a["psi_old"].conj()) + 0.2 * np.eye(25) / 25
    a["t1"], a["t2"] = np.linspace(0, 0.3, 4), np.linspace(0, 0.25, 3)
    return a


def _two_atom_model(pkg, a, free, parallel):
    from coldatoms_b200 import legacy
    return legacy.legacy_two_atom_model(a["atom"], a["trap"], a["red1"], a["blue1"], a["red2"], a["blue2"], a["det1"], a["det2"], a["dec"],
                                        a["V_rr"], a["bc"], free_motion=free, parallel=parallel)


def _direct_model(pkg, a, free):
    from coldatoms_b200 import legacy
    return legacy.legacy_direct_model(a["atom"], a["trap"], a["b1"], a["b2"], [2 * np.pi * 0.3], [-2 * np.pi * 0.2], a["dec"], 2 * np.pi * 5.0, 1.234,
                                      free_motion=free)


@pytest.mark.parametrize("free", [True, False])
def test_oracle_legacy_drivers_match_the_independent_old_basis_restatement(pkg, oracle, free):
    """Pins the oracle's legacy parameterisations (flags, level map, signs, beam offsets, direct coupling, second-segment phase) to
    tests/pyref.py written in the reference's old basis the way src/two_atom_model.jl reads."""
    import pyref
    from coldatoms_b200 import legacy
    a = _legacy_setup(pkg, oracle)
    tight = pkg._abi.ode_opts(1e-10, 1e-12, maxiters=10 ** 7)
    for par in (False, True):
        m = _two_atom_model(pkg, a, free, par)
        o, o2, _ = oracle.rydberg2_run_ex(m, a["tspan"], psi0=legacy._state_in(a["psi_old"]), n_traj=1, samples1=a["s1"][:1], samples2=a["s2"][:1], ode=tight)
        ref, _ = pyref.legacy_two_atom_traj(a["tspan"], a["psi_old"], a["atom"], a["trap"], a["s1"][0], a["s2"][0], a["red1"], a["blue1"], a["red2"],
                                            a["blue2"], a["det1"], a["det2"], a["dec"], a["V_rr"], a["bc"], free_motion=free, parallel=par,
                                            rtol=1e-11, atol=1e-13)
        assert np.abs(legacy._rho_out(o) - ref).max() < 5e-9
    m = _direct_model(pkg, a, free)
    t = np.concatenate([a["t1"], a["t2"]])
    o, o2, _ = oracle.rydberg2_run_ex(m, t, rho0=legacy._rho_in(a["rho1"]), nt_seg1=len(a["t1"]), n_traj=1, samples1=a["s1"][:1], samples2=a["s2"][:1], ode=tight)
    ref, _ = pyref.legacy_direct_cz_traj(a["t1"], a["t2"], a["rho1"], a["atom"], a["trap"], a["s1"][0], a["s2"][0], a["b1"], a["b2"], [2 * np.pi * 0.3],
                                         [-2 * np.pi * 0.2], a["dec"], 2 * np.pi * 5.0, 1.234, free_motion=free, rtol=1e-11, atol=1e-13)
    assert np.abs(legacy._rho_out(o) - ref).max() < 5e-9
    assert np.abs(o2 - o @ o).max() < 1e-13


def _cases(pkg, oracle):
    """(name, model, run kwargs) of everything the full-form path must cover."""
    from coldatoms_b200 import legacy
    a = _legacy_setup(pkg, oracle)
    _, cz = pkg.get_default_configs()
    cz.laser_noise = True
    cz.tspan = np.array([0.0, 0.004, 0.008])
    cz.ξ = 1.3
    noise = dict(f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes)
    psi_l = pkg.tensor((pkg.ket_l + pkg.ket_1 + 1j * pkg.ket_0) / np.sqrt(3), (pkg.ket_1 + pkg.ket_r) / np.sqrt(2))
    out = [("czlp_psi0_with_l", pkg.model_from_config(cz, two_atom=True), dict(tspan=cz.tspan, psi0=psi_l, **noise)),
           ("czlp_block_state_on_full_kernel", pkg.model_from_config(cz, two_atom=True), dict(tspan=cz.tspan, rho0=np.outer(cz.ψ0, np.conj(cz.ψ0)), **noise))]
    czt = cz.copy()
    czt.free_motion = False
    out.append(("czlp_trapped_psi0_with_l", pkg.model_from_config(czt, two_atom=True, rho2_mode=pkg._abi.CA_RHO2_ABS2), dict(tspan=cz.tspan, psi0=psi_l, **noise)))
    for free in (True, False):
        out.append((f"two_atom_simulation_free{int(free)}", _two_atom_model(pkg, a, free, False),
                    dict(tspan=a["tspan"], psi0=legacy._state_in(a["psi_old"]), samples=(a["s1"], a["s2"]))))
        out.append((f"direct_cz_free{int(free)}", _direct_model(pkg, a, free),
                    dict(tspan=np.concatenate([a["t1"], a["t2"]]), rho0=legacy._rho_in(a["rho1"]), nt_seg1=len(a["t1"]), samples=(a["s1"], a["s2"]))))
    return out


def test_full_form_device_source_matches_oracle(pkg, oracle, he):
    for name, model, kw in _cases(pkg, oracle):
        kw = dict(kw)
        tspan, smp = kw.pop("tspan"), kw.pop("samples", None)
        nz = {k: kw.pop(k) for k in ("f", "amp_red", "amp_blue") if k in kw}
        dt_fixed = min((tspan[1] - tspan[0]) / 64, 0.25 / max(abs(model.Delta0), 1.0))   # inside the stability region of the explicit pair
        for ode, tol in ((pkg._abi.ode_opts(), 1e-8), (pkg._abi.ode_opts(dt=dt_fixed, adaptive=False), 1e-12)):
            hk = dict(s1=None if smp is None else smp[0][1], s2=None if smp is None else smp[1][1], f=nz.get("f"), ar=nz.get("amp_red"), ab=nz.get("amp_blue"))
            rc, r, r2, st = he_traj2f(he, model, tspan, ode=ode, seed=17, gi=5, **hk, **kw)
            assert rc == 0, name
            ok = dict(samples1=None if smp is None else smp[0][1:2], samples2=None if smp is None else smp[1][1:2], seed=17, traj_offset=5, **nz)
            o, o2, ost = oracle.rydberg2_run_ex(model, tspan, n_traj=1, ode=ode, **ok, **kw)
            assert st == [ost["n_rhs"], ost["n_accept"], ost["n_reject"]], name
            assert np.abs(r - o).max() < tol and np.abs(r2 - o2).max() < tol, (name, np.abs(r - o).max(), np.abs(r2 - o2).max())
            assert abs(np.trace(r[-1]).real - 1.0) < 1e-9


@pytest.mark.gpu
def test_full_form_kernel_matches_oracle_through_the_c_abi(pkg, oracle, ctx):
    for name, model, kw in _cases(pkg, oracle):
        kw = dict(kw)
        tspan, smp = kw.pop("tspan"), kw.pop("samples", None)
        n = 3 if smp is not None else 7
        sk = {} if smp is None else dict(samples1=smp[0], samples2=smp[1])
        dt_fixed = min((tspan[1] - tspan[0]) / 64, 0.25 / max(abs(model.Delta0), 1.0))
        for ode, tol in ((pkg._abi.ode_opts(), 1e-6), (pkg._abi.ode_opts(dt=dt_fixed, adaptive=False), 1e-11)):
            rho, rho2, st = ctx.rydberg2_run_ex(model, tspan, n_traj=n, ode=ode, seed=17, traj_offset=5, **sk, **kw)
            o, o2, ost = oracle.rydberg2_run_ex(model, tspan, n_traj=n, ode=ode, seed=17, traj_offset=5, **sk, **kw)
            assert st["n_rhs"] == ost["n_rhs"] and st["n_sample_attempts"] == ost["n_sample_attempts"], name
            assert np.abs(rho - o).max() < tol and np.abs(rho2 - o2).max() < tol, (name, np.abs(rho - o).max())
    # the blocked and the full kernel agree on a state both can run (ψ0 inside the {0,1,r,p}² block)
    _, cz = pkg.get_default_configs()
    cz.laser_noise = True
    cz.tspan = np.array([0.0, 0.01, 0.02])
    m = pkg.model_from_config(cz, two_atom=True)
    noise = dict(f=cz.f, amp_red=cz.red_laser_phase_amplitudes, amp_blue=cz.blue_laser_phase_amplitudes, seed=3)
    a, a2, _ = ctx.rydberg2_run(m, cz.tspan, cz.ψ0, 40, **noise)
    b, b2, _ = ctx.rydberg2_run_ex(m, cz.tspan, rho0=np.outer(cz.ψ0, np.conj(cz.ψ0)), n_traj=40, **noise)
    assert np.abs(a - b).max() < 1e-8 and np.abs(a2 - b2).max() < 1e-8


@pytest.mark.gpu
def test_legacy_drivers_public_api(pkg, oracle):
    """two_atom_simulation / direct_CZ_simulation (src/two_atom_model.jl:89-190, :192-319) through the public entry points, in the
    reference's old basis order, against the independent scipy restatement."""
    import pyref
    a = _legacy_setup(pkg, oracle)
    n = 3
    ρ, ρ2 = pkg.two_atom_simulation(a["tspan"], a["psi_old"], a["atom"], a["trap"], a["s1"], a["s2"], a["red1"], a["blue1"], a["red2"], a["blue2"],
                                    a["det1"], a["det2"], a["dec"], a["V_rr"], a["bc"], reltol=1e-9, abstol=1e-11)
    ref = np.mean([pyref.legacy_two_atom_traj(a["tspan"], a["psi_old"], a["atom"], a["trap"], a["s1"][i], a["s2"][i], a["red1"], a["blue1"], a["red2"],
                                              a["blue2"], a["det1"], a["det2"], a["dec"], a["V_rr"], a["bc"], rtol=1e-11, atol=1e-13)[0] for i in range(n)], axis=0)
    assert ρ.shape == (len(a["tspan"]), 25, 25) and np.abs(ρ - ref).max() < 1e-7
    ρ, ρ2 = pkg.direct_CZ_simulation(a["t1"], a["t2"], a["rho1"], a["atom"], a["trap"], a["s1"], a["s2"], a["red1"], a["b1"], a["red2"], a["b2"],
                                     [2 * np.pi * 0.3, 0.0], [-2 * np.pi * 0.2, 0.0], a["dec"], 2 * np.pi * 5.0, 1.234, reltol=1e-9, abstol=1e-11)
    ref = np.mean([pyref.legacy_direct_cz_traj(a["t1"], a["t2"], a["rho1"], a["atom"], a["trap"], a["s1"][i], a["s2"][i], a["b1"], a["b2"], [2 * np.pi * 0.3],
                                               [-2 * np.pi * 0.2], a["dec"], 2 * np.pi * 5.0, 1.234, rtol=1e-11, atol=1e-13)[0] for i in range(n)], axis=0)
    assert ρ.shape == (len(a["t1"]) + len(a["t2"]), 25, 25) and np.abs(ρ - ref).max() < 1e-7
