"""ctypes wrapper of the CPU oracle (oracle/coldatoms_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs.  The product package never imports this module.
"""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def _load_abi():
    import sys
    if "coldatoms_b200" in sys.modules:
        return sys.modules["coldatoms_b200"]._abi
    spec = importlib.util.spec_from_file_location("_coldatoms_abi", os.path.join(_ROOT, "coldatoms.jl_b200", "_abi.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


abi = _load_abi()
_LIB = None
dp = C.POINTER(C.c_double)


def build(force=False):
    so = os.path.join(_HERE, "libcoldatoms_oracle.so")
    src = os.path.join(_HERE, "coldatoms_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libcoldatoms_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.co_vconst.restype = C.c_double
        L.co_LG_coeff.restype = C.c_double
        L.co_HG_coeff.restype = C.c_double
        L.co_sample_atoms.restype = C.c_int64
        L.co_sample_atoms_metropolis.restype = C.c_double
        L.co_sample_atoms_metropolis_chains.restype = C.c_double
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(dp)


def _arr(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _psi(psi0):
    v = np.ascontiguousarray(np.asarray(psi0, dtype=np.complex128))
    return v.view(np.float64)


def set_dense_emulation(on):
    """Switch the RHS to the reference's own dense `mul!` arithmetic (32 products of 5x5 / 66 of 25x25 per RHS): timing baseline only."""
    lib().co_set_dense_emulation(C.c_int(1 if on else 0))


def num_threads():
    return int(lib().co_num_threads())


def eval_beam(beam, xyz):
    xyz = _arr(xyz).reshape(-1, 3)
    out = np.empty((len(xyz), 2))
    lib().co_eval_beam(C.byref(beam), _p(xyz), C.c_int64(len(xyz)), _p(out))
    return out[:, 0] + 1j * out[:, 1]


def HG_coeff(k, n):
    return lib().co_HG_coeff(int(k), int(n))


def LG_coeff(k, n):
    return lib().co_LG_coeff(int(k), int(n))


def philox(seed, index, draw, stream):
    out = (C.c_uint32 * 4)()
    lib().co_philox(C.c_uint64(seed), C.c_uint64(index), C.c_uint32(draw), C.c_uint32(stream), out)
    return list(out)


def sample_atoms(trap, atom, n, offset=0, seed=abi.DEFAULT_SEED, stream=0, eps=1e-2):
    trap, atom = _arr(trap), _arr(atom)
    out = np.empty((n, 6))
    att = lib().co_sample_atoms(_p(trap), _p(atom), C.c_int64(n), C.c_int64(offset), C.c_uint64(seed), C.c_int32(stream),
                                C.c_double(eps), _p(out))
    return out, int(att)


def sample_atoms_metropolis(trap, atom, N, freq=10, skip=1000, harmonic=False, eps=1e-2, seed=abi.DEFAULT_SEED, stream=0):
    trap, atom = _arr(trap), _arr(atom)
    out = np.empty((N, 6))
    acc = lib().co_sample_atoms_metropolis(_p(trap), _p(atom), C.c_int64(N), C.c_int64(freq), C.c_int64(skip),
                                           C.c_int(int(harmonic)), C.c_double(eps), C.c_uint64(seed), C.c_int32(stream), _p(out))
    return out, float(acc)


def sample_atoms_metropolis_chains(trap, atom, N, freq=10, skip=1000, harmonic=False, eps=1e-2, n_chains=1, seed=abi.DEFAULT_SEED, stream=0):
    """The Metropolis kernel run as n_chains independent chains (the GPU path's decomposition; n_chains=1 = the reference chain)."""
    trap, atom = _arr(trap), _arr(atom)
    out = np.empty((N, 6))
    acc = lib().co_sample_atoms_metropolis_chains(_p(trap), _p(atom), C.c_int64(N), C.c_int64(freq), C.c_int64(skip), C.c_int(int(harmonic)),
                                                  C.c_double(eps), C.c_int64(n_chains), C.c_uint64(seed), C.c_int32(stream), _p(out))
    return out, float(acc)


def phase_draw(seed, index, stream, nf):
    out = np.empty(nf)
    lib().co_phase_draw(C.c_uint64(seed), C.c_uint64(index), C.c_uint32(stream), C.c_int(nf), _p(out))
    return out


def phi(tnodes, f, amp, phases):
    tnodes, f, amp, phases = _arr(tnodes), _arr(f), _arr(amp), _arr(phases)
    out = np.empty(len(tnodes))
    lib().co_phi(_p(tnodes), C.c_int(len(tnodes)), _p(f), _p(amp), _p(phases), C.c_int(len(f)), _p(out))
    return out


def rydberg1_run(model, tspan, psi0, n_traj, traj_offset=0, n_total=0, seed=abi.DEFAULT_SEED, samples=None,
                 phases_red=None, phases_blue=None, f=None, amp_red=None, amp_blue=None, ode=None, out_flags=0, nthreads=0):
    tspan = _arr(tspan)
    nt = len(tspan)
    samples, phases_red, phases_blue = _arr(samples), _arr(phases_red), _arr(phases_blue)
    f, amp_red, amp_blue = _arr(f), _arr(amp_red), _arr(amp_blue)
    nf = 0 if f is None else len(f)
    rho = np.zeros((nt, 5, 5), dtype=np.complex128)
    rho2 = np.zeros((nt, 5, 5), dtype=np.complex128)
    st = abi.ca_stats()
    ode = ode if ode is not None else abi.ode_opts()
    rc = lib().co_rydberg1_run(C.byref(model), _p(tspan), C.c_int32(nt), _p(_psi(psi0)), C.c_int64(n_traj),
                               C.c_int64(traj_offset), C.c_int64(n_total), C.c_uint64(seed), _p(samples), _p(phases_red),
                               _p(phases_blue), _p(f), _p(amp_red), _p(amp_blue), C.c_int32(nf), C.byref(ode),
                               C.c_int32(out_flags), rho.ctypes.data_as(dp), rho2.ctypes.data_as(dp), C.byref(st),
                               C.c_int32(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle rydberg1 failed rc={rc}")
    # column-major d x d per time point -> numpy [t, row, col]
    return rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), st.as_dict()


def rydberg2_run(model, tspan, psi0, n_traj, traj_offset=0, n_total=0, seed=abi.DEFAULT_SEED, samples1=None, samples2=None,
                 phases4=None, f=None, amp_red=None, amp_blue=None, ode=None, out_flags=0, nthreads=0):
    tspan = _arr(tspan)
    nt = len(tspan)
    samples1, samples2 = _arr(samples1), _arr(samples2)
    f, amp_red, amp_blue = _arr(f), _arr(amp_red), _arr(amp_blue)
    nf = 0 if f is None else len(f)
    ph_arr = None
    keep = []
    if phases4 is not None:
        keep = [_arr(p) for p in phases4]
        ph_arr = (dp * 4)(*[_p(p) for p in keep])
    rho = np.zeros((nt, 25, 25), dtype=np.complex128)
    rho2 = np.zeros((nt, 25, 25), dtype=np.complex128)
    st = abi.ca_stats()
    ode = ode if ode is not None else abi.ode_opts()
    rc = lib().co_rydberg2_run(C.byref(model), _p(tspan), C.c_int32(nt), _p(_psi(psi0)), C.c_int64(n_traj),
                               C.c_int64(traj_offset), C.c_int64(n_total), C.c_uint64(seed), _p(samples1), _p(samples2), ph_arr,
                               _p(f), _p(amp_red), _p(amp_blue), C.c_int32(nf), C.byref(ode), C.c_int32(out_flags),
                               rho.ctypes.data_as(dp), rho2.ctypes.data_as(dp), C.byref(st), C.c_int32(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle rydberg2 failed rc={rc}")
    return rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), st.as_dict()


def rydberg2_run_ex(model, tspan, psi0=None, rho0=None, n_traj=1, nt_seg1=0, traj_offset=0, n_total=0, seed=abi.DEFAULT_SEED, samples1=None,
                    samples2=None, phases4=None, f=None, amp_red=None, amp_blue=None, ode=None, out_flags=0, nthreads=0):
    """co_rydberg2_run_ex: general initial state (rho0[25,25] numpy [row, col]), two time segments, legacy two-atom flags."""
    tspan = _arr(tspan)
    nt = len(tspan)
    samples1, samples2 = _arr(samples1), _arr(samples2)
    f, amp_red, amp_blue = _arr(f), _arr(amp_red), _arr(amp_blue)
    nf = 0 if f is None else len(f)
    ph_arr, keep = None, []
    if phases4 is not None:
        keep = [_arr(p) for p in phases4]
        ph_arr = (dp * 4)(*[_p(p) for p in keep])
    r0 = None if rho0 is None else np.ascontiguousarray(np.asarray(rho0, dtype=np.complex128).T).view(np.float64)   # column-major
    p0 = None if psi0 is None else _psi(psi0)
    rho = np.zeros((nt, 25, 25), dtype=np.complex128)
    rho2 = np.zeros((nt, 25, 25), dtype=np.complex128)
    st = abi.ca_stats()
    ode = ode if ode is not None else abi.ode_opts()
    rc = lib().co_rydberg2_run_ex(C.byref(model), _p(tspan), C.c_int32(nt), C.c_int32(nt_seg1), _p(p0), _p(r0), C.c_int64(n_traj),
                                  C.c_int64(traj_offset), C.c_int64(n_total), C.c_uint64(seed), _p(samples1), _p(samples2), ph_arr,
                                  _p(f), _p(amp_red), _p(amp_blue), C.c_int32(nf), C.byref(ode), C.c_int32(out_flags),
                                  rho.ctypes.data_as(dp), rho2.ctypes.data_as(dp), C.byref(st), C.c_int32(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle rydberg2_ex failed rc={rc}")
    return rho.transpose(0, 2, 1).copy(), rho2.transpose(0, 2, 1).copy(), st.as_dict()


def release_recapture(trap, atom, tspan, n, traj_offset=0, n_total=0, eps=1e-3, seed=abi.DEFAULT_SEED, samples=None,
                      out_flags=0, want_indicators=False):
    trap, atom, tspan, samples = _arr(trap), _arr(atom), _arr(tspan), _arr(samples)
    nt = len(tspan)
    probs = np.zeros(nt)
    ind = np.zeros((n, nt), dtype=np.uint8) if want_indicators else None
    lib().co_release_recapture(_p(trap), _p(atom), _p(tspan), C.c_int32(nt), C.c_int64(n), C.c_int64(traj_offset),
                               C.c_int64(n_total), C.c_double(eps), C.c_uint64(seed), _p(samples), C.c_int32(out_flags),
                               _p(probs), None if ind is None else ind.ctypes.data_as(C.POINTER(C.c_uint8)))
    return probs, ind


# ---- thermal averages (numpy restatement of the reference's sample loops; the Ω values come from the C oracle's beam code) ----
def doppler_red(vx, vz, beam):
    """Δ(vx, vz, laser_params) -- src/rydberg_model.jl:77-84."""
    k = 2.0 * beam.z0 / beam.w0 ** 2
    return k * np.sin(beam.theta) * vx + k * np.cos(beam.theta) * vz


def twophoton_terms(red, blue, samples):
    """Per-sample Ω_twophoton and δ_twophoton as calibrate_two_photon evaluates them (src/rydberg_model.jl:293 and :300, with
    Ω_twophoton :267-269 = abs(Ωb Ωr / 2Δ) and δ_twophoton :275-277 = (|Ωr|² − |Ωb|²)/4Δ; Δ = the red Doppler shift alone)."""
    s = _arr(samples).reshape(-1, 6)
    Or, Ob = eval_beam(red, s[:, :3]), eval_beam(blue, s[:, :3])
    D = doppler_red(s[:, 3], s[:, 5], red)
    return np.abs(Ob * Or / (2.0 * D)), (np.abs(Or) ** 2 - np.abs(Ob) ** 2) / (4.0 * D)


def rm6_terms(samples1, samples2, centers):
    """1/(1e-18 + sum((s1[1:3] - s2[1:3]).^2)^3) after the centre shifts -- src/cz_model.jl:98-101."""
    s1 = _arr(samples1).reshape(-1, 6)[:, :3] + np.asarray(centers[0], float)
    s2 = _arr(samples2).reshape(-1, 6)[:, :3] + np.asarray(centers[1], float)
    return 1.0 / (1e-18 + np.sum((s1 - s2) ** 2, axis=1) ** 3)
