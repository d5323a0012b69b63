"""Legacy two-atom drivers of the reference (src/two_atom_model.jl; commented out of the module at src/NeutralAtoms.jl:47 but part of
the hot path's scope): `two_atom_simulation` (:89-190) and `direct_CZ_simulation` (:192-319).

Both are marshalling layers over ONE `ca_rydberg2_run_ex` call (the full 25x25 kernel): the legacy parameterisation travels in
`ca_model.two_atom_flags` (per-atom lasers and detunings, constant V_rr, beam-frame offsets, direct g<->r coupling with a phase on the
second time segment).  The reference writes these drivers in its OLD level order (g, p, r, gt, zero) (two_atom_model.jl:3-8); the
library works in (0, 1, r, p, l), so states go in and density matrices come out through the level map g->1, p->p, r->r, gt->0,
zero->l.  Deviations from the (dead) reference code, both documented in SURVEY App. B11 / include/coldatoms_b200.h: the physically
correct J† is used for atom 2, and the first segment of the direct drive couples σgr / σrg with Ω/2 and conj(Ω)/2 (Hermitian).
"""
import numpy as np

from . import _abi, _lib
from .api import _last_stats, _next_seed, _ode
from .config import beam_from_params

__all__ = ["two_atom_simulation", "direct_CZ_simulation", "legacy_two_atom_model", "legacy_direct_model", "old_to_new_index"]

_OLD2NEW = np.array([1, 3, 2, 0, 4])   # old level (g, p, r, gt, zero) -> level index in (0, 1, r, p, l)
BEAM_DIST = 3.4                        # `dist = 3.4` (two_atom_model.jl:100): x offset of atom 2 in the blue beam's frame


def old_to_new_index():
    """new composite index a1 + 5 a2 of every old composite index."""
    i1, i2 = np.meshgrid(np.arange(5), np.arange(5), indexing="ij")
    new = np.zeros(25, dtype=int)
    new[(i1 + 5 * i2).ravel()] = (_OLD2NEW[i1] + 5 * _OLD2NEW[i2]).ravel()
    return new


def _state_in(psi_old):
    out = np.zeros(25, dtype=np.complex128)
    out[old_to_new_index()] = np.asarray(psi_old, dtype=np.complex128).reshape(25)
    return out


def _rho_in(rho_old):
    p = old_to_new_index()
    out = np.zeros((25, 25), dtype=np.complex128)
    out[np.ix_(p, p)] = np.asarray(rho_old, dtype=np.complex128).reshape(25, 25)
    return out


def _rho_out(rho_new):
    p = old_to_new_index()
    return np.asarray(rho_new)[..., p[:, None], p[None, :]]


def _base(atom_params, trap_params, decay_params, atom_motion, free_motion, spontaneous_decay):
    m = _abi.ca_model()
    m.atom_m, m.atom_T = map(float, atom_params)
    m.trap_U0, m.trap_w0, m.trap_z0 = map(float, trap_params)
    Γg, Γgt = decay_params if spontaneous_decay else (0.0, 0.0)
    m.Gamma1, m.Gamma0, m.Gammal, m.Gammar = float(Γg), float(Γgt), 0.0, 0.0   # J = sqrt(Γg) σgp, sqrt(Γgt) σgtp (:117)
    m.atom_motion, m.free_motion, m.laser_noise = int(atom_motion), int(free_motion), 0
    m.decay_intermediate, m.decay_rydberg = 1, 0
    m.detuning_sign, m.compat, m.rho2_mode, m.n_noise_nodes = _abi.CA_SIGN_RYDBERG1, 0, _abi.CA_RHO2_MATSQ, 1001
    return m


def legacy_two_atom_model(atom_params, trap_params, red_laser_params_first, blue_laser_params_first, red_laser_params_second,
                          blue_laser_params_second, detuning_params_first, detuning_params_second, decay_params, V_rr, beam_coords,
                          atom_motion=True, free_motion=True, spontaneous_decay=True, parallel=False):
    """ca_model of `two_atom_simulation` (two_atom_model.jl:89-190)."""
    m = _base(atom_params, trap_params, decay_params, atom_motion, free_motion, spontaneous_decay)
    m.red = beam_from_params(list(red_laser_params_first)[:3], arb_bms=True)
    m.blue = beam_from_params(blue_laser_params_first, arb_bms=True)
    m.red2 = beam_from_params(list(red_laser_params_second)[:3], arb_bms=True)
    m.blue2 = beam_from_params(blue_laser_params_second, arb_bms=True)
    m.Delta0, m.delta0 = map(float, detuning_params_first)
    m.Delta0_2, m.delta0_2 = map(float, detuning_params_second)
    m.doppler_kind, m.parallel = _abi.CA_DOPPLER_Z, int(parallel)
    X0, Y0 = beam_coords
    m.center1[:] = [-float(X0), -float(Y0), 0.0]              # Ω_blue(X1 - X0, Y1 - Y0, Z1, ...)        (:160-161)
    m.center2[:] = [BEAM_DIST - float(X0), -float(Y0), 0.0]   # Ω_blue(dist + X2 - X0, Y2 - Y0, Z2, ...) (:167-168)
    m.V_rr = float(V_rr)
    m.two_atom_flags = _abi.CA_2A_PER_ATOM_LASERS | _abi.CA_2A_CONST_VRR | _abi.CA_2A_SHIFT_AFTER_MOTION
    return m


def legacy_direct_model(atom_params, trap_params, blue_laser_params_first, blue_laser_params_second, detuning_params_first,
                        detuning_params_second, decay_params, V_rr, phase, atom_motion=True, free_motion=True, spontaneous_decay=True):
    """ca_model of `direct_CZ_simulation` (two_atom_model.jl:192-319): Δ0 on nr, Ω(blue)/2 on σgr / σrg, V_rr, phase on segment 2."""
    m = _base(atom_params, trap_params, decay_params, atom_motion, free_motion, spontaneous_decay)
    m.red.kind = m.red2.kind = _abi.CA_BEAM_UNIFORM          # unused
    m.red.w0 = m.red.z0 = m.red2.w0 = m.red2.z0 = 1.0
    m.blue = beam_from_params(blue_laser_params_first)        # Ω(x, y, z, laser_params): the 5-parameter Gaussian beam (:268)
    m.blue2 = beam_from_params(blue_laser_params_second)
    m.Delta0, m.Delta0_2 = float(detuning_params_first[0]), float(detuning_params_second[0])
    m.doppler_kind = _abi.CA_DOPPLER_Z
    m.V_rr, m.seg2_phase = float(V_rr), float(phase)
    m.two_atom_flags = _abi.CA_2A_PER_ATOM_LASERS | _abi.CA_2A_CONST_VRR | _abi.CA_2A_DIRECT
    return m


def _samples(s):
    return np.ascontiguousarray(np.asarray(s, dtype=float).reshape(-1, 6))


def two_atom_simulation(tspan, ψ0, atom_params, trap_params, samples_first, samples_second, red_laser_params_first,
                        blue_laser_params_first, red_laser_params_second, blue_laser_params_second, detuning_params_first,
                        detuning_params_second, decay_params, V_rr, beam_coords, atom_motion=True, free_motion=True,
                        spontaneous_decay=True, parallel=False, device=None, **ode_kwargs):
    """`two_atom_simulation(...)` (src/two_atom_model.jl:89-190) -> (ρ_mean, ρ2_mean), 25x25 per time point in the reference's OLD
    basis order (g, p, r, gt, zero) ⊗ (same).  ψ0: 25-vector in that order."""
    m = legacy_two_atom_model(atom_params, trap_params, red_laser_params_first, blue_laser_params_first, red_laser_params_second,
                              blue_laser_params_second, detuning_params_first, detuning_params_second, decay_params, V_rr, beam_coords,
                              atom_motion, free_motion, spontaneous_decay, parallel)
    s1, s2 = _samples(samples_first), _samples(samples_second)
    if len(s1) != len(s2):
        raise ValueError("lengths are not equal")   # the reference prints this and runs on (two_atom_model.jl:103-105)
    ctx = _lib.default_context(device)
    ρ, ρ2, st = ctx.rydberg2_run_ex(m, tspan, psi0=_state_in(ψ0), n_traj=len(s1), samples1=s1, samples2=s2, ode=_ode(ode_kwargs))
    _last_stats.clear()
    _last_stats.update(st)
    return _rho_out(ρ), _rho_out(ρ2)


def direct_CZ_simulation(tspan1, tspan2, ρ_1, atom_params, trap_params, samples_first, samples_second, red_laser_params_first,
                         blue_laser_params_first, red_laser_params_second, blue_laser_params_second, detuning_params_first,
                         detuning_params_second, decay_params, V_rr, phase, atom_motion=True, free_motion=True,
                         spontaneous_decay=True, parallel=False, device=None, **ode_kwargs):
    """`direct_CZ_simulation(...)` (src/two_atom_model.jl:192-319) -> (ρ_mean, ρ2_mean) at the length(tspan1) + length(tspan2)
    concatenated output times; ρ_1: 25x25 initial density matrix in the OLD basis order.  The red laser parameters and `parallel`
    are accepted and ignored, as in the reference."""
    m = legacy_direct_model(atom_params, trap_params, blue_laser_params_first, blue_laser_params_second, detuning_params_first,
                            detuning_params_second, decay_params, V_rr, phase, atom_motion, free_motion, spontaneous_decay)
    s1, s2 = _samples(samples_first), _samples(samples_second)
    if len(s1) != len(s2):
        raise ValueError("lengths are not equal")
    t = np.concatenate([np.asarray(tspan1, float), np.asarray(tspan2, float)])
    ctx = _lib.default_context(device)
    ρ, ρ2, st = ctx.rydberg2_run_ex(m, t, rho0=_rho_in(ρ_1), nt_seg1=len(tspan1), n_traj=len(s1), samples1=s1, samples2=s2, ode=_ode(ode_kwargs))
    _last_stats.clear()
    _last_stats.update(st)
    return _rho_out(ρ), _rho_out(ρ2)
