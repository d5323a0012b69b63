"""Parameter structs of the reference, same field names and order (src/utilities.jl:177-233), and their
marshalling into the C-ABI POD `ca_model` (include/coldatoms_b200.h)."""
import copy
import math
from dataclasses import dataclass, field
from typing import Any, List

import numpy as np

from . import _abi
from .physics import T_twophoton, w0_to_z0, δ_twophoton, ϕ_amplitudes

# level order, src/rydberg_model.jl:2-7
LEVELS = ("0", "1", "r", "p", "l")


def nlevelstate(i, n=5):
    v = np.zeros(n, dtype=np.complex128)
    v[i] = 1.0
    return v


ket_0, ket_1, ket_r, ket_p, ket_l = (nlevelstate(i) for i in range(5))


def tensor(a, b):
    """QuantumOptics `a ⊗ b`: first factor is the fastest index (data = kron(b, a))."""
    return np.kron(np.asarray(b), np.asarray(a))


@dataclass
class RydbergConfig:
    """src/utilities.jl:177-200 (positional order identical)."""
    tspan: Any
    ψ0: Any
    atom_params: List[float]
    trap_params: List[float]
    n_samples: int
    f: Any
    red_laser_phase_amplitudes: Any
    blue_laser_phase_amplitudes: Any
    red_laser_params: List[Any]
    blue_laser_params: List[Any]
    detuning_params: List[float]
    decay_params: List[float]
    atom_motion: bool
    free_motion: bool
    laser_noise: bool
    spontaneous_decay_intermediate: bool
    spontaneous_decay_rydberg: bool

    def copy(self):
        return copy.deepcopy(self)


@dataclass
class CZLPConfig(RydbergConfig):
    """src/utilities.jl:203-233."""
    atom_centers: List[List[float]] = field(default_factory=lambda: [[-1.0, 0.0, 0.0], [1.0, 0.0, 0.0]])
    c6: float = 0.0
    ΔtoΩ: float = 0.0
    Ωτ: float = 0.0
    ξ: float = 0.0
    ϕ_RZ: float = 0.0


def beam_from_params(p, arb_bms=False):
    """Ω dispatch on the laser_params vector: src/rydberg_model.jl:51-74; Ω_red / Ω_blue of
    src/rydberg_model_arb_bms.jl:1-13 when arb_bms=True."""
    b = _abi.ca_beam()
    b.Omega0, b.w0, b.z0 = float(p[0]), float(p[1]), float(p[2])
    b.theta, b.n, b.sqz = 0.0, 1.0, 1.0
    b.kind = _abi.CA_BEAM_GAUSS
    if len(p) == 3:
        b.kind = _abi.CA_BEAM_UNIFORM if arb_bms else _abi.CA_BEAM_GAUSS
    elif len(p) == 5:
        b.theta, b.n = float(p[3]), float(p[4])
    elif len(p) in (6, 7):
        tag, n, m = p[3], int(p[4]), int(p[5])
        if n == 0 and m == 0 and not arb_bms:
            pass  # TEM00 at θ=0 (rydberg_model.jl:58-59,66-67)
        elif tag == "simp_flattop_LG" and (len(p) == 6 or arb_bms):
            b.kind, b.lg_l = _abi.CA_BEAM_FLATTOP_LG, n
        elif tag == "simp_flattop_HG" and len(p) == 7:
            b.kind, b.hg_n, b.hg_m, b.sqz = _abi.CA_BEAM_FLATTOP_HG, n, m, float(p[6])
        else:
            # the reference silently returns `nothing` here; we refuse instead
            raise ValueError(f"unsupported laser_params / beam_type {p!r}")
    else:
        raise ValueError("laser_params must have 3, 5, 6 or 7 entries")
    return b


def model_from_config(cfg, two_atom=False, compat=_abi.CA_COMPAT_DEFAULT, rho2_mode=_abi.CA_RHO2_MATSQ):
    m = _abi.ca_model()
    m.atom_m, m.atom_T = map(float, cfg.atom_params)
    m.trap_U0, m.trap_w0, m.trap_z0 = map(float, cfg.trap_params)
    m.red = beam_from_params(cfg.red_laser_params)
    m.blue = beam_from_params(cfg.blue_laser_params)
    m.Delta0, m.delta0 = map(float, cfg.detuning_params)
    m.Gamma0, m.Gamma1, m.Gammal, m.Gammar = map(float, cfg.decay_params)
    m.atom_motion, m.free_motion, m.laser_noise = int(cfg.atom_motion), int(cfg.free_motion), int(cfg.laser_noise)
    m.decay_intermediate = int(cfg.spontaneous_decay_intermediate)
    m.decay_rydberg = int(cfg.spontaneous_decay_rydberg)
    m.doppler_kind = _abi.CA_DOPPLER_XZ
    m.parallel = 0
    m.detuning_sign = _abi.CA_SIGN_CZ if two_atom else _abi.CA_SIGN_RYDBERG1
    m.compat = compat
    m.rho2_mode = rho2_mode
    m.n_noise_nodes = 1001
    if two_atom:
        m.center1[:] = [float(v) for v in cfg.atom_centers[0]]
        m.center2[:] = [float(v) for v in cfg.atom_centers[1]]
        m.c6 = float(cfg.c6)
        m.xi = float(cfg.ξ)
    return m


def get_default_configs():
    """`get_default_configs()` of /default.jl:6-157 -> (RydbergConfig, CZLPConfig)."""
    m, T = 86.9091835, 100.0
    atom_params = [m, T]
    U0, w0, λ0, M2 = 1000.0, 1.0, 0.813, 1.0
    z0 = w0_to_z0(w0, λ0, M2)
    trap_params = [U0, w0, z0]

    h0 = 13.0 * 1e-6
    hg1, hg2 = 25.0 * 1e-6, 10.0e3 * 1e-6
    fg1, fg2 = 130.0 * 1e-3, 234.0 * 1e-3
    σg1, σg2 = 18.0 * 1e-3, 1.5 * 1e-3
    red_laser_phase_params = [h0, [hg1, hg2], [σg1, σg2], [fg1, fg2]]
    blue_laser_phase_params = [h0, [hg1, hg2], [σg1, σg2], [fg1, fg2]]
    f = 0.01 + 0.0025 * np.arange(397)  # 0.01:0.0025:1.0
    red_amp = ϕ_amplitudes(f, red_laser_phase_params)
    blue_amp = ϕ_amplitudes(f, blue_laser_phase_params)

    λr, λb, wr, wb = 0.795, 0.475, 50.0, 10.0
    zr, zb = w0_to_z0(wr, λr), w0_to_z0(wb, λb)
    Δ0 = 2.0 * math.pi * 1000.0
    Ω = 2 * math.pi * 2.0
    Ωr = math.sqrt(2 * Δ0 * Ω)
    Ωb = math.sqrt(2 * Δ0 * Ω)
    θr, θb, nr, nb = math.pi / 2, 0.0, 1, 1
    red_laser_params = [Ωr, wr, zr, θr, nr]
    blue_laser_params = [Ωb, wb, zb, θb, nb]
    detuning_params = [Δ0, δ_twophoton(Ωr, Ωb, Δ0)]
    Γ = 2.0 * math.pi * 5.75
    Γ0, Γ1, Γl = Γ / 4, Γ / 4, 2 * Γ / 4
    τr = 111.74
    Γr = 1 / τr
    decay_params = [Γ0, Γ1, Γl, Γr]

    T0 = T_twophoton(Ωr, Ωb, Δ0)
    tspan = (T0 / 20) * np.arange(101)  # 0:T0/20:5T0
    ψ0 = ket_1.copy()
    n_samples = 20
    flags = dict(atom_motion=True, free_motion=True, laser_noise=False,
                 spontaneous_decay_intermediate=True, spontaneous_decay_rydberg=True)
    cfg = RydbergConfig(tspan, ψ0, list(atom_params), list(trap_params), n_samples, f, red_amp, blue_amp,
                        list(red_laser_params), list(blue_laser_params), list(detuning_params),
                        list(decay_params), **flags)

    d = 2.0
    atom_centers = [[-d / 2, 0.0, 0.0], [d / 2, 0.0, 0.0]]
    c6 = 2 * math.pi * 135298
    ΔtoΩ, Ωτ, ξ = 0.377371, 4.29268, 3.90242
    ket_pos = (ket_0 + ket_1) / math.sqrt(2)
    ψ0_cz = tensor(ket_pos, ket_pos)
    Ω2 = 2 * math.pi / T0
    τ = 2 * math.pi / (Ω2 * math.sqrt(ΔtoΩ ** 2 + 2.0))
    ϕ2 = 2 * τ * ΔtoΩ * Ω2
    ϕ1 = (ϕ2 - math.pi) / 2
    tspan_cz = np.array([0.0, 2 * τ])
    cfg_cz = CZLPConfig(tspan_cz, ψ0_cz, list(atom_params), list(trap_params), n_samples, f, red_amp.copy(),
                        blue_amp.copy(), list(red_laser_params), list(blue_laser_params), list(detuning_params),
                        list(decay_params), **flags, atom_centers=atom_centers, c6=c6, ΔtoΩ=ΔtoΩ, Ωτ=Ωτ, ξ=ξ,
                        ϕ_RZ=ϕ1)
    return cfg, cfg_cz
