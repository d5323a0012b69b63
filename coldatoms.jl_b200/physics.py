"""Host-side scalar helpers of the reference that sit next to the hot path (exported names kept).

These are O(1) or O(Nf) host computations the reference runs once per call (never per trajectory); they are
part of the user-facing API (`src/NeutralAtoms.jl:22-38`), not of the accelerated path.
"""
import math

import numpy as np

# src/atom_sampler.jl:3-10 with PhysicalConstants.CODATA2018
c = 299792458.0
kB = 1.380649e-23
mu = 1.66053906660e-27
E0 = kB * 1e-6
g0 = 9.81 * 1e-6
vconst = math.sqrt(E0 / mu)
r0 = 1e-6


def w(z, w0, z0):
    """Beam radius, src/utilities.jl:2-4."""
    return w0 * np.sqrt(1.0 + (np.asarray(z) / z0) ** 2)


def w0_to_z0(w0, λ, M2=1.0):
    """Rayleigh length, src/utilities.jl:22-24."""
    return math.pi * w0 ** 2 / λ / M2


def _rot(x, y, z, θ):
    xt = np.sqrt(np.asarray(x, dtype=float) ** 2 + np.asarray(y, dtype=float) ** 2)
    zt = np.asarray(z, dtype=float)
    return xt * math.cos(θ) - zt * math.sin(θ), xt * math.sin(θ) + zt * math.cos(θ)


def A(x, y, z, w0, z0, n=1, θ=0.0):
    """Gaussian-beam amplitude with |E0|=1, src/utilities.jl:27-31."""
    xt, zt = _rot(x, y, z, θ)
    return (w0 / w(zt, w0, z0)) * np.exp(-((xt ** 2) / (w(zt, w0, z0) ** 2)) ** n)


def I(x, y, z, w0, z0, n=1, θ=0.0):  # noqa: E743
    """Intensity, src/utilities.jl:35-39."""
    return A(x, y, z, w0, z0, n=n, θ=θ) ** 2


def A_phase(x, y, z, w0, z0, θ=0.0):
    """Gaussian-beam phase factor, src/utilities.jl:43-48."""
    xt, zt = _rot(x, y, z, θ)
    k = 2.0 * z0 / w0 ** 2
    return np.exp(-1j * k * zt * (0.5 * xt ** 2 / (zt ** 2 + z0 ** 2)) + 1j * np.arctan(zt / z0))


def E(x, y, z, w0, z0, n=1, θ=0.0):
    """Complex amplitude, src/utilities.jl:53-55."""
    return A(x, y, z, w0, z0, n=n, θ=θ) * A_phase(x, y, z, w0, z0, θ=θ)


def trap_frequencies(atom_params, trap_params):
    """(ωr, ωz), src/utilities.jl:73-79 (note ωz uses w0)."""
    m, T = atom_params
    U0, w0, z0 = trap_params
    ω = vconst / w0 * math.sqrt(U0 / m)
    return 2 * ω, math.sqrt(2) * ω


def get_trap_params(ωr, ωz, U0, λ, m=87.0, dif=True):
    """src/atom_sampler.jl:166-178."""
    if dif:
        Δω = ωr - ωz
        w0 = λ / (1.0 - Δω / ωr) / (math.sqrt(2.0) * math.pi)
        z0 = math.pi * w0 ** 2 / λ
        Utemp = m * (ωr / (2.0 * vconst)) ** 2
    else:
        w0 = vconst * math.sqrt(4.0 * U0 / m) / ωr
        z0 = vconst * math.sqrt(2.0 * U0 / m) / ωz
        Utemp = U0
    return w0, z0, Utemp


def Sϕ(f, laser_phase_params):
    """Phase-noise spectral density, src/lasernoise_sampler.jl:2-13."""
    h0, hg, σg, fg = laser_phase_params
    f = np.asarray(f, dtype=float)
    res = 2.0 * h0 * np.ones(len(f))
    for i in range(len(hg)):
        res = res + 2 * hg[i] * np.exp(-(f - fg[i]) ** 2 / (2 * σg[i] ** 2))
    return res / f ** 2


def ϕ_amplitudes(f, laser_phase_params):
    """Spectral amplitudes, src/lasernoise_sampler.jl:16-28."""
    h0, hg, σg, fg = laser_phase_params
    f = np.asarray(f, dtype=float)
    Δf = f[1] - f[0]
    res = 2.0 * h0 * np.ones(len(f))
    for i in range(len(hg)):
        res = res + 2 * hg[i] * np.exp(-(f - fg[i]) ** 2 / (2 * σg[i] ** 2))
    return 2.0 * np.sqrt(Δf * res) / f


def Ω_twophoton(Ωr, Ωb, Δ):
    """src/rydberg_model.jl:267-269."""
    return abs(Ωb * Ωr / (2.0 * Δ))


def T_twophoton(Ωr, Ωb, Δ):
    """src/rydberg_model.jl:271-273."""
    return 2.0 * math.pi / Ω_twophoton(Ωr, Ωb, Δ)


def δ_twophoton(Ωr, Ωb, Δ):
    """src/rydberg_model.jl:275-277."""
    return (abs(Ωr) ** 2 - abs(Ωb) ** 2) / (4.0 * Δ)


def Ωr_required(Ω, Ωb, Δ):
    """src/rydberg_model.jl:279-281."""
    return 2.0 * Δ * Ω / abs(Ωb)


# ASCII aliases
S_phi, phi_amplitudes = Sϕ, ϕ_amplitudes
Omega_twophoton, delta_twophoton, Omegar_required = Ω_twophoton, δ_twophoton, Ωr_required
