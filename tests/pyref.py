"""Independent numpy/scipy restatement of the reference model builders (dense operators + closures, the way the
Julia source reads), used ONLY to pin the C oracle: scipy DOP853 at rtol 1e-12 is the converged truth.

Follows src/rydberg_model.jl:2-25,28-98,126-191 and src/cz_model.jl:2-69 of the reference.
"""
import math

import numpy as np
from scipy.integrate import solve_ivp

vconst = math.sqrt(1.380649e-23 * 1e-6 / 1.66053906660e-27)


def ket(i, n=5):
    v = np.zeros(n, dtype=complex)
    v[i] = 1
    return v


K0, K1, KR, KP, KL = (ket(i) for i in range(5))


def outer(a, b):
    return np.outer(a, b.conj())


Id = np.eye(5, dtype=complex)
np_, nr_ = outer(KP, KP), outer(KR, KR)
s1p, sp1, spr, srp = outer(K1, KP), outer(KP, K1), outer(KP, KR), outer(KR, KP)
s0p, slp, slr = outer(K0, KP), outer(KL, KP), outer(KL, KR)
operators = [np_, nr_, s1p, sp1, spr, srp]


def kron_q(a, b):  # QuantumOptics a ⊗ b
    return np.kron(b, a)


def A(x, y, z, w0, z0, θ=0.0):
    xt, zt = math.sqrt(x * x + y * y), z
    xt, zt = xt * math.cos(θ) - zt * math.sin(θ), xt * math.sin(θ) + zt * math.cos(θ)
    w = w0 * math.sqrt(1 + (zt / z0) ** 2)
    return (w0 / w) * math.exp(-(xt ** 2) / w ** 2)


def A_phase(x, y, z, w0, z0, θ=0.0):
    xt, zt = math.sqrt(x * x + y * y), z
    xt, zt = xt * math.cos(θ) - zt * math.sin(θ), xt * math.sin(θ) + zt * math.cos(θ)
    k = 2 * z0 / w0 ** 2
    return np.exp(-1j * k * zt * (0.5 * xt ** 2 / (zt ** 2 + z0 ** 2)) + 1j * math.atan(zt / z0))


def Omega(x, y, z, p):
    Ω0, w0, z0, θ, n = p
    return Ω0 * A(x, y, z, w0, z0, θ) * A_phase(x, y, z, w0, z0, θ)


def R(t, ri, vi, ω, free):
    return ri + vi * t if free else ri * math.cos(ω * t) + vi / ω * math.sin(ω * t)


def V(t, ri, vi, ω, free):
    return vi if free else vi * math.cos(ω * t) - ri * ω * math.sin(ω * t)


def Delta(vx, vz, red):
    w0, z0, θ = red[1:4]
    k = 2 * z0 / w0 ** 2
    return k * math.sin(θ) * vx + k * math.cos(θ) * vz


def delta(vx, vz, red, blue):
    wr, zr, θr = red[1:4]
    wb, zb, θb = blue[1:4]
    kr, kb = 2 * zr / wr ** 2, 2 * zb / wb ** 2
    return (kr * math.sin(θr) + kb * math.sin(θb)) * vx + (kr * math.cos(θr) + kb * math.cos(θb)) * vz


def trap_frequencies(atom, trap):
    ω = vconst / trap[1] * math.sqrt(trap[0] / atom[0])
    return 2 * ω, math.sqrt(2) * ω


def phi_table(tn, f, amp, phases):
    return np.sum(amp[:, None] * np.cos(2 * np.pi * f[:, None] * tn[None, :] + phases[:, None]), axis=0)


def closures(sample, cfg, phi_red_tab, phi_blue_tab, tn, extra_red_phase=None, sign=-1.0):
    ωr, ωz = trap_frequencies(cfg.atom_params, cfg.trap_params)
    x, y, z, vx, vy, vz = sample if cfg.atom_motion else np.zeros(6)
    fr = cfg.free_motion
    red, blue = cfg.red_laser_params, cfg.blue_laser_params
    Δ0, δ0 = cfg.detuning_params

    def coefs(t):
        X, Y, Z = R(t, x, vx, ωr, fr), R(t, y, vy, ωr, fr), R(t, z, vz, ωz, fr)
        Vx, Vz = V(t, x, vx, ωr, fr), V(t, z, vy, ωz, fr)  # Vz from vy: rydberg_model.jl:41
        pr = np.interp(t, tn, phi_red_tab) if phi_red_tab is not None else 0.0
        pb = np.interp(t, tn, phi_blue_tab) if phi_blue_tab is not None else 0.0
        if extra_red_phase is not None:
            pr += extra_red_phase(t)
        Ωr = np.exp(1j * pr) * Omega(X, Y, Z, red)
        Ωb = np.exp(1j * pb) * Omega(X, Y, Z, blue)
        return [sign * Delta(Vx, Vz, red) - Δ0, sign * delta(Vx, Vz, red, blue) - δ0,
                Ωr / 2, np.conj(Ωr) / 2, Ωb / 2, np.conj(Ωb) / 2]

    return coefs


def gated_decay(cfg):
    Γ0, Γ1, Γl = cfg.decay_params[:3] if cfg.spontaneous_decay_intermediate else (0, 0, 0)
    Γr = cfg.decay_params[3] if cfg.spontaneous_decay_rydberg else 0.0
    return Γ0, Γ1, Γl, Γr


def master(tspan, rho0, Hfun, J, rtol=1e-12, atol=1e-14, method="DOP853"):
    d = rho0.shape[0]
    JdJ = sum(j.conj().T @ j for j in J)

    def rhs(t, y):
        rho = y.view(complex).reshape(d, d)
        H = Hfun(t)
        out = -1j * (H @ rho - rho @ H)
        for j in J:
            out = out + j @ rho @ j.conj().T
        out = out - 0.5 * (JdJ @ rho + rho @ JdJ)
        return out.reshape(-1).view(float)

    sol = solve_ivp(rhs, (tspan[0], tspan[-1]), rho0.astype(complex).reshape(-1).view(float), t_eval=tspan,
                    method=method, rtol=rtol, atol=atol)
    assert sol.success
    return sol.y.T.copy().view(complex).reshape(len(tspan), d, d), sol.nfev


def one_atom_traj(cfg, sample, phases_red=None, phases_blue=None, **kw):
    tend = cfg.tspan[-1]
    tn = np.arange(1001) * (tend / 1000)
    tr = tb = None
    if cfg.laser_noise:
        tr = phi_table(tn, np.asarray(cfg.f), np.asarray(cfg.red_laser_phase_amplitudes), phases_red)
        tb = phi_table(tn, np.asarray(cfg.f), np.asarray(cfg.blue_laser_phase_amplitudes), phases_blue)
    coefs = closures(sample, cfg, tr, tb, tn)
    Γ0, Γ1, Γl, Γr = gated_decay(cfg)
    J = [math.sqrt(Γ0) * s0p, math.sqrt(Γ1) * s1p, math.sqrt(Γl) * slp, math.sqrt(Γr) * slr]
    psi = np.asarray(cfg.ψ0, dtype=complex)

    def H(t):
        c = coefs(t)
        return sum(ci * op for ci, op in zip(c, operators))

    return master(np.asarray(cfg.tspan), np.outer(psi, psi.conj()), H, J, **kw)


def two_atom_traj(cfg, sample1, sample2, phases4=None, **kw):
    tend = cfg.tspan[-1]
    τ = tend / 2
    tn = np.arange(1001) * (tend / 1000)
    s = [np.array(sample1, dtype=float), np.array(sample2, dtype=float)]
    for a in range(2):
        s[a][:3] += np.asarray(cfg.atom_centers[a])
    tabs = [None] * 4
    if cfg.laser_noise and phases4 is not None:
        for k in range(4):
            amp = cfg.blue_laser_phase_amplitudes if k & 1 else cfg.red_laser_phase_amplitudes
            tabs[k] = phi_table(tn, np.asarray(cfg.f), np.asarray(amp), phases4[k])
    jump = lambda t: 0.0 if t < τ else cfg.ξ  # noqa: E731
    cf = [closures(s[a], cfg, tabs[2 * a], tabs[2 * a + 1], tn, extra_red_phase=jump, sign=+1.0) for a in range(2)]
    ops2 = [kron_q(o, Id) for o in operators] + [kron_q(Id, o) for o in operators] + [kron_q(nr_, nr_)]
    Γ0, Γ1, Γl, Γr = gated_decay(cfg)
    j1 = [math.sqrt(Γ0) * s0p, math.sqrt(Γ1) * s1p, math.sqrt(Γl) * slp, math.sqrt(Γr) * slr]
    J = [kron_q(j, Id) for j in j1] + [kron_q(Id, j) for j in j1]
    ωr, ωz = trap_frequencies(cfg.atom_params, cfg.trap_params)
    fr = cfg.free_motion
    g = [si if cfg.atom_motion else np.zeros(6) for si in s]

    def Vt(t):
        dx = R(t, g[0][0], g[0][3], ωr, fr) - R(t, g[1][0], g[1][3], ωr, fr)
        dy = R(t, g[0][1], g[0][4], ωr, fr) - R(t, g[1][1], g[1][4], ωr, fr)
        dz = R(t, g[0][2], g[0][5], ωz, fr) - R(t, g[1][2], g[1][5], ωz, fr)
        return cfg.c6 / (1e-18 + (dx * dx + dy * dy + dz * dz) ** 3)

    def H(t):
        c = cf[0](t) + cf[1](t) + [Vt(t)]
        return sum(ci * op for ci, op in zip(c, ops2))

    psi = np.asarray(cfg.ψ0, dtype=complex)
    return master(np.asarray(cfg.tspan), np.outer(psi, psi.conj()), H, J, **kw)


# ---------------------------------------------------------------------------------------------------------------------------
# Legacy two-atom drivers, src/two_atom_model.jl, restated in the reference's OLD basis (g, p, r, gt, zero) with dense operators:
# two_atom_simulation (:89-190) and direct_CZ_simulation (:192-319).  Used only to pin the oracle's / the kernels' handling of the
# legacy parameterisations (level map g->1, p->p, r->r, gt->0, zero->l done by the HOST wrappers, not here).
# ---------------------------------------------------------------------------------------------------------------------------
from scipy.special import eval_genlaguerre, eval_hermite  # noqa: E402

G_, P_, R_, GT_, Z_ = (ket(i) for i in range(5))   # two_atom_model.jl:3-8


def _fact(n):
    return float(math.factorial(n))


def LG_coeff(k, n):   # arbitrary_beams.jl:24-30
    return (-1.0) ** k * sum(_fact(i) / (2 ** i * _fact(i - k)) for i in range(k, n + 1)) / _fact(k)


def HG_coeff(k, n):   # arbitrary_beams.jl:32-38
    return sum(_fact(2 * i) / (_fact(i) * 2 ** (3 * i) * _fact(i - k) * _fact(2 * k)) for i in range(k, n + 1))


def flattopHG(x, y, z, p):   # arbitrary_beams.jl:50-67
    Ω, w0, zr, _, n, m, sqz = p
    w = w0 * math.sqrt(1 + (z / zr) ** 2)
    k = 2 * zr / w0 ** 2
    phase0 = k * z + k / 2 * z * (x * x + y * y) / (z * z + zr * zr)
    E = 0j
    for i in range(0, int(n) + 1, 2):
        for j in range(0, int(m) + 1, 2):
            gouy = (i + j + 1) * math.atan(z / zr)
            cij = HG_coeff(i // 2, int(n) // 2) * HG_coeff(j // 2, int(m) // 2)
            hx = math.exp(-x * x / w ** 2) * eval_hermite(i, math.sqrt(2) * x / w)
            hy = math.exp(-y * y / (w * sqz) ** 2) * eval_hermite(j, math.sqrt(2) * y / (w * sqz))
            E += cij * hx * hy * np.exp(-1j * gouy)
    return Ω * w0 * E / w * np.exp(-1j * phase0)


def flattopLG(x, y, z, p):   # arbitrary_beams.jl:69-81
    Ω, w0, zr, _, l, _m = p
    w = w0 * math.sqrt(1 + (z / zr) ** 2)
    k = 2 * zr / w0 ** 2
    r2 = x * x + y * y
    phase0 = k * z + k / 2 * z * r2 / (z * z + zr * zr)
    E = 0j
    for i in range(int(l) // 2 + 1):
        E += LG_coeff(i, int(l) // 2) * math.exp(-r2 / w ** 2) * eval_genlaguerre(i, 0, 2 * r2 / w ** 2) * np.exp(-1j * (i + 1) * math.atan(z / zr))
    return Ω * w0 / w * np.exp(-1j * phase0) * E


def Omega_blue(x, y, z, p):   # rydberg_model_arb_bms.jl:6-13
    return flattopLG(x, y, z, p) if p[3] == "simp_flattop_LG" else flattopHG(x, y, z, p)


def _legacy_detunings(vz, red, blue, parallel):
    """Δ(Vz, red) and δ(Vz, red, blue[1:3]; parallel): absent from the snapshot, semantics of SURVEY App. B7."""
    kr, kb = 2 * red[2] / red[1] ** 2, 2 * blue[2] / blue[1] ** 2
    return kr * vz, (kr + kb if parallel else kr - kb) * vz


def legacy_two_atom_traj(tspan, psi0, atom_params, trap_params, s1, s2, red1, blue1, red2, blue2, det1, det2, decay_params, V_rr, beam_coords,
                         atom_motion=True, free_motion=True, spontaneous_decay=True, parallel=False, **kw):
    """two_atom_simulation (two_atom_model.jl:89-190) for ONE sample pair; ρ(t) in the old basis, index a1 + 5 a2."""
    dist = 3.4
    X0, Y0 = beam_coords
    ωr, ωz = trap_frequencies(atom_params, trap_params)
    Δ01, δ01 = det1
    Δ02, δ02 = det2
    Γg, Γgt = decay_params if spontaneous_decay else (0.0, 0.0)
    one = [outer(P_, P_), outer(R_, R_), outer(G_, P_), outer(P_, G_), outer(P_, R_), outer(R_, P_)]   # np, nr, σgp, σpg, σpr, σrp
    ops = [kron_q(o, Id) for o in one] + [kron_q(Id, o) for o in one] + [kron_q(outer(R_, R_), outer(R_, R_))]
    j1 = [math.sqrt(Γg) * outer(G_, P_), math.sqrt(Γgt) * outer(GT_, P_)]
    J = [kron_q(j, Id) for j in j1] + [kron_q(Id, j) for j in j1]   # physically correct Jdagger
    a = [np.array(s1, float) if atom_motion else np.zeros(6), np.array(s2, float) if atom_motion else np.zeros(6)]
    fr = free_motion

    def H(t):
        c = []
        for k, (red, blue, Δ0, δ0, off) in enumerate(((red1, blue1, Δ01, δ01, 0.0), (red2, blue2, Δ02, δ02, dist))):
            x, y, z, vx, vy, vz = a[k]
            X, Y, Z, Vz = R(t, x, vx, ωr, fr), R(t, y, vy, ωr, fr), R(t, z, vz, ωz, fr), V(t, z, vz, ωz, fr)
            D, d = _legacy_detunings(Vz, red1, blue1, parallel)    # both atoms: the FIRST atom's laser parameters (:158-159, :168-169)
            Ωr, Ωb = red[0], Omega_blue(off + X - X0, Y - Y0, Z, blue)
            c += [-D - Δ0, -d - δ0, Ωr / 2, np.conj(Ωr) / 2, Ωb / 2, np.conj(Ωb) / 2]
        c.append(V_rr)
        return sum(ci * op for ci, op in zip(c, ops))

    psi = np.asarray(psi0, dtype=complex)
    return master(np.asarray(tspan), np.outer(psi, psi.conj()), H, J, **kw)


def legacy_direct_cz_traj(tspan1, tspan2, rho1, atom_params, trap_params, s1, s2, blue1, blue2, det1, det2, decay_params, V_rr, phase,
                          atom_motion=True, free_motion=True, spontaneous_decay=True, **kw):
    """direct_CZ_simulation (two_atom_model.jl:192-319) for ONE sample pair: two consecutive solves, the second with exp(i phase)
    on the couplings; Hermitian couplings in BOTH segments (the reference's first segment is not, see include/coldatoms_b200.h)."""
    ωr, ωz = trap_frequencies(atom_params, trap_params)
    Γg, Γgt = decay_params if spontaneous_decay else (0.0, 0.0)
    one = [outer(R_, R_), outer(G_, R_), outer(R_, G_)]   # nr, σgr, σrg
    ops = [kron_q(o, Id) for o in one] + [kron_q(Id, o) for o in one] + [kron_q(outer(R_, R_), outer(R_, R_))]
    j1 = [math.sqrt(Γg) * outer(G_, P_), math.sqrt(Γgt) * outer(GT_, P_)]
    J = [kron_q(j, Id) for j in j1] + [kron_q(Id, j) for j in j1]
    a = [np.array(s1, float) if atom_motion else np.zeros(6), np.array(s2, float) if atom_motion else np.zeros(6)]
    fr = free_motion

    def Hseg(ph):
        def H(t):
            c = []
            for k, (blue, Δ0) in enumerate(((blue1, det1[0]), (blue2, det2[0]))):
                x, y, z, vx, vy, vz = a[k]
                Ωd = np.exp(1j * ph) * Omega(R(t, x, vx, ωr, fr), R(t, y, vy, ωr, fr), R(t, z, vz, ωz, fr), blue)
                c += [Δ0, Ωd / 2, np.conj(Ωd) / 2]
            c.append(V_rr)
            return sum(ci * op for ci, op in zip(c, ops))
        return H

    r1, n1 = master(np.asarray(tspan1), np.asarray(rho1, dtype=complex), Hseg(0.0), J, **kw)
    r2, n2 = master(np.asarray(tspan2), r1[-1], Hseg(phase), J, **kw)
    return np.concatenate([r1, r2]), n1 + n2
