"""Fidelity callers of the hot path (src/fidelity.jl), SURVEY §8f-1.

`get_rydberg_infidelity` issues its 5 configurations x 6 initial states as TWO batched `ca_rydberg1_sweep` launches
(the points that share a trajectory count go together) instead of 30 serial ensembles; `get_parity_fidelity` evaluates
the 62 832-point global-RZ scan of src/fidelity.jl:166-169 in closed form (16 phase terms) instead of 62 832 products of
25x25 matrices.  Return values and names follow the reference."""
from collections import OrderedDict

import numpy as np

from . import _abi, _lib
from .api import _next_seed, _ode, simulation_czlp
from .config import ket_0, ket_1, model_from_config, tensor
from .gates import Hadamard, Id, RZ, op_tensor

basis_fidelity_states = [ket_0, ket_1, (ket_0 + ket_1) / np.sqrt(2), (ket_0 - ket_1) / np.sqrt(2),
                         (ket_0 + 1j * ket_1) / np.sqrt(2), (ket_0 - 1j * ket_1) / np.sqrt(2)]   # src/fidelity.jl:1-8


def get_rydberg_fidelity_configs(cfg, n_samples=20):
    """src/fidelity.jl:11-64 (key spelling included)."""
    configs = OrderedDict()

    def variant(T1, inter, ryd, noise, free, n):
        c = cfg.copy()
        if T1:
            c.atom_params[1] = 1.0
        c.spontaneous_decay_intermediate, c.spontaneous_decay_rydberg = inter, ryd
        c.laser_noise, c.free_motion, c.n_samples = noise, free, n
        return c

    configs["Intermdeiate state decay"] = variant(True, True, False, False, True, 1)
    configs["Rydberg state decay"] = variant(True, False, True, False, True, 1)
    configs["Laser noise"] = variant(True, False, False, True, False, n_samples)
    configs["Atom motion"] = variant(False, False, False, False, True, n_samples)
    configs["Total"] = variant(False, True, True, True, True, n_samples)
    return configs


def get_rydberg_infidelity(cfg, U=None, states=None, n_samples=100, seed=None, device=None, batched=True, **ode_kwargs):
    """src/fidelity.jl:67-95 -> {name: mean over states of 1 - <ψ_ideal|ρ(t_end)|ψ_ideal>}.

    batched=True: all (configuration, state) points that share n_samples run in one `ca_rydberg1_sweep` launch;
    batched=False: one `ca_rydberg1_run` per point on the SAME global trajectory indices (so both paths agree to rounding)."""
    U = np.eye(5, dtype=np.complex128) if U is None else np.asarray(U, dtype=np.complex128)
    states = basis_fidelity_states if states is None else states
    configs = get_rydberg_fidelity_configs(cfg, n_samples)
    ctx = _lib.default_context(device)
    ode = _ode(ode_kwargs)
    sd = _next_seed(seed)
    points = [(name, c, np.asarray(s, dtype=np.complex128)) for name, c in configs.items() for s in states]
    rho_end = [None] * len(points)
    kw = dict(f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes)
    offset = 0
    for n in sorted({c.n_samples for _, c, _ in points}):
        idx = [i for i, (_, c, _) in enumerate(points) if c.n_samples == n]
        models = [model_from_config(points[i][1]) for i in idx]
        if batched:
            rho, _, _ = ctx.rydberg1_sweep(models, [points[i][2] for i in idx], [points[i][1].tspan[-1] for i in idx], n,
                                           traj_offset=offset, seed=sd, ode=ode, **kw)
            for k, i in enumerate(idx):
                rho_end[i] = rho[k]
        else:
            for k, i in enumerate(idx):
                c = points[i][1]
                noise = kw if c.laser_noise else {}
                rho, _, _ = ctx.rydberg1_run(models[k], [c.tspan[0], c.tspan[-1]], points[i][2], n, traj_offset=offset + k * n, seed=sd,
                                             ode=ode, **noise)
                rho_end[i] = rho[-1]
        offset += n * len(idx)
    out = OrderedDict()
    for name in configs:
        acc = 0.0
        for (nm, _, s), ρ in zip(points, rho_end):
            if nm == name:
                ψ = U @ s
                acc += 1.0 - float(np.real(np.conj(ψ) @ ρ @ ψ))
        out[name] = acc / len(states)
    return out


# ---- CZ parity pipeline ---------------------------------------------------------------------------------------------
_ket_pos = (ket_0 + ket_1) / np.sqrt(2)
_Phi_p = (tensor(ket_0, ket_0) + tensor(ket_1, ket_1)) / np.sqrt(2)
_Had2 = op_tensor(Id, Hadamard)


def _rz_diag(ϕ):
    """diag of RZ(ϕ) ⊗ RZ(ϕ) for an array of angles: [len(ϕ), 25]."""
    ϕ = np.atleast_1d(np.asarray(ϕ, dtype=float))
    d1 = np.ones((len(ϕ), 5), dtype=np.complex128)
    d1[:, 0] = np.exp(-0.5j * ϕ)
    d1[:, 1] = np.exp(0.5j * ϕ)
    return (d1[:, None, :] * d1[:, :, None]).reshape(len(ϕ), 25)   # index a1 + 5 a2


def parity_scan(ρ, ϕ_list):
    """F(ϕ) = <Φ+| Had RZZ(ϕ) ρ (Had RZZ(ϕ))† |Φ+> for all ϕ at once (src/fidelity.jl:166): with h = Had†Φ+ and the diagonal
    d(ϕ) of RZ⊗RZ, F = Σ_IJ d_I conj(h_I) ρ_IJ conj(d_J) h_J -- only the four qubit components of h are non-zero."""
    h = _Had2.conj().T @ _Phi_p
    nz = np.flatnonzero(np.abs(h) > 0)
    d = _rz_diag(ϕ_list)[:, nz]
    M = np.conj(h[nz])[:, None] * np.asarray(ρ)[np.ix_(nz, nz)] * h[nz][None, :]
    return np.real(np.einsum("pi,ij,pj->p", d, M, np.conj(d)))


def get_parity_fidelity(cfg, seed=None, device=None, **ode_kwargs):
    """src/fidelity.jl:147-176 -> (ϕ_list, F_list, ϕ at the maximum)."""
    c = cfg.copy()
    c.ψ0 = tensor(_ket_pos, _ket_pos)
    ρ1 = simulation_czlp(c, seed=seed, device=device, **ode_kwargs)[0][-1]
    ϕ_list = np.arange(0.0, 2 * np.pi, 1e-4)   # 0.0:0.0001:2π
    F = parity_scan(ρ1, ϕ_list)
    return ϕ_list, F, float(ϕ_list[int(np.argmax(F))])


def get_parity_fidelity_temp(ρ, ϕ_RZ):
    """src/fidelity.jl:179-190 -> (F, ρt)."""
    g = op_tensor(RZ(ϕ_RZ), RZ(ϕ_RZ))
    ρt = g @ np.asarray(ρ) @ g.conj().T
    ρt = _Had2 @ ρt @ _Had2.conj().T
    return float(np.real(np.conj(_Phi_p) @ ρt @ _Phi_p)), ρt


def cz_infidelity_points(cfg, n_samples=1):
    """The seven simulation_czlp ensembles behind get_cz_infidelity (src/fidelity.jl:193-240), in the reference's order: the parity-scan
    run that fixes ϕ_RZ (:205, n = 1), the calibration-error run (:225, n = 1), then the five error-budget configurations."""
    c = cfg.copy()
    c.spontaneous_decay_intermediate = c.spontaneous_decay_rydberg = c.laser_noise = False
    c.atom_params[1] = 1.0
    c.n_samples = 1
    pts = [("parity scan", c), ("calibration", c.copy())] + list(get_rydberg_fidelity_configs(cfg, n_samples).items())
    out = []
    for name, cc in pts:
        cc = cc.copy()
        cc.ψ0 = tensor(_ket_pos, _ket_pos)
        out.append((name, cc))
    return out


def get_cz_infidelity(cfg, n_samples=1, seed=None, device=None, batched=True, **ode_kwargs):
    """src/fidelity.jl:193-240 -> ({name: infidelity above the calibration error}, calibration_error).

    batched=True (default): the seven ensembles run as ONE `ca_rydberg2_sweep` launch, trajectories of all of them handed
    out longest first; point p uses the global trajectory indices Σ_{q<p} n_q + i of one Philox key.  batched=False: seven serial
    `ca_rydberg2_run` calls on the SAME indices (so both paths agree to rounding).  ode_kwargs apply to every run."""
    pts = cz_infidelity_points(cfg, n_samples)
    sd = _next_seed(seed)
    ctx = _lib.default_context(device)
    ode = _ode(ode_kwargs)
    kw = dict(f=cfg.f, amp_red=cfg.red_laser_phase_amplitudes, amp_blue=cfg.blue_laser_phase_amplitudes)
    models = [model_from_config(c, two_atom=True) for _, c in pts]
    counts = [int(c.n_samples) for _, c in pts]
    if batched:
        rho_end, _, _ = ctx.rydberg2_sweep(models, [c.ψ0 for _, c in pts], [c.tspan[-1] for _, c in pts], counts, seed=sd, ode=ode, **kw)
    else:
        rho_end, off = [], 0
        for (_, c), m, n in zip(pts, models, counts):
            noise = kw if c.laser_noise else {}
            ρ, _, _ = ctx.rydberg2_run(m, [c.tspan[0], c.tspan[-1]], c.ψ0, n, traj_offset=off, seed=sd, ode=ode, **noise)
            rho_end.append(ρ[-1])
            off += n
    ϕ_list = np.arange(0.0, 2 * np.pi, 1e-4)
    ϕ_RZ = float(ϕ_list[int(np.argmax(parity_scan(rho_end[0], ϕ_list)))])
    A = _Had2 @ op_tensor(RZ(ϕ_RZ), RZ(ϕ_RZ))

    def infid(ρ):
        ρ = A @ ρ @ A.conj().T
        return 1.0 - float(np.real(np.conj(_Phi_p) @ ρ @ _Phi_p))

    calibration_error = infid(rho_end[1])
    out = OrderedDict()
    for (name, _), ρ in zip(pts[2:], rho_end[2:]):
        out[name] = max(infid(ρ) - calibration_error, 0.0)
    return out, calibration_error
