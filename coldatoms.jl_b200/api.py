"""Host-side mirror of the reference's ensemble drivers: same names, argument meaning and return shapes.

Every function here is a thin marshalling layer over one C-ABI call into libcoldatoms_b200.so; the ensemble
loop, the samplers and the ODE solves all run on the GPU.  Density matrices come back as numpy arrays
`ρ[t, row, col]` (the Python analogue of Julia's `Vector{Operator}`: `simulation(cfg)[0][-1]` is the final 5x5
matrix), already divided by `n_samples` like the reference (`src/rydberg_model.jl:264`).

Multi-GPU: when `torch.distributed` is initialised (one process per GPU, NCCL), pass `distributed=True`; each
rank integrates the contiguous slice [rank*N/W, (rank+1)*N/W) of the trajectory index space and the
un-normalised sums are combined with ONE all-reduce (the payload the reference's dead `simulation_mpi`
intended to `MPI.Reduce`, src/rydberg_model.jl:379).
"""
import math
from collections import OrderedDict

import numpy as np

from . import _abi
from . import _lib
from .config import CZLPConfig, RydbergConfig, beam_from_params, model_from_config
from .physics import Ω_twophoton, δ_twophoton

__all__ = [
    "simulation", "simulation_czlp", "simulation_blue_intens", "release_recapture", "samples_generate", "ϕ", "phi",
    "calibrate_two_photon", "get_blockade_stark_shift_factor", "get_rydberg_probs", "get_two_qubit_probs", "seed_b200",
    "Ω", "simple_flattopHG_field", "simple_flattopLG_field", "HG_coeff", "LG_coeff", "last_stats", "blue_intens_model",
]
# Python NFKC-normalises identifiers (ϕ U+03D5 -> φ U+03C6): normalise the export list the same way
__all__ = [__import__("unicodedata").normalize("NFKC", _n) for _n in __all__]

_seed_state = [_abi.DEFAULT_SEED]
_last_stats = {}


def seed_b200(seed):
    """Seed the Philox key sequence used when a call does not pass `seed=` (analogue of `Random.seed!`)."""
    _seed_state[0] = int(seed) & 0xFFFFFFFFFFFFFFFF


def _next_seed(seed):
    if seed is not None:
        return int(seed) & 0xFFFFFFFFFFFFFFFF
    s = _seed_state[0]
    # splitmix64 step: successive calls draw independent ensembles, like Julia's advancing global RNG
    z = (s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    _seed_state[0] = z
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def last_stats():
    """ca_stats of the most recent ensemble call (RHS evaluations, accepted/rejected steps, kernel ms)."""
    return dict(_last_stats)


def _ode(ode_kwargs):
    kw = dict(ode_kwargs)
    alg = kw.pop("alg", "DP5")
    if str(alg).rstrip("()") not in ("DP5", "Tsit5"):
        raise ValueError("only the embedded 5(4) pair is implemented (alg=DP5; Tsit5 is accepted as an alias)")
    allowed = {"reltol", "abstol", "dt", "dtmax", "maxiters", "adaptive", "dense"}
    bad = set(kw) - allowed
    if bad:
        raise TypeError(f"unsupported ode_kwargs: {sorted(bad)} (supported: {sorted(allowed)} and alg)")
    return _abi.ode_opts(**kw)


def _shard(n, distributed):
    """Contiguous slice of the trajectory index space for this rank."""
    if not distributed:
        return 0, n, 1
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("distributed=True needs an initialised torch.distributed process group")
    r, w = dist.get_rank(), dist.get_world_size()
    lo, hi = (n * r) // w, (n * (r + 1)) // w
    return lo, hi - lo, w


def _allreduce_mean(ctx, run, nt, d, n_total):
    """Run with device-resident un-normalised sums, all-reduce them over NCCL, normalise.

    The 8 shard-additive counters of the call (ca_stats_to_device: RHS / step counts and the two failure flags) ride at the tail of
    the SAME buffer, so one all-reduce carries the sums and the status: a trajectory that hit maxiters or went non-finite on ANY rank
    raises on EVERY rank, exactly like the single-GPU call does (such a trajectory is left out of the sums)."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", ctx.device)
    n_out = 2 * nt * d * d * 2
    flat = torch.zeros(n_out + 8, dtype=torch.float64, device=dev)
    buf = flat[:n_out].view(2, nt, d, d, 2)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    try:
        st = run(_abi.CA_OUT_SUMS | _abi.CA_OUT_DEVICE, (buf[0].data_ptr(), buf[1].data_ptr()))
        ctx.stats_to_device(flat[n_out:].data_ptr())
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        host = flat.cpu().numpy()
    finally:
        ctx.reset_stream()
    cnt = host[n_out:]
    st = dict(st, n_traj=int(cnt[7]), n_rhs=int(cnt[0]), n_accept=int(cnt[1]), n_reject=int(cnt[2]), n_sample_attempts=int(cnt[3]),
              n_exact_steps=int(cnt[6]))
    if cnt[5] > 0:
        raise _lib.ColdAtomsError(_abi.CA_ERR_NONFINITE, "a trajectory produced a non-finite state or the step size underflowed (on some rank)")
    if cnt[4] > 0:
        raise _lib.ColdAtomsError(_abi.CA_ERR_MAXITERS, "a trajectory hit maxiters (on some rank)")
    out = host[:n_out].reshape(2, nt, d, d, 2) / float(n_total)
    cm = out[..., 0] + 1j * out[..., 1]
    return cm[0].transpose(0, 2, 1).copy(), cm[1].transpose(0, 2, 1).copy(), st


# ---------------------------------------------------------------------------------------------------------
# samplers (exported helpers, docs/src/examples.md:139)
# ---------------------------------------------------------------------------------------------------------
def samples_generate(trap_params, atom_params, N, freq=10, skip=1000, harmonic=True, eps=1e-2, seed=None, device=None, n_chains=0):
    """`samples_generate` (src/atom_sampler.jl:80-120) -> (samples[N,6], acc_rate).

    harmonic=True is the on-device Philox sampler (MvNormal + energy rejection, :84-96).  harmonic=False is the Metropolis
    branch (:97-120) run as `n_chains` independent chains, one per GPU thread (each with its own `skip` burn-in, emitting
    every `freq`-th state); n_chains=1 is the reference's single sequential chain, the default fills the GPU.
    """
    ctx = _lib.default_context(device)
    if not harmonic:
        return ctx.sample_atoms_metropolis(trap_params, atom_params, int(N), freq=freq, skip=skip, harmonic=False, eps=eps,
                                           n_chains=n_chains, seed=_next_seed(seed))
    out, _ = ctx.sample_atoms(trap_params, atom_params, int(N), seed=_next_seed(seed), eps=eps)
    return out, 1


def ϕ(tspan, f, amplitudes, seed=None, device=None):
    """Phase-noise trajectory `ϕ(tspan, f, amplitudes)` (src/lasernoise_sampler.jl:31-37) on a uniform grid from 0."""
    ctx = _lib.default_context(device)
    return ctx.phase_noise(tspan, f, amplitudes, 1, seed=_next_seed(seed))[0]


phi = ϕ


def Ω(x, y, z, laser_params, device=None):
    """Complex Rabi frequency of a beam at points (x,y,z), evaluated by the device beam code
    (`Ω`, src/rydberg_model.jl:51-74)."""
    ctx = _lib.default_context(device)
    xyz = np.stack(np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float), np.asarray(z, float)), -1)
    return ctx.eval_beam(beam_from_params(laser_params), xyz.reshape(-1, 3)).reshape(xyz.shape[:-1])


def simple_flattopHG_field(x, y, z, laser_params, device=None):
    """src/arbitrary_beams.jl:50-67 (7-element laser_params)."""
    ctx = _lib.default_context(device)
    xyz = np.stack(np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float), np.asarray(z, float)), -1)
    p = list(laser_params)
    p[3] = "simp_flattop_HG"
    return ctx.eval_beam(beam_from_params(p, arb_bms=True), xyz.reshape(-1, 3)).reshape(xyz.shape[:-1])


def simple_flattopLG_field(x, y, z, laser_params, device=None):
    """src/arbitrary_beams.jl:69-81 (6-element laser_params)."""
    ctx = _lib.default_context(device)
    xyz = np.stack(np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float), np.asarray(z, float)), -1)
    p = list(laser_params)[:6]
    p[3] = "simp_flattop_LG"
    return ctx.eval_beam(beam_from_params(p, arb_bms=True), xyz.reshape(-1, 3)).reshape(xyz.shape[:-1])


def LG_coeff(k, n):
    """src/arbitrary_beams.jl:24-30."""
    res = sum(math.factorial(i) / (2 ** i * math.factorial(i - k)) for i in range(k, n + 1))
    return (-1.0) ** k * res / math.factorial(k)


def HG_coeff(k, n):
    """src/arbitrary_beams.jl:32-38."""
    return sum(math.factorial(2 * i) / (math.factorial(i) * 2 ** (3 * i) * math.factorial(i - k) * math.factorial(2 * k))
               for i in range(k, n + 1))


# ---------------------------------------------------------------------------------------------------------
# one-atom ensemble
# ---------------------------------------------------------------------------------------------------------
def calibrate_two_photon(cfg, n_samples=1000, seed=None, device=None, samples=None):
    """`calibrate_two_photon` (src/rydberg_model.jl:285-306).  The two sample loops (:293, :300) are device reductions
    (`ca_thermal_average`, CA_AVG_TWOPHOTON) over the same thermal ensemble: Philox-drawn on the device, or host `samples[n,6]`."""
    from .config import model_from_config
    ctx = _lib.default_context(device)
    sd = _next_seed(seed)
    c = cfg.copy()
    Ωr, Ωb = cfg.red_laser_params[0], cfg.blue_laser_params[0]
    n = n_samples if samples is None else len(samples)
    sums, _ = ctx.thermal_average(_abi.CA_AVG_TWOPHOTON, model_from_config(c), n, seed=sd, samples1=samples)
    factor = math.sqrt((sums[0] / n) / Ω_twophoton(Ωr, Ωb, cfg.detuning_params[0]))
    c.red_laser_params[0] = Ωr / factor
    c.blue_laser_params[0] = Ωb / factor
    sums, _ = ctx.thermal_average(_abi.CA_AVG_TWOPHOTON, model_from_config(c), n, seed=sd, samples1=samples)
    c.detuning_params[1] += sums[1] / n - δ_twophoton(Ωr, Ωb, cfg.detuning_params[0])
    return c


def simulation(cfg: RydbergConfig, temperature_calibrate=False, seed=None, device=None, distributed=False, samples=None,
               phases_red=None, phases_blue=None, compat=_abi.CA_COMPAT_DEFAULT, rho2_mode=_abi.CA_RHO2_MATSQ, **ode_kwargs):
    """`simulation(cfg; temperature_calibrate=false, ode_kwargs...)` (src/rydberg_model.jl:194-265) -> (ρ, ρ2).

    Extra keywords (not in the reference): `seed`, `device`, `distributed`, host-supplied `samples[N,6]` /
    `phases_red|blue[N,Nf]` (the parity protocol of SURVEY §8c), `compat`, `rho2_mode`.
    """
    sd = _next_seed(seed)
    # the calibration draws its own thermal ensemble (rydberg_model.jl:289 vs :206-212): a sub-seed derived from the call's seed, so
    # that it is not the first 1000 atoms of the simulated ensemble
    cfg_t = calibrate_two_photon(cfg, seed=(sd ^ 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF, device=device) if temperature_calibrate else cfg
    model = model_from_config(cfg_t, compat=compat, rho2_mode=rho2_mode)
    ctx = _lib.default_context(device)
    n = int(cfg_t.n_samples)
    lo, cnt, world = _shard(n, distributed)
    kw = dict(traj_offset=lo, n_total=n, seed=sd,
              samples=None if samples is None else np.asarray(samples)[lo:lo + cnt],
              phases_red=None if phases_red is None else np.asarray(phases_red)[lo:lo + cnt],
              phases_blue=None if phases_blue is None else np.asarray(phases_blue)[lo:lo + cnt],
              f=cfg_t.f if cfg_t.laser_noise else None,
              amp_red=cfg_t.red_laser_phase_amplitudes if cfg_t.laser_noise else None,
              amp_blue=cfg_t.blue_laser_phase_amplitudes if cfg_t.laser_noise else None, ode=_ode(ode_kwargs))
    if not distributed:
        ρ, ρ2, st = ctx.rydberg1_run(model, cfg_t.tspan, cfg_t.ψ0, cnt, **kw)
    else:
        ρ, ρ2, st = _allreduce_mean(
            ctx, lambda fl, dev: ctx.rydberg1_run(model, cfg_t.tspan, cfg_t.ψ0, cnt, out_flags=fl, dev_out=dev, **kw)[2],
            len(cfg_t.tspan), 5, n)
    _last_stats.clear()
    _last_stats.update(st)
    return ρ, ρ2


def blue_intens_model(atom_params, trap_params, red_laser_params, blue_laser_params, detuning_params, decay_params,
                      atom_motion=True, free_motion=True, spontaneous_decay=True, parallel=False):
    """The ca_model behind `simulation_blue_intens` (src/rydberg_model_arb_bms.jl:15-86): uniform red, flat-top HG/LG blue,
    Doppler from vz only, two decay channels, no phase noise.  The reference file uses symbols that do not exist in the
    snapshot; the semantics implemented are the ones inferred in SURVEY App. B7 (Δ=k_r vz, δ=(k_r ± k_b) vz, jumps
    sqrt(Γg)|1><p| and sqrt(Γgt)|0><p|)."""
    m = _abi.ca_model()
    m.atom_m, m.atom_T = map(float, atom_params)
    m.trap_U0, m.trap_w0, m.trap_z0 = map(float, trap_params)
    m.red = beam_from_params(list(red_laser_params)[:3], arb_bms=True)
    m.blue = beam_from_params(blue_laser_params, arb_bms=True)
    m.Delta0, m.delta0 = map(float, detuning_params)
    Γg, Γgt = decay_params if spontaneous_decay else (0.0, 0.0)
    m.Gamma1, m.Gamma0, m.Gammal, m.Gammar = float(Γg), float(Γgt), 0.0, 0.0
    m.atom_motion, m.free_motion, m.laser_noise = int(atom_motion), int(free_motion), 0
    m.decay_intermediate, m.decay_rydberg = 1, 0
    m.doppler_kind, m.parallel = _abi.CA_DOPPLER_Z, int(parallel)
    m.detuning_sign, m.compat, m.rho2_mode, m.n_noise_nodes = _abi.CA_SIGN_RYDBERG1, 0, _abi.CA_RHO2_MATSQ, 1001
    return m


def simulation_blue_intens(tspan, ψ0, atom_params, trap_params, samples, red_laser_params, blue_laser_params, detuning_params,
                           decay_params, atom_motion=True, free_motion=True, spontaneous_decay=True, parallel=False, device=None,
                           **ode_kwargs):
    """`simulation_blue_intens` (src/rydberg_model_arb_bms.jl:15-86) -> (ρ_mean, ρ2_mean) over the host-supplied samples."""
    m = blue_intens_model(atom_params, trap_params, red_laser_params, blue_laser_params, detuning_params, decay_params,
                          atom_motion, free_motion, spontaneous_decay, parallel)
    samples = np.asarray(samples, dtype=float).reshape(-1, 6)
    ctx = _lib.default_context(device)
    ρ, ρ2, st = ctx.rydberg1_run(m, tspan, ψ0, len(samples), samples=samples, ode=_ode(ode_kwargs))
    _last_stats.clear()
    _last_stats.update(st)
    return ρ, ρ2


# ---------------------------------------------------------------------------------------------------------
# two-atom CZ ensemble
# ---------------------------------------------------------------------------------------------------------
def simulation_czlp(cfg: CZLPConfig, seed=None, device=None, distributed=False, samples1=None, samples2=None, phases4=None,
                    compat=_abi.CA_COMPAT_DEFAULT, rho2_mode=_abi.CA_RHO2_MATSQ, **ode_kwargs):
    """`simulation_czlp(cfg; ode_kwargs...)` (src/cz_model.jl:107-171) -> (ρ, ρ2), 25x25 per time point."""
    model = model_from_config(cfg, two_atom=True, compat=compat, rho2_mode=rho2_mode)
    ctx = _lib.default_context(device)
    n = int(cfg.n_samples)
    lo, cnt, world = _shard(n, distributed)
    sl = slice(lo, lo + cnt)
    kw = dict(traj_offset=lo, n_total=n, seed=_next_seed(seed),
              samples1=None if samples1 is None else np.asarray(samples1)[sl],
              samples2=None if samples2 is None else np.asarray(samples2)[sl],
              phases4=None if phases4 is None else [np.asarray(p)[sl] for p in phases4],
              f=cfg.f if cfg.laser_noise else None,
              amp_red=cfg.red_laser_phase_amplitudes if cfg.laser_noise else None,
              amp_blue=cfg.blue_laser_phase_amplitudes if cfg.laser_noise else None, ode=_ode(ode_kwargs))
    if not distributed:
        ρ, ρ2, st = ctx.rydberg2_run(model, cfg.tspan, cfg.ψ0, cnt, **kw)
    else:
        ρ, ρ2, st = _allreduce_mean(
            ctx, lambda fl, dev: ctx.rydberg2_run(model, cfg.tspan, cfg.ψ0, cnt, out_flags=fl, dev_out=dev, **kw)[2],
            len(cfg.tspan), 25, n)
    _last_stats.clear()
    _last_stats.update(st)
    return ρ, ρ2


def get_blockade_stark_shift_factor(trap_params, atom_params, atom_centers, Ω, c6, n_samples=10000, seed=None, device=None,
                                    samples1=None, samples2=None):
    """src/cz_model.jl:77-104: thermal mean of R^-6 -> -Ω / (2 c6 <R^-6>); the map/mean of :101 is a device reduction
    (`ca_thermal_average`, CA_AVG_RM6) over Philox-drawn pairs (streams atom1/atom2) or host `samples1/2[n,6]`."""
    ctx = _lib.default_context(device)
    m = _abi.ca_model()
    m.atom_m, m.atom_T = map(float, atom_params)
    m.trap_U0, m.trap_w0, m.trap_z0 = map(float, trap_params)
    m.red.kind = m.blue.kind = _abi.CA_BEAM_UNIFORM
    m.blue.w0 = m.blue.z0 = 1.0
    m.center1[:] = [float(v) for v in atom_centers[0]]
    m.center2[:] = [float(v) for v in atom_centers[1]]
    n = n_samples if samples1 is None else len(samples1)
    sums, _ = ctx.thermal_average(_abi.CA_AVG_RM6, m, n, seed=_next_seed(seed), samples1=samples1, samples2=samples2)
    return -Ω / (2.0 * c6 * (sums[0] / n))


# ---------------------------------------------------------------------------------------------------------
# release and recapture
# ---------------------------------------------------------------------------------------------------------
def release_recapture(tspan, trap_params, atom_params, N, freq=10, skip=1000, eps=1e-3, harmonic=True, seed=None, device=None,
                      distributed=False):
    """`release_recapture` (src/basic_experiments.jl:72-81) -> (probs[nt], acc_rate).  harmonic=False draws the atoms with the
    multi-chain Metropolis sampler (sampler eps stays 1e-2, :73) and feeds them to the recapture kernel."""
    ctx = _lib.default_context(device)
    n = int(N)
    lo, cnt, world = _shard(n, distributed)
    sd = _next_seed(seed)
    if not harmonic:
        # Metropolis branch: every rank runs its own set of independent chains for its share of the N atoms (the multi-chain sampler is
        # already a decomposition of the reference's single chain; ranks differ by the Philox key) and the counts are all-reduced
        rk = 0
        if distributed:
            import torch.distributed as dist
            rk = dist.get_rank()
        smp, acc = ctx.sample_atoms_metropolis(trap_params, atom_params, cnt, freq=freq, skip=skip, harmonic=False,
                                               seed=(sd + 0x632BE59BD9B4E019 * rk) & 0xFFFFFFFFFFFFFFFF)
        if not distributed:
            probs, _, st = ctx.release_recapture(trap_params, atom_params, tspan, n, eps=eps, seed=sd, samples=smp)
            _last_stats.clear()
            _last_stats.update(st)
            return probs, acc
        probs, st = _allreduce_counts(ctx, lambda fl, dev: ctx.release_recapture(trap_params, atom_params, tspan, cnt, n_total=n, eps=eps, seed=sd,
                                                                                 samples=smp, out_flags=fl, dev_out=dev)[2], len(tspan), n, extra=acc * cnt)
        _last_stats.clear()
        _last_stats.update(st)
        return probs[:-1], float(probs[-1])   # acceptance rate: mean over the ranks' chains, weighted by their sample counts
    if not distributed:
        probs, _, st = ctx.release_recapture(trap_params, atom_params, tspan, cnt, traj_offset=lo, n_total=n, eps=eps, seed=sd)
    else:
        probs, st = _allreduce_counts(ctx, lambda fl, dev: ctx.release_recapture(trap_params, atom_params, tspan, cnt, traj_offset=lo, n_total=n,
                                                                                 eps=eps, seed=sd, out_flags=fl, dev_out=dev)[2], len(tspan), n)
        probs = probs[:-1]
    _last_stats.clear()
    _last_stats.update(st)
    return probs, 1.0


def _allreduce_counts(ctx, run, nt, n_total, extra=0.0):
    """Recapture counts stay on the device (CA_OUT_SUMS | CA_OUT_DEVICE), one NCCL all-reduce, then the division by N
    (`recapture ./ N`, basic_experiments.jl:80).  `extra` rides along as an additional shard-additive slot."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", ctx.device)
    buf = torch.zeros(nt + 1, dtype=torch.float64, device=dev)
    buf[nt] = float(extra)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    try:
        st = run(_abi.CA_OUT_SUMS | _abi.CA_OUT_DEVICE, buf.data_ptr())
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        out = buf.cpu().numpy() / float(n_total)
    finally:
        ctx.reset_stream()
    return out, st


# ---------------------------------------------------------------------------------------------------------
# host post-processing on the returned ρ (unchanged semantics, src/utilities.jl:82-117)
# ---------------------------------------------------------------------------------------------------------
def get_rydberg_probs(ρ, ρ2, eps=1e-12):
    out = OrderedDict()
    ρ, ρ2 = np.asarray(ρ), np.asarray(ρ2)
    for i, name in enumerate(["0", "1", "r", "p", "l"]):
        P, P2 = np.real(ρ[:, i, i]), np.real(ρ2[:, i, i])
        out["P" + name] = P
        out["S" + name] = np.sqrt(P2 - P ** 2 + eps) / len(ρ)  # divides by length(ρ) = nt
    return out


def get_two_qubit_probs(ρ, ρ2, eps=1e-12):
    out = OrderedDict()
    ρ, ρ2 = np.asarray(ρ), np.asarray(ρ2)
    for name, (a, b) in zip(["00", "01", "10", "11"], [(0, 0), (0, 1), (1, 0), (1, 1)]):
        i = a + 5 * b
        P, P2 = np.real(ρ[:, i, i]), np.real(ρ2[:, i, i])
        out["P" + name] = P
        out["S" + name] = np.sqrt(P2 - P ** 2 + eps) / len(ρ)
    return out
