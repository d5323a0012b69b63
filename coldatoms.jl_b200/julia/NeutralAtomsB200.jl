# NeutralAtomsB200.jl -- drop-in `ccall` shim: the reference's ensemble drivers on top of libcoldatoms_b200.so.
#
# Usage (in a Julia session that has the reference package loaded):
#
#     using NeutralAtoms                      # the unmodified reference (module name of this snapshot)
#     include("NeutralAtomsB200.jl")          # defines module NeutralAtomsB200
#     using .NeutralAtomsB200
#     NeutralAtomsB200.init!("/path/to/libcoldatoms_b200.so"; device = 0)
#     ρ, ρ2 = NeutralAtomsB200.simulation(cfg)            # same signature/return as NeutralAtoms.simulation
#     NeutralAtomsB200.override!()                          # optional: make NeutralAtoms.simulation & co. call the GPU path
#
# Every function keeps the reference's signature, argument meaning and return shapes:
#   simulation(cfg::RydbergConfig; temperature_calibrate=false, ode_kwargs...)   src/rydberg_model.jl:194-265
#   simulation_czlp(cfg::CZLPConfig; ode_kwargs...)                              src/cz_model.jl:107-171
#   simulation_blue_intens(tspan, ψ0, atom_params, trap_params, samples, ...)    src/rydberg_model_arb_bms.jl:15-86
#   release_recapture(tspan, trap_params, atom_params, N; ...)                   src/basic_experiments.jl:72-81
#   samples_generate(trap_params, atom_params, N; harmonic=true, ...)            src/atom_sampler.jl:80-96
#   ϕ(tspan, f, amplitudes)                                                      src/lasernoise_sampler.jl:31-37
#   calibrate_two_photon(cfg, n_samples=1000)                                    src/rydberg_model.jl:285-306
#   get_blockade_stark_shift_factor(trap_params, atom_params, atom_centers, Ω, c6, n_samples=10000)   src/cz_model.jl:77-104
# The ensemble loop (rydberg_model.jl:260-262, cz_model.jl:151), the samplers and every master_dynamic call run
# inside ONE ccall.  There is no CUDA.jl fallback and no CPU fallback: if the library or a B200 is missing, init!
# throws.  Julia is not available in the build container of this repository, so this file is exercised off-box
# only (tests/julia_parity.jl); the same C ABI is driven by the Python mirror in the test-suite.
module NeutralAtomsB200

using NeutralAtoms
using QuantumOptics
using Libdl

export simulation, simulation_czlp, simulation_blue_intens, release_recapture, samples_generate, ϕ,
       calibrate_two_photon, get_blockade_stark_shift_factor

# ---- mirror of include/coldatoms_b200.h ---------------------------------------------------------------
struct CaBeam
    Omega0::Float64; w0::Float64; z0::Float64; theta::Float64; n::Float64; sqz::Float64
    kind::Int32; hg_n::Int32; hg_m::Int32; lg_l::Int32; _pad::Int32
end
struct CaModel
    atom_m::Float64; atom_T::Float64
    trap_U0::Float64; trap_w0::Float64; trap_z0::Float64
    red::CaBeam; blue::CaBeam
    Delta0::Float64; delta0::Float64
    Gamma0::Float64; Gamma1::Float64; Gammal::Float64; Gammar::Float64
    atom_motion::Int32; free_motion::Int32; laser_noise::Int32
    decay_intermediate::Int32; decay_rydberg::Int32
    doppler_kind::Int32; parallel::Int32; detuning_sign::Int32
    compat::Int32; rho2_mode::Int32; n_noise_nodes::Int32; _pad::Int32
    center1::NTuple{3,Float64}; center2::NTuple{3,Float64}
    c6::Float64; xi::Float64
    two_atom_flags::Int32; _pad2::Int32
    red2::CaBeam; blue2::CaBeam
    Delta0_2::Float64; delta0_2::Float64; V_rr::Float64; seg2_phase::Float64
end
struct CaOde
    reltol::Float64; abstol::Float64; dt::Float64; dtmax::Float64
    maxiters::Int64; adaptive::Int32; dense::Int32
end
mutable struct CaStats
    n_traj::Int64; n_rhs::Int64; n_accept::Int64; n_reject::Int64; n_sample_attempts::Int64
    kernel_ms::Float64; total_ms::Float64; noise_ms::Float64
    n_launches::Int32; status::Int32
    n_exact_steps::Int64
    CaStats() = new(0, 0, 0, 0, 0, 0.0, 0.0, 0.0, 0, 0, 0)
end

const BEAM_GAUSS, BEAM_HG, BEAM_LG, BEAM_UNIFORM = Int32(0), Int32(1), Int32(2), Int32(3)
const DOPPLER_XZ, DOPPLER_Z = Int32(0), Int32(1)
const SIGN_RYDBERG1, SIGN_CZ = Int32(0), Int32(1)
const COMPAT_DEFAULT = Int32(1)
const TWOATOM_PER_ATOM_LASERS, TWOATOM_CONST_VRR, TWOATOM_SHIFT_AFTER_MOTION, TWOATOM_DIRECT = Int32(1), Int32(2), Int32(4), Int32(8)   # CA_2A_*

# The rydberg_model.jl / cz_model.jl parameterisation (29 leading fields): the legacy two-atom fields of two_atom_model.jl are zeroed.
const NOBEAM = CaBeam(0.0, 1.0, 1.0, 0.0, 1.0, 1.0, BEAM_GAUSS, 0, 0, 0, 0)
CaModel(a::Vararg{Any,29}) = CaModel(a..., Int32(0), Int32(0), NOBEAM, NOBEAM, 0.0, 0.0, 0.0, 0.0)

const LIB = Ref{Ptr{Cvoid}}(C_NULL)
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
const MULTI = Ref{Ptr{Cvoid}}(C_NULL)   # ca_multi* when several GPUs are driven from this process (init!(...; devices = ...))
const SEED = Ref{UInt64}(0xC01DA705)

sym(name) = Libdl.dlsym(LIB[], name)
lasterr() = unsafe_string(ccall(sym(:ca_last_error), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("coldatoms_b200 error $rc: $(lasterr())")

"""
Load the library and create the context on `device`. Throws if there is no sm_100 GPU (no fallback).
`devices = :all` (every visible GPU) or a list of device ordinals additionally creates the multi-GPU handle (`ca_init_multi`: one context
per GPU in THIS process and an NCCL clique): `simulation`, `simulation_czlp`, `release_recapture` and the batched `get_rydberg_infidelity`
then shard their trajectories (grid points) over those GPUs inside the one `ccall` they make, with one `ncclAllReduce` of the sums.
"""
function init!(path::AbstractString; device::Integer = 0, devices = nothing)
    LIB[] = Libdl.dlopen(path)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall(sym(:ca_init), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ctx))
    CTX[] = ctx[]
    if devices !== nothing
        devs = devices === :all ? Cint[] : Cint.(collect(devices))
        m = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve devs check(ccall(sym(:ca_init_multi), Cint, (Ptr{Cint}, Int32, Ref{Ptr{Cvoid}}),
                                      isempty(devs) ? Ptr{Cint}(C_NULL) : pointer(devs), length(devs), m))
        MULTI[] = m[]
    end
    return nothing
end
function finalize!()
    MULTI[] != C_NULL && ccall(sym(:ca_finalize_multi), Cint, (Ptr{Cvoid},), MULTI[])
    CTX[] != C_NULL && ccall(sym(:ca_finalize), Cint, (Ptr{Cvoid},), CTX[])
    MULTI[] = C_NULL; CTX[] = C_NULL
    return nothing
end
# entry point and handle of an ensemble call: the *_multi functions take the argument list of their single-GPU counterparts
entry(single::Symbol, multi::Symbol) = MULTI[] == C_NULL ? (sym(single), CTX[]) : (sym(multi), MULTI[])
seed!(s::Integer) = (SEED[] = UInt64(s); nothing)
function next_seed()   # successive calls draw independent ensembles, like Julia's advancing global RNG
    SEED[] += 0x9E3779B97F4A7C15
    z = SEED[]
    z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9
    z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB
    return z ⊻ (z >> 31)
end

# Ω dispatch on the laser_params vector (src/rydberg_model.jl:51-74; Ω_red/Ω_blue of rydberg_model_arb_bms.jl:1-13)
function beam(p; arb_bms = false)
    Ω0, w0, z0 = Float64(p[1]), Float64(p[2]), Float64(p[3])
    if length(p) == 3
        return CaBeam(Ω0, w0, z0, 0.0, 1.0, 1.0, arb_bms ? BEAM_UNIFORM : BEAM_GAUSS, 0, 0, 0, 0)
    elseif length(p) == 5
        return CaBeam(Ω0, w0, z0, Float64(p[4]), Float64(p[5]), 1.0, BEAM_GAUSS, 0, 0, 0, 0)
    elseif length(p) in (6, 7)
        tag, n, m = p[4], Int32(p[5]), Int32(p[6])
        if n == 0 && m == 0 && !arb_bms
            return CaBeam(Ω0, w0, z0, 0.0, 1.0, 1.0, BEAM_GAUSS, 0, 0, 0, 0)
        elseif tag == "simp_flattop_LG"
            return CaBeam(Ω0, w0, z0, 0.0, 1.0, 1.0, BEAM_LG, 0, 0, n, 0)
        elseif tag == "simp_flattop_HG" && length(p) == 7
            return CaBeam(Ω0, w0, z0, 0.0, 1.0, Float64(p[7]), BEAM_HG, n, m, 0, 0)
        end
    end
    error("unsupported laser_params $p")   # the reference silently returns `nothing` here
end

function model(cfg; two_atom = false)
    c1 = two_atom ? Tuple(Float64.(cfg.atom_centers[1])) : (0.0, 0.0, 0.0)
    c2 = two_atom ? Tuple(Float64.(cfg.atom_centers[2])) : (0.0, 0.0, 0.0)
    CaModel(cfg.atom_params[1], cfg.atom_params[2], cfg.trap_params[1], cfg.trap_params[2], cfg.trap_params[3],
            beam(cfg.red_laser_params), beam(cfg.blue_laser_params), cfg.detuning_params[1], cfg.detuning_params[2],
            cfg.decay_params[1], cfg.decay_params[2], cfg.decay_params[3], cfg.decay_params[4],
            cfg.atom_motion, cfg.free_motion, cfg.laser_noise, cfg.spontaneous_decay_intermediate, cfg.spontaneous_decay_rydberg,
            DOPPLER_XZ, 0, two_atom ? SIGN_CZ : SIGN_RYDBERG1, COMPAT_DEFAULT, 0, 1001, 0, c1, c2,
            two_atom ? cfg.c6 : 0.0, two_atom ? cfg.ξ : 0.0)
end

# `ode_kwargs...` are forwarded verbatim to OrdinaryDiffEq by the reference (rydberg_model.jl:255); the subset the
# library understands: reltol, abstol, dt, dtmax, maxiters, adaptive (alg = DP5, Tsit5 accepted as alias)
function ode(; reltol = 1e-6, abstol = 1e-8, dt = 0.0, dtmax = 0.0, maxiters = 100000, adaptive = true, alg = nothing, kw...)
    isempty(kw) || error("unsupported ode_kwargs: $(keys(kw))")
    CaOde(reltol, abstol, dt, dtmax, maxiters, adaptive, 1)
end

# column-major d×d complex blocks -> Vector{Operator} on basis b (what master_dynamic returns)
function operators(buf::Vector{ComplexF64}, d::Int, nt::Int, b)
    [Operator(b, b, reshape(buf[(j-1)*d*d+1:j*d*d], d, d)) for j in 1:nt]
end

null() = Ptr{Float64}(C_NULL)

function run1(m::CaModel, tspan, ψ0data, n; samples = nothing, seed = next_seed(), f = nothing, ar = nothing, ab = nothing, o = ode())
    nt = length(tspan)
    ρ = zeros(ComplexF64, 25 * nt); ρ2 = zeros(ComplexF64, 25 * nt)
    st = CaStats()
    smp = samples === nothing ? null() : pointer(samples)
    nf = f === nothing ? 0 : length(f)
    fp, h = entry(:ca_rydberg1_run, :ca_rydberg1_run_multi)
    GC.@preserve samples f ar ab begin
        check(ccall(fp, Cint,
            (Ptr{Cvoid}, Ref{CaModel}, Ptr{Float64}, Int32, Ptr{ComplexF64}, Int64, Int64, Int64, UInt64, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ref{CaOde}, Int32, Ptr{ComplexF64}, Ptr{ComplexF64}, Ref{CaStats}),
            h, m, Float64.(tspan), nt, ComplexF64.(ψ0data), n, 0, n, seed, smp, null(), null(),
            f === nothing ? null() : pointer(f), ar === nothing ? null() : pointer(ar), ab === nothing ? null() : pointer(ab),
            nf, o, 0, ρ, ρ2, st))
    end
    return ρ, ρ2, st
end

# un-normalised device-reduced sums of the thermal-average functionals (ca_thermal_average)
function thermal_sums(kind::Integer, m::CaModel, n::Integer; seed = next_seed())
    sums = zeros(Float64, 2)
    att = Ref{Int64}(0)
    check(ccall(sym(:ca_thermal_average), Cint,
        (Ptr{Cvoid}, Int32, Ref{CaModel}, Int64, Int64, UInt64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}),
        CTX[], kind, m, n, 0, seed, 1e-2, null(), null(), sums, att))
    return sums
end

"`calibrate_two_photon(cfg, n_samples=1000)` -- src/rydberg_model.jl:285-306; the two sample loops (:293, :300) run on the device"
function calibrate_two_photon(cfg::NeutralAtoms.RydbergConfig, n_samples = 1000)
    c = deepcopy(cfg)
    Ωr, Ωb = cfg.red_laser_params[1], cfg.blue_laser_params[1]
    seed = next_seed()                      # one thermal ensemble for both loops, as in the reference
    s = thermal_sums(0, model(c), n_samples; seed = seed)
    factor = sqrt((s[1] / n_samples) / NeutralAtoms.Ω_twophoton(Ωr, Ωb, cfg.detuning_params[1]))
    c.red_laser_params[1] = Ωr / factor
    c.blue_laser_params[1] = Ωb / factor
    s = thermal_sums(0, model(c), n_samples; seed = seed)
    c.detuning_params[2] += s[2] / n_samples - NeutralAtoms.δ_twophoton(Ωr, Ωb, cfg.detuning_params[1])
    return c
end

"`get_blockade_stark_shift_factor(trap_params, atom_params, atom_centers, Ω, c6, n_samples=10000)` -- src/cz_model.jl:77-104"
function get_blockade_stark_shift_factor(trap_params, atom_params, atom_centers, Ω, c6, n_samples = 10000)
    b = CaBeam(0.0, 1.0, 1.0, 0.0, 1.0, 1.0, BEAM_UNIFORM, 0, 0, 0, 0)
    m = CaModel(atom_params[1], atom_params[2], trap_params[1], trap_params[2], trap_params[3], b, b, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
                1, 1, 0, 0, 0, DOPPLER_XZ, 0, SIGN_CZ, COMPAT_DEFAULT, 0, 1001, 0,
                Tuple(Float64.(atom_centers[1])), Tuple(Float64.(atom_centers[2])), Float64(c6), 0.0)
    Rm6 = thermal_sums(1, m, n_samples)[1] / n_samples
    return -Ω / (2.0 * c6 * Rm6)
end

"`simulation(cfg::RydbergConfig; temperature_calibrate=false, ode_kwargs...)` -- src/rydberg_model.jl:194-265"
function simulation(cfg::NeutralAtoms.RydbergConfig; temperature_calibrate = false, ode_kwargs...)
    cfg_t = temperature_calibrate ? calibrate_two_photon(cfg) : deepcopy(cfg)
    noise = cfg_t.laser_noise
    ρ, ρ2, _ = run1(model(cfg_t), cfg_t.tspan, cfg_t.ψ0.data, cfg_t.n_samples;
                    f = noise ? cfg_t.f : nothing, ar = noise ? cfg_t.red_laser_phase_amplitudes : nothing,
                    ab = noise ? cfg_t.blue_laser_phase_amplitudes : nothing, o = ode(; ode_kwargs...))
    b = cfg_t.ψ0.basis
    return operators(ρ, 5, length(cfg_t.tspan), b), operators(ρ2, 5, length(cfg_t.tspan), b)
end

"`simulation_blue_intens(...)` -- src/rydberg_model_arb_bms.jl:15-86 (semantics of SURVEY App. B7)"
function simulation_blue_intens(tspan, ψ0, atom_params, trap_params, samples, red_laser_params, blue_laser_params,
                                detuning_params, decay_params; atom_motion = true, free_motion = true,
                                spontaneous_decay = true, parallel = false)
    Γg, Γgt = spontaneous_decay ? decay_params : [0.0, 0.0]
    m = CaModel(atom_params[1], atom_params[2], trap_params[1], trap_params[2], trap_params[3],
                beam(red_laser_params[1:3]; arb_bms = true), beam(blue_laser_params; arb_bms = true),
                detuning_params[1], detuning_params[2], Γgt, Γg, 0.0, 0.0, atom_motion, free_motion, false, true, false,
                DOPPLER_Z, parallel, SIGN_RYDBERG1, 0, 0, 1001, 0, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), 0.0, 0.0)
    flat = collect(Iterators.flatten(samples))            # Vector{Vector} -> n×6 row-major
    ρ, ρ2, _ = run1(m, tspan, ψ0.data, length(samples); samples = flat)
    b = ψ0.basis
    return operators(ρ, 5, length(tspan), b), operators(ρ2, 5, length(tspan), b)
end

"`simulation_czlp(cfg::CZLPConfig; ode_kwargs...)` -- src/cz_model.jl:107-171"
function simulation_czlp(cfg::NeutralAtoms.CZLPConfig; ode_kwargs...)
    nt = length(cfg.tspan)
    ρ = zeros(ComplexF64, 625 * nt); ρ2 = zeros(ComplexF64, 625 * nt)
    st = CaStats()
    noise = cfg.laser_noise
    f, ar, ab = cfg.f, cfg.red_laser_phase_amplitudes, cfg.blue_laser_phase_amplitudes
    fp, h = entry(:ca_rydberg2_run, :ca_rydberg2_run_multi)
    GC.@preserve f ar ab begin
        check(ccall(fp, Cint,
            (Ptr{Cvoid}, Ref{CaModel}, Ptr{Float64}, Int32, Ptr{ComplexF64}, Int64, Int64, Int64, UInt64, Ptr{Float64}, Ptr{Float64},
             Ptr{Ptr{Float64}}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ref{CaOde}, Int32, Ptr{ComplexF64}, Ptr{ComplexF64}, Ref{CaStats}),
            h, model(cfg; two_atom = true), Float64.(cfg.tspan), nt, ComplexF64.(cfg.ψ0.data), cfg.n_samples, 0, cfg.n_samples,
            next_seed(), null(), null(), Ptr{Ptr{Float64}}(C_NULL), noise ? pointer(f) : null(), noise ? pointer(ar) : null(),
            noise ? pointer(ab) : null(), noise ? length(f) : 0, ode(; ode_kwargs...), 0, ρ, ρ2, st))
    end
    b = cfg.ψ0.basis
    mk(buf) = [Operator(b, b, reshape(buf[(j-1)*625+1:j*625], 25, 25)) for j in 1:nt]
    return mk(ρ), mk(ρ2)
end

"`release_recapture(tspan, trap_params, atom_params, N; freq, skip, eps, harmonic)` -- src/basic_experiments.jl:72-81"
function release_recapture(tspan, trap_params, atom_params, N; freq = 10, skip = 1000, eps = 1e-3, harmonic = true)
    probs = zeros(Float64, length(tspan))
    st = CaStats()
    acc = 1.0
    smp = null()
    if !harmonic   # Metropolis branch: multi-chain sampler on the device, its samples go to the recapture kernel
        samples, acc = samples_generate(trap_params, atom_params, N; freq = freq, skip = skip, harmonic = false)
        smp = reduce(hcat, samples)   # 6 x N column-major = N x 6 row-major
    end
    GC.@preserve smp begin
        if MULTI[] == C_NULL
            check(ccall(sym(:ca_release_recapture), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int64, Int64, Int64, Float64, UInt64, Ptr{Float64}, Int32,
                 Ptr{Float64}, Ptr{UInt8}, Ref{CaStats}),
                CTX[], Float64.(trap_params), Float64.(atom_params), Float64.(tspan), length(tspan), N, 0, N, eps, next_seed(), smp, 0,
                probs, Ptr{UInt8}(C_NULL), st))
        else   # (the multi-GPU entry has no per-atom indicator output)
            check(ccall(sym(:ca_release_recapture_multi), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Int64, Int64, Int64, Float64, UInt64, Ptr{Float64}, Int32,
                 Ptr{Float64}, Ref{CaStats}),
                MULTI[], Float64.(trap_params), Float64.(atom_params), Float64.(tspan), length(tspan), N, 0, N, eps, next_seed(), smp, 0,
                probs, st))
        end
    end
    return probs, acc
end

"`samples_generate(trap_params, atom_params, N; freq, skip, harmonic, eps)` -- src/atom_sampler.jl:80-120 -> (Vector{Vector{Float64}}, acc_rate)"
function samples_generate(trap_params, atom_params, N; freq = 10, skip = 1000, harmonic = true, eps = 1e-2, n_chains = 0)
    out = zeros(Float64, 6, N)
    if !harmonic   # Metropolis (:97-120) as n_chains independent chains, one per GPU thread (n_chains = 1: the reference chain)
        acc = Ref{Float64}(0.0)
        check(ccall(sym(:ca_sample_atoms_metropolis), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Int32, Float64, Int64, UInt64, Int32, Ptr{Float64}, Ref{Float64}),
            CTX[], Float64.(trap_params), Float64.(atom_params), N, freq, skip, 0, eps, n_chains, next_seed(), 0, out, acc))
        return [out[:, i] for i in 1:N], acc[]
    end
    att = Ref{Int64}(0)
    check(ccall(sym(:ca_sample_atoms), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, UInt64, Int32, Float64, Ptr{Float64}, Ref{Int64}),
        CTX[], Float64.(trap_params), Float64.(atom_params), N, 0, next_seed(), 0, eps, out, att))
    return [out[:, i] for i in 1:N], 1
end

"`ϕ(tspan, f, amplitudes)` -- src/lasernoise_sampler.jl:31-37 (tspan must be a uniform grid starting at 0)"
function ϕ(tspan, f, amplitudes)
    out = zeros(Float64, length(tspan))
    check(ccall(sym(:ca_phase_noise), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Int64, Int64, UInt64, Int32, Ptr{Float64}, Ptr{Float64}),
        CTX[], Float64.(tspan), length(tspan), Float64.(f), Float64.(amplitudes), length(f), 1, 0, next_seed(), 2, null(), out))
    return out
end

"""
`get_rydberg_infidelity(cfg; U, states, n_samples, ode_kwargs...)` -- src/fidelity.jl:67-95 with the 5 configurations x
`length(states)` initial states batched into ONE `ca_rydberg1_sweep` launch per trajectory-count group (instead of 30 serial
ensembles).  Returns the same `Dict(name => infidelity)`.
"""
function get_rydberg_infidelity(cfg::NeutralAtoms.RydbergConfig;
        U = NeutralAtoms.dense(NeutralAtoms.identityoperator(cfg.ψ0.basis)),
        states = NeutralAtoms.basis_fidelity_states, n_samples = 100, ode_kwargs...)
    configs = NeutralAtoms.get_rydberg_fidelity_configs(cfg, n_samples)
    names = collect(keys(configs))
    points = [(name, configs[name], st) for name in names for st in states]
    ρend = Vector{Matrix{ComplexF64}}(undef, length(points))
    seed = next_seed(); offset = 0
    f, ar, ab = Float64.(cfg.f), Float64.(cfg.red_laser_phase_amplitudes), Float64.(cfg.blue_laser_phase_amplitudes)
    for n in sort(unique([c.n_samples for (_, c, _) in points]))
        idx = [i for (i, (_, c, _)) in enumerate(points) if c.n_samples == n]
        models = [model(points[i][2]) for i in idx]
        psis = reduce(vcat, [ComplexF64.(points[i][3].data) for i in idx])
        tends = Float64[points[i][2].tspan[end] for i in idx]
        ρ = zeros(ComplexF64, 25 * length(idx)); ρ2 = similar(ρ)
        st = CaStats()
        fp, h = entry(:ca_rydberg1_sweep, :ca_rydberg1_sweep_multi)
        GC.@preserve models psis tends f ar ab check(ccall(fp, Cint,
            (Ptr{Cvoid}, Ptr{CaModel}, Ptr{ComplexF64}, Ptr{Float64}, Int32, Int64, Int64, Int64, UInt64, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Int32, Ref{CaOde}, Int32, Ptr{ComplexF64}, Ptr{ComplexF64}, Ref{CaStats}),
            h, models, psis, tends, length(idx), n, offset, 0, seed, f, ar, ab, length(f), ode(; ode_kwargs...), 0, ρ, ρ2, st))
        for (k, i) in enumerate(idx)
            ρend[i] = reshape(ρ[(k-1)*25+1:k*25], 5, 5)
        end
        offset += n * length(idx)
    end
    infidelities = Dict()
    for name in names
        acc = 0.0
        for (i, (nm, _, st)) in enumerate(points)
            nm == name || continue
            ψ = (U * st).data
            acc += 1.0 - real(ψ' * ρend[i] * ψ)
        end
        infidelities[name] = acc / length(states)
    end
    return infidelities
end

# F(ϕ) = <Φ+| Had RZZ(ϕ) ρ (Had RZZ(ϕ))† |Φ+> of src/fidelity.jl:166 for all ϕ: RZ(ϕ) ⊗ RZ(ϕ) is diagonal, so with h = Had† Φ+ (four non-zero
# qubit components) F(ϕ) = v(ϕ)' ρ v(ϕ), v = conj(d(ϕ)) .* h -- 62 832 four-term forms instead of 62 832 products of 25×25 operators.
function parity_scan(ρ1, ϕ_list)
    Had = NeutralAtoms.Id ⊗ NeutralAtoms.Hadamard
    Φp = (NeutralAtoms.ket_0 ⊗ NeutralAtoms.ket_0 + NeutralAtoms.ket_1 ⊗ NeutralAtoms.ket_1) / sqrt(2)
    h = (dagger(Had) * Φp).data
    nz = findall(x -> abs(x) > 0, h)
    R = ρ1.data[nz, nz]
    F = zeros(Float64, length(ϕ_list))
    for (k, ϕ) in enumerate(ϕ_list)
        d1 = [NeutralAtoms.RZ(ϕ).data[i, i] for i in 1:5]
        d = kron(d1, d1)                      # diagonal of RZ(ϕ) ⊗ RZ(ϕ) (both factors equal: the ⊗ ordering does not matter)
        v = conj.(d[nz]) .* h[nz]
        F[k] = real(v' * R * v)
    end
    return F
end

"`get_parity_fidelity(cfg::CZLPConfig; ode_kwargs...)` -- src/fidelity.jl:147-176 -> (ϕ_list, F_list, ϕ at the maximum); the scan in closed form"
function get_parity_fidelity(cfg::NeutralAtoms.CZLPConfig; ode_kwargs...)
    c = deepcopy(cfg)
    ket_pos = (NeutralAtoms.ket_0 + NeutralAtoms.ket_1) / sqrt(2)
    c.ψ0 = ket_pos ⊗ ket_pos
    ρ1 = simulation_czlp(c; ode_kwargs...)[1][end]
    ϕ_list = [0.0:0.0001:2π;]
    F = parity_scan(ρ1, ϕ_list)
    return ϕ_list, F, ϕ_list[argmax(F)]
end

"""
`get_cz_infidelity(cfg::CZLPConfig; n_samples=1, ode_kwargs...)` -- src/fidelity.jl:193-240 with its seven `simulation_czlp` ensembles (the
parity-scan run that fixes ϕ_RZ (:205), the calibration-error run (:209), the five error-budget configurations (:218-231)) in ONE
`ca_rydberg2_sweep` launch, trajectories of all of them handed out longest first.  Returns `(infidelities::Dict, calibration_error)`.
"""
function get_cz_infidelity(cfg::NeutralAtoms.CZLPConfig; n_samples = 1, ode_kwargs...)
    configs = NeutralAtoms.get_rydberg_fidelity_configs(cfg, n_samples)
    names = collect(keys(configs))
    ket_pos = (NeutralAtoms.ket_0 + NeutralAtoms.ket_1) / sqrt(2)
    ψ0 = ket_pos ⊗ ket_pos
    cfg_t = deepcopy(cfg)
    cfg_t.spontaneous_decay_intermediate = false
    cfg_t.spontaneous_decay_rydberg = false
    cfg_t.laser_noise = false
    cfg_t.atom_params[2] = 1.0
    cfg_t.n_samples = 1
    pts = vcat([cfg_t, deepcopy(cfg_t)], [deepcopy(configs[name]) for name in names])
    models = [model(c; two_atom = true) for c in pts]
    psis = repeat(ComplexF64.(ψ0.data), length(pts))
    tends = Float64[c.tspan[end] for c in pts]
    counts = Int64[c.n_samples for c in pts]
    ρ = zeros(ComplexF64, 625 * length(pts)); ρ2 = similar(ρ)
    st = CaStats()
    f, ar, ab = Float64.(cfg.f), Float64.(cfg.red_laser_phase_amplitudes), Float64.(cfg.blue_laser_phase_amplitudes)
    GC.@preserve models psis tends counts f ar ab check(ccall(sym(:ca_rydberg2_sweep), Cint,
        (Ptr{Cvoid}, Ptr{CaModel}, Ptr{ComplexF64}, Ptr{Float64}, Int32, Ptr{Int64}, Int64, UInt64, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Int32, Ref{CaOde}, Int32, Ptr{ComplexF64}, Ptr{ComplexF64}, Ref{CaStats}),
        CTX[], models, psis, tends, length(pts), counts, 0, next_seed(), f, ar, ab, length(f), ode(; ode_kwargs...), 0, ρ, ρ2, st))
    b = ψ0.basis
    ρend = [Operator(b, b, reshape(ρ[(k-1)*625+1:k*625], 25, 25)) for k in 1:length(pts)]
    ϕ_list = [0.0:0.0001:2π;]
    ϕ_RZ = ϕ_list[argmax(parity_scan(ρend[1], ϕ_list))]
    Had = NeutralAtoms.Id ⊗ NeutralAtoms.Hadamard
    global_RZ = NeutralAtoms.RZ(ϕ_RZ) ⊗ NeutralAtoms.RZ(ϕ_RZ)
    Φp = (NeutralAtoms.ket_0 ⊗ NeutralAtoms.ket_0 + NeutralAtoms.ket_1 ⊗ NeutralAtoms.ket_1) / sqrt(2)
    infid(ρk) = 1.0 - real(dagger(Φp) * (Had * global_RZ * ρk * dagger(Had * global_RZ)) * Φp)
    calibration_error = infid(ρend[2])
    infidelities = Dict()
    for (k, name) in enumerate(names)
        infidelities[name] = maximum([infid(ρend[k+2]) - calibration_error, 0.0])
    end
    return infidelities, calibration_error
end

"Route the reference's own entry points through the GPU path (fidelity.jl and user scripts keep working unchanged)."
function override!()
    @eval NeutralAtoms begin
        simulation(cfg::RydbergConfig; kw...) = Main.NeutralAtomsB200.simulation(cfg; kw...)
        simulation_czlp(cfg::CZLPConfig; kw...) = Main.NeutralAtomsB200.simulation_czlp(cfg; kw...)
        release_recapture(tspan, trap_params, atom_params, N; kw...) = Main.NeutralAtomsB200.release_recapture(tspan, trap_params, atom_params, N; kw...)
        calibrate_two_photon(cfg::RydbergConfig, n_samples=1000) = Main.NeutralAtomsB200.calibrate_two_photon(cfg, n_samples)
        get_blockade_stark_shift_factor(trap_params, atom_params, atom_centers, Ω, c6, n_samples=10000) =
            Main.NeutralAtomsB200.get_blockade_stark_shift_factor(trap_params, atom_params, atom_centers, Ω, c6, n_samples)
        # the fidelity callers as batched launches (src/fidelity.jl:67-95, 147-176, 193-240)
        get_rydberg_infidelity(cfg::RydbergConfig; kw...) = Main.NeutralAtomsB200.get_rydberg_infidelity(cfg; kw...)
        get_parity_fidelity(cfg::CZLPConfig; kw...) = Main.NeutralAtomsB200.get_parity_fidelity(cfg; kw...)
        get_cz_infidelity(cfg::CZLPConfig; kw...) = Main.NeutralAtomsB200.get_cz_infidelity(cfg; kw...)
    end
    return nothing
end

end # module
