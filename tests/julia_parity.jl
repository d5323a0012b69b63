# julia_parity.jl -- OFF-BOX parity script (Julia is not available in the build container; this file has not been run).
#
# Feeds identical host-generated samples and phases to (a) the unmodified reference (NeutralAtoms.simulation with
# samples_generate and ϕ replaced by array readers) and (b) the GPU library through the ccall shim, and prints the
# maximum deviation of ρ and ρ2.  Expected: < 1e-5 on elements, < 1e-6 on populations (the reference's own solver error
# at reltol 1e-6; see DESIGN.md §2-3).
#
#   julia --project=/path/to/ColdAtoms.jl tests/julia_parity.jl /path/to/libcoldatoms_b200.so
using NeutralAtoms, QuantumOptics, Random, Libdl
include(joinpath(@__DIR__, "..", "coldatoms.jl_b200", "julia", "NeutralAtomsB200.jl"))
using .NeutralAtomsB200
include(joinpath(dirname(pathof(NeutralAtoms)), "..", "default.jl"))

NeutralAtomsB200.init!(ARGS[1])
cfg, _ = get_default_configs()
cfg.laser_noise = true
cfg.n_samples = 16
rng = MersenneTwister(1)
samples = NeutralAtoms.samples_generate(cfg.trap_params, cfg.atom_params, cfg.n_samples)[1]
phases = [2π .* rand(rng, length(cfg.f)) for _ in 1:2cfg.n_samples]

# (a) reference with injected randomness
const SAMPLES, PHASES, COUNTER = samples, phases, Ref(0)
@eval NeutralAtoms begin
    samples_generate(trap_params, atom_params, N; kw...) = (Main.SAMPLES[1:N], 1)
    function ϕ(tspan, f, amplitudes)
        ϕf = Main.PHASES[Main.COUNTER[] += 1]
        vec(sum(amplitudes .* cos.(2π * f .* tspan' .+ ϕf), dims = 1))
    end
end
ρ_ref, ρ2_ref = NeutralAtoms.simulation(cfg)

# (b) GPU path with the same arrays: red phases are the odd draws, blue the even ones (call order in GenerateHamiltonian)
flat = collect(Iterators.flatten(samples))
pr = collect(Iterators.flatten(phases[1:2:end])); pb = collect(Iterators.flatten(phases[2:2:end]))
m = NeutralAtomsB200.model(cfg)
nt = length(cfg.tspan)
ρ = zeros(ComplexF64, 25nt); ρ2 = zeros(ComplexF64, 25nt); st = NeutralAtomsB200.CaStats()
NeutralAtomsB200.check(ccall(NeutralAtomsB200.sym(:ca_rydberg1_run), Cint,
    (Ptr{Cvoid}, Ref{NeutralAtomsB200.CaModel}, Ptr{Float64}, Int32, Ptr{ComplexF64}, Int64, Int64, Int64, UInt64, Ptr{Float64}, Ptr{Float64},
     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ref{NeutralAtomsB200.CaOde}, Int32, Ptr{ComplexF64}, Ptr{ComplexF64},
     Ref{NeutralAtomsB200.CaStats}),
    NeutralAtomsB200.CTX[], m, cfg.tspan, nt, cfg.ψ0.data, cfg.n_samples, 0, cfg.n_samples, 0, flat, pr, pb, cfg.f,
    cfg.red_laser_phase_amplitudes, cfg.blue_laser_phase_amplitudes, length(cfg.f), NeutralAtomsB200.ode(), 0, ρ, ρ2, st))
dev = maximum(abs.(vcat([vec(ρ_ref[j].data) for j in 1:nt]...) .- ρ))
dev2 = maximum(abs.(vcat([vec(ρ2_ref[j].data) for j in 1:nt]...) .- ρ2))
println("max |Δρ| = $dev   max |Δρ2| = $dev2   (rhs evals on GPU: $(st.n_rhs))")
