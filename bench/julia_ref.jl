# julia_ref.jl -- OFF-BOX timing of the UNMODIFIED reference (BASELINE.md §3 item 3; SURVEY.md §8d).  Julia is not available in the
# build container or on the GPU box, so this file has NOT been run; `bench.py --impl reference` times the C oracle port instead and says
# so (`cpu_baseline.kind = "port"`).  On a machine that has Julia and the package's dependencies:
#
#   julia --project=/path/to/ColdAtoms.jl -t auto bench/julia_ref.jl [workload] [n_samples]
#
# workload: r1 (default; BASELINE config 2: get_default_configs()[1], laser_noise = true, tspan = 0:T0/20:5T0), r1b (pi pulse, [0, T0/2]),
#           z1 (BASELINE config 4: simulation_czlp, tspan = [0, 2τ]), c1 (BASELINE config 1: release_recapture, 10^4 atoms).
# Prints one JSON line in the format of bench.py (metric, value = trajectories/s through the reference's own public API and stock code
# path, cores = Threads.nthreads(); the reference loop itself is serial: src/rydberg_model.jl:260-262, src/cz_model.jl:151).
using NeutralAtoms, QuantumOptics
include(joinpath(dirname(pathof(NeutralAtoms)), "..", "default.jl"))

workload = length(ARGS) >= 1 ? ARGS[1] : "r1"
cfg, cfg_cz = get_default_configs()

function timed(f)
    f()                       # compile
    t0 = time_ns(); f(); (time_ns() - t0) / 1e9
end

if workload == "c1"
    n = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 10_000
    trap = [1000.0, 1.0, w0_to_z0(1.0, 0.852)]
    secs = timed(() -> release_recapture(collect(0.0:1.0:50.0), trap, [86.9091835, 50.0], n; eps = 1e-3, harmonic = true))
    println("{\"impl\": \"julia_ref\", \"metric\": \"release-recapture atoms/sec\", \"value\": $(n / secs), \"unit\": \"atoms/s\", ",
            "\"cores\": $(Threads.nthreads()), \"sample\": \"$n atoms, tspan = 0:1:50, harmonic sampler\"}")
elseif workload == "z1"
    n = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 8
    cfg_cz.laser_noise = true
    cfg_cz.n_samples = n
    secs = timed(() -> simulation_czlp(cfg_cz; maxiters = 10^7))
    println("{\"impl\": \"julia_ref\", \"metric\": \"master-eq trajectories/sec\", \"value\": $(n / secs), \"unit\": \"trajectories/s\", ",
            "\"cores\": $(Threads.nthreads()), \"sample\": \"$n trajectories of Z1 (simulation_czlp, default.jl cfg_czlp, laser_noise = true)\"}")
else
    n = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 64
    cfg.laser_noise = true
    cfg.n_samples = n
    if workload == "r1b"
        cfg.tspan = [0.0, cfg.tspan[end] / 10]   # [0, T0/2] of tspan = 0:T0/20:5T0
    end
    secs = timed(() -> simulation(cfg))
    println("{\"impl\": \"julia_ref\", \"metric\": \"master-eq trajectories/sec\", \"value\": $(n / secs), \"unit\": \"trajectories/s\", ",
            "\"cores\": $(Threads.nthreads()), \"sample\": \"$n trajectories of $(uppercase(workload)) (simulation, default.jl cfg, laser_noise = true)\"}")
end
