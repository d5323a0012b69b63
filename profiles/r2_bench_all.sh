#!/bin/bash
# round-2 measurement pass on one GPU: every BASELINE config through bench.py (one JSON line each, profiles/r2/bench_*.json)
cd /root/repo; mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --workload r1b --steps 3 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
python bench.py --workload r2 --steps 2 --warmup 3 > gpurun_out/bench_r2_hg4.json 2> gpurun_out/bench_r2_hg4.err
python bench.py --workload r2 --beam lg6 --steps 2 --warmup 3 > gpurun_out/bench_r2_lg6.json 2> gpurun_out/bench_r2_lg6.err
python bench.py --workload z1 --steps 2 --warmup 3 > gpurun_out/bench_z1.json 2> gpurun_out/bench_z1.err
python bench.py --workload c1 --steps 20 --warmup 5 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
python bench.py --workload c1 --trajectories 10000000 --steps 5 --warmup 3 > gpurun_out/bench_c1_1e7.json 2> gpurun_out/bench_c1_1e7.err
python bench.py --workload s1 --steps 1 --warmup 3 --traj-per-point ${S1_TPP:-10000} > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref_r1.err
tail -c 400 gpurun_out/bench_*.err
