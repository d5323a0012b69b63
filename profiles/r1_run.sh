#!/bin/bash
# Round-1 measurement pass (run on the GPU box through gpurun): tests, smoke, bench lines, ncu launch list, ncu full captures.
mkdir -p gpurun_out/r1
O=gpurun_out/r1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -1 | tee $O/smoke.txt
python bench.py > $O/bench_r1.json 2> $O/bench_r1.err; cut -c1-300 $O/bench_r1.json
python bench.py --workload z1 --steps 2 --warmup 3 --cpu-seconds 8 > $O/bench_z1.json 2> $O/bench_z1.err; cut -c1-300 $O/bench_z1.json
python bench.py --workload r1b --steps 3 --warmup 3 --cpu-seconds 5 > $O/bench_r1b.json 2> $O/bench_r1b.err; cut -c1-300 $O/bench_r1b.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --no-cpu > $O/ncu_launches.log 2>&1
python profiles/prof_r1.py 37888 1 > $O/plain_r1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_lindblad1 -c 1 -f -o $O/k_lindblad1 python profiles/prof_r1.py 37888 1 > $O/ncu_r1.log 2>&1
python profiles/prof_z1.py 2368 0.05 > $O/plain_z1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_lindblad2 -c 1 -f -o $O/k_lindblad2 python profiles/prof_z1.py 2368 0.05 > $O/ncu_z1.log 2>&1
tail -n 2 $O/plain_r1.log $O/plain_z1.log
