#!/bin/bash
# last round-2 call: ncu launch list of the default bench.py run on the final library (after its plain run exited 0), NS = 15 rate at 12 waves
cd /root/repo; O=gpurun_out/last; mkdir -p $O
python bench.py --no-cpu --steps 2 --warmup 3 > $O/bench_plain.json 2> $O/bench_plain.err && \
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_py.csv python bench.py --no-cpu --steps 2 --warmup 3 > $O/ncu_launches.log 2>&1
timeout 40 python profiles/prof_ns15.py 397824 > $O/ns15_rate.txt 2>&1
cut -c1-200 $O/bench_plain.json; tail -n 2 $O/ns15_rate.txt; wc -l $O/launches_bench_py.csv
