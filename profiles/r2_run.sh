#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun, one part per call: the .ncu-rep files must stay below gpurun's 64 MiB):
#   bash profiles/r2_run.sh a   ncu launch list of the default bench.py run + `ncu --set full` of k_lindblad1 (R1, NS = 15)
#   bash profiles/r2_run.sh b   `ncu --set full` of k_lindblad1 (R2 flat-top) and k_lindblad2 (Z1)
#   bash profiles/r2_run.sh c   `ncu --set full` of k_lindblad2_full
# Every capture runs after its plain run exited 0.  The reports come back in gpurun_out/r2/ and are summarised HERE with
# profiles/ncu_summary.py / ncu_phases.py into profiles/r2/*.txt.
mkdir -p gpurun_out/r2
O=gpurun_out/r2
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
if [ "$1" = "a" ]; then
python bench.py --no-cpu --steps 2 --warmup 3 > $O/bench_plain.json 2> $O/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_py.csv python bench.py --no-cpu --steps 2 --warmup 3 > $O/ncu_launches.log 2>&1
python profiles/prof_r1.py 37888 1 > $O/plain_r1.log 2>&1 && $NCU -k regex:k_lindblad1 -o $O/k_lindblad1_r1 python profiles/prof_r1.py 37888 1 > $O/ncu_r1.log 2>&1
python profiles/prof_ns15.py 33152 > $O/plain_ns15.log 2>&1 && $NCU -k regex:k_lindblad1 -o $O/k_lindblad1_ns15 python profiles/prof_ns15.py 33152 > $O/ncu_ns15.log 2>&1
elif [ "$1" = "b" ]; then
python profiles/prof_r2.py 37888 > $O/plain_r2.log 2>&1 && $NCU -k regex:k_lindblad1 -o $O/k_lindblad1_r2 python profiles/prof_r2.py 37888 > $O/ncu_r2.log 2>&1
python profiles/prof_z1.py 1184 0.2 > $O/plain_z1.log 2>&1 && $NCU -k regex:k_lindblad2 -o $O/k_lindblad2_z1 python profiles/prof_z1.py 1184 0.2 > $O/ncu_z1.log 2>&1
else
python profiles/prof_z1f.py 444 0.02 > $O/plain_z1f.log 2>&1 && $NCU -k regex:k_lindblad2_full -o $O/k_lindblad2_full python profiles/prof_z1f.py 444 0.02 > $O/ncu_z1f.log 2>&1
fi
tail -n 2 $O/plain_*.log
ls -la $O
