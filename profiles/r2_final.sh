#!/bin/bash
# Final round-2 pass, ONE gpurun call: (1) A/B of per-translation-unit compiler variants (profiles/final_ab.py) -> variants/cand.so,
# (2) the GPU test-suite, smoke() and the R1 / Z1 bench lines on that candidate library.  Everything lands in gpurun_out/final/.
cd /root/repo; mkdir -p gpurun_out/final; O=gpurun_out/final
timeout 420 python profiles/final_ab.py $O/final_ab.json > $O/final_ab.log 2>&1
export COLDATOMS_B200_LIB=/root/repo/variants/cand.so
[ -f $COLDATOMS_B200_LIB ] || export COLDATOMS_B200_LIB=/root/repo/coldatoms.jl_b200/csrc/libcoldatoms_b200.so
timeout 420 python -m pytest tests -m gpu -x -q -rs > $O/pytest_gpu.txt 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
timeout 240 python bench.py --steps 3 --warmup 3 > $O/bench_r1.json 2> $O/bench_r1.err
timeout 240 python bench.py --workload z1 --steps 2 --warmup 3 > $O/bench_z1.json 2> $O/bench_z1.err
tail -n 3 $O/final_ab.log $O/pytest_gpu.txt $O/smoke.txt; cut -c1-400 $O/bench_r1.json $O/bench_z1.json
