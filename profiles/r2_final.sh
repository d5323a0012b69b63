#!/bin/bash
# Final round-2 passes, ONE gpurun call each: (1) A/B of per-translation-unit compiler variants (profiles/final_ab.py) -> variants/cand.so,
# (2) if anything won: the GPU test-suite and smoke() on that candidate library, (3) the R1 bench line, (4) an ncu capture of the NS = 15
# kernel as it is built now.  Everything lands in gpurun_out/$1/ (default: final).
cd /root/repo; O=gpurun_out/${1:-final}; mkdir -p $O
timeout 300 python profiles/final_ab.py $O/final_ab.json > $O/final_ab.log 2>&1
export COLDATOMS_B200_LIB=/root/repo/variants/cand.so
[ -f $COLDATOMS_B200_LIB ] || export COLDATOMS_B200_LIB=/root/repo/coldatoms.jl_b200/csrc/libcoldatoms_b200.so
if ! grep -q '"winners": \[\]' $O/final_ab.json; then
timeout 420 python -m pytest tests -m gpu -x -q -rs > $O/pytest_gpu.txt 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
fi
timeout 240 python bench.py --steps 3 --warmup 3 > $O/bench_r1.json 2> $O/bench_r1.err
python profiles/prof_ns15.py 33152 > $O/plain_ns15.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -c 1 -f -k regex:k_lindblad1 -o $O/k_lindblad1_ns15 python profiles/prof_ns15.py 33152 > $O/ncu_ns15.log 2>&1
tail -n 3 $O/final_ab.log $O/pytest_gpu.txt $O/smoke.txt $O/plain_ns15.log; cut -c1-300 $O/bench_r1.json; ls -la $O
