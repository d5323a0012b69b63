#!/bin/bash
# round-2 multi-GPU pass (gpurun --gpus 8): weak and strong scaling of R1, Z1 weak, the full S1 sweep; one JSON line each
N=${1:-8}
O=gpurun_out/scale$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > $O/bench_r1_weak.json 2> $O/bench_r1_weak.err
$TR --master-port 29542 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --scaling strong > $O/bench_r1_strong.json 2> $O/bench_r1_strong.err
$TR --master-port 29543 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu --workload z1 > $O/bench_z1_weak.json 2> $O/bench_z1_weak.err
$TR --master-port 29544 bench.py --gpus $N --steps 1 --warmup 3 --no-cpu --workload s1 > $O/bench_s1.json 2> $O/bench_s1.err
for f in $O/*.json; do echo $f; cut -c1-160 $f; done
tail -c 300 $O/*.err
