#!/bin/bash
# developer loop for K3 experiments: compiles ONLY one k_lindblad1 variant (default NS = 9, BEAMS = 1: the R1 kernel; ~40 s) with extra
# -D switches (CA_DEFS) and links it with the objects of the last full build into variants/<name>.so.
#   CA_DEFS="-DCA_NOISE_CHUNK=64" profiles/dev_k3.sh a [ns] [beams];  python profiles/ab.py 151552 variants/a.so variants/b.so
set -e
R=/root/repo
C=$R/coldatoms.jl_b200/csrc
N=${1:-libk3}; NS=${2:-9}; BM=${3:-1}
mkdir -p $R/variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v ${CA_DEFS} \
    -DCA_INST_NS=$NS -DCA_INST_BEAMS=$BM -c -o $R/variants/$N.o $C/ca_l1_inst.cu 2> $R/variants/$N.log
OBJS=$(ls $C/build/*.o | grep -v l1_${NS}_${BM}.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -cudart shared -shared -o $R/variants/$N.so $R/variants/$N.o $OBJS -ldl
grep -A2 "Function properties for _ZN2ca11k_lindblad1ILi${NS}ELb0" $R/variants/$N.log | grep -v Compiling | paste - - - | sed 's/ptxas info    : //g; s/Function properties for //'
