#!/bin/bash
# developer loop for K4 experiments: compiles ONLY ca_lindblad2.cu (13 s) with extra -D switches (CA_DEFS) and links it with the
# objects of the last full build into variants/<name>.so.  Use on the GPU box as
#   AB_WORKLOAD=z1 AB_TEND=0.1 python profiles/ab.py 2368 variants/a.so variants/b.so
set -e
R=/root/repo
C=$R/coldatoms.jl_b200/csrc
N=${1:-libk4}
mkdir -p $R/variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v ${CA_DEFS} \
    -c -o $R/variants/$N.o $C/ca_lindblad2.cu 2> $R/variants/$N.log
OBJS=$(ls $C/build/*.o | grep -v ca_lindblad2.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -cudart shared -shared -o $R/variants/$N.so $R/variants/$N.o $OBJS -ldl
grep -A2 "Function properties for _ZN2ca11k_lindblad2ILb0" $R/variants/$N.log | grep -v Compiling | paste - - - | sed 's/ptxas info    : //g; s/Function properties for //'
