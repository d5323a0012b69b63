#!/bin/bash
# developer loop for K3 experiments: builds variants/libfast.so with ONLY k_lindblad1<9,false,1> (the R1 variant; ~25 s instead of
# ~130 s) from the working tree.  Use on the GPU box as  COLDATOMS_B200_LIB=variants/libfast.so python profiles/prof_r1.py ...
set -e
R=/root/repo
mkdir -p $R/variants
cd $R/coldatoms.jl_b200/csrc
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v -DCA_FAST_BUILD=${CA_FAST:-1} ${CA_DEFS} \
    -shared -o $R/variants/${1:-libfast}.so ca_kernels.cu ca_lindblad2.cu ca_model.cpp 2> $R/variants/${1:-libfast}.log
grep -A2 "k_lindblad1ILi9ELb0ELi${CA_FAST:-1}" $R/variants/${1:-libfast}.log | grep -v Compiling | paste - - | sed 's/ptxas info    : //g; s/Function properties for //'
