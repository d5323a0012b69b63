#!/bin/bash
# developer loop: host-emulation build + CPU tests of the device source, then the sm_100a build with register report
set -e
R=/root/repo
make -s -C $R/tests/hostemu
(cd $R && python -m pytest tests/test_hostemu.py -x -q 2>&1 | tail -3)
make -C $R/coldatoms.jl_b200/csrc 2>&1 | grep -E "error|warning" || true
grep -A2 "k_lindblad1ILi9ELb0\|k_lindblad2" $R/coldatoms.jl_b200/csrc/build.log | grep -v Compiling | paste - - | sed 's/ptxas info    : //g; s/Function properties for //'
