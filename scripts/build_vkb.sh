#!/bin/bash
# usage: scripts/build_vkb.sh tag [-D flags...]   -> scripts/_build/vkb_<tag>
tag=$1; shift
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false --expt-relaxed-constexpr \
     -I hepmc_b200/csrc "$@" scripts/vegas_kernel_bench.cu -o scripts/_build/vkb_$tag
