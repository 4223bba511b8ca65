#!/bin/bash
# dev: build a phase-timing variant of the library (separate .so) and print the per-phase cycle breakdown of CTA 0
set -e
cd energybalancemodel.jl_b200/csrc
mkdir -p build_pt
for f in ebm_capi classic_uniform classic_bands classic_strict miz_kernel util_kernels; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include -DEBM_PHASE_TIMING -c $f.cu -o build_pt/$f.o &
done
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include -fmad=false -c miz_strict.cu -o build_pt/miz_strict.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libebm_cuda_pt.so build_pt/*.o -lcudart
