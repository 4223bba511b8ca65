#!/bin/bash
# usage: devlib.sh <tag> [extra nvcc flags] -> energybalancemodel.jl_b200/lib/libebm_dev_<tag>.so (UPAR instance only)
tag=$1; shift
cd /root/repo/energybalancemodel.jl_b200/csrc
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include -DEBM_DEV_FAST_BUILD "$@" -c classic_uniform.cu -o /tmp/cu_$tag.o || exit 1
objs=$(ls build/*.o | grep -v classic_uniform.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libebm_dev_$tag.so $objs /tmp/cu_$tag.o -lcudart -ldl
