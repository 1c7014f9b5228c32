#!/bin/bash
# build a variant of libmpcgpu.so with extra nvcc flags for the c2_tmpc12 object only: tools/build_variant.sh <tag> [flags...]
set -e
cd /root/repo
P=oscar_mpc_planner_mr_modification_b200
tag=$1; shift
mkdir -p scratch
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v -I$P/generated/c2_tmpc12 \
  '-DMPC_MODEL_HEADER="model.cuh"' -DMPC_CFG_TAG=c2_tmpc12 "$@" -c $P/csrc/mpc_config_impl.cu -o scratch/cfg_c2_$tag.o 2> scratch/ptxas_$tag.log
grep -A2 "Compiling entry.*split_kernel\|Compiling entry.*solve_kernelILi8" scratch/ptxas_$tag.log | grep "Used" || true
nvcc -shared -o scratch/libmpcgpu_$tag.so scratch/cfg_c2_$tag.o $P/lib/obj/cfg_c1_basic.o $P/lib/obj/cfg_tmpc_shipped.o $P/lib/obj/cfg_c5_ccmpc.o $P/lib/obj/capi.o $P/lib/obj/multi.o $P/lib/obj/synth.o -lcudart -lpthread
echo built scratch/libmpcgpu_$tag.so
