#!/bin/bash
# build the c2_tmpc12 object only (extra nvcc flags in "$@") and dump the SASS of the throughput kernel to /tmp/k8.sass
set -e
cd /root/repo
P=oscar_mpc_planner_mr_modification_b200
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v -I$P/generated/c2_tmpc12 \
  '-DMPC_MODEL_HEADER="model.cuh"' -DMPC_CFG_TAG=c2_tmpc12 "$@" -c $P/csrc/mpc_config_impl.cu -o /tmp/c2.o 2>&1 | grep -A2 "mpc_solve_kernelILi8" | grep -i "spill\|Used"
cuobjdump -sass /tmp/c2.o | awk '/Function : .*mpc_solve_kernelILi8/{f=1} /Function : .*mpc_solve_kernelILi1/{f=0} f' | grep -v "^\s*/\* 0x" | cut -c1-110 > /tmp/k8.sass
wc -l /tmp/k8.sass
