#!/bin/bash
# one full-size launch of a kernel under ncu (--set full, with source); usage: tools/ncu_k1.sh <kernel> <skip> [extra bench args]
K=${1:-k_smem}; SKIP=${2:-4}; shift; shift
mkdir -p gpurun_out
BWA_B200_LANES=1 BWA_B200_CHUNK=1000000 timeout 1500 ncu --set full --import-source on --clock-control none -k regex:^$K\$ -s $SKIP -c 1 \
  -o gpurun_out/ncu_$K -f python bench.py --layout se --steps 1 --warmup 3 --cpu-sample 2000 "$@" > gpurun_out/ncu_$K.log 2>&1
tail -3 gpurun_out/ncu_$K.log
ls -la gpurun_out/ncu_$K.ncu-rep
