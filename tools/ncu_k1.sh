#!/bin/bash
# one full-size launch of each named kernel under ncu (--set full, with source); usage: tools/ncu_k1.sh <skip> <kernel>...
SKIP=${1:-4}; shift
mkdir -p gpurun_out
python bench.py --layout se --steps 1 --warmup 1 --cpu-sample 2000 > /dev/null 2>&1   # build the index once, outside ncu
for K in "$@"; do
BWA_B200_LANES=1 BWA_B200_CHUNK=1000000 timeout 1500 ncu --set full --import-source on --clock-control none -k regex:^$K\$ -s $SKIP -c 1 \
  -o gpurun_out/ncu_$K -f python bench.py --worker --layout se --steps 1 --warmup 3 --cpu-sample 2000 > gpurun_out/ncu_$K.log 2>&1
tail -2 gpurun_out/ncu_$K.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
