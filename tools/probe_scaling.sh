#!/bin/bash
# host-side scaling probe on the GPU box: reference bwa mem and bwa-b200 at several thread counts (20 Mbp ref, 200k reads)
set -e
cd /root/repo
python bench.py --ref-mbp 20 --reads 200000 --steps 1 --warmup 1 --cpu-sample 1000 > /dev/null 2>&1 || true
D=/tmp/bwa_b200_bench
lscpu | grep -E "^CPU\(s\)|Thread|Core|Socket|NUMA node\(s\)|Model name" 
cat /sys/fs/cgroup/cpu.max 2>/dev/null || true
for t in 1 8 32 64 128; do
  /usr/bin/time -f "ref -t $t wall %e s cpu %U+%S" oracle/_ref/bwa mem -t $t -K 100000000 -v 3 $D/ref_20.fa $D/reads_200000_150_r0.fq 2>&1 >/dev/null | grep -E "Processed|wall" | tail -2
done
for t in 8 16 32 64 128; do
  echo "== bwa-b200 -t $t"
  BWA_B200_PROFILE=1 bwa_b200/bwa-b200 mem -t $t -K 100000000 -v 3 $D/ref_20.fa $D/reads_200000_150_r0.fq 2>&1 >/dev/null | grep -E "Processed|prof" | awk '{printf "%s ", $0} END {print ""}' | sed 's/\[prof\]//g; s/  */ /g'
done
