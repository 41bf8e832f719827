#!/bin/bash
# SAM of bwa-b200 vs the reference on a sample of the bench workload (3 Gbp index): tools/parity_big.sh [n_pairs]
N=${1:-10000}
D=/tmp/bwa_b200_bench
python bench.py --layout pe --steps 1 --warmup 1 --cpu-sample 2000 > /dev/null 2>&1    # makes the index and the reads
FA=$D/ref_3000.fa; FQ1=$(ls $D/*_1.fq | head -1); FQ2=$(ls $D/*_2.fq | head -1)
head -n $((4*N)) $FQ1 > /tmp/s_1.fq; head -n $((4*N)) $FQ2 > /tmp/s_2.fq
oracle/_ref/bwa mem -t 16 -K 100000000 $FA /tmp/s_1.fq /tmp/s_2.fq 2>/dev/null | grep -v '^@PG' > /tmp/ref.sam
for intv in 32 8; do
  BWA_B200_SA_INTV=$intv bwa_b200/bwa-b200 mem -t 16 -K 100000000 $FA /tmp/s_1.fq /tmp/s_2.fq 2>/tmp/b200.err | grep -v '^@PG' > /tmp/b200.sam
  echo "sa_intv $intv: ref lines $(wc -l < /tmp/ref.sam) b200 lines $(wc -l < /tmp/b200.sam) differing $(diff /tmp/ref.sam /tmp/b200.sam | grep -c '^<')"
done
diff /tmp/ref.sam /tmp/b200.sam | head -8 | cut -c1-400
tail -3 /tmp/b200.err
mkdir -p gpurun_out; diff /tmp/ref.sam /tmp/b200.sam | head -400 > gpurun_out/parity_big.diff
