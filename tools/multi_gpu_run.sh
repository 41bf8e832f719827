#!/bin/bash
# N-GPU check (default 2): the weak-scaling benchmark line under torchrun, then the multi-GPU `mem` launcher: parity against the
# reference on a small case, and the wall time of its alignment phase on a 16 M-read file with striped ingest on and off
N=${1:-2}
REP=${REP:-16}          # the big file = the bench's 1 M-read PE files REP times over
MODES=${MODES:-"1 0"}   # striped ingest on / off
SINGLE=${SINGLE:-1}      # also run the big file on one GPU
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 > $O/scale_n$N.json 2> $O/scale_n$N.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/scale_n$N.json').read().strip().splitlines()[-1]); print('N=%d: e2e %.0f reads/s, %.1f ms/step, value %.0f' % (d['n_gpus'], d['e2e']['value'], d['ms_per_step'], d['value']))"; grep "replicated" $O/scale_n$N.err; lap bench
D=/tmp/mg; mkdir -p $D
python tools/gen_data.py ref --out $D/ref.fa --contigs 3 --len 150000 --seed 11 && oracle/_ref/bwa index $D/ref.fa 2>/dev/null
python tools/gen_data.py reads --ref $D/ref.fa --out $D/r -n 20000 --len 150 --seed 12 --paired
oracle/_ref/bwa mem -v 1 -t 8 -K 600000 $D/ref.fa $D/r_1.fq $D/r_2.fq 2>/dev/null | grep -v '^@PG' > $D/ref.sam
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 -m bwa_b200.multi -v 1 -t 8 -K 600000 -o $D/out.sam $D/ref.fa $D/r_1.fq $D/r_2.fq > $O/multi_n$N.log 2>&1
echo "multi rc=$?"; grep -v '^@PG' $D/out.sam > $D/out.nopg.sam; grep "striped" $O/multi_n$N.log
echo "multi-GPU mem: ref lines $(wc -l < $D/ref.sam), ours $(wc -l < $D/out.nopg.sam), differing $(diff $D/ref.sam $D/out.nopg.sam | grep -c '^<')" | tee -a $O/multi_n$N.log
lap multi_parity
# throughput of the launcher's alignment phase: 16 M reads (the bench's 1 M-read PE files, 16 times over) against the 3 Gbp index
W=/tmp/bwa_b200_bench; FA=$W/ref_3000.fa; A=$W/reads_pe1000000_150_e10_r0_1.fq; B=$W/reads_pe1000000_150_e10_r0_2.fq
if [ -f $A ]; then
  for k in $(seq $REP); do cat $A; done > $D/big_1.fq; for k in $(seq $REP); do cat $B; done > $D/big_2.fq; ls -la $D/big_1.fq | awk '{print $5, $9}'; lap bigfile
  for striped in $MODES; do
    BWA_B200_STRIPED=$striped timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$striped -m bwa_b200.multi -v 1 -t 16 -o $D/big$striped.sam $FA $D/big_1.fq $D/big_2.fq > $O/multi_big_n${N}_striped$striped.log 2>&1
    echo "N=$N striped=$striped rc=$?: $(grep -h 'striped ingest\|spent' $O/multi_big_n${N}_striped$striped.log | tr '\n' ' ')"; lap big_striped$striped
  done
  [ -f $D/big0.sam ] && cmp <(grep -v '^@PG' $D/big1.sam) <(grep -v '^@PG' $D/big0.sam) && echo "striped and unstriped outputs identical ($(wc -l < $D/big1.sam) lines)"
  if [ "$SINGLE" = 1 ]; then
    timeout 900 python -m bwa_b200.multi -v 1 -t 16 -o $D/big_single.sam $FA $D/big_1.fq $D/big_2.fq > $O/single_big.log 2>&1; echo "N=1 rc=$?: $(grep -h 'spent' $O/single_big.log)"
    cmp <(grep -v '^@PG' $D/big1.sam) <(grep -v '^@PG' $D/big_single.sam) && echo "N=$N output identical to the single-GPU output"
    lap single
  fi
fi
ls -la $O/ | awk '{print $5, $9}' | grep -i "multi\|scale\|single"
