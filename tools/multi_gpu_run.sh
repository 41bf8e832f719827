#!/bin/bash
# N-GPU check (default 2): the weak-scaling benchmark line under torchrun, then the multi-GPU `mem` launcher against the reference's SAM
N=${1:-2}
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 4 --warmup 3 > $O/scale_n$N.json 2> $O/scale_n$N.err
echo "bench rc=$?"; tail -c 900 $O/scale_n$N.json; echo; grep "replicated" $O/scale_n$N.err; lap bench
D=/tmp/mg; mkdir -p $D
python tools/gen_data.py ref --out $D/ref.fa --contigs 3 --len 150000 --seed 11 && oracle/_ref/bwa index $D/ref.fa 2>/dev/null
python tools/gen_data.py reads --ref $D/ref.fa --out $D/r -n 20000 --len 150 --seed 12 --paired
oracle/_ref/bwa mem -v 1 -t 8 -K 600000 $D/ref.fa $D/r_1.fq $D/r_2.fq 2>/dev/null | grep -v '^@PG' > $D/ref.sam
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 -m bwa_b200.multi -v 1 -t 8 -K 600000 -o $D/out.sam $D/ref.fa $D/r_1.fq $D/r_2.fq > $O/multi_n$N.log 2>&1
echo "multi rc=$?"; grep -v '^@PG' $D/out.sam > $D/out.nopg.sam
echo "multi-GPU mem: ref lines $(wc -l < $D/ref.sam), ours $(wc -l < $D/out.nopg.sam), differing $(diff $D/ref.sam $D/out.nopg.sam | grep -c '^<')" | tee -a $O/multi_n$N.log
lap multi
