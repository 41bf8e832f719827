#!/bin/bash
# Round 2, GPU call 11: the index kept resident across processes (CUDA IPC): test, and what attaching costs next to loading; default line
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 600 python -m pytest tests/test_resident.py tests/test_tail.py -x -q -m gpu > $O/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2l_pytest.log; lap pytest
python bench.py --worker --steps 12 --warmup 3 --cpu-sample 20000 > $O/r2l_pe.json 2>$O/r2l_pe.err; python -c "
import json; d=json.loads(open('$O/r2l_pe.json').read().strip().splitlines()[-1]); print('default: e2e %.0f reads/s, %.1f ms/step' % (d['e2e']['value'], d['ms_per_step']), {k: round(v,1) for k,v in d['kernels_ms_per_step'].items()}, d['cpu_baseline'].get('sam_identical_on_sample'))"; lap pe
# 3 Gbp: bwa-b200 mem on 20 k reads, loading the index vs attaching to the resident copy
W=/tmp/bwa_b200_bench; FA=$W/ref_3000.fa; R1=$(ls $W/reads_pe1000000_150_e10_r0_1.fq.sample10000 2>/dev/null); R2=$(ls $W/reads_pe1000000_150_e10_r0_2.fq.sample10000 2>/dev/null)
ls $W | head -20
if [ -n "$R1" ]; then
  /usr/bin/time -f "load+align: %e s wall" bwa_b200/bwa-b200 mem -v 1 -t 16 $FA $R1 $R2 > /tmp/a.sam 2> $O/r2l_cold.err; tail -1 $O/r2l_cold.err
  /usr/bin/time -f "shm (make resident): %e s wall" bwa_b200/bwa-b200 shm $FA 2> $O/r2l_shm.err; tail -2 $O/r2l_shm.err
  bwa_b200/bwa-b200 shm -l
  /usr/bin/time -f "attach+align: %e s wall" bwa_b200/bwa-b200 mem -v 3 -t 16 $FA $R1 $R2 > /tmp/b.sam 2> $O/r2l_warm.err; grep -c resident $O/r2l_warm.err; tail -1 $O/r2l_warm.err
  /usr/bin/time -f "attach+align again: %e s wall" bwa_b200/bwa-b200 mem -v 1 -t 16 $FA $R1 $R2 > /tmp/c.sam 2> $O/r2l_warm2.err; tail -1 $O/r2l_warm2.err
  grep -v '^@PG' /tmp/a.sam | md5sum; grep -v '^@PG' /tmp/b.sam | md5sum; grep -v '^@PG' /tmp/c.sam | md5sum
  bwa_b200/bwa-b200 shm -d; bwa_b200/bwa-b200 shm -l; nvidia-smi --query-gpu=memory.used --format=csv
fi
lap resident
ls -la $O/r2l_* | awk '{print $5, $9}'
