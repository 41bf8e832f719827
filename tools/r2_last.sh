#!/bin/bash
# Round 2, last GPU call (3 minutes of budget): device-clock timeline of the final pipeline (k_smem_c) -- how busy is the GPU in the steady part of a run
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
BWA_B200_BENCH_VERIFY=0 BWA_B200_GPUTRACE=1 timeout 200 python bench.py --worker --steps 12 --warmup 3 --cpu-sample 2000 > $O/r2z_pe_trace.json 2>$O/r2z_pe_trace.err
python -c "
import json; d=json.loads(open('$O/r2z_pe_trace.json').read().strip().splitlines()[-1]); print('traced default: e2e %.1f ms/step' % d['ms_per_step'], d['kernels_ms_per_step'])"
python tools/gpu_timeline.py --window 24 $O/r2z_pe_trace.err | tee $O/r2z_pe_timeline.txt | tail -30; lap default
grep "gputrace" $O/r2z_pe_trace.err | gzip > $O/r2z_pe_trace.gz; rm -f $O/r2z_pe_trace.err
