#!/bin/bash
# Round 2, GPU call 8: device-clock timeline of the default run (where does the GPU idle?), and of 4 calls in flight (why it collapses)
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
BWA_B200_GPUTRACE=1 BWA_B200_TRACE=1 python bench.py --worker --steps 6 --warmup 3 --cpu-sample 2000 > $O/r2i_pe_trace.json 2>$O/r2i_pe_trace.err
python -c "
import json; d=json.loads(open('$O/r2i_pe_trace.json').read().strip().splitlines()[-1]); print('traced default: e2e %.1f ms/step' % d['ms_per_step'])"
python tools/gpu_timeline.py $O/r2i_pe_trace.err 600 | tee $O/r2i_pe_timeline.txt; lap default
BWA_B200_GPUTRACE=1 BWA_B200_TRACE=1 python bench.py --worker --inflight 4 --steps 6 --warmup 3 --cpu-sample 2000 > $O/r2i_pe_if4_trace.json 2>$O/r2i_pe_if4_trace.err
python -c "
import json; d=json.loads(open('$O/r2i_pe_if4_trace.json').read().strip().splitlines()[-1]); print('traced inflight 4: e2e %.1f ms/step' % d['ms_per_step'])"
python tools/gpu_timeline.py $O/r2i_pe_if4_trace.err 600 | tee $O/r2i_pe_if4_timeline.txt; lap if4
BWA_B200_GPUTRACE=1 python bench.py --worker --inflight 2 --steps 6 --warmup 3 --cpu-sample 2000 > $O/r2i_pe_if2_trace.json 2>$O/r2i_pe_if2_trace.err
python -c "
import json; d=json.loads(open('$O/r2i_pe_if2_trace.json').read().strip().splitlines()[-1]); print('traced inflight 2: e2e %.1f ms/step' % d['ms_per_step'])"
python tools/gpu_timeline.py $O/r2i_pe_if2_trace.err 600 | tee $O/r2i_pe_if2_timeline.txt; lap if2
gzip -f $O/r2i_pe_trace.err $O/r2i_pe_if4_trace.err $O/r2i_pe_if2_trace.err
ls -la $O/r2i_* | awk '{print $5, $9}'
