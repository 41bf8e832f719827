#!/bin/bash
# phase timeline of one PE and one SE batch on the GPU box: tools/trace_run.sh > gpurun_out/trace.log
mkdir -p gpurun_out
for layout in pe se; do
  BWA_B200_TRACE=1 python bench.py --layout $layout --steps 2 --warmup 3 > gpurun_out/trace_$layout.json 2> gpurun_out/trace_$layout.err
  echo "== $layout"; cat gpurun_out/trace_$layout.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e'], d['ms_per_step'], d['value'])"
done
