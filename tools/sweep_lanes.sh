#!/bin/bash
# e2e vs (lanes, host threads) on the GPU box; assumes the 3 Gbp workload already exists in /tmp/bwa_b200_bench
cd /root/repo
for layout in se pe; do
 for cfg in "2 16" "3 16" "4 16" "3 12" "4 20"; do
  set -- $cfg
  BWA_B200_LANES=$1 timeout 600 python bench.py --layout $layout --threads $2 --steps 4 --warmup 2 --cpu-sample 2000 > /tmp/sw.json 2> /tmp/sw.err
  python -c "import json; d=json.load(open('/tmp/sw.json')); print('$layout lanes $1 threads $2: e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1))"
  grep Processed /tmp/sw.err | tail -7 | head -4 | sed 's/.*in \([0-9.]*\) CPU sec, \([0-9.]*\) real.*/\1cpu \2real/' | tr '\n' ' '; echo
 done
done
