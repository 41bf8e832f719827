#!/bin/bash
# e2e vs (lanes, host threads per lane) on the GPU box; assumes the 3 Gbp workload already exists in /tmp/bwa_b200_bench
cd /root/repo
for layout in se pe; do
 for cfg in "2 8" "2 12" "3 6" "3 8" "3 12" "4 6" "4 8"; do
  set -- $cfg
  BWA_B200_LANES=$1 timeout 600 python bench.py --layout $layout --threads $2 --steps 4 --warmup 2 --cpu-sample 2000 > /tmp/sw.json 2> /tmp/sw.err
  python -c "import json; d=json.load(open('/tmp/sw.json')); print('$layout lanes $1 threads $2: e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1))"
  grep Processed /tmp/sw.err | tail -4 | awk '{printf \"%s/%s \", \$6, \$10} END {print \"\"}'
 done
done
