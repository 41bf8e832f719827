#!/bin/bash
# kernel-only times with the L2 fetch granularity hint at its default and at 32 bytes
cd /root/repo
for g in 0 32; do
  BWA_B200_L2_FETCH=$g BWA_B200_LANES=1 BWA_B200_CHUNK=100000000 python bench.py --worker --layout se --inflight 1 --steps 2 --warmup 1 --cpu-sample 2000 > /tmp/v.json 2>/dev/null
  python -c "import json; d=json.load(open('/tmp/v.json')); print('l2 fetch $g:', {k: round(v,1) for k,v in d['kernels_ms_per_step'].items()}, 'e2e %.0f' % d['e2e']['value'])"
done
