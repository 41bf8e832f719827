#!/bin/bash
# e2e throughput vs (calls in flight, lanes per call, chunk size) on the GPU box; usage: tools/sweep_lanes.sh "<inflight lanes chunk>" ...
cd /root/repo; mkdir -p gpurun_out
python bench.py --layout se --steps 1 --warmup 1 --cpu-sample 2000 > /dev/null 2>&1
for layout in pe se; do
for cfg in "$@"; do
  set -- $cfg
  BWA_B200_LANES=$2 BWA_B200_CHUNK=$3 python bench.py --layout $layout --inflight $1 --steps 8 --warmup 3 --cpu-sample 2000 > /tmp/s.json 2>/dev/null
  python -c "import json; d=json.load(open('/tmp/s.json')); print('$layout inflight $1 lanes $2 chunk $3: e2e %.0f reads/s, %.1f ms/step' % (d['e2e']['value'], d['ms_per_step']))"
done; done
