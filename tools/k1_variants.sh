#!/bin/bash
# K1 occupancy variants on the GPU box: rebuild with different __launch_bounds__ min-blocks and time the kernels alone
cd /root/repo
python bench.py --layout se --steps 1 --warmup 1 --cpu-sample 2000 > /dev/null 2>&1
for mb in 1 6 8; do
  touch bwa_b200/csrc/cuda/bwag_smem.cu
  make NVEXTRA=-DK1_MIN_BLOCKS=$mb all 2>&1 | grep -E " error" 
  BWA_B200_LANES=1 BWA_B200_CHUNK=100000000 python bench.py --layout se --steps 2 --warmup 2 --cpu-sample 2000 > /tmp/v.json 2>/dev/null
  python -c "import json; d=json.load(open('/tmp/v.json')); print('min_blocks $mb:', {k: round(v,1) for k,v in d['kernels_ms_per_step'].items()}, 'roof %.0f GB/s' % d['roofline']['achieved'])"
done
touch bwa_b200/csrc/cuda/bwag_smem.cu; make all 2>&1 | grep " error"
