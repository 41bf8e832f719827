#!/bin/bash
# kernel variants on the GPU box: rebuild with different macros and time the kernels alone (same box, same index)
cd /root/repo
python bench.py --layout se --steps 1 --warmup 1 --cpu-sample 2000 > /dev/null 2>&1
for v in "" "$@" ; do   # e.g. tools/k1_variants.sh -DK1_MIN_BLOCKS=4 -DK1_SLOTS=8 -DK1_NO_QSMEM
  touch bwa_b200/csrc/cuda/*.cu
  make NVEXTRA="$v" all 2>&1 | grep -E " error" 
  BWA_B200_LANES=1 BWA_B200_CHUNK=100000000 python bench.py --layout se --steps 2 --warmup 2 --cpu-sample 2000 > /tmp/v.json 2>/dev/null
  python -c "import json; d=json.load(open('/tmp/v.json')); print('variant [$v]:', {k: round(v,1) for k,v in d['kernels_ms_per_step'].items()}, 'roof %.0f GB/s' % d['roofline']['achieved'])"
done
touch bwa_b200/csrc/cuda/*.cu; make all 2>&1 | grep " error"
