#!/bin/bash
# 2 GPUs: why is a rank slower beside another one (88.8 vs 74.4 ms/step)?  The weak-scaling line with the ranks bound to their GPU's NUMA node
# (the default for N > 1) and unbound, and with fewer / more host threads per rank.
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out; N=2
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
nvidia-smi topo -m 2>/dev/null | head -8; lscpu | grep -i "numa\|socket" | head -6; cat /sys/fs/cgroup/cpu.max 2>/dev/null
run() { tag=$1; port=$2; bind=$3; shift 3; BWA_B200_BIND=$bind timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$port bench.py --gpus $N --steps 12 --warmup 4 --cpu-sample 20000 "$@" > $O/n2b_$tag.json 2> $O/n2b_$tag.err
  python -c "
import json; d=json.loads(open('$O/n2b_$tag.json').read().strip().splitlines()[-1]); print('$tag: e2e %.0f reads/s, %.1f ms/step, bound %s, threads %s' % (d['e2e']['value'], d['ms_per_step'], d['config'].get('numa_bound_cpus'), d['config'].get('host_threads_per_rank')))"; grep "bound" $O/n2b_$tag.err | head -2; lap $tag; }
run bound 31 1
run unbound 32 0
run bound_t6 33 1 --threads 6
run bound_t24 34 1 --threads 24
ls -la $O/n2b_* | awk '{print $5, $9}'
