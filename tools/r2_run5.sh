#!/bin/bash
# Round 2, GPU call 5: K1 occupancy variants (read kept as packed copy + N bitmap, 6/7/8 blocks per SM), calls-in-flight sweep.
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "e2e %.0f reads/s, %.1f ms/step, value %.0f" % (d["e2e"]["value"], d["ms_per_step"], d["value"]), {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items()}, d.get("cpu_baseline", {}).get("sam_identical_on_sample"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
make -j16 all > $O/r2e_build.log 2>&1; lap build
K="python bench.py --worker --layout se --steps 2 --warmup 2 --cpu-sample 20000"
export BWA_B200_BENCH_VERIFY=0
$K > $O/r2e_k1_default.json 2>/dev/null; echo "K1 default:"; line $O/r2e_k1_default.json; lap k1_default
for v in "6" "7" "8"; do
  rm -f build/cuda/bwag_smem.o build/cuda/bwag_api.o
  make -j16 NVEXTRA="-DK1_PACKED8 -DK1_SLOTS=4 -DK1_MIN_BLOCKS=$v" all >> $O/r2e_build.log 2>&1
  $K > $O/r2e_k1_packed_mb$v.json 2>$O/r2e_k1_packed_mb$v.err; echo "K1 packed read, $v blocks/SM:"; line $O/r2e_k1_packed_mb$v.json; lap k1_mb$v
done
for v in "-DK1_PREFETCH" "-DK1_PREFETCH -DK1_PACKED8 -DK1_SLOTS=4 -DK1_MIN_BLOCKS=6"; do
  rm -f build/cuda/bwag_smem.o build/cuda/bwag_api.o
  make -j16 NVEXTRA="$v" all >> $O/r2e_build.log 2>&1
  t=$(echo "$v" | tr -d ' =-' | tr 'D' '_'); $K > $O/r2e_k1$t.json 2>/dev/null; echo "K1 [$v]:"; line $O/r2e_k1$t.json; lap k1_pf
done
rm -f build/cuda/bwag_smem.o build/cuda/bwag_api.o; make -j16 all >> $O/r2e_build.log 2>&1; lap rebuild_default
B="python bench.py --worker --steps 8 --warmup 3 --cpu-sample 20000"
for i in 2 3 4; do BWA_B200_INFLIGHT=$i $B --inflight $i > $O/r2e_pe_if$i.json 2>/dev/null; echo "inflight=$i:"; line $O/r2e_pe_if$i.json; done; lap inflight
BWA_B200_INFLIGHT=3 BWA_B200_LANES=3 $B --inflight 3 > $O/r2e_pe_if3_l3.json 2>/dev/null; echo "inflight=3 lanes=3:"; line $O/r2e_pe_if3_l3.json
BWA_B200_CHUNK=131072 $B > $O/r2e_pe_chunk128k.json 2>/dev/null; echo "chunk=131072:"; line $O/r2e_pe_chunk128k.json; lap chunk
unset BWA_B200_BENCH_VERIFY
timeout 900 python bench.py --worker --workload pacbio --steps 2 --warmup 1 > $O/r2e_wl_pacbio.json 2>$O/r2e_wl_pacbio.err; echo "workload pacbio:"; line $O/r2e_wl_pacbio.json; python -c "import json; d=json.loads(open('$O/r2e_wl_pacbio.json').read().strip().splitlines()[-1]); print(json.dumps(d['cpu_baseline'].get('sam_diff'), indent=1)[:3000])"; lap pacbio
BWA_B200_PROFILE=1 BWA_B200_LANES=1 timeout 1500 python bench.py --worker --inflight 1 --workload stress --steps 2 --warmup 1 > $O/r2e_wl_stress_prof.json 2>$O/r2e_wl_stress_prof.err; echo "workload stress (1 lane, profile):"; line $O/r2e_wl_stress_prof.json; grep "\[prof\]" $O/r2e_wl_stress_prof.err | grep -v "loop\|extension:\|batch counters" | awk '{a[$2]+=$3; n[$2]++} END {for (k in a) printf "%-16s %10.1f ms  x%d\n", k, a[k], n[k]}' | sort -k2 -n -r | head -24; grep "loop" $O/r2e_wl_stress_prof.err | awk '{a[$3]+=$4} END {for (k in a) printf "loop %-12s %8.2f CPU-s\n", k, a[k]}' | sort -k3 -n -r | head; lap stress_prof
timeout 900 python bench.py --worker --workload stress --steps 3 --warmup 1 > $O/r2e_wl_stress.json 2>$O/r2e_wl_stress.err; echo "workload stress:"; line $O/r2e_wl_stress.json; lap stress
ls -la $O/r2e_* $O/sam_diff* 2>/dev/null | awk '{print $5, $9}'
