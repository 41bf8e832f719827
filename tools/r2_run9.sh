#!/bin/bash
# Round 2, GPU call 9: the seed-level filter of long reads on the device (parity, then the long-read workloads with it on and off)
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests/test_tail.py tests/test_gpu_parity.py tests/test_localsw.py -x -q -m gpu > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2j_pytest.log; lap pytest
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    cb = d.get("cpu_baseline", {})
    print(sys.argv[1], "e2e %.0f reads/s, %.1f ms/step, value %.0f" % (d["e2e"]["value"], d["ms_per_step"], d["value"]), {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items()},
          "cpu %s x%s" % (round(cb.get("value") or 0), round(d["e2e"]["value"] / cb["value"], 1) if cb.get("value") else "?"), cb.get("sam_identical_on_sample"), cb.get("sam_diff"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
for wl in len1000 pacbio; do
  timeout 900 python bench.py --worker --workload $wl --steps 3 --warmup 1 > $O/r2j_wl_$wl.json 2>$O/r2j_wl_$wl.err; echo "workload $wl:"; line $O/r2j_wl_$wl.json
  BWA_B200_DEVICE_SEEDSW=0 timeout 900 python bench.py --worker --workload $wl --steps 3 --warmup 1 --cpu-sample 200 > $O/r2j_wl_${wl}_hostflt.json 2>/dev/null; echo "workload $wl, filter on the host:"; line $O/r2j_wl_${wl}_hostflt.json
  lap $wl
done
BWA_B200_PROFILE=1 BWA_B200_LANES=1 timeout 600 python bench.py --worker --inflight 1 --workload len1000 --steps 2 --warmup 1 --cpu-sample 200 > $O/r2j_wl_len1000_prof.json 2>$O/r2j_wl_len1000_prof.err; grep "\[prof\]" $O/r2j_wl_len1000_prof.err | grep -v "loop\|extension:\|batch counters" | awk '{a[$2]+=$3; n[$2]++} END {for (k in a) printf "%-16s %10.1f ms  x%d\n", k, a[k], n[k]}' | sort -k2 -n -r | head -12; grep "batch counters" $O/r2j_wl_len1000_prof.err | tail -2 | cut -c1-300; lap len1000_prof
ls -la $O/r2j_* 2>/dev/null | awk '{print $5, $9}'
