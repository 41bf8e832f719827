#!/bin/bash
# Round 2, GPU call 6: after the seeding-pool fix, K3's dense filter, region hand-over, heavy-read routing, 3 calls in flight: parity, pacbio, stress, default.
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2g_pytest.log; lap pytest
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    cb = d.get("cpu_baseline", {})
    print(sys.argv[1], "e2e %.0f reads/s, %.1f ms/step, value %.0f" % (d["e2e"]["value"], d["ms_per_step"], d["value"]), {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items()},
          "cpu %s x%s" % (round(cb.get("value") or 0), round(d["e2e"]["value"] / cb["value"], 1) if cb.get("value") else "?"), cb.get("sam_identical_on_sample"), cb.get("sam_identical_on_se_sample_with_options"), cb.get("sam_diff"), d.get("device_tail", {}).get("handed_back_to_host_postprocessing"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
python bench.py --worker --steps 8 --warmup 3 --cpu-sample 20000 > $O/r2g_pe.json 2>$O/r2g_pe.err; echo "default:"; line $O/r2g_pe.json; lap pe
timeout 900 python bench.py --worker --workload pacbio --steps 3 --warmup 1 > $O/r2g_wl_pacbio.json 2>$O/r2g_wl_pacbio.err; echo "workload pacbio:"; line $O/r2g_wl_pacbio.json; tail -3 $O/r2g_wl_pacbio.err | cut -c1-300; lap pacbio
timeout 900 python bench.py --worker --workload len1000 --steps 3 --warmup 1 > $O/r2g_wl_len1000.json 2>/dev/null; echo "workload len1000:"; line $O/r2g_wl_len1000.json; lap len1000
BWA_B200_PROFILE=1 BWA_B200_LANES=1 timeout 1500 python bench.py --worker --inflight 1 --workload stress --steps 2 --warmup 1 > $O/r2g_wl_stress_prof.json 2>$O/r2g_wl_stress_prof.err; echo "workload stress (1 lane, profile):"; line $O/r2g_wl_stress_prof.json; grep "\[prof\]" $O/r2g_wl_stress_prof.err | grep -v "loop\|extension:\|batch counters" | awk '{a[$2]+=$3; n[$2]++} END {for (k in a) printf "%-16s %10.1f ms  x%d\n", k, a[k], n[k]}' | sort -k2 -n -r | head -16; grep "loop" $O/r2g_wl_stress_prof.err | awk '{a[$3]+=$4} END {for (k in a) printf "loop %-12s %8.2f CPU-s\n", k, a[k]}' | sort -k3 -n -r | head -5; lap stress_prof
timeout 900 python bench.py --worker --workload stress --steps 3 --warmup 1 > $O/r2g_wl_stress.json 2>$O/r2g_wl_stress.err; echo "workload stress:"; line $O/r2g_wl_stress.json; lap stress
ls -la $O/r2g_* $O/sam_diff* 2>/dev/null | awk '{print $5, $9}'
