#!/bin/bash
# Round 2, GPU call 7: warp-per-alignment K6 (parity + the repeat-rich workload), and K1 leaving room on the SMs for other lanes' K4/K5.
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests/test_localsw.py tests/test_tail.py tests/test_gpu_parity.py -x -q -m gpu > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2h_pytest.log; lap pytest
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    cb = d.get("cpu_baseline", {})
    print(sys.argv[1], "e2e %.0f reads/s, %.1f ms/step, value %.0f" % (d["e2e"]["value"], d["ms_per_step"], d["value"]), {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items()},
          "cpu %s x%s" % (round(cb.get("value") or 0), round(d["e2e"]["value"] / cb["value"], 1) if cb.get("value") else "?"), cb.get("sam_identical_on_sample"), d.get("device_tail", {}).get("handed_back_to_host_postprocessing"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
python bench.py --worker --steps 6 --warmup 3 --cpu-sample 20000 > $O/r2h_pe.json 2>$O/r2h_pe.err; echo "default:"; line $O/r2h_pe.json; lap pe
for nb in 4 3 2; do
  BWA_B200_K1_BLOCKS=$nb python bench.py --worker --steps 6 --warmup 3 --cpu-sample 2000 > $O/r2h_pe_k1b$nb.json 2>/dev/null; echo "K1_BLOCKS=$nb:"; line $O/r2h_pe_k1b$nb.json
done; lap k1blocks
BWA_B200_K1_BLOCKS=3 python bench.py --worker --inflight 4 --steps 6 --warmup 3 --cpu-sample 2000 > $O/r2h_pe_k1b3_if4.json 2>/dev/null; echo "K1_BLOCKS=3 inflight 4:"; line $O/r2h_pe_k1b3_if4.json
BWA_B200_K1_BLOCKS=3 BWA_B200_LANES=4 python bench.py --worker --steps 6 --warmup 3 --cpu-sample 2000 > $O/r2h_pe_k1b3_l4.json 2>/dev/null; echo "K1_BLOCKS=3 lanes 4:"; line $O/r2h_pe_k1b3_l4.json; lap k1b3
BWA_B200_PROFILE=1 BWA_B200_LANES=1 timeout 1500 python bench.py --worker --inflight 1 --workload stress --steps 2 --warmup 1 --cpu-sample 2000 > $O/r2h_wl_stress_prof.json 2>$O/r2h_wl_stress_prof.err; echo "workload stress (1 lane, profile):"; line $O/r2h_wl_stress_prof.json; grep "\[prof\]" $O/r2h_wl_stress_prof.err | grep -v "loop\|extension:\|batch counters" | awk '{a[$2]+=$3; n[$2]++} END {for (k in a) printf "%-16s %10.1f ms  x%d\n", k, a[k], n[k]}' | sort -k2 -n -r | head -12; grep "loop" $O/r2h_wl_stress_prof.err | awk '{a[$3]+=$4} END {for (k in a) printf "loop %-12s %8.2f CPU-s\n", k, a[k]}' | sort -k3 -n -r | head -4; lap stress_prof
timeout 900 python bench.py --worker --workload stress --steps 3 --warmup 1 > $O/r2h_wl_stress.json 2>$O/r2h_wl_stress.err; echo "workload stress:"; line $O/r2h_wl_stress.json; lap stress
ls -la $O/r2h_* 2>/dev/null | awk '{print $5, $9}'
