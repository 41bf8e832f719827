#!/bin/bash
# Round 2, GPU call 4: lane kernel with warp-cooperative first rows and 3 blocks/SM, K5 backtrack bytes in shared memory, K6 (mate rescue on the device),
# then the other workloads of BASELINE.json (length sweep, pacbio) and a repeat-rich reference.
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2d_pytest.log; lap pytest
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
B="python bench.py --worker --steps 8 --warmup 3 --cpu-sample 20000"
$B > $O/r2d_pe.json 2>$O/r2d_pe.err; echo "default:"; line $O/r2d_pe.json; lap pe
BWA_B200_K4_LANE=0 $B > $O/r2d_pe_nolane.json 2>/dev/null; echo "K4_LANE=0:"; line $O/r2d_pe_nolane.json
BWA_B200_K5_ZSM=0 $B > $O/r2d_pe_nozsm.json 2>/dev/null; echo "K5_ZSM=0:"; line $O/r2d_pe_nozsm.json; lap ab
for w in len36 len75 len300 len1000 pacbio stress; do
  timeout 900 python bench.py --worker --workload $w --steps 4 --warmup 2 > $O/r2d_wl_$w.json 2>$O/r2d_wl_$w.err; echo "workload $w:"; line $O/r2d_wl_$w.json; lap wl_$w
done
BWA_B200_SELFCHECK=0 BWA_B200_LANES=1 BWA_B200_CHUNK=1000000 timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:^(k_extend_lane|k_global_sm_fast)$' -s 4 -c 2 -o $O/r2d_ncu -f python bench.py --worker --inflight 1 --steps 1 --warmup 2 --cpu-sample 2000 > $O/r2d_ncu.log 2>&1; lap ncu_full
ls -la $O/r2d_* | awk '{print $5, $9}'
