#!/bin/bash
# Round 2, GPU call 3: lane kernel after the converged first-row/band/cut-off rewrite, single-symbol extend + pre-packed reads in K1, small-superblock build.
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2c_pytest.log; lap pytest
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "e2e %.0f reads/s, %.1f ms/step, value %.0f" % (d["e2e"]["value"], d["ms_per_step"], d["value"]), {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items()}, d.get("cpu_baseline", {}).get("sam_identical_on_sample"), d.get("device_tail", {}).get("handed_back_to_host_postprocessing"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
B="python bench.py --worker --steps 8 --warmup 3 --cpu-sample 20000"
$B > $O/r2c_pe.json 2>$O/r2c_pe.err; echo "default:"; line $O/r2c_pe.json; lap pe
BWA_B200_K4_LANE=0 $B > $O/r2c_pe_nolane.json 2>/dev/null; echo "K4_LANE=0:"; line $O/r2c_pe_nolane.json
BWA_B200_SYNC=spin $B > $O/r2c_pe_spin.json 2>/dev/null; echo "spin:"; line $O/r2c_pe_spin.json; lap ab
$B --layout se > $O/r2c_se.json 2>/dev/null; echo "SE:"; line $O/r2c_se.json; lap se
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2c_launches.csv python bench.py --worker --steps 1 --warmup 1 --cpu-sample 2000 > $O/r2c_launches.log 2>&1; lap ncu_launches
BWA_B200_SELFCHECK=0 BWA_B200_LANES=1 BWA_B200_CHUNK=1000000 timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:^(k_smem|k_smem_fwd|k_extend_lane|k_pack_reads)$' -s 8 -c 4 -o $O/r2c_ncu -f python bench.py --worker --inflight 1 --steps 1 --warmup 2 --cpu-sample 2000 > $O/r2c_ncu.log 2>&1; lap ncu_full
ls -la $O/r2c_* | awk '{print $5, $9}'
