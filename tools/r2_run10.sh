#!/bin/bash
# Round 2, GPU call 10: K5L (lane-per-request global alignment) -- parity, A/B timing, one ncu capture of it; pacbio after skipping stage 4
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests/test_tail.py tests/test_gpu_parity.py -x -q -m gpu > $O/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2k_pytest.log; lap pytest
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    cb = d.get("cpu_baseline", {})
    print(sys.argv[1], "e2e %.0f reads/s, %.1f ms/step, value %.0f" % (d["e2e"]["value"], d["ms_per_step"], d["value"]), {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items()},
          "cpu %s x%s" % (round(cb.get("value") or 0), round(d["e2e"]["value"] / cb["value"], 1) if cb.get("value") else "?"), cb.get("sam_identical_on_sample"), cb.get("sam_identical_on_se_sample_with_options"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
python bench.py --worker --steps 12 --warmup 3 --cpu-sample 20000 > $O/r2k_pe.json 2>$O/r2k_pe.err; echo "default, 12 steps:"; line $O/r2k_pe.json; lap pe
BWA_B200_K5_LANE=0 python bench.py --worker --steps 12 --warmup 3 --cpu-sample 2000 > $O/r2k_pe_nok5l.json 2>/dev/null; echo "K5_LANE=0, 12 steps:"; line $O/r2k_pe_nok5l.json; lap nok5l
python bench.py --worker --layout se --steps 12 --warmup 3 --cpu-sample 2000 > $O/r2k_se.json 2>/dev/null; echo "SE, 12 steps:"; line $O/r2k_se.json; lap se
timeout 900 python bench.py --worker --workload pacbio --steps 3 --warmup 1 > $O/r2k_wl_pacbio.json 2>$O/r2k_wl_pacbio.err; echo "workload pacbio:"; line $O/r2k_wl_pacbio.json; lap pacbio
timeout 900 python bench.py --worker --workload len300 --steps 4 --warmup 2 --cpu-sample 2000 > $O/r2k_wl_len300.json 2>/dev/null; echo "workload len300:"; line $O/r2k_wl_len300.json; lap len300
BWA_B200_SELFCHECK=0 BWA_B200_BENCH_VERIFY=0 BWA_B200_LANES=1 BWA_B200_CHUNK=1000000 timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:^(k_global_lane|k_global_sm_fast)$' -s 2 -c 2 -o $O/r2k_ncu -f python bench.py --worker --inflight 1 --steps 1 --warmup 1 --cpu-sample 2000 > $O/r2k_ncu.log 2>&1; lap ncu_full
ls -la $O/r2k_* | awk '{print $5, $9}'
