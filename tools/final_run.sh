#!/bin/bash
# round-end verification on one B200: parity tests, smoke, the default benchmark lines, two A/B probes, ncu launch list + full capture of K1
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests -x -q -m gpu > $O/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/final_pytest.log; lap pytest
if ! grep -q " passed" $O/final_pytest.log || grep -q "failed" $O/final_pytest.log; then
  BWA_B200_K5_SM=0 timeout 900 python -m pytest tests -x -q -m gpu > $O/final_pytest_k5off.log 2>&1; echo "pytest (K5 global scratch) rc=$?"; tail -2 $O/final_pytest_k5off.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; lap smoke
python bench.py > $O/final_pe.json 2> $O/final_pe.err; lap bench_pe
python bench.py --layout se > $O/final_se.json 2> $O/final_se.err; lap bench_se
BWA_B200_K5_SM=0 BWA_B200_K4_SM=0 python bench.py --worker --layout se --steps 2 --warmup 2 --cpu-sample 2000 > $O/final_se_nosm.json 2>/dev/null; lap bench_nosm
BWA_B200_LANES=3 python bench.py > $O/final_pe_l3.json 2>/dev/null; lap bench_l3
for f in final_pe final_se final_se_nosm final_pe_l3; do python - $O/$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "e2e %.0f reads/s, %.1f ms/step, kernels-only %.0f reads/s" % (d["e2e"]["value"], d["ms_per_step"], d["value"]), {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items()},
          "roofline %.0f GB/s" % d["roofline"]["achieved"], d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("sam_identical_on_sample"), d["config"].get("watchdog"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches.csv python bench.py --worker --steps 1 --warmup 1 --cpu-sample 2000 > $O/final_launches.log 2>&1; lap ncu_launches
BWA_B200_LANES=1 BWA_B200_CHUNK=1000000 timeout 600 ncu --set full --import-source on --clock-control none -k regex:^k_smem\$ -s 4 -c 1 -o $O/final_ncu_k_smem -f python bench.py --worker --layout se --steps 1 --warmup 3 --cpu-sample 2000 > $O/final_ncu_k_smem.log 2>&1; lap ncu_full
ls -la $O/final_* | awk '{print $5, $9}'
