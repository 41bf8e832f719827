#!/bin/bash
# Round 2, second final call on one B200: the K1 with compact candidate lists (k_smem_c) against k_smem -- parity, A/B bench lines, the official line,
# ncu launch list + full capture of the changed kernels, and the long-read workloads it also changes (table lookups and more lanes for reads > 350 bases).
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2g_pytest.log; lap pytest
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    cb = d.get("cpu_baseline", {})
    print(sys.argv[1], "e2e %.0f reads/s, %.1f ms/step, value %.0f" % (d["e2e"]["value"], d["ms_per_step"], d["value"]), {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items()},
          "cpu %s x%s" % (round(cb.get("value") or 0), round(d["e2e"]["value"] / cb["value"], 1) if cb.get("value") else "?"), cb.get("sam_identical_on_sample"), cb.get("sam_identical_on_se_sample_with_options"), d.get("device_tail", {}).get("handed_back_to_host_postprocessing"),
          "roofline %.3f" % d["roofline"]["frac"], d["roofline"].get("kernel", "")[:40], {k: round(v["frac"], 3) for k, v in d.get("roofline_sw", {}).items() if isinstance(v, dict)})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
BWA_B200_K1_COMPACT=0 python bench.py --worker --steps 8 --warmup 3 --cpu-sample 20000 > $O/r2g_pe_k1old.json 2>$O/r2g_pe_k1old.err; echo "k_smem (BWA_B200_K1_COMPACT=0), 8 steps:"; line $O/r2g_pe_k1old.json; lap k1old
python bench.py --worker --steps 8 --warmup 3 --cpu-sample 20000 > $O/r2g_pe_k1c.json 2>$O/r2g_pe_k1c.err; echo "k_smem_c (default), 8 steps:"; line $O/r2g_pe_k1c.json; lap k1c
python bench.py --steps 20 --warmup 5 > $O/r2g_pe.json 2>$O/r2g_pe.err; echo "default workload, the driver's 20 steps (python bench.py --steps 20 --warmup 5):"; line $O/r2g_pe.json; lap pe
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2g_launches.csv python bench.py --worker --steps 1 --warmup 1 --cpu-sample 2000 > $O/r2g_launches.log 2>&1; lap ncu_launches
BWA_B200_SELFCHECK=0 BWA_B200_BENCH_VERIFY=0 BWA_B200_LANES=1 BWA_B200_CHUNK=1000000 timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:^(k_pack_reads|k_smem_c|k_seed_post)$' -s 3 -c 3 -o $O/r2g_ncu -f python bench.py --worker --inflight 1 --steps 1 --warmup 1 --cpu-sample 2000 > $O/r2g_ncu.log 2>&1; lap ncu_full
for w in len300 len1000 pacbio; do
  timeout 600 python bench.py --worker --workload $w --steps 4 --warmup 2 > $O/r2g_wl_$w.json 2>$O/r2g_wl_$w.err; echo "workload $w:"; line $O/r2g_wl_$w.json; lap wl_$w
done
BWA_B200_BENCH_WL_READS=8000 timeout 600 python bench.py --worker --workload pacbio --steps 3 --warmup 2 > $O/r2g_wl_pacbio8k.json 2>$O/r2g_wl_pacbio8k.err; echo "workload pacbio, 8000 reads per step:"; line $O/r2g_wl_pacbio8k.json; tail -2 $O/r2g_wl_pacbio8k.err | cut -c1-300; lap wl_pacbio8k
python bench.py --layout se --steps 12 --warmup 4 --cpu-sample 100000 > $O/r2g_se.json 2>/dev/null; echo "SE:"; line $O/r2g_se.json; lap se
ls -la $O/r2g_* | awk '{print $5, $9}'
