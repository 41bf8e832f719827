#!/bin/bash
# Round 2, final GPU call on one B200: parity, smoke, the default bench line, the other workloads, ncu launch list + full capture of every shipping kernel.
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2f_pytest.log; lap pytest
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-400; lap smoke
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    cb = d.get("cpu_baseline", {})
    print(sys.argv[1], "e2e %.0f reads/s, %.1f ms/step, value %.0f" % (d["e2e"]["value"], d["ms_per_step"], d["value"]), {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items()},
          "cpu %s x%s" % (round(cb.get("value") or 0), round(d["e2e"]["value"] / cb["value"], 1) if cb.get("value") else "?"), cb.get("sam_identical_on_sample"), cb.get("sam_identical_on_se_sample_with_options"), d.get("device_tail", {}).get("handed_back_to_host_postprocessing"),
          "roofline %.3f" % d["roofline"]["frac"], {k: round(v["frac"], 3) for k, v in d.get("roofline_sw", {}).items() if isinstance(v, dict)}, d["config"].get("index_verified"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
python bench.py > $O/r2f_pe.json 2>$O/r2f_pe.err; echo "default (python bench.py):"; line $O/r2f_pe.json; lap pe
python bench.py --layout se > $O/r2f_se.json 2>/dev/null; echo "SE:"; line $O/r2f_se.json; lap se
for w in len36 len75 len300 len1000 pacbio stress; do
  timeout 900 python bench.py --worker --workload $w --steps 4 --warmup 2 > $O/r2f_wl_$w.json 2>$O/r2f_wl_$w.err; echo "workload $w:"; line $O/r2f_wl_$w.json; lap wl_$w
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches.csv python bench.py --worker --steps 1 --warmup 1 --cpu-sample 2000 > $O/r2f_launches.log 2>&1; lap ncu_launches
BWA_B200_SELFCHECK=0 BWA_B200_BENCH_VERIFY=0 BWA_B200_LANES=1 BWA_B200_CHUNK=1000000 timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:^(k_pack_reads|k_smem|k_smem_fwd|k_seed_post|k_sa|k_chain|k_extend_lane|k_extend_sm_fast|k_tail_regs|k_global_sm_fast|k_tail_sam)$' -s 11 -c 11 -o $O/r2f_ncu -f python bench.py --worker --inflight 1 --steps 1 --warmup 1 --cpu-sample 2000 > $O/r2f_ncu.log 2>&1; lap ncu_full
ls -la $O/r2f_* | awk '{print $5, $9}'
