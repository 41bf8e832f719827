#!/bin/bash
# Round 2, GPU call 2: device tail (stage 4) + lane-per-read extension kernel: parity, A/B of both, host-thread budgets, ncu.
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r2b_pytest.log; lap pytest
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300; lap smoke
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
$B > $O/r2b_pe.json 2>$O/r2b_pe.err; echo "default:"; line $O/r2b_pe.json; lap pe
BWA_B200_K4_LANE=0 $B > $O/r2b_pe_nolane.json 2>/dev/null; echo "K4_LANE=0:"; line $O/r2b_pe_nolane.json
BWA_B200_TAIL=0 $B > $O/r2b_pe_notail.json 2>/dev/null; echo "TAIL=0:"; line $O/r2b_pe_notail.json
BWA_B200_TAIL=0 BWA_B200_K4_LANE=0 $B > $O/r2b_pe_r1path.json 2>/dev/null; echo "TAIL=0 K4_LANE=0:"; line $O/r2b_pe_r1path.json; lap ab
$B --layout se > $O/r2b_se.json 2>/dev/null; echo "SE:"; line $O/r2b_se.json; lap se
for t in 12 6; do $B --threads $t > $O/r2b_pe_t$t.json 2>/dev/null; echo "threads=$t:"; line $O/r2b_pe_t$t.json; done; lap threads
for m in yield sleep; do BWA_B200_SYNC=$m $B > $O/r2b_pe_$m.json 2>/dev/null; echo "wait=$m:"; line $O/r2b_pe_$m.json; done; lap sync
BWA_B200_LANES=3 $B > $O/r2b_pe_l3.json 2>/dev/null; echo "lanes=3:"; line $O/r2b_pe_l3.json
BWA_B200_INFLIGHT=3 $B --inflight 3 > $O/r2b_pe_if3.json 2>/dev/null; echo "inflight=3:"; line $O/r2b_pe_if3.json; lap lanes
BWA_B200_PROFILE=1 BWA_B200_LANES=1 python bench.py --worker --inflight 1 --steps 2 --warmup 2 --cpu-sample 2000 > $O/r2b_prof.json 2> $O/r2b_prof.err; grep prof $O/r2b_prof.err | tail -40 > $O/r2b_prof_phases.txt; lap prof
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2b_launches.csv python bench.py --worker --steps 1 --warmup 1 --cpu-sample 2000 > $O/r2b_launches.log 2>&1; lap ncu_launches
BWA_B200_SELFCHECK=0 BWA_B200_LANES=1 BWA_B200_CHUNK=1000000 timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:^(k_extend_lane|k_tail_regs|k_tail_sam|k_global_sm_fast)$' -s 8 -c 4 -o $O/r2b_ncu -f python bench.py --worker --inflight 1 --steps 1 --warmup 2 --cpu-sample 2000 > $O/r2b_ncu.log 2>&1; lap ncu_full
ls -la $O/r2b_* | awk '{print $5, $9}'
