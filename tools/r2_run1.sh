#!/bin/bash
# Round 2, GPU call 1: parity on the shipped kernels, e2e vs stream-wait mode and host-thread budget, ncu of every shipped kernel.
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 600 python -m pytest tests -x -q -m gpu > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2a_pytest.log; lap pytest
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "e2e %.0f reads/s, %.1f ms/step" % (d["e2e"]["value"], d["ms_per_step"]), {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items()}, d.get("cpu_baseline", {}).get("sam_identical_on_sample"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
B="python bench.py --worker --steps 6 --warmup 3 --cpu-sample 2000"
for m in spin sleep yield; do BWA_B200_SYNC=$m $B > $O/r2a_pe_$m.json 2>$O/r2a_pe_$m.err; echo "wait=$m:"; line $O/r2a_pe_$m.json; done; lap sync_modes
for m in spin sleep; do BWA_B200_SYNC=$m $B --threads 12 > $O/r2a_pe_t12_$m.json 2>/dev/null; echo "wait=$m threads=12:"; line $O/r2a_pe_t12_$m.json; done; lap t12
BWA_B200_PROFILE=1 BWA_B200_LANES=1 BWA_B200_INFLIGHT=1 python bench.py --worker --inflight 1 --steps 2 --warmup 2 --cpu-sample 2000 > $O/r2a_prof.json 2> $O/r2a_prof.err; grep prof $O/r2a_prof.err | tail -60 > $O/r2a_prof_phases.txt; lap prof
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2a_launches.csv python bench.py --worker --steps 1 --warmup 1 --cpu-sample 2000 > $O/r2a_launches.log 2>&1; lap ncu_launches
BWA_B200_SELFCHECK=0 BWA_B200_LANES=1 BWA_B200_CHUNK=1000000 timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:^(k_smem|k_smem_fwd|k_sa|k_chain|k_extend_sm_fast|k_global_sm_fast)$' -s 12 -c 6 -o $O/r2a_ncu_all -f python bench.py --worker --inflight 1 --layout se --steps 1 --warmup 2 --cpu-sample 2000 > $O/r2a_ncu_all.log 2>&1; lap ncu_full
ls -la $O/r2a_* | awk '{print $5, $9}'
