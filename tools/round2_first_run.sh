#!/bin/bash
# First GPU call of round 2: the lean K4/K5 row sweeps were written after round 1's GPU budget was spent, so (1) parity
# on the real kernels, (2) A/B of the kernels alone (same box, same index): lean vs first sweep, short-string table depths, (3) resident-block
# variants of the lean kernels (rebuilds bwag_extend.o / bwag_global.o with -DK4_MINB/-DK5_MINB), (4) ncu full capture
# of the lean K4 for the source-line view.  Everything lands in gpurun_out/r2_first_*.
cd /root/repo; mkdir -p gpurun_out; O=gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1 $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2_first_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_first_pytest.log; lap pytest
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "e2e %.0f reads/s, %.1f ms/step" % (d["e2e"]["value"], d["ms_per_step"]), {k: round(v, 1) for k, v in d["kernels_ms_per_step"].items()}, d.get("cpu_baseline", {}).get("sam_identical_on_sample"), "| self-check:", d["config"].get("device_selfcheck", "?")[:40])
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
B="python bench.py --worker --layout se --steps 3 --warmup 2 --cpu-sample 2000"
$B > $O/r2_first_lean.json 2>/dev/null; line $O/r2_first_lean.json; lap lean
BWA_B200_K4_FAST=0 BWA_B200_K5_FAST=0 $B > $O/r2_first_firstsweep.json 2>/dev/null; line $O/r2_first_firstsweep.json; lap first_sweep
for k in 0 10 12 13; do BWA_B200_KTAB=$k $B > $O/r2_first_ktab$k.json 2>/dev/null; echo "short-string table depth $k (default 14):"; line $O/r2_first_ktab$k.json; done; lap ktab
for mb in 4 5 8; do
  make -s NVEXTRA="-DK4_MINB=$mb -DK5_MINB=$mb" build/cuda/bwag_extend.o build/cuda/bwag_global.o -B > /dev/null 2>&1 && make -s all > /dev/null 2>&1
  $B > $O/r2_first_minb$mb.json 2>/dev/null; echo "min blocks $mb:"; line $O/r2_first_minb$mb.json; lap minb$mb
done
make -s build/cuda/bwag_extend.o build/cuda/bwag_global.o -B > /dev/null 2>&1 && make -s all > /dev/null 2>&1
for sa in 8 4 1; do BWA_B200_SA_INTV=$sa $B > $O/r2_first_sa$sa.json 2>/dev/null; echo "SA sample interval $sa (default 2):"; line $O/r2_first_sa$sa.json; done; lap sa_interval
for m in spin yield sleep; do BWA_B200_SYNC=$m python bench.py --worker --steps 4 --warmup 2 --cpu-sample 2000 > $O/r2_first_pe_$m.json 2>/dev/null; echo "stream wait = $m:"; line $O/r2_first_pe_$m.json; done; lap sync_modes
BWA_B200_LANES=1 BWA_B200_CHUNK=1000000 timeout 600 ncu --set full --import-source on --clock-control none -k regex:^k_extend_sm_fast\$ -s 2 -c 1 -o $O/r2_first_ncu_k_extend -f python bench.py --worker --layout se --steps 1 --warmup 2 --cpu-sample 2000 > $O/r2_first_ncu_k_extend.log 2>&1; lap ncu_k4
ls -la $O/r2_first_* | awk '{print $5, $9}'
