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
python bench.py --steps 20 --warmup 5 > $O/r2f_pe.json 2>$O/r2f_pe.err; echo "default workload, the driver's 20 steps (python bench.py --steps 20 --warmup 5):"; line $O/r2f_pe.json; lap pe
python bench.py > $O/r2f_pe_default.json 2>/dev/null; echo "python bench.py (no flags):"; line $O/r2f_pe_default.json; lap pe_default
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2f_ref.json 2>/dev/null; echo "reference arm:"; cut -c1-400 $O/r2f_ref.json; lap ref
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches.csv python bench.py --worker --steps 1 --warmup 1 --cpu-sample 2000 > $O/r2f_launches.log 2>&1; lap ncu_launches
BWA_B200_SELFCHECK=0 BWA_B200_BENCH_VERIFY=0 BWA_B200_LANES=1 BWA_B200_CHUNK=1000000 timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:^(k_pack_reads|k_smem|k_smem_fwd|k_seed_post|k_sa|k_chain|k_extend_lane|k_extend_sm_fast|k_tail_regs|k_global_sm_fast|k_tail_sam)$' -s 11 -c 11 -o $O/r2f_ncu -f python bench.py --worker --inflight 1 --steps 1 --warmup 1 --cpu-sample 2000 > $O/r2f_ncu.log 2>&1; lap ncu_full
python bench.py --layout se --steps 20 --warmup 5 --cpu-sample 100000 > $O/r2f_se.json 2>/dev/null; echo "SE:"; line $O/r2f_se.json; lap se
for w in len36 len75 len300 len1000 pacbio stress; do
  timeout 900 python bench.py --worker --workload $w --steps 4 --warmup 2 > $O/r2f_wl_$w.json 2>$O/r2f_wl_$w.err; echo "workload $w:"; line $O/r2f_wl_$w.json; lap wl_$w
done
# the index kept resident across processes: bwa-b200 mem on 20 k reads of the 3 Gbp workload, loading the index vs attaching to the keeper's copy
W=/tmp/bwa_b200_bench; FA=$W/ref_3000.fa; R1=$W/reads_pe1000000_150_e10_r0_1.fq; R2=$W/reads_pe1000000_150_e10_r0_2.fq
if [ -f $R1 ]; then
  head -40000 $R1 > /tmp/s_1.fq; head -40000 $R2 > /tmp/s_2.fq
  now() { date +%s.%N; }
  a=$(now); bwa_b200/bwa-b200 mem -v 1 -t 16 $FA /tmp/s_1.fq /tmp/s_2.fq > /tmp/a.sam 2> $O/r2f_res_cold.err; b=$(now); echo "bwa-b200 mem, index from the files: $(python -c "print('%.1f' % ($b-$a))") s wall"
  a=$(now); bwa_b200/bwa-b200 shm $FA 2> $O/r2f_res_shm.err; b=$(now); echo "bwa-b200 shm (make resident): $(python -c "print('%.1f' % ($b-$a))") s wall"; bwa_b200/bwa-b200 shm -l
  for k in 1 2; do a=$(now); bwa_b200/bwa-b200 mem -v 3 -t 16 $FA /tmp/s_1.fq /tmp/s_2.fq > /tmp/b$k.sam 2> $O/r2f_res_warm$k.err; b=$(now); echo "bwa-b200 mem, attached to the resident index: $(python -c "print('%.1f' % ($b-$a))") s wall ($(grep -c 'resident on the GPU' $O/r2f_res_warm$k.err) attach message)"; done
  echo "md5 of the three SAMs: $(grep -v '^@PG' /tmp/a.sam | md5sum | cut -c1-12) $(grep -v '^@PG' /tmp/b1.sam | md5sum | cut -c1-12) $(grep -v '^@PG' /tmp/b2.sam | md5sum | cut -c1-12)"
  bwa_b200/bwa-b200 shm -d; echo "after shm -d: $(bwa_b200/bwa-b200 shm -l | wc -l) resident, GPU memory used: $(nvidia-smi --query-gpu=memory.used --format=csv,noheader)"
fi
lap resident
BWA_B200_SELFCHECK=0 BWA_B200_BENCH_VERIFY=0 BWA_B200_LANES=1 timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:^(k_localsw_warp|k_chain_emit)$' -c 2 -o $O/r2f_ncu_k6 -f python bench.py --worker --workload len1000 --inflight 1 --steps 1 --warmup 1 --cpu-sample 200 > $O/r2f_ncu_k6.log 2>&1; lap ncu_k6
ls -la $O/r2f_* | awk '{print $5, $9}'
