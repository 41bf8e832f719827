"""bench.py's B200 arm from end to end without a GPU: the CUDA kernels run in the SIMT emulator (tests/_build/libbwa_b200_cusim.so) and the
few torch.cuda calls the harness makes are stubbed.  Checks the contract of the JSON line (keys, the e2e block, the kernel named in
`roofline`, SAM identity of the sample against the reference binary) and that the pacing of the concurrent calls ran.  Numbers mean nothing here."""
import json
import os
import subprocess
import sys

from conftest import ROOT

DRIVER = r'''
import sys, os
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
_tensor = torch.tensor
torch.tensor = lambda *a, **k: _tensor(*a, **{x: y for x, y in k.items() if x != "device"})
import bwa_b200
bwa_b200.LIB_PATH = os.path.join(ROOT, "tests", "_build", "libbwa_b200_cusim.so")
bwa_b200.CLI_PATH = os.path.join(ROOT, "tests", "_build", "bwa-b200-cusim")
import importlib.util
spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
sys.argv = ["bench.py", "--worker", "--ref-mbp", "1", "--reads", "1200", "--steps", "4", "--warmup", "3", "--cpu-sample", "400", "--workdir", WORKDIR]
bench.main()
'''


def test_bench_worker_line_on_the_emulator(tmp_path):
    code = "ROOT = %r\nWORKDIR = %r\n" % (ROOT, str(tmp_path / "w")) + DRIVER
    env = dict(os.environ, BWA_B200_SELFCHECK="0")
    p = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=800, env=env, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr.decode()[-3000:]
    d = json.loads(p.stdout.decode().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "e2e", "gpu_launches", "clocks", "roofline", "roofline_sw", "cpu_baseline", "kernels_ms_per_step"):
        assert k in d, k
    assert d["steps"] == 4 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0 and "k_smem_c" in d["roofline"]["kernel"] and d["roofline"]["bound"] == "hbm"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["sam_identical_on_sample"] is True
    assert "paced" in d["config"]["pipeline"] and "workload" in d["config"]
