"""The N>1 path on CPU: two processes (torch.distributed, gloo), the index blob broadcast from rank 0, batches dealt
round-robin, parts merged by rank 0.  The kernels run in the CPU SIMT emulator build of the library, so the whole
multi-GPU control flow of bwa_b200/multi.py is exercised without a GPU; the merged SAM must equal the reference's."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, ref_sam

CUSIM_SO = os.path.join(ROOT, "tests/_build/libbwa_b200_cusim.so")


def _launch(world, args, out, port):
    env = dict(os.environ, BWA_B200_LIB=CUSIM_SO, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), "-m", "bwa_b200.multi"] + args[:-len(args) + args.index("--")] + ["-o", out] + args[args.index("--") + 1:]
    r = subprocess.run(cmd, env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-3000:]
    return open(out, "rb").read()


@pytest.mark.parametrize("paired", [False, True], ids=["se", "pe"])
def test_two_ranks_equal_reference(data, tmp_path, paired):
    fa, fqs = data.reads("stress", tag="mg" + ("pe" if paired else "se"), n=90, seed=77, paired=paired, err=(0.016, 0.002, 0.002), chimeric=0.05)
    opts = ["-K", "6000", "-t", "2"]           # ~40 reads per batch: several batches per rank
    got = _launch(2, opts + ["--", fa] + fqs, str(tmp_path / "out.sam"), 29631 + int(paired))
    want = ref_sam(opts + [fa] + fqs)
    assert got == want
    assert got.count(b"\n") > 90


def test_merge_parts_orders_batches(tmp_path):
    from bwa_b200.multi import merge_parts, shard_of
    parts = []
    text = {0: b"@HD\n", 1: b""}
    idx = {0: "-1 4\n", 1: "-1 0\n"}
    for b in range(5):
        r = shard_of(b, 2)
        rec = (b"batch%d\n" % b) * (b + 1)
        text[r] += rec
        idx[r] += "%d %d\n" % (b, len(rec))
    for r in (0, 1):
        p, q = tmp_path / ("p%d" % r), tmp_path / ("p%d.idx" % r)
        p.write_bytes(text[r]); q.write_text(idx[r])
        parts.append((str(p), str(q)))
    merge_parts(str(tmp_path / "m"), parts)
    assert (tmp_path / "m").read_bytes() == b"@HD\n" + b"".join((b"batch%d\n" % b) * (b + 1) for b in range(5))
