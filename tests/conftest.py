import hashlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

REF_BWA = os.path.join(ROOT, "oracle", "_ref", "bwa")
KATDUMP = os.path.join(ROOT, "oracle", "_ref", "katdump")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
TESTBIN = os.path.join(ROOT, "tests", "_build", "bwa-b200-oracle")
CUSIMBIN = os.path.join(ROOT, "tests", "_build", "bwa-b200-cusim")
DATA = os.path.join(ROOT, "tests", "_build", "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _sh(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build the checkers (and, on the CPU box, the reference into oracle/_ref) once per session."""
    if not (os.path.exists(REF_BWA) and os.path.exists(ORACLE_SO)):
        _sh(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "all"])
    if not (os.path.exists(TESTBIN) and os.path.exists(CUSIMBIN)):
        _sh(["make", "-C", ROOT, "-j8", "testbin", "cusim"])
    return True


def strip_pg(b):
    return b"\n".join(l for l in b.split(b"\n") if not l.startswith(b"@PG"))


def ref_sam(args):
    """SAM of the unmodified reference binary (oracle/_ref/bwa mem ...), @PG removed."""
    r = subprocess.run([REF_BWA, "mem", "-v", "1"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
    return strip_pg(r.stdout)


def run_sam(binary, args):
    r = subprocess.run([binary, "mem", "-v", "1"] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    return strip_pg(r.stdout)


class DataSets:
    """Seeded synthetic datasets (tools/gen_data.py), generated and indexed (reference `bwa index`) on first use."""

    def __init__(self):
        os.makedirs(DATA, exist_ok=True)
        self._done = set()

    def ref(self, name):
        import gen_data
        fa = os.path.join(DATA, name + ".fa")
        if name not in self._done:
            if not os.path.exists(fa + ".bwt"):
                if name == "c1":
                    contigs = gen_data.random_contigs(1, 1000000, 7)
                elif name == "two":
                    contigs = gen_data.random_contigs(2, 150000, 9)
                elif name == "stress":
                    contigs = gen_data.stress_contigs(1500000, 5)
                else:
                    raise KeyError(name)
                gen_data.write_fasta(fa, contigs)
                _sh([REF_BWA, "index", fa])
            self._done.add(name)
        return fa

    def reads(self, ref, tag, n, length=150, seed=11, paired=False, err=(0.008, 0.001, 0.001), chimeric=0.0):
        import gen_data
        fa = self.ref(ref)
        base = os.path.join(DATA, "%s_%s" % (ref, tag))
        outs = [base + "_1.fq", base + "_2.fq"] if paired else [base + ".fq"]
        if not all(os.path.exists(o) for o in outs):
            contigs = gen_data.read_fasta(fa)
            r1, r2 = gen_data.gen_reads(contigs, n, length, seed, err=err, paired=paired, chimeric=chimeric)
            gen_data.write_fastq(outs[0], r1)
            if paired:
                gen_data.write_fastq(outs[1], r2)
        return fa, outs


@pytest.fixture(scope="session")
def data(built):
    return DataSets()


def md5(b):
    return hashlib.md5(b).hexdigest()
