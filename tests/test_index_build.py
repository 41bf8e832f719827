"""bwa_b200/index_build.py (input preparation for 3 Gbp benchmark references, where `bwa index` needs hours) against the
reference's `bwa index`, byte for byte, on small adversarial references: texts whose fwd+revcomp ends in runs of A (a
reference starting with T's), so that several suffixes shorter than the 32-base sort key have identical zero-padded
keys and only their lengths order them; periodic and single-base texts (deep tie refinement)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import REF_BWA

CASES = {
    "tt": "TT", "ttt": "TTT", "ttc": "TTC", "t40": "T" * 40, "aa_tail": "", "period": "", "polyA": "",
}


def _ref(name, rng):
    body = "".join("ACGT"[i] for i in rng.integers(0, 4, 3000))
    if name in ("tt", "ttt", "ttc", "t40"):
        return CASES[name] + body
    if name == "aa_tail":
        return "TTTT" + body + "AAAAAA"
    if name == "period":
        return "TTAGGG" * 300 + body[:500] + "TTAGGG" * 40
    return "A" * 700


def _build_both(tmp_path, name, device):
    import bwa_b200.index_build as ib
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000)
    seq = _ref(name, rng)
    fa = str(tmp_path / (name + ".fa"))
    with open(fa, "w") as f:
        f.write(">c1\n%s\n>c2 second contig\n%s\n" % (seq, seq[::-1][:800] if len(seq) > 900 else "ACGTTTGACA" * 7))
    mine = str(tmp_path / (name + "_mine.fa"))
    shutil.copy(fa, mine)
    subprocess.run([REF_BWA, "index", fa], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    recs = ib.read_fasta(mine)
    codes = ib.pack_reference(recs, mine)
    ib.build_from_codes(codes, mine, device=device, verbose=False)
    for ext in ("pac", "ann", "amb", "bwt", "sa"):
        assert open(fa + "." + ext, "rb").read() == open(mine + "." + ext, "rb").read(), (name, ext)


@pytest.mark.parametrize("name", sorted(CASES))
def test_index_builder_cpu_identical_to_bwa_index(built, tmp_path, name):
    _build_both(tmp_path, name, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_index_builder_gpu_identical_to_bwa_index(tmp_path, name):
    _build_both(tmp_path, name, "cuda")
