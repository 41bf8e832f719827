"""K6 (bwag_localsw.cu): batched local Smith-Waterman with start recovery == the reference's ksw_align2 on the 400 committed
known-answer vectors (tests/golden/local.kat.gz: u8 and i16 kernels, KSW_XSUBO/XSTART/XBYTE combinations, several
scoring schemes), through the C ABI (bwag_localsw).  CPU: the SIMT-emulated kernel; -m gpu: the CUDA kernel."""
import ctypes as C
import gzip
import os

import numpy as np
import pytest

from conftest import ROOT

CODE = {c: i for i, c in enumerate("ACGT")}


class SwPar(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "w", "zdrop", "pen_clip5", "pen_clip3")] + [("mat", C.c_int8 * 25)]


class SwTask(C.Structure):
    _fields_ = [("t_beg", C.c_int64), ("q_beg", C.c_int64), ("tlen", C.c_int32), ("qlen", C.c_int32), ("xtra", C.c_uint32), ("flags", C.c_int32)]


class SwRes(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("score", "te", "qe", "score2", "te2", "tb", "qb")]


def _run(S, data):
    import bwa_b200
    fa = data.ref("two")
    idx = bwa_b200.lib().bwa_idx_load(fa.encode(), 7).contents
    l_pac = C.cast(idx.bns, C.POINTER(C.c_int64))[0]
    S.bwag_ctx_create.restype = C.c_void_p
    S.bwag_ctx_create.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    S.bwag_batch_begin.restype = C.c_void_p
    S.bwag_batch_begin.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    S.bwag_localsw.argtypes = [C.c_void_p, C.POINTER(SwPar), C.c_int, C.POINTER(SwTask), C.c_void_p, C.c_size_t, C.POINTER(C.POINTER(SwRes))]
    S.bwag_batch_end.argtypes = [C.c_void_p]
    S.bwag_ctx_destroy.argtypes = [C.c_void_p]
    S.bwag_last_error.restype = C.c_char_p
    ctx = S.bwag_ctx_create(-1, idx.bwt, l_pac, idx.pac)
    assert ctx
    read = np.zeros(32, dtype=np.uint8)
    off = np.array([0, 32], dtype=np.int64)
    b = S.bwag_batch_begin(ctx, 1, read.ctypes.data, off.ctypes.data)
    assert b
    # group the vectors by scoring scheme: one call (one launch) per scheme, many tasks each
    groups = {}
    for line in gzip.open(os.path.join(ROOT, "tests/golden/local.kat.gz"), "rt"):
        f = line.split()
        key = tuple(map(int, f[1:7]))
        groups.setdefault(key, []).append((int(f[7]), f[10], f[11], tuple(map(int, f[13:20]))))
    n = 0
    for (a, bb, od, ed, oi, ei), vecs in groups.items():
        par = SwPar(a, bb, od, ed, oi, ei, 100, 100, 5, 5)
        for i in range(5):
            for j in range(5):
                par.mat[i * 5 + j] = -1 if (i == 4 or j == 4) else (a if i == j else -bb)
        pool, tasks = bytearray(), (SwTask * len(vecs))()
        for k, (xtra, q, t, _) in enumerate(vecs):
            qo = len(pool); pool += bytes(CODE.get(c, 4) for c in q)
            to = len(pool); pool += bytes(CODE.get(c, 4) for c in t)
            tasks[k] = SwTask(to, qo, len(t), len(q), xtra, 0)
        buf = (C.c_uint8 * len(pool)).from_buffer(pool)
        out = C.POINTER(SwRes)()
        rc = S.bwag_localsw(b, C.byref(par), len(vecs), tasks, buf, len(pool), C.byref(out))
        assert rc == 0, S.bwag_last_error()
        for k, (_, q, _, want) in enumerate(vecs):
            r = out[k]
            assert (r.score, r.te, r.qe, r.score2, r.te2, r.tb, r.qb) == want, (a, bb, od, ed, oi, ei, q[:30])
            n += 1
    assert n == 400
    S.bwag_batch_end(b)
    S.bwag_ctx_destroy(ctx)


@pytest.mark.parametrize("warp", ["1", "0"])
def test_localsw_kernel_emulated_equals_ksw_align2(built, data, warp, monkeypatch):
    """warp=1: a warp per alignment (vectors in shared memory); warp=0: a lane per alignment (the fallback for long queries)."""
    monkeypatch.setenv("BWA_B200_K6_WARP", warp)
    _run(C.CDLL(os.path.join(ROOT, "tests/_build/libbwa_b200_cusim.so"), mode=C.RTLD_LOCAL), data)


@pytest.mark.gpu
@pytest.mark.parametrize("warp", ["1", "0"])
def test_localsw_kernel_equals_ksw_align2(data, warp, monkeypatch):
    import bwa_b200
    monkeypatch.setenv("BWA_B200_K6_WARP", warp)
    _run(bwa_b200.lib(), data)
