"""ctypes view of the device-batch ABI (include/bwa_b200_dev.h) shared by the CPU and GPU tests."""
import ctypes as C

import numpy as np


class Intv(C.Structure):
    _fields_ = [("x", C.c_uint64 * 3), ("info", C.c_uint64)]


class SeedPar(C.Structure):
    _fields_ = [("min_seed_len", C.c_int), ("split_len", C.c_int), ("split_width", C.c_int), ("max_occ", C.c_int), ("max_mem_intv", C.c_uint64)]


class Seeds(C.Structure):
    _fields_ = [("intv_beg", C.POINTER(C.c_int64)), ("intv_n", C.POINTER(C.c_int32)), ("intv", C.POINTER(Intv)), ("seed_beg", C.POINTER(C.c_int64)),
                ("rbeg", C.POINTER(C.c_int64)), ("n_intv", C.c_int64), ("n_seeds", C.c_int64)]


def seed_stage(L, bwt, l_pac, pac, codes, off, par, ktab=None, touches=None):
    """ktab: depth of the short-string table to build first (None: no table, the plain rank path); touches: a list that
    receives the run's occ_touches counter"""
    L.bwag_ctx_create.restype = C.c_void_p
    L.bwag_ctx_create.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    L.bwag_batch_begin.restype = C.c_void_p
    L.bwag_batch_begin.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.bwag_seed.argtypes = [C.c_void_p, C.POINTER(SeedPar), C.POINTER(Seeds)]
    L.bwag_batch_end.argtypes = [C.c_void_p]
    L.bwag_ctx_destroy.argtypes = [C.c_void_p]
    ctx = L.bwag_ctx_create(-1, bwt, l_pac, pac)
    assert ctx
    if ktab is not None:
        L.bwag_ctx_build_ktab.argtypes = [C.c_void_p, C.c_int]
        assert L.bwag_ctx_build_ktab(ctx, ktab) == 0
    b = L.bwag_batch_begin(ctx, len(off) - 1, codes.ctypes.data, off.ctypes.data)
    out = Seeds()
    assert L.bwag_seed(b, C.byref(par), C.byref(out)) == 0
    res = []
    for r in range(len(off) - 1):
        iv = []
        for k in range(out.intv_beg[r], out.intv_beg[r] + out.intv_n[r]):
            x = out.intv[k]
            cnt = min(x.x[2], par.max_occ)
            iv.append((x.x[0], x.x[1], x.x[2], x.info, tuple(out.rbeg[out.seed_beg[k] + c] for c in range(cnt))))
        res.append(iv)
    L.bwag_batch_end(b)
    if touches is not None:      # the reference-equivalent Occ-block count of this run (bwag_stats_t.occ_touches, first field)
        st = (C.c_uint64 * 32)()
        L.bwag_stats_get.argtypes = [C.c_void_p, C.c_void_p]
        L.bwag_stats_get(ctx, st)
        touches.append(int(st[0]))
    L.bwag_ctx_destroy(ctx)
    return res




def load_reads_as_codes(fq, n):
    tab = np.full(256, 4, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        tab[ch] = i
        tab[ch + 32] = i
    seqs = [l.strip() for i, l in enumerate(open(fq, "rb")) if i % 4 == 1][:n]
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    codes = tab[np.frombuffer(b"".join(seqs), dtype=np.uint8)].copy()
    return codes, off
