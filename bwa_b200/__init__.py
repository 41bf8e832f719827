"""bwa_b200 -- thin Python binding (ctypes) over libbwa_b200.so, the B200-native BWA-MEM seed-and-extend path.

The product is the C-ABI shared library (include/bwa_b200.h, include/bwa_b200_dev.h) and the
`bwa-b200 mem` command line; this module only loads it for tests, the benchmark and multi-GPU
launching (torch.distributed).  Nothing here computes: every call goes to the library, and the library
has no CPU fallback -- importing works without a GPU, aligning does not.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libbwa_b200.so")
CLI_PATH = os.path.join(_HERE, "bwa-b200")


class MemOpt(C.Structure):
    """mem_opt_t (reference bwamem.h:52-84), same field order and types."""
    _fields_ = [
        ("a", C.c_int), ("b", C.c_int), ("o_del", C.c_int), ("e_del", C.c_int), ("o_ins", C.c_int), ("e_ins", C.c_int),
        ("pen_unpaired", C.c_int), ("pen_clip5", C.c_int), ("pen_clip3", C.c_int), ("w", C.c_int), ("zdrop", C.c_int),
        ("max_mem_intv", C.c_uint64), ("T", C.c_int), ("flag", C.c_int), ("min_seed_len", C.c_int),
        ("min_chain_weight", C.c_int), ("max_chain_extend", C.c_int), ("split_factor", C.c_float),
        ("split_width", C.c_int), ("max_occ", C.c_int), ("max_chain_gap", C.c_int), ("n_threads", C.c_int),
        ("chunk_size", C.c_int), ("mask_level", C.c_float), ("drop_ratio", C.c_float), ("XA_drop_ratio", C.c_float),
        ("mask_level_redun", C.c_float), ("mapQ_coef_len", C.c_float), ("mapQ_coef_fac", C.c_int), ("max_ins", C.c_int),
        ("max_matesw", C.c_int), ("max_XA_hits", C.c_int), ("max_XA_hits_alt", C.c_int), ("mat", C.c_int8 * 25),
    ]


class Bseq1(C.Structure):
    """bseq1_t (reference bwa.h:58-61)."""
    _fields_ = [("l_seq", C.c_int), ("id", C.c_int), ("name", C.c_void_p), ("comment", C.c_void_p),
                ("seq", C.c_void_p), ("qual", C.c_void_p), ("sam", C.c_void_p)]


class BwaIdx(C.Structure):
    """bwaidx_t (reference bwa.h:48-56)."""
    _fields_ = [("bwt", C.c_void_p), ("bns", C.c_void_p), ("pac", C.c_void_p), ("is_shm", C.c_int), ("l_mem", C.c_int64), ("mem", C.c_void_p)]


class Stats(C.Structure):
    """bwag_stats_t (include/bwa_b200_dev.h)."""
    _fields_ = [("occ_touches", C.c_uint64), ("sa_touches", C.c_uint64), ("sa_touches_algo", C.c_uint64), ("ext_cells", C.c_uint64),
                ("glb_cells", C.c_uint64), ("ms_smem", C.c_double), ("ms_sa", C.c_double), ("ms_extend", C.c_double), ("ms_global", C.c_double),
                ("ms_h2d", C.c_double), ("ms_d2h", C.c_double), ("n_launch", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("ms_chain", C.c_double),
                ("ms_tail", C.c_double), ("tail_reads", C.c_uint64), ("tail_complex", C.c_uint64),
                ("ms_localsw", C.c_double), ("sw_tasks", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


MEM_F_PE = 0x2

_lib = None


def lib(path=None):
    """Load the CUDA library.  Raises if it has not been built -- there is no other implementation."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError("%s is missing: build it with `make` (nvcc, sm_100a); bwa_b200 has no CPU fallback" % p)
    L = C.CDLL(p, mode=C.RTLD_LOCAL)
    L.mem_opt_init.restype = C.POINTER(MemOpt)
    L.bwa_idx_load.restype = C.POINTER(BwaIdx)
    L.bwa_idx_load.argtypes = [C.c_char_p, C.c_int]
    L.bwa_idx_destroy.argtypes = [C.POINTER(BwaIdx)]
    L.bwa_fill_scmat.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int8)]
    L.mem_process_seqs.argtypes = [C.POINTER(MemOpt), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(Bseq1), C.c_void_p]
    L.mem_process_seqs.restype = None
    L.bb_fq_open.restype = C.c_void_p
    L.bb_fq_open.argtypes = [C.c_char_p]
    L.bb_fq_close.argtypes = [C.c_void_p]
    L.bseq_read.restype = C.POINTER(Bseq1)
    L.bseq_read.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    L.bb_device_attach.restype = C.c_void_p
    L.bb_device_attach.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.bb_device_adopt.argtypes = [C.c_void_p, C.c_void_p]
    L.bwag_stats_get.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.bwag_stats_reset.argtypes = [C.c_void_p]
    L.bwag_ctx_densify_sa.argtypes = [C.c_void_p, C.c_int]
    L.bwag_last_error.restype = C.c_char_p
    L.main_mem.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    if path is None:
        _lib = L
    return L


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


class Index:
    """The on-disk index of the reference's `bwa index`, loaded with bwa_idx_load and kept resident in HBM."""

    def __init__(self, prefix, library=None):
        self.L = library or lib()
        self.p = self.L.bwa_idx_load(prefix.encode(), 7)
        if not self.p:
            raise RuntimeError("cannot load index %s" % prefix)
        self.ctx = None

    @classmethod
    def wrap(cls, p, library, keep=None):
        """An index that is already loaded (and possibly resident, see multi.replicate_index)."""
        self = cls.__new__(cls)
        self.L, self.p, self.ctx, self._keep = library, p, None, keep
        return self

    def attach(self):
        """Upload to the current CUDA device (no-op if already resident); returns the device context handle."""
        if self.ctx is None:
            i = self.p.contents
            self.ctx = self.L.bb_device_attach(i.bwt, i.bns, i.pac)
        return self.ctx

    def stats(self, reset=False):
        s = Stats()
        self.L.bwag_stats_get(self.attach(), C.byref(s))
        if reset:
            self.L.bwag_stats_reset(self.ctx)
        return s.as_dict()

    def densify_sa(self, intv):
        if self.L.bwag_ctx_densify_sa(self.attach(), intv) != 0:
            raise RuntimeError(self.L.bwag_last_error().decode())

    def verify(self, first=0, stride=1):
        """Check the resident FM-index against the resident text (bwag_ctx_verify); stride 1 = every row."""
        out = (C.c_uint64 * 4)()
        self.L.bwag_ctx_verify.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        if self.L.bwag_ctx_verify(self.attach(), first, stride, out) != 0:
            raise RuntimeError(self.L.bwag_last_error().decode())
        return {"rows": int(out[0]), "bwt_text_sa_mismatches": int(out[1]), "order_violations": int(out[2]), "undecided": int(out[3])}

    def close(self):
        if self.p:
            self.L.bwa_idx_destroy(self.p)
            self.p = None


class ReadBatch:
    """A batch of reads in host memory as the reference's bseq1_t array (bseq_read)."""

    def __init__(self, fq1, fq2=None, chunk=2**31 - 1, library=None):
        self.L = library or lib()
        f1 = self.L.bb_fq_open(fq1.encode())
        f2 = self.L.bb_fq_open(fq2.encode()) if fq2 else None
        if not f1 or (fq2 and not f2):
            raise RuntimeError("cannot open reads")
        n = C.c_int(0)
        self.seqs = self.L.bseq_read(chunk, C.byref(n), f1, f2)
        self.n = n.value
        self.paired = fq2 is not None
        self.L.bb_fq_close(f1)
        if f2:
            self.L.bb_fq_close(f2)

    def sam(self, free=True):
        """Concatenated SAM text produced by the last mem_process_seqs call."""
        self.L.bb_batch_cat_sam.restype = C.c_int64
        self.L.bb_batch_cat_sam.argtypes = [C.c_int, C.POINTER(Bseq1), C.c_void_p]
        n = self.L.bb_batch_cat_sam(self.n, self.seqs, None)
        buf = C.create_string_buffer(n + 1)
        self.L.bb_batch_cat_sam(self.n, self.seqs, buf)
        if free:
            self.L.bb_batch_free_sam(self.n, self.seqs)
        return buf.raw[:n]

    def total_bases(self):
        return sum(self.seqs[i].l_seq for i in range(self.n))


def mem_process_seqs(opt, index, batch, n_processed=0):
    """The drop-in boundary (reference bwamem.h:161): host buffers in, SAM text out, GPU in between."""
    i = index.p.contents
    index.L.mem_process_seqs(opt, i.bwt, i.bns, i.pac, n_processed, batch.n, batch.seqs, None)


def bind_to_gpu_node(device_index):
    """Restrict this process (and the threads it starts afterwards: host workers, lanes) to the CPUs of the NUMA node the GPU hangs
    off (sysfs local_cpulist of its PCI function), so that pinned buffers are first-touched next to the GPU and the copies do not cross
    the socket link.  Multi-GPU plumbing: one process per GPU.  Returns the CPU list used, or None if it could not be determined."""
    import torch
    try:
        pr = torch.cuda.get_device_properties(device_index)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= set(os.sched_getaffinity(0))
        if len(cpus) < 4:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def run_cli(args, stdout_path, stderr_path=None, binary=None, timeout=None):
    """Run `bwa-b200 mem ...` as a process."""
    with open(stdout_path, "wb") as so:
        se = open(stderr_path, "wb") if stderr_path else subprocess.DEVNULL
        try:
            return subprocess.call([binary or CLI_PATH] + list(args), stdout=so, stderr=se, timeout=timeout)
        finally:
            if stderr_path:
                se.close()
