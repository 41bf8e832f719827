"""Multi-GPU `mem`: one process per GPU (torch.distributed), reads dealt over the ranks, no hot-path collective.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        -m bwa_b200.multi [bwa mem options] -o out.sam ref.fa reads_1.fq [reads_2.fq]

What is distributed:
  - the index: rank 0 loads the files of `bwa index` and fills the device blob; ONE broadcast (NCCL over
    NVLink) gives every GPU its copy; the other ranks only read .ann/.amb/.pac for the SAM text;
  - the reads: batch b (the batches a single-GPU run would form, -K bases each) is aligned by rank b % world, so
    batch boundaries -- and the per-batch insert-size model of paired-end data -- do not depend on the number
    of GPUs: the merged SAM is the single-GPU SAM.  Striped ingest (plan_batches): each rank scans one byte
    stripe of the input for record lengths, the ranks exchange those, compute the same boundaries and then
    parse only the byte ranges of their own batches -- nobody parses more than 1/world of the input.  Inputs
    that cannot be cut by byte offset (gzip, stdin, wrapped lines) fall back to every rank parsing everything;
  - the output: each rank writes its batches to a part file, rank 0 merges the parts in batch order.
The alignment itself is the C library's `main_mem` (bb_cli.c), unchanged; this module is only plumbing.
With BWA_B200_LIB pointing at the CPU-emulated build (tests/_build/libbwa_b200_cusim.so) the same code runs
on the gloo backend, which is how the N>1 logic is tested without GPUs.
"""
import ctypes as C
import os
import sys

import bwa_b200

BWA_IDX_BNS, BWA_IDX_PAC, BWA_IDX_ALL = 2, 4, 7


def shard_of(batch_no, world):
    """Rank that aligns batch `batch_no` (mirrors bb_cli.c)."""
    return batch_no % world


def _bind(L):
    L.bwag_blob_bytes.restype = C.c_size_t
    L.bwag_blob_bytes.argtypes = [C.c_void_p, C.c_int64]
    L.bwag_blob_fill.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.bwag_ctx_from_blob.restype = C.c_void_p
    L.bwag_ctx_from_blob.argtypes = [C.c_int, C.c_void_p, C.c_int]
    L.bb_device_adopt2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bb_cli_set_index.argtypes = [C.c_void_p]
    L.bwa_idx_load.restype = C.POINTER(bwa_b200.BwaIdx)
    L.bwa_idx_load.argtypes = [C.c_char_p, C.c_int]


def replicate_index(L, prefix, rank, device, dist=None, on_gpu=True):
    """Load the index on rank 0, broadcast its device blob, make it resident on every rank.

    Returns (idx pointer, keep-alive objects).  `device` is the CUDA ordinal of this rank."""
    import torch
    _bind(L)
    world = dist.get_world_size() if dist is not None else 1
    idx = L.bwa_idx_load(prefix.encode(), BWA_IDX_ALL if rank == 0 else BWA_IDX_BNS | BWA_IDX_PAC)
    if not idx:
        raise RuntimeError("cannot load index %s" % prefix)
    i = idx.contents
    l_pac = C.cast(i.bns, C.POINTER(C.c_int64))[0]
    keep = []
    if world == 1:
        L.bb_device_attach.restype = C.c_void_p
        L.bb_device_attach.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bb_device_attach(i.bwt, i.bns, i.pac)
        return idx, keep
    dev = torch.device("cuda", device) if on_gpu else torch.device("cpu")
    nbytes = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == 0:
        nbytes[0] = L.bwag_blob_bytes(i.bwt, l_pac)
    dist.broadcast(nbytes, 0)
    blob = torch.empty(int(nbytes[0]), dtype=torch.uint8, device=dev)
    if rank == 0 and L.bwag_blob_fill(device, blob.data_ptr(), i.bwt, l_pac, i.pac) != 0:
        raise RuntimeError(L.bwag_last_error().decode())
    dist.broadcast(blob, 0)          # the only collective of a run
    ctx = L.bwag_ctx_from_blob(device, blob.data_ptr(), 0)
    if not ctx:
        raise RuntimeError(L.bwag_last_error().decode())
    if rank != 0:                    # no FM-index in host memory here: the host code only needs a key for the resident copy
        key = C.create_string_buffer(256)
        i.bwt = C.cast(key, C.c_void_p)
        keep.append(key)
    L.bb_device_adopt2(i.bwt, i.bns, i.pac, ctx)   # densifies the SA sample, builds the short-string table, runs the start-up self-check
    keep.append(blob)
    return idx, keep


class Stripe(C.Structure):
    _fields_ = [("n", C.c_int64), ("len", C.POINTER(C.c_int32)), ("off", C.POINTER(C.c_int64))]


class PlannedBatch(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("no", "n_before", "beg1", "end1", "beg2", "end2")]


def _getopt(argv):
    """(options {flag: last value or True}, positionals) the way getopt sees argv for bb_cli.c's option string."""
    flags_with_arg = set("kcvsrtRABOEUwLdTQDmINofWxGhyKXHFz")   # the ':' options of bb_cli.c's getopt string
    opts, pos, k = {}, [], 0
    while k < len(argv):
        a = argv[k]
        if a == "--":
            pos += argv[k + 1:]
            break
        if a.startswith("-") and len(a) > 1:
            j = 1
            while j < len(a):
                c = a[j]
                if c in flags_with_arg:
                    if j + 1 < len(a):
                        opts[c] = a[j + 1:]
                    elif k + 1 < len(argv):
                        k += 1
                        opts[c] = argv[k]
                    break
                opts[c] = True
                j += 1
        else:
            pos.append(a)
        k += 1
    return opts, pos


def batch_bounds(pair_bases, chunk):
    """Batch boundaries from the bases of each record PAIR (two files: mate 1 + mate 2; one file: records 2k and 2k+1),
    the rule of bseq_read (bwa.c:79-112): a batch ends after the first pair that brings its bases to >= chunk.
    Returns the pair index each batch starts at, plus the total as the last element."""
    import numpy as np
    cum = np.cumsum(np.asarray(pair_bases, dtype=np.int64))
    starts, s = [0], 0
    n = len(cum)
    while s < n:
        before = int(cum[s - 1]) if s else 0
        e = int(np.searchsorted(cum, before + chunk, side="left"))   # first pair p with cum[p] - before >= chunk
        s = min(e, n - 1) + 1
        starts.append(s)
    return starts


def plan_batches(L, files, chunk, rank, world, dist=None, device="cpu"):
    """Striped ingest: the (batch no, reads before, byte ranges) of the batches of `rank`, and the number of batches in all.
    None when the input does not qualify (every rank must then parse everything).  Collective: every rank calls it."""
    import numpy as np
    import torch
    L.bb_fq_plain_size.restype = C.c_int64
    L.bb_fq_plain_size.argtypes = [C.c_char_p]
    L.bb_fq_scan_stripe.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.POINTER(Stripe)]
    L.bb_fq_stripe_free.argtypes = [C.POINTER(Stripe)]

    def gather(t):   # variable-length all-gather of a 1-d int64 tensor
        if world == 1:
            return [t]
        n = torch.tensor([t.numel()], dtype=torch.int64, device=device)
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n)
        m = max(int(x) for x in ns)
        pad = torch.zeros(m, dtype=torch.int64, device=device)
        pad[:t.numel()] = t
        outs = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad)
        return [o[:int(k)] for o, k in zip(outs, ns)]

    sizes = [L.bb_fq_plain_size(f.encode()) for f in files]
    lens, offs, fit = [], [], all(s >= 0 for s in sizes)
    for f, size in zip(files, sizes):
        st = Stripe()
        if fit and L.bb_fq_scan_stripe(f.encode(), size * rank // world, size * (rank + 1) // world, C.byref(st)) != 0:
            fit = False
        if fit and st.n:
            lens.append(np.ctypeslib.as_array(st.len, shape=(st.n,)).astype(np.int64))
            offs.append(np.ctypeslib.as_array(st.off, shape=(st.n,)).copy())
        else:
            lens.append(np.zeros(0, dtype=np.int64))
            offs.append(np.zeros(0, dtype=np.int64))
        L.bb_fq_stripe_free(C.byref(st))
    ok = torch.tensor([1 if fit else 0], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok[0]) == 0:
        return None
    all_lens, first_rec = [], []     # per file: every record's length in file order; index of this rank's first record
    for k in range(len(files)):
        parts = [p.cpu().numpy() for p in gather(torch.from_numpy(lens[k]).to(device))]
        first_rec.append(sum(len(p) for p in parts[:rank]))
        all_lens.append(np.concatenate(parts))
    if len(files) == 2:
        if len(all_lens[0]) != len(all_lens[1]):
            return None                  # the reference stops at the shorter file with a warning: leave that to the classic reader
        pair_bases, per_pair = all_lens[0] + all_lens[1], [1, 1]
    else:
        a = all_lens[0]
        pair_bases = a[0::2].copy()
        pair_bases[:len(a) // 2] += a[1::2]
        per_pair = [2]
    if len(pair_bases) == 0:
        return None
    starts = batch_bounds(pair_bases, chunk)
    n_batches = len(starts) - 1
    # byte offset of every batch's first record, per file: the stripe owner knows it; max-reduce over the ranks
    begs = []
    for k in range(len(files)):
        rec = np.asarray(starts[:-1], dtype=np.int64) * per_pair[k]
        t = torch.full((n_batches,), -1, dtype=torch.int64)
        mine = (rec >= first_rec[k]) & (rec < first_rec[k] + len(offs[k]))
        t[torch.from_numpy(mine)] = torch.from_numpy(offs[k][rec[mine] - first_rec[k]])
        t = t.to(device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        b = t.cpu().numpy()
        if (b < 0).any():
            raise RuntimeError("striped ingest: a batch start was found by no rank")
        begs.append(np.concatenate([b, [sizes[k]]]))
    recs_per_pair = sum(per_pair)
    n_records = sum(len(x) for x in all_lens)
    mine = []
    for b in range(rank, n_batches, world):
        pb = PlannedBatch(no=b, n_before=min(starts[b] * recs_per_pair, n_records), beg1=int(begs[0][b]), end1=int(begs[0][b + 1]),
                          beg2=int(begs[1][b]) if len(files) == 2 else 0, end2=int(begs[1][b + 1]) if len(files) == 2 else 0)
        mine.append(pb)
    return (PlannedBatch * max(len(mine), 1))(*mine), len(mine), n_batches


def merge_parts(out_path, parts):
    """parts: [(sam part file, its index file)] by rank; writes the header, then the batches in batch order."""
    where = {}
    hdr = 0
    for r, (part, idxf) in enumerate(parts):
        pos = 0
        for line in open(idxf):
            no, nbytes = (int(x) for x in line.split())
            if no < 0:
                if r == 0:
                    hdr = nbytes
                pos = nbytes
            else:
                where[no] = (r, pos, nbytes)
                pos += nbytes
    files = [open(p, "rb") for p, _ in parts]
    with open(out_path, "wb") as out:
        out.write(files[0].read(hdr))
        for no in sorted(where):
            r, pos, nbytes = where[no]
            files[r].seek(pos)
            left = nbytes
            while left > 0:
                buf = files[r].read(min(left, 1 << 24))
                if not buf:
                    raise RuntimeError("part file %s is shorter than its index says" % parts[r][0])
                out.write(buf)
                left -= len(buf)
    for f in files:
        f.close()
    if sorted(where) != list(range(len(where))):
        raise RuntimeError("batches missing from the parts: have %s" % sorted(where))


def main(argv=None):
    import torch
    import torch.distributed as dist
    argv = list(sys.argv[1:] if argv is None else argv)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lib_path = os.environ.get("BWA_B200_LIB")
    L = bwa_b200.lib(lib_path)
    on_gpu = lib_path is None or "cusim" not in os.path.basename(lib_path)
    if on_gpu:
        if not torch.cuda.is_available():
            raise SystemExit("bwa_b200.multi: no CUDA device; there is no CPU path")
        torch.cuda.set_device(local_rank)
        if world > 1 and os.environ.get("BWA_B200_BIND", "0") == "1":   # opt-in (no gain measured on a 2-GPU box whose GPUs share a node)
            bwa_b200.bind_to_gpu_node(local_rank)   # host workers, lanes and pinned buffers next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl" if on_gpu else "gloo", **({"device_id": torch.device("cuda", local_rank)} if on_gpu else {}))
    # the output file and the index prefix are the only arguments this launcher looks at
    out = None
    for k, a in enumerate(argv):
        if a in ("-o", "-f") and k + 1 < len(argv):
            out = argv[k + 1]
    if out is None:
        raise SystemExit("bwa_b200.multi: -o FILE is required (every rank writes a part of it)")
    positional = _positionals(argv)
    if len(positional) < 2:
        raise SystemExit("usage: python -m bwa_b200.multi [bwa mem options] -o out.sam ref.fa reads_1.fq [reads_2.fq]")
    idx, keep = replicate_index(L, positional[0], rank, local_rank if on_gpu else 0, dist if world > 1 else None, on_gpu)
    part, part_idx = "%s.part%d" % (out, rank), "%s.part%d.idx" % (out, rank)
    os.environ["BWA_B200_RANK"], os.environ["BWA_B200_WORLD"], os.environ["BWA_B200_SHARD_IDX"] = str(rank), str(world), part_idx
    L.bb_cli_set_index(idx)
    # striped ingest when the input allows it (BWA_B200_STRIPED=0: every rank parses everything, the round-1 behaviour)
    plan = None
    opts, _ = _getopt(argv)
    if world > 1 and os.environ.get("BWA_B200_STRIPED", "1") != "0":
        chunk = int(opts["K"]) if "K" in opts and int(opts["K"]) > 0 else 10000000 * max(1, int(opts.get("t", 1)))
        files = positional[1:2] if ("p" in opts or len(positional) < 3) else positional[1:3]
        plan = plan_batches(L, files, chunk, rank, world, dist, torch.device("cuda", local_rank) if on_gpu else torch.device("cpu"))
        if plan is not None:
            L.bb_cli_set_plan.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
            L.bb_cli_set_plan(C.cast(plan[0], C.c_void_p), plan[1], plan[2])
            keep.append(plan)
    if world > 1 and rank == 0:
        print("[bwa_b200.multi] striped ingest: %s" % ("%d batches over %d ranks" % (plan[2], world) if plan else "off (switched off, or the input cannot be cut by byte offset): every rank parses everything"), file=sys.stderr)
    args = ["mem"] + [part if (k > 0 and argv[k - 1] in ("-o", "-f")) else a for k, a in enumerate(argv)]
    arr = (C.c_char_p * (len(args) + 1))(*[a.encode() for a in args], None)
    import time
    t_align = time.time()
    rc = L.main_mem(len(args), arr)
    t_align = time.time() - t_align
    ok = torch.tensor([0 if rc == 0 else 1])
    if world > 1:
        if on_gpu:
            ok = ok.cuda()
        dist.all_reduce(ok)      # also the barrier before the merge
    if int(ok[0]) != 0:
        raise SystemExit("bwa_b200.multi: a rank failed")
    if rank == 0:
        print("[bwa_b200.multi] rank 0 spent %.2f s in main_mem (index already replicated)" % t_align, file=sys.stderr)
    if rank == 0:
        parts = [("%s.part%d" % (out, r), "%s.part%d.idx" % (out, r)) for r in range(world)]
        merge_parts(out, parts)
        for p, q in parts:
            os.remove(p)
            os.remove(q)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _positionals(argv):
    """Non-option arguments the way getopt sees them for bb_cli.c's option string."""
    return _getopt(argv)[1]


if __name__ == "__main__":
    sys.exit(main())
