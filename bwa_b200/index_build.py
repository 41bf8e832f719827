"""index_build.py -- GPU construction of the reference's on-disk index (.bwt .sa .pac .ann .amb).

Why it exists: BASELINE's headline workload is a 3 Gbp reference; the reference's own `bwa index` needs hours
for that (bwa.1:767), far beyond a benchmark run, and a GPU box keeps nothing between runs.  This module
produces THE SAME FILES (byte-identical, checked against `bwa index` in tests/) in seconds to minutes so that
both `bwa-b200 mem` and the unmodified `bwa mem` (CPU baseline) can load them.  It is input preparation, not
part of the timed hot path, and therefore leans on library
code (torch.sort) instead of hand-written kernels.

Method: the text is T = forward + reverse complement (bwtindex.c:267-274).  Its suffix array is obtained by
sorting suffixes on their first 32 bases (one 64-bit key) in buckets of the first two bases, then refining the
rare groups of equal keys with the following 32-base keys until no ties remain (a suffix that runs off the end
of the text sorts first, as with a terminal '$').  BWT = T[SA-1]; `primary` is the row of suffix 0; Occ
checkpoints every 128 symbols are interleaved as in bwt_bwtupdate_core (bwtindex.c:150-172); SA is sampled
every 32nd row (bwt.c:62-84).  'N' bases become lrand48() draws with seed 11 (bntseq.c:266,295).
"""
import os
import sys
import struct
import time

import numpy as np

OCC_INTERVAL = 128
SA_INTV = 32


def _lrand48_stream(n, seed=11):
    """glibc srand48(seed); n x (lrand48() & 3)."""
    x = (seed << 16) | 0x330E
    a, c, m = 0x5DEECE66D, 0xB, (1 << 48) - 1
    out = np.empty(n, dtype=np.uint8)
    for i in range(n):
        x = (a * x + c) & m
        out[i] = (x >> 17) & 3
    return out


def read_fasta(path):
    """[(name, comment, ascii uint8 array)] with the header split as kseq does (name up to the first white space)."""
    recs = []
    data = open(path, "rb").read()
    for block in data.split(b">")[1:]:
        nl = block.index(b"\n")
        hdr = block[:nl].rstrip(b"\r")
        parts = hdr.split(None, 1)
        name = parts[0] if parts else b""
        comment = b""
        if len(hdr) > len(name):
            comment = hdr[len(name) + 1:]
        seq = np.frombuffer(block[nl + 1:], dtype=np.uint8)
        seq = seq[(seq != 10) & (seq != 13)].copy()
        recs.append((name, comment, seq))
    return recs


_NT4 = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _NT4[_c] = _i
    _NT4[_c + 32] = _i
_NT4[ord("-")] = 5


def pack_reference(recs, prefix):
    """.pac/.ann/.amb exactly as bns_fasta2bntseq(for_only=1) + bns_dump write them; returns forward codes (uint8 0..3)."""
    n_n = 0
    for _, _, s in recs:
        n_n += int((_NT4[s] >= 4).sum())
    rnd = _lrand48_stream(n_n) if n_n else None
    rpos = 0
    anns, ambs, fwd = [], [], []
    offset = 0
    for name, comment, s in recs:
        c = _NT4[s]
        isn = c >= 4
        n_ambs = 0
        if isn.any():
            # holes: maximal runs of the same non-ACGT character (bntseq.c:246-262)
            idx = np.flatnonzero(isn)
            start = np.ones(len(idx), dtype=bool)
            start[1:] = (idx[1:] != idx[:-1] + 1) | (s[idx[1:]] != s[idx[:-1]])
            st = idx[start]
            ends = np.append(idx[np.flatnonzero(start)[1:] - 1], idx[-1])
            for b, e in zip(st, ends):
                ambs.append((offset + int(b), int(e - b + 1), chr(s[b])))
            n_ambs = len(st)
            c = c.copy()
            k = int(isn.sum())
            c[isn] = rnd[rpos:rpos + k]
            rpos += k
        anns.append((name, comment, offset, len(s), n_ambs))
        fwd.append(c)
        offset += len(s)
    l_pac = offset
    codes = np.concatenate(fwd) if len(fwd) > 1 else fwd[0]
    # .pac: 4 bases per byte, first base in the top bits; size always l_pac/4+1(+1)
    pad = (-l_pac) % 4
    cp = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)]) if pad else codes
    q = cp.reshape(-1, 4)
    pac = (q[:, 0] << 6 | q[:, 1] << 4 | q[:, 2] << 2 | q[:, 3]).astype(np.uint8)
    with open(prefix + ".pac", "wb") as f:
        f.write(pac.tobytes())
        if l_pac % 4 == 0:
            f.write(b"\0")
        f.write(bytes([l_pac % 4]))
    with open(prefix + ".ann", "w") as f:
        f.write("%d %d %u\n" % (l_pac, len(anns), 11))
        for name, comment, off, ln, na in anns:
            f.write("0 %s" % name.decode())
            anno = comment.decode() if len(comment) else "(null)"
            f.write(" %s\n" % anno)
            f.write("%d %d %d\n" % (off, ln, na))
    with open(prefix + ".amb", "w") as f:
        f.write("%d %d %u\n" % (l_pac, len(anns), len(ambs)))
        for off, ln, ch in ambs:
            f.write("%d %d %c\n" % (off, ln, ch))
    return codes


def _suffix_keys(W, pos, n, torch):
    """64-bit key (as order-preserving signed int64) of the 32 bases starting at pos; bases beyond n read as A."""
    w = pos >> 5
    s = (pos & 31) << 1
    hi = W[w]
    lo = W[w + 1]
    # (hi << s) | (lo >>> (64 - s)); torch shifts are arithmetic on int64, so mask the logical part
    left = hi << s
    rs = 64 - s
    lo_shift = torch.where(s == 0, torch.zeros_like(lo), (lo >> rs.clamp(max=63)) & ((torch.ones_like(lo) << s) - 1))
    key = left | lo_shift
    return key ^ (-0x8000000000000000)


def build_from_codes(codes_fwd, prefix, device="cuda", verbose=True):
    """Suffix-sort fwd+revcomp on the GPU and write <prefix>.bwt and <prefix>.sa."""
    import torch
    t0 = time.time()
    l_pac = len(codes_fwd)
    n = 2 * l_pac
    dev = torch.device(device)
    fwd = torch.from_numpy(codes_fwd).to(dev)
    T = torch.empty(n + 96, dtype=torch.uint8, device=dev)
    T[:l_pac] = fwd
    T[l_pac:n] = 3 - fwd.flip(0)
    T[n:] = 0
    del fwd
    # packed words: 32 bases per int64, first base in the top bits
    nw = (n + 96) // 32
    W = torch.zeros(nw + 1, dtype=torch.int64, device=dev)
    CH = 1 << 24
    sh = torch.arange(31, -1, -1, device=dev, dtype=torch.int64) * 2
    for b in range(0, nw, CH):
        e = min(nw, b + CH)
        blk = T[b * 32:e * 32].view(-1, 32).to(torch.int64)
        W[b:e] = (blk << sh).sum(dim=1)   # disjoint bit fields: the sum is the OR (wraps into the sign bit as intended)
        del blk
    if verbose:
        print("[index_build] text of %d bases packed in %.1fs" % (n, time.time() - t0), file=sys.stderr, flush=True)

    # BWT rows: row 0 is the empty suffix ('$'), rows 1..n the sorted suffixes
    bwt_chars = torch.empty(n + 1, dtype=torch.uint8, device=dev)
    bwt_chars[0] = T[n - 1]
    n_sa = (n + SA_INTV) // SA_INTV
    sa_samp = torch.zeros(n_sa, dtype=torch.int64, device=dev)
    sa_samp[0] = n
    primary = -1
    row = 1
    # buckets by the first `kb` bases, sized so that one bucket's sort fits comfortably in memory
    kb = 1
    while (n >> (2 * kb)) > (1 << 28) and kb < 6:
        kb += 1
    nb = 4 ** kb
    pos_all = None
    if nb <= 4:
        code_k = T[:n].to(torch.int16)
        for j in range(1, kb):
            code_k = code_k * 4 + T[j:n + j].to(torch.int16)
    else:
        code_k = T[:n].to(torch.int16)
        for j in range(1, kb):
            code_k = code_k * 4 + T[j:n + j].to(torch.int16)
    for b in range(nb):
        pos = torch.nonzero(code_k == b, as_tuple=False).flatten()
        if pos.numel() == 0:
            continue
        key = _suffix_keys(W, pos, n, torch)
        key, order = torch.sort(key)
        pos = pos[order]
        del order
        pos = _refine_ties(W, T, key, pos, n, torch, verbose)
        del key
        m = pos.numel()
        prev = pos - 1
        isz = pos == 0
        if bool(isz.any()):
            primary = row + int(torch.nonzero(isz).flatten()[0])
        prev = torch.where(isz, torch.zeros_like(prev), prev)
        bwt_chars[row:row + m] = T[prev]
        # suffix-array samples at rows that are multiples of 32
        first = (-row) % SA_INTV
        if first < m:
            sel = torch.arange(first, m, SA_INTV, device=dev)
            sa_samp[(row + sel) // SA_INTV] = pos[sel]
        row += m
        del pos, prev, isz
        if verbose:
            print("[index_build] bucket %d/%d done, %d rows, %.1fs" % (b + 1, nb, row, time.time() - t0), file=sys.stderr, flush=True)
    assert row == n + 1 and primary > 0
    del code_k, W

    # remove the '$' row, count, interleave Occ checkpoints (bwtindex.c:150-172)
    B = torch.cat([bwt_chars[:primary], bwt_chars[primary + 1:]])
    del bwt_chars
    L2 = [0] * 5
    cnt = torch.bincount(B.to(torch.int64), minlength=4).tolist()
    for c in range(4):
        L2[c + 1] = L2[c] + cnt[c]
    n_blk = (n + OCC_INTERVAL - 1) // OCC_INTERVAL
    padn = n_blk * OCC_INTERVAL - n
    Bp = torch.cat([B, torch.zeros(padn, dtype=torch.uint8, device=dev)]) if padn else B
    blk = Bp.view(n_blk, OCC_INTERVAL)
    occ = torch.zeros(n_blk + 1, 4, dtype=torch.int64, device=dev)
    valid = None
    if padn:
        valid = torch.ones(n_blk * OCC_INTERVAL, dtype=torch.bool, device=dev)
        valid[n:] = False
        valid = valid.view(n_blk, OCC_INTERVAL)
    for c in range(4):
        eq = blk == c
        if valid is not None:
            eq = eq & valid
        occ[1:, c] = eq.sum(dim=1).cumsum(0)
    out = torch.zeros(n_blk, 16, dtype=torch.int64, device=dev)   # 16 x u32 per block, kept in int64 lanes
    for c in range(4):
        out[:, 2 * c] = occ[:-1, c] & 0xffffffff
        out[:, 2 * c + 1] = occ[:-1, c] >> 32
    sh16 = torch.arange(15, -1, -1, device=dev, dtype=torch.int64) * 2
    for b0 in range(0, n_blk, 1 << 21):          # 16 symbols -> one 32-bit word, first symbol in the top bits
        b1 = min(n_blk, b0 + (1 << 21))
        out[b0:b1, 8:] = (blk[b0:b1].view(-1, 8, 16).to(torch.int64) << sh16).sum(dim=2)
    out32 = out.to(torch.int32).cpu().numpy().view(np.uint32).reshape(-1)
    # words of the last (partial) block beyond ceil(rem/16) are not part of the reference's file
    rem = n - (n_blk - 1) * OCC_INTERVAL
    n_last_words = (rem + 15) // 16
    body = out32[:(n_blk - 1) * 16 + 8 + n_last_words]
    tail = np.zeros(8, dtype=np.uint32)
    last = occ[-1].cpu().numpy()
    for c in range(4):
        tail[2 * c] = int(last[c]) & 0xffffffff
        tail[2 * c + 1] = int(last[c]) >> 32
    with open(prefix + ".bwt", "wb") as f:
        f.write(struct.pack("<5Q", primary, *L2[1:]))
        f.write(body.tobytes())
        f.write(tail.tobytes())
    sa = sa_samp.cpu().numpy().astype(np.uint64)
    with open(prefix + ".sa", "wb") as f:
        f.write(struct.pack("<5Q", primary, *L2[1:]))
        f.write(struct.pack("<2Q", SA_INTV, n))
        f.write(sa[1:].tobytes())
    if verbose:
        print("[index_build] %s.{bwt,sa} written, total %.1fs" % (prefix, time.time() - t0), file=sys.stderr, flush=True)


def _refine_ties(W, T, key, pos, n, torch, verbose):
    """Order suffixes whose first 32 bases are equal by their following bases."""
    m = pos.numel()
    if m < 2:
        return pos
    same = key[1:] == key[:-1]
    if not bool(same.any()):
        return pos
    # group id = index of the first element of each run of equal keys
    idx = torch.arange(m, device=pos.device)
    start = torch.ones(m, dtype=torch.bool, device=pos.device)
    start[1:] = ~same
    gid = torch.cummax(torch.where(start, idx, torch.zeros_like(idx)), 0).values
    tied = torch.zeros(m, dtype=torch.bool, device=pos.device)
    tied[1:] |= same
    tied[:-1] |= same
    sel = torch.nonzero(tied).flatten()          # slots (sorted) that belong to a tie group
    g = gid[sel]
    p = pos[sel]
    off = 32
    rounds = 0
    while sel.numel():
        q = p + off
        ended = q >= n                           # ran off the text: sorts before everything else in its group
        k2 = _suffix_keys(W, torch.where(ended, torch.zeros_like(q), q), n, torch)
        # order by (group, ended first -- the shorter suffix first: '$' sorts below every base, and two suffixes that ran off
        # the text in the same group are A-padded copies of each other --, then key): three stable sorts, least significant first
        o = torch.sort(k2, stable=True).indices
        ekey = torch.where(ended, -p, torch.full_like(p, 0x7fffffffffffffff))
        o = o[torch.sort(ekey[o], stable=True).indices]
        o = o[torch.sort(g[o], stable=True).indices]
        g, p, k2, ended = g[o], p[o], k2[o], ended[o]
        pos[sel] = p                             # slots of a group are contiguous and sel is sorted, so this writes each group in its new order
        # still tied: same group, same key, neither ended
        eq = (g[1:] == g[:-1]) & (k2[1:] == k2[:-1]) & ~ended[1:] & ~ended[:-1]
        if not bool(eq.any()):
            break
        # new groups = runs of eq; their ids must stay unique and ordered: use the slot index of the run start
        mm = sel.numel()
        st = torch.ones(mm, dtype=torch.bool, device=pos.device)
        st[1:] = ~eq
        ii = torch.arange(mm, device=pos.device)
        run = torch.cummax(torch.where(st, ii, torch.zeros_like(ii)), 0).values
        still = torch.zeros(mm, dtype=torch.bool, device=pos.device)
        still[1:] |= eq
        still[:-1] |= eq
        keep = torch.nonzero(still).flatten()
        sel, p, g = sel[keep], p[keep], sel[run[keep]]
        off += 32
        rounds += 1
        if rounds > 1 << 20:
            raise RuntimeError("tie refinement does not terminate")
    return pos


def build(fa, prefix=None, device="cuda", verbose=True):
    """All five index files for a FASTA, byte-identical to `bwa index <fa>`."""
    prefix = prefix or fa
    recs = read_fasta(fa)
    codes = pack_reference(recs, prefix)
    build_from_codes(codes, prefix, device=device, verbose=verbose)


def build_from_contigs(contigs, prefix, names=None, device="cuda", verbose=True):
    """Same, from in-memory ASCII contigs (tools/gen_data.py) without writing a FASTA."""
    recs = [((names[i] if names else ("chr%d" % (i + 1))).encode(), b"", c) for i, c in enumerate(contigs)]
    codes = pack_reference(recs, prefix)
    build_from_codes(codes, prefix, device=device, verbose=verbose)


if __name__ == "__main__":
    import sys
    build(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, device=sys.argv[3] if len(sys.argv) > 3 else "cuda")
