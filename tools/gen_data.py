#!/usr/bin/env python
"""Seeded synthetic references and reads for parity tests and benchmarks.

Recipes follow SURVEY.md section 8(d) / BASELINE.md section 3:
  * reference: uniform i.i.d. ACGT, one or more contigs (each < 2^31 bp);
  * `stress` reference: repeat family + tandem repeats + N runs + 2 contigs;
  * reads: uniform start, strand 50/50, per-base error split sub/ins/del, qualities 'I',
    names r<i>_<contig>_<pos> (truth encoded); PE: FR, insert ~N(400,50) clipped >= 160;
  * pacbio: 10 kbp, 10 % error split 20 % sub / 30 % del / 50 % ins.

Everything is vectorised numpy so that 3 Gbp / millions of reads generate in seconds to minutes.
"""
import argparse
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.zeros(256, dtype=np.uint8)
COMP[:] = ord("N")
for a, b in zip(b"ACGTN", b"TGCAN"):
    COMP[a] = b


def random_contigs(n_contigs, contig_len, seed):
    rng = np.random.default_rng(seed)
    return [ACGT[rng.integers(0, 4, size=contig_len, dtype=np.uint8)] for _ in range(n_contigs)]


def stress_contigs(total_len, seed):
    """Repeat-rich reference: 1 kb family x N copies at 3 % divergence, tandem repeats, N runs, 2 contigs."""
    rng = np.random.default_rng(seed)
    lens = [total_len * 6 // 10, total_len - total_len * 6 // 10]
    contigs = [ACGT[rng.integers(0, 4, size=l, dtype=np.uint8)].copy() for l in lens]
    fam = ACGT[rng.integers(0, 4, size=1000, dtype=np.uint8)]
    n_copies = max(20, total_len // 6700)
    for _ in range(n_copies):
        c = contigs[rng.integers(0, 2)]
        pos = rng.integers(0, len(c) - 1000)
        cp = fam.copy()
        mut = rng.random(1000) < 0.03
        cp[mut] = ACGT[rng.integers(0, 4, size=int(mut.sum()))]
        if rng.random() < 0.5:
            cp = COMP[cp[::-1]]
        c[pos:pos + 1000] = cp
    for _ in range(max(10, total_len // 50000)):
        c = contigs[rng.integers(0, 2)]
        unit = ACGT[rng.integers(0, 4, size=int(rng.integers(2, 30)))]
        reps = int(rng.integers(5, 60))
        tr = np.tile(unit, reps)
        pos = rng.integers(0, len(c) - len(tr))
        c[pos:pos + len(tr)] = tr
    for _ in range(max(4, total_len // 500000)):
        c = contigs[rng.integers(0, 2)]
        l = int(rng.integers(10, 500))
        pos = rng.integers(0, len(c) - l)
        c[pos:pos + l] = ord("N")
    return contigs


def write_fasta(path, contigs, prefix="chr"):
    with open(path, "wb") as f:
        for i, c in enumerate(contigs):
            f.write(b">%s%d\n" % (prefix.encode(), i + 1))
            n = len(c)
            w = 80
            full = n // w * w
            if full:
                body = np.empty((full // w, w + 1), dtype=np.uint8)
                body[:, :w] = c[:full].reshape(-1, w)
                body[:, w] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(c[full:].tobytes() + b"\n")


def _mutate(frag, rng, sub, ins, dele, out_len):
    """Apply substitution/insertion/deletion errors to one fragment (uint8 ASCII); returns exactly out_len bases
    when the fragment is long enough, else what is available."""
    n = len(frag)
    r = rng.random(n)
    keep = r >= dele  # deleted bases dropped
    is_sub = (r >= dele) & (r < dele + sub)
    f = frag.copy()
    ns = int(is_sub.sum())
    if ns:
        # substitute with a *different* base
        cur = np.searchsorted(ACGT, f[is_sub])
        cur = np.where(f[is_sub] == ord("N"), 0, cur)
        f[is_sub] = ACGT[(cur + rng.integers(1, 4, size=ns)) % 4]
    is_ins = (r >= dele + sub) & (r < dele + sub + ins)
    counts = keep.astype(np.int64) + is_ins.astype(np.int64)
    out = np.repeat(f, counts)
    # inserted base = the copy that precedes the original base; randomise it
    ends = np.cumsum(counts)
    ins_pos = (ends - counts)[is_ins & keep]
    if len(ins_pos):
        out[ins_pos] = ACGT[rng.integers(0, 4, size=len(ins_pos))]
    return out[:out_len]


def gen_reads(contigs, n, read_len, seed, err=(0.008, 0.001, 0.001), paired=False, ins_mean=400, ins_sd=50,
              chimeric=0.0, prefix="r"):
    rng = np.random.default_rng(seed)
    clens = np.array([len(c) for c in contigs], dtype=np.int64)
    sub, ins, dele = err
    slack = int(read_len * (dele * 3 + 0.02)) + 8
    recs1, recs2 = [], []
    span = read_len + slack
    cid = rng.choice(len(contigs), size=n, p=clens / clens.sum())
    if paired:
        isz = np.maximum(160, np.rint(rng.normal(ins_mean, ins_sd, size=n))).astype(np.int64)
        isz = np.maximum(isz, span)
    else:
        isz = np.full(n, span, dtype=np.int64)
    pos = (rng.random(n) * (clens[cid] - isz - 1)).astype(np.int64)
    strand = rng.integers(0, 2, size=n)
    chim = rng.random(n) < chimeric
    for i in range(n):
        c = contigs[cid[i]]
        frag = c[pos[i]:pos[i] + isz[i]]
        if strand[i]:
            frag = COMP[frag[::-1]]
        r1 = _mutate(frag[:span], rng, sub, ins, dele, read_len)
        if chim[i]:  # splice in a piece from elsewhere
            c2 = contigs[rng.integers(0, len(contigs))]
            p2 = int(rng.integers(0, len(c2) - read_len))
            cut = int(rng.integers(read_len // 3, 2 * read_len // 3))
            r1 = np.concatenate([r1[:cut], c2[p2:p2 + read_len - cut]])
        name = b"%s%d_%d_%d" % (prefix.encode(), i, cid[i] + 1, pos[i] + 1)
        recs1.append((name, r1))
        if paired:
            mate = COMP[frag[::-1]][:span]
            r2 = _mutate(mate, rng, sub, ins, dele, read_len)
            recs2.append((name, r2))
    return recs1, recs2


def write_fastq(path, recs):
    with open(path, "wb") as f:
        for name, seq in recs:
            s = seq.tobytes()
            f.write(b"@" + name + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    r = sub.add_parser("ref")
    r.add_argument("--out", required=True)
    r.add_argument("--contigs", type=int, default=1)
    r.add_argument("--len", type=int, default=1_000_000)
    r.add_argument("--seed", type=int, default=7)
    r.add_argument("--stress", action="store_true")
    q = sub.add_parser("reads")
    q.add_argument("--ref", required=True)
    q.add_argument("--out", required=True, help="output prefix: <out>.fq or <out>_1.fq/<out>_2.fq")
    q.add_argument("-n", type=int, default=10000)
    q.add_argument("--len", type=int, default=150)
    q.add_argument("--seed", type=int, default=11)
    q.add_argument("--paired", action="store_true")
    q.add_argument("--pacbio", action="store_true")
    q.add_argument("--err", type=float, default=0.01)
    q.add_argument("--chimeric", type=float, default=0.0)
    a = ap.parse_args()
    if a.cmd == "ref":
        contigs = stress_contigs(a.len, a.seed) if a.stress else random_contigs(a.contigs, a.len, a.seed)
        write_fasta(a.out, contigs)
    else:
        contigs = read_fasta(a.ref)
        if a.pacbio:
            err = (a.err * 0.2, a.err * 0.5, a.err * 0.3)
        else:
            err = (a.err * 0.8, a.err * 0.1, a.err * 0.1)
        r1, r2 = gen_reads(contigs, a.n, a.len, a.seed, err=err, paired=a.paired, chimeric=a.chimeric)
        if a.paired:
            write_fastq(a.out + "_1.fq", r1)
            write_fastq(a.out + "_2.fq", r2)
        else:
            write_fastq(a.out + ".fq", r1)


def read_fasta(path):
    contigs, cur = [], []
    with open(path, "rb") as f:
        data = f.read()
    for block in data.split(b">")[1:]:
        nl = block.index(b"\n")
        seq = np.frombuffer(block[nl + 1:], dtype=np.uint8)
        contigs.append(seq[seq != 10].copy())
    return contigs


if __name__ == "__main__":
    main()
