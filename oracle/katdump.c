/* TEST INFRASTRUCTURE ONLY.
 *
 * katdump -- known-answer dumper.  Links the UNMODIFIED reference library (oracle/_ref/libbwa.a,
 * built from /root/reference by oracle/Makefile) and prints, as plain text, what the reference's
 * own functions return on given inputs.  tests/ diff these dumps against (a) the CPU restatement
 * in oracle/oracle_*.c and (b) the CUDA path.
 *
 *   katdump smem   <idx> <reads.fq> [max_reads]   bwt_smem1 / bwt_seed_strategy1 at every x  (bwt.c:289-379)
 *   katdump sa     <idx> <n> <seed>               bwt_sa at n pseudo-random rows               (bwt.c:86-96)
 *   katdump chain  <idx> <reads.fq> [opts]        mem_chain, then mem_chain_flt                 (bwamem.c:277-411)
 *   katdump regs   <idx> <reads.fq> [opts]        mem_align1_core                               (bwamem.c:1081-1117)
 *   katdump aln    <idx> <reads.fq> [opts]        mem_mark_primary_se + mem_reg2aln             (bwamem.c:547,1119)
 *   katdump extend <n> <seed>                     ksw_extend2 on seeded random inputs           (ksw.c:416-515)
 *   katdump global <n> <seed>                     ksw_global2 on seeded random inputs           (ksw.c:540-642)
 *   katdump local  <n> <seed>                     ksw_align2 on seeded random inputs            (ksw.c:379-401)
 *
 * [opts] = "pacbio" switches to the -x pacbio scoring preset (fastmap.c:337-354).
 * The random-input generators (splitmix64) are re-implemented identically in tests/katgen.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <zlib.h>
#include "bwa.h"
#include "bwamem.h"
#include "bwt.h"
#include "bntseq.h"
#include "ksw.h"
#include "kvec.h"
#include "kseq.h"
KSEQ_DECLARE(gzFile)

/* non-static reference functions that are not in the public headers */
typedef struct { int64_t rbeg; int32_t qbeg, len; int score; } kd_seed_t;               /* mem_seed_t  bwamem.c:194 */
typedef struct { int n, m, first, rid; uint32_t w:29, kept:2, is_alt:1; float frac_rep; int64_t pos; kd_seed_t *seeds; } kd_chain_t; /* bwamem.c:200 */
typedef struct { size_t n, m; kd_chain_t *a; } kd_chain_v;
extern kd_chain_v mem_chain(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, int len, const uint8_t *seq, void *buf);
extern int mem_chain_flt(const mem_opt_t *opt, int n_chn, kd_chain_t *a);
extern mem_alnreg_v mem_align1_core(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac, int l_seq, char *seq, void *buf);
extern int mem_mark_primary_se(const mem_opt_t *opt, int n, mem_alnreg_t *a, int64_t id);

static uint64_t sm_state;
static uint64_t sm_next(void)
{
	uint64_t z = (sm_state += 0x9e3779b97f4a7c15ULL);
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
	return z ^ (z >> 31);
}
static int sm_range(int lo, int hi) { return lo + (int)(sm_next() % (uint64_t)(hi - lo + 1)); } /* inclusive */

static mem_opt_t *make_opt(const char *mode)
{
	mem_opt_t *opt = mem_opt_init();
	if (mode && strcmp(mode, "pacbio") == 0) { /* fastmap.c:337-354 */
		opt->o_del = opt->e_del = opt->o_ins = opt->e_ins = 1; opt->b = 1;
		opt->split_factor = 10.; opt->min_chain_weight = 40; opt->min_seed_len = 17;
		opt->pen_clip5 = opt->pen_clip3 = 0;
	}
	bwa_fill_scmat(opt->a, opt->b, opt->mat);
	return opt;
}

static void to_codes(kseq_t *ks)
{
	size_t i;
	for (i = 0; i < ks->seq.l; ++i) ks->seq.s[i] = nst_nt4_table[(int)(unsigned char)ks->seq.s[i]];
}

static int do_smem(int argc, char **argv)
{
	bwaidx_t *idx = bwa_idx_load(argv[2], BWA_IDX_BWT);
	gzFile fp = gzopen(argv[3], "r");
	kseq_t *ks = kseq_init(fp);
	int max_reads = argc > 4 ? atoi(argv[4]) : 1 << 30, nr = 0;
	bwtintv_v mem = {0, 0, 0};
	while (nr < max_reads && kseq_read(ks) >= 0) {
		int x, len = ks->seq.l;
		size_t i;
		to_codes(ks);
		printf("R %s %d\n", ks->name.s, len);
		for (x = 0; x < len; ++x) {
			int mi, ret;
			bwtintv_t m;
			for (mi = 1; mi <= 3; mi += 2) { /* min_intv 1 and 3 */
				ret = bwt_smem1(idx->bwt, len, (uint8_t *)ks->seq.s, x, mi, &mem, 0);
				printf("S %d %d %d %d", x, mi, ret, (int)mem.n);
				for (i = 0; i < mem.n; ++i)
					printf(" %llu,%llu,%llu,%llu", (unsigned long long)mem.a[i].x[0], (unsigned long long)mem.a[i].x[1],
						   (unsigned long long)mem.a[i].x[2], (unsigned long long)mem.a[i].info);
				printf("\n");
			}
			ret = bwt_seed_strategy1(idx->bwt, len, (uint8_t *)ks->seq.s, x, 19, 20, &m);
			printf("T %d %d %llu,%llu,%llu,%llu\n", x, ret, (unsigned long long)m.x[0], (unsigned long long)m.x[1],
				   (unsigned long long)m.x[2], (unsigned long long)m.info);
		}
		++nr;
	}
	return 0;
}

static int do_sa(int argc, char **argv)
{
	bwaidx_t *idx = bwa_idx_load(argv[2], BWA_IDX_BWT);
	int i, n = atoi(argv[3]);
	sm_state = strtoull(argv[4], 0, 10);
	for (i = 0; i < n; ++i) {
		uint64_t k = sm_next() % (idx->bwt->seq_len + 1);
		printf("A %llu %llu\n", (unsigned long long)k, (unsigned long long)bwt_sa(idx->bwt, k));
	}
	return 0;
}

static void print_chains(const char *tag, int n, kd_chain_t *a)
{
	int i, j;
	for (i = 0; i < n; ++i) {
		printf("%s %d pos=%lld rid=%d n=%d w=%u kept=%u frac_rep=%.6f", tag, i, (long long)a[i].pos, a[i].rid, a[i].n, (unsigned)a[i].w, (unsigned)a[i].kept, a[i].frac_rep);
		for (j = 0; j < a[i].n; ++j)
			printf(" %lld,%d,%d,%d", (long long)a[i].seeds[j].rbeg, a[i].seeds[j].qbeg, a[i].seeds[j].len, a[i].seeds[j].score);
		printf("\n");
	}
}

static int do_chain(int argc, char **argv)
{
	bwaidx_t *idx = bwa_idx_load(argv[2], BWA_IDX_ALL);
	mem_opt_t *opt = make_opt(argc > 4 ? argv[4] : 0);
	gzFile fp = gzopen(argv[3], "r");
	kseq_t *ks = kseq_init(fp);
	while (kseq_read(ks) >= 0) {
		kd_chain_v c;
		int n;
		to_codes(ks);
		c = mem_chain(opt, idx->bwt, idx->bns, ks->seq.l, (uint8_t *)ks->seq.s, 0);
		printf("R %s %d\n", ks->name.s, (int)ks->seq.l);
		print_chains("C", c.n, c.a); /* w/kept uninitialised here -> printed but ignored by the differ for tag C */
		n = mem_chain_flt(opt, c.n, c.a);
		print_chains("F", n, c.a);
	}
	return 0;
}

static void print_regs(const mem_alnreg_v *r)
{
	size_t i;
	for (i = 0; i < r->n; ++i) {
		const mem_alnreg_t *p = &r->a[i];
		printf("G %d rb=%lld re=%lld qb=%d qe=%d rid=%d score=%d truesc=%d sub=%d csub=%d sub_n=%d w=%d seedcov=%d secondary=%d secondary_all=%d seedlen0=%d n_comp=%d is_alt=%d frac_rep=%.6f\n",
			   (int)i, (long long)p->rb, (long long)p->re, p->qb, p->qe, p->rid, p->score, p->truesc, p->sub, p->csub, p->sub_n, p->w, p->seedcov,
			   p->secondary, p->secondary_all, p->seedlen0, p->n_comp, p->is_alt, p->frac_rep);
	}
}

static int do_regs(int argc, char **argv, int with_aln)
{
	bwaidx_t *idx = bwa_idx_load(argv[2], BWA_IDX_ALL);
	mem_opt_t *opt = make_opt(argc > 4 ? argv[4] : 0);
	gzFile fp = gzopen(argv[3], "r");
	kseq_t *ks = kseq_init(fp);
	int64_t id = 0;
	while (kseq_read(ks) >= 0) {
		mem_alnreg_v r;
		size_t i;
		r = mem_align1_core(opt, idx->bwt, idx->bns, idx->pac, ks->seq.l, ks->seq.s, 0);
		printf("R %s %d\n", ks->name.s, (int)ks->seq.l);
		if (with_aln) mem_mark_primary_se(opt, r.n, r.a, id);
		print_regs(&r);
		if (with_aln) {
			for (i = 0; i < r.n; ++i) {
				mem_aln_t a = mem_reg2aln(opt, idx->bns, idx->pac, ks->seq.l, ks->seq.s, &r.a[i]);
				int k;
				printf("L %d rid=%d pos=%lld rev=%d mapq=%d NM=%d score=%d sub=%d flag=%d cigar=", (int)i, a.rid, (long long)a.pos, a.is_rev, a.mapq, a.NM, a.score, a.sub, a.flag);
				for (k = 0; k < a.n_cigar; ++k) printf("%d%c", a.cigar[k] >> 4, "MIDSH"[a.cigar[k] & 0xf]);
				printf(" MD=%s\n", a.n_cigar ? (char *)(a.cigar + a.n_cigar) : "");
				free(a.cigar);
			}
		}
		free(r.a);
		++id;
	}
	return 0;
}

/* seeded random DP inputs; mirrored in tests/katgen.py */
static void gen_pair(int qlen, int tlen, uint8_t *q, uint8_t *t, int related)
{
	int i, j;
	for (i = 0; i < tlen; ++i) t[i] = (sm_next() % 100) < 2 ? 4 : (uint8_t)(sm_next() & 3);
	if (!related) { for (i = 0; i < qlen; ++i) q[i] = (uint8_t)(sm_next() & 3); return; }
	/* query = mutated copy of target: 6% sub, 2% ins, 2% del; padded with random */
	for (i = j = 0; i < qlen; ) {
		int r = (int)(sm_next() % 100);
		if (j >= tlen) { q[i++] = (uint8_t)(sm_next() & 3); continue; }
		if (r < 6) { q[i++] = (uint8_t)(sm_next() & 3); ++j; }
		else if (r < 8) { q[i++] = (uint8_t)(sm_next() & 3); }
		else if (r < 10) { ++j; }
		else { q[i++] = t[j++]; }
	}
}

static int do_extend(int argc, char **argv)
{
	int it, n = atoi(argv[2]);
	sm_state = strtoull(argv[3], 0, 10);
	for (it = 0; it < n; ++it) {
		int8_t mat[25];
		int a = sm_range(1, 2), b = sm_range(1, 6), o_del = sm_range(1, 8), e_del = sm_range(1, 3), o_ins = sm_range(1, 8), e_ins = sm_range(1, 3);
		int qlen = sm_range(1, (it % 8 == 0) ? 400 : 160), tlen = qlen + sm_range(0, 120), w = sm_range(1, 120);
		int end_bonus = sm_range(0, 7), zdrop = (sm_next() & 1) ? sm_range(10, 120) : 0, h0 = sm_range(1, 300);
		int related = (int)(sm_next() % 4) != 0, qle, tle, gtle, gscore, max_off, sc, i;
		uint8_t *q = malloc(qlen), *t = malloc(tlen);
		bwa_fill_scmat(a, b, mat);
		gen_pair(qlen, tlen, q, t, related);
		sc = ksw_extend2(qlen, q, tlen, t, 5, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, &qle, &tle, &gtle, &gscore, &max_off);
		printf("E %d %d %d %d %d %d %d %d %d %d %d %d ", a, b, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0, qlen, tlen);
		for (i = 0; i < qlen; ++i) putchar("ACGTN"[q[i]]);
		putchar(' ');
		for (i = 0; i < tlen; ++i) putchar("ACGTN"[t[i]]);
		printf(" -> %d %d %d %d %d %d\n", sc, qle, tle, gtle, gscore, max_off);
		free(q); free(t);
	}
	return 0;
}

static int do_global(int argc, char **argv)
{
	int it, n = atoi(argv[2]);
	sm_state = strtoull(argv[3], 0, 10);
	for (it = 0; it < n; ++it) {
		int8_t mat[25];
		int a = sm_range(1, 2), b = sm_range(1, 6), o_del = sm_range(1, 8), e_del = sm_range(1, 3), o_ins = sm_range(1, 8), e_ins = sm_range(1, 3);
		int qlen = sm_range(1, (it % 8 == 0) ? 400 : 160), d = sm_range(-20, 20), tlen = qlen + d > 0 ? qlen + d : 1;
		int w = abs(tlen - qlen) + sm_range(1, 60), n_cigar, sc, i;
		uint32_t *cigar = 0;
		uint8_t *q = malloc(qlen), *t = malloc(tlen);
		bwa_fill_scmat(a, b, mat);
		gen_pair(qlen, tlen, q, t, (int)(sm_next() % 8) != 0);
		sc = ksw_global2(qlen, q, tlen, t, 5, mat, o_del, e_del, o_ins, e_ins, w, &n_cigar, &cigar);
		printf("B %d %d %d %d %d %d %d %d %d ", a, b, o_del, e_del, o_ins, e_ins, w, qlen, tlen);
		for (i = 0; i < qlen; ++i) putchar("ACGTN"[q[i]]);
		putchar(' ');
		for (i = 0; i < tlen; ++i) putchar("ACGTN"[t[i]]);
		printf(" -> %d ", sc);
		for (i = 0; i < n_cigar; ++i) printf("%d%c", cigar[i] >> 4, "MIDSH"[cigar[i] & 0xf]);
		printf("\n");
		free(cigar); free(q); free(t);
	}
	return 0;
}

static int do_local(int argc, char **argv)
{
	int it, n = atoi(argv[2]);
	sm_state = strtoull(argv[3], 0, 10);
	for (it = 0; it < n; ++it) {
		int8_t mat[25];
		int a = 1, b = sm_range(1, 5), o_del = sm_range(1, 7), e_del = sm_range(1, 2), o_ins = sm_range(1, 7), e_ins = sm_range(1, 2);
		int qlen = sm_range(20, 251), tlen = sm_range(qlen, qlen + 600), i, off;
		int xtra = KSW_XSUBO | KSW_XSTART | (qlen * a < 250 ? KSW_XBYTE : 0) | (19 * a);
		uint8_t *q = malloc(qlen), *t = malloc(tlen), *tmp = malloc(qlen + 64);
		kswr_t r;
		bwa_fill_scmat(a, b, mat);
		for (i = 0; i < tlen; ++i) t[i] = (sm_next() % 100) < 1 ? 4 : (uint8_t)(sm_next() & 3);
		off = sm_range(0, tlen - qlen);
		gen_pair(qlen, qlen + 32 < tlen - off ? qlen + 32 : tlen - off, q, t + off, (int)(sm_next() % 6) != 0);
		r = ksw_align2(qlen, q, tlen, t, 5, mat, o_del, e_del, o_ins, e_ins, xtra, 0);
		printf("W %d %d %d %d %d %d %d %d %d ", a, b, o_del, e_del, o_ins, e_ins, xtra, qlen, tlen);
		for (i = 0; i < qlen; ++i) putchar("ACGTN"[q[i]]);
		putchar(' ');
		for (i = 0; i < tlen; ++i) putchar("ACGTN"[t[i]]);
		printf(" -> %d %d %d %d %d %d %d\n", r.score, r.te, r.qe, r.score2, r.te2, r.tb, r.qb);
		free(q); free(t); free(tmp);
	}
	return 0;
}

int main(int argc, char **argv)
{
	bwa_verbose = 1;
	if (argc < 4) { fprintf(stderr, "usage: katdump <smem|sa|chain|regs|aln|extend|global|local> ...\n"); return 1; }
	if (!strcmp(argv[1], "smem")) return do_smem(argc, argv);
	if (!strcmp(argv[1], "sa")) return do_sa(argc, argv);
	if (!strcmp(argv[1], "chain")) return do_chain(argc, argv);
	if (!strcmp(argv[1], "regs")) return do_regs(argc, argv, 0);
	if (!strcmp(argv[1], "aln")) return do_regs(argc, argv, 1);
	if (!strcmp(argv[1], "extend")) return do_extend(argc, argv);
	if (!strcmp(argv[1], "global")) return do_global(argc, argv);
	if (!strcmp(argv[1], "local")) return do_local(argc, argv);
	fprintf(stderr, "unknown mode %s\n", argv[1]);
	return 1;
}
