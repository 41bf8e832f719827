/* bb_pair.c -- paired-end logic on the host: insert-size model per batch, mate rescue, pairing of
 * single-end hits, pair MAPQ and the paired SAM records (reference: bwamem_pair.c:48-419).
 *
 * All arithmetic that feeds an integer decision keeps the reference's operand types (double
 * products truncated with +.499, erfc/log from libm), see SURVEY.md section 7 item 8.
 */
#include <math.h>
#include <assert.h>
#include "bb_host.h"

#define MIN_RATIO     0.8
#define MIN_DIR_CNT   10
#define MIN_DIR_RATIO 0.05
#define OUTLIER_BOUND 2.0
#define MAPPING_BOUND 3.0
#define MAX_STDDEV    4.0

/* orientation class (0 FF, 1 FR, 2 RF, 3 RR) and distance of two hits given their doubled-coordinate starts */
static inline int infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist)
{
	int r1 = b1 >= l_pac, r2 = b2 >= l_pac;
	int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
	*dist = p2 > b1 ? p2 - b1 : b1 - p2;
	return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

static int best_overlapping_sub(const mem_opt_t *opt, const mem_alnreg_v *r) /* bwamem_pair.c:58-70 */
{
	size_t j;
	for (j = 1; j < r->n; ++j) {
		int b_max = r->a[j].qb > r->a[0].qb ? r->a[j].qb : r->a[0].qb;
		int e_min = r->a[j].qe < r->a[0].qe ? r->a[j].qe : r->a[0].qe;
		if (e_min > b_max) {
			int lj = r->a[j].qe - r->a[j].qb, l0 = r->a[0].qe - r->a[0].qb;
			int min_l = lj < l0 ? lj : l0;
			if (e_min - b_max >= min_l * opt->mask_level) break;
		}
	}
	return j < r->n ? r->a[j].score : opt->min_seed_len * opt->a;
}

/* one pair's contribution to the insert-size model (the loop body of bwamem_pair.c:88-101): 0 if the pair is
 * not a confident unique one, else (orientation+1) << 48 | distance.  Independent per pair, so the batch
 * driver evaluates it in parallel while the chunks are still in flight. */
uint64_t bb_pestat_pair(const mem_opt_t *opt, int64_t l_pac, const mem_alnreg_v *r0, const mem_alnreg_v *r1)
{
	int64_t is;
	int dir;
	if (r0->n == 0 || r1->n == 0) return 0;
	if (best_overlapping_sub(opt, r0) > MIN_RATIO * r0->a[0].score) return 0;
	if (best_overlapping_sub(opt, r1) > MIN_RATIO * r1->a[0].score) return 0;
	if (r0->a[0].rid != r1->a[0].rid) return 0;
	dir = infer_dir(l_pac, r0->a[0].rb, r1->a[0].rb, &is);
	return is && is <= opt->max_ins ? (uint64_t)(dir + 1) << 48 | (uint64_t)is : 0;
}

/* ascending order of insert sizes; they are bounded by max_ins, so a counting pass replaces the comparison sort */
static void sort_isizes(size_t n, uint64_t *q, int max_ins)
{
	size_t k, o = 0;
	uint32_t *cnt;
	int v;
	if (n < 4096 || max_ins > 1 << 24) { bb_sort_u64(n, q); return; }
	cnt = bb_calloc((size_t)max_ins + 1, sizeof(uint32_t));
	for (k = 0; k < n; ++k) ++cnt[q[k]];
	for (v = 0; v <= max_ins; ++v) { uint32_t c = cnt[v]; while (c--) q[o++] = (uint64_t)v; }
	free(cnt);
}

void mem_pestat(const mem_opt_t *opt, int64_t l_pac, int n, const mem_alnreg_v *regs, mem_pestat_t pes[4])
{
	uint64_t *v = bb_malloc(sizeof(uint64_t) * ((size_t)(n >> 1) + 1));
	int i;
	for (i = 0; i < n >> 1; ++i) v[i] = bb_pestat_pair(opt, l_pac, &regs[i << 1], &regs[i << 1 | 1]);
	bb_pestat_from_pairs(opt, n >> 1, v, pes);
	free(v);
}

void bb_pestat_from_pairs(const mem_opt_t *opt, long n_pairs, const uint64_t *v, mem_pestat_t pes[4])
{
	BB_VEC(uint64_t) isize[4];
	long i;
	int d;
	size_t max;
	memset(pes, 0, 4 * sizeof(mem_pestat_t));
	memset(isize, 0, sizeof(isize));
	for (i = 0; i < n_pairs; ++i)
		if (v[i]) bb_vec_push(isize[(v[i] >> 48) - 1], v[i] & 0xffffffffffffULL);
	if (bwa_verbose >= 3) fprintf(stderr, "[M::%s] # candidate unique pairs for (FF, FR, RF, RR): (%ld, %ld, %ld, %ld)\n", "mem_pestat", (long)isize[0].n, (long)isize[1].n, (long)isize[2].n, (long)isize[3].n);
	for (d = 0; d < 4; ++d) {
		mem_pestat_t *r = &pes[d];
		uint64_t *q = isize[d].a;
		size_t qn = isize[d].n, k;
		int p25, p50, p75, x;
		if (qn < MIN_DIR_CNT) {
			fprintf(stderr, "[M::%s] skip orientation %c%c as there are not enough pairs\n", "mem_pestat", "FR"[d >> 1 & 1], "FR"[d & 1]);
			r->failed = 1;
			continue;
		} else fprintf(stderr, "[M::%s] analyzing insert size distribution for orientation %c%c...\n", "mem_pestat", "FR"[d >> 1 & 1], "FR"[d & 1]);
		sort_isizes(qn, q, opt->max_ins);
		p25 = (int)q[(int)(.25 * qn + .499)];
		p50 = (int)q[(int)(.50 * qn + .499)];
		p75 = (int)q[(int)(.75 * qn + .499)];
		r->low = (int)(p25 - OUTLIER_BOUND * (p75 - p25) + .499);
		if (r->low < 1) r->low = 1;
		r->high = (int)(p75 + OUTLIER_BOUND * (p75 - p25) + .499);
		fprintf(stderr, "[M::%s] (25, 50, 75) percentile: (%d, %d, %d)\n", "mem_pestat", p25, p50, p75);
		fprintf(stderr, "[M::%s] low and high boundaries for computing mean and std.dev: (%d, %d)\n", "mem_pestat", r->low, r->high);
		for (k = 0, x = 0, r->avg = 0; k < qn; ++k)
			if (q[k] >= (uint64_t)r->low && q[k] <= (uint64_t)r->high) { r->avg += q[k]; ++x; }
		r->avg /= x;
		for (k = 0, r->std = 0; k < qn; ++k)
			if (q[k] >= (uint64_t)r->low && q[k] <= (uint64_t)r->high) r->std += (q[k] - r->avg) * (q[k] - r->avg);
		r->std = sqrt(r->std / x);
		fprintf(stderr, "[M::%s] mean and std.dev: (%.2f, %.2f)\n", "mem_pestat", r->avg, r->std);
		r->low = (int)(p25 - MAPPING_BOUND * (p75 - p25) + .499);
		r->high = (int)(p75 + MAPPING_BOUND * (p75 - p25) + .499);
		if (r->low > r->avg - MAX_STDDEV * r->std) r->low = (int)(r->avg - MAX_STDDEV * r->std + .499);
		if (r->high < r->avg + MAX_STDDEV * r->std) r->high = (int)(r->avg + MAX_STDDEV * r->std + .499);
		if (r->low < 1) r->low = 1;
		fprintf(stderr, "[M::%s] low and high boundaries for proper pairs: (%d, %d)\n", "mem_pestat", r->low, r->high);
	}
	for (d = 0, max = 0; d < 4; ++d) if (isize[d].n > max) max = isize[d].n;
	for (d = 0; d < 4; ++d)
		if (pes[d].failed == 0 && isize[d].n < max * MIN_DIR_RATIO) {
			pes[d].failed = 1;
			fprintf(stderr, "[M::%s] skip orientation %c%c\n", "mem_pestat", "FR"[d >> 1 & 1], "FR"[d & 1]);
		}
	for (d = 0; d < 4; ++d) free(isize[d].a);
}

/* mate rescue for one anchor region (bwamem_pair.c:137-206): local SW of the mate inside the window the
 * insert-size model predicts; hits are inserted into ma (kept sorted by score) and de-duplicated */
static const bb_swr_t *swcache_get(bb_swcache_t *c, int which, int is_rev, int64_t rb, int64_t re)
{
	size_t k;
	bb_swent_t e;
	for (k = 0; k < c->v.n; ++k) {
		const bb_swent_t *x = &c->v.a[k];
		if (x->which == which && x->is_rev == is_rev && x->rb == rb && x->re == re) {
			if (x->done) return &x->res;
			++c->pending;
			return 0;
		}
	}
	memset(&e, 0, sizeof(e));
	e.which = which; e.is_rev = is_rev; e.rb = rb; e.re = re;
	bb_vec_push(c->v, e);
	++c->pending;
	return 0;
}

int bb_matesw(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const mem_pestat_t pes[4], const mem_alnreg_t *a, int l_ms, const uint8_t *ms, mem_alnreg_v *ma, bb_swcache_t *swc, int which)
{
	int64_t l_pac = bns->l_pac;
	int i, r, skip[4], n = 0, rid = -1;
	for (r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
	for (i = 0; i < (int)ma->n; ++i) {
		int64_t dist;
		r = infer_dir(l_pac, a->rb, ma->a[i].rb, &dist);
		if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
	}
	if (skip[0] + skip[1] + skip[2] + skip[3] == 4) return 0;
	for (r = 0; r < 4; ++r) {
		int is_rev, is_larger;
		uint8_t *seq, *rev = 0, *ref = 0;
		int64_t rb, re;
		if (skip[r]) continue;
		is_rev = (r >> 1 != (r & 1));
		is_larger = !(r >> 1);
		if (is_rev && !swc) {
			rev = bb_malloc(l_ms);
			for (i = 0; i < l_ms; ++i) rev[l_ms - 1 - i] = ms[i] < 4 ? 3 - ms[i] : 4;
			seq = rev;
		} else seq = (uint8_t *)ms;
		if (!is_rev) {
			rb = is_larger ? a->rb + pes[r].low : a->rb - pes[r].high;
			re = (is_larger ? a->rb + pes[r].high : a->rb - pes[r].low) + l_ms;
		} else {
			rb = (is_larger ? a->rb + pes[r].low : a->rb - pes[r].high) - l_ms;
			re = is_larger ? a->rb + pes[r].high : a->rb - pes[r].low;
		}
		if (rb < 0) rb = 0;
		if (re > l_pac << 1) re = l_pac << 1;
		if (rb < re) { if (swc) bb_clamp_to_contig(bns, &rb, (rb + re) >> 1, &re, &rid); else ref = bb_fetch_seq(bns, pac, &rb, (rb + re) >> 1, &re, &rid); }
		if (a->rid == rid && re - rb >= opt->min_seed_len) {
			bb_swr_t aln;
			mem_alnreg_t b;
			int tmp, xtra = BB_SW_XSUBO | BB_SW_XSTART | (l_ms * opt->a < 250 ? BB_SW_XBYTE : 0) | (opt->min_seed_len * opt->a);
			if (swc) {   /* the alignment comes from the device: (query read, strand, window) identifies it */
				const bb_swr_t *got = swcache_get(swc, which, is_rev, rb, re);
				if (!got) { if (swc->probe) continue; return -1; }
				aln = *got;
			} else aln = bb_local_sw(l_ms, seq, (int)(re - rb), ref, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, xtra);
			memset(&b, 0, sizeof(b));
			if (aln.score >= opt->min_seed_len && aln.qb >= 0) {
				b.rid = a->rid;
				b.is_alt = a->is_alt;
				b.qb = is_rev ? l_ms - (aln.qe + 1) : aln.qb;
				b.qe = is_rev ? l_ms - aln.qb : aln.qe + 1;
				b.rb = is_rev ? (l_pac << 1) - (rb + aln.te + 1) : rb + aln.tb;
				b.re = is_rev ? (l_pac << 1) - (rb + aln.tb) : rb + aln.te + 1;
				b.score = aln.score;
				b.csub = aln.score2;
				b.secondary = -1;
				b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
				bb_regs_make_room(ma);
				ma->a[ma->n++] = b;
				for (i = 0; i < (int)ma->n - 1; ++i)
					if (ma->a[i].score < b.score) break;
				tmp = i;
				for (i = (int)ma->n - 1; i > tmp; --i) ma->a[i] = ma->a[i - 1];
				ma->a[i] = b;
			}
			++n;
		}
		if (n) ma->n = bb_sort_dedup_patch(opt, 0, 0, 0, (int)ma->n, ma->a);
		free(rev); free(ref);
	}
	return n;
}

/* the rescue block at the top of mem_sam_pe (bwamem_pair.c:289-301); modifies a[0], a[1] */
int bb_rescue_pe(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const mem_pestat_t pes[4], bseq1_t s[2], mem_alnreg_v a[2], bb_swcache_t *swc)
{
	int i, n = 0;
	size_t j;
	mem_alnreg_v b[2];
	mem_alnreg_t st[2][4];
	if (opt->flag & MEM_F_NO_RESCUE) return 0;
	memset(b, 0, sizeof(b));
	for (i = 0; i < 2; ++i) {
		if (a[i].n <= 4) { b[i].a = st[i]; b[i].m = 4; }   /* anchors: copies, because a[] changes while we rescue */
		for (j = 0; j < a[i].n; ++j)
			if (a[i].a[j].score >= a[i].a[0].score - opt->pen_unpaired) bb_vec_push(b[i], a[i].a[j]);
	}
	/* First attempt with device alignments: ask for the alignment of EVERY anchor and orientation that the hits present now do not
	 * rule out.  Rescued hits can only rule out more (bwamem_pair.c:143-147), so this is a superset of what the pass will use, and
	 * one device round serves the pair instead of one round per alignment. */
	if (swc) swc->probe = swc->v.n == 0;
	for (i = 0; i < 2 && n >= 0; ++i)
		for (j = 0; j < b[i].n && (int)j < opt->max_matesw; ++j) {
			int k = bb_matesw(opt, bns, pac, pes, &b[i].a[j], s[!i].l_seq, (uint8_t *)s[!i].seq, &a[!i], swc, !i);
			if (k < 0) { n = -1; break; }   /* an alignment was requested from the device: this pass is void */
			n += k;
		}
	if (b[0].a != st[0]) free(b[0].a);
	if (b[1].a != st[1]) free(b[1].a);
	if (swc && swc->probe) { swc->probe = 0; if (swc->pending > 0) return -1; }   /* requests collected: the pass is void */
	return n;
}

/* best proper pair among the primary-assembly hits of both ends (bwamem_pair.c:208-269) */
/* The insert-size term of a pair's score, .721 * log(2 * erfc(|dist - avg| / std / sqrt 2)) * a (bwamem_pair.c:266), depends
 * on the distance only and distances are integers in [low, high]: each thread tabulates the term once per insert-size
 * model (same expression, same doubles) instead of calling erfc and log for every candidate pair. */
typedef struct { double avg, std; int low, high, a, cap, valid; double *t; } pair_memo_t;
static __thread pair_memo_t tl_pair_memo[4];
static inline double pair_term(const mem_opt_t *opt, const mem_pestat_t *pe, int dir, int64_t dist)
{
	pair_memo_t *m = &tl_pair_memo[dir];
	if (!m->valid || m->avg != pe->avg || m->std != pe->std || m->low != pe->low || m->high != pe->high || m->a != opt->a) {
		int64_t n = (int64_t)pe->high - pe->low + 1, k;
		if (n <= 0 || n > 65536) { double ns = (dist - pe->avg) / pe->std; return .721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * opt->a; }
		if (n > m->cap) { m->t = bb_realloc(m->t, (size_t)n * sizeof(double)); m->cap = (int)n; }
		for (k = 0; k < n; ++k) { double ns = ((pe->low + k) - pe->avg) / pe->std; m->t[k] = .721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * opt->a; }
		m->avg = pe->avg; m->std = pe->std; m->low = pe->low; m->high = pe->high; m->a = opt->a; m->valid = 1;
	}
	return m->t[dist - pe->low];
}

static int pair_hits(const mem_opt_t *opt, const bntseq_t *bns, const mem_pestat_t pes[4], mem_alnreg_v a[2], int id, int *sub, int *n_sub, int z[2], int n_pri[2])
{
	BB_VEC(bb_pair64_t) v = {0, 0, 0}, u = {0, 0, 0};
	bb_pair64_t vst[16], ust[32];
	int r, i, k, y[4], ret;
	int64_t l_pac = bns->l_pac;
	if (n_pri[0] + n_pri[1] <= 16) { v.a = vst; v.m = 16; }
	for (r = 0; r < 2; ++r)
		for (i = 0; i < n_pri[r]; ++i) {
			bb_pair64_t key;
			mem_alnreg_t *e = &a[r].a[i];
			key.x = e->rb < l_pac ? e->rb : (l_pac << 1) - 1 - e->rb;
			key.x = (uint64_t)e->rid << 32 | (key.x - bns->anns[e->rid].offset);
			key.y = (uint64_t)e->score << 32 | i << 2 | (e->rb >= l_pac) << 1 | r;
			bb_vec_push(v, key);
		}
	bb_sort_pair64(v.n, v.a);
	y[0] = y[1] = y[2] = y[3] = -1;
	for (i = 0; i < (int)v.n; ++i) {
		for (r = 0; r < 2; ++r) {
			int dir = r << 1 | (v.a[i].y >> 1 & 1), which;
			if (pes[dir].failed) continue;
			which = r << 1 | ((v.a[i].y & 1) ^ 1);
			if (y[which] < 0) continue;
			for (k = y[which]; k >= 0; --k) {
				int64_t dist;
				int q;
				bb_pair64_t p;
				if ((v.a[k].y & 3) != (uint64_t)which) continue;
				dist = (int64_t)v.a[i].x - v.a[k].x;
				if (dist > pes[dir].high) break;
				if (dist < pes[dir].low) continue;
				q = (int)((v.a[i].y >> 32) + (v.a[k].y >> 32) + pair_term(opt, &pes[dir], dir, dist) + .499);
				if (q < 0) q = 0;
				p.y = (uint64_t)k << 32 | i;
				p.x = (uint64_t)q << 32 | (bb_mix64(p.y ^ id << 8) & 0xffffffffU);
				if (u.a == 0 && u.n == 0) { u.a = ust; u.m = 32; }
				if (u.a == ust && u.n == 32) { bb_pair64_t *h_ = bb_malloc(64 * sizeof(bb_pair64_t)); memcpy(h_, ust, sizeof(ust)); u.a = h_; u.m = 64; }
				bb_vec_push(u, p);
			}
		}
		y[v.a[i].y & 3] = i;
	}
	if (u.n) {
		int tmp = opt->a + opt->b;
		if (opt->o_del + opt->e_del > tmp) tmp = opt->o_del + opt->e_del;
		if (opt->o_ins + opt->e_ins > tmp) tmp = opt->o_ins + opt->e_ins;
		bb_sort_pair64(u.n, u.a);
		i = (int)(u.a[u.n - 1].y >> 32); k = (int)(u.a[u.n - 1].y << 32 >> 32);
		z[v.a[i].y & 1] = (int)(v.a[i].y << 32 >> 34);
		z[v.a[k].y & 1] = (int)(v.a[k].y << 32 >> 34);
		ret = (int)(u.a[u.n - 1].x >> 32);
		*sub = u.n > 1 ? (int)(u.a[u.n - 2].x >> 32) : 0;
		for (i = (int)u.n - 2, *n_sub = 0; i >= 0; --i)
			if (*sub - (int)(u.a[i].x >> 32) <= tmp) ++*n_sub;
	} else { ret = 0; *sub = 0; *n_sub = 0; }
	if (u.a != ust) free(u.a);
	if (v.a != vst) free(v.a);
	return ret;
}

#define RAW_MAPQ(diff, a) ((int)(6.02 * (diff) / (a) + .499))

/* everything of mem_sam_pe after the rescue block (bwamem_pair.c:302-419).  sc[i] carries the
 * alignment cache of read i; when sc[0].dry is set no text is produced. */
int bb_sam_pe(bb_samctx_t sc[2], const mem_pestat_t pes[4], uint64_t id, bseq1_t s[2], mem_alnreg_v a[2], int rescue_done)
{
	const mem_opt_t *opt = sc[0].opt;
	const bntseq_t *bns = sc[0].bns;
	int i, j, z[2], o, subo, n_sub, extra_flag = 1, n_pri[2], n_aa[2], dry = sc[0].dry;
	bb_str_t str = {0, 0, 0};
	mem_aln_t h[2], g[2], aa[2][2];
	(void)rescue_done;
	memset(h, 0, sizeof(h)); memset(g, 0, sizeof(g));
	n_aa[0] = n_aa[1] = 0;
	n_pri[0] = bb_mark_primary_se(opt, (int)a[0].n, a[0].a, id << 1 | 0);
	n_pri[1] = bb_mark_primary_se(opt, (int)a[1].n, a[1].a, id << 1 | 1);
	if (opt->flag & MEM_F_PRIMARY5) { bb_reorder_primary5(opt->T, &a[0]); bb_reorder_primary5(opt->T, &a[1]); }
	if (opt->flag & MEM_F_NOPAIRING) goto no_pairing;
	if (n_pri[0] && n_pri[1] && (o = pair_hits(opt, bns, pes, a, (int)id, &subo, &n_sub, z, n_pri)) > 0) {
		int is_multi[2], q_pe, score_un, q_se[2];
		char **XA[2];
		for (i = 0; i < 2; ++i) {
			for (j = 1; j < n_pri[i]; ++j)
				if (a[i].a[j].secondary < 0 && a[i].a[j].score >= opt->T) break;
			is_multi[i] = j < n_pri[i] ? 1 : 0;
		}
		if (is_multi[0] || is_multi[1]) goto no_pairing;
		score_un = a[0].a[0].score + a[1].a[0].score - opt->pen_unpaired;
		subo = subo > score_un ? subo : score_un;
		q_pe = RAW_MAPQ(o - subo, opt->a);
		if (n_sub > 0) q_pe -= (int)(4.343 * log(n_sub + 1) + .499);
		if (q_pe < 0) q_pe = 0;
		if (q_pe > 60) q_pe = 60;
		q_pe = (int)(q_pe * (1. - .5 * (a[0].a[0].frac_rep + a[1].a[0].frac_rep)) + .499);
		if (o > score_un) {
			mem_alnreg_t *c[2];
			c[0] = &a[0].a[z[0]]; c[1] = &a[1].a[z[1]];
			for (i = 0; i < 2; ++i) {
				if (c[i]->secondary >= 0) { c[i]->sub = a[i].a[c[i]->secondary].score; c[i]->secondary = -2; }
				q_se[i] = bb_approx_mapq_se(opt, c[i]);
			}
			q_se[0] = q_se[0] > q_pe ? q_se[0] : q_pe < q_se[0] + 40 ? q_pe : q_se[0] + 40;
			q_se[1] = q_se[1] > q_pe ? q_se[1] : q_pe < q_se[1] + 40 ? q_pe : q_se[1] + 40;
			extra_flag |= 2;
			q_se[0] = q_se[0] < RAW_MAPQ(c[0]->score - c[0]->csub, opt->a) ? q_se[0] : RAW_MAPQ(c[0]->score - c[0]->csub, opt->a);
			q_se[1] = q_se[1] < RAW_MAPQ(c[1]->score - c[1]->csub, opt->a) ? q_se[1] : RAW_MAPQ(c[1]->score - c[1]->csub, opt->a);
		} else {
			z[0] = z[1] = 0;
			q_se[0] = bb_approx_mapq_se(opt, &a[0].a[0]);
			q_se[1] = bb_approx_mapq_se(opt, &a[1].a[0]);
		}
		for (i = 0; i < 2; ++i) {
			int k = a[i].a[z[i]].secondary_all;
			if (k >= 0 && k < n_pri[i]) {
				assert(a[i].a[k].secondary_all < 0);
				for (j = 0; j < (int)a[i].n; ++j)
					if (a[i].a[j].secondary_all == k || j == k) a[i].a[j].secondary_all = z[i];
				a[i].a[z[i]].secondary_all = -1;
			}
		}
		if (!(opt->flag & MEM_F_ALL)) {
			for (i = 0; i < 2; ++i) XA[i] = bb_gen_alt(&sc[i], &a[i], s[i].l_seq, s[i].seq);
		} else XA[0] = XA[1] = 0;
		for (i = 0; i < 2; ++i) {
			h[i] = bb_reg2aln(&sc[i], s[i].l_seq, s[i].seq, &a[i].a[z[i]]);
			h[i].mapq = q_se[i];
			h[i].flag |= 0x40 << i | extra_flag;
			h[i].XA = XA[i] ? XA[i][z[i]] : 0;
			aa[i][n_aa[i]++] = h[i];
			if (n_pri[i] < (int)a[i].n) {
				mem_alnreg_t *p = &a[i].a[n_pri[i]];
				if (p->score < opt->T || p->secondary >= 0 || !p->is_alt) continue;
				g[i] = bb_reg2aln(&sc[i], s[i].l_seq, s[i].seq, p);
				g[i].flag |= 0x800 | 0x40 << i | extra_flag;
				g[i].XA = XA[i] ? XA[i][n_pri[i]] : 0;
				aa[i][n_aa[i]++] = g[i];
			}
		}
		if (!dry) {
			/* one buffer per read, sized for the usual single record up front instead of a chain of doublings and a copy */
			bb_str_need(&str, (size_t)s[0].l_seq * 2 + strlen(s[0].name) + 448);   /* covers what bb_aln2sam reserves for one ordinary record: no second allocation */
			for (i = 0; i < n_aa[0]; ++i) bb_aln2sam(opt, bns, &str, &s[0], n_aa[0], aa[0], i, &h[1]);
			s[0].sam = str.s;
			str.s = 0; str.l = str.m = 0;
			bb_str_need(&str, (size_t)s[1].l_seq * 2 + strlen(s[1].name) + 448);
			for (i = 0; i < n_aa[1]; ++i) bb_aln2sam(opt, bns, &str, &s[1], n_aa[1], aa[1], i, &h[0]);
			s[1].sam = str.s;
			if (strcmp(s[0].name, s[1].name) != 0) bb_fatal("mem_sam_pe", "paired reads have different names: \"%s\", \"%s\"\n", s[0].name, s[1].name);
		}
		for (i = 0; i < 2; ++i) {
			bb_cigar_free(&sc[i], h[i].cigar); bb_cigar_free(&sc[i], g[i].cigar);
			if (XA[i] == 0) continue;
			for (j = 0; j < (int)a[i].n; ++j) free(XA[i][j]);
			free(XA[i]);
		}
	} else goto no_pairing;
	return 0;

no_pairing:
	for (i = 0; i < 2; ++i) {
		int which = -1;
		if (a[i].n) {
			if (a[i].a[0].score >= opt->T) which = 0;
			else if (n_pri[i] < (int)a[i].n && a[i].a[n_pri[i]].score >= opt->T) which = n_pri[i];
		}
		if (which >= 0) h[i] = bb_reg2aln(&sc[i], s[i].l_seq, s[i].seq, &a[i].a[which]);
		else h[i] = bb_reg2aln(&sc[i], s[i].l_seq, s[i].seq, 0);
	}
	if (!(opt->flag & MEM_F_NOPAIRING) && h[0].rid == h[1].rid && h[0].rid >= 0) {
		int64_t dist;
		int d = infer_dir(bns->l_pac, a[0].a[0].rb, a[1].a[0].rb, &dist);
		if (!pes[d].failed && dist >= pes[d].low && dist <= pes[d].high) extra_flag |= 2;
	}
	bb_reg2sam(&sc[0], &s[0], &a[0], 0x41 | extra_flag, &h[1]);
	bb_reg2sam(&sc[1], &s[1], &a[1], 0x81 | extra_flag, &h[0]);
	if (!dry && strcmp(s[0].name, s[1].name) != 0) bb_fatal("mem_sam_pe", "paired reads have different names: \"%s\", \"%s\"\n", s[0].name, s[1].name);
	bb_cigar_free(&sc[0], h[0].cigar); bb_cigar_free(&sc[1], h[1].cigar);
	return 0;
}
