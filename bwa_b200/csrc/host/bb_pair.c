/* bb_pair.c -- paired-end logic on the host, for the pairs the device tail (bwag_tail.cu) hands back and for the per-batch
 * insert-size model: the model itself, mate rescue, choice of the best proper pair, pair MAPQ and the paired SAM records.
 * Behaviour to match: bwamem_pair.c:48-419.
 *
 * What bit-exactness pins down: the operand types and evaluation order of every floating-point expression that is truncated to an
 * int or compared with one (marked "order matters"), libm's erfc/log/sqrt, the unstable introsorts, and the hash that breaks ties
 * between equally good pairs.  The decomposition into steps, the data carried between them and the names are this file's own; the
 * device formulation for the common small case is t_pair / t_matesw_would_align in bwag_tail.cu.
 */
#include <math.h>
#include <assert.h>
#include "bb_host.h"

/* ------------------------------------------------------------------------------------------------------ orientation classes
 * A pair falls in one of four classes by the strands of its two hits as seen from the first: 0 FF, 1 FR, 2 RF, 3 RR.  Positions are
 * in the doubled coordinate system (forward strand [0, l_pac), reverse strand [l_pac, 2 l_pac)). */
static const char ORI_NAME[4][3] = { "FF", "FR", "RF", "RR" };

static inline int orientation_of(int64_t l_pac, int64_t first, int64_t second, int64_t *dist)
{
	const int rev1 = first >= l_pac, rev2 = second >= l_pac;
	const int64_t proj = rev1 == rev2 ? second : (l_pac << 1) - 1 - second;   /* the second hit projected on the strand of the first */
	const int after = proj > first;
	*dist = after ? proj - first : first - proj;
	return (rev1 == rev2 ? 0 : 1) ^ (after ? 0 : 3);
}
static inline int within(const mem_pestat_t *m, int64_t dist) { return dist >= m->low && dist <= m->high; }

/* ------------------------------------------------------------------------------------------------------ insert-size model */
static const double UNIQUE_RATIO = 0.8;     /* a hit is unique if its best overlapping rival scores at most this share */
static const double FENCE_MOMENTS = 2.0;    /* interquartile fences for the values that enter mean and deviation ... */
static const double FENCE_PROPER = 3.0;     /* ... and for the proper-pair range */
static const double SIGMA_PROPER = 4.0;     /* the proper-pair range also reaches this many deviations from the mean */
static const double MIN_CLASS_SHARE = 0.05; /* a class with fewer pairs than this share of the largest class is dropped */
enum { MIN_CLASS_PAIRS = 10 };

/* score of the best hit that overlaps the top hit on the read by mask_level of the shorter one; a seed's worth if there is none */
static int rival_of_top(const mem_opt_t *opt, const mem_alnreg_v *v)
{
	const mem_alnreg_t *top = &v->a[0];
	size_t j;
	for (j = 1; j < v->n; ++j) {
		const mem_alnreg_t *r = &v->a[j];
		const int lo = r->qb > top->qb ? r->qb : top->qb, hi = r->qe < top->qe ? r->qe : top->qe;
		if (hi > lo) {
			const int lr = r->qe - r->qb, lt = top->qe - top->qb;
			if (hi - lo >= (lr < lt ? lr : lt) * opt->mask_level) return r->score;
		}
	}
	return opt->min_seed_len * opt->a;
}

/* One pair's vote for the model: 0 if either end is missing, ambiguous or on another contig, or the distance is out of range; else
 * (class + 1) << 48 | distance.  Independent per pair: the batch driver collects the votes while chunks are still in flight. */
uint64_t bb_pestat_pair(const mem_opt_t *opt, int64_t l_pac, const mem_alnreg_v *r0, const mem_alnreg_v *r1)
{
	int64_t dist;
	int cls;
	if (r0->n == 0 || r1->n == 0) return 0;
	if (rival_of_top(opt, r0) > UNIQUE_RATIO * r0->a[0].score) return 0;
	if (rival_of_top(opt, r1) > UNIQUE_RATIO * r1->a[0].score) return 0;
	if (r0->a[0].rid != r1->a[0].rid) return 0;
	cls = orientation_of(l_pac, r0->a[0].rb, r1->a[0].rb, &dist);
	if (dist == 0 || dist > opt->max_ins) return 0;
	return (uint64_t)(cls + 1) << 48 | (uint64_t)dist;
}

/* ascending order of the distances; they are bounded by max_ins, so a counting pass replaces the comparison sort */
static void sort_distances(size_t n, uint64_t *q, int max_ins)
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

static inline int quantile(const uint64_t *sorted, size_t n, double f) { return (int)sorted[(int)(f * n + .499)]; }   /* order matters */

/* the model of one class from its sorted distances (messages as the reference prints them: users grep for these) */
static void fit_class(const uint64_t *q, size_t n, mem_pestat_t *m)
{
	const int q1 = quantile(q, n, .25), q2 = quantile(q, n, .50), q3 = quantile(q, n, .75), iqr = q3 - q1;
	size_t k;
	int used = 0;
	m->low = (int)(q1 - FENCE_MOMENTS * iqr + .499);   /* order matters */
	if (m->low < 1) m->low = 1;
	m->high = (int)(q3 + FENCE_MOMENTS * iqr + .499);
	fprintf(stderr, "[M::%s] (25, 50, 75) percentile: (%d, %d, %d)\n", "mem_pestat", q1, q2, q3);
	fprintf(stderr, "[M::%s] low and high boundaries for computing mean and std.dev: (%d, %d)\n", "mem_pestat", m->low, m->high);
	m->avg = 0;
	for (k = 0; k < n; ++k)
		if (q[k] >= (uint64_t)m->low && q[k] <= (uint64_t)m->high) { m->avg += q[k]; ++used; }
	m->avg /= used;
	m->std = 0;
	for (k = 0; k < n; ++k)
		if (q[k] >= (uint64_t)m->low && q[k] <= (uint64_t)m->high) m->std += (q[k] - m->avg) * (q[k] - m->avg);
	m->std = sqrt(m->std / used);
	fprintf(stderr, "[M::%s] mean and std.dev: (%.2f, %.2f)\n", "mem_pestat", m->avg, m->std);
	m->low = (int)(q1 - FENCE_PROPER * iqr + .499);
	m->high = (int)(q3 + FENCE_PROPER * iqr + .499);
	if (m->low > m->avg - SIGMA_PROPER * m->std) m->low = (int)(m->avg - SIGMA_PROPER * m->std + .499);
	if (m->high < m->avg + SIGMA_PROPER * m->std) m->high = (int)(m->avg + SIGMA_PROPER * m->std + .499);
	if (m->low < 1) m->low = 1;
	fprintf(stderr, "[M::%s] low and high boundaries for proper pairs: (%d, %d)\n", "mem_pestat", m->low, m->high);
}

void bb_pestat_from_pairs(const mem_opt_t *opt, long n_pairs, const uint64_t *votes, mem_pestat_t pes[4])
{
	BB_VEC(uint64_t) by_class[4];
	size_t largest = 0;
	long i;
	int c;
	memset(pes, 0, 4 * sizeof(mem_pestat_t));
	memset(by_class, 0, sizeof(by_class));
	for (i = 0; i < n_pairs; ++i)
		if (votes[i]) bb_vec_push(by_class[(votes[i] >> 48) - 1], votes[i] & 0xffffffffffffULL);
	if (bwa_verbose >= 3) fprintf(stderr, "[M::%s] # candidate unique pairs for (FF, FR, RF, RR): (%ld, %ld, %ld, %ld)\n", "mem_pestat", (long)by_class[0].n, (long)by_class[1].n, (long)by_class[2].n, (long)by_class[3].n);
	for (c = 0; c < 4; ++c) {
		if (by_class[c].n < MIN_CLASS_PAIRS) {
			fprintf(stderr, "[M::%s] skip orientation %s as there are not enough pairs\n", "mem_pestat", ORI_NAME[c]);
			pes[c].failed = 1;
			continue;
		}
		fprintf(stderr, "[M::%s] analyzing insert size distribution for orientation %s...\n", "mem_pestat", ORI_NAME[c]);
		sort_distances(by_class[c].n, by_class[c].a, opt->max_ins);
		fit_class(by_class[c].a, by_class[c].n, &pes[c]);
	}
	for (c = 0; c < 4; ++c) if (by_class[c].n > largest) largest = by_class[c].n;
	for (c = 0; c < 4; ++c)
		if (!pes[c].failed && by_class[c].n < largest * MIN_CLASS_SHARE) {
			pes[c].failed = 1;
			fprintf(stderr, "[M::%s] skip orientation %s\n", "mem_pestat", ORI_NAME[c]);
		}
	for (c = 0; c < 4; ++c) free(by_class[c].a);
}

void mem_pestat(const mem_opt_t *opt, int64_t l_pac, int n, const mem_alnreg_v *regs, mem_pestat_t pes[4])
{
	const long n_pairs = n >> 1;
	uint64_t *votes = bb_malloc(sizeof(uint64_t) * ((size_t)n_pairs + 1));
	long i;
	for (i = 0; i < n_pairs; ++i) votes[i] = bb_pestat_pair(opt, l_pac, &regs[2 * i], &regs[2 * i + 1]);
	bb_pestat_from_pairs(opt, n_pairs, votes, pes);
	free(votes);
}

/* ------------------------------------------------------------------------------------------------------------ mate rescue
 * For an anchor hit of one read, the model predicts where the mate lies in each orientation class; a local alignment of the mate
 * inside that window may find a hit seeding missed (bwamem_pair.c:137-206).  With a cache (swc) the alignments come from the device
 * (K6): a miss records the request and the pass over the pair is replayed after the device has served the batch's requests. */
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


/* the window of class `cls` around the anchor, in doubled coordinates; the mate is aligned reverse-complemented when *mate_rc */
static void rescue_window(const mem_pestat_t *m, int cls, int64_t anchor, int l_mate, int64_t l_pac, int64_t *beg, int64_t *end, int *mate_rc)
{
	const int mate_after = !(cls >> 1);                 /* the mate lies at larger coordinates than the anchor */
	const int64_t near = mate_after ? anchor + m->low : anchor - m->high, far = mate_after ? anchor + m->high : anchor - m->low;
	*mate_rc = (cls >> 1) != (cls & 1);
	if (*mate_rc) { *beg = near - l_mate; *end = far; }
	else { *beg = near; *end = far + l_mate; }
	if (*beg < 0) *beg = 0;
	if (*end > l_pac << 1) *end = l_pac << 1;
}

/* the region a local alignment inside [win_beg, ...) stands for */
static mem_alnreg_t region_of(const bb_swr_t *sw, const mem_alnreg_t *anchor, int mate_rc, int l_mate, int64_t win_beg, int64_t l_pac)
{
	mem_alnreg_t r;
	int64_t ref_len, qry_len;
	memset(&r, 0, sizeof(r));
	r.rid = anchor->rid;
	r.is_alt = anchor->is_alt;
	if (mate_rc) {
		r.qb = l_mate - (sw->qe + 1); r.qe = l_mate - sw->qb;
		r.rb = (l_pac << 1) - (win_beg + sw->te + 1); r.re = (l_pac << 1) - (win_beg + sw->tb);
	} else {
		r.qb = sw->qb; r.qe = sw->qe + 1;
		r.rb = win_beg + sw->tb; r.re = win_beg + sw->te + 1;
	}
	r.score = sw->score;
	r.csub = sw->score2;
	r.secondary = -1;
	ref_len = r.re - r.rb; qry_len = r.qe - r.qb;
	r.seedcov = (int)((ref_len < qry_len ? ref_len : qry_len) >> 1);
	return r;
}

static void insert_by_score(mem_alnreg_v *v, const mem_alnreg_t *r)   /* before the first hit that scores less */
{
	size_t at = 0;
	bb_regs_make_room(v);
	while (at < v->n && v->a[at].score >= r->score) ++at;
	memmove(&v->a[at + 1], &v->a[at], (v->n - at) * sizeof(mem_alnreg_t));
	v->a[at] = *r;
	++v->n;
}

/* Returns the number of alignments made for this anchor (hits found are merged into `mate_hits`), or -1 when one had to be
 * requested from the device.  `which`: the read of the pair that is the mate (the cache key). */
int bb_matesw(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const mem_pestat_t pes[4], const mem_alnreg_t *anchor, int l_mate, const uint8_t *mate, mem_alnreg_v *mate_hits, bb_swcache_t *swc, int which)
{
	const int64_t l_pac = bns->l_pac;
	int cls, n_aligned = 0, closed[4], open = 0;
	size_t k;
	/* a class is closed if its model failed or a mate hit already sits where it predicts */
	for (cls = 0; cls < 4; ++cls) closed[cls] = pes[cls].failed ? 1 : 0;
	for (k = 0; k < mate_hits->n; ++k) {
		int64_t dist;
		cls = orientation_of(l_pac, anchor->rb, mate_hits->a[k].rb, &dist);
		if (within(&pes[cls], dist)) closed[cls] = 1;
	}
	for (cls = 0; cls < 4; ++cls) open += !closed[cls];
	if (!open) return 0;
	for (cls = 0; cls < 4; ++cls) {
		int64_t beg, end;
		int mate_rc, rid = -1;
		uint8_t *rc_copy = 0, *window = 0;
		if (closed[cls]) continue;
		rescue_window(&pes[cls], cls, anchor->rb, l_mate, l_pac, &beg, &end, &mate_rc);
		if (beg < end) {
			if (swc) bb_clamp_to_contig(bns, &beg, (beg + end) >> 1, &end, &rid);
			else window = bb_fetch_seq(bns, pac, &beg, (beg + end) >> 1, &end, &rid);
		}
		if (anchor->rid == rid && end - beg >= opt->min_seed_len) {
			const int xtra = BB_SW_XSUBO | BB_SW_XSTART | (l_mate * opt->a < 250 ? BB_SW_XBYTE : 0) | (opt->min_seed_len * opt->a);
			bb_swr_t sw;
			if (swc) {   /* from the device: (mate read, strand, window) identifies the alignment */
				const bb_swr_t *got = swcache_get(swc, which, mate_rc, beg, end);
				if (!got) { if (swc->probe) continue; return -1; }
				sw = *got;
			} else {
				const uint8_t *query = mate;
				if (mate_rc) {
					int x;
					rc_copy = bb_malloc(l_mate);
					for (x = 0; x < l_mate; ++x) rc_copy[l_mate - 1 - x] = mate[x] < 4 ? 3 - mate[x] : 4;
					query = rc_copy;
				}
				sw = bb_local_sw(l_mate, (uint8_t *)query, (int)(end - beg), window, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, xtra);
			}
			if (sw.score >= opt->min_seed_len && sw.qb >= 0) {
				const mem_alnreg_t hit = region_of(&sw, anchor, mate_rc, l_mate, beg, l_pac);
				insert_by_score(mate_hits, &hit);
			}
			++n_aligned;
		}
		if (n_aligned) mate_hits->n = bb_sort_dedup_patch(opt, 0, 0, 0, (int)mate_hits->n, mate_hits->a);
		free(rc_copy); free(window);
	}
	return n_aligned;
}

/* Rescue for both reads of a pair: every hit of a read within pen_unpaired of its best (at most max_matesw of them) anchors a search
 * for the other read.  Modifies hits[0], hits[1].  -1: alignments were requested from the device, call again once they are in. */
int bb_rescue_pe(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const mem_pestat_t pes[4], bseq1_t s[2], mem_alnreg_v hits[2], bb_swcache_t *swc)
{
	mem_alnreg_v anchors[2];
	mem_alnreg_t few[2][4];
	int r, total = 0;
	size_t j;
	if (opt->flag & MEM_F_NO_RESCUE) return 0;
	memset(anchors, 0, sizeof(anchors));
	for (r = 0; r < 2; ++r) {   /* copies: the hit lists change while we rescue */
		if (hits[r].n <= 4) { anchors[r].a = few[r]; anchors[r].m = 4; }
		for (j = 0; j < hits[r].n; ++j)
			if (hits[r].a[j].score >= hits[r].a[0].score - opt->pen_unpaired) bb_vec_push(anchors[r], hits[r].a[j]);
	}
	/* First attempt with device alignments: ask for the alignment of EVERY anchor and class that the hits present now do not
	 * close.  Rescued hits can only close more, so this is a superset of what the pass will use, and one device round serves the
	 * pair instead of one round per alignment. */
	if (swc) swc->probe = swc->v.n == 0;
	for (r = 0; r < 2 && total >= 0; ++r)
		for (j = 0; j < anchors[r].n && (int)j < opt->max_matesw; ++j) {
			const int k = bb_matesw(opt, bns, pac, pes, &anchors[r].a[j], s[!r].l_seq, (uint8_t *)s[!r].seq, &hits[!r], swc, !r);
			if (k < 0) { total = -1; break; }   /* an alignment was requested from the device: this pass is void */
			total += k;
		}
	if (anchors[0].a != few[0]) free(anchors[0].a);
	if (anchors[1].a != few[1]) free(anchors[1].a);
	if (swc && swc->probe) { swc->probe = 0; if (swc->pending > 0) return -1; }   /* requests collected: the pass is void */
	return total;
}

/* ------------------------------------------------------------------------------------------------------ the best proper pair */
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
		memset(m->t, 0xff, (size_t)n * sizeof(double));   /* all-ones = NaN = "not computed yet": the model changes with every batch, and a thread with a handful of pairs must not pay for 65 k erfc + log (the term itself is never NaN: erfc >= 0, log(0) = -inf) */
		(void)k;
		m->avg = pe->avg; m->std = pe->std; m->low = pe->low; m->high = pe->high; m->a = opt->a; m->valid = 1;
	}
	{
		double *slot = &m->t[dist - pe->low], v = *slot;
		if (v != v) { double ns = (dist - pe->avg) / pe->std; v = *slot = .721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * opt->a; }
		return v;
	}
}


/* An end = one primary-assembly hit of either read, keyed for a sweep along the genome.
 *   pos:  contig << 32 | forward-strand offset in the contig
 *   info: score << 32 | index in its read's hit list << 2 | reverse strand << 1 | read */
#define END_READ(e)  ((int)((e).y & 1))
#define END_REV(e)   ((int)((e).y >> 1 & 1))
#define END_SLOT(e)  ((int)((e).y & 3))                 /* (strand, read): the four kinds of ends */
#define END_INDEX(e) ((int)((e).y << 32 >> 34))
#define END_SCORE(e) ((int)((e).y >> 32))

typedef struct { int score, runner_up, n_close; int pick[2]; } proper_t;   /* pick[r]: index of read r's hit in the best pair */

/* Among the first n_pri[r] hits of each read: every (hit of read 0, hit of read 1) on one contig whose distance fits a class of the
 * model is a candidate; its score is the two hit scores plus the insert-size term.  The best candidate wins, ties broken by a hash
 * of the pair and the read id.  Returns 0 when there is no candidate. */
static int best_proper_pair(const mem_opt_t *opt, const bntseq_t *bns, const mem_pestat_t pes[4], mem_alnreg_v hits[2], const int n_pri[2], int id, proper_t *out)
{
	BB_VEC(bb_pair64_t) ends = {0, 0, 0}, cands = {0, 0, 0};
	bb_pair64_t ends_few[16], cands_few[32];
	const int64_t l_pac = bns->l_pac;
	int latest[4] = { -1, -1, -1, -1 };   /* per kind of end: the last one the sweep has passed */
	int r, i;
	if (n_pri[0] + n_pri[1] <= 16) { ends.a = ends_few; ends.m = 16; }
	for (r = 0; r < 2; ++r)
		for (i = 0; i < n_pri[r]; ++i) {
			const mem_alnreg_t *h = &hits[r].a[i];
			const int rev = h->rb >= l_pac;
			const uint64_t fwd = rev ? (l_pac << 1) - 1 - h->rb : h->rb;
			bb_pair64_t e;
			e.x = (uint64_t)h->rid << 32 | (fwd - bns->anns[h->rid].offset);
			e.y = (uint64_t)h->score << 32 | i << 2 | rev << 1 | r;
			bb_vec_push(ends, e);
		}
	bb_sort_pair64(ends.n, ends.a);
	for (i = 0; i < (int)ends.n; ++i) {
		const bb_pair64_t *cur = &ends.a[i];
		int partner_rev;
		for (partner_rev = 0; partner_rev < 2; ++partner_rev) {   /* partners: ends of the other read on either strand, behind the sweep */
			const int cls = partner_rev << 1 | END_REV(*cur), kind = partner_rev << 1 | (END_READ(*cur) ^ 1);
			int k;
			if (pes[cls].failed) continue;
			for (k = latest[kind]; k >= 0; --k) {
				const bb_pair64_t *old = &ends.a[k];
				int64_t dist;
				int q;
				bb_pair64_t c;
				if (END_SLOT(*old) != kind) continue;
				dist = (int64_t)cur->x - old->x;
				if (dist > pes[cls].high) break;
				if (dist < pes[cls].low) continue;
				q = (int)(END_SCORE(*cur) + END_SCORE(*old) + pair_term(opt, &pes[cls], cls, dist) + .499);   /* order matters */
				if (q < 0) q = 0;
				c.y = (uint64_t)k << 32 | i;
				c.x = (uint64_t)q << 32 | (bb_mix64(c.y ^ id << 8) & 0xffffffffU);
				if (cands.a == 0) { cands.a = cands_few; cands.m = 32; }
				else if (cands.a == cands_few && cands.n == 32) { bb_pair64_t *h_ = bb_malloc(64 * sizeof(bb_pair64_t)); memcpy(h_, cands_few, sizeof(cands_few)); cands.a = h_; cands.m = 64; }
				bb_vec_push(cands, c);
			}
		}
		latest[END_SLOT(*cur)] = i;
	}
	memset(out, 0, sizeof(*out));
	if (cands.n) {
		const bb_pair64_t *win;
		int slack = opt->a + opt->b, k;
		if (opt->o_del + opt->e_del > slack) slack = opt->o_del + opt->e_del;
		if (opt->o_ins + opt->e_ins > slack) slack = opt->o_ins + opt->e_ins;
		bb_sort_pair64(cands.n, cands.a);
		win = &cands.a[cands.n - 1];
		{
			const bb_pair64_t *e1 = &ends.a[win->y >> 32], *e2 = &ends.a[win->y << 32 >> 32];
			out->pick[END_READ(*e1)] = END_INDEX(*e1);
			out->pick[END_READ(*e2)] = END_INDEX(*e2);
		}
		out->score = (int)(win->x >> 32);
		out->runner_up = cands.n > 1 ? (int)(cands.a[cands.n - 2].x >> 32) : 0;
		for (k = (int)cands.n - 2; k >= 0; --k)
			if (out->runner_up - (int)(cands.a[k].x >> 32) <= slack) ++out->n_close;
	}
	if (cands.a != cands_few) free(cands.a);
	if (ends.a != ends_few) free(ends.a);
	return out->score;
}

/* ------------------------------------------------------------------------------------------------------ records of a pair */
static inline int mapq_of_gap(int score_gap, int match) { return (int)(6.02 * score_gap / match + .499); }   /* order matters */

/* does read r have a second reportable primary hit (a supplementary candidate)?  Such pairs are not treated as proper pairs */
static int has_second_primary(const mem_opt_t *opt, const mem_alnreg_v *v, int n_pri)
{
	int j;
	for (j = 1; j < n_pri; ++j)
		if (v->a[j].secondary < 0 && v->a[j].score >= opt->T) return 1;
	return 0;
}

/* hit `pick` takes over as the head of its family in the all-hits numbering (it may have been a child of an ALT or of a better hit) */
static void make_family_head(mem_alnreg_v *v, int n_pri, int pick)
{
	const int head = v->a[pick].secondary_all;
	size_t j;
	if (head < 0 || head >= n_pri) return;
	assert(v->a[head].secondary_all < 0);
	for (j = 0; j < v->n; ++j)
		if (v->a[j].secondary_all == head || (int)j == head) v->a[j].secondary_all = pick;
	v->a[pick].secondary_all = -1;
}

/* the proper pair `pp` is reported: MAPQs, flags and text of both reads */
static void emit_proper(bb_samctx_t sc[2], bseq1_t s[2], mem_alnreg_v hits[2], const int n_pri[2], const proper_t *pp)
{
	const mem_opt_t *opt = sc[0].opt;
	const bntseq_t *bns = sc[0].bns;
	const int unpaired = hits[0].a[0].score + hits[1].a[0].score - opt->pen_unpaired;   /* what the two best hits score as a non-pair */
	const int rival = pp->runner_up > unpaired ? pp->runner_up : unpaired;
	int pick[2], q_read[2], q_pair, flags = 1, r, j, n_rec[2] = { 0, 0 };
	mem_aln_t main_rec[2], alt_rec[2], rec[2][2];
	char **xa[2];
	memset(main_rec, 0, sizeof(main_rec)); memset(alt_rec, 0, sizeof(alt_rec));
	q_pair = mapq_of_gap(pp->score - rival, opt->a);
	if (pp->n_close > 0) q_pair -= (int)(4.343 * log(pp->n_close + 1) + .499);
	if (q_pair < 0) q_pair = 0;
	if (q_pair > 60) q_pair = 60;
	q_pair = (int)(q_pair * (1. - .5 * (hits[0].a[0].frac_rep + hits[1].a[0].frac_rep)) + .499);   /* order matters (float sum, double product) */
	if (pp->score > unpaired) {   /* the pair beats the two best hits taken alone: report the pair's hits, lifted by the pair's quality */
		for (r = 0; r < 2; ++r) {
			mem_alnreg_t *h = &hits[r].a[pp->pick[r]];
			int q, cap;
			pick[r] = pp->pick[r];
			if (h->secondary >= 0) { h->sub = hits[r].a[h->secondary].score; h->secondary = -2; }
			q = bb_approx_mapq_se(opt, h);
			if (q < q_pair) q = q_pair < q + 40 ? q_pair : q + 40;
			cap = mapq_of_gap(h->score - h->csub, opt->a);
			q_read[r] = q < cap ? q : cap;
		}
		flags |= 2;
	} else {
		for (r = 0; r < 2; ++r) { pick[r] = 0; q_read[r] = bb_approx_mapq_se(opt, &hits[r].a[0]); }
	}
	for (r = 0; r < 2; ++r) make_family_head(&hits[r], n_pri[r], pick[r]);
	for (r = 0; r < 2; ++r) xa[r] = (opt->flag & MEM_F_ALL) ? 0 : bb_gen_alt(&sc[r], &hits[r], s[r].l_seq, s[r].seq);
	for (r = 0; r < 2; ++r) {
		main_rec[r] = bb_reg2aln(&sc[r], s[r].l_seq, s[r].seq, &hits[r].a[pick[r]]);
		main_rec[r].mapq = q_read[r];
		main_rec[r].flag |= 0x40 << r | flags;
		main_rec[r].XA = xa[r] ? xa[r][pick[r]] : 0;
		rec[r][n_rec[r]++] = main_rec[r];
		if (n_pri[r] < (int)hits[r].n) {   /* the best ALT hit rides along as a supplementary record */
			const mem_alnreg_t *alt = &hits[r].a[n_pri[r]];
			if (alt->score < opt->T || alt->secondary >= 0 || !alt->is_alt) continue;
			alt_rec[r] = bb_reg2aln(&sc[r], s[r].l_seq, s[r].seq, alt);
			alt_rec[r].flag |= 0x800 | 0x40 << r | flags;
			alt_rec[r].XA = xa[r] ? xa[r][n_pri[r]] : 0;
			rec[r][n_rec[r]++] = alt_rec[r];
		}
	}
	if (!sc[0].dry) {
		for (r = 0; r < 2; ++r) {
			bb_str_t text = {0, 0, 0};
			/* one buffer per read, sized for the usual single record up front (what bb_aln2sam reserves for one ordinary record) */
			bb_str_need(&text, (size_t)s[r].l_seq * 2 + strlen(s[r].name) + 448);
			for (j = 0; j < n_rec[r]; ++j) bb_aln2sam(opt, bns, &text, &s[r], n_rec[r], rec[r], j, &main_rec[!r]);
			s[r].sam = text.s;
		}
		if (strcmp(s[0].name, s[1].name) != 0) bb_fatal("mem_sam_pe", "paired reads have different names: \"%s\", \"%s\"\n", s[0].name, s[1].name);
	}
	for (r = 0; r < 2; ++r) {
		bb_cigar_free(&sc[r], main_rec[r].cigar); bb_cigar_free(&sc[r], alt_rec[r].cigar);
		if (!xa[r]) continue;
		for (j = 0; j < (int)hits[r].n; ++j) free(xa[r][j]);
		free(xa[r]);
	}
}

/* no proper pair: each read is reported the single-end way, with its mate's position filled in */
static void emit_separately(bb_samctx_t sc[2], const mem_pestat_t pes[4], bseq1_t s[2], mem_alnreg_v hits[2], const int n_pri[2])
{
	const mem_opt_t *opt = sc[0].opt;
	mem_aln_t top[2];
	int r, flags = 1;
	for (r = 0; r < 2; ++r) {   /* the record the mate fields of the other read point at: the best hit if reportable, else the best ALT hit */
		const mem_alnreg_t *h = 0;
		if (hits[r].n) {
			if (hits[r].a[0].score >= opt->T) h = &hits[r].a[0];
			else if (n_pri[r] < (int)hits[r].n && hits[r].a[n_pri[r]].score >= opt->T) h = &hits[r].a[n_pri[r]];
		}
		top[r] = bb_reg2aln(&sc[r], s[r].l_seq, s[r].seq, h);
	}
	if (!(opt->flag & MEM_F_NOPAIRING) && top[0].rid == top[1].rid && top[0].rid >= 0) {   /* still flagged proper if the two best hits fit the model */
		int64_t dist;
		const int cls = orientation_of(sc[0].bns->l_pac, hits[0].a[0].rb, hits[1].a[0].rb, &dist);
		if (!pes[cls].failed && within(&pes[cls], dist)) flags |= 2;
	}
	bb_reg2sam(&sc[0], &s[0], &hits[0], 0x41 | flags, &top[1]);
	bb_reg2sam(&sc[1], &s[1], &hits[1], 0x81 | flags, &top[0]);
	if (!sc[0].dry && strcmp(s[0].name, s[1].name) != 0) bb_fatal("mem_sam_pe", "paired reads have different names: \"%s\", \"%s\"\n", s[0].name, s[1].name);
	bb_cigar_free(&sc[0], top[0].cigar); bb_cigar_free(&sc[1], top[1].cigar);
}

/* Everything of a pair after rescue (bwamem_pair.c:302-419).  sc[r] carries the alignment cache of read r; when sc[0].dry is set no
 * text is produced (a pass that only collects the alignments the records will need). */
int bb_sam_pe(bb_samctx_t sc[2], const mem_pestat_t pes[4], uint64_t id, bseq1_t s[2], mem_alnreg_v hits[2], int rescue_done)
{
	const mem_opt_t *opt = sc[0].opt;
	proper_t pp;
	int n_pri[2], proper;
	(void)rescue_done;
	n_pri[0] = bb_mark_primary_se(opt, (int)hits[0].n, hits[0].a, id << 1 | 0);
	n_pri[1] = bb_mark_primary_se(opt, (int)hits[1].n, hits[1].a, id << 1 | 1);
	if (opt->flag & MEM_F_PRIMARY5) { bb_reorder_primary5(opt->T, &hits[0]); bb_reorder_primary5(opt->T, &hits[1]); }
	proper = !(opt->flag & MEM_F_NOPAIRING) && n_pri[0] && n_pri[1]
	      && best_proper_pair(opt, sc[0].bns, pes, hits, n_pri, (int)id, &pp) > 0
	      && !has_second_primary(opt, &hits[0], n_pri[0]) && !has_second_primary(opt, &hits[1], n_pri[1]);
	if (proper) emit_proper(sc, s, hits, n_pri, &pp);
	else emit_separately(sc, pes, s, hits, n_pri);
	return 0;
}
