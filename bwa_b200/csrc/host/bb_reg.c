/* bb_reg.c -- host side of region post-processing, for the reads the device tail (bwag_tail.cu) hands back:
 * removal of contained regions and joining of colinear neighbours, parent/child (primary/secondary) marking,
 * single-end MAPQ.  Behaviour to match: bwamem.c:417-584 and 982-1030.
 *
 * What bit-exactness pins down here, and nothing more: the float/double operand types of every threshold test, the
 * evaluation order of the floating-point expressions (marked "order matters"), and the unstable introsort of bb_sort.h
 * (equal keys must land where the reference's sort puts them).  Passes, helpers and names are this file's own; the
 * device formulation of the same rules for the common small case is t_dedup / t_mark_primary / t_mapq_se.
 */
#include <math.h>
#include <pthread.h>
#include <limits.h>
#include "bb_host.h"
#include "bb_sort.h"

#define by_ref_end(a, b) ((a).re < (b).re)
BB_SORT_DEFINE(static, sort_regs_by_end, mem_alnreg_t, by_ref_end)

#define by_score_pos(a, b) ((a).score > (b).score || ((a).score == (b).score && ((a).rb < (b).rb || ((a).rb == (b).rb && (a).qb < (b).qb))))
BB_SORT_DEFINE(static, sort_regs_by_score, mem_alnreg_t, by_score_pos)

#define by_score_hash(a, b) ((a).score > (b).score || ((a).score == (b).score && ((a).is_alt < (b).is_alt || ((a).is_alt == (b).is_alt && (a).hash < (b).hash))))
BB_SORT_DEFINE(static, sort_regs_score_hash, mem_alnreg_t, by_score_hash)

#define by_alt_score_hash(a, b) ((a).is_alt < (b).is_alt || ((a).is_alt == (b).is_alt && ((a).score > (b).score || ((a).score == (b).score && (a).hash < (b).hash))))
BB_SORT_DEFINE(static, sort_regs_alt_score_hash, mem_alnreg_t, by_alt_score_hash)

static inline int imin(int x, int y) { return x < y ? x : y; }
static inline int imax(int x, int y) { return x > y ? x : y; }
static inline int64_t lmin(int64_t x, int64_t y) { return x < y ? x : y; }
static inline int qspan(const mem_alnreg_t *r) { return r->qe - r->qb; }
static inline int64_t rspan(const mem_alnreg_t *r) { return r->re - r->rb; }
static inline int is_dropped(const mem_alnreg_t *r) { return r->qe <= r->qb; }   /* an emptied query interval marks a region for removal */
static inline void drop(mem_alnreg_t *r) { r->qe = r->qb; }

/* largest single-event penalty of the scoring scheme: two hits whose scores differ by no more count as "equally good" */
static inline int tie_slack(const mem_opt_t *o)
{
	return imax(imax(o->a + o->b, o->o_del + o->e_del), o->o_ins + o->e_ins);
}

/* ---------------------------------------------------------------------------------------------------- joining neighbours */

/* Geometry test for joining `left` and `right` (left starts first on the reference): band width for the joint global
 * alignment, or -1 when the two are not colinear enough.  Thresholds are float constants against a double (bwamem.c:441-449). */
static int join_band(const mem_opt_t *opt, int64_t l_pac, const mem_alnreg_t *left, const mem_alnreg_t *right)
{
	int skew;
	double slope_gap;
	if (left->rb < l_pac && right->rb >= l_pac) return -1;                            /* different strands */
	if (left->qb >= right->qb || left->qe >= right->qe || left->re >= right->re) return -1;   /* not in order on both axes */
	skew = (int)((left->re - right->rb) - (left->qe - right->qb));
	if (skew < 0) skew = -skew;
	slope_gap = (double)(left->re - right->rb) / (right->re - left->rb) - (double)(left->qe - right->qb) / (right->qe - left->qb);   /* order matters */
	if (slope_gap < 0.) slope_gap = -slope_gap;
	if (left->re < right->rb || left->qe < right->qb) {   /* a gap between them: stricter */
		if (skew > opt->w << 1 || slope_gap >= 0.05f) return -1;
	} else if (skew > opt->w << 2 || slope_gap >= 0.05f * 2) return -1;
	return imin(skew + left->w + right->w, opt->w << 2);
}

/* Score of the joint alignment if joining pays: >0 score (band in *band), 0 no, -1 the global alignment is not in the cache
 * yet (it has been requested; the caller replays the read after the next device round). */
static int join_score(const mem_opt_t *opt, const bntseq_t *bns, bb_gcache_t *gc, const mem_alnreg_t *left, const mem_alnreg_t *right, int *band)
{
	const bb_galn_t *g;
	int w, by_query, by_ref, sum;
	if (!bns || !gc) return 0;
	if ((w = join_band(opt, bns->l_pac, left, right)) < 0) return 0;
	if ((g = bb_gcache_get(gc, BWAG_G_SCORE, left->qb, right->qe, left->rb, right->re, w, 0)) == 0) return -1;
	/* the score the two would have if scores scaled with the joint span: the joint alignment must reach 90% of it */
	sum = right->score + left->score;
	by_query = (int)((double)(right->qe - left->qb) / (qspan(right) + qspan(left)) * sum + .499);   /* order matters */
	by_ref = (int)((double)(right->re - left->rb) / (rspan(right) + rspan(left)) * sum + .499);
	if ((double)g->score / imax(by_query, by_ref) < 0.90f) return 0;
	*band = w;
	return g->score;
}

static void absorb(mem_alnreg_t *keep, mem_alnreg_t *gone, int score, int band)   /* keep := keep + gone joined; gone is dropped */
{
	keep->n_comp += gone->n_comp + 1;
	keep->seedcov = imax(keep->seedcov, gone->seedcov);
	keep->sub = imax(keep->sub, gone->sub);
	keep->csub = imax(keep->csub, gone->csub);
	keep->qb = gone->qb; keep->rb = gone->rb;
	keep->truesc = keep->score = score;
	keep->w = band;
	gone->qb = gone->qe;
}

static int compact(int n, mem_alnreg_t *a)   /* close the holes left by dropped regions */
{
	int i, kept = 0;
	for (i = 0; i < n; ++i)
		if (!is_dropped(&a[i])) { if (kept != i) a[kept] = a[i]; ++kept; }
	return kept;
}

/* Sweep in order of reference end: each region meets the earlier ones that end within max_chain_gap of its start.  Of two that
 * overlap by more than mask_level_redun on both axes the lower-scoring goes; an earlier one that is colinear is joined. */
static int sweep_redundant(const mem_opt_t *opt, const bntseq_t *bns, bb_gcache_t *gc, int n, mem_alnreg_t *a)
{
	int i, j;
	for (i = 1; i < n; ++i) {
		mem_alnreg_t *cur = &a[i];
		for (j = i - 1; j >= 0; --j) {
			mem_alnreg_t *old = &a[j];
			int64_t ov_ref, ov_qry;
			if (cur->rid != old->rid || cur->rb >= old->re + opt->max_chain_gap) break;
			if (old->qe == old->qb) continue;
			ov_ref = old->re - cur->rb;
			ov_qry = old->qb < cur->qb ? old->qe - cur->qb : cur->qe - old->qb;
			if (ov_ref > opt->mask_level_redun * lmin(rspan(old), rspan(cur)) && ov_qry > opt->mask_level_redun * imin(qspan(old), qspan(cur))) {
				if (cur->score < old->score) { drop(cur); break; }
				drop(old);
			} else if (old->rb < cur->rb) {
				int band, sc = join_score(opt, bns, gc, old, cur, &band);
				if (sc < 0) return -1;
				if (sc > 0) absorb(cur, old, sc, band);
			}
		}
	}
	return 0;
}

/* Returns the new count, or -1 if a joint alignment is pending on the device (the caller restores the array and retries). */
int bb_sort_dedup_patch(const mem_opt_t *opt, const bntseq_t *bns, bb_gcache_t *gc, int l_query, int n, mem_alnreg_t *a)
{
	int i;
	(void)l_query;
	if (n <= 1) return n;
	sort_regs_by_end(n, a);
	for (i = 0; i < n; ++i) a[i].n_comp = 1;
	if (sweep_redundant(opt, bns, gc, n, a) < 0) return -1;
	n = compact(n, a);
	sort_regs_by_score(n, a);
	for (i = 1; i < n; ++i)   /* exact duplicates (same score, same starts) are neighbours now */
		if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) drop(&a[i]);
	return n > 1 ? 1 + compact(n - 1, a + 1) : n;
}

void bb_regs_make_room(mem_alnreg_v *v) /* room for one more region */
{
	if (v->m & BB_BORROWED) {
		size_t cap = v->n + 4;
		mem_alnreg_t *na = bb_malloc(cap * sizeof(mem_alnreg_t));
		memcpy(na, v->a, v->n * sizeof(mem_alnreg_t));
		v->a = na; v->m = cap;
	} else bb_vec_reserve(*v, v->n + 1);
}

/* ------------------------------------------------------------------------------------------------- parents and children */

/* do the query intervals of x and y overlap by mask_level of the shorter one? (float product against an int) */
static inline int covers(const mem_opt_t *opt, const mem_alnreg_t *x, const mem_alnreg_t *y)
{
	const int lo = imax(x->qb, y->qb), hi = imin(x->qe, y->qe);
	return hi > lo && hi - lo >= imin(qspan(x), qspan(y)) * opt->mask_level;
}

/* Regions a[0..n) are in rank order.  A region becomes the child of the first earlier parent it overlaps, else a parent itself.
 * The parent records its first child's score (sub) and counts the children that are as good as itself (sub_n). */
static void assign_parents(const mem_opt_t *opt, int n, mem_alnreg_t *a, bb_int_v *parents_)
{
	bb_int_v parents = *parents_;
	const int slack = tie_slack(opt);
	int i;
	parents.n = 0;
	bb_vec_push(parents, 0);
	for (i = 1; i < n; ++i) {
		size_t k;
		for (k = 0; k < parents.n; ++k) if (covers(opt, &a[parents.a[k]], &a[i])) break;
		if (k == parents.n) { bb_vec_push(parents, i); continue; }
		{
			mem_alnreg_t *par = &a[parents.a[k]];
			if (par->sub == 0) par->sub = a[i].score;
			if (par->score - a[i].score <= slack && (par->is_alt || !a[i].is_alt)) ++par->sub_n;
			a[i].secondary = parents.a[k];
		}
	}
	*parents_ = parents;
}

/* Ranks the regions (score, then primary assembly before ALT, then a per-read hash), marks children, and -- when ALT hits are
 * present -- ranks again with the primary-assembly hits first and marks among those alone.  Returns their number. */
int bb_mark_primary_se(const mem_opt_t *opt, int n, mem_alnreg_t *a, int64_t id)
{
	int small[32];
	bb_int_v scratch = {0, 0, 0};
	int i, n_pri = 0;
	if (n == 0) return 0;
	if (n <= 32) { scratch.a = small; scratch.m = 32; }   /* the usual case: no heap traffic */
	for (i = 0; i < n; ++i) {
		mem_alnreg_t *r = &a[i];
		r->sub = r->alt_sc = 0;
		r->secondary = r->secondary_all = -1;
		r->hash = bb_mix64((uint64_t)(id + i));
		n_pri += !r->is_alt;
	}
	sort_regs_score_hash(n, a);
	assign_parents(opt, n, a, &scratch);
	for (i = 0; i < n; ++i) {
		mem_alnreg_t *r = &a[i];
		r->secondary_all = i;   /* for now: this region's rank in the all-hits order */
		if (!r->is_alt && r->secondary >= 0 && a[r->secondary].is_alt) r->alt_sc = a[r->secondary].score;
	}
	if (n_pri == n) {
		for (i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
	} else {
		int *new_rank;
		if (scratch.a != small) bb_vec_reserve(scratch, (size_t)n);
		new_rank = scratch.a;
		if (n_pri > 0) sort_regs_alt_score_hash(n, a);
		for (i = 0; i < n; ++i) new_rank[a[i].secondary_all] = i;
		for (i = 0; i < n; ++i) {
			mem_alnreg_t *r = &a[i];
			if (r->secondary < 0) { r->secondary_all = -1; continue; }
			r->secondary_all = new_rank[r->secondary];
			if (r->is_alt) r->secondary = INT_MAX;
		}
		if (n_pri > 0) {
			for (i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
			assign_parents(opt, n_pri, a, &scratch);
		}
	}
	if (scratch.a != small) free(scratch.a);
	return n_pri;
}

/* -5: of the reportable primary hits the one that starts leftmost on the read goes first (bwamem.c:1008-1030) */
void bb_reorder_primary5(int T, mem_alnreg_v *v)
{
	size_t k;
	int reportable = 0, best_qb = INT_MAX, pick = -1;
	for (k = 0; k < v->n; ++k) {
		const mem_alnreg_t *r = &v->a[k];
		if (r->secondary >= 0 || r->is_alt || r->score < T) continue;
		++reportable;
		if (r->qb < best_qb) { best_qb = r->qb; pick = (int)k; }
	}
	if (reportable <= 1 || pick == 0) return;
	{ mem_alnreg_t t = v->a[0]; v->a[0] = v->a[pick]; v->a[pick] = t; }
	for (k = 1; k < v->n; ++k) {   /* references to the two swapped slots follow them */
		mem_alnreg_t *r = &v->a[k];
		if (r->secondary == 0) r->secondary = pick; else if (r->secondary == pick) r->secondary = 0;
		if (r->secondary_all == 0) r->secondary_all = pick; else if (r->secondary_all == pick) r->secondary_all = 0;
	}
}

/* ------------------------------------------------------------------------------------------------------------------ MAPQ */

/* log of a small non-negative integer (seed coverage, number of sub-optimal hits + 1): the same libm values, tabulated once */
#define LOGTAB_N 4096
static double g_logtab[LOGTAB_N];
static pthread_once_t g_logtab_once = PTHREAD_ONCE_INIT;
static void logtab_init(void) { int i; for (i = 0; i < LOGTAB_N; ++i) g_logtab[i] = log(i); }
static inline double log_of_int(int n)
{
	if ((unsigned)n >= LOGTAB_N) return log(n);
	pthread_once(&g_logtab_once, logtab_init);
	return g_logtab[n];
}

/* Single-end mapping quality of a region from its score, the best competing score and the seed coverage (bwamem.c:982-1006).
 * Every line that says "order matters" is a double expression truncated to int: operands and order are the reference's. */
int bb_approx_mapq_se(const mem_opt_t *opt, const mem_alnreg_t *r)
{
	const int span = qspan(r) > rspan(r) ? qspan(r) : (int)rspan(r);
	int rival = r->sub ? r->sub : opt->min_seed_len * opt->a, q;
	double identity;
	if (r->csub > rival) rival = r->csub;
	if (rival >= r->score) return 0;
	identity = 1. - (double)(span * opt->a - r->score) / (opt->a + opt->b) / span;   /* order matters */
	if (r->score == 0) q = 0;
	else if (opt->mapQ_coef_len > 0) {
		double scale = span < opt->mapQ_coef_len ? 1. : opt->mapQ_coef_fac / log(span);
		scale *= identity * identity;
		q = (int)(6.02 * (r->score - rival) / opt->a * scale * scale + .499);   /* order matters */
	} else {
		q = (int)(MEM_MAPQ_COEF * (1. - (double)rival / r->score) * log_of_int(r->seedcov) + .499);   /* order matters */
		if (identity < 0.95) q = (int)(q * identity * identity + .499);
	}
	if (r->sub_n > 0) q -= (int)(4.343 * log_of_int(r->sub_n + 1) + .499);
	if (q > 60) q = 60;
	if (q < 0) q = 0;
	return (int)(q * (1. - r->frac_rep) + .499);
}
