/* bb_reg.c -- alignment regions of one read after extension: redundancy removal / merging of colinear
 * neighbours, primary-vs-secondary marking, approximate single-end MAPQ.
 *
 * Integer and floating-point expressions keep the reference's operand types and order
 * (bwamem.c:417-584, 982-1030): thresholds are float products compared against ints, and the
 * sorts are the unstable introsort of bb_sort.h, so equal keys end up in the same order.
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

/* Can regions a (left) and b (right) be joined by one banded global alignment?  Returns the joint
 * score (>0) and the band in *w_out, 0 if not, -1 if the alignment had to be requested from the
 * device first (bwamem.c:432-461). */
static int try_patch(const mem_opt_t *opt, const bntseq_t *bns, bb_gcache_t *gc, const mem_alnreg_t *a, const mem_alnreg_t *b, int *w_out)
{
	int w, score, q_s, r_s;
	double r;
	const bb_galn_t *g;
	if (bns == 0 || gc == 0) return 0;
	if (a->rb < bns->l_pac && b->rb >= bns->l_pac) return 0;
	if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;
	w = (int)((a->re - b->rb) - (a->qe - b->qb));
	w = w > 0 ? w : -w;
	r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);
	r = r > 0. ? r : -r;
	if (a->re < b->rb || a->qe < b->qb) {
		if (w > opt->w << 1 || r >= 0.05f) return 0;
	} else if (w > opt->w << 2 || r >= 0.05f * 2) return 0;
	w += a->w + b->w;
	w = w < opt->w << 2 ? w : opt->w << 2;
	g = bb_gcache_get(gc, BWAG_G_SCORE, a->qb, b->qe, a->rb, b->re, w, 0);
	if (!g) return -1;
	score = g->score;
	q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);
	r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);
	if ((double)score / (q_s > r_s ? q_s : r_s) < 0.90f) return 0;
	*w_out = w;
	return score;
}

/* bwamem.c:463-515.  Returns the new count, or -1 if a patch alignment is pending on the device
 * (the caller restores the array and retries after the next device round). */
int bb_sort_dedup_patch(const mem_opt_t *opt, const bntseq_t *bns, bb_gcache_t *gc, int l_query, int n, mem_alnreg_t *a)
{
	int m, i, j;
	(void)l_query;
	if (n <= 1) return n;
	sort_regs_by_end(n, a);
	for (i = 0; i < n; ++i) a[i].n_comp = 1;
	for (i = 1; i < n; ++i) {
		mem_alnreg_t *p = &a[i];
		if (p->rid != a[i - 1].rid || p->rb >= a[i - 1].re + opt->max_chain_gap) continue;
		for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + opt->max_chain_gap; --j) {
			mem_alnreg_t *q = &a[j];
			int64_t o_r, o_q, m_r, m_q;
			int score, w;
			if (q->qe == q->qb) continue;
			o_r = q->re - p->rb;
			o_q = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
			m_r = q->re - q->rb < p->re - p->rb ? q->re - q->rb : p->re - p->rb;
			m_q = q->qe - q->qb < p->qe - p->qb ? q->qe - q->qb : p->qe - p->qb;
			if (o_r > opt->mask_level_redun * m_r && o_q > opt->mask_level_redun * m_q) {
				if (p->score < q->score) { p->qe = p->qb; break; }
				else q->qe = q->qb;
			} else if (q->rb < p->rb && (score = try_patch(opt, bns, gc, q, p, &w)) != 0) {
				if (score < 0) return -1;
				p->n_comp += q->n_comp + 1;
				p->seedcov = p->seedcov > q->seedcov ? p->seedcov : q->seedcov;
				p->sub = p->sub > q->sub ? p->sub : q->sub;
				p->csub = p->csub > q->csub ? p->csub : q->csub;
				p->qb = q->qb; p->rb = q->rb;
				p->truesc = p->score = score;
				p->w = w;
				q->qb = q->qe;
			}
		}
	}
	for (i = 0, m = 0; i < n; ++i)
		if (a[i].qe > a[i].qb) { if (m != i) a[m] = a[i]; ++m; }
	n = m;
	sort_regs_by_score(n, a);
	for (i = 1; i < n; ++i)
		if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
	for (i = 1, m = 1; i < n; ++i)
		if (a[i].qe > a[i].qb) { if (m != i) a[m] = a[i]; ++m; }
	return m;
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

static void mark_core(const mem_opt_t *opt, int n, mem_alnreg_t *a, bb_int_v *z_)
{
	bb_int_v z = *z_;
	int i, tmp;
	size_t k;
	tmp = opt->a + opt->b;
	if (opt->o_del + opt->e_del > tmp) tmp = opt->o_del + opt->e_del;
	if (opt->o_ins + opt->e_ins > tmp) tmp = opt->o_ins + opt->e_ins;
	z.n = 0;
	bb_vec_push(z, 0);
	for (i = 1; i < n; ++i) {
		for (k = 0; k < z.n; ++k) {
			int j = z.a[k];
			int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
			int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
			if (e_min > b_max) {
				int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
				if (e_min - b_max >= min_l * opt->mask_level) {
					if (a[j].sub == 0) a[j].sub = a[i].score;
					if (a[j].score - a[i].score <= tmp && (a[j].is_alt || !a[i].is_alt)) ++a[j].sub_n;
					break;
				}
			}
		}
		if (k == z.n) bb_vec_push(z, i);
		else a[i].secondary = z.a[k];
	}
	*z_ = z;
}

/* bwamem.c:547-584 */
int bb_mark_primary_se(const mem_opt_t *opt, int n, mem_alnreg_t *a, int64_t id)
{
	int zstack[32];
	bb_int_v z = {0, 0, 0};
	int i, n_pri = 0;
	if (n == 0) return 0;
	if (n <= 32) { z.a = zstack; z.m = 32; }   /* the usual case: no heap traffic */
	for (i = 0; i < n; ++i) {
		a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1;
		a[i].hash = bb_mix64((uint64_t)(id + i));
		if (!a[i].is_alt) ++n_pri;
	}
	sort_regs_score_hash(n, a);
	mark_core(opt, n, a, &z);
	for (i = 0; i < n; ++i) {
		mem_alnreg_t *p = &a[i];
		p->secondary_all = i;
		if (!p->is_alt && p->secondary >= 0 && a[p->secondary].is_alt) p->alt_sc = a[p->secondary].score;
	}
	if (n_pri >= 0 && n_pri < n) {
		if (z.a != zstack) bb_vec_reserve(z, (size_t)n);
		if (n_pri > 0) sort_regs_alt_score_hash(n, a);
		for (i = 0; i < n; ++i) z.a[a[i].secondary_all] = i;
		for (i = 0; i < n; ++i) {
			if (a[i].secondary >= 0) {
				a[i].secondary_all = z.a[a[i].secondary];
				if (a[i].is_alt) a[i].secondary = INT_MAX;
			} else a[i].secondary_all = -1;
		}
		if (n_pri > 0) {
			for (i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
			mark_core(opt, n_pri, a, &z);
		}
	} else for (i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
	if (z.a != zstack) free(z.a);
	return n_pri;
}

/* bwamem.c:1008-1030 */
void bb_reorder_primary5(int T, mem_alnreg_v *a)
{
	int n_pri = 0, left_st = INT_MAX, left_k = -1;
	size_t k;
	mem_alnreg_t t;
	for (k = 0; k < a->n; ++k)
		if (a->a[k].secondary < 0 && !a->a[k].is_alt && a->a[k].score >= T) ++n_pri;
	if (n_pri <= 1) return;
	for (k = 0; k < a->n; ++k) {
		mem_alnreg_t *p = &a->a[k];
		if (p->secondary >= 0 || p->is_alt || p->score < T) continue;
		if (p->qb < left_st) { left_st = p->qb; left_k = (int)k; }
	}
	if (left_k == 0) return;
	t = a->a[0]; a->a[0] = a->a[left_k]; a->a[left_k] = t;
	for (k = 1; k < a->n; ++k) {
		mem_alnreg_t *p = &a->a[k];
		if (p->secondary == 0) p->secondary = left_k;
		else if (p->secondary == left_k) p->secondary = 0;
		if (p->secondary_all == 0) p->secondary_all = left_k;
		else if (p->secondary_all == left_k) p->secondary_all = 0;
	}
}

/* bwamem.c:982-1006 */
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

int bb_approx_mapq_se(const mem_opt_t *opt, const mem_alnreg_t *a)
{
	int mapq, l, sub = a->sub ? a->sub : opt->min_seed_len * opt->a;
	double identity;
	sub = a->csub > sub ? a->csub : sub;
	if (sub >= a->score) return 0;
	l = a->qe - a->qb > a->re - a->rb ? a->qe - a->qb : (int)(a->re - a->rb);
	identity = 1. - (double)(l * opt->a - a->score) / (opt->a + opt->b) / l;
	if (a->score == 0) mapq = 0;
	else if (opt->mapQ_coef_len > 0) {
		double tmp;
		tmp = l < opt->mapQ_coef_len ? 1. : opt->mapQ_coef_fac / log(l);
		tmp *= identity * identity;
		mapq = (int)(6.02 * (a->score - sub) / opt->a * tmp * tmp + .499);
	} else {
		mapq = (int)(MEM_MAPQ_COEF * (1. - (double)sub / a->score) * log_of_int(a->seedcov) + .499);
		mapq = identity < 0.95 ? (int)(mapq * identity * identity + .499) : mapq;
	}
	if (a->sub_n > 0) mapq -= (int)(4.343 * log_of_int(a->sub_n + 1) + .499);
	if (mapq > 60) mapq = 60;
	if (mapq < 0) mapq = 0;
	mapq = (int)(mapq * (1. - a->frac_rep) + .499);
	return mapq;
}
