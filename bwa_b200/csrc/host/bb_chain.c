/* bb_chain.c -- seeds -> colinear chains -> filtered chains (host side of the hot path).
 *
 * Semantics follow the reference exactly (bwamem.c:216-411, 597-654) because which seed joins which
 * chain, and the order of equal-position / equal-weight chains, decide which extensions are run:
 *   - chains live in an ordered multimap keyed by the reference position of their first seed; a new
 *     seed looks up "the equal key, else the predecessor" and either merges into that chain or opens
 *     a new one.  With duplicate keys the element found depends on the shape of klib's B-tree
 *     (kbtree.h:113-127 search, 184-224 insert with pre-emptive splits, node order t=5 for 40-byte
 *     keys in 512-byte nodes), so the same tree is kept here, over chain indices;
 *   - chains are weighted, sorted by weight with the unstable introsort (bb_sort.h) and filtered by
 *     the overlap rules of mem_chain_flt.
 */
#include <math.h>
#include "bb_host.h"
#include "bb_sort.h"

/* ---------------------------------------------------------------- ordered multimap of chains */
#define BT_T 5                 /* minimum degree: ((512-4-8)/(8+40)+1)>>1  (kbtree.h:59) */
#define BT_MAXK (2 * BT_T - 1)

typedef struct {
	int n, internal;
	int key[BT_MAXK];          /* chain indices, ordered by chain pos */
	int child[BT_MAXK + 1];    /* node indices */
} bt_node_t;

struct bb_chainer {
	BB_VEC(bt_node_t) nodes;
	int root, n_keys;
	bb_chain_v chains;         /* storage of the chains being built */
	/* bump arena for seed arrays */
	char **blocks; size_t *block_cap; size_t n_blocks, m_blocks, cur_block, cur_off, block_sz;
	BB_VEC(int) order;         /* in-order traversal scratch */
};

bb_chainer_t *bb_chainer_new(void)
{
	bb_chainer_t *c = bb_calloc(1, sizeof(*c));
	c->block_sz = 1 << 16;
	return c;
}

void bb_chainer_free(bb_chainer_t *c)
{
	size_t i;
	if (!c) return;
	for (i = 0; i < c->n_blocks; ++i) free(c->blocks[i]);
	free(c->blocks); free(c->block_cap); free(c->nodes.a); free(c->chains.a); free(c->order.a);
	free(c);
}

static void *arena_alloc(bb_chainer_t *c, size_t bytes)
{
	void *p;
	bytes = (bytes + 15) & ~(size_t)15;
	/* advance to a block with room; blocks are kept (and reused) across reads */
	while (c->cur_block < c->n_blocks && c->cur_off + bytes > c->block_cap[c->cur_block]) { ++c->cur_block; c->cur_off = 0; }
	if (c->cur_block >= c->n_blocks) {
		size_t sz = bytes > c->block_sz ? bytes : c->block_sz;
		if (c->n_blocks == c->m_blocks) {
			c->m_blocks = c->m_blocks ? c->m_blocks << 1 : 8;
			c->blocks = bb_realloc(c->blocks, c->m_blocks * sizeof(char *));
			c->block_cap = bb_realloc(c->block_cap, c->m_blocks * sizeof(size_t));
		}
		c->blocks[c->n_blocks] = bb_malloc(sz);
		c->block_cap[c->n_blocks] = sz;
		c->cur_block = c->n_blocks++; c->cur_off = 0;
	}
	p = c->blocks[c->cur_block] + c->cur_off;
	c->cur_off += bytes;
	return p;
}

static void arena_reset(bb_chainer_t *c) { c->cur_block = 0; c->cur_off = 0; }

static int bt_new_node(bb_chainer_t *c, int internal)
{
	bt_node_t z;
	memset(&z, 0, sizeof(z));
	z.internal = internal;
	bb_vec_push(c->nodes, z);
	return (int)c->nodes.n - 1;
}

#define CPOS(c, k) ((c)->chains.a[k].pos)

/* index of the first key == pos if present, else of the last key < pos (may be -1); *cmp = sign(pos - key) at that slot */
static int bt_locate(const bb_chainer_t *c, const bt_node_t *x, int64_t pos, int *cmp)
{
	int lo = 0, hi = x->n;
	if (x->n == 0) return -1;
	while (lo < hi) {
		int mid = (lo + hi) >> 1;
		if (CPOS(c, x->key[mid]) < pos) lo = mid + 1; else hi = mid;
	}
	if (lo == x->n) { *cmp = 1; return x->n - 1; }
	*cmp = pos < CPOS(c, x->key[lo]) ? -1 : pos > CPOS(c, x->key[lo]) ? 1 : 0;
	if (*cmp < 0) --lo;
	return lo;
}

/* the chain with the equal key if the descent meets one, else the closest smaller key seen (kbtree.h:151-168) */
static int bt_floor(const bb_chainer_t *c, int64_t pos)
{
	int x = c->root, lower = -1;
	for (;;) {
		const bt_node_t *nd = &c->nodes.a[x];
		int cmp = 0, i = bt_locate(c, nd, pos, &cmp);
		if (i >= 0 && cmp == 0) return nd->key[i];
		if (i >= 0) lower = nd->key[i];
		if (!nd->internal) return lower;
		x = nd->child[i + 1];
	}
}

/* split the full child y = x.child[i]; its median key moves up into x at slot i */
static void bt_split(bb_chainer_t *c, int xi, int i, int yi)
{
	int zi = bt_new_node(c, c->nodes.a[yi].internal);
	bt_node_t *x = &c->nodes.a[xi], *y = &c->nodes.a[yi], *z = &c->nodes.a[zi];
	z->n = BT_T - 1;
	memcpy(z->key, y->key + BT_T, sizeof(int) * (BT_T - 1));
	if (y->internal) memcpy(z->child, y->child + BT_T, sizeof(int) * BT_T);
	y->n = BT_T - 1;
	memmove(x->child + i + 2, x->child + i + 1, sizeof(int) * (x->n - i));
	x->child[i + 1] = zi;
	memmove(x->key + i + 1, x->key + i, sizeof(int) * (x->n - i));
	x->key[i] = y->key[BT_T - 1];
	++x->n;
}

static void bt_insert(bb_chainer_t *c, int chain)
{
	int64_t pos = CPOS(c, chain);
	int xi, cmp;
	++c->n_keys;
	if (c->nodes.a[c->root].n == BT_MAXK) {
		int s = bt_new_node(c, 1);
		c->nodes.a[s].child[0] = c->root;
		bt_split(c, s, 0, c->root);
		c->root = s;
	}
	xi = c->root;
	for (;;) {
		bt_node_t *x = &c->nodes.a[xi];
		int i = bt_locate(c, x, pos, &cmp);
		if (!x->internal) {
			if (i != x->n - 1) memmove(x->key + i + 2, x->key + i + 1, sizeof(int) * (x->n - i - 1));
			x->key[i + 1] = chain;
			++x->n;
			return;
		}
		++i;
		if (c->nodes.a[x->child[i]].n == BT_MAXK) {
			bt_split(c, xi, i, x->child[i]);
			x = &c->nodes.a[xi]; /* nodes vector may have moved */
			if (pos > CPOS(c, x->key[i])) ++i;
		}
		xi = x->child[i];
	}
}

static void bt_inorder(bb_chainer_t *c, int xi)
{
	const bt_node_t *x = &c->nodes.a[xi];
	int i;
	for (i = 0; i < x->n; ++i) {
		if (x->internal) { bt_inorder(c, x->child[i]); x = &c->nodes.a[xi]; }
		bb_vec_push(c->order, x->key[i]);
	}
	if (x->internal) bt_inorder(c, x->child[x->n]);
}

/* ---------------------------------------------------------------- chaining */

/* 1 if seed p was absorbed by (or is redundant with) chain c (bwamem.c:216-237) */
static int try_merge(bb_chainer_t *ws, const mem_opt_t *opt, int64_t l_pac, bb_chain_t *c, const bb_seed_t *p, int seed_rid)
{
	const bb_seed_t *last = &c->seeds[c->n - 1], *first = &c->seeds[0];
	int64_t qend = last->qbeg + last->len, rend = last->rbeg + last->len, x, y;
	if (seed_rid != c->rid) return 0;
	if (p->qbeg >= first->qbeg && p->qbeg + p->len <= qend && p->rbeg >= first->rbeg && p->rbeg + p->len <= rend) return 1;
	if ((last->rbeg < l_pac || first->rbeg < l_pac) && p->rbeg >= l_pac) return 0;
	x = p->qbeg - last->qbeg;
	y = p->rbeg - last->rbeg;
	if (y >= 0 && x - y <= opt->w && y - x <= opt->w && x - last->len < opt->max_chain_gap && y - last->len < opt->max_chain_gap) {
		if (c->n == c->m) {
			bb_seed_t *ns = arena_alloc(ws, sizeof(bb_seed_t) * (size_t)c->m * 2);
			memcpy(ns, c->seeds, sizeof(bb_seed_t) * c->n);
			c->seeds = ns; c->m <<= 1;
		}
		c->seeds[c->n++] = *p;
		return 1;
	}
	return 0;
}

void bb_chain_build(bb_chainer_t *ws, const mem_opt_t *opt, const bntseq_t *bns, int l_query,
                    int n_intv, const bwtintv_t *intv, const int64_t *seed_beg, const int64_t *rbeg, bb_chain_v *out)
{
	int i, b = 0, e = 0, l_rep = 0;
	int64_t l_pac = bns->l_pac;
	size_t k;
	out->n = 0;
	if (l_query < opt->min_seed_len) return;
	ws->nodes.n = 0; ws->chains.n = 0; ws->n_keys = 0; ws->order.n = 0;
	arena_reset(ws);
	ws->root = bt_new_node(ws, 0);
	/* bases covered by over-represented seeds (bwamem.c:291-298) */
	for (i = 0; i < n_intv; ++i) {
		int sb = (int)(intv[i].info >> 32), se = (int)(uint32_t)intv[i].info;
		if (intv[i].x[2] <= (uint64_t)opt->max_occ) continue;
		if (sb > e) { l_rep += e - b; b = sb; e = se; }
		else if (se > e) e = se;
	}
	l_rep += e - b;
	for (i = 0; i < n_intv; ++i) {
		int slen = (int)((uint32_t)intv[i].info - (uint32_t)(intv[i].info >> 32));
		int64_t s0 = seed_beg[i], s1 = s0 + (int64_t)(intv[i].x[2] < (uint64_t)opt->max_occ ? intv[i].x[2] : (uint64_t)opt->max_occ), j;
		for (j = s0; j < s1; ++j) { /* the device already applied the max_occ subsampling (bwamem.c:304-305) */
			bb_seed_t s;
			int rid, lower;
			s.rbeg = rbeg[j];
			s.qbeg = (int)(intv[i].info >> 32);
			s.score = s.len = slen;
			rid = bb_intv2rid(bns, s.rbeg, s.rbeg + s.len);
			if (rid < 0) continue;
			lower = ws->n_keys ? bt_floor(ws, s.rbeg) : -1;
			if (lower < 0 || !try_merge(ws, opt, l_pac, &ws->chains.a[lower], &s, rid)) {
				bb_chain_t nc;
				memset(&nc, 0, sizeof(nc));
				nc.n = 1; nc.m = 4;
				nc.seeds = arena_alloc(ws, sizeof(bb_seed_t) * 4);
				nc.seeds[0] = s;
				nc.rid = rid;
				nc.pos = s.rbeg;
				nc.is_alt = !!bns->anns[rid].is_alt;
				bb_vec_push(ws->chains, nc);
				bt_insert(ws, (int)ws->chains.n - 1);
			}
		}
	}
	bt_inorder(ws, ws->root);
	bb_vec_reserve(*out, ws->order.n);
	for (k = 0; k < ws->order.n; ++k) {
		out->a[k] = ws->chains.a[ws->order.a[k]];
		out->a[k].frac_rep = (float)l_rep / l_query;
	}
	out->n = ws->order.n;
}

/* min(query bases, reference bases) covered by the seeds of a chain (bwamem.c:239-258) */
int bb_chain_weight(const bb_chain_t *c)
{
	int64_t end;
	int j, wq = 0, wr = 0;
	for (j = 0, end = 0; j < c->n; ++j) {
		const bb_seed_t *s = &c->seeds[j];
		if (s->qbeg >= end) wq += s->len;
		else if (s->qbeg + s->len > end) wq += (int)(s->qbeg + s->len - end);
		if (s->qbeg + s->len > end) end = s->qbeg + s->len;
	}
	for (j = 0, end = 0; j < c->n; ++j) {
		const bb_seed_t *s = &c->seeds[j];
		if (s->rbeg >= end) wr += s->len;
		else if (s->rbeg + s->len > end) wr += (int)(s->rbeg + s->len - end);
		if (s->rbeg + s->len > end) end = s->rbeg + s->len;
	}
	if (wr < wq) wq = wr;
	return wq < 1 << 30 ? wq : (1 << 30) - 1;
}

#define chain_heavier(a, b) ((a).w > (b).w)
BB_SORT_DEFINE(static, sort_chains_by_weight, bb_chain_t, chain_heavier)

#define Q_BEG(ch) ((ch).seeds[0].qbeg)
#define Q_END(ch) ((ch).seeds[(ch).n - 1].qbeg + (ch).seeds[(ch).n - 1].len)

/* bwamem.c:353-411; returns the number of chains kept (compacted to the front) */
int bb_chain_filter(const mem_opt_t *opt, int n, bb_chain_t *a)
{
	int i, k, n_kept = 0, *kept_idx;
	if (n == 0) return 0;
	for (i = k = 0; i < n; ++i) {
		bb_chain_t *c = &a[i];
		c->first = -1; c->kept = 0;
		c->w = (uint32_t)bb_chain_weight(c) & 0x1fffffffu; /* 29-bit field in the reference (bwamem.c:202) */
		if ((int)c->w >= opt->min_chain_weight) a[k++] = *c;
	}
	n = k;
	if (n == 0) return 0; /* (the reference would touch a[0] of an empty array here) */
	sort_chains_by_weight(n, a);
	kept_idx = bb_malloc(sizeof(int) * n);
	a[0].kept = 3;
	kept_idx[n_kept++] = 0;
	for (i = 1; i < n; ++i) {
		int large_ovlp = 0;
		for (k = 0; k < n_kept; ++k) {
			int j = kept_idx[k];
			int b_max = Q_BEG(a[j]) > Q_BEG(a[i]) ? Q_BEG(a[j]) : Q_BEG(a[i]);
			int e_min = Q_END(a[j]) < Q_END(a[i]) ? Q_END(a[j]) : Q_END(a[i]);
			if (e_min > b_max && (!a[j].is_alt || a[i].is_alt)) {
				int li = Q_END(a[i]) - Q_BEG(a[i]), lj = Q_END(a[j]) - Q_BEG(a[j]);
				int min_l = li < lj ? li : lj;
				if (e_min - b_max >= min_l * opt->mask_level && min_l < opt->max_chain_gap) {
					large_ovlp = 1;
					if (a[j].first < 0) a[j].first = i;
					if (a[i].w < a[j].w * opt->drop_ratio && (int)(a[j].w - a[i].w) >= opt->min_seed_len << 1) break;
				}
			}
		}
		if (k == n_kept) {
			kept_idx[n_kept++] = i;
			a[i].kept = large_ovlp ? 2 : 3;
		}
	}
	for (i = 0; i < n_kept; ++i) {
		bb_chain_t *c = &a[kept_idx[i]];
		if (c->first >= 0) a[c->first].kept = 1;
	}
	free(kept_idx);
	for (i = k = 0; i < n; ++i) {
		if (a[i].kept == 0 || a[i].kept == 3) continue;
		if (++k >= opt->max_chain_extend) break;
	}
	for (; i < n; ++i)
		if (a[i].kept < 3) a[i].kept = 0;
	for (i = k = 0; i < n; ++i)
		if (a[i].kept) a[k++] = a[i];
	return k;
}

int bb_cal_max_gap(const mem_opt_t *opt, int qlen) /* bwamem.c:647-654 */
{
	int l_del = (int)((double)(qlen * opt->a - opt->o_del) / opt->e_del + 1.);
	int l_ins = (int)((double)(qlen * opt->a - opt->o_ins) / opt->e_ins + 1.);
	int l = l_del > l_ins ? l_del : l_ins;
	if (l < 1) l = 1;
	return l < opt->w << 1 ? l : opt->w << 1;
}

/* ---------------------------------------------------------------- seed filter for long reads (bwamem.c:590-641) */
#define SHORT_EXT 50
#define SHORT_LEN 200

static int seed_sw_score(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, int l_query, const uint8_t *query, const bb_seed_t *s)
{
	int qb, qe, rid;
	int64_t rb, re, mid, l_pac = bns->l_pac;
	uint8_t *rseq;
	bb_swr_t x;
	if (s->len >= SHORT_LEN) return -1;
	qb = s->qbeg; qe = s->qbeg + s->len;
	rb = s->rbeg; re = s->rbeg + s->len;
	mid = (rb + re) >> 1;
	qb -= SHORT_EXT; if (qb < 0) qb = 0;
	qe += SHORT_EXT; if (qe > l_query) qe = l_query;
	rb -= SHORT_EXT; if (rb < 0) rb = 0;
	re += SHORT_EXT; if (re > l_pac << 1) re = l_pac << 1;
	if (rb < l_pac && l_pac < re) { if (mid < l_pac) re = l_pac; else rb = l_pac; }
	if (qe - qb >= SHORT_LEN || re - rb >= SHORT_LEN) return -1;
	rseq = bb_fetch_seq(bns, pac, &rb, mid, &re, &rid);
	x = bb_local_sw(qe - qb, (uint8_t *)query + qb, (int)(re - rb), rseq, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, BB_SW_XSTART);
	free(rseq);
	return x.score;
}

void bb_chain_seed_sw(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, int l_query, const uint8_t *query, int n, bb_chain_t *a)
{
	double min_l = opt->min_chain_weight ? 1.1f * opt->min_chain_weight : 5.5f * log(l_query);
	int i, j, k, min_hsp = (int)(opt->a * min_l + .499);
	if (min_l > 0.05f * l_query) return;
	for (i = 0; i < n; ++i) {
		bb_chain_t *c = &a[i];
		for (j = k = 0; j < c->n; ++j) {
			bb_seed_t *s = &c->seeds[j];
			s->score = seed_sw_score(opt, bns, pac, l_query, query, s);
			if (s->score < 0 || s->score >= min_hsp) {
				if (s->score < 0) s->score = s->len * opt->a;
				c->seeds[k++] = *s;
			}
		}
		c->n = k;
	}
}
