/* TEST INFRASTRUCTURE ONLY -- never linked into the product (libbwa_b200.so / bwa-b200).
 *
 * oracle_fm.c -- sequential CPU restatement of the FM-index half of the BWA-MEM hot path:
 * rank (Occ) queries, bidirectional interval extension, SMEM enumeration, the three-pass seeding
 * driver and suffix-array lookup.  Each function cites the reference lines it follows.  Pinned
 * against the real reference by tests/test_oracle_pin.py (oracle/_ref/katdump smem|sa|chain).
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* 64-byte block covering BWT row k: 4 x u64 counts, then 8 x u32 of 16 symbols, MSB first (bwt.h:74-82) */
static inline const uint32_t *blk(const bwt_t *b, uint64_t k) { return b->bwt + ((k >> 7) << 4); }

static inline int popc2(uint32_t w, int c) /* how many of the 16 2-bit symbols in w equal c */
{
	uint32_t x = w ^ (0x55555555u * (uint32_t)(3 - c)); /* lanes equal to c become 11 */
	x = x & (x >> 1) & 0x55555555u;
	return __builtin_popcount(x);
}

/* Occ(c, k) for all four c, k inclusive; k == -1 -> zeros (bwt.c:169-186) */
void orc_occ4(const bwt_t *b, uint64_t k, uint64_t cnt[4])
{
	const uint32_t *p;
	int c, j, nw, rem;
	if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	k -= (k >= b->primary); /* '$' is not stored */
	p = blk(b, k);
	memcpy(cnt, p, 32);
	p += 8;
	nw = (int)((k & 127) >> 4); /* full words before the one holding k */
	rem = (int)(k & 15);        /* symbols 0..rem of the last word count */
	for (c = 0; c < 4; ++c) {
		uint64_t n = 0;
		for (j = 0; j < nw; ++j) n += popc2(p[j], c);
		n += popc2(p[nw] >> ((15 - rem) << 1), c);        /* keep the rem+1 leading symbols */
		if (c == 0) n -= 15 - rem;                         /* the shifted-in zeros look like 'A' */
		cnt[c] += n;
	}
}

/* single-symbol rank used by the LF walk (bwt.c:107-129) */
uint64_t orc_occ(const bwt_t *b, uint64_t k, int c)
{
	uint64_t cnt[4];
	if (k == b->seq_len) return b->L2[c + 1] - b->L2[c];
	if (k == (uint64_t)-1) return 0;
	orc_occ4(b, k, cnt);
	return cnt[c];
}

/* bwt.c:262-275; touches counts how many 64-byte blocks a real implementation reads (1 if both ranks
 * fall into the same block, else 2; bwt.c:194-197) */
void orc_extend(const bwt_t *b, const bwtintv_t *ik, bwtintv_t ok[4], int is_back, uint64_t *touches)
{
	uint64_t tk[4], tl[4], k = ik->x[!is_back] - 1, l = ik->x[!is_back] - 1 + ik->x[2];
	int i;
	orc_occ4(b, k, tk);
	orc_occ4(b, l, tl);
	if (touches) {
		uint64_t k2 = k - (k >= b->primary), l2 = l - (l >= b->primary);
		*touches += (k == (uint64_t)-1 || l == (uint64_t)-1 || (k2 >> 7) != (l2 >> 7)) ? 2 : 1;
	}
	for (i = 0; i < 4; ++i) {
		ok[i].x[!is_back] = b->L2[i] + 1 + tk[i];
		ok[i].x[2] = tl[i] - tk[i];
	}
	ok[3].x[is_back] = ik->x[is_back] + (ik->x[!is_back] <= b->primary && ik->x[!is_back] + ik->x[2] - 1 >= b->primary);
	ok[2].x[is_back] = ok[3].x[is_back] + ok[3].x[2];
	ok[1].x[is_back] = ok[2].x[is_back] + ok[2].x[2];
	ok[0].x[is_back] = ok[1].x[is_back] + ok[1].x[2];
}

static void set_intv(const bwt_t *b, int c, bwtintv_t *ik) /* bwt.h:82 */
{
	ik->x[0] = b->L2[c] + 1; ik->x[2] = b->L2[c + 1] - b->L2[c]; ik->x[1] = b->L2[3 - c] + 1; ik->info = 0;
}

static void vpush(orc_intv_v *v, const bwtintv_t *x)
{
	if (v->n == v->m) { v->m = v->m ? v->m << 1 : 16; v->a = realloc(v->a, v->m * sizeof(bwtintv_t)); }
	v->a[v->n++] = *x;
}
static void vrev(orc_intv_v *v)
{
	size_t i;
	for (i = 0; i < v->n >> 1; ++i) { bwtintv_t t = v->a[i]; v->a[i] = v->a[v->n - 1 - i]; v->a[v->n - 1 - i] = t; }
}

/* all SMEMs through query position x with interval size >= min_intv; returns the end of the longest
 * match starting at x (bwt.c:289-351, max_intv fixed to 0 as in bwt_smem1) */
int orc_smem1(const bwt_t *b, int len, const uint8_t *q, int x, int min_intv, orc_intv_v *mem, uint64_t *touches)
{
	orc_intv_v cur = {0, 0, 0}, prev = {0, 0, 0}, sw;
	bwtintv_t ik, ok[4];
	int i, ret;
	size_t j;
	mem->n = 0;
	if (q[x] > 3) return x + 1;
	if (min_intv < 1) min_intv = 1;
	set_intv(b, q[x], &ik);
	ik.info = (uint64_t)x + 1;
	for (i = x + 1; i < len; ++i) { /* forward: remember an interval whenever its size is about to change */
		if (q[i] < 4) {
			int c = 3 - q[i];
			orc_extend(b, &ik, ok, 0, touches);
			if (ok[c].x[2] != ik.x[2]) {
				vpush(&cur, &ik);
				if (ok[c].x[2] < (uint64_t)min_intv) break;
			}
			ik = ok[c]; ik.info = (uint64_t)i + 1;
		} else { vpush(&cur, &ik); break; }
	}
	if (i == len) vpush(&cur, &ik);
	vrev(&cur);
	ret = (int)cur.a[0].info;
	sw = cur; cur = prev; prev = sw;
	for (i = x - 1; i >= -1; --i) { /* backward: extend every candidate; the first that dies is an SMEM */
		int c = i < 0 ? -1 : q[i] < 4 ? q[i] : -1;
		cur.n = 0;
		for (j = 0; j < prev.n; ++j) {
			bwtintv_t *p = &prev.a[j];
			if (c >= 0) orc_extend(b, p, ok, 1, touches);
			if (c < 0 || ok[c].x[2] < (uint64_t)min_intv) {
				if (cur.n == 0 && (mem->n == 0 || (uint64_t)(i + 1) < mem->a[mem->n - 1].info >> 32)) {
					ik = *p; ik.info |= (uint64_t)(i + 1) << 32;
					vpush(mem, &ik);
				}
			} else if (cur.n == 0 || ok[c].x[2] != cur.a[cur.n - 1].x[2]) {
				ok[c].info = p->info;
				vpush(&cur, &ok[c]);
			}
		}
		if (cur.n == 0) break;
		sw = cur; cur = prev; prev = sw;
	}
	vrev(mem);
	free(cur.a); free(prev.a);
	return ret;
}

/* forward-only seed: first point with fewer than max_intv occurrences and length > min_len (bwt.c:358-379) */
int orc_seed_strategy1(const bwt_t *b, int len, const uint8_t *q, int x, int min_len, int max_intv, bwtintv_t *mem, uint64_t *touches)
{
	bwtintv_t ik, ok[4];
	int i;
	memset(mem, 0, sizeof(*mem));
	if (q[x] > 3) return x + 1;
	set_intv(b, q[x], &ik);
	for (i = x + 1; i < len; ++i) {
		int c;
		if (q[i] > 3) return i + 1;
		c = 3 - q[i];
		orc_extend(b, &ik, ok, 0, touches);
		if (ok[c].x[2] < (uint64_t)max_intv && i - x >= min_len) {
			*mem = ok[c];
			mem->info = (uint64_t)x << 32 | (uint64_t)(i + 1);
			return i + 1;
		}
		ik = ok[c];
	}
	return len;
}

static int cmp_info(const void *a, const void *b)
{
	const bwtintv_t *x = a, *y = b;
	return x->info < y->info ? -1 : x->info > y->info ? 1 : 0; /* equal info => identical intervals, any order */
}

/* the three seeding passes + sort by (start,end) (bwamem.c:140-188) */
void orc_collect_intv(const bwt_t *b, int len, const uint8_t *seq, int min_seed_len, int split_len, int split_width, uint64_t max_mem_intv,
                      orc_intv_v *out, uint64_t *touches)
{
	orc_intv_v m1 = {0, 0, 0};
	int x = 0;
	size_t i, k, old_n;
	out->n = 0;
	while (x < len) {
		if (seq[x] < 4) {
			x = orc_smem1(b, len, seq, x, 1, &m1, touches);
			for (i = 0; i < m1.n; ++i)
				if ((int)((uint32_t)m1.a[i].info - (uint32_t)(m1.a[i].info >> 32)) >= min_seed_len) vpush(out, &m1.a[i]);
		} else ++x;
	}
	old_n = out->n;
	for (k = 0; k < old_n; ++k) {
		bwtintv_t p = out->a[k];
		int start = (int)(p.info >> 32), end = (int)(uint32_t)p.info;
		if (end - start < split_len || p.x[2] > (uint64_t)split_width) continue;
		orc_smem1(b, len, seq, (start + end) >> 1, (int)p.x[2] + 1, &m1, touches);
		for (i = 0; i < m1.n; ++i)
			if ((int)((uint32_t)m1.a[i].info - (uint32_t)(m1.a[i].info >> 32)) >= min_seed_len) vpush(out, &m1.a[i]);
	}
	if (max_mem_intv > 0) {
		x = 0;
		while (x < len) {
			if (seq[x] < 4) {
				bwtintv_t m;
				x = orc_seed_strategy1(b, len, seq, x, min_seed_len, (int)max_mem_intv, &m, touches);
				if (m.x[2] > 0) vpush(out, &m);
			} else ++x;
		}
	}
	qsort(out->a, out->n, sizeof(bwtintv_t), cmp_info);
	free(m1.a);
}

/* text position of BWT row k: LF-walk to the nearest sampled row (bwt.c:53-59, 86-96) */
uint64_t orc_sa(const bwt_t *b, uint64_t k, uint64_t *steps)
{
	uint64_t sa = 0, mask = (uint64_t)b->sa_intv - 1;
	while (k & mask) {
		uint64_t x;
		int c;
		++sa;
		if (steps) ++*steps;
		if (k == b->primary) { k = 0; continue; }
		x = k - (k > b->primary);
		c = blk(b, x)[8 + ((x & 127) >> 4)] >> ((~x & 15) << 1) & 3;
		k = b->L2[c] + orc_occ(b, k, c);
	}
	return sa + b->sa[k / b->sa_intv];
}
