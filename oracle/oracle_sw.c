/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * oracle_sw.c -- sequential CPU restatement of the dynamic-programming half of the BWA-MEM hot path:
 * banded extension with adaptive band and Z-drop, banded global alignment with backtrack, CIGAR/NM/MD
 * generation, the band-doubling loop of mem_reg2aln, and the per-chain extension driver.  Pinned
 * against the real reference by tests/test_oracle_pin.py (oracle/_ref/katdump extend|global|regs|aln).
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "oracle.h"

#define NEG_INF (-0x40000000)

/* ksw.c:416-515.  Row i = target base, columns = query; eh[j] carries H(i-1,j-1) and E(i,j). */
int orc_extend_sw(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
               int w, int end_bonus, int zdrop, int h0, int *qle, int *tle, int *gtle, int *gscore_, int *max_off_, uint64_t *cells)
{
	int *H = calloc(qlen + 1, sizeof(int)), *E = calloc(qlen + 1, sizeof(int));
	int i, j, k, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, beg, end, max, max_i, max_j, max_ins, max_del, max_ie, gscore, max_off;
	H[0] = h0; H[1] = h0 > oe_ins ? h0 - oe_ins : 0;
	for (j = 2; j <= qlen && H[j - 1] > e_ins; ++j) H[j] = H[j - 1] - e_ins;
	for (i = 0, max = 0; i < 25; ++i) max = max > mat[i] ? max : mat[i];
	max_ins = (int)((double)(qlen * max + end_bonus - o_ins) / e_ins + 1.); if (max_ins < 1) max_ins = 1;
	if (w > max_ins) w = max_ins;
	max_del = (int)((double)(qlen * max + end_bonus - o_del) / e_del + 1.); if (max_del < 1) max_del = 1;
	if (w > max_del) w = max_del;
	max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
	beg = 0; end = qlen;
	for (i = 0; i < tlen; ++i) {
		int f = 0, h1, m = 0, mj = -1;
		const int8_t *row = mat + target[i] * 5;
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; } else h1 = 0;
		if (cells && end > beg) *cells += (uint64_t)(end - beg);
		for (j = beg; j < end; ++j) {
			int M = H[j], e = E[j], h, t;
			H[j] = h1;
			M = M ? M + row[query[j]] : 0;
			h = M > e ? M : e; h = h > f ? h : f;
			h1 = h;
			mj = m > h ? mj : j;       /* ties go to the larger column */
			m = m > h ? m : h;
			t = M - oe_del; t = t > 0 ? t : 0; e -= e_del; e = e > t ? e : t; E[j] = e;
			t = M - oe_ins; t = t > 0 ? t : 0; f -= e_ins; f = f > t ? f : t;
		}
		H[end] = h1; E[end] = 0;
		if (j == qlen) { max_ie = gscore > h1 ? max_ie : i; gscore = gscore > h1 ? gscore : h1; }
		if (m == 0) break;
		if (m > max) {
			max = m; max_i = i; max_j = mj;
			k = mj - i; if (k < 0) k = -k;
			if (k > max_off) max_off = k;
		} else if (zdrop > 0) {
			if (i - max_i > mj - max_j) { if (max - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break; }
			else { if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break; }
		}
		for (j = beg; j < end && H[j] == 0 && E[j] == 0; ++j) {}
		beg = j;
		for (j = end; j >= beg && H[j] == 0 && E[j] == 0; --j) {}
		end = j + 2 < qlen ? j + 2 : qlen;
	}
	free(H); free(E);
	if (qle) *qle = max_j + 1;
	if (tle) *tle = max_i + 1;
	if (gtle) *gtle = max_ie + 1;
	if (gscore_) *gscore_ = gscore;
	if (max_off_) *max_off_ = max_off;
	return max;
}

static void cigar_push(orc_u32_v *c, int op, int len)
{
	if (c->n && (c->a[c->n - 1] & 0xf) == (uint32_t)op) { c->a[c->n - 1] += (uint32_t)len << 4; return; }
	if (c->n == c->m) { c->m = c->m ? c->m << 1 : 8; c->a = realloc(c->a, c->m * 4); }
	c->a[c->n++] = (uint32_t)len << 4 | (uint32_t)op;
}

/* ksw.c:540-642.  cigar == NULL -> score only. */
int orc_global(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
               int w, orc_u32_v *cigar, uint64_t *cells)
{
	int *H = malloc(sizeof(int) * (qlen + 1)), *E = malloc(sizeof(int) * (qlen + 1));
	int i, j, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, score, n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
	uint8_t *z = cigar ? malloc((size_t)n_col * tlen + 1) : 0;
	if (cigar) cigar->n = 0;
	H[0] = 0; E[0] = NEG_INF;
	for (j = 1; j <= qlen && j <= w; ++j) { H[j] = -(o_ins + e_ins * j); E[j] = NEG_INF; }
	for (; j <= qlen; ++j) H[j] = E[j] = NEG_INF;
	for (i = 0; i < tlen; ++i) {
		int f = NEG_INF, h1, beg = i > w ? i - w : 0, end = i + w + 1 < qlen ? i + w + 1 : qlen;
		const int8_t *row = mat + target[i] * 5;
		uint8_t *zi = z ? z + (size_t)i * n_col : 0;
		h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : NEG_INF;
		if (cells && end > beg) *cells += (uint64_t)(end - beg);
		for (j = beg; j < end; ++j) {
			int m = H[j] + row[query[j]], e = E[j], h, t;
			uint8_t d;
			H[j] = h1;
			d = m >= e ? 0 : 1; h = m >= e ? m : e;
			d = h >= f ? d : 2; h = h >= f ? h : f;
			h1 = h;
			t = m - oe_del; e -= e_del; d |= e > t ? 1 << 2 : 0; e = e > t ? e : t; E[j] = e;
			t = m - oe_ins; f -= e_ins; d |= f > t ? 2 << 4 : 0; f = f > t ? f : t;
			if (zi) zi[j - beg] = d;
		}
		H[end] = h1; E[end] = NEG_INF;
	}
	score = H[qlen];
	if (cigar) {
		int which = 0, k;
		size_t a, n;
		i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
		while (i >= 0 && k >= 0) {
			which = z[(size_t)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
			if (which == 0) { cigar_push(cigar, 0, 1); --i; --k; }
			else if (which == 1) { cigar_push(cigar, 2, 1); --i; }
			else { cigar_push(cigar, 1, 1); --k; }
		}
		if (i >= 0) cigar_push(cigar, 2, i + 1);
		if (k >= 0) cigar_push(cigar, 1, k + 1);
		for (a = 0, n = cigar->n; a < n >> 1; ++a) { uint32_t t = cigar->a[a]; cigar->a[a] = cigar->a[n - 1 - a]; cigar->a[n - 1 - a] = t; }
	}
	free(H); free(E); free(z);
	return score;
}

static inline int pac_base(const uint8_t *pac, int64_t k) { return pac[k >> 2] >> ((~k & 3) << 1) & 3; }

/* bases [beg,end) of the doubled coordinate system (bntseq.c:403-424); 0 if the range straddles l_pac */
uint8_t *orc_get_seq(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, int64_t *len)
{
	uint8_t *s;
	int64_t k, l = 0;
	if (end < beg) { int64_t t = beg; beg = end; end = t; }
	if (end > l_pac << 1) end = l_pac << 1;
	if (beg < 0) beg = 0;
	*len = 0;
	if (!(beg >= l_pac || end <= l_pac)) return 0;
	*len = end - beg;
	s = malloc(end - beg + 1);
	if (beg >= l_pac) for (k = (l_pac << 1) - 1 - beg; k > (l_pac << 1) - 1 - end; --k) s[l++] = 3 - pac_base(pac, k);
	else for (k = beg; k < end; ++k) s[l++] = pac_base(pac, k);
	return s;
}

static void md_putnum(orc_str_t *s, int v)
{
	char buf[16];
	int n = snprintf(buf, sizeof(buf), "%d", v);
	if (s->l + n + 2 > s->m) { s->m = (s->l + n + 2) * 2; s->s = realloc(s->s, s->m); }
	memcpy(s->s + s->l, buf, n); s->l += n; s->s[s->l] = 0;
}
static void md_putc(orc_str_t *s, int c)
{
	if (s->l + 2 > s->m) { s->m = (s->l + 2) * 2; s->s = realloc(s->s, s->m); }
	s->s[s->l++] = (char)c; s->s[s->l] = 0;
}

/* bwa.c:148-234.  want_cigar==0 -> score only.  Returns 0 on success, -1 if the reference would
 * return no alignment (empty query/ref or a range bridging the strands). */
int orc_gen_cigar(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins, int w_, int64_t l_pac, const uint8_t *pac,
                  int l_query, const uint8_t *query_, int64_t rb, int64_t re, int want_cigar,
                  int *score, orc_u32_v *cigar, int *NM, orc_str_t *md, uint64_t *cells)
{
	uint8_t *rseq, *query;
	int64_t rlen;
	int i;
	if (cigar) cigar->n = 0;
	if (NM) *NM = -1;
	if (md) { md->l = 0; if (md->s) md->s[0] = 0; }
	if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return -1;
	rseq = orc_get_seq(l_pac, pac, rb, re, &rlen);
	if (re - rb != rlen) { free(rseq); return -1; }
	query = malloc(l_query);
	memcpy(query, query_, l_query);
	if (rb >= l_pac) { /* reverse both so that gaps are left-aligned on the forward strand */
		for (i = 0; i < l_query >> 1; ++i) { uint8_t t = query[i]; query[i] = query[l_query - 1 - i]; query[l_query - 1 - i] = t; }
		for (i = 0; i < rlen >> 1; ++i) { uint8_t t = rseq[i]; rseq[i] = rseq[rlen - 1 - i]; rseq[rlen - 1 - i] = t; }
	}
	if (l_query == re - rb && w_ == 0) {
		if (want_cigar) cigar_push(cigar, 0, l_query);
		for (i = 0, *score = 0; i < l_query; ++i) *score += mat[rseq[i] * 5 + query[i]];
	} else {
		int w, max_gap, max_ins, max_del, min_w, d = (int)rlen - l_query;
		if (d < 0) d = -d;
		max_ins = (int)((double)(((l_query + 1) >> 1) * mat[0] - o_ins) / e_ins + 1.);
		max_del = (int)((double)(((l_query + 1) >> 1) * mat[0] - o_del) / e_del + 1.);
		max_gap = max_ins > max_del ? max_ins : max_del;
		max_gap = max_gap > 1 ? max_gap : 1;
		w = (max_gap + d + 1) >> 1;
		w = w < w_ ? w : w_;
		min_w = d + 3;
		w = w > min_w ? w : min_w;
		*score = orc_global(l_query, query, (int)rlen, rseq, mat, o_del, e_del, o_ins, e_ins, w, want_cigar ? cigar : 0, cells);
	}
	if (want_cigar && NM && md) {
		int k, x = 0, y = 0, u = 0, n_mm = 0, n_gap = 0;
		const char *b2c = rb < l_pac ? "ACGTN" : "TGCAN";
		for (k = 0; k < (int)cigar->n; ++k) {
			int op = cigar->a[k] & 0xf, len = cigar->a[k] >> 4;
			if (op == 0) {
				for (i = 0; i < len; ++i) {
					if (query[x + i] != rseq[y + i]) { md_putnum(md, u); md_putc(md, b2c[rseq[y + i]]); ++n_mm; u = 0; }
					else ++u;
				}
				x += len; y += len;
			} else if (op == 2) {
				if (k > 0 && k < (int)cigar->n - 1) {
					md_putnum(md, u); md_putc(md, '^');
					for (i = 0; i < len; ++i) md_putc(md, b2c[rseq[y + i]]);
					u = 0; n_gap += len;
				}
				y += len;
			} else if (op == 1) { x += len; n_gap += len; }
		}
		md_putnum(md, u);
		*NM = n_mm + n_gap;
	}
	free(query); free(rseq);
	return 0;
}

/* the do-while of mem_reg2aln (bwamem.c:1143-1152): start from band w2, double up to 3 times */
int orc_reg2aln_core(const int8_t mat[25], int a, int o_del, int e_del, int o_ins, int e_ins, int opt_w, int64_t l_pac, const uint8_t *pac,
                     int l_query, const uint8_t *query, int64_t rb, int64_t re, int w2, int truesc,
                     int *score, orc_u32_v *cigar, int *NM, orc_str_t *md, uint64_t *cells)
{
	int i = 0, last_sc = -(1 << 30);
	do {
		w2 = w2 < opt_w << 2 ? w2 : opt_w << 2;
		if (orc_gen_cigar(mat, o_del, e_del, o_ins, e_ins, w2, l_pac, pac, l_query, query, rb, re, 1, score, cigar, NM, md, cells) < 0) return -1;
		if (*score == last_sc || w2 == opt_w << 2) break;
		last_sc = *score;
		w2 <<= 1;
	} while (++i < 3 && *score < truesc - a);
	return 0;
}

static int max_gap_of(const orc_swpar_t *p, int qlen) /* bwamem.c:647-654 */
{
	int l_del = (int)((double)(qlen * p->a - p->o_del) / p->e_del + 1.);
	int l_ins = (int)((double)(qlen * p->a - p->o_ins) / p->e_ins + 1.);
	int l = l_del > l_ins ? l_del : l_ins;
	l = l > 1 ? l : 1;
	return l < p->w << 1 ? l : p->w << 1;
}

/* The seed loop of mem_chain2aln (bwamem.c:693-810) for one chain whose reference window [rmax0,rmax1)
 * and seed order (ascending sort key) were prepared by the caller.  Appends to regs[*n_regs..]. */
void orc_chain2aln(const orc_swpar_t *p, int64_t l_pac, const uint8_t *pac, int l_query, const uint8_t *query,
                   int64_t rmax0, int64_t rmax1, int n_seeds, const orc_xseed_t *seeds, int chain_idx,
                   orc_xreg_t *regs, int *n_regs, uint64_t *cells)
{
	int64_t rlen;
	uint8_t *rseq = orc_get_seq(l_pac, pac, rmax0, rmax1, &rlen);
	char *dead = calloc(n_seeds + 1, 1); /* "srt[i] == 0": skipped, or the sort key itself was 0 */
	int k, i;
	for (k = 0; k < n_seeds; ++k) dead[k] = (seeds[k].len & 0x80000000u) != 0;
	for (k = n_seeds - 1; k >= 0; --k) {
		int64_t s_rbeg = seeds[k].rbeg;
		int s_qbeg = seeds[k].qbeg, s_len = (int)(seeds[k].len & 0x7fffffffu), aw[2], max_off[2];
		orc_xreg_t *a;
		for (i = 0; i < *n_regs; ++i) { /* was this seed already covered by an earlier extension of this read? */
			const orc_xreg_t *q = &regs[i];
			int64_t rd;
			int qd, w, mg;
			if (s_rbeg < q->rb || s_rbeg + s_len > q->re || s_qbeg < q->qb || s_qbeg + s_len > q->qe) continue;
			if (s_len - q->seedlen0 > .1 * l_query) continue;
			qd = s_qbeg - q->qb; rd = s_rbeg - q->rb;
			mg = max_gap_of(p, qd < rd ? qd : (int)rd);
			w = mg < q->w ? mg : q->w;
			if (qd - rd < w && rd - qd < w) break;
			qd = q->qe - (s_qbeg + s_len); rd = q->re - (s_rbeg + s_len);
			mg = max_gap_of(p, qd < rd ? qd : (int)rd);
			w = mg < q->w ? mg : q->w;
			if (qd - rd < w && rd - qd < w) break;
		}
		if (i < *n_regs) { /* contained: extend anyway only if an overlapping, already-extended seed disagrees on the diagonal */
			for (i = k + 1; i < n_seeds; ++i) {
				int64_t t_rbeg = seeds[i].rbeg;
				int t_qbeg = seeds[i].qbeg, t_len = (int)(seeds[i].len & 0x7fffffffu);
				if (dead[i]) continue;
				if (t_len < s_len * .95) continue;
				if (s_qbeg <= t_qbeg && s_qbeg + s_len - t_qbeg >= s_len >> 2 && t_qbeg - s_qbeg != t_rbeg - s_rbeg) break;
				if (t_qbeg <= s_qbeg && t_qbeg + t_len - s_qbeg >= s_len >> 2 && s_qbeg - t_qbeg != s_rbeg - t_rbeg) break;
			}
			if (i == n_seeds) { dead[k] = 1; continue; }
		}
		a = &regs[(*n_regs)++];
		memset(a, 0, sizeof(*a));
		a->w = aw[0] = aw[1] = p->w;
		a->score = a->truesc = -1;
		a->chain = chain_idx;
		if (s_qbeg) { /* left extension on reversed sequences */
			int qle, tle, gtle, gscore, tl = (int)(s_rbeg - rmax0);
			uint8_t *qs = malloc(s_qbeg), *rs = malloc(tl + 1);
			for (i = 0; i < s_qbeg; ++i) qs[i] = query[s_qbeg - 1 - i];
			for (i = 0; i < tl; ++i) rs[i] = rseq[tl - 1 - i];
			for (i = 0; i < 2; ++i) {
				int prev = a->score;
				aw[0] = p->w << i;
				a->score = orc_extend_sw(s_qbeg, qs, tl, rs, p->mat, p->o_del, p->e_del, p->o_ins, p->e_ins, aw[0], p->pen_clip5, p->zdrop, s_len * p->a, &qle, &tle, &gtle, &gscore, &max_off[0], cells);
				if (a->score == prev || max_off[0] < (aw[0] >> 1) + (aw[0] >> 2)) break;
			}
			if (gscore <= 0 || gscore <= a->score - p->pen_clip5) { a->qb = s_qbeg - qle; a->rb = s_rbeg - tle; a->truesc = a->score; }
			else { a->qb = 0; a->rb = s_rbeg - gtle; a->truesc = gscore; }
			free(qs); free(rs);
		} else { a->score = a->truesc = s_len * p->a; a->qb = 0; a->rb = s_rbeg; }
		if (s_qbeg + s_len != l_query) { /* right extension */
			int qle, tle, gtle, gscore, sc0 = a->score, qe = s_qbeg + s_len;
			int64_t re = s_rbeg + s_len - rmax0;
			for (i = 0; i < 2; ++i) {
				int prev = a->score;
				aw[1] = p->w << i;
				a->score = orc_extend_sw(l_query - qe, query + qe, (int)(rmax1 - rmax0 - re), rseq + re, p->mat, p->o_del, p->e_del, p->o_ins, p->e_ins, aw[1], p->pen_clip3, p->zdrop, sc0, &qle, &tle, &gtle, &gscore, &max_off[1], cells);
				if (a->score == prev || max_off[1] < (aw[1] >> 1) + (aw[1] >> 2)) break;
			}
			if (gscore <= 0 || gscore <= a->score - p->pen_clip3) { a->qe = qe + qle; a->re = rmax0 + re + tle; a->truesc += a->score - sc0; }
			else { a->qe = l_query; a->re = rmax0 + re + gtle; a->truesc += gscore - sc0; }
		} else { a->qe = l_query; a->re = s_rbeg + s_len; }
		for (i = 0, a->seedcov = 0; i < n_seeds; ++i) {
			int t_len = (int)(seeds[i].len & 0x7fffffffu);
			if (seeds[i].qbeg >= a->qb && seeds[i].qbeg + t_len <= a->qe && seeds[i].rbeg >= a->rb && seeds[i].rbeg + t_len <= a->re) a->seedcov += t_len;
		}
		a->w = aw[0] > aw[1] ? aw[0] : aw[1];
		a->seedlen0 = s_len;
	}
	free(dead); free(rseq);
}
