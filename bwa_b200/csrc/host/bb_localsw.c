/* bb_localsw.c -- local Smith-Waterman with start recovery, used off the main path by mate rescue
 * (bwamem_pair.c:137-206) and the long-read seed filter (bwamem.c:597-622).
 *
 * The reference computes this with Farrar's striped SIMD kernels (ksw.c:122-370), whose result
 * (score, end points, second-best score) depends on details of the striping: query positions are
 * padded to a multiple of the lane count with zero-scoring columns, F is propagated lazily, E is
 * derived before the lazy pass, and the 8-bit kernel works on biased saturating bytes.  To return
 * the same numbers in every corner this file evaluates the same recurrence lane by lane in plain
 * integer arithmetic ("one vector" = an array of P lanes, lane l of stripe j = query position
 * j + l*slen).  SURVEY.md section 8(f) lists a batched CUDA version of this routine as the next
 * component after the seed/extend/global path; until then it is host code.
 */
#include "bb_host.h"

typedef struct {
	int P, slen, qlen, is8;
	int shift, maxsc;
	int *prof;   /* [5][slen][P] */
} profile_t;

static int sat_u8(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
static int sat_i16(int v) { return v < -32768 ? -32768 : v > 32767 ? 32767 : v; }

static void profile_init(profile_t *q, int is8, int qlen, const uint8_t *query, const int8_t *mat)
{
	int a, i, l, mn = 127, mx = 0;
	q->is8 = is8; q->P = is8 ? 16 : 8; q->qlen = qlen;
	q->slen = (qlen + q->P - 1) / q->P;
	for (a = 0; a < 25; ++a) { if (mat[a] < mn) mn = mat[a]; if (mat[a] > mx) mx = mat[a]; }
	q->maxsc = mx;
	q->shift = (256 - (mn & 0xff)) & 0xff; /* bias that makes every matrix entry non-negative */
	q->prof = bb_malloc(sizeof(int) * 5 * q->slen * q->P);
	for (a = 0; a < 5; ++a)
		for (i = 0; i < q->slen; ++i)
			for (l = 0; l < q->P; ++l) {
				int k = i + l * q->slen;
				int v = k >= qlen ? 0 : mat[a * 5 + query[k]];
				q->prof[(a * q->slen + i) * q->P + l] = is8 ? (v + q->shift) & 0xff : v;
			}
}

static bb_swr_t striped_sw(const profile_t *q, int tlen, const uint8_t *target, int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
	const int P = q->P, slen = q->slen, is8 = q->is8;
	const int minsc = (xtra & BB_SW_XSUBO) ? xtra & 0xffff : 0x10000;
	const int endsc = (xtra & BB_SW_XSTOP) ? xtra & 0xffff : 0x10000;
	const int oe_del = is8 ? (o_del + e_del) & 0xff : (o_del + e_del) & 0xffff;
	const int oe_ins = is8 ? (o_ins + e_ins) & 0xff : (o_ins + e_ins) & 0xffff;
	int *H0 = bb_calloc((size_t)slen * P, sizeof(int)), *H1 = bb_calloc((size_t)slen * P, sizeof(int));
	int *E = bb_calloc((size_t)slen * P, sizeof(int)), *Hmax = bb_calloc((size_t)slen * P, sizeof(int));
	int h[16], f[16], mx[16];
	BB_VEC(uint64_t) b = {0, 0, 0};
	int i, j, k, l, te = -1, gmax = 0;
	bb_swr_t r = {0, -1, -1, -1, -1, -1, -1};

	for (i = 0; i < tlen; ++i) {
		const int *S = q->prof + (size_t)target[i] * slen * P;
		int imax = 0, done = 0;
		for (l = P - 1; l > 0; --l) h[l] = H0[(slen - 1) * P + l - 1]; /* previous row, shifted by one lane */
		h[0] = 0;
		for (l = 0; l < P; ++l) f[l] = mx[l] = 0;
		for (j = 0; j < slen; ++j) {
			for (l = 0; l < P; ++l) {
				int hv, e = E[j * P + l], t;
				if (is8) { hv = sat_u8(h[l] + S[j * P + l]); hv = sat_u8(hv - q->shift); }
				else hv = sat_i16(h[l] + S[j * P + l]);
				if (e > hv) hv = e;
				if (f[l] > hv) hv = f[l];
				if (hv > mx[l]) mx[l] = hv;
				H1[j * P + l] = hv;
				e -= e_del; if (e < 0) e = 0;
				t = hv - oe_del; if (t < 0) t = 0;
				E[j * P + l] = e > t ? e : t;
				f[l] -= e_ins; if (f[l] < 0) f[l] = 0;
				t = hv - oe_ins; if (t < 0) t = 0;
				if (t > f[l]) f[l] = t;
				h[l] = H0[j * P + l];
			}
		}
		for (k = 0; k < 16 && !done; ++k) { /* lazy F: at most 16 sweeps in both kernels */
			for (l = P - 1; l > 0; --l) f[l] = f[l - 1];
			f[0] = 0;
			for (j = 0; j < slen; ++j) {
				int any = 0;
				for (l = 0; l < P; ++l) {
					int hv = H1[j * P + l];
					if (f[l] > hv) hv = f[l];
					H1[j * P + l] = hv;
					hv -= oe_ins; if (hv < 0) hv = 0;
					f[l] -= e_ins; if (f[l] < 0) f[l] = 0;
					if (f[l] > hv) any = 1;
				}
				if (!any) { done = 1; break; }
			}
		}
		for (l = 0; l < P; ++l) if (mx[l] > imax) imax = mx[l];
		if (imax >= minsc) {
			if (b.n == 0 || (int32_t)b.a[b.n - 1] + 1 != i) bb_vec_push(b, (uint64_t)imax << 32 | (uint32_t)i);
			else if ((int)(b.a[b.n - 1] >> 32) < imax) b.a[b.n - 1] = (uint64_t)imax << 32 | (uint32_t)i;
		}
		if (imax > gmax) {
			gmax = imax; te = i;
			memcpy(Hmax, H1, sizeof(int) * slen * P);
			if ((is8 && gmax + q->shift >= 255) || gmax >= endsc) break;
		}
		{ int *t = H0; H0 = H1; H1 = t; }
	}
	r.score = is8 ? (gmax + q->shift < 255 ? gmax : 255) : gmax;
	r.te = te;
	if (!is8 || r.score != 255) {
		int best = -1, n = slen * P;
		if (!is8) r.qe = -1;
		for (i = 0; i < n; ++i) {
			int v = Hmax[i], pos = i / P + i % P * slen;
			if (v > best) { best = v; r.qe = pos; }
			else if (v == best && pos < r.qe) r.qe = pos;
		}
		if (b.a) {
			int low, high;
			size_t x;
			i = (r.score + q->maxsc - 1) / q->maxsc;
			low = te - i; high = te + i;
			for (x = 0; x < b.n; ++x) {
				int e = (int32_t)b.a[x];
				if ((e < low || e > high) && (int)(b.a[x] >> 32) > r.score2) { r.score2 = (int)(b.a[x] >> 32); r.te2 = e; }
			}
		}
	}
	free(b.a); free(H0); free(H1); free(E); free(Hmax);
	return r;
}

static void reverse_bytes(int l, uint8_t *s)
{
	int i;
	for (i = 0; i < l >> 1; ++i) { uint8_t t = s[i]; s[i] = s[l - 1 - i]; s[l - 1 - i] = t; }
}

/* same contract as ksw_align2 (ksw.c:379-401) with m=5 and no cached profile */
bb_swr_t bb_local_sw(int qlen, uint8_t *query, int tlen, uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
	profile_t q;
	bb_swr_t r, rr;
	int is8 = (xtra & BB_SW_XBYTE) ? 1 : 0;
	profile_init(&q, is8, qlen, query, mat);
	r = striped_sw(&q, tlen, target, o_del, e_del, o_ins, e_ins, xtra);
	free(q.prof);
	if ((xtra & BB_SW_XSTART) == 0 || ((xtra & BB_SW_XSUBO) && r.score < (xtra & 0xffff))) return r;
	reverse_bytes(r.qe + 1, query); reverse_bytes(r.te + 1, target);
	profile_init(&q, is8, r.qe + 1, query, mat);
	rr = striped_sw(&q, tlen, target, o_del, e_del, o_ins, e_ins, BB_SW_XSTOP | r.score);
	free(q.prof);
	reverse_bytes(r.qe + 1, query); reverse_bytes(r.te + 1, target);
	if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
	return r;
}
