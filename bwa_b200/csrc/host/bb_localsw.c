/* bb_localsw.c -- local Smith-Waterman with start recovery, used off the main path by mate rescue
 * (bwamem_pair.c:137-206) and the long-read seed filter (bwamem.c:597-622).
 *
 * The reference computes this with Farrar's striped SIMD kernels (ksw.c:122-370), whose result
 * (score, end points, second-best score) depends on details of the striping: query positions are
 * padded to a multiple of the lane count with zero-scoring columns, F is propagated lazily, E is
 * derived before the lazy pass, and the 8-bit kernel works on biased saturating bytes.  To return
 * the same numbers in every corner this file evaluates the same striped recurrence ("one vector" = P lanes,
 * lane l of stripe j = query position j + l*slen): with SSE2 vectors where the compiler targets x86
 * (16 unsigned bytes or 8 signed words per vector), and lane by lane in plain integer arithmetic elsewhere
 * -- the scalar version is also the executable specification the vector version is tested against
 * (tests/test_oracle_pin.py: known-answer vectors of the reference, and BWA_B200_SCALAR_SW=1).
 * SURVEY.md section 8(f) lists a batched CUDA version of this routine as the next component after the
 * seed/extend/global path; until then it is host code.
 */
#include "bb_host.h"
#if defined(__SSE2__)
#include <emmintrin.h>
#define BB_HAVE_SSE2 1
#endif

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

static bb_swr_t striped_sw_scalar(const profile_t *q, int tlen, const uint8_t *target, int o_del, int e_del, int o_ins, int e_ins, int xtra)
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


#ifdef BB_HAVE_SSE2
/* ---- the same two kernels on SSE2 vectors ---- */
typedef struct { int is8, P, slen, qlen, shift, maxsc; void *mem; __m128i *prof, *H0, *H1, *E, *Hmax; } vprofile_t;

static void vprofile_init(vprofile_t *q, int is8, int qlen, const uint8_t *query, const int8_t *mat)
{
	int a, i, l, mn = 127, mx = 0;
	size_t n_vec;
	q->is8 = is8; q->P = is8 ? 16 : 8; q->qlen = qlen;
	q->slen = (qlen + q->P - 1) / q->P;
	for (a = 0; a < 25; ++a) { if (mat[a] < mn) mn = mat[a]; if (mat[a] > mx) mx = mat[a]; }
	q->maxsc = mx;
	q->shift = (256 - (mn & 0xff)) & 0xff;
	n_vec = (size_t)q->slen * 9;                     /* 5 profile rows + H0, H1, E, Hmax */
	q->mem = bb_malloc(n_vec * 16 + 16);
	q->prof = (__m128i *)(((uintptr_t)q->mem + 15) & ~(uintptr_t)15);
	q->H0 = q->prof + 5 * (size_t)q->slen; q->H1 = q->H0 + q->slen; q->E = q->H1 + q->slen; q->Hmax = q->E + q->slen;
	for (a = 0; a < 5; ++a)
		for (i = 0; i < q->slen; ++i) {
			if (is8) {
				uint8_t v[16];
				for (l = 0; l < 16; ++l) { int k = i + l * q->slen; v[l] = (uint8_t)((k >= qlen ? 0 : mat[a * 5 + query[k]]) + q->shift); }
				q->prof[a * q->slen + i] = _mm_loadu_si128((const __m128i *)v);
			} else {
				int16_t v[8];
				for (l = 0; l < 8; ++l) { int k = i + l * q->slen; v[l] = (int16_t)(k >= qlen ? 0 : mat[a * 5 + query[k]]); }
				q->prof[a * q->slen + i] = _mm_loadu_si128((const __m128i *)v);
			}
		}
}

static inline int hmax_u8(__m128i v)
{
	v = _mm_max_epu8(v, _mm_srli_si128(v, 8)); v = _mm_max_epu8(v, _mm_srli_si128(v, 4));
	v = _mm_max_epu8(v, _mm_srli_si128(v, 2)); v = _mm_max_epu8(v, _mm_srli_si128(v, 1));
	return _mm_cvtsi128_si32(v) & 0xff;
}
static inline int hmax_i16(__m128i v)
{
	v = _mm_max_epi16(v, _mm_srli_si128(v, 8)); v = _mm_max_epi16(v, _mm_srli_si128(v, 4)); v = _mm_max_epi16(v, _mm_srli_si128(v, 2));
	return (int16_t)(_mm_cvtsi128_si32(v) & 0xffff);
}

static bb_swr_t striped_sw_sse2(vprofile_t *q, int tlen, const uint8_t *target, int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
	const int slen = q->slen, is8 = q->is8, P = q->P;
	const int minsc = (xtra & BB_SW_XSUBO) ? xtra & 0xffff : 0x10000;
	const int endsc = (xtra & BB_SW_XSTOP) ? xtra & 0xffff : 0x10000;
	const __m128i zero = _mm_setzero_si128();
	__m128i *H0 = q->H0, *H1 = q->H1, *E = q->E, *Hmax = q->Hmax;
	BB_VEC(uint64_t) b = {0, 0, 0};
	int i, j, k, te = -1, gmax = 0;
	bb_swr_t r = {0, -1, -1, -1, -1, -1, -1};
	for (j = 0; j < slen; ++j) H0[j] = H1[j] = E[j] = Hmax[j] = zero;

	if (is8) {
		const __m128i v_oe_del = _mm_set1_epi8((char)(o_del + e_del)), v_e_del = _mm_set1_epi8((char)e_del);
		const __m128i v_oe_ins = _mm_set1_epi8((char)(o_ins + e_ins)), v_e_ins = _mm_set1_epi8((char)e_ins), v_shift = _mm_set1_epi8((char)q->shift);
		for (i = 0; i < tlen; ++i) {
			const __m128i *S = q->prof + (size_t)target[i] * slen;
			__m128i h = _mm_slli_si128(H0[slen - 1], 1), f = zero, mx = zero, e, t;
			int imax;
			for (j = 0; j < slen; ++j) {
				h = _mm_subs_epu8(_mm_adds_epu8(h, S[j]), v_shift);
				e = E[j];
				h = _mm_max_epu8(_mm_max_epu8(h, e), f);
				mx = _mm_max_epu8(mx, h);
				H1[j] = h;
				t = _mm_subs_epu8(h, v_oe_del);
				E[j] = _mm_max_epu8(_mm_subs_epu8(e, v_e_del), t);
				t = _mm_subs_epu8(h, v_oe_ins);
				f = _mm_max_epu8(_mm_subs_epu8(f, v_e_ins), t);
				h = H0[j];
			}
			for (k = 0; k < 16; ++k) {     /* lazy F */
				f = _mm_slli_si128(f, 1);
				for (j = 0; j < slen; ++j) {
					h = _mm_max_epu8(H1[j], f);
					H1[j] = h;
					h = _mm_subs_epu8(h, v_oe_ins);
					f = _mm_subs_epu8(f, v_e_ins);
					if (_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_subs_epu8(f, h), zero)) == 0xffff) goto f_done8;
				}
			}
f_done8:
			imax = hmax_u8(mx);
			if (imax >= minsc) {
				if (b.n == 0 || (int32_t)b.a[b.n - 1] + 1 != i) bb_vec_push(b, (uint64_t)imax << 32 | (uint32_t)i);
				else if ((int)(b.a[b.n - 1] >> 32) < imax) b.a[b.n - 1] = (uint64_t)imax << 32 | (uint32_t)i;
			}
			if (imax > gmax) {
				gmax = imax; te = i;
				for (j = 0; j < slen; ++j) Hmax[j] = H1[j];
				if (gmax + q->shift >= 255 || gmax >= endsc) break;
			}
			{ __m128i *x = H0; H0 = H1; H1 = x; }
		}
	} else {
		const __m128i v_oe_del = _mm_set1_epi16((short)(o_del + e_del)), v_e_del = _mm_set1_epi16((short)e_del);
		const __m128i v_oe_ins = _mm_set1_epi16((short)(o_ins + e_ins)), v_e_ins = _mm_set1_epi16((short)e_ins);
		for (i = 0; i < tlen; ++i) {
			const __m128i *S = q->prof + (size_t)target[i] * slen;
			__m128i h = _mm_slli_si128(H0[slen - 1], 2), f = zero, mx = zero, e, t;
			int imax;
			for (j = 0; j < slen; ++j) {
				h = _mm_adds_epi16(h, S[j]);
				e = E[j];
				h = _mm_max_epi16(_mm_max_epi16(h, e), f);
				mx = _mm_max_epi16(mx, h);
				H1[j] = h;
				t = _mm_subs_epu16(h, v_oe_del);
				E[j] = _mm_max_epi16(_mm_subs_epu16(e, v_e_del), t);
				t = _mm_subs_epu16(h, v_oe_ins);
				f = _mm_max_epi16(_mm_subs_epu16(f, v_e_ins), t);
				h = H0[j];
			}
			for (k = 0; k < 16; ++k) {
				f = _mm_slli_si128(f, 2);
				for (j = 0; j < slen; ++j) {
					h = _mm_max_epi16(H1[j], f);
					H1[j] = h;
					h = _mm_subs_epu16(h, v_oe_ins);
					f = _mm_subs_epu16(f, v_e_ins);
					if (_mm_movemask_epi8(_mm_cmpgt_epi16(f, h)) == 0) goto f_done16;
				}
			}
f_done16:
			imax = hmax_i16(mx);
			if (imax >= minsc) {
				if (b.n == 0 || (int32_t)b.a[b.n - 1] + 1 != i) bb_vec_push(b, (uint64_t)imax << 32 | (uint32_t)i);
				else if ((int)(b.a[b.n - 1] >> 32) < imax) b.a[b.n - 1] = (uint64_t)imax << 32 | (uint32_t)i;
			}
			if (imax > gmax) {
				gmax = imax; te = i;
				for (j = 0; j < slen; ++j) Hmax[j] = H1[j];
				if (gmax >= endsc) break;
			}
			{ __m128i *x = H0; H0 = H1; H1 = x; }
		}
	}
	r.score = is8 ? (gmax + q->shift < 255 ? gmax : 255) : gmax;
	r.te = te;
	if (!is8 || r.score != 255) {
		int best = -1, n = slen * P;
		if (!is8) r.qe = -1;
		for (i = 0; i < n; ++i) {   /* memory order of the striped layout: element i is lane i % P of stripe i / P */
			int v = is8 ? ((const uint8_t *)Hmax)[i] : ((const int16_t *)Hmax)[i], pos = i / P + i % P * slen;
			if (v > best) { best = v; r.qe = pos; }
			else if (v == best && pos < r.qe) r.qe = pos;
		}
		if (b.a) {
			int low, high;
			size_t x;
			i = (r.score + q->maxsc - 1) / q->maxsc;
			low = te - i; high = te + i;
			for (x = 0; x < b.n; ++x) {
				int e2 = (int32_t)b.a[x];
				if ((e2 < low || e2 > high) && (int)(b.a[x] >> 32) > r.score2) { r.score2 = (int)(b.a[x] >> 32); r.te2 = e2; }
			}
		}
	}
	free(b.a);
	return r;
}
#endif

/* one forward (or reverse) pass with whichever implementation is in use */
static bb_swr_t striped_pass(int is8, int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
	bb_swr_t r;
#ifdef BB_HAVE_SSE2
	static int scalar = -1;
	int use_scalar = __atomic_load_n(&scalar, __ATOMIC_RELAXED);
	if (use_scalar < 0) { const char *e = getenv("BWA_B200_SCALAR_SW"); use_scalar = e && atoi(e) != 0; __atomic_store_n(&scalar, use_scalar, __ATOMIC_RELAXED); }
	if (!use_scalar) {
		vprofile_t vq;
		vprofile_init(&vq, is8, qlen, query, mat);
		r = striped_sw_sse2(&vq, tlen, target, o_del, e_del, o_ins, e_ins, xtra);
		free(vq.mem);
		return r;
	}
#endif
	{
		profile_t q;
		profile_init(&q, is8, qlen, query, mat);
		r = striped_sw_scalar(&q, tlen, target, o_del, e_del, o_ins, e_ins, xtra);
		free(q.prof);
	}
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
	bb_swr_t r, rr;
	int is8 = (xtra & BB_SW_XBYTE) ? 1 : 0;
	r = striped_pass(is8, qlen, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, xtra);
	if ((xtra & BB_SW_XSTART) == 0 || ((xtra & BB_SW_XSUBO) && r.score < (xtra & 0xffff))) return r;
	reverse_bytes(r.qe + 1, query); reverse_bytes(r.te + 1, target);
	rr = striped_pass(is8, r.qe + 1, query, tlen, target, mat, o_del, e_del, o_ins, e_ins, BB_SW_XSTOP | r.score);
	reverse_bytes(r.qe + 1, query); reverse_bytes(r.te + 1, target);
	if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
	return r;
}
