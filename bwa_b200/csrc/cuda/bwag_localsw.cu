/* bwag_localsw.cu -- K6: batched local Smith-Waterman with start recovery (ksw_align2, ksw.c:379-401, over the striped
 * kernels ksw_u8 / ksw_i16, ksw.c:122-370), the alignment behind mate rescue (mem_matesw, bwamem_pair.c:137-206) and the
 * long-read seed filter (mem_seed_sw, bwamem.c:597-622).
 *
 * What has to be reproduced is not "a" local alignment but the numbers Farrar's striped kernels return: score, end points,
 * second-best score (the `b` list of row maxima, ksw.c:215-223,241-249), smallest query end among equal scores (237-239),
 * the early stop (228), the biased saturating bytes of the 8-bit kernel, and E taken from H before the lazy-F pass
 * (189-192,200), which makes E depend on the stripe length.  So each task evaluates the striped recurrence itself: "one
 * vector" = P values (16 bytes or 8 words), value l of stripe j = query position j + l*slen.
 *
 * Mapping to the machine.  One LANE per task; the P values of a vector are a short inner loop of the lane, the vectors of a row
 * (H, E, the row before, the best row) live in the lane's slice of a global scratch area that stays L1/L2 resident (~1.3 KB per
 * task for 150-bp queries), and a warp works on 32 tasks.  Tasks are independent and few on unique references (mate rescue
 * triggers for pairs without a proper mate; 0.8 per read on repeat-rich data), so the kernel favours exactness and simplicity over
 * the last factor in speed: ~30 integer instructions per (query position x reference position).
 * Query: the batch's read (optionally reverse-complemented) or bytes of a caller pool; target: a window of the reference in the
 * doubled coordinate system or bytes of the pool (known-answer tests).
 */
#include "bwag_dev.cuh"
#include "bwag_kernels.h"

__device__ __forceinline__ int sw_sat_u8(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
__device__ __forceinline__ int sw_sat_i16(int v) { return v < -32768 ? -32768 : v > 32767 ? 32767 : v; }

struct SwOut { int score, te, qe, score2, te2; };

/* one pass of ksw_u8 (is8) / ksw_i16 over query q[0..qlen), target t[0..tlen) */
__device__ SwOut sw_pass(int is8, int qlen, const uint8_t *q, int tlen, const uint8_t *t, const int8_t *mat, int shift, int maxsc,
                         int o_del, int e_del, int o_ins, int e_ins, u32 xtra, short *H0, short *H1, short *E, short *Hmax, u64 *b)
{
	const int P = is8 ? 16 : 8, slen = (qlen + P - 1) / P, n = slen * P;
	const int minsc = (xtra & BWAG_SW_XSUBO) ? (int)(xtra & 0xffff) : 0x10000;
	const int endsc = (xtra & BWAG_SW_XSTOP) ? (int)(xtra & 0xffff) : 0x10000;
	const int oe_del = is8 ? (o_del + e_del) & 0xff : (o_del + e_del) & 0xffff;
	const int oe_ins = is8 ? (o_ins + e_ins) & 0xff : (o_ins + e_ins) & 0xffff;
	int h[16], f[16], mx[16];
	int te = -1, gmax = 0, nb = 0;
	SwOut r; r.score = 0; r.te = -1; r.qe = -1; r.score2 = -1; r.te2 = -1;
	for (int x = 0; x < n; ++x) { H0[x] = 0; H1[x] = 0; E[x] = 0; Hmax[x] = 0; }
	for (int i = 0; i < tlen; ++i) {
		const int8_t *mrow = mat + t[i] * 5;
		int imax = 0, done = 0;
		for (int l = P - 1; l > 0; --l) h[l] = H0[(slen - 1) * P + l - 1];   /* previous row, shifted by one value */
		h[0] = 0;
		for (int l = 0; l < P; ++l) f[l] = mx[l] = 0;
		for (int j = 0; j < slen; ++j) {
			for (int l = 0; l < P; ++l) {
				const int k = j + l * slen;
				const int s = k >= qlen ? 0 : mrow[q[k]];
				int hv, e = E[j * P + l], tt;
				if (is8) { hv = sw_sat_u8(h[l] + ((s + shift) & 0xff)); hv = sw_sat_u8(hv - shift); }
				else hv = sw_sat_i16(h[l] + s);
				if (e > hv) hv = e;
				if (f[l] > hv) hv = f[l];
				if (hv > mx[l]) mx[l] = hv;
				H1[j * P + l] = (short)hv;
				e -= e_del; if (e < 0) e = 0;
				tt = hv - oe_del; if (tt < 0) tt = 0;
				E[j * P + l] = (short)(e > tt ? e : tt);
				f[l] -= e_ins; if (f[l] < 0) f[l] = 0;
				tt = hv - oe_ins; if (tt < 0) tt = 0;
				if (tt > f[l]) f[l] = tt;
				h[l] = H0[j * P + l];
			}
		}
		for (int k = 0; k < 16 && !done; ++k) {   /* lazy F: at most 16 sweeps in both kernels */
			for (int l = P - 1; l > 0; --l) f[l] = f[l - 1];
			f[0] = 0;
			for (int j = 0; j < slen; ++j) {
				int any = 0;
				for (int l = 0; l < P; ++l) {
					int hv = H1[j * P + l];
					if (f[l] > hv) hv = f[l];
					H1[j * P + l] = (short)hv;
					hv -= oe_ins; if (hv < 0) hv = 0;
					f[l] -= e_ins; if (f[l] < 0) f[l] = 0;
					if (f[l] > hv) any = 1;
				}
				if (!any) { done = 1; break; }
			}
		}
		for (int l = 0; l < P; ++l) if (mx[l] > imax) imax = mx[l];
		if (imax >= minsc) {
			if (nb == 0 || (int)(u32)b[nb - 1] + 1 != i) b[nb++] = (u64)imax << 32 | (u32)i;
			else if ((int)(b[nb - 1] >> 32) < imax) b[nb - 1] = (u64)imax << 32 | (u32)i;
		}
		if (imax > gmax) {
			gmax = imax; te = i;
			for (int x = 0; x < n; ++x) Hmax[x] = H1[x];
			if ((is8 && gmax + shift >= 255) || gmax >= endsc) break;
		}
		{ short *sw = H0; H0 = H1; H1 = sw; }
	}
	r.score = is8 ? (gmax + shift < 255 ? gmax : 255) : gmax;
	r.te = te;
	if (!is8 || r.score != 255) {
		int best = -1;
		if (!is8) r.qe = -1;
		for (int x = 0; x < n; ++x) {
			const int v = Hmax[x], pos = x / P + x % P * slen;
			if (v > best) { best = v; r.qe = pos; }
			else if (v == best && pos < r.qe) r.qe = pos;
		}
		if (nb) {
			const int d = (r.score + maxsc - 1) / maxsc, low = te - d, high = te + d;
			for (int x = 0; x < nb; ++x) {
				const int e2 = (int)(u32)b[x];
				if ((e2 < low || e2 > high) && (int)(b[x] >> 32) > r.score2) { r.score2 = (int)(b[x] >> 32); r.te2 = e2; }
			}
		}
	}
	return r;
}

__device__ __forceinline__ void sw_reverse(int l, uint8_t *s) { for (int i = 0; i < l >> 1; ++i) { const uint8_t x = s[i]; s[i] = s[l - 1 - i]; s[l - 1 - i] = x; } }

__global__ void __launch_bounds__(64) k_localsw(DevIndex ix, SwArgs a)
{
	const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned char *mine = a.scratch + tid * a.per_thread;
	/* this lane's scratch: four vectors rows (short[cap_n]), the row-maximum list (u64[cap_t]), query and target bytes */
	short *H0 = reinterpret_cast<short *>(mine), *H1 = H0 + a.cap_n, *E = H1 + a.cap_n, *Hmax = E + a.cap_n;
	u64 *b = reinterpret_cast<u64 *>(Hmax + a.cap_n);
	uint8_t *qb = reinterpret_cast<uint8_t *>(b + a.cap_t), *tb = qb + a.cap_q;
	const bwag_sw_par_t &p = a.par;
	int mn = 127, maxsc = 0;
	for (int k = 0; k < 25; ++k) { mn = mn < p.mat[k] ? mn : p.mat[k]; maxsc = maxsc > p.mat[k] ? maxsc : p.mat[k]; }
	const int shift = (256 - (mn & 0xff)) & 0xff;   /* bias that makes every matrix entry non-negative (ksw.c:84-86) */
	for (;;) {
		const int tix = atomicAdd(a.next_task, 1);
		if (tix >= a.n_tasks) break;
		const bwag_swtask_t tk = a.tasks[tix];
		bwag_swres_t out;
		out.score = 0; out.te = out.qe = out.score2 = out.te2 = out.tb = out.qb = -1;
		const int qlen = tk.qlen, tlen = tk.tlen, is8 = (tk.xtra & BWAG_SW_XBYTE) ? 1 : 0;
		const int P = is8 ? 16 : 8;
		if (qlen <= 0 || tlen <= 0 || qlen > a.cap_q || tlen > a.cap_t || ((qlen + P - 1) / P) * P > a.cap_n) { out.score = -1; a.res[tix] = out; if (qlen > 0 && tlen > 0) atomicOr(a.flags, 32u); continue; }
		{   /* stage the two sequences */
			const uint8_t *qs = (tk.flags & BWAG_SWF_QREAD) ? a.codes + tk.q_beg : a.pool + tk.q_beg;
			if (tk.flags & BWAG_SWF_QREV) for (int x = 0; x < qlen; ++x) { const int c = qs[x]; qb[qlen - 1 - x] = (uint8_t)(c < 4 ? 3 - c : 4); }
			else for (int x = 0; x < qlen; ++x) { const int c = qs[x]; qb[x] = (uint8_t)(c > 4 ? 4 : c); }
			if (tk.flags & BWAG_SWF_TREF) for (int x = 0; x < tlen; ++x) tb[x] = (uint8_t)bwag_ref_base(ix, tk.t_beg + x);
			else for (int x = 0; x < tlen; ++x) { const int c = a.pool[tk.t_beg + x]; tb[x] = (uint8_t)(c > 4 ? 4 : c); }
		}
		const SwOut r = sw_pass(is8, qlen, qb, tlen, tb, p.mat, shift, maxsc, p.o_del, p.e_del, p.o_ins, p.e_ins, tk.xtra, H0, H1, E, Hmax, b);
		out.score = r.score; out.te = r.te; out.qe = r.qe; out.score2 = r.score2; out.te2 = r.te2;
		if ((tk.xtra & BWAG_SW_XSTART) && r.qe >= 0 && r.te >= 0 && !((tk.xtra & BWAG_SW_XSUBO) && r.score < (int)(tk.xtra & 0xffff))) {   /* start: the same kernel backwards (ksw.c:393-399) */
			sw_reverse(r.qe + 1, qb); sw_reverse(r.te + 1, tb);
			const SwOut rr = sw_pass(is8, r.qe + 1, qb, tlen, tb, p.mat, shift, maxsc, p.o_del, p.e_del, p.o_ins, p.e_ins, BWAG_SW_XSTOP | (u32)r.score, H0, H1, E, Hmax, b);
			if (r.score == rr.score) { out.tb = r.te - rr.te; out.qb = r.qe - rr.qe; }
		}
		a.res[tix] = out;
	}
}
