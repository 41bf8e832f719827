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

/* ------------------------------------------------------------------------------------------------ K6, warp per task
 * The same striped recurrence with the P values of a vector on P LANES of a warp (16 or 8; the other lanes idle): vectors of a row
 * in shared memory ([stripe][lane] shorts), the shift by one value is a shuffle, "any lane still improves" of the lazy-F pass is a
 * ballot, the row maximum a warp reduction.  About 100 k warp instructions per 150 x 630 rescue alignment instead of ~3 M serial
 * lane instructions through global memory: the repeat-rich workload asks for 4 such alignments per read (profiles/r2_call5_*). */
__device__ SwOut sw_pass_warp(int lane, int is8, int qlen, const uint8_t *q, int tlen, const uint8_t *t, const int8_t *mat, int shift, int maxsc,
                              int o_del, int e_del, int o_ins, int e_ins, u32 xtra, short *H0, short *H1, short *E, short *Hmax, u64 *b)
{
	const int P = is8 ? 16 : 8, slen = (qlen + P - 1) / P, n = slen * P;
	const bool act = lane < P;
	const int minsc = (xtra & BWAG_SW_XSUBO) ? (int)(xtra & 0xffff) : 0x10000;
	const int endsc = (xtra & BWAG_SW_XSTOP) ? (int)(xtra & 0xffff) : 0x10000;
	const int oe_del = is8 ? (o_del + e_del) & 0xff : (o_del + e_del) & 0xffff;
	const int oe_ins = is8 ? (o_ins + e_ins) & 0xff : (o_ins + e_ins) & 0xffff;
	int te = -1, gmax = 0, nb = 0, last_i = -2, last_sc = 0;
	SwOut r; r.score = 0; r.te = -1; r.qe = -1; r.score2 = -1; r.te2 = -1;
	for (int x = lane; x < n; x += 32) { H0[x] = 0; H1[x] = 0; E[x] = 0; Hmax[x] = 0; }
	__syncwarp();
	for (int i = 0; i < tlen; ++i) {
		const int8_t *mrow = mat + t[i] * 5;
		int h, f = 0, mx = 0;
		{
			const int last = act ? H0[(slen - 1) * P + lane] : 0;      /* previous row, shifted by one value */
			h = __shfl_up_sync(FULL_MASK, last, 1);
			if (lane == 0) h = 0;
		}
		if (act) for (int j = 0; j < slen; ++j) {
			const int k = j + lane * slen;
			const int s = k >= qlen ? 0 : mrow[q[k]];
			int hv, e = E[j * P + lane], tt;
			if (is8) { hv = sw_sat_u8(h + ((s + shift) & 0xff)); hv = sw_sat_u8(hv - shift); }
			else hv = sw_sat_i16(h + s);
			if (e > hv) hv = e;
			if (f > hv) hv = f;
			if (hv > mx) mx = hv;
			H1[j * P + lane] = (short)hv;
			e -= e_del; if (e < 0) e = 0;
			tt = hv - oe_del; if (tt < 0) tt = 0;
			E[j * P + lane] = (short)(e > tt ? e : tt);
			f -= e_ins; if (f < 0) f = 0;
			tt = hv - oe_ins; if (tt < 0) tt = 0;
			if (tt > f) f = tt;
			h = H0[j * P + lane];
		}
		{   /* lazy F: at most 16 sweeps in both kernels (ksw.c:200-212, 330-340) */
			bool done = false;
			for (int k = 0; k < 16 && !done; ++k) {
				f = __shfl_up_sync(FULL_MASK, f, 1);
				if (lane == 0) f = 0;
				for (int j = 0; j < slen; ++j) {
					bool more = false;
					if (act) {
						int hv = H1[j * P + lane];
						if (f > hv) hv = f;
						H1[j * P + lane] = (short)hv;
						hv -= oe_ins; if (hv < 0) hv = 0;
						f -= e_ins; if (f < 0) f = 0;
						more = f > hv;
					}
					if (!__any_sync(FULL_MASK, more)) { done = true; break; }
				}
			}
		}
		const int imax = __reduce_max_sync(FULL_MASK, act ? mx : 0);
		if (imax >= minsc) {   /* the list's last entry lives in registers (the same in every lane); lane 0 mirrors it to memory */
			if (nb == 0 || last_i + 1 != i) { ++nb; last_i = i; last_sc = imax; if (lane == 0) b[nb - 1] = (u64)imax << 32 | (u32)i; }
			else if (last_sc < imax) { last_i = i; last_sc = imax; if (lane == 0) b[nb - 1] = (u64)imax << 32 | (u32)i; }
		}
		bool stop = false;
		if (imax > gmax) {
			gmax = imax; te = i;
			if (act) for (int j = 0; j < slen; ++j) Hmax[j * P + lane] = H1[j * P + lane];
			if ((is8 && gmax + shift >= 255) || gmax >= endsc) stop = true;
		}
		__syncwarp();
		if (stop) break;
		{ short *sw = H0; H0 = H1; H1 = sw; }
	}
	r.score = is8 ? (gmax + shift < 255 ? gmax : 255) : gmax;
	r.te = te;
	if (!is8 || r.score != 255) {
		int best = -1, bpos = 0x7fffffff;
		if (act) for (int j = 0; j < slen; ++j) {
			const int v = Hmax[j * P + lane], pos = j + lane * slen;
			if (v > best || (v == best && pos < bpos)) { best = v; bpos = pos; }
		}
		const int gb = __reduce_max_sync(FULL_MASK, best);
		r.qe = __reduce_min_sync(FULL_MASK, best == gb ? bpos : 0x7fffffff);
		if (n == 0) r.qe = -1;
		if (nb) {   /* second best: the first entry of the largest score outside [te - d, te + d] (ksw.c:241-249) */
			const int d = (r.score + maxsc - 1) / maxsc, low = te - d, high = te + d;
			int s2 = -1, x2 = 0x7fffffff;
			for (int x = lane; x < nb; x += 32) {
				const int e2 = (int)(u32)b[x], sc = (int)(b[x] >> 32);
				if ((e2 < low || e2 > high) && sc > s2) { s2 = sc; x2 = x; }
			}
			const int g2 = __reduce_max_sync(FULL_MASK, s2);
			if (g2 > -1) {
				const int gx = __reduce_min_sync(FULL_MASK, s2 == g2 ? x2 : 0x7fffffff);
				r.score2 = g2; r.te2 = (int)(u32)b[gx];
			}
		}
	}
	__syncwarp();
	return r;
}

__global__ void __launch_bounds__(128) k_localsw_warp(DevIndex ix, SwArgs a)
{
#ifdef BWAG_CUSIM
	unsigned char *dyn = cusim_dyn_smem;
#else
	extern __shared__ int4 k6_dyn[];
	unsigned char *dyn = reinterpret_cast<unsigned char *>(k6_dyn);
#endif
	const int lane = threadIdx.x & 31;
	const i64 wid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	/* shared per warp: four vector rows (short[cap_n]) and the query bytes; global per warp: the target bytes and the row-maximum list */
	unsigned char *mine = dyn + (size_t)(threadIdx.x >> 5) * (size_t)(8 * a.cap_n + a.cap_q);
	short *H0 = reinterpret_cast<short *>(mine), *H1 = H0 + a.cap_n, *E = H1 + a.cap_n, *Hmax = E + a.cap_n;
	uint8_t *qb = reinterpret_cast<uint8_t *>(Hmax + a.cap_n);
	unsigned char *gm = a.scratch + wid * a.per_thread;
	u64 *b = reinterpret_cast<u64 *>(gm);
	uint8_t *tb = reinterpret_cast<uint8_t *>(b + a.cap_t);
	const bwag_sw_par_t &p = a.par;
	__shared__ int8_t s_mat[32];
	if (threadIdx.x < 25) s_mat[threadIdx.x] = p.mat[threadIdx.x];
	__syncthreads();
	int mn = 127, maxsc = 0;
	for (int k = 0; k < 25; ++k) { mn = mn < p.mat[k] ? mn : p.mat[k]; maxsc = maxsc > p.mat[k] ? maxsc : p.mat[k]; }
	const int shift = (256 - (mn & 0xff)) & 0xff;
	for (;;) {
		int tix = 0;
		if (lane == 0) tix = atomicAdd(a.next_task, 1);
		tix = __shfl_sync(FULL_MASK, tix, 0);
		if (tix >= a.n_tasks) break;
		const bwag_swtask_t tk = a.tasks[tix];
		bwag_swres_t out;
		out.score = 0; out.te = out.qe = out.score2 = out.te2 = out.tb = out.qb = -1;
		const int qlen = tk.qlen, tlen = tk.tlen, is8 = (tk.xtra & BWAG_SW_XBYTE) ? 1 : 0;
		const int P = is8 ? 16 : 8;
		if (qlen <= 0 || tlen <= 0 || qlen > a.cap_q || tlen > a.cap_t || ((qlen + P - 1) / P) * P > a.cap_n) {
			out.score = -1;
			if (lane == 0) { a.res[tix] = out; if (qlen > 0 && tlen > 0) atomicOr(a.flags, 32u); }
			continue;
		}
		__syncwarp();
		{   /* stage the two sequences */
			const uint8_t *qs = (tk.flags & BWAG_SWF_QREAD) ? a.codes + tk.q_beg : a.pool + tk.q_beg;
			if (tk.flags & BWAG_SWF_QREV) for (int x = lane; x < qlen; x += 32) { const int c = qs[x]; qb[qlen - 1 - x] = (uint8_t)(c < 4 ? 3 - c : 4); }
			else for (int x = lane; x < qlen; x += 32) { const int c = qs[x]; qb[x] = (uint8_t)(c > 4 ? 4 : c); }
			if (tk.flags & BWAG_SWF_TREF) for (int x = lane; x < tlen; x += 32) tb[x] = (uint8_t)bwag_ref_base(ix, tk.t_beg + x);
			else for (int x = lane; x < tlen; x += 32) { const int c = a.pool[tk.t_beg + x]; tb[x] = (uint8_t)(c > 4 ? 4 : c); }
		}
		__syncwarp();
		const SwOut r = sw_pass_warp(lane, is8, qlen, qb, tlen, tb, s_mat, shift, maxsc, p.o_del, p.e_del, p.o_ins, p.e_ins, tk.xtra, H0, H1, E, Hmax, b);
		out.score = r.score; out.te = r.te; out.qe = r.qe; out.score2 = r.score2; out.te2 = r.te2;
		if ((tk.xtra & BWAG_SW_XSTART) && r.qe >= 0 && r.te >= 0 && !((tk.xtra & BWAG_SW_XSUBO) && r.score < (int)(tk.xtra & 0xffff))) {   /* start: the same kernel backwards (ksw.c:393-399) */
			for (int x = lane; x < (r.qe + 1) >> 1; x += 32) { const uint8_t y = qb[x]; qb[x] = qb[r.qe - x]; qb[r.qe - x] = y; }
			for (int x = lane; x < (r.te + 1) >> 1; x += 32) { const uint8_t y = tb[x]; tb[x] = tb[r.te - x]; tb[r.te - x] = y; }
			__syncwarp();
			const SwOut rr = sw_pass_warp(lane, is8, r.qe + 1, qb, tlen, tb, s_mat, shift, maxsc, p.o_del, p.e_del, p.o_ins, p.e_ins, BWAG_SW_XSTOP | (u32)r.score, H0, H1, E, Hmax, b);
			if (r.score == rr.score) { out.tb = r.te - rr.te; out.qb = r.qe - rr.qe; }
		}
		if (lane == 0) a.res[tix] = out;
		__syncwarp();
	}
}
