/* bwag_seed.cu -- stage 1 kernels: SMEM seeding (K1) and suffix-array lookup (K2).
 *
 * K1 replaces mem_collect_intv (bwamem.c:140-188) and everything under it: bwt_smem1a (bwt.c:289-351),
 * bwt_seed_strategy1 (bwt.c:358-379), bwt_extend (bwt.c:262-275), bwt_2occ4/bwt_occ4 (bwt.c:169-220).
 *
 * Mapping to the machine.  The work is a chain of dependent FM-index steps per read (~700 for a
 * 150-bp read), each of which needs two 64-byte Occ blocks at effectively random addresses of a
 * multi-GB table: the kernel is bound by HBM latency x parallelism, not by arithmetic.  So:
 *   - a read is owned by a GROUP of 8 lanes (4 reads per warp).  One bwt_extend = one converged step
 *     of the whole warp: lane g*8+t loads the t-th 16-byte quarter of block(k) (t<4) or of block(l)
 *     (t>=4) -- one LDG.128 per lane, 32 lanes = 8 full blocks = 16 HBM sectors per instruction --
 *     counts symbols in its quarter, and the group combines the pieces with shuffles;
 *   - each group runs the seeding algorithm as a state machine that is advanced until it needs the
 *     next bwt_extend, so all four groups of a warp (which are in different phases of different
 *     reads) meet at the same load instruction every iteration: the memory system always sees full
 *     warps of independent requests, and divergent bookkeeping stays short;
 *   - groups are persistent and pull the next read from an atomic counter, so long and short reads
 *     balance; the grid is sized to fill every SM (host side);
 *   - interval lists live in a per-group scratch area in global memory (L1/L2 resident); the result
 *     of a read is sorted by (start,end) on the spot and appended to the batch-wide output pools with
 *     one atomicAdd, together with the BWT rows whose suffix-array values K2 must resolve.
 *
 * K2 replaces bwt_sa/bwt_invPsi/bwt_occ (bwt.c:53-59,86-129): one lane per seed walks LF-steps until
 * it hits a sampled row.  Lanes that finish pull new seeds (ballot + one atomicAdd per warp), so a
 * warp keeps 32 independent 64-byte requests in flight despite the geometric walk lengths.
 */
#include "bwag_dev.cuh"
#include "bwag_kernels.h"

#define GRP 8                      /* lanes per read */
#define GRP_PER_WARP (32 / GRP)

enum { ST_IDLE = 0, ST_FWD, ST_BWD, ST_STR, ST_NONE };

struct Intv { u64 x0, x1, x2, info; };

__device__ __forceinline__ Intv ld_intv(const Intv *p)
{
	const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(p);
	ulonglong2 a = q[0], b = q[1];
	Intv r; r.x0 = a.x; r.x1 = a.y; r.x2 = b.x; r.info = b.y;
	return r;
}
__device__ __forceinline__ void st_intv(Intv *p, const Intv &v)
{
	ulonglong2 *q = reinterpret_cast<ulonglong2 *>(p);
	ulonglong2 a, b; a.x = v.x0; a.y = v.x1; b.x = v.x2; b.y = v.info;
	q[0] = a; q[1] = b;
}

/* One bidirectional extension for every group of the warp (bwt.c:262-275 on top of bwt.c:189-220).
 * Must be called by all 32 lanes converged.  need: this group wants an extension of (x0,x1,x2);
 * back: backward (1) or forward (0).  Results: ok_x2[c] = size of the interval after adding c,
 * ok_s[c] = its start in the searched direction, ok_o[c] = start in the other direction.
 * Returns the number of 64-byte blocks the reference would touch for this call (0 if !need). */
__device__ __forceinline__ int group_extend(const DevIndex &ix, int lane, bool need, int back, u64 x0, u64 x1, u64 x2,
                                            u64 ok_s[4], u64 ok_o[4], u64 ok_x2[4])
{
	const int t = lane & (GRP - 1), half = t >> 2, part = t & 3, gbase = lane & ~(GRP - 1);
	const u64 xs = back ? x0 : x1;             /* x[!is_back] */
	const u64 xo = back ? x1 : x0;             /* x[is_back]  */
	const u64 k = xs - 1, l = xs - 1 + x2;
	const u64 kk = half ? l : k;
	u64 c0 = 0, c1 = 0;                        /* part 0: counts A,C   part 1: counts G,T */
	u32 pc = 0;                                /* part 2,3: packed symbol counts */
	if (need && kk != (u64)-1) {
		u64 kp = kk - (kk >= ix.primary);
		uint4 v = __ldg(ix.bwt + ((kp >> 7) << 2) + part);
		if (part == 0 || part == 1) { c0 = (u64)v.y << 32 | v.x; c1 = (u64)v.w << 32 | v.z; }
		else pc = bwag_quad_counts(v, part - 2, (int)(kp & 127));
	}
	u64 tk[4], tl[4];
	{
		u64 a, b;
		u32 p;
		a = __shfl_sync(FULL_MASK, c0, gbase + 0); b = __shfl_sync(FULL_MASK, c1, gbase + 0); tk[0] = a; tk[1] = b;
		a = __shfl_sync(FULL_MASK, c0, gbase + 1); b = __shfl_sync(FULL_MASK, c1, gbase + 1); tk[2] = a; tk[3] = b;
		p = __shfl_sync(FULL_MASK, pc, gbase + 2) + __shfl_sync(FULL_MASK, pc, gbase + 3);
		tk[0] += p & 0xff; tk[1] += p >> 8 & 0xff; tk[2] += p >> 16 & 0xff; tk[3] += p >> 24;
		a = __shfl_sync(FULL_MASK, c0, gbase + 4); b = __shfl_sync(FULL_MASK, c1, gbase + 4); tl[0] = a; tl[1] = b;
		a = __shfl_sync(FULL_MASK, c0, gbase + 5); b = __shfl_sync(FULL_MASK, c1, gbase + 5); tl[2] = a; tl[3] = b;
		p = __shfl_sync(FULL_MASK, pc, gbase + 6) + __shfl_sync(FULL_MASK, pc, gbase + 7);
		tl[0] += p & 0xff; tl[1] += p >> 8 & 0xff; tl[2] += p >> 16 & 0xff; tl[3] += p >> 24;
	}
#pragma unroll
	for (int c = 0; c < 4; ++c) { ok_s[c] = ix.L2[c] + 1 + tk[c]; ok_x2[c] = tl[c] - tk[c]; }
	ok_o[3] = xo + ((xs <= ix.primary && xs + x2 - 1 >= ix.primary) ? 1 : 0);
	ok_o[2] = ok_o[3] + ok_x2[3];
	ok_o[1] = ok_o[2] + ok_x2[2];
	ok_o[0] = ok_o[1] + ok_x2[1];
	if (!need) return 0;
	if (k == (u64)-1 || l == (u64)-1) return 2;
	return ((k - (k >= ix.primary)) >> 7) != ((l - (l >= ix.primary)) >> 7) ? 2 : 1;
}

__device__ __forceinline__ void init_intv(const DevIndex &ix, int c, u64 &x0, u64 &x1, u64 &x2)
{
	x0 = ix.L2[c] + 1; x2 = ix.L2[c + 1] - ix.L2[c]; x1 = ix.L2[3 - c] + 1;
}

/* scratch of one group: two interval lists of cap_list entries, the raw results of the current
 * bwt_smem1 call (cap_list) and the accumulated list of the read (cap_mem) */
__global__ void __launch_bounds__(K1_THREADS)
k_smem(DevIndex ix, SeedArgs a)
{
	const int lane = threadIdx.x & 31, t = lane & (GRP - 1);
	const u32 gmask = ((1u << GRP) - 1) << (lane & ~(GRP - 1));
	const bool leader = t == 0;
	const i64 gid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) / GRP;
	Intv *listA = a.scratch + gid * (i64)(3 * a.cap_list + a.cap_mem), *listB = listA + a.cap_list, *m1 = listB + a.cap_list, *mem = m1 + a.cap_list;

	/* group-uniform state (every lane of the group holds the same values) */
	int rid = -1, len = 0, pass = 3, st = ST_IDLE, x = 0, k2 = 0, old_n = 0;
	int sx = 0, min_intv = 1, i = 0, j = 0, n_prev = 0, n_curr = 0, rev_first = 0, mem_n = 0, m1_n = 0, last_start = 0, ret = 0;
	const uint8_t *q = 0;
	Intv *prev = listA, *curr = listB;
	u64 ik0 = 0, ik1 = 0, ik2 = 0, ikinfo = 0, curr_last_x2 = 0, pinfo = 0;
	u64 e0 = 0, e1 = 0, e2 = 0;     /* interval to extend this step */
	u64 touches = 0;
	int overflow = 0;

	for (;;) {
		/* ---- advance this group's state machine until it needs a bwt_extend (or has no read) ---- */
		bool need = false;
		int back = 0;
		for (;;) {
			if (st == ST_IDLE) {
				if (pass == 0) {            /* first pass: all SMEMs (bwamem.c:147-157) */
					while (x < len && q[x] > 3) ++x;
					if (x >= len) { pass = 1; k2 = 0; old_n = mem_n; continue; }
					sx = x; min_intv = 1;
				} else if (pass == 1) {     /* second pass: re-seed inside long, rare SMEMs (bwamem.c:159-168) */
					bool found = false;
					while (k2 < old_n) {
						Intv p = ld_intv(mem + k2); ++k2;
						int s = (int)(p.info >> 32), e = (int)(u32)p.info;
						if (e - s < a.split_len || p.x2 > (u64)a.split_width) continue;
						sx = (s + e) >> 1; min_intv = (int)p.x2 + 1; found = true;
						break;
					}
					if (!found) { pass = 2; x = 0; continue; }
				} else if (pass == 2) {     /* third pass: forward-only seeds (bwamem.c:170-185) */
					if (a.max_mem_intv == 0) { pass = 3; continue; }
					while (x < len && q[x] > 3) ++x;
					if (x >= len) { pass = 3; continue; }
					init_intv(ix, q[x], ik0, ik1, ik2);
					i = x + 1; st = ST_STR;
					continue;
				} else {                    /* read finished (or none yet): emit, fetch the next one */
					if (rid >= 0) {
						/* sort by info (rank sort; equal infos are identical intervals), append to the pools */
						int n = mem_n, tot_seeds = 0;
						for (int e = 0; e < n; ++e) { u64 occ = mem[e].x2; tot_seeds += (int)(occ < (u64)a.max_occ ? occ : (u64)a.max_occ); }
						i64 base = 0, sbase = 0;
						if (leader) {
							base = (i64)atomicAdd(a.n_intv, (u64)n);
							sbase = (i64)atomicAdd(a.n_seeds, (u64)tot_seeds);
							a.intv_beg[rid] = base; a.intv_n[rid] = n;
						}
						base = __shfl_sync(gmask, base, lane & ~(GRP - 1));
						sbase = __shfl_sync(gmask, sbase, lane & ~(GRP - 1));
						if (base + n > a.cap_intv || sbase + tot_seeds > a.cap_seeds) overflow |= 1;
						else {
							for (int e = t; e < n; e += GRP) {
								Intv p = ld_intv(mem + e);
								int rank = 0; i64 sb = 0;
								for (int f = 0; f < n; ++f) {
									u64 fi = mem[f].info;
									if (fi < p.info || (fi == p.info && f < e)) { u64 occ = mem[f].x2; ++rank; sb += (i64)(occ < (u64)a.max_occ ? occ : (u64)a.max_occ); }
								}
								st_intv(reinterpret_cast<Intv *>(a.intv) + base + rank, p);
								a.seed_beg[base + rank] = sbase + sb;
								/* BWT rows whose SA value is wanted: x0 + c*step (bwamem.c:304-309) */
								i64 cnt = (i64)(p.x2 < (u64)a.max_occ ? p.x2 : (u64)a.max_occ);
								u64 step = p.x2 > (u64)a.max_occ ? p.x2 / (u64)a.max_occ : 1;
								for (i64 c = 0; c < cnt; ++c) a.rbeg[sbase + sb + c] = (i64)(p.x0 + (u64)c * step);
							}
						}
					}
					int nr = 0;
					if (leader) nr = atomicAdd(a.next_read, 1);
					nr = __shfl_sync(gmask, nr, lane & ~(GRP - 1));
					if (nr >= a.n_reads) { rid = -1; st = ST_NONE; break; }
					rid = nr; q = a.codes + a.off[rid]; len = (int)(a.off[rid + 1] - a.off[rid]);
					pass = 0; x = 0; mem_n = 0;
					if (len > a.cap_list) { overflow |= 8; pass = 3; mem_n = 0; }
					continue;
				}
				/* start bwt_smem1(sx, min_intv) (bwt.c:289-303) */
				init_intv(ix, q[sx], ik0, ik1, ik2);
				ikinfo = (u64)sx + 1;
				i = sx + 1; n_curr = 0; m1_n = 0; st = ST_FWD;
				continue;
			}
			if (st == ST_FWD) {
				if (i < len && q[i] < 4) { e0 = ik0; e1 = ik1; e2 = ik2; need = true; back = 0; break; }
				/* end of read or ambiguous base: keep the current interval, then turn around */
				if (leader) { Intv v; v.x0 = ik0; v.x1 = ik1; v.x2 = ik2; v.info = ikinfo; st_intv(curr + n_curr, v); }
				++n_curr;
				goto turn_around;
			}
			if (st == ST_STR) {
				if (i >= len) { x = len; st = ST_IDLE; continue; }
				if (q[i] > 3) { x = i + 1; st = ST_IDLE; continue; }
				e0 = ik0; e1 = ik1; e2 = ik2; need = true; back = 0;
				break;
			}
			if (st == ST_BWD) {
				int c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
				if (c < 0) {
					/* nothing extends: only the longest candidate (first in the list) can be an SMEM */
					if (m1_n == 0 || i + 1 < last_start) {
						Intv p = ld_intv(prev + (rev_first ? n_prev - 1 : 0));
						p.info |= (u64)(i + 1) << 32;
						if (leader) st_intv(m1 + m1_n, p);
						++m1_n; last_start = i + 1;
					}
					goto call_done;
				}
				if (j < n_prev) {
					Intv p = ld_intv(prev + (rev_first ? n_prev - 1 - j : j));
					e0 = p.x0; e1 = p.x1; e2 = p.x2; pinfo = p.info; need = true; back = 1;
					break;
				}
				if (n_curr == 0) goto call_done;
				{ Intv *tmp = prev; prev = curr; curr = tmp; }
				n_prev = n_curr; n_curr = 0; rev_first = 0; --i; j = 0;
				__syncwarp(gmask);          /* list entries written by the leader become visible to the group */
				continue;
			}
			break; /* ST_NONE */

turn_around: /* forward sweep finished (bwt.c:323-326): candidates are visited longest match first */
			ret = (int)(u32)ikinfo;     /* info of the last pushed interval == end of the longest match */
			{ Intv *tmp = prev; prev = curr; curr = tmp; }
			n_prev = n_curr; n_curr = 0; rev_first = 1; i = sx - 1; j = 0; st = ST_BWD;
			__syncwarp(gmask);
			continue;

call_done:  /* bwt_smem1 returns: keep matches of at least min_seed_len, in ascending start order */
			__syncwarp(gmask);
			for (int e = m1_n - 1; e >= 0; --e) {
				Intv p = ld_intv(m1 + e);
				if ((int)((u32)p.info - (u32)(p.info >> 32)) >= a.min_seed_len) {
					if (mem_n < a.cap_mem) { if (leader) st_intv(mem + mem_n, p); ++mem_n; }
					else overflow |= 8;
				}
			}
			__syncwarp(gmask);
			if (pass == 0) x = ret;
			st = ST_IDLE;
			continue;
		}

		/* ---- one converged bwt_extend for all groups of the warp ---- */
		__syncwarp();
		if (__all_sync(FULL_MASK, st == ST_NONE)) break;
		u64 ok_s[4], ok_o[4], ok_x2[4];
		touches += (u64)group_extend(ix, lane, need, back, e0, e1, e2, ok_s, ok_o, ok_x2);
		if (!need) continue;

		/* ---- consume the result ---- */
		if (st == ST_FWD) {                 /* bwt.c:307-316; forward extension by base b uses ok[3-b] */
			int c = 3 - q[i];
			bool stop = false;
			if (ok_x2[c] != ik2) {
				if (n_curr < a.cap_list) { if (leader) { Intv v; v.x0 = ik0; v.x1 = ik1; v.x2 = ik2; v.info = ikinfo; st_intv(curr + n_curr, v); } ++n_curr; }
				else overflow |= 8;
				if (ok_x2[c] < (u64)min_intv) stop = true;
			}
			if (stop) {
				ret = (int)(u32)ikinfo;
				{ Intv *tmp = prev; prev = curr; curr = tmp; }
				n_prev = n_curr; n_curr = 0; rev_first = 1; i = sx - 1; j = 0; st = ST_BWD;
				__syncwarp(gmask);
			} else {
				ik0 = ok_o[c]; ik1 = ok_s[c]; ik2 = ok_x2[c]; ikinfo = (u64)i + 1;  /* forward: x[1] is the searched side */
				++i;
				if (i == len) { /* reached the end: the last interval is a candidate too (bwt.c:322) */
					if (n_curr < a.cap_list) { if (leader) { Intv v; v.x0 = ik0; v.x1 = ik1; v.x2 = ik2; v.info = ikinfo; st_intv(curr + n_curr, v); } ++n_curr; }
					else overflow |= 8;
					ret = (int)(u32)ikinfo;
					{ Intv *tmp = prev; prev = curr; curr = tmp; }
					n_prev = n_curr; n_curr = 0; rev_first = 1; i = sx - 1; j = 0; st = ST_BWD;
					__syncwarp(gmask);
				}
			}
		} else if (st == ST_STR) {          /* bwt.c:366-375 */
			int c = 3 - q[i];
			if (ok_x2[c] < a.max_mem_intv && i - x >= a.min_seed_len) {
				if (ok_x2[c] > 0) {
					if (mem_n < a.cap_mem) {
						if (leader) { Intv v; v.x0 = ok_o[c]; v.x1 = ok_s[c]; v.x2 = ok_x2[c]; v.info = (u64)x << 32 | (u64)(i + 1); st_intv(mem + mem_n, v); }
						++mem_n;
					} else overflow |= 8;
				}
				x = i + 1; st = ST_IDLE;
			} else { ik0 = ok_o[c]; ik1 = ok_s[c]; ik2 = ok_x2[c]; ++i; }
		} else if (st == ST_BWD) {          /* bwt.c:331-343; backward extension by base b uses ok[b] */
			int c = q[i];
			if (ok_x2[c] < (u64)min_intv) {
				if (n_curr == 0 && (m1_n == 0 || i + 1 < last_start)) {
					if (leader) { Intv v; v.x0 = e0; v.x1 = e1; v.x2 = e2; v.info = pinfo | (u64)(i + 1) << 32; st_intv(m1 + m1_n, v); }
					++m1_n; last_start = i + 1;
				}
			} else if (n_curr == 0 || ok_x2[c] != curr_last_x2) {
				if (leader) { Intv v; v.x0 = ok_s[c]; v.x1 = ok_o[c]; v.x2 = ok_x2[c]; v.info = pinfo; st_intv(curr + n_curr, v); }
				++n_curr; curr_last_x2 = ok_x2[c];
			}
			++j;
		}
	}
	if (leader && touches) atomicAdd(a.occ_touches, touches);
	if (overflow && leader) atomicOr(a.flags, (u32)overflow);
}

/* ------------------------------------------------------------------------------------------------ K2 */

/* one LF step: row of the preceding text position (bwt.c:53-59 with bwt_occ bwt.c:107-129) */
__device__ __forceinline__ u64 lf_step(const DevIndex &ix, u64 k)
{
	if (k == ix.primary) return 0;
	u64 kp = k - (k > ix.primary);                 /* row in the '$'-less BWT == what bwt_occ uses since k != primary */
	const uint4 *blk = ix.bwt + ((kp >> 7) << 2);
	int pos = (int)(kp & 127);
	uint4 w = __ldg(blk + 2 + (pos >> 6));         /* the 64-symbol half holding kp */
	u32 word = (pos >> 4 & 3) == 0 ? w.x : (pos >> 4 & 3) == 1 ? w.y : (pos >> 4 & 3) == 2 ? w.z : w.w;
	int c = word >> ((~pos & 15) << 1) & 3;
	uint4 cn = __ldg(blk + (c >> 1));
	u64 n = (c & 1) ? ((u64)cn.w << 32 | cn.z) : ((u64)cn.y << 32 | cn.x);
	u32 pc = bwag_quad_counts(w, pos >> 6, pos);
	if (pos >= 64) pc += bwag_quad_counts(__ldg(blk + 2), 0, pos);
	return ix.L2[c] + n + (pc >> (c << 3) & 0xff);
}

__global__ void __launch_bounds__(K2_THREADS)
k_sa(DevIndex ix, SaArgs a)
{
	const int lane = threadIdx.x & 31;
	const u64 mask = ((u64)1 << ix.sa_shift) - 1;
	i64 idx = -1;
	u64 k = 0, steps = 0, touches = 0, algo = 0;
	for (;;) {
		/* refill idle lanes: one atomicAdd per warp for all of them */
		bool idle = idx < 0;
		u32 bal = __ballot_sync(FULL_MASK, idle);
		if (bal) {
			i64 base = 0;
			int leader = __ffs(bal) - 1;
			if (lane == leader) base = (i64)atomicAdd(a.next, (u64)__popc(bal));
			base = __shfl_sync(FULL_MASK, base, leader);
			if (idle) {
				i64 mine = base + __popc(bal & ((1u << lane) - 1));
				if (mine < a.n) { idx = mine; k = (u64)a.rbeg[idx]; steps = 0; }
			}
		}
		if (__all_sync(FULL_MASK, idx < 0)) break;
		if (idx >= 0) {
			if ((k & mask) == 0) {
				a.rbeg[idx] = (i64)(steps + ix.sa[k >> ix.sa_shift]);
				idx = -1;
			} else { k = lf_step(ix, k); ++steps; ++touches; }
			/* what the walk would have cost with the on-disk sample (every 32nd row): counted separately */
			(void)algo;
		}
	}
	if (touches) atomicAdd(a.sa_touches, touches);
}

/* densify the suffix-array sample: out[r] = SA[r << out_shift] for every r, walking from the existing sample */
__global__ void k_sa_densify(DevIndex ix, u64 *out, int out_shift, u64 n_out)
{
	const u64 mask = ((u64)1 << ix.sa_shift) - 1;
	for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_out; r += (u64)gridDim.x * blockDim.x) {
		u64 k = r << out_shift, steps = 0;
		while (k & mask) { k = lf_step(ix, k); ++steps; }
		out[r] = r == 0 ? (u64)-1 : steps + ix.sa[k >> ix.sa_shift];
	}
}
