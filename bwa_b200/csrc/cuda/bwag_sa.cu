/* bwag_sa.cu -- suffix-array lookup (K2) and the optional on-device densification of the SA sample.
 *
 * K2 replaces bwt_sa/bwt_invPsi/bwt_occ (bwt.c:53-59,86-129): one lane per seed walks LF-steps until it
 * hits a sampled row.  Walk lengths are geometric (mean = sampling interval - 1), so lanes that finish
 * pull new seeds (ballot + one atomicAdd per warp): a warp keeps 32 independent 64-byte requests in
 * flight regardless of the spread.  Each step reads the 16-byte quarter holding the symbol, the 16 bytes
 * holding its cumulative count and, for positions in the second half, the first symbol quarter pair --
 * all inside one 64-byte Occ block, i.e. two HBM sectors per step.
 */
#include "bwag_dev.cuh"
#include "bwag_kernels.h"

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
