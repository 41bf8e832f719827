/* bwag_dev.cuh -- device-side view of the index, launch/portability macros, FM-index block arithmetic.
 *
 * Index layout in HBM (one blob, see bwag_api.cu): the reference's own structures, unchanged:
 *   - Occ/BWT blocks exactly as in bwt_t::bwt (bwt.h:74-82): one 64-byte block per 128 BWT symbols =
 *     4 x u64 cumulative counts (A,C,G,T; '$' excluded) followed by 8 x u32 words of 16 2-bit symbols,
 *     first symbol in the top bits.  64-byte aligned so a block is two 32-byte HBM sectors / four
 *     16-byte vector loads;
 *   - the sampled suffix array bwt_t::sa (every sa_intv-th row, sa[0] = -1), optionally re-sampled
 *     more densely on the device;
 *   - the 2-bit forward reference pac (4 bases per byte, first base in the top bits).
 */
#ifndef BWAG_DEV_CUH
#define BWAG_DEV_CUH

#ifndef BWAG_CUSIM
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include "bwa_b200_dev.h"

#ifdef BWAG_CUSIM
#define BWAG_LAUNCH(kern, grid, block, smem, stream, ...) cusim_launch(dim3(grid), dim3(block), (smem), [&] { kern(__VA_ARGS__); })
#else
#define BWAG_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

struct DevIndex {
	const uint4 *bwt;   /* 4 x uint4 per block */
	const u64 *sa;
	const uint8_t *pac;
	u64 primary, seq_len;
	u64 L2[5];
	u64 n_sa;
	i64 l_pac;
	int sa_shift;       /* log2(sampling interval of sa[]) */
};

#define FULL_MASK 0xffffffffu

/* number of symbols == c among the first n (1..16) symbols of a BWT word (symbol 0 in the top bits);
 * returns the four counts packed one per byte (A in bits 0-7 ... T in bits 24-31) */
__device__ __forceinline__ u32 bwag_word_counts(u32 w, int n)
{
	u32 s = w >> ((16 - n) << 1);              /* drop the symbols after the n-th; zeros enter at the top */
	u32 lo = s & 0x55555555u, hi = (s >> 1) & 0x55555555u;
	u32 nT = __popc(hi & lo), nG = __popc(hi & ~lo), nC = __popc(~hi & lo & 0x55555555u);
	u32 nA = (u32)n - nT - nG - nC;             /* everything else among the n real symbols */
	return nA | nC << 8 | nG << 16 | nT << 24;
}

/* counts over symbols [0, pos] (pos in 0..127) of a block restricted to the four words held in v,
 * which are words 4*half .. 4*half+3 of the block's symbol area (half = 0 or 1) */
__device__ __forceinline__ u32 bwag_quad_counts(uint4 v, int half, int pos)
{
	int n = pos + 1 - (half << 6);              /* symbols of this 64-symbol half that count */
	u32 r = 0;
	if (n <= 0) return 0;
	if (n > 64) n = 64;
	r += bwag_word_counts(v.x, n >= 16 ? 16 : n);
	if (n > 16) r += bwag_word_counts(v.y, n >= 32 ? 16 : n - 16);
	if (n > 32) r += bwag_word_counts(v.z, n >= 48 ? 16 : n - 32);
	if (n > 48) r += bwag_word_counts(v.w, n - 48);
	return r;
}

__device__ __forceinline__ int bwag_pac_base(const uint8_t *pac, i64 k) { return pac[k >> 2] >> ((~k & 3) << 1) & 3; }

/* base at position p of the doubled (forward + reverse-complement) coordinate system */
__device__ __forceinline__ int bwag_ref_base(const DevIndex &ix, i64 p)
{
	return p < ix.l_pac ? bwag_pac_base(ix.pac, p) : 3 - bwag_pac_base(ix.pac, (ix.l_pac << 1) - 1 - p);
}

#endif
