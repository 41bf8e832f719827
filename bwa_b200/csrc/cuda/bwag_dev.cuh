/* bwag_dev.cuh -- device-side view of the index, launch/portability macros, FM-index block arithmetic.
 *
 * Index layout in HBM (one blob, see bwag_api.cu):
 *   - Occ/BWT blocks: one 64-byte block per 128 BWT symbols like bwt_t::bwt (bwt.h:74-82), with the same
 *     4 x u64 cumulative counts (A,C,G,T; '$' excluded) in the first 32 bytes, but the 128 symbols stored
 *     as two bit planes instead of the file's 2-bit packing (k_occ_planes converts in place after upload):
 *     bytes 32-47 = bit 1 of every symbol, bytes 48-63 = bit 0, symbol p of the block at bit p&31 of
 *     word p>>5 of its plane.  Ranks of all four symbols up to a position are then three masked popcounts
 *     per 32 symbols with no data-dependent branch: popc(hi), popc(lo), popc(hi&lo) give T = both,
 *     G = hi-only, C = lo-only, A = the rest.  64-byte aligned: a block is two 32-byte HBM sectors;
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

/* low t bits set, t clamped to [0,32]: which symbols of a 32-symbol plane word lie in [0,pos] */
__device__ __forceinline__ u32 bwag_plane_mask(int t) { return __funnelshift_lc(0xffffffffu, 0u, (u32)(t > 0 ? t : 0)); }

/* ranks of all four symbols over positions [0,pos] of one Occ block given its four 16-byte quarters:
 * cA = counts A,C ; cG = counts G,T ; ph, pl = the two bit planes (bwt_occ4, bwt.c:169-186) */
__device__ __forceinline__ void bwag_block_counts(const uint4 &cA, const uint4 &cG, const uint4 &ph, const uint4 &pl, int pos, u64 out[4])
{
	const int n = pos + 1;
	const u32 m0 = bwag_plane_mask(n), m1 = bwag_plane_mask(n - 32), m2 = bwag_plane_mask(n - 64), m3 = bwag_plane_mask(n - 96);
	const u32 h0 = ph.x & m0, h1 = ph.y & m1, h2 = ph.z & m2, h3 = ph.w & m3;
	const u32 l0 = pl.x & m0, l1 = pl.y & m1, l2 = pl.z & m2, l3 = pl.w & m3;
	const u32 nH = __popc(h0) + __popc(h1) + __popc(h2) + __popc(h3);
	const u32 nL = __popc(l0) + __popc(l1) + __popc(l2) + __popc(l3);
	const u32 nT = __popc(h0 & l0) + __popc(h1 & l1) + __popc(h2 & l2) + __popc(h3 & l3);
	out[0] = ((u64)cA.y << 32 | cA.x) + (u32)(n + nT - nH - nL);
	out[1] = ((u64)cA.w << 32 | cA.z) + (nL - nT);
	out[2] = ((u64)cG.y << 32 | cG.x) + (nH - nT);
	out[3] = ((u64)cG.w << 32 | cG.z) + nT;
}

/* symbol at position pos of a block and the number of its occurrences in [0,pos] */
__device__ __forceinline__ int bwag_block_symbol_rank(const uint4 &ph, const uint4 &pl, int pos, u32 *rank)
{
	const int w = pos >> 5, n = pos + 1;
	const u32 hw = w == 0 ? ph.x : w == 1 ? ph.y : w == 2 ? ph.z : ph.w, lw = w == 0 ? pl.x : w == 1 ? pl.y : w == 2 ? pl.z : pl.w;
	const int c = (int)(hw >> (pos & 31) & 1) << 1 | (int)(lw >> (pos & 31) & 1);
	const u32 fh = (c & 2) ? 0u : 0xffffffffu, fl = (c & 1) ? 0u : 0xffffffffu;   /* flip a plane where the symbol's bit is 0 */
	*rank = __popc((ph.x ^ fh) & (pl.x ^ fl) & bwag_plane_mask(n)) + __popc((ph.y ^ fh) & (pl.y ^ fl) & bwag_plane_mask(n - 32))
	      + __popc((ph.z ^ fh) & (pl.z ^ fl) & bwag_plane_mask(n - 64)) + __popc((ph.w ^ fh) & (pl.w ^ fl) & bwag_plane_mask(n - 96));
	return c;
}

__device__ __forceinline__ int bwag_pac_base(const uint8_t *pac, i64 k) { return pac[k >> 2] >> ((~k & 3) << 1) & 3; }

/* base at position p of the doubled (forward + reverse-complement) coordinate system */
__device__ __forceinline__ int bwag_ref_base(const DevIndex &ix, i64 p)
{
	return p < ix.l_pac ? bwag_pac_base(ix.pac, p) : 3 - bwag_pac_base(ix.pac, (ix.l_pac << 1) - 1 - p);
}

#endif
