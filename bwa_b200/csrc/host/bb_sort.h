/* bb_sort.h -- the unstable sort whose exact permutation the SAM output depends on.
 *
 * BWA-MEM sorts chains by weight (bwamem.c:367) and regions by end/score (bwamem.c:467,504) with
 * klib's ks_introsort (ksort.h:176-226).  Keys tie, the sort is not stable, and later steps keep
 * "the first of equals", so a bit-exact re-implementation must visit and swap elements in the same
 * sequence.  The procedure restated here:
 *   - n == 2: one compare/swap; otherwise quicksort with an explicit stack and a depth budget of
 *     2*ceil(log2 n) (floor 2 levels -> 4);
 *   - pivot: median of first, last and the element just right of the middle, moved to the right end;
 *   - Hoare-style scan (left index pre-incremented past "< pivot", right index pre-decremented past
 *     "> pivot" while not crossed), pivot swapped into place;
 *   - the larger side is deferred on the stack only if it has more than 16 gaps; a side with <= 16 gaps
 *     is left unsorted;  the smaller side is iterated on under the same rule;
 *   - when the depth budget hits zero the range is comb-sorted (shrink 1.2473..., gap 9/10 -> 11,
 *     finishing insertion sort when the last gap was not 1);
 *   - one insertion sort over the whole array finishes the job.
 */
#ifndef BB_SORT_H
#define BB_SORT_H
#include <stddef.h>
#include <stdlib.h>

#define BB_SORT_DEFINE(SCOPE, name, T, LT)                                                        \
	static inline void name##_ins(T *a, long lo, long hi) /* [lo,hi) */                             \
	{                                                                                               \
		long p, q;                                                                                  \
		for (p = lo + 1; p < hi; ++p)                                                               \
			for (q = p; q > lo && LT(a[q], a[q - 1]); --q) { T x = a[q]; a[q] = a[q - 1]; a[q - 1] = x; } \
	}                                                                                               \
	static void name##_comb(T *a, size_t n)                                                         \
	{                                                                                               \
		const double shrink = 1.2473309501039786540366528676643;                                    \
		size_t gap = n, p;                                                                          \
		int moved;                                                                                  \
		do {                                                                                        \
			if (gap > 2) { gap = (size_t)(gap / shrink); if (gap == 9 || gap == 10) gap = 11; }     \
			moved = 0;                                                                              \
			for (p = 0; p + gap < n; ++p)                                                           \
				if (LT(a[p + gap], a[p])) { T x = a[p]; a[p] = a[p + gap]; a[p + gap] = x; moved = 1; } \
		} while (moved || gap > 2);                                                                 \
		if (gap != 1) name##_ins(a, 0, (long)n);                                                    \
	}                                                                                               \
	SCOPE void name(size_t n, T *a)                                                                 \
	{                                                                                               \
		struct { long lo, hi; int depth; } *stk;                                                    \
		long lo, hi, i, j, k, top = 0;                                                              \
		int d;                                                                                      \
		if (n < 1) return;                                                                          \
		if (n == 2) { if (LT(a[1], a[0])) { T x = a[0]; a[0] = a[1]; a[1] = x; } return; }          \
		for (d = 2; (1ul << d) < n; ++d) {}                                                         \
		stk = malloc(sizeof(*stk) * (sizeof(size_t) * d + 2));                                      \
		lo = 0; hi = (long)n - 1; d <<= 1;                                                          \
		for (;;) {                                                                                  \
			if (lo < hi) {                                                                          \
				T piv;                                                                              \
				if (--d == 0) { name##_comb(a + lo, (size_t)(hi - lo + 1)); hi = lo; continue; }    \
				i = lo; j = hi; k = i + ((j - i) >> 1) + 1;                                         \
				if (LT(a[k], a[i])) { if (LT(a[k], a[j])) k = j; }                                  \
				else k = LT(a[j], a[i]) ? i : j;                                                    \
				piv = a[k];                                                                         \
				if (k != hi) { T x = a[k]; a[k] = a[hi]; a[hi] = x; }                               \
				for (;;) {                                                                          \
					do ++i; while (LT(a[i], piv));                                                  \
					do --j; while (i <= j && LT(piv, a[j]));                                        \
					if (j <= i) break;                                                              \
					{ T x = a[i]; a[i] = a[j]; a[j] = x; }                                          \
				}                                                                                   \
				{ T x = a[i]; a[i] = a[hi]; a[hi] = x; }                                            \
				if (i - lo > hi - i) {                                                              \
					if (i - lo > 16) { stk[top].lo = lo; stk[top].hi = i - 1; stk[top].depth = d; ++top; } \
					lo = hi - i > 16 ? i + 1 : hi;                                                  \
				} else {                                                                            \
					if (hi - i > 16) { stk[top].lo = i + 1; stk[top].hi = hi; stk[top].depth = d; ++top; } \
					hi = i - lo > 16 ? i - 1 : lo;                                                  \
				}                                                                                   \
			} else if (top == 0) {                                                                  \
				free(stk);                                                                          \
				name##_ins(a, 0, (long)n);                                                          \
				return;                                                                             \
			} else { --top; lo = stk[top].lo; hi = stk[top].hi; d = stk[top].depth; }               \
		}                                                                                           \
	}
#endif
