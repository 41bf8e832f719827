/* TEST INFRASTRUCTURE ONLY -- declarations of the CPU restatement (oracle_fm.c, oracle_sw.c). */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#include <stddef.h>
#include "bwa_b200.h"      /* reference-layout structs: bwt_t, bwtintv_t */

typedef struct { size_t n, m; bwtintv_t *a; } orc_intv_v;
typedef struct { size_t n, m; uint32_t *a; } orc_u32_v;
typedef struct { size_t l, m; char *s; } orc_str_t;
typedef struct { int a, b, o_del, e_del, o_ins, e_ins, w, zdrop, pen_clip5, pen_clip3; int8_t mat[25]; } orc_swpar_t; /* == bwag_sw_par_t */
typedef struct { int64_t rbeg; int32_t qbeg; uint32_t len; } orc_xseed_t;                                            /* == bwag_xseed_t */
typedef struct { int64_t rb, re; int32_t qb, qe, score, truesc, w, seedcov, seedlen0, chain; } orc_xreg_t;           /* == bwag_xreg_t */

void orc_occ4(const bwt_t *b, uint64_t k, uint64_t cnt[4]);
uint64_t orc_occ(const bwt_t *b, uint64_t k, int c);
void orc_extend(const bwt_t *b, const bwtintv_t *ik, bwtintv_t ok[4], int is_back, uint64_t *touches);
int orc_smem1(const bwt_t *b, int len, const uint8_t *q, int x, int min_intv, orc_intv_v *mem, uint64_t *touches);
int orc_seed_strategy1(const bwt_t *b, int len, const uint8_t *q, int x, int min_len, int max_intv, bwtintv_t *mem, uint64_t *touches);
void orc_collect_intv(const bwt_t *b, int len, const uint8_t *seq, int min_seed_len, int split_len, int split_width, uint64_t max_mem_intv, orc_intv_v *out, uint64_t *touches);
uint64_t orc_sa(const bwt_t *b, uint64_t k, uint64_t *steps);

int orc_extend_sw(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                  int w, int end_bonus, int zdrop, int h0, int *qle, int *tle, int *gtle, int *gscore, int *max_off, uint64_t *cells);
int orc_global(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
               int w, orc_u32_v *cigar, uint64_t *cells);
uint8_t *orc_get_seq(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, int64_t *len);
int orc_gen_cigar(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins, int w_, int64_t l_pac, const uint8_t *pac,
                  int l_query, const uint8_t *query, int64_t rb, int64_t re, int want_cigar,
                  int *score, orc_u32_v *cigar, int *NM, orc_str_t *md, uint64_t *cells);
int orc_reg2aln_core(const int8_t mat[25], int a, int o_del, int e_del, int o_ins, int e_ins, int opt_w, int64_t l_pac, const uint8_t *pac,
                     int l_query, const uint8_t *query, int64_t rb, int64_t re, int w2, int truesc,
                     int *score, orc_u32_v *cigar, int *NM, orc_str_t *md, uint64_t *cells);
void orc_chain2aln(const orc_swpar_t *p, int64_t l_pac, const uint8_t *pac, int l_query, const uint8_t *query,
                   int64_t rmax0, int64_t rmax1, int n_seeds, const orc_xseed_t *seeds, int chain_idx,
                   orc_xreg_t *regs, int *n_regs, uint64_t *cells);
#endif
