/* bwa_b200.h -- C ABI of the B200-native BWA-MEM seed-and-extend path.
 *
 * Part 1 re-declares, layout-for-layout, the reference's library interface for this path so that a
 * caller written against libbwa.a (fastmap.c:process, example.c) can link against libbwa_b200.so
 * unchanged.  Each declaration cites the reference interface it replaces.
 *
 * Part 2 (bwa_b200_dev.h) declares the device-batch entry points the host glue calls; they are
 * implemented by hand-written sm_100a CUDA kernels and have no CPU implementation in the product.
 */
#ifndef BWA_B200_H
#define BWA_B200_H

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- FM-index (reference: bwt.h:46-66) */
typedef uint64_t bwtint_t;

typedef struct {                 /* bwt.h:48-60, sizeof == 1120 */
	bwtint_t primary;            /* row of the removed '$' */
	bwtint_t L2[5];              /* cumulative symbol counts */
	bwtint_t seq_len;            /* length of fwd+revcomp text */
	bwtint_t bwt_size;           /* number of uint32 words in bwt[] */
	uint32_t *bwt;               /* interleaved Occ + 2-bit BWT: one 64-byte block per 128 symbols */
	uint32_t cnt_table[256];
	int sa_intv;
	bwtint_t n_sa;
	bwtint_t *sa;
} bwt_t;

typedef struct { bwtint_t x[3], info; } bwtintv_t;        /* bwt.h:62-64 */
typedef struct { size_t n, m; bwtintv_t *a; } bwtintv_v;  /* bwt.h:66 */

/* ---------------------------------------------------------------- reference metadata (bntseq.h:41-64) */
typedef struct { int64_t offset; int32_t len; int32_t n_ambs; uint32_t gi; int32_t is_alt; char *name, *anno; } bntann1_t;
typedef struct { int64_t offset; int32_t len; char amb; } bntamb1_t;
typedef struct {
	int64_t l_pac;
	int32_t n_seqs;
	uint32_t seed;
	bntann1_t *anns;
	int32_t n_holes;
	bntamb1_t *ambs;
	FILE *fp_pac;
} bntseq_t;

/* ---------------------------------------------------------------- index + read batch (bwa.h:48-61) */
#define BWA_IDX_BWT 0x1
#define BWA_IDX_BNS 0x2
#define BWA_IDX_PAC 0x4
#define BWA_IDX_ALL 0x7

typedef struct {                 /* bwa.h:48-56 */
	bwt_t *bwt;
	bntseq_t *bns;
	uint8_t *pac;
	int is_shm;
	int64_t l_mem;
	uint8_t *mem;
} bwaidx_t;

typedef struct {                 /* bwa.h:58-61 */
	int l_seq, id;
	char *name, *comment, *seq, *qual, *sam;
} bseq1_t;

extern int bwa_verbose;          /* bwa.c:42 */
extern char bwa_rg_id[256];      /* bwa.c:44 */
extern char *bwa_pg;             /* bwa.c:45 */

bwaidx_t *bwa_idx_load(const char *hint, int which);                 /* bwa.h:84,  bwa.c:318 */
void bwa_idx_destroy(bwaidx_t *idx);                                 /* bwa.h:85,  bwa.c:323 */
void bwa_fill_scmat(int a, int b, int8_t mat[25]);                   /* bwa.h:72,  bwa.c:136 */
void bwa_print_sam_hdr(const bntseq_t *bns, const char *hdr_line);   /* bwa.h:89,  bwa.c:407 */
char *bwa_set_rg(const char *s);                                     /* bwa.h:90,  bwa.c:460 */
char *bwa_insert_header(const char *s, char *hdr);                   /* bwa.h:91,  bwa.c:489 */
bseq1_t *bseq_read(int chunk_size, int *n_, void *ks1_, void *ks2_); /* bwa.h:69,  bwa.c:79  (ks*: bb_fq_t*) */
void bseq_classify(int n, bseq1_t *seqs, int m[2], bseq1_t *sep[2]); /* bwa.h:70,  bwa.c:114 */

/* ---------------------------------------------------------------- BWA-MEM (bwamem.h:40-211) */
#define MEM_MAPQ_COEF 30.0
#define MEM_MAPQ_MAX  60

#define MEM_F_PE        0x2
#define MEM_F_NOPAIRING 0x4
#define MEM_F_ALL       0x8
#define MEM_F_NO_MULTI  0x10
#define MEM_F_NO_RESCUE 0x20
#define MEM_F_REF_HDR   0x100
#define MEM_F_SOFTCLIP  0x200
#define MEM_F_SMARTPE   0x400
#define MEM_F_PRIMARY5  0x800
#define MEM_F_KEEP_SUPP_MAPQ 0x1000
#define MEM_F_XB        0x2000

typedef struct {                 /* bwamem.h:52-84, sizeof == 168 */
	int a, b;
	int o_del, e_del;
	int o_ins, e_ins;
	int pen_unpaired;
	int pen_clip5, pen_clip3;
	int w;
	int zdrop;
	uint64_t max_mem_intv;
	int T;
	int flag;
	int min_seed_len;
	int min_chain_weight;
	int max_chain_extend;
	float split_factor;
	int split_width;
	int max_occ;
	int max_chain_gap;
	int n_threads;
	int chunk_size;
	float mask_level;
	float drop_ratio;
	float XA_drop_ratio;
	float mask_level_redun;
	float mapQ_coef_len;
	int mapQ_coef_fac;
	int max_ins;
	int max_matesw;
	int max_XA_hits, max_XA_hits_alt;
	int8_t mat[25];
} mem_opt_t;

typedef struct {                 /* bwamem.h:86-104, sizeof == 88 */
	int64_t rb, re;
	int qb, qe;
	int rid;
	int score;
	int truesc;
	int sub;
	int alt_sc;
	int csub;
	int sub_n;
	int w;
	int seedcov;
	int secondary;
	int secondary_all;
	int seedlen0;
	int n_comp:30, is_alt:2;
	float frac_rep;
	uint64_t hash;
} mem_alnreg_t;

typedef struct { size_t n, m; mem_alnreg_t *a; } mem_alnreg_v;

typedef struct {                 /* bwamem.h:108-112 */
	int low, high;
	int failed;
	double avg, std;
} mem_pestat_t;

typedef struct {                 /* bwamem.h:114-124, sizeof == 56 */
	int64_t pos;
	int rid;
	int flag;
	uint32_t is_rev:1, is_alt:1, mapq:8, NM:22;
	int n_cigar;
	uint32_t *cigar;             /* len<<4|op, MIDSH; the MD string follows the last op in the same allocation */
	char *XA;
	int score, sub, alt_sc;
} mem_aln_t;

/* the layouts above are the reference's ([measured sizeof] in SURVEY.md 8b); a caller compiled against bwamem.h/bwa.h/bwt.h
 * passes these structs by pointer and by value */
#if defined(__cplusplus)
#define BB_SIZE_CHECK(t, n) static_assert(sizeof(t) == (n), "layout of " #t " differs from the reference")
#else
#define BB_SIZE_CHECK(t, n) _Static_assert(sizeof(t) == (n), "layout of " #t " differs from the reference")
#endif
BB_SIZE_CHECK(bwt_t, 1120); BB_SIZE_CHECK(bwtintv_t, 32); BB_SIZE_CHECK(bntann1_t, 40); BB_SIZE_CHECK(bntseq_t, 48);
BB_SIZE_CHECK(bwaidx_t, 48); BB_SIZE_CHECK(bseq1_t, 48); BB_SIZE_CHECK(mem_opt_t, 168); BB_SIZE_CHECK(mem_alnreg_t, 88);
BB_SIZE_CHECK(mem_alnreg_v, 24); BB_SIZE_CHECK(mem_pestat_t, 32); BB_SIZE_CHECK(mem_aln_t, 56);

mem_opt_t *mem_opt_init(void);                                       /* bwamem.h:136, bwamem.c:74 */

/* The drop-in boundary (bwamem.h:161, bwamem.c:1235).  Same contract as the reference:
 * reads seqs[i].{l_seq,seq,name,qual,comment}, rewrites seqs[i].seq in place to 0..4 codes,
 * mallocs seqs[i].sam (caller frees); PE reads interleaved 2i/2i+1 when MEM_F_PE is set;
 * n_processed seeds the hash tie-breaks; pes0 != NULL fixes the insert-size model.
 * All failures are fatal (exit), as in the reference -- including "no CUDA device".
 * The FM-index / pac given here are uploaded to HBM on first use and cached by pointer. */
void mem_process_seqs(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac,
                      int64_t n_processed, int n, bseq1_t *seqs, const mem_pestat_t *pes0);

/* bwamem.h:178 / bwamem_extra.c:102 : regions for one read (no CIGAR), primary marked with a random id */
mem_alnreg_v mem_align1(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac, int l_seq, const char *seq);
/* bwamem.h:192 / bwamem.c:1119 : CIGAR, strand, MAPQ and forward position for one region */
mem_aln_t mem_reg2aln(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, int l_seq, const char *seq, const mem_alnreg_t *ar);
/* bwamem.h:207 / bwamem_pair.c:72 */
void mem_pestat(const mem_opt_t *opt, int64_t l_pac, int n, const mem_alnreg_v *regs, mem_pestat_t pes[4]);

/* command-line entry (fastmap.c:141 main_mem; main.c:87) */
int main_mem(int argc, char *argv[]);
/* extension for multi-GPU launchers (bwa_b200/multi.py): main_mem uses this index, which the caller has already made
 * resident on the GPU (bb_device_adopt), instead of loading argv's prefix; with BWA_B200_RANK / BWA_B200_WORLD set it
 * aligns only the batches b with b % world == rank */
void bb_cli_set_index(bwaidx_t *idx);

/* Striped ingest for multi-GPU runs (no counterpart in the reference, whose step 0 is one reader thread, fastmap.c:64-123).
 * bb_fq_scan_stripe lists the records whose header starts in the bytes [beg, end) of a strictly laid out uncompressed
 * FASTQ/FASTA file: sequence length and byte offset of each.  BB_SCAN_UNFIT: the file is not of that kind (gzip, wrapped
 * lines, truncated): the launcher then lets every rank parse the whole input as before.
 * bb_cli_set_plan hands main_mem the batches THIS process aligns: batch number, reads before it, and the byte range of the
 * batch in each input file; main_mem then reads only those ranges.  n_batches_total: batches of the whole run. */
#define BB_SCAN_UNFIT 1
typedef struct { int64_t n; int32_t *len; int64_t *off; } bb_fqstripe_t;
int64_t bb_fq_plain_size(const char *fn);
int bb_fq_scan_stripe(const char *fn, int64_t beg, int64_t end, bb_fqstripe_t *out);
void bb_fq_stripe_free(bb_fqstripe_t *s);
typedef struct { int64_t no, n_before, beg1, end1, beg2, end2; } bb_planned_batch_t;
void bb_cli_set_plan(const bb_planned_batch_t *mine, int64_t n_mine, int64_t n_batches_total);

#ifdef __cplusplus
}
#endif
#endif
