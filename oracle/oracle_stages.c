/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * oracle_stages.c -- the device-batch ABI of include/bwa_b200_dev.h answered by the CPU restatement
 * (oracle_fm.c / oracle_sw.c).  Two uses, both in tests/:
 *   1. checker: tests call the CUDA library and this library with the same inputs and compare the
 *      result buffers bit for bit;
 *   2. host-glue tests without a GPU: tests/_build/bwa-b200-oracle links the host glue against this
 *      file so that chaining/SAM logic can be diffed against the reference `bwa mem` on a CPU box.
 * The product library (libbwa_b200.so) contains none of this and fails loudly without a CUDA device.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include "bwa_b200_dev.h"
#include "oracle.h"

struct bwag_ctx {
	const bwt_t *bwt;
	int64_t l_pac;
	const uint8_t *pac;
	bwag_stats_t st;
};

struct bwag_batch {
	bwag_ctx_t *ctx;
	int n;
	const uint8_t *codes;
	const int64_t *off;
	/* stage outputs */
	int64_t *intv_off, *seed_off, *rbeg; int32_t *intv_n;
	bwtintv_t *intv;
	int32_t *n_regs;
	bwag_xreg_t *regs;
	bwag_gres_t *gres;
	uint32_t *cig;
	char *md;
};

static int n_threads(void)
{
	const char *e = getenv("BWA_B200_ORACLE_THREADS");
	int n = e ? atoi(e) : 8;
	return n < 1 ? 1 : n;
}

typedef struct { void (*fn)(void *, long); void *d; long n, next; } pf_t;
static void *pf_main(void *a)
{
	pf_t *p = a;
	for (;;) {
		long b = __sync_fetch_and_add(&p->next, 64), e, i;
		if (b >= p->n) break;
		e = b + 64 < p->n ? b + 64 : p->n;
		for (i = b; i < e; ++i) p->fn(p->d, i);
	}
	return 0;
}
static void pfor(void (*fn)(void *, long), void *d, long n)
{
	pf_t p = { fn, d, n, 0 };
	int t, nt = n_threads();
	pthread_t th[64];
	if (nt > 64) nt = 64;
	if (n < 256) nt = 1;
	for (t = 1; t < nt; ++t) pthread_create(&th[t], 0, pf_main, &p);
	pf_main(&p);
	for (t = 1; t < nt; ++t) pthread_join(th[t], 0);
}

void *bwag_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void bwag_host_free(void *p) { free(p); }
const char *bwag_last_error(void) { return "oracle (CPU) stages"; }
size_t bwag_blob_bytes(const bwt_t *bwt, int64_t l_pac) { (void)bwt; (void)l_pac; return 0; }
int bwag_blob_fill(int device, void *d_blob, const bwt_t *bwt, int64_t l_pac, const uint8_t *pac) { (void)device; (void)d_blob; (void)bwt; (void)l_pac; (void)pac; return 1; }
bwag_ctx_t *bwag_ctx_from_blob(int device, void *d_blob, int own) { (void)device; (void)d_blob; (void)own; return 0; }
int bwag_ctx_densify_sa(bwag_ctx_t *ctx, int intv) { (void)ctx; (void)intv; return 0; }
int bwag_ctx_build_ktab(bwag_ctx_t *ctx, int depth) { (void)ctx; (void)depth; return 0; }
void bwag_ctx_baseline(bwag_ctx_t *ctx, int on) { (void)ctx; (void)on; }
int bwag_is_emulator(void) { return 2; }
int bwag_ctx_verify(bwag_ctx_t *ctx, uint64_t first, uint64_t stride, uint64_t out[4]) { (void)ctx; (void)first; (void)stride; (void)out; return BWAG_UNSUPPORTED; }
int bwag_ctx_export(bwag_ctx_t *ctx, const char *path) { (void)ctx; (void)path; return BWAG_UNSUPPORTED; }
bwag_ctx_t *bwag_ctx_import(const char *path, int64_t l_pac) { (void)path; (void)l_pac; return 0; }
void bwag_ctx_unexport(const char *path) { (void)path; }
/* stage 4 (the post-processing on the device) has no oracle restatement: the host-side post-processing IS the checker for it */
int bwag_ctx_set_contigs(bwag_ctx_t *ctx, int n_seqs, const int64_t *offset, const int32_t *len, const uint8_t *is_alt, const char *const *names) { (void)ctx; (void)n_seqs; (void)offset; (void)len; (void)is_alt; (void)names; return BWAG_UNSUPPORTED; }
int bwag_localsw(bwag_batch_t *b, const bwag_sw_par_t *par, int n_tasks, const bwag_swtask_t *tasks, const uint8_t *pool, size_t pool_bytes, const bwag_swres_t **out) { (void)b; (void)par; (void)n_tasks; (void)tasks; (void)pool; (void)pool_bytes; (void)out; return BWAG_UNSUPPORTED; }
int bwag_fetch_cregs(bwag_batch_t *b, int n_sel, const int32_t *sel, bwag_cregs_t *out) { (void)b; (void)n_sel; (void)sel; (void)out; return BWAG_UNSUPPORTED; }
int bwag_tail_regs(bwag_batch_t *b, const mem_opt_t *opt, const bwag_sw_par_t *sp, const uint64_t **pe_is, const uint8_t **cflag) { (void)b; (void)opt; (void)sp; (void)pe_is; (void)cflag; return BWAG_UNSUPPORTED; }
int bwag_tail_sam(bwag_batch_t *b, const mem_opt_t *opt, const mem_pestat_t pes[4], const double *const pair_tab[4], const double *log_tab, int64_t n_processed, const char *rg_id, bwag_sam_t *out) { (void)b; (void)opt; (void)pes; (void)pair_tab; (void)log_tab; (void)n_processed; (void)rg_id; (void)out; return BWAG_UNSUPPORTED; }

bwag_ctx_t *bwag_ctx_create(int device, const bwt_t *bwt, int64_t l_pac, const uint8_t *pac)
{
	bwag_ctx_t *c = calloc(1, sizeof(*c));
	(void)device;
	c->bwt = bwt; c->l_pac = l_pac; c->pac = pac;
	return c;
}
void bwag_ctx_destroy(bwag_ctx_t *c) { free(c); }
void bwag_stats_get(bwag_ctx_t *c, bwag_stats_t *s) { *s = c->st; }
void bwag_stats_reset(bwag_ctx_t *c) { memset(&c->st, 0, sizeof(c->st)); }

bwag_batch_t *bwag_batch_begin(bwag_ctx_t *ctx, int n, const uint8_t *codes, const int64_t *off)
{
	bwag_batch_t *b = calloc(1, sizeof(*b));
	b->ctx = ctx; b->n = n; b->codes = codes; b->off = off;
	return b;
}
void bwag_batch_end(bwag_batch_t *b)
{
	if (!b) return;
	free(b->intv_off); free(b->intv_n); free(b->seed_off); free(b->rbeg); free(b->intv); free(b->n_regs); free(b->regs); free(b->gres); free(b->cig); free(b->md);
	free(b);
}

/* ---------------------------------------------------------------- stage 1 */
typedef struct { bwag_batch_t *b; const bwag_seed_par_t *par; orc_intv_v *per; uint64_t *touch; } s1_t;
static void s1_read(void *d, long i)
{
	s1_t *s = d;
	bwag_batch_t *b = s->b;
	orc_collect_intv(b->ctx->bwt, (int)(b->off[i + 1] - b->off[i]), b->codes + b->off[i], s->par->min_seed_len, s->par->split_len, s->par->split_width,
	                 s->par->max_mem_intv, &s->per[i], &s->touch[i]);
}
typedef struct { bwag_batch_t *b; const bwag_seed_par_t *par; uint64_t *steps; int64_t *row; } s2_t;
static void s2_seed(void *d, long i)
{
	s2_t *s = d;
	s->b->rbeg[i] = (int64_t)orc_sa(s->b->ctx->bwt, (uint64_t)s->row[i], &s->steps[i]);
}

int bwag_seed(bwag_batch_t *b, const bwag_seed_par_t *par, bwag_seeds_t *out)
{
	s1_t s1 = { b, par, calloc(b->n + 1, sizeof(orc_intv_v)), calloc(b->n + 1, sizeof(uint64_t)) };
	s2_t s2;
	int64_t i, n_intv = 0, n_seeds = 0, k;
	uint64_t *steps;
	pfor(s1_read, &s1, b->n);
	free(b->intv_off); free(b->intv_n); free(b->seed_off); free(b->rbeg); free(b->intv);
	b->intv_off = malloc(sizeof(int64_t) * (b->n + 1));
	b->intv_n = malloc(sizeof(int32_t) * (b->n + 1));
	for (i = 0; i < b->n; ++i) { b->intv_off[i] = n_intv; b->intv_n[i] = (int32_t)s1.per[i].n; n_intv += (int64_t)s1.per[i].n; b->ctx->st.occ_touches += s1.touch[i]; }
	b->intv_off[b->n] = n_intv;
	b->intv = malloc(sizeof(bwtintv_t) * (n_intv + 1));
	b->seed_off = malloc(sizeof(int64_t) * (n_intv + 1));
	for (i = 0, k = 0; i < b->n; ++i) {
		size_t j;
		for (j = 0; j < s1.per[i].n; ++j, ++k) {
			const bwtintv_t *p = &s1.per[i].a[j];
			int64_t cnt = (int64_t)p->x[2] < par->max_occ ? (int64_t)p->x[2] : par->max_occ; /* bwamem.c:304-305 */
			b->intv[k] = *p;
			b->seed_off[k] = n_seeds;
			n_seeds += cnt;
		}
		free(s1.per[i].a);
	}
	b->seed_off[n_intv] = n_seeds;
	b->rbeg = malloc(sizeof(int64_t) * (n_seeds + 1));
	steps = calloc(n_seeds + 1, sizeof(uint64_t));
	s2.b = b; s2.par = par; s2.steps = steps; s2.row = malloc(sizeof(int64_t) * (n_seeds + 1));
	for (k = 0; k < n_intv; ++k) {
		const bwtintv_t *p = &b->intv[k];
		int64_t step = (int64_t)p->x[2] > par->max_occ ? (int64_t)p->x[2] / par->max_occ : 1, cnt = b->seed_off[k + 1] - b->seed_off[k], c;
		for (c = 0; c < cnt; ++c) s2.row[b->seed_off[k] + c] = (int64_t)p->x[0] + c * step;
	}
	pfor(s2_seed, &s2, n_seeds);
	for (i = 0; i < n_seeds; ++i) { b->ctx->st.sa_touches += steps[i]; b->ctx->st.sa_touches_algo += steps[i]; }
	free(steps); free(s2.row); free(s1.per); free(s1.touch);
	if (!out) return 0;
	out->intv_beg = b->intv_off; out->intv_n = b->intv_n; out->intv = b->intv; out->seed_beg = b->seed_off; out->rbeg = b->rbeg; out->n_intv = n_intv; out->n_seeds = n_seeds;
	return 0;
}

/* ---------------------------------------------------------------- stage 2 */
typedef struct { bwag_batch_t *b; const bwag_sw_par_t *par; const int32_t *chain_off; const bwag_xchain_t *chains; const bwag_xseed_t *seeds; uint64_t *cells; } s3_t;
static void s3_read(void *d, long i)
{
	s3_t *s = d;
	bwag_batch_t *b = s->b;
	int c, n = 0;
	bwag_xreg_t *out;
	b->n_regs[i] = 0;
	if (s->chain_off[i] == s->chain_off[i + 1]) return;
	out = b->regs + s->chains[s->chain_off[i]].seed_off;
	for (c = s->chain_off[i]; c < s->chain_off[i + 1]; ++c) {
		const bwag_xchain_t *ch = &s->chains[c];
		orc_chain2aln((const orc_swpar_t *)s->par, b->ctx->l_pac, b->ctx->pac, (int)(b->off[i + 1] - b->off[i]), b->codes + b->off[i],
		              ch->rmax0, ch->rmax1, ch->n_seeds, (const orc_xseed_t *)(s->seeds + ch->seed_off), c - s->chain_off[i], (orc_xreg_t *)out, &n, &s->cells[i]);
	}
	b->n_regs[i] = n;
}

int bwag_extend(bwag_batch_t *b, const bwag_sw_par_t *par, const int32_t *chain_off, const bwag_xchain_t *chains, int64_t n_seeds, const bwag_xseed_t *seeds, bwag_regs_t *out)
{
	uint64_t *cells = calloc(b->n + 1, sizeof(uint64_t));
	s3_t s = { b, par, chain_off, chains, seeds, cells };
	int i;
	free(b->n_regs); free(b->regs);
	b->n_regs = calloc(b->n + 1, sizeof(int32_t));
	b->regs = calloc(n_seeds + 1, sizeof(bwag_xreg_t));
	pfor(s3_read, &s, b->n);
	for (i = 0; i < b->n; ++i) b->ctx->st.ext_cells += cells[i];
	free(cells);
	out->n_regs = b->n_regs; out->regs = b->regs;
	return 0;
}

/* the oracle has no device-side chaining: the host glue then chains on the host (bb_chain.c), which is what the
 * CPU tests of the host glue are about */
int bwag_chain_extend(bwag_batch_t *b, const bwag_chain_par_t *cp, const bwag_sw_par_t *sp, const bwag_contigs_t *ctg, bwag_cregs_t *out)
{
	(void)b; (void)cp; (void)sp; (void)ctg; (void)out;
	return BWAG_UNSUPPORTED;
}

/* ---------------------------------------------------------------- stage 3 */
typedef struct { bwag_batch_t *b; const bwag_sw_par_t *par; const bwag_gtask_t *t; orc_u32_v *cig; orc_str_t *md; uint64_t *cells; } s4_t;
static void s4_task(void *d, long i)
{
	s4_t *s = d;
	bwag_batch_t *b = s->b;
	const bwag_gtask_t *t = &s->t[i];
	const bwag_sw_par_t *p = s->par;
	const uint8_t *q = b->codes + b->off[t->read] + t->qb;
	bwag_gres_t *r = &b->gres[i];
	int score = 0, NM = -1;
	memset(r, 0, sizeof(*r));
	if (t->mode == BWAG_G_SCORE)
		orc_gen_cigar(p->mat, p->o_del, p->e_del, p->o_ins, p->e_ins, t->w, b->ctx->l_pac, b->ctx->pac, t->qe - t->qb, q, t->rb, t->re, 0, &score, 0, 0, 0, &s->cells[i]);
	else
		orc_reg2aln_core(p->mat, p->a, p->o_del, p->e_del, p->o_ins, p->e_ins, p->w, b->ctx->l_pac, b->ctx->pac, t->qe - t->qb, q, t->rb, t->re, t->w, t->truesc,
		                 &score, &s->cig[i], &NM, &s->md[i], &s->cells[i]);
	r->score = score; r->NM = NM; r->n_cigar = (int32_t)s->cig[i].n;
	r->l_md = t->mode == BWAG_G_SCORE ? 0 : (int32_t)s->md[i].l + 1;
}

int bwag_global(bwag_batch_t *b, const bwag_sw_par_t *par, int n_tasks, const bwag_gtask_t *tasks, bwag_galn_t *out)
{
	uint64_t *cells = calloc(n_tasks + 1, sizeof(uint64_t));
	s4_t s = { b, par, tasks, calloc(n_tasks + 1, sizeof(orc_u32_v)), calloc(n_tasks + 1, sizeof(orc_str_t)), cells };
	int64_t i, nc = 0, nm = 0;
	free(b->gres); free(b->cig); free(b->md);
	b->gres = calloc(n_tasks + 1, sizeof(bwag_gres_t));
	pfor(s4_task, &s, n_tasks);
	for (i = 0; i < n_tasks; ++i) { b->gres[i].cigar_off = nc; b->gres[i].md_off = nm; nc += b->gres[i].n_cigar; nm += b->gres[i].l_md; }
	b->cig = malloc(4 * (nc + 1)); b->md = malloc(nm + 1);
	for (i = 0; i < n_tasks; ++i) {
		if (b->gres[i].n_cigar) memcpy(b->cig + b->gres[i].cigar_off, s.cig[i].a, 4 * (size_t)b->gres[i].n_cigar);
		if (b->gres[i].l_md) { if (s.md[i].s) memcpy(b->md + b->gres[i].md_off, s.md[i].s, b->gres[i].l_md); else b->md[b->gres[i].md_off] = 0; }
		free(s.cig[i].a); free(s.md[i].s);
	}
	for (i = 0; i < n_tasks; ++i) b->ctx->st.glb_cells += cells[i];
	free(cells); free(s.cig); free(s.md);
	out->res = b->gres; out->cigar = b->cig; out->md = b->md;
	return 0;
}
