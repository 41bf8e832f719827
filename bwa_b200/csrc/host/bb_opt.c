/* bb_opt.c -- option defaults, scoring matrix, SAM header helpers (reference bwamem.c:74-110,
 * bwa.c:114-145, 407-502). */
#include <math.h>
#include "bb_host.h"

mem_opt_t *mem_opt_init(void)
{
	mem_opt_t *o = bb_calloc(1, sizeof(mem_opt_t));
	o->a = 1; o->b = 4;
	o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1;
	o->w = 100; o->T = 30; o->zdrop = 100;
	o->pen_unpaired = 17; o->pen_clip5 = o->pen_clip3 = 5;
	o->max_mem_intv = 20;
	o->min_seed_len = 19; o->split_width = 10; o->max_occ = 500;
	o->max_chain_gap = 10000; o->max_ins = 10000;
	o->mask_level = 0.50f; o->drop_ratio = 0.50f; o->XA_drop_ratio = 0.80f;
	o->split_factor = 1.5f;
	o->chunk_size = 10000000; o->n_threads = 1;
	o->max_XA_hits = 5; o->max_XA_hits_alt = 200; o->max_matesw = 50;
	o->mask_level_redun = 0.95f;
	o->min_chain_weight = 0; o->max_chain_extend = 1 << 30;
	o->mapQ_coef_len = 50;
	o->mapQ_coef_fac = (int)log(o->mapQ_coef_len); /* an int in the reference: log(50)=3.91 -> 3 (bwamem.h:79) */
	bwa_fill_scmat(o->a, o->b, o->mat);
	return o;
}

void bwa_fill_scmat(int a, int b, int8_t mat[25])
{
	int i, j;
	for (i = 0; i < 5; ++i)
		for (j = 0; j < 5; ++j)
			mat[i * 5 + j] = (int8_t)((i == 4 || j == 4) ? -1 : i == j ? a : -b);
}

void bseq_classify(int n, bseq1_t *seqs, int m[2], bseq1_t *sep[2])
{
	BB_VEC(bseq1_t) a[2] = {{0, 0, 0}, {0, 0, 0}};
	int i, has_last = 1;
	for (i = 1; i < n; ++i) {
		if (has_last) {
			if (strcmp(seqs[i].name, seqs[i - 1].name) == 0) { bb_vec_push(a[1], seqs[i - 1]); bb_vec_push(a[1], seqs[i]); has_last = 0; }
			else bb_vec_push(a[0], seqs[i - 1]);
		} else has_last = 1;
	}
	if (has_last && n > 0) bb_vec_push(a[0], seqs[n - 1]);
	sep[0] = a[0].a; m[0] = (int)a[0].n;
	sep[1] = a[1].a; m[1] = (int)a[1].n;
}

static int count_tag(const char *hdr, const char *tag)
{
	int n = 0;
	const char *p = hdr;
	while ((p = strstr(p, tag)) != 0) { if (p == hdr || p[-1] == '\n') ++n; p += 4; }
	return n;
}

void bwa_print_sam_hdr(const bntseq_t *bns, const char *hdr_line)
{
	int i, n_HD = 0, n_SQ = 0;
	if (hdr_line) { n_HD = count_tag(hdr_line, "@HD\t"); n_SQ = count_tag(hdr_line, "@SQ\t"); }
	if (n_HD == 0) printf("@HD\tVN:1.5\tSO:unsorted\tGO:query\n");
	else if (bwa_verbose >= 2) fprintf(stderr, "[W::%s] please don't include @HD with option -H. Continue anyway.\n", __func__);
	if (n_SQ == 0) {
		for (i = 0; i < bns->n_seqs; ++i) {
			printf("@SQ\tSN:%s\tLN:%d", bns->anns[i].name, bns->anns[i].len);
			if (bns->anns[i].is_alt) printf("\tAH:*\n"); else putchar('\n');
		}
	} else if (n_SQ != bns->n_seqs && bwa_verbose >= 2)
		fprintf(stderr, "[W::%s] %d @SQ lines provided with -H; %d sequences in the index. Continue anyway.\n", __func__, n_SQ, bns->n_seqs);
	if (hdr_line) printf("%s\n", hdr_line);
	if (bwa_pg) printf("%s\n", bwa_pg);
}

static char *unescape(char *s) /* \t \n \r \\ (bwa.c:441-458) */
{
	char *p, *q;
	for (p = q = s; *p; ++p) {
		if (*p == '\\') {
			++p;
			if (*p == 't') *q++ = '\t';
			else if (*p == 'n') *q++ = '\n';
			else if (*p == 'r') *q++ = '\r';
			else if (*p == '\\') *q++ = '\\';
		} else *q++ = *p;
	}
	*q = 0;
	return s;
}

char *bwa_set_rg(const char *s)
{
	char *p, *q, *r, *rg_line = 0;
	memset(bwa_rg_id, 0, 256);
	if (strstr(s, "@RG") != s) {
		if (bwa_verbose >= 1) fprintf(stderr, "[E::%s] the read group line is not started with @RG\n", __func__);
		return 0;
	}
	if (strstr(s, "\t") != NULL) {
		if (bwa_verbose >= 1) fprintf(stderr, "[E::%s] the read group line contained literal <tab> characters -- replace with escaped tabs: \\t\n", __func__);
		return 0;
	}
	rg_line = bb_strdup(s);
	unescape(rg_line);
	if ((p = strstr(rg_line, "\tID:")) == 0) {
		if (bwa_verbose >= 1) fprintf(stderr, "[E::%s] no ID within the read group line\n", __func__);
		free(rg_line); return 0;
	}
	p += 4;
	for (q = p; *q && *q != '\t' && *q != '\n'; ++q) {}
	if (q - p + 1 > 256) {
		if (bwa_verbose >= 1) fprintf(stderr, "[E::%s] @RG:ID is longer than 255 characters\n", __func__);
		free(rg_line); return 0;
	}
	for (q = p, r = bwa_rg_id; *q && *q != '\t' && *q != '\n'; ++q) *r++ = *q;
	return rg_line;
}

char *bwa_insert_header(const char *s, char *hdr)
{
	size_t len = 0;
	if (s == 0 || s[0] != '@') return hdr;
	if (hdr) {
		len = strlen(hdr);
		hdr = bb_realloc(hdr, len + strlen(s) + 2);
		hdr[len++] = '\n';
		strcpy(hdr + len, s);
	} else hdr = bb_strdup(s);
	unescape(hdr + len);
	return hdr;
}
