/* bb_sam.c -- region -> alignment record -> SAM text.
 *
 * Mirrors the reference's emitter field by field (bwamem.c:818-976, 1033-1079, 1119-1189;
 * bwamem_extra.c:116-172) so that output diffs byte-for-byte.  The one
 * structural difference: the CIGAR/NM/MD of a region is not computed here but looked up in the
 * read's alignment cache, which the batch driver fills from the device global-alignment stage.
 * A first "dry" pass over a read only discovers which alignments it needs.
 */
#include <math.h>
#include <limits.h>
#include <assert.h>
#include "bb_host.h"

static int infer_bw(int l1, int l2, int score, int a, int q, int r) /* bwamem.c:818-825 */
{
	int w;
	if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
	w = (int)(((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.));
	if (w < abs(l1 - l2)) w = abs(l1 - l2);
	return w;
}

static int cigar_ref_len(int n_cigar, const uint32_t *cigar)
{
	int k, l = 0;
	for (k = 0; k < n_cigar; ++k) {
		int op = cigar[k] & 0xf;
		if (op == 0 || op == 2) l += cigar[k] >> 4;
	}
	return l;
}

/* starting band of the band-doubling loop of mem_reg2aln for region ar (bwamem.c:1138-1142) */
int bb_reg2aln_band(const mem_opt_t *opt, const mem_alnreg_t *ar)
{
	int tmp = infer_bw(ar->qe - ar->qb, (int)(ar->re - ar->rb), ar->truesc, opt->a, opt->o_del, opt->e_del);
	int w2 = infer_bw(ar->qe - ar->qb, (int)(ar->re - ar->rb), ar->truesc, opt->a, opt->o_ins, opt->e_ins);
	w2 = w2 > tmp ? w2 : tmp;
	if (w2 > opt->w) w2 = w2 < ar->w ? w2 : ar->w;
	return w2;
}

/* bwamem.c:1119-1189 */
mem_aln_t bb_reg2aln(bb_samctx_t *sc, int l_query, const char *query_, const mem_alnreg_t *ar)
{
	const mem_opt_t *opt = sc->opt;
	const bntseq_t *bns = sc->bns;
	mem_aln_t a;
	int w2, qb, qe, is_rev, l_MD;
	int64_t pos, rb, re;
	const bb_galn_t *g;
	(void)query_;
	memset(&a, 0, sizeof(a));
	if (ar == 0 || ar->rb < 0 || ar->re < 0) {
		a.rid = -1; a.pos = -1; a.flag |= 0x4;
		return a;
	}
	qb = ar->qb; qe = ar->qe; rb = ar->rb; re = ar->re;
	a.mapq = ar->secondary < 0 ? bb_approx_mapq_se(opt, ar) : 0;
	if (ar->secondary >= 0) a.flag |= 0x100;
	w2 = bb_reg2aln_band(opt, ar);
	/* the band-doubling loop (bwamem.c:1144-1152) runs on the device; w2 is its starting band */
	g = bb_gcache_get(sc->gc, BWAG_G_REG2ALN, qb, qe, rb, re, w2, ar->truesc);
	pos = bb_depos(bns, rb < bns->l_pac ? rb : re - 1, &is_rev);
	a.is_rev = is_rev;
	if (g) {
		l_MD = g->l_md;
		a.n_cigar = g->n_cigar;
		{
			const int words = g->n_cigar + 2 + ((l_MD + 3) >> 2);
			if (sc->scratch && sc->scratch_used + words <= sc->scratch_cap) { a.cigar = sc->scratch + sc->scratch_used; sc->scratch_used += words; }
			else a.cigar = bb_malloc(4 * (size_t)words);
		}
		memcpy(a.cigar, g->cigar, 4 * (size_t)g->n_cigar + l_MD);
		a.NM = g->NM;
		if (a.n_cigar > 0) { /* drop a leading or trailing deletion */
			if ((a.cigar[0] & 0xf) == 2) {
				pos += a.cigar[0] >> 4;
				--a.n_cigar;
				memmove(a.cigar, a.cigar + 1, a.n_cigar * 4 + l_MD);
			} else if ((a.cigar[a.n_cigar - 1] & 0xf) == 2) {
				--a.n_cigar;
				memmove(a.cigar + a.n_cigar, a.cigar + a.n_cigar + 1, l_MD);
			}
		}
		if (qb != 0 || qe != l_query) {
			int clip5 = is_rev ? l_query - qe : qb, clip3 = is_rev ? qb : l_query - qe;
			if (clip5) {
				memmove(a.cigar + 1, a.cigar, a.n_cigar * 4 + l_MD);
				a.cigar[0] = clip5 << 4 | 3;
				++a.n_cigar;
			}
			if (clip3) {
				memmove(a.cigar + a.n_cigar + 1, a.cigar + a.n_cigar, l_MD);
				a.cigar[a.n_cigar++] = clip3 << 4 | 3;
			}
		}
	} else {
		/* alignment not available yet (dry pass): the leading-deletion shift of pos is unknown, but a
		 * deletion cannot move pos to another contig of a region that lies within one contig */
		a.n_cigar = 0; a.cigar = 0;
	}
	a.rid = bb_pos2rid(bns, pos);
	if (g) assert(a.rid == ar->rid);
	else a.rid = ar->rid;
	a.pos = pos - bns->anns[a.rid].offset;
	a.score = ar->score; a.sub = ar->sub > ar->csub ? ar->sub : ar->csub;
	a.is_alt = ar->is_alt; a.alt_sc = ar->alt_sc;
	return a;
}

static void put_cigar(const mem_opt_t *opt, const mem_aln_t *p, bb_str_t *str, int which) /* bwamem.c:838-849 */
{
	int i;
	if (p->n_cigar) {
		for (i = 0; i < p->n_cigar; ++i) {
			int c = p->cigar[i] & 0xf;
			if (!(opt->flag & MEM_F_SOFTCLIP) && !p->is_alt && (c == 3 || c == 4)) c = which ? 4 : 3;
			bb_putl(str, p->cigar[i] >> 4); bb_putc(str, "MIDSH"[c]);
		}
	} else bb_putc(str, '*');
}

/* SEQ and QUAL columns: codes 0..4 -> "ACGTN" (reverse strand: complement, right to left), quality copied or reversed.
 * 16 bases per step with byte shuffles where the CPU has SSSE3 (checked once at run time), else byte by byte. */
#if defined(__x86_64__) && defined(__GNUC__)
#include <tmmintrin.h>
__attribute__((target("ssse3"))) static void codes_to_text_ssse3(char *dst, const uint8_t *c, int n, int rev)
{
	const __m128i fwd = _mm_setr_epi8('A', 'C', 'G', 'T', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N');
	const __m128i cmp = _mm_setr_epi8('T', 'G', 'C', 'A', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N', 'N');
	const __m128i flip = _mm_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
	int i = 0;
	if (!rev) {
		for (; i + 16 <= n; i += 16) _mm_storeu_si128((__m128i *)(dst + i), _mm_shuffle_epi8(fwd, _mm_loadu_si128((const __m128i *)(c + i))));
		for (; i < n; ++i) dst[i] = "ACGTN"[c[i] > 4 ? 4 : c[i]];
	} else {
		for (; i + 16 <= n; i += 16)   /* the 16 codes that end at n-i, emitted in reverse order */
			_mm_storeu_si128((__m128i *)(dst + i), _mm_shuffle_epi8(cmp, _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(c + n - i - 16)), flip)));
		for (; i < n; ++i) { uint8_t v = c[n - 1 - i]; dst[i] = "TGCAN"[v > 4 ? 4 : v]; }
	}
}
__attribute__((target("ssse3"))) static void reverse_text_ssse3(char *dst, const char *src, int n)
{
	const __m128i flip = _mm_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
	int i = 0;
	for (; i + 16 <= n; i += 16) _mm_storeu_si128((__m128i *)(dst + i), _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(src + n - i - 16)), flip));
	for (; i < n; ++i) dst[i] = src[n - 1 - i];
}
static int have_ssse3(void) { static int v = -1; int x = __atomic_load_n(&v, __ATOMIC_RELAXED); if (x < 0) { x = __builtin_cpu_supports("ssse3") ? 1 : 0; __atomic_store_n(&v, x, __ATOMIC_RELAXED); } return x; }
#else
static int have_ssse3(void) { return 0; }
static void codes_to_text_ssse3(char *dst, const uint8_t *c, int n, int rev) { (void)dst; (void)c; (void)n; (void)rev; }
static void reverse_text_ssse3(char *dst, const char *src, int n) { (void)dst; (void)src; (void)n; }
#endif

void bb_codes_to_text(char *dst, const uint8_t *codes, int n, int rev)
{
	int i;
	if (have_ssse3()) { codes_to_text_ssse3(dst, codes, n, rev); return; }
	if (!rev) for (i = 0; i < n; ++i) dst[i] = "ACGTN"[codes[i] > 4 ? 4 : codes[i]];
	else for (i = 0; i < n; ++i) { uint8_t v = codes[n - 1 - i]; dst[i] = "TGCAN"[v > 4 ? 4 : v]; }
}

void bb_copy_text(char *dst, const char *src, int n, int rev)
{
	int i;
	if (!rev) { memcpy(dst, src, (size_t)n); return; }
	if (have_ssse3()) { reverse_text_ssse3(dst, src, n); return; }
	for (i = 0; i < n; ++i) dst[i] = src[n - 1 - i];
}

/* bwamem.c:851-976 */
void bb_aln2sam(const mem_opt_t *opt, const bntseq_t *bns, bb_str_t *str, bseq1_t *s, int n, const mem_aln_t *list, int which, const mem_aln_t *m_)
{
	int i;
	mem_aln_t ptmp = list[which], *p = &ptmp, mtmp, *m = 0;
	if (m_) { mtmp = *m_; m = &mtmp; }
	p->flag |= m ? 0x1 : 0;
	p->flag |= p->rid < 0 ? 0x4 : 0;
	p->flag |= m && m->rid < 0 ? 0x8 : 0;
	if (p->rid < 0 && m && m->rid >= 0) { p->rid = m->rid; p->pos = m->pos; p->is_rev = m->is_rev; p->n_cigar = 0; }
	if (m && m->rid < 0 && p->rid >= 0) { m->rid = p->rid; m->pos = p->pos; m->is_rev = p->is_rev; m->n_cigar = 0; }
	p->flag |= p->is_rev ? 0x10 : 0;
	p->flag |= m && m->is_rev ? 0x20 : 0;

	{   /* the fixed columns and tags: one capacity check for all of them, then plain stores through a local pointer
	     * (the text is the reference's, field by field: bwamem.c:881-948) */
		const char *rname = p->rid >= 0 ? bns->anns[p->rid].name : 0, *mname = m && m->rid >= 0 && p->rid != m->rid ? bns->anns[m->rid].name : 0;
		const char *md = p->n_cigar ? (const char *)(p->cigar + p->n_cigar) : 0;
		const size_t l_name = strlen(s->name), l_rname = rname ? strlen(rname) : 0, l_mname = mname ? strlen(mname) : 0, l_md = md ? strlen(md) : 0;
		const size_t l_rg = bwa_rg_id[0] ? strlen(bwa_rg_id) : 0;
		int qb = 0, qe = s->l_seq;
		char *w;
		if (!(p->flag & 0x100)) {
			const int hard = p->n_cigar && which && !(opt->flag & MEM_F_SOFTCLIP) && !p->is_alt;
			if (hard) { /* trim what a hard clip removes; on the reverse strand the CIGAR runs the other way */
				int c0 = p->cigar[0] & 0xf, c1 = p->cigar[p->n_cigar - 1] & 0xf;
				if (!p->is_rev) {
					if (c0 == 4 || c0 == 3) qb += p->cigar[0] >> 4;
					if (c1 == 4 || c1 == 3) qe -= p->cigar[p->n_cigar - 1] >> 4;
				} else {
					if (c0 == 4 || c0 == 3) qe -= p->cigar[0] >> 4;
					if (c1 == 4 || c1 == 3) qb += p->cigar[p->n_cigar - 1] >> 4;
				}
			}
		}
		bb_str_need(str, l_name + l_rname + l_mname + l_md + l_rg + 12 * ((size_t)p->n_cigar + (m ? (size_t)m->n_cigar : 0)) + 2 * (size_t)(qe > qb ? qe - qb : 0) + 320);
		w = str->s + str->l;
#define W_C(c) (*w++ = (char)(c))
#define W_S(ptr, len) do { memcpy(w, (ptr), (len)); w += (len); } while (0)
#define W_L(v) (w = bb_fmt_l(w, (int64_t)(v)))
#define W_CIGAR(al) do { const mem_aln_t *al_ = (al); int i_; \
			if (al_->n_cigar) { for (i_ = 0; i_ < al_->n_cigar; ++i_) { int c_ = al_->cigar[i_] & 0xf; \
				if (!(opt->flag & MEM_F_SOFTCLIP) && !al_->is_alt && (c_ == 3 || c_ == 4)) c_ = which ? 4 : 3; \
				W_L(al_->cigar[i_] >> 4); W_C("MIDSH"[c_]); } } else W_C('*'); } while (0)
		W_S(s->name, l_name); W_C('\t');
		W_L((p->flag & 0xffff) | (p->flag & 0x10000 ? 0x100 : 0)); W_C('\t');
		if (p->rid >= 0) {
			W_S(rname, l_rname); W_C('\t');
			W_L(p->pos + 1); W_C('\t');
			W_L(p->mapq); W_C('\t');
			W_CIGAR(p);
		} else W_S("*\t0\t0\t*", 7);
		W_C('\t');

		if (m && m->rid >= 0) {
			if (p->rid == m->rid) W_C('=');
			else W_S(mname, l_mname);
			W_C('\t');
			W_L(m->pos + 1); W_C('\t');
			if (p->rid == m->rid) {
				int64_t p0 = p->pos + (p->is_rev ? cigar_ref_len(p->n_cigar, p->cigar) - 1 : 0);
				int64_t p1 = m->pos + (m->is_rev ? cigar_ref_len(m->n_cigar, m->cigar) - 1 : 0);
				if (m->n_cigar == 0 || p->n_cigar == 0) W_C('0');
				else W_L(-(p0 - p1 + (p0 > p1 ? 1 : p0 < p1 ? -1 : 0)));
			} else W_C('0');
		} else W_S("*\t0\t0", 5);
		W_C('\t');

		if (p->flag & 0x100) W_S("*\t*", 3);
		else {
			bb_codes_to_text(w, (const uint8_t *)s->seq + qb, qe - qb, p->is_rev);
			w += qe - qb;
			W_C('\t');
			if (s->qual) { bb_copy_text(w, s->qual + qb, qe - qb, p->is_rev); w += qe - qb; }
			else W_C('*');
		}

		if (p->n_cigar) {
			W_S("\tNM:i:", 6); W_L(p->NM);
			W_S("\tMD:Z:", 6); W_S(md, l_md);
		}
		if (m && m->n_cigar) { W_S("\tMC:Z:", 6); W_CIGAR(m); }
		if (m) { W_S("\tMQ:i:", 6); W_L(m->mapq); }
		if (p->score >= 0) { W_S("\tAS:i:", 6); W_L(p->score); }
		if (p->sub >= 0) { W_S("\tXS:i:", 6); W_L(p->sub); }
		if (l_rg) { W_S("\tRG:Z:", 6); W_S(bwa_rg_id, l_rg); }
#undef W_C
#undef W_S
#undef W_L
#undef W_CIGAR
		*w = 0;
		str->l = (size_t)(w - str->s);
	}
	if (!(p->flag & 0x100)) {
		for (i = 0; i < n; ++i)
			if (i != which && !(list[i].flag & 0x100)) break;
		if (i < n) {
			bb_putsn(str, "\tSA:Z:", 6);
			for (i = 0; i < n; ++i) {
				const mem_aln_t *r = &list[i];
				int k;
				if (i == which || (r->flag & 0x100)) continue;
				bb_puts(str, bns->anns[r->rid].name); bb_putc(str, ',');
				bb_putl(str, r->pos + 1); bb_putc(str, ',');
				bb_putc(str, "+-"[r->is_rev]); bb_putc(str, ',');
				for (k = 0; k < r->n_cigar; ++k) { bb_putl(str, r->cigar[k] >> 4); bb_putc(str, "MIDSH"[r->cigar[k] & 0xf]); }
				bb_putc(str, ','); bb_putl(str, r->mapq);
				bb_putc(str, ','); bb_putl(str, r->NM);
				bb_putc(str, ';');
			}
		}
		if (p->alt_sc > 0) {
			char buf[64];
			snprintf(buf, sizeof(buf), "\tpa:f:%.3f", (double)p->score / p->alt_sc);
			bb_puts(str, buf);
		}
	}
	if (p->XA) {
		bb_putsn(str, (opt->flag & MEM_F_XB) ? "\tXB:Z:" : "\tXA:Z:", 6);
		bb_puts(str, p->XA);
	}
	if (s->comment) { bb_putc(str, '\t'); bb_puts(str, s->comment); }
	if ((opt->flag & MEM_F_REF_HDR) && p->rid >= 0 && bns->anns[p->rid].anno != 0 && bns->anns[p->rid].anno[0] != 0) {
		size_t from;
		bb_putsn(str, "\tXR:Z:", 6);
		from = str->l;
		bb_puts(str, bns->anns[p->rid].anno);
		for (; from < str->l; ++from)
			if (str->s[from] == '\t') str->s[from] = ' ';
	}
	bb_putc(str, '\n');
}

static int xa_parent(double ratio, const mem_alnreg_t *a, int i) /* bwamem_extra.c:116-121 */
{
	int k = a[i].secondary_all;
	if (k >= 0 && a[i].score >= a[k].score * ratio) return k;
	return -1;
}

/* XA/XB strings per primary region (bwamem_extra.c:124-172); NULL when no region has alternatives */
char **bb_gen_alt(bb_samctx_t *sc, const mem_alnreg_v *a, int l_query, const char *query)
{
	const mem_opt_t *opt = sc->opt;
	const bntseq_t *bns = sc->bns;
	int i, k, r, tot = 0, *cnt;
	bb_str_t *aln, one = {0, 0, 0};
	char **XA, *has_alt;
	if (a->n <= 1) return 0;      /* a lone region has no parent: nothing to list (tot == 0 below) */
	cnt = bb_calloc(a->n, sizeof(int));
	has_alt = bb_calloc(a->n, 1);
	for (i = 0; i < (int)a->n; ++i) {
		r = xa_parent(opt->XA_drop_ratio, a->a, i);
		if (r >= 0) { ++cnt[r]; ++tot; if (a->a[i].is_alt) has_alt[r] = 1; }
	}
	if (tot == 0) { free(cnt); free(has_alt); return 0; }
	aln = bb_calloc(a->n, sizeof(bb_str_t));
	for (i = 0; i < (int)a->n; ++i) {
		mem_aln_t t;
		if ((r = xa_parent(opt->XA_drop_ratio, a->a, i)) < 0) continue;
		if (cnt[r] > opt->max_XA_hits_alt || (!has_alt[r] && cnt[r] > opt->max_XA_hits)) continue;
		t = bb_reg2aln(sc, l_query, query, &a->a[i]);
		if (sc->dry) { bb_cigar_free(sc, t.cigar); continue; }
		one.l = 0;
		bb_puts(&one, bns->anns[t.rid].name);
		bb_putc(&one, ','); bb_putc(&one, "+-"[t.is_rev]); bb_putl(&one, t.pos + 1);
		bb_putc(&one, ',');
		for (k = 0; k < t.n_cigar; ++k) { bb_putl(&one, t.cigar[k] >> 4); bb_putc(&one, "MIDSHN"[t.cigar[k] & 0xf]); }
		bb_putc(&one, ','); bb_putl(&one, t.NM);
		if (opt->flag & MEM_F_XB) { bb_putc(&one, ','); bb_putl(&one, t.score); bb_putc(&one, ','); bb_putl(&one, t.mapq); }
		bb_putc(&one, ';');
		bb_cigar_free(sc, t.cigar);
		bb_putsn(&aln[r], one.s, one.l);
	}
	XA = bb_calloc(a->n, sizeof(char *));
	for (k = 0; k < (int)a->n; ++k) XA[k] = aln[k].s;
	free(has_alt); free(cnt); free(aln); free(one.s);
	return XA;
}

/* bwamem.c:1033-1079 */
void bb_reg2sam(bb_samctx_t *sc, bseq1_t *s, mem_alnreg_v *a, int extra_flag, const mem_aln_t *m)
{
	const mem_opt_t *opt = sc->opt;
	bb_str_t str = {0, 0, 0};
	BB_VEC(mem_aln_t) aa = {0, 0, 0};
	size_t k;
	int l = 0;
	char **XA = 0;
	if (!(opt->flag & MEM_F_ALL)) XA = bb_gen_alt(sc, a, s->l_seq, s->seq);
	if (!sc->dry) bb_str_need(&str, (size_t)s->l_seq * 2 + strlen(s->name) + 448);   /* one allocation for the common single-record case */
	for (k = 0; k < a->n; ++k) {
		mem_alnreg_t *p = &a->a[k];
		mem_aln_t q;
		if (p->score < opt->T) continue;
		if (p->secondary >= 0 && (p->is_alt || !(opt->flag & MEM_F_ALL))) continue;
		if (p->secondary >= 0 && p->secondary < INT_MAX && p->score < a->a[p->secondary].score * opt->drop_ratio) continue;
		q = bb_reg2aln(sc, s->l_seq, s->seq, p);
		q.XA = XA ? XA[k] : 0;
		q.flag |= extra_flag;
		if (p->secondary >= 0) q.sub = -1;
		if (l && p->secondary < 0) q.flag |= (opt->flag & MEM_F_NO_MULTI) ? 0x10000 : 0x800;
		if (!(opt->flag & MEM_F_KEEP_SUPP_MAPQ) && l && !p->is_alt && q.mapq > aa.a[0].mapq) q.mapq = aa.a[0].mapq;
		bb_vec_push(aa, q);
		++l;
	}
	if (!sc->dry) {
		if (aa.n == 0) {
			mem_aln_t t = bb_reg2aln(sc, s->l_seq, s->seq, 0);
			t.flag |= extra_flag;
			bb_aln2sam(opt, sc->bns, &str, s, 1, &t, 0, m);
		} else for (k = 0; k < aa.n; ++k) bb_aln2sam(opt, sc->bns, &str, s, (int)aa.n, aa.a, (int)k, m);
		s->sam = str.s;
	}
	for (k = 0; k < aa.n; ++k) bb_cigar_free(sc, aa.a[k].cigar);
	free(aa.a);
	if (XA) { for (k = 0; k < a->n; ++k) free(XA[k]); free(XA); }
}
