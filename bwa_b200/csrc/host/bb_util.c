/* bb_util.c -- fatal errors, checked allocation, timers, thread helpers, plain-key sorts. */
#include <stdarg.h>
#include <pthread.h>
#include <sys/time.h>
#include <sys/resource.h>
#include <limits.h>
#include "bb_util.h"
#include "bb_sort.h"

void bb_fatal(const char *where, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	fprintf(stderr, "[%s] ", where);
	vfprintf(stderr, fmt, ap);
	fprintf(stderr, "\n");
	va_end(ap);
	exit(EXIT_FAILURE); /* reference: err_fatal -> exit(EXIT_FAILURE), utils.c:90-99 */
}

void *bb_malloc(size_t n)
{
	void *p = malloc(n ? n : 1);
	if (!p) bb_fatal("bb_malloc", "out of memory allocating %zu bytes", n);
	return p;
}
void *bb_calloc(size_t n, size_t sz)
{
	void *p = calloc(n ? n : 1, sz ? sz : 1);
	if (!p) bb_fatal("bb_calloc", "out of memory allocating %zu x %zu bytes", n, sz);
	return p;
}
void *bb_realloc(void *q, size_t n)
{
	void *p = realloc(q, n ? n : 1);
	if (!p) bb_fatal("bb_realloc", "out of memory allocating %zu bytes", n);
	return p;
}
char *bb_strdup(const char *s)
{
	size_t l = strlen(s);
	char *p = bb_malloc(l + 1);
	memcpy(p, s, l + 1);
	return p;
}

double bb_cputime(void)
{
	struct rusage r;
	getrusage(RUSAGE_SELF, &r);
	return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec);
}
double bb_realtime(void)
{
	struct timeval tp;
	gettimeofday(&tp, 0);
	return tp.tv_sec + tp.tv_usec * 1e-6;
}

/* ---- sorts on plain keys ---- */
#define u64_lt(a, b) ((a) < (b))
BB_SORT_DEFINE(, bb_sort_u64, uint64_t, u64_lt)
#define p64_lt(a, b) ((a).x < (b).x || ((a).x == (b).x && (a).y < (b).y))
BB_SORT_DEFINE(, bb_sort_pair64, bb_pair64_t, p64_lt)

/* ---- parallel for: dynamic chunks off one atomic counter ---- */
typedef struct {
	void (*fn)(void *, long, int);
	void *data;
	long n, chunk;
	long next;
} pf_shared_t;
typedef struct { pf_shared_t *sh; int tid; } pf_arg_t;

static void *pf_worker(void *a_)
{
	pf_arg_t *a = a_;
	pf_shared_t *sh = a->sh;
	for (;;) {
		long b = __sync_fetch_and_add(&sh->next, sh->chunk), e, i;
		if (b >= sh->n) break;
		e = b + sh->chunk < sh->n ? b + sh->chunk : sh->n;
		for (i = b; i < e; ++i) sh->fn(sh->data, i, a->tid);
	}
	return 0;
}

void bb_parallel_for(int nt, void (*fn)(void *, long, int), void *data, long n)
{
	pf_shared_t sh;
	int t;
	if (n <= 0) return;
	if (nt < 1) nt = 1;
	if (nt == 1 || n == 1) { long i; for (i = 0; i < n; ++i) fn(data, i, 0); return; }
	sh.fn = fn; sh.data = data; sh.n = n; sh.next = 0;
	sh.chunk = n / (nt * 16L); if (sh.chunk < 1) sh.chunk = 1; if (sh.chunk > 256) sh.chunk = 256;
	{
		pthread_t *th = bb_malloc(sizeof(pthread_t) * nt);
		pf_arg_t *args = bb_malloc(sizeof(pf_arg_t) * nt);
		for (t = 0; t < nt; ++t) { args[t].sh = &sh; args[t].tid = t; }
		for (t = 1; t < nt; ++t)
			if (pthread_create(&th[t], 0, pf_worker, &args[t]) != 0) bb_fatal("bb_parallel_for", "pthread_create failed");
		pf_worker(&args[0]);
		for (t = 1; t < nt; ++t) pthread_join(th[t], 0);
		free(th); free(args);
	}
}

