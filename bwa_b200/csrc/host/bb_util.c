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

/* ---- parallel for on a persistent pool of worker threads ----
 * Workers sleep on a condition variable between jobs; a job hands out dynamic chunks from one atomic
 * counter.  Pools are keyed by size and created on first use (a batch issues ~10 parallel loops, so
 * creating 100+ threads per loop would cost more than the loops themselves). */
typedef struct bb_pool {
	int nt;                       /* workers including the caller */
	pthread_t *th;
	pthread_mutex_t mu, job_mu;
	pthread_cond_t cv_go, cv_done;
	long gen;                     /* job generation */
	int n_idle_done;              /* workers finished with the current job */
	void (*fn)(void *, long, int);
	void *data;
	long n, chunk;
	volatile long next;
	int lane;
	struct bb_pool *link;
} bb_pool_t;
typedef struct { bb_pool_t *p; int tid; } pool_arg_t;

static void pool_run(bb_pool_t *p, int tid)
{
	for (;;) {
		long b = __sync_fetch_and_add(&p->next, p->chunk), e, i;
		if (b >= p->n) break;
		e = b + p->chunk < p->n ? b + p->chunk : p->n;
		for (i = b; i < e; ++i) p->fn(p->data, i, tid);
	}
}

static void *pool_worker(void *a_)
{
	pool_arg_t *a = a_;
	bb_pool_t *p = a->p;
	long seen = 0;
	for (;;) {
		pthread_mutex_lock(&p->mu);
		while (p->gen == seen) pthread_cond_wait(&p->cv_go, &p->mu);
		seen = p->gen;
		pthread_mutex_unlock(&p->mu);
		pool_run(p, a->tid);
		pthread_mutex_lock(&p->mu);
		if (++p->n_idle_done == p->nt - 1) pthread_cond_signal(&p->cv_done);
		pthread_mutex_unlock(&p->mu);
	}
	return 0;
}

static bb_pool_t *g_pools;
static pthread_mutex_t g_pools_mu = PTHREAD_MUTEX_INITIALIZER;

static bb_pool_t *pool_get(int nt, int lane)
{
	bb_pool_t *p;
	int t;
	pthread_mutex_lock(&g_pools_mu);
	for (p = g_pools; p; p = p->link) if (p->nt == nt && p->lane == lane) break;
	if (!p) {
		p = bb_calloc(1, sizeof(*p));
		p->nt = nt; p->lane = lane;
		pthread_mutex_init(&p->mu, 0); pthread_mutex_init(&p->job_mu, 0);
		pthread_cond_init(&p->cv_go, 0); pthread_cond_init(&p->cv_done, 0);
		p->th = bb_malloc(sizeof(pthread_t) * nt);
		for (t = 1; t < nt; ++t) {
			pool_arg_t *a = bb_malloc(sizeof(*a));
			a->p = p; a->tid = t;
			if (pthread_create(&p->th[t], 0, pool_worker, a) != 0) bb_fatal("bb_parallel_for", "pthread_create failed");
			pthread_detach(p->th[t]);
		}
		p->link = g_pools; g_pools = p;
	}
	pthread_mutex_unlock(&g_pools_mu);
	return p;
}

void bb_parallel_for(int nt, void (*fn)(void *, long, int), void *data, long n) { bb_parallel_for_lane(0, nt, fn, data, n); }

/* independent pools per lane: two batches in flight run their host loops side by side */
void bb_parallel_for_lane(int lane, int nt, void (*fn)(void *, long, int), void *data, long n)
{
	bb_pool_t *p;
	if (n <= 0) return;
	if (nt < 1) nt = 1;
	if (nt == 1 || n == 1) { long i; for (i = 0; i < n; ++i) fn(data, i, 0); return; }
	p = pool_get(nt, lane);
	pthread_mutex_lock(&p->job_mu);       /* one job at a time per pool */
	p->fn = fn; p->data = data; p->n = n; p->next = 0;
	p->chunk = n / (nt * 8L); if (p->chunk < 1) p->chunk = 1; if (p->chunk > 512) p->chunk = 512;
	pthread_mutex_lock(&p->mu);
	p->n_idle_done = 0; ++p->gen;
	pthread_cond_broadcast(&p->cv_go);
	pthread_mutex_unlock(&p->mu);
	pool_run(p, 0);
	pthread_mutex_lock(&p->mu);
	while (p->n_idle_done < p->nt - 1) pthread_cond_wait(&p->cv_done, &p->mu);
	pthread_mutex_unlock(&p->mu);
	pthread_mutex_unlock(&p->job_mu);
}
