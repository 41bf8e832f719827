/* TEST INFRASTRUCTURE ONLY -- fiber runtime of the SIMT emulator (see cusim.h). x86-64 SysV only. */
#include "cusim.h"
#include <sys/mman.h>
#include <pthread.h>
#include <vector>
#include <atomic>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
thread_local unsigned char *cusim_dyn_smem;

extern "C" void cusim_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl cusim_switch
.type cusim_switch,@function
cusim_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
.size cusim_switch,.-cusim_switch
)");

namespace {

const size_t STACK_SZ = 512 << 10;

struct Fiber { void *sp; char *stack; unsigned tid; bool done; uint3 tix; };
struct Warp { uint64_t slot[32]; unsigned arrived, read; };

struct BlockCtx {
	std::vector<Fiber> fibers;
	std::vector<Warp> warps;
	void *sched_sp;
	int cur;
	unsigned n_threads;
	const std::function<void()> *body;
	unsigned long progress;
	unsigned bar_count; unsigned long bar_gen;
	std::vector<char *> stack_pool;
};
thread_local BlockCtx *g_ctx;

void fiber_main()
{
	BlockCtx *c = g_ctx;
	(*c->body)();
	c->fibers[c->cur].done = true;
	++c->progress;
	cusim_switch(&c->fibers[c->cur].sp, c->sched_sp);
	abort(); /* never resumed */
}

char *get_stack(BlockCtx *c, unsigned i)
{
	while (c->stack_pool.size() <= i) {
		void *p = mmap(0, STACK_SZ, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
		if (p == MAP_FAILED) { perror("cusim: mmap"); abort(); }
		c->stack_pool.push_back((char *)p);
	}
	return c->stack_pool[i];
}

void run_block(BlockCtx *c, dim3 grid, dim3 block, uint3 bidx, size_t smem, const std::function<void()> &body)
{
	unsigned n = block.x * block.y * block.z;
	std::vector<unsigned char> dyn(smem + 16);
	c->n_threads = n; c->body = &body; c->progress = 0; c->bar_count = 0; c->bar_gen = 0;
	c->fibers.assign(n, Fiber());
	c->warps.assign((n + 31) / 32, Warp());
	for (auto &w : c->warps) { w.arrived = w.read = 0; }
	gridDim = grid; blockDim = block; blockIdx = bidx;
	cusim_dyn_smem = dyn.data();
	for (unsigned i = 0; i < n; ++i) {
		Fiber &f = c->fibers[i];
		f.stack = get_stack(c, i);
		f.tid = i; f.done = false;
		f.tix.x = i % block.x; f.tix.y = (i / block.x) % block.y; f.tix.z = i / (block.x * block.y);
		uintptr_t top = ((uintptr_t)f.stack + STACK_SZ) & ~(uintptr_t)15;
		void **sp = (void **)(top - 64);
		for (int k = 0; k < 6; ++k) sp[k] = 0;
		sp[6] = (void *)fiber_main;  /* return address consumed by cusim_switch's ret */
		sp[7] = 0;
		f.sp = sp;
	}
	unsigned alive = n;
	unsigned long idle_rounds = 0;
	while (alive) {
		unsigned long before = c->progress;
		for (unsigned i = 0; i < n; ++i) {
			Fiber &f = c->fibers[i];
			if (f.done) continue;
			c->cur = (int)i;
			threadIdx = f.tix;
			cusim_switch(&c->sched_sp, f.sp);
			if (f.done) --alive;
		}
		if (c->progress == before) {
			if (++idle_rounds > 100000) { fprintf(stderr, "cusim: deadlock in block (%u,%u,%u): no fiber makes progress\n", bidx.x, bidx.y, bidx.z); abort(); }
		} else idle_rounds = 0;
	}
}

struct LaunchShared { dim3 grid, block; size_t smem; const std::function<void()> *body; std::atomic<unsigned long> next; };

void *launch_worker(void *a)
{
	LaunchShared *ls = (LaunchShared *)a;
	BlockCtx ctx;
	g_ctx = &ctx;
	unsigned long total = (unsigned long)ls->grid.x * ls->grid.y * ls->grid.z;
	for (;;) {
		unsigned long b = ls->next.fetch_add(1);
		if (b >= total) break;
		uint3 bi; bi.x = (unsigned)(b % ls->grid.x); bi.y = (unsigned)((b / ls->grid.x) % ls->grid.y); bi.z = (unsigned)(b / ((unsigned long)ls->grid.x * ls->grid.y));
		run_block(&ctx, ls->grid, ls->block, bi, ls->smem, *ls->body);
	}
	for (char *s : ctx.stack_pool) munmap(s, STACK_SZ);
	g_ctx = 0;
	return 0;
}

} // namespace

int cusim_lane() { return (int)(g_ctx->fibers[g_ctx->cur].tid & 31); }

void cusim_yield()
{
	BlockCtx *c = g_ctx;
	cusim_switch(&c->fibers[c->cur].sp, c->sched_sp);
}

uint64_t cusim_warp_gather(unsigned mask, uint64_t v, uint64_t out[32])
{
	BlockCtx *c = g_ctx;
	unsigned tid = c->fibers[c->cur].tid, lane = tid & 31, bit = 1u << lane;
	Warp &w = c->warps[tid >> 5];
	unsigned nl = c->n_threads - (tid & ~31u);
	if (nl < 32) mask &= (1u << nl) - 1;
	if (!(mask & bit)) { fprintf(stderr, "cusim: lane %u calls a collective with mask %08x that excludes it\n", lane, mask); abort(); }
	while (w.arrived & bit) cusim_yield();
	w.slot[lane] = v; w.arrived |= bit;
	++c->progress;
	while ((w.arrived & mask) != mask) cusim_yield();
	for (int i = 0; i < 32; ++i) out[i] = w.slot[i];
	w.read |= bit;
	if ((w.read & mask) == mask) { w.arrived &= ~mask; w.read &= ~mask; ++c->progress; }
	else while (w.read & bit) cusim_yield();
	return v;
}

void cusim_block_barrier()
{
	BlockCtx *c = g_ctx;
	unsigned long gen = c->bar_gen;
	++c->progress;
	if (++c->bar_count == c->n_threads) { c->bar_count = 0; ++c->bar_gen; return; }
	while (c->bar_gen == gen) cusim_yield();
}

void cusim_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
	LaunchShared ls;
	ls.grid = grid; ls.block = block; ls.smem = smem; ls.body = &body; ls.next = 0;
	const char *e = getenv("CUSIM_THREADS");
	int nt = e ? atoi(e) : 8;
	unsigned long total = (unsigned long)grid.x * grid.y * grid.z;
	if (nt < 1) nt = 1;
	if ((unsigned long)nt > total) nt = (int)total;
	if (nt <= 1) { launch_worker(&ls); return; }
	std::vector<pthread_t> th(nt);
	for (int i = 0; i < nt; ++i) pthread_create(&th[i], 0, launch_worker, &ls);
	for (int i = 0; i < nt; ++i) pthread_join(th[i], 0);
}
