// Random-gather ceiling of HBM for the FM-index access pattern: every lane reads whole 64-byte blocks (4 x LDG.128,
// or 32-byte half blocks) at independent random addresses of a table much larger than L2.  This is what K1/K2 can at
// best approach; a streaming-copy peak is not reachable with 64-byte random requests (one DRAM row activation each).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench tools/gather_bench.cu -lcuda && ./gather_bench [table GB]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda.h>
typedef unsigned long long u64;
template <int BYTES, int ILP, bool WIDE>
__global__ void k_gather(const uint4 *tab, u64 n_blocks, int iters, u64 *sink)
{
	u64 s = (u64)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
	u64 acc = 0;
	for (int it = 0; it < iters; ++it) {
		uint4 v[ILP][BYTES / 16];
#pragma unroll
		for (int u = 0; u < ILP; ++u) {
			s = s * 6364136223846793005ull + 1442695040888963407ull;
			const u64 b = (s >> 20) & (n_blocks - 1);     /* n_blocks is a power of two */
			if (WIDE) {     /* one 256-bit load per 32 bytes */
#pragma unroll
				for (int q = 0; q < BYTES / 32; ++q)
					asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[u][2 * q].x), "=r"(v[u][2 * q].y), "=r"(v[u][2 * q].z), "=r"(v[u][2 * q].w),
					             "=r"(v[u][2 * q + 1].x), "=r"(v[u][2 * q + 1].y), "=r"(v[u][2 * q + 1].z), "=r"(v[u][2 * q + 1].w) : "l"(tab + b * 4 + 2 * q));
			} else {
#pragma unroll
				for (int q = 0; q < BYTES / 16; ++q) v[u][q] = __ldg(tab + b * 4 + q);
			}
		}
#pragma unroll
		for (int u = 0; u < ILP; ++u)
#pragma unroll
			for (int q = 0; q < BYTES / 16; ++q) acc += v[u][q].x ^ v[u][q].w;
		s ^= acc & 1;     // the next address depends on the data, like an FM-index walk
	}
	if (acc == 0x1234567) *sink = acc;
}
template <int BYTES, int ILP, bool WIDE = false> void run(const uint4 *tab, u64 n_blocks, u64 *sink, int threads_per_sm_target)
{
	int dev, nsm; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
	const int block = 128, grid = nsm * threads_per_sm_target / block, iters = 2000 / ILP;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_gather<BYTES, ILP, WIDE><<<grid, block>>>(tab, n_blocks, iters / 4, sink);
	cudaEventRecord(e0);
	k_gather<BYTES, ILP, WIDE><<<grid, block>>>(tab, n_blocks, iters, sink);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	const double n = (double)grid * block * iters * ILP;
	printf("bytes/request %3d (%s loads)  requests in flight/lane %d  lanes/SM %4d : %7.1f G requests/s  %7.1f GB/s\n", BYTES, WIDE ? "256-bit" : "128-bit", ILP, threads_per_sm_target, n / ms / 1e6, n * BYTES / ms / 1e6);
}
// the same table through the virtual-memory API with 512 MiB alignment: does the driver map it with larger pages?
static uint4 *vmm_alloc(size_t bytes)
{
	CUmemAllocationProp prop = {};
	prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
	size_t gmin = 0, grec = 0;
	cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
	cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
	printf("VMM granularity: minimum %zu, recommended %zu\n", gmin, grec);
	const size_t big = (size_t)512 << 20;
	bytes = (bytes + big - 1) / big * big;
	CUmemGenericAllocationHandle h;
	CUdeviceptr va = 0;
	if (cuMemCreate(&h, bytes, &prop, 0) != CUDA_SUCCESS) { printf("cuMemCreate failed\n"); return 0; }
	if (cuMemAddressReserve(&va, bytes, big, 0, 0) != CUDA_SUCCESS) { printf("cuMemAddressReserve failed\n"); return 0; }
	if (cuMemMap(va, bytes, 0, h, 0) != CUDA_SUCCESS) { printf("cuMemMap failed\n"); return 0; }
	CUmemAccessDesc acc = {};
	acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
	if (cuMemSetAccess(va, bytes, &acc, 1) != CUDA_SUCCESS) { printf("cuMemSetAccess failed\n"); return 0; }
	return (uint4 *)va;
}

int main(int argc, char **argv)
{
	uint4 *tab; u64 *sink;
	const u64 max_blocks = (u64)1 << 27;          /* 8 GiB of 64-byte blocks */
	cudaMalloc(&tab, max_blocks * 64); cudaMalloc(&sink, 8);
	cudaMemset(tab, 1, max_blocks * 64);
	(void)argc; (void)argv;
	for (u64 n_blocks = (u64)1 << 22; n_blocks <= max_blocks; n_blocks <<= 1) {   /* 256 MiB .. 8 GiB */
		if (n_blocks != (u64)1 << 22 && n_blocks != (u64)1 << 24 && n_blocks != (u64)1 << 26 && n_blocks != max_blocks) continue;
		printf("table %.2f GiB, dependent random requests\n", n_blocks * 64.0 / (1 << 30));
		run<64, 1>(tab, n_blocks, sink, 640);  run<64, 1>(tab, n_blocks, sink, 2048);
		run<64, 2>(tab, n_blocks, sink, 640);  run<64, 2>(tab, n_blocks, sink, 2048);
		run<64, 4>(tab, n_blocks, sink, 2048);
		run<32, 1>(tab, n_blocks, sink, 640);  run<32, 1>(tab, n_blocks, sink, 2048);
		run<32, 2>(tab, n_blocks, sink, 640);  run<32, 2>(tab, n_blocks, sink, 2048);
		run<32, 4>(tab, n_blocks, sink, 2048);
		run<32, 1, true>(tab, n_blocks, sink, 640); run<32, 2, true>(tab, n_blocks, sink, 640); run<32, 2, true>(tab, n_blocks, sink, 2048);
		run<64, 1, true>(tab, n_blocks, sink, 640); run<64, 2, true>(tab, n_blocks, sink, 2048);
	}
	cudaFree(tab);
	uint4 *vt = vmm_alloc(max_blocks * 64);
	if (vt) {
		cudaMemset(vt, 1, max_blocks * 64);
		printf("table 8.00 GiB through cuMemCreate/cuMemMap, 512 MiB aligned\n");
		run<64, 1>(vt, max_blocks, sink, 640); run<32, 1>(vt, max_blocks, sink, 640); run<32, 2, true>(vt, max_blocks, sink, 640);
	}
	return 0;
}
