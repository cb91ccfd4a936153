// Roofline denominator the driver's MEASURED_PEAKS.json does not carry: the bandwidth of fully divergent gathers from a
// buffer that is L2-resident (the 51 MB hash table, the per-frame BVH) or not (a 2^22-entry table).  Every lane of a warp loads
// `chunk` bytes (16 = one hash-table entry, 128 = one BVH quad fetched by 8 cooperating lanes) from a pseudo-random, chunk-aligned
// offset, `per_thread` times with independent addresses (8 loads in flight per lane, like the 8 corners of one level).
#include <algorithm>

#include "ctx.h"

namespace insta {
namespace {

__device__ __forceinline__ uint32_t mix(uint32_t x) {
	x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
	return x;
}
__device__ __forceinline__ uint4 ld_gather(const uint4* p) {
	uint4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

// COOP = lanes per chunk (1: 16-byte chunks, 8: 128-byte chunks read by 8 adjacent lanes)
template <int COOP>
__global__ void __launch_bounds__(512) gather_kernel(const uint4* __restrict__ buf, uint32_t n_chunks, uint32_t per_thread, uint32_t seed, uint32_t* __restrict__ sink) {
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t group = tid / COOP, sub = tid % COOP;
	uint32_t acc = 0;
	for (uint32_t it = 0; it < per_thread; it += 8) {
		uint4 v[8];
#pragma unroll
		for (uint32_t k = 0; k < 8; ++k) {
			const uint32_t c = mix(group * 0x9e3779b9u + (it + k) * 0x85ebca6bu + seed) % n_chunks;
			v[k] = ld_gather(buf + (size_t)c * COOP + sub);
		}
#pragma unroll
		for (uint32_t k = 0; k < 8; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
	}
	if (acc == 0x12345678u) sink[0] = acc;  // never true in practice; keeps the loads alive
}

// The same access pattern as reductions: every lane adds 8 halves to a pseudo-random 16-byte entry (red.global.add.noftz.v4.f16x2, what hash_bwd_kernel issues
// per corner); COOP adjacent lanes hit adjacent entries of one COOP x 16-byte chunk (2 = the two x-sides of a cell sharing a 32-byte sector).
template <int COOP>
__global__ void __launch_bounds__(512) scatter_kernel(uint4* __restrict__ buf, uint32_t n_chunks, uint32_t per_thread, uint32_t seed) {
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t group = tid / COOP, sub = tid % COOP;
	const uint32_t one = 0x04000400u;  // two small fp16 values: the sums stay finite
	for (uint32_t it = 0; it < per_thread; ++it) {
		const uint32_t c = mix(group * 0x9e3779b9u + it * 0x85ebca6bu + seed) % n_chunks;
		asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1,%1,%1,%1};" ::"l"(buf + (size_t)c * COOP + sub), "r"(one) : "memory");
	}
}
}  // namespace

int microbench_scatter(insta_ctx* ctx, uint64_t buffer_bytes, uint32_t chunk_bytes, uint32_t iters, float* h_gbs, cudaStream_t s) {
	if (chunk_bytes != 16 && chunk_bytes != 32 && chunk_bytes != 64) INSTA_FAIL(INSTA_ERR_INVALID, "insta_microbench_scatter: chunk_bytes must be 16, 32 or 64");
	if (buffer_bytes < 4096 || buffer_bytes % 128 || buffer_bytes / chunk_bytes > 0xffffffffull) INSTA_FAIL(INSTA_ERR_INVALID, "insta_microbench_scatter: bad buffer size");
	uint4* buf = nullptr;
	INSTA_CUDA(cudaMalloc(&buf, buffer_bytes));
	const uint32_t n_chunks = (uint32_t)(buffer_bytes / chunk_bytes);
	const uint32_t threads = 512, per_thread = 32;
	const uint32_t blocks = (uint32_t)ctx->sm_count * 4u * 8u;
	cudaEvent_t e0, e1;
	INSTA_CUDA(cudaEventCreate(&e0));
	INSTA_CUDA(cudaEventCreate(&e1));
	float best_ms = 1e30f;
	for (uint32_t it = 0; it < iters + 2; ++it) {
		INSTA_CUDA(cudaMemsetAsync(buf, 0, buffer_bytes, s));
		INSTA_CUDA(cudaEventRecord(e0, s));
		if (chunk_bytes == 16) scatter_kernel<1><<<blocks, threads, 0, s>>>(buf, n_chunks, per_thread, 17u * it + 1u);
		else if (chunk_bytes == 32) scatter_kernel<2><<<blocks, threads, 0, s>>>(buf, n_chunks, per_thread, 17u * it + 1u);
		else scatter_kernel<4><<<blocks, threads, 0, s>>>(buf, n_chunks, per_thread, 17u * it + 1u);
		ctx->launches++;
		INSTA_CUDA(cudaEventRecord(e1, s));
		INSTA_CUDA(cudaEventSynchronize(e1));
		float ms = 0.f;
		INSTA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
		if (it >= 2) best_ms = std::min(best_ms, ms);
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(buf);
	const double bytes = (double)blocks * threads * per_thread * 16.0;
	*h_gbs = (float)(bytes / (best_ms * 1e-3) / 1e9);
	return INSTA_OK;
}

int microbench_gather(insta_ctx* ctx, uint64_t buffer_bytes, uint32_t chunk_bytes, uint32_t iters, float* h_gbs, cudaStream_t s) {
	if (chunk_bytes != 16 && chunk_bytes != 32 && chunk_bytes != 64 && chunk_bytes != 128) INSTA_FAIL(INSTA_ERR_INVALID, "insta_microbench_gather: chunk_bytes must be 16, 32, 64 or 128");
	if (buffer_bytes < 4096 || buffer_bytes % 128 || buffer_bytes / chunk_bytes > 0xffffffffull) INSTA_FAIL(INSTA_ERR_INVALID, "insta_microbench_gather: bad buffer size");
	uint4* buf = nullptr;
	uint32_t* sink = nullptr;
	INSTA_CUDA(cudaMalloc(&buf, buffer_bytes));
	INSTA_CUDA(cudaMalloc(&sink, 4));
	INSTA_CUDA(cudaMemsetAsync(buf, 1, buffer_bytes, s));
	const uint32_t n_chunks = (uint32_t)(buffer_bytes / chunk_bytes);
	const uint32_t threads = 512, per_thread = 64;
	const uint32_t blocks = (uint32_t)ctx->sm_count * 4u * 8u;  // 8 waves of 4 CTAs per SM
	cudaEvent_t e0, e1;
	INSTA_CUDA(cudaEventCreate(&e0));
	INSTA_CUDA(cudaEventCreate(&e1));
	float best_ms = 1e30f;
	for (uint32_t it = 0; it < iters + 2; ++it) {  // two warm-up launches bring the buffer into L2 when it fits
		INSTA_CUDA(cudaEventRecord(e0, s));
		if (chunk_bytes == 16) gather_kernel<1><<<blocks, threads, 0, s>>>(buf, n_chunks, per_thread, 17u * it + 1u, sink);
		else if (chunk_bytes == 32) gather_kernel<2><<<blocks, threads, 0, s>>>(buf, n_chunks, per_thread, 17u * it + 1u, sink);  // 2 adjacent lanes share a 32-byte sector
		else if (chunk_bytes == 64) gather_kernel<4><<<blocks, threads, 0, s>>>(buf, n_chunks, per_thread, 17u * it + 1u, sink);
		else gather_kernel<8><<<blocks, threads, 0, s>>>(buf, n_chunks, per_thread, 17u * it + 1u, sink);
		ctx->launches++;
		INSTA_CUDA(cudaEventRecord(e1, s));
		INSTA_CUDA(cudaEventSynchronize(e1));
		float ms = 0.f;
		INSTA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
		if (it >= 2) best_ms = std::min(best_ms, ms);
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(buf);
	cudaFree(sink);
	const double bytes = (double)blocks * threads * per_thread * 16.0;
	*h_gbs = (float)(bytes / (best_ms * 1e-3) / 1e9);
	return INSTA_OK;
}

}  // namespace insta
