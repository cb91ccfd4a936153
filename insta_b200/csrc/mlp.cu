// The NeRF network of the INSTA hot path on sm_100a, inference direction: hash-grid encoding, density MLP, spherical
// harmonics of the view direction and colour MLP fused into ONE persistent kernel.  Replaces
//   ngp::NerfNetwork::inference_mixed_precision_impl (ngp nerf_network.h:159-232) and what it launches:
//   kernel_grid (tcnn encodings/grid.h:138-310), extract_expressions (nerf_network.h:86-105), kernel_mlp_fused x2
//   (tcnn src/fully_fused_mlp.cu:500-634), kernel_sh (tcnn encodings/spherical_harmonics.h:47-153), extract_density (34-46).
// The reference materialises the 128 x N fp16 feature matrix (+ 16 expression rows) in HBM -- 1.2 GB per 4M-sample batch
// written by the encoder and read back by the MLP; here the features go from the gathers straight into the shared-memory
// operand tile of the first GEMM and never leave the SM.
//
// Structure of a CTA (persistent, one per SM, 384 threads):
//   warps 4-11 "encoders": per 128-sample tile, 2 threads per sample x 8 levels each: 8 corner gathers (128-bit loads, table
//              L2-resident), trilinear interpolation with the reference's fp16 accumulation, one 16-byte store per level
//              into the K-major operand tile A0 (double buffered); the 16 expression floats fill K = 128..143.
//   warps 0-3  "MLP": one elected thread issues tcgen05.mma (M = 128 samples, N = 64 / 16, K = 16 per instruction, fp16
//              operands from shared memory, fp32 accumulators in TENSOR MEMORY); all four warps run the epilogues
//              (tcgen05.ld of their 32 TMEM lanes, ReLU, fp16 pairs, tcgen05.st into the columns the next layer's GEMM reads as its A operand:
//              after the first layer the activations never leave tensor memory).
//   density MLP 144 -> 64 -> 64 -> 16, colour MLP (16 density features | 16 SH) 32 -> 64 -> 64 -> 64 -> 16 (3 used): seven
//   dependent GEMMs per tile (~51 kFLOP/sample), hidden behind the gathers of the next tile through the mbarrier ring.
// Operand tiles use the canonical no-swizzle K-major UMMA layout: 8 rows x 16 bytes core matrices, 128 bytes apart along
// M/N (SBO), row-count x 16 bytes apart along K (LBO); a thread that owns a row writes 16-byte chunks that are conflict-free
// across the warp.
//
// Numerics: activations are rounded to fp16 between layers exactly where the reference stores them (its wmma accumulators are
// fp16 as well, so it rounds more often); accumulation here is fp32.  tests/test_gpu_network.py compares against an fp32
// PyTorch evaluation of the same fp16 weights / fp16 activations; tolerance stated there.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "ctx.h"
#include "grid_device.cuh"
#include "umma.cuh"

namespace insta {
namespace {
using namespace griddev;
using namespace umma;

constexpr int kTileM = 128;
constexpr int kMlpWarps = 4, kEncWarps = 8;
constexpr int kMlpThreads = kMlpWarps * 32, kEncThreads = kEncWarps * 32, kNetThreads = kMlpThreads + kEncThreads;
constexpr int kWidth = 64;        // n_neurons
constexpr int kInDensity = 144;   // 128 grid features + 16 expression dims
constexpr int kInRgb = 32;        // 16 density outputs + 16 SH
constexpr int kOutPad = 16;       // padded output width
// tensor memory: accumulators in columns 0-63; the A operands of the GEMMs after the first one: hidden activations (K = 64 -> 32 columns of fp16 pairs)
// and the colour MLP's input (16 density outputs | 16 SH -> 16 columns)
constexpr uint32_t kTmemCols = 128, kColH = 64, kColA3 = 96;

// parameter offsets (halves) in the trainer's order: density MLP, rgb MLP, grid (ngp nerf_network.h:492-525)
constexpr uint32_t kOffW1 = 0, kOffW2 = kOffW1 + kWidth * kInDensity, kOffW3 = kOffW2 + kWidth * kWidth;
constexpr uint32_t kOffV1 = kOffW3 + kOutPad * kWidth, kOffV2 = kOffV1 + kWidth * kInRgb, kOffV3 = kOffV2 + kWidth * kWidth;
constexpr uint32_t kOffV4 = kOffV3 + kWidth * kWidth, kOffGrid = kOffV4 + kOutPad * kWidth;
static_assert(kOffGrid == 25600, "MLP parameter count of the shipped config");

// shared-memory map (bytes)
constexpr uint32_t kChunkA = kTileM * 16;                       // one K-chunk (8 halves) of a 128-row operand tile
constexpr uint32_t kBytesA0 = (kInDensity / 8) * kChunkA;       // 36,864
constexpr uint32_t kSmA0 = 0;                                   // two buffers
constexpr uint32_t kSmW1 = kSmA0 + 2 * kBytesA0;
constexpr uint32_t kSmW2 = kSmW1 + kWidth * kInDensity * 2;
constexpr uint32_t kSmW3 = kSmW2 + kWidth * kWidth * 2;
constexpr uint32_t kSmV1 = kSmW3 + kOutPad * kWidth * 2;
constexpr uint32_t kSmV2 = kSmV1 + kWidth * kInRgb * 2;
constexpr uint32_t kSmV3 = kSmV2 + kWidth * kWidth * 2;
constexpr uint32_t kSmV4 = kSmV3 + kWidth * kWidth * 2;
constexpr uint32_t kSmBar = kSmV4 + kOutPad * kWidth * 2;       // 5 mbarriers + tmem base
constexpr uint32_t kSmTotal = kSmBar + 64;

__device__ __forceinline__ void mlp_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kMlpThreads) : "memory"); }

// D[128 x N] (TMEM) = A[128 x K] * B[N x K]^T, operands in the canonical layout described in the header
template <int N, int K>
__device__ __forceinline__ void issue_gemm(uint32_t tmem, uint32_t a_saddr, uint32_t b_saddr, uint32_t mma_bar) {
	constexpr uint32_t idesc = umma_idesc(kTileM, N);
#pragma unroll
	for (int j = 0; j < K / 16; ++j) {
		const uint64_t ad = umma_desc(a_saddr + j * 2 * kChunkA, kChunkA, 128);
		const uint64_t bd = umma_desc(b_saddr + j * 2 * (N * 16), N * 16, 128);
		umma_f16(tmem, ad, bd, idesc, j > 0 ? 1u : 0u);
	}
	umma_commit(mma_bar);
}

// the same with A in tensor memory (row in lane, two K elements per 32-bit column): the layers after the first one
template <int N, int K>
__device__ __forceinline__ void issue_gemm_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_saddr, uint32_t mma_bar) {
	constexpr uint32_t idesc = umma_idesc(kTileM, N);
#pragma unroll
	for (int j = 0; j < K / 16; ++j) {
		const uint64_t bd = umma_desc(b_saddr + j * 2 * (N * 16), N * 16, 128);
		umma_f16_ts(tmem_d, tmem_a + j * 8, bd, idesc, j > 0 ? 1u : 0u);
	}
	umma_commit(mma_bar);
}

// TMEM accumulators of this thread's row -> ReLU -> fp16 pairs -> the A-operand columns of the next layer, again in tensor memory (tcgen05.st): no
// shared-memory tile and no fence.proxy.async between two layers (with the tile in shared memory that fence -- MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, queued
// behind the encoder warps' gathers -- was 40 % of the MLP warps' time in the training kernel; profiles/r2_net_fwd_bwd_bulk_ncu_summary.txt)
template <int N_COLS, bool RELU>
__device__ __forceinline__ void epilogue_to_tmem(uint32_t taddr_d, uint32_t taddr_a, float* keep0 /* nullable: receives column 0 as fp16-rounded */) {
	constexpr int W = N_COLS >= 32 ? 32 : 16;
#pragma unroll
	for (int c0 = 0; c0 < N_COLS; c0 += W) {
		float v[W];
		if (W == 32) tmem_ld32(taddr_d + c0, v);
		else tmem_ld16(taddr_d + c0, v);
		if (RELU) {
#pragma unroll
			for (int k = 0; k < W; ++k) v[k] = fmaxf(v[k], 0.0f);
		}
		uint32_t p[W / 2];
#pragma unroll
		for (int k = 0; k < W / 2; ++k) p[k] = pack2(v[2 * k], v[2 * k + 1]);
		if (W == 32) tmem_st16(taddr_a + c0 / 2, p);
		else tmem_st8(taddr_a + c0 / 2, p);
		if (keep0 && c0 == 0) *keep0 = __half2float(__float2half_rn(v[0]));
	}
}

// degree-4 spherical harmonics of dir * 2 - 1 (tcnn spherical_harmonics.h:62-100), 16 coefficients as fp16
__device__ __forceinline__ void sh4(float x, float y, float z, uint4& lo, uint4& hi) {
	const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
	float o[16];
	o[0] = 0.28209479177387814f;
	o[1] = -0.48860251190291987f * y;
	o[2] = 0.48860251190291987f * z;
	o[3] = -0.48860251190291987f * x;
	o[4] = 1.0925484305920792f * xy;
	o[5] = -1.0925484305920792f * yz;
	o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
	o[7] = -1.0925484305920792f * xz;
	o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
	o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
	o[10] = 2.8906114426405538f * xy * z;
	o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
	o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
	o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
	o[14] = 1.4453057213202769f * z * (x2 - y2);
	o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
	lo = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
	hi = make_uint4(pack2(o[8], o[9]), pack2(o[10], o[11]), pack2(o[12], o[13]), pack2(o[14], o[15]));
}

// weights [N][K] row-major (tcnn fully_fused_mlp.cu:662-681) -> canonical K-major operand layout
__device__ __forceinline__ void stage_weights(unsigned char* dst, const __half* __restrict__ w, int N, int K) {
	const int chunks = N * (K / 8);
	for (int t = threadIdx.x; t < chunks; t += blockDim.x) {
		const int n = t / (K / 8), kc = t - n * (K / 8);
		const uint4 v = __ldg(reinterpret_cast<const uint4*>(w + (size_t)n * K + kc * 8));
		*reinterpret_cast<uint4*>(dst + kc * (N * 16) + (n >> 3) * 128 + (n & 7) * 16) = v;
	}
}

struct NetArgs {
	const float* records;
	uint32_t stride, n, n_extra;
	const __half* params;   // MLP weights then the grid table
	__half* out;
	uint32_t out_stride;    // 0: 4 halves per sample; else out[i + c * out_stride]
	const uint32_t* index;       // nullable: sample k of the launch is record / output index[k] (the trainer's lazily evaluated rounds)
	const float* uniform_extra;  // nullable: records carry no expression floats; every sample takes these n_extra values, or all ones when its holder is -2
	const uint32_t* n_ptr;  // nullable: device-side count, only min(n, *n_ptr) samples are evaluated
	uint32_t density_only;  // NerfNetwork::density: records are bare positions, expression inputs = 1, out = one density logit per sample
	GridLevels L;
};

__global__ void __launch_bounds__(kNetThreads, 1) nerf_network_fwd_kernel(const __grid_constant__ NetArgs a) {
	extern __shared__ __align__(1024) unsigned char smem[];
	const uint32_t sbase = smem_u32(smem);
	const uint32_t bar_full0 = sbase + kSmBar, bar_empty0 = sbase + kSmBar + 16, bar_mma = sbase + kSmBar + 32;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmBar + 40);
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint32_t n_samples = a.n_ptr ? min(a.n, __ldg(a.n_ptr)) : a.n;
	const uint32_t n_tiles = (n_samples + kTileM - 1) / kTileM;

	// ---- one-time setup: weights, barriers, tensor memory ----
	stage_weights(smem + kSmW1, a.params + kOffW1, kWidth, kInDensity);
	stage_weights(smem + kSmW2, a.params + kOffW2, kWidth, kWidth);
	stage_weights(smem + kSmW3, a.params + kOffW3, kOutPad, kWidth);
	stage_weights(smem + kSmV1, a.params + kOffV1, kWidth, kInRgb);
	stage_weights(smem + kSmV2, a.params + kOffV2, kWidth, kWidth);
	stage_weights(smem + kSmV3, a.params + kOffV3, kWidth, kWidth);
	stage_weights(smem + kSmV4, a.params + kOffV4, kOutPad, kWidth);
	if (threadIdx.x == 0) {
		mbar_init(bar_full0, kEncWarps);
		mbar_init(bar_full0 + 8, kEncWarps);
		mbar_init(bar_empty0, 1);
		mbar_init(bar_empty0 + 8, 1);
		mbar_init(bar_mma, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0) {
		__syncwarp();
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	fence_proxy_async();  // the staged weights are read by the tensor core through the async proxy
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp >= kMlpWarps) {
		// =================================== encoders ===================================
		const uint32_t e = threadIdx.x - kMlpThreads;
		const uint32_t row = e & (kTileM - 1), half_id = e >> 7;  // levels [8 half_id, 8 half_id + 8)
		const uint4* table = reinterpret_cast<const uint4*>(a.params + kOffGrid);
		uint32_t it = 0;
		// the position of this row's sample in the NEXT tile is requested a tile ahead: the records come from HBM, and a DRAM round trip at the head
		// of every tile (~1 of ~9 us) is time the encoder warps -- which bound the kernel -- do not gather
		const float* nrec = nullptr;
		float npx = 0.f, npy = 0.f, npz = 0.f;
		auto request = [&](uint32_t tile) {
			const uint32_t i = tile * kTileM + row;
			nrec = nullptr;
			if (tile < n_tiles && i < n_samples) {
				nrec = a.records + (size_t)(a.index ? __ldg(a.index + i) : i) * a.stride;
				npx = __ldg(nrec); npy = __ldg(nrec + 1); npz = __ldg(nrec + 2);
			}
		};
		request(blockIdx.x);
		for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
			const uint32_t b = it & 1u;
			const float* rec = nrec;
			const float px = npx, py = npy, pz = npz;
			request(tile + gridDim.x);
			mbar_wait(bar_empty0 + 8 * b, ((it >> 1) & 1u) ^ 1u);
			unsigned char* A0 = smem + kSmA0 + b * kBytesA0 + (row >> 3) * 128 + (row & 7) * 16;
			if (rec) {
#pragma unroll 2
				for (uint32_t l = 0; l < 8; ++l) {
					const uint32_t level = half_id * 8 + l;
					__half2 acc[4] = {__half2(), __half2(), __half2(), __half2()};
					if (level < a.L.n_levels) {
						const Corner c = pos_fract3(px, py, pz, a.L.scale[level]);
						if (a.L.hashed[level]) encode_one<true>(table + a.L.offset[level], c, a.L.size[level], 0, 0, acc);
						else encode_one<false>(table + a.L.offset[level], c, a.L.size[level], a.L.mul1[level], a.L.mul2[level], acc);
					}
					*reinterpret_cast<uint4*>(A0 + level * kChunkA) =
					    make_uint4(*reinterpret_cast<uint32_t*>(&acc[0]), *reinterpret_cast<uint32_t*>(&acc[1]), *reinterpret_cast<uint32_t*>(&acc[2]),
					               *reinterpret_cast<uint32_t*>(&acc[3]));
				}
				// expression dims: floats [stride - n_extra, stride) -> K = 128 + 8 half_id .. (extract_expressions)
				float ex[8];
				const bool ones = a.density_only || (a.uniform_extra && __ldg(rec + 4) == -2.0f);
				const float* exsrc = a.uniform_extra ? a.uniform_extra : rec + a.stride - a.n_extra;
#pragma unroll
				for (uint32_t k = 0; k < 8; ++k) {
					const uint32_t d = half_id * 8 + k;
					ex[k] = d < a.n_extra ? (ones ? 1.0f : __ldg(exsrc + d)) : 0.0f;
				}
				*reinterpret_cast<uint4*>(A0 + (16 + half_id) * kChunkA) = make_uint4(pack2(ex[0], ex[1]), pack2(ex[2], ex[3]), pack2(ex[4], ex[5]), pack2(ex[6], ex[7]));
			} else {
				const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
				for (uint32_t l = 0; l < 8; ++l) *reinterpret_cast<uint4*>(A0 + (half_id * 8 + l) * kChunkA) = z;
				*reinterpret_cast<uint4*>(A0 + (16 + half_id) * kChunkA) = z;
			}
			fence_proxy_async();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_full0 + 8 * b);
		}
	} else {
		// =================================== MLP warps ===================================
		const uint32_t row = threadIdx.x;                        // == TMEM lane
		const uint32_t taddr = tmem_base + ((warp * 32u) << 16);  // this warp's lane quadrant
		const uint32_t tH = tmem_base + kColH, tA3 = tmem_base + kColA3;
		uint32_t mma_phase = 0, it = 0;
		const bool issuer = threadIdx.x == 0;
// what the epilogue stored into tensor memory is complete before the issuer's next GEMM reads it
#define STAGE_DONE()   \
	tmem_st_wait();    \
	tc_fence_before(); \
	mlp_bar_sync();    \
	tc_fence_after()
		for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
			const uint32_t b = it & 1u;
			const uint32_t i = tile * kTileM + row;
			const uint32_t si = (a.index && i < n_samples) ? __ldg(a.index + i) : i;  // record / output slot of this row's sample
			// SH of the view direction while the first GEMM is in flight
			float dx = 0.5f, dy = 0.5f, dz = 0.5f;
			if (i < n_samples && !a.density_only) {
				const float* rec = a.records + (size_t)si * a.stride;
				dx = __ldg(rec + 8); dy = __ldg(rec + 9); dz = __ldg(rec + 10);
			}
			// ---- density layer 1: A0 (144, shared memory) -> 64 ----
			mbar_wait(bar_full0 + 8 * b, (it >> 1) & 1u);
			tc_fence_after();
			if (issuer) issue_gemm<kWidth, kInDensity>(tmem_base, sbase + kSmA0 + b * kBytesA0, sbase + kSmW1, bar_mma);
			{
				uint4 lo, hi;
				sh4(dx * 2.f - 1.f, dy * 2.f - 1.f, dz * 2.f - 1.f, lo, hi);
				const uint32_t p[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
				tmem_st8(taddr + kColA3 + 8, p);
			}
			mbar_wait(bar_mma, mma_phase); mma_phase ^= 1u;
			tc_fence_after();
			if (issuer) mbar_arrive(bar_empty0 + 8 * b);  // A0[b] has been consumed: the encoders may refill it
			epilogue_to_tmem<kWidth, true>(taddr, taddr + kColH, nullptr);
			STAGE_DONE();
			// ---- density layer 2: 64 -> 64 ----
			if (issuer) issue_gemm_ts<kWidth, kWidth>(tmem_base, tH, sbase + kSmW2, bar_mma);
			mbar_wait(bar_mma, mma_phase); mma_phase ^= 1u;
			tc_fence_after();
			epilogue_to_tmem<kWidth, true>(taddr, taddr + kColH, nullptr);
			STAGE_DONE();
			// ---- density output: 64 -> 16, no activation; column 0 is the density logit ----
			if (issuer) issue_gemm_ts<kOutPad, kWidth>(tmem_base, tH, sbase + kSmW3, bar_mma);
			mbar_wait(bar_mma, mma_phase); mma_phase ^= 1u;
			tc_fence_after();
			float sigma = 0.f;
			epilogue_to_tmem<kOutPad, false>(taddr, taddr + kColA3, &sigma);
			STAGE_DONE();
			// ---- colour layer 1: (16 density | 16 SH) -> 64 ----
			if (issuer) issue_gemm_ts<kWidth, kInRgb>(tmem_base, tA3, sbase + kSmV1, bar_mma);
			mbar_wait(bar_mma, mma_phase); mma_phase ^= 1u;
			tc_fence_after();
			epilogue_to_tmem<kWidth, true>(taddr, taddr + kColH, nullptr);
			STAGE_DONE();
			// ---- colour layers 2, 3 ----
			if (issuer) issue_gemm_ts<kWidth, kWidth>(tmem_base, tH, sbase + kSmV2, bar_mma);
			mbar_wait(bar_mma, mma_phase); mma_phase ^= 1u;
			tc_fence_after();
			epilogue_to_tmem<kWidth, true>(taddr, taddr + kColH, nullptr);
			STAGE_DONE();
			if (issuer) issue_gemm_ts<kWidth, kWidth>(tmem_base, tH, sbase + kSmV3, bar_mma);
			mbar_wait(bar_mma, mma_phase); mma_phase ^= 1u;
			tc_fence_after();
			epilogue_to_tmem<kWidth, true>(taddr, taddr + kColH, nullptr);
			STAGE_DONE();
			// ---- colour output: 64 -> 16 (3 used) ----
			if (issuer) issue_gemm_ts<kOutPad, kWidth>(tmem_base, tH, sbase + kSmV4, bar_mma);
			mbar_wait(bar_mma, mma_phase); mma_phase ^= 1u;
			tc_fence_after();
			float o[16];
			tmem_ld16(taddr, o);
			tc_fence_before();
			if (i < n_samples) {
				const __half2 rg = __floats2half2_rn(o[0], o[1]), bs = __floats2half2_rn(o[2], sigma);
				if (a.density_only) {
					a.out[si] = __high2half(bs);
				} else if (a.out_stride == 0) {
					*reinterpret_cast<uint2*>(a.out + (size_t)si * 4) = make_uint2(*reinterpret_cast<const uint32_t*>(&rg), *reinterpret_cast<const uint32_t*>(&bs));
				} else {
					a.out[si] = __low2half(rg);
					a.out[si + (size_t)a.out_stride] = __high2half(rg);
					a.out[si + 2 * (size_t)a.out_stride] = __low2half(bs);
					a.out[si + 3 * (size_t)a.out_stride] = __high2half(bs);
				}
			}
			// the next tile's first GEMM overwrites the accumulator columns: all reads above must have completed
			mlp_bar_sync();
		}
#undef STAGE_DONE
	}

	tc_fence_before();
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}


// ======================================================================================================================
// The same network with LANE-PAIR encoders (default; nerf_network_fwd_kernel above = INSTA_NET_FWD=v1).
// The kernel above is bound by the L1's sector rate (88 % of peak: 32 lanes = 32 samples = 32 sectors per gather instruction).  hash_fwd_pairs_kernel's mapping --
// lanes 2j / 2j + 1 = the two x-sides of a cell: neighbouring table entries, one sector half of the time -- removes a quarter of that work but costs shuffles and
// issue slots, which 8 encoder warps do not have (measured in round 2: 1.92 -> 2.28 ms).  With the activations in tensor memory the MLP side needs 4 light warps and no
// shared-memory tiles, so the CTA affords 12 encoder warps.  They no longer own rows: the work of a tile is 32 TASKS (level x half of the tile's rows: 64 samples x
// 4 gathers per lane), a warp takes the tasks w, w + 12, w + 24, ... of the endless sequence tile 0 task 0..31, tile 1 task 0..31, ... -- no assignment problem (32 is not
// a multiple of 12) and no atomics; a tile is complete when its 32 tasks have arrived on its mbarrier.  Positions and the 16 expression inputs of a tile are staged by the
// MLP threads (thread = row; they read the record anyway for the view direction), two tiles ahead, behind the first layer's epilogue.
constexpr int kEncWarpsDefault = 12;
constexpr uint32_t kTasksPerTile = 32;
constexpr uint32_t kSmPos = kSmBar + 64, kSmTotalP = kSmPos + 2 * kTileM * 3 * 4;

template <bool HASHED>
__device__ __forceinline__ void encoder_task(const GridLevels& L, const uint4* __restrict__ table, uint32_t level, const float* __restrict__ pos, uint32_t hf, uint32_t lane,
                                             unsigned char* A0) {
	const uint32_t b = lane & 1u, g = lane >> 1;
	const uint32_t size = L.size[level], mul1 = L.mul1[level], mul2 = L.mul2[level];
	const float scale = L.scale[level];
	table += L.offset[level];
	// this lane's rows: 64 hf + g + 16 p -- for a fixed p the 16 lane pairs write 256 contiguous bytes of the operand tile
	__half2 acc[4][2];
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		uint4 v[2][4];
		float w[2][4];
#pragma unroll
		for (int q = 0; q < 2; ++q) {
			const float* sp = pos + 3u * (64u * hf + g + 16u * (2 * h + q));
			pair_load<HASHED>(table, pos_fract3(sp[0], sp[1], sp[2], scale), b, size, mul1, mul2, v[q], w[q]);
		}
#pragma unroll
		for (int q = 0; q < 2; ++q) pair_accumulate(v[q], w[q], b, acc[2 * h + q][0], acc[2 * h + q][1]);
	}
#pragma unroll
	for (int p = 0; p < 4; ++p) {
		const uint32_t row = 64u * hf + g + 16u * p;
		*reinterpret_cast<uint2*>(A0 + level * kChunkA + row * 16u + 8u * b) =
		    make_uint2(*reinterpret_cast<const uint32_t*>(&acc[p][0]), *reinterpret_cast<const uint32_t*>(&acc[p][1]));  // features 4b .. 4b+3 of the level
	}
}

template <int kEncWarpsP>
__global__ void __launch_bounds__(kMlpThreads + kEncWarpsP * 32, 1) nerf_network_fwd_pairs_kernel(const __grid_constant__ NetArgs a) {
	extern __shared__ __align__(1024) unsigned char smem[];
	const uint32_t sbase = smem_u32(smem);
	const uint32_t bar_full0 = sbase + kSmBar, bar_empty0 = sbase + kSmBar + 16, bar_mma = sbase + kSmBar + 32;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmBar + 40);
	float* pos = reinterpret_cast<float*>(smem + kSmPos);  // [2][128][3]
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint32_t n_samples = a.n_ptr ? min(a.n, __ldg(a.n_ptr)) : a.n;
	const uint32_t n_tiles = (n_samples + kTileM - 1) / kTileM;

	stage_weights(smem + kSmW1, a.params + kOffW1, kWidth, kInDensity);
	stage_weights(smem + kSmW2, a.params + kOffW2, kWidth, kWidth);
	stage_weights(smem + kSmW3, a.params + kOffW3, kOutPad, kWidth);
	stage_weights(smem + kSmV1, a.params + kOffV1, kWidth, kInRgb);
	stage_weights(smem + kSmV2, a.params + kOffV2, kWidth, kWidth);
	stage_weights(smem + kSmV3, a.params + kOffV3, kWidth, kWidth);
	stage_weights(smem + kSmV4, a.params + kOffV4, kOutPad, kWidth);
	if (threadIdx.x == 0) {
		mbar_init(bar_full0, kTasksPerTile);
		mbar_init(bar_full0 + 8, kTasksPerTile);
		mbar_init(bar_empty0, 1);
		mbar_init(bar_empty0 + 8, 1);
		mbar_init(bar_mma, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0) {
		__syncwarp();
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	// MLP thread = tile row: position and the 16 expression inputs (extract_expressions) of that row's sample in tile `tile`
	struct RowIn {
		float px, py, pz;
		uint32_t ex[8];
	};
	auto fetch = [&](uint32_t tile, uint32_t row) {
		RowIn r{};
		const uint32_t i = tile * kTileM + row;
		if (tile < n_tiles && i < n_samples) {
			const float* rec = a.records + (size_t)(a.index ? __ldg(a.index + i) : i) * a.stride;
			r.px = __ldg(rec); r.py = __ldg(rec + 1); r.pz = __ldg(rec + 2);
			const bool ones = a.density_only || (a.uniform_extra && __ldg(rec + 4) == -2.0f);
			const float* exsrc = a.uniform_extra ? a.uniform_extra : rec + a.stride - a.n_extra;
#pragma unroll
			for (uint32_t k = 0; k < 8; ++k) {
				const float e0 = 2 * k < a.n_extra ? (ones ? 1.0f : __ldg(exsrc + 2 * k)) : 0.0f;
				const float e1 = 2 * k + 1 < a.n_extra ? (ones ? 1.0f : __ldg(exsrc + 2 * k + 1)) : 0.0f;
				r.ex[k] = pack2(e0, e1);
			}
		}
		return r;
	};
	auto stage = [&](uint32_t b, uint32_t row, const RowIn& r) {
		float* P = pos + (b * kTileM + row) * 3u;
		P[0] = r.px; P[1] = r.py; P[2] = r.pz;
		unsigned char* A = smem + kSmA0 + b * kBytesA0 + row * 16u;
		*reinterpret_cast<uint4*>(A + 16 * kChunkA) = make_uint4(r.ex[0], r.ex[1], r.ex[2], r.ex[3]);
		*reinterpret_cast<uint4*>(A + 17 * kChunkA) = make_uint4(r.ex[4], r.ex[5], r.ex[6], r.ex[7]);
	};
	if (warp < kMlpWarps) {
		stage(0, threadIdx.x, fetch(blockIdx.x, threadIdx.x));
		stage(1, threadIdx.x, fetch(blockIdx.x + gridDim.x, threadIdx.x));
	}
	fence_proxy_async();  // the staged weights and expression chunks are read by the tensor core through the async proxy
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp >= kMlpWarps) {
		// =================================== encoders: tasks w, w + 12, ... ===================================
		const uint4* table = reinterpret_cast<const uint4*>(a.params + kOffGrid);
		for (uint32_t id = warp - kMlpWarps;; id += kEncWarpsP) {
			const uint32_t it = id / kTasksPerTile, sub = id % kTasksPerTile;
			if ((uint64_t)blockIdx.x + (uint64_t)it * gridDim.x >= n_tiles) break;
			const uint32_t b = it & 1u, level = sub & 15u, hf = sub >> 4;
			mbar_wait(bar_empty0 + 8 * b, ((it >> 1) & 1u) ^ 1u);
			unsigned char* A0 = smem + kSmA0 + b * kBytesA0;
			const float* P = pos + b * kTileM * 3u;
			if (level < a.L.n_levels) {
				if (a.L.hashed[level]) encoder_task<true>(a.L, table, level, P, hf, lane, A0);
				else encoder_task<false>(a.L, table, level, P, hf, lane, A0);
			} else {
#pragma unroll
				for (int p = 0; p < 4; ++p)
					*reinterpret_cast<uint2*>(A0 + level * kChunkA + (64u * hf + (lane >> 1) + 16u * p) * 16u + 8u * (lane & 1u)) = make_uint2(0u, 0u);
			}
			fence_proxy_async();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_full0 + 8 * b);
		}
	} else {
		// =================================== MLP warps ===================================
		const uint32_t row = threadIdx.x;                        // == TMEM lane
		const uint32_t taddr = tmem_base + ((warp * 32u) << 16);  // this warp's lane quadrant
		const uint32_t tH = tmem_base + kColH, tA3 = tmem_base + kColA3;
		uint32_t mma_phase = 0, it = 0;
		const bool issuer = threadIdx.x == 0;
#define STAGE_DONE()   \
	tmem_st_wait();    \
	tc_fence_before(); \
	mlp_bar_sync();    \
	tc_fence_after()
#define MMA_WAIT()                 \
	mbar_wait(bar_mma, mma_phase); \
	mma_phase ^= 1u;               \
	tc_fence_after()
		for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
			const uint32_t b = it & 1u;
			const uint32_t i = tile * kTileM + row;
			const uint32_t si = (a.index && i < n_samples) ? __ldg(a.index + i) : i;  // record / output slot of this row's sample
			float dx = 0.5f, dy = 0.5f, dz = 0.5f;
			if (i < n_samples && !a.density_only) {
				const float* rec = a.records + (size_t)si * a.stride;
				dx = __ldg(rec + 8); dy = __ldg(rec + 9); dz = __ldg(rec + 10);
			}
			const RowIn next = fetch(tile + 2 * gridDim.x, row);  // the tile that will use this tile's buffers; consumed behind the first epilogue
			// ---- density layer 1: A0 (144, shared memory) -> 64 ----
			mbar_wait(bar_full0 + 8 * b, (it >> 1) & 1u);
			tc_fence_after();
			if (issuer) issue_gemm<kWidth, kInDensity>(tmem_base, sbase + kSmA0 + b * kBytesA0, sbase + kSmW1, bar_mma);
			{
				uint4 lo, hi;
				sh4(dx * 2.f - 1.f, dy * 2.f - 1.f, dz * 2.f - 1.f, lo, hi);
				const uint32_t p[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
				tmem_st8(taddr + kColA3 + 8, p);
			}
			MMA_WAIT();
			epilogue_to_tmem<kWidth, true>(taddr, taddr + kColH, nullptr);
			// A0[b] and its positions have been consumed (all 32 tasks arrived before the GEMM was issued, and the GEMM is complete): restage them for tile + 2 * grid
			stage(b, row, next);
			fence_proxy_async();
			STAGE_DONE();
			if (issuer) {
				mbar_arrive(bar_empty0 + 8 * b);
				issue_gemm_ts<kWidth, kWidth>(tmem_base, tH, sbase + kSmW2, bar_mma);
			}
			MMA_WAIT();
			epilogue_to_tmem<kWidth, true>(taddr, taddr + kColH, nullptr);
			STAGE_DONE();
			// ---- density output: 64 -> 16, no activation; column 0 is the density logit ----
			if (issuer) issue_gemm_ts<kOutPad, kWidth>(tmem_base, tH, sbase + kSmW3, bar_mma);
			MMA_WAIT();
			float sigma = 0.f;
			epilogue_to_tmem<kOutPad, false>(taddr, taddr + kColA3, &sigma);
			STAGE_DONE();
			// ---- colour layer 1: (16 density | 16 SH) -> 64 ----
			if (issuer) issue_gemm_ts<kWidth, kInRgb>(tmem_base, tA3, sbase + kSmV1, bar_mma);
			MMA_WAIT();
			epilogue_to_tmem<kWidth, true>(taddr, taddr + kColH, nullptr);
			STAGE_DONE();
			// ---- colour layers 2, 3 ----
			if (issuer) issue_gemm_ts<kWidth, kWidth>(tmem_base, tH, sbase + kSmV2, bar_mma);
			MMA_WAIT();
			epilogue_to_tmem<kWidth, true>(taddr, taddr + kColH, nullptr);
			STAGE_DONE();
			if (issuer) issue_gemm_ts<kWidth, kWidth>(tmem_base, tH, sbase + kSmV3, bar_mma);
			MMA_WAIT();
			epilogue_to_tmem<kWidth, true>(taddr, taddr + kColH, nullptr);
			STAGE_DONE();
			// ---- colour output: 64 -> 16 (3 used) ----
			if (issuer) issue_gemm_ts<kOutPad, kWidth>(tmem_base, tH, sbase + kSmV4, bar_mma);
			MMA_WAIT();
			float o[16];
			tmem_ld16(taddr, o);
			tc_fence_before();
			if (i < n_samples) {
				const __half2 rg = __floats2half2_rn(o[0], o[1]), bs = __floats2half2_rn(o[2], sigma);
				if (a.density_only) {
					a.out[si] = __high2half(bs);
				} else if (a.out_stride == 0) {
					*reinterpret_cast<uint2*>(a.out + (size_t)si * 4) = make_uint2(*reinterpret_cast<const uint32_t*>(&rg), *reinterpret_cast<const uint32_t*>(&bs));
				} else {
					a.out[si] = __low2half(rg);
					a.out[si + (size_t)a.out_stride] = __high2half(rg);
					a.out[si + 2 * (size_t)a.out_stride] = __low2half(bs);
					a.out[si + 3 * (size_t)a.out_stride] = __high2half(bs);
				}
			}
			// the next tile's first GEMM overwrites the accumulator columns: all reads above must have completed
			mlp_bar_sync();
		}
#undef STAGE_DONE
#undef MMA_WAIT
	}

	tc_fence_before();
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

// both launchers: the lane-pair kernel unless INSTA_NET_FWD=v1
static int launch_net(const NetArgs& a, uint32_t grid, cudaStream_t s) {
	static const bool v1 = [] { const char* e = getenv("INSTA_NET_FWD"); return e && e[0] == 'v'; }();
	if (v1) {
		INSTA_CUDA(cudaFuncSetAttribute(nerf_network_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmTotal));
		nerf_network_fwd_kernel<<<grid, kNetThreads, kSmTotal, s>>>(a);
	} else {
		static const int enc = [] { const char* e = getenv("INSTA_NET_ENC_WARPS"); return e ? atoi(e) : kEncWarpsDefault; }();  // A/B: 8, 10, 12, 14, 16
#define LAUNCH_P(W)                                                                                                                       \
	INSTA_CUDA(cudaFuncSetAttribute(nerf_network_fwd_pairs_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmTotalP)); \
	nerf_network_fwd_pairs_kernel<W><<<grid, kMlpThreads + W * 32, kSmTotalP, s>>>(a)
		if (enc == 8) { LAUNCH_P(8); }
		else if (enc == 10) { LAUNCH_P(10); }
		else if (enc == 14) { LAUNCH_P(14); }
		else if (enc == 16) { LAUNCH_P(16); }
		else { LAUNCH_P(12); }
#undef LAUNCH_P
	}
	return INSTA_OK;
}

}  // namespace

uint32_t network_n_params(const insta_ctx* ctx) { return kOffGrid + ctx->levels.offset[ctx->levels.n_levels] * 8u; }

int launch_network_fwd(insta_ctx* ctx, const float* d_records, uint32_t stride, uint32_t n, const void* d_params, void* d_out, uint32_t out_stride, cudaStream_t s,
                       const uint32_t* d_n, const float* d_uniform_extra, const uint32_t* d_index) {
	if (n == 0) return INSTA_OK;
	NetArgs a{};
	a.records = d_records;
	a.stride = stride;
	a.n = n;
	a.n_extra = ctx->p.exp_params + ctx->p.geo_params;
	a.params = (const __half*)d_params;
	a.out = (__half*)d_out;
	a.out_stride = out_stride;
	a.n_ptr = d_n;
	a.uniform_extra = d_uniform_extra;
	a.index = d_index;
	a.L = ctx->levels;
	if (a.n_extra > 16) INSTA_FAIL(INSTA_ERR_INVALID, "network: more than 16 extra dims are not supported by the density MLP input (144 = 128 + 16)");
	const uint32_t n_tiles = (n + kTileM - 1) / kTileM;
	const uint32_t grid = std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count);
	if (int rc = launch_net(a, grid, s)) return rc;
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

// NerfNetwork::density (ngp nerf_network.h:385-414): hash grid + expression rows (all 1, extract_expressions_random) + density MLP; the first
// output of the density MLP is the density logit.  Runs the same fused kernel (the colour MLP's result is dropped; the kernel is bound by the
// grid gathers, not by the GEMMs).
int launch_network_density(insta_ctx* ctx, const float* d_xyz, uint32_t stride, uint32_t n, const void* d_params, void* d_logits, cudaStream_t s) {
	if (n == 0) return INSTA_OK;
	NetArgs a{};
	a.records = d_xyz;
	a.stride = stride;
	a.n = n;
	a.n_extra = ctx->p.exp_params + ctx->p.geo_params;
	a.params = (const __half*)d_params;
	a.out = (__half*)d_logits;
	a.out_stride = 0;
	a.n_ptr = nullptr;
	a.density_only = 1;
	a.L = ctx->levels;
	if (a.n_extra > 16) INSTA_FAIL(INSTA_ERR_INVALID, "network: more than 16 extra dims are not supported by the density MLP input (144 = 128 + 16)");
	const uint32_t n_tiles = (n + kTileM - 1) / kTileM;
	if (int rc = launch_net(a, std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count), s)) return rc;
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

}  // namespace insta
