// The NeRF network of the INSTA hot path on sm_100a, training direction: forward + backward of hash grid, density MLP, SH and
// colour MLP on the compacted batch (2^18 samples).  Replaces ngp::NerfNetwork::forward_impl / backward_impl (ngp nerf_network.h:
// 238-383) and what they launch: kernel_grid, kernel_mlp_fused (forward, activations kept), kernel_mlp_fused_backward
// (tcnn src/fully_fused_mlp.cu:151-313), the seven split-k CUTLASS weight-gradient GEMMs (755-855), add_density_gradient /
// extract_rgb glue, and kernel_grid_backward (tcnn encodings/grid.h:312-435, via hashgrid.cu).
//
// Two tcgen05 kernels:
//   net_fwd_bwd_kernel  persistent, same encoder / MLP warp split as the inference kernel (mlp.cu).  Per 128-sample tile: the
//                       forward chain (7 GEMMs), then the data-gradient chain (7 GEMMs against TRANSPOSED weights staged once
//                       per CTA).  ReLU masks stay in registers (64 bits per hidden layer and sample); every layer's input X_l
//                       and output gradient dY_l is written once to a global fp16 scratch (row-major) for the weight gradients;
//                       dL/d(grid features) leaves as the SoA matrix kernel hash_bwd_kernel scatters.
//   net_wgrad_kernel    dW_l = dY_l^T X_l for all seven layers in one pass over the samples: both operands are used in their
//                       MN-major (transposed) UMMA form, which is exactly the scratch copied in 16-byte chunks; all
//                       seven fp32 accumulators (496 columns) live in tensor memory for the whole kernel; one fp32 atomicAdd
//                       pass per CTA at the end.  M is padded to 128 by letting the descriptor run into the next layers' tiles
//                       (rows >= n_out are never read back).
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
constexpr int kEncWarps = 8;
// MLP warps: 4 x SPLIT.  A warp reads the 32 tensor-memory lanes (tile rows) 32 (w % 4) ..; with SPLIT = 2 the warps w and w + 4 share those rows and take
// half of a layer's columns each, so an epilogue -- the part of the 15-stage dependent chain of a tile that is not the GEMM -- takes half as long.
constexpr int kSplit = 2;
constexpr int kWidth = 64, kInDensity = 144, kInRgb = 32, kOutPad = 16;
constexpr uint32_t kOffW1 = 0, kOffW2 = kOffW1 + kWidth * kInDensity, kOffW3 = kOffW2 + kWidth * kWidth;
constexpr uint32_t kOffV1 = kOffW3 + kOutPad * kWidth, kOffV2 = kOffV1 + kWidth * kInRgb, kOffV3 = kOffV2 + kWidth * kWidth;
constexpr uint32_t kOffV4 = kOffV3 + kWidth * kWidth, kOffGrid = kOffV4 + kOutPad * kWidth;

// per-layer shapes, layer order = parameter order
__constant__ const uint32_t c_nout[7] = {64, 64, 16, 64, 64, 64, 16};
__constant__ const uint32_t c_kin[7] = {144, 64, 64, 32, 64, 64, 64};
__constant__ const uint32_t c_woff[7] = {kOffW1, kOffW2, kOffW3, kOffV1, kOffV2, kOffV3, kOffV4};
// column offsets (halves) of X_l / dY_l inside one scratch row
constexpr uint32_t kXCols = 496, kDyCols = 352;
// scratch layout: [tile][chunk][row][8 halves]; halves per chunk of a tile and chunks per tile
constexpr uint32_t kChunkHalves = kTileM * 8, kXChunks = kXCols / 8, kDyChunks = kDyCols / 8, kFeatChunks = 16;
__constant__ const uint32_t c_xoff[7] = {0, 144, 208, 272, 304, 368, 432};
__constant__ const uint32_t c_dyoff[7] = {0, 64, 128, 144, 208, 272, 336};

constexpr uint32_t kChunkA = kTileM * 16;
constexpr uint32_t kBytesA0 = (kInDensity / 8) * kChunkA, kBytesH = (kWidth / 8) * kChunkA, kBytesA3 = (kInRgb / 8) * kChunkA;
constexpr uint32_t kWBytes = kOffGrid * 2;  // 51,200
constexpr uint32_t kSmA0 = 0, kSmH = kSmA0 + 2 * kBytesA0, kSmA3 = kSmH + 2 * kBytesH, kSmW = kSmA3 + kBytesA3, kSmWt = kSmW + kWBytes;
constexpr uint32_t kSmBar = kSmWt + kWBytes, kSmTotal = kSmBar + 64;
constexpr uint32_t kTmemCols = 256;

template <int THREADS>
__device__ __forceinline__ void mlp_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }

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

// weights [N][K] row-major -> K-major operand tile (B of the forward GEMMs)
__device__ __forceinline__ void stage_w(unsigned char* dst, const __half* __restrict__ w, int N, int K) {
	for (int t = threadIdx.x; t < N * (K / 8); t += blockDim.x) {
		const int n = t / (K / 8), kc = t - n * (K / 8);
		*reinterpret_cast<uint4*>(dst + kc * (N * 16) + (n >> 3) * 128 + (n & 7) * 16) = __ldg(reinterpret_cast<const uint4*>(w + (size_t)n * K + kc * 8));
	}
}
// the same weights transposed: operand B of dX = dY W is Bt[n = in][k = out] = W[out][in], K-major in `out`
__device__ __forceinline__ void stage_wt(unsigned char* dst, const __half* __restrict__ w, int N_out, int K_in) {
	for (int t = threadIdx.x; t < K_in * (N_out / 8); t += blockDim.x) {
		const int i = t / (N_out / 8), oc = t - i * (N_out / 8);
		__half v[8];
#pragma unroll
		for (int e = 0; e < 8; ++e) v[e] = w[(size_t)(oc * 8 + e) * K_in + i];
		*reinterpret_cast<uint4*>(dst + oc * (K_in * 16) + (i >> 3) * 128 + (i & 7) * 16) = *reinterpret_cast<const uint4*>(v);
	}
}

// W (16 or 32) consecutive accumulator columns of this thread's row
template <int W>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
	if (W == 32) tmem_ld32(taddr, v);
	else tmem_ld16(taddr, v);
}

// forward epilogue: N_COLS columns of a TMEM row -> (ReLU, mask bits: one word per 32 columns) -> fp16 -> next operand tile (16-byte chunks from `chunk0`)
template <int N_COLS, bool RELU>
__device__ __forceinline__ void epi_fwd(uint32_t taddr, unsigned char* tile, uint32_t chunk0, uint32_t row, uint32_t* mask, float* keep0) {
	constexpr int W = N_COLS >= 32 ? 32 : 16;
	unsigned char* dst = tile + (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
	for (int c0 = 0; c0 < N_COLS; c0 += W) {
		float v[W];
		tmem_ld_cols<W>(taddr + c0, v);
		if (RELU) {
			uint32_t bits = 0;
#pragma unroll
			for (int k = 0; k < W; ++k) {
				bits |= (v[k] > 0.0f ? 1u : 0u) << k;
				v[k] = fmaxf(v[k], 0.0f);
			}
			mask[c0 / 32] = bits;
		}
#pragma unroll
		for (int j = 0; j < W / 8; ++j) {
			const uint4 q = make_uint4(pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]), pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
			*reinterpret_cast<uint4*>(dst + (chunk0 + c0 / 8 + j) * kChunkA) = q;
		}
		if (keep0 && c0 == 0) *keep0 = v[0];
	}
}

// backward epilogue: TMEM row -> (mask) -> fp16 -> operand tile of the next data-gradient GEMM.
// add0 is added to column 0 before rounding (the density gradient joining the colour MLP's input gradient).
template <int N_COLS, bool MASK>
__device__ __forceinline__ void epi_bwd(uint32_t taddr, unsigned char* tile, uint32_t chunk0, uint32_t row, const uint32_t* mask, float add0) {
	constexpr int W = N_COLS >= 32 ? 32 : 16;
	unsigned char* dst = tile + (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
	for (int c0 = 0; c0 < N_COLS; c0 += W) {
		float v[W];
		tmem_ld_cols<W>(taddr + c0, v);
		if (MASK) {
			const uint32_t bits = mask[c0 / 32];
#pragma unroll
			for (int k = 0; k < W; ++k) v[k] = ((bits >> k) & 1u) ? v[k] : 0.0f;
		}
		if (c0 == 0) v[0] += add0;
#pragma unroll
		for (int j = 0; j < W / 8; ++j) {
			const uint4 q = make_uint4(pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]), pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
			*reinterpret_cast<uint4*>(dst + (chunk0 + c0 / 8 + j) * kChunkA) = q;
		}
	}
}

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

struct BwdArgs {
	const float* records;
	uint32_t stride, n, n_extra, n_pad;  // n_pad: rows of the scratch matrices (multiple of 128)
	const __half* params;
	const __half* dloss;    // dloss_width halves per sample: d(r, g, b, sigma)
	uint32_t dloss_width;
	__half* out;            // nullable: forward output, 4 halves per sample
	__half* xs;             // [n_pad / 128][kXChunks][128][8]
	__half* dys;            // [n_pad / 128][kDyChunks][128][8]
	__half* dfeat;          // [n_pad / 128][16 levels][128][8]: dL/d(grid features)
	GridLevels L;
};

template <int SPLIT>
__global__ void __launch_bounds__((4 * SPLIT + kEncWarps) * 32, 1) net_fwd_bwd_kernel(const __grid_constant__ BwdArgs a) {
	constexpr int kMlpWarps = 4 * SPLIT, kMlpThreads = kMlpWarps * 32;
	extern __shared__ __align__(1024) unsigned char smem[];
	const uint32_t sbase = smem_u32(smem);
	const uint32_t bar_full0 = sbase + kSmBar, bar_empty0 = sbase + kSmBar + 16, bar_mma = sbase + kSmBar + 32;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmBar + 40);
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint32_t n_tiles = (a.n + kTileM - 1) / kTileM;

	for (int l = 0; l < 7; ++l) {
		stage_w(smem + kSmW + c_woff[l] * 2, a.params + c_woff[l], (int)c_nout[l], (int)c_kin[l]);
		stage_wt(smem + kSmWt + c_woff[l] * 2, a.params + c_woff[l], (int)c_nout[l], (int)c_kin[l]);
	}
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
	fence_proxy_async();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp >= kMlpWarps) {
		// =================================== encoders (as in mlp.cu; the MLP side sends the finished tile to the X scratch) ===================================
		const uint32_t e = threadIdx.x - kMlpThreads;
		const uint32_t row = e & (kTileM - 1), half_id = e >> 7;
		const uint4* table = reinterpret_cast<const uint4*>(a.params + kOffGrid);
		uint32_t it = 0;
		for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
			const uint32_t b = it & 1u;
			mbar_wait(bar_empty0 + 8 * b, ((it >> 1) & 1u) ^ 1u);
			unsigned char* A0 = smem + kSmA0 + b * kBytesA0 + (row >> 3) * 128 + (row & 7) * 16;
			const uint32_t i = tile * kTileM + row;
			if (i < a.n) {
				const float* rec = a.records + (size_t)i * a.stride;
				const float px = __ldg(rec), py = __ldg(rec + 1), pz = __ldg(rec + 2);
#pragma unroll 2
				for (uint32_t l = 0; l < 8; ++l) {
					const uint32_t level = half_id * 8 + l;
					__half2 acc[4] = {__half2(), __half2(), __half2(), __half2()};
					if (level < a.L.n_levels) {
						const Corner c = pos_fract3(px, py, pz, a.L.scale[level]);
						if (a.L.hashed[level]) encode_one<true>(table + a.L.offset[level], c, a.L.size[level], 0, 0, acc);
						else encode_one<false>(table + a.L.offset[level], c, a.L.size[level], a.L.mul1[level], a.L.mul2[level], acc);
					}
					const uint4 v = make_uint4(*reinterpret_cast<uint32_t*>(&acc[0]), *reinterpret_cast<uint32_t*>(&acc[1]), *reinterpret_cast<uint32_t*>(&acc[2]),
					                           *reinterpret_cast<uint32_t*>(&acc[3]));
					*reinterpret_cast<uint4*>(A0 + level * kChunkA) = v;
				}
				float ex[8];
#pragma unroll
				for (uint32_t k = 0; k < 8; ++k) {
					const uint32_t d = half_id * 8 + k;
					ex[k] = d < a.n_extra ? __ldg(rec + a.stride - a.n_extra + d) : 0.0f;
				}
				const uint4 v = make_uint4(pack2(ex[0], ex[1]), pack2(ex[2], ex[3]), pack2(ex[4], ex[5]), pack2(ex[6], ex[7]));
				*reinterpret_cast<uint4*>(A0 + (16 + half_id) * kChunkA) = v;
			} else {
				const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
				for (uint32_t l = 0; l < 8; ++l) {
					*reinterpret_cast<uint4*>(A0 + (half_id * 8 + l) * kChunkA) = z;
				}
				*reinterpret_cast<uint4*>(A0 + (16 + half_id) * kChunkA) = z;
			}
			fence_proxy_async();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_full0 + 8 * b);
		}
	} else {
		// =================================== MLP warps: forward chain, then data-gradient chain ===================================
		// thread = (tile row, column part h): NC columns of every 64-wide layer from column cb on; the 16-wide layers belong to part 0
		constexpr int NC = kWidth / SPLIT;
		const uint32_t row = threadIdx.x & (kTileM - 1), h = threadIdx.x >> 7;
		const uint32_t cb = h * NC;
		const uint32_t taddr = tmem_base + (((warp & 3u) * 32u) << 16);
		unsigned char* H0 = smem + kSmH;
		unsigned char* H1 = smem + kSmH + kBytesH;
		unsigned char* A3 = smem + kSmA3;
		const uint32_t sH0 = sbase + kSmH, sH1 = sH0 + kBytesH, sA3 = sbase + kSmA3, sW = sbase + kSmW, sWt = sbase + kSmWt;
		uint32_t mma_phase = 0, it = 0;
		const bool issuer = threadIdx.x == 0;
#define MMA_WAIT()                        \
	mbar_wait(bar_mma, mma_phase);        \
	mma_phase ^= 1u;                      \
	tc_fence_after()
// the issuer's earlier bulk copies have read their tiles (each was issued a whole stage ago) before anybody may overwrite one
#define TILE_READY()                   \
	if (issuer) bulk_wait_read0();     \
	fence_proxy_async();               \
	tc_fence_before();                 \
	mlp_bar_sync<kMlpThreads>();       \
	tc_fence_after()
// issuer only: `chunks` finished 16-byte-column chunks of a shared-memory tile -> the same chunks of this tile in a scratch matrix
#define SEND(gbase, tile_chunks, chunk0, saddr, chunks)                                                                                             \
	bulk_store((gbase) + ((size_t)tile * (tile_chunks) + (chunk0)) * kChunkHalves, (saddr), (chunks) * kChunkA);                                    \
	bulk_commit()
		for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
			const uint32_t b = it & 1u;
			const uint32_t i = tile * kTileM + row;
			const bool valid = i < a.n;
			const uint32_t sA0 = sbase + kSmA0 + b * kBytesA0;
			float dx = 0.5f, dy = 0.5f, dz = 0.5f;
			float g_rgb[3] = {0.f, 0.f, 0.f}, g_sigma = 0.f;
			if (valid) {
				const float* rec = a.records + (size_t)i * a.stride;
				dx = __ldg(rec + 8); dy = __ldg(rec + 9); dz = __ldg(rec + 10);
				const uint2 raw = *reinterpret_cast<const uint2*>(a.dloss + (size_t)i * a.dloss_width);
				const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
				g_rgb[0] = lo.x; g_rgb[1] = lo.y; g_rgb[2] = hi.x; g_sigma = hi.y;
			}
			uint32_t m1[2] = {0, 0}, m2[2] = {0, 0}, m4[2] = {0, 0}, m5[2] = {0, 0}, m6[2] = {0, 0};
			// ---------------- forward ----------------
			mbar_wait(bar_full0 + 8 * b, (it >> 1) & 1u);
			tc_fence_after();
			if (issuer) {
				issue_gemm<kWidth, kInDensity>(tmem_base, sA0, sW + kOffW1 * 2, bar_mma);
				SEND(a.xs, kXChunks, 0, sA0, kInDensity / 8);  // X0: grid features + expression inputs, as the encoders wrote them
			}
			if (h == 0) {
				uint4 lo, hi;
				sh4(dx * 2.f - 1.f, dy * 2.f - 1.f, dz * 2.f - 1.f, lo, hi);
				unsigned char* dst = A3 + (row >> 3) * 128 + (row & 7) * 16;
				*reinterpret_cast<uint4*>(dst + 2 * kChunkA) = lo;
				*reinterpret_cast<uint4*>(dst + 3 * kChunkA) = hi;
			}
			MMA_WAIT();
			epi_fwd<NC, true>(taddr + cb, H0, cb / 8, row, m1, nullptr);
			TILE_READY();
			if (issuer) {
				if (it > 0) mbar_arrive(bar_empty0 + 8 * (b ^ 1u));
				issue_gemm<kWidth, kWidth>(tmem_base, sH0, sW + kOffW2 * 2, bar_mma);
				SEND(a.xs, kXChunks, 18, sH0, 8);
			}
			MMA_WAIT();
			epi_fwd<NC, true>(taddr + cb, H1, cb / 8, row, m2, nullptr);
			TILE_READY();
			if (issuer) {
				issue_gemm<kOutPad, kWidth>(tmem_base, sH1, sW + kOffW3 * 2, bar_mma);
				SEND(a.xs, kXChunks, 26, sH1, 8);
			}
			MMA_WAIT();
			float sigma = 0.f;
			if (h == 0) epi_fwd<kOutPad, false>(taddr, A3, 0, row, nullptr, &sigma);
			TILE_READY();
			if (issuer) {
				issue_gemm<kWidth, kInRgb>(tmem_base, sA3, sW + kOffV1 * 2, bar_mma);
				SEND(a.xs, kXChunks, 34, sA3, 4);  // density network outputs + SH of the direction
			}
			MMA_WAIT();
			epi_fwd<NC, true>(taddr + cb, H0, cb / 8, row, m4, nullptr);
			TILE_READY();
			if (issuer) {
				issue_gemm<kWidth, kWidth>(tmem_base, sH0, sW + kOffV2 * 2, bar_mma);
				SEND(a.xs, kXChunks, 38, sH0, 8);
			}
			MMA_WAIT();
			epi_fwd<NC, true>(taddr + cb, H1, cb / 8, row, m5, nullptr);
			TILE_READY();
			if (issuer) {
				issue_gemm<kWidth, kWidth>(tmem_base, sH1, sW + kOffV3 * 2, bar_mma);
				SEND(a.xs, kXChunks, 46, sH1, 8);
			}
			MMA_WAIT();
			epi_fwd<NC, true>(taddr + cb, H0, cb / 8, row, m6, nullptr);
			TILE_READY();
			if (issuer) { SEND(a.xs, kXChunks, 54, sH0, 8); }
			if (a.out) {  // forward output, when the caller wants it
				if (issuer) issue_gemm<kOutPad, kWidth>(tmem_base, sH0, sW + kOffV4 * 2, bar_mma);
				MMA_WAIT();
				if (h == 0) {
					float o[16];
					tmem_ld16(taddr, o);
					if (valid) {
						const __half2 rg = __floats2half2_rn(o[0], o[1]), bs = __floats2half2_rn(o[2], __half2float(__float2half_rn(sigma)));
						*reinterpret_cast<uint2*>(a.out + (size_t)i * 4) = make_uint2(*reinterpret_cast<const uint32_t*>(&rg), *reinterpret_cast<const uint32_t*>(&bs));
					}
				}
				tc_fence_before();
				mlp_bar_sync<kMlpThreads>();
				tc_fence_after();
			}
			// ---------------- backward: data gradients ----------------
			if (h == SPLIT - 1) {
				// dOut (colour output layer): d(r, g, b) in columns 0-2, the rest 0
				const uint4 lo = make_uint4(pack2(g_rgb[0], g_rgb[1]), pack2(g_rgb[2], 0.f), 0u, 0u), z = make_uint4(0, 0, 0, 0);
				unsigned char* dst = H1 + (row >> 3) * 128 + (row & 7) * 16;
				*reinterpret_cast<uint4*>(dst) = lo;
				*reinterpret_cast<uint4*>(dst + kChunkA) = z;
			}
			TILE_READY();
			// dG3 = dOut V4  (K = 16)
			if (issuer) {
				issue_gemm<kWidth, kOutPad>(tmem_base, sH1, sWt + kOffV4 * 2, bar_mma);
				SEND(a.dys, kDyChunks, 42, sH1, 2);
			}
			MMA_WAIT();
			epi_bwd<NC, true>(taddr + cb, H0, cb / 8, row, m6, 0.f);
			TILE_READY();
			if (issuer) {
				issue_gemm<kWidth, kWidth>(tmem_base, sH0, sWt + kOffV3 * 2, bar_mma);
				SEND(a.dys, kDyChunks, 34, sH0, 8);
			}
			MMA_WAIT();
			epi_bwd<NC, true>(taddr + cb, H1, cb / 8, row, m5, 0.f);
			TILE_READY();
			if (issuer) {
				issue_gemm<kWidth, kWidth>(tmem_base, sH1, sWt + kOffV2 * 2, bar_mma);
				SEND(a.dys, kDyChunks, 26, sH1, 8);
			}
			MMA_WAIT();
			epi_bwd<NC, true>(taddr + cb, H0, cb / 8, row, m4, 0.f);
			TILE_READY();
			// d(colour MLP input) = dG1 V1: columns 0-15 are the density network's outputs (+ d sigma on column 0), 16-31 the SH
			if (issuer) {
				issue_gemm<kInRgb, kWidth>(tmem_base, sH0, sWt + kOffV1 * 2, bar_mma);
				SEND(a.dys, kDyChunks, 18, sH0, 8);
			}
			MMA_WAIT();
			if (h == 0) epi_bwd<kOutPad, false>(taddr, H1, 0, row, nullptr, g_sigma);
			TILE_READY();
			// dH2 = dD W3 (K = 16)
			if (issuer) {
				issue_gemm<kWidth, kOutPad>(tmem_base, sH1, sWt + kOffW3 * 2, bar_mma);
				SEND(a.dys, kDyChunks, 16, sH1, 2);
			}
			MMA_WAIT();
			epi_bwd<NC, true>(taddr + cb, H0, cb / 8, row, m2, 0.f);
			TILE_READY();
			if (issuer) {
				issue_gemm<kWidth, kWidth>(tmem_base, sH0, sWt + kOffW2 * 2, bar_mma);
				SEND(a.dys, kDyChunks, 8, sH0, 8);
			}
			MMA_WAIT();
			epi_bwd<NC, true>(taddr + cb, H1, cb / 8, row, m1, 0.f);
			TILE_READY();
			// dX0 = dH1 W1: 144 columns; the first 128 are the grid features
			if (issuer) {
				issue_gemm<kInDensity, kWidth>(tmem_base, sH1, sWt + kOffW1 * 2, bar_mma);
				SEND(a.dys, kDyChunks, 0, sH1, 8);
			}
			MMA_WAIT();
			// staged in this tile's A0 buffer (its GEMM and its bulk copy are long done; the encoders get the buffer back only after this)
			{
				unsigned char* T = smem + kSmA0 + b * kBytesA0;
#pragma unroll
				for (int part = 0; part < 2 / SPLIT; ++part) epi_bwd<64, false>(taddr + 64 * (h + part), T, 8 * (h + part), row, nullptr, 0.f);
			}
			TILE_READY();
			// the encoders get this buffer back behind the next tile's first barrier, where the issuer has waited for this copy (they are a tile ahead anyway)
			if (issuer) { SEND(a.dfeat, kFeatChunks, 0, sA0, kFeatChunks); }
		}
		if (issuer) bulk_wait0();
#undef SEND
#undef MMA_WAIT
#undef TILE_READY
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

// ======================================================================================================================
// The same chain with the activations kept in TENSOR MEMORY between the GEMMs (default; net_fwd_bwd_kernel above = INSTA_NET_BWD=ss).
// With the operand tiles in shared memory every epilogue ends in fence.proxy.async (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC) before the barrier, and the MLP warps
// spent 40 % of their time there even with no global store outstanding (profiles/r2_net_fwd_bwd_bulk_ncu_summary.txt) -- the fence queues behind the encoder
// warps' gathers.  Here an epilogue packs its fp16 pairs and writes them with tcgen05.st into the columns the next GEMM reads as its A operand (".ts" form:
// row in lane, two K elements per 32-bit column), so a stage is tcgen05.ld -> ALU -> tcgen05.st -> wait::st -> barrier: no shared-memory tile, no proxy fence.
// The same registers go to the scratch matrices with 16-byte stores that are contiguous across a warp in the tile layout (row = lane) and that nobody waits
// for.  Only the first GEMM of a tile takes A from shared memory (the encoders' tile), which also leaves for the X scratch by one bulk copy.
constexpr uint32_t kTsSmW = 2 * kBytesA0, kTsSmWt = kTsSmW + kWBytes, kTsSmBar = kTsSmWt + kWBytes, kTsSmTotal = kTsSmBar + 64;
// tensor-memory columns of the A operands: hidden activations / their gradients (K = 64: 32 columns), colour-MLP input (K = 32: 16), 16-wide gradients (K = 16: 8)
constexpr uint32_t kColH = 160, kColA3 = 192, kColS = 208;

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

// N_COLS (16 or 32) accumulator columns of this thread's row -> MODE 1: ReLU, mask bits recorded / MODE 2: mask applied / MODE 0: as is -> fp16 pairs ->
// the A-operand columns at taddr_a + one 16-byte chunk per 8 columns of the scratch tile at g (this row's slot of the first chunk)
template <int N_COLS, int MODE>
__device__ __forceinline__ void epi_ts(uint32_t taddr_d, uint32_t taddr_a, uint32_t& mask, __half* g, float add0, float* keep0) {
	float v[N_COLS];
	tmem_ld_cols<N_COLS>(taddr_d, v);
	if (MODE == 1) {
		uint32_t bits = 0;
#pragma unroll
		for (int k = 0; k < N_COLS; ++k) {
			bits |= (v[k] > 0.0f ? 1u : 0u) << k;
			v[k] = fmaxf(v[k], 0.0f);
		}
		mask = bits;
	}
	if (MODE == 2) {
#pragma unroll
		for (int k = 0; k < N_COLS; ++k) v[k] = ((mask >> k) & 1u) ? v[k] : 0.0f;
	}
	v[0] += add0;
	if (keep0) *keep0 = v[0];
	uint32_t p[N_COLS / 2];
#pragma unroll
	for (int k = 0; k < N_COLS / 2; ++k) p[k] = pack2(v[2 * k], v[2 * k + 1]);
	if (N_COLS == 32) tmem_st16(taddr_a, p);
	else tmem_st8(taddr_a, p);
#pragma unroll
	for (int j = 0; j < N_COLS / 8; ++j) *reinterpret_cast<uint4*>(g + (size_t)j * kChunkHalves) = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
}

template <int SPLIT>
__global__ void __launch_bounds__((4 * SPLIT + kEncWarps) * 32, 1) net_fwd_bwd_ts_kernel(const __grid_constant__ BwdArgs a) {
	constexpr int kMlpWarps = 4 * SPLIT, kMlpThreads = kMlpWarps * 32;
	extern __shared__ __align__(1024) unsigned char smem[];
	const uint32_t sbase = smem_u32(smem);
	const uint32_t bar_full0 = sbase + kTsSmBar, bar_empty0 = sbase + kTsSmBar + 16, bar_mma = sbase + kTsSmBar + 32;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kTsSmBar + 40);
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint32_t n_tiles = (a.n + kTileM - 1) / kTileM;

	for (int l = 0; l < 7; ++l) {
		stage_w(smem + kTsSmW + c_woff[l] * 2, a.params + c_woff[l], (int)c_nout[l], (int)c_kin[l]);
		stage_wt(smem + kTsSmWt + c_woff[l] * 2, a.params + c_woff[l], (int)c_nout[l], (int)c_kin[l]);
	}
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
	fence_proxy_async();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp >= kMlpWarps) {
		// =================================== encoders: the tile of grid features + expression inputs, as in mlp.cu ===================================
		const uint32_t e = threadIdx.x - kMlpThreads;
		const uint32_t row = e & (kTileM - 1), half_id = e >> 7;
		const uint4* table = reinterpret_cast<const uint4*>(a.params + kOffGrid);
		uint32_t it = 0;
		for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
			const uint32_t b = it & 1u;
			mbar_wait(bar_empty0 + 8 * b, ((it >> 1) & 1u) ^ 1u);
			unsigned char* A0 = smem + b * kBytesA0 + (row >> 3) * 128 + (row & 7) * 16;
			const uint32_t i = tile * kTileM + row;
			if (i < a.n) {
				const float* rec = a.records + (size_t)i * a.stride;
				const float px = __ldg(rec), py = __ldg(rec + 1), pz = __ldg(rec + 2);
#pragma unroll 2
				for (uint32_t l = 0; l < 8; ++l) {
					const uint32_t level = half_id * 8 + l;
					__half2 acc[4] = {__half2(), __half2(), __half2(), __half2()};
					if (level < a.L.n_levels) {
						const Corner c = pos_fract3(px, py, pz, a.L.scale[level]);
						if (a.L.hashed[level]) encode_one<true>(table + a.L.offset[level], c, a.L.size[level], 0, 0, acc);
						else encode_one<false>(table + a.L.offset[level], c, a.L.size[level], a.L.mul1[level], a.L.mul2[level], acc);
					}
					*reinterpret_cast<uint4*>(A0 + level * kChunkA) = make_uint4(*reinterpret_cast<uint32_t*>(&acc[0]), *reinterpret_cast<uint32_t*>(&acc[1]),
					                                                            *reinterpret_cast<uint32_t*>(&acc[2]), *reinterpret_cast<uint32_t*>(&acc[3]));
				}
				float ex[8];
#pragma unroll
				for (uint32_t k = 0; k < 8; ++k) {
					const uint32_t d = half_id * 8 + k;
					ex[k] = d < a.n_extra ? __ldg(rec + a.stride - a.n_extra + d) : 0.0f;
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
		// thread = (tile row, column part h): NC columns of every 64-wide layer from column cb on; the 16-wide layers belong to part 0
		constexpr int NC = kWidth / SPLIT, NP = NC / 32;  // columns per thread, in NP pieces of 32 (one mask word each)
		const uint32_t row = threadIdx.x & (kTileM - 1), h = threadIdx.x >> 7;
		const uint32_t cb = h * NC;
		const uint32_t tlane = tmem_base + (((warp & 3u) * 32u) << 16);  // this warp's 32 lanes (tile rows), column 0
		const uint32_t sW = sbase + kTsSmW, sWt = sbase + kTsSmWt;
		const uint32_t tH = tmem_base + kColH, tA3 = tmem_base + kColA3, tS = tmem_base + kColS;
		uint32_t mma_phase = 0, it = 0;
		const bool issuer = threadIdx.x == 0;
#define MMA_WAIT()                        \
	mbar_wait(bar_mma, mma_phase);        \
	mma_phase ^= 1u;                      \
	tc_fence_after()
// what the epilogue stored into tensor memory is complete before anybody -- the issuer's next GEMM -- may read it
#define STAGE_DONE()             \
	tmem_st_wait();              \
	tc_fence_before();           \
	mlp_bar_sync<kMlpThreads>(); \
	tc_fence_after()
// a 64-wide layer: this thread's NC columns -> A-operand columns kColH.. + scratch chunks from `chunk0` of matrix g
#define HIDDEN(MODE, m, g, chunk0)                                                                                                                 \
	_Pragma("unroll") for (int part = 0; part < NP; ++part)                                                                                         \
	    epi_ts<32, MODE>(tlane + cb + 32 * part, tlane + kColH + (cb + 32 * part) / 2, m[part], (g) + (size_t)((chunk0) + (cb + 32 * part) / 8) * kChunkHalves, 0.f, nullptr)
		for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
			const uint32_t b = it & 1u;
			const uint32_t i = tile * kTileM + row;
			const bool valid = i < a.n;
			const uint32_t sA0 = sbase + b * kBytesA0;
			__half* xs_t = a.xs + (size_t)tile * kXChunks * kChunkHalves + row * 8;
			__half* dys_t = a.dys + (size_t)tile * kDyChunks * kChunkHalves + row * 8;
			float dx = 0.5f, dy = 0.5f, dz = 0.5f;
			float g_rgb[3] = {0.f, 0.f, 0.f}, g_sigma = 0.f;
			if (valid) {
				const float* rec = a.records + (size_t)i * a.stride;
				dx = __ldg(rec + 8); dy = __ldg(rec + 9); dz = __ldg(rec + 10);
				const uint2 raw = *reinterpret_cast<const uint2*>(a.dloss + (size_t)i * a.dloss_width);
				const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
				g_rgb[0] = lo.x; g_rgb[1] = lo.y; g_rgb[2] = hi.x; g_sigma = hi.y;
			}
			uint32_t m1[2] = {0, 0}, m2[2] = {0, 0}, m4[2] = {0, 0}, m5[2] = {0, 0}, m6[2] = {0, 0}, no_mask = 0;
			// ---------------- forward ----------------
			mbar_wait(bar_full0 + 8 * b, (it >> 1) & 1u);
			tc_fence_after();
			if (issuer) {
				issue_gemm<kWidth, kInDensity>(tmem_base, sA0, sW + kOffW1 * 2, bar_mma);
				bulk_store(a.xs + (size_t)tile * kXChunks * kChunkHalves, sA0, kBytesA0);  // X0: grid features + expression inputs, as the encoders wrote them
				bulk_commit();
			}
			if (h == 0) {
				// SH of the direction: the second half of the colour MLP's input
				uint4 lo, hi;
				sh4(dx * 2.f - 1.f, dy * 2.f - 1.f, dz * 2.f - 1.f, lo, hi);
				const uint32_t p[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
				tmem_st8(tlane + kColA3 + 8, p);
				*reinterpret_cast<uint4*>(xs_t + (size_t)36 * kChunkHalves) = lo;
				*reinterpret_cast<uint4*>(xs_t + (size_t)37 * kChunkHalves) = hi;
			}
			MMA_WAIT();
			HIDDEN(1, m1, xs_t, 18);
			STAGE_DONE();
			if (issuer) {
				// the first GEMM and the bulk copy have read the encoders' tile
				bulk_wait_read0();
				mbar_arrive(bar_empty0 + 8 * b);
				issue_gemm_ts<kWidth, kWidth>(tmem_base, tH, sW + kOffW2 * 2, bar_mma);
			}
			MMA_WAIT();
			HIDDEN(1, m2, xs_t, 26);
			STAGE_DONE();
			if (issuer) issue_gemm_ts<kOutPad, kWidth>(tmem_base, tH, sW + kOffW3 * 2, bar_mma);
			MMA_WAIT();
			float sigma = 0.f;
			if (h == 0) epi_ts<kOutPad, 0>(tlane, tlane + kColA3, no_mask, xs_t + (size_t)34 * kChunkHalves, 0.f, &sigma);
			STAGE_DONE();
			if (issuer) issue_gemm_ts<kWidth, kInRgb>(tmem_base, tA3, sW + kOffV1 * 2, bar_mma);
			MMA_WAIT();
			HIDDEN(1, m4, xs_t, 38);
			STAGE_DONE();
			if (issuer) issue_gemm_ts<kWidth, kWidth>(tmem_base, tH, sW + kOffV2 * 2, bar_mma);
			MMA_WAIT();
			HIDDEN(1, m5, xs_t, 46);
			STAGE_DONE();
			if (issuer) issue_gemm_ts<kWidth, kWidth>(tmem_base, tH, sW + kOffV3 * 2, bar_mma);
			MMA_WAIT();
			HIDDEN(1, m6, xs_t, 54);
			STAGE_DONE();
			if (a.out) {  // forward output, when the caller wants it
				if (issuer) issue_gemm_ts<kOutPad, kWidth>(tmem_base, tH, sW + kOffV4 * 2, bar_mma);
				MMA_WAIT();
				if (h == 0) {
					float o[16];
					tmem_ld16(tlane, o);
					if (valid) {
						const __half2 rg = __floats2half2_rn(o[0], o[1]), bs = __floats2half2_rn(o[2], __half2float(__float2half_rn(sigma)));
						*reinterpret_cast<uint2*>(a.out + (size_t)i * 4) = make_uint2(*reinterpret_cast<const uint32_t*>(&rg), *reinterpret_cast<const uint32_t*>(&bs));
					}
				}
				tc_fence_before();
				mlp_bar_sync<kMlpThreads>();
				tc_fence_after();
			}
			// ---------------- backward: data gradients ----------------
			if (h == SPLIT - 1) {
				// dOut (colour output layer): d(r, g, b) in columns 0-2, the rest 0
				const uint32_t p[8] = {pack2(g_rgb[0], g_rgb[1]), pack2(g_rgb[2], 0.f), 0u, 0u, 0u, 0u, 0u, 0u};
				tmem_st8(tlane + kColS, p);
				*reinterpret_cast<uint4*>(dys_t + (size_t)42 * kChunkHalves) = make_uint4(p[0], p[1], 0u, 0u);
				*reinterpret_cast<uint4*>(dys_t + (size_t)43 * kChunkHalves) = make_uint4(0u, 0u, 0u, 0u);
			}
			STAGE_DONE();
			// dG3 = dOut V4  (K = 16)
			if (issuer) issue_gemm_ts<kWidth, kOutPad>(tmem_base, tS, sWt + kOffV4 * 2, bar_mma);
			MMA_WAIT();
			HIDDEN(2, m6, dys_t, 34);
			STAGE_DONE();
			if (issuer) issue_gemm_ts<kWidth, kWidth>(tmem_base, tH, sWt + kOffV3 * 2, bar_mma);
			MMA_WAIT();
			HIDDEN(2, m5, dys_t, 26);
			STAGE_DONE();
			if (issuer) issue_gemm_ts<kWidth, kWidth>(tmem_base, tH, sWt + kOffV2 * 2, bar_mma);
			MMA_WAIT();
			HIDDEN(2, m4, dys_t, 18);
			STAGE_DONE();
			// d(colour MLP input) = dG1 V1: columns 0-15 are the density network's outputs (+ d sigma on column 0), 16-31 the SH
			if (issuer) issue_gemm_ts<kInRgb, kWidth>(tmem_base, tH, sWt + kOffV1 * 2, bar_mma);
			MMA_WAIT();
			if (h == 0) epi_ts<kOutPad, 0>(tlane, tlane + kColS, no_mask, dys_t + (size_t)16 * kChunkHalves, g_sigma, nullptr);
			STAGE_DONE();
			// dH2 = dD W3 (K = 16)
			if (issuer) issue_gemm_ts<kWidth, kOutPad>(tmem_base, tS, sWt + kOffW3 * 2, bar_mma);
			MMA_WAIT();
			HIDDEN(2, m2, dys_t, 8);
			STAGE_DONE();
			if (issuer) issue_gemm_ts<kWidth, kWidth>(tmem_base, tH, sWt + kOffW2 * 2, bar_mma);
			MMA_WAIT();
			HIDDEN(2, m1, dys_t, 0);
			STAGE_DONE();
			// dX0 = dH1 W1: 144 columns; the first 128 are the grid features -> [tile][level][row][8 features], what hash_bwd_pairs_kernel<true> reads
			if (issuer) issue_gemm_ts<kInDensity, kWidth>(tmem_base, tH, sWt + kOffW1 * 2, bar_mma);
			MMA_WAIT();
			{
				__half* df_t = a.dfeat + (size_t)tile * kFeatChunks * kChunkHalves + row * 8;
#pragma unroll 1
				for (uint32_t c0 = h * (128 / SPLIT); c0 < (h + 1) * (128 / SPLIT); c0 += 32) {
					float v[32];
					tmem_ld32(tlane + c0, v);
#pragma unroll
					for (int j = 0; j < 4; ++j)
						*reinterpret_cast<uint4*>(df_t + (size_t)(c0 / 8 + j) * kChunkHalves) =
						    make_uint4(pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]), pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
				}
			}
			tc_fence_before();
			mlp_bar_sync<kMlpThreads>();
			tc_fence_after();
		}
		if (issuer) bulk_wait0();
#undef MMA_WAIT
#undef STAGE_DONE
#undef HIDDEN
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

// ======================================================================================================================
// Weight gradients.
constexpr int kWgThreads = 256;
constexpr uint32_t kWgK = 32;                          // samples per stage
constexpr uint32_t kWgSbo = (kWgK / 8) * 128 + 16;     // 528: stride between 8-feature groups (+16 B: bank spread of the copy)
constexpr uint32_t kWgXGroups = kXCols / 8, kWgDyGroups = kDyCols / 8;
constexpr uint32_t kWgXBytes = kWgXGroups * kWgSbo, kWgDyBytes = (kWgDyGroups + 16) * kWgSbo;  // + 16 groups the M = 128 descriptor may touch
constexpr uint32_t kWgStage = kWgXBytes + kWgDyBytes;
constexpr uint32_t kWgStages = 3;                      // blocks of kWgK samples in shared memory: two of them being copied while the MMAs read the third
constexpr uint32_t kWgBar = kWgStages * kWgStage, kWgSmTotal = kWgBar + 64;
constexpr uint32_t kWgTmemCols = 512;
__constant__ const uint32_t c_wgcol[7] = {0, 144, 208, 272, 304, 368, 432};  // accumulator column of layer l (= c_xoff)

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory"); }

struct WgArgs {
	const __half* xs;
	const __half* dys;
	uint32_t n_pad;
	float* dw32;  // 25,600 floats, accumulated with atomics
};

__global__ void __launch_bounds__(kWgThreads, 1) net_wgrad_kernel(const WgArgs a) {
	extern __shared__ __align__(1024) unsigned char smem[];
	const uint32_t sbase = smem_u32(smem);
	const uint32_t bar_free0 = sbase + kWgBar, bar_done = bar_free0 + 8 * kWgStages;  // one mbarrier per stage, then the final one
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kWgBar + 8 * kWgStages + 8);
	const uint32_t warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) {
		for (uint32_t st = 0; st < kWgStages; ++st) mbar_init(bar_free0 + 8 * st, 1);
		mbar_init(bar_done, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	// the padding groups behind the dY tiles are read by the M = 128 descriptors: keep them finite
	for (uint32_t t = threadIdx.x; t < kWgStages * kWgStage / 16; t += blockDim.x) reinterpret_cast<uint4*>(smem)[t] = make_uint4(0, 0, 0, 0);
	if (warp == 0) {
		__syncwarp();
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kWgTmemCols) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	fence_proxy_async();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	const uint32_t n_blocks = a.n_pad / kWgK;
	// copies of block `blk` (kWgK samples of X and dY) into stage st; chunk (sample k, feature group g): 16 bytes of the row-major scratch -> MN-major unit.
	// Always closes a cp.async group, also past the last block, so that "all but the newest group" means the same thing in every iteration
	auto request = [&](uint32_t blk, uint32_t st) {
		if (blk < n_blocks) {
			const uint32_t sx = sbase + st * kWgStage, sdy = sx + kWgXBytes;
			// scratch layout [tile][chunk g][row][8 halves]: the kWgK rows of a chunk are 512 contiguous bytes on both sides
			const uint32_t tile = blk / (kTileM / kWgK), r0 = (blk % (kTileM / kWgK)) * kWgK;
			const __half* xsrc = a.xs + (size_t)tile * kXChunks * kChunkHalves + r0 * 8;
			const __half* dysrc = a.dys + (size_t)tile * kDyChunks * kChunkHalves + r0 * 8;
			for (uint32_t t = threadIdx.x; t < kWgK * kWgXGroups; t += blockDim.x) {
				const uint32_t g = t / kWgK, k = t % kWgK;
				cp_async16(sx + g * kWgSbo + k * 16, xsrc + (size_t)g * kChunkHalves + k * 8);
			}
			for (uint32_t t = threadIdx.x; t < kWgK * kWgDyGroups; t += blockDim.x) {
				const uint32_t g = t / kWgK, k = t % kWgK;
				cp_async16(sdy + g * kWgSbo + k * 16, dysrc + (size_t)g * kChunkHalves + k * 8);
			}
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
	request(blockIdx.x, 0);
	request(blockIdx.x + gridDim.x, 1);
	uint32_t it = 0;
	for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
		const uint32_t st = it % kWgStages;
		asm volatile("cp.async.wait_group 1;" ::: "memory");  // this block's copies have landed (the next block's may still be in flight)
		fence_proxy_async();
		tc_fence_before();
		__syncthreads();
		tc_fence_after();
		if (threadIdx.x == 0) {
			const uint32_t sx = sbase + st * kWgStage, sdy = sx + kWgXBytes;
#pragma unroll
			for (int l = 0; l < 7; ++l) {
				const uint32_t idesc = umma_idesc(128, c_kin[l], 1, 1);
#pragma unroll
				for (uint32_t j = 0; j < kWgK / 16; ++j) {
					const uint64_t ad = umma_desc(sdy + (c_dyoff[l] / 8) * kWgSbo + j * 256, 128, kWgSbo);
					const uint64_t bd = umma_desc(sx + (c_xoff[l] / 8) * kWgSbo + j * 256, 128, kWgSbo);
					umma_f16(tmem_base + c_wgcol[l], ad, bd, idesc, (it > 0 || j > 0) ? 1u : 0u);
				}
			}
			umma_commit(bar_free0 + 8 * st);
		}
		// the block after next goes where the PREVIOUS block was: its MMAs were issued an iteration ago
		const uint32_t nst = (it + 2) % kWgStages;
		if (it > 0) mbar_wait(bar_free0 + 8 * nst, ((it - 1) / kWgStages) & 1u);
		request(blk + 2 * gridDim.x, nst);
	}
	// all MMAs done -> epilogue
	if (threadIdx.x == 0) umma_commit(bar_done);
	mbar_wait(bar_done, 0);
	tc_fence_after();
	if (it > 0 && warp < 2) {
		const uint32_t o = threadIdx.x;  // output row = TMEM lane, 0..63
		const uint32_t taddr = tmem_base + ((warp * 32u) << 16);
#pragma unroll 1
		for (int l = 0; l < 7; ++l) {
			const uint32_t kin = c_kin[l], nout = c_nout[l];
			for (uint32_t c0 = 0; c0 < kin; c0 += 16) {
				float v[16];
				tmem_ld16(taddr + c_wgcol[l] + c0, v);
				if (o < nout) {
					float* dst = a.dw32 + c_woff[l] + (size_t)o * kin + c0;
#pragma unroll
					for (int k = 0; k < 16; ++k) atomicAdd(dst + k, v[k]);
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kWgTmemCols) : "memory");
}

__global__ void wgrad_finish_kernel(uint32_t n, float* __restrict__ dw32, __half* __restrict__ grads) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	grads[i] = __float2half_rn(dw32[i]);
	dw32[i] = 0.f;
}

}  // namespace

struct BwdScratch {
	uint32_t n_pad = 0;
	__half* xs = nullptr;
	__half* dys = nullptr;
	__half* dfeat = nullptr;
	float* dw32 = nullptr;
};

void network_bwd_destroy(insta_ctx* ctx) {
	BwdScratch* s = (BwdScratch*)ctx->bwd_scratch;
	if (!s) return;
	cudaFree(s->xs); cudaFree(s->dys); cudaFree(s->dfeat); cudaFree(s->dw32);
	delete s;
	ctx->bwd_scratch = nullptr;
}

// NerfNetwork::forward + backward on the compacted batch: fills the fp16 gradient buffer (MLP weights, then grid) in parameter order.
// `ex` (training with more than one rank): the grid gradients -- 99.9 % of the gradient vector, all produced by the scatter kernel -- are computed FIRST and
// their NCCL sum runs on ex->side while this stream computes the MLP weight gradients; the caller's optimizer step waits for ex->ev_grid_summed.
int launch_network_fwd_bwd(insta_ctx* ctx, const float* d_records, uint32_t stride, uint32_t n, const void* d_params, const void* d_dloss, uint32_t dloss_width,
                           void* d_grads, void* d_out, cudaStream_t s, const GradExchange* ex) {
	if (n == 0) return INSTA_OK;
	if (!ctx->bwd_scratch) ctx->bwd_scratch = new BwdScratch();
	BwdScratch* sc = (BwdScratch*)ctx->bwd_scratch;
	const uint32_t n_pad = (n + kTileM - 1) / kTileM * kTileM;
	if (n_pad > sc->n_pad) {
		INSTA_CUDA(cudaDeviceSynchronize());
		cudaFree(sc->xs); cudaFree(sc->dys); cudaFree(sc->dfeat);
		sc->xs = sc->dys = sc->dfeat = nullptr;
		sc->n_pad = 0;
		INSTA_CUDA(cudaMalloc(&sc->xs, (size_t)n_pad * kXCols * 2));
		INSTA_CUDA(cudaMalloc(&sc->dys, (size_t)n_pad * kDyCols * 2));
		INSTA_CUDA(cudaMalloc(&sc->dfeat, (size_t)n_pad * 128 * 2));
		sc->n_pad = n_pad;
	}
	if (!sc->dw32) {
		INSTA_CUDA(cudaMalloc(&sc->dw32, kOffGrid * sizeof(float)));
		INSTA_CUDA(cudaMemsetAsync(sc->dw32, 0, kOffGrid * sizeof(float), s));
	}
	BwdArgs a{};
	a.records = d_records;
	a.stride = stride;
	a.n = n;
	a.n_extra = ctx->p.exp_params + ctx->p.geo_params;
	a.n_pad = n_pad;
	a.params = (const __half*)d_params;
	a.dloss = (const __half*)d_dloss;
	a.dloss_width = dloss_width;
	a.out = (__half*)d_out;
	a.xs = sc->xs;
	a.dys = sc->dys;
	a.dfeat = sc->dfeat;
	a.L = ctx->levels;
	// rows [n, n_pad) of the scratch matrices are read by the weight-gradient kernel: the kernel writes every row of every tile, and a row without a sample
	// (zero inputs, zero loss gradient) is all zeros in X and dY
	const uint32_t n_tiles = n_pad / kTileM;
	static const bool split1 = [] { const char* e = getenv("INSTA_NET_BWD_SPLIT"); return e && atoi(e) == 1; }();  // A/B: four MLP warps, a whole row per thread (round 1)
	static const bool ss = [] { const char* e = getenv("INSTA_NET_BWD"); return e && e[0] == 's'; }();                   // A/B: activations through shared-memory operand tiles
	const uint32_t grid = std::min<uint32_t>(n_tiles, (uint32_t)ctx->sm_count);
	if (ss) {
		if (split1) {
			INSTA_CUDA(cudaFuncSetAttribute(net_fwd_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmTotal));
			net_fwd_bwd_kernel<1><<<grid, (4 + kEncWarps) * 32, kSmTotal, s>>>(a);
		} else {
			INSTA_CUDA(cudaFuncSetAttribute(net_fwd_bwd_kernel<kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmTotal));
			net_fwd_bwd_kernel<kSplit><<<grid, (4 * kSplit + kEncWarps) * 32, kSmTotal, s>>>(a);
		}
	} else if (split1) {
		INSTA_CUDA(cudaFuncSetAttribute(net_fwd_bwd_ts_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTsSmTotal));
		net_fwd_bwd_ts_kernel<1><<<grid, (4 + kEncWarps) * 32, kTsSmTotal, s>>>(a);
	} else {
		INSTA_CUDA(cudaFuncSetAttribute(net_fwd_bwd_ts_kernel<kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTsSmTotal));
		net_fwd_bwd_ts_kernel<kSplit><<<grid, (4 * kSplit + kEncWarps) * 32, kTsSmTotal, s>>>(a);
	}
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	const uint64_t n_grid = (uint64_t)ctx->levels.offset[ctx->levels.n_levels] * 8u;
	if (ex) {
		// Grid gradients (kernel_grid_backward on the SoA dL/d(features)) in groups of levels of about equal size, each group's sum across ranks on the side
		// stream while the next group is scattered: the scatter kernel's CTAs are small, so NCCL's kernel gets SMs next to it (it does not next to the
		// persistent weight-gradient kernel), and only the last group's sum is left when the scatter ends.
		const GridLevels& L = ctx->levels;
		const uint32_t n_groups = std::min<uint32_t>(4u, L.n_levels);
		uint32_t lb = 0;
		for (uint32_t g = 0; g < n_groups; ++g) {
			// group g ends at the level where the running entry count passes (g + 1) / n_groups of the table
			uint32_t le = lb + 1;
			while (le < L.n_levels && (g + 1 == n_groups || (uint64_t)L.offset[le] * n_groups < (uint64_t)L.offset[L.n_levels] * (g + 1))) ++le;
			if (g + 1 == n_groups) le = L.n_levels;
			int rc = launch_hash_bwd_levels(ctx, d_records, stride, n, sc->dfeat, 0 /* tiled */, (__half*)d_grads + kOffGrid, 1, lb, le, s);
			if (rc) return rc;
			INSTA_CUDA(cudaEventRecord(ex->ev_grid_ready[g], s));
			INSTA_CUDA(cudaStreamWaitEvent(ex->side, ex->ev_grid_ready[g], 0));
			rc = grad_allreduce(ctx, (__half*)d_grads + kOffGrid + (size_t)L.offset[lb] * 8u, (uint64_t)(L.offset[le] - L.offset[lb]) * 8u, ex->side);
			if (rc) return rc;
			lb = le;
			if (lb >= L.n_levels) break;
		}
		(void)n_grid;
		INSTA_CUDA(cudaEventRecord(ex->ev_grid_summed, ex->side));
	}
	WgArgs w{sc->xs, sc->dys, n_pad, sc->dw32};
	INSTA_CUDA(cudaFuncSetAttribute(net_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmTotal));
	// With the collective in flight on the side stream, leave a few SMs to NCCL's kernel: the weight-gradient kernel is persistent (one CTA per SM with
	// most of its shared memory), and a collective that cannot get an SM simply runs after it
	static const uint32_t reserve = [] { const char* e = getenv("INSTA_WGRAD_SM_RESERVE"); return e ? (uint32_t)atoi(e) : 16u; }();
	const uint32_t wg_sms = ex ? std::max<uint32_t>((uint32_t)ctx->sm_count > reserve ? (uint32_t)ctx->sm_count - reserve : 1u, 1u) : (uint32_t)ctx->sm_count;
	net_wgrad_kernel<<<std::min<uint32_t>(n_pad / kWgK, wg_sms), kWgThreads, kWgSmTotal, s>>>(w);
	ctx->launches++;
	wgrad_finish_kernel<<<(kOffGrid + 255) / 256, 256, 0, s>>>(kOffGrid, sc->dw32, (__half*)d_grads);
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	if (ex) {
		// the 25,600 MLP gradients: NCCL runs the two sums in the order they were issued (the same on every rank)
		int rc = grad_allreduce(ctx, d_grads, kOffGrid, s);
		if (rc) return rc;
		INSTA_CUDA(cudaStreamWaitEvent(s, ex->ev_grid_summed, 0));
		return INSTA_OK;
	}
	// grid gradients: kernel_grid_backward on the SoA dL/d(features)
	return launch_hash_bwd_pitched(ctx, d_records, stride, n, sc->dfeat, 0 /* tiled */, (__half*)d_grads + kOffGrid, 1, s);
}

}  // namespace insta
