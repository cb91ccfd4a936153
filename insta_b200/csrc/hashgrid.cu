// Stage 2 of the INSTA hot path on sm_100a: multiresolution hash-grid encoding, forward and scatter-add backward.
// Replaces tcnn::GridEncodingTemplated<__half,3,8>::forward_impl / backward_impl and the kernels kernel_grid /
// kernel_grid_backward (tcnn include/tiny-cuda-nn/encodings/grid.h:138-310, 312-435, 874-1110), with fast_hash /
// grid_index (grid.h:82-117) and pos_fract (tcnn common_device.h:415-431).
//
// Layout: the fp16 parameter table is borrowed from the caller in the reference's own order (entries x 8 features,
// levels concatenated at the ctor's offsets), so a snapshot's params_binary can be used as is.  One entry = 16 bytes =
// one 128-bit load per corner.  Forward: a CTA owns a tile of 64 samples and all levels (warp = level, lane = sample pair),
// so the strided sample records are read once; the 49 MB table stays L2-resident (evict_last gathers, streaming loads /
// stores for everything that is touched once); outputs are written SoA, one full 128-byte line per (tile, feature row).
//
// Exactness: slot indices use the same uint32 arithmetic as grid_index, including its loop-exit rule and wrap-around;
// the per-level scale is evaluated ON DEVICE with the reference kernel's expression (exp2f(l * log2(s)) * base - 1,
// contracted by nvcc to one FFMA exactly as in the reference build) so floor() sees identical values; the
// interpolation reproduces `result[f] += (half)(w * (float)table[f])` in corner order with fp16 accumulation.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>

#include "ctx.h"
#include "grid_device.cuh"

namespace insta {
namespace {

constexpr int kFwdWarps = 16;
constexpr int kFwdThreads = kFwdWarps * 32;
constexpr int kBwdThreads = kFwdThreads;
constexpr uint32_t kTile = 64;  // samples per CTA: 32 lanes x 2

__global__ void level_scale_kernel(uint32_t n_levels, float log2_per_level_scale, uint32_t base_resolution, float* out) {
	const uint32_t level = threadIdx.x;
	// same expression (and the same default contraction) as tcnn grid.h:188
	if (level < n_levels) out[level] = exp2f(level * log2_per_level_scale) * base_resolution - 1.0f;
}

using namespace griddev;

// One CTA = one tile of kTile samples x ALL levels: warp w handles level w (levels w, w + kFwdWarps, ... when there are
// more levels than warps), lane l the sample pair (2l, 2l+1) of the tile.  The tile's positions are read from the strided
// records ONCE (staged in shared memory) -- with one level per CTA the 108-byte records streamed through L2 once per
// level (16 x 453 MB; measured 7.3 GB of DRAM reads per 4M-sample launch, profiles/r1_hash_fwd_ncu_summary.txt).
template <bool HASHED>
__device__ __forceinline__ void fwd_level(const GridLevels& L, uint32_t level, uint32_t n, uint32_t i0, bool one, bool two, const float* __restrict__ sp,
                                          const uint4* __restrict__ table, __half* __restrict__ out) {
	const uint32_t size = L.size[level], mul1 = L.mul1[level], mul2 = L.mul2[level];
	const float scale = L.scale[level];
	table += L.offset[level];
	const Corner c0 = pos_fract3(sp[0], sp[1], sp[2], scale);
	const Corner c1 = pos_fract3(sp[3], sp[4], sp[5], scale);
	__half2 a0[4], a1[4];
	encode_one<HASHED>(table, c0, size, mul1, mul2, a0);
	encode_one<HASHED>(table, c1, size, mul1, mul2, a1);
	if (!one) return;
	__half* o = out + (size_t)(level * 8) * n + i0;
	if (two && (n % 2u == 0)) {
		// rows start at even element offsets -> 4-byte aligned pair stores
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const __half2 lo = __lows2half2(a0[k], a1[k]), hi = __highs2half2(a0[k], a1[k]);
			st_stream_u32(o + (size_t)(2 * k) * n, *reinterpret_cast<const uint32_t*>(&lo));
			st_stream_u32(o + (size_t)(2 * k + 1) * n, *reinterpret_cast<const uint32_t*>(&hi));
		}
	} else {
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			st_stream_u16(o + (size_t)(2 * k) * n, __half_as_ushort(__low2half(a0[k])));
			st_stream_u16(o + (size_t)(2 * k + 1) * n, __half_as_ushort(__high2half(a0[k])));
			if (two) {
				st_stream_u16(o + (size_t)(2 * k) * n + 1, __half_as_ushort(__low2half(a1[k])));
				st_stream_u16(o + (size_t)(2 * k + 1) * n + 1, __half_as_ushort(__high2half(a1[k])));
			}
		}
	}
}

// stage the xyz of the tile's samples: s_pos[3 * j + d], zero for samples past n
__device__ __forceinline__ void stage_positions(float* s_pos, const float* __restrict__ pos, uint32_t stride, uint32_t n, uint32_t tile_first) {
	for (uint32_t t = threadIdx.x; t < kTile * 3u; t += blockDim.x) {
		const uint32_t j = t / 3u, d = t - 3u * j, i = tile_first + j;
		s_pos[t] = i < n ? ld_stream_f32(pos + (size_t)i * stride + d) : 0.0f;
	}
	__syncthreads();
}

__global__ void __launch_bounds__(kFwdThreads) hash_fwd_kernel(const __grid_constant__ GridLevels L, uint32_t n, const float* __restrict__ pos,
                                                               uint32_t stride, const uint4* __restrict__ table, __half* __restrict__ out) {
	__shared__ float s_pos[kTile * 3];
	const uint32_t tile_first = blockIdx.x * kTile;
	stage_positions(s_pos, pos, stride, n, tile_first);
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	const uint32_t i0 = tile_first + 2u * lane;
	const bool one = i0 < n, two = i0 + 1 < n;
	for (uint32_t level = warp; level < L.n_levels; level += kFwdWarps) {
		if (L.hashed[level]) fwd_level<true>(L, level, n, i0, one, two, s_pos + 6 * lane, table, out);
		else fwd_level<false>(L, level, n, i0, one, two, s_pos + 6 * lane, table, out);
	}
}

// Lane-pair version (default): lane l = x-side (l & 1) of the samples 4 (l >> 1) + {0, 1, 2, 3} of the tile; see pair_load / pair_accumulate.  A lane ends up
// with 4 features x 4 consecutive samples: one 8-byte store per feature row, 16 lanes of a warp complete a 128-byte line per row.
template <bool HASHED>
__device__ __forceinline__ void fwd_level_pairs(const GridLevels& L, uint32_t level, uint32_t n, uint32_t tile_first, uint32_t lane, const float* __restrict__ s_pos,
                                                const uint4* __restrict__ table, __half* __restrict__ out) {
	const uint32_t b = lane & 1u, g = lane >> 1;
	const uint32_t size = L.size[level], mul1 = L.mul1[level], mul2 = L.mul2[level];
	const float scale = L.scale[level];
	table += L.offset[level];
	__half2 acc[4][2];
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		uint4 v[2][4];
		float w[2][4];
#pragma unroll
		for (int q = 0; q < 2; ++q) {
			const float* sp = s_pos + 3u * (4u * g + 2u * h + q);
			pair_load<HASHED>(table, pos_fract3(sp[0], sp[1], sp[2], scale), b, size, mul1, mul2, v[q], w[q]);
		}
#pragma unroll
		for (int q = 0; q < 2; ++q) pair_accumulate(v[q], w[q], b, acc[2 * h + q][0], acc[2 * h + q][1]);
	}
	const uint32_t i0 = tile_first + 4u * g;
	if (i0 >= n) return;
#pragma unroll
	for (int j = 0; j < 2; ++j) {
		__half* row = out + (size_t)(level * 8u + 4u * b + 2u * j) * n + i0;  // features 4b + 2j (low halves) and 4b + 2j + 1 (high halves)
		if (i0 + 3u < n && (n % 4u == 0)) {
			const __half2 l01 = __lows2half2(acc[0][j], acc[1][j]), l23 = __lows2half2(acc[2][j], acc[3][j]);
			const __half2 h01 = __highs2half2(acc[0][j], acc[1][j]), h23 = __highs2half2(acc[2][j], acc[3][j]);
			st_stream_v2_u32(row, *reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
			st_stream_v2_u32(row + n, *reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
		} else {
#pragma unroll
			for (uint32_t p = 0; p < 4; ++p)
				if (i0 + p < n) {
					st_stream_u16(row + p, __half_as_ushort(__low2half(acc[p][j])));
					st_stream_u16(row + n + p, __half_as_ushort(__high2half(acc[p][j])));
				}
		}
	}
}

// TILES consecutive tiles per CTA, the positions of tile t + 1 requested before the gathers of tile t.  Measured (profiles/r2_hash_fwd_tiles_ab.txt): more
// than one tile per CTA is SLOWER (1.51 -> 1.61 ms per 4M samples) -- the barrier per tile couples the 16 level-warps, while a one-tile CTA has no barrier
// after its staging and a finished warp's slot goes to the next CTA -- so TILES = 1 ships.  What helps the one-tile CTA is an L2 prefetch of the records
// a CTA `pf_tiles` further on will stage (one cp.async.bulk.prefetch per CTA): its 16 warps otherwise sit through a DRAM round trip at the barrier
// (ncu: barrier stall 1.9 per issue); 300 tiles ahead: -2 %.
template <int TILES>
__global__ void __launch_bounds__(kFwdThreads, 2) hash_fwd_pairs_kernel(const __grid_constant__ GridLevels L, uint32_t n, const float* __restrict__ pos,
                                                                     uint32_t stride, const uint4* __restrict__ table, __half* __restrict__ out, uint32_t pf_tiles) {
	__shared__ float s_pos[TILES > 1 ? 2 : 1][kTile * 3];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	uint32_t tile_first = blockIdx.x * (kTile * TILES);
	if (tile_first >= n) return;
	if (pf_tiles && threadIdx.x == 0) {
		const uint64_t first = (uint64_t)(blockIdx.x + pf_tiles) * (kTile * TILES);
		if (first + kTile * TILES <= n) {
			const uint64_t a0 = reinterpret_cast<uint64_t>(pos + first * stride) & ~15ull;
			const uint32_t bytes = (kTile * TILES * stride * 4u + 31u) & ~15u;
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(bytes) : "memory");
		}
	}
	stage_positions(s_pos[0], pos, stride, n, tile_first);
#pragma unroll 1
	for (int t = 0; t < TILES; ++t, tile_first += kTile) {
		const uint32_t next_first = tile_first + kTile;
		const bool more = t + 1 < TILES && next_first < n;  // CTA-uniform
		float nxt = 0.0f;
		if (more && threadIdx.x < kTile * 3u) {
			const uint32_t j = threadIdx.x / 3u, d = threadIdx.x - 3u * j, i = next_first + j;
			if (i < n) nxt = ld_stream_f32(pos + (size_t)i * stride + d);
		}
		const float* sp = s_pos[TILES > 1 ? (t & 1) : 0];
		for (uint32_t level = warp; level < L.n_levels; level += kFwdWarps) {
			if (L.hashed[level]) fwd_level_pairs<true>(L, level, n, tile_first, lane, sp, table, out);
			else fwd_level_pairs<false>(L, level, n, tile_first, lane, sp, table, out);
		}
		if (!more) break;
		if (threadIdx.x < kTile * 3u) s_pos[TILES > 1 ? ((t + 1) & 1) : 0][threadIdx.x] = nxt;
		__syncthreads();
	}
}

__global__ void __launch_bounds__(256) hash_slots_kernel(const __grid_constant__ GridLevels L, uint32_t n, const float* __restrict__ pos,
                                                         uint32_t stride, uint32_t* __restrict__ slots) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t level = blockIdx.y;
	if (i >= n) return;
	const float* p = pos + (size_t)i * stride;
	const Corner c = pos_fract3(__ldg(p), __ldg(p + 1), __ldg(p + 2), L.scale[level]);
	for (uint32_t k = 0; k < 8; ++k) {
		const uint32_t gx = c.g[0] + (k & 1u), gy = c.g[1] + ((k >> 1) & 1u), gz = c.g[2] + ((k >> 2) & 1u);
		const uint32_t s = L.hashed[level] ? slot_of<true>(gx, gy, gz, L.size[level], 0, 0) : slot_of<false>(gx, gy, gz, L.size[level], L.mul1[level], L.mul2[level]);
		slots[((size_t)i * L.n_levels + level) * 8 + k] = L.offset[level] + s;
	}
}

// one 16-byte vector reduction per corner: 8 fp16 adds performed atomically at L2 (sm_90+)
__device__ __forceinline__ void red_add_v4_f16x2(void* addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
	asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1,%2,%3,%4};" ::"l"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Same tiling as the forward kernel: warp = level, lane = sample pair, positions staged once per tile.
template <bool HASHED>
__device__ __forceinline__ void bwd_sample(const uint4* grad_base, const Corner& c, const float g[8], uint32_t size, uint32_t mul1, uint32_t mul2) {
	uint4* grad = const_cast<uint4*>(grad_base);
#pragma unroll
	for (uint32_t k = 0; k < 8; ++k) {
		float wk = 1.0f;
		uint32_t q[3];
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			if ((k & (1u << d)) == 0) {
				wk = __fmul_rn(wk, __fsub_rn(1.0f, c.fr[d]));
				q[d] = c.g[d];
			} else {
				wk = __fmul_rn(wk, c.fr[d]);
				q[d] = c.g[d] + 1;
			}
		}
		uint32_t packed[4];
#pragma unroll
		for (int h = 0; h < 4; ++h) {
			const __half2 v = __floats2half2_rn(__fmul_rn(g[2 * h], wk), __fmul_rn(g[2 * h + 1], wk));  // grid.h:352
			packed[h] = *reinterpret_cast<const uint32_t*>(&v);
		}
		red_add_v4_f16x2(grad + slot_of<HASHED>(q[0], q[1], q[2], size, mul1, mul2), packed[0], packed[1], packed[2], packed[3]);
	}
}

template <bool HASHED>
__device__ __forceinline__ void bwd_level(const GridLevels& L, uint32_t level, uint32_t n, uint32_t i0, bool one, bool two, const float* __restrict__ sp,
                                          const __half* __restrict__ dLdy, uint4* __restrict__ grad) {
	// n = row pitch of dLdy (samples per feature row)
	if (!one) return;
	const uint32_t size = L.size[level], mul1 = L.mul1[level], mul2 = L.mul2[level];
	const float scale = L.scale[level];
	grad += L.offset[level];
	float g0[8], g1[8];
	const __half* row = dLdy + (size_t)(level * 8) * n + i0;
	if (two && (n % 2u == 0)) {
#pragma unroll
		for (int f = 0; f < 8; ++f) {
			const float2 v = __half22float2(*reinterpret_cast<const __half2*>(row + (size_t)f * n));
			g0[f] = v.x;
			g1[f] = v.y;
		}
	} else {
#pragma unroll
		for (int f = 0; f < 8; ++f) {
			g0[f] = __half2float(row[(size_t)f * n]);
			g1[f] = two ? __half2float(row[(size_t)f * n + 1]) : 0.0f;
		}
	}
	bool any0 = false, any1 = false;
#pragma unroll
	for (int f = 0; f < 8; ++f) {
		any0 |= (g0[f] != 0.0f);
		any1 |= (g1[f] != 0.0f);
	}
	// adding +0 to every slot is a no-op
	if (any0) bwd_sample<HASHED>(grad, pos_fract3(sp[0], sp[1], sp[2], scale), g0, size, mul1, mul2);
	if (two && any1) bwd_sample<HASHED>(grad, pos_fract3(sp[3], sp[4], sp[5], scale), g1, size, mul1, mul2);
}

// levels [level_begin, level_end): one warp per level (the block has that many warps, at most kFwdWarps)
__global__ void __launch_bounds__(kBwdThreads) hash_bwd_kernel(const __grid_constant__ GridLevels L, uint32_t n, uint32_t pitch, const float* __restrict__ pos,
                                                               uint32_t stride, const __half* __restrict__ dLdy, uint4* __restrict__ grad, uint32_t level_begin,
                                                               uint32_t level_end) {
	__shared__ float s_pos[kTile * 3];
	const uint32_t tile_first = blockIdx.x * kTile;
	stage_positions(s_pos, pos, stride, n, tile_first);
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, n_warps = blockDim.x >> 5;
	const uint32_t i0 = tile_first + 2u * lane;
	const bool one = i0 < n, two = i0 + 1 < n;
	for (uint32_t level = level_begin + warp; level < level_end; level += n_warps) {
		if (L.hashed[level]) bwd_level<true>(L, level, pitch, i0, one, two, s_pos + 6 * lane, dLdy, grad);
		else bwd_level<false>(L, level, pitch, i0, one, two, s_pos + 6 * lane, dLdy, grad);
	}
}

// Lane-pair version (default), the mapping of hash_fwd_pairs_kernel: lane l = x-side (l & 1) of the samples 4 (l >> 1) + {0, 1, 2, 3} of the tile.  In every
// reduction instruction the two lanes of a pair address the entries (x, y, z) and (x + 1, y, z): neighbours in the table, the same 32-byte sector whenever the
// slot of the x-side 0 is even, which the memory system then handles as ONE request (insta_microbench_scatter: 32-byte chunks shared by two lanes against
// 16-byte chunks).  Same products and the same fp16 rounding as kernel_grid_backward (grid.h:352); the order of the additions is not defined in the reference either.
// TILED: dLdy in the layout the fused network kernel (mlp_bwd.cu) produces -- [tile of 128 samples][level][sample][8 features], the 8 gradients of a level in
// one 16-byte load -- instead of kernel_grid_backward's SoA rows.
template <bool HASHED, bool TILED>
__device__ __forceinline__ void bwd_level_pairs(const GridLevels& L, uint32_t level, uint32_t n, uint32_t pitch, uint32_t tile_first, uint32_t lane,
                                                const float* __restrict__ s_pos, const __half* __restrict__ dLdy, uint4* __restrict__ grad) {
	const uint32_t b = lane & 1u, g = lane >> 1;
	const uint32_t i0 = tile_first + 4u * g;
	if (i0 >= n) return;
	const uint32_t size = L.size[level], mul1 = L.mul1[level], mul2 = L.mul2[level];
	const float scale = L.scale[level];
	grad += L.offset[level];
	const __half* row = dLdy + (size_t)(level * 8) * pitch + i0;
	uint32_t gr[4][4];  // [sample][feature pair] as half2 bits
	if (TILED) {
#pragma unroll
		for (uint32_t q = 0; q < 4; ++q) {
			const uint32_t i = i0 + q;
			uint4 v = make_uint4(0, 0, 0, 0);
			if (i < n) v = __ldg(reinterpret_cast<const uint4*>(dLdy + ((size_t)(i >> 7) * 16u + level) * 1024u + (i & 127u) * 8u));
			gr[q][0] = v.x; gr[q][1] = v.y; gr[q][2] = v.z; gr[q][3] = v.w;
		}
	} else if (i0 + 3u < n && (pitch % 4u == 0)) {
		uint2 raw[8];
#pragma unroll
		for (int f = 0; f < 8; ++f) raw[f] = __ldg(reinterpret_cast<const uint2*>(row + (size_t)f * pitch));  // feature f of the 4 samples
#pragma unroll
		for (int h = 0; h < 4; ++h) {
			gr[0][h] = __byte_perm(raw[2 * h].x, raw[2 * h + 1].x, 0x5410);
			gr[1][h] = __byte_perm(raw[2 * h].x, raw[2 * h + 1].x, 0x7632);
			gr[2][h] = __byte_perm(raw[2 * h].y, raw[2 * h + 1].y, 0x5410);
			gr[3][h] = __byte_perm(raw[2 * h].y, raw[2 * h + 1].y, 0x7632);
		}
	} else {
#pragma unroll
		for (uint32_t q = 0; q < 4; ++q)
#pragma unroll
			for (int h = 0; h < 4; ++h) {
				const uint32_t lo = i0 + q < n ? __half_as_ushort(row[(size_t)(2 * h) * pitch + q]) : 0u;
				const uint32_t hi = i0 + q < n ? __half_as_ushort(row[(size_t)(2 * h + 1) * pitch + q]) : 0u;
				gr[q][h] = lo | (hi << 16);
			}
	}
#pragma unroll
	for (uint32_t q = 0; q < 4; ++q) {
		// adding +0 to every slot is a no-op (and -0 + x = x): skip samples whose gradient is all zeros
		if (((gr[q][0] | gr[q][1] | gr[q][2] | gr[q][3]) & 0x7fff7fffu) == 0u) continue;
		const float* sp = s_pos + 3u * (4u * g + q);
		const Corner c = pos_fract3(sp[0], sp[1], sp[2], scale);
		const float wx = b ? c.fr[0] : __fsub_rn(1.0f, c.fr[0]);  // 1 * wx is exact
		const uint32_t gx = c.g[0] + b;
		float gf[8];
#pragma unroll
		for (int h = 0; h < 4; ++h) {
			const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&gr[q][h]));
			gf[2 * h] = f.x;
			gf[2 * h + 1] = f.y;
		}
#pragma unroll
		for (uint32_t k = 0; k < 4; ++k) {
			const uint32_t dy = k & 1u, dz = k >> 1;
			float wk = __fmul_rn(wx, dy ? c.fr[1] : __fsub_rn(1.0f, c.fr[1]));
			wk = __fmul_rn(wk, dz ? c.fr[2] : __fsub_rn(1.0f, c.fr[2]));
			uint32_t packed[4];
#pragma unroll
			for (int h = 0; h < 4; ++h) {
				const __half2 v = __floats2half2_rn(__fmul_rn(gf[2 * h], wk), __fmul_rn(gf[2 * h + 1], wk));  // grid.h:352
				packed[h] = *reinterpret_cast<const uint32_t*>(&v);
			}
			red_add_v4_f16x2(grad + slot_of<HASHED>(gx, c.g[1] + dy, c.g[2] + dz, size, mul1, mul2), packed[0], packed[1], packed[2], packed[3]);
		}
	}
}

template <bool TILED>
__global__ void __launch_bounds__(kBwdThreads) hash_bwd_pairs_kernel(const __grid_constant__ GridLevels L, uint32_t n, uint32_t pitch, const float* __restrict__ pos,
                                                                     uint32_t stride, const __half* __restrict__ dLdy, uint4* __restrict__ grad, uint32_t level_begin,
                                                                     uint32_t level_end) {
	__shared__ float s_pos[kTile * 3];
	const uint32_t tile_first = blockIdx.x * kTile;
	stage_positions(s_pos, pos, stride, n, tile_first);
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, n_warps = blockDim.x >> 5;
	for (uint32_t level = level_begin + warp; level < level_end; level += n_warps) {
		if (L.hashed[level]) bwd_level_pairs<true, TILED>(L, level, n, pitch, tile_first, lane, s_pos, dLdy, grad);
		else bwd_level_pairs<false, TILED>(L, level, n, pitch, tile_first, lane, s_pos, dLdy, grad);
	}
}

}  // namespace

// Offset table exactly as the ctor builds it (grid.h:898-922), with the per-level scale taken from the device.
int init_grid_levels(insta_ctx* ctx) {
	const insta_params& p = ctx->p;
	if (p.n_levels == 0 || p.n_levels > (uint32_t)kMaxLevels) INSTA_FAIL(INSTA_ERR_INVALID, "n_levels must be in 1..16");
	if (p.n_features_per_level != 8) INSTA_FAIL(INSTA_ERR_INVALID, "only n_features_per_level == 8 is implemented");
	if (p.log2_hashmap_size < 4 || p.log2_hashmap_size > 28) INSTA_FAIL(INSTA_ERR_INVALID, "log2_hashmap_size out of range");
	float pls = p.per_level_scale;
	if (pls <= 0.f) {
		// ngp/src/testbed.cu:2246-2258: desired_resolution 2048
		pls = p.n_levels > 1 ? std::exp(std::log(2048.0f * (float)p.aabb_scale / (float)p.base_resolution) / (float)(p.n_levels - 1)) : 1.0f;  // ngp testbed.cu:2255 guards the same way
		ctx->p.per_level_scale = pls;
	}
	float* d_scales = nullptr;
	INSTA_CUDA(cudaMalloc(&d_scales, sizeof(float) * kMaxLevels));
	level_scale_kernel<<<1, 32>>>(p.n_levels, std::log2(pls), p.base_resolution, d_scales);
	ctx->launches++;
	float scales[kMaxLevels];
	cudaError_t e = cudaMemcpy(scales, d_scales, sizeof(float) * kMaxLevels, cudaMemcpyDeviceToHost);
	cudaFree(d_scales);
	INSTA_CUDA(e);

	GridLevels& L = ctx->levels;
	L = GridLevels{};
	L.n_levels = p.n_levels;
	uint32_t offset = 0;
	for (uint32_t l = 0; l < p.n_levels; ++l) {
		const float scale = scales[l];
		const uint32_t res = (uint32_t)std::ceil(scale) + 1;
		const uint32_t max_params = 0xffffffffu / 2;
		uint32_t params = std::pow((float)res, 3) > (float)max_params ? max_params : res * res * res;
		params = (params + 7u) / 8u * 8u;
		params = std::min(params, 1u << p.log2_hashmap_size);
		// grid_index's loop (grid.h:105-110) in the same uint32 arithmetic, to learn which dims it accumulates
		uint32_t stride = 1, mul[3] = {0, 0, 0};
		for (uint32_t dim = 0; dim < 3 && stride <= params; ++dim) {
			mul[dim] = stride;
			stride *= res;
		}
		L.offset[l] = offset;
		L.size[l] = params;
		L.res[l] = res;
		L.mul1[l] = mul[1];
		L.mul2[l] = mul[2];
		L.hashed[l] = params < stride ? 1u : 0u;
		L.scale[l] = scale;
		if (L.hashed[l] && (params & (params - 1)) != 0) INSTA_FAIL(INSTA_ERR_INVALID, "hashed level with a non power-of-two size");
		offset += params;
	}
	L.offset[p.n_levels] = offset;
	return INSTA_OK;
}

int launch_hash_fwd(insta_ctx* ctx, const float* d_pos, uint32_t stride, uint32_t n, const void* table, void* out, cudaStream_t s) {
	if (n == 0) return INSTA_OK;
	const dim3 grid((n + kTile - 1) / kTile);
	static const bool v1 = [] { const char* e = getenv("INSTA_HASH_FWD"); return e && std::string(e) == "v1"; }();  // A/B: the sample-pair-per-lane kernel of round 1
	if (v1) hash_fwd_kernel<<<grid, kFwdThreads, 0, s>>>(ctx->levels, n, d_pos, stride, (const uint4*)table, (__half*)out);
	else {
		static const uint32_t pf = [] { const char* e = getenv("INSTA_HASH_PF"); return e ? (uint32_t)atoi(e) : 300u; }();  // tiles ahead (0 = no prefetch; A/B)
		static const bool t4 = [] { const char* e = getenv("INSTA_HASH_FWD"); return e && std::string(e) == "t4"; }();       // A/B: four tiles per CTA
		if (t4) hash_fwd_pairs_kernel<4><<<(grid.x + 3) / 4, kFwdThreads, 0, s>>>(ctx->levels, n, d_pos, stride, (const uint4*)table, (__half*)out, pf / 4);
		else hash_fwd_pairs_kernel<1><<<grid, kFwdThreads, 0, s>>>(ctx->levels, n, d_pos, stride, (const uint4*)table, (__half*)out, pf);
	}
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

int launch_hash_bwd(insta_ctx* ctx, const float* d_pos, uint32_t stride, uint32_t n, const void* dLdy, void* grad, int zero_first, cudaStream_t s) {
	return launch_hash_bwd_pitched(ctx, d_pos, stride, n, dLdy, n, grad, zero_first, s);
}

// dLdy rows are `pitch` samples apart (pitch >= n); pitch == 0: the tiled layout of the fused network kernel (bwd_level_pairs)
int launch_hash_bwd_pitched(insta_ctx* ctx, const float* d_pos, uint32_t stride, uint32_t n, const void* dLdy, uint32_t pitch, void* grad, int zero_first, cudaStream_t s) {
	return launch_hash_bwd_levels(ctx, d_pos, stride, n, dLdy, pitch, grad, zero_first, 0, ctx->levels.n_levels, s);
}

// the same for the levels [level_begin, level_end) only (zero_first clears just their part of the gradient table): the trainer's pipelined exchange
int launch_hash_bwd_levels(insta_ctx* ctx, const float* d_pos, uint32_t stride, uint32_t n, const void* dLdy, uint32_t pitch, void* grad, int zero_first, uint32_t level_begin,
                           uint32_t level_end, cudaStream_t s) {
	const GridLevels& L = ctx->levels;
	if (level_end > L.n_levels || level_begin >= level_end) INSTA_FAIL(INSTA_ERR_INVALID, "launch_hash_bwd_levels: bad level range");
	if (zero_first) INSTA_CUDA(cudaMemsetAsync((unsigned char*)grad + (size_t)L.offset[level_begin] * 16, 0, (size_t)(L.offset[level_end] - L.offset[level_begin]) * 16, s));
	if (n == 0) return INSTA_OK;
	const dim3 grid((n + kTile - 1) / kTile);
	const uint32_t warps = std::min<uint32_t>(level_end - level_begin, (uint32_t)kFwdWarps);
	static const bool v1 = [] { const char* e = getenv("INSTA_HASH_BWD"); return e && std::string(e) == "v1"; }();  // A/B: the sample-pair-per-lane kernel of round 1
	if (pitch == 0) hash_bwd_pairs_kernel<true><<<grid, warps * 32, 0, s>>>(L, n, pitch, d_pos, stride, (const __half*)dLdy, (uint4*)grad, level_begin, level_end);
	else if (v1) hash_bwd_kernel<<<grid, warps * 32, 0, s>>>(L, n, pitch, d_pos, stride, (const __half*)dLdy, (uint4*)grad, level_begin, level_end);
	else hash_bwd_pairs_kernel<false><<<grid, warps * 32, 0, s>>>(L, n, pitch, d_pos, stride, (const __half*)dLdy, (uint4*)grad, level_begin, level_end);
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

int launch_hash_slots(insta_ctx* ctx, const float* d_pos, uint32_t stride, uint32_t n, uint32_t* slots, cudaStream_t s) {
	if (n == 0) return INSTA_OK;
	const dim3 grid((n + 255) / 256, ctx->levels.n_levels);
	hash_slots_kernel<<<grid, 256, 0, s>>>(ctx->levels, n, d_pos, stride, slots);
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

}  // namespace insta
