// Device functions of the multiresolution hash grid shared by hashgrid.cu (stand-alone encode) and mlp.cu (encode fused
// into the network): pos_fract (tcnn common_device.h:415-431), fast_hash / grid_index (tcnn encodings/grid.h:82-117) and the
// trilinear interpolation with fp16 accumulation of kernel_grid (grid.h:207-266).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace insta {
namespace griddev {

struct Corner {
	uint32_t g[3];
	float fr[3];
};

__device__ __forceinline__ Corner pos_fract3(float x, float y, float z, float scale) {
	Corner c;
	const float p[3] = {x, y, z};
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		const float v = fmaf(p[d], scale, 0.5f);
		const float fl = floorf(v);
		c.g[d] = (uint32_t)(int)fl;
		c.fr[d] = v - fl;
	}
	return c;
}

template <bool HASHED>
__device__ __forceinline__ uint32_t slot_of(uint32_t gx, uint32_t gy, uint32_t gz, uint32_t size, uint32_t mul1, uint32_t mul2) {
	uint32_t idx;
	if (HASHED) {
		idx = gx ^ (gy * 2654435761u) ^ (gz * 805459861u);
		return idx & (size - 1);  // hashed levels always have size == 2^log2_hashmap_size
	}
	idx = gx + gy * mul1 + gz * mul2;
	return idx % size;
}

__device__ __forceinline__ uint4 ld_entry(const uint4* p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
	float r;
	asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
	return r;
}
__device__ __forceinline__ void st_stream_u32(void* p, uint32_t v) { asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_stream_v2_u32(void* p, uint32_t a, uint32_t b) { asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void st_stream_u16(void* p, uint16_t v) { asm volatile("st.global.cs.u16 [%0], %1;" ::"l"(p), "h"(v) : "memory"); }

// acc += (half2)(w * (float2)v)  per feature pair: fp32 multiply, round to fp16, fp16 add -- grid.h:257-259
__device__ __forceinline__ void accum(__half2 acc[4], const uint4 v, float w) {
	const uint32_t raw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&raw[k]));
		acc[k] = __hadd2(acc[k], __floats2half2_rn(__fmul_rn(w, f.x), __fmul_rn(w, f.y)));
	}
}

// One x-side of a sample (4 of its 8 corners) per lane: lanes 2j and 2j+1 of a warp hold the corners (x, y+dy, z+dz) and (x+1, y+dy, z+dz) of the same
// sample, so in every load instruction the two lanes of a pair hit neighbouring table entries -- the SAME 32-byte sector whenever x is even (dense index
// x + ..., hashed index x ^ ...: bit 0 of x is bit 0 of the slot).  L1 charges a divergent gather per distinct sector, not per lane (measured with
// insta_microbench_gather on the 49 MB table: 4.3 TB/s for 16-byte chunks, 8.7 TB/s when two adjacent lanes share a 32-byte chunk), so this removes a
// quarter of the tag-stage work of kernel_grid's access pattern.  The fp16 accumulation keeps kernel_grid's corner order (grid.h:207-266: corner index
// bit 0 = x): each lane accumulates HALF of the features over all 8 corners and gets the partner's products for those features by shuffle.
//   b = x-side of this lane (0 / 1); result = features 4b .. 4b+3 of the sample as two half2
template <bool HASHED>
__device__ __forceinline__ void pair_load(const uint4* __restrict__ table, const Corner& c, uint32_t b, uint32_t size, uint32_t mul1, uint32_t mul2, uint4 v[4], float w[4]) {
	const float wx = b ? c.fr[0] : __fsub_rn(1.0f, c.fr[0]);  // 1 * wx is exact
	const uint32_t gx = c.g[0] + b;
#pragma unroll
	for (uint32_t k = 0; k < 4; ++k) {
		const uint32_t dy = k & 1u, dz = k >> 1;
		float wk = __fmul_rn(wx, dy ? c.fr[1] : __fsub_rn(1.0f, c.fr[1]));
		wk = __fmul_rn(wk, dz ? c.fr[2] : __fsub_rn(1.0f, c.fr[2]));
		w[k] = wk;
		v[k] = ld_entry(table + slot_of<HASHED>(gx, c.g[1] + dy, c.g[2] + dz, size, mul1, mul2));
	}
}
__device__ __forceinline__ void pair_accumulate(const uint4 v[4], const float w[4], uint32_t b, __half2& out0, __half2& out1) {
	__half2 a0 = __half2(__ushort_as_half(0), __ushort_as_half(0)), a1 = a0;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const uint32_t raw[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
		uint32_t p[4];
#pragma unroll
		for (int h = 0; h < 4; ++h) {
			const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&raw[h]));
			const __half2 pr = __floats2half2_rn(__fmul_rn(w[k], f.x), __fmul_rn(w[k], f.y));
			p[h] = *reinterpret_cast<const uint32_t*>(&pr);
		}
		// keep the pair of half2 with this lane's features, hand the other pair to the partner
		const uint32_t r0 = __shfl_xor_sync(0xffffffffu, b ? p[0] : p[2], 1), r1 = __shfl_xor_sync(0xffffffffu, b ? p[1] : p[3], 1);
		const uint32_t own0 = b ? p[2] : p[0], own1 = b ? p[3] : p[1];
		// corner order: x-side 0 first
		const uint32_t f0 = b ? r0 : own0, s0 = b ? own0 : r0, f1 = b ? r1 : own1, s1 = b ? own1 : r1;
		a0 = __hadd2(__hadd2(a0, *reinterpret_cast<const __half2*>(&f0)), *reinterpret_cast<const __half2*>(&s0));
		a1 = __hadd2(__hadd2(a1, *reinterpret_cast<const __half2*>(&f1)), *reinterpret_cast<const __half2*>(&s1));
	}
	out0 = a0;
	out1 = a1;
}

template <bool HASHED>
__device__ __forceinline__ void encode_one(const uint4* __restrict__ table, const Corner& c, uint32_t size, uint32_t mul1, uint32_t mul2, __half2 acc[4]) {
	uint4 v[8];
	float w[8];
#pragma unroll
	for (uint32_t k = 0; k < 8; ++k) {
		float wk = 1.0f;
		uint32_t g[3];
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			if ((k & (1u << d)) == 0) {
				wk = __fmul_rn(wk, __fsub_rn(1.0f, c.fr[d]));
				g[d] = c.g[d];
			} else {
				wk = __fmul_rn(wk, c.fr[d]);
				g[d] = c.g[d] + 1;
			}
		}
		w[k] = wk;
		v[k] = ld_entry(table + slot_of<HASHED>(g[0], g[1], g[2], size, mul1, mul2));
	}
#pragma unroll
	for (int k = 0; k < 4; ++k) acc[k] = __half2(__ushort_as_half(0), __ushort_as_half(0));
#pragma unroll
	for (int k = 0; k < 8; ++k) accum(acc, v[k], w[k]);
}

}  // namespace griddev
}  // namespace insta
