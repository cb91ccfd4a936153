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
