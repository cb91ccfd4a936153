// Device functions shared by nerf.cu (render tracer) and train.cu (training sample generation / loss): step size, the
// reference's low-discrepancy sequence, occupancy-grid lookup and the DDA skip (ngp/src/testbed_nerf.cu:48-93, 186-208,
// 313-337, 471-485; ngp random_val.cuh:160-288; ngp bounding_box.cuh:163-221).  Expressions follow the reference's evaluation
// order with nvcc's contractions written out (see nerf.cu).
#pragma once
#include <cuda_fp16.h>
#include <float.h>
#include <stdint.h>

namespace insta {
namespace nerfdev {

// pcg32 (tcnn pcg32.h): the generator every reference kernel draws from
struct Pcg32 {
	uint64_t state, inc;
	__device__ __forceinline__ uint32_t next_uint() {
		const uint64_t old = state;
		state = old * 0x5851f42d4c957f2dULL + inc;
		const uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
		const uint32_t rot = (uint32_t)(old >> 59u);
		return (xs >> rot) | (xs << ((~rot + 1u) & 31));
	}
	__device__ __forceinline__ float next_float() { return __uint_as_float((next_uint() >> 9) | 0x3f800000u) - 1.0f; }
	__device__ __forceinline__ void advance(uint64_t delta) {
		uint64_t cur_mult = 0x5851f42d4c957f2dULL, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
		while (delta > 0) {
			if (delta & 1) {
				acc_mult *= cur_mult;
				acc_plus = acc_plus * cur_mult + cur_plus;
			}
			cur_plus = (cur_mult + 1) * cur_plus;
			cur_mult *= cur_mult;
			delta >>= 1;
		}
		state = acc_mult * state + acc_plus;
	}
};

constexpr uint32_t kGrid = 128;        // NERF_GRIDSIZE (ngp nerf.h:24-26)
constexpr uint32_t kNerfSteps = 1024;  // NERF_STEPS
constexpr uint32_t kCascades = 8;      // NERF_CASCADES
__device__ __forceinline__ float sqrt3() { return 1.73205080757f; }
__device__ __forceinline__ float min_stepsize() { return sqrt3() / kNerfSteps; }
__device__ __forceinline__ float max_stepsize() { return min_stepsize() * (1 << (kCascades - 1)) * kNerfSteps / kGrid; }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }  // tcnn::clamp = max(lo, min(hi, v))
// tcnn::clamp(val, lower, upper) = val < lower ? lower : (upper < val ? upper : val)   (tcnn common.h:243-246)
__device__ __forceinline__ float tclamp(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }
__device__ __forceinline__ int tclampi(int v, int lo, int hi) { return v < lo ? lo : (hi < v ? hi : v); }
__device__ __forceinline__ float calc_dt(float t, float cone_angle) { return tclamp(t * cone_angle, min_stepsize(), max_stepsize()); }
__device__ __forceinline__ float sgn(float x) { return copysignf(1.0f, x); }

// ---- low-discrepancy sequence of the reference (ngp random_val.cuh:160-288): Sobol dims 0/1 + Laine-Karras scrambling ----
__device__ __forceinline__ uint32_t lk_perm(uint32_t x, uint32_t seed) {
	x += seed;
	x ^= x * 0x6c50b47cu;
	x ^= x * 0xb82f1e52u;
	x ^= x * 0xc7afe638u;
	x ^= x * 0x8d22f6e6u;
	return x;
}
__device__ __forceinline__ uint32_t scramble(uint32_t x, uint32_t seed) { return __brev(lk_perm(__brev(x), seed)); }
__device__ __forceinline__ uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
// direction numbers of dimension 0 are 1 << (31 - bit): the van der Corput sequence
__device__ __forceinline__ uint32_t sobol0(uint32_t index) { return __brev(index); }
// dimension 1: v[0] = 0x80000000, v[k] = v[k-1] ^ (v[k-1] >> 1)   (same table as random_val.cuh:171-178)
__device__ __forceinline__ uint32_t sobol1(uint32_t index) {
	uint32_t x = 0, v = 0x80000000u;
	for (uint32_t bit = 0; bit < 32; ++bit) {
		if ((index >> bit) & 1u) x ^= v;
		v ^= v >> 1;
	}
	return x;
}
__device__ __forceinline__ float ld_random_val(uint32_t index, uint32_t seed) {
	const float S = float(1.0 / (1ull << 32));
	index = scramble(index, seed);
	return (float)scramble(sobol0(index), hash_combine(seed, 0)) * S;
}
// the two raw 32-bit values of ld_random_val_2d (value = (float)raw * 2^-32)
__device__ __forceinline__ uint2 ld_random_raw_2d(uint32_t index, uint32_t seed) {
	index = scramble(index, seed);
	return make_uint2(scramble(sobol0(index), hash_combine(seed, 0)), scramble(sobol1(index), hash_combine(seed, 1)));
}
__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }
// fract(0.5 - ld(0) + ld(spp)): the reference build folds 0.5 - ld(0) into a constant and contracts the rest into one FFMA
__device__ __forceinline__ float2 ld_random_pixel_offset(uint32_t spp) {
	const float S = float(1.0 / (1ull << 32));
	const uint2 a = ld_random_raw_2d(0, 0xdeadbeefu), b = ld_random_raw_2d(spp, 0xdeadbeefu);
	const float cx = __fsub_rn(0.5f, __fmul_rn((float)a.x, S)), cy = __fsub_rn(0.5f, __fmul_rn((float)a.y, S));
	return make_float2(fractf(fmaf((float)b.x, S, cx)), fractf(fmaf((float)b.y, S, cy)));
}

// ---- occupancy grid (testbed_nerf.cu:313-337, 471-485; tcnn common_device.h:339-354) ----
__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}
__device__ __forceinline__ uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }

__device__ __forceinline__ int mip_from_pos(float px, float py, float pz, uint32_t max_cascade = kCascades - 1) {
	int exponent;
	const float maxval = fmaxf(fmaxf(fabsf(px - 0.5f), fabsf(py - 0.5f)), fabsf(pz - 0.5f));
	frexpf(maxval, &exponent);
	return min((int)max_cascade - 1, max(0, exponent + 1));
}
__device__ __forceinline__ int mip_from_dt(float dt, float px, float py, float pz, uint32_t max_cascade = kCascades - 1) {
	const int mip = mip_from_pos(px, py, pz, max_cascade);
	dt *= 2 * kGrid;
	if (dt < 1.f) return mip;
	int exponent;
	frexpf(dt, &exponent);
	return min((int)max_cascade - 1, max(exponent, mip));
}
__device__ __forceinline__ uint32_t cascaded_grid_idx_at(float px, float py, float pz, uint32_t mip) {
	const float mip_scale = scalbnf(1.0f, -(int)mip);
	// pos -= 0.5; pos *= mip_scale; pos += 0.5  -- the reference build contracts the last two into one FFMA
	const float x = fmaf(px - 0.5f, mip_scale, 0.5f), y = fmaf(py - 0.5f, mip_scale, 0.5f), z = fmaf(pz - 0.5f, mip_scale, 0.5f);
	const int ix = (int)(x * (float)kGrid), iy = (int)(y * (float)kGrid), iz = (int)(z * (float)kGrid);
	return morton3D((uint32_t)tclampi(ix, 0, (int)kGrid - 1), (uint32_t)tclampi(iy, 0, (int)kGrid - 1), (uint32_t)tclampi(iz, 0, (int)kGrid - 1));
}
__device__ __forceinline__ bool occupied_at(float px, float py, float pz, const uint8_t* __restrict__ bitfield, uint32_t mip) {
	const uint32_t idx = cascaded_grid_idx_at(px, py, pz, mip);
	return (__ldg(bitfield + idx / 8 + (size_t)mip * (kGrid * kGrid * kGrid / 8)) >> (idx % 8)) & 1u;
}
// testbed_nerf.cu:186-208
__device__ __forceinline__ float distance_to_next_voxel(float px, float py, float pz, float dx, float dy, float dz, float ix, float iy, float iz, uint32_t res) {
	const float r = (float)res;
	const float x = r * px, y = r * py, z = r * pz;
	const float tx = (floorf(fmaf(0.5f, sgn(dx), x + 0.5f)) - x) * ix;
	const float ty = (floorf(fmaf(0.5f, sgn(dy), y + 0.5f)) - y) * iy;
	const float tz = (floorf(fmaf(0.5f, sgn(dz), z + 0.5f)) - z) * iz;
	const float t = fminf(fminf(tx, ty), tz);
	return fmaxf(t / r, 0.0f);
}
__device__ __forceinline__ float advance_to_next_voxel(float t, float cone_angle, float px, float py, float pz, float dx, float dy, float dz, float ix, float iy, float iz,
                                                       uint32_t res) {
	const float t_target = t + distance_to_next_voxel(px, py, pz, dx, dy, dz, ix, iy, iz, res);
	do {
		t += calc_dt(t, cone_angle);
	} while (t < t_target);
	return t;
}
__device__ __forceinline__ bool aabb_contains(const float* a, float x, float y, float z) {
	return x >= a[0] && x <= a[3] && y >= a[1] && y <= a[4] && z >= a[2] && z <= a[5];
}
// BoundingBox::ray_intersect (ngp bounding_box.cuh:163-208), returns tmin (FLT_MAX when missed)
__device__ __forceinline__ float aabb_ray_tmin(const float* a, float ox, float oy, float oz, float dx, float dy, float dz) {
	float tmin = (a[0] - ox) / dx, tmax = (a[3] - ox) / dx;
	if (tmin > tmax) { const float s = tmin; tmin = tmax; tmax = s; }
	float tymin = (a[1] - oy) / dy, tymax = (a[4] - oy) / dy;
	if (tymin > tymax) { const float s = tymin; tymin = tymax; tymax = s; }
	if (tmin > tymax || tymin > tmax) return FLT_MAX;
	if (tymin > tmin) tmin = tymin;
	if (tymax < tmax) tmax = tymax;
	float tzmin = (a[2] - oz) / dz, tzmax = (a[5] - oz) / dz;
	if (tzmin > tzmax) { const float s = tzmin; tzmin = tzmax; tzmax = s; }
	if (tmin > tzmax || tzmin > tmax) return FLT_MAX;
	if (tzmin > tmin) tmin = tzmin;
	return tmin;
}

}  // namespace nerfdev
}  // namespace insta
