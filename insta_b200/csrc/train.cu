// Training side of stage 3 on sm_100a: ray / sample generation with occupancy skipping, per-ray compositing + loss +
// gradient seed + compaction, batch roll-over, and the optimizer step.  Replaces (ngp/src/testbed_nerf.cu)
//   generate_training_samples_nerf 1105-1285 (+ image_idx 1082-1103, nerf_random_image_pos_training 1066-1080),
//   compute_loss_kernel_train_nerf 1305-1643 (Huber / 5 colour loss, L1 depth loss, segmentation weighting, density
//   regularisers), fill_rollover / fill_rollover_and_rescale (tcnn common_device.h:503-523),
// and tcnn's adam_step + ema_step_half_precision (optimizers/adam.h:48-114, ema.h:64-77) as ONE pass over the parameters.
//
// Design notes (not a port):
//   * the sample generator keeps the reference's two marches (count, reserve a contiguous span with one atomic, emit) because
//     the span layout is the contract with the loss kernel, but it never writes outside its own span (the reference's
//     first pass stomps `holder` of records 0..n of the GLOBAL buffer, a data race, testbed_nerf.cu:1233 vs 1253);
//   * the optimizer reads each parameter's state once: gradient un-scaling, Adam with per-parameter step counters, the
//     zero-gradient early-out for hash-grid entries, and the EMA of the fp16 weights are fused; untouched grid entries move
//     8 bytes instead of 2 passes x 6-10 bytes.
// Exactness: image choice, pixel choice (pcg32 stream), per-ray step counts and emitted records are bit-exact against the
// reference kernel for rays given explicitly (tests/test_gpu_train.py); loss and gradients are toleranced (fp16 outputs).
#include <cuda_fp16.h>
#include <float.h>

#include <algorithm>
#include <cmath>

#include "ctx.h"
#include "nerf_device.cuh"

namespace insta {
namespace {
using namespace nerfdev;

__device__ __forceinline__ float srgb_to_linear(float s) { return s <= 0.04045f ? s / 12.92f : powf((s + 0.055f) / 1.055f, 2.4f); }
__device__ __forceinline__ float linear_to_srgb(float l) { return l < 0.0031308f ? 12.92f * l : 1.055f * powf(l, 0.41666f) - 0.055f; }

// read_rgba (ngp common_device.cuh:653-685)
__device__ __forceinline__ float4 read_rgba(const ImageDev& im, uint32_t px, uint32_t py) {
	const size_t idx = (size_t)px + (size_t)py * im.res[0];
	if (im.type == 1u) {
		const uint32_t raw = reinterpret_cast<const uint32_t*>(im.pixels)[idx];
		if (raw == 0x00FF00FFu) return make_float4(-1.f, -1.f, -1.f, -1.f);
		const float alpha = (float)(raw >> 24) * (1.0f / 255.0f);
		return make_float4(srgb_to_linear((float)(raw & 0xff) * (1.0f / 255.0f)) * alpha, srgb_to_linear((float)((raw >> 8) & 0xff) * (1.0f / 255.0f)) * alpha,
		                   srgb_to_linear((float)((raw >> 16) & 0xff) * (1.0f / 255.0f)) * alpha, alpha);
	}
	if (im.type == 2u) {
		const uint2 raw = reinterpret_cast<const uint2*>(im.pixels)[idx];
		const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
		return make_float4(lo.x, lo.y, hi.x, hi.y);
	}
	return reinterpret_cast<const float4*>(im.pixels)[idx];
}

// image_idx without an error CDF (testbed_nerf.cu:1098-1102): wrapping uint32 arithmetic
__device__ __forceinline__ uint32_t image_idx(uint32_t base_idx, uint32_t n_rays, uint32_t n_rays_total, uint32_t n_images) {
	return (((base_idx + n_rays_total) * n_images) / n_rays) % n_images;
}
// nerf_random_image_pos_training + image_pos (testbed_nerf.cu:1066-1080; common_device.cuh:609-611)
__device__ __forceinline__ void image_pos_training(Pcg32& rng, const ImageDev& im, bool snap, float& x, float& y, uint32_t& px, uint32_t& py) {
	x = rng.next_float();
	y = rng.next_float();
	const float rx = (float)im.res[0], ry = (float)im.res[1];
	if (snap) {
		const int ix = min(max((int)(x * rx), 0), (int)im.res[0] - 1), iy = min(max((int)(y * ry), 0), (int)im.res[1] - 1);
		x = ((float)ix + 0.5f) / rx;
		y = ((float)iy + 0.5f) / ry;
	}
	px = (uint32_t)max(min((int)(x * rx), (int)im.res[0] - 1), 0);
	py = (uint32_t)max(min((int)(y * ry), (int)im.res[1] - 1), 0);
}

constexpr uint32_t kGenKeep = 64;  // ray parameters kept from the counting march (32 KB of shared memory per 128-thread CTA)
struct GenArgs {
	uint32_t n_rays, n_rays_total, max_samples, n_images, stride, n_extra;
	uint64_t rng_state, rng_inc;
	float aabb[6];
	float cone_angle;
	int snap;
	const ImageDev* images;
	const uint8_t* bitfield;
	const float* extra_dims;  // n_images x n_extra
	float* coords;
	uint32_t* ray_counter;
	uint32_t* numsteps_counter;
	uint32_t* ray_indices;
	float* rays;       // 6 floats per accepted ray
	uint32_t* numsteps;  // (count, base) per accepted ray
	const uint32_t* emit_index;  // train_emit_kernel<true>: the records to expand (a round's list) ...
	const uint32_t* emit_count;  // ... and how many
};

__global__ void __launch_bounds__(128) train_generate_kernel(const GenArgs a) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.n_rays) return;
	const uint32_t img = image_idx(i, a.n_rays, a.n_rays_total, a.n_images);
	const ImageDev im = a.images[img];
	Pcg32 rng{a.rng_state, a.rng_inc};
	rng.advance((uint64_t)(uint32_t)(i * 8u));  // N_MAX_RANDOM_SAMPLES_PER_RAY
	float x, y;
	uint32_t px, py;
	image_pos_training(rng, im, a.snap != 0, x, y, px, py);
	if (read_rgba(im, px, py).x < 0.0f) return;  // masked-away pixel
	rng.next_float();                             // motion-blur time (no rolling shutter: unused, but it advances the stream)
	float ox, oy, oz, ux, uy, uz;
	if (im.rays) {
		const float* r = im.rays + ((size_t)px + (size_t)py * im.res[0]) * 6;
		ox = r[0]; oy = r[1]; oz = r[2]; ux = r[3]; uy = r[4]; uz = r[5];
	} else {
		const float* m = im.xform;
		const float lx = (x - im.pp[0]) * (float)im.res[0] / im.focal[0], ly = (y - im.pp[1]) * (float)im.res[1] / im.focal[1];
		ox = m[9]; oy = m[10]; oz = m[11];
		ux = fmaf(m[0], lx, fmaf(m[3], ly, m[6]));
		uy = fmaf(m[1], lx, fmaf(m[4], ly, m[7]));
		uz = fmaf(m[2], lx, fmaf(m[5], ly, m[8]));
	}
	float dx = ux, dy = uy, dz = uz;
	{
		const float n2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
		if (n2 > 0.f) {
			const float n = sqrtf(n2);
			dx /= n; dy /= n; dz /= n;
		}
	}
	const float tmin = fmaxf(aabb_ray_tmin(a.aabb, ox, oy, oz, dx, dy, dz), 0.0f);
	const float startt = fmaf(calc_dt(tmin, a.cone_angle), rng.next_float(), tmin);
	const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;

	// first march: number of occupied samples; the ray parameters of the first kGenKeep of them are kept (shared memory, [step][thread]) so
	// that the emit loop below re-marches only the rays that are longer than that
	__shared__ float t_keep[kGenKeep * 128];
	float t_resume = 0.f;
	uint32_t j = 0;
	float t = startt;
	while (j < kNerfSteps) {
		const float qx = fmaf(t, dx, ox), qy = fmaf(t, dy, oy), qz = fmaf(t, dz, oz);
		if (!aabb_contains(a.aabb, qx, qy, qz)) break;
		const float dt = calc_dt(t, a.cone_angle);
		const uint32_t mip = (uint32_t)mip_from_dt(dt, qx, qy, qz);
		if (occupied_at(qx, qy, qz, a.bitfield, mip)) {
			if (j < kGenKeep) t_keep[j * 128u + threadIdx.x] = t;
			++j;
			t += dt;
			if (j == kGenKeep) t_resume = t;
		} else {
			t = advance_to_next_voxel(t, a.cone_angle, qx, qy, qz, dx, dy, dz, ix, iy, iz, kGrid >> mip);
		}
	}
	if (j == 0) return;
	const uint32_t numsteps = j;
	const uint32_t base = atomicAdd(a.numsteps_counter, numsteps);
	if (base + numsteps > a.max_samples) {
		// dropped (testbed_nerf.cu:1248-1251).  The part of the span below the capacity stays a hole: mark it so that the emit pass skips it
		for (uint32_t k = base; k < a.max_samples; ++k) reinterpret_cast<uint32_t*>(a.coords + (size_t)k * a.stride)[1] = 0xffffffffu;
		return;
	}
	const uint32_t ray_idx = atomicAdd(a.ray_counter, 1u);
	a.ray_indices[ray_idx] = i;
	float* ro = a.rays + (size_t)ray_idx * 6;
	ro[0] = ox; ro[1] = oy; ro[2] = oz; ro[3] = ux; ro[4] = uy; ro[5] = uz;
	a.numsteps[ray_idx * 2 + 0] = numsteps;
	a.numsteps[ray_idx * 2 + 1] = base;

	// second march: only the ray parameter t and the ray's slot per sample (8 bytes); train_emit_kernel expands them into records with
	// coalesced stores.  (Writing the 27 floats from here costs 27 store instructions per sample, each touching 32 different lines.)
	float* rec = a.coords + (size_t)base * a.stride;
	const uint32_t kept = min(numsteps, kGenKeep);
	for (j = 0; j < kept; ++j) {
		rec[0] = t_keep[j * 128u + threadIdx.x];
		reinterpret_cast<uint32_t*>(rec)[1] = ray_idx;
		rec += a.stride;
	}
	t = t_resume;
	while (j < numsteps) {
		const float qx = fmaf(t, dx, ox), qy = fmaf(t, dy, oy), qz = fmaf(t, dz, oz);
		if (!aabb_contains(a.aabb, qx, qy, qz)) break;
		const float dt = calc_dt(t, a.cone_angle);
		const uint32_t mip = (uint32_t)mip_from_dt(dt, qx, qy, qz);
		if (occupied_at(qx, qy, qz, a.bitfield, mip)) {
			rec[0] = t;
			reinterpret_cast<uint32_t*>(rec)[1] = ray_idx;
			rec += a.stride;
			++j;
			t += dt;
		} else {
			t = advance_to_next_voxel(t, a.cone_angle, qx, qy, qz, dx, dy, dz, ix, iy, iz, kGrid >> mip);
		}
	}
}

// Expands (t, ray slot) into the NerfCoordinate records (testbed_nerf.cu:1262-1272).  A warp owns 32 consecutive records: every lane
// computes the fields of one sample, the 32 x stride floats leave through shared memory as full-line stores.
// INDEXED: only the records a round of insta_train_nerf is about to evaluate (emit_index / emit_count: the list train_round_kernel wrote).  The loss kernel, the
// compaction and the backward pass never read a sample behind the one where a ray became opaque, and with a trained model that is most of them: expanding
// all 2.6 M generated samples cost 280 MB of record writes per step for 0.9 M that are read.
constexpr uint32_t kEmitWarps = 4;
template <bool INDEXED>
__global__ void __launch_bounds__(kEmitWarps * 32) train_emit_kernel(const GenArgs a) {
	extern __shared__ float emit_smem[];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	float* tile = emit_smem + (size_t)warp * 32u * a.stride;
	const uint32_t n = INDEXED ? min(*a.emit_count, a.max_samples) : min(*a.numsteps_counter, a.max_samples);
	const uint32_t first = (blockIdx.x * kEmitWarps + warp) * 32u;
	if (first >= n) return;
	const uint32_t cnt = min(32u, n - first);
	float* rec0 = a.coords + (size_t)first * a.stride;
	const uint32_t my_id = INDEXED ? (lane < cnt ? a.emit_index[first + lane] : 0u) : first + lane;
	bool valid = false;
	if (lane < cnt) {
		const float* rec = a.coords + (size_t)my_id * a.stride;
		const float t = rec[0];
		const uint32_t ray_idx = reinterpret_cast<const uint32_t*>(rec)[1];
		valid = ray_idx != 0xffffffffu;
		if (valid) {
			const float* ro = a.rays + (size_t)ray_idx * 6;
			const float ox = ro[0], oy = ro[1], oz = ro[2];
			float dx = ro[3], dy = ro[4], dz = ro[5];
			{
				const float n2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
				if (n2 > 0.f) {
					const float nn = sqrtf(n2);
					dx /= nn; dy /= nn; dz /= nn;
				}
			}
			const uint32_t img = image_idx(a.ray_indices[ray_idx], a.n_rays, a.n_rays_total, a.n_images);
			const float qx = fmaf(t, dx, ox), qy = fmaf(t, dy, oy), qz = fmaf(t, dz, oz);
			const float dt = calc_dt(t, a.cone_angle);
			const float max_step = min_stepsize() * (1 << (kCascades - 1));
			const float wx = (qx - a.aabb[0]) / (a.aabb[3] - a.aabb[0]), wy = (qy - a.aabb[1]) / (a.aabb[4] - a.aabb[1]), wz = (qz - a.aabb[2]) / (a.aabb[5] - a.aabb[2]);
			float* o = tile + (size_t)lane * a.stride;
			o[0] = wx; o[1] = wy; o[2] = wz;
			o[3] = (dt - min_stepsize()) / (max_step - min_stepsize());
			o[4] = (float)img;
			o[5] = wx; o[6] = wy; o[7] = wz;
			o[8] = (dx + 1.0f) * 0.5f; o[9] = (dy + 1.0f) * 0.5f; o[10] = (dz + 1.0f) * 0.5f;
			const float* extra = a.extra_dims ? a.extra_dims + (size_t)img * a.n_extra : nullptr;
			for (uint32_t e = 0; e < a.n_extra; ++e) o[11 + e] = extra ? __ldg(extra + e) : 0.0f;
			for (uint32_t e = 11 + a.n_extra; e < a.stride; ++e) o[e] = rec[e];  // floats past the extra dims are not the generator's
		}
	}
	const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
	__syncwarp();
	if (INDEXED) {
		// the list holds runs of a ray's consecutive samples, not one contiguous range: one record (stride contiguous floats) per store instruction
		for (uint32_t r = 0; r < cnt; ++r) {
			const uint32_t id = __shfl_sync(0xffffffffu, my_id, r);
			if ((vmask >> r) & 1u)
				for (uint32_t k = lane; k < a.stride; k += 32) a.coords[(size_t)id * a.stride + k] = tile[(size_t)r * a.stride + k];
		}
		return;
	}
	if (vmask == (cnt == 32u ? 0xffffffffu : ((1u << cnt) - 1u))) {
		for (uint32_t k = lane; k < cnt * a.stride; k += 32) rec0[k] = tile[k];
	} else {
		// a hole of a dropped ray inside this tile (at most one per step): per-record copies of the valid ones
		for (uint32_t r = 0; r < cnt; ++r)
			if ((vmask >> r) & 1u)
				for (uint32_t k = lane; k < a.stride; k += 32) rec0[(size_t)r * a.stride + k] = tile[(size_t)r * a.stride + k];
	}
}

// ---------------------------------------------------------------------------------------------------------------------
struct LossArgs {
	uint32_t n_rays, n_rays_total, max_samples_compacted, n_images, stride, out_width;
	uint64_t rng_state, rng_inc;
	float aabb[6];
	float loss_scale;
	int snap, random_bg, penalize_density;
	float bg[3];
	float depth_lambda, seg_lambda, near_distance, beta;
	const ImageDev* images;
	const uint32_t* ray_counter;
	const __half* net_out;       // out_width halves per sample (r, g, b, sigma first)
	uint32_t* compacted_counter;
	const uint32_t* ray_indices;
	const float* rays;
	uint32_t* numsteps;          // in: (count, base); out: (compacted count, compacted base)
	const float* coords;
	float* coords_out;
	__half* dloss;               // out_width halves per compacted sample
	float* loss_out;             // per ray, nullable
	const float* mean_density;   // density_grid_mean
};

__device__ __forceinline__ float logistic(float x) { return 1.0f / (1.0f + expf(-x)); }
// huber_loss(target, prediction, 0.1) / 5   (testbed_nerf.cu:139-155, 1299)
__device__ __forceinline__ void huber5(float target, float pred, float& loss, float& grad) {
	const float alpha = 0.1f, d = pred - target, ad = fabsf(d);
	loss = (ad > alpha ? (ad - 0.5f * alpha) : (0.5f / alpha * d * d)) / 5.0f;
	grad = (ad > alpha ? (d > 0 ? 1.0f : -1.0f) : (d / alpha)) / 5.0f;
}

// One WARP per ray (the reference runs one thread per ray, so a step of 10^4..10^5 rays leaves most of the chip idle and every thread
// walks its samples through a chain of dependent, 108-byte-strided loads).  Lanes take consecutive samples of the ray: network outputs and
// record fields are loaded and activated in parallel (coalesced), then the order-dependent recurrences -- transmittance, colour / depth
// accumulation, the early exit at T < 1e-4 -- run as a warp-uniform loop over the 32 staged samples with exactly the reference's
// operation order, so the compacted step count and every rounding are those of the sequential loop.  The compacted records are copied
// cooperatively.
__global__ void __launch_bounds__(128) train_loss_kernel(const LossArgs a) {
	constexpr uint32_t FULL = 0xffffffffu;
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
	if (i >= *a.ray_counter) return;
	const uint32_t numsteps = a.numsteps[i * 2 + 0];
	const uint32_t base = a.numsteps[i * 2 + 1];
	const float* cin = a.coords + (size_t)base * a.stride;
	const __half* nout = a.net_out + (size_t)base * a.out_width;
	const float* ray = a.rays + (size_t)i * 6;
	const float rox = ray[0], roy = ray[1], roz = ray[2];
	const float sx = a.aabb[3] - a.aabb[0], sy = a.aabb[4] - a.aabb[1], sz = a.aabb[5] - a.aabb[2];
	const float max_step = min_stepsize() * (1 << (kCascades - 1));

	// per-sample quantities of sample j (one lane each)
	auto sample = [&](uint32_t j, float& r, float& g, float& b, float& o3, float& density, float& dt, float& depth) {
		const uint2 raw = *reinterpret_cast<const uint2*>(nout + (size_t)j * a.out_width);
		const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
		o3 = hi.y;
		r = logistic(lo.x); g = logistic(lo.y); b = logistic(hi.x);
		const float* c = cin + (size_t)j * a.stride;
		const float px = fmaf(c[5], sx, a.aabb[0]), py = fmaf(c[6], sy, a.aabb[1]), pz = fmaf(c[7], sz, a.aabb[2]);  // deformed_pos
		dt = fmaf(c[3], max_step - min_stepsize(), min_stepsize());
		const float ex = px - rox, ey = py - roy, ez = pz - roz;
		depth = sqrtf(ex * ex + (ey * ey + ez * ez));
		density = __expf(o3);
	};

	// forward composite (1379-1401)
	float T = 1.f;
	const float EPSILON = 1e-4f;
	float rr = 0.f, rg = 0.f, rb = 0.f, depth_ray = 0.f;
	uint32_t compacted = 0;
	bool stop = false;
	for (uint32_t j0 = 0; j0 < numsteps && !stop; j0 += 32) {
		const uint32_t cnt = min(32u, numsteps - j0);
		float r = 0.f, g = 0.f, b = 0.f, o3, density, dt, depth = 0.f, alpha = 0.f;
		if (lane < cnt) {
			sample(j0 + lane, r, g, b, o3, density, dt, depth);
			alpha = 1.f - __expf(-density * dt);
		}
		for (uint32_t k = 0; k < cnt; ++k) {
			if (T < EPSILON) {
				stop = true;
				break;
			}
			const float ak = __shfl_sync(FULL, alpha, k), rk = __shfl_sync(FULL, r, k), gk = __shfl_sync(FULL, g, k), bk = __shfl_sync(FULL, b, k),
			            dk = __shfl_sync(FULL, depth, k);
			const float w = ak * T;
			rr += w * rk; rg += w * gk; rb += w * bk;
			depth_ray += w * dk;
			T *= (1.f - ak);
			++compacted;
		}
	}
	const float T_ray = T;

	// target colour: same pcg32 stream as the generator (every lane computes the same values)
	const uint32_t ray_idx = a.ray_indices[i];
	Pcg32 rng{a.rng_state, a.rng_inc};
	rng.advance((uint64_t)(uint32_t)(ray_idx * 8u));
	const uint32_t img = image_idx(ray_idx, a.n_rays, a.n_rays_total, a.n_images);
	const ImageDev im = a.images[img];
	float x, y;
	uint32_t px, py;
	image_pos_training(rng, im, a.snap != 0, x, y, px, py);
	float bgx = a.bg[0], bgy = a.bg[1], bgz = a.bg[2];
	if (a.random_bg) {
		bgx = rng.next_float(); bgy = rng.next_float(); bgz = rng.next_float();
	}
	// srgb_to_linear then (SRGB colour space, not training in linear colours) linear_to_srgb again
	bgx = linear_to_srgb(srgb_to_linear(bgx)); bgy = linear_to_srgb(srgb_to_linear(bgy)); bgz = linear_to_srgb(srgb_to_linear(bgz));
	const float4 tex = read_rgba(im, px, py);
	float tr, tg, tb;
	if (tex.w > 0) {
		tr = linear_to_srgb(tex.x / tex.w) * tex.w + (1.0f - tex.w) * bgx;
		tg = linear_to_srgb(tex.y / tex.w) * tex.w + (1.0f - tex.w) * bgy;
		tb = linear_to_srgb(tex.z / tex.w) * tex.w + (1.0f - tex.w) * bgz;
	} else {
		tr = bgx; tg = bgy; tb = bgz;
	}
	if (compacted == numsteps) {
		rr += T * bgx; rg += T * bgy; rb += T * bgz;
	}

	uint32_t cbase = 0;
	if (lane == 0) cbase = atomicAdd(a.compacted_counter, compacted);
	cbase = __shfl_sync(FULL, cbase, 0);
	compacted = min(a.max_samples_compacted - min(a.max_samples_compacted, cbase), compacted);
	if (lane == 0) {
		a.numsteps[i * 2 + 0] = compacted;
		a.numsteps[i * 2 + 1] = cbase;
	}

	float seg_supervision = 1.f, depth_lambda = a.depth_lambda;
	if (im.seg) {
		const uint32_t raw = reinterpret_cast<const uint32_t*>(im.seg)[(size_t)px + (size_t)py * im.res[0]];
		const int mask_id = (int)(raw & 0xff);
		if (mask_id == 120 || mask_id == 130 || mask_id == 140) seg_supervision = a.seg_lambda;
		if (mask_id <= 90 || mask_id > 250) depth_lambda = 0.f;
		if (tex.w < 0.001f) depth_lambda = 0.f;
	}
	float lr_, lg_, lb_, gr, gg, gb;
	huber5(tr, rr, lr_, gr);
	huber5(tg, rg, lg_, gg);
	huber5(tb, rb, lb_, gb);
	gr *= seg_supervision; gg *= seg_supervision; gb *= seg_supervision;
	// the reference returns before writing the loss of a ray whose samples no longer fit the compacted batch (1468-1470)
	if (compacted == 0) return;
	if (a.loss_out && lane == 0) a.loss_out[i] = ((lr_ + lg_ + lb_) / 3.0f) / (float)a.n_rays;
	float* cout = a.coords_out + (size_t)cbase * a.stride;
	__half* dl = a.dloss + (size_t)cbase * a.out_width;

	const float dnorm = sqrtf(ray[3] * ray[3] + (ray[4] * ray[4] + ray[5] * ray[5]));
	const float target_depth = dnorm * ((depth_lambda > 0.0f && im.depth) ? im.depth[(size_t)px + (size_t)py * im.res[0]] : -1.0f);
	// L1 depth loss gradient: sign(prediction - target)
	float depth_grad = target_depth > 0.0f ? copysignf(1.0f, depth_ray - target_depth) : 0.f;
	depth_grad *= depth_lambda;

	const float loss_scale = a.loss_scale / (float)a.n_rays;
	const float l1_reg_density = *a.mean_density < 0.01f ? 1e-4f : 0.0f;  // NERF_MIN_OPTICAL_THICKNESS

	// backward sweep (1547-1602)
	float r2 = 0.f, g2 = 0.f, b2 = 0.f, depth_ray2 = 0.f;
	T = 1.f;
	for (uint32_t j0 = 0; j0 < compacted; j0 += 32) {
		const uint32_t cnt = min(32u, compacted - j0);
		// the records of the chunk, copied as one contiguous run of floats
		{
			const float* src = cin + (size_t)j0 * a.stride;
			float* dst = cout + (size_t)j0 * a.stride;
			for (uint32_t k = lane; k < cnt * a.stride; k += 32) dst[k] = src[k];
		}
		float r = 0.f, g = 0.f, b = 0.f, o3 = 0.f, density = 0.f, dt = 0.f, depth = 0.f, alpha = 0.f;
		if (lane < cnt) {
			sample(j0 + lane, r, g, b, o3, density, dt, depth);
			alpha = 1.f - __expf(-density * dt);
		}
		// prefix state of THIS lane's sample, taken from the warp-uniform sequential recurrence
		float my_w = 0.f, my_T = 0.f, my_r2 = 0.f, my_g2 = 0.f, my_b2 = 0.f, my_d2 = 0.f;
		for (uint32_t k = 0; k < cnt; ++k) {
			const float ak = __shfl_sync(FULL, alpha, k), rk = __shfl_sync(FULL, r, k), gk = __shfl_sync(FULL, g, k), bk = __shfl_sync(FULL, b, k),
			            dk = __shfl_sync(FULL, depth, k);
			const float w = ak * T;
			r2 += w * rk; g2 += w * gk; b2 += w * bk;
			depth_ray2 += w * dk;
			T *= (1.f - ak);
			if (lane == k) {
				my_w = w; my_T = T; my_r2 = r2; my_g2 = g2; my_b2 = b2; my_d2 = depth_ray2;
			}
		}
		if (lane < cnt) {
			const float w = my_w, Tj = my_T;
			const float sfx_r = rr - my_r2, sfx_g = rg - my_g2, sfx_b = rb - my_b2;
			// logistic derivative s (1 - s); output_l2_reg is 0 for logistic colours
			const float d0 = loss_scale * (w * gr * (r * (1 - r))), d1 = loss_scale * (w * gg * (g * (1 - g))), d2 = loss_scale * (w * gb * (b * (1 - b)));
			const float density_derivative = __expf(tclamp(o3, -15.0f, 15.0f));
			const float depth_suffix = depth_ray - my_d2;
			const float depth_supervision = depth_grad * (Tj * depth - depth_suffix);
			const float beta_supervision = dt / (T_ray - 1.f + 1e-3f);
			float dmlp = density_derivative * (dt * (gr * (Tj * r - sfx_r) + (gg * (Tj * g - sfx_g) + gb * (Tj * b - sfx_b)) + depth_supervision) + a.beta * beta_supervision);
			if (a.penalize_density) dmlp += (tex.w <= 0.01f) ? w : 0.f;
			const float d3 = loss_scale * dmlp + (o3 < 0.0f ? -l1_reg_density : 0.0f) + (o3 > -10.0f && depth < a.near_distance ? 1e-4f : 0.0f);
			const __half2 h01 = __floats2half2_rn(d0, d1), h23 = __floats2half2_rn(d2, d3);
			const size_t j = j0 + lane;
			*reinterpret_cast<uint2*>(dl + j * a.out_width) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
			for (uint32_t k = 4; k < a.out_width; ++k) dl[j * a.out_width + k] = __float2half(0.f);
		}
	}
}

// ---- lazily evaluated rounds ------------------------------------------------------------------------------------------------------------------
// The reference deforms and evaluates EVERY generated sample before the loss kernel walks each ray front to back and stops at T < 1e-4
// (testbed_nerf.cu:3290-3292, 1379-1401): nothing behind that sample is ever read -- not by the loss, not by the compaction, not by the backward pass.
// With a trained model that is most of them (bench workload: 2.6 M generated, 0.27 M used).  The trainer therefore evaluates a ray's samples in rounds
// ([0, 8), [8, 32), [32, end)): after each round this kernel replays the loss kernel's transmittance recurrence over the samples evaluated so far --
// same expressions, same order, hence the same bits and the same stopping sample -- and only rays that are still transparent contribute their next
// chunk to the index list the deformation and network kernels of the next round run over.  Everything the loss kernel reads has been evaluated by then.
struct RoundArgs {
	uint32_t stride, out_width;
	uint32_t prev_lo, prev_hi, lo, hi;   // sample ranges (per ray) evaluated by the previous round / to be evaluated by this one; prev_hi == 0: first round
	const uint32_t* ray_counter;
	const uint32_t* numsteps;            // (count, base)
	const float* coords;
	const __half* net_out;
	float* ray_T;                        // transmittance after the evaluated samples; < 0: the ray needs nothing more
	uint32_t* index;                     // out: sample indices of this round
	uint32_t* count;                     // out: their number
};

__global__ void __launch_bounds__(128) train_round_kernel(const RoundArgs a) {
	constexpr uint32_t FULL = 0xffffffffu;
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
	if (i >= *a.ray_counter) return;
	const uint32_t numsteps = a.numsteps[i * 2 + 0];
	const uint32_t base = a.numsteps[i * 2 + 1];
	const float EPSILON = 1e-4f;
	const float max_step = min_stepsize() * (1 << (kCascades - 1));
	float T = 1.f;
	if (a.prev_hi > 0) {
		T = a.ray_T[i];
		if (T < 0.f) return;
		const uint32_t end = min(a.prev_hi, numsteps);
		bool stop = false;
		for (uint32_t j0 = a.prev_lo; j0 < end && !stop; j0 += 32) {
			const uint32_t cnt = min(32u, end - j0);
			float alpha = 0.f;
			if (lane < cnt) {
				// exactly train_loss_kernel's `sample`: o3 = fourth half of the sample's output, density = __expf(o3), alpha = 1 - __expf(-density dt)
				const size_t j = (size_t)base + j0 + lane;
				const uint2 raw = *reinterpret_cast<const uint2*>(a.net_out + j * a.out_width);
				const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
				const float o3 = hi.y;
				const float dt = fmaf(a.coords[j * a.stride + 3], max_step - min_stepsize(), min_stepsize());
				const float density = __expf(o3);
				alpha = 1.f - __expf(-density * dt);
			}
			for (uint32_t k = 0; k < cnt; ++k) {
				if (T < EPSILON) {
					stop = true;
					break;
				}
				const float ak = __shfl_sync(FULL, alpha, k);
				T *= (1.f - ak);
			}
		}
		// the loss kernel tests T before it touches the next sample
		if (stop || T < EPSILON || end >= numsteps) {
			if (lane == 0) a.ray_T[i] = -1.f;
			return;
		}
	} else if (numsteps == 0) {
		if (lane == 0) a.ray_T[i] = -1.f;
		return;
	}
	const uint32_t n = min(a.hi, numsteps) - a.lo;
	uint32_t pos = 0;
	if (lane == 0) pos = atomicAdd(a.count, n);
	pos = __shfl_sync(FULL, pos, 0);
	for (uint32_t k = lane; k < n; k += 32) a.index[pos + k] = base + a.lo + k;
	if (lane == 0) a.ray_T[i] = T;
}

// fill_rollover(_and_rescale): pad the compacted batch to exactly n_elements samples by repeating it
template <typename T, bool RESCALE>
__global__ void rollover_kernel(uint32_t n_elements, uint32_t stride, const uint32_t* __restrict__ n_input_ptr, T* __restrict__ inout) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t n_in = *n_input_ptr;
	if (i < n_in * stride || i >= n_elements * stride || n_in == 0) return;
	T v = inout[i % (n_in * stride)];
	if (RESCALE) v = (T)((float)v * n_in / n_elements);
	inout[i] = v;
}

// ---------------------------------------------------------------------------------------------------------------------
struct OptArgs {
	uint32_t n, n_matrix;
	float loss_scale, lr, beta1, beta2, eps, l2_reg, ema_decay, ema_debias_old, ema_debias_new;
	float* master;
	__half* weights;
	const __half* grads;
	float* m1;
	float* m2;
	uint32_t* steps;
	__half* ema;
	// adam.h:94-97 evaluates lr * sqrt(1 - beta2^t) / (1 - beta1^t) with two powf per parameter; t is the parameter's own update count.  The same expression
	// for t < table_n comes from a table (adam_lr_table_kernel, filled every step with this step's lr / betas: bit-identical), and from t = t_flat on both
	// powers are below half an ulp of 1, so the expression IS lr.  A step count in between (only with betas much closer to 1 than the shipped 0.9 / 0.99)
	// takes the two powf.  With ~10 % of the grid entries touched per step every warp ran the powf path for a few lanes: 0.19 ms of a 2.2 ms step.
	const float* lr_table;
	uint32_t table_n, t_flat;
};
constexpr uint32_t kAdamLrTable = 2048;

__device__ __forceinline__ float adam_lr_expr(float lr, float beta1, float beta2, uint32_t step) {
	return lr * (sqrtf(1 - powf(beta2, (float)step)) / (1 - powf(beta1, (float)step)));
}
__global__ void adam_lr_table_kernel(float lr, float beta1, float beta2, uint32_t n, float* __restrict__ out) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) out[t] = adam_lr_expr(lr, beta1, beta2, t);
}
__device__ __forceinline__ float adam_lr(const OptArgs& a, uint32_t step) {
	if (step < a.table_n) return __ldg(a.lr_table + step);
	if (step >= a.t_flat) return a.lr;
	return adam_lr_expr(a.lr, a.beta1, a.beta2, step);
}

__global__ void __launch_bounds__(256) adam_ema_kernel(const OptArgs a) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.n) return;
	float g = __half2float(a.grads[i]) / a.loss_scale;
	__half w16 = a.weights[i];
	const bool matrix = i < a.n_matrix;
	if (matrix || g != 0.f) {
		const float w = a.master[i];
		if (matrix) g += a.l2_reg * w;
		const float m1 = a.m1[i] = a.beta1 * a.m1[i] + (1 - a.beta1) * g;
		const float m2 = a.m2[i] = a.beta2 * a.m2[i] + (1 - a.beta2) * (g * g);
		const uint32_t step = ++a.steps[i];
		const float lr = adam_lr(a, step);
		const float eff = fminf(fmaxf(lr / (sqrtf(m2) + a.eps), 0.0f), FLT_MAX);
		const float nw = w - eff * m1;  // weight decay terms are zero in the shipped config
		a.master[i] = nw;
		w16 = __float2half(nw);
		a.weights[i] = w16;
	}
	const float f = (__half2float(a.ema[i]) * a.ema_decay * a.ema_debias_old + __half2float(w16) * (1 - a.ema_decay)) * a.ema_debias_new;
	a.ema[i] = __float2half(f);
}

// Eight parameters per thread (one 128-bit load each of the fp16 gradients / weights / EMA weights).  Most hash-grid entries are not touched
// in a step: their group costs 48 bytes of traffic in three vector accesses + one vector store instead of 8 x 4 two-byte accesses; the
// fp32 state of a group is only read when one of its gradients is non-zero (or for the MLP weights, which always take the l2 term).
__global__ void __launch_bounds__(256) adam_ema_vec8_kernel(const OptArgs a) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t i0 = v * 8u;
	if (i0 >= a.n) return;
	const uint4 gq = __ldg(reinterpret_cast<const uint4*>(a.grads) + v);
	uint4 wq = reinterpret_cast<const uint4*>(a.weights)[v];
	uint4 eq = reinterpret_cast<const uint4*>(a.ema)[v];
	const bool matrix = i0 < a.n_matrix;
	__half g16[8], w16[8], e16[8];
	*reinterpret_cast<uint4*>(g16) = gq;
	*reinterpret_cast<uint4*>(w16) = wq;
	*reinterpret_cast<uint4*>(e16) = eq;
	if (matrix || ((gq.x | gq.y | gq.z | gq.w) & 0x7fff7fffu) != 0u) {
		float4* mp = reinterpret_cast<float4*>(a.master) + 2 * (size_t)v;
		float4* p1 = reinterpret_cast<float4*>(a.m1) + 2 * (size_t)v;
		float4* p2 = reinterpret_cast<float4*>(a.m2) + 2 * (size_t)v;
		uint4* ps = reinterpret_cast<uint4*>(a.steps) + 2 * (size_t)v;
		float w[8], m1[8], m2[8];
		uint32_t st[8];
		*reinterpret_cast<float4*>(w) = mp[0]; *reinterpret_cast<float4*>(w + 4) = mp[1];
		*reinterpret_cast<float4*>(m1) = p1[0]; *reinterpret_cast<float4*>(m1 + 4) = p1[1];
		*reinterpret_cast<float4*>(m2) = p2[0]; *reinterpret_cast<float4*>(m2 + 4) = p2[1];
		*reinterpret_cast<uint4*>(st) = ps[0]; *reinterpret_cast<uint4*>(st + 4) = ps[1];
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			float g = __half2float(g16[k]) / a.loss_scale;
			if (matrix || g != 0.f) {
				if (matrix) g += a.l2_reg * w[k];
				m1[k] = a.beta1 * m1[k] + (1 - a.beta1) * g;
				m2[k] = a.beta2 * m2[k] + (1 - a.beta2) * (g * g);
				const uint32_t step = ++st[k];
				const float lr = adam_lr(a, step);
				const float eff = fminf(fmaxf(lr / (sqrtf(m2[k]) + a.eps), 0.0f), FLT_MAX);
				w[k] = w[k] - eff * m1[k];
				w16[k] = __float2half(w[k]);
			}
		}
		mp[0] = *reinterpret_cast<float4*>(w); mp[1] = *reinterpret_cast<float4*>(w + 4);
		p1[0] = *reinterpret_cast<float4*>(m1); p1[1] = *reinterpret_cast<float4*>(m1 + 4);
		p2[0] = *reinterpret_cast<float4*>(m2); p2[1] = *reinterpret_cast<float4*>(m2 + 4);
		ps[0] = *reinterpret_cast<uint4*>(st); ps[1] = *reinterpret_cast<uint4*>(st + 4);
		reinterpret_cast<uint4*>(a.weights)[v] = *reinterpret_cast<uint4*>(w16);
	}
#pragma unroll
	for (int k = 0; k < 8; ++k)
		e16[k] = __float2half((__half2float(e16[k]) * a.ema_decay * a.ema_debias_old + __half2float(w16[k]) * (1 - a.ema_decay)) * a.ema_debias_new);
	reinterpret_cast<uint4*>(a.ema)[v] = *reinterpret_cast<uint4*>(e16);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
int training_set_images(insta_ctx* ctx, const insta_train_image* h, uint32_t n) {
	if (ctx->d_train_images) {
		INSTA_CUDA(cudaDeviceSynchronize());
		cudaFree(ctx->d_train_images);
		ctx->d_train_images = nullptr;
	}
	ctx->n_train_images = 0;
	if (n == 0) return INSTA_OK;
	std::vector<ImageDev> host(n);
	for (uint32_t k = 0; k < n; ++k) {
		ImageDev& d = host[k];
		d.pixels = h[k].d_pixels;
		d.seg = h[k].d_seg_mask;
		d.depth = h[k].d_depth;
		d.rays = h[k].d_rays;
		for (int c = 0; c < 2; ++c) {
			d.res[c] = h[k].resolution[c];
			d.focal[c] = h[k].focal_length[c];
			d.pp[c] = h[k].principal_point[c];
		}
		for (int c = 0; c < 12; ++c) d.xform[c] = h[k].xform[c];
		d.type = h[k].image_data_type;
		d.pad = 0;
		if (!d.pixels || d.res[0] == 0 || d.res[1] == 0 || d.type < 1 || d.type > 3) INSTA_FAIL(INSTA_ERR_INVALID, "insta_training_set_images: bad image " + std::to_string(k));
	}
	INSTA_CUDA(cudaMalloc(&ctx->d_train_images, sizeof(ImageDev) * n));
	INSTA_CUDA(cudaMemcpy(ctx->d_train_images, host.data(), sizeof(ImageDev) * n, cudaMemcpyHostToDevice));
	ctx->n_train_images = n;
	return INSTA_OK;
}

void pcg_seed(uint64_t seed, uint64_t advance_steps, uint64_t& state, uint64_t& inc) {
	// pcg32(seed): state = 0; inc = 3; next; state += seed; next   (tcnn pcg32.h:57-63), then `advance_steps` x advance()
	const uint64_t mult = 0x5851f42d4c957f2dULL;
	inc = 3;
	state = 0 * mult + inc;
	state += seed;
	state = state * mult + inc;
	for (uint64_t k = 0; k < advance_steps; ++k) {
		// advance(1 << 32)
		uint64_t cur_mult = mult, cur_plus = inc, acc_mult = 1, acc_plus = 0, delta = 1ull << 32;
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
}

static GenArgs gen_args(insta_ctx* ctx, const insta_train_step* t) {
	GenArgs a{};
	a.n_rays = t->n_rays;
	a.n_rays_total = t->n_rays_total;
	a.max_samples = t->max_samples;
	a.n_images = ctx->n_train_images;
	a.stride = t->stride_floats;
	a.n_extra = ctx->p.exp_params + ctx->p.geo_params;
	pcg_seed(t->rng_seed, t->rng_advances, a.rng_state, a.rng_inc);
	for (int k = 0; k < 6; ++k) a.aabb[k] = ctx->aabb[k];
	a.cone_angle = t->cone_angle_constant;
	a.snap = (int)t->snap_to_pixel_centers;
	a.images = (const ImageDev*)ctx->d_train_images;
	a.bitfield = (const uint8_t*)t->d_density_grid_bitfield;
	a.extra_dims = t->d_extra_dims;
	a.coords = t->d_coords;
	a.ray_counter = t->d_counters;
	a.numsteps_counter = t->d_counters + 1;
	a.ray_indices = t->d_ray_indices;
	a.rays = t->d_rays;
	a.numsteps = t->d_numsteps;
	return a;
}

// emit = false: the records stay (ray parameter, ray slot) pairs; insta_train_nerf expands them round by round (train_emit_round)
int train_generate(insta_ctx* ctx, const insta_train_step* t, cudaStream_t s, bool emit) {
	const GenArgs a = gen_args(ctx, t);
	INSTA_CUDA(cudaMemsetAsync(t->d_counters, 0, 3 * sizeof(uint32_t), s));
	if (a.n_rays == 0) return INSTA_OK;
	train_generate_kernel<<<(a.n_rays + 127) / 128, 128, 0, s>>>(a);
	ctx->launches++;
	if (emit) {
		// the sample count is only known on the device: the emit grid covers the capacity, warps past the count exit at once
		const size_t emit_smem = (size_t)kEmitWarps * 32u * a.stride * sizeof(float);
		if (emit_smem > 48 * 1024) INSTA_CUDA(cudaFuncSetAttribute(train_emit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emit_smem));
		train_emit_kernel<false><<<(a.max_samples + kEmitWarps * 32 - 1) / (kEmitWarps * 32), kEmitWarps * 32, emit_smem, s>>>(a);
		ctx->launches++;
	}
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

// the records of one round's list (d_index[0 .. *d_count), at most `bound`)
int train_emit_round(insta_ctx* ctx, const insta_train_step* t, const uint32_t* d_index, const uint32_t* d_count, uint32_t bound, cudaStream_t s) {
	if (t->n_rays == 0 || bound == 0) return INSTA_OK;
	GenArgs a = gen_args(ctx, t);
	a.emit_index = d_index;
	a.emit_count = d_count;
	a.max_samples = bound;
	const size_t emit_smem = (size_t)kEmitWarps * 32u * a.stride * sizeof(float);
	if (emit_smem > 48 * 1024) INSTA_CUDA(cudaFuncSetAttribute(train_emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emit_smem));
	train_emit_kernel<true><<<(bound + kEmitWarps * 32 - 1) / (kEmitWarps * 32), kEmitWarps * 32, emit_smem, s>>>(a);
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

int train_round(insta_ctx* ctx, const insta_train_step* t, const void* d_net_out, uint32_t out_width, uint32_t prev_lo, uint32_t prev_hi, uint32_t lo, uint32_t hi,
                float* d_ray_T, uint32_t* d_index, uint32_t* d_count, cudaStream_t s) {
	if (t->n_rays == 0) return INSTA_OK;
	RoundArgs a{};
	a.stride = t->stride_floats;
	a.out_width = out_width;
	a.prev_lo = prev_lo; a.prev_hi = prev_hi; a.lo = lo; a.hi = hi;
	a.ray_counter = t->d_counters;
	a.numsteps = t->d_numsteps;
	a.coords = t->d_coords;
	a.net_out = (const __half*)d_net_out;
	a.ray_T = d_ray_T;
	a.index = d_index;
	a.count = d_count;
	train_round_kernel<<<(t->n_rays + 3) / 4, 128, 0, s>>>(a);  // one warp per ray
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

int train_loss(insta_ctx* ctx, const insta_train_step* t, const void* d_net_out, uint32_t out_width, float* d_coords_compacted, void* d_dloss, float* d_loss,
               const float* d_mean_density, cudaStream_t s) {
	LossArgs a{};
	a.n_rays = t->n_rays;
	a.n_rays_total = t->n_rays_total;
	a.max_samples_compacted = t->target_batch_size;
	a.n_images = ctx->n_train_images;
	a.stride = t->stride_floats;
	a.out_width = out_width;
	pcg_seed(t->rng_seed, t->rng_advances, a.rng_state, a.rng_inc);
	for (int k = 0; k < 6; ++k) a.aabb[k] = ctx->aabb[k];
	a.loss_scale = t->loss_scale;
	a.snap = (int)t->snap_to_pixel_centers;
	a.random_bg = (int)t->random_bg_color;
	a.penalize_density = (int)t->penalize_density_alpha;
	for (int k = 0; k < 3; ++k) a.bg[k] = t->background_color[k];
	a.depth_lambda = t->depth_weight;
	a.seg_lambda = t->seg_weight;
	a.near_distance = t->near_distance;
	a.beta = t->beta_loss;
	a.images = (const ImageDev*)ctx->d_train_images;
	a.ray_counter = t->d_counters;
	a.net_out = (const __half*)d_net_out;
	a.compacted_counter = t->d_counters + 2;
	a.ray_indices = t->d_ray_indices;
	a.rays = t->d_rays;
	a.numsteps = t->d_numsteps;
	a.coords = t->d_coords;
	a.coords_out = d_coords_compacted;
	a.dloss = (__half*)d_dloss;
	a.loss_out = d_loss;
	a.mean_density = d_mean_density;
	if (a.n_rays == 0) return INSTA_OK;
	train_loss_kernel<<<(a.n_rays + 3) / 4, 128, 0, s>>>(a);  // one warp per ray
	ctx->launches++;
	// pad the compacted batch to exactly target_batch_size samples
	const uint32_t nb = t->target_batch_size;
	rollover_kernel<__half, true><<<((size_t)nb * out_width + 127) / 128, 128, 0, s>>>(nb, out_width, a.compacted_counter, (__half*)d_dloss);
	rollover_kernel<float, false><<<((size_t)nb * a.stride + 127) / 128, 128, 0, s>>>(nb, a.stride, a.compacted_counter, d_coords_compacted);
	ctx->launches += 2;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

int optimizer_step(insta_ctx* ctx, const insta_optimizer* o, uint32_t n_params, uint32_t n_matrix, cudaStream_t s) {
	OptArgs a{};
	a.n = n_params;
	a.n_matrix = n_matrix;
	a.loss_scale = o->loss_scale;
	a.lr = o->learning_rate;
	a.beta1 = o->beta1;
	a.beta2 = o->beta2;
	a.eps = o->epsilon;
	a.l2_reg = o->l2_reg;
	a.ema_decay = o->ema_decay;
	// EmaOptimizer::step (ema.h:103-109) with the nested optimizer's step count AFTER this step
	const uint32_t step = o->current_step + 1;
	a.ema_debias_old = 1 - (float)std::pow(o->ema_decay, step - 1);
	a.ema_debias_new = 1.0f / (1 - (float)std::pow(o->ema_decay, step));
	a.master = o->d_params_fp32;
	a.weights = (__half*)o->d_params_f16;
	a.grads = (const __half*)o->d_gradients_f16;
	a.m1 = o->d_first_moments;
	a.m2 = o->d_second_moments;
	a.steps = o->d_param_steps;
	a.ema = (__half*)o->d_params_ema_f16;
	// bias-corrected learning rates of this step (see OptArgs)
	if (!ctx->d_adam_lr_table) INSTA_CUDA(cudaMalloc(&ctx->d_adam_lr_table, sizeof(float) * kAdamLrTable));
	adam_lr_table_kernel<<<kAdamLrTable / 256, 256, 0, s>>>(a.lr, a.beta1, a.beta2, kAdamLrTable, ctx->d_adam_lr_table);
	ctx->launches++;
	a.lr_table = ctx->d_adam_lr_table;
	a.table_n = kAdamLrTable;
	{
		// smallest t from which beta^t < 2^-26 for both betas (then 1 - beta^t rounds to 1 whatever the last bits of powf are); betas outside (0, 1): never
		auto flat_from = [](float beta) -> double { return beta > 0.f && beta < 1.f ? std::ceil(std::log(std::ldexp(1.0, -26)) / std::log((double)beta)) + 2.0 : 4.0e9; };
		const double t = std::max(flat_from(a.beta1), flat_from(a.beta2));
		a.t_flat = t < 4.0e9 ? (uint32_t)t : 0xffffffffu;
	}
	auto al = [](const void* p, uintptr_t m) { return (reinterpret_cast<uintptr_t>(p) & (m - 1)) == 0; };
	if (n_params % 8 == 0 && n_matrix % 8 == 0 && al(a.grads, 16) && al(a.weights, 16) && al(a.ema, 16) && al(a.master, 32) && al(a.m1, 32) && al(a.m2, 32) && al(a.steps, 32))
		adam_ema_vec8_kernel<<<(n_params / 8 + 255) / 256, 256, 0, s>>>(a);
	else
		adam_ema_kernel<<<(n_params + 255) / 256, 256, 0, s>>>(a);
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

}  // namespace insta
