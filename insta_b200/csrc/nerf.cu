// Stage 3 of the INSTA hot path on sm_100a: ray setup, occupancy-grid skipping, sample generation and front-to-back
// compositing of the render path.  Replaces (ngp/src/testbed_nerf.cu):
//   init_rays_with_payload_kernel_nerf 1854-1960 (+ pixel_to_ray, ngp common_device.cuh:260-316, pinhole branch)
//   advance_pos_nerf 634-690, generate_next_nerf_network_inputs 722-795, composite_kernel_nerf 797-1009 (Shade mode),
//   compact_kernel_nerf 1829-1852, shade_kernel_nerf 1794-1827, and the driver NerfTracer::trace 2112-2244.
//
// Design (not a port):
//   * ray state is SoA (two float4 + one uint2 + rgba float4 + depth per ray), ping-ponged between two sets; the 40-byte
//     AoS NerfPayload of the reference would cost 2.5 uncoalesced sectors per field access;
//   * ray setup and the skip to the first occupied cell are one kernel; compositing, stream compaction of the survivors
//     (warp-aggregated atomics: one atomic per warp, not per ray) and the write of finished rays to the frame buffer are
//     one kernel -- the reference runs compact + composite + shade as three passes over the ray state;
//   * the occupancy bitfield (256 KB) is read through the read-only path and stays L1/L2-resident.
//
// Exactness: which cells are skipped and how many steps a ray takes are decided by fp32 comparisons on t and pos, so the
// expressions below follow the reference's evaluation order, with the contractions nvcc applies to the reference build
// written out as explicit fmaf (pos = o + d t is one FFMA per component, etc.).  tests/test_gpu_nerf.py checks ray
// state, per-ray step counts and the generated records bit-for-bit against the reference's own kernels.
#include <cuda_fp16.h>
#include <float.h>

#include <algorithm>
#include <cstring>
#include <cmath>

#include "ctx.h"
#include "nerf_device.cuh"

namespace insta {
namespace {

using namespace nerfdev;

struct RenderArgs {
	insta_camera cam;
	float render_aabb[6];
	float train_aabb[6];
	float cone_angle;
	float min_transmittance;
	const uint8_t* bitfield;
	uint32_t min_mip;
};

struct RaySet {
	float4* o_t;     // origin, t
	float4* d_w;     // dir, max_weight
	uint2* meta;     // pixel idx, n_steps
	float4* rgba;
	float* depth;
};

// one warp-aggregated append: returns the slot of this lane in the list counted by *counter (all lanes must call)
__device__ __forceinline__ uint32_t warp_append(bool want, uint32_t* counter) {
	const uint32_t mask = __ballot_sync(0xffffffffu, want);
	if (mask == 0) return 0;
	const int leader = __ffs(mask) - 1;
	uint32_t base = 0;
	if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(counter, __popc(mask));
	base = __shfl_sync(0xffffffffu, base, leader);
	return base + __popc(mask & ((1u << (threadIdx.x & 31)) - 1u));
}

// init_rays_with_payload_kernel_nerf + advance_pos_nerf.  Every pixel's final state goes to the debug arrays when given
// (parity tests); alive rays are appended to `out`.
__global__ void __launch_bounds__(128) nerf_init_rays_kernel(const RenderArgs a, RaySet out, uint32_t* __restrict__ n_alive, float4* __restrict__ frame_rgba,
                                                             float* __restrict__ frame_depth, float4* __restrict__ dbg_o_t, float4* __restrict__ dbg_d_alive) {
	const uint32_t x = threadIdx.x % 16 + 16 * blockIdx.x, y = threadIdx.x / 16 + 8 * blockIdx.y;
	const uint32_t W = a.cam.resolution[0], H = a.cam.resolution[1];
	const bool inside = x < W && y < H;
	const uint32_t idx = x + W * y;
	bool alive = false;
	float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 1, t = 0;
	if (inside) {
		const float* m = a.cam.camera_matrix;  // column-major 3x4
		// pixel_to_ray, pinhole branch (common_device.cuh:275-305); parallax shift is (0, 0, scale) in INSTA
		const float2 off = ld_random_pixel_offset(a.cam.snap_to_pixel_centers ? 0u : a.cam.sample_index);
		const float u = ((float)x + off.x) / (float)W, v = ((float)y + off.y) / (float)H;
		const float lx = (u - a.cam.screen_center[0]) * (float)W / a.cam.focal_length[0];
		const float ly = (v - a.cam.screen_center[1]) * (float)H / a.cam.focal_length[1];
		// dir = R * (lx, ly, 1) as the reference build evaluates it: fma(col0, x, fma(col1, y, col2 * z))
		float rx = fmaf(m[0], lx, fmaf(m[3], ly, m[6]));
		float ry = fmaf(m[1], lx, fmaf(m[4], ly, m[7]));
		float rz = fmaf(m[2], lx, fmaf(m[5], ly, m[8]));
		ox = m[9]; oy = m[10]; oz = m[11];
		frame_depth[idx] = 1e10f;
		// ray.d.normalized(): d / sqrt(squaredNorm) when the squared norm is positive
		const float n2 = fmaf(rx, rx, fmaf(ry, ry, rz * rz));
		if (n2 > 0.f) {
			const float n = sqrtf(n2);
			rx /= n; ry /= n; rz /= n;
		}
		dx = rx; dy = ry; dz = rz;
		t = fmaxf(aabb_ray_tmin(a.render_aabb, ox, oy, oz, dx, dy, dz), 0.05f) + 1e-6f;  // NERF_RENDERING_NEAR_DISTANCE
		alive = aabb_contains(a.render_aabb, fmaf(dx, t, ox), fmaf(dy, t, oy), fmaf(dz, t, oz));
		if (alive) {
			// advance_pos_nerf: jitter the start, then skip empty cells
			const float idx_ = 1.0f / dx, idy_ = 1.0f / dy, idz_ = 1.0f / dz;
			float dt = calc_dt(t, a.cone_angle);
			t = fmaf(ld_random_val(a.cam.sample_index, idx * 786433u), dt, t);
			while (true) {
				const float px = fmaf(dx, t, ox), py = fmaf(dy, t, oy), pz = fmaf(dz, t, oz);
				if (!aabb_contains(a.render_aabb, px, py, pz)) {
					alive = false;
					break;
				}
				dt = calc_dt(t, a.cone_angle);
				const uint32_t mip = max(a.min_mip, (uint32_t)mip_from_dt(dt, px, py, pz));
				if (!a.bitfield || occupied_at(px, py, pz, a.bitfield, mip)) break;
				t = advance_to_next_voxel(t, a.cone_angle, px, py, pz, dx, dy, dz, idx_, idy_, idz_, kGrid >> mip);
			}
		}
		if (dbg_o_t) {
			dbg_o_t[idx] = make_float4(ox, oy, oz, t);
			dbg_d_alive[idx] = make_float4(dx, dy, dz, alive ? 1.0f : 0.0f);
		}
	}
	const uint32_t slot = warp_append(alive, n_alive);
	if (alive) {
		out.o_t[slot] = make_float4(ox, oy, oz, t);
		out.d_w[slot] = make_float4(dx, dy, dz, 0.0f);
		out.meta[slot] = make_uint2(idx, 0u);
		out.rgba[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
		out.depth[slot] = 0.0f;
	}
}

// generate_next_nerf_network_inputs: up to n_steps occupied samples per alive ray, sample (i + j n_alive) of the record
// buffer; records of steps the ray never reaches keep holder = -1 (so the deformation skips them).
// The 32 records a warp produces for one step are contiguous in the buffer ((i0 + j n_alive) .. + 31): they are assembled in shared
// memory and leave as full-line stores (27 scalar stores per record, each touching 32 different lines, otherwise).
constexpr int kMarchChunk = 16;  // steps marched per lane before the warp writes their records

// Device-side schedule of insta_render_frame: what NerfTracer::trace keeps on the host between two march iterations (testbed_nerf.cu:2156-2172: live-ray
// count read back, steps per iteration derived from it) lives here, so that the iterations of a frame are enqueued without a host round trip.  The
// kernels of an iteration take their sizes from this block when they are handed one (launch grids are sized for the upper bounds).
struct RenderSched {
	uint32_t n_alive;     // rays entering the iteration
	uint32_t n_steps;     // steps marched per ray in it
	uint32_t n_records;   // = n_alive * n_steps: what the deformation and the network process
	uint32_t step;        // NerfTracer's running step counter (composite: current_step)
	uint32_t iterations;  // iterations with work so far
	uint32_t pad;
	unsigned long long samples;
};

// one thread: closes iteration k - 1 (survivor count -> n_alive) and sets up iteration k
__global__ void render_schedule_kernel(RenderSched* __restrict__ sc, const uint32_t* __restrict__ n_survivors, uint32_t* __restrict__ counter_to_zero, uint32_t budget,
                                       uint32_t max_steps, int first, uint32_t tail_budget, uint32_t tail_max_steps) {
	const uint32_t n = *n_survivors;
	const uint32_t prev_steps = first ? 0u : sc->n_steps;
	uint32_t ns = n ? min(max(budget / n, 1u), max_steps) : 0u;
	// the tail of a frame: few rays left -- let them march further, so that the last iterations (each a full chain of launches for a few ten thousand records)
	// collapse into one
	if (n && n * ns < tail_budget) ns = min(max(tail_budget / n, ns), max(tail_max_steps, ns));
	sc->step = first ? 1u : sc->step + prev_steps;
	sc->n_alive = n;
	sc->n_steps = ns;
	sc->n_records = n * ns;
	if (first) {
		sc->iterations = 0;
		sc->samples = 0ull;
	}
	if (n) {
		sc->iterations += 1;
		sc->samples += (unsigned long long)n * ns;
	}
	*counter_to_zero = 0;
}

__global__ void __launch_bounds__(128) nerf_march_kernel(const RenderArgs a, RaySet rays, const uint32_t n_alive_arg, float* __restrict__ records, const uint32_t stride,
                                                         const uint32_t n_steps_arg, const float frame_id, const float* __restrict__ extra_dims, const uint32_t n_extra,
                                                         const RenderSched* __restrict__ sched) {
	extern __shared__ float march_smem[];
	const uint32_t n_alive = sched ? sched->n_alive : n_alive_arg, n_steps = sched ? sched->n_steps : n_steps_arg;
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	float* tile = march_smem + (size_t)warp * 32u * stride;
	const uint32_t i0 = blockIdx.x * blockDim.x + warp * 32u;
	if (i0 >= n_alive) return;
	const uint32_t cnt = min(32u, n_alive - i0);
	const uint32_t i = i0 + lane;
	const bool mine = lane < cnt;
	float ox = 0.f, oy = 0.f, oz = 0.f, dx = 1.f, dy = 1.f, dz = 1.f, t = 0.f;
	if (mine) {
		const float4 o_t = rays.o_t[i], d_w = rays.d_w[i];
		ox = o_t.x; oy = o_t.y; oz = o_t.z; dx = d_w.x; dy = d_w.y; dz = d_w.z;
		t = o_t.w;
	}
	const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
	const float max_step = min_stepsize() * (1 << (kCascades - 1));
	float* my = tile + (size_t)lane * stride;
	for (uint32_t e = 0; e < stride; ++e) my[e] = 0.0f;
	if (mine) {
		my[8] = (dx + 1.0f) * 0.5f; my[9] = (dy + 1.0f) * 0.5f; my[10] = (dz + 1.0f) * 0.5f;
		for (uint32_t e = 0; e < n_extra; ++e) my[11 + e] = extra_dims ? extra_dims[e] : 0.0f;
	}
	uint32_t steps = 0;   // samples this ray generated in this round
	bool left = !mine;    // steps after the one where the ray left the box are never composited: holder = -1 marks them as padding
	// Phase 1 (per lane, no warp synchronisation): the ray's next <= kMarchChunk samples.  ONE loop whose iteration is one probe of the occupancy
	// grid -- it either emits a sample or skips to the next voxel -- so lanes that are at different steps still execute the same instructions
	// together and the warp takes as long as its longest ray.  (A loop over steps with the skip loop inside made every step wait for the one lane
	// that was crossing empty space at that step: ncu showed the skip loop running with 1 of 32 threads, 215 us per launch whatever the ray count.)
	// Phase 2: the 32 records of a step are assembled in shared memory and leave as full-line stores.
	float4* stepbuf = reinterpret_cast<float4*>(march_smem + (size_t)(blockDim.x >> 5) * 32u * stride) + (size_t)warp * kMarchChunk * 32u;
	for (uint32_t j0 = 0; j0 < n_steps; j0 += kMarchChunk) {
		const uint32_t cn = min((uint32_t)kMarchChunk, n_steps - j0);
		uint32_t k = 0;
		while (!left && k < cn) {
			const float px = fmaf(dx, t, ox), py = fmaf(dy, t, oy), pz = fmaf(dz, t, oz);
			if (!aabb_contains(a.render_aabb, px, py, pz)) {
				left = true;
				break;
			}
			const float dt = calc_dt(t, a.cone_angle);
			const uint32_t mip = max(a.min_mip, (uint32_t)mip_from_dt(dt, px, py, pz));
			if (!a.bitfield || occupied_at(px, py, pz, a.bitfield, mip)) {
				// warp_position(pos, train_aabb) = (pos - min) / diag; warp_dt
				stepbuf[k * 32u + lane] = make_float4((px - a.train_aabb[0]) / (a.train_aabb[3] - a.train_aabb[0]), (py - a.train_aabb[1]) / (a.train_aabb[4] - a.train_aabb[1]),
				                                      (pz - a.train_aabb[2]) / (a.train_aabb[5] - a.train_aabb[2]), (dt - min_stepsize()) / (max_step - min_stepsize()));
				t += dt;
				++steps;
				++k;
			} else {
				t = advance_to_next_voxel(t, a.cone_angle, px, py, pz, dx, dy, dz, ix, iy, iz, kGrid >> mip);
			}
		}
		for (; k < cn; ++k) stepbuf[k * 32u + lane] = make_float4(0.f, 0.f, 0.f, -1.0f);  // warped dt is never negative: marks a step the ray does not reach
		__syncwarp();
		for (k = 0; k < cn; ++k) {
			const float4 v = stepbuf[k * 32u + lane];
			if (v.w >= 0.0f) {
				my[0] = v.x; my[1] = v.y; my[2] = v.z;
				my[3] = v.w;
				my[4] = frame_id;
				my[5] = v.x; my[6] = v.y; my[7] = v.z;
			} else {
				my[4] = -1.0f;
			}
			__syncwarp();
			float* dst = records + (size_t)(i0 + (j0 + k) * n_alive) * stride;
			for (uint32_t q = lane; q < cnt * stride; q += 32) dst[q] = tile[q];
			__syncwarp();
		}
	}
	if (mine) {
		rays.o_t[i].w = t;
		rays.meta[i].y = (rays.meta[i].y & 0xffff0000u) | steps;  // low half: steps generated in this round (payload.n_steps)
	}
}

__device__ __forceinline__ float logistic(float x) { return 1.0f / (1.0f + expf(-x)); }

// composite_kernel_nerf (Shade mode, exponential density, logistic rgb) + compact_kernel_nerf + shade_kernel_nerf.
// out_stride == 0: network output is 4 halves per sample (r, g, b, sigma); else out[i + c * out_stride] as the reference.
__global__ void __launch_bounds__(128) nerf_composite_kernel(const RenderArgs a, RaySet rays, RaySet next, const uint32_t n_alive_arg, const float* __restrict__ records,
                                                             const uint32_t stride, const __half* __restrict__ net_out, const uint32_t out_stride, const uint32_t n_steps_arg,
                                                             const uint32_t current_step_arg, uint32_t* __restrict__ n_next, float4* __restrict__ frame_rgba,
                                                             float* __restrict__ frame_depth, const int linear_colors, const RenderSched* __restrict__ sched) {
	const uint32_t n_alive = sched ? sched->n_alive : n_alive_arg, n_steps = sched ? sched->n_steps : n_steps_arg,
	               current_step = sched ? sched->step : current_step_arg;
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	bool alive = false, finished = false;
	float4 rgba = make_float4(0, 0, 0, 0), o_t = rgba, d_w = rgba;
	float depth = 0.f;
	uint2 meta = make_uint2(0, 0);
	if (i < n_alive) {
		o_t = rays.o_t[i];
		d_w = rays.d_w[i];
		meta = rays.meta[i];
		rgba = rays.rgba[i];
		depth = rays.depth[i];
		const uint32_t actual = meta.y & 0xffffu;
		const float* cm = a.cam.camera_matrix;
		const float max_step = min_stepsize() * (1 << (kCascades - 1));
		uint32_t j = 0;
		for (; j < actual; ++j) {
			const size_t s = (size_t)i + (size_t)j * n_alive;
			float o0, o1, o2, o3;
			if (out_stride == 0) {
				const uint2 raw = *reinterpret_cast<const uint2*>(net_out + s * 4);
				const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
				o0 = lo.x; o1 = lo.y; o2 = hi.x; o3 = hi.y;
			} else {
				o0 = __half2float(net_out[s]); o1 = __half2float(net_out[s + out_stride]);
				o2 = __half2float(net_out[s + 2 * (size_t)out_stride]); o3 = __half2float(net_out[s + 3 * (size_t)out_stride]);
			}
			const float* rec = records + s * stride;
			// deformed_pos (floats 5-7) un-warped with the training box
			const float qx = fmaf(rec[5], a.train_aabb[3] - a.train_aabb[0], a.train_aabb[0]);
			const float qy = fmaf(rec[6], a.train_aabb[4] - a.train_aabb[1], a.train_aabb[1]);
			const float qz = fmaf(rec[7], a.train_aabb[5] - a.train_aabb[2], a.train_aabb[2]);
			const float T = 1.f - rgba.w;
			const float dt = fmaf(rec[3], max_step - min_stepsize(), min_stepsize());  // unwarp_dt
			const float alpha = 1.f - __expf(-__expf(o3) * dt);
			const float weight = alpha * T;
			rgba.x = fmaf(logistic(o0), weight, rgba.x);
			rgba.y = fmaf(logistic(o1), weight, rgba.y);
			rgba.z = fmaf(logistic(o2), weight, rgba.z);
			rgba.w += weight;
			if (weight > d_w.w) {
				d_w.w = weight;
				depth = fmaf(cm[6], qx - cm[9], fmaf(cm[7], qy - cm[10], cm[8] * (qz - cm[11])));
			}
			if (rgba.w > (1.0f - a.min_transmittance)) {
				const float inv = rgba.w;
				rgba.x /= inv; rgba.y /= inv; rgba.z /= inv; rgba.w /= inv;
				break;
			}
		}
		alive = !(j < n_steps);
		if (!alive) meta.y = (j + current_step) << 16 | (meta.y & 0xffffu);
		finished = !alive && rgba.w > 0.001f;
	}
	const uint32_t slot = warp_append(alive, n_next);
	if (alive) {
		next.o_t[slot] = o_t;
		next.d_w[slot] = d_w;
		next.meta[slot] = meta;
		next.rgba[slot] = rgba;
		next.depth[slot] = depth;
	}
	if (finished) {
		// shade_kernel_nerf, Shade mode: accumulate in linear colours over the frame buffer
		float4 c = rgba;
		if (!linear_colors) {
			auto s2l = [](float v) { return v <= 0.04045f ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f); };
			c.x = s2l(c.x); c.y = s2l(c.y); c.z = s2l(c.z);
		}
		const float4 fb = frame_rgba[meta.x];
		const float k = 1.0f - c.w;
		frame_rgba[meta.x] = make_float4(fmaf(fb.x, k, c.x), fmaf(fb.y, k, c.y), fmaf(fb.z, k, c.z), fmaf(fb.w, k, c.w));
		if (c.w > 0.2f) frame_depth[meta.x] = depth;
	}
}

}  // namespace

// ---- tracer state owned by the context ---------------------------------------------------------------------------------
struct Tracer {
	uint32_t cap = 0;
	float* d_records = nullptr;   // render_frame scratch: n_pixels records
	void* d_net_out = nullptr;    // n_pixels x 4 halves
	float* d_zero_extra = nullptr;
	uint32_t scratch_cap = 0, scratch_stride = 0;
	void* slab = nullptr;
	RaySet set[2]{};
	uint32_t* d_counters = nullptr;  // [0] alive (current), [1] alive (next)
	uint32_t* h_counters = nullptr;  // pinned
	uint32_t n_alive = 0, n_init = 0, cur = 0, step = 1;
	RenderSched* d_sched = nullptr;
	RenderSched* h_sched = nullptr;  // pinned
	uint32_t last_iterations = 0;    // of the previous frame: how many iterations render_frame enqueues before it first looks at the device's count
	RenderArgs args{};
};

static Tracer* tracer_of(insta_ctx* ctx) {
	if (!ctx->tracer) ctx->tracer = new Tracer();
	return (Tracer*)ctx->tracer;
}

void tracer_destroy(insta_ctx* ctx) {
	Tracer* t = (Tracer*)ctx->tracer;
	if (!t) return;
	if (t->slab) cudaFree(t->slab);
	if (t->d_records) cudaFree(t->d_records);
	if (t->d_net_out) cudaFree(t->d_net_out);
	if (t->d_zero_extra) cudaFree(t->d_zero_extra);
	if (t->d_counters) cudaFree(t->d_counters);
	if (t->h_counters) cudaFreeHost(t->h_counters);
	if (t->d_sched) cudaFree(t->d_sched);
	if (t->h_sched) cudaFreeHost(t->h_sched);
	delete t;
	ctx->tracer = nullptr;
}

static int tracer_reserve(insta_ctx* ctx, Tracer* t, uint32_t n) {
	if (!t->d_counters) {
		INSTA_CUDA(cudaMalloc(&t->d_counters, 64));
		INSTA_CUDA(cudaHostAlloc(&t->h_counters, 64, cudaHostAllocDefault));
		INSTA_CUDA(cudaMalloc(&t->d_sched, sizeof(RenderSched)));
		INSTA_CUDA(cudaHostAlloc(&t->h_sched, sizeof(RenderSched), cudaHostAllocDefault));
	}
	if (n <= t->cap) return INSTA_OK;
	INSTA_CUDA(cudaDeviceSynchronize());
	if (t->slab) cudaFree(t->slab);
	t->slab = nullptr;
	const size_t per = (size_t)n * (16 + 16 + 8 + 16 + 4);
	INSTA_CUDA(cudaMalloc(&t->slab, per * 2));
	unsigned char* p = (unsigned char*)t->slab;
	for (int k = 0; k < 2; ++k) {
		t->set[k].o_t = (float4*)p; p += (size_t)n * 16;
		t->set[k].d_w = (float4*)p; p += (size_t)n * 16;
		t->set[k].rgba = (float4*)p; p += (size_t)n * 16;
		t->set[k].meta = (uint2*)p; p += (size_t)n * 8;
		t->set[k].depth = (float*)p; p += (size_t)n * 4;
	}
	t->cap = n;
	return INSTA_OK;
}

static void fill_args(const insta_ctx* ctx, const insta_camera* cam, const insta_render_opts* o, RenderArgs& a) {
	a.cam = *cam;
	for (int k = 0; k < 6; ++k) {
		a.render_aabb[k] = o && o->use_render_aabb ? o->render_aabb[k] : ctx->render_aabb[k];
		a.train_aabb[k] = ctx->aabb[k];
	}
	a.cone_angle = o ? o->cone_angle_constant : 0.0f;
	a.min_transmittance = o ? o->min_transmittance : 0.01f;
	a.bitfield = o ? (const uint8_t*)o->d_density_grid_bitfield : nullptr;
	a.min_mip = 0;
}

int tracer_init_rays(insta_ctx* ctx, const insta_camera* cam, const insta_render_opts* opts, float* d_frame_rgba, float* d_frame_depth, float* d_dbg_origin_t,
                     float* d_dbg_dir_alive, uint32_t* h_n_alive, cudaStream_t s) {
	Tracer* t = tracer_of(ctx);
	const uint32_t W = cam->resolution[0], H = cam->resolution[1];
	if (W == 0 || H == 0) INSTA_FAIL(INSTA_ERR_INVALID, "insta_tracer_init_rays: empty resolution");
	int rc = tracer_reserve(ctx, t, W * H);
	if (rc) return rc;
	fill_args(ctx, cam, opts, t->args);
	t->cur = 0;
	t->step = 1;
	t->n_init = W * H;
	INSTA_CUDA(cudaMemsetAsync(t->d_counters, 0, 64, s));
	const dim3 grid((W + 15) / 16, (H + 7) / 8);
	nerf_init_rays_kernel<<<grid, 128, 0, s>>>(t->args, t->set[0], t->d_counters, (float4*)d_frame_rgba, d_frame_depth, (float4*)d_dbg_origin_t, (float4*)d_dbg_dir_alive);
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	if (!h_n_alive) return INSTA_OK;  // insta_render_frame: the count stays on the device
	INSTA_CUDA(cudaMemcpyAsync(t->h_counters, t->d_counters, 4, cudaMemcpyDeviceToHost, s));
	INSTA_CUDA(cudaStreamSynchronize(s));
	t->n_alive = t->h_counters[0];
	*h_n_alive = t->n_alive;
	return INSTA_OK;
}

uint32_t tracer_n_steps(const insta_ctx* ctx) {
	const Tracer* t = (const Tracer*)ctx->tracer;
	if (!t || t->n_alive == 0) return 0;
	// NerfTracer::trace: clamp(n_rays_initialized / n_alive, 1, 8)  (testbed_nerf.cu:2172)
	return std::min(std::max(t->n_init / t->n_alive, 1u), 8u);
}

int tracer_generate(insta_ctx* ctx, float* d_records, uint32_t stride, uint32_t n_steps, int32_t frame_id, const float* d_extra_dims, uint32_t* n_records,
                    cudaStream_t s) {
	Tracer* t = tracer_of(ctx);
	if (n_records) *n_records = t->n_alive * n_steps;
	if (t->n_alive == 0) return INSTA_OK;
	if (stride < 11) INSTA_FAIL(INSTA_ERR_INVALID, "insta_tracer_generate_samples: stride < 11");
	const size_t march_smem = (size_t)4 * 32 * stride * sizeof(float) + (size_t)4 * kMarchChunk * 32 * sizeof(float4);
	if (march_smem > 48 * 1024) INSTA_CUDA(cudaFuncSetAttribute(nerf_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)march_smem));
	nerf_march_kernel<<<(t->n_alive + 127) / 128, 128, march_smem, s>>>(t->args, t->set[t->cur], t->n_alive, d_records, stride, n_steps, (float)frame_id,
	                                                                    d_extra_dims, stride - 11, nullptr);
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

int tracer_composite(insta_ctx* ctx, const float* d_records, uint32_t stride, const void* d_net_out, uint32_t out_stride, uint32_t n_steps, float* d_frame_rgba,
                     float* d_frame_depth, int linear_colors, uint32_t* h_n_alive, cudaStream_t s) {
	Tracer* t = tracer_of(ctx);
	if (t->n_alive == 0) {
		if (h_n_alive) *h_n_alive = 0;
		return INSTA_OK;
	}
	uint32_t* cnt_next = t->d_counters + 1 + (t->cur ^ 1u);  // [1] / [2], alternating
	INSTA_CUDA(cudaMemsetAsync(cnt_next, 0, 4, s));
	nerf_composite_kernel<<<(t->n_alive + 127) / 128, 128, 0, s>>>(t->args, t->set[t->cur], t->set[t->cur ^ 1u], t->n_alive, d_records, stride, (const __half*)d_net_out,
	                                                               out_stride, n_steps, t->step, cnt_next, (float4*)d_frame_rgba, d_frame_depth, linear_colors, nullptr);
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	INSTA_CUDA(cudaMemcpyAsync(t->h_counters, cnt_next, 4, cudaMemcpyDeviceToHost, s));
	INSTA_CUDA(cudaStreamSynchronize(s));
	t->cur ^= 1u;
	t->step += n_steps;
	t->n_alive = t->h_counters[0];
	if (h_n_alive) *h_n_alive = t->n_alive;
	return INSTA_OK;
}


// Testbed::render_nerf -> NerfTracer::init_rays_from_camera + trace + shade (testbed_nerf.cu:2320-2440, 2112-2244) with this
// library's own deformation and network between generate and composite.  One stream sync per march iteration (the live-ray
// count decides the number of steps of the next one), as in the reference.
int render_frame(insta_ctx* ctx, const insta_camera* cam, const insta_render_opts* opts, int32_t frame_id, const void* d_params, const float* d_extra_dims,
                 float* d_frame_rgba, float* d_frame_depth, insta_render_stats* stats, cudaStream_t s) {
	Tracer* t = tracer_of(ctx);
	const uint32_t n_pix = cam->resolution[0] * cam->resolution[1];
	// Between this function's own kernels a record needs no expression floats: they are the same n_extra values for every sample of the frame, except
	// that `warp` overwrites them with ones for samples whose triangle is outside the lips region (core.cu:515-519).  The compact record is the 11
	// NerfCoordinate floats; the deformation flags "ones" in the holder slot (-2) and the network kernel reads the frame's values from d_extra_dims:
	// 44 instead of 108 bytes per sample through march / deformation / network.  The frame is bit-identical (tests/test_gpu_nerf.py).
	const bool compact = ctx->p.geo_params == 0 && !ctx->render_full_records;
	const uint32_t stride = compact ? 11u : 11 + ctx->p.exp_params + ctx->p.geo_params;
	// Steps per iteration: the reference marches clamp(rays initialised / n_alive, 1, 8) steps between two compactions, i.e. at most one record per
	// pixel and launch -- a few hundred thousand records, which neither the deformation nor the network kernel can turn into throughput.  The frame
	// does not depend on that schedule, so by default an iteration fills a budget of max(pixels, 2^21) records with up to 16 steps per ray
	// (512^2, B200: 9.55 ms per frame with the reference's schedule, 7.00 with 16 steps, 7.24 with 32, 7.76 with 64 -- profiles/r1_render_schedule.txt).
	const uint32_t max_steps = opts && opts->max_steps_per_iteration ? std::min(opts->max_steps_per_iteration, 0xffffu) : 16u;
	const uint32_t budget = max_steps <= 8u ? n_pix : std::max(n_pix, 1u << 21);
	// opt-in (INSTA_RENDER_TAIL="log2 records,steps"): an iteration that would carry fewer than 2^tail records marches up to `steps` steps instead.  Measured
	// (profiles/r2_bwd_ab_v1_scatter_roofline_render_tail.log): "19,64" turns 6 iterations into 4 but evaluates 7 % more samples -- 4.92 ms per 512^2 frame either way, so it is off
	static const uint32_t tail_log2 = [] { const char* e = getenv("INSTA_RENDER_TAIL"); return e ? (uint32_t)atoi(e) : 0u; }();
	static const uint32_t tail_steps_env = [] { const char* e = getenv("INSTA_RENDER_TAIL"); const char* c = e ? strchr(e, ',') : nullptr; return c ? (uint32_t)atoi(c + 1) : 64u; }();
	const bool tail_on = !(opts && opts->max_steps_per_iteration) && tail_log2 > 0 && tail_log2 < 31;
	const uint32_t tail_budget = tail_on ? std::min(1u << tail_log2, budget) : 0u, tail_steps = tail_on ? tail_steps_env : 0u;
	if (budget > t->scratch_cap || stride != t->scratch_stride) {
		INSTA_CUDA(cudaDeviceSynchronize());
		if (t->d_records) cudaFree(t->d_records);
		if (t->d_net_out) cudaFree(t->d_net_out);
		t->d_records = nullptr;
		t->d_net_out = nullptr;
		t->scratch_cap = 0;
		INSTA_CUDA(cudaMalloc(&t->d_records, (size_t)budget * stride * 4));
		INSTA_CUDA(cudaMalloc(&t->d_net_out, (size_t)budget * 8));
		t->scratch_cap = budget;
		t->scratch_stride = stride;
	}
	if (!t->d_zero_extra) {
		INSTA_CUDA(cudaMalloc(&t->d_zero_extra, 64 * 4));
		INSTA_CUDA(cudaMemset(t->d_zero_extra, 0, 64 * 4));
	}
	INSTA_CUDA(cudaMemsetAsync(d_frame_rgba, 0, (size_t)n_pix * 16, s));
	int rc = tracer_init_rays(ctx, cam, opts, d_frame_rgba, d_frame_depth, nullptr, nullptr, nullptr, s);
	if (rc) return rc;
	const bool deform = frame_id >= 0 && !ctx->frames.empty();
	if (deform) {
		rc = sync_frame_table(ctx, s);
		if (rc) return rc;
	}
	// The iterations are enqueued WITHOUT reading the live-ray count back (NerfTracer::trace synchronises once per iteration, testbed_nerf.cu:2156-2165):
	// render_schedule_kernel turns the survivor count of iteration k - 1 into the sizes of iteration k on the device, the kernels of an iteration are
	// launched for the upper bounds and take their sizes from that block.  The host looks at the count after as many iterations as the previous frame
	// needed (an iteration without rays costs five empty launches), then every two.
	const float* extra = d_extra_dims ? d_extra_dims : t->d_zero_extra;
	const size_t march_smem = (size_t)4 * 32 * stride * sizeof(float) + (size_t)4 * kMarchChunk * 32 * sizeof(float4);
	if (march_smem > 48 * 1024) INSTA_CUDA(cudaFuncSetAttribute(nerf_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)march_smem));
	const uint32_t ray_blocks = (n_pix + 127) / 128;
	uint32_t cur = 0, enqueued = 0;
	bool first = true;
	uint32_t group = std::min(std::max(t->last_iterations, 4u), 16u);
	while (true) {
		for (uint32_t g = 0; g < group; ++g, ++enqueued) {
			uint32_t* cnt_next = t->d_counters + 1 + (cur ^ 1u);
			render_schedule_kernel<<<1, 1, 0, s>>>(t->d_sched, first ? t->d_counters : t->d_counters + 1 + cur, cnt_next, budget, max_steps, first ? 1 : 0, tail_budget, tail_steps);
			nerf_march_kernel<<<ray_blocks, 128, march_smem, s>>>(t->args, t->set[cur], 0, t->d_records, stride, 0, (float)frame_id, extra, stride - 11, t->d_sched);
			ctx->launches += 2;
			INSTA_CUDA(cudaGetLastError());
			if (deform) {
				rc = launch_deform(ctx, t->d_records, budget, stride, frame_id, nullptr, false, nullptr, s, &t->d_sched->n_records, /*coherent_hint=*/true, compact);
				if (rc) return rc;
			}
			rc = launch_network_fwd(ctx, t->d_records, stride, budget, d_params, t->d_net_out, 0, s, &t->d_sched->n_records, compact ? extra : nullptr);
			if (rc) return rc;
			nerf_composite_kernel<<<ray_blocks, 128, 0, s>>>(t->args, t->set[cur], t->set[cur ^ 1u], 0, t->d_records, stride, (const __half*)t->d_net_out, 0, 0, 0, cnt_next,
			                                                (float4*)d_frame_rgba, d_frame_depth, opts ? (int)opts->linear_colors : 0, t->d_sched);
			ctx->launches++;
			INSTA_CUDA(cudaGetLastError());
			cur ^= 1u;
			first = false;
		}
		INSTA_CUDA(cudaMemcpyAsync(t->h_sched, t->d_sched, sizeof(RenderSched), cudaMemcpyDeviceToHost, s));
		INSTA_CUDA(cudaMemcpyAsync(t->h_counters, t->d_counters + 1 + cur, 4, cudaMemcpyDeviceToHost, s));
		INSTA_CUDA(cudaStreamSynchronize(s));
		if (t->h_counters[0] == 0 || enqueued >= 4096u) break;
		group = 2;
	}
	const uint32_t iterations = t->h_sched->iterations;
	const uint64_t samples = t->h_sched->samples;
	t->last_iterations = iterations;
	t->n_alive = 0;
	t->cur = cur;
	if (stats) {
		stats->iterations = iterations;
		stats->samples = samples;
	}
	return INSTA_OK;
}

}  // namespace insta
