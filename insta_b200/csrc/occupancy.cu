// Occupancy-grid maintenance on sm_100a: Testbed::update_density_grid_nerf / update_density_grid_mean_and_bitfield
// (ngp/src/testbed_nerf.cu:2774-2878) with the kernels they launch (391-438 mark_untrained_density_grid, 487-516
// generate_grid_samples_nerf_nonuniform, 518-534 splat_grid_samples_nerf_max_nearest_neighbor, 554-577 ema_grid_samples_nerf,
// 585-632 grid_to_bitfield / bitfield_max_pool), the INSTA hook in the middle (rta::Core::surface_closest_point on bare xyz ->
// `warp_random_samples`, src/core.cu:528-579, called at testbed_nerf.cu:2850) and NerfNetwork::density (nerf_network.h:385-414).
//
// Not a port: the two sample sets (uniform / non-uniform threshold) come from ONE launch, the posed points are deformed by the same
// warp-queue kernel as the training samples (random frame per point), the density comes from the fused tcgen05 network kernel in its
// density-only mode (the encoded features never exist in memory), and the EMA update is fused with the mean reduction.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>

#include "ctx.h"
#include "nerf_device.cuh"

namespace insta {

struct OccScratch {
	uint32_t cap_samples = 0, cap_cells = 0;
	float* positions = nullptr;    // n_samples x 3 (NerfPosition, ngp nerf.h:62-74)
	uint32_t* indices = nullptr;   // n_samples
	float* grid_tmp = nullptr;     // n_cells
	__half* mlp_out = nullptr;     // n_samples density logits
};

namespace {
using namespace nerfdev;
constexpr uint32_t kCells = kGrid * kGrid * kGrid;
constexpr float kMinOpticalThickness = 0.01f;  // NERF_MIN_OPTICAL_THICKNESS (testbed_nerf.cu:63)

// tcnn::morton3D_invert (tcnn common_device.h:356-363)
__device__ __forceinline__ uint32_t morton3D_invert(uint32_t x) {
	x = x & 0x49249249u;
	x = (x | (x >> 2)) & 0xc30c30c3u;
	x = (x | (x >> 4)) & 0x0f00f00fu;
	x = (x | (x >> 8)) & 0xff0000ffu;
	x = (x | (x >> 16)) & 0x0000ffffu;
	return x;
}

// mark_untrained_density_grid (testbed_nerf.cu:391-438): -1 for cells no training camera sees, 0 for the visible ones
__global__ void __launch_bounds__(128) occ_mark_untrained_kernel(uint32_t n_elements, float* __restrict__ grid, uint32_t n_images, const ImageDev* __restrict__ images,
                                                                 int clear_visible) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const uint32_t level = i / kCells, pos_idx = i % kCells;
	const float s = scalbnf(1.0f, (int)level);
	const float px = ((float)morton3D_invert(pos_idx) + 0.5f) / (float)kGrid, py = ((float)morton3D_invert(pos_idx >> 1) + 0.5f) / (float)kGrid,
	            pz = ((float)morton3D_invert(pos_idx >> 2) + 0.5f) / (float)kGrid;
	const float x0 = (px - 0.5f) * s + 0.5f, y0 = (py - 0.5f) * s + 0.5f, z0 = (pz - 0.5f) * s + 0.5f;
	const float voxel_radius = 0.5f * sqrt3() * s / (float)kGrid;
	int count = 0;
	for (uint32_t j = 0; j < n_images; ++j) {
		const ImageDev& im = images[j];
		const float half_resx = (float)im.res[0] * 0.5f, half_resy = (float)im.res[1] * 0.5f;
		const float lx = x0 - im.xform[9], ly = y0 - im.xform[10], lz = z0 - im.xform[11];
		const float x = lx * im.xform[0] + ly * im.xform[1] + lz * im.xform[2];
		const float y = lx * im.xform[3] + ly * im.xform[4] + lz * im.xform[5];
		const float z = lx * im.xform[6] + ly * im.xform[7] + lz * im.xform[8];
		if (z > 0.f && fabsf(x) - voxel_radius < z / im.focal[0] * half_resx && fabsf(y) - voxel_radius < z / im.focal[1] * half_resy) {
			count = 1;
			break;
		}
	}
	if (clear_visible || (grid[i] < 0) != (count <= 0)) grid[i] = count > 0 ? 0.f : -1.f;
}

// generate_grid_samples_nerf_nonuniform (testbed_nerf.cu:487-516) for both sample sets of one update: samples [0, n_uniform) take any
// cell that is not marked untrained (thresh -0.01), samples [n_uniform, n_uniform + n_nonuniform) prefer occupied cells (thresh
// NERF_MIN_OPTICAL_THICKNESS).  Both sets index the SAME rng state with their local index, as the reference's two launches do.
__global__ void __launch_bounds__(128) occ_generate_kernel(uint32_t n_uniform, uint32_t n_nonuniform, uint64_t rng_state, uint64_t rng_inc, uint32_t step, float ax, float ay, float az, float sx, float sy, float sz,
                                                           const float* __restrict__ grid_in, float* __restrict__ out_xyz, uint32_t* __restrict__ indices,
                                                           uint32_t n_cascades) {
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= n_uniform + n_nonuniform) return;
	const bool uniform = g < n_uniform;
	const uint32_t i = uniform ? g : g - n_uniform;
	const uint32_t n_elements = uniform ? n_uniform : n_nonuniform;
	const float thresh = uniform ? -0.01f : kMinOpticalThickness;
	Pcg32 rng{rng_state, rng_inc};
	rng.advance((uint64_t)(int64_t)(uint32_t)(i * 4u));
	const uint32_t level = (uint32_t)(rng.next_float() * (float)n_cascades) % n_cascades;
	uint32_t idx = 0;
	for (uint32_t j = 0; j < 10; ++j) {
		idx = ((i + step * n_elements) * 56924617u + j * 19349663u + 96925573u) % kCells;
		idx += level * kCells;
		if (__ldg(grid_in + idx) > thresh) break;
	}
	const uint32_t pos_idx = idx % kCells;
	const float rx = rng.next_float(), ry = rng.next_float(), rz = rng.next_float();
	const float s = scalbnf(1.0f, (int)level);
	// ((cell + r) / 128 - 0.5) * 2^level + 0.5: the scale factors are powers of two, so contraction cannot change the result
	const float px = (((float)morton3D_invert(pos_idx) + rx) / (float)kGrid - 0.5f) * s + 0.5f;
	const float py = (((float)morton3D_invert(pos_idx >> 1) + ry) / (float)kGrid - 0.5f) * s + 0.5f;
	const float pz = (((float)morton3D_invert(pos_idx >> 2) + rz) / (float)kGrid - 0.5f) * s + 0.5f;
	// warp_position = aabb.relative_pos (testbed_nerf.cu:262-267)
	out_xyz[(size_t)g * 3 + 0] = (px - ax) / sx;
	out_xyz[(size_t)g * 3 + 1] = (py - ay) / sy;
	out_xyz[(size_t)g * 3 + 2] = (pz - az) / sz;
	indices[g] = idx;
}

// splat_grid_samples_nerf_max_nearest_neighbor (testbed_nerf.cu:518-534), density activation = exponential (__expf)
__global__ void __launch_bounds__(256) occ_splat_kernel(uint32_t n, const uint32_t* __restrict__ indices, const __half* __restrict__ logits, float* __restrict__ grid_tmp) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float optical_thickness = __expf(__half2float(logits[i])) * min_stepsize();
	// positive floats order like their bit patterns
	atomicMax(reinterpret_cast<uint32_t*>(grid_tmp) + indices[i], __float_as_uint(optical_thickness));
}

// ema_grid_samples_nerf (testbed_nerf.cu:554-577; maximum instead of EMA) fused with the mean of max(val, 0) over the first cascade
// (update_density_grid_mean_and_bitfield, 2867)
__global__ void __launch_bounds__(256) occ_ema_mean_kernel(uint32_t n_elements, float decay, float* __restrict__ grid, const float* __restrict__ grid_tmp,
                                                           float* __restrict__ mean) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	float contrib = 0.f;
	if (i < n_elements) {
		float val = grid[i];
		if (grid_tmp) {
			const float importance = grid_tmp[i];
			val = val < 0.f ? val : fmaxf(val * decay, importance);
			grid[i] = val;
		}
		if (i < kCells) contrib = fmaxf(val, 0.f) / (float)kCells;
	}
	for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
	__shared__ float part[8];
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = contrib;
	__syncthreads();
	if (threadIdx.x < 8) {
		float v = part[threadIdx.x];
		for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
		if (threadIdx.x == 0 && blockIdx.x * blockDim.x < kCells) atomicAdd(mean, v);
	}
}

// grid_to_bitfield (testbed_nerf.cu:585-609): one thread = 8 cells = 32 bytes in, 1 byte out
__global__ void __launch_bounds__(256) occ_bitfield_kernel(uint32_t n_elements, uint32_t n_nonzero, const float* __restrict__ grid, uint8_t* __restrict__ bitfield,
                                                           const float* __restrict__ mean) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	if (i >= n_nonzero) {
		bitfield[i] = 0;
		return;
	}
	const float thresh = fminf(kMinOpticalThickness, __ldg(mean));
	const float4 a = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)i), b = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)i + 1);
	uint32_t bits = 0;
	bits |= a.x > thresh ? 1u : 0u;
	bits |= a.y > thresh ? 2u : 0u;
	bits |= a.z > thresh ? 4u : 0u;
	bits |= a.w > thresh ? 8u : 0u;
	bits |= b.x > thresh ? 16u : 0u;
	bits |= b.y > thresh ? 32u : 0u;
	bits |= b.z > thresh ? 64u : 0u;
	bits |= b.w > thresh ? 128u : 0u;
	bitfield[i] = (uint8_t)bits;
}

// bitfield_max_pool (testbed_nerf.cu:611-632): 8 bytes of the finer level -> 8 bits, OR-ed into the centre of the coarser level
__global__ void __launch_bounds__(256) occ_max_pool_kernel(uint32_t n_elements, const uint8_t* __restrict__ prev, uint8_t* __restrict__ next) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const uint2 raw = __ldg(reinterpret_cast<const uint2*>(prev) + i);
	uint32_t bits = 0;
#pragma unroll
	for (uint32_t j = 0; j < 4; ++j) {
		bits |= ((raw.x >> (8 * j)) & 0xffu) ? (1u << j) : 0u;
		bits |= ((raw.y >> (8 * j)) & 0xffu) ? (16u << j) : 0u;
	}
	const uint32_t x = morton3D_invert(i) + kGrid / 8, y = morton3D_invert(i >> 1) + kGrid / 8, z = morton3D_invert(i >> 2) + kGrid / 8;
	next[morton3D(x, y, z)] |= (uint8_t)bits;
}
}  // namespace

void occupancy_destroy(insta_ctx* ctx) {
	OccScratch* s = (OccScratch*)ctx->occ_scratch;
	if (!s) return;
	cudaFree(s->positions); cudaFree(s->indices); cudaFree(s->grid_tmp); cudaFree(s->mlp_out);
	delete s;
	ctx->occ_scratch = nullptr;
}

static int occ_reserve(insta_ctx* ctx, uint32_t n_samples, uint32_t n_cells) {
	if (!ctx->occ_scratch) ctx->occ_scratch = new OccScratch();
	OccScratch* s = (OccScratch*)ctx->occ_scratch;
	if (n_samples > s->cap_samples) {
		INSTA_CUDA(cudaDeviceSynchronize());
		cudaFree(s->positions); cudaFree(s->indices); cudaFree(s->mlp_out);
		s->positions = nullptr; s->indices = nullptr; s->mlp_out = nullptr; s->cap_samples = 0;
		INSTA_CUDA(cudaMalloc(&s->positions, sizeof(float) * 3 * (size_t)n_samples));
		INSTA_CUDA(cudaMalloc(&s->indices, sizeof(uint32_t) * (size_t)n_samples));
		INSTA_CUDA(cudaMalloc(&s->mlp_out, sizeof(__half) * (size_t)n_samples));
		s->cap_samples = n_samples;
	}
	if (n_cells > s->cap_cells) {
		INSTA_CUDA(cudaDeviceSynchronize());
		cudaFree(s->grid_tmp);
		s->grid_tmp = nullptr; s->cap_cells = 0;
		INSTA_CUDA(cudaMalloc(&s->grid_tmp, sizeof(float) * (size_t)n_cells));
		s->cap_cells = n_cells;
	}
	return INSTA_OK;
}

// update_density_grid_mean_and_bitfield (testbed_nerf.cu:2860-2878); d_grid_tmp != null fuses the EMA / max update into the mean pass
static int mean_and_bitfield(insta_ctx* ctx, const insta_density_grid* g, const float* d_grid_tmp, float decay, cudaStream_t s) {
	const uint32_t n_all = kCells * g->n_cascades;
	INSTA_CUDA(cudaMemsetAsync(g->d_density_grid_mean, 0, sizeof(float), s));
	occ_ema_mean_kernel<<<((d_grid_tmp ? n_all : kCells) + 255) / 256, 256, 0, s>>>(d_grid_tmp ? n_all : kCells, decay, g->d_density_grid, d_grid_tmp, g->d_density_grid_mean);
	uint8_t* bf = (uint8_t*)g->d_density_grid_bitfield;
	occ_bitfield_kernel<<<(kCells / 8 * kCascades + 255) / 256, 256, 0, s>>>(kCells / 8 * kCascades, kCells / 8 * g->n_cascades, g->d_density_grid, bf, g->d_density_grid_mean);
	ctx->launches += 2;
	for (uint32_t level = 1; level < kCascades; ++level) {
		occ_max_pool_kernel<<<(kCells / 64 + 255) / 256, 256, 0, s>>>(kCells / 64, bf + (size_t)(level - 1) * (kCells / 8), bf + (size_t)level * (kCells / 8));
		ctx->launches++;
	}
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

int density_grid_mean_and_bitfield(insta_ctx* ctx, const insta_density_grid* g, cudaStream_t s) { return mean_and_bitfield(ctx, g, nullptr, 0.f, s); }

int density_grid_update(insta_ctx* ctx, insta_density_grid* g, const void* d_params_f16, uint32_t n_uniform, uint32_t n_nonuniform, cudaStream_t s) {
	const uint32_t n_cells = kCells * g->n_cascades;
	const uint32_t n_samples = n_uniform + n_nonuniform;
	int rc = occ_reserve(ctx, n_samples, n_cells);
	if (rc) return rc;
	OccScratch* sc = (OccScratch*)ctx->occ_scratch;

	// first step (or a changed training set): cull the cells no camera sees (testbed_nerf.cu:2797-2813)
	if (g->training_step == 0 || g->images_changed) {
		if (g->training_step == 0) g->ema_step = 0;
		if (!g->has_rays && ctx->n_train_images > 0) {
			occ_mark_untrained_kernel<<<(n_cells + 127) / 128, 128, 0, s>>>(n_cells, g->d_density_grid, ctx->n_train_images, (const ImageDev*)ctx->d_train_images,
			                                                                g->training_step == 0);
			ctx->launches++;
		} else {
			INSTA_CUDA(cudaMemsetAsync(g->d_density_grid, 0, sizeof(float) * n_cells, s));
		}
		g->images_changed = 0;
	}
	INSTA_CUDA(cudaMemsetAsync(sc->grid_tmp, 0, sizeof(float) * n_cells, s));
	if (n_samples > 0) {
		uint64_t state, inc;
		pcg_seed(g->rng_seed, g->rng_advances, state, inc);
		const float* a = ctx->aabb;
		occ_generate_kernel<<<(n_samples + 127) / 128, 128, 0, s>>>(n_uniform, n_nonuniform, state, inc, g->ema_step, a[0], a[1], a[2], a[3] - a[0], a[4] - a[1], a[5] - a[2], g->d_density_grid, sc->positions, sc->indices, g->n_cascades);
		ctx->launches++;
		INSTA_CUDA(cudaGetLastError());
		// rta::Core::surface_closest_point, 3-row path -> warp_random_samples: a random frame per point (core.cu:528-579)
		if (!ctx->frames.empty()) {
			{
				std::lock_guard<std::mutex> lock(ctx->mu);
				rc = sync_frame_table(ctx, s);
			}
			if (rc) return rc;
			rc = launch_deform(ctx, sc->positions, n_samples, 3, -1, nullptr, true, nullptr, s);
			if (rc) return rc;
		}
		// NerfNetwork::density with the training parameters; the expression inputs are all 1 (extract_expressions_random, nerf_network.h:108-128)
		rc = launch_network_density(ctx, sc->positions, 3, n_samples, d_params_f16, sc->mlp_out, s);
		if (rc) return rc;
		occ_splat_kernel<<<(n_samples + 255) / 256, 256, 0, s>>>(n_samples, sc->indices, sc->mlp_out, sc->grid_tmp);
		ctx->launches++;
	}
	g->rng_advances += 1;  // density_grid_rng.advance() (testbed_nerf.cu:2845)
	trainer_rng_advance(ctx);  // m_rng.advance() between the two sample sets (2831): the TRAINING generator moves too
	rc = mean_and_bitfield(ctx, g, sc->grid_tmp, g->decay, s);
	g->ema_step += 1;
	return rc;
}

}  // namespace insta
