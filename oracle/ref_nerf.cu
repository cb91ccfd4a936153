// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI launcher around the REFERENCE's own stage-3 CUDA kernels (ngp/src/testbed_nerf.cu), compiled for sm_100 from the
// source where it lies: oracle/Makefile extracts the kernel definitions (that file also contains the Testbed / GUI code)
// into oracle/_ref/nerf_extracted.inc at build time and this harness includes them unmodified:
//   init_rays_with_payload_kernel_nerf 1854, advance_pos_nerf 634, generate_next_nerf_network_inputs 722,
//   composite_kernel_nerf 797, compact_kernel_nerf 1829, shade_kernel_nerf 1794, generate_training_samples_nerf 1105,
//   compute_loss_kernel_train_nerf 1305 and the device helpers they call.
// Launch shapes are the reference's (NerfTracer::init_rays_from_camera / trace, testbed_nerf.cu:2030-2244).
#include <neural-graphics-primitives/common.h>
#include <neural-graphics-primitives/common_device.cuh>
#include <neural-graphics-primitives/bounding_box.cuh>
#include <neural-graphics-primitives/nerf.h>
#include <neural-graphics-primitives/nerf_loader.h>
#include <neural-graphics-primitives/random_val.cuh>
#include <neural-graphics-primitives/triangle_bvh.cuh>
#include <neural-graphics-primitives/envmap.cuh>

#include <tiny-cuda-nn/common.h>
#include <tiny-cuda-nn/common_device.h>
#include <tiny-cuda-nn/gpu_matrix.h>
#include <tiny-cuda-nn/reduce_sum.h>

#include <vector>

using namespace Eigen;
using namespace tcnn;

NGP_NAMESPACE_BEGIN
#include "nerf_extracted.inc"
NGP_NAMESPACE_END

using namespace ngp;

namespace {
struct RefTracer {
	uint32_t n = 0;
	Array4f* rgba[3] = {nullptr, nullptr, nullptr};
	float* depth[3] = {nullptr, nullptr, nullptr};
	NerfPayload* payload[3] = {nullptr, nullptr, nullptr};
	uint32_t* counters = nullptr;  // [0] alive, [1] hit
	uint32_t n_alive = 0, n_init = 0, dbi = 0, step = 1;
	BoundingBox render_aabb, train_aabb;
	Matrix<float, 3, 4> cam;
	Vector2f focal;
	float cone = 0.f, min_T = 0.01f;
	const uint8_t* grid = nullptr;
};
RefTracer g;

BoundingBox bb6(const float* a) {
	BoundingBox b;
	b.min = Vector3f(a[0], a[1], a[2]);
	b.max = Vector3f(a[3], a[4], a[5]);
	return b;
}
}  // namespace

extern "C" {

// init_rays_from_camera with the parameters Testbed::render_nerf passes in INSTA's setting (no distortion, no envmap,
// parallax (0,0,scale=1), aperture 0, plane_z = 1, no rolling shutter, identity render_aabb_to_local, Shade mode)
int refnerf_init_rays(const float* cam12, const float* focal2, const float* center2, const uint32_t* res2, uint32_t spp, int snap, const float* render_aabb6,
                      const float* train_aabb6, const uint8_t* d_bitfield, float cone_angle, float min_transmittance, float* d_frame_rgba, float* d_frame_depth,
                      cudaStream_t stream) {
	const uint32_t n = res2[0] * res2[1];
	if (n > g.n) {
		for (int k = 0; k < 3; ++k) {
			cudaFree(g.rgba[k]); cudaFree(g.depth[k]); cudaFree(g.payload[k]);
			cudaMalloc(&g.rgba[k], sizeof(Array4f) * n);
			cudaMalloc(&g.depth[k], sizeof(float) * n);
			cudaMalloc(&g.payload[k], sizeof(NerfPayload) * n);
		}
		if (!g.counters) cudaMalloc(&g.counters, 16);
		g.n = n;
	}
	g.render_aabb = bb6(render_aabb6);
	g.train_aabb = bb6(train_aabb6);
	for (int c = 0; c < 4; ++c)
		for (int r = 0; r < 3; ++r) g.cam(r, c) = cam12[c * 3 + r];
	g.focal = Vector2f(focal2[0], focal2[1]);
	g.cone = cone_angle;
	g.min_T = min_transmittance;
	g.grid = d_bitfield;
	const Vector2i resolution((int)res2[0], (int)res2[1]);
	const dim3 threads = {16, 8, 1};
	const dim3 blocks = {div_round_up(res2[0], threads.x), div_round_up(res2[1], threads.y), 1};
	BoundingSphere bs;
	init_rays_with_payload_kernel_nerf<<<blocks, threads, 0, stream>>>(spp, g.payload[0], resolution, g.focal, g.cam, g.cam, Vector4f::Zero(),
	                                                                  Vector2f(center2[0], center2[1]), Vector3f(0.f, 0.f, 1.f), (bool)snap, g.render_aabb, bs,
	                                                                  Matrix3f::Identity(), 1.0f, 0.0f, CameraDistortion{}, nullptr, Vector2i::Zero(),
	                                                                  (Array4f*)d_frame_rgba, d_frame_depth, nullptr, Vector2i::Zero(), ERenderMode::Shade);
	g.n_init = n;
	cudaMemsetAsync(g.rgba[0], 0, n * sizeof(Array4f), stream);
	cudaMemsetAsync(g.depth[0], 0, n * sizeof(float), stream);
	linear_kernel(advance_pos_nerf, 0, stream, n, g.render_aabb, bs, Matrix3f::Identity(), Vector3f(g.cam.col(2)), g.focal, spp, g.payload[0], d_bitfield, 0u,
	              cone_angle);
	cudaMemsetAsync(g.counters, 0, 16, stream);
	g.n_alive = n;
	g.dbi = 0;
	g.step = 1;
	return (int)cudaGetLastError();
}

// per-pixel payload state (origin, t) / (dir, alive) of the freshly initialised rays
__global__ void refnerf_export_kernel(uint32_t n, const NerfPayload* p, float4* o_t, float4* d_alive) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	o_t[i] = make_float4(p[i].origin.x(), p[i].origin.y(), p[i].origin.z(), p[i].t);
	d_alive[i] = make_float4(p[i].dir.x(), p[i].dir.y(), p[i].dir.z(), p[i].alive ? 1.f : 0.f);
}
int refnerf_export_init(float* d_o_t, float* d_d_alive, cudaStream_t stream) {
	refnerf_export_kernel<<<div_round_up(g.n_init, 128u), 128, 0, stream>>>(g.n_init, g.payload[0], (float4*)d_o_t, (float4*)d_d_alive);
	return (int)cudaGetLastError();
}

// first half of one iteration of NerfTracer::trace: compact, read back n_alive, generate the next network inputs.
// Returns n_alive; *n_steps_out = clamp(n_init / n_alive, 1, 8).
int refnerf_compact_and_generate(float* d_records, uint32_t n_extra, const float* d_extra_dims, uint32_t frame_id, uint32_t* n_steps_out, cudaStream_t stream) {
	const uint32_t cur = (g.dbi + 1) % 2, tmp = g.dbi % 2;
	++g.dbi;
	cudaMemsetAsync(g.counters, 0, sizeof(uint32_t), stream);
	linear_kernel(compact_kernel_nerf, 0, stream, g.n_alive, g.rgba[tmp], g.depth[tmp], g.payload[tmp], g.rgba[cur], g.depth[cur], g.payload[cur], g.rgba[2],
	              g.depth[2], g.payload[2], g.counters, g.counters + 1);
	cudaMemcpyAsync(&g.n_alive, g.counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
	cudaStreamSynchronize(stream);
	if (g.n_alive == 0) return 0;
	const uint32_t n_steps = tcnn::clamp(g.n_init / g.n_alive, 1u, 8u);
	*n_steps_out = n_steps;
	const uint32_t extra_stride = n_extra * sizeof(float);
	PitchedPtr<NerfCoordinate> input_data((NerfCoordinate*)d_records, 1, 0, extra_stride);
	BoundingSphere bs;
	linear_kernel(generate_next_nerf_network_inputs, 0, stream, g.n_alive, g.render_aabb, bs, Matrix3f::Identity(), g.train_aabb, g.focal, Vector3f(g.cam.col(2)),
	              g.payload[cur], input_data, n_steps, g.grid, 0u, g.cone, d_extra_dims, frame_id, 0.f, (ngp::TriangleBvhNode**)nullptr, (ngp::Triangle**)nullptr);
	return (int)g.n_alive;
}

// second half: composite with the network output (row-major / SoA, element stride out_stride)
int refnerf_composite(const float* d_records, uint32_t n_extra, const void* d_net_out, uint32_t out_stride, uint32_t n_steps, cudaStream_t stream) {
	const uint32_t cur = g.dbi % 2;
	PitchedPtr<NerfCoordinate> input_data((NerfCoordinate*)d_records, 1, 0, n_extra * sizeof(float));
	linear_kernel(composite_kernel_nerf, 0, stream, g.n_alive, out_stride, g.step, g.train_aabb, 0.f, 0, 0u, (const TrainingXForm*)nullptr, g.cam, g.focal, 1.0f,
	              g.rgba[cur], g.depth[cur], g.payload[cur], input_data, (const network_precision_t*)d_net_out, 16u, n_steps, ERenderMode::Shade, g.grid,
	              ENerfActivation::Logistic, ENerfActivation::Exponential, -1, g.min_T);
	g.step += n_steps;
	return (int)cudaGetLastError();
}

// after the loop: shade the finished rays into the frame buffer; returns n_hit
int refnerf_shade(float* d_frame_rgba, float* d_frame_depth, int linear_colors, cudaStream_t stream) {
	uint32_t n_hit = 0;
	cudaMemcpyAsync(&n_hit, g.counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
	cudaStreamSynchronize(stream);
	if (n_hit)
		linear_kernel(shade_kernel_nerf, 0, stream, n_hit, g.rgba[2], g.depth[2], g.payload[2], ERenderMode::Shade, (bool)linear_colors, (Array4f*)d_frame_rgba,
		              d_frame_depth);
	return (int)n_hit;
}


// ---- training kernels -------------------------------------------------------------------------------------------
static TrainingImageMetadata* g_meta = nullptr;
static TrainingXForm* g_xforms = nullptr;
static uint32_t g_n_images = 0;
static float* g_exposure = nullptr;

// per image: device pointers (pixels, seg, depth, rays: 4 per image), resolution (2), focal (2), principal point (2), xform (12), type
int refnerf_train_set_images(uint32_t n, const void* const* ptrs4, const uint32_t* res2, const float* focal2, const float* pp2, const float* xform12, const uint32_t* type) {
	std::vector<TrainingImageMetadata> meta(n);
	std::vector<TrainingXForm> xf(n);
	for (uint32_t k = 0; k < n; ++k) {
		meta[k].pixels = ptrs4[4 * k + 0];
		meta[k].seg_mask = ptrs4[4 * k + 1];
		meta[k].depth = (const float*)ptrs4[4 * k + 2];
		meta[k].rays = (const Ray*)ptrs4[4 * k + 3];
		meta[k].image_data_type = (EImageDataType)type[k];
		meta[k].resolution = Vector2i((int)res2[2 * k], (int)res2[2 * k + 1]);
		meta[k].focal_length = Vector2f(focal2[2 * k], focal2[2 * k + 1]);
		meta[k].principal_point = Vector2f(pp2[2 * k], pp2[2 * k + 1]);
		for (int c = 0; c < 4; ++c)
			for (int r = 0; r < 3; ++r) xf[k].start(r, c) = xform12[12 * k + c * 3 + r];
		xf[k].end = xf[k].start;
	}
	cudaFree(g_meta); cudaFree(g_xforms); cudaFree(g_exposure);
	cudaMalloc(&g_meta, sizeof(TrainingImageMetadata) * n);
	cudaMalloc(&g_xforms, sizeof(TrainingXForm) * n);
	cudaMalloc(&g_exposure, sizeof(float) * 3 * n);
	cudaMemcpy(g_meta, meta.data(), sizeof(TrainingImageMetadata) * n, cudaMemcpyHostToDevice);
	cudaMemcpy(g_xforms, xf.data(), sizeof(TrainingXForm) * n, cudaMemcpyHostToDevice);
	cudaMemset(g_exposure, 0, sizeof(float) * 3 * n);
	g_n_images = n;
	return (int)cudaGetLastError();
}

// generate_training_samples_nerf as train_nerf_step launches it (testbed_nerf.cu:3252-3284); counters = {rays, samples, compacted}
int refnerf_train_generate(uint32_t n_rays, uint32_t n_rays_total, uint32_t max_samples, uint64_t rng_seed, uint64_t rng_advances, const float* aabb6,
                           const uint8_t* d_bitfield, float cone_angle, int snap, const float* d_extra_dims, uint32_t n_extra, float* d_coords, uint32_t* d_counters,
                           uint32_t* d_ray_indices, float* d_rays, uint32_t* d_numsteps, cudaStream_t stream) {
	default_rng_t rng{rng_seed};
	for (uint64_t k = 0; k < rng_advances; ++k) rng.advance();
	cudaMemsetAsync(d_counters, 0, 3 * sizeof(uint32_t), stream);
	BoundingSphere bs;
	linear_kernel(generate_training_samples_nerf, 0, stream, n_rays, bb6(aabb6), bb6(aabb6), bs, max_samples, n_rays_total, rng, d_counters, d_counters + 1,
	              d_ray_indices, (Ray*)d_rays, d_numsteps, PitchedPtr<NerfCoordinate>((NerfCoordinate*)d_coords, 1, 0, n_extra * sizeof(float)), g_n_images, g_meta,
	              g_xforms, d_bitfield, false, (float*)nullptr, (bool)snap, false, cone_angle, (const float*)nullptr, Vector2i::Zero(), (const float*)nullptr,
	              (const float*)nullptr, (const float*)nullptr, Vector2i::Zero(), d_extra_dims, n_extra);
	return (int)cudaGetLastError();
}

// compute_loss_kernel_train_nerf as train_nerf_step launches it (testbed_nerf.cu:3298-3351) + the roll-over kernels (3353-3358)
int refnerf_train_loss(uint32_t n_rays, uint32_t n_rays_total, uint32_t target_batch, uint64_t rng_seed, uint64_t rng_advances, const float* aabb6, float loss_scale,
                       uint32_t padded_output_width, const float* bg3, int random_bg, const void* d_net_out, uint32_t* d_counters, const uint32_t* d_ray_indices,
                       const float* d_rays, uint32_t* d_numsteps, const float* d_coords, float* d_coords_compacted, uint32_t n_extra, void* d_dloss, float* d_loss,
                       int snap, const float* d_mean_density, float depth_lambda, float seg_lambda, float near_distance, int penalize_density, float beta,
                       cudaStream_t stream) {
	default_rng_t rng{rng_seed};
	for (uint64_t k = 0; k < rng_advances; ++k) rng.advance();
	const uint32_t extra_stride = n_extra * sizeof(float);
	linear_kernel(compute_loss_kernel_train_nerf, 0, stream, n_rays, bb6(aabb6), n_rays_total, rng, target_batch, d_counters, loss_scale, (int)padded_output_width,
	              (const float*)nullptr, (float*)nullptr, Vector2i::Zero(), ELossType::L2, Array3f(bg3[0], bg3[1], bg3[2]), EColorSpace::SRGB, (bool)random_bg, false,
	              g_n_images, g_meta, (const network_precision_t*)d_net_out, d_counters + 2, d_ray_indices, (const Ray*)d_rays, d_numsteps,
	              PitchedPtr<const NerfCoordinate>((NerfCoordinate*)d_coords, 1, 0, extra_stride), PitchedPtr<NerfCoordinate>((NerfCoordinate*)d_coords_compacted, 1, 0, extra_stride),
	              (network_precision_t*)d_dloss, ELossType::Huber, ELossType::L1, d_loss, false, (float*)nullptr, ENerfActivation::Logistic, ENerfActivation::Exponential,
	              (bool)snap, (float*)nullptr, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, Vector2i::Zero(), Vector2i::Zero(), (const float*)nullptr,
	              Vector2i::Zero(), (float*)nullptr, (float*)nullptr, d_mean_density, (const Eigen::Array3f*)g_exposure, (Eigen::Array3f*)nullptr, depth_lambda, seg_lambda,
	              near_distance, (bool)penalize_density, beta);
	const uint32_t floats_per_coord = sizeof(NerfCoordinate) / sizeof(float) + n_extra;
	fill_rollover_and_rescale<network_precision_t><<<n_blocks_linear(target_batch * padded_output_width), n_threads_linear, 0, stream>>>(target_batch, padded_output_width,
	                                                                                                                          d_counters + 2, (network_precision_t*)d_dloss);
	fill_rollover<float><<<n_blocks_linear(target_batch * floats_per_coord), n_threads_linear, 0, stream>>>(target_batch, floats_per_coord, d_counters + 2, d_coords_compacted);
	return (int)cudaGetLastError();
}

// ---- occupancy-grid maintenance: the kernels of update_density_grid_nerf / update_density_grid_mean_and_bitfield (testbed_nerf.cu:2774-2878) ----
int refnerf_grid_mark_untrained(uint32_t n_elements, float* d_grid, int clear_visible, cudaStream_t stream) {
	linear_kernel(mark_untrained_density_grid, 0, stream, n_elements, d_grid, g_n_images, g_meta, g_xforms, (bool)clear_visible);
	return (int)cudaGetLastError();
}

int refnerf_grid_generate(uint32_t n, uint64_t rng_seed, uint64_t rng_advances, uint32_t step, const float* aabb6, const float* d_grid, float* d_positions,
                          uint32_t* d_indices, uint32_t n_cascades, float thresh, cudaStream_t stream) {
	default_rng_t rng{rng_seed};
	for (uint64_t k = 0; k < rng_advances; ++k) rng.advance();
	static_assert(sizeof(NerfPosition) == 12, "NerfPosition is a bare position in this build");
	linear_kernel(generate_grid_samples_nerf_nonuniform, 0, stream, n, rng, step, bb6(aabb6), d_grid, (NerfPosition*)d_positions, d_indices, n_cascades, thresh);
	return (int)cudaGetLastError();
}

int refnerf_grid_splat_ema(uint32_t n_samples, const uint32_t* d_indices, const void* d_logits_f16, uint32_t n_elements, float decay, uint32_t ema_step, float* d_grid,
                           float* d_grid_tmp, cudaStream_t stream) {
	cudaMemsetAsync(d_grid_tmp, 0, sizeof(float) * n_elements, stream);
	linear_kernel(splat_grid_samples_nerf_max_nearest_neighbor, 0, stream, n_samples, d_indices, (const network_precision_t*)d_logits_f16, d_grid_tmp,
	              ENerfActivation::Logistic, ENerfActivation::Exponential);
	linear_kernel(ema_grid_samples_nerf, 0, stream, n_elements, decay, ema_step, d_grid, d_grid_tmp);
	return (int)cudaGetLastError();
}

int refnerf_grid_mean_and_bitfield(float* d_grid, uint8_t* d_bitfield, float* d_mean, uint32_t n_cascades, cudaStream_t stream) {
	const uint32_t n_elements = NERF_GRIDSIZE() * NERF_GRIDSIZE() * NERF_GRIDSIZE();
	cudaMemsetAsync(d_mean, 0, sizeof(float), stream);
	reduce_sum(d_grid, [n_elements] __device__(float val) { return fmaxf(val, 0.f) / (n_elements); }, d_mean, n_elements, stream);
	linear_kernel(grid_to_bitfield, 0, stream, n_elements / 8 * NERF_CASCADES(), n_elements / 8 * n_cascades, d_grid, d_bitfield, d_mean);
	for (uint32_t level = 1; level < NERF_CASCADES(); ++level)
		linear_kernel(bitfield_max_pool, 0, stream, n_elements / 64, d_bitfield + grid_mip_offset(level - 1) / 8, d_bitfield + grid_mip_offset(level) / 8);
	return (int)cudaGetLastError();
}

}  // extern "C"
