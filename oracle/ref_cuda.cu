// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI launcher around the REFERENCE's own CUDA kernels, compiled for sm_100 from the sources where they lie
// (oracle/Makefile target `refcuda`; output oracle/_ref/libref_cuda.so):
//   * `warp`, `warp_random_samples`            src/core.cu:442-579, extracted at build time into
//                                              oracle/_ref/warp_extracted.inc (that TU also contains the GUI)
//   * kernel_grid<__half,3,8>                  tcnn include/tiny-cuda-nn/encodings/grid.h:138-310
//   * kernel_grid_backward<__half,__half,3,8,2>                                 grid.h:312-435
// with the launch shapes the reference uses (linear_kernel: 128 threads; grid.h:974-975, 1071-1074).
// Used on the GPU box for (a) bit-parity of slot indices / encodings / triangle ids against the reference's CUDA
// path and (b) the "reference CUDA path" timing that bench.py reports beside ours.
#include <neural-graphics-primitives/common.h>
#include <neural-graphics-primitives/common_device.cuh>
#include <neural-graphics-primitives/random_val.cuh>
#include <neural-graphics-primitives/triangle_bvh.cuh>
#include <neural-graphics-primitives/bvh.cuh>
#include <rta/affine.cuh>

#include <tiny-cuda-nn/common.h>
#include <tiny-cuda-nn/encodings/grid.h>

using namespace Eigen;
using namespace tcnn;

namespace ngp {
struct TrainingImageMetadata;
struct TrainingXForm;
}  // namespace ngp

// the few file-scope helpers of src/core.cu the kernels use (core.cu:42, 62-68)
constexpr int MAX_UINT = std::numeric_limits<uint32_t>::max();
__host__ __device__ Vector3f warp_direction(const Vector3f& dir) { return (dir + Vector3f::Ones()) * 0.5f; }
__device__ Vector3f unwarp_direction(const Vector3f& dir) { return dir * 2.0f - Vector3f::Ones(); }

#include "warp_extracted.inc"

static GridOffsetTable make_offsets(uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale) {
	// the ctor's table (grid.h:898-922) -- GridEncodingTemplated itself cannot be instantiated without linking all of tcnn
	GridOffsetTable t{};
	uint32_t offset = 0;
	for (uint32_t i = 0; i < n_levels; ++i) {
		const float scale = exp2f(i * std::log2(per_level_scale)) * base_resolution - 1.0f;
		const uint32_t resolution = (uint32_t)(ceilf(scale)) + 1;
		uint32_t max_params = std::numeric_limits<uint32_t>::max() / 2;
		uint32_t params_in_level = std::pow((float)resolution, 3) > (float)max_params ? max_params : powi(resolution, 3);
		params_in_level = next_multiple(params_in_level, 8u);
		params_in_level = std::min(params_in_level, (1u << log2_hashmap_size));
		t.data[i] = offset;
		offset += params_in_level;
	}
	t.data[n_levels] = offset;
	t.size = n_levels + 1;
	return t;
}

// glue kernels that call the reference's own device functions directly (for bit-level checks of ids and distances)
__global__ void refcuda_closest_kernel(uint32_t n, const float* __restrict__ pts, const ngp::TriangleBvhNode* __restrict__ nodes,
                                       const ngp::Triangle* __restrict__ tris, int* __restrict__ idx, float* __restrict__ dist) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	auto r = closest_triangle(Vector3f(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), nodes, tris);
	idx[i] = r.first;
	dist[i] = r.second;
}
__global__ void refcuda_distance_kernel(uint32_t n, const float* __restrict__ pts, const ngp::Triangle* __restrict__ tri, const ngp::TriangleBvhNode* __restrict__ node,
                                        float* __restrict__ out_tri, float* __restrict__ out_box) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const Vector3f p(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
	out_tri[i] = tri->distance_sq(p);
	out_box[i] = node->bb.distance_sq(p);
}

extern "C" {

int refcuda_closest_triangle(uint32_t n, const float* d_pts, const void* d_nodes, const void* d_tris, int* d_idx, float* d_dist, cudaStream_t stream) {
	linear_kernel(refcuda_closest_kernel, 0, stream, n, d_pts, (const ngp::TriangleBvhNode*)d_nodes, (const ngp::Triangle*)d_tris, d_idx, d_dist);
	return (int)cudaGetLastError();
}
int refcuda_distances(uint32_t n, const float* d_pts, const void* d_tri, const void* d_node, float* d_out_tri, float* d_out_box, cudaStream_t stream) {
	linear_kernel(refcuda_distance_kernel, 0, stream, n, d_pts, (const ngp::Triangle*)d_tri, (const ngp::TriangleBvhNode*)d_node, d_out_tri, d_out_box);
	return (int)cudaGetLastError();
}

int refcuda_warp(float* d_samples, uint32_t n, uint32_t stride, int32_t frame_id, uint32_t n_meshes, void** d_nodes_ptrs, void** d_tris_bvh_ptrs,
                 void** d_tris_topo_ptrs, void* d_canon, int* d_adjacency, int disable_interp, const float* aabb6, uint32_t extra_dims, int use_geo,
                 cudaStream_t stream) {
	ngp::BoundingBox bb;
	bb.min = Vector3f(aabb6[0], aabb6[1], aabb6[2]);
	bb.max = Vector3f(aabb6[3], aabb6[4], aabb6[5]);
	linear_kernel(warp, 0, stream, n, d_samples, (ngp::TriangleBvhNode**)d_nodes_ptrs, (ngp::Triangle**)d_tris_bvh_ptrs,
	              (ngp::Triangle**)d_tris_topo_ptrs, (ngp::Triangle*)d_canon, stride, frame_id < 0 ? MAX_UINT : frame_id, n_meshes, (bool)use_geo,
	              (uint32_t)use_geo, bb, (const ngp::TrainingImageMetadata*)nullptr, (const ngp::TrainingXForm*)nullptr, default_rng_t{42},
	              (float*)nullptr, extra_dims, false, d_adjacency, (bool)disable_interp, false);
	return (int)cudaGetLastError();
}

int refcuda_warp_random_samples(float* d_xyz, uint32_t n, uint32_t n_meshes, void** d_nodes_ptrs, void** d_tris_bvh_ptrs, void** d_tris_topo_ptrs,
                                void* d_canon, int* d_adjacency, int disable_interp, cudaStream_t stream) {
	linear_kernel(warp_random_samples, 0, stream, n, d_xyz, (ngp::TriangleBvhNode**)d_nodes_ptrs, (ngp::Triangle**)d_tris_bvh_ptrs,
	              (ngp::Triangle**)d_tris_topo_ptrs, (ngp::Triangle*)d_canon, default_rng_t{42}, n_meshes, d_adjacency, (bool)disable_interp);
	return (int)cudaGetLastError();
}

int refcuda_grid_n_entries(uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale) {
	return (int)make_offsets(n_levels, log2_hashmap_size, base_resolution, per_level_scale).data[n_levels];
}

// forward_impl's launch (grid.h:974-1000), SoA output
int refcuda_grid_forward(uint32_t n, const float* d_pos, uint32_t pos_stride, const void* d_grid, void* d_out_soa, uint32_t n_levels,
                         uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale, cudaStream_t stream) {
	const GridOffsetTable offs = make_offsets(n_levels, log2_hashmap_size, base_resolution, per_level_scale);
	static constexpr uint32_t N_THREADS_HASHGRID = 512;
	const dim3 blocks = {div_round_up(n, N_THREADS_HASHGRID), n_levels, 1};
	kernel_grid<__half, 3, 8><<<blocks, N_THREADS_HASHGRID, 0, stream>>>(n, n_levels * 8, offs, base_resolution, std::log2(per_level_scale), 0.f, 1000.f,
	                                                                      nullptr, InterpolationType::Linear, GridType::Hash, (const __half*)d_grid,
	                                                                      MatrixView<const float>(d_pos, 1, pos_stride), (__half*)d_out_soa, nullptr);
	return (int)cudaGetLastError();
}

// backward_impl's launch (grid.h:1067-1093), SoA gradients, fp16 table zeroed first
int refcuda_grid_backward(uint32_t n, const float* d_pos, uint32_t pos_stride, const void* d_dLdy_soa, void* d_grad, uint32_t n_levels,
                          uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale, cudaStream_t stream) {
	const GridOffsetTable offs = make_offsets(n_levels, log2_hashmap_size, base_resolution, per_level_scale);
	cudaMemsetAsync(d_grad, 0, (size_t)offs.data[n_levels] * 8 * sizeof(__half), stream);
	static constexpr uint32_t N_THREADS_HASHGRID = 256;
	static constexpr uint32_t N_FEATURES_PER_THREAD = 2;
	const dim3 blocks = {div_round_up(n * 8 / N_FEATURES_PER_THREAD, N_THREADS_HASHGRID), n_levels, 1};
	kernel_grid_backward<__half, __half, 3, 8, N_FEATURES_PER_THREAD><<<blocks, N_THREADS_HASHGRID, 0, stream>>>(
	    n, n_levels * 8, offs, base_resolution, std::log2(per_level_scale), 1000.f, nullptr, false, InterpolationType::Linear, GridType::Hash,
	    (__half*)d_grad, MatrixView<const float>(d_pos, 1, pos_stride), (const __half*)d_dLdy_soa);
	return (int)cudaGetLastError();
}

}  // extern "C"
