// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Thin C-ABI harness around the REFERENCE's own host-callable code, compiled from the sources where
// they lie under /root/reference (see oracle/Makefile; output: oracle/_ref/libref_host.so).
// It exposes, unmodified:
//   * ngp::TriangleBvh::make()/build()          instant-ngp/src/triangle_bvh.cu:548-631
//   * closest_triangle()                        instant-ngp/include/neural-graphics-primitives/bvh.cuh:48-96
//   * ray_intersect()                           instant-ngp/include/neural-graphics-primitives/bvh.cuh:98-144
//   * ngp::Triangle::{distance_sq,tbn,centroid} instant-ngp/include/neural-graphics-primitives/triangle.cuh:43-107,220
//   * tri2projection / interpolate_transformation / transform_point / transform_direction
//                                               include/rta/affine.cuh:26-116
// The only glue written here is the per-sample driver `ref_warp_cpu`, which calls those functions in
// the order src/core.cu:442-526 (`warp`) does.  interpolate_transformation is __device__-only in
// the reference because of __expf; for the host build it is re-tagged __host__ __device__ and __expf
// maps to expf (documented tolerance source, SURVEY.md 7 hard part 2).
//
// No GPU is needed: TriangleBvh::build() finishes the host tree before it tries the device upload
// (triangle_bvh.cu:628); the upload exception is swallowed and nodes() is used.

#include <neural-graphics-primitives/triangle_bvh.cuh>
#include <neural-graphics-primitives/bvh.cuh>

__host__ __device__ inline float ref_expf_hd(float x) {
#ifdef __CUDA_ARCH__
	return __expf(x);
#else
	return expf(x);
#endif
}
#undef __device__
#define __device__ __location__(device) __location__(host)
#define __expf(x) ref_expf_hd(x)
#include <rta/affine.cuh>
#undef __expf
#undef __device__
#define __device__ __location__(device)

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

static_assert(sizeof(ngp::Triangle) == 48, "Triangle ABI");
static_assert(sizeof(ngp::TriangleBvhNode) == 32, "Node ABI");

extern "C" {

int ref_sizeof_triangle() { return (int)sizeof(ngp::Triangle); }
int ref_sizeof_node() { return (int)sizeof(ngp::TriangleBvhNode); }
int ref_max_threads() {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

// tris: n x 48 B records, permuted IN PLACE into BVH order (as TinyMesh::build_tries relies on).
// Returns the node count (or -needed if max_nodes is too small).
int ref_bvh_build(void* tris_inout, int n_tris, int leaf_size, void* nodes_out, int max_nodes) {
	std::vector<ngp::Triangle> tris(n_tris);
	memcpy(tris.data(), tris_inout, sizeof(ngp::Triangle) * n_tris);
	auto bvh = ngp::TriangleBvh::make();
	try {
		bvh->build(tris, (uint32_t)leaf_size);
	} catch (const std::exception&) {
		// device upload failed (no GPU here); the host tree is complete.
		cudaGetLastError();
	}
	// node count is not exposed; recover it by walking the tree.
	const ngp::TriangleBvhNode* nodes = bvh->nodes();
	int n_nodes = 1;
	{
		std::vector<int> st{0};
		while (!st.empty()) {
			int i = st.back(); st.pop_back();
			if (nodes[i].left_idx >= 0) {
				if (nodes[i].right_idx > n_nodes) n_nodes = nodes[i].right_idx;
				for (int c = nodes[i].left_idx; c < nodes[i].right_idx; ++c) st.push_back(c);
			}
		}
	}
	if (n_nodes > max_nodes) return -n_nodes;
	memcpy(nodes_out, nodes, sizeof(ngp::TriangleBvhNode) * n_nodes);
	memcpy(tris_inout, tris.data(), sizeof(ngp::Triangle) * n_tris);
	return n_nodes;
}

// TriangleBvh4::ray_intersect (instant-ngp/src/triangle_bvh.cu:271-312, what raycast_flame's trace_bvh runs) is local to its translation unit; INSTA's header
// copy of the same loop (include/neural-graphics-primitives/bvh.cuh:98-144) is callable and takes the start value of mint as a parameter:
// 10 = MAX_DIST (triangle_bvh.cu:49).  Triangle::ray_intersect and BoundingBox::ray_intersect are the reference's own (triangle.cuh:66-81, bounding_box.cuh:163-205).
void ref_ray_intersect(const float* ro, const float* rd, int n, const void* nodes, const void* tris, int* out_idx, float* out_t) {
	const auto* N = (const ngp::TriangleBvhNode*)nodes;
	const auto* T = (const ngp::Triangle*)tris;
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; ++i) {
		auto r = ray_intersect(Vector3f(ro[3 * i], ro[3 * i + 1], ro[3 * i + 2]), Vector3f(rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]), N, T, 10.0f);
		out_idx[i] = r.first;
		if (out_t) out_t[i] = r.second;
	}
}

void ref_closest_triangle(const float* pts, int n, const void* nodes, const void* tris, int* out_idx, float* out_dist, int n_threads) {
	const auto* N = (const ngp::TriangleBvhNode*)nodes;
	const auto* T = (const ngp::Triangle*)tris;
#pragma omp parallel for schedule(static) num_threads(n_threads)
	for (int i = 0; i < n; ++i) {
		auto r = closest_triangle(Vector3f(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), N, T);
		out_idx[i] = r.first;
		if (out_dist) out_dist[i] = r.second;
	}
}

void ref_triangle_distance_sq(const float* pts, int n, const void* tri, float* out) {
	const auto* T = (const ngp::Triangle*)tri;
	for (int i = 0; i < n; ++i) out[i] = T->distance_sq(Vector3f(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
}

void ref_bbox_distance_sq(const float* pts, int n, const float* mn, const float* mx, float* out) {
	ngp::BoundingBox bb;
	bb.min = Vector3f(mn[0], mn[1], mn[2]);
	bb.max = Vector3f(mx[0], mx[1], mx[2]);
	for (int i = 0; i < n; ++i) out[i] = bb.distance_sq(Vector3f(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
}

// 4x4 column-major out
void ref_tri2projection(const void* tri, float* out16) {
	Matrix4f M = tri2projection(Vector3f::Zero(), *(const ngp::Triangle*)tri);
	memcpy(out16, M.data(), 64);
}

// Host replay of `warp` (src/core.cu:442-526) / `warp_random_samples` (528-579, when stride == 3 and
// frame ids are supplied in frame_ids) using the reference's own functions.
//   samples      : n x stride floats, mutated in place exactly as the kernel does
//   frame_id     : >= 0 fixed frame (render_deformed), -1 = per-sample holder (samples[j+4])
//   frame_ids    : optional explicit per-sample frame (xyz path), overrides everything when non-null
//   out_tri      : optional, topology id of the winning triangle (-1 skipped)
//   out_bvh_idx  : optional, index in BVH order
void ref_warp_cpu(float* samples, uint32_t n, uint32_t stride, int32_t frame_id, uint32_t n_meshes,
                  const void* const* nodes_ptrs, const void* const* tris_bvh_ptrs, const void* const* tris_topo_ptrs,
                  const void* tris_canon, const int* adjacency, int disable_interp, const float* aabb6,
                  uint32_t extra_dims, int use_geo, const int32_t* frame_ids,
                  int32_t* out_tri, int32_t* out_bvh_idx, int n_threads) {
	ngp::BoundingBox render_bbox;
	render_bbox.min = Vector3f(aabb6[0], aabb6[1], aabb6[2]);
	render_bbox.max = Vector3f(aabb6[3], aabb6[4], aabb6[5]);
	const auto* canon = (const ngp::Triangle*)tris_canon;
	const bool xyz_only = stride == 3;
#pragma omp parallel for schedule(static) num_threads(n_threads)
	for (int64_t i = 0; i < (int64_t)n; ++i) {
		size_t j = (size_t)i * stride;
		if (out_tri) out_tri[i] = -1;
		if (out_bvh_idx) out_bvh_idx[i] = -1;
		uint32_t f;
		if (frame_ids) {
			f = (uint32_t)frame_ids[i];
		} else {
			if (samples[j + 4] < 0.f) continue;
			f = frame_id < 0 ? uint32_t(samples[j + 4]) : (uint32_t)frame_id;
		}
		if (f >= n_meshes) continue;
		Vector3f point = {samples[j + 0], samples[j + 1], samples[j + 2]};
		if (!xyz_only && !render_bbox.contains(point)) continue;

		const auto* nodes = (const ngp::TriangleBvhNode*)nodes_ptrs[f];
		const auto* tbvh = (const ngp::Triangle*)tris_bvh_ptrs[f];
		const auto* ttopo = (const ngp::Triangle*)tris_topo_ptrs[f];
		auto nn = closest_triangle(point, nodes, tbvh);
		int id = nn.first;
		if (id == -1) continue;
		const ngp::Triangle& deformed = tbvh[id];
		const ngp::Triangle& canonical = canon[deformed.id];
		int mask_id = deformed.mask_id;
		if (out_tri) out_tri[i] = (int32_t)deformed.id;
		if (out_bvh_idx) out_bvh_idx[i] = id;

		Matrix4f world_to_tri, tri_to_canon;
		if (mask_id != 2 || disable_interp) {
			world_to_tri = tri2projection(point, deformed).inverse();
			tri_to_canon = tri2projection(point, canonical);
		} else {
			world_to_tri = interpolate_transformation(point, adjacency, ttopo, deformed, true);
			tri_to_canon = interpolate_transformation(point, adjacency, canon, canonical, false);
		}
		Vector3f cp = transform_point(point, world_to_tri, tri_to_canon);
		samples[j + 0] = cp.x(); samples[j + 1] = cp.y(); samples[j + 2] = cp.z();
		if (xyz_only) continue;

		Vector3f direction = {samples[j + 8], samples[j + 9], samples[j + 10]};
		Vector3f un = direction * 2.0f - Vector3f::Ones();                 // unwarp_direction, core.cu:66
		Vector3f cd = transform_direction(un, world_to_tri, tri_to_canon);
		cd = (cd + Vector3f::Ones()) * 0.5f;                               // warp_direction, core.cu:62
		samples[j + 8] = cd.x(); samples[j + 9] = cd.y(); samples[j + 10] = cd.z();
		if (mask_id != 7) {
			for (uint32_t m = 0; m < extra_dims; ++m) samples[j + stride - extra_dims + m] = 1.f;
		}
		if (use_geo) {
			float cosine = (point - deformed.centroid()).normalized().dot(deformed.normal());
			samples[j + stride - 1] = cosine;
		}
	}
}

} // extern "C"
