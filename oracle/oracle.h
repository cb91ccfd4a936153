/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference algorithm for the INSTA
 * deformation + hash-encoding hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libinsta_b200.so) never does.
 *
 * Parity status: PINNED.  Every stage-1 function is checked bit-for-bit against the reference's own host code
 * (oracle/_ref/libref_host.so, built from /root/reference by oracle/Makefile) in tests/test_oracle_vs_ref.py and
 * against committed golden vectors generated from it (tests/golden/, tests/golden/make_golden.py).  The hash-grid
 * functions have no host implementation in the reference; they are pinned on the GPU box against the reference's
 * own CUDA kernels (oracle/_ref/libref_cuda.so, tests/test_gpu_vs_ref_cuda.py).
 *
 * All floating point here is plain IEEE fp32, left-to-right, never contracted (the build uses
 * -ffp-contract=off): that is what the reference's host build computes (CMakeLists.txt:71 passes no FMA flag).
 */
#ifndef INSTA_ORACLE_H
#define INSTA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same 48-byte layout as ngp::Triangle (triangle.cuh:234-238). */
typedef struct orc_tri {
	float a[3], b[3], c[3];
	float scaling;
	uint32_t id;
	uint32_t mask_id;
} orc_tri;

/* Same 32-byte layout as ngp::TriangleBvhNode (triangle_bvh.cuh:29-33). */
typedef struct orc_node {
	float mn[3], mx[3];
	int32_t left_idx, right_idx;
} orc_node;

int orc_max_threads(void);
/* 0 = reference HOST build arithmetic (default), 1 = reference DEVICE build (nvcc -fmad contraction); see oracle.cpp */
void orc_set_arith(int mode);
int orc_get_arith(void);

/* ---- per-frame mesh preparation (tiny_mesh.cu:134-178, 316-371; masking.cpp:118-148) ---- */
void orc_build_triangles(const float* verts, const uint32_t* faces, uint32_t n_faces, const uint32_t* mask_ids,
                         const orc_tri* canonical /* nullable */, orc_tri* out);
void orc_mesh_aabb(const float* verts, uint32_t n_verts, float* aabb6);
void orc_adjacency(const uint32_t* faces, uint32_t n_faces, uint32_t n_verts, int32_t* out3);
/* region_ids/region_offsets/region_verts: regions in ascending FlameMasks order (std::map iteration). */
void orc_triangle_masks(const uint32_t* faces, uint32_t n_faces, const int32_t* region_ids, const uint32_t* region_offsets,
                        const uint32_t* region_verts, uint32_t n_regions, uint32_t* out_mask);

/* ---- BVH (triangle_bvh.cu:548-631, bvh.cuh:48-96) ---- */
int orc_bvh_build(orc_tri* tris_inout, int n_tris, int leaf_size, orc_node* nodes_out, int max_nodes);
void orc_closest_triangle(const float* pts, int n, const orc_node* nodes, const orc_tri* tris, int* out_idx, float* out_dist,
                          int n_threads);
/* closest_triangle started from max_distance_sq = init_best[i] instead of FLT_MAX (bvh.cuh:48) */
/* TriangleBvh4::ray_intersect (ngp src/triangle_bvh.cu:271-312): first hit of n rays, BVH-order index (-1 = none within MAX_DIST = 10) and ray parameter */
void orc_ray_intersect(const float* ro, const float* rd, int n, const orc_node* nodes, const orc_tri* tris, int* out_idx, float* out_t);
void orc_closest_triangle_bounded(const float* pts, int n, const orc_node* nodes, const orc_tri* tris, const float* init_best, int* out_idx, int n_threads);
/* diagnostic: traversal statistics (see oracle.cpp) */
void orc_closest_stats(const float* pts, int n, const orc_node* nodes, const orc_tri* tris, double* out8);
void orc_closest_stats_bounded(const float* pts, int n, const orc_node* nodes, const orc_tri* tris, const float* init_best, double* out8);
/* The product's cylinder bounds restated (checker of the pruning argument; not in the reference): 8 halves per node */
void orc_cyl_build(const orc_node* nodes, int n_nodes, const orc_tri* tris, uint16_t* out);
void orc_closest_triangle_cyl(const float* pts, int n, const orc_node* nodes, const orc_tri* tris, const uint16_t* cyl, const float* init_best, int* out_idx,
                              double* out_stats4, int check);
float orc_triangle_distance_sq(const orc_tri* t, const float* p);
float orc_bbox_distance_sq(const float* mn, const float* mx, const float* p);
void orc_tri2projection(const orc_tri* t, float* out16_colmajor);

/* ---- deformation (core.cu:442-579, affine.cuh:26-116); argument meaning as ref_warp_cpu ---- */
void orc_warp(float* samples, uint32_t n, uint32_t stride, int32_t frame_id, uint32_t n_meshes, const orc_node* const* nodes_ptrs,
              const orc_tri* const* tris_bvh_ptrs, const orc_tri* const* tris_topo_ptrs, const orc_tri* tris_canon,
              const int32_t* adjacency, int disable_interp, const float* aabb6, uint32_t extra_dims, int use_geo,
              const int32_t* frame_ids, int32_t* out_tri, int32_t* out_bvh_idx, int n_threads);
/* frame id drawn per point by warp_random_samples (core.cu:544-549): pcg32(seed 42 default stream), advance(i*(n_meshes-1)) */
void orc_random_frame_ids(uint32_t n, uint32_t n_meshes, uint64_t rng_state, uint64_t rng_inc, int32_t* out);

/* ---- multiresolution hash grid (tcnn grid.h:82-117, 138-310, 312-435, 874-939; common_device.h:415-431) ---- */
/* offsets: n_levels+1 entries (in grid entries, not features); scales/res: n_levels. Host evaluation of the ctor. */
void orc_grid_layout(uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale,
                     uint32_t* offsets, float* scales, uint32_t* resolutions);
/* slot index of the 8 corners of each (sample, level): out[(i*n_levels + l)*8 + corner], entry units incl. level offset */
void orc_grid_slots(uint32_t n, const float* pos, uint32_t pos_stride, uint32_t n_levels, const uint32_t* offsets,
                    const float* scales, const uint32_t* resolutions, uint32_t* out_slots, float* out_weights /* nullable */);
/* F = 8 features per level, fp16 table, fp16 accumulate; out_soa[i + (l*8+f)*n] (fp16 bits) */
void orc_grid_forward(uint32_t n, const float* pos, uint32_t pos_stride, uint32_t n_levels, const uint32_t* offsets,
                      const float* scales, const uint32_t* resolutions, const uint16_t* table_f16, uint16_t* out_soa_f16,
                      int n_threads);
/* exact (double) sum of the fp16-rounded per-corner contributions, plus hit counts per feature slot */
void orc_grid_backward(uint32_t n, const float* pos, uint32_t pos_stride, uint32_t n_levels, const uint32_t* offsets,
                       const float* scales, const uint32_t* resolutions, const uint16_t* dLdy_soa_f16, double* grad_out,
                       uint32_t* hits_out /* nullable, per entry */);

uint32_t orc_f16_selfcheck(uint32_t n, uint32_t seed);
uint16_t orc_f32_to_f16(float x);
float orc_f16_to_f32(uint16_t h);

#ifdef __cplusplus
}
#endif
#endif
