// Internal definitions shared by the translation units of libinsta_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/insta_b200.h"

namespace insta {

// ---- error plumbing: status codes across the ABI, message kept per thread ---------------------------------
void set_error(const std::string& msg);
#define INSTA_FAIL(code, msg)            \
	do {                                 \
		::insta::set_error(msg);         \
		return (code);                   \
	} while (0)
#define INSTA_CUDA(expr)                                                                                           \
	do {                                                                                                           \
		cudaError_t e__ = (expr);                                                                                  \
		if (e__ != cudaSuccess) {                                                                                  \
			::insta::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
			                   std::to_string(__LINE__) + ")");                                                    \
			return INSTA_ERR_CUDA;                                                                                 \
		}                                                                                                          \
	} while (0)

// ---- host-side mesh records -------------------------------------------------------------------------------
// Byte-compatible with ngp::Triangle (ngp triangle.cuh:234-238) so that a caller may also hand over the
// reference's own `triangles_gpu` buffers.
struct Tri {
	float a[3], b[3], c[3];
	float scaling;
	uint32_t id;
	uint32_t mask_id;
};
static_assert(sizeof(Tri) == 48, "Tri must match ngp::Triangle");

// Reference-layout BVH node (ngp triangle_bvh.cuh:29-33); used on the host during the build only.
struct Node {
	float mn[3], mx[3];
	int32_t left_idx, right_idx;
};

// Device BVH node: the four children of one inner node of the reference's 4-ary tree, boxes stored SoA so that each
// plane is one 128-bit load and the whole quad is exactly one 128-byte line.
//   ref[k] >= 0 : index of the child's own Quad
//   ref[k] <  0 : leaf, -(ref+1) = (first_triangle << 3) | count   (count <= 7)
struct __align__(128) Quad {
	float mnx[4], mny[4], mnz[4], mxx[4], mxy[4], mxz[4];
	int32_t code[4];  // the same references as 16-bit stack codes (bit 15 = leaf; quad index | first << 2 | count), -1 = n/a
	int32_t ref[4];   // last, so that the first 112 bytes are all a compact-code traversal reads (deform_pf_kernel stages exactly those)
};
static_assert(sizeof(Quad) == 128, "Quad must be one cache line");

// Per-triangle transform record (topology order): M = [ TBN * (scaling if mask_id == 1 else 1) | a ] exactly as
// tri2projection builds it (include/rta/affine.cuh:26-39), plus the centroid used by the neighbour blend weights.
struct __align__(16) TriXf {
	float m[9];      // column-major 3x3
	float a[3];      // translation (vertex a)
	float c[3];      // centroid
	uint32_t mask_id;
};
static_assert(sizeof(TriXf) == 64, "TriXf is four 128-bit loads");

// Tight bounds of the four children of one quad (deform.cu, "cylinder bounds"): per child the unit axis n (fp16, rounded away from zero), the slab
// centre / half width along n and the radius around the axis through the centre of the child's BOX, the last three in doubled units.  A child is a
// leaf or an inner node; the records of a frame form a block of their own (FrameDev::cyl), kCylStride bytes per quad.
struct __align__(16) QuadCyl {
	uint16_t child[4][8];  // fp16 bits: nx, ny, nz, zc2, h2, r2, 0, 0
};
static_assert(sizeof(QuadCyl) == 64, "QuadCyl is four 128-bit chunks");
constexpr uint32_t kCylStride = 64;  // packed: with 1,700 frames in a batch the L2 footprint per frame decides the traversal kernel's time

struct FrameDev {
	const Quad* quads;
	const Tri* tris;     // BVH order
	const TriXf* xf;     // topology order
	const unsigned char* cyl;  // nullable: QuadCyl per quad at kCylStride
};

struct HostBvh {
	std::vector<Node> nodes;   // reference layout
	std::vector<Quad> quads;   // device layout
	std::vector<Tri> tris;     // BVH order
	std::vector<uint32_t> child_ranges;  // per quad and child: [begin, end) of the child's subtree in `tris` (depends on the triangle count only, not on the vertices)
};

// mesh_host.cpp
bool load_obj(const char* path, std::vector<float>& verts, std::vector<float>& colors, std::vector<uint32_t>& faces, std::string& err);
void build_triangles(const float* verts, const uint32_t* faces, uint32_t n_faces, const uint32_t* mask_ids, const Tri* canonical, Tri* out);
void build_adjacency(const uint32_t* faces, uint32_t n_faces, uint32_t n_verts, std::vector<int32_t>& out3);
void masks_from_regions(const uint32_t* faces, uint32_t n_faces, const std::vector<std::pair<int, std::vector<uint32_t>>>& regions,
                        std::vector<uint32_t>& out);
void mesh_aabb(const float* verts, uint32_t n_verts, float* aabb6);
void build_bvh(std::vector<Tri>& tris, uint32_t leaf_size, HostBvh& out);

// ---- hash grid level descriptors (kernel parameter; <= 16 levels) --------------------------------------------
constexpr int kMaxLevels = 16;
struct GridLevels {
	uint32_t n_levels;
	uint32_t offset[kMaxLevels + 1];  // entries
	uint32_t size[kMaxLevels];        // hashmap_size of the level (entries)
	uint32_t res[kMaxLevels];         // grid_resolution
	uint32_t mul1[kMaxLevels];        // dense stride of y (0 when the reference's loop stopped before it)
	uint32_t mul2[kMaxLevels];        // dense stride of z (uint32 wrap-around preserved)
	uint32_t hashed[kMaxLevels];      // 1 = fast_hash
	float scale[kMaxLevels];
};

// One training image on the device (ngp::TrainingImageMetadata + TrainingXForm, ngp nerf_loader.h:48-64)
struct ImageDev {
	const void* pixels;
	const void* seg;
	const float* depth;
	const float* rays;  // 6 floats per pixel (origin, unnormalised direction) or null
	uint32_t res[2];
	float focal[2], pp[2];
	float xform[12];
	uint32_t type;  // EImageDataType: 1 = Byte, 2 = Half, 3 = Float (ngp common.h)
	uint32_t pad;
};

}  // namespace insta

struct insta_ctx {
	insta_params p{};
	int device = 0;
	int sm_count = 0;
	// canonical
	uint32_t n_verts = 0, n_faces = 0;
	std::vector<float> canon_verts;
	std::vector<uint32_t> faces, mask_ids;
	std::vector<int32_t> adjacency;
	std::vector<insta::Tri> canon_tris;  // topology order
	float render_aabb[6] = {0, 0, 0, 1, 1, 1};
	float aabb[6] = {0, 0, 0, 1, 1, 1};
	uint32_t* d_faces = nullptr;
	uint32_t* d_mask_ids = nullptr;
	int32_t* d_adjacency = nullptr;
	float* d_canon_area = nullptr;
	insta::TriXf* d_canon_xf = nullptr;
	// frames
	struct Frame {
		void* slab = nullptr;
		insta::FrameDev dev{};
		uint32_t n_nodes = 0, n_quads = 0;
		std::vector<uint32_t> bvh_order_ids;
		float* d_verts = nullptr;          // inside the frame's slab (insta_replace_frames / insta_refit_frames write it again)
		uint32_t* d_order = nullptr;       // refit only: topology id of every BVH-order triangle (allocated on first refit)
	};
	std::vector<Frame> frames;
	insta::FrameDev* d_frame_table = nullptr;
	uint32_t frame_table_cap = 0;
	bool frame_table_dirty = false;
	uint32_t stack_need = 0;          // traversal stack entries a query can need for the loaded topology
	bool compact_entries = false;     // every reference of the loaded topology fits the 16-bit stack code
	bool force_legacy_deform = false; // INSTA_DEFORM_KERNEL=legacy: thread-per-query kernel everywhere (A/B measurements)
	bool force_wq_deform = false;     // INSTA_DEFORM_KERNEL=wq: warp-queue kernel everywhere
	bool no_prefetch_deform = false;  // INSTA_DEFORM_KERNEL=wq1: round-1 warp-queue kernel (deform_wq_kernel) instead of deform_pf_kernel (A/B measurements)
	bool layout_uniform = false;      // every resident frame has its triangle / transform tables at the same byte offsets from its quads (sync_frame_table)
	uint32_t tris_off = 0, xf_off = 0;
	uint32_t cyl_off = 0;             // byte offset of a frame's cylinder bounds from its quads (0 = none / not uniform)
	uint32_t* d_child_ranges = nullptr;  // [n_quads][4][2]: triangle range of every quad child (same for every frame of a topology); build_cyl_kernel
	uint32_t child_ranges_quads = 0;
	bool disable_cyl = false;         // INSTA_DEFORM_CYL=0: traversals without the cylinder bounds (A/B measurements)
	const unsigned char* arena = nullptr;  // lowest frame allocation; layout_uniform also requires every frame to end within 64 GB of it (32-bit offsets in 16-byte units)
	void* d_recorder = nullptr;       // recorder read-back staging: 8-bit RGBA + millimetre depth (recorder.cu)
	size_t recorder_cap = 0;          // pixels
	uint16_t* d_seed_grid = nullptr;  // 32^3 cells over render_aabb -> topology id of a nearby canonical triangle (deform.cu seed_bound)
	float seed_lo[3] = {0, 0, 0}, seed_scale[3] = {0, 0, 0};
	bool disable_seed_bound = false;  // INSTA_DEFORM_SEED=0: traversals start from FLT_MAX like the reference's (A/B measurements)
	std::vector<uint32_t> train_round_edges{0u, 8u, 32u, 0xffffffffu};  // per-ray sample ranges of the training step's rounds (at most 7 rounds)
	bool train_all_samples = false;   // INSTA_TRAIN_ROUNDS=0: insta_train_nerf deforms / evaluates every generated sample like the reference (A/B, parity tests)
	bool render_full_records = false; // INSTA_RENDER_RECORDS=full: insta_render_frame keeps the 27-float records between its kernels (A/B, parity tests)
	uint32_t coherent_max_n = 1u << 19; // tracer batches below this many records use the thread-per-query kernel (INSTA_DEFORM_COHERENT_MAX)
	uint32_t* d_counter = nullptr;    // 32 sample counters (128 B apart) of the persistent deform kernel
	std::atomic<uint32_t> counter_rr{0};
	// grid
	insta::GridLevels levels{};
	// host-path staging
	cudaStream_t streams[2] = {nullptr, nullptr};
	cudaEvent_t events[2] = {nullptr, nullptr};
	float* d_stage_rec[2] = {nullptr, nullptr};
	void* d_stage_enc[2] = {nullptr, nullptr};
	uint32_t stage_cap = 0, stage_stride = 0;
	std::atomic<uint64_t> launches{0};
	std::mutex mu;
	void* tracer = nullptr;           // insta::Tracer (nerf.cu)
	void* bwd_scratch = nullptr;      // insta::BwdScratch (mlp_bwd.cu)
	insta::Quad* d_canon_quads = nullptr;  // BVH of the canonical mesh, built at the first insta_raycast_mesh / insta_mesh_ray_intersect with frame_id < 0 (raycast.cu)
	insta::Tri* d_canon_tris_bvh = nullptr;
	float canon_centroid[3] = {0, 0, 0};   // TinyMesh::m_centorid (tiny_mesh.cu:64-86): mean of the canonical vertices
	float* d_adam_lr_table = nullptr; // bias-corrected learning rate per parameter step count (train.cu optimizer_step), kAdamLrTable floats
	void* d_train_images = nullptr;   // insta::ImageDev[n_train_images] (train.cu)
	uint32_t n_train_images = 0;
	void* trainer = nullptr;          // insta::Trainer (trainer.cu)
	void* comm = nullptr;             // insta::Comm (trainer.cu)
	void* occ_scratch = nullptr;      // insta::OccScratch (occupancy.cu)
	insta_render_opts render_opts{};  // last options given to insta_tracer_init_rays
};

namespace insta {
// deform.cu
int launch_build_xf(insta_ctx* ctx, const float* d_verts, TriXf* d_xf, const float* d_canon_area, float* d_area_out, cudaStream_t s);
int launch_build_seed_grid(insta_ctx* ctx, cudaStream_t s);
int launch_refit(insta_ctx* ctx, const FrameDev& fr, uint32_t n_quads, const float* d_verts, const uint32_t* d_order, cudaStream_t s);
int launch_build_cyl(insta_ctx* ctx, const FrameDev& fr, uint32_t n_quads, cudaStream_t s);
// recorder.cu
int recorder_readback(insta_ctx* ctx, const float* d_rgba, const float* d_depth, uint32_t w, uint32_t h, int srgb_to_linear, uint8_t* h_rgba8, uint32_t* h_depth_mm,
                      cudaStream_t s);
void recorder_destroy(insta_ctx* ctx);
int write_png_rgba8(const char* path, uint32_t w, uint32_t h, const uint8_t* rgba8);
int write_depth_bin(const char* path, const uint32_t* depth_mm, size_t n);
int write_trajectory_json(const char* path, uint32_t w, uint32_t h, float fl_x, float fl_y, float cx, float cy, const insta_trajectory_frame* frames, uint32_t n);
int launch_deform(insta_ctx* ctx, float* d_coords, uint32_t n, uint32_t stride, int32_t frame_id, const int32_t* d_frame_ids,
                  bool random_frames, int32_t* d_tri_ids, cudaStream_t s, const uint32_t* d_n = nullptr, bool coherent_hint = false,
                  bool compact_extra = false, const uint32_t* d_index = nullptr);
bool deform_supports_index(const insta_ctx* ctx);
// hashgrid.cu
int init_grid_levels(insta_ctx* ctx);
int launch_hash_fwd(insta_ctx* ctx, const float* d_pos, uint32_t stride, uint32_t n, const void* table, void* out, cudaStream_t s);
int launch_hash_bwd(insta_ctx* ctx, const float* d_pos, uint32_t stride, uint32_t n, const void* dLdy, void* grad, int zero_first, cudaStream_t s);
int launch_hash_bwd_pitched(insta_ctx* ctx, const float* d_pos, uint32_t stride, uint32_t n, const void* dLdy, uint32_t pitch, void* grad, int zero_first, cudaStream_t s);
int launch_hash_bwd_levels(insta_ctx* ctx, const float* d_pos, uint32_t stride, uint32_t n, const void* dLdy, uint32_t pitch, void* grad, int zero_first, uint32_t level_begin,
                           uint32_t level_end, cudaStream_t s);
int launch_hash_slots(insta_ctx* ctx, const float* d_pos, uint32_t stride, uint32_t n, uint32_t* slots, cudaStream_t s);
int sync_frame_table(insta_ctx* ctx, cudaStream_t s);
// mlp.cu
uint32_t network_n_params(const insta_ctx* ctx);
int launch_network_fwd(insta_ctx* ctx, const float* d_records, uint32_t stride, uint32_t n, const void* d_params, void* d_out, uint32_t out_stride, cudaStream_t s,
                       const uint32_t* d_n = nullptr,
                       const float* d_uniform_extra = nullptr, const uint32_t* d_index = nullptr);
// mlp_bwd.cu
void network_bwd_destroy(insta_ctx* ctx);
struct GradExchange {  // trainer.cu: side stream + events of the overlapped gradient sum (world > 1)
	cudaStream_t side;
	cudaEvent_t ev_grid_ready[4], ev_grid_summed;
};
int launch_network_fwd_bwd(insta_ctx* ctx, const float* d_records, uint32_t stride, uint32_t n, const void* d_params, const void* d_dloss, uint32_t dloss_width,
                           void* d_grads, void* d_out, cudaStream_t s, const GradExchange* ex = nullptr);
// train.cu
void pcg_seed(uint64_t seed, uint64_t advance_steps, uint64_t& state, uint64_t& inc);  // pcg32{seed} advanced `advance_steps` x advance()
int training_set_images(insta_ctx* ctx, const insta_train_image* h, uint32_t n);
int train_generate(insta_ctx* ctx, const insta_train_step* t, cudaStream_t s, bool emit = true);
int train_emit_round(insta_ctx* ctx, const insta_train_step* t, const uint32_t* d_index, const uint32_t* d_count, uint32_t bound, cudaStream_t s);
int train_round(insta_ctx* ctx, const insta_train_step* t, const void* d_net_out, uint32_t out_width, uint32_t prev_lo, uint32_t prev_hi, uint32_t lo, uint32_t hi,
                float* d_ray_T, uint32_t* d_index, uint32_t* d_count, cudaStream_t s);
int train_loss(insta_ctx* ctx, const insta_train_step* t, const void* d_net_out, uint32_t out_width, float* d_coords_compacted, void* d_dloss, float* d_loss,
               const float* d_mean_density, cudaStream_t s);
int optimizer_step(insta_ctx* ctx, const insta_optimizer* o, uint32_t n_params, uint32_t n_matrix, cudaStream_t s);
int launch_network_density(insta_ctx* ctx, const float* d_xyz, uint32_t stride, uint32_t n, const void* d_params, void* d_logits_f16, cudaStream_t s);
// snapshot.cpp
struct SnapshotFile;
int snapshot_save(const char* path, const insta_params* p, const insta_snapshot* s);
int snapshot_open(const char* path, SnapshotFile** out_file, insta_snapshot* s);
void snapshot_close(SnapshotFile* f);
// raycast.cu
int launch_mesh_ray_intersect(insta_ctx* ctx, const Quad* quads, const Tri* tris, const float* d_origins, const float* d_dirs, uint32_t n, int32_t* d_tri_ids, float* d_t,
                              cudaStream_t s);
int launch_mesh_overlay(insta_ctx* ctx, const insta_camera* cam, const float* aabb, const float* center, int color_normal, const Quad* quads, const Tri* tris, float* d_rgba,
                        float* d_depth, cudaStream_t s);
// microbench.cu
int microbench_gather(insta_ctx* ctx, uint64_t buffer_bytes, uint32_t chunk_bytes, uint32_t iters, float* h_gbs, cudaStream_t s);
int microbench_scatter(insta_ctx* ctx, uint64_t buffer_bytes, uint32_t chunk_bytes, uint32_t iters, float* h_gbs, cudaStream_t s);
// occupancy.cu
void occupancy_destroy(insta_ctx* ctx);
int density_grid_update(insta_ctx* ctx, insta_density_grid* g, const void* d_params_f16, uint32_t n_uniform, uint32_t n_nonuniform, cudaStream_t s);
int density_grid_mean_and_bitfield(insta_ctx* ctx, const insta_density_grid* g, cudaStream_t s);
// trainer.cu
void trainer_rng_advance(insta_ctx* ctx);
void trainer_destroy(insta_ctx* ctx);
void comm_destroy(insta_ctx* ctx);
int comm_unique_id(void* h_id128);
int comm_init(insta_ctx* ctx, int world, int rank, const void* h_id128);
int comm_world(const insta_ctx* ctx);
int comm_exchange_mode(const insta_ctx* ctx);
int grad_allreduce(insta_ctx* ctx, void* d_grads_f16, uint64_t n, cudaStream_t s);
int trainer_create(insta_ctx* ctx, const insta_trainer_config* cfg);
int trainer_state(const insta_ctx* ctx, insta_train_stats* out);
int train_nerf(insta_ctx* ctx, const insta_optimizer* opt, const void* d_bitfield, const float* d_extra_dims, const float* d_mean_density, int get_loss_scalar,
               insta_train_stats* h_stats, cudaStream_t s);
// nerf.cu
void tracer_destroy(insta_ctx* ctx);
int tracer_init_rays(insta_ctx* ctx, const insta_camera* cam, const insta_render_opts* opts, float* d_frame_rgba, float* d_frame_depth, float* d_dbg_origin_t,
                     float* d_dbg_dir_alive, uint32_t* h_n_alive, cudaStream_t s);
uint32_t tracer_n_steps(const insta_ctx* ctx);
int tracer_generate(insta_ctx* ctx, float* d_records, uint32_t stride, uint32_t n_steps, int32_t frame_id, const float* d_extra_dims, uint32_t* n_records,
                    cudaStream_t s);
int tracer_composite(insta_ctx* ctx, const float* d_records, uint32_t stride, const void* d_net_out, uint32_t out_stride, uint32_t n_steps, float* d_frame_rgba,
                     float* d_frame_depth, int linear_colors, uint32_t* h_n_alive, cudaStream_t s);
int render_frame(insta_ctx* ctx, const insta_camera* cam, const insta_render_opts* opts, int32_t frame_id, const void* d_params, const float* d_extra_dims,
                 float* d_frame_rgba, float* d_frame_depth, insta_render_stats* stats, cudaStream_t s);
}  // namespace insta
