// insta::Hooks -- header-only C++ adapter over the C ABI (insta_b200.h) with the reference's method names.
//
// This is the host-side mirror of the rta::Core hook surface for the hot path (include/rta/core.h:54-68 of the
// reference; hooks declared in ngp/include/neural-graphics-primitives/testbed.h:72-78): same names, same argument
// meaning, same error behaviour (std::runtime_error, as CUDA_CHECK_THROW raises).  It contains no computation: every
// method is one call into libinsta_b200.so.  INTEGRATION.md shows where the reference calls it.
#pragma once
#include <stdexcept>
#include <string>

#include "insta_b200.h"

namespace insta {

class Hooks {
public:
	Hooks() = default;
	Hooks(const Hooks&) = delete;
	Hooks& operator=(const Hooks&) = delete;
	~Hooks() { insta_ctx_destroy(m_ctx); }

	// rta::Core ctor + the INSTA keys read in src/core.cu:201-215
	void create(const insta_params& p, int device = -1) {
		insta_ctx_destroy(m_ctx);
		m_ctx = nullptr;
		check(insta_ctx_create(&p, device, &m_ctx));
	}
	insta_ctx* ctx() const { return m_ctx; }

	// rta::Core::load_meshes (src/core.cu:234-337)
	void load_canonical_obj(const std::string& canonical_obj, const std::string& mesh_masks_dir) {
		check(insta_load_canonical_obj(m_ctx, canonical_obj.c_str(), mesh_masks_dir.empty() ? nullptr : mesh_masks_dir.c_str()));
	}
	void set_canonical(const float* verts, uint32_t n_verts, const uint32_t* faces, uint32_t n_faces, const uint32_t* mask_ids) {
		check(insta_set_canonical(m_ctx, verts, n_verts, faces, n_faces, mask_ids));
	}
	int32_t add_frame(const float* verts_ngp_space) {
		int32_t idx = -1;
		check(insta_add_frame(m_ctx, verts_ngp_space, &idx));
		return idx;
	}
	int32_t add_frame_obj(const std::string& obj_path) {
		int32_t idx = -1;
		check(insta_add_frame_obj(m_ctx, obj_path.c_str(), &idx));
		return idx;
	}
	void add_frames(const float* verts, uint32_t n, uint32_t n_threads = 0) { check(insta_add_frames(m_ctx, verts, n, n_threads)); }
	void clear_frames() { check(insta_clear_frames(m_ctx)); }
	uint32_t n_frames() const { return insta_n_frames(m_ctx); }
	// Testbed::m_aabb (tested inclusively by `warp`, core.cu:477)
	void set_aabb(const float aabb6[6]) { check(insta_set_aabb(m_ctx, aabb6)); }
	void render_aabb(float aabb6[6]) const { check(insta_get_render_aabb(m_ctx, aabb6)); }

	// rta::Core::surface_closest_point, 27-row path (src/core.cu:339-344, 581-608)
	void surface_closest_point(float* d_coords, uint32_t n, uint32_t floats_per_coord, void* stream, int32_t* d_tri_ids = nullptr) {
		check(insta_deform_samples(m_ctx, d_coords, n, floats_per_coord, -1, d_tri_ids, stream));
	}
	// rta::Core::render_deformed (src/core.cu:346-348, 610-637)
	void render_deformed(float* d_coords, uint32_t n, uint32_t floats_per_coord, uint32_t target_deform_frame, void* stream, int32_t* d_tri_ids = nullptr) {
		check(insta_deform_samples(m_ctx, d_coords, n, floats_per_coord, (int32_t)target_deform_frame, d_tri_ids, stream));
	}
	// rta::Core::surface_closest_point, 3-row path (src/core.cu:639-655)
	void surface_closest_point_3d(float* d_xyz, uint32_t n, void* stream, const int32_t* d_frame_ids = nullptr, int32_t* d_tri_ids = nullptr) {
		check(insta_deform_points(m_ctx, d_xyz, n, d_frame_ids, d_tri_ids, stream));
	}

	// tcnn::GridEncodingTemplated<__half,3,8>::forward_impl / backward_impl (tcnn encodings/grid.h:941-1110)
	void grid_forward(const float* d_pos, uint32_t stride_floats, uint32_t n, const void* d_params_f16, void* d_out_soa_f16, void* stream) {
		check(insta_hash_encode_fwd(m_ctx, d_pos, stride_floats, n, d_params_f16, d_out_soa_f16, stream));
	}
	void grid_backward(const float* d_pos, uint32_t stride_floats, uint32_t n, const void* d_dLdy_soa_f16, void* d_grad_f16, bool overwrite, void* stream) {
		check(insta_hash_encode_bwd(m_ctx, d_pos, stride_floats, n, d_dLdy_soa_f16, d_grad_f16, overwrite ? 1 : 0, stream));
	}
	uint32_t grid_n_params() const { return insta_grid_n_params(m_ctx); }

	// ngp::NerfNetwork<__half>: inference_mixed_precision_impl / density / forward + backward (ngp nerf_network.h:159-232, 385-414, 238-383)
	uint32_t network_n_params() const { return insta_network_n_params(m_ctx); }
	void network_inference(const float* d_records, uint32_t floats_per_coord, uint32_t n, const void* d_params_f16, void* d_rgbsigma_f16, uint32_t out_stride_elems,
	                       void* stream) {
		check(insta_network_inference(m_ctx, d_records, floats_per_coord, n, d_params_f16, d_rgbsigma_f16, out_stride_elems, stream));
	}
	void network_density(const float* d_xyz, uint32_t stride_floats, uint32_t n, const void* d_params_f16, void* d_logits_f16, void* stream) {
		check(insta_network_density(m_ctx, d_xyz, stride_floats, n, d_params_f16, d_logits_f16, stream));
	}
	void network_forward_backward(const float* d_records, uint32_t floats_per_coord, uint32_t n, const void* d_params_f16, const void* d_dloss_f16, uint32_t dloss_width,
	                              void* d_gradients_f16, void* d_out_f16, void* stream) {
		check(insta_network_forward_backward(m_ctx, d_records, floats_per_coord, n, d_params_f16, d_dloss_f16, dloss_width, d_gradients_f16, d_out_f16, stream));
	}

	// Testbed::render_nerf -> NerfTracer (ngp/src/testbed_nerf.cu:2320-2440, 2030-2244) for frame `target_deform_frame`
	insta_render_stats render_nerf(const insta_camera& cam, const insta_render_opts& opts, int32_t target_deform_frame, const void* d_params_f16,
	                               const float* d_extra_dims, float* d_frame_rgba, float* d_frame_depth, void* stream) {
		insta_render_stats st{};
		check(insta_render_frame(m_ctx, &cam, &opts, target_deform_frame, d_params_f16, d_extra_dims, d_frame_rgba, d_frame_depth, &st, stream));
		return st;
	}

	// rta::Core::raycast_flame (src/core.cu:350-402; called from Testbed::render_frame, ngp/src/testbed.cu:2626-2628): the shaded mesh overlay.
	// render_deformed / target_deform_frame / raycast_normal = Core::m_render_deformed / m_target_deform_frame / m_raycast_normal
	void raycast_flame(const insta_camera& cam, const insta_render_opts& opts, bool render_deformed, uint32_t target_deform_frame, bool raycast_normal, float* d_frame_rgba,
	                   float* d_frame_depth, void* stream) {
		check(insta_raycast_mesh(m_ctx, &cam, &opts, render_deformed ? (int32_t)target_deform_frame : -1, raycast_normal ? 1 : 0, d_frame_rgba, d_frame_depth, stream));
	}

	// rta::Recorder::dump_frame_buffer (src/recorder.cu:299-359): the frame as the recorder stores it, converted on the GPU; host buffers valid after the
	// caller's next stream synchronisation
	void read_back_frame(const float* d_frame_rgba, const float* d_frame_depth, uint32_t width, uint32_t height, bool normals_mode, uint8_t* h_rgba8,
	                     uint32_t* h_depth_mm, void* stream) {
		check(insta_recorder_readback(m_ctx, d_frame_rgba, d_frame_depth, width, height, normals_mode ? 1u : 0u, h_rgba8, h_depth_mm, stream));
	}

	// Testbed::load_nerf -> metadata_gpu / transforms_gpu (testbed_nerf.cu:2698-2760)
	void set_training_images(const insta_train_image* images, uint32_t n) { check(insta_training_set_images(m_ctx, images, n)); }
	// Testbed::train_nerf (testbed_nerf.cu:2916-2992): one step; returns the counters update_after_training would have set
	void create_trainer(const insta_trainer_config& cfg) { check(insta_trainer_create(m_ctx, &cfg)); }
	insta_train_stats train_nerf(const insta_optimizer& opt, const void* d_density_grid_bitfield, const float* d_extra_dims, const float* d_density_grid_mean,
	                             bool get_loss_scalar, void* stream) {
		insta_train_stats st{};
		check(insta_train_nerf(m_ctx, &opt, d_density_grid_bitfield, d_extra_dims, d_density_grid_mean, get_loss_scalar ? 1 : 0, &st, stream));
		return st;
	}
	// Testbed::training_prep_nerf / update_density_grid_nerf / update_density_grid_mean_and_bitfield (testbed_nerf.cu:3428-3444, 2774-2878)
	bool training_prep_nerf(insta_density_grid& grid, const void* d_params_f16, void* stream) {
		int did = 0;
		check(insta_training_prep_nerf(m_ctx, &grid, d_params_f16, &did, stream));
		return did != 0;
	}
	void update_density_grid_nerf(insta_density_grid& grid, const void* d_params_f16, uint32_t n_uniform, uint32_t n_nonuniform, void* stream) {
		check(insta_update_density_grid(m_ctx, &grid, d_params_f16, n_uniform, n_nonuniform, stream));
	}
	void update_density_grid_mean_and_bitfield(const insta_density_grid& grid, void* stream) {
		check(insta_update_density_grid_mean_and_bitfield(m_ctx, &grid, stream));
	}
	// data-parallel training (no counterpart in the single-GPU reference)
	void init_communicator(int world_size, int rank, const void* nccl_unique_id_128) { check(insta_comm_init_rank(m_ctx, world_size, rank, nccl_unique_id_128)); }

	// Testbed::save_snapshot / load_snapshot (ngp/src/testbed.cu:2916-3016); host buffers
	static void save_snapshot(const std::string& path, const insta_params* p, const insta_snapshot& s) { check(insta_snapshot_save(path.c_str(), p, &s)); }
	struct LoadedSnapshot {
		insta_snapshot_file* file = nullptr;
		insta_snapshot data{};
		LoadedSnapshot() = default;
		LoadedSnapshot(const LoadedSnapshot&) = delete;
		LoadedSnapshot& operator=(const LoadedSnapshot&) = delete;
		LoadedSnapshot(LoadedSnapshot&& o) noexcept : file(o.file), data(o.data) { o.file = nullptr; }
		~LoadedSnapshot() { insta_snapshot_close(file); }
	};
	static LoadedSnapshot load_snapshot(const std::string& path) {
		LoadedSnapshot l;
		check(insta_snapshot_open(path.c_str(), &l.file, &l.data));
		return l;
	}

	static void check(int status) {
		if (status != INSTA_OK) throw std::runtime_error(std::string("insta_b200: ") + insta_last_error());
	}

private:
	insta_ctx* m_ctx = nullptr;
};

}  // namespace insta
