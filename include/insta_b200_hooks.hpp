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

	static void check(int status) {
		if (status != INSTA_OK) throw std::runtime_error(std::string("insta_b200: ") + insta_last_error());
	}

private:
	insta_ctx* m_ctx = nullptr;
};

}  // namespace insta
