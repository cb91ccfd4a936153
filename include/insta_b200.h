/* insta_b200 -- C ABI of the B200-native (sm_100a) implementation of INSTA's per-sample deformation-and-encoding
 * hot path.  This is the drop-in boundary: plain pointers and sizes, no C++/torch types, no exceptions.
 *
 * Conventions
 *   - every function returns INSTA_OK (0) or a negative insta_status; insta_last_error() gives the message of the
 *     last failure on the calling thread (the reference throws std::runtime_error via CUDA_CHECK_THROW instead;
 *     src/main.cu:261 catches it).
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers.  Nothing is allocated inside a hot-path call.
 *   - every hot-path entry takes the caller's cudaStream_t (as void*), is stream-ordered and never synchronises --
 *     same contract as the reference hooks, which all run on Testbed::m_stream (testbed.cu:2488).
 *   - one context per GPU; a context is bound to the device it was created on.
 *
 * Reference interface replaced by each entry point is cited beside it (paths relative to the INSTA repository;
 * ngp/ = instant-ngp/, tcnn/ = instant-ngp/dependencies/tiny-cuda-nn/).
 */
#ifndef INSTA_B200_H
#define INSTA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define INSTA_API __declspec(dllexport)
#else
#define INSTA_API __attribute__((visibility("default")))
#endif

typedef struct insta_ctx insta_ctx;
typedef void* insta_stream; /* cudaStream_t */

typedef enum insta_status {
	INSTA_OK = 0,
	INSTA_ERR_INVALID = -1,   /* bad argument / state */
	INSTA_ERR_CUDA = -2,      /* a CUDA runtime call failed */
	INSTA_ERR_IO = -3,        /* file could not be read / parsed */
	INSTA_ERR_NO_DEVICE = -4, /* no sm_100 device: there is NO CPU fallback */
	INSTA_ERR_CAPACITY = -5   /* frame cache (max_cached_bvh) exhausted */
} insta_status;

/* Plain-old-data parameters, filled from the reference's JSON keys.
 *   INSTA keys read in src/core.cu:201-215 (configs/nerf/insta.json, main.json)
 *   "encoding" block of configs/nerf/main.json:45-53 and ngp/src/testbed.cu:2246-2258 (per_level_scale). */
typedef struct insta_params {
	uint32_t struct_size;            /* = sizeof(insta_params); ABI versioning */
	/* ---- INSTA keys ---- */
	int32_t  disable_interp;         /* "disable_interp"  (core.cu:201) : 1 = never blend neighbour TBNs */
	uint32_t max_cached_bvh;         /* "max_cached_bvh"  (core.cu:204) : capacity of the per-frame BVH cache, default 4000 */
	uint32_t exp_params;             /* "exp_params"      (core.cu:210) : expression floats appended to each sample, 16 */
	uint32_t geo_params;             /* "geo_params"      (core.cu:209) : 0 = off; >0 writes the cosine feature */
	uint32_t n_primitives_per_leaf;  /* N_PRIMITIVES_LEAF (src/tiny_mesh.cu:28-32) : 2 in Release */
	float    aabb_min[3], aabb_max[3]; /* Testbed::m_aabb used by `warp` to skip samples (core.cu:477); unit cube */
	/* ---- "encoding" (HashGrid) ---- */
	uint32_t n_levels;               /* 16 */
	uint32_t n_features_per_level;   /* 8 (only 8 is implemented: one 128-bit load per corner) */
	uint32_t log2_hashmap_size;      /* 18 */
	uint32_t base_resolution;        /* 16 */
	float    per_level_scale;        /* 0 = derive as testbed.cu:2246-2258 does: exp(log(2048*aabb_scale/base)/(n_levels-1)) */
	uint32_t aabb_scale;             /* 1 */
	/* ---- warp_random_samples RNG (ngp/src/testbed.cu:2185: default_rng_t{42}) ---- */
	uint64_t rng_seed;
	/* ---- rounding flavour of the closest-triangle distance arithmetic (it breaks ties) ----
	 * 0 = bit-for-bit what the reference's CUDA build computes (nvcc's FMA contraction of bvh.cuh / triangle.cuh), default;
	 * 1 = bit-for-bit what the reference's HOST build of the same functions computes (nothing contracted). */
	uint32_t arith;
	uint32_t reserved_;
} insta_params;

INSTA_API void        insta_params_default(insta_params* p);
INSTA_API const char* insta_last_error(void);
INSTA_API const char* insta_version(void);

/* rta::Core ctor + Testbed device setup.  device < 0 = current device.  Fails with INSTA_ERR_NO_DEVICE when no
 * CUDA device of compute capability 10.x is present. */
INSTA_API int  insta_ctx_create(const insta_params* p, int device, insta_ctx** out);
INSTA_API void insta_ctx_destroy(insta_ctx* ctx);

/* ------------------------------------------------------------------------------------------------------------
 * Mesh ingest -- replaces rta::Core::load_meshes (src/core.cu:234-337), rta::TinyMesh (src/tiny_mesh.cu:39-178,
 * 316-371), rta::Masking (src/masking.cpp:118-180) and ngp::TriangleBvh4::build (ngp/src/triangle_bvh.cu:548-631).
 * Vertices are in ngp space, i.e. ALREADY offset by +0.5 (tiny_mesh.cu:63-74); the *_obj variants add it themselves.
 * ------------------------------------------------------------------------------------------------------------ */
/* canonical template: topology, per-triangle FlameMasks ids (NULL = all Torso/15), adjacency, render AABB */
INSTA_API int insta_set_canonical(insta_ctx* ctx, const float* h_verts, uint32_t n_verts, const uint32_t* h_faces,
                                  uint32_t n_faces, const uint32_t* h_mask_ids);
/* Masking ctor + TinyMesh ctor on "<dir>/canonical.obj" with region masks in "<mesh_masks_dir>/{lips,face,...}.obj" */
INSTA_API int insta_load_canonical_obj(insta_ctx* ctx, const char* canonical_obj, const char* mesh_masks_dir);
/* one tracked frame: triangles (+scaling vs canonical), BVH (reference-identical topology and leaf order),
 * per-triangle TBN table; returns its frame index in *frame_index (== image index in the reference) */
INSTA_API int insta_add_frame(insta_ctx* ctx, const float* h_verts, int32_t* frame_index);
/* n frames at once, h_verts = [n][n_verts][3]; host BVH builds run on n_threads threads (0 = all cores), like the
 * thread pool in core.cu:309 */
INSTA_API int insta_add_frames(insta_ctx* ctx, const float* h_verts, uint32_t n, uint32_t n_threads);
INSTA_API int insta_add_frame_obj(insta_ctx* ctx, const char* obj_path, int32_t* frame_index);
INSTA_API int insta_clear_frames(insta_ctx* ctx);
/* Frame-cache turnover = rta::Core::reload_training_data (src/core.cu:663-706: every 1,500 training steps the reference synchronises the device, drops the
 * dataset and loads another subset of images + meshes, load_meshes core.cu:294-334).  Replaces the meshes of n resident frames in place: host BVH builds on
 * n_threads threads (the reference's build order, so the winners among tied triangles stay the reference's), uploads into the frames' existing device
 * storage (no allocation, pointers in the frame table unchanged).  Synchronises the device at entry like the reference.  h_verts: n vertex sets in the
 * canonical topology; h_frame_indices: which resident frame each one replaces (distinct). */
INSTA_API int insta_replace_frames(insta_ctx* ctx, const int32_t* h_frame_indices, const float* h_verts, uint32_t n, uint32_t n_threads);
/* The same turnover WITHOUT the host build (not in the reference, non-parity): the frames keep the tree they have and its boxes are refitted
 * on the GPU to the new vertices.  The nearest triangle of a query is still found (same distance), but where several triangles are exactly equidistant the
 * winner may differ from the reference's, because the visit order follows the old frame's median splits (tests/test_gpu_frames.py reports the rate).
 * Does not synchronise: copies and kernels are ordered on `stream`. */
INSTA_API int insta_refit_frames(insta_ctx* ctx, const int32_t* h_frame_indices, const float* h_verts, uint32_t n, insta_stream stream);
INSTA_API uint32_t insta_n_frames(const insta_ctx* ctx);
/* TinyMesh::m_aabb of the canonical mesh (tiny_mesh.cu:157-162) -> Testbed::m_render_aabb (core.cu:289) */
INSTA_API int insta_get_render_aabb(const insta_ctx* ctx, float* h_aabb6);
/* Testbed::m_aabb, tested inclusively by `warp` */
INSTA_API int insta_set_aabb(insta_ctx* ctx, const float* h_aabb6);
/* host copies for inspection / parity tests (any pointer may be NULL):
 *   mask ids [n_faces], adjacency [n_faces*3] (-1 padded), BVH-order topology ids of a frame [n_faces] */
INSTA_API int insta_get_topology(const insta_ctx* ctx, uint32_t* h_mask_ids, int32_t* h_adjacency);
INSTA_API int insta_get_frame_bvh_order(const insta_ctx* ctx, int32_t frame, uint32_t* h_tri_ids, uint32_t* n_nodes);
/* The traversal's cylinder bounds of a frame (this library's own pruning structure, not in the reference): per quad (= inner node of
 * the reference tree, ngp triangle_bvh.cu:548-631, in node order) and child 8 fp16 values (axis x y z, slab centre, slab half width, radius, 0, 0; the
 * last three in doubled units).  h_fp16_bits: n_quads * 4 * 8 uint16 (NULL: only the count).  Synchronises.  *n_quads = 0 when the topology has none. */
INSTA_API int insta_get_frame_cylinders(insta_ctx* ctx, int32_t frame, uint16_t* h_fp16_bits, uint32_t* n_quads);

/* ------------------------------------------------------------------------------------------------------------
 * Stage 1 -- deformation.
 * ------------------------------------------------------------------------------------------------------------ */
/* Replaces rta::Core::surface_closest_point (27-row path -> kernel `warp`, src/core.cu:339-344, 442-526, 581-608;
 * called at ngp/src/testbed_nerf.cu:3290) when frame_id == -1 (frame = per-sample `holder`, float #4), and
 * rta::Core::render_deformed (core.cu:346-348, 610-637; called at testbed_nerf.cu:2199-2200) when frame_id >= 0.
 * d_coords: n records of stride_floats floats (NerfCoordinate + extra dims, ngp nerf.h:86-115), mutated IN PLACE:
 * floats 0-2 <- canonical position, 8-10 <- canonical direction (stored (d+1)/2), the last exp_params floats <- 1.0
 * unless the closest triangle is in the Lips region, last float <- cosine feature if geo_params > 0.
 * Records with holder < 0, frame >= n_frames or position outside m_aabb are left untouched.
 * d_tri_ids (nullable): topology id of the closest triangle per sample, -1 for skipped samples.
 * d_coords must be 16-byte aligned. */
INSTA_API int insta_deform_samples(insta_ctx* ctx, float* d_coords, uint32_t n, uint32_t stride_floats, int32_t frame_id,
                                   int32_t* d_tri_ids, insta_stream stream);
/* Same with the record count on the DEVICE: only min(n_max, *d_n_valid) records are processed (d_n_valid NULL = n_max).  For callers whose
 * sample count is produced by a kernel (generate_training_samples_nerf's counter): no host round trip, where the reference deforms
 * `max_inference` records -- last step's count rounded up (ngp/src/testbed_nerf.cu:3217-3222, 3290). */
INSTA_API int insta_deform_samples_n(insta_ctx* ctx, float* d_coords, uint32_t n_max, const uint32_t* d_n_valid, uint32_t stride_floats, int32_t frame_id,
                                     int32_t* d_tri_ids, insta_stream stream);
/* Replaces the 3-row path of surface_closest_point (kernel `warp_random_samples`, core.cu:528-579, 639-655; called
 * at testbed_nerf.cu:2850 for the occupancy-grid refresh).  d_xyz: n x 3 floats in place.  d_frame_ids NULL = draw
 * the frame per point exactly as the reference: pcg32(rng_seed).advance(i*(n_frames-1)).next_float()*(n_frames-1). */
INSTA_API int insta_deform_points(insta_ctx* ctx, float* d_xyz, uint32_t n, const int32_t* d_frame_ids, int32_t* d_tri_ids,
                                  insta_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * Stage 2 -- multiresolution hash grid (tcnn::GridEncodingTemplated<__half,3,8>, tcnn encodings/grid.h).
 * The parameter table is BORROWED from the caller (trainer buffer order is the snapshot ABI, tcnn trainer.h:282-297):
 * fp16, n_entries x 8, entry offsets per level as insta_grid_layout reports.
 * ------------------------------------------------------------------------------------------------------------ */
/* ctor offset table (grid.h:874-939).  h_offsets: n_levels+1 (entries), h_scales / h_resolutions: n_levels; the scale
 * is the value the device kernels use (evaluated once ON DEVICE with the reference kernel's expression, grid.h:188). */
INSTA_API int      insta_grid_layout(const insta_ctx* ctx, uint32_t* h_offsets, float* h_scales, uint32_t* h_resolutions);
INSTA_API uint32_t insta_grid_n_params(const insta_ctx* ctx); /* = n_entries * 8  (25,716,288 for the shipped config) */
/* forward_impl / kernel_grid (grid.h:138-310, 941-1017).  d_pos: xyz at d_pos[i*pos_stride_floats + {0,1,2}];
 * d_out_soa: fp16 [n_levels*8][n]  (out[i + (l*8+f)*n], grid.h:265 -- tcnn row-major == SoA) */
INSTA_API int insta_hash_encode_fwd(insta_ctx* ctx, const float* d_pos, uint32_t pos_stride_floats, uint32_t n,
                                    const void* d_table_f16, void* d_out_soa_f16, insta_stream stream);
/* backward_impl / kernel_grid_backward (grid.h:312-435, 1019-1110).  d_dLdy_soa: fp16 [n_levels*8][n];
 * d_grad_table_f16: n_entries x 8 fp16, zeroed first when zero_first != 0 (EGradientMode::Overwrite, grid.h:1067) */
INSTA_API int insta_hash_encode_bwd(insta_ctx* ctx, const float* d_pos, uint32_t pos_stride_floats, uint32_t n,
                                    const void* d_dLdy_soa_f16, void* d_grad_table_f16, int zero_first, insta_stream stream);
/* parity diagnostic: global entry index of the 8 corners, d_slots[(i*n_levels + l)*8 + corner] (grid_index, grid.h:100) */
INSTA_API int insta_hash_slots(insta_ctx* ctx, const float* d_pos, uint32_t pos_stride_floats, uint32_t n, uint32_t* d_slots,
                               insta_stream stream);

/* NerfNetwork::forward + backward on the compacted training batch (nerf_network.h:238-383; testbed_nerf.cu:3369-3370):
 * d_dloss_f16 = dL/d(r, g, b, sigma logits), dloss_width halves per sample (already multiplied by the loss scale);
 * d_gradients_f16 receives dL/dparams in parameter order (n_params halves, EGradientMode::Overwrite); d_out_f16 (nullable):
 * the forward output, 4 halves per sample.  Scratch for the layer activations is owned by the context. */
INSTA_API int insta_network_forward_backward(insta_ctx* ctx, const float* d_records, uint32_t stride_floats, uint32_t n, const void* d_params_f16,
                                             const void* d_dloss_f16, uint32_t dloss_width, void* d_gradients_f16, void* d_out_f16, insta_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * Stage 3 -- render tracer (ngp::Testbed::NerfTracer, ngp/src/testbed_nerf.cu:2030-2244 and the kernels it launches).
 * The ray state lives in the context (NerfTracer::m_rays, testbed.h:146-230); the sample records and the network
 * output are the caller's buffers, exactly as in the reference, so that any network can sit between generate and
 * composite.  insta_render_frame runs the whole loop with this library's own network.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct insta_camera {
	float    camera_matrix[12];      /* Eigen::Matrix<float,3,4>, column-major: columns right, down, forward, origin (ngp space) */
	float    focal_length[2];        /* pixels; Testbed::calc_focal_length (ngp/src/testbed.cu:2531) */
	float    screen_center[2];       /* Testbed::render_screen_center, 0.5 / 0.5 by default */
	uint32_t resolution[2];
	uint32_t sample_index;           /* render_buffer.spp() */
	uint32_t snap_to_pixel_centers;  /* Testbed::m_snap_to_pixel_centers (default true, testbed.h:829) */
} insta_camera;

typedef struct insta_render_opts {
	const void* d_density_grid_bitfield; /* Testbed::Nerf::density_grid_bitfield: 128^3 bits per cascade, Morton order; NULL = no skipping */
	float    cone_angle_constant;        /* 0 for aabb_scale 1 (testbed_nerf.cu:2691) */
	float    min_transmittance;          /* rendering_min_transmittance, 0.01 (testbed.h:669) */
	uint32_t use_render_aabb;            /* 0 = the canonical mesh's render AABB (core.cu:289), 1 = render_aabb below */
	float    render_aabb[6];
	uint32_t linear_colors;              /* Testbed::Nerf::Training::linear_colors (false: sRGB -> linear when accumulating) */
	uint32_t max_steps_per_iteration;    /* insta_render_frame only.  March steps per ray between two compactions: clamp(budget / n_alive, 1, this).
	                                      * 8 = the reference's schedule (MAX_STEPS_INBETWEEN_COMPACTION, budget = rays initialised,
	                                      * testbed_nerf.cu:68, 2172); 0 = library default (16, budget = max(pixels, 2^21) records): fewer,
	                                      * larger deformation / network batches.  The frame does not depend on the schedule (a ray's samples
	                                      * and their compositing order are the same), only the number of samples evaluated past a ray's
	                                      * termination does. */
} insta_render_opts;

typedef struct insta_render_stats {
	uint32_t iterations; /* march iterations (= stream syncs) */
	uint32_t reserved_;
	uint64_t samples;    /* sample records generated (incl. padding of rays that left the box) */
} insta_render_stats;

/* Testbed::render_nerf (ngp/src/testbed_nerf.cu:2320-2440): rays from the camera, march / deform with frame frame_id
 * (rta::Core::render_deformed; frame_id < 0 = no deformation) / network / composite until no ray is alive, result
 * accumulated into d_frame_rgba (res.x*res.y float4, zeroed first) and d_frame_depth.  d_extra_dims: the frame's exp_params
 * expression coefficients (Testbed::get_inference_extra_dims), NULL = zeros.  Synchronises the stream once per march
 * iteration like NerfTracer::trace (testbed_nerf.cu:2164-2165); opts->max_steps_per_iteration sets how many steps an
 * iteration marches. */
INSTA_API int insta_render_frame(insta_ctx* ctx, const insta_camera* cam, const insta_render_opts* opts, int32_t frame_id, const void* d_params_f16,
                                 const float* d_extra_dims, float* d_frame_rgba, float* d_frame_depth, insta_render_stats* h_stats, insta_stream stream);

/* NerfTracer::init_rays_from_camera: init_rays_with_payload_kernel_nerf + advance_pos_nerf (testbed_nerf.cu:1854-1960,
 * 634-690, 2030-2110).  d_frame_rgba (res.x*res.y float4) is left untouched, d_frame_depth is set to 1e10 like the
 * reference.  d_dbg_* (nullable, res.x*res.y float4 each): per-pixel (origin, t) and (dir, alive) after the skip to the first
 * occupied cell, for parity tests.  *h_n_alive receives the number of live rays (one stream sync, as in the reference). */
INSTA_API int insta_tracer_init_rays(insta_ctx* ctx, const insta_camera* cam, const insta_render_opts* opts, float* d_frame_rgba, float* d_frame_depth,
                                     float* d_dbg_origin_t, float* d_dbg_dir_alive, uint32_t* h_n_alive, insta_stream stream);
/* clamp(n_rays_initialized / n_alive, 1, 8) -- the reference's steps between compactions (testbed_nerf.cu:2172) */
INSTA_API uint32_t insta_tracer_n_steps(const insta_ctx* ctx);
/* generate_next_nerf_network_inputs (testbed_nerf.cu:722-795): n_alive * n_steps records, record of (ray i, step j) at
 * index i + j * n_alive, holder = frame_id (-1 for steps a ray does not reach), d_extra_dims = the frame's expression
 * coefficients copied into floats 11.. of every record.  *n_records = n_alive * n_steps. */
INSTA_API int insta_tracer_generate_samples(insta_ctx* ctx, float* d_records, uint32_t stride_floats, uint32_t n_steps, int32_t frame_id,
                                            const float* d_extra_dims, uint32_t* n_records, insta_stream stream);
/* composite_kernel_nerf + compact_kernel_nerf + shade_kernel_nerf (testbed_nerf.cu:797-1009, 1829-1852, 1794-1827).
 * d_net_out fp16: out_stride_elems == 0 -> 4 halves per sample (r, g, b, sigma logits); else out[i + c * out_stride_elems]
 * (the reference's row-major network output, testbed_nerf.cu:842-845). */
INSTA_API int insta_tracer_composite(insta_ctx* ctx, const float* d_records, uint32_t stride_floats, const void* d_net_out_f16, uint32_t out_stride_elems,
                                     uint32_t n_steps, float* d_frame_rgba, float* d_frame_depth, uint32_t* h_n_alive, insta_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * Recorder read-back -- rta::Recorder::dump_frame_buffer (src/recorder.cu:299-359).  The reference copies the float4 frame
 * buffer and the float depth buffer to the host with two blocking cudaMemcpy (20 bytes per pixel) and converts them on one
 * host thread; here the conversion runs on the GPU and 8 bytes per pixel leave the device with asynchronous copies on
 * `stream` (no synchronisation: the host buffers are valid after the caller's next sync; pinned buffers keep it asynchronous).
 *   h_rgba8    : width*height*4 bytes, (uint8_t) clamp(transform(v) * 255, 0, 255) per channel incl. alpha (save_rgba,
 *                recorder.cu:36-60); transform = identity, or srgb_to_linear when srgb_to_linear != 0 (Normals mode, 341-343)
 *   h_depth_mm : width*(height-1) values, uint32(d * 1000) with d = 0 outside [0, 5] (save_depth, recorder.cu:75-105; the
 *                reference reads depth[x + width + y*width] for output pixel (x, y) and skips indices past the buffer, so
 *                the file holds rows 1 .. height-1 -- kept, the .bin files are what downstream scripts read)
 * Either host pointer may be NULL. */
INSTA_API int insta_recorder_readback(insta_ctx* ctx, const float* d_frame_rgba, const float* d_frame_depth, uint32_t width, uint32_t height,
                                      uint32_t srgb_to_linear, uint8_t* h_rgba8, uint32_t* h_depth_mm, insta_stream stream);
/* Host only.  8-bit RGBA PNG (stb's stbi_write_png in the reference, recorder.cu:59; stored deflate blocks here -- the decoded
 * pixels are what agrees) and the depth .bin (raw little-endian uint32 millimetres, recorder.cu:88-103). */
INSTA_API int insta_host_write_png(const char* path, uint32_t width, uint32_t height, const uint8_t* h_rgba8);
INSTA_API int insta_host_write_depth_bin(const char* path, const uint32_t* h_depth_mm, uint64_t n_values);
/* Host only.  The camera trajectory of a recording -- rta::Recorder::dump_cameras_json (src/recorder.cu:648-695) over the JsonFrame entries that
 * dump_frame_buffer appends per recorded frame (recorder.cu:318-343): image size, focal lengths in pixels, principal point, and per frame the 4x4
 * camera-to-world matrix (the 3x4 camera with 0.5 subtracted from its translation: INSTA's scene offset) plus the paths "flame/exp/<id>.txt",
 * "flame/eyes/<id>.txt", "meshes/<id>.obj" with <id> = the 5-digit frame index + suffix.  Frames are sorted by mesh path, keys and the 4-space
 * indentation are nlohmann::json's.  (Recorder::raycast_flame -- the shaded mesh overlay of the GUI, src/core.cu:350-402 -- is not part of this library.) */
typedef struct insta_trajectory_frame {
	float       camera_matrix[12]; /* 3x4, row-major: rta::Core::m_camera at the time the frame was recorded */
	uint32_t    frame_index;       /* Recorder::m_index_frame */
	const char* suffix;            /* dump_frame_buffer's suffix argument, NULL = "" */
} insta_trajectory_frame;
INSTA_API int insta_host_write_trajectory_json(const char* path, uint32_t width, uint32_t height, float fl_x, float fl_y, float cx, float cy,
                                               const insta_trajectory_frame* h_frames, uint32_t n_frames);

/* ------------------------------------------------------------------------------------------------------------
 * Network -- ngp::NerfNetwork<__half> (ngp/include/neural-graphics-primitives/nerf_network.h) with the shipped
 * configuration (configs/nerf/main.json:33-69): HashGrid 16x8 -> (+16 expression rows) -> density MLP 144-64-64-16 ->
 * (16 density outputs | SH degree 4 of the direction) -> colour MLP 32-64-64-64-16.  d_params_f16 is the trainer's fp16
 * parameter buffer in its own order (density MLP, colour MLP, grid; nerf_network.h:492-525; weights row-major
 * [out][in], tcnn src/fully_fused_mlp.cu:662-681), i.e. exactly `params_binary` of a snapshot.
 * ------------------------------------------------------------------------------------------------------------ */
INSTA_API uint32_t insta_network_n_params(const insta_ctx* ctx); /* 25,600 + grid = 25,741,888 for the shipped config */
/* NerfNetwork::inference_mixed_precision_impl (nerf_network.h:159-232): one fused kernel (encode + both MLPs + SH).
 * d_records: n sample records (position floats 0-2, direction floats 8-10 stored as (d+1)/2, the last exp_params floats =
 * expression coefficients).  Output fp16: out_stride_elems == 0 -> 4 halves per sample (r, g, b, sigma logits);
 * else out[i + c * out_stride_elems], c = 0..3 -- the reference's row-major rgbsigma matrix (testbed_nerf.cu:2197). */
INSTA_API int insta_network_inference(insta_ctx* ctx, const float* d_records, uint32_t stride_floats, uint32_t n, const void* d_params_f16, void* d_out_f16,
                                      uint32_t out_stride_elems, insta_stream stream);

/* insta_network_inference over min(n_max, *d_n_valid) records (d_n_valid: device pointer, NULL = n_max) */
INSTA_API int insta_network_inference_n(insta_ctx* ctx, const float* d_records, uint32_t stride_floats, uint32_t n_max, const uint32_t* d_n_valid,
                                        const void* d_params_f16, void* d_out_f16, uint32_t out_stride_elems, insta_stream stream);
/* NerfNetwork::density (nerf_network.h:385-414): hash grid + density MLP on bare positions (d_xyz[i * stride_floats + {0,1,2}]), expression
 * inputs all 1 (extract_expressions_random, nerf_network.h:108-128); one fp16 density logit per sample. */
INSTA_API int insta_network_density(insta_ctx* ctx, const float* d_xyz, uint32_t stride_floats, uint32_t n, const void* d_params_f16, void* d_logits_f16,
                                    insta_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * Stage 3 -- training step pieces (ngp::Testbed::train_nerf_step, ngp/src/testbed_nerf.cu:3170-3426).
 * ------------------------------------------------------------------------------------------------------------ */
/* ngp::TrainingImageMetadata + TrainingXForm of one training image (ngp nerf_loader.h:48-64, common.h:168-171) */
typedef struct insta_train_image {
	const void*  d_pixels;         /* RGBA, image_data_type; negative red (0x00FF00FF for bytes) = masked away */
	const void*  d_seg_mask;       /* nullable; RGBA8, red channel = region id (read_seg_mask, common_device.cuh:687) */
	const float* d_depth;          /* nullable */
	const float* d_rays;           /* nullable; explicit per-pixel rays: origin + unnormalised direction, 6 floats */
	uint32_t resolution[2];
	float    focal_length[2];
	float    principal_point[2];
	float    xform[12];            /* camera to world 3x4, column-major (start == end: no rolling shutter) */
	uint32_t image_data_type;      /* ngp EImageDataType: 1 = Byte, 2 = Half, 3 = Float */
	uint32_t reserved_;
} insta_train_image;
/* Testbed::load_nerf -> dataset.metadata_gpu / transforms_gpu (testbed_nerf.cu:2698-2760) */
INSTA_API int insta_training_set_images(insta_ctx* ctx, const insta_train_image* h_images, uint32_t n_images);

/* The buffers and scalars train_nerf_step hands to its kernels.  All d_* are the caller's device buffers. */
typedef struct insta_train_step {
	uint32_t n_rays;               /* n_rays_per_batch, <= 2^18 (testbed_nerf.cu:2910) */
	uint32_t n_rays_total;         /* rays consumed by earlier steps (image choice, testbed_nerf.cu:1102) */
	uint32_t max_samples;          /* capacity of d_coords in samples (max_inference, 16 * 2^18) */
	uint32_t target_batch_size;    /* samples of the compacted batch, 2^18 (testbed.h:857) */
	uint32_t stride_floats;        /* 11 + exp_params */
	uint32_t snap_to_pixel_centers;/* training.snap_to_pixel_centers (true) */
	uint32_t random_bg_color;      /* training.random_bg_color (true) */
	uint32_t penalize_density_alpha; /* "penalize_density_alpha" (core.cu:191) */
	uint64_t rng_seed;             /* Testbed::m_rng = default_rng_t{seed} ... */
	uint64_t rng_advances;         /* ... advanced once per finished step (testbed_nerf.cu:3419) */
	float    cone_angle_constant;
	float    loss_scale;           /* LOSS_SCALE = 128 (testbed.h:256) */
	float    background_color[3];
	float    depth_weight;         /* "depth_weight" -> depth_supervision_lambda (core.cu:208) */
	float    seg_weight;           /* "seg_weight" -> seg_mask_supervision_lambda (core.cu:189) */
	float    near_distance;        /* training.near_distance */
	float    beta_loss;            /* "beta_loss" (core.cu:193) */
	const void*  d_density_grid_bitfield;
	const float* d_extra_dims;     /* n_images x exp_params expression coefficients (extra_dims_gpu) */
	float*    d_coords;            /* max_samples records */
	uint32_t* d_counters;          /* [0] rays accepted, [1] samples generated, [2] samples after compaction */
	uint32_t* d_ray_indices;       /* n_rays */
	float*    d_rays;              /* n_rays x 6 */
	uint32_t* d_numsteps;          /* n_rays x 2: (count, base) */
} insta_train_step;
/* generate_training_samples_nerf (testbed_nerf.cu:1105-1285); zeroes d_counters first */
INSTA_API int insta_generate_training_samples(insta_ctx* ctx, const insta_train_step* step, insta_stream stream);
/* compute_loss_kernel_train_nerf + fill_rollover(_and_rescale) (testbed_nerf.cu:1305-1643, 3353-3361).
 * d_net_out_f16: out_width halves per generated sample (r, g, b, sigma logits first); d_dloss_f16: out_width halves per
 * compacted sample; d_coords_compacted: target_batch_size records; d_loss (nullable): per-ray loss / n_rays;
 * d_mean_density: Testbed::Nerf::density_grid_mean (one float). */
INSTA_API int insta_compute_loss(insta_ctx* ctx, const insta_train_step* step, const void* d_net_out_f16, uint32_t out_width, float* d_coords_compacted,
                                 void* d_dloss_f16, float* d_loss, const float* d_mean_density, insta_stream stream);

/* tcnn::EmaOptimizer{AdamOptimizer} step (tcnn optimizers/adam.h:48-114, 157-190; ema.h:64-77, 103-135) over the trainer's
 * buffers (tcnn trainer.h:282-297): fp32 master params, fp16 params, fp16 gradients (scaled by loss_scale), Adam moments,
 * per-parameter step counters, fp16 EMA params (the inference / snapshot weights). */
typedef struct insta_optimizer {
	float learning_rate, beta1, beta2, epsilon, l2_reg, ema_decay, loss_scale; /* configs/nerf/main.json:21-32 */
	uint32_t current_step;         /* optimizer steps taken before this one */
	float*    d_params_fp32;
	void*     d_params_f16;
	const void* d_gradients_f16;
	float*    d_first_moments;
	float*    d_second_moments;
	uint32_t* d_param_steps;
	void*     d_params_ema_f16;
} insta_optimizer;
INSTA_API int insta_optimizer_step(insta_ctx* ctx, const insta_optimizer* opt, insta_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * One full training step -- ngp::Testbed::train_nerf (ngp/src/testbed_nerf.cu:2916-2992): Counters::prepare_for_training_steps,
 * train_nerf_step (3170-3426: sample generation, rta::Core::surface_closest_point, inference, loss + compaction, forward + backward),
 * the data-parallel gradient exchange (none in the single-GPU reference), tcnn::Trainer::optimizer_step (Adam + EMA) and
 * Counters::update_after_training (2889-2914: adapts rays_per_batch so that the compacted batch approaches target_batch_size).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct insta_trainer_config {
	uint32_t struct_size;            /* = sizeof(insta_trainer_config) */
	uint32_t target_batch_size;      /* 2^18 (ngp testbed.h:857) */
	uint32_t rays_per_batch;         /* initial Counters::rays_per_batch, 2^12 (ngp testbed.h:640) */
	uint32_t snap_to_pixel_centers;  /* training.snap_to_pixel_centers */
	uint32_t random_bg_color;        /* training.random_bg_color */
	uint32_t penalize_density_alpha; /* "penalize_density_alpha" (core.cu:191) */
	uint64_t rng_seed;               /* Testbed::m_rng seed (ngp/src/testbed.cu:2185: 1337) */
	float    cone_angle_constant;
	float    loss_scale;             /* LOSS_SCALE = 128 (ngp testbed.h:256) */
	float    background_color[3];
	float    depth_weight, seg_weight, near_distance, beta_loss; /* core.cu:189-208 */
	/* measurement controls (not in the reference): a fixed workload per step */
	uint32_t fixed_rays_per_batch;   /* 0 = Counters::update_after_training adapts rays_per_batch (reference behaviour); else every step shoots this many rays */
	uint32_t freeze_parameters;      /* 1 = the step does everything except Trainer::optimizer_step (gradients and their allreduce included): the model, hence the
	                                    samples per ray and the compacted batch, stay what they are */
} insta_trainer_config;
typedef struct insta_train_stats {
	uint32_t training_step;          /* steps finished */
	uint32_t n_rays;                 /* rays of THIS step */
	uint32_t n_rays_accepted;        /* rays that hit occupied cells */
	uint32_t rays_per_batch;         /* rays of the NEXT step */
	uint32_t n_rays_total;
	uint32_t measured_batch_size;    /* samples after compaction */
	uint32_t measured_batch_size_before_compaction;
	float    loss;                   /* loss scalar of update_after_training when requested, else 0 */
} insta_train_stats;
INSTA_API void insta_trainer_config_default(insta_trainer_config* cfg);
/* allocates the step's workspace once (the reference takes it from the tcnn arena every step, testbed_nerf.cu:3175-3202): ~0.55 GB at 2^18 */
INSTA_API int insta_trainer_create(insta_ctx* ctx, const insta_trainer_config* cfg);
INSTA_API int insta_trainer_destroy(insta_ctx* ctx);
INSTA_API int insta_trainer_state(const insta_ctx* ctx, insta_train_stats* h_stats);
/* one step.  opt: the trainer's buffers + hyper-parameters (opt->current_step = optimizer steps taken so far; opt->loss_scale is ignored in
 * favour of the trainer's).  d_density_grid_bitfield / d_extra_dims ([n_images][exp_params]) / d_mean_density as in insta_train_step and
 * insta_compute_loss.  Needs insta_training_set_images and frames (frame i deforms the samples of image i).  Synchronises the stream once
 * at the end to read the counters back, like update_after_training. */
INSTA_API int insta_train_nerf(insta_ctx* ctx, const insta_optimizer* opt, const void* d_density_grid_bitfield, const float* d_extra_dims,
                               const float* d_mean_density, int get_loss_scalar, insta_train_stats* h_stats, insta_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * Occupancy-grid maintenance -- ngp::Testbed::training_prep_nerf / update_density_grid_nerf / update_density_grid_mean_and_bitfield
 * (ngp/src/testbed_nerf.cu:3428-3444, 2774-2878) including the INSTA hook on the sampled points (rta::Core::surface_closest_point on
 * bare xyz -> kernel `warp_random_samples`, src/core.cu:528-579; called at testbed_nerf.cu:2850) and NerfNetwork::density
 * (ngp nerf_network.h:385-414).  The grid buffers are the caller's (Testbed::Nerf::density_grid / density_grid_bitfield /
 * density_grid_mean, ngp testbed.h:642-646); the sample scratch is the context's.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct insta_density_grid {
	float*   d_density_grid;          /* 128^3 * n_cascades floats, Morton order */
	void*    d_density_grid_bitfield; /* 8 levels x 128^3 bits = 2 MiB (grid_mip_offset(NERF_CASCADES) / 8, testbed_nerf.cu:78-80) */
	float*   d_density_grid_mean;     /* 1 float */
	uint32_t n_cascades;              /* Testbed::Nerf::max_cascade + 1; 1 for aabb_scale 1 */
	uint32_t ema_step;                /* density_grid_ema_step; advanced by every update */
	uint64_t rng_seed;                /* density_grid_rng = default_rng_t{m_rng.next_uint()} (ngp/src/testbed.cu:2198) */
	uint64_t rng_advances;            /* advanced by every update */
	float    decay;                   /* training.density_grid_decay, 0.95 (ngp testbed.h:617) */
	uint32_t training_step;           /* Testbed::m_training_step: 0 resets the grid, < 256 samples every cell */
	uint32_t has_rays;                /* dataset.has_rays: skip the camera-frustum culling of step 0 */
	uint32_t images_changed;          /* n_images_for_training != n_images_for_training_prev; cleared by the update */
} insta_density_grid;
/* update_density_grid_nerf with explicit sample counts; g->ema_step / rng_advances are advanced. d_params_f16: the TRAINING parameters
 * (density(..., use_inference_params = false)). */
INSTA_API int insta_update_density_grid(insta_ctx* ctx, insta_density_grid* g, const void* d_params_f16, uint32_t n_uniform_samples,
                                        uint32_t n_nonuniform_samples, insta_stream stream);
/* training_prep_nerf with Testbed::train's schedule (ngp/src/testbed.cu:2480-2481): does nothing unless
 * training_step % clamp(training_step / 16, 1, 16) == 0; 128^3 * n_cascades uniform samples before step 256, a quarter uniform + a
 * quarter non-uniform after.  *did_update (nullable) reports whether the grid was refreshed. */
INSTA_API int insta_training_prep_nerf(insta_ctx* ctx, insta_density_grid* g, const void* d_params_f16, int* did_update, insta_stream stream);
/* update_density_grid_mean_and_bitfield alone (e.g. after loading a snapshot's density grid) */
INSTA_API int insta_update_density_grid_mean_and_bitfield(insta_ctx* ctx, const insta_density_grid* g, insta_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * Data-parallel training across the GPUs of one box: one context (process) per GPU, each with its own shard of frames / images and its
 * own ray batches; the only exchange step is the sum of the fp16 gradient buffer between backward and the optimizer step (SURVEY.md 8e;
 * 25,741,888 halves = 51.5 MB), NCCL over NVLink on the caller's stream.  NCCL is resolved at run time (dlopen of libnccl.so.2); without
 * a communicator every call below is a no-op for world size 1.
 * ------------------------------------------------------------------------------------------------------------ */
INSTA_API int insta_comm_unique_id(void* h_id128);  /* ncclGetUniqueId on rank 0; distribute the 128 bytes to the other ranks */
INSTA_API int insta_comm_init_rank(insta_ctx* ctx, int world_size, int rank, const void* h_id128);
INSTA_API int insta_comm_destroy(insta_ctx* ctx);
/* How insta_train_nerf sums the gradients of this context's ranks: 0 = single rank / no communicator, 1 = ncclAllReduce, 2 = this library's own kernel over
 * peer memory (NVLink / NVSwitch; one node, <= 8 ranks, buffers shared with cudaIpc: set up at the first training step, ncclAllReduce when that fails on any
 * rank).  INSTA_GRAD_EXCHANGE = p2p | serial | overlap | none selects the mode at insta_trainer_create. */
INSTA_API int insta_comm_exchange_mode(const insta_ctx* ctx);
/* in-place sum over the ranks; insta_train_nerf calls it itself, pre-scaling each rank's loss by 1 / world_size */
INSTA_API int insta_grad_allreduce(insta_ctx* ctx, void* d_gradients_f16, uint64_t n_halves, insta_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * Host-buffer convenience (the end-to-end path a caller without device buffers uses): H2D of the records, stage 1,
 * stage 2, D2H of the mutated records and of the fp16 encodings, chunked and double-buffered on internal streams.
 * h_coords / h_out_soa_f16 should be pinned (cudaHostAlloc / insta_host_alloc) for full PCIe speed.
 * Blocks until the results are in host memory.
 * ------------------------------------------------------------------------------------------------------------ */
INSTA_API int insta_deform_encode_host(insta_ctx* ctx, float* h_coords, uint32_t n, uint32_t stride_floats, int32_t frame_id,
                                       const void* d_table_f16, void* h_out_soa_f16);
/* The same for a caller that wants what the reference's render / training loops consume after `warp`: deformation (rta::Core::surface_closest_point /
 * render_deformed, src/core.cu:581-637) followed by NerfNetwork::inference_mixed_precision (ngp nerf_network.h:159-232) on host records; out = 4 halves
 * (r, g, b, sigma logits) per sample.  The records are NOT copied back (h_coords is read only): 108 B in, 8 B out per sample over PCIe instead of
 * 108 in + 364 out. */
INSTA_API int insta_deform_infer_host(insta_ctx* ctx, const float* h_coords, uint32_t n, uint32_t stride_floats, int32_t frame_id,
                                      const void* d_params_f16, void* h_out_rgbsigma_f16);
INSTA_API int insta_host_alloc(void** p, uint64_t bytes);
INSTA_API int insta_host_free(void* p);

/* ------------------------------------------------------------------------------------------------------------
 * Host-only utilities (no GPU needed): the host halves of the mesh ingest, exposed so that callers and tests can
 * reuse / inspect them.  Layouts are the reference's: triangles = ngp::Triangle (48 B: a, b, c, scaling, id, mask_id;
 * ngp triangle.cuh:234-238), nodes = ngp::TriangleBvhNode (32 B: min, max, left_idx, right_idx; triangle_bvh.cuh:29).
 * ------------------------------------------------------------------------------------------------------------ */
/* ngp::TriangleBvh::build (ngp/src/triangle_bvh.cu:548-631): permutes h_tris IN PLACE into BVH order and writes the
 * node array; *n_nodes receives the node count (INSTA_ERR_CAPACITY if max_nodes is too small). */
INSTA_API int insta_host_bvh_build(void* h_tris, uint32_t n_tris, uint32_t n_primitives_per_leaf, void* h_nodes, uint32_t max_nodes,
                                   uint32_t* n_nodes);
/* TinyMesh::build_triangles (src/tiny_mesh.cu:134-155); h_canonical_tris NULL = canonical mesh itself (scaling 1) */
INSTA_API int insta_host_build_triangles(const float* h_verts, const uint32_t* h_faces, uint32_t n_faces, const uint32_t* h_mask_ids,
                                         const void* h_canonical_tris, void* h_tris_out);
/* TinyMesh::get_triangle_3_neighbours (src/tiny_mesh.cu:316-371): h_out = n_faces x 3, -1 padded */
INSTA_API int insta_host_adjacency(const uint32_t* h_faces, uint32_t n_faces, uint32_t n_verts, int32_t* h_out);
/* TinyMesh::load + Masking (src/tiny_mesh.cu:53-84, src/masking.cpp:32-180): parses the OBJ (+0.5 offset applied) and
 * the region masks.  Pass NULL outputs with the counts to size the buffers first. */
INSTA_API int insta_host_load_canonical_obj(const char* canonical_obj, const char* mesh_masks_dir, float* h_verts, uint32_t* h_faces,
                                            uint32_t* h_mask_ids, uint32_t* n_verts, uint32_t* n_faces);
/* TinyMesh::m_aabb (src/tiny_mesh.cu:157-162) */
INSTA_API int insta_host_mesh_aabb(const float* h_verts, uint32_t n_verts, float* h_aabb6);

/* ------------------------------------------------------------------------------------------------------------
 * Snapshots in the reference's MessagePack layout -- ngp::Testbed::save_snapshot / load_snapshot (ngp/src/testbed.cu:2916-3016),
 * tcnn::Trainer::serialize / deserialize (tcnn trainer.h:255-279), EmaOptimizer / AdamOptimizer::serialize (tcnn optimizers/ema.h:182-194,
 * adam.h:271-292).  Host-only (no GPU needed): binaries are HOST pointers; copy them to / from the trainer's device buffers yourself.
 * A file written here loads in the reference and vice versa ("nerf"/"dataset" is not written; the reference then keeps its loaded dataset).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct insta_snapshot {
	uint64_t    n_params;              /* snapshot.n_params */
	const void* h_params_f16;          /* snapshot.params_binary: the INFERENCE (= EMA) weights, trainer order (density MLP, colour MLP, grid) */
	uint64_t    n_density_grid;        /* elements of snapshot.density_grid_binary: 128^3 * n_cascades, 0 = never populated */
	const void* h_density_grid_f16;
	uint32_t    training_step;         /* snapshot.training_step (the reference resumes at training_step + 1) */
	float       loss;
	float       aabb[6], render_aabb[6]; /* min xyz, max xyz */
	float       bounding_radius;
	uint32_t    aabb_scale;            /* snapshot.nerf.aabb_scale */
	uint32_t    rays_per_batch, measured_batch_size, measured_batch_size_before_compaction; /* snapshot.nerf.rgb.* */
	/* optimizer: hyper-parameters from the embedded network config; state only when has_optimizer (include_optimizer_state) */
	float       ema_decay, beta1, beta2, epsilon, l2_reg;
	float       opt_base_learning_rate;
	uint32_t    has_optimizer;
	uint32_t    opt_current_step;
	const float*    h_first_moments;   /* n_params; when read from a file: NOT necessarily 4-byte aligned, copy with memcpy / cudaMemcpy */
	const float*    h_second_moments;
	const uint32_t* h_param_steps;     /* NULL in files of older tcnn versions = all zero */
	const void*     h_weights_ema_f16;
} insta_snapshot;
typedef struct insta_snapshot_file insta_snapshot_file;
/* writes "encoding" from p (NULL = defaults), the shipped network blocks, and the "snapshot" object */
INSTA_API int insta_snapshot_save(const char* path, const insta_params* p, const insta_snapshot* s);
/* parses the file; the pointers in *s stay valid until insta_snapshot_close(*file) */
INSTA_API int insta_snapshot_open(const char* path, insta_snapshot_file** file, insta_snapshot* s);
INSTA_API void insta_snapshot_close(insta_snapshot_file* file);

/* rta::Core::raycast_flame (src/core.cu:350-402): the shaded FLAME mesh drawn into the frame buffer -- the viewer's / recorder's mesh overlay.
 * One ray per pixel through the pixel centre (SphereTracer::init_rays_from_camera, ngp src/testbed_sdf.cu:493-577: normalised, advanced to the render box),
 * first hit on the mesh of frame `frame_id` (>= 0: m_meshes[frame], the `m_render_deformed` case; -1: m_canonical_shape) with the reference's BVH traversal
 * (SphereTracer::trace_bvh -> raytrace_kernel -> TriangleBvh4::ray_intersect, ngp src/triangle_bvh.cu:271-312, 721-738; rays start with mint = MAX_DIST = 10),
 * shading as shade_kernel_sdf (src/core.cu:74-109): color_normal = Core::m_raycast_normal (0: three fixed lights around the canonical mesh's centroid, 1: the
 * normal as a colour), alpha 1, depth = z of the hit.  d_frame_depth is set to 1e10 for every pixel first, like the reference; pixels without a hit keep
 * their colour.  Render box: the canonical mesh's (insta_set_canonical) unless opts->use_render_aabb.  cam->snap_to_pixel_centers must be 1. */
INSTA_API int insta_raycast_mesh(insta_ctx* ctx, const insta_camera* cam, const insta_render_opts* opts, int32_t frame_id, int color_normal, float* d_frame_rgba,
                                 float* d_frame_depth, insta_stream stream);
/* TriangleBvh::ray_trace_gpu (ngp src/triangle_bvh.cu:495-512) as a query: n rays (origin, direction: 3 floats each, directions as given), d_tri_ids[i] = topology
 * id of the first triangle hit within MAX_DIST = 10 (-1: none), d_t[i] (nullable) = the ray parameter (10 when nothing is hit).  Same traversal order and
 * IEEE arithmetic as the reference's host build: ids and t are bit-exact against it (tests/test_gpu_raycast.py). */
INSTA_API int insta_mesh_ray_intersect(insta_ctx* ctx, int32_t frame_id, const float* d_origins, const float* d_dirs, uint32_t n, int32_t* d_tri_ids, float* d_t,
                                       insta_stream stream);

/* Measurement aid (not on the hot path): bandwidth of fully divergent gathers of chunk_bytes (16 = one hash-table entry, 128 = one BVH quad
 * read by 8 cooperating lanes) from an internal buffer of buffer_bytes -- L2-resident when it fits the 126 MB L2, HBM-bound otherwise.
 * Best of `iters` launches after warm-up, GB/s of gathered bytes.  bench.py reports it as the L2-gather roofline. */
INSTA_API int insta_microbench_gather(insta_ctx* ctx, uint64_t buffer_bytes, uint32_t chunk_bytes, uint32_t iters, float* h_gbs, insta_stream stream);
/* The same for the scatter side (roofline of kernel_grid_backward, tcnn encodings/grid.h:312-435): every lane adds 8 halves to a pseudo-random 16-byte entry with one
 * vector reduction at L2; chunk_bytes 32 / 64 = 2 / 4 adjacent lanes on adjacent entries.  GB/s of reduced bytes (16 per lane and reduction). */
INSTA_API int insta_microbench_scatter(insta_ctx* ctx, uint64_t buffer_bytes, uint32_t chunk_bytes, uint32_t iters, float* h_gbs, insta_stream stream);

/* Host only: the sharding rule of SURVEY 8e for C / C++ callers (the Python helper insta_b200/shard.py states the same rule).  One process per GPU; rank r of
 * `world` holds the frames r, r + world, r + 2 world, ... (their BVHs, transform tables and images never reach the other GPUs), draws its own ray batches
 * over its frames and renders the frames it holds; the only exchange is insta_train_nerf's gradient sum.
 *   insta_shard_frame_count : how many of n_frames global frames rank holds
 *   insta_shard_owner       : owner rank and local index of a global frame id
 *   insta_shard_local_ids   : per-sample holder values for this rank: local frame index where the sample's frame lives here, -1 (skipped by `warp`) elsewhere */
INSTA_API uint32_t insta_shard_frame_count(uint32_t n_frames, uint32_t rank, uint32_t world);
INSTA_API int insta_shard_owner(uint32_t global_frame, uint32_t world, uint32_t* rank, uint32_t* local_index);
INSTA_API int insta_shard_local_ids(const int32_t* h_global_ids, uint64_t n, uint32_t rank, uint32_t world, int32_t* h_local_ids);

/* number of kernels this library has launched on behalf of ctx since creation (bench.py's gpu_launches) */
INSTA_API uint64_t insta_launch_count(const insta_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* INSTA_B200_H */
