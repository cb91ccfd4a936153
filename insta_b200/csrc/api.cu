// C ABI of libinsta_b200.so (include/insta_b200.h): context, mesh ingest and the thin launch wrappers.
// Host code only; the kernels live in deform.cu and hashgrid.cu.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "ctx.h"

namespace insta {
const char* last_error_cstr();

// The device copy of the frame table (FrameDev per frame) is brought up to date by the ingest entry points (insta_add_frame[s], insta_replace_frames,
// insta_clear_frames), which synchronise anyway; it is allocated once at max_cached_bvh entries, so it never moves while kernels read it.  Caller holds ctx->mu.
static int upload_frame_table(insta_ctx* ctx) {
	const uint32_t n = (uint32_t)ctx->frames.size();
	const uint32_t cap = std::max<uint32_t>(std::max<uint32_t>(ctx->p.max_cached_bvh, n), 64u);
	if (cap > ctx->frame_table_cap) {
		// only when max_cached_bvh was raised behind our back: nothing may be in flight then
		INSTA_CUDA(cudaDeviceSynchronize());
		if (ctx->d_frame_table) cudaFree(ctx->d_frame_table);
		ctx->d_frame_table = nullptr;
		ctx->frame_table_cap = 0;
		INSTA_CUDA(cudaMalloc(&ctx->d_frame_table, sizeof(FrameDev) * cap));
		ctx->frame_table_cap = cap;
	}
	std::vector<FrameDev> host(n);
	for (uint32_t i = 0; i < n; ++i) host[i] = ctx->frames[i].dev;
	ctx->layout_uniform = n > 0;
	uint64_t lo = ~0ull, hi = 0;
	for (uint32_t i = 0; i < n; ++i) {
		const uint64_t q = (uint64_t)host[i].quads, to = (uint64_t)host[i].tris - q, xo = (uint64_t)host[i].xf - q;
		const uint64_t co = host[i].cyl ? (uint64_t)host[i].cyl - q : 0;
		if (i == 0) {
			ctx->tris_off = (uint32_t)to;
			ctx->xf_off = (uint32_t)xo;
			ctx->cyl_off = (uint32_t)co;
		}
		if (co != ctx->cyl_off || (co & 15u) || co >= (1ull << 32)) ctx->cyl_off = 0;  // bounds of some frame missing or elsewhere: traversals run without them
		if (to != ctx->tris_off || xo != ctx->xf_off || (uint64_t)host[i].tris < q || (uint64_t)host[i].xf < (uint64_t)host[i].tris) ctx->layout_uniform = false;
		lo = std::min(lo, q);
		hi = std::max(hi, (uint64_t)host[i].xf + (uint64_t)ctx->n_faces * sizeof(TriXf));
	}
	ctx->arena = (const unsigned char*)lo;
	if (n && ((lo & 15u) || ((uint64_t)ctx->tris_off & 15u) || hi - lo >= (1ull << 36))) ctx->layout_uniform = false;
	if (n) INSTA_CUDA(cudaMemcpy(ctx->d_frame_table, host.data(), sizeof(FrameDev) * n, cudaMemcpyHostToDevice));
	ctx->frame_table_dirty = false;
	return INSTA_OK;
}

// Launch paths only check (no lock, no synchronisation): every ingest call leaves the table current, also when it fails half-way.
int sync_frame_table(insta_ctx* ctx, cudaStream_t) {
	if (ctx->frame_table_dirty) INSTA_FAIL(INSTA_ERR_INVALID, "the device frame table is out of date (an ingest call failed while uploading it)");
	return INSTA_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Cylinder bounds (deform.cu) exist for topologies the compact-code traversal kernel handles; their block sits between a frame's quads and triangles.
static size_t cyl_bytes(const insta_ctx* ctx, size_t n_quads) { return ctx->compact_entries && !ctx->disable_cyl ? align_up(n_quads * kCylStride, 256) : 0; }

// The triangle range of every quad child: a function of the triangle count and the leaf size alone, uploaded with the first frame.
static int ensure_child_ranges(insta_ctx* ctx, const HostBvh& bvh) {
	if (!cyl_bytes(ctx, bvh.quads.size())) return INSTA_OK;
	if (ctx->d_child_ranges && ctx->child_ranges_quads == bvh.quads.size()) return INSTA_OK;
	if (ctx->d_child_ranges) {
		INSTA_CUDA(cudaDeviceSynchronize());
		cudaFree(ctx->d_child_ranges);
		ctx->d_child_ranges = nullptr;
	}
	INSTA_CUDA(cudaMalloc(&ctx->d_child_ranges, bvh.child_ranges.size() * 4));
	INSTA_CUDA(cudaMemcpy(ctx->d_child_ranges, bvh.child_ranges.data(), bvh.child_ranges.size() * 4, cudaMemcpyHostToDevice));
	ctx->child_ranges_quads = (uint32_t)bvh.quads.size();
	return INSTA_OK;
}

// Upload one prepared frame: BVH quads + BVH-order triangles from the host build, transform table computed on device.
static int upload_frame(insta_ctx* ctx, const float* h_verts, const HostBvh& bvh, insta_ctx::Frame& fr) {
	const size_t cyl_b = cyl_bytes(ctx, bvh.quads.size());
	const size_t quads_b = align_up(bvh.quads.size() * sizeof(Quad), 256) + cyl_b;
	const size_t tris_b = align_up(bvh.tris.size() * sizeof(Tri), 256);
	const size_t xf_b = align_up((size_t)ctx->n_faces * sizeof(TriXf), 256);
	const size_t verts_b = align_up((size_t)ctx->n_verts * 12, 256);
	int rcr = ensure_child_ranges(ctx, bvh);
	if (rcr) return rcr;
	INSTA_CUDA(cudaMalloc(&fr.slab, quads_b + tris_b + xf_b + verts_b));
	unsigned char* base = (unsigned char*)fr.slab;
	Quad* d_quads = (Quad*)base;
	Tri* d_tris = (Tri*)(base + quads_b);
	TriXf* d_xf = (TriXf*)(base + quads_b + tris_b);
	float* d_verts = (float*)(base + quads_b + tris_b + xf_b);
	fr.dev.cyl = cyl_b ? base + quads_b - cyl_b : nullptr;
	cudaStream_t s = ctx->streams[0];
	INSTA_CUDA(cudaMemcpyAsync(d_quads, bvh.quads.data(), bvh.quads.size() * sizeof(Quad), cudaMemcpyHostToDevice, s));
	INSTA_CUDA(cudaMemcpyAsync(d_tris, bvh.tris.data(), bvh.tris.size() * sizeof(Tri), cudaMemcpyHostToDevice, s));
	INSTA_CUDA(cudaMemcpyAsync(d_verts, h_verts, (size_t)ctx->n_verts * 12, cudaMemcpyHostToDevice, s));
	int rc = launch_build_xf(ctx, d_verts, d_xf, ctx->d_canon_area, nullptr, s);
	if (rc) return rc;
	fr.dev.quads = d_quads;
	fr.dev.tris = d_tris;
	fr.dev.xf = d_xf;
	rc = launch_build_cyl(ctx, fr.dev, (uint32_t)bvh.quads.size(), s);
	if (rc) return rc;
	INSTA_CUDA(cudaStreamSynchronize(s));
	fr.d_verts = d_verts;
	fr.n_nodes = (uint32_t)bvh.nodes.size();
	fr.n_quads = (uint32_t)bvh.quads.size();
	fr.bvh_order_ids.resize(bvh.tris.size());
	for (size_t i = 0; i < bvh.tris.size(); ++i) fr.bvh_order_ids[i] = bvh.tris[i].id;
	return INSTA_OK;
}

// Upload a batch of prepared frames: one allocation (owned by the batch's first frame), asynchronous copies, one sync.
static int upload_frames(insta_ctx* ctx, const float* h_verts, const std::vector<HostBvh>& bvhs) {
	if (bvhs.empty()) return INSTA_OK;
	const size_t xf_b = align_up((size_t)ctx->n_faces * sizeof(TriXf), 256);
	const size_t verts_b = align_up((size_t)ctx->n_verts * 12, 256);
	std::vector<size_t> off(bvhs.size() + 1, 0);
	for (size_t i = 0; i < bvhs.size(); ++i)
		off[i + 1] = off[i] + align_up(bvhs[i].quads.size() * sizeof(Quad), 256) + cyl_bytes(ctx, bvhs[i].quads.size()) + align_up(bvhs[i].tris.size() * sizeof(Tri), 256) + xf_b + verts_b;
	int rcr = ensure_child_ranges(ctx, bvhs[0]);
	if (rcr) return rcr;
	void* slab = nullptr;
	INSTA_CUDA(cudaMalloc(&slab, off.back()));
	cudaStream_t s = ctx->streams[0];
	std::vector<insta_ctx::Frame> frames(bvhs.size());
	int rc = INSTA_OK;
	cudaError_t e = cudaSuccess;
	for (size_t i = 0; i < bvhs.size() && !rc && e == cudaSuccess; ++i) {
		const HostBvh& bvh = bvhs[i];
		unsigned char* base = (unsigned char*)slab + off[i];
		const size_t cyl_b = cyl_bytes(ctx, bvh.quads.size());
		const size_t quads_b = align_up(bvh.quads.size() * sizeof(Quad), 256) + cyl_b, tris_b = align_up(bvh.tris.size() * sizeof(Tri), 256);
		Quad* d_quads = (Quad*)base;
		Tri* d_tris = (Tri*)(base + quads_b);
		TriXf* d_xf = (TriXf*)(base + quads_b + tris_b);
		float* d_verts = (float*)(base + quads_b + tris_b + xf_b);
		const float* hv = h_verts + (size_t)ctx->n_verts * 3 * i;
		e = cudaMemcpyAsync(d_quads, bvh.quads.data(), bvh.quads.size() * sizeof(Quad), cudaMemcpyHostToDevice, s);
		if (e == cudaSuccess) e = cudaMemcpyAsync(d_tris, bvh.tris.data(), bvh.tris.size() * sizeof(Tri), cudaMemcpyHostToDevice, s);
		if (e == cudaSuccess) e = cudaMemcpyAsync(d_verts, hv, (size_t)ctx->n_verts * 12, cudaMemcpyHostToDevice, s);
		if (e != cudaSuccess) break;
		rc = launch_build_xf(ctx, d_verts, d_xf, ctx->d_canon_area, nullptr, s);
		insta_ctx::Frame& fr = frames[i];
		fr.slab = i == 0 ? slab : nullptr;
		fr.dev.quads = d_quads;
		fr.dev.tris = d_tris;
		fr.dev.xf = d_xf;
		fr.dev.cyl = cyl_b ? base + quads_b - cyl_b : nullptr;
		if (!rc) rc = launch_build_cyl(ctx, fr.dev, (uint32_t)bvh.quads.size(), s);
		fr.d_verts = d_verts;
		fr.n_nodes = (uint32_t)bvh.nodes.size();
		fr.n_quads = (uint32_t)bvh.quads.size();
		fr.bvh_order_ids.resize(bvh.tris.size());
		for (size_t k = 0; k < bvh.tris.size(); ++k) fr.bvh_order_ids[k] = bvh.tris[k].id;
	}
	const cudaError_t es = cudaStreamSynchronize(s);
	if (rc || e != cudaSuccess || es != cudaSuccess) {
		cudaFree(slab);
		if (rc) return rc;
		INSTA_CUDA(e != cudaSuccess ? e : es);
	}
	for (auto& fr : frames) ctx->frames.push_back(std::move(fr));
	return INSTA_OK;
}

static void prepare_frame(const insta_ctx* ctx, const float* h_verts, HostBvh& bvh) {
	std::vector<Tri> tris(ctx->n_faces);
	build_triangles(h_verts, ctx->faces.data(), ctx->n_faces, ctx->mask_ids.data(), ctx->canon_tris.data(), tris.data());
	build_bvh(tris, ctx->p.n_primitives_per_leaf, bvh);
}

static int ensure_stage(insta_ctx* ctx, uint32_t chunk, uint32_t stride) {
	if (ctx->stage_cap >= chunk && ctx->stage_stride >= stride) return INSTA_OK;
	for (int k = 0; k < 2; ++k) {
		if (ctx->d_stage_rec[k]) cudaFree(ctx->d_stage_rec[k]);
		if (ctx->d_stage_enc[k]) cudaFree(ctx->d_stage_enc[k]);
		ctx->d_stage_rec[k] = nullptr;
		ctx->d_stage_enc[k] = nullptr;
	}
	for (int k = 0; k < 2; ++k) {
		INSTA_CUDA(cudaMalloc(&ctx->d_stage_rec[k], (size_t)chunk * stride * 4));
		INSTA_CUDA(cudaMalloc(&ctx->d_stage_enc[k], (size_t)chunk * ctx->levels.n_levels * 8 * 2));
	}
	ctx->stage_cap = chunk;
	ctx->stage_stride = stride;
	return INSTA_OK;
}

// TinyMesh::load (+0.5 offset) and Masking on the canonical OBJ; host only.
static int load_canonical_obj_host(const char* canonical_obj, const char* mesh_masks_dir, std::vector<float>& verts, std::vector<uint32_t>& faces,
                                   std::vector<uint32_t>& masks) {
	std::vector<float> colors;
	std::string err;
	if (!load_obj(canonical_obj, verts, colors, faces, err)) INSTA_FAIL(INSTA_ERR_IO, err);
	for (float& v : verts) v = v * 1.0f + 0.5f;  // MESH_SCALE = 1, offset 0.5 (tiny_mesh.cu:63-74)
	masks.clear();
	if (mesh_masks_dir) {
		// regions loaded by Masking::mesh_to_mask (masking.cpp:150-157) with their FlameMasks enum values
		static const std::pair<const char*, int> kRegions[] = {{"boundary", 0}, {"eye_region", 1}, {"face", 2}, {"left_ear", 4}, {"lips", 7}, {"right_ear", 10}};
		std::vector<std::pair<int, std::vector<uint32_t>>> regions;
		for (const auto& r : kRegions) {
			std::vector<float> v2, c2;
			std::vector<uint32_t> f2;
			const std::string path = std::string(mesh_masks_dir) + "/" + r.first + ".obj";
			if (!load_obj(path.c_str(), v2, c2, f2, err)) INSTA_FAIL(INSTA_ERR_IO, err);
			std::vector<uint32_t> red;
			for (size_t i = 0; i + 2 < c2.size(); i += 3)
				if (c2[i] >= 1.0f && c2[i + 1] == 0.0f && c2[i + 2] == 0.0f) red.push_back((uint32_t)(i / 3));
			regions.emplace_back(r.second, std::move(red));
		}
		masks_from_regions(faces.data(), (uint32_t)(faces.size() / 3), regions, masks);
	}
	return INSTA_OK;
}

}  // namespace insta

using namespace insta;

extern "C" {

void insta_params_default(insta_params* p) {
	if (!p) return;
	memset(p, 0, sizeof(*p));
	p->struct_size = sizeof(insta_params);
	p->disable_interp = 0;
	p->max_cached_bvh = 4000;  // configs/nerf/insta.json:4
	p->exp_params = 16;        // configs/nerf/main.json:17
	p->geo_params = 0;
	p->n_primitives_per_leaf = 2;
	for (int k = 0; k < 3; ++k) {
		p->aabb_min[k] = 0.f;
		p->aabb_max[k] = 1.f;
	}
	p->n_levels = 16;
	p->n_features_per_level = 8;
	p->log2_hashmap_size = 18;
	p->base_resolution = 16;
	p->per_level_scale = 0.f;
	p->aabb_scale = 1;
	p->rng_seed = 42;
	p->arith = 0;
}

const char* insta_last_error(void) { return last_error_cstr(); }
const char* insta_version(void) { return "insta_b200 0.1 (sm_100a)"; }

int insta_ctx_create(const insta_params* p, int device, insta_ctx** out) {
	if (!out) INSTA_FAIL(INSTA_ERR_INVALID, "out is null");
	*out = nullptr;
	insta_params def;
	insta_params_default(&def);
	if (p) {
		if (p->struct_size != sizeof(insta_params)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_params.struct_size mismatch");
		def = *p;
	}
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
		cudaGetLastError();
		INSTA_FAIL(INSTA_ERR_NO_DEVICE, "no CUDA device visible: insta_b200 has no CPU fallback");
	}
	if (device < 0) INSTA_CUDA(cudaGetDevice(&device));
	INSTA_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	INSTA_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) INSTA_FAIL(INSTA_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", this library carries sm_100a code only");
	if (def.arith > 1) INSTA_FAIL(INSTA_ERR_INVALID, "insta_params.arith must be 0 (reference CUDA build) or 1 (reference host build)");
	// 0 never terminates the host BVH build (a 1-triangle range is split for ever); the leaf references carry the count in 3 bits
	if (def.n_primitives_per_leaf < 1 || def.n_primitives_per_leaf > 7) INSTA_FAIL(INSTA_ERR_INVALID, "insta_params.n_primitives_per_leaf must be in 1..7");
	if (def.max_cached_bvh == 0) INSTA_FAIL(INSTA_ERR_INVALID, "insta_params.max_cached_bvh must be positive");
	insta_ctx* ctx = new insta_ctx();
	ctx->p = def;
	ctx->device = device;
	ctx->sm_count = prop.multiProcessorCount;
	for (int k = 0; k < 3; ++k) {
		ctx->aabb[k] = def.aabb_min[k];
		ctx->aabb[3 + k] = def.aabb_max[k];
	}
	for (int k = 0; k < 2; ++k) {
		if (cudaStreamCreateWithFlags(&ctx->streams[k], cudaStreamNonBlocking) != cudaSuccess ||
		    cudaEventCreateWithFlags(&ctx->events[k], cudaEventDisableTiming) != cudaSuccess) {
			set_error("stream/event creation failed");
			insta_ctx_destroy(ctx);
			return INSTA_ERR_CUDA;
		}
	}
	{
		// launch_deform takes its per-launch scratch from the device's default stream-ordered pool: keep freed blocks cached instead of returning them to
		// the driver at every synchronisation
		cudaMemPool_t pool = nullptr;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
			uint64_t keep = ~0ull;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
		cudaGetLastError();
	}
	if (cudaMalloc(&ctx->d_counter, 32 * 128) != cudaSuccess) {
		set_error("counter allocation failed");
		insta_ctx_destroy(ctx);
		return INSTA_ERR_CUDA;
	}
	{
		const char* k = getenv("INSTA_DEFORM_KERNEL");
		ctx->force_legacy_deform = k && std::string(k) == "legacy";
		ctx->force_wq_deform = k && (std::string(k) == "wq" || std::string(k) == "wq1");
		ctx->no_prefetch_deform = k && std::string(k) == "wq1";
		if (const char* tr = getenv("INSTA_TRAIN_ROUNDS")) ctx->train_all_samples = std::string(tr) == "0";
		if (const char* te = getenv("INSTA_TRAIN_ROUND_EDGES")) {  // e.g. "4,16,48": rounds [0,4) [4,16) [16,48) [48,end)
			std::vector<uint32_t> e{0u};
			for (const char* p = te; *p && e.size() < 7;) {
				char* end = nullptr;
				const unsigned long v = strtoul(p, &end, 10);
				if (end == p) break;
				if (v > e.back() && v < 0xffffffffu) e.push_back((uint32_t)v);
				p = *end ? end + 1 : end;
			}
			e.push_back(0xffffffffu);
			ctx->train_round_edges = e;
		}
		if (const char* rr = getenv("INSTA_RENDER_RECORDS")) ctx->render_full_records = std::string(rr) == "full";
		if (const char* sd = getenv("INSTA_DEFORM_SEED")) ctx->disable_seed_bound = std::string(sd) == "0";
		if (const char* sd = getenv("INSTA_DEFORM_CYL")) ctx->disable_cyl = std::string(sd) == "0";
		if (const char* m = getenv("INSTA_DEFORM_COHERENT_MAX")) ctx->coherent_max_n = (uint32_t)strtoul(m, nullptr, 10);
	}
	int rc = init_grid_levels(ctx);
	if (rc) {
		insta_ctx_destroy(ctx);
		return rc;
	}
	*out = ctx;
	return INSTA_OK;
}

void insta_ctx_destroy(insta_ctx* ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	for (auto& f : ctx->frames) {
		if (f.slab) cudaFree(f.slab);
		if (f.d_order) cudaFree(f.d_order);
	}
	if (ctx->d_frame_table) cudaFree(ctx->d_frame_table);
	if (ctx->d_faces) cudaFree(ctx->d_faces);
	if (ctx->d_mask_ids) cudaFree(ctx->d_mask_ids);
	if (ctx->d_adjacency) cudaFree(ctx->d_adjacency);
	if (ctx->d_canon_area) cudaFree(ctx->d_canon_area);
	if (ctx->d_canon_xf) cudaFree(ctx->d_canon_xf);
	if (ctx->d_seed_grid) cudaFree(ctx->d_seed_grid);
	if (ctx->d_child_ranges) cudaFree(ctx->d_child_ranges);
	if (ctx->d_counter) cudaFree(ctx->d_counter);
	tracer_destroy(ctx);
	recorder_destroy(ctx);
	network_bwd_destroy(ctx);
	if (ctx->d_adam_lr_table) cudaFree(ctx->d_adam_lr_table);
	if (ctx->d_canon_quads) cudaFree(ctx->d_canon_quads);
	if (ctx->d_canon_tris_bvh) cudaFree(ctx->d_canon_tris_bvh);
	trainer_destroy(ctx);
	comm_destroy(ctx);
	occupancy_destroy(ctx);
	if (ctx->d_train_images) cudaFree(ctx->d_train_images);
	for (int k = 0; k < 2; ++k) {
		if (ctx->d_stage_rec[k]) cudaFree(ctx->d_stage_rec[k]);
		if (ctx->d_stage_enc[k]) cudaFree(ctx->d_stage_enc[k]);
		if (ctx->streams[k]) cudaStreamDestroy(ctx->streams[k]);
		if (ctx->events[k]) cudaEventDestroy(ctx->events[k]);
	}
	delete ctx;
}

// ---------------------------------------------------------------------------------------------------------------
int insta_set_canonical(insta_ctx* ctx, const float* h_verts, uint32_t n_verts, const uint32_t* h_faces, uint32_t n_faces,
                        const uint32_t* h_mask_ids) {
	if (!ctx || !h_verts || !h_faces || n_verts == 0 || n_faces == 0) INSTA_FAIL(INSTA_ERR_INVALID, "insta_set_canonical: bad arguments");
	for (uint32_t i = 0; i < n_faces * 3; ++i)
		if (h_faces[i] >= n_verts) INSTA_FAIL(INSTA_ERR_INVALID, "insta_set_canonical: face index out of range");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	int rc = insta_clear_frames(ctx);
	if (rc) return rc;
	ctx->n_verts = n_verts;
	ctx->n_faces = n_faces;
	ctx->canon_verts.assign(h_verts, h_verts + (size_t)n_verts * 3);
	{
		// TinyMesh::m_centorid (tiny_mesh.cu:64-86): the vertices summed in file order in single precision, divided by their number; a new canonical mesh also
		// invalidates its ray-casting BVH (raycast.cu)
		float c[3] = {0.f, 0.f, 0.f};
		for (uint32_t i = 0; i < n_verts; ++i)
			for (int d = 0; d < 3; ++d) c[d] += h_verts[3 * (size_t)i + d];
		for (int d = 0; d < 3; ++d) ctx->canon_centroid[d] = c[d] / (float)n_verts;
		if (ctx->d_canon_quads) cudaFree(ctx->d_canon_quads);
		if (ctx->d_canon_tris_bvh) cudaFree(ctx->d_canon_tris_bvh);
		ctx->d_canon_quads = nullptr;
		ctx->d_canon_tris_bvh = nullptr;
	}
	ctx->faces.assign(h_faces, h_faces + (size_t)n_faces * 3);
	if (h_mask_ids) ctx->mask_ids.assign(h_mask_ids, h_mask_ids + n_faces);
	else ctx->mask_ids.assign(n_faces, 15u);
	ctx->canon_tris.resize(n_faces);
	build_triangles(h_verts, h_faces, n_faces, ctx->mask_ids.data(), nullptr, ctx->canon_tris.data());
	build_adjacency(h_faces, n_faces, n_verts, ctx->adjacency);
	mesh_aabb(h_verts, n_verts, ctx->render_aabb);
	{
		// inner levels of the 4-ary median-split tree: a node with more than leaf_size triangles is split into 4
		uint32_t levels = 0;
		for (uint64_t cnt = n_faces; cnt > ctx->p.n_primitives_per_leaf || levels == 0; cnt = (cnt + 3) / 4) ++levels;
		ctx->stack_need = 3 * (levels > 0 ? levels - 1 : 0) + 4;
		const char* wide = getenv("INSTA_DEFORM_STACK");
		ctx->compact_entries = n_faces < 8192u && ctx->p.n_primitives_per_leaf <= 3u && !(wide && std::string(wide) == "wide");
	}

	for (void* p : {(void*)ctx->d_faces, (void*)ctx->d_mask_ids, (void*)ctx->d_adjacency, (void*)ctx->d_canon_area, (void*)ctx->d_canon_xf})
		if (p) cudaFree(p);
	ctx->d_faces = nullptr; ctx->d_mask_ids = nullptr; ctx->d_adjacency = nullptr; ctx->d_canon_area = nullptr; ctx->d_canon_xf = nullptr;
	INSTA_CUDA(cudaMalloc(&ctx->d_faces, (size_t)n_faces * 12));
	INSTA_CUDA(cudaMalloc(&ctx->d_mask_ids, (size_t)n_faces * 4));
	INSTA_CUDA(cudaMalloc(&ctx->d_adjacency, (size_t)n_faces * 12));
	INSTA_CUDA(cudaMalloc(&ctx->d_canon_area, (size_t)n_faces * 4));
	INSTA_CUDA(cudaMalloc(&ctx->d_canon_xf, (size_t)n_faces * sizeof(TriXf)));
	float* d_verts = nullptr;
	INSTA_CUDA(cudaMalloc(&d_verts, (size_t)n_verts * 12));
	cudaStream_t s = ctx->streams[0];
	INSTA_CUDA(cudaMemcpyAsync(ctx->d_faces, h_faces, (size_t)n_faces * 12, cudaMemcpyHostToDevice, s));
	INSTA_CUDA(cudaMemcpyAsync(ctx->d_mask_ids, ctx->mask_ids.data(), (size_t)n_faces * 4, cudaMemcpyHostToDevice, s));
	INSTA_CUDA(cudaMemcpyAsync(ctx->d_adjacency, ctx->adjacency.data(), (size_t)n_faces * 12, cudaMemcpyHostToDevice, s));
	INSTA_CUDA(cudaMemcpyAsync(d_verts, h_verts, (size_t)n_verts * 12, cudaMemcpyHostToDevice, s));
	rc = launch_build_xf(ctx, d_verts, ctx->d_canon_xf, nullptr, ctx->d_canon_area, s);
	if (!rc) rc = launch_build_seed_grid(ctx, s);
	cudaError_t e = cudaStreamSynchronize(s);
	cudaFree(d_verts);
	if (rc) return rc;
	INSTA_CUDA(e);
	return INSTA_OK;
}

int insta_load_canonical_obj(insta_ctx* ctx, const char* canonical_obj, const char* mesh_masks_dir) {
	if (!ctx || !canonical_obj) INSTA_FAIL(INSTA_ERR_INVALID, "insta_load_canonical_obj: bad arguments");
	std::vector<float> verts;
	std::vector<uint32_t> faces, masks;
	int rc = load_canonical_obj_host(canonical_obj, mesh_masks_dir, verts, faces, masks);
	if (rc) return rc;
	return insta_set_canonical(ctx, verts.data(), (uint32_t)(verts.size() / 3), faces.data(), (uint32_t)(faces.size() / 3),
	                           masks.empty() ? nullptr : masks.data());
}

int insta_add_frame(insta_ctx* ctx, const float* h_verts, int32_t* frame_index) {
	if (!ctx || !h_verts) INSTA_FAIL(INSTA_ERR_INVALID, "insta_add_frame: bad arguments");
	if (ctx->n_faces == 0) INSTA_FAIL(INSTA_ERR_INVALID, "insta_add_frame: set the canonical mesh first");
	if (ctx->frames.size() >= ctx->p.max_cached_bvh) INSTA_FAIL(INSTA_ERR_CAPACITY, "insta_add_frame: max_cached_bvh frames already resident");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	HostBvh bvh;
	prepare_frame(ctx, h_verts, bvh);
	insta_ctx::Frame fr;
	int rc = upload_frame(ctx, h_verts, bvh, fr);
	if (rc) {
		if (fr.slab) cudaFree(fr.slab);
		return rc;
	}
	std::lock_guard<std::mutex> lock(ctx->mu);
	ctx->frames.push_back(std::move(fr));
	ctx->frame_table_dirty = true;
	if (frame_index) *frame_index = (int32_t)ctx->frames.size() - 1;
	return upload_frame_table(ctx);
}

int insta_add_frames(insta_ctx* ctx, const float* h_verts, uint32_t n, uint32_t n_threads) {
	if (!ctx || !h_verts) INSTA_FAIL(INSTA_ERR_INVALID, "insta_add_frames: bad arguments");
	if (ctx->n_faces == 0) INSTA_FAIL(INSTA_ERR_INVALID, "insta_add_frames: set the canonical mesh first");
	if (ctx->frames.size() + n > ctx->p.max_cached_bvh) INSTA_FAIL(INSTA_ERR_CAPACITY, "insta_add_frames: exceeds max_cached_bvh");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
	n_threads = std::min(n_threads, std::max(1u, n));
	const size_t per = (size_t)ctx->n_verts * 3;
	// Host BVH builds (the reference's build order needs libstdc++'s nth_element, SURVEY hard part 1) run on a thread pool one batch AHEAD of the uploads:
	// while batch k+1 is being built, batch k goes to the device as ONE allocation with asynchronous copies and one stream sync (a cudaMalloc, three
	// blocking copies and a sync per frame made the upload, not the build, the bottleneck: 523 frames/s on 16 threads).
	const uint32_t batch = std::max(32u, n_threads * 4);
	std::atomic<bool> build_failed{false};  // an exception must not leave a worker thread (std::terminate): it becomes an error return
	auto build_batch = [&](uint32_t b0, uint32_t b1, std::vector<HostBvh>& out) {
		try {
			out.clear();
			out.resize(b1 - b0);
		} catch (...) {
			build_failed = true;
			return;
		}
		std::atomic<uint32_t> next{b0};
		std::vector<std::thread> pool;
		const uint32_t nt = std::min(n_threads, b1 - b0);
		auto work = [&]() {
			try {
				for (uint32_t i; (i = next.fetch_add(1)) < b1;) prepare_frame(ctx, h_verts + per * i, out[i - b0]);
			} catch (...) {
				build_failed = true;
			}
		};
		try {
			for (uint32_t t = 1; t < nt; ++t) pool.emplace_back(work);
		} catch (...) {  // thread creation failed: the calling thread finishes the batch alone
		}
		work();
		for (auto& t : pool) t.join();
	};
	std::lock_guard<std::mutex> lock(ctx->mu);
	std::vector<HostBvh> cur, nxt;
	build_batch(0, std::min(n, batch), cur);
	int rc = INSTA_OK;
	for (uint32_t b0 = 0; b0 < n && !rc; b0 += batch) {
		const uint32_t b1 = std::min(n, b0 + batch), b2 = std::min(n, b1 + batch);
		std::thread builder;
		if (b1 < n) builder = std::thread([&]() { build_batch(b1, b2, nxt); });
		if (!build_failed) {
			rc = upload_frames(ctx, h_verts + per * b0, cur);
			ctx->frame_table_dirty = true;  // earlier batches stay appended when a later one fails: the table must follow ctx->frames on every exit
		}
		if (builder.joinable()) builder.join();
		if (build_failed && !rc) {
			set_error("insta_add_frames: a host BVH build failed (out of memory?)");
			rc = INSTA_ERR_INVALID;
		}
		cur.swap(nxt);
	}
	const int rc_table = upload_frame_table(ctx);
	return rc ? rc : rc_table;
}

int insta_add_frame_obj(insta_ctx* ctx, const char* obj_path, int32_t* frame_index) {
	if (!ctx || !obj_path) INSTA_FAIL(INSTA_ERR_INVALID, "insta_add_frame_obj: bad arguments");
	std::vector<float> verts, colors;
	std::vector<uint32_t> faces;
	std::string err;
	if (!load_obj(obj_path, verts, colors, faces, err)) INSTA_FAIL(INSTA_ERR_IO, err);
	if (verts.size() / 3 != ctx->n_verts) INSTA_FAIL(INSTA_ERR_INVALID, "insta_add_frame_obj: vertex count differs from the canonical mesh");
	for (float& v : verts) v = v * 1.0f + 0.5f;
	return insta_add_frame(ctx, verts.data(), frame_index);
}

int insta_clear_frames(insta_ctx* ctx) {
	if (!ctx) INSTA_FAIL(INSTA_ERR_INVALID, "ctx is null");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	INSTA_CUDA(cudaDeviceSynchronize());
	std::lock_guard<std::mutex> lock(ctx->mu);
	for (auto& f : ctx->frames) {
		if (f.slab) cudaFree(f.slab);
		if (f.d_order) cudaFree(f.d_order);
	}
	ctx->frames.clear();
	ctx->frame_table_dirty = true;
	return upload_frame_table(ctx);
}

static int check_frame_indices(const insta_ctx* ctx, const int32_t* idx, uint32_t n, const char* who) {
	std::vector<char> seen(ctx->frames.size(), 0);
	for (uint32_t i = 0; i < n; ++i) {
		if (idx[i] < 0 || (size_t)idx[i] >= ctx->frames.size()) INSTA_FAIL(INSTA_ERR_INVALID, std::string(who) + ": frame index out of range");
		if (seen[idx[i]]) INSTA_FAIL(INSTA_ERR_INVALID, std::string(who) + ": frame index given twice");
		seen[idx[i]] = 1;
	}
	return INSTA_OK;
}

int insta_replace_frames(insta_ctx* ctx, const int32_t* h_frame_indices, const float* h_verts, uint32_t n, uint32_t n_threads) {
	if (!ctx || !h_frame_indices || !h_verts) INSTA_FAIL(INSTA_ERR_INVALID, "insta_replace_frames: bad arguments");
	if (n == 0) return INSTA_OK;
	INSTA_CUDA(cudaSetDevice(ctx->device));
	std::lock_guard<std::mutex> lock(ctx->mu);
	int rc = check_frame_indices(ctx, h_frame_indices, n, "insta_replace_frames");
	if (rc) return rc;
	if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
	n_threads = std::min(n_threads, n);
	const size_t per = (size_t)ctx->n_verts * 3;
	// host builds on the pool (batches of 64 frames, the upload of a batch overlaps the build of the next one)
	const uint32_t batch = std::max(32u, n_threads * 4);
	std::atomic<bool> build_failed{false};
	auto build_batch = [&](uint32_t b0, uint32_t b1, std::vector<HostBvh>& out) {
		try {
			out.clear();
			out.resize(b1 - b0);
		} catch (...) {
			build_failed = true;
			return;
		}
		std::atomic<uint32_t> next{b0};
		std::vector<std::thread> pool;
		auto work = [&]() {
			try {
				for (uint32_t i; (i = next.fetch_add(1)) < b1;) prepare_frame(ctx, h_verts + per * i, out[i - b0]);
			} catch (...) {
				build_failed = true;
			}
		};
		try {
			for (uint32_t t = 1; t < std::min(n_threads, b1 - b0); ++t) pool.emplace_back(work);
		} catch (...) {
		}
		work();
		for (auto& t : pool) t.join();
	};
	std::vector<HostBvh> cur, nxt;
	build_batch(0, std::min(n, batch), cur);
	INSTA_CUDA(cudaDeviceSynchronize());  // nothing may still read the frames that are about to change (reload_training_data does the same, core.cu:696)
	cudaStream_t s = ctx->streams[0];
	for (uint32_t b0 = 0; b0 < n && !rc; b0 += batch) {
		const uint32_t b1 = std::min(n, b0 + batch), b2 = std::min(n, b1 + batch);
		std::thread builder;
		if (b1 < n) builder = std::thread([&]() { build_batch(b1, b2, nxt); });
		for (uint32_t i = b0; i < b1 && !rc && !build_failed; ++i) {
			insta_ctx::Frame& fr = ctx->frames[h_frame_indices[i]];
			const HostBvh& bvh = cur[i - b0];
			if (bvh.quads.size() != fr.n_quads || bvh.tris.size() != (size_t)ctx->n_faces) {
				set_error("insta_replace_frames: the new tree does not fit the frame's storage");
				rc = INSTA_ERR_INVALID;
				break;
			}
			cudaError_t e = cudaMemcpyAsync(const_cast<Quad*>(fr.dev.quads), bvh.quads.data(), bvh.quads.size() * sizeof(Quad), cudaMemcpyHostToDevice, s);
			if (e == cudaSuccess) e = cudaMemcpyAsync(const_cast<Tri*>(fr.dev.tris), bvh.tris.data(), bvh.tris.size() * sizeof(Tri), cudaMemcpyHostToDevice, s);
			if (e == cudaSuccess) e = cudaMemcpyAsync(fr.d_verts, h_verts + per * i, per * 4, cudaMemcpyHostToDevice, s);
			if (e != cudaSuccess) {
				set_error(std::string("insta_replace_frames: upload failed: ") + cudaGetErrorString(e));
				rc = INSTA_ERR_CUDA;
				break;
			}
			rc = launch_build_xf(ctx, fr.d_verts, const_cast<TriXf*>(fr.dev.xf), ctx->d_canon_area, nullptr, s);
			if (!rc) rc = launch_build_cyl(ctx, fr.dev, fr.n_quads, s);
			for (size_t k = 0; k < bvh.tris.size(); ++k) fr.bvh_order_ids[k] = bvh.tris[k].id;
			if (fr.d_order) {  // a refit would now follow a stale order
				cudaFree(fr.d_order);
				fr.d_order = nullptr;
			}
		}
		const cudaError_t es = cudaStreamSynchronize(s);  // `cur` is pageable and is reused by the next build
		if (builder.joinable()) builder.join();
		if (!rc && es != cudaSuccess) INSTA_CUDA(es);
		if (build_failed && !rc) {
			set_error("insta_replace_frames: a host BVH build failed (out of memory?)");
			rc = INSTA_ERR_INVALID;
		}
		cur.swap(nxt);
	}
	return rc;
}

int insta_refit_frames(insta_ctx* ctx, const int32_t* h_frame_indices, const float* h_verts, uint32_t n, insta_stream stream) {
	if (!ctx || !h_frame_indices || !h_verts) INSTA_FAIL(INSTA_ERR_INVALID, "insta_refit_frames: bad arguments");
	if (n == 0) return INSTA_OK;
	INSTA_CUDA(cudaSetDevice(ctx->device));
	std::lock_guard<std::mutex> lock(ctx->mu);
	int rc = check_frame_indices(ctx, h_frame_indices, n, "insta_refit_frames");
	if (rc) return rc;
	cudaStream_t s = (cudaStream_t)stream;
	const size_t per = (size_t)ctx->n_verts * 3;
	for (uint32_t i = 0; i < n; ++i) {
		insta_ctx::Frame& fr = ctx->frames[h_frame_indices[i]];
		if (!fr.d_order) {
			INSTA_CUDA(cudaMalloc(&fr.d_order, fr.bvh_order_ids.size() * 4));
			INSTA_CUDA(cudaMemcpy(fr.d_order, fr.bvh_order_ids.data(), fr.bvh_order_ids.size() * 4, cudaMemcpyHostToDevice));
		}
		INSTA_CUDA(cudaMemcpyAsync(fr.d_verts, h_verts + per * i, per * 4, cudaMemcpyHostToDevice, s));
		rc = launch_build_xf(ctx, fr.d_verts, const_cast<TriXf*>(fr.dev.xf), ctx->d_canon_area, nullptr, s);
		if (!rc) rc = launch_refit(ctx, fr.dev, fr.n_quads, fr.d_verts, fr.d_order, s);
		if (!rc) rc = launch_build_cyl(ctx, fr.dev, fr.n_quads, s);
		if (rc) return rc;
	}
	return INSTA_OK;
}

uint32_t insta_n_frames(const insta_ctx* ctx) { return ctx ? (uint32_t)ctx->frames.size() : 0; }

int insta_get_render_aabb(const insta_ctx* ctx, float* h_aabb6) {
	if (!ctx || !h_aabb6) INSTA_FAIL(INSTA_ERR_INVALID, "bad arguments");
	memcpy(h_aabb6, ctx->render_aabb, 24);
	return INSTA_OK;
}

int insta_set_aabb(insta_ctx* ctx, const float* h_aabb6) {
	if (!ctx || !h_aabb6) INSTA_FAIL(INSTA_ERR_INVALID, "bad arguments");
	memcpy(ctx->aabb, h_aabb6, 24);
	return INSTA_OK;
}

int insta_get_topology(const insta_ctx* ctx, uint32_t* h_mask_ids, int32_t* h_adjacency) {
	if (!ctx || ctx->n_faces == 0) INSTA_FAIL(INSTA_ERR_INVALID, "no canonical mesh");
	if (h_mask_ids) memcpy(h_mask_ids, ctx->mask_ids.data(), (size_t)ctx->n_faces * 4);
	if (h_adjacency) memcpy(h_adjacency, ctx->adjacency.data(), (size_t)ctx->n_faces * 12);
	return INSTA_OK;
}

int insta_get_frame_bvh_order(const insta_ctx* ctx, int32_t frame, uint32_t* h_tri_ids, uint32_t* n_nodes) {
	if (!ctx || frame < 0 || (size_t)frame >= ctx->frames.size()) INSTA_FAIL(INSTA_ERR_INVALID, "frame out of range");
	const auto& f = ctx->frames[frame];
	if (h_tri_ids) memcpy(h_tri_ids, f.bvh_order_ids.data(), f.bvh_order_ids.size() * 4);
	if (n_nodes) *n_nodes = f.n_nodes;
	return INSTA_OK;
}

int insta_get_frame_cylinders(insta_ctx* ctx, int32_t frame, uint16_t* h_fp16_bits, uint32_t* n_quads) {
	if (!ctx || frame < 0 || (size_t)frame >= ctx->frames.size() || !n_quads) INSTA_FAIL(INSTA_ERR_INVALID, "insta_get_frame_cylinders: bad arguments");
	const auto& f = ctx->frames[frame];
	*n_quads = f.dev.cyl ? f.n_quads : 0u;
	if (!f.dev.cyl || !h_fp16_bits) return INSTA_OK;
	INSTA_CUDA(cudaSetDevice(ctx->device));
	INSTA_CUDA(cudaDeviceSynchronize());
	INSTA_CUDA(cudaMemcpy(h_fp16_bits, f.dev.cyl, (size_t)f.n_quads * kCylStride, cudaMemcpyDeviceToHost));
	return INSTA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
int insta_deform_samples(insta_ctx* ctx, float* d_coords, uint32_t n, uint32_t stride_floats, int32_t frame_id, int32_t* d_tri_ids,
                         insta_stream stream) {
	return insta_deform_samples_n(ctx, d_coords, n, nullptr, stride_floats, frame_id, d_tri_ids, stream);
}

int insta_deform_samples_n(insta_ctx* ctx, float* d_coords, uint32_t n, const uint32_t* d_n_valid, uint32_t stride_floats, int32_t frame_id, int32_t* d_tri_ids,
                           insta_stream stream) {
	if (!ctx || (!d_coords && n)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_deform_samples: bad arguments");
	const uint32_t extra = ctx->p.exp_params + ctx->p.geo_params;
	if (stride_floats < 11 + extra) INSTA_FAIL(INSTA_ERR_INVALID, "insta_deform_samples: stride smaller than NerfCoordinate + extra dims");
	if (ctx->frames.empty()) INSTA_FAIL(INSTA_ERR_INVALID, "insta_deform_samples: no frames loaded");
	if (frame_id < -1) INSTA_FAIL(INSTA_ERR_INVALID, "insta_deform_samples: frame_id must be >= -1");
	int rc = sync_frame_table(ctx, (cudaStream_t)stream);
	if (rc) return rc;
	return launch_deform(ctx, d_coords, n, stride_floats, frame_id, nullptr, false, d_tri_ids, (cudaStream_t)stream, d_n_valid);
}

int insta_deform_points(insta_ctx* ctx, float* d_xyz, uint32_t n, const int32_t* d_frame_ids, int32_t* d_tri_ids, insta_stream stream) {
	if (!ctx || (!d_xyz && n)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_deform_points: bad arguments");
	if (ctx->frames.empty()) INSTA_FAIL(INSTA_ERR_INVALID, "insta_deform_points: no frames loaded");
	if (!d_frame_ids && ctx->frames.size() < 2) INSTA_FAIL(INSTA_ERR_INVALID, "insta_deform_points: the reference's frame draw needs >= 2 frames");
	int rc = sync_frame_table(ctx, (cudaStream_t)stream);
	if (rc) return rc;
	return launch_deform(ctx, d_xyz, n, 3, -1, d_frame_ids, true, d_tri_ids, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------
int insta_grid_layout(const insta_ctx* ctx, uint32_t* h_offsets, float* h_scales, uint32_t* h_resolutions) {
	if (!ctx) INSTA_FAIL(INSTA_ERR_INVALID, "ctx is null");
	const GridLevels& L = ctx->levels;
	if (h_offsets) memcpy(h_offsets, L.offset, sizeof(uint32_t) * (L.n_levels + 1));
	if (h_scales) memcpy(h_scales, L.scale, sizeof(float) * L.n_levels);
	if (h_resolutions) memcpy(h_resolutions, L.res, sizeof(uint32_t) * L.n_levels);
	return INSTA_OK;
}

uint32_t insta_grid_n_params(const insta_ctx* ctx) { return ctx ? ctx->levels.offset[ctx->levels.n_levels] * 8u : 0u; }

int insta_hash_encode_fwd(insta_ctx* ctx, const float* d_pos, uint32_t pos_stride_floats, uint32_t n, const void* d_table_f16, void* d_out_soa_f16,
                          insta_stream stream) {
	if (!ctx || ((!d_pos || !d_table_f16 || !d_out_soa_f16) && n) || pos_stride_floats < 3) INSTA_FAIL(INSTA_ERR_INVALID, "insta_hash_encode_fwd: bad arguments");
	if (reinterpret_cast<uintptr_t>(d_table_f16) % 16u) INSTA_FAIL(INSTA_ERR_INVALID, "insta_hash_encode_fwd: table must be 16-byte aligned");
	return launch_hash_fwd(ctx, d_pos, pos_stride_floats, n, d_table_f16, d_out_soa_f16, (cudaStream_t)stream);
}

int insta_hash_encode_bwd(insta_ctx* ctx, const float* d_pos, uint32_t pos_stride_floats, uint32_t n, const void* d_dLdy_soa_f16,
                          void* d_grad_table_f16, int zero_first, insta_stream stream) {
	if (!ctx || !d_grad_table_f16 || ((!d_pos || !d_dLdy_soa_f16) && n) || pos_stride_floats < 3) INSTA_FAIL(INSTA_ERR_INVALID, "insta_hash_encode_bwd: bad arguments");
	if (reinterpret_cast<uintptr_t>(d_grad_table_f16) % 16u) INSTA_FAIL(INSTA_ERR_INVALID, "insta_hash_encode_bwd: gradient table must be 16-byte aligned");
	return launch_hash_bwd(ctx, d_pos, pos_stride_floats, n, d_dLdy_soa_f16, d_grad_table_f16, zero_first, (cudaStream_t)stream);
}

int insta_hash_slots(insta_ctx* ctx, const float* d_pos, uint32_t pos_stride_floats, uint32_t n, uint32_t* d_slots, insta_stream stream) {
	if (!ctx || ((!d_pos || !d_slots) && n) || pos_stride_floats < 3) INSTA_FAIL(INSTA_ERR_INVALID, "insta_hash_slots: bad arguments");
	return launch_hash_slots(ctx, d_pos, pos_stride_floats, n, d_slots, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------
int insta_network_forward_backward(insta_ctx* ctx, const float* d_records, uint32_t stride_floats, uint32_t n, const void* d_params_f16, const void* d_dloss_f16,
                                   uint32_t dloss_width, void* d_gradients_f16, void* d_out_f16, insta_stream stream) {
	if (!ctx || !d_gradients_f16 || ((!d_records || !d_params_f16 || !d_dloss_f16) && n)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_forward_backward: bad arguments");
	if (stride_floats < 11 + ctx->p.exp_params + ctx->p.geo_params) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_forward_backward: stride smaller than NerfCoordinate + extra dims");
	if (reinterpret_cast<uintptr_t>(d_params_f16) % 16u || reinterpret_cast<uintptr_t>(d_gradients_f16) % 16u) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_forward_backward: parameters / gradients must be 16-byte aligned");
	if (dloss_width < 4 || dloss_width % 4) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_forward_backward: dloss_width must be a multiple of 4");
	if (ctx->p.n_levels != 16 || ctx->p.n_features_per_level != 8) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_forward_backward: the density MLP expects 16 levels x 8 features");
	return launch_network_fwd_bwd(ctx, d_records, stride_floats, n, d_params_f16, d_dloss_f16, dloss_width, d_gradients_f16, d_out_f16, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------
int insta_training_set_images(insta_ctx* ctx, const insta_train_image* h_images, uint32_t n_images) {
	if (!ctx || (!h_images && n_images)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_training_set_images: bad arguments");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	return training_set_images(ctx, h_images, n_images);
}

static int check_step(const insta_ctx* ctx, const insta_train_step* t, const char* who) {
	if (!ctx || !t) INSTA_FAIL(INSTA_ERR_INVALID, std::string(who) + ": bad arguments");
	if (ctx->n_train_images == 0) INSTA_FAIL(INSTA_ERR_INVALID, std::string(who) + ": no training images (insta_training_set_images)");
	if (!t->d_coords || !t->d_counters || !t->d_ray_indices || !t->d_rays || !t->d_numsteps || !t->d_density_grid_bitfield)
		INSTA_FAIL(INSTA_ERR_INVALID, std::string(who) + ": a buffer of insta_train_step is null");
	if (t->stride_floats < 11 + ctx->p.exp_params + ctx->p.geo_params) INSTA_FAIL(INSTA_ERR_INVALID, std::string(who) + ": stride smaller than NerfCoordinate + extra dims");
	return INSTA_OK;
}

int insta_generate_training_samples(insta_ctx* ctx, const insta_train_step* step, insta_stream stream) {
	int rc = check_step(ctx, step, "insta_generate_training_samples");
	if (rc) return rc;
	return train_generate(ctx, step, (cudaStream_t)stream);
}

int insta_compute_loss(insta_ctx* ctx, const insta_train_step* step, const void* d_net_out_f16, uint32_t out_width, float* d_coords_compacted, void* d_dloss_f16,
                       float* d_loss, const float* d_mean_density, insta_stream stream) {
	int rc = check_step(ctx, step, "insta_compute_loss");
	if (rc) return rc;
	if (!d_net_out_f16 || !d_coords_compacted || !d_dloss_f16 || !d_mean_density || out_width < 4 || out_width % 4) INSTA_FAIL(INSTA_ERR_INVALID, "insta_compute_loss: bad arguments");
	return train_loss(ctx, step, d_net_out_f16, out_width, d_coords_compacted, d_dloss_f16, d_loss, d_mean_density, (cudaStream_t)stream);
}

int insta_optimizer_step(insta_ctx* ctx, const insta_optimizer* o, insta_stream stream) {
	if (!ctx || !o || !o->d_params_fp32 || !o->d_params_f16 || !o->d_gradients_f16 || !o->d_first_moments || !o->d_second_moments || !o->d_param_steps || !o->d_params_ema_f16)
		INSTA_FAIL(INSTA_ERR_INVALID, "insta_optimizer_step: bad arguments");
	if (!(o->loss_scale > 0.f)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_optimizer_step: loss_scale must be positive");
	return optimizer_step(ctx, o, network_n_params(ctx), 25600u, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------
void insta_trainer_config_default(insta_trainer_config* c) {
	if (!c) return;
	std::memset(c, 0, sizeof(*c));
	c->struct_size = sizeof(*c);
	c->target_batch_size = 1u << 18;
	c->rays_per_batch = 1u << 12;
	c->snap_to_pixel_centers = 1;
	c->random_bg_color = 1;
	c->penalize_density_alpha = 0;
	c->rng_seed = 1337;
	c->cone_angle_constant = 0.f;
	c->loss_scale = 128.f;
	c->depth_weight = 0.f;
	c->seg_weight = 0.f;
	c->near_distance = 0.2f;
	c->beta_loss = 0.f;
}

int insta_trainer_create(insta_ctx* ctx, const insta_trainer_config* cfg) {
	if (!ctx || !cfg || cfg->struct_size != sizeof(insta_trainer_config)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_trainer_create: bad arguments / struct_size");
	if (cfg->target_batch_size == 0 || cfg->target_batch_size % 128 || cfg->target_batch_size > (1u << 20)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_trainer_create: target_batch_size must be a multiple of 128, at most 2^20");
	if (!(cfg->loss_scale > 0.f)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_trainer_create: loss_scale must be positive");
	if (ctx->p.n_levels != 16 || ctx->p.n_features_per_level != 8) INSTA_FAIL(INSTA_ERR_INVALID, "insta_trainer_create: the density MLP expects 16 levels x 8 features");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	return trainer_create(ctx, cfg);
}

int insta_trainer_destroy(insta_ctx* ctx) {
	if (!ctx) INSTA_FAIL(INSTA_ERR_INVALID, "insta_trainer_destroy: null context");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	INSTA_CUDA(cudaDeviceSynchronize());
	trainer_destroy(ctx);
	return INSTA_OK;
}

int insta_trainer_state(const insta_ctx* ctx, insta_train_stats* h_stats) {
	if (!ctx || !h_stats) INSTA_FAIL(INSTA_ERR_INVALID, "insta_trainer_state: bad arguments");
	return trainer_state(ctx, h_stats);
}

int insta_train_nerf(insta_ctx* ctx, const insta_optimizer* o, const void* d_bitfield, const float* d_extra_dims, const float* d_mean_density, int get_loss_scalar,
                     insta_train_stats* h_stats, insta_stream stream) {
	if (!ctx || !o || !o->d_params_fp32 || !o->d_params_f16 || !o->d_gradients_f16 || !o->d_first_moments || !o->d_second_moments || !o->d_param_steps || !o->d_params_ema_f16)
		INSTA_FAIL(INSTA_ERR_INVALID, "insta_train_nerf: bad arguments");
	if (!d_bitfield || !d_mean_density) INSTA_FAIL(INSTA_ERR_INVALID, "insta_train_nerf: density grid bitfield / mean are required");
	if (ctx->n_train_images == 0) INSTA_FAIL(INSTA_ERR_INVALID, "insta_train_nerf: no training images (insta_training_set_images)");
	if (ctx->frames.empty()) INSTA_FAIL(INSTA_ERR_INVALID, "insta_train_nerf: no frames loaded");
	if (reinterpret_cast<uintptr_t>(o->d_params_f16) % 16u || reinterpret_cast<uintptr_t>(o->d_gradients_f16) % 16u) INSTA_FAIL(INSTA_ERR_INVALID, "insta_train_nerf: parameters / gradients must be 16-byte aligned");
	return train_nerf(ctx, o, d_bitfield, d_extra_dims, d_mean_density, get_loss_scalar, h_stats, (cudaStream_t)stream);
}

static int check_grid(const insta_ctx* ctx, const insta_density_grid* g, const char* who) {
	if (!ctx || !g || !g->d_density_grid || !g->d_density_grid_bitfield || !g->d_density_grid_mean) INSTA_FAIL(INSTA_ERR_INVALID, std::string(who) + ": bad arguments");
	if (g->n_cascades < 1 || g->n_cascades > 8) INSTA_FAIL(INSTA_ERR_INVALID, std::string(who) + ": n_cascades must be 1..8");
	if (reinterpret_cast<uintptr_t>(g->d_density_grid) % 16u || reinterpret_cast<uintptr_t>(g->d_density_grid_bitfield) % 8u) INSTA_FAIL(INSTA_ERR_INVALID, std::string(who) + ": grid buffers must be 16-byte aligned");
	return INSTA_OK;
}

int insta_update_density_grid(insta_ctx* ctx, insta_density_grid* g, const void* d_params_f16, uint32_t n_uniform, uint32_t n_nonuniform, insta_stream stream) {
	int rc = check_grid(ctx, g, "insta_update_density_grid");
	if (rc) return rc;
	if (!d_params_f16 || reinterpret_cast<uintptr_t>(d_params_f16) % 16u) INSTA_FAIL(INSTA_ERR_INVALID, "insta_update_density_grid: parameters must be non-null and 16-byte aligned");
	if (ctx->p.n_levels != 16 || ctx->p.n_features_per_level != 8) INSTA_FAIL(INSTA_ERR_INVALID, "insta_update_density_grid: the density MLP expects 16 levels x 8 features");
	if ((uint64_t)n_uniform + n_nonuniform > (1ull << 27)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_update_density_grid: too many samples");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	return density_grid_update(ctx, g, d_params_f16, n_uniform, n_nonuniform, (cudaStream_t)stream);
}

int insta_training_prep_nerf(insta_ctx* ctx, insta_density_grid* g, const void* d_params_f16, int* did_update, insta_stream stream) {
	if (did_update) *did_update = 0;
	int rc = check_grid(ctx, g, "insta_training_prep_nerf");
	if (rc) return rc;
	const uint32_t n_prep_to_skip = std::min(std::max(g->training_step / 16u, 1u), 16u);
	if (g->training_step % n_prep_to_skip != 0) return INSTA_OK;
	const uint32_t cells = 128u * 128u * 128u * g->n_cascades;
	if (g->training_step < 256) rc = insta_update_density_grid(ctx, g, d_params_f16, cells, 0, stream);
	else rc = insta_update_density_grid(ctx, g, d_params_f16, cells / 4, cells / 4, stream);
	if (rc == INSTA_OK && did_update) *did_update = 1;
	return rc;
}

int insta_update_density_grid_mean_and_bitfield(insta_ctx* ctx, const insta_density_grid* g, insta_stream stream) {
	int rc = check_grid(ctx, g, "insta_update_density_grid_mean_and_bitfield");
	if (rc) return rc;
	return density_grid_mean_and_bitfield(ctx, g, (cudaStream_t)stream);
}

int insta_snapshot_save(const char* path, const insta_params* p, const insta_snapshot* s) {
	if (!path || !s || !s->h_params_f16 || s->n_params == 0 || (s->n_density_grid && !s->h_density_grid_f16)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_snapshot_save: bad arguments");
	if (s->has_optimizer && (!s->h_first_moments || !s->h_second_moments || !s->h_param_steps || !s->h_weights_ema_f16)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_snapshot_save: optimizer state incomplete");
	if (s->n_params * 4 > 0xffffffffull) INSTA_FAIL(INSTA_ERR_INVALID, "insta_snapshot_save: parameter count exceeds a MessagePack bin32");
	insta_params dflt;
	if (!p) {
		insta_params_default(&dflt);
		p = &dflt;
	}
	return snapshot_save(path, p, s);
}

int insta_snapshot_open(const char* path, insta_snapshot_file** file, insta_snapshot* s) {
	if (!path || !file || !s) INSTA_FAIL(INSTA_ERR_INVALID, "insta_snapshot_open: bad arguments");
	*file = nullptr;
	SnapshotFile* f = nullptr;
	const int rc = snapshot_open(path, &f, s);
	if (rc == INSTA_OK) *file = reinterpret_cast<insta_snapshot_file*>(f);
	return rc;
}

void insta_snapshot_close(insta_snapshot_file* file) { snapshot_close(reinterpret_cast<SnapshotFile*>(file)); }

int insta_microbench_gather(insta_ctx* ctx, uint64_t buffer_bytes, uint32_t chunk_bytes, uint32_t iters, float* h_gbs, insta_stream stream) {
	if (!ctx || !h_gbs || iters == 0) INSTA_FAIL(INSTA_ERR_INVALID, "insta_microbench_gather: bad arguments");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	return microbench_gather(ctx, buffer_bytes, chunk_bytes, iters, h_gbs, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------
// Mesh ray casting: the FLAME overlay (raycast.cu)
static int mesh_of(insta_ctx* ctx, int32_t frame_id, const Quad** quads, const Tri** tris, const char* who) {
	if (ctx->n_faces == 0) INSTA_FAIL(INSTA_ERR_INVALID, std::string(who) + ": no canonical mesh (insta_set_canonical)");
	if (frame_id >= 0) {
		if ((size_t)frame_id >= ctx->frames.size()) INSTA_FAIL(INSTA_ERR_INVALID, std::string(who) + ": frame_id out of range");
		*quads = ctx->frames[frame_id].dev.quads;
		*tris = ctx->frames[frame_id].dev.tris;
		return INSTA_OK;
	}
	std::lock_guard<std::mutex> lock(ctx->mu);
	if (!ctx->d_canon_quads) {
		// m_canonical_shape->triangle_bvh (TinyMesh::build_tries, tiny_mesh.cu:134-178): the same host build as a frame's
		std::vector<Tri> tris_h(ctx->canon_tris);
		HostBvh bvh;
		build_bvh(tris_h, ctx->p.n_primitives_per_leaf, bvh);
		Quad* dq = nullptr;
		Tri* dt = nullptr;
		INSTA_CUDA(cudaMalloc(&dq, bvh.quads.size() * sizeof(Quad)));
		cudaError_t e = cudaMalloc(&dt, bvh.tris.size() * sizeof(Tri));
		if (e == cudaSuccess) e = cudaMemcpy(dq, bvh.quads.data(), bvh.quads.size() * sizeof(Quad), cudaMemcpyHostToDevice);
		if (e == cudaSuccess) e = cudaMemcpy(dt, bvh.tris.data(), bvh.tris.size() * sizeof(Tri), cudaMemcpyHostToDevice);
		if (e != cudaSuccess) {
			cudaFree(dq);
			if (dt) cudaFree(dt);
			INSTA_CUDA(e);
		}
		ctx->d_canon_quads = dq;
		ctx->d_canon_tris_bvh = dt;
	}
	*quads = ctx->d_canon_quads;
	*tris = ctx->d_canon_tris_bvh;
	return INSTA_OK;
}

int insta_mesh_ray_intersect(insta_ctx* ctx, int32_t frame_id, const float* d_origins, const float* d_dirs, uint32_t n, int32_t* d_tri_ids, float* d_t, insta_stream stream) {
	if (!ctx || !d_origins || !d_dirs || !d_tri_ids) INSTA_FAIL(INSTA_ERR_INVALID, "insta_mesh_ray_intersect: bad arguments");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	const Quad* quads = nullptr;
	const Tri* tris = nullptr;
	if (int rc = mesh_of(ctx, frame_id, &quads, &tris, "insta_mesh_ray_intersect")) return rc;
	return launch_mesh_ray_intersect(ctx, quads, tris, d_origins, d_dirs, n, d_tri_ids, d_t, (cudaStream_t)stream);
}

int insta_raycast_mesh(insta_ctx* ctx, const insta_camera* cam, const insta_render_opts* opts, int32_t frame_id, int color_normal, float* d_frame_rgba, float* d_frame_depth,
                       insta_stream stream) {
	if (!ctx || !cam || !d_frame_rgba || !d_frame_depth) INSTA_FAIL(INSTA_ERR_INVALID, "insta_raycast_mesh: bad arguments");
	if (cam->resolution[0] == 0 || cam->resolution[1] == 0) INSTA_FAIL(INSTA_ERR_INVALID, "insta_raycast_mesh: empty resolution");
	if (!cam->snap_to_pixel_centers) INSTA_FAIL(INSTA_ERR_INVALID, "insta_raycast_mesh: the overlay is drawn through pixel centres (snap_to_pixel_centers = 1)");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	const Quad* quads = nullptr;
	const Tri* tris = nullptr;
	if (int rc = mesh_of(ctx, frame_id, &quads, &tris, "insta_raycast_mesh")) return rc;
	const float* aabb = opts && opts->use_render_aabb ? opts->render_aabb : ctx->render_aabb;
	return launch_mesh_overlay(ctx, cam, aabb, ctx->canon_centroid, color_normal, quads, tris, d_frame_rgba, d_frame_depth, (cudaStream_t)stream);
}

int insta_microbench_scatter(insta_ctx* ctx, uint64_t buffer_bytes, uint32_t chunk_bytes, uint32_t iters, float* h_gbs, insta_stream stream) {
	if (!ctx || !h_gbs || iters == 0) INSTA_FAIL(INSTA_ERR_INVALID, "insta_microbench_scatter: bad arguments");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	return microbench_scatter(ctx, buffer_bytes, chunk_bytes, iters, h_gbs, (cudaStream_t)stream);
}

int insta_comm_unique_id(void* h_id128) {
	if (!h_id128) INSTA_FAIL(INSTA_ERR_INVALID, "insta_comm_unique_id: null buffer");
	return comm_unique_id(h_id128);
}

int insta_comm_init_rank(insta_ctx* ctx, int world_size, int rank, const void* h_id128) {
	if (!ctx || world_size < 1 || rank < 0 || rank >= world_size || (world_size > 1 && !h_id128)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_comm_init_rank: bad arguments");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	return comm_init(ctx, world_size, rank, h_id128);
}

int insta_comm_destroy(insta_ctx* ctx) {
	if (!ctx) INSTA_FAIL(INSTA_ERR_INVALID, "insta_comm_destroy: null context");
	comm_destroy(ctx);
	return INSTA_OK;
}

int insta_comm_exchange_mode(const insta_ctx* ctx) { return ctx ? comm_exchange_mode(ctx) : 0; }

int insta_grad_allreduce(insta_ctx* ctx, void* d_gradients_f16, uint64_t n_halves, insta_stream stream) {
	if (!ctx || (!d_gradients_f16 && n_halves)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_grad_allreduce: bad arguments");
	return grad_allreduce(ctx, d_gradients_f16, n_halves, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------
uint32_t insta_network_n_params(const insta_ctx* ctx) { return ctx ? network_n_params(ctx) : 0; }

int insta_network_inference(insta_ctx* ctx, const float* d_records, uint32_t stride_floats, uint32_t n, const void* d_params_f16, void* d_out_f16,
                            uint32_t out_stride_elems, insta_stream stream) {
	return insta_network_inference_n(ctx, d_records, stride_floats, n, nullptr, d_params_f16, d_out_f16, out_stride_elems, stream);
}

int insta_network_inference_n(insta_ctx* ctx, const float* d_records, uint32_t stride_floats, uint32_t n, const uint32_t* d_n_valid, const void* d_params_f16,
                              void* d_out_f16, uint32_t out_stride_elems, insta_stream stream) {
	if (!ctx || ((!d_records || !d_params_f16 || !d_out_f16) && n)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_inference: bad arguments");
	if (stride_floats < 11 + ctx->p.exp_params + ctx->p.geo_params) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_inference: stride smaller than NerfCoordinate + extra dims");
	if (reinterpret_cast<uintptr_t>(d_params_f16) % 16u) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_inference: parameters must be 16-byte aligned");
	if (ctx->p.n_levels != 16 || ctx->p.n_features_per_level != 8) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_inference: the density MLP expects 16 levels x 8 features");
	if (out_stride_elems != 0 && out_stride_elems < n) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_inference: out_stride_elems < n");
	return launch_network_fwd(ctx, d_records, stride_floats, n, d_params_f16, d_out_f16, out_stride_elems, (cudaStream_t)stream, d_n_valid);
}

int insta_network_density(insta_ctx* ctx, const float* d_xyz, uint32_t stride_floats, uint32_t n, const void* d_params_f16, void* d_logits_f16, insta_stream stream) {
	if (!ctx || ((!d_xyz || !d_params_f16 || !d_logits_f16) && n)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_density: bad arguments");
	if (stride_floats < 3) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_density: stride smaller than a position");
	if (reinterpret_cast<uintptr_t>(d_params_f16) % 16u) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_density: parameters must be 16-byte aligned");
	if (ctx->p.n_levels != 16 || ctx->p.n_features_per_level != 8) INSTA_FAIL(INSTA_ERR_INVALID, "insta_network_density: the density MLP expects 16 levels x 8 features");
	return launch_network_density(ctx, d_xyz, stride_floats, n, d_params_f16, d_logits_f16, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------
int insta_tracer_init_rays(insta_ctx* ctx, const insta_camera* cam, const insta_render_opts* opts, float* d_frame_rgba, float* d_frame_depth,
                           float* d_dbg_origin_t, float* d_dbg_dir_alive, uint32_t* h_n_alive, insta_stream stream) {
	if (!ctx || !cam || !d_frame_depth) INSTA_FAIL(INSTA_ERR_INVALID, "insta_tracer_init_rays: bad arguments");
	if ((d_dbg_origin_t == nullptr) != (d_dbg_dir_alive == nullptr)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_tracer_init_rays: give both debug arrays or none");
	if (ctx->n_faces == 0 && !(opts && opts->use_render_aabb)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_tracer_init_rays: no canonical mesh and no render_aabb");
	ctx->render_opts = opts ? *opts : insta_render_opts{nullptr, 0.f, 0.01f, 0, {0, 0, 0, 1, 1, 1}, 0, 0};
	uint32_t n_alive = 0;  // the step-by-step entry points keep the live-ray count on the host, as NerfTracer does
	return tracer_init_rays(ctx, cam, &ctx->render_opts, d_frame_rgba, d_frame_depth, d_dbg_origin_t, d_dbg_dir_alive, h_n_alive ? h_n_alive : &n_alive, (cudaStream_t)stream);
}

int insta_render_frame(insta_ctx* ctx, const insta_camera* cam, const insta_render_opts* opts, int32_t frame_id, const void* d_params_f16, const float* d_extra_dims,
                       float* d_frame_rgba, float* d_frame_depth, insta_render_stats* h_stats, insta_stream stream) {
	if (!ctx || !cam || !d_params_f16 || !d_frame_rgba || !d_frame_depth) INSTA_FAIL(INSTA_ERR_INVALID, "insta_render_frame: bad arguments");
	if (ctx->n_faces == 0 && !(opts && opts->use_render_aabb)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_render_frame: no canonical mesh and no render_aabb");
	if (frame_id >= (int32_t)ctx->frames.size()) INSTA_FAIL(INSTA_ERR_INVALID, "insta_render_frame: frame_id out of range");
	if (reinterpret_cast<uintptr_t>(d_params_f16) % 16u) INSTA_FAIL(INSTA_ERR_INVALID, "insta_render_frame: parameters must be 16-byte aligned");
	ctx->render_opts = opts ? *opts : insta_render_opts{nullptr, 0.f, 0.01f, 0, {0, 0, 0, 1, 1, 1}, 0, 0};
	return render_frame(ctx, cam, &ctx->render_opts, frame_id, d_params_f16, d_extra_dims, d_frame_rgba, d_frame_depth, h_stats, (cudaStream_t)stream);
}

uint32_t insta_tracer_n_steps(const insta_ctx* ctx) { return ctx ? tracer_n_steps(ctx) : 0; }

int insta_tracer_generate_samples(insta_ctx* ctx, float* d_records, uint32_t stride_floats, uint32_t n_steps, int32_t frame_id, const float* d_extra_dims,
                                  uint32_t* n_records, insta_stream stream) {
	if (!ctx || !ctx->tracer || !d_records || n_steps == 0 || n_steps > 0xffffu) INSTA_FAIL(INSTA_ERR_INVALID, "insta_tracer_generate_samples: bad arguments");
	return tracer_generate(ctx, d_records, stride_floats, n_steps, frame_id, d_extra_dims, n_records, (cudaStream_t)stream);
}

int insta_tracer_composite(insta_ctx* ctx, const float* d_records, uint32_t stride_floats, const void* d_net_out_f16, uint32_t out_stride_elems, uint32_t n_steps,
                           float* d_frame_rgba, float* d_frame_depth, uint32_t* h_n_alive, insta_stream stream) {
	if (!ctx || !ctx->tracer || !d_records || !d_net_out_f16 || !d_frame_rgba || !d_frame_depth) INSTA_FAIL(INSTA_ERR_INVALID, "insta_tracer_composite: bad arguments");
	return tracer_composite(ctx, d_records, stride_floats, d_net_out_f16, out_stride_elems, n_steps, d_frame_rgba, d_frame_depth, (int)ctx->render_opts.linear_colors,
	                        h_n_alive, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------
// Recorder read-back
int insta_recorder_readback(insta_ctx* ctx, const float* d_frame_rgba, const float* d_frame_depth, uint32_t width, uint32_t height, uint32_t srgb_to_linear,
                            uint8_t* h_rgba8, uint32_t* h_depth_mm, insta_stream stream) {
	if (!ctx || !d_frame_rgba || width == 0 || height == 0 || (!h_rgba8 && !h_depth_mm)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_recorder_readback: bad arguments");
	if (h_depth_mm && !d_frame_depth) INSTA_FAIL(INSTA_ERR_INVALID, "insta_recorder_readback: depth requested without a depth buffer");
	if ((uint64_t)width * height > 0x7fffffffu) INSTA_FAIL(INSTA_ERR_INVALID, "insta_recorder_readback: frame too large");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	return recorder_readback(ctx, d_frame_rgba, d_frame_depth, width, height, (int)srgb_to_linear, h_rgba8, h_depth_mm, (cudaStream_t)stream);
}

int insta_host_write_png(const char* path, uint32_t width, uint32_t height, const uint8_t* h_rgba8) {
	if (!path || !h_rgba8 || width == 0 || height == 0) INSTA_FAIL(INSTA_ERR_INVALID, "insta_host_write_png: bad arguments");
	return write_png_rgba8(path, width, height, h_rgba8);
}

int insta_host_write_depth_bin(const char* path, const uint32_t* h_depth_mm, uint64_t n_values) {
	if (!path || (!h_depth_mm && n_values)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_host_write_depth_bin: bad arguments");
	return write_depth_bin(path, h_depth_mm, (size_t)n_values);
}

uint32_t insta_shard_frame_count(uint32_t n_frames, uint32_t rank, uint32_t world) {
	if (world == 0 || rank >= world || rank >= n_frames) return 0;
	return (n_frames - rank + world - 1) / world;
}

int insta_shard_owner(uint32_t global_frame, uint32_t world, uint32_t* rank, uint32_t* local_index) {
	if (world == 0) INSTA_FAIL(INSTA_ERR_INVALID, "insta_shard_owner: world must be positive");
	if (rank) *rank = global_frame % world;
	if (local_index) *local_index = global_frame / world;
	return INSTA_OK;
}

int insta_shard_local_ids(const int32_t* h_global_ids, uint64_t n, uint32_t rank, uint32_t world, int32_t* h_local_ids) {
	if (world == 0 || rank >= world || (n && (!h_global_ids || !h_local_ids))) INSTA_FAIL(INSTA_ERR_INVALID, "insta_shard_local_ids: bad arguments");
	for (uint64_t i = 0; i < n; ++i) {
		const int32_t g = h_global_ids[i];
		h_local_ids[i] = (g >= 0 && (uint32_t)g % world == rank) ? (int32_t)((uint32_t)g / world) : -1;
	}
	return INSTA_OK;
}

int insta_host_write_trajectory_json(const char* path, uint32_t width, uint32_t height, float fl_x, float fl_y, float cx, float cy, const insta_trajectory_frame* h_frames,
                                     uint32_t n_frames) {
	if (!path || (!h_frames && n_frames)) INSTA_FAIL(INSTA_ERR_INVALID, "insta_host_write_trajectory_json: bad arguments");
	return write_trajectory_json(path, width, height, fl_x, fl_y, cx, cy, h_frames, n_frames);
}

// ---------------------------------------------------------------------------------------------------------------
// Host-buffer path: chunks of the batch ping-pong between two streams so the H2D of chunk k+1 and the D2H of chunk
// k-1 overlap the kernels of chunk k.  The SoA output is written per chunk with a 2-D copy (row pitch n).
int insta_deform_encode_host(insta_ctx* ctx, float* h_coords, uint32_t n, uint32_t stride_floats, int32_t frame_id, const void* d_table_f16,
                             void* h_out_soa_f16) {
	if (!ctx || !h_coords || !d_table_f16 || !h_out_soa_f16) INSTA_FAIL(INSTA_ERR_INVALID, "insta_deform_encode_host: bad arguments");
	if (ctx->frames.empty()) INSTA_FAIL(INSTA_ERR_INVALID, "insta_deform_encode_host: no frames loaded");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	static const uint32_t chunk_log2 = [] { const char* e = getenv("INSTA_HOST_CHUNK_LOG2"); const int v = e ? atoi(e) : 19; return (uint32_t)(v >= 12 && v <= 22 ? v : 19); }();
	const uint32_t chunk = 1u << chunk_log2;  // default 524,288 samples: 56.6 MB of records, 134 MB of encodings per chunk
	int rc = ensure_stage(ctx, std::min(chunk, std::max(n, 1u)), stride_floats);
	if (rc) return rc;
	rc = sync_frame_table(ctx, ctx->streams[0]);
	if (rc) return rc;
	const uint32_t rows = ctx->levels.n_levels * 8;
	uint32_t k = 0;
	for (uint32_t first = 0; first < n; first += chunk, ++k) {
		const uint32_t m = std::min(chunk, n - first);
		const int b = k & 1;
		cudaStream_t s = ctx->streams[b];
		float* d_rec = ctx->d_stage_rec[b];
		INSTA_CUDA(cudaMemcpyAsync(d_rec, h_coords + (size_t)first * stride_floats, (size_t)m * stride_floats * 4, cudaMemcpyHostToDevice, s));
		rc = launch_deform(ctx, d_rec, m, stride_floats, frame_id, nullptr, false, nullptr, s);
		if (rc) return rc;
		rc = launch_hash_fwd(ctx, d_rec, stride_floats, m, d_table_f16, ctx->d_stage_enc[b], s);
		if (rc) return rc;
		INSTA_CUDA(cudaMemcpyAsync(h_coords + (size_t)first * stride_floats, d_rec, (size_t)m * stride_floats * 4, cudaMemcpyDeviceToHost, s));
		INSTA_CUDA(cudaMemcpy2DAsync((__half*)h_out_soa_f16 + first, (size_t)n * 2, ctx->d_stage_enc[b], (size_t)m * 2, (size_t)m * 2, rows,
		                             cudaMemcpyDeviceToHost, s));
	}
	INSTA_CUDA(cudaStreamSynchronize(ctx->streams[0]));
	INSTA_CUDA(cudaStreamSynchronize(ctx->streams[1]));
	return INSTA_OK;
}

int insta_deform_infer_host(insta_ctx* ctx, const float* h_coords, uint32_t n, uint32_t stride_floats, int32_t frame_id, const void* d_params_f16,
                            void* h_out_rgbsigma_f16) {
	if (!ctx || !h_coords || !d_params_f16 || !h_out_rgbsigma_f16) INSTA_FAIL(INSTA_ERR_INVALID, "insta_deform_infer_host: bad arguments");
	if (ctx->frames.empty()) INSTA_FAIL(INSTA_ERR_INVALID, "insta_deform_infer_host: no frames loaded");
	if (stride_floats < 11 + ctx->p.exp_params + ctx->p.geo_params) INSTA_FAIL(INSTA_ERR_INVALID, "insta_deform_infer_host: records too short for the network's inputs");
	INSTA_CUDA(cudaSetDevice(ctx->device));
	const uint32_t chunk = 1u << 19;
	int rc = ensure_stage(ctx, std::min(chunk, std::max(n, 1u)), stride_floats);  // the encoding stage buffer (128 halves per sample) holds the 4 output halves
	if (rc) return rc;
	rc = sync_frame_table(ctx, ctx->streams[0]);
	if (rc) return rc;
	uint32_t k = 0;
	for (uint32_t first = 0; first < n; first += chunk, ++k) {
		const uint32_t m = std::min(chunk, n - first);
		const int b = k & 1;
		cudaStream_t s = ctx->streams[b];
		float* d_rec = ctx->d_stage_rec[b];
		INSTA_CUDA(cudaMemcpyAsync(d_rec, h_coords + (size_t)first * stride_floats, (size_t)m * stride_floats * 4, cudaMemcpyHostToDevice, s));
		rc = launch_deform(ctx, d_rec, m, stride_floats, frame_id, nullptr, false, nullptr, s);
		if (rc) return rc;
		rc = launch_network_fwd(ctx, d_rec, stride_floats, m, d_params_f16, ctx->d_stage_enc[b], 0, s);
		if (rc) return rc;
		INSTA_CUDA(cudaMemcpyAsync((__half*)h_out_rgbsigma_f16 + (size_t)first * 4, ctx->d_stage_enc[b], (size_t)m * 8, cudaMemcpyDeviceToHost, s));
	}
	INSTA_CUDA(cudaStreamSynchronize(ctx->streams[0]));
	INSTA_CUDA(cudaStreamSynchronize(ctx->streams[1]));
	return INSTA_OK;
}

int insta_host_alloc(void** p, uint64_t bytes) {
	if (!p) INSTA_FAIL(INSTA_ERR_INVALID, "p is null");
	INSTA_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
	return INSTA_OK;
}

int insta_host_free(void* p) {
	INSTA_CUDA(cudaFreeHost(p));
	return INSTA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// host-only utilities
int insta_host_bvh_build(void* h_tris, uint32_t n_tris, uint32_t leaf, void* h_nodes, uint32_t max_nodes, uint32_t* n_nodes) {
	if (!h_tris || n_tris == 0 || leaf == 0 || leaf > 7) INSTA_FAIL(INSTA_ERR_INVALID, "insta_host_bvh_build: bad arguments");
	std::vector<Tri> tris((Tri*)h_tris, (Tri*)h_tris + n_tris);
	HostBvh bvh;
	build_bvh(tris, leaf, bvh);
	if (n_nodes) *n_nodes = (uint32_t)bvh.nodes.size();
	if (bvh.nodes.size() > max_nodes) INSTA_FAIL(INSTA_ERR_CAPACITY, "insta_host_bvh_build: node buffer too small");
	if (h_nodes) memcpy(h_nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(Node));
	memcpy(h_tris, bvh.tris.data(), (size_t)n_tris * sizeof(Tri));
	return INSTA_OK;
}

int insta_host_build_triangles(const float* h_verts, const uint32_t* h_faces, uint32_t n_faces, const uint32_t* h_mask_ids,
                               const void* h_canonical_tris, void* h_tris_out) {
	if (!h_verts || !h_faces || !h_tris_out) INSTA_FAIL(INSTA_ERR_INVALID, "insta_host_build_triangles: bad arguments");
	build_triangles(h_verts, h_faces, n_faces, h_mask_ids, (const Tri*)h_canonical_tris, (Tri*)h_tris_out);
	return INSTA_OK;
}

int insta_host_adjacency(const uint32_t* h_faces, uint32_t n_faces, uint32_t n_verts, int32_t* h_out) {
	if (!h_faces || !h_out) INSTA_FAIL(INSTA_ERR_INVALID, "insta_host_adjacency: bad arguments");
	std::vector<int32_t> adj;
	build_adjacency(h_faces, n_faces, n_verts, adj);
	memcpy(h_out, adj.data(), adj.size() * 4);
	return INSTA_OK;
}

int insta_host_load_canonical_obj(const char* canonical_obj, const char* mesh_masks_dir, float* h_verts, uint32_t* h_faces, uint32_t* h_mask_ids,
                                  uint32_t* n_verts, uint32_t* n_faces) {
	if (!canonical_obj) INSTA_FAIL(INSTA_ERR_INVALID, "insta_host_load_canonical_obj: bad arguments");
	std::vector<float> verts;
	std::vector<uint32_t> faces, masks;
	int rc = load_canonical_obj_host(canonical_obj, mesh_masks_dir, verts, faces, masks);
	if (rc) return rc;
	if (n_verts) *n_verts = (uint32_t)(verts.size() / 3);
	if (n_faces) *n_faces = (uint32_t)(faces.size() / 3);
	if (h_verts) memcpy(h_verts, verts.data(), verts.size() * 4);
	if (h_faces) memcpy(h_faces, faces.data(), faces.size() * 4);
	if (h_mask_ids) {
		if (masks.empty()) masks.assign(faces.size() / 3, 15u);
		memcpy(h_mask_ids, masks.data(), masks.size() * 4);
	}
	return INSTA_OK;
}

int insta_host_mesh_aabb(const float* h_verts, uint32_t n_verts, float* h_aabb6) {
	if (!h_verts || !h_aabb6 || n_verts == 0) INSTA_FAIL(INSTA_ERR_INVALID, "insta_host_mesh_aabb: bad arguments");
	mesh_aabb(h_verts, n_verts, h_aabb6);
	return INSTA_OK;
}

uint64_t insta_launch_count(const insta_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

}  // extern "C"
