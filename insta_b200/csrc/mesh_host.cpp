// Host-side mesh ingest for libinsta_b200: OBJ parsing, region masks, per-frame triangles, edge adjacency and the
// 4-ary BVH whose topology and leaf order must be identical to the reference's, because closest_triangle breaks
// distance ties by visit order (hard part 1).
//
// Behaviour follows (file:line in the INSTA repository):
//   TinyMesh::load / build_triangles / get_triangle_3_neighbours   src/tiny_mesh.cu:53-84, 134-162, 316-371
//   Masking::load_mesh_color_mask / triangle_to_mask               src/masking.cpp:118-180
//   TriangleBvhWithBranchingFactor<4>::build                       instant-ngp/src/triangle_bvh.cu:548-631
// All arithmetic that decides a comparison is plain fp32 without contraction (this file is compiled with
// -ffp-contract=off), which is what the reference's host build evaluates.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>

#include "ctx.h"

namespace insta {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
const char* last_error_cstr() { return g_last_error.c_str(); }

// ---------------------------------------------------------------------------------------------------------------
// Wavefront OBJ: "v x y z [r g b]" and "f i[/..] j[/..] k[/..]" (triangles or fans); 1-based, negative = relative.
bool load_obj(const char* path, std::vector<float>& verts, std::vector<float>& colors, std::vector<uint32_t>& faces, std::string& err) {
	std::ifstream in(path);
	if (!in) {
		err = std::string("cannot open ") + path;
		return false;
	}
	verts.clear();
	colors.clear();
	faces.clear();
	std::string line;
	std::vector<long> poly;
	while (std::getline(in, line)) {
		const char* s = line.c_str();
		while (*s == ' ' || *s == '\t') ++s;
		if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
			char* end;
			float vals[6];
			int k = 0;
			const char* q = s + 1;
			for (; k < 6; ++k) {
				float v = strtof(q, &end);
				if (end == q) break;
				vals[k] = v;
				q = end;
			}
			if (k < 3) {
				err = std::string("bad vertex line in ") + path;
				return false;
			}
			verts.insert(verts.end(), vals, vals + 3);
			if (k >= 6) colors.insert(colors.end(), vals + 3, vals + 6);
		} else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
			poly.clear();
			const char* q = s + 1;
			while (*q) {
				while (*q == ' ' || *q == '\t' || *q == '\r') ++q;
				if (!*q) break;
				char* end;
				long idx = strtol(q, &end, 10);
				if (end == q) break;
				long nv = (long)(verts.size() / 3);
				poly.push_back(idx > 0 ? idx - 1 : nv + idx);
				q = end;
				while (*q && *q != ' ' && *q != '\t' && *q != '\r') ++q;  // skip /vt/vn
			}
			for (size_t k = 2; k < poly.size(); ++k) {
				faces.push_back((uint32_t)poly[0]);
				faces.push_back((uint32_t)poly[k - 1]);
				faces.push_back((uint32_t)poly[k]);
			}
		}
	}
	if (verts.empty()) {
		err = std::string("no vertices in ") + path;
		return false;
	}
	const uint32_t nv = (uint32_t)(verts.size() / 3);
	for (uint32_t f : faces)
		if (f >= nv) {
			err = std::string("face index out of range in ") + path;
			return false;
		}
	return true;
}

// ---------------------------------------------------------------------------------------------------------------
static inline float area_of(const Tri& t) {
	const float ux = t.b[0] - t.a[0], uy = t.b[1] - t.a[1], uz = t.b[2] - t.a[2];
	const float vx = t.c[0] - t.a[0], vy = t.c[1] - t.a[1], vz = t.c[2] - t.a[2];
	const float cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
	return 0.5f * std::sqrt(cx * cx + (cy * cy + cz * cz));
}

void build_triangles(const float* verts, const uint32_t* faces, uint32_t n_faces, const uint32_t* mask_ids, const Tri* canonical, Tri* out) {
	for (uint32_t f = 0; f < n_faces; ++f) {
		Tri& t = out[f];
		memcpy(t.a, verts + 3 * (size_t)faces[3 * f + 0], 12);
		memcpy(t.b, verts + 3 * (size_t)faces[3 * f + 1], 12);
		memcpy(t.c, verts + 3 * (size_t)faces[3 * f + 2], 12);
		t.id = f;
		t.mask_id = mask_ids ? mask_ids[f] : 15u;
		t.scaling = canonical ? area_of(t) / area_of(canonical[f]) : 1.0f;
	}
}

void mesh_aabb(const float* verts, uint32_t n_verts, float* aabb6) {
	float mn[3], mx[3];
	for (int k = 0; k < 3; ++k) {
		mn[k] = std::numeric_limits<float>::infinity();
		mx[k] = -mn[k];
	}
	for (uint32_t i = 0; i < n_verts; ++i)
		for (int k = 0; k < 3; ++k) {
			mn[k] = std::min(mn[k], verts[3 * (size_t)i + k]);
			mx[k] = std::max(mx[k], verts[3 * (size_t)i + k]);
		}
	const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
	const float inflate = std::sqrt(dx * dx + (dy * dy + dz * dz)) * 0.15f;
	for (int k = 0; k < 3; ++k) {
		aabb6[k] = mn[k] - inflate;
		aabb6[3 + k] = mx[k] + inflate;
	}
}

// Up to three edge-sharing faces per face, ascending, -1 padded.  Built from a sorted (edge -> faces) list rather
// than per-vertex std::set stars; the result is the same set the reference computes for a manifold mesh.
void build_adjacency(const uint32_t* faces, uint32_t n_faces, uint32_t n_verts, std::vector<int32_t>& out3) {
	(void)n_verts;
	struct EdgeFace {
		uint64_t key;
		uint32_t face;
	};
	std::vector<EdgeFace> ef;
	ef.reserve((size_t)n_faces * 3);
	for (uint32_t f = 0; f < n_faces; ++f)
		for (int e = 0; e < 3; ++e) {
			uint32_t u = faces[3 * f + e], v = faces[3 * f + (e + 1) % 3];
			if (u == v) continue;
			if (u > v) std::swap(u, v);
			ef.push_back({((uint64_t)u << 32) | v, f});
		}
	std::sort(ef.begin(), ef.end(), [](const EdgeFace& a, const EdgeFace& b) { return a.key < b.key || (a.key == b.key && a.face < b.face); });
	std::vector<std::vector<uint32_t>> nb(n_faces);
	for (size_t i = 0; i < ef.size();) {
		size_t j = i;
		while (j < ef.size() && ef[j].key == ef[i].key) ++j;
		for (size_t a = i; a < j; ++a)
			for (size_t b = i; b < j; ++b)
				if (ef[a].face != ef[b].face) nb[ef[a].face].push_back(ef[b].face);
		i = j;
	}
	out3.assign((size_t)n_faces * 3, -1);
	for (uint32_t f = 0; f < n_faces; ++f) {
		auto& v = nb[f];
		std::sort(v.begin(), v.end());
		v.erase(std::unique(v.begin(), v.end()), v.end());
		for (size_t k = 0; k < v.size() && k < 3; ++k) out3[3 * (size_t)f + k] = (int32_t)v[k];
	}
}

void masks_from_regions(const uint32_t* faces, uint32_t n_faces, const std::vector<std::pair<int, std::vector<uint32_t>>>& regions,
                        std::vector<uint32_t>& out) {
	uint32_t nv = 0;
	for (uint32_t i = 0; i < n_faces * 3; ++i) nv = std::max(nv, faces[i] + 1);
	out.assign(n_faces, 15u);  // FlameMasks::Torso
	std::vector<uint8_t> in_region(nv);
	auto sorted = regions;
	std::sort(sorted.begin(), sorted.end(), [](const auto& a, const auto& b) { return a.first < b.first; });  // std::map order
	for (const auto& r : sorted) {
		std::fill(in_region.begin(), in_region.end(), 0);
		for (uint32_t v : r.second)
			if (v < nv) in_region[v] = 1;
		for (uint32_t f = 0; f < n_faces; ++f)
			if (in_region[faces[3 * f]] | in_region[faces[3 * f + 1]] | in_region[faces[3 * f + 2]]) out[f] = (uint32_t)r.first;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// BVH.  The permutation is produced by std::nth_element over 16-byte centroid keys; libstdc++'s introselect only
// looks at comparison results, so the order equals the one the reference obtains by selecting on whole triangles.
namespace {
struct Key {
	float c[3];
	uint32_t tri;
};
struct Range {
	uint32_t begin, end;
};

inline void bounds(const std::vector<Tri>& tris, const Key* keys, Range r, uint32_t n_total, float* mn, float* mx) {
	// BoundingBox(begin, end): min = max = begin->a, then enlarge by every vertex (ngp bounding_box.cuh:54-59);
	// for an empty range `begin` still designates an element of the array.
	const uint32_t first = r.begin < n_total ? r.begin : n_total - 1;
	const Tri& t0 = tris[keys[first].tri];
	for (int k = 0; k < 3; ++k) mn[k] = mx[k] = t0.a[k];
	for (uint32_t i = r.begin; i < r.end; ++i) {
		const Tri& t = tris[keys[i].tri];
		for (int k = 0; k < 3; ++k) {
			mn[k] = std::min(std::min(std::min(mn[k], t.a[k]), t.b[k]), t.c[k]);
			mx[k] = std::max(std::max(std::max(mx[k], t.a[k]), t.b[k]), t.c[k]);
		}
	}
}

inline uint32_t split_axis(const Key* keys, Range r) {
	const float cnt = (float)(r.end - r.begin);
	float mean[3] = {0, 0, 0};
	for (uint32_t i = r.begin; i < r.end; ++i)
		for (int k = 0; k < 3; ++k) mean[k] += keys[i].c[k];
	for (int k = 0; k < 3; ++k) mean[k] /= cnt;
	float var[3] = {0, 0, 0};
	for (uint32_t i = r.begin; i < r.end; ++i)
		for (int k = 0; k < 3; ++k) {
			const float d = keys[i].c[k] - mean[k];
			var[k] += d * d;
		}
	for (int k = 0; k < 3; ++k) var[k] /= cnt;
	uint32_t axis = 0;
	if (var[1] > var[axis]) axis = 1;
	if (var[2] > var[axis]) axis = 2;
	return axis;
}
}  // namespace

void build_bvh(std::vector<Tri>& tris_in, uint32_t leaf_size, HostBvh& out) {
	const uint32_t n = (uint32_t)tris_in.size();
	std::vector<Key> keys(n);
	for (uint32_t i = 0; i < n; ++i) {
		const Tri& t = tris_in[i];
		for (int k = 0; k < 3; ++k) keys[i].c[k] = (t.a[k] + t.b[k] + t.c[k]) / 3.0f;
		keys[i].tri = i;
	}
	auto& nodes = out.nodes;
	nodes.clear();
	nodes.reserve(n + n / 2 + 8);
	nodes.emplace_back();
	bounds(tris_in, keys.data(), {0, n}, n, nodes[0].mn, nodes[0].mx);
	std::vector<Range> node_range;
	node_range.reserve(n + n / 2 + 8);
	node_range.push_back({0, n});

	struct Work {
		uint32_t node;
		Range r;
	};
	std::vector<Work> todo;
	todo.push_back({0, {0, n}});
	while (!todo.empty()) {
		const Work w = todo.back();
		todo.pop_back();
		// two rounds of median split: whole range, then second half, then first half
		Range ch[4];
		ch[0] = w.r;
		for (int n_children = 1; n_children < 4; n_children *= 2)
			for (int i = n_children - 1; i >= 0; --i) {
				const Range r = ch[i];
				const uint32_t axis = split_axis(keys.data(), r);
				const uint32_t mid = r.begin + (r.end - r.begin) / 2;
				std::nth_element(keys.begin() + r.begin, keys.begin() + mid, keys.begin() + r.end,
				                 [axis](const Key& a, const Key& b) { return a.c[axis] < b.c[axis]; });
				ch[2 * i] = {r.begin, mid};
				ch[2 * i + 1] = {mid, r.end};
			}
		nodes[w.node].left_idx = (int32_t)nodes.size();
		for (int i = 0; i < 4; ++i) {
			const uint32_t idx = (uint32_t)nodes.size();
			nodes.emplace_back();
			node_range.push_back(ch[i]);
			bounds(tris_in, keys.data(), ch[i], n, nodes[idx].mn, nodes[idx].mx);
			if (ch[i].end - ch[i].begin <= leaf_size) {
				nodes[idx].left_idx = -(int32_t)ch[i].begin - 1;
				nodes[idx].right_idx = -(int32_t)ch[i].end - 1;
			} else {
				todo.push_back({idx, ch[i]});
			}
		}
		nodes[w.node].right_idx = (int32_t)nodes.size();
	}

	out.tris.resize(n);
	for (uint32_t i = 0; i < n; ++i) out.tris[i] = tris_in[keys[i].tri];

	// device layout: one Quad per inner node, numbered in node order (quad 0 = children of the root)
	std::vector<int32_t> quad_of(nodes.size(), -1);
	uint32_t n_quads = 0;
	for (size_t i = 0; i < nodes.size(); ++i)
		if (nodes[i].left_idx >= 0) quad_of[i] = (int32_t)n_quads++;
	out.quads.assign(n_quads, Quad{});
	out.child_ranges.assign((size_t)n_quads * 8, 0u);
	for (size_t i = 0; i < nodes.size(); ++i) {
		if (nodes[i].left_idx < 0) continue;
		Quad& q = out.quads[quad_of[i]];
		for (int k = 0; k < 4; ++k) {
			const Node& c = nodes[nodes[i].left_idx + k];
			out.child_ranges[(size_t)quad_of[i] * 8 + 2 * k] = node_range[nodes[i].left_idx + k].begin;
			out.child_ranges[(size_t)quad_of[i] * 8 + 2 * k + 1] = node_range[nodes[i].left_idx + k].end;
			q.mnx[k] = c.mn[0]; q.mny[k] = c.mn[1]; q.mnz[k] = c.mn[2];
			q.mxx[k] = c.mx[0]; q.mxy[k] = c.mx[1]; q.mxz[k] = c.mx[2];
			if (c.left_idx >= 0) {
				q.ref[k] = quad_of[nodes[i].left_idx + k];
				q.code[k] = q.ref[k] < 0x7fff ? q.ref[k] : -1;  // 16-bit stack code of the warp-queue kernel (deform.cu)
			} else {
				const uint32_t first = (uint32_t)(-c.left_idx - 1), count = (uint32_t)(-c.right_idx - 1) - first;
				// An empty leaf (count 0) keeps its degenerate box: it takes part in the child sort exactly as in the
				// reference, because the order among tied box distances decides the order in which tied triangles are met.
				q.ref[k] = -(int32_t)((first << 3) | (count & 7u)) - 1;
				q.code[k] = (first < 8192u && count <= 3u) ? (int32_t)(0x8000u | (first << 2) | count) : -1;
			}
		}
	}
}

}  // namespace insta
