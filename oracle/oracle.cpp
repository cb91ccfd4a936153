// TEST INFRASTRUCTURE ONLY -- see oracle.h.  CPU restatement of the reference algorithm; every function cites
// the reference file:line it follows.  Plain IEEE fp32, no contraction (-ffp-contract=off), no Eigen.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#ifdef __F16C__
#include <immintrin.h>
#endif

namespace {

struct V3 {
	float x, y, z;
	float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 v3(const float* p) { return {p[0], p[1], p[2]}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
// Eigen fixed-size-3 reduction order is x0 + (x1 + x2)  (Eigen/src/Core/Redux.h:101-115)
inline float dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
inline float sqn(V3 a) { return dot(a, a); }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float sgn(float x) { return copysignf(1.0f, x); }                                 // ngp common.h:191
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }  // tcnn common.h:243
// Eigen normalized(): z = squaredNorm; if (z > 0) v / sqrt(z) else v   (Eigen/src/Core/Dot.h)
inline V3 normalized(V3 a) {
	float z = sqn(a);
	if (z > 0.f) return a / std::sqrt(z);
	return a;
}
inline V3 centroid(const orc_tri& t) { return (v3(t.a) + v3(t.b) + v3(t.c)) / 3.0f; }            // triangle.cuh:220
inline float centroid_axis(const orc_tri& t, int ax) { return (t.a[ax] + t.b[ax] + t.c[ax]) / 3; }  // triangle.cuh:224

// Arithmetic flavour of the two distance functions that decide the closest triangle:
//   0 = the reference's HOST build (plain IEEE fp32, nothing contracted)               -- pinned here against libref_host.so
//   1 = the reference's DEVICE build (nvcc 12.9 -O3, default -fmad=true): the FMUL/FADD pairs nvcc fuses into FFMA,
//       read off the SASS of `warp` / `warp_random_samples` (identical in both kernels)  -- pinned on the GPU box against
//       the reference's own CUDA kernels (tests/test_gpu_vs_ref_cuda.py)
// fmaf() here is the exact single-rounding fused multiply-add.
int g_arith = 0;

// x0*y0 + (x1*y1 + x2*y2) as the device build evaluates it: pattern A fuses x and y and rounds the z product,
// pattern B fuses x and z and rounds the y product.
inline float dotA(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
inline float dotB(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.z, b.z, a.y * b.y)); }
// a x b, each component fma(first product, -(second product rounded))
inline V3 crossF(V3 a, V3 b) { return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))}; }

// triangle.cuh:88-107, device-build contraction
inline float tri_distance_sq_dev(const orc_tri& t, V3 pos) {
	V3 a = v3(t.a), b = v3(t.b), c = v3(t.c);
	V3 v21 = b - a, p1 = pos - a;
	V3 v32 = c - b, p2 = pos - b;
	V3 v13 = a - c, p3 = pos - c;
	V3 nor = crossF(v21, v13);
	if (sgn(dotA(crossF(v21, nor), p1)) + sgn(dotA(crossF(v32, nor), p2)) + sgn(dotA(crossF(v13, nor), p3)) < 2.0f) {
		auto edge = [](V3 v, V3 q) {
			float tt = clampf(dotB(v, q) / dotA(v, v), 0.0f, 1.0f);
			V3 e = {fmaf(v.x, tt, -q.x), fmaf(v.y, tt, -q.y), fmaf(v.z, tt, -q.z)};
			return dotA(e, e);
		};
		float m = edge(v21, p1);
		float e2 = edge(v32, p2), e3 = edge(v13, p3);
		if (e2 < m) m = e2;
		if (e3 < m) m = e3;
		return m;
	}
	float d = dotB(nor, p1);
	return d * d / dotB(nor, nor);
}

// triangle.cuh:88-107, host build
inline float tri_distance_sq_host(const orc_tri& t, V3 pos) {
	V3 a = v3(t.a), b = v3(t.b), c = v3(t.c);
	V3 v21 = b - a, p1 = pos - a;
	V3 v32 = c - b, p2 = pos - b;
	V3 v13 = a - c, p3 = pos - c;
	V3 nor = cross(v21, v13);
	if (sgn(dot(cross(v21, nor), p1)) + sgn(dot(cross(v32, nor), p2)) + sgn(dot(cross(v13, nor), p3)) < 2.0f) {
		float e1 = sqn(v21 * clampf(dot(v21, p1) / sqn(v21), 0.0f, 1.0f) - p1);
		float e2 = sqn(v32 * clampf(dot(v32, p2) / sqn(v32), 0.0f, 1.0f) - p2);
		float e3 = sqn(v13 * clampf(dot(v13, p3) / sqn(v13), 0.0f, 1.0f) - p3);
		float m = e1;             // std::min({..}) keeps the first smallest
		if (e2 < m) m = e2;
		if (e3 < m) m = e3;
		return m;
	}
	return dot(nor, p1) * dot(nor, p1) / sqn(nor);
}
inline float tri_distance_sq(const orc_tri& t, V3 pos) { return g_arith ? tri_distance_sq_dev(t, pos) : tri_distance_sq_host(t, pos); }

// bounding_box.cuh:228-230
inline float box_distance_sq(const float* mn, const float* mx, V3 p) {
	float d[3];
	for (int k = 0; k < 3; ++k) {
		float lo = mn[k] - p[k], hi = p[k] - mx[k];
		float m = std::max(lo, hi);
		d[k] = std::max(m, 0.0f);
	}
	if (g_arith) return fmaf(d[0], d[0], fmaf(d[1], d[1], d[2] * d[2]));
	return d[0] * d[0] + (d[1] * d[1] + d[2] * d[2]);
}

inline void box_of(const orc_tri* begin, const orc_tri* end, float* mn, float* mx) {  // bounding_box.cuh:54-59
	for (int k = 0; k < 3; ++k) mn[k] = mx[k] = begin->a[k];
	for (const orc_tri* it = begin; it != end; ++it) {
		const float* vs[3] = {it->a, it->b, it->c};
		for (auto v : vs)
			for (int k = 0; k < 3; ++k) {
				mn[k] = std::min(mn[k], v[k]);
				mx[k] = std::max(mx[k], v[k]);
			}
	}
}

// closest_triangle, bvh.cuh:48-96.  Returns BVH-order index; *dist = sqrt(best).  init_best: the reference's max_distance_sq argument (bvh.cuh:48,
// FLT_MAX at every call site of the path); the product starts from a finite upper bound instead, tests/test_oracle_seed_bound.py checks that this
// cannot change the triangle.
inline int closest_tri(V3 p, const orc_node* nodes, const orc_tri* tris, float* dist, float init_best = std::numeric_limits<float>::max()) {
	int stack[64];
	int sp = 0;
	stack[sp++] = 0;
	float best = init_best;
	int best_i = -1;
	while (sp > 0) {
		int idx = stack[--sp];
		const orc_node& nd = nodes[idx];
		if (nd.left_idx < 0) {
			int end = -nd.right_idx - 1;
			for (int i = -nd.left_idx - 1; i < end; ++i) {
				float d = tri_distance_sq(tris[i], p);
				if (d <= best) { best = d; best_i = i; }
			}
		} else {
			float cd[4];
			uint32_t ci[4];
			for (uint32_t k = 0; k < 4; ++k) {
				cd[k] = box_distance_sq(nodes[nd.left_idx + k].mn, nodes[nd.left_idx + k].mx, p);
				ci[k] = nd.left_idx + k;
			}
			auto cas = [&](int i, int j) {  // swap iff values[i] < values[j]  (descending)
				if (cd[i] < cd[j]) { std::swap(cd[i], cd[j]); std::swap(ci[i], ci[j]); }
			};
			cas(0, 2); cas(1, 3); cas(0, 1); cas(2, 3); cas(1, 2);
			for (int k = 0; k < 4; ++k)
				if (cd[k] <= best && sp < 63) stack[sp++] = (int)ci[k];
		}
	}
	if (best_i == -1) best = 0.0f;
	if (dist) *dist = std::sqrt(best);
	return best_i;
}

// ---- the mesh ray cast of rta::Core::raycast_flame (src/core.cu:350-402) ------------------------------------------------------------------
// BoundingBox::ray_intersect(pos, dir).x()  (ngp bounding_box.cuh:163-205): entry parameter of the slab test, FLT_MAX on a miss
inline float box_ray_tmin(const float* mn, const float* mx, V3 o, V3 d) {
	float tmin = (mn[0] - o.x) / d.x, tmax = (mx[0] - o.x) / d.x;
	if (tmin > tmax) std::swap(tmin, tmax);
	float tymin = (mn[1] - o.y) / d.y, tymax = (mx[1] - o.y) / d.y;
	if (tymin > tymax) std::swap(tymin, tymax);
	if (tmin > tymax || tymin > tmax) return std::numeric_limits<float>::max();
	if (tymin > tmin) tmin = tymin;
	if (tymax < tmax) tmax = tymax;
	float tzmin = (mn[2] - o.z) / d.z, tzmax = (mx[2] - o.z) / d.z;
	if (tzmin > tzmax) std::swap(tzmin, tzmax);
	if (tmin > tzmax || tzmin > tmax) return std::numeric_limits<float>::max();
	if (tzmin > tmin) tmin = tzmin;
	return tmin;
}
// Triangle::ray_intersect  (ngp triangle.cuh:66-81)
inline float tri_ray_t(const orc_tri& t, V3 ro, V3 rd) {
	V3 v1v0 = v3(t.b) - v3(t.a), v2v0 = v3(t.c) - v3(t.a), rov0 = ro - v3(t.a);
	V3 n = cross(v1v0, v2v0), q = cross(rov0, rd);
	float d = 1.0f / dot(rd, n);
	float u = d * -dot(q, v2v0), v = d * dot(q, v1v0), tt = d * -dot(n, rov0);
	if (u < 0.0f || u > 1.0f || v < 0.0f || (u + v) > 1.0f || tt < 0.0f) tt = std::numeric_limits<float>::max();
	return tt;
}
// TriangleBvh4::ray_intersect  (ngp src/triangle_bvh.cu:271-312; INSTA's header copy of the same loop: ngp bvh.cuh:98-144); mint starts at MAX_DIST = 10
// (triangle_bvh.cu:49).  Returns the BVH-order index of the first hit, -1 = none.
inline int ray_tri(V3 ro, V3 rd, const orc_node* nodes, const orc_tri* tris, float* t_out) {
	int stack[64];
	int sp = 0;
	stack[sp++] = 0;
	float mint = 10.0f;
	int best = -1;
	while (sp > 0) {
		int idx = stack[--sp];
		const orc_node& nd = nodes[idx];
		if (nd.left_idx < 0) {
			int end = -nd.right_idx - 1;
			for (int i = -nd.left_idx - 1; i < end; ++i) {
				float t = tri_ray_t(tris[i], ro, rd);
				if (t < mint) { mint = t; best = i; }
			}
		} else {
			float cd[4];
			uint32_t ci[4];
			for (uint32_t k = 0; k < 4; ++k) {
				cd[k] = box_ray_tmin(nodes[nd.left_idx + k].mn, nodes[nd.left_idx + k].mx, ro, rd);
				ci[k] = nd.left_idx + k;
			}
			auto cas = [&](int i, int j) {  // swap iff values[i] < values[j]  (descending)
				if (cd[i] < cd[j]) { std::swap(cd[i], cd[j]); std::swap(ci[i], ci[j]); }
			};
			cas(0, 2); cas(1, 3); cas(0, 1); cas(2, 3); cas(1, 2);
			for (int k = 0; k < 4; ++k)
				if (cd[k] < mint && sp < 63) stack[sp++] = (int)ci[k];
		}
	}
	if (t_out) *t_out = mint;
	return best;
}

// triangle.cuh:43-57 ; 3x3 column-major (col0 tangent, col1 bitangent, col2 normal)
inline void tbn(const orc_tri& tr, float* m9) {
	V3 t = v3(tr.b) - v3(tr.a);
	V3 e = v3(tr.c) - v3(tr.a);
	V3 n = normalized(cross(t, e));
	V3 bt = normalized(cross(t, n));
	V3 tg = normalized(cross(bt, n));
	m9[0] = tg.x; m9[1] = tg.y; m9[2] = tg.z;
	m9[3] = bt.x; m9[4] = bt.y; m9[5] = bt.z;
	m9[6] = n.x;  m9[7] = n.y;  m9[8] = n.z;
}

// affine.cuh:26-39 ; 4x4 column-major
inline void tri2proj(const orc_tri& tr, float* M) {
	float r[9];
	tbn(tr, r);
	if (tr.mask_id == 1)
		for (int i = 0; i < 9; ++i) r[i] = r[i] * tr.scaling;
	for (int c = 0; c < 3; ++c) {
		M[c * 4 + 0] = r[c * 3 + 0]; M[c * 4 + 1] = r[c * 3 + 1]; M[c * 4 + 2] = r[c * 3 + 2]; M[c * 4 + 3] = 0.f;
	}
	M[12] = tr.a[0]; M[13] = tr.a[1]; M[14] = tr.a[2]; M[15] = 1.f;
}

// General 4x4 inverse in double (the reference uses Eigen's fp32 cofactor inverse; tolerance-class quantity).
inline void inv4(const float* Mf, float* out) {
	double m[16], inv[16];
	for (int i = 0; i < 16; ++i) m[i] = Mf[i];
	inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
	inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
	inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
	inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
	inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
	inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
	inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
	inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
	inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
	inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
	inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
	inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
	inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
	inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
	inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
	inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
	double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
	double id = 1.0 / det;
	for (int i = 0; i < 16; ++i) out[i] = (float)(inv[i] * id);
}

// affine.cuh:45-100
inline void interp_transform(V3 point, const int32_t* adjacency, const orc_tri* tris, const orc_tri& tri, bool invert, float* out) {
	uint32_t id = tri.id;
	float sum = 0.f;
	const float factor = 4.f;
	for (int i = 0; i < 3; ++i) {
		int e = adjacency[id * 3 + i];
		if (e >= 0) sum += expf(-factor * std::sqrt(sqn(centroid(tris[e]) - point)));
	}
	sum += expf(-1.f * std::sqrt(sqn(centroid(tri) - point)));
	float M[16] = {0};
	float P[16];
	for (int i = 0; i < 3; ++i) {
		int e = adjacency[id * 3 + i];
		if (e >= 0) {
			float a = expf(-factor * std::sqrt(sqn(centroid(tris[e]) - point)));
			float w = a / sum;
			tri2proj(tris[e], P);
			for (int k = 0; k < 16; ++k) M[k] += w * P[k];
		}
	}
	float a = expf(-1.f * std::sqrt(sqn(centroid(tri) - point)));
	float w = a / sum;
	tri2proj(tri, P);
	for (int k = 0; k < 16; ++k) M[k] += w * P[k];
	if (invert) inv4(M, out);
	else memcpy(out, M, 64);
}

inline void mat4_mul_vec(const float* M, const float* x, float* y) {
	for (int r = 0; r < 4; ++r) y[r] = M[r] * x[0] + M[4 + r] * x[1] + M[8 + r] * x[2] + M[12 + r] * x[3];
}

struct Pcg32 {  // tcnn/dependencies/pcg32/pcg32.h
	uint64_t state, inc;
	uint32_t next_uint() {
		uint64_t old = state;
		state = old * 0x5851f42d4c957f2dULL + inc;
		uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
		uint32_t rot = (uint32_t)(old >> 59u);
		return (xs >> rot) | (xs << ((~rot + 1u) & 31));
	}
	float next_float() {
		union { uint32_t u; float f; } x;
		x.u = (next_uint() >> 9) | 0x3f800000u;
		return x.f - 1.0f;
	}
	void advance(int64_t delta_) {
		uint64_t cur_mult = 0x5851f42d4c957f2dULL, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
		uint64_t delta = (uint64_t)delta_;
		while (delta > 0) {
			if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
			cur_plus = (cur_mult + 1) * cur_plus;
			cur_mult *= cur_mult;
			delta /= 2;
		}
		state = acc_mult * state + acc_plus;
	}
};

}  // namespace

extern "C" {

void orc_set_arith(int mode) { g_arith = mode ? 1 : 0; }
int orc_get_arith(void) { return g_arith; }

int orc_max_threads(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

// ---------------------------------------------------------------------------------------------------------
// fp16 <-> fp32 (IEEE binary16, round-to-nearest-even, subnormals kept) -- what __float2half_rn / __half2float do.
uint16_t orc_f32_to_f16(float x) {
	uint32_t u;
	memcpy(&u, &x, 4);
	uint32_t sign = (u >> 16) & 0x8000u;
	uint32_t em = u & 0x7fffffffu;
	if (em >= 0x7f800000u) return (uint16_t)(sign | (em > 0x7f800000u ? 0x7e00u : 0x7c00u));
	if (em >= 0x477ff000u) {  // >= 65520 rounds to inf
		return (uint16_t)(sign | 0x7c00u);
	}
	if (em < 0x38800000u) {  // < 2^-14 : subnormal half
		if (em < 0x33000000u) return (uint16_t)sign;  // < 2^-25 -> 0 (ties at exactly 2^-25 go to even = 0)
		uint32_t e = em >> 23;
		uint32_t mant = (em & 0x7fffffu) | 0x800000u;
		uint32_t shift = 126 - e;  // 14..24 -> value = mant * 2^(e-150); half subnormal unit 2^-24
		uint32_t h = mant >> shift;
		uint32_t rem = mant & ((1u << shift) - 1);
		uint32_t half = 1u << (shift - 1);
		if (rem > half || (rem == half && (h & 1))) ++h;
		return (uint16_t)(sign | h);
	}
	uint32_t h = ((em >> 13) - (112u << 10));
	uint32_t rem = em & 0x1fffu;
	if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;
	return (uint16_t)(sign | h);
}

float orc_f16_to_f32(uint16_t h) {
	uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
	uint32_t e = (h >> 10) & 0x1f;
	uint32_t m = h & 0x3ffu;
	uint32_t u;
	if (e == 0) {
		if (m == 0) u = sign;
		else {
			int sh = 0;
			while (!(m & 0x400u)) { m <<= 1; ++sh; }
			m &= 0x3ffu;
			u = sign | ((uint32_t)(113 - sh) << 23) | (m << 13);
		}
	} else if (e == 31) u = sign | 0x7f800000u | (m << 13);
	else u = sign | ((e + 112) << 23) | (m << 13);
	float f;
	memcpy(&f, &u, 4);
	return f;
}

}  // extern "C"

namespace {
// fp16 addition with a single rounding (what HADD does): the exact sum of two halves fits a double.
inline uint16_t f64_to_f16(double x) {
	// round-to-nearest-even directly from double to avoid double rounding through float
	uint64_t u;
	memcpy(&u, &x, 8);
	uint16_t sign = (uint16_t)((u >> 48) & 0x8000u);
	double ax = std::fabs(x);
	if (std::isnan(x)) return sign | 0x7e00u;
	if (ax >= 65520.0) return sign | 0x7c00u;
	if (ax < 5.9604644775390625e-08 * 0.5) return sign;  // < 2^-25
	int e;
	double m = std::frexp(ax, &e);  // ax = m * 2^e, m in [0.5,1)
	int he = e - 1;                  // unbiased exponent of the leading bit
	double scaled;
	int base;
	if (he < -14) {  // subnormal: unit 2^-24
		scaled = std::ldexp(ax, 24);
		base = 0;
		double r = std::nearbyint(scaled);  // default rounding mode: ties-to-even
		return sign | (uint16_t)r;
	}
	(void)m;
	scaled = std::ldexp(ax, 10 - he);  // in [1024, 2048)
	double r = std::nearbyint(scaled);
	uint32_t q = (uint32_t)r;
	base = he + 15;
	if (q == 2048) { q = 1024; ++base; }
	if (base >= 31) return sign | 0x7c00u;
	return sign | (uint16_t)((base << 10) | (q - 1024));
}
inline uint16_t hadd_soft(uint16_t a, uint16_t b) { return f64_to_f16((double)orc_f16_to_f32(a) + (double)orc_f16_to_f32(b)); }
#ifdef __F16C__
// Hardware conversions (round-to-nearest-even, subnormals kept).  fp16 add = fp32 add + one conversion: rounding twice
// is innocuous because fp32 carries 24 >= 2*11+2 significand bits.
inline float h2f(uint16_t h) { return _cvtsh_ss(h); }
inline uint16_t f2h(float f) { return _cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC); }
inline uint16_t hadd(uint16_t a, uint16_t b) { return f2h(h2f(a) + h2f(b)); }
#else
inline float h2f(uint16_t h) { return orc_f16_to_f32(h); }
inline uint16_t f2h(float f) { return orc_f32_to_f16(f); }
inline uint16_t hadd(uint16_t a, uint16_t b) { return hadd_soft(a, b); }
#endif

// pos_fract, tcnn common_device.h:425-431.  `input * scale + 0.5f` is a single FFMA in the reference's CUDA build
// (the only build of this function that exists), hence fmaf.
inline void pos_fract(float input, float scale, float* frac, uint32_t* grid) {
	float p = fmaf(input, scale, 0.5f);
	int tmp = (int)floorf(p);
	*grid = (uint32_t)tmp;
	*frac = p - (float)tmp;
}

// grid_index, grid.h:100-117 (feature 0; entry units, F multiplied by the caller)
inline uint32_t grid_entry(uint32_t hashmap_size, uint32_t res, const uint32_t g[3]) {
	uint32_t stride = 1, index = 0;
	for (uint32_t dim = 0; dim < 3 && stride <= hashmap_size; ++dim) {
		index += g[dim] * stride;
		stride *= res;
	}
	if (hashmap_size < stride) index = (g[0] * 1u) ^ (g[1] * 2654435761u) ^ (g[2] * 805459861u);  // fast_hash, grid.h:82-98
	return index % hashmap_size;
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------------------
// mesh preparation
void orc_build_triangles(const float* verts, const uint32_t* faces, uint32_t n_faces, const uint32_t* mask_ids,
                         const orc_tri* canonical, orc_tri* out) {  // tiny_mesh.cu:134-155
	auto area = [](const orc_tri& t) {  // triangle.cuh:35-37
		return 0.5f * std::sqrt(sqn(cross(v3(t.b) - v3(t.a), v3(t.c) - v3(t.a))));
	};
	for (uint32_t f = 0; f < n_faces; ++f) {
		orc_tri t;
		for (int k = 0; k < 3; ++k) {
			t.a[k] = verts[faces[3 * f + 0] * 3 + k];
			t.b[k] = verts[faces[3 * f + 1] * 3 + k];
			t.c[k] = verts[faces[3 * f + 2] * 3 + k];
		}
		t.id = f;
		t.mask_id = mask_ids ? mask_ids[f] : 15u;
		float s = 1;
		if (canonical) s *= area(t) / area(canonical[f]);
		t.scaling = s;
		out[f] = t;
	}
}

void orc_mesh_aabb(const float* verts, uint32_t n_verts, float* aabb6) {  // tiny_mesh.cu:157-162
	float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
	for (uint32_t i = 0; i < n_verts; ++i)
		for (int k = 0; k < 3; ++k) {
			mn[k] = std::min(mn[k], verts[3 * i + k]);
			mx[k] = std::max(mx[k], verts[3 * i + k]);
		}
	V3 d = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
	float amount = std::sqrt(sqn(d)) * 0.15f;
	for (int k = 0; k < 3; ++k) {
		aabb6[k] = mn[k] - amount;
		aabb6[3 + k] = mx[k] + amount;
	}
}

void orc_adjacency(const uint32_t* faces, uint32_t n_faces, uint32_t n_verts, int32_t* out3) {  // tiny_mesh.cu:316-371
	std::vector<std::set<int>> star(n_verts);
	for (uint32_t f = 0; f < n_faces; ++f)
		for (int k = 0; k < 3; ++k) star[faces[3 * f + k]].insert((int)f);
	for (uint32_t f = 0; f < n_faces; ++f) {
		const auto &na = star[faces[3 * f]], &nb = star[faces[3 * f + 1]], &nc = star[faces[3 * f + 2]];
		std::set<int> uniq;
		std::set_intersection(na.begin(), na.end(), nb.begin(), nb.end(), std::inserter(uniq, uniq.end()));
		std::set_intersection(na.begin(), na.end(), nc.begin(), nc.end(), std::inserter(uniq, uniq.end()));
		std::set_intersection(nb.begin(), nb.end(), nc.begin(), nc.end(), std::inserter(uniq, uniq.end()));
		std::vector<int> ids(uniq.begin(), uniq.end());
		while (ids.size() < 4) ids.push_back(-1);
		out3[3 * f] = out3[3 * f + 1] = out3[3 * f + 2] = -1;
		int k = 0;
		for (int n : ids)
			if (n != (int)f && k < 3) out3[3 * f + k++] = n;
	}
}

void orc_triangle_masks(const uint32_t* faces, uint32_t n_faces, const int32_t* region_ids, const uint32_t* region_offsets,
                        const uint32_t* region_verts, uint32_t n_regions, uint32_t* out_mask) {  // masking.cpp:118-148
	for (uint32_t f = 0; f < n_faces; ++f) {
		uint32_t m = 15;  // FlameMasks::Torso
		for (uint32_t r = 0; r < n_regions; ++r) {
			for (uint32_t k = region_offsets[r]; k < region_offsets[r + 1]; ++k) {
				uint32_t v = region_verts[k];
				if (faces[3 * f] == v || faces[3 * f + 1] == v || faces[3 * f + 2] == v) {
					m = (uint32_t)region_ids[r];
					break;
				}
			}
		}
		out_mask[f] = m;
	}
}

// ---------------------------------------------------------------------------------------------------------
// BVH build, triangle_bvh.cu:548-631
int orc_bvh_build(orc_tri* tris, int n_tris, int leaf_size, orc_node* nodes_out, int max_nodes) {
	std::vector<orc_node> nodes;
	nodes.emplace_back();
	box_of(tris, tris + n_tris, nodes[0].mn, nodes[0].mx);
	struct BuildNode { int node_idx; orc_tri *begin, *end; };
	std::vector<BuildNode> stack;
	stack.push_back({0, tris, tris + n_tris});
	while (!stack.empty()) {
		BuildNode curr = stack.back();
		stack.pop_back();
		BuildNode ch[4];
		ch[0].begin = curr.begin;
		ch[0].end = curr.end;
		int n_children = 1;
		while (n_children < 4) {
			for (int i = n_children - 1; i >= 0; --i) {
				BuildNode& c = ch[i];
				float cnt = (float)(c.end - c.begin);
				V3 mean = {0, 0, 0};
				for (orc_tri* it = c.begin; it != c.end; ++it) mean = mean + centroid(*it);
				mean = mean / cnt;
				V3 var = {0, 0, 0};
				for (orc_tri* it = c.begin; it != c.end; ++it) {
					V3 d = centroid(*it) - mean;
					var = var + V3{d.x * d.x, d.y * d.y, d.z * d.z};
				}
				var = var / cnt;
				int axis = 0;  // Eigen maxCoeff: first strictly-greatest coefficient
				float best = var.x;
				if (var.y > best) { best = var.y; axis = 1; }
				if (var.z > best) { best = var.z; axis = 2; }
				orc_tri* m = c.begin + (c.end - c.begin) / 2;
				std::nth_element(c.begin, m, c.end, [&](const orc_tri& t1, const orc_tri& t2) { return centroid_axis(t1, axis) < centroid_axis(t2, axis); });
				ch[i * 2].begin = ch[i].begin;
				ch[i * 2 + 1].end = ch[i].end;
				ch[i * 2].end = ch[i * 2 + 1].begin = m;
			}
			n_children *= 2;
		}
		nodes[curr.node_idx].left_idx = (int)nodes.size();
		for (int i = 0; i < 4; ++i) {
			BuildNode& c = ch[i];
			c.node_idx = (int)nodes.size();
			nodes.emplace_back();
			box_of(c.begin, c.end, nodes.back().mn, nodes.back().mx);
			if (c.end - c.begin <= leaf_size) {
				nodes.back().left_idx = -(int)(c.begin - tris) - 1;
				nodes.back().right_idx = -(int)(c.end - tris) - 1;
			} else {
				stack.push_back(c);
			}
		}
		nodes[curr.node_idx].right_idx = (int)nodes.size();
	}
	if ((int)nodes.size() > max_nodes) return -(int)nodes.size();
	memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(orc_node));
	return (int)nodes.size();
}

void orc_closest_triangle(const float* pts, int n, const orc_node* nodes, const orc_tri* tris, int* out_idx, float* out_dist,
                          int n_threads) {
#pragma omp parallel for schedule(static) num_threads(n_threads)
	for (int i = 0; i < n; ++i) {
		float d;
		out_idx[i] = closest_tri(v3(pts + 3 * i), nodes, tris, &d);
		if (out_dist) out_dist[i] = d;
	}
}


void orc_ray_intersect(const float* ro, const float* rd, int n, const orc_node* nodes, const orc_tri* tris, int* out_idx, float* out_t) {
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; ++i) {
		float t;
		out_idx[i] = ray_tri(v3(ro + 3 * i), v3(rd + 3 * i), nodes, tris, &t);
		if (out_t) out_t[i] = t;
	}
}

void orc_closest_triangle_bounded(const float* pts, int n, const orc_node* nodes, const orc_tri* tris, const float* init_best, int* out_idx, int n_threads) {
#pragma omp parallel for schedule(static) num_threads(n_threads)
	for (int i = 0; i < n; ++i) out_idx[i] = closest_tri(v3(pts + 3 * i), nodes, tris, nullptr, init_best[i]);
}

// Diagnostic (design studies only): traversal statistics of closest_tri with the product's pop-time rule for inner nodes.
// out8 = {expands, inner entries dropped at pop, leaf visits, triangle tests, leaf visits whose box distance exceeded
//         `best` at pop (would be skippable), stack high-water mark, queries, 0}
void orc_closest_stats(const float* pts, int n, const orc_node* nodes, const orc_tri* tris, double* out8) {
	orc_closest_stats_bounded(pts, n, nodes, tris, nullptr, out8);
}

// Same traversal started from an initial upper bound per query (init_best[q] >= the true minimum, e.g. the distance to a
// candidate triangle).  Pruned subtrees cannot hold the winner and the relative visit order of the rest is unchanged, so the
// winner (last visited among the tied minima, bvh.cuh:82-85) is the reference's.  Diagnostic for the cost model in DESIGN.md.
void orc_closest_stats_bounded(const float* pts, int n, const orc_node* nodes, const orc_tri* tris, const float* init_best, double* out8) {
	double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	for (int q = 0; q < n; ++q) {
		V3 p = v3(pts + 3 * q);
		int stack[64];
		float sd[64];
		int sp = 0;
		stack[sp] = 0; sd[sp] = 0.f; ++sp;
		float best = init_best ? init_best[q] : std::numeric_limits<float>::max();
		while (sp > 0) {
			--sp;
			int idx = stack[sp];
			float dpush = sd[sp];
			const orc_node& nd = nodes[idx];
			if (nd.left_idx < 0) {
				acc[2] += 1;
				if (dpush > best) acc[4] += 1;
				int end = -nd.right_idx - 1;
				for (int i = -nd.left_idx - 1; i < end; ++i) {
					acc[3] += 1;
					{  // how many exact tests a conservative per-triangle bound (its own AABB) would have skipped
						const orc_tri& t = tris[i];
						float mn[3], mx[3];
						for (int k = 0; k < 3; ++k) {
							mn[k] = std::min(t.a[k], std::min(t.b[k], t.c[k]));
							mx[k] = std::max(t.a[k], std::max(t.b[k], t.c[k]));
						}
						if (box_distance_sq(mn, mx, p) > best) acc[7] += 1;
					}
					float d = tri_distance_sq(tris[i], p);
					if (d <= best) best = d;
				}
			} else {
				if (dpush > best) { acc[1] += 1; continue; }
				acc[0] += 1;
				float cd[4];
				uint32_t ci[4];
				for (uint32_t k = 0; k < 4; ++k) {
					cd[k] = box_distance_sq(nodes[nd.left_idx + k].mn, nodes[nd.left_idx + k].mx, p);
					ci[k] = nd.left_idx + k;
				}
				auto cas = [&](int i, int j) {
					if (cd[i] < cd[j]) { std::swap(cd[i], cd[j]); std::swap(ci[i], ci[j]); }
				};
				cas(0, 2); cas(1, 3); cas(0, 1); cas(2, 3); cas(1, 2);
				for (int k = 0; k < 4; ++k)
					if (cd[k] <= best && sp < 63) { stack[sp] = (int)ci[k]; sd[sp] = cd[k]; ++sp; }
				if (sp > acc[5]) acc[5] = sp;
			}
		}
		acc[6] += 1;
	}
	for (int k = 0; k < 8; ++k) out8[k] = acc[k];
}


// ---------------------------------------------------------------------------------------------------------
// The product's cylinder bounds (insta_b200/csrc/deform.cu: build_cyl_kernel, cyl_cull), restated on the CPU.  NOT in the reference: this is
// the checker of the product's pruning argument -- a child whose bound exceeds `best` holds no triangle the reference would accept (d <= best),
// so skipping it changes neither `best` nor the winner.  tests/test_oracle_cyl_bound.py runs the reference traversal with the extra cull and
// (a) compares the winners with the plain traversal, (b) checks every culled subtree exhaustively.
namespace {
// nearest half >= v (up = true) or <= v (up = false); v finite
inline uint16_t f2h_dir(float v, bool up) {
	if (v == 0.f) return 0;
	const bool neg = v < 0.f;
	const float a = std::fabs(v);
	const bool mag_up = neg ? !up : up;  // round the magnitude up?
	if (a > 65504.f) return (uint16_t)((neg ? 0x8000u : 0u) | (mag_up ? 0x7c00u : 0x7bffu));
	int ex;
	std::frexp(a, &ex);  // a = f * 2^ex, f in [0.5, 1)
	int e = ex - 1;      // a in [2^e, 2^(e+1))
	if (e < -14) e = -14;
	const double ulp = std::ldexp(1.0, e - 10);
	double k = (double)a / ulp;  // exact (power-of-two scaling)
	double kk = mag_up ? std::ceil(k) : std::floor(k);
	uint32_t bits;
	if (e == -14 && kk < 1024.0) bits = (uint32_t)kk;  // subnormal
	else {
		if (kk >= 2048.0) { kk = 1024.0; ++e; }
		bits = ((uint32_t)(e + 15) << 10) | ((uint32_t)kk - 1024u);
	}
	return (uint16_t)((neg ? 0x8000u : 0u) | bits);
}
inline uint16_t f2h_rn(float v) {
	const uint16_t lo = f2h_dir(v, false), hi = f2h_dir(v, true);
	const float fl = h2f(lo), fh = h2f(hi);
	if (lo == hi) return lo;
	const float dl = v - fl, dh = fh - v;
	if (dl < dh) return lo;
	if (dh < dl) return hi;
	return (lo & 1u) ? hi : lo;
}
}  // namespace

// out: 8 halves per NODE (nx, ny, nz, zc2, h2, r2, 0, 0); node 0 (the root) is never tested and gets the infinite cylinder
void orc_cyl_build(const orc_node* nodes, int n_nodes, const orc_tri* tris, uint16_t* out) {
	std::vector<int> lo(n_nodes), hi(n_nodes);
	for (int i = n_nodes - 1; i >= 0; --i) {
		if (nodes[i].left_idx < 0) { lo[i] = -nodes[i].left_idx - 1; hi[i] = -nodes[i].right_idx - 1; }
		else {
			lo[i] = 1 << 30; hi[i] = 0;
			for (int k = 0; k < 4; ++k) { lo[i] = std::min(lo[i], lo[nodes[i].left_idx + k]); hi[i] = std::max(hi[i], hi[nodes[i].left_idx + k]); }
		}
	}
	for (int i = 0; i < n_nodes; ++i) {
		uint16_t* rec = out + 8 * i;
		const uint16_t inf_rec[8] = {0, 0, 0x3c00, 0, 0x7c00, 0x7c00, 0, 0};
		bool never = i == 0 || lo[i] >= hi[i];
		float nx = 0.f, ny = 0.f, nz = 0.f;
		for (int t = lo[i]; t < hi[i] && !never; ++t) {
			const orc_tri& T = tris[t];
			const float ux = T.b[0] - T.a[0], uy = T.b[1] - T.a[1], uz = T.b[2] - T.a[2];
			const float vx = T.c[0] - T.a[0], vy = T.c[1] - T.a[1], vz = T.c[2] - T.a[2];
			const float cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
			const float c2 = cx * cx + cy * cy + cz * cz, uv = (ux * ux + uy * uy + uz * uz) * (vx * vx + vy * vy + vz * vz);
			if (!(c2 >= 1e-4f * uv) || !(c2 > 0.f)) never = true;
			nx += cx; ny += cy; nz += cz;
		}
		if (never) { memcpy(rec, inf_rec, 16); continue; }
		const float l = std::sqrt(nx * nx + ny * ny + nz * nz);
		if (l > 0.f && l < std::numeric_limits<float>::max()) { nx /= l; ny /= l; nz /= l; } else { nx = 0.f; ny = 0.f; nz = 1.f; }
		const uint16_t hx = f2h_dir(nx, nx >= 0.f), hy = f2h_dir(ny, ny >= 0.f), hz = f2h_dir(nz, nz >= 0.f);
		const double ax = h2f(hx), ay = h2f(hy), az = h2f(hz), s2 = ax * ax + ay * ay + az * az;
		const orc_node& nd = nodes[i];
		const double c2x = (double)nd.mn[0] + nd.mx[0], c2y = (double)nd.mn[1] + nd.mx[1], c2z = (double)nd.mn[2] + nd.mx[2];
		double zlo = 1e300, zhi = -1e300, lat = 0.0;
		for (int t = lo[i]; t < hi[i]; ++t) {
			const float* v[3] = {tris[t].a, tris[t].b, tris[t].c};
			for (int j = 0; j < 3; ++j) {
				const double ux = 2.0 * v[j][0] - c2x, uy = 2.0 * v[j][1] - c2y, uz = 2.0 * v[j][2] - c2z;
				const double z = ax * ux + ay * uy + az * uz;
				zlo = std::min(zlo, z); zhi = std::max(zhi, z);
				lat = std::max(lat, ux * ux + uy * uy + uz * uz - z * z / s2);
			}
		}
		const double size2 = 2.0 * (((double)nd.mx[0] - nd.mn[0]) + ((double)nd.mx[1] - nd.mn[1]) + ((double)nd.mx[2] - nd.mn[2]));
		const double slack = 2e-6 * size2 + 1e-12;
		const uint16_t hzc = f2h_rn((float)(0.5 * (zlo + zhi)));
		const double zc = h2f(hzc);
		const double h2 = std::max(zhi - zc, zc - zlo) + slack, r2 = std::sqrt(std::max(lat, 0.0)) + slack;
		if (!(size2 >= 0.0) || !(h2 < 6e4) || !(r2 < 6e4) || !(std::fabs(zc) < 6e4)) { memcpy(rec, inf_rec, 16); continue; }
		auto up = [](double v) { float f = (float)v; if ((double)f < v) f = std::nextafter(f, std::numeric_limits<float>::infinity()); return f2h_dir(f, true); };
		rec[0] = hx; rec[1] = hy; rec[2] = hz; rec[3] = hzc; rec[4] = up(h2); rec[5] = up(r2); rec[6] = rec[7] = 0;
	}
}

// cyl_cull of deform.cu in fp32
static inline bool cyl_cull_cpu(const uint16_t* rec, const float* mn, const float* mx, V3 p, float best4) {
	const float pp[3] = {p.x, p.y, p.z};
	float D[3];
	for (int k = 0; k < 3; ++k) D[k] = (pp[k] - mx[k]) - (mn[k] - pp[k]);
	const float nx = h2f(rec[0]), ny = h2f(rec[1]), nz = h2f(rec[2]), zc = h2f(rec[3]), h = h2f(rec[4]), r = h2f(rec[5]);
	const float q = fmaf(nx, D[0], fmaf(ny, D[1], nz * D[2]));
	const float DD = fmaf(D[0], D[0], fmaf(D[1], D[1], D[2] * D[2]));
	const float rr = std::sqrt(std::max(fmaf(-q, q, DD), 0.f));
	const float dz = std::max(std::fabs(q - zc) - h, 0.f), dr = std::max(rr - r, 0.f);
	return fmaf(dz, dz, dr * dr) > fmaf(DD, 4e-6f, best4);
}

// The reference traversal (bvh.cuh:48-96) started from init_best (nullable: FLT_MAX) with the product's extra cull.
// out_stats4 (nullable) += {quad expansions, leaf visits, children culled by the cylinder that the box test had passed, violations};
// check != 0: every culled subtree is searched exhaustively for a triangle with d <= best (a violation of the argument).
void orc_closest_triangle_cyl(const float* pts, int n, const orc_node* nodes, const orc_tri* tris, const uint16_t* cyl, const float* init_best, int* out_idx,
                              double* out_stats4, int check) {
	double acc[4] = {0, 0, 0, 0};
	for (int q = 0; q < n; ++q) {
		const V3 p = v3(pts + 3 * q);
		int stack[64];
		int sp = 0;
		stack[sp++] = 0;
		float best = init_best ? init_best[q] : std::numeric_limits<float>::max();
		int best_i = -1;
		while (sp > 0) {
			const int idx = stack[--sp];
			const orc_node& nd = nodes[idx];
			if (nd.left_idx < 0) {
				acc[1] += 1;
				const int end = -nd.right_idx - 1;
				for (int i = -nd.left_idx - 1; i < end; ++i) {
					const float d = tri_distance_sq(tris[i], p);
					if (d <= best) { best = d; best_i = i; }
				}
			} else {
				acc[0] += 1;
				float cd[4];
				uint32_t ci[4];
				const float best4 = best * 4.0088f;
				for (uint32_t k = 0; k < 4; ++k) {
					const orc_node& c = nodes[nd.left_idx + k];
					cd[k] = box_distance_sq(c.mn, c.mx, p);
					ci[k] = nd.left_idx + k;
					if (cyl_cull_cpu(cyl + 8 * ci[k], c.mn, c.mx, p, best4)) ci[k] |= 0x80000000u;
				}
				auto cas = [&](int i, int j) {
					if (cd[i] < cd[j]) { std::swap(cd[i], cd[j]); std::swap(ci[i], ci[j]); }
				};
				cas(0, 2); cas(1, 3); cas(0, 1); cas(2, 3); cas(1, 2);
				for (int k = 0; k < 4; ++k) {
					if (!(cd[k] <= best)) continue;
					if (ci[k] & 0x80000000u) {
						acc[2] += 1;
						if (check) {
							std::vector<int> todo{(int)(ci[k] & 0x7fffffffu)};
							while (!todo.empty()) {
								const int j = todo.back();
								todo.pop_back();
								if (nodes[j].left_idx < 0) {
									for (int i = -nodes[j].left_idx - 1; i < -nodes[j].right_idx - 1; ++i)
										if (tri_distance_sq(tris[i], p) <= best) acc[3] += 1;
								} else
									for (int c = 0; c < 4; ++c) todo.push_back(nodes[j].left_idx + c);
							}
						}
						continue;
					}
					if (sp < 63) stack[sp++] = (int)ci[k];
				}
			}
		}
		out_idx[q] = best_i;
	}
	if (out_stats4)
		for (int k = 0; k < 4; ++k) out_stats4[k] += acc[k];
}

float orc_triangle_distance_sq(const orc_tri* t, const float* p) { return tri_distance_sq(*t, v3(p)); }
float orc_bbox_distance_sq(const float* mn, const float* mx, const float* p) { return box_distance_sq(mn, mx, v3(p)); }
void orc_tri2projection(const orc_tri* t, float* out16) { tri2proj(*t, out16); }

// ---------------------------------------------------------------------------------------------------------
// warp / warp_random_samples, core.cu:442-579
void orc_warp(float* samples, uint32_t n, uint32_t stride, int32_t frame_id, uint32_t n_meshes, const orc_node* const* nodes_ptrs,
              const orc_tri* const* tris_bvh_ptrs, const orc_tri* const* tris_topo_ptrs, const orc_tri* canon,
              const int32_t* adjacency, int disable_interp, const float* aabb6, uint32_t extra_dims, int use_geo,
              const int32_t* frame_ids, int32_t* out_tri, int32_t* out_bvh_idx, int n_threads) {
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
			if (samples[j + 4] < 0.f) continue;                                   // core.cu:470
			f = frame_id < 0 ? (uint32_t)samples[j + 4] : (uint32_t)frame_id;     // core.cu:471
		}
		if (f >= n_meshes) continue;                                              // core.cu:472
		V3 p = v3(samples + j);
		if (!xyz_only) {                                                          // core.cu:477, bounding_box.cuh:216-221
			if (!(p.x >= aabb6[0] && p.x <= aabb6[3] && p.y >= aabb6[1] && p.y <= aabb6[4] && p.z >= aabb6[2] && p.z <= aabb6[5])) continue;
		}
		const orc_tri* tbvh = tris_bvh_ptrs[f];
		int id = closest_tri(p, nodes_ptrs[f], tbvh, nullptr);
		if (id == -1) continue;
		const orc_tri& deformed = tbvh[id];
		const orc_tri& canonical = canon[deformed.id];
		int mask_id = (int)deformed.mask_id;
		if (out_tri) out_tri[i] = (int32_t)deformed.id;
		if (out_bvh_idx) out_bvh_idx[i] = id;

		float w2t[16], t2c[16];
		if (mask_id != 2 || disable_interp) {                                      // core.cu:495-498
			float M[16];
			tri2proj(deformed, M);
			inv4(M, w2t);
			tri2proj(canonical, t2c);
		} else {                                                                  // core.cu:492-493
			interp_transform(p, adjacency, tris_topo_ptrs[f], deformed, true, w2t);
			interp_transform(p, adjacency, canon, canonical, false, t2c);
		}
		float x[4] = {p.x, p.y, p.z, 1.f}, xl[4], xc[4];                          // affine.cuh:102-108
		mat4_mul_vec(w2t, x, xl);
		mat4_mul_vec(t2c, xl, xc);
		samples[j + 0] = xc[0]; samples[j + 1] = xc[1]; samples[j + 2] = xc[2];
		if (xyz_only) continue;

		float d[4] = {samples[j + 8] * 2.0f - 1.0f, samples[j + 9] * 2.0f - 1.0f, samples[j + 10] * 2.0f - 1.0f, 0.f};  // core.cu:66
		float dl[4], dc[4];
		mat4_mul_vec(w2t, d, dl);  // w = 0 -> only the 3x3 block acts (affine.cuh:110-116)
		dl[3] = 0.f;
		mat4_mul_vec(t2c, dl, dc);
		V3 cd = normalized(V3{dc[0], dc[1], dc[2]});
		samples[j + 8] = (cd.x + 1.0f) * 0.5f;                                     // core.cu:62
		samples[j + 9] = (cd.y + 1.0f) * 0.5f;
		samples[j + 10] = (cd.z + 1.0f) * 0.5f;
		if (mask_id != 7)                                                          // core.cu:515-519
			for (uint32_t m = 0; m < extra_dims; ++m) samples[j + stride - extra_dims + m] = 1.f;
		if (use_geo) {                                                             // core.cu:522-525
			V3 v = normalized(p - centroid(deformed));
			V3 nrm = normalized(cross(v3(deformed.b) - v3(deformed.a), v3(deformed.c) - v3(deformed.a)));
			samples[j + stride - 1] = dot(v, nrm);
		}
	}
}

void orc_random_frame_ids(uint32_t n, uint32_t n_meshes, uint64_t rng_state, uint64_t rng_inc, int32_t* out) {  // core.cu:543-549
	uint32_t nm = n_meshes - 1;
	for (uint32_t i = 0; i < n; ++i) {
		Pcg32 rng{rng_state, rng_inc};
		rng.advance((int64_t)(uint32_t)(i * nm));
		float r = rng.next_float();
		out[i] = (int32_t)(uint32_t)(r * (float)nm);
	}
}

// ---------------------------------------------------------------------------------------------------------
// hash grid
void orc_grid_layout(uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale, uint32_t* offsets,
                     float* scales, uint32_t* resolutions) {  // grid.h:898-922 (+ the kernel's own evaluation, 188-189)
	uint32_t offset = 0;
	const float log2_pls = std::log2(per_level_scale);
	for (uint32_t i = 0; i < n_levels; ++i) {
		// kernel form: exp2f(level * log2_per_level_scale) * base_resolution - 1.0f, contracted to one FFMA by nvcc
		const float scale = fmaf(exp2f((float)i * log2_pls), (float)base_resolution, -1.0f);
		const uint32_t res = (uint32_t)(ceilf(scale)) + 1;
		uint32_t max_params = 0xffffffffu / 2;
		uint32_t params = std::pow((float)res, 3) > (float)max_params ? max_params : res * res * res;
		params = (params + 7u) / 8u * 8u;
		params = std::min(params, 1u << log2_hashmap_size);
		offsets[i] = offset;
		offset += params;
		scales[i] = scale;
		resolutions[i] = res;
	}
	offsets[n_levels] = offset;
}

void orc_grid_slots(uint32_t n, const float* pos, uint32_t pos_stride, uint32_t n_levels, const uint32_t* offsets, const float* scales,
                    const uint32_t* resolutions, uint32_t* out_slots, float* out_weights) {
	for (uint32_t i = 0; i < n; ++i)
		for (uint32_t l = 0; l < n_levels; ++l) {
			const uint32_t hs = offsets[l + 1] - offsets[l];
			float fr[3];
			uint32_t g[3];
			for (int d = 0; d < 3; ++d) pos_fract(pos[(size_t)i * pos_stride + d], scales[l], &fr[d], &g[d]);
			for (uint32_t c = 0; c < 8; ++c) {
				float w = 1;
				uint32_t gl[3];
				for (int d = 0; d < 3; ++d) {
					if ((c & (1u << d)) == 0) { w *= 1 - fr[d]; gl[d] = g[d]; }
					else { w *= fr[d]; gl[d] = g[d] + 1; }
				}
				size_t o = ((size_t)i * n_levels + l) * 8 + c;
				out_slots[o] = offsets[l] + grid_entry(hs, resolutions[l], gl);
				if (out_weights) out_weights[o] = w;
			}
		}
}

void orc_grid_forward(uint32_t n, const float* pos, uint32_t pos_stride, uint32_t n_levels, const uint32_t* offsets, const float* scales,
                      const uint32_t* resolutions, const uint16_t* table, uint16_t* out, int n_threads) {  // grid.h:138-266
#pragma omp parallel for schedule(static) num_threads(n_threads)
	for (int64_t i = 0; i < (int64_t)n; ++i)
		for (uint32_t l = 0; l < n_levels; ++l) {
			const uint32_t hs = offsets[l + 1] - offsets[l];
			float fr[3];
			uint32_t g[3];
			for (int d = 0; d < 3; ++d) pos_fract(pos[(size_t)i * pos_stride + d], scales[l], &fr[d], &g[d]);
			uint16_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
			for (uint32_t c = 0; c < 8; ++c) {
				float w = 1;
				uint32_t gl[3];
				for (int d = 0; d < 3; ++d) {
					if ((c & (1u << d)) == 0) { w *= 1 - fr[d]; gl[d] = g[d]; }
					else { w *= fr[d]; gl[d] = g[d] + 1; }
				}
				const uint16_t* e = table + ((size_t)offsets[l] + grid_entry(hs, resolutions[l], gl)) * 8;
				for (int f = 0; f < 8; ++f) acc[f] = hadd(acc[f], f2h(w * h2f(e[f])));  // grid.h:257-259
			}
			for (int f = 0; f < 8; ++f) out[(size_t)i + (size_t)(l * 8 + f) * n] = acc[f];                   // grid.h:265
		}
}

// software vs hardware fp16 paths (self-check for tests): returns the number of mismatches over a sweep
uint32_t orc_f16_selfcheck(uint32_t n, uint32_t seed) {
	uint32_t bad = 0, x = seed * 2654435761u + 1u;
	for (uint32_t i = 0; i < n; ++i) {
		x = x * 1664525u + 1013904223u;
		uint16_t a = (uint16_t)(x >> 16);
		x = x * 1664525u + 1013904223u;
		uint16_t b = (uint16_t)(x >> 16);
		if ((a & 0x7c00u) == 0x7c00u || (b & 0x7c00u) == 0x7c00u) continue;
		if (hadd(a, b) != hadd_soft(a, b)) ++bad;
		float f;
		uint32_t u = x ^ (x << 13);
		memcpy(&f, &u, 4);
		if (f == f && f2h(f) != orc_f32_to_f16(f)) ++bad;
		if (h2f(a) != orc_f16_to_f32(a)) ++bad;
	}
	return bad;
}

void orc_grid_backward(uint32_t n, const float* pos, uint32_t pos_stride, uint32_t n_levels, const uint32_t* offsets, const float* scales,
                       const uint32_t* resolutions, const uint16_t* dLdy, double* grad_out, uint32_t* hits_out) {  // grid.h:312-435
	for (uint32_t i = 0; i < n; ++i)
		for (uint32_t l = 0; l < n_levels; ++l) {
			const uint32_t hs = offsets[l + 1] - offsets[l];
			float fr[3];
			uint32_t g[3];
			for (int d = 0; d < 3; ++d) pos_fract(pos[(size_t)i * pos_stride + d], scales[l], &fr[d], &g[d]);
			for (uint32_t c = 0; c < 8; ++c) {
				float w = 1;
				uint32_t gl[3];
				for (int d = 0; d < 3; ++d) {
					if ((c & (1u << d)) == 0) { w *= 1 - fr[d]; gl[d] = g[d]; }
					else { w *= fr[d]; gl[d] = g[d] + 1; }
				}
				size_t e = (size_t)offsets[l] + grid_entry(hs, resolutions[l], gl);
				if (hits_out) hits_out[e]++;
				for (int f = 0; f < 8; ++f) {
					float gy = orc_f16_to_f32(dLdy[(size_t)i + (size_t)(l * 8 + f) * n]);
					grad_out[e * 8 + f] += (double)orc_f16_to_f32(orc_f32_to_f16(gy * w));  // grid.h:352-353
				}
			}
		}
}

}  // extern "C"
