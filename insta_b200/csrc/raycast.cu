// The FLAME-mesh overlay of the INSTA viewer / recorder on sm_100a: rta::Core::raycast_flame (src/core.cu:350-402) = SphereTracer::init_rays_from_camera
// (ngp src/testbed_sdf.cu:493-577, 600-642: one ray per pixel, advanced to the render box) + SphereTracer::trace_bvh (testbed_sdf.cu:650-665) ->
// TriangleBvh::ray_trace_gpu -> raytrace_kernel (ngp src/triangle_bvh.cu:721-738) -> TriangleBvh4::ray_intersect (271-312; INSTA's header copy of the same loop:
// ngp include/neural-graphics-primitives/bvh.cuh:98-144) with BoundingBox::ray_intersect (bounding_box.cuh:163-205) and Triangle::ray_intersect / normal
// (triangle.cuh:39-41, 66-85) + shade_kernel_sdf (src/core.cu:74-109).
//
// One kernel per entry point, one thread per ray.  The traversal walks this library's Quad layout (the four children of a reference node in one 128-byte
// line, ctx.h) in the reference's order: the children of a node sorted by their box-entry parameter with the reference's 5-comparator network (descending,
// so the nearest is pushed last and popped first), pushed while `t_box < mint`, a leaf's triangles tested in BVH order with `t < mint` -- so among triangles
// hit at the same parameter (a ray through a shared edge) the winner is the reference's.
// Arithmetic: plain IEEE single precision, one rounding per operation in the reference's expression order (the flavour of the reference's HOST build, which
// is what the parity tests can run without the viewer).  nvcc contracts some of these products into FMAs in the reference's device build; the two
// differ in the last bit of t and only between triangles a ray hits at (nearly) the same parameter.
#include <cfloat>

#include "ctx.h"

namespace insta {
namespace {

__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
// Eigen's fixed-size-3 dot reduces as a0 b0 + (a1 b1 + a2 b2)  (Eigen/src/Core/Redux.h)
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) { return fadd(fmul(ax, bx), fadd(fmul(ay, by), fmul(az, bz))); }

// BoundingBox::ray_intersect(pos, dir).x(): entry parameter of the slab test, FLT_MAX when the ray misses.  Negative when the origin is inside or the box
// lies behind the origin -- the reference pushes such children too (t_box < mint), so does the traversal below.
__device__ __forceinline__ float box_ray_tmin(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, float ox, float oy, float oz, float dx, float dy, float dz) {
	float tmin = fdiv(fsub(mnx, ox), dx), tmax = fdiv(fsub(mxx, ox), dx);
	if (tmin > tmax) { const float s = tmin; tmin = tmax; tmax = s; }
	float tymin = fdiv(fsub(mny, oy), dy), tymax = fdiv(fsub(mxy, oy), dy);
	if (tymin > tymax) { const float s = tymin; tymin = tymax; tymax = s; }
	if (tmin > tymax || tymin > tmax) return FLT_MAX;
	if (tymin > tmin) tmin = tymin;
	if (tymax < tmax) tmax = tymax;
	float tzmin = fdiv(fsub(mnz, oz), dz), tzmax = fdiv(fsub(mxz, oz), dz);
	if (tzmin > tzmax) { const float s = tzmin; tzmin = tzmax; tzmax = s; }
	if (tmin > tzmax || tzmin > tmax) return FLT_MAX;
	if (tzmin > tmin) tmin = tzmin;
	return tmin;
}

// Triangle::ray_intersect (triangle.cuh:66-81)
__device__ __forceinline__ float tri_ray_t(const Tri* __restrict__ t, float ox, float oy, float oz, float dx, float dy, float dz) {
	const float4 r0 = __ldg(reinterpret_cast<const float4*>(t)), r1 = __ldg(reinterpret_cast<const float4*>(t) + 1);
	const float cz = __ldg(reinterpret_cast<const float*>(t) + 8);
	const float ax = r0.x, ay = r0.y, az = r0.z, bx = r0.w, by = r1.x, bz = r1.y, cx = r1.z, cy = r1.w;
	const float e1x = fsub(bx, ax), e1y = fsub(by, ay), e1z = fsub(bz, az);  // v1v0
	const float e2x = fsub(cx, ax), e2y = fsub(cy, ay), e2z = fsub(cz, az);  // v2v0
	const float rx = fsub(ox, ax), ry = fsub(oy, ay), rz = fsub(oz, az);     // rov0
	const float nx = fsub(fmul(e1y, e2z), fmul(e1z, e2y)), ny = fsub(fmul(e1z, e2x), fmul(e1x, e2z)), nz = fsub(fmul(e1x, e2y), fmul(e1y, e2x));
	const float qx = fsub(fmul(ry, dz), fmul(rz, dy)), qy = fsub(fmul(rz, dx), fmul(rx, dz)), qz = fsub(fmul(rx, dy), fmul(ry, dx));
	const float d = fdiv(1.0f, dot3(dx, dy, dz, nx, ny, nz));
	const float u = fmul(d, -dot3(qx, qy, qz, e2x, e2y, e2z));
	const float v = fmul(d, dot3(qx, qy, qz, e1x, e1y, e1z));
	float tt = fmul(d, -dot3(nx, ny, nz, rx, ry, rz));
	if (u < 0.0f || u > 1.0f || v < 0.0f || fadd(u, v) > 1.0f || tt < 0.0f) tt = FLT_MAX;
	return tt;
}

struct Cand {
	float d;
	int ref;
};
// swap iff a.d < b.d: descending order, the comparator of ngp bvh.cuh:16-27 / triangle_bvh.cu:152-165
__device__ __forceinline__ void cas(Cand& a, Cand& b) {
	if (a.d < b.d) {
		const Cand s = a;
		a = b;
		b = s;
	}
}

constexpr float kMaxDist = 10.0f;  // MAX_DIST, triangle_bvh.cu:49: start value of mint
constexpr int kRayStack = 64;

// returns the BVH-order index of the first triangle the ray hits (-1: none within kMaxDist); t receives mint
__device__ int mesh_ray_intersect(const Quad* __restrict__ quads, const Tri* __restrict__ tris, float ox, float oy, float oz, float dx, float dy, float dz, float& t_out) {
	int stack[kRayStack];
	int sp = 0, cur = 0, best = -1;  // quad 0 = the root's children
	float mint = kMaxDist;
	while (true) {
		if (cur >= 0) {
			const float4* q = reinterpret_cast<const float4*>(quads + cur);
			const float4 mnx = __ldg(q), mny = __ldg(q + 1), mnz = __ldg(q + 2), mxx = __ldg(q + 3), mxy = __ldg(q + 4), mxz = __ldg(q + 5);
			const int4 ref = __ldg(reinterpret_cast<const int4*>(q + 7));
			Cand c[4];
			c[0] = {box_ray_tmin(mnx.x, mny.x, mnz.x, mxx.x, mxy.x, mxz.x, ox, oy, oz, dx, dy, dz), ref.x};
			c[1] = {box_ray_tmin(mnx.y, mny.y, mnz.y, mxx.y, mxy.y, mxz.y, ox, oy, oz, dx, dy, dz), ref.y};
			c[2] = {box_ray_tmin(mnx.z, mny.z, mnz.z, mxx.z, mxy.z, mxz.z, ox, oy, oz, dx, dy, dz), ref.z};
			c[3] = {box_ray_tmin(mnx.w, mny.w, mnz.w, mxx.w, mxy.w, mxz.w, ox, oy, oz, dx, dy, dz), ref.w};
			cas(c[0], c[2]); cas(c[1], c[3]); cas(c[0], c[1]); cas(c[2], c[3]); cas(c[1], c[2]);
#pragma unroll
			for (int k = 0; k < 4; ++k)
				if (c[k].d < mint && sp < kRayStack) stack[sp++] = c[k].ref;
		}
		if (sp == 0) break;
		const int e = stack[--sp];
		if (e >= 0) {
			cur = e;
			continue;
		}
		cur = -1;
		const uint32_t code = (uint32_t)(-(e + 1));
		const int first = (int)(code >> 3), count = (int)(code & 7u);
		for (int i = first; i < first + count; ++i) {
			const float t = tri_ray_t(tris + i, ox, oy, oz, dx, dy, dz);
			if (t < mint) {
				mint = t;
				best = i;
			}
		}
	}
	t_out = mint;
	return best;
}

__global__ void __launch_bounds__(128) mesh_ray_intersect_kernel(uint32_t n, const float* __restrict__ origins, const float* __restrict__ dirs, const Quad* __restrict__ quads,
                                                                 const Tri* __restrict__ tris, int32_t* __restrict__ tri_ids, float* __restrict__ ts) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float t;
	const int b = mesh_ray_intersect(quads, tris, origins[3 * i], origins[3 * i + 1], origins[3 * i + 2], dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], t);
	// the caller gets topology ids (the reference's Triangle::id), like every other entry point
	tri_ids[i] = b >= 0 ? (int32_t)__ldg(&tris[b].id) : -1;
	if (ts) ts[i] = t;
}

// the Van der Corput / Sobol pair of ld_random_pixel_offset (ngp common_device.cuh:232-258, random_val.cuh) is only needed when the caller does not snap to pixel
// centres; the overlay is drawn with snapping (Testbed::m_snap_to_pixel_centers, default true), other values are rejected by the launcher
__device__ __forceinline__ float srgb_to_linear(float x) { return x <= 0.04045f ? x / 12.92f : powf((x + 0.055f) / 1.055f, 2.4f); }

struct OverlayArgs {
	insta_camera cam;
	float aabb[6];
	float center[3];
	int color_normal;
	const Quad* quads;
	const Tri* tris;
	float4* rgba;
	float* depth;
};

__global__ void __launch_bounds__(128) mesh_overlay_kernel(const OverlayArgs a) {
	const uint32_t x = threadIdx.x % 16 + 16 * blockIdx.x, y = threadIdx.x / 16 + 8 * blockIdx.y;
	const uint32_t W = a.cam.resolution[0], H = a.cam.resolution[1];
	if (x >= W || y >= H) return;
	const uint32_t idx = x + W * y;
	// init_rays_with_payload_kernel_sdf, plane_z >= 0, no octree, no envmap: pixel_to_ray (pinhole, pixel centre), normalised, advanced to the box
	const float* m = a.cam.camera_matrix;
	const float u = ((float)x + 0.5f) / (float)W, v = ((float)y + 0.5f) / (float)H;
	const float lx = (u - a.cam.screen_center[0]) * (float)W / a.cam.focal_length[0];
	const float ly = (v - a.cam.screen_center[1]) * (float)H / a.cam.focal_length[1];
	float dx = fmaf(m[0], lx, fmaf(m[3], ly, m[6])), dy = fmaf(m[1], lx, fmaf(m[4], ly, m[7])), dz = fmaf(m[2], lx, fmaf(m[5], ly, m[8]));
	float ox = m[9], oy = m[10], oz = m[11];
	a.depth[idx] = 1e10f;
	const float n2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
	if (n2 > 0.f) {
		const float n = sqrtf(n2);
		dx /= n; dy /= n; dz /= n;
	}
	const float t0 = fmaxf(box_ray_tmin(a.aabb[0], a.aabb[1], a.aabb[2], a.aabb[3], a.aabb[4], a.aabb[5], ox, oy, oz, dx, dy, dz), 0.0f);
	const float adv = t0 + 1e-6f;
	ox = fadd(ox, fmul(adv, dx)); oy = fadd(oy, fmul(adv, dy)); oz = fadd(oz, fmul(adv, dz));
	auto inside = [&](float px, float py, float pz) { return px >= a.aabb[0] && px <= a.aabb[3] && py >= a.aabb[1] && py <= a.aabb[4] && pz >= a.aabb[2] && pz <= a.aabb[5]; };
	if (!inside(ox, oy, oz)) return;  // payload.alive = false.  (The reference traces these pixels too, with whatever direction the payload buffer held.)
	// raytrace_kernel
	float t;
	const int b = mesh_ray_intersect(a.quads, a.tris, ox, oy, oz, dx, dy, dz, t);
	if (b < 0) return;  // positions = orig + 10 dir: outside any render box
	const float px = fadd(ox, fmul(t, dx)), py = fadd(oy, fmul(t, dy)), pz = fadd(oz, fmul(t, dz));
	if (!inside(px, py, pz)) return;  // shade_kernel_sdf: aabb.contains(position)
	// Triangle::normal() = (b - a).cross(c - a).normalized(), normalised once more by the shader
	const Tri tr = a.tris[b];
	const float e1x = tr.b[0] - tr.a[0], e1y = tr.b[1] - tr.a[1], e1z = tr.b[2] - tr.a[2];
	const float e2x = tr.c[0] - tr.a[0], e2y = tr.c[1] - tr.a[1], e2z = tr.c[2] - tr.a[2];
	float nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
	for (int pass = 0; pass < 2; ++pass) {
		const float l2 = nx * nx + ny * ny + nz * nz;
		if (l2 > 0.f) {
			const float l = sqrtf(l2);
			nx /= l; ny /= l; nz /= l;
		}
	}
	float r, g, bl;
	if (a.color_normal) {
		r = srgb_to_linear(0.5f * nx + 0.5f); g = srgb_to_linear(0.5f * ny + 0.5f); bl = srgb_to_linear(0.5f * nz + 0.5f);
	} else {
		// three lights AT center + offset (the reference dots the normal with the light's position, not with a direction towards it)
		const float L[3][3] = {{0.f, 0.f, 1.f}, {1.5f, 0.25f, 0.5f}, {-1.5f, 0.25f, 0.5f}};
		float c = 0.f;
		for (int k = 0; k < 3; ++k) {
			const float d = nx * (a.center[0] + L[k][0]) + ny * (a.center[1] + L[k][1]) + nz * (a.center[2] + L[k][2]);
			c += 1 / 3.f * fminf(1.f, fmaxf(d, 0.f)) * 0.75f;
		}
		r = g = bl = srgb_to_linear(c);
	}
	a.rgba[idx] = make_float4(r, g, bl, 1.0f);
	a.depth[idx] = pz;
}

}  // namespace

int launch_mesh_ray_intersect(insta_ctx* ctx, const Quad* quads, const Tri* tris, const float* d_origins, const float* d_dirs, uint32_t n, int32_t* d_tri_ids, float* d_t,
                              cudaStream_t s) {
	if (n == 0) return INSTA_OK;
	mesh_ray_intersect_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, d_origins, d_dirs, quads, tris, d_tri_ids, d_t);
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

int launch_mesh_overlay(insta_ctx* ctx, const insta_camera* cam, const float* aabb, const float* center, int color_normal, const Quad* quads, const Tri* tris, float* d_rgba,
                        float* d_depth, cudaStream_t s) {
	OverlayArgs a{};
	a.cam = *cam;
	for (int k = 0; k < 6; ++k) a.aabb[k] = aabb[k];
	for (int k = 0; k < 3; ++k) a.center[k] = center[k];
	a.color_normal = color_normal;
	a.quads = quads;
	a.tris = tris;
	a.rgba = (float4*)d_rgba;
	a.depth = d_depth;
	const dim3 grid((cam->resolution[0] + 15) / 16, (cam->resolution[1] + 7) / 8);
	mesh_overlay_kernel<<<grid, 128, 0, s>>>(a);
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

}  // namespace insta
