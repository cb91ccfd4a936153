// Read-back of a rendered frame in the form rta::Recorder writes to disk (src/recorder.cu:299-359): the frame buffer is converted ON THE GPU to
// 8-bit RGBA (save_rgba, recorder.cu:36-60) and integer-millimetre depth (save_depth, recorder.cu:75-105) and leaves the device as 8 bytes per pixel
// with asynchronous copies on the caller's stream -- the reference moves 20 bytes per pixel with two blocking cudaMemcpy and converts on one host
// thread.  Plus the two file writers (host only): a PNG with stored (uncompressed) deflate blocks and the depth .bin.
#include <cstdio>
#include <cstring>
#include <vector>

#include <algorithm>
#include <charconv>
#include <string>
#include <vector>

#include "ctx.h"

namespace insta {
namespace {

// (uint8_t) tcnn::clamp(transform(v) * 255.f, 0.f, 255.f)   (recorder.cu:46-49); transform = identity or srgb_to_linear (Normals mode, 341-343)
__device__ __forceinline__ uint32_t to_u8(float v, int srgb_to_linear) {
	if (srgb_to_linear) v = v <= 0.04045f ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f);
	const float s = __fmul_rn(v, 255.f);
	const float c = s < 0.f ? 0.f : (255.f < s ? 255.f : s);  // tcnn::clamp; a NaN passes through and converts to 0 (as cvttss2si's low byte on the host)
	return (uint32_t)c;
}

__global__ void __launch_bounds__(256) recorder_convert_kernel(const float4* __restrict__ rgba, const float* __restrict__ depth, const uint32_t w, const uint32_t h,
                                                               const int srgb_to_linear, uint32_t* __restrict__ rgba8, uint32_t* __restrict__ depth_mm) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t n = w * h;
	if (i >= n) return;
	if (rgba8) {
		const float4 c = rgba[i];
		rgba8[i] = to_u8(c.x, srgb_to_linear) | (to_u8(c.y, srgb_to_linear) << 8) | (to_u8(c.z, srgb_to_linear) << 16) | (to_u8(c.w, srgb_to_linear) << 24);
	}
	// save_depth reads depth[x + w + y w] for output pixel (x, y) and skips indices past the buffer (recorder.cu:93-94): the file holds rows 1 .. h-1
	if (depth_mm && i + w < n) {
		float d = depth[i + w];
		if (d < 0.f || d > 5.f) d = 0.f;
		depth_mm[i] = (uint32_t)__fmul_rn(d, 1000.0f);
	}
}

const uint32_t* crc_table() {
	static uint32_t t[256];
	static bool init = false;
	if (!init) {
		for (uint32_t n = 0; n < 256; ++n) {
			uint32_t c = n;
			for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
			t[n] = c;
		}
		init = true;
	}
	return t;
}
uint32_t crc32(uint32_t crc, const uint8_t* p, size_t n) {
	const uint32_t* t = crc_table();
	crc = ~crc;
	for (size_t i = 0; i < n; ++i) crc = t[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
	return ~crc;
}
void put32(std::vector<uint8_t>& v, uint32_t x) {
	v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}
void chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& data) {
	put32(out, (uint32_t)data.size());
	const size_t at = out.size();
	out.insert(out.end(), type, type + 4);
	out.insert(out.end(), data.begin(), data.end());
	put32(out, crc32(0, out.data() + at, out.size() - at));
}

}  // namespace

int recorder_readback(insta_ctx* ctx, const float* d_rgba, const float* d_depth, uint32_t w, uint32_t h, int srgb_to_linear, uint8_t* h_rgba8, uint32_t* h_depth_mm,
                      cudaStream_t s) {
	const size_t n = (size_t)w * h;
	if (n > ctx->recorder_cap) {
		INSTA_CUDA(cudaStreamSynchronize(s));
		if (ctx->d_recorder) cudaFree(ctx->d_recorder);
		ctx->d_recorder = nullptr;
		ctx->recorder_cap = 0;
		INSTA_CUDA(cudaMalloc(&ctx->d_recorder, n * 8));
		ctx->recorder_cap = n;
	}
	uint32_t* d_rgba8 = (uint32_t*)ctx->d_recorder;
	uint32_t* d_mm = d_rgba8 + n;
	recorder_convert_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>((const float4*)d_rgba, d_depth, w, h, srgb_to_linear, h_rgba8 ? d_rgba8 : nullptr,
	                                                                     h_depth_mm && d_depth ? d_mm : nullptr);
	ctx->launches++;
	INSTA_CUDA(cudaGetLastError());
	if (h_rgba8) INSTA_CUDA(cudaMemcpyAsync(h_rgba8, d_rgba8, n * 4, cudaMemcpyDeviceToHost, s));
	if (h_depth_mm && d_depth && h > 1) INSTA_CUDA(cudaMemcpyAsync(h_depth_mm, d_mm, (size_t)w * (h - 1) * 4, cudaMemcpyDeviceToHost, s));
	return INSTA_OK;
}

void recorder_destroy(insta_ctx* ctx) {
	if (ctx->d_recorder) cudaFree(ctx->d_recorder);
	ctx->d_recorder = nullptr;
	ctx->recorder_cap = 0;
}

// 8-bit RGBA PNG, filter 0 on every row, zlib stream of stored blocks (stb's writer in the reference compresses; the decoded pixels are what must agree)
int write_png_rgba8(const char* path, uint32_t w, uint32_t h, const uint8_t* rgba8) {
	std::vector<uint8_t> raw;
	raw.reserve((size_t)h * (1 + (size_t)w * 4));
	for (uint32_t y = 0; y < h; ++y) {
		raw.push_back(0);
		raw.insert(raw.end(), rgba8 + (size_t)y * w * 4, rgba8 + (size_t)(y + 1) * w * 4);
	}
	std::vector<uint8_t> z;
	z.push_back(0x78); z.push_back(0x01);
	uint32_t a = 1, b = 0;
	for (size_t pos = 0; pos < raw.size() || pos == 0;) {
		const size_t len = std::min<size_t>(65535, raw.size() - pos);
		z.push_back(pos + len >= raw.size() ? 1 : 0);
		z.push_back((uint8_t)(len & 0xff)); z.push_back((uint8_t)(len >> 8));
		z.push_back((uint8_t)(~len & 0xff)); z.push_back((uint8_t)((~len >> 8) & 0xff));
		z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + len);
		for (size_t i = pos; i < pos + len; ++i) {
			a = (a + raw[i]) % 65521u;
			b = (b + a) % 65521u;
		}
		pos += len;
		if (len == 0) break;
	}
	put32(z, (b << 16) | a);
	std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
	std::vector<uint8_t> ihdr;
	put32(ihdr, w); put32(ihdr, h);
	ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
	chunk(out, "IHDR", ihdr);
	chunk(out, "IDAT", z);
	chunk(out, "IEND", {});
	FILE* f = fopen(path, "wb");
	if (!f) INSTA_FAIL(INSTA_ERR_IO, std::string("cannot write ") + path);
	const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
	fclose(f);
	if (!ok) INSTA_FAIL(INSTA_ERR_IO, std::string("short write to ") + path);
	return INSTA_OK;
}

int write_depth_bin(const char* path, const uint32_t* depth_mm, size_t n) {
	FILE* f = fopen(path, "wb");
	if (!f) INSTA_FAIL(INSTA_ERR_IO, std::string("cannot write ") + path);
	const bool ok = fwrite(depth_mm, 4, n, f) == n;
	fclose(f);
	if (!ok) INSTA_FAIL(INSTA_ERR_IO, std::string("short write to ") + path);
	return INSTA_OK;
}

// ---- camera trajectory (Recorder::dump_cameras_json, recorder.cu:648-695) ------------------------------------------------------------------------
namespace {
// nlohmann::json prints a float (stored as double) with the shortest digits that round-trip the DOUBLE, "x.0" for integral values
std::string json_number(float v) {
	char buf[64];
	const auto r = std::to_chars(buf, buf + sizeof(buf), (double)v);
	std::string s(buf, r.ptr);
	if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
	return s;
}
std::string json_string(const std::string& v) {
	std::string s = "\"";
	for (char c : v) {
		if (c == '"' || c == '\\') s += '\\';
		s += c;
	}
	return s + "\"";
}
}  // namespace

int write_trajectory_json(const char* path, uint32_t w, uint32_t h, float fl_x, float fl_y, float cx, float cy, const insta_trajectory_frame* frames, uint32_t n) {
	struct Row {
		std::string id, mesh;
		float m[16];
	};
	std::vector<Row> rows(n);
	for (uint32_t i = 0; i < n; ++i) {
		const std::string idx = std::to_string(frames[i].frame_index);
		rows[i].id = std::string(5 - std::min<size_t>(5, idx.size()), '0') + idx + (frames[i].suffix ? frames[i].suffix : "");
		rows[i].mesh = "meshes/" + rows[i].id + ".obj";
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 4; ++c) rows[i].m[4 * r + c] = frames[i].camera_matrix[4 * r + c] - (c == 3 ? 0.5f : 0.0f);
		rows[i].m[12] = rows[i].m[13] = rows[i].m[14] = 0.0f;
		rows[i].m[15] = 1.0f;
	}
	std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.mesh < b.mesh; });
	std::string o = "{\n";
	const std::string i1(4, ' '), i2(8, ' '), i3(12, ' '), i4(16, ' '), i5(20, ' ');
	o += i1 + "\"cx\": " + json_number(cx) + ",\n" + i1 + "\"cy\": " + json_number(cy) + ",\n";
	o += i1 + "\"fl_x\": " + json_number(fl_x) + ",\n" + i1 + "\"fl_y\": " + json_number(fl_y) + ",\n";
	o += i1 + "\"frames\": [";
	for (uint32_t i = 0; i < n; ++i) {
		const Row& R = rows[i];
		o += (i ? ",\n" : "\n") + i2 + "{\n";
		o += i3 + "\"depth_path\": \"\",\n";
		o += i3 + "\"exp_path\": " + json_string("flame/exp/" + R.id + ".txt") + ",\n";
		o += i3 + "\"eyes_path\": " + json_string("flame/eyes/" + R.id + ".txt") + ",\n";
		o += i3 + "\"file_path\": \"\",\n";
		o += i3 + "\"mesh_path\": " + json_string(R.mesh) + ",\n";
		o += i3 + "\"seg_mask_path\": \"\",\n";
		o += i3 + "\"transform_matrix\": [\n";
		for (int r = 0; r < 4; ++r) {
			o += i4 + "[\n";
			for (int c = 0; c < 4; ++c) o += i5 + json_number(R.m[4 * r + c]) + (c < 3 ? ",\n" : "\n");
			o += i4 + (r < 3 ? "],\n" : "]\n");
		}
		o += i3 + "]\n" + i2 + "}";
	}
	o += n ? "\n" + i1 + "],\n" : "],\n";
	o += i1 + "\"h\": " + std::to_string(h) + ",\n";
	o += i1 + "\"integer_depth_scale\": " + json_number(0.001f) + ",\n";
	o += i1 + "\"synthetic\": true,\n";
	o += i1 + "\"w\": " + std::to_string(w) + "\n}\n";
	FILE* f = fopen(path, "wb");
	if (!f) INSTA_FAIL(INSTA_ERR_IO, std::string("cannot write ") + path);
	const bool ok = fwrite(o.data(), 1, o.size(), f) == o.size();
	fclose(f);
	if (!ok) INSTA_FAIL(INSTA_ERR_IO, std::string("short write to ") + path);
	return INSTA_OK;
}

}  // namespace insta
