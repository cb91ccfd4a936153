// Snapshot files in the reference's layout: ngp::Testbed::save_snapshot / load_snapshot (ngp/src/testbed.cu:2916-3016) write the network
// config as MessagePack (nlohmann::json::to_msgpack) with a "snapshot" object holding tcnn::Trainer::serialize (tcnn trainer.h:255-279:
// n_params, params_binary = the fp16 INFERENCE weights, i.e. the EMA weights, in trainer order; optionally the optimizer state of
// EmaOptimizer{AdamOptimizer}, tcnn optimizers/ema.h:182-194, adam.h:271-292), density_grid_binary (fp16), the training counters and boxes.
// This file is host-only: a MessagePack reader for everything the reference may have written and a writer for the keys above.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ctx.h"

namespace insta {
namespace {

// ---- writer -----------------------------------------------------------------------------------------------------------------------------
struct Writer {
	std::vector<uint8_t> out;
	void u8(uint8_t v) { out.push_back(v); }
	void be(uint64_t v, int bytes) {
		for (int k = bytes - 1; k >= 0; --k) out.push_back((uint8_t)(v >> (8 * k)));
	}
	void map(uint32_t n) {
		if (n < 16) u8(0x80 | n);
		else if (n < 65536) { u8(0xde); be(n, 2); }
		else { u8(0xdf); be(n, 4); }
	}
	void array(uint32_t n) {
		if (n < 16) u8(0x90 | n);
		else if (n < 65536) { u8(0xdc); be(n, 2); }
		else { u8(0xdd); be(n, 4); }
	}
	void str(const char* s) {
		const size_t n = std::strlen(s);
		if (n < 32) u8(0xa0 | (uint8_t)n);
		else if (n < 256) { u8(0xd9); u8((uint8_t)n); }
		else { u8(0xda); be(n, 2); }
		out.insert(out.end(), s, s + n);
	}
	void uint(uint64_t v) {
		if (v < 128) u8((uint8_t)v);
		else if (v < 256) { u8(0xcc); u8((uint8_t)v); }
		else if (v < 65536) { u8(0xcd); be(v, 2); }
		else if (v <= 0xffffffffull) { u8(0xce); be(v, 4); }
		else { u8(0xcf); be(v, 8); }
	}
	void boolean(bool b) { u8(b ? 0xc3 : 0xc2); }
	// nlohmann writes a double as float32 when that is lossless, else float64
	void real(double d) {
		const float f = (float)d;
		if ((double)f == d) {
			uint32_t bits;
			std::memcpy(&bits, &f, 4);
			u8(0xca); be(bits, 4);
		} else {
			uint64_t bits;
			std::memcpy(&bits, &d, 8);
			u8(0xcb); be(bits, 8);
		}
	}
	void bin(const void* p, uint64_t n) {
		if (n < 256) { u8(0xc4); u8((uint8_t)n); }
		else if (n < 65536) { u8(0xc5); be(n, 2); }
		else { u8(0xc6); be(n, 4); }
		const uint8_t* b = (const uint8_t*)p;
		out.insert(out.end(), b, b + n);
	}
	void kv_str(const char* k, const char* v) { str(k); str(v); }
	void kv_uint(const char* k, uint64_t v) { str(k); uint(v); }
	void kv_real(const char* k, double v) { str(k); real(v); }
	void kv_bool(const char* k, bool v) { str(k); boolean(v); }
	void vec3(const float* v) { array(3); for (int i = 0; i < 3; ++i) real(v[i]); }
	void box(const char* k, const float* b6) {  // to_json(BoundingBox), ngp json_binding.h:73-76 (keys sorted: max, min)
		str(k); map(2); str("max"); vec3(b6 + 3); str("min"); vec3(b6);
	}
};

// ---- reader -----------------------------------------------------------------------------------------------------------------------------
struct Value {
	enum Kind { NIL, BOOL, INT, REAL, STR, BIN, ARRAY, MAP } kind = NIL;
	int64_t i = 0;
	double d = 0;
	const uint8_t* p = nullptr;  // STR / BIN payload inside the file buffer
	uint64_t n = 0;
	std::vector<Value> items;    // ARRAY elements; MAP: key, value, key, value ...
	const Value* get(const char* key) const {
		if (kind != MAP) return nullptr;
		const size_t len = std::strlen(key);
		for (size_t k = 0; k + 1 < items.size(); k += 2)
			if (items[k].kind == STR && items[k].n == len && std::memcmp(items[k].p, key, len) == 0) return &items[k + 1];
		return nullptr;
	}
	double number(double dflt) const { return kind == INT ? (double)i : (kind == REAL ? d : (kind == BOOL ? (double)i : dflt)); }
};

struct Reader {
	const uint8_t* p;
	const uint8_t* end;
	bool ok = true;
	uint64_t be(int bytes) {
		if (end - p < bytes) { ok = false; return 0; }
		uint64_t v = 0;
		for (int k = 0; k < bytes; ++k) v = (v << 8) | *p++;
		return v;
	}
	void blob(Value& v, Value::Kind kind, uint64_t n) {
		if ((uint64_t)(end - p) < n) { ok = false; return; }
		v.kind = kind; v.p = p; v.n = n; p += n;
	}
	void seq(Value& v, Value::Kind kind, uint64_t n, int depth) {
		v.kind = kind;
		const uint64_t count = kind == Value::MAP ? 2 * n : n;
		if (count > (uint64_t)(end - p)) { ok = false; return; }
		v.items.resize(count);
		for (uint64_t k = 0; k < count && ok; ++k) parse(v.items[k], depth + 1);
	}
	void parse(Value& v, int depth = 0) {
		if (!ok || p >= end || depth > 64) { ok = false; return; }
		const uint8_t t = *p++;
		if (t < 0x80) { v.kind = Value::INT; v.i = t; }
		else if (t < 0x90) seq(v, Value::MAP, t & 0x0f, depth);
		else if (t < 0xa0) seq(v, Value::ARRAY, t & 0x0f, depth);
		else if (t < 0xc0) blob(v, Value::STR, t & 0x1f);
		else if (t >= 0xe0) { v.kind = Value::INT; v.i = (int8_t)t; }
		else switch (t) {
			case 0xc0: v.kind = Value::NIL; break;
			case 0xc2: v.kind = Value::BOOL; v.i = 0; break;
			case 0xc3: v.kind = Value::BOOL; v.i = 1; break;
			case 0xc4: blob(v, Value::BIN, be(1)); break;
			case 0xc5: blob(v, Value::BIN, be(2)); break;
			case 0xc6: blob(v, Value::BIN, be(4)); break;
			case 0xc7: { const uint64_t n = be(1); be(1); blob(v, Value::BIN, n); break; }  // ext 8 (nlohmann: binary with subtype)
			case 0xc8: { const uint64_t n = be(2); be(1); blob(v, Value::BIN, n); break; }
			case 0xc9: { const uint64_t n = be(4); be(1); blob(v, Value::BIN, n); break; }
			case 0xca: { const uint32_t b = (uint32_t)be(4); float f; std::memcpy(&f, &b, 4); v.kind = Value::REAL; v.d = f; break; }
			case 0xcb: { const uint64_t b = be(8); std::memcpy(&v.d, &b, 8); v.kind = Value::REAL; break; }
			case 0xcc: v.kind = Value::INT; v.i = (int64_t)be(1); break;
			case 0xcd: v.kind = Value::INT; v.i = (int64_t)be(2); break;
			case 0xce: v.kind = Value::INT; v.i = (int64_t)be(4); break;
			case 0xcf: v.kind = Value::INT; v.i = (int64_t)be(8); break;
			case 0xd0: v.kind = Value::INT; v.i = (int8_t)be(1); break;
			case 0xd1: v.kind = Value::INT; v.i = (int16_t)be(2); break;
			case 0xd2: v.kind = Value::INT; v.i = (int32_t)be(4); break;
			case 0xd3: v.kind = Value::INT; v.i = (int64_t)be(8); break;
			case 0xd4: be(1); blob(v, Value::BIN, 1); break;   // fixext
			case 0xd5: be(1); blob(v, Value::BIN, 2); break;
			case 0xd6: be(1); blob(v, Value::BIN, 4); break;
			case 0xd7: be(1); blob(v, Value::BIN, 8); break;
			case 0xd8: be(1); blob(v, Value::BIN, 16); break;
			case 0xd9: blob(v, Value::STR, be(1)); break;
			case 0xda: blob(v, Value::STR, be(2)); break;
			case 0xdb: blob(v, Value::STR, be(4)); break;
			case 0xdc: seq(v, Value::ARRAY, be(2), depth); break;
			case 0xdd: seq(v, Value::ARRAY, be(4), depth); break;
			case 0xde: seq(v, Value::MAP, be(2), depth); break;
			case 0xdf: seq(v, Value::MAP, be(4), depth); break;
			default: ok = false;
		}
	}
};

void read_box(const Value* v, float* b6) {
	if (!v) return;
	const Value* mn = v->get("min");
	const Value* mx = v->get("max");
	for (int k = 0; k < 3; ++k) {
		if (mn && mn->kind == Value::ARRAY && mn->items.size() == 3) b6[k] = (float)mn->items[k].number(b6[k]);
		if (mx && mx->kind == Value::ARRAY && mx->items.size() == 3) b6[3 + k] = (float)mx->items[k].number(b6[3 + k]);
	}
}
}  // namespace

struct SnapshotFile {
	std::vector<uint8_t> bytes;
	Value root;
};

int snapshot_save(const char* path, const insta_params* p, const insta_snapshot* s) {
	Writer w;
	w.out.reserve((size_t)s->n_params * 2 + (size_t)s->n_density_grid * 2 + (s->has_optimizer ? (size_t)s->n_params * 14 : 0) + 4096);
	// The network config the reference would have loaded -- EVERY key of configs/nerf/main.json, in nlohmann's (sorted) order: Testbed::reset_network reads
	// "nerf_network" as a string without a guard (testbed.cu:2285), rta::Core reads seg_weight / penalize_density_alpha / beta_loss / depth_weight /
	// geo_params / exp_params / use_eyes (src/core.cu:187-213) and the recorder max_steps (src/recorder.cu:63); a file without them makes the reference's
	// loader throw a json type_error.  Values: insta_params where it carries the key, main.json's otherwise.
	w.map(24);
	w.kv_uint("beta_loss", 0);
	w.kv_real("depth_weight", 1.25);
	w.str("dir_encoding"); w.map(2);
	w.str("nested"); w.array(2);
	w.map(3); w.kv_uint("degree", 4); w.kv_uint("n_dims_to_encode", 3); w.kv_str("otype", "SphericalHarmonics");
	w.map(3); w.kv_uint("degree", 4); w.kv_uint("n_bins", 4); w.kv_str("otype", "Identity");
	w.kv_str("otype", "Composite");
	w.kv_bool("disable_interp", p->disable_interp != 0);
	w.str("encoding"); w.map(5);
	w.kv_uint("base_resolution", p->base_resolution); w.kv_uint("log2_hashmap_size", p->log2_hashmap_size); w.kv_uint("n_features_per_level", p->n_features_per_level);
	w.kv_uint("n_levels", p->n_levels); w.kv_str("otype", "HashGrid");
	w.kv_uint("exp_params", p->exp_params);
	w.kv_uint("geo_params", p->geo_params);
	w.str("loss"); w.map(1); w.kv_str("otype", "Huber");
	w.kv_uint("max_cached_bvh", p->max_cached_bvh);
	w.kv_uint("max_images_gpu", 1700);
	w.kv_uint("max_steps", 33000);
	w.kv_str("nerf_network", "concatenation");
	w.str("network"); w.map(6);
	w.kv_str("activation", "ReLU"); w.kv_uint("n_hidden_layers", 2); w.kv_uint("n_neurons", 64); w.kv_uint("n_output_dims", 16); w.kv_str("otype", "FullyFusedMLP");
	w.kv_str("output_activation", "None");
	w.kv_bool("optimize_latent_code", false);
	w.str("optimizer"); w.map(3);
	w.kv_real("decay", (double)s->ema_decay);
	w.str("nested"); w.map(6);
	w.kv_real("beta1", (double)s->beta1); w.kv_real("beta2", (double)s->beta2); w.kv_real("epsilon", (double)s->epsilon); w.kv_real("l2_reg", (double)s->l2_reg);
	w.kv_real("learning_rate", (double)s->opt_base_learning_rate); w.kv_str("otype", "Adam");
	w.kv_str("otype", "Ema");
	w.kv_bool("penalize_density_alpha", false);
	w.kv_bool("render_from_snapshot", false);
	w.kv_bool("render_novel_trajectory", false);
	w.str("rgb_network"); w.map(5);
	w.kv_str("activation", "ReLU"); w.kv_uint("n_hidden_layers", 3); w.kv_uint("n_neurons", 64); w.kv_str("otype", "FullyFusedMLP"); w.kv_str("output_activation", "None");
	w.kv_uint("seg_weight", 40);
	// ---- "snapshot" ----
	w.str("snapshot");
	w.map(11 + (s->has_optimizer ? 1 : 0));
	w.box("aabb", s->aabb);
	w.kv_real("bounding_radius", (double)s->bounding_radius);
	w.str("density_grid_binary"); w.bin(s->h_density_grid_f16, s->n_density_grid * 2);
	w.kv_uint("density_grid_size", 128);
	w.kv_real("loss", (double)s->loss);
	w.kv_uint("n_params", s->n_params);
	w.str("nerf"); w.map(2);
	w.kv_uint("aabb_scale", s->aabb_scale);
	w.str("rgb"); w.map(3);
	w.kv_uint("measured_batch_size", s->measured_batch_size); w.kv_uint("measured_batch_size_before_compaction", s->measured_batch_size_before_compaction);
	w.kv_uint("rays_per_batch", s->rays_per_batch);
	if (s->has_optimizer) {
		w.str("optimizer"); w.map(2);
		w.str("nested"); w.map(5);
		w.kv_real("base_learning_rate", (double)s->opt_base_learning_rate);
		w.kv_uint("current_step", s->opt_current_step);
		w.str("first_moments_binary"); w.bin(s->h_first_moments, s->n_params * 4);
		w.str("param_steps_binary"); w.bin(s->h_param_steps, s->n_params * 4);
		w.str("second_moments_binary"); w.bin(s->h_second_moments, s->n_params * 4);
		w.str("weights_ema_binary"); w.bin(s->h_weights_ema_f16, s->n_params * 2);
	}
	w.str("params_binary"); w.bin(s->h_params_f16, s->n_params * 2);
	w.box("render_aabb", s->render_aabb);
	w.kv_uint("training_step", s->training_step);
	w.kv_uint("version", 1);
	w.kv_bool("use_dataset_cache", true);
	w.kv_bool("use_eyes", false);
	w.kv_bool("use_ngp", false);
	FILE* f = std::fopen(path, "wb");
	if (!f) INSTA_FAIL(INSTA_ERR_IO, std::string("insta_snapshot_save: cannot open ") + path);
	const size_t written = std::fwrite(w.out.data(), 1, w.out.size(), f);
	std::fclose(f);
	if (written != w.out.size()) INSTA_FAIL(INSTA_ERR_IO, std::string("insta_snapshot_save: short write to ") + path);
	return INSTA_OK;
}

int snapshot_open(const char* path, SnapshotFile** out_file, insta_snapshot* s) {
	FILE* f = std::fopen(path, "rb");
	if (!f) INSTA_FAIL(INSTA_ERR_IO, std::string("insta_snapshot_open: cannot open ") + path);
	std::unique_ptr<SnapshotFile> sf(new SnapshotFile());
	std::fseek(f, 0, SEEK_END);
	const long size = std::ftell(f);
	std::fseek(f, 0, SEEK_SET);
	sf->bytes.resize(size > 0 ? (size_t)size : 0);
	const size_t got = sf->bytes.empty() ? 0 : std::fread(sf->bytes.data(), 1, sf->bytes.size(), f);
	std::fclose(f);
	if (got != sf->bytes.size() || got == 0) INSTA_FAIL(INSTA_ERR_IO, std::string("insta_snapshot_open: cannot read ") + path);
	Reader r{sf->bytes.data(), sf->bytes.data() + sf->bytes.size()};
	r.parse(sf->root);
	if (!r.ok) INSTA_FAIL(INSTA_ERR_IO, std::string("insta_snapshot_open: not a MessagePack document: ") + path);
	const Value* snap = sf->root.get("snapshot");
	if (!snap || snap->kind != Value::MAP) INSTA_FAIL(INSTA_ERR_IO, std::string("insta_snapshot_open: file does not contain a snapshot: ") + path);
	const Value* ver = snap->get("version");
	if (!ver || ver->number(0) < 1) INSTA_FAIL(INSTA_ERR_IO, "insta_snapshot_open: snapshot uses an old format");
	if (const Value* gs = snap->get("density_grid_size"))
		if (gs->number(128) != 128) INSTA_FAIL(INSTA_ERR_IO, "insta_snapshot_open: incompatible grid size");
	std::memset(s, 0, sizeof(*s));
	const float unit[6] = {0, 0, 0, 1, 1, 1};
	std::memcpy(s->aabb, unit, sizeof(unit));
	std::memcpy(s->render_aabb, unit, sizeof(unit));
	const Value* pb = snap->get("params_binary");
	if (!pb || pb->kind != Value::BIN) INSTA_FAIL(INSTA_ERR_IO, "insta_snapshot_open: params_binary missing");
	s->h_params_f16 = pb->p;
	s->n_params = pb->n / 2;
	if (const Value* np = snap->get("n_params"))
		if ((uint64_t)np->number(0) != s->n_params) INSTA_FAIL(INSTA_ERR_IO, "insta_snapshot_open: n_params does not match params_binary");
	if (const Value* dg = snap->get("density_grid_binary")) {
		if (dg->kind != Value::BIN) INSTA_FAIL(INSTA_ERR_IO, "insta_snapshot_open: density_grid_binary is not binary");
		s->h_density_grid_f16 = dg->p;
		s->n_density_grid = dg->n / 2;
	}
	if (const Value* v = snap->get("training_step")) s->training_step = (uint32_t)v->number(0);
	if (const Value* v = snap->get("loss")) s->loss = (float)v->number(0);
	if (const Value* v = snap->get("bounding_radius")) s->bounding_radius = (float)v->number(0);
	read_box(snap->get("aabb"), s->aabb);
	read_box(snap->get("render_aabb"), s->render_aabb);
	s->aabb_scale = 1;
	if (const Value* nerf = snap->get("nerf")) {
		if (const Value* v = nerf->get("aabb_scale")) s->aabb_scale = (uint32_t)v->number(1);
		if (const Value* rgb = nerf->get("rgb")) {
			if (const Value* v = rgb->get("rays_per_batch")) s->rays_per_batch = (uint32_t)v->number(0);
			if (const Value* v = rgb->get("measured_batch_size")) s->measured_batch_size = (uint32_t)v->number(0);
			if (const Value* v = rgb->get("measured_batch_size_before_compaction")) s->measured_batch_size_before_compaction = (uint32_t)v->number(0);
		}
	}
	// optimizer hyper-parameters from the embedded network config, state from snapshot.optimizer
	s->ema_decay = 0.95f; s->beta1 = 0.9f; s->beta2 = 0.99f; s->epsilon = 1e-15f; s->l2_reg = 1e-6f; s->opt_base_learning_rate = 1.75e-3f;
	if (const Value* oc = sf->root.get("optimizer")) {
		if (const Value* v = oc->get("decay")) s->ema_decay = (float)v->number(s->ema_decay);
		if (const Value* n = oc->get("nested")) {
			if (const Value* v = n->get("beta1")) s->beta1 = (float)v->number(s->beta1);
			if (const Value* v = n->get("beta2")) s->beta2 = (float)v->number(s->beta2);
			if (const Value* v = n->get("epsilon")) s->epsilon = (float)v->number(s->epsilon);
			if (const Value* v = n->get("l2_reg")) s->l2_reg = (float)v->number(s->l2_reg);
			if (const Value* v = n->get("learning_rate")) s->opt_base_learning_rate = (float)v->number(s->opt_base_learning_rate);
		}
	}
	if (const Value* o = snap->get("optimizer")) {
		const Value* n = o->get("nested");
		const Value* ema = o->get("weights_ema_binary");
		const Value* m1 = n ? n->get("first_moments_binary") : nullptr;
		const Value* m2 = n ? n->get("second_moments_binary") : nullptr;
		const Value* ps = n ? n->get("param_steps_binary") : nullptr;
		if (ema && m1 && m2 && ema->kind == Value::BIN && m1->kind == Value::BIN && m2->kind == Value::BIN && ema->n == s->n_params * 2 && m1->n == s->n_params * 4 &&
		    m2->n == s->n_params * 4) {
			s->has_optimizer = 1;
			s->h_weights_ema_f16 = ema->p;
			s->h_first_moments = (const float*)m1->p;
			s->h_second_moments = (const float*)m2->p;
			if (ps && ps->kind == Value::BIN && ps->n == s->n_params * 4) s->h_param_steps = (const uint32_t*)ps->p;  // absent in old files: all zero (adam.h:284-289)
			if (const Value* v = n->get("current_step")) s->opt_current_step = (uint32_t)v->number(0);
			if (const Value* v = n->get("base_learning_rate")) s->opt_base_learning_rate = (float)v->number(s->opt_base_learning_rate);
		}
	}
	*out_file = sf.release();
	return INSTA_OK;
}

void snapshot_close(SnapshotFile* f) { delete f; }

}  // namespace insta
