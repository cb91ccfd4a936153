// TEST INFRASTRUCTURE ONLY.  Pins the snapshot layout (insta_b200/csrc/snapshot.cpp) against the REFERENCE's own serializer: this file is compiled
// against the nlohmann::json header the reference vendors (tcnn/dependencies/json/json.hpp, header-only) and restates, key access by key access and
// with the same C++ target types, what the reference does with a snapshot file:
//   refjson_load_like_reference : load_network_config (from_msgpack) -> Testbed::load_snapshot (ngp/src/testbed.cu:2951-3016) -> reset_network's reads of
//                                 the embedded config (testbed.cu:2207-2285, 2312-2324) -> rta::Core::load_training_data / post_rendering (src/core.cu:
//                                 187-213) -> Recorder (src/recorder.cu:63) -> tcnn Trainer / Ema / Adam ::deserialize (trainer.h:267-279, ema.h:189-194,
//                                 adam.h:281-292).  Any missing key or wrong type throws exactly as it would in the reference; the message comes back.
//   refjson_save_like_reference : Testbed::save_snapshot (testbed.cu:2916-2949) over configs/nerf/main.json (text passed in) with tcnn's serialize()
//                                 (trainer.h:254-265, ema.h:182-187, adam.h:271-279) and json::to_msgpack -- the bytes our reader must accept.
// Built by oracle/Makefile (target refjson) into oracle/_ref/libref_json.so; used by tests/test_snapshot_ref.py and tests/golden/make_golden_snapshot.py.
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

#include <json/json.hpp>

using json = nlohmann::json;

namespace {
thread_local std::string g_msg;

// ngp json_binding.h:73-76 (BoundingBox) with Eigen vectors as arrays (json_binding.h:33-60)
void box_from_json(const json& j, float* b6) {
	for (int k = 0; k < 3; ++k) {
		b6[k] = j.at("min").at(k).get<float>();
		b6[3 + k] = j.at("max").at(k).get<float>();
	}
}
json box_to_json(const float* b6) {
	json j;
	j["min"] = {b6[0], b6[1], b6[2]};
	j["max"] = {b6[3], b6[4], b6[5]};
	return j;
}
// tcnn gpu_memory_json.h:53-70: a GPUMemory<T> comes out of a binary value (or, for old files, an array of numbers)
size_t binary_elements(const json& j, size_t elem_size) {
	if (j.is_binary()) {
		const json::binary_t& b = j;
		return b.size() / elem_size;
	}
	if (j.is_object() && j.contains("bytes")) {
		const json::array_t& a = j["bytes"];
		return a.size() / elem_size;
	}
	throw std::runtime_error("Invalid json type: must be either binary or object");
}
json to_binary(const void* p, size_t n_bytes) {
	json::binary_t b;
	b.resize(n_bytes);
	if (n_bytes) std::memcpy(b.data(), p, n_bytes);
	return b;
}
}  // namespace

extern "C" {

const char* refjson_last_message() { return g_msg.c_str(); }

// out8: n_params, n_density_grid, training_step + 1, rays_per_batch, has optimizer state, current_step, exp_params, geo_params
int refjson_load_like_reference(const char* path, uint64_t* out8) {
	try {
		std::ifstream f(path, std::ios::in | std::ios::binary);
		if (!f) throw std::runtime_error("cannot open file");
		json config = json::from_msgpack(f);  // load_network_config, testbed.cu:155-159
		if (!config.contains("snapshot")) throw std::runtime_error("File does not contain a snapshot.");
		const auto& snapshot = config["snapshot"];
		if (snapshot.value("version", 0) < 1) throw std::runtime_error("Snapshot uses an old format.");
		float aabb[6] = {0, 0, 0, 1, 1, 1}, render_aabb[6] = {0, 0, 0, 1, 1, 1};
		if (snapshot.contains("aabb")) box_from_json(snapshot["aabb"], aabb);
		if (snapshot.contains("render_aabb")) box_from_json(snapshot["render_aabb"], render_aabb);
		const float bounding_radius = snapshot.value("bounding_radius", 1.0f);
		(void)bounding_radius;
		if (snapshot["density_grid_size"] != 128) throw std::runtime_error("Incompatible grid size.");
		const uint32_t rays_per_batch = snapshot["nerf"]["rgb"]["rays_per_batch"];
		const uint32_t measured_batch_size = snapshot["nerf"]["rgb"]["measured_batch_size"];
		const uint32_t measured_batch_size_before_compaction = snapshot["nerf"]["rgb"]["measured_batch_size_before_compaction"];
		(void)measured_batch_size; (void)measured_batch_size_before_compaction;
		if (!snapshot["nerf"].contains("dataset") && snapshot["nerf"].contains("aabb_scale")) {
			const int aabb_scale = snapshot["nerf"]["aabb_scale"];
			(void)aabb_scale;
		}
		const size_t n_grid = binary_elements(snapshot["density_grid_binary"], 2);
		if (n_grid != 0 && n_grid % (128ull * 128 * 128) != 0) throw std::runtime_error("Incompatible number of grid cascades.");
		// reset_network(false): the embedded config
		json& encoding_config = config["encoding"];
		json& loss_config = config["loss"];
		json& optimizer_config = config["optimizer"];
		json& network_config = config["network"];
		const uint32_t n_features_per_level = encoding_config.value("n_features_per_level", 2u);
		const uint32_t n_levels = encoding_config.value("n_levels", 16u);
		const uint32_t log2_hashmap_size = encoding_config.value("log2_hashmap_size", 15);
		const uint32_t base_resolution = encoding_config.value("base_resolution", 0);
		(void)n_features_per_level; (void)n_levels; (void)log2_hashmap_size; (void)base_resolution;
		json& dir_encoding_config = config["dir_encoding"];
		json& rgb_network_config = config["rgb_network"];
		std::string nerf_network_type = config["nerf_network"];
		std::string s1 = std::string(encoding_config["otype"]), s2 = std::string(network_config["otype"]), s3 = std::string(dir_encoding_config["otype"]),
		            s4 = std::string(rgb_network_config["otype"]), s5 = std::string(loss_config["otype"]), s6 = std::string(optimizer_config["otype"]);
		const int n1 = (int)network_config["n_neurons"], l1 = (int)network_config["n_hidden_layers"], n2 = (int)rgb_network_config["n_neurons"],
		          l2 = (int)rgb_network_config["n_hidden_layers"];
		(void)n1; (void)l1; (void)n2; (void)l2;
		if (nerf_network_type.empty() || s1.empty() || s2.empty() || s3.empty() || s4.empty() || s5.empty() || s6.empty()) throw std::runtime_error("empty otype");
		// tcnn create_optimizer / update_hyperparams (ema.h, adam.h): value() with defaults, nested object required by Ema
		const float decay = optimizer_config.value("decay", 0.99f);
		const json& nested = optimizer_config.at("nested");
		const float lr = nested.value("learning_rate", 1e-3f), b1 = nested.value("beta1", 0.9f), b2 = nested.value("beta2", 0.999f), eps = nested.value("epsilon", 1e-8f),
		            l2r = nested.value("l2_reg", 1e-8f);
		(void)decay; (void)lr; (void)b1; (void)b2; (void)eps; (void)l2r;
		// rta::Core::load_training_data (core.cu:200-213), post_rendering (187-196), Recorder (recorder.cu:63)
		if (config.contains("disable_interp")) { bool v = config["disable_interp"]; (void)v; }
		if (config.contains("optimize_latent_code")) { bool v = config["optimize_latent_code"]; (void)v; }
		if (config.contains("max_cached_bvh")) { uint32_t v = config["max_cached_bvh"]; (void)v; }
		if (config.contains("max_images_gpu")) { int v = config["max_images_gpu"]; (void)v; }
		if (config.contains("use_dataset_cache")) { bool v = config["use_dataset_cache"]; (void)v; }
		const float depth_weight = config["depth_weight"];
		const uint32_t geo_params = config["geo_params"], exp_params = config["exp_params"];
		const bool use_eyes = config["use_eyes"];
		const float seg_weight = config["seg_weight"];
		const bool penalize_density_alpha = config["penalize_density_alpha"];
		const float beta_loss = config["beta_loss"];
		const uint32_t max_steps = config["max_steps"];
		(void)depth_weight; (void)use_eyes; (void)seg_weight; (void)penalize_density_alpha; (void)beta_loss; (void)max_steps;
		const uint32_t training_step = uint32_t(config["snapshot"]["training_step"]) + 1;
		const float loss = config["snapshot"]["loss"];
		(void)loss;
		// Trainer::deserialize
		const json& data = config["snapshot"];
		const size_t n_params = binary_elements(data["params_binary"], 2);
		uint64_t has_opt = 0, current_step = 0;
		if (data.contains("optimizer")) {
			const json& o = data["optimizer"];
			const size_t n_ema = binary_elements(o["weights_ema_binary"], 2);
			const json& n = o["nested"];
			const size_t n_m1 = binary_elements(n["first_moments_binary"], 4), n_m2 = binary_elements(n["second_moments_binary"], 4);
			if (n.contains("param_steps_binary") && binary_elements(n["param_steps_binary"], 4) != n_m2) throw std::runtime_error("param_steps size");
			current_step = (uint32_t)n["current_step"];
			const float base_lr = n["base_learning_rate"];
			(void)base_lr;
			if (n_ema != n_params || n_m1 != n_params || n_m2 != n_params) throw std::runtime_error("optimizer state size does not match n_params");
			has_opt = 1;
		}
		out8[0] = n_params; out8[1] = n_grid; out8[2] = training_step; out8[3] = rays_per_batch; out8[4] = has_opt; out8[5] = current_step; out8[6] = exp_params;
		out8[7] = geo_params;
		return 0;
	} catch (const std::exception& e) {
		g_msg = e.what();
		return 1;
	}
}

// Recorder::dump_cameras_json (src/recorder.cu:648-695) over JsonFrame entries built as dump_frame_buffer builds them (recorder.cu:318-343).
// cams: n x 12 floats (3x4 row-major m_camera); ids: n frame indices; the text the reference would write comes back in g_msg (refjson_last_message).
int refjson_dump_cameras(int w, int h, float fl_x, float fl_y, float cx, float cy, const float* cams, const uint32_t* ids, int n) {
	try {
		struct Frame { std::string depth_path, exp_path, eyes_path, file_path, mesh_path, seg_mask_path; float m[4][4]; };
		std::vector<Frame> frames;
		for (int i = 0; i < n; ++i) {
			Frame f{};
			std::string str = std::to_string(ids[i]);
			auto id = std::string(5 - std::min(5, int(str.length())), '0') + str;
			for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) f.m[r][c] = r == c ? 1.0f : 0.0f;
			for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) f.m[r][c] = cams[12 * i + 4 * r + c];
			for (int r = 0; r < 3; ++r) f.m[r][3] -= 0.5f;
			f.exp_path = "flame/exp/" + id + ".txt";
			f.eyes_path = "flame/eyes/" + id + ".txt";
			f.mesh_path = "meshes/" + id + ".obj";
			frames.push_back(f);
		}
		nlohmann::json main;
		main["synthetic"] = true;
		main["h"] = h;
		main["w"] = w;
		main["fl_x"] = fl_x;
		main["fl_y"] = fl_y;
		main["cx"] = cx;
		main["cy"] = cy;
		main["integer_depth_scale"] = 0.001f;
		auto data_array = nlohmann::json::array();
		for (auto& frame : frames) {
			nlohmann::json data;
			data["depth_path"] = frame.depth_path;
			data["exp_path"] = frame.exp_path;
			data["eyes_path"] = frame.eyes_path;
			data["file_path"] = frame.file_path;
			data["mesh_path"] = frame.mesh_path;
			data["seg_mask_path"] = frame.seg_mask_path;
			auto camera = nlohmann::json::array();
			for (int m = 0; m < 4; ++m) {
				auto row = nlohmann::json::array();
				for (int k = 0; k < 4; ++k) row.push_back(frame.m[m][k]);
				camera.push_back(row);
			}
			data["transform_matrix"] = camera;
			data_array.push_back(data);
		}
		std::sort(data_array.begin(), data_array.end(), [](const json& a, const json& b) { return a["mesh_path"] < b["mesh_path"]; });
		main["frames"] = data_array;
		std::ostringstream o;
		o << std::setw(4) << main << std::endl;
		g_msg = o.str();
		return 0;
	} catch (const std::exception& e) {
		g_msg = e.what();
		return 1;
	}
}

int refjson_save_like_reference(const char* path, const char* main_json_text, uint64_t n_params, const void* params_f16, uint64_t n_grid, const void* grid_f16,
                                uint32_t training_step, float loss, const float* aabb6, const float* render_aabb6, float bounding_radius, int aabb_scale,
                                uint32_t rays_per_batch, uint32_t measured, uint32_t measured_before, int with_optimizer, uint32_t current_step, float base_lr,
                                const float* m1, const float* m2, const uint32_t* steps, const void* ema_f16) {
	try {
		json config = json::parse(main_json_text, nullptr, true, true);
		json data;  // Trainer::serialize
		data["n_params"] = (size_t)n_params;
		data["params_binary"] = to_binary(params_f16, n_params * 2);
		if (with_optimizer) {
			json adam;  // AdamOptimizer::serialize
			adam["current_step"] = current_step;
			adam["base_learning_rate"] = base_lr;
			adam["first_moments_binary"] = to_binary(m1, n_params * 4);
			adam["second_moments_binary"] = to_binary(m2, n_params * 4);
			adam["param_steps_binary"] = to_binary(steps, n_params * 4);
			json ema;  // EmaOptimizer::serialize
			ema["nested"] = adam;
			ema["weights_ema_binary"] = to_binary(ema_f16, n_params * 2);
			data["optimizer"] = ema;
		}
		config["snapshot"] = data;
		auto& snapshot = config["snapshot"];
		snapshot["version"] = 1;
		snapshot["density_grid_size"] = 128;
		snapshot["density_grid_binary"] = to_binary(grid_f16, n_grid * 2);
		snapshot["nerf"]["aabb_scale"] = aabb_scale;
		snapshot["training_step"] = training_step;
		snapshot["loss"] = loss;
		snapshot["aabb"] = box_to_json(aabb6);
		snapshot["render_aabb"] = box_to_json(render_aabb6);
		snapshot["bounding_radius"] = bounding_radius;
		snapshot["nerf"]["rgb"]["rays_per_batch"] = rays_per_batch;
		snapshot["nerf"]["rgb"]["measured_batch_size"] = measured;
		snapshot["nerf"]["rgb"]["measured_batch_size_before_compaction"] = measured_before;
		std::ofstream f(path, std::ios::out | std::ios::binary);
		json::to_msgpack(config, f);
		return f.good() ? 0 : 1;
	} catch (const std::exception& e) {
		g_msg = e.what();
		return 1;
	}
}
}
