// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI launcher around the REFERENCE's own network: ngp::NerfNetwork<__half> (ngp/include/neural-graphics-primitives/nerf_network.h) on top of
// tiny-cuda-nn's GridEncoding, SphericalHarmonics and FullyFusedMLP, compiled for sm_100 from the sources where they lie (oracle/Makefile:
// refnet; tcnn's translation units are compiled unmodified into oracle/_ref/tcnn/*.o).  Built with the configuration of
// configs/nerf/main.json as ngp::Testbed::reset_network builds it (ngp/src/testbed.cu:2246-2300).  Used by tests/test_gpu_network_ref.py
// (parity of the fused tcgen05 network kernels against the reference's inference, density and forward + backward) and by bench.py's
// ref_cuda leg (reference timings of the network).
#include <neural-graphics-primitives/common.h>
#include <neural-graphics-primitives/nerf.h>
#include <neural-graphics-primitives/random_val.cuh>
#include <neural-graphics-primitives/nerf_network.h>

#include <tiny-cuda-nn/common.h>
#include <tiny-cuda-nn/gpu_matrix.h>
#include <tiny-cuda-nn/gpu_memory.h>
#include <tiny-cuda-nn/optimizer.h>

#include <memory>
#include <string>

using namespace tcnn;
using namespace ngp;
using precision_t = network_precision_t;

namespace {
std::shared_ptr<NerfNetwork<precision_t>> g_net;
GPUMemory<precision_t> g_params, g_grads;
GPUMemory<float> g_params_fp32;
std::shared_ptr<Optimizer<precision_t>> g_opt;
uint32_t g_stride = 0;
std::string g_err;
}  // namespace

extern "C" {

const char* refnet_last_error() { return g_err.c_str(); }

// returns n_params (> 0) or -1
long refnet_create(uint32_t n_extra_dims, uint32_t n_levels, uint32_t n_features, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale) {
	try {
		nlohmann::json enc = {{"otype", "HashGrid"}, {"n_levels", n_levels}, {"n_features_per_level", n_features}, {"log2_hashmap_size", log2_hashmap_size},
		                      {"base_resolution", base_resolution}, {"per_level_scale", per_level_scale}};
		nlohmann::json dir = {{"otype", "Composite"},
		                      {"nested", {{{"n_dims_to_encode", 3}, {"otype", "SphericalHarmonics"}, {"degree", 4}}, {{"otype", "Identity"}, {"n_bins", 4}, {"degree", 4}}}}};
		nlohmann::json net = {{"otype", "FullyFusedMLP"}, {"activation", "ReLU"}, {"output_activation", "None"}, {"n_neurons", 64}, {"n_hidden_layers", 2},
		                      {"n_output_dims", 16}};
		nlohmann::json rgb = {{"otype", "FullyFusedMLP"}, {"activation", "ReLU"}, {"output_activation", "None"}, {"n_neurons", 64}, {"n_hidden_layers", 3}};
		nlohmann::json mapper = nlohmann::json::object();
		g_stride = sizeof(NerfCoordinate) / sizeof(float) + n_extra_dims;
		// dir_offset = n_pos + dt + holder + deformed_pos (testbed.cu:2292)
		g_net = std::make_shared<NerfNetwork<precision_t>>(3, 3, n_extra_dims, 3 + 1 + 1 + 3, enc, dir, net, rgb, mapper);
		const size_t n = g_net->n_params();
		g_params.resize(n);
		g_grads.resize(n);
		g_params.memset(0);
		g_grads.memset(0);
		g_net->set_params(g_params.data(), g_params.data(), g_params.data(), g_grads.data());
		return (long)n;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

int refnet_set_params(const void* d_params_f16) {
	if (!g_net) return -1;
	return (int)cudaMemcpy(g_params.data(), d_params_f16, g_params.get_bytes(), cudaMemcpyDeviceToDevice);
}

uint32_t refnet_padded_output_width() { return g_net ? g_net->padded_output_width() : 0; }

// NerfNetwork::inference_mixed_precision on n records (n a multiple of 128); d_out: row-major [padded_output_width][n] as NerfTracer uses it
int refnet_inference(const float* d_records, uint32_t n, void* d_out_rm, cudaStream_t stream) {
	try {
		GPUMatrix<float> in((float*)d_records, g_stride, n);
		GPUMatrix<precision_t, RM> out((precision_t*)d_out_rm, g_net->padded_output_width(), n);
		g_net->inference_mixed_precision(stream, in, out, false);
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

// NerfNetwork::density on bare positions (n x 3 floats): row-major [padded_density_output_width][n], row 0 = density logit
int refnet_density(const float* d_xyz, uint32_t n, void* d_out_rm, cudaStream_t stream) {
	try {
		GPUMatrix<float> in((float*)d_xyz, 3, n);
		GPUMatrix<precision_t, RM> out((precision_t*)d_out_rm, g_net->padded_density_output_width(), n);
		default_rng_t rng{1337};
		g_net->density(stream, in, out, rng, false, nullptr, 1);
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

// train_nerf_step's forward + backward on the compacted batch (testbed_nerf.cu:3366-3371): column-major output / dL_doutput
// ([padded_output_width] halves per sample), gradients (EGradientMode::Overwrite) copied to d_grads_f16 in parameter order
int refnet_forward_backward(const float* d_records, uint32_t n, const void* d_dloss_cm, void* d_out_cm, void* d_grads_f16, cudaStream_t stream) {
	try {
		GPUMatrix<float> in((float*)d_records, g_stride, n);
		GPUMatrix<precision_t> out((precision_t*)d_out_cm, g_net->padded_output_width(), n);
		GPUMatrix<precision_t> dl((precision_t*)d_dloss_cm, g_net->padded_output_width(), n);
		auto ctx = g_net->forward(stream, in, &out, false, false);
		g_net->backward(stream, *ctx, in, out, dl, nullptr, false, EGradientMode::Overwrite);
		if (d_grads_f16) cudaMemcpyAsync(d_grads_f16, g_grads.data(), g_grads.get_bytes(), cudaMemcpyDeviceToDevice, stream);
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

// inference with the column-major output train_nerf_step uses ([padded_output_width] halves per sample, testbed_nerf.cu:3225)
int refnet_inference_cm(const float* d_records, uint32_t n, void* d_out_cm, cudaStream_t stream) {
	try {
		GPUMatrix<float> in((float*)d_records, g_stride, n);
		GPUMatrix<precision_t> out((precision_t*)d_out_cm, g_net->padded_output_width(), n);
		g_net->inference_mixed_precision(stream, in, out, false);
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

// tcnn::Trainer::optimizer_step with the optimizer of configs/nerf/main.json:21-32 (Ema{Adam}); the fp32 master copy is made on first use
int refnet_optimizer_step(float loss_scale, cudaStream_t stream) {
	try {
		if (!g_opt) {
			nlohmann::json cfg = {{"otype", "Ema"}, {"decay", 0.95},
			                      {"nested", {{"otype", "Adam"}, {"learning_rate", 0.00175}, {"beta1", 0.9}, {"beta2", 0.99}, {"epsilon", 1e-15}, {"l2_reg", 1e-6}}}};
			g_opt.reset(create_optimizer<precision_t>(cfg));
			g_opt->allocate(g_net);
			g_params_fp32.resize(g_params.size());
			parallel_for_gpu(stream, g_params.size(), [dst = g_params_fp32.data(), src = g_params.data()] __device__(size_t i) { dst[i] = (float)src[i]; });
		}
		g_opt->step(stream, loss_scale, g_params_fp32.data(), g_params.data(), g_grads.data());
		return 0;
	} catch (const std::exception& e) {
		g_err = e.what();
		return -1;
	}
}

// For the optimizer parity test: feed a gradient vector, read back what Ema{Adam} made of it
int refnet_set_grads(const void* d_grads_f16, cudaStream_t stream) {
	cudaMemcpyAsync(g_grads.data(), d_grads_f16, g_grads.get_bytes(), cudaMemcpyDeviceToDevice, stream);
	return 0;
}
int refnet_get_optimizer_state(float* d_params_fp32, void* d_params_f16, void* d_ema_f16, cudaStream_t stream) {
	if (!g_opt) {
		g_err = "no optimizer step has run yet";
		return -1;
	}
	if (d_params_fp32) cudaMemcpyAsync(d_params_fp32, g_params_fp32.data(), g_params_fp32.get_bytes(), cudaMemcpyDeviceToDevice, stream);
	if (d_params_f16) cudaMemcpyAsync(d_params_f16, g_params.data(), g_params.get_bytes(), cudaMemcpyDeviceToDevice, stream);
	if (d_ema_f16) cudaMemcpyAsync(d_ema_f16, g_opt->custom_weights(), g_params.get_bytes(), cudaMemcpyDeviceToDevice, stream);
	return 0;
}
// forget the optimizer (its moments, step count and fp32 master copy): the next refnet_optimizer_step starts from the current fp16 parameters
int refnet_reset_optimizer() {
	g_opt.reset();
	return 0;
}

}  // extern "C"
