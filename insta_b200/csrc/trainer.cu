// One full training step behind the C ABI: ngp::Testbed::train_nerf / train_nerf_step (ngp/src/testbed_nerf.cu:2916-2992, 3170-3426)
// + tcnn::Trainer::optimizer_step (tcnn trainer.h:113-131), with rta::Core::surface_closest_point hooked in between sample generation and
// the network (testbed_nerf.cu:3290).  The gradient exchange of a data-parallel run sits between backward and the optimizer
// step and is the only collective of the path: ncclAllReduce(sum, fp16) over the gradient buffer on the caller's stream.
//
// B200-first differences to the reference's step (results for the records the loss reads are unchanged):
//   * the reference deforms and evaluates `max_inference` records (last step's sample count rounded up, 16 x batch on the first step) and
//     needs a host round trip to learn the real count.  Here the deformation and inference kernels read the generator's sample counter on
//     the device (`n_ptr`), so exactly the generated samples are processed, with no host synchronisation before the loss.
//   * the workspace the reference carves out of the tcnn arena every step (testbed_nerf.cu:3175-3202) is allocated once per trainer.
#include <cuda_fp16.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ctx.h"

namespace insta {

// ---- NCCL through dlopen: libinsta_b200.so must load on a machine without NCCL (single-GPU callers) -------------------------------------
namespace {
struct NcclId { char internal[128]; };  // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;
struct NcclApi {
	void* lib = nullptr;
	int (*GetUniqueId)(NcclId*) = nullptr;
	int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
	int (*CommDestroy)(NcclComm) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
	int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;  // optional: peer-memory exchange setup
	const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kNcclHalf = 6, kNcclFloat = 7, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values (nccl.h)

NcclApi* nccl_api(std::string& err) {
	static NcclApi api;
	static bool tried = false;
	static std::string load_err;
	if (!tried) {
		tried = true;
		// a process that already imported torch has its bundled libnccl.so.2 loaded: the soname lookup returns that copy
		for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
			api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if (api.lib) break;
		}
		if (!api.lib) {
			load_err = std::string("NCCL not found: ") + dlerror();
		} else {
			api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
			api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
			api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
			api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
			api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
			api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
			if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) {
				load_err = "NCCL library lacks a required symbol";
				api.lib = nullptr;
			}
		}
	}
	if (!api.lib) {
		err = load_err;
		return nullptr;
	}
	return &api;
}
}  // namespace

// Peer-memory gradient exchange (one node, world <= 8): every rank owns an exchange buffer E (the gradient vector) and a flag line, both exported with
// cudaIpcGetMemHandle and mapped by every other rank.  See p2p_reduce_kernel.
constexpr int kMaxPeers = 8;
struct PeerExchange {
	bool tried = false, ready = false;
	uint64_t n_halves = 0;
	__half* E = nullptr;              // this rank's exchange buffer
	uint32_t* flags = nullptr;        // this rank's flag line: flags[p] = last phase rank p has signalled
	__half* peer_E[kMaxPeers] = {};   // peer_E[rank] = E
	uint32_t* peer_flags[kMaxPeers] = {};
	uint32_t phase = 0;
	void* d_tables = nullptr;         // PeerTables on the device
};
struct Comm;
static void peer_exchange_free(Comm* c);
struct Comm {
	NcclComm comm = nullptr;
	int rank = 0, world = 1;
	PeerExchange px;
};

struct Trainer {
	insta_trainer_config cfg{};
	// workspace of train_nerf_step (testbed_nerf.cu:3175-3202)
	uint32_t* ray_indices = nullptr;
	float* rays = nullptr;
	uint32_t* numsteps = nullptr;
	float* coords = nullptr;
	void* mlp_out = nullptr;
	void* dloss = nullptr;
	float* coords_compacted = nullptr;
	uint32_t* counters = nullptr;   // [0] rays accepted, [1] samples generated, [2] samples after compaction
	float* loss = nullptr;          // per ray
	float* loss_sum = nullptr;      // one float
	uint32_t* h_counters = nullptr; // pinned: 3 counters + loss sum
	// lazily evaluated rounds (train.cu train_round_kernel)
	float* ray_T = nullptr;         // per ray
	uint32_t* round_index = nullptr; // sample indices of the current round (max_samples)
	uint32_t* round_count = nullptr; // one counter per round
	// Testbed::Nerf::Training::Counters (testbed.h:638-650)
	uint32_t rays_per_batch = 1u << 12;
	uint32_t n_rays_total = 0;
	uint32_t measured_batch_size = 0, measured_batch_size_before_compaction = 0;
	uint64_t rng_advances = 0;
	uint32_t training_step = 0;
	// gradient exchange (world > 1)
	GradExchange ex{};
	bool serial_exchange = false;  // INSTA_GRAD_EXCHANGE=serial
	bool p2p_exchange = false;     // INSTA_GRAD_EXCHANGE=p2p: peer-memory exchange (peer_exchange_run); falls back to ncclAllReduce when the setup fails on any rank
	bool no_exchange = false;      // INSTA_GRAD_EXCHANGE=none: the step without its collective (measurement only: the ranks' parameters diverge)
	// INSTA_TRAIN_PROFILE=1: CUDA events between the phases of a step, averaged and printed to stderr every 20 steps (measurement aid; ncu's serialised,
	// cold-cache launch times overstate the latency-bound kernels)
	bool profile = false;
	cudaEvent_t pev[7] = {};
	double pms[6] = {};
	uint32_t psteps = 0;
};

namespace {
constexpr uint32_t kMaxRays = 1u << 18;   // testbed_nerf.cu:2910
constexpr uint32_t kGranularity = 128;    // tcnn::batch_size_granularity
constexpr uint32_t kOutWidth = 4;         // halves per sample of this library's network output (r, g, b, sigma logits)

__global__ void __launch_bounds__(256) loss_sum_kernel(const float* __restrict__ loss, uint32_t n, float* __restrict__ out) {
	float acc = 0.f;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += loss[i];
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	__shared__ float part[8];
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x < 8) {
		acc = part[threadIdx.x];
		for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffu, acc, o);
		if (threadIdx.x == 0) atomicAdd(out, acc);
	}
}

void free_trainer(Trainer* t) {
	if (!t) return;
	cudaFree(t->ray_indices); cudaFree(t->rays); cudaFree(t->numsteps); cudaFree(t->coords); cudaFree(t->mlp_out); cudaFree(t->dloss);
	cudaFree(t->coords_compacted); cudaFree(t->counters); cudaFree(t->loss); cudaFree(t->loss_sum);
	cudaFree(t->ray_T); cudaFree(t->round_index); cudaFree(t->round_count);
	if (t->h_counters) cudaFreeHost(t->h_counters);
	if (t->ex.side) cudaStreamDestroy(t->ex.side);
	for (int g = 0; g < 4; ++g)
		if (t->ex.ev_grid_ready[g]) cudaEventDestroy(t->ex.ev_grid_ready[g]);
	if (t->ex.ev_grid_summed) cudaEventDestroy(t->ex.ev_grid_summed);
	for (auto& ev : t->pev)
		if (ev) cudaEventDestroy(ev);
	delete t;
}
}  // namespace

void trainer_destroy(insta_ctx* ctx) {
	free_trainer((Trainer*)ctx->trainer);
	ctx->trainer = nullptr;
}

void comm_destroy(insta_ctx* ctx) {
	Comm* c = (Comm*)ctx->comm;
	if (!c) return;
	std::string err;
	cudaDeviceSynchronize();
	peer_exchange_free(c);
	if (c->comm)
		if (NcclApi* api = nccl_api(err)) api->CommDestroy(c->comm);
	delete c;
	ctx->comm = nullptr;
}

int comm_unique_id(void* h_id128) {
	std::string err;
	NcclApi* api = nccl_api(err);
	if (!api) INSTA_FAIL(INSTA_ERR_INVALID, "insta_comm_unique_id: " + err);
	NcclId id;
	const int rc = api->GetUniqueId(&id);
	if (rc != 0) INSTA_FAIL(INSTA_ERR_CUDA, std::string("ncclGetUniqueId: ") + (api->GetErrorString ? api->GetErrorString(rc) : "error"));
	std::memcpy(h_id128, &id, sizeof(id));
	return INSTA_OK;
}

int comm_init(insta_ctx* ctx, int world, int rank, const void* h_id128) {
	comm_destroy(ctx);
	Comm* c = new Comm();
	c->rank = rank;
	c->world = world;
	if (world > 1) {
		std::string err;
		NcclApi* api = nccl_api(err);
		if (!api) {
			delete c;
			INSTA_FAIL(INSTA_ERR_INVALID, "insta_comm_init_rank: " + err);
		}
		NcclId id;
		std::memcpy(&id, h_id128, sizeof(id));
		const int rc = api->CommInitRank(&c->comm, world, id, rank);
		if (rc != 0) {
			delete c;
			INSTA_FAIL(INSTA_ERR_CUDA, std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(rc) : "error"));
		}
	}
	ctx->comm = c;
	return INSTA_OK;
}

int comm_world(const insta_ctx* ctx) { return ctx->comm ? ((Comm*)ctx->comm)->world : 1; }
int comm_exchange_mode(const insta_ctx* ctx) {
	const Comm* c = (const Comm*)ctx->comm;
	if (!c || c->world <= 1) return 0;
	return c->px.ready ? 2 : 1;
}

int grad_allreduce(insta_ctx* ctx, void* d_grads_f16, uint64_t n, cudaStream_t s) {
	Comm* c = (Comm*)ctx->comm;
	if (!c || c->world <= 1) return INSTA_OK;  // a single rank owns the whole batch
	std::string err;
	NcclApi* api = nccl_api(err);
	if (!api) INSTA_FAIL(INSTA_ERR_INVALID, "insta_grad_allreduce: " + err);
	const int rc = api->AllReduce(d_grads_f16, d_grads_f16, (size_t)n, kNcclHalf, kNcclSum, c->comm, s);
	if (rc != 0) INSTA_FAIL(INSTA_ERR_CUDA, std::string("ncclAllReduce: ") + (api->GetErrorString ? api->GetErrorString(rc) : "error"));
	return INSTA_OK;
}


// ---- gradient exchange over peer memory (NVLink / NVSwitch) ---------------------------------------------------------------------------------------
// The backward pass of rank r writes its gradient vector into E_r.  Then, instead of ncclAllReduce:
//   barrier (every rank's E is final) -> p2p_reduce_kernel -> barrier (every rank's writes have landed) -> optimizer step on E_r.
// p2p_reduce_kernel on rank r owns slice r of the vector: per 16-byte chunk it loads the chunk from all `world` exchange buffers (world - 1 of the loads
// cross NVLink), adds the eight halves in fp32 in rank order -- every rank would compute the same bits for a chunk, so the replicas stay identical -- and
// stores the fp16 sum into slice r of ALL buffers (world - 1 remote stores).  Per rank 2 (world - 1) / world of the vector cross the link, as in a ring
// all-reduce, but in one pass of independent 16-byte accesses with nothing between the two directions except the kernel itself.  Slice r of any buffer is
// read and written by rank r only; a buffer is not overwritten by the next step's backward pass before every peer has passed the second barrier.
__global__ void __launch_bounds__(256) p2p_reduce_kernel(const uint4* const* __restrict__ src, uint4* const* __restrict__ dst, uint32_t world, uint64_t first, uint64_t count) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	uint4 v[kMaxPeers];
#pragma unroll
	for (int p = 0; p < kMaxPeers; ++p)
		if ((uint32_t)p < world) v[p] = src[p][first + i];
	float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
	for (int p = 0; p < kMaxPeers; ++p)
		if ((uint32_t)p < world) {
			const __half2* h = reinterpret_cast<const __half2*>(&v[p]);
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const float2 f = __half22float2(h[k]);
				acc[2 * k] += f.x;
				acc[2 * k + 1] += f.y;
			}
		}
	uint4 out;
	__half2* o = reinterpret_cast<__half2*>(&out);
#pragma unroll
	for (int k = 0; k < 4; ++k) o[k] = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
#pragma unroll
	for (int p = 0; p < kMaxPeers; ++p)
		if ((uint32_t)p < world) dst[p][first + i] = out;
}

// One warp: lane p tells rank p "rank `rank` has reached `phase`", then waits until every rank has told this one.  A kernel boundary precedes it on the
// stream, so this rank's earlier writes (local and remote) have been issued; the fence orders them before the flag at system scope.
__global__ void p2p_barrier_kernel(uint32_t* const* __restrict__ peer_flags, volatile uint32_t* __restrict__ my_flags, uint32_t rank, uint32_t world, uint32_t phase) {
	const uint32_t p = threadIdx.x;
	__threadfence_system();
	if (p < world) {
		asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[p] + rank), "r"(phase) : "memory");
		uint32_t seen;
		do {
			asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(my_flags + p) : "memory");
		} while ((int32_t)(seen - phase) < 0);
	}
	__syncwarp();
	__threadfence_system();
}

struct PeerTables {  // device copies of the pointer tables (kernel arguments by pointer keep the kernels' parameter space small)
	const uint4* src[kMaxPeers];
	uint4* dst[kMaxPeers];
	uint32_t* flags[kMaxPeers];
};

static void peer_exchange_free(Comm* c) {
	PeerExchange& x = c->px;
	for (int p = 0; p < kMaxPeers; ++p) {
		if (p != c->rank && x.peer_E[p]) cudaIpcCloseMemHandle(x.peer_E[p]);
		if (p != c->rank && x.peer_flags[p]) cudaIpcCloseMemHandle(x.peer_flags[p]);
		x.peer_E[p] = nullptr;
		x.peer_flags[p] = nullptr;
	}
	if (x.E) cudaFree(x.E);
	if (x.flags) cudaFree(x.flags);
	if (x.d_tables) cudaFree(x.d_tables);
	x.d_tables = nullptr;
	x.E = nullptr;
	x.flags = nullptr;
	x.ready = false;
}

// Collective: every rank calls it at the same point (first training step).  Any failure on any rank (no IPC, world > 8, old NCCL) makes ALL ranks fall back
// to ncclAllReduce -- the outcome is agreed upon with an all-reduce of a flag.
static bool peer_exchange_setup(insta_ctx* ctx, uint64_t n_halves, cudaStream_t s) {
	Comm* c = (Comm*)ctx->comm;
	PeerExchange& x = c->px;
	if (x.tried) return x.ready && x.n_halves == n_halves;
	x.tried = true;
	std::string err;
	NcclApi* api = nccl_api(err);
	if (!api || !api->AllGather || c->world > kMaxPeers) return false;
	struct Blob { cudaIpcMemHandle_t e, f; int ok; int pad[3]; };
	Blob mine{};
	bool ok = cudaMalloc(&x.E, n_halves * 2) == cudaSuccess && cudaMalloc(&x.flags, 256) == cudaSuccess;
	ok = ok && cudaMemset(x.flags, 0, 256) == cudaSuccess && cudaMemset(x.E, 0, n_halves * 2) == cudaSuccess;
	ok = ok && cudaIpcGetMemHandle(&mine.e, x.E) == cudaSuccess && cudaIpcGetMemHandle(&mine.f, x.flags) == cudaSuccess;
	mine.ok = ok ? 1 : 0;
	cudaGetLastError();
	Blob* d_blobs = nullptr;
	std::vector<Blob> all(c->world);
	bool coll = cudaMalloc(&d_blobs, sizeof(Blob) * (c->world + 1)) == cudaSuccess;
	if (coll) {
		cudaMemcpy(d_blobs + c->world, &mine, sizeof(Blob), cudaMemcpyHostToDevice);
		coll = api->AllGather(d_blobs + c->world, d_blobs, sizeof(Blob), /*ncclChar*/ 0, c->comm, s) == 0 && cudaStreamSynchronize(s) == cudaSuccess;
		if (coll) cudaMemcpy(all.data(), d_blobs, sizeof(Blob) * c->world, cudaMemcpyDeviceToHost);
	}
	bool all_ok = coll;
	for (int p = 0; p < c->world && all_ok; ++p) all_ok = all[p].ok == 1;
	if (all_ok) {
		for (int p = 0; p < c->world && all_ok; ++p) {
			if (p == c->rank) {
				x.peer_E[p] = x.E;
				x.peer_flags[p] = x.flags;
				continue;
			}
			void *pe = nullptr, *pf = nullptr;
			all_ok = cudaIpcOpenMemHandle(&pe, all[p].e, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
			         cudaIpcOpenMemHandle(&pf, all[p].f, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
			x.peer_E[p] = (__half*)pe;
			x.peer_flags[p] = (uint32_t*)pf;
		}
		cudaGetLastError();
	}
	// agree: min over the ranks of "I mapped everybody" (the blob buffer doubles as the flag's storage)
	int agreed = 0;
	if (coll) {
		const float mine_ok = all_ok ? 0.f : 1.f;  // sum of failures
		cudaMemcpy(d_blobs, &mine_ok, 4, cudaMemcpyHostToDevice);
		float total = 1.f;
		if (api->AllReduce(d_blobs, d_blobs, 1, kNcclFloat, kNcclSum, c->comm, s) == 0 && cudaStreamSynchronize(s) == cudaSuccess &&
		    cudaMemcpy(&total, d_blobs, 4, cudaMemcpyDeviceToHost) == cudaSuccess)
			agreed = total == 0.f ? 1 : 0;
	}
	if (d_blobs) cudaFree(d_blobs);
	if (agreed) {
		PeerTables t{};
		for (int p = 0; p < c->world; ++p) {
			t.src[p] = reinterpret_cast<const uint4*>(x.peer_E[p]);
			t.dst[p] = reinterpret_cast<uint4*>(x.peer_E[p]);
			t.flags[p] = x.peer_flags[p];
		}
		if (cudaMalloc(&x.d_tables, sizeof(PeerTables)) != cudaSuccess || cudaMemcpy(x.d_tables, &t, sizeof(PeerTables), cudaMemcpyHostToDevice) != cudaSuccess) agreed = 0;
		// (a rank failing here would leave the others waiting at the first barrier: 100 bytes of device memory do not fail after 51 MB succeeded)
	}
	if (!agreed) {
		peer_exchange_free(c);
		return false;
	}
	x.n_halves = n_halves;
	x.ready = true;
	return true;
}

// barrier -> reduce -> barrier on stream s; E holds the summed gradients on every rank afterwards
static int peer_exchange_run(insta_ctx* ctx, cudaStream_t s) {
	Comm* c = (Comm*)ctx->comm;
	PeerExchange& x = c->px;
	PeerTables* d_tables = (PeerTables*)x.d_tables;
	const uint64_t chunks = x.n_halves / 8;  // the parameter count is a multiple of 8 (16-byte grid entries, 25,600 MLP weights)
	const uint64_t per = (chunks + c->world - 1) / c->world;
	const uint64_t first = std::min<uint64_t>(per * c->rank, chunks), count = std::min<uint64_t>(per, chunks - first);
	p2p_barrier_kernel<<<1, 32, 0, s>>>(d_tables->flags, x.flags, (uint32_t)c->rank, (uint32_t)c->world, ++x.phase);
	if (count) p2p_reduce_kernel<<<(uint32_t)((count + 255) / 256), 256, 0, s>>>(d_tables->src, d_tables->dst, (uint32_t)c->world, first, count);
	p2p_barrier_kernel<<<1, 32, 0, s>>>(d_tables->flags, x.flags, (uint32_t)c->rank, (uint32_t)c->world, ++x.phase);
	ctx->launches += 3;
	INSTA_CUDA(cudaGetLastError());
	return INSTA_OK;
}

int trainer_create(insta_ctx* ctx, const insta_trainer_config* cfg) {
	INSTA_CUDA(cudaDeviceSynchronize());
	trainer_destroy(ctx);
	Trainer* t = new Trainer();
	t->cfg = *cfg;
	// Default: one ncclAllReduce after the backward pass.  The pipelined form (INSTA_GRAD_EXCHANGE=overlap: scatter kernel per level group, each group's sum on
	// a side stream) measured no better on 2 GPUs (3.17 vs 3.14 ms per step, 3.01 without the collective; profiles/r2_train_exchange.txt): four short scatter
	// launches and four latency-bound collectives cost what the overlap gains, and next to the persistent weight-gradient kernel NCCL's kernel finds no SM.
	t->serial_exchange = true;
	if (const char* e = getenv("INSTA_TRAIN_PROFILE")) t->profile = atoi(e) != 0;
	if (t->profile)
		for (auto& ev : t->pev) cudaEventCreate(&ev);
	if (const char* e = getenv("INSTA_GRAD_EXCHANGE")) {
		t->serial_exchange = std::string(e) != "overlap";
		t->no_exchange = std::string(e) == "none";
		t->p2p_exchange = std::string(e) == "p2p";
	}
	const uint32_t target = cfg->target_batch_size;
	const uint32_t max_samples = target * 16;  // "somewhat of a worst case" (testbed_nerf.cu:3172)
	const uint32_t stride = 11 + ctx->p.exp_params + ctx->p.geo_params;
	t->rays_per_batch = std::min(std::max(cfg->rays_per_batch, 1u), kMaxRays);
	auto fail = [&](cudaError_t e) {
		set_error(std::string("insta_trainer_create: ") + cudaGetErrorString(e));
		free_trainer(t);
		return INSTA_ERR_CUDA;
	};
	cudaError_t e;
	if ((e = cudaMalloc(&t->ray_indices, sizeof(uint32_t) * kMaxRays)) != cudaSuccess) return fail(e);
	if ((e = cudaMalloc(&t->rays, sizeof(float) * 6 * kMaxRays)) != cudaSuccess) return fail(e);
	if ((e = cudaMalloc(&t->numsteps, sizeof(uint32_t) * 2 * kMaxRays)) != cudaSuccess) return fail(e);
	if ((e = cudaMalloc(&t->coords, sizeof(float) * (size_t)max_samples * stride)) != cudaSuccess) return fail(e);
	if ((e = cudaMalloc(&t->mlp_out, 2 * (size_t)max_samples * kOutWidth)) != cudaSuccess) return fail(e);
	if ((e = cudaMalloc(&t->dloss, 2 * (size_t)target * kOutWidth)) != cudaSuccess) return fail(e);
	if ((e = cudaMalloc(&t->coords_compacted, sizeof(float) * (size_t)target * stride)) != cudaSuccess) return fail(e);
	if ((e = cudaMalloc(&t->counters, sizeof(uint32_t) * 4)) != cudaSuccess) return fail(e);
	if ((e = cudaMalloc(&t->loss, sizeof(float) * kMaxRays)) != cudaSuccess) return fail(e);
	if ((e = cudaMalloc(&t->ray_T, sizeof(float) * kMaxRays)) != cudaSuccess) return fail(e);
	if ((e = cudaMalloc(&t->round_index, sizeof(uint32_t) * (size_t)max_samples)) != cudaSuccess) return fail(e);
	if ((e = cudaMalloc(&t->round_count, sizeof(uint32_t) * 8)) != cudaSuccess) return fail(e);
	if ((e = cudaMalloc(&t->loss_sum, sizeof(float))) != cudaSuccess) return fail(e);
	if ((e = cudaHostAlloc(&t->h_counters, sizeof(uint32_t) * 4, cudaHostAllocDefault)) != cudaSuccess) return fail(e);
	// the reference's arena memory is not cleared either, but stale records must at least be finite for the first step's kernels
	if ((e = cudaMemset(t->coords, 0, sizeof(float) * (size_t)max_samples * stride)) != cudaSuccess) return fail(e);
	// a step that accepts no ray leaves the compacted batch untouched (the reference's fill_rollover divides by zero there): zero records and
	// zero loss gradients then give zero parameter gradients instead of whatever the allocation held
	if ((e = cudaMemset(t->coords_compacted, 0, sizeof(float) * (size_t)target * stride)) != cudaSuccess) return fail(e);
	if ((e = cudaMemset(t->dloss, 0, 2 * (size_t)target * kOutWidth)) != cudaSuccess) return fail(e);
	ctx->trainer = t;
	return INSTA_OK;
}

void trainer_rng_advance(insta_ctx* ctx) {
	if (Trainer* t = (Trainer*)ctx->trainer) t->rng_advances += 1;
}

int trainer_state(const insta_ctx* ctx, insta_train_stats* out) {
	const Trainer* t = (const Trainer*)ctx->trainer;
	if (!t) INSTA_FAIL(INSTA_ERR_INVALID, "no trainer (insta_trainer_create)");
	std::memset(out, 0, sizeof(*out));
	out->training_step = t->training_step;
	out->rays_per_batch = t->rays_per_batch;
	out->n_rays_total = t->n_rays_total;
	out->measured_batch_size = t->measured_batch_size;
	out->measured_batch_size_before_compaction = t->measured_batch_size_before_compaction;
	return INSTA_OK;
}

int train_nerf(insta_ctx* ctx, const insta_optimizer* opt, const void* d_bitfield, const float* d_extra_dims, const float* d_mean_density, int get_loss_scalar,
               insta_train_stats* h_stats, cudaStream_t s) {
	Trainer* t = (Trainer*)ctx->trainer;
	if (!t) INSTA_FAIL(INSTA_ERR_INVALID, "insta_train_nerf: no trainer (insta_trainer_create)");
	const insta_trainer_config& c = t->cfg;
	const uint32_t target = c.target_batch_size;
	const uint32_t n_rays = c.fixed_rays_per_batch ? std::min(c.fixed_rays_per_batch, kMaxRays) : t->rays_per_batch;
	const uint32_t world = (uint32_t)comm_world(ctx);
	const uint32_t stride = 11 + ctx->p.exp_params + ctx->p.geo_params;

	// Counters::prepare_for_training_steps (testbed_nerf.cu:2880-2887) -- the generator clears the counters itself
	INSTA_CUDA(cudaMemsetAsync(t->loss, 0, sizeof(float) * n_rays, s));

	insta_train_step st{};
	st.n_rays = n_rays;
	st.n_rays_total = t->n_rays_total;
	// max_inference (testbed_nerf.cu:3217-3222): the generator's capacity is last step's sample count rounded up, the full workspace on step 0
	st.max_samples = t->measured_batch_size_before_compaction == 0
	                     ? target * 16
	                     : (std::min(t->measured_batch_size_before_compaction, target * 16) + kGranularity - 1) / kGranularity * kGranularity;
	st.target_batch_size = target;
	st.stride_floats = stride;
	st.snap_to_pixel_centers = c.snap_to_pixel_centers;
	st.random_bg_color = c.random_bg_color;
	st.penalize_density_alpha = c.penalize_density_alpha;
	st.rng_seed = c.rng_seed;
	st.rng_advances = t->rng_advances;
	st.cone_angle_constant = c.cone_angle_constant;
	// every rank contributes the mean over ITS rays; the summed gradient must be the mean over all ranks (pre-scale by 1/R
	// so that the fp16 sum cannot overflow where a single rank's gradient would not)
	st.loss_scale = c.loss_scale / (float)world;
	for (int k = 0; k < 3; ++k) st.background_color[k] = c.background_color[k];
	st.depth_weight = c.depth_weight;
	st.seg_weight = c.seg_weight;
	st.near_distance = c.near_distance;
	st.beta_loss = c.beta_loss;
	st.d_density_grid_bitfield = d_bitfield;
	st.d_extra_dims = d_extra_dims;
	st.d_coords = t->coords;
	st.d_counters = t->counters;
	st.d_ray_indices = t->ray_indices;
	st.d_rays = t->rays;
	st.d_numsteps = t->numsteps;
	t->n_rays_total += n_rays;  // testbed_nerf.cu:3235

	auto mark = [&](int k) { if (t->profile) cudaEventRecord(t->pev[k], s); };
	mark(0);
	// with lazily evaluated rounds the generator leaves (ray parameter, ray slot) pairs and every round expands the records it is about to evaluate
	const bool lazy_rounds = !(ctx->train_all_samples || !deform_supports_index(ctx));
	int rc = train_generate(ctx, &st, s, /*emit=*/!lazy_rounds);
	if (rc) return rc;
	mark(1);
	// rta::Core::surface_closest_point on the generated records (testbed_nerf.cu:3290), then the network on the same records (3292)
	{
		std::lock_guard<std::mutex> lock(ctx->mu);
		rc = sync_frame_table(ctx, s);
	}
	if (rc) return rc;
	if (ctx->train_all_samples || !deform_supports_index(ctx)) {
		rc = launch_deform(ctx, t->coords, st.max_samples, stride, -1, nullptr, false, nullptr, s, t->counters + 1);
		if (rc) return rc;
		rc = launch_network_fwd(ctx, t->coords, stride, st.max_samples, opt->d_params_f16, t->mlp_out, 0, s, t->counters + 1);
		if (rc) return rc;
	} else {
		// Only the samples the loss kernel will read: rounds of [0, 8), [8, 32), [32, end) samples per ray, a ray drops out once its transmittance is
		// below the loss kernel's threshold (train_round_kernel).  Device-side counts, no host round trip.
		const std::vector<uint32_t>& edges = ctx->train_round_edges;  // {0, 8, 32, UINT32_MAX} by default (INSTA_TRAIN_ROUND_EDGES)
		const int n_rounds = (int)edges.size() - 1;
		INSTA_CUDA(cudaMemsetAsync(t->round_count, 0, sizeof(uint32_t) * 8, s));
		for (int r = 0; r < n_rounds; ++r) {
			rc = train_round(ctx, &st, t->mlp_out, kOutWidth, r ? edges[r - 1] : 0u, r ? edges[r] : 0u, edges[r], edges[r + 1], t->ray_T, t->round_index,
			                 t->round_count + r, s);
			if (rc) return rc;
			// upper bound of the round's size for the launch geometry; the kernels read the real count on the device
			const uint64_t bound = edges[r + 1] != 0xffffffffu ? std::min<uint64_t>((uint64_t)n_rays * (edges[r + 1] - edges[r]), st.max_samples) : st.max_samples;
			rc = train_emit_round(ctx, &st, t->round_index, t->round_count + r, (uint32_t)bound, s);
			if (rc) return rc;
			rc = launch_deform(ctx, t->coords, (uint32_t)bound, stride, -1, nullptr, false, nullptr, s, t->round_count + r, false, false, t->round_index);
			if (rc) return rc;
			rc = launch_network_fwd(ctx, t->coords, stride, (uint32_t)bound, opt->d_params_f16, t->mlp_out, 0, s, t->round_count + r, nullptr, t->round_index);
			if (rc) return rc;
		}
	}
	mark(2);
	rc = train_loss(ctx, &st, t->mlp_out, kOutWidth, t->coords_compacted, t->dloss, t->loss, d_mean_density, s);
	if (rc) return rc;
	mark(3);
	// forward + backward on the compacted, rolled-over batch (testbed_nerf.cu:3366-3371)
	// data-parallel exchange step (world > 1): one ncclAllReduce over the whole gradient vector after the backward pass, or (INSTA_GRAD_EXCHANGE=overlap)
	// the grid gradients summed group of levels by group of levels on a side stream while the scatter kernel works on the next group
	const bool overlap = world > 1 && !t->serial_exchange && !t->no_exchange;
	if (overlap && !t->ex.side) {
		INSTA_CUDA(cudaStreamCreateWithFlags(&t->ex.side, cudaStreamNonBlocking));
		for (int g = 0; g < 4; ++g) INSTA_CUDA(cudaEventCreateWithFlags(&t->ex.ev_grid_ready[g], cudaEventDisableTiming));
		INSTA_CUDA(cudaEventCreateWithFlags(&t->ex.ev_grid_summed, cudaEventDisableTiming));
	}
	// INSTA_GRAD_EXCHANGE=p2p: the backward pass writes into this rank's exported exchange buffer, the sum is formed over peer memory (peer_exchange_run)
	const bool p2p = world > 1 && t->p2p_exchange && !t->no_exchange && peer_exchange_setup(ctx, network_n_params(ctx), s);
	void* grads = p2p ? (void*)((Comm*)ctx->comm)->px.E : (void*)opt->d_gradients_f16;
	rc = launch_network_fwd_bwd(ctx, t->coords_compacted, stride, target, opt->d_params_f16, t->dloss, kOutWidth, grads, nullptr, s, overlap && !p2p ? &t->ex : nullptr);
	if (rc) return rc;
	t->rng_advances += 1;  // m_rng.advance() (testbed_nerf.cu:3419)
	mark(4);
	if (p2p) {
		rc = peer_exchange_run(ctx, s);
		if (rc) return rc;
		// the caller's gradient buffer holds the summed gradients after the step, as with the collective
		INSTA_CUDA(cudaMemcpyAsync((void*)opt->d_gradients_f16, grads, (size_t)network_n_params(ctx) * 2, cudaMemcpyDeviceToDevice, s));
	} else if (!overlap && !t->no_exchange) {
		rc = grad_allreduce(ctx, (void*)opt->d_gradients_f16, network_n_params(ctx), s);
		if (rc) return rc;
	}
	mark(5);
	// Trainer::optimizer_step(stream, LOSS_SCALE) (testbed_nerf.cu:2974)
	if (!c.freeze_parameters) {
		insta_optimizer o = *opt;
		o.d_gradients_f16 = (decltype(o.d_gradients_f16))grads;
		o.loss_scale = c.loss_scale;
		rc = optimizer_step(ctx, &o, network_n_params(ctx), 25600u, s);
		if (rc) return rc;
	}
	t->training_step += 1;
	mark(6);

	// Counters::update_after_training (testbed_nerf.cu:2889-2914): one read-back per step, as in the reference
	if (get_loss_scalar) {
		INSTA_CUDA(cudaMemsetAsync(t->loss_sum, 0, sizeof(float), s));
		loss_sum_kernel<<<std::min<uint32_t>((n_rays + 255) / 256, 148u), 256, 0, s>>>(t->loss, n_rays, t->loss_sum);
		ctx->launches++;
		INSTA_CUDA(cudaMemcpyAsync(t->h_counters + 3, t->loss_sum, sizeof(float), cudaMemcpyDeviceToHost, s));
	}
	INSTA_CUDA(cudaMemcpyAsync(t->h_counters, t->counters, sizeof(uint32_t) * 3, cudaMemcpyDeviceToHost, s));
	INSTA_CUDA(cudaStreamSynchronize(s));
	if (t->profile) {
		static const char* const names[6] = {"generate", "deform+network", "loss+compaction", "forward+backward", "exchange", "optimizer"};
		for (int k = 0; k < 6; ++k) {
			float ms = 0.f;
			cudaEventElapsedTime(&ms, t->pev[k], t->pev[k + 1]);
			t->pms[k] += ms;
		}
		if (++t->psteps == 20) {
			fprintf(stderr, "[insta_train_nerf] step %u, mean of 20 steps (ms):", t->training_step);
			for (int k = 0; k < 6; ++k) fprintf(stderr, " %s %.3f", names[k], t->pms[k] / 20.0), t->pms[k] = 0.0;
			fprintf(stderr, "\n");
			t->psteps = 0;
		}
	}
	const uint32_t generated = t->h_counters[1], compacted = t->h_counters[2];
	float loss_scalar = 0.f;
	t->measured_batch_size = t->measured_batch_size_before_compaction = 0;
	if (generated != 0 && compacted != 0) {
		t->measured_batch_size_before_compaction = generated;
		t->measured_batch_size = compacted;
		if (get_loss_scalar) {
			float sum;
			std::memcpy(&sum, t->h_counters + 3, sizeof(float));
			loss_scalar = sum * (float)compacted / (float)target;
		}
		if (!c.fixed_rays_per_batch) {
			uint32_t r = (uint32_t)((float)t->rays_per_batch * (float)target / (float)compacted);
			r = (r + kGranularity - 1) / kGranularity * kGranularity;
			t->rays_per_batch = std::min(r, kMaxRays);
		}
	}
	if (h_stats) {
		trainer_state(ctx, h_stats);
		h_stats->n_rays = n_rays;
		h_stats->n_rays_accepted = t->h_counters[0];
		h_stats->loss = loss_scalar;
	}
	return INSTA_OK;
}

}  // namespace insta
