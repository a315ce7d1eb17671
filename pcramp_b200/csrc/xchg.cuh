// xchg.cuh -- the multi-GPU exchange of a batch's results, fused into the tail of pair scoring (included by pcramp_gpu.cu).
//
// Replaces the BitSet gather of reduce_best_assay (main.cpp:1421-1601; MPI_Send / MPI_Recv of every rank's BitSet).
// Sequences are sharded across the GPUs of one box in contiguous index ranges whose boundaries are multiples of 32
// (SURVEY.md section 8e), so the shards own disjoint WORDS of a pair's LSB-first bitset.  Every rank keeps a
// "global" result buffer (n_pairs x ceil(N_seq / 32) words, double buffered) that its peers can address over NVLink
// (cudaIpc handles between the processes of a torchrun job, or plain pointers between contexts of one process):
//
//   push_kernel      right after the scoring kernels: every rank stores its shard's words into the global buffer of EVERY
//                    rank (peer stores through NVSwitch, coalesced), then the last CTA to finish publishes the step
//                    number in each peer's flag slot (__threadfence_system + volatile store)
//   wait_kernel      spins (bounded) until all ranks' flags have reached this step
//   coverage kernels the ones of score.cuh, on the global bitsets: population count for unit weights, else the reference's
//                    summation order (pass-1 detections ascending, then pass-2-only) over the global weights
//
// No NCCL call, no host synchronisation and no staging copy sits between the scoring kernels and the merged result; a
// step's exchange is three small launches on the library's stream.  Double buffering makes re-use safe: a rank can run
// at most one step ahead of a peer (its own wait for step k+1 needs the peer's push of step k+1, which the peer issues
// after its coverage of step k).
#pragma once
#include "ctx.cuh"
#include "score.cuh"

namespace pcr {
namespace xchg {

constexpr uint32_t MAX_WORLD = 16;
constexpr uint32_t FLAG_STRIDE = 32; // one 128-byte line per writer

struct Peers {
	uint32_t *base[MAX_WORLD];
};

struct State {
	uint32_t rank = 0, world = 0, max_pairs = 0, total_seq = 0, words_global = 0;
	std::vector<uint32_t> shard_lo; // world + 1 sequence offsets
	void *buf = nullptr;            // this rank's buffer: [2][any | pass1][max_pairs][words_global] words, then flags[MAX_WORLD][FLAG_STRIDE]
	size_t buf_bytes = 0, plane_words = 0, flag_off_words = 0;
	Peers peers = {};
	bool connected = false;
	std::vector<void *> ipc_opened;
	DevBuf d_weight, d_cov, d_done, d_err;
	bool unit_weights = true;
	uint32_t step = 0, last_pairs = 0;
};

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t *p) { return *(const volatile uint32_t *)p; }

// src: this shard's n_pairs x words_local words; every destination gets them at word offset word_off of its rows
__global__ void __launch_bounds__(256) push_kernel(const uint32_t *__restrict__ src_any, const uint32_t *__restrict__ src_p1, uint32_t n_pairs,
	uint32_t words_local, uint32_t words_global, uint32_t word_off, Peers peers, uint32_t world, size_t any_off, size_t p1_off, size_t flag_off,
	uint32_t my_rank, uint32_t step, unsigned int *done)
{
	const uint64_t total = (uint64_t)n_pairs * words_local;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t p = (uint32_t)(i / words_local), w = (uint32_t)(i % words_local);
		const uint32_t a = src_any[i];
		const size_t o = (size_t)p * words_global + word_off + w;
		if (src_p1) {
			const uint32_t b = src_p1[i];
			for (uint32_t d = 0; d < world; ++d) {
				peers.base[d][any_off + o] = a;
				peers.base[d][p1_off + o] = b;
			}
		} else {
			for (uint32_t d = 0; d < world; ++d) peers.base[d][any_off + o] = a;
		}
	}
	// publish: the last CTA to get here knows every store of this launch was issued and fenced
	__threadfence_system();
	__syncthreads();
	__shared__ unsigned int s_last;
	if (threadIdx.x == 0) s_last = (atomicAdd(done, 1u) == gridDim.x - 1u) ? 1u : 0u;
	__syncthreads();
	if (s_last) {
		__threadfence_system();
		if (threadIdx.x < world) *(volatile uint32_t *)(peers.base[threadIdx.x] + flag_off + (size_t)my_rank * FLAG_STRIDE) = step;
		if (threadIdx.x == 0) *done = 0u;
	}
}

// one CTA; thread s waits for rank s.  ~2 s of polling at most, then the error word is set and the host call reports it.
__global__ void wait_kernel(const uint32_t *flags, uint32_t world, uint32_t step, unsigned int *err)
{
	if (threadIdx.x < world) {
		const uint32_t *f = flags + (size_t)threadIdx.x * FLAG_STRIDE;
		const long long t0 = clock64();
		bool ok = false;
		for (;;) {
			const uint32_t v = ld_volatile(f);
			if ((int32_t)(v - step) >= 0) { ok = true; break; }
			if (clock64() - t0 > 4000000000ll) break;
			__nanosleep(200);
		}
		if (!ok) atomicOr(err, 1u << threadIdx.x);
	}
	__threadfence_system();
}

} // namespace xchg
} // namespace pcr

struct pcramp_gpu_xchg : pcr::xchg::State {};

extern "C" {

int pcramp_gpu_exchange_create(pcramp_gpu_ctx *ctx, uint32_t rank, uint32_t world, const uint32_t *shard_nseq, uint32_t max_pairs,
	const float *weight_all)
{
	using namespace pcr::xchg;
	if (!ctx) return 1;
	if (world == 0 || world > MAX_WORLD || rank >= world || !shard_nseq || max_pairs == 0) return fail(ctx, "pcramp_gpu_exchange_create: bad arguments");
	CK(cudaSetDevice(ctx->device));
	if (ctx->xchg) pcramp_gpu_exchange_destroy(ctx);
	pcramp_gpu_xchg *x = new pcramp_gpu_xchg();
	ctx->xchg = x;
	x->rank = rank;
	x->world = world;
	x->max_pairs = max_pairs;
	x->shard_lo.assign(world + 1, 0);
	for (uint32_t s = 0; s < world; ++s) {
		if (s + 1 < world && (shard_nseq[s] % 32u) != 0u)
			return fail(ctx, "pcramp_gpu_exchange_create: every shard but the last must hold a multiple of 32 sequences (disjoint bitset words)");
		x->shard_lo[s + 1] = x->shard_lo[s] + shard_nseq[s];
	}
	x->total_seq = x->shard_lo[world];
	x->words_global = (x->total_seq + 31u) / 32u;
	x->plane_words = (size_t)max_pairs * std::max<uint32_t>(1, x->words_global);
	x->flag_off_words = 4 * x->plane_words;
	x->buf_bytes = (x->flag_off_words + (size_t)MAX_WORLD * FLAG_STRIDE) * 4;
	CK(cudaMalloc(&x->buf, x->buf_bytes));
	CK(cudaMemsetAsync(x->buf, 0, x->buf_bytes, ctx->stream));
	x->unit_weights = true;
	std::vector<float> w(std::max<uint32_t>(1, x->total_seq), 1.0f);
	if (weight_all)
		for (uint32_t i = 0; i < x->total_seq; ++i) {
			w[i] = weight_all[i];
			if (w[i] != 1.0f) x->unit_weights = false;
		}
	CK(x->d_weight.ensure(w.size() * 4));
	CK(cudaMemcpyAsync(x->d_weight.p, w.data(), w.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
	CK(x->d_cov.ensure((size_t)max_pairs * 4));
	CK(x->d_done.ensure(8));
	CK(x->d_err.ensure(8));
	CK(cudaMemsetAsync(x->d_done.p, 0, 8, ctx->stream));
	CK(cudaMemsetAsync(x->d_err.p, 0, 8, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

/* this rank's buffer: as a raw device pointer (contexts of one process) or as a cudaIpcMemHandle_t (64 bytes, other processes) */
void *pcramp_gpu_exchange_buffer(pcramp_gpu_ctx *ctx) { return (ctx && ctx->xchg) ? ctx->xchg->buf : nullptr; }

int pcramp_gpu_exchange_ipc_handle(pcramp_gpu_ctx *ctx, void *handle64)
{
	if (!ctx) return 1;
	if (!ctx->xchg || !handle64) return fail(ctx, "pcramp_gpu_exchange_ipc_handle: no exchange / null argument");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
	CK(cudaSetDevice(ctx->device));
	cudaIpcMemHandle_t h;
	CK(cudaIpcGetMemHandle(&h, ctx->xchg->buf));
	memcpy(handle64, &h, 64);
	return 0;
}

/* peers: world entries.  from_ipc = 1: world x 64 bytes of handles gathered from all ranks (own entry ignored);
 * from_ipc = 0: world device pointers (void *) valid in this process. */
int pcramp_gpu_exchange_connect(pcramp_gpu_ctx *ctx, const void *peers, int from_ipc)
{
	using namespace pcr::xchg;
	if (!ctx) return 1;
	pcramp_gpu_xchg *x = ctx->xchg;
	if (!x || !peers) return fail(ctx, "pcramp_gpu_exchange_connect: no exchange / null argument");
	CK(cudaSetDevice(ctx->device));
	for (uint32_t s = 0; s < x->world; ++s) {
		if (s == x->rank) { x->peers.base[s] = (uint32_t *)x->buf; continue; }
		if (from_ipc) {
			cudaIpcMemHandle_t h;
			memcpy(&h, (const char *)peers + 64 * (size_t)s, 64);
			void *p = nullptr;
			CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
			x->ipc_opened.push_back(p);
			x->peers.base[s] = (uint32_t *)p;
		} else {
			void *p = ((void *const *)peers)[s];
			if (!p) return fail(ctx, "pcramp_gpu_exchange_connect: null peer pointer");
			cudaPointerAttributes at;
			CK(cudaPointerGetAttributes(&at, p));
			if (at.device != ctx->device) {
				int can = 0;
				CK(cudaDeviceCanAccessPeer(&can, ctx->device, at.device));
				if (!can) return fail(ctx, "pcramp_gpu_exchange_connect: no peer access between the two devices");
				const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
				(void)cudaGetLastError();
			}
			x->peers.base[s] = (uint32_t *)p;
		}
	}
	x->connected = true;
	return 0;
}

/* After pcramp_gpu_score_pairs_staged on `kind` (this rank's shard): push, wait for every rank, coverage over all sequences.
 * Asynchronous on the library's stream. */
int pcramp_gpu_exchange_step(pcramp_gpu_ctx *ctx, int kind)
{
	using namespace pcr::xchg;
	if (check_kind(ctx, kind)) return 1;
	pcramp_gpu_xchg *x = ctx->xchg;
	if (!x || !x->connected) return fail(ctx, "pcramp_gpu_exchange_step: exchange not created / connected");
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[kind];
	const uint32_t n_pairs = ctx->n_pairs;
	if (n_pairs > x->max_pairs) return fail(ctx, "pcramp_gpu_exchange_step: batch larger than max_pairs");
	if (s.n != x->shard_lo[x->rank + 1] - x->shard_lo[x->rank]) return fail(ctx, "pcramp_gpu_exchange_step: the collection is not this rank's shard");
	cudaStream_t st = ctx->stream;
	const uint32_t words_local = (s.n + 31u) / 32u;
	if (ctx->res_words != words_local) return fail(ctx, "pcramp_gpu_exchange_step: no staged scoring result for this collection");
	x->step += 1;
	x->last_pairs = n_pairs;
	const uint32_t b = x->step & 1u;
	const size_t any_off = (size_t)(2 * b) * x->plane_words, p1_off = any_off + x->plane_words;
	const uint64_t total = (uint64_t)n_pairs * words_local;
	if (n_pairs) {
		const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((total + 255) / 256, (uint64_t)ctx->sm_count * 4));
		push_kernel<<<grid, 256, 0, st>>>(ctx->d_bits.as<uint32_t>(), x->unit_weights ? nullptr : ctx->d_bits1.as<uint32_t>(), n_pairs, words_local,
			x->words_global, x->shard_lo[x->rank] / 32u, x->peers, x->world, any_off, p1_off, x->flag_off_words, x->rank, x->step,
			x->d_done.as<unsigned int>());
		CK(cudaGetLastError());
		wait_kernel<<<1, 32, 0, st>>>((const uint32_t *)x->buf + x->flag_off_words, x->world, x->step, x->d_err.as<unsigned int>());
		CK(cudaGetLastError());
		const uint32_t *g_any = (const uint32_t *)x->buf + any_off, *g_p1 = (const uint32_t *)x->buf + p1_off;
		if (x->unit_weights)
			coverage_count_kernel<<<grid_for(32ull * n_pairs, 256), 256, 0, st>>>(g_any, n_pairs, x->words_global, x->d_cov.as<float>());
		else
			coverage_kernel<<<grid_for(n_pairs, 128), 128, 0, st>>>(g_any, g_p1, x->d_weight.as<float>(), n_pairs, x->words_global, x->total_seq,
				x->d_cov.as<float>());
		CK(cudaGetLastError());
		ctx->stats.kernel_launches += 3;
	}
	return 0;
}

/* device pointers of the merged result of the last step: coverage float[n_pairs], bitsets uint32[n_pairs x words_global] */
void *pcramp_gpu_exchange_coverage(pcramp_gpu_ctx *ctx) { return (ctx && ctx->xchg) ? ctx->xchg->d_cov.p : nullptr; }
void *pcramp_gpu_exchange_bitsets(pcramp_gpu_ctx *ctx)
{
	if (!ctx || !ctx->xchg) return nullptr;
	pcramp_gpu_xchg *x = ctx->xchg;
	return (uint32_t *)x->buf + (size_t)(2 * (x->step & 1u)) * x->plane_words;
}
uint32_t pcramp_gpu_exchange_words(pcramp_gpu_ctx *ctx) { return (ctx && ctx->xchg) ? ctx->xchg->words_global : 0; }

/* host copies of the merged result (synchronises the stream; reports a peer that never arrived) */
int pcramp_gpu_exchange_fetch(pcramp_gpu_ctx *ctx, float *coverage, uint32_t *bitsets)
{
	if (!ctx) return 1;
	pcramp_gpu_xchg *x = ctx->xchg;
	if (!x) return fail(ctx, "pcramp_gpu_exchange_fetch: no exchange");
	CK(cudaSetDevice(ctx->device));
	unsigned int err = 0;
	CK(cudaMemcpyAsync(&err, x->d_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (coverage && x->last_pairs) CK(cudaMemcpyAsync(coverage, x->d_cov.p, (size_t)x->last_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (bitsets && x->last_pairs)
		CK(cudaMemcpyAsync(bitsets, pcramp_gpu_exchange_bitsets(ctx), (size_t)x->last_pairs * x->words_global * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (err) {
		char b[128];
		snprintf(b, sizeof(b), "pcramp_gpu_exchange: timed out waiting for rank mask 0x%x", err);
		return fail(ctx, b);
	}
	return 0;
}

int pcramp_gpu_exchange_destroy(pcramp_gpu_ctx *ctx)
{
	if (!ctx || !ctx->xchg) return 0;
	pcramp_gpu_xchg *x = ctx->xchg;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	for (void *p : x->ipc_opened) cudaIpcCloseMemHandle(p);
	if (x->buf) cudaFree(x->buf);
	delete x;
	ctx->xchg = nullptr;
	return 0;
}

} // extern "C"
