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
//
// Failure model.  A rank that never arrives is detected by wait_kernel after a configurable bound (default 120 s:
// pcramp_gpu_exchange_set_timeout_ms; ranks legitimately skew by seconds -- a first-step index build, cudaMalloc growth, rank-0
// I/O).  On a timeout the error word is set, the step's coverage is POISONED (NaN in every slot, so a consumer of the device
// pointers cannot mistake a partial merge for a result), and -- the double-buffer invariant being gone -- every later
// exchange_step / reduce_best on this context fails once the host has seen the error (exchange_status, exchange_fetch): the
// exchange must be destroyed on all ranks and created again.  Ranks must not share a device (a spinning wait_kernel blocks
// device-synchronising calls such as cudaMalloc / cudaFree of other contexts on the same device): allocate scratch before
// the first step where contexts of one process play the ranks (the tests).
//
// reduce_best (main.cpp:1421-1601, rule :1455-1480): every rank stores its 32-byte winner record {Score, total degeneracy,
// global trial index} into a slot of every peer's buffer, publishes a second flag, waits for all ranks and folds the records
// in rank order with the reference's rule -- ONE single-CTA kernel, no NCCL, no host staging.
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

constexpr uint32_t REC_WORDS = 8;    // a winner record: {target, background, overlap (float), valid, degeneracy (double), trial (int64)}

struct State {
	uint32_t rank = 0, world = 0, max_pairs = 0, total_seq = 0, words_global = 0;
	std::vector<uint32_t> shard_lo; // world + 1 sequence offsets
	// this rank's buffer: [2][any | pass1][max_pairs][words_global] words, flags[MAX_WORLD][FLAG_STRIDE], record flags[MAX_WORLD][FLAG_STRIDE],
	// records[2][MAX_WORLD][REC_WORDS]
	void *buf = nullptr;
	size_t buf_bytes = 0, plane_words = 0, flag_off_words = 0, rflag_off_words = 0, rec_off_words = 0;
	Peers peers = {};
	bool connected = false, failed = false;
	std::vector<void *> ipc_opened;
	DevBuf d_weight, d_cov, d_done, d_err, d_best;
	bool unit_weights = true;
	uint32_t step = 0, last_pairs = 0, rstep = 0;
	unsigned long long timeout_cycles = 240000000000ull; // ~120 s at 2 GHz (clock64 ticks at the SM clock)
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

// thread s (s < world) polls the flag of rank s until it reaches `step`; false after `timeout` SM cycles
__device__ __forceinline__ bool wait_flag(const uint32_t *f, uint32_t step, unsigned long long timeout)
{
	const long long t0 = clock64();
	for (;;) {
		const uint32_t v = ld_volatile(f);
		if ((int32_t)(v - step) >= 0) return true;
		if ((unsigned long long)(clock64() - t0) > timeout) return false;
		__nanosleep(200);
	}
}

// one CTA; thread s waits for rank s.  Bounded: after `timeout` cycles the error word is set (bit s = rank s missing).
__global__ void wait_kernel(const uint32_t *flags, uint32_t world, uint32_t step, unsigned long long timeout, unsigned int *err)
{
	if (threadIdx.x < world && !wait_flag(flags + (size_t)threadIdx.x * FLAG_STRIDE, step, timeout)) atomicOr(err, 1u << threadIdx.x);
	__threadfence_system();
}

// after a timeout the merged buffer is partial: no consumer of the device pointers may take its coverage for a result
__global__ void poison_kernel(const unsigned int *err, float *cov, uint32_t n)
{
	if (*err == 0u) return;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) cov[i] = __int_as_float(0x7fc00000);
}

struct BestRecord {
	float target, background, overlap;
	uint32_t valid;
	double degeneracy;
	long long trial;
};
static_assert(sizeof(BestRecord) == REC_WORDS * 4, "a winner record is 32 bytes");

// Score::operator< / == (pcramp.h:180-201) on records
__device__ __forceinline__ bool score_less(const BestRecord &a, const BestRecord &b)
{
	const float x = a.target - a.background, y = b.target - b.background;
	return (x == y) ? (a.overlap < b.overlap) : (x < y);
}
__device__ __forceinline__ bool score_equal(const BestRecord &a, const BestRecord &b)
{
	return (a.target - a.background) == (b.target - b.background) && a.overlap == b.overlap;
}

// reduce_best_assay: push this rank's record to every rank, wait for all, fold in rank order with the rule of main.cpp:1455-1480
// (a later record replaces the running best unless its Score is lower, or equal with a total degeneracy that is not smaller).
// out: {owner rank, 0, record}
__global__ void reduce_best_kernel(BestRecord mine, Peers peers, uint32_t world, uint32_t my_rank, uint32_t rstep, size_t rec_off, size_t rflag_off,
	unsigned long long timeout, unsigned int *err, uint32_t *out)
{
	const uint32_t t = threadIdx.x;
	const size_t slot = rec_off + ((size_t)(rstep & 1u) * MAX_WORLD + my_rank) * REC_WORDS;
	if (t < world) {
		const uint32_t *src = (const uint32_t *)&mine;
		for (uint32_t k = 0; k < REC_WORDS; ++k) *(volatile uint32_t *)(peers.base[t] + slot + k) = src[k];
		__threadfence_system();
		*(volatile uint32_t *)(peers.base[t] + rflag_off + (size_t)my_rank * FLAG_STRIDE) = rstep;
		if (!wait_flag(peers.base[my_rank] + rflag_off + (size_t)t * FLAG_STRIDE, rstep, timeout)) atomicOr(err, 1u << t);
	}
	__threadfence_system();
	__syncthreads();
	if (t == 0) {
		const uint32_t *recs = peers.base[my_rank] + rec_off + (size_t)(rstep & 1u) * MAX_WORLD * REC_WORDS;
		BestRecord best;
		uint32_t owner = 0;
		for (uint32_t k = 0; k < REC_WORDS; ++k) ((uint32_t *)&best)[k] = ld_volatile(recs + k);
		for (uint32_t s = 1; s < world; ++s) {
			BestRecord r;
			for (uint32_t k = 0; k < REC_WORDS; ++k) ((uint32_t *)&r)[k] = ld_volatile(recs + (size_t)s * REC_WORDS + k);
			if (score_less(r, best)) continue;
			if (score_equal(r, best) && best.degeneracy <= r.degeneracy) continue;
			best = r;
			owner = s;
		}
		out[0] = owner;
		out[1] = *err;
		for (uint32_t k = 0; k < REC_WORDS; ++k) out[2 + k] = ((const uint32_t *)&best)[k];
	}
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
	// re-creating under connected peers would free a buffer they still store into and restart the step count under their flags
	if (ctx->xchg) return fail(ctx, "pcramp_gpu_exchange_create: an exchange exists; pcramp_gpu_exchange_destroy it on ALL ranks first");
	for (uint32_t s = 0; s + 1 < world; ++s)
		if ((shard_nseq[s] % 32u) != 0u)
			return fail(ctx, "pcramp_gpu_exchange_create: every shard but the last must hold a multiple of 32 sequences (disjoint bitset words)");
	pcramp_gpu_xchg *x = new pcramp_gpu_xchg();
	ctx->xchg = x;
	x->rank = rank;
	x->world = world;
	x->max_pairs = max_pairs;
	x->shard_lo.assign(world + 1, 0);
	for (uint32_t s = 0; s < world; ++s) x->shard_lo[s + 1] = x->shard_lo[s] + shard_nseq[s];
	x->total_seq = x->shard_lo[world];
	x->words_global = (x->total_seq + 31u) / 32u;
	x->plane_words = (size_t)max_pairs * std::max<uint32_t>(1, x->words_global);
	x->flag_off_words = 4 * x->plane_words;
	x->rflag_off_words = x->flag_off_words + (size_t)MAX_WORLD * FLAG_STRIDE;
	x->rec_off_words = x->rflag_off_words + (size_t)MAX_WORLD * FLAG_STRIDE;
	x->buf_bytes = (x->rec_off_words + 2 * (size_t)MAX_WORLD * REC_WORDS) * 4;
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
	CK(x->d_best.ensure(64));
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
	if (x->connected) return fail(ctx, "pcramp_gpu_exchange_connect: already connected (destroy and create the exchange to reconnect)");
	CK(cudaSetDevice(ctx->device));
	// a failure half way must not leave opened handles or a half-filled peer table behind
	auto undo = [&](const std::string &why) {
		for (void *p : x->ipc_opened) cudaIpcCloseMemHandle(p);
		x->ipc_opened.clear();
		x->peers = Peers{};
		(void)cudaGetLastError();
		return fail(ctx, why);
	};
	for (uint32_t s = 0; s < x->world; ++s) {
		if (s == x->rank) { x->peers.base[s] = (uint32_t *)x->buf; continue; }
		if (from_ipc) {
			cudaIpcMemHandle_t h;
			memcpy(&h, (const char *)peers + 64 * (size_t)s, 64);
			void *p = nullptr;
			const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
			if (e != cudaSuccess) return undo(std::string("pcramp_gpu_exchange_connect: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
			x->ipc_opened.push_back(p);
			x->peers.base[s] = (uint32_t *)p;
		} else {
			void *p = ((void *const *)peers)[s];
			if (!p) return undo("pcramp_gpu_exchange_connect: null peer pointer");
			cudaPointerAttributes at;
			if (cudaPointerGetAttributes(&at, p) != cudaSuccess) return undo("pcramp_gpu_exchange_connect: a peer pointer is not device memory");
			if (at.device != ctx->device) {
				int can = 0;
				if (cudaDeviceCanAccessPeer(&can, ctx->device, at.device) != cudaSuccess || !can)
					return undo("pcramp_gpu_exchange_connect: no peer access between the two devices");
				const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
				if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
					return undo(std::string("pcramp_gpu_exchange_connect: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
				(void)cudaGetLastError();
			}
			x->peers.base[s] = (uint32_t *)p;
		}
	}
	x->connected = true;
	return 0;
}

/* the bound of every wait, in milliseconds of GPU time at the SM's maximum clock (default 120 000) */
int pcramp_gpu_exchange_set_timeout_ms(pcramp_gpu_ctx *ctx, uint32_t ms)
{
	if (!ctx) return 1;
	if (!ctx->xchg) return fail(ctx, "pcramp_gpu_exchange_set_timeout_ms: no exchange");
	cudaDeviceProp pr;
	CK(cudaGetDeviceProperties(&pr, ctx->device));
	ctx->xchg->timeout_cycles = (unsigned long long)std::max<uint32_t>(1, ms) * (unsigned long long)std::max(1, pr.clockRate);
	return 0;
}

/* synchronises the stream and reports the ranks that never arrived (bit s = rank s); a non-zero mask makes the exchange unusable */
int pcramp_gpu_exchange_status(pcramp_gpu_ctx *ctx, uint32_t *timed_out_mask)
{
	if (!ctx) return 1;
	pcramp_gpu_xchg *x = ctx->xchg;
	if (!x) return fail(ctx, "pcramp_gpu_exchange_status: no exchange");
	CK(cudaSetDevice(ctx->device));
	unsigned int err = 0;
	CK(cudaMemcpyAsync(&err, x->d_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (err) x->failed = true;
	if (timed_out_mask) *timed_out_mask = err;
	return 0;
}

uint32_t pcramp_gpu_exchange_pairs(pcramp_gpu_ctx *ctx) { return (ctx && ctx->xchg) ? ctx->xchg->last_pairs : 0; }

/* After pcramp_gpu_score_pairs_staged on `kind` (this rank's shard): push, wait for every rank, coverage over all sequences.
 * Asynchronous on the library's stream. */
int pcramp_gpu_exchange_step(pcramp_gpu_ctx *ctx, int kind)
{
	using namespace pcr::xchg;
	if (check_kind(ctx, kind)) return 1;
	pcramp_gpu_xchg *x = ctx->xchg;
	if (!x || !x->connected) return fail(ctx, "pcramp_gpu_exchange_step: exchange not created / connected");
	if (x->failed) return fail(ctx, "pcramp_gpu_exchange_step: a rank timed out earlier; destroy the exchange on all ranks and create it again");
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
		wait_kernel<<<1, 32, 0, st>>>((const uint32_t *)x->buf + x->flag_off_words, x->world, x->step, x->timeout_cycles, x->d_err.as<unsigned int>());
		CK(cudaGetLastError());
		const uint32_t *g_any = (const uint32_t *)x->buf + any_off, *g_p1 = (const uint32_t *)x->buf + p1_off;
		if (x->unit_weights)
			coverage_count_kernel<<<grid_for(32ull * n_pairs, 256), 256, 0, st>>>(g_any, n_pairs, x->words_global, x->d_cov.as<float>());
		else
			coverage_kernel<<<grid_for(n_pairs, 128), 128, 0, st>>>(g_any, g_p1, x->d_weight.as<float>(), n_pairs, x->words_global, x->total_seq,
				x->d_cov.as<float>());
		CK(cudaGetLastError());
		poison_kernel<<<4, 256, 0, st>>>(x->d_err.as<unsigned int>(), x->d_cov.as<float>(), n_pairs);
		CK(cudaGetLastError());
		ctx->stats.kernel_launches += 4;
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
		x->failed = true;
		char b[160];
		snprintf(b, sizeof(b), "pcramp_gpu_exchange: timed out waiting for rank mask 0x%x (coverage poisoned; re-create the exchange)", err);
		return fail(ctx, b);
	}
	return 0;
}

/* reduce_best_assay (main.cpp:1421-1601) for the winner of this rank's trials (pcramp_gpu_best_assay): target / background coverage
 * and oligo overlap of its Score, PCR::total_degeneracy, its GLOBAL trial index (< 0: this rank has no assay; it then competes with
 * the reference's default Score, pcramp.h:176-179).  Every rank receives the same winner: the rank that owns it, its Score, degeneracy
 * and trial index.  The fold runs in rank order with the rule of main.cpp:1455-1480 (a tie on Score and degeneracy keeps the lower
 * rank: what the root keeps when messages arrive in rank order).  Synchronises the stream. */
int pcramp_gpu_reduce_best(pcramp_gpu_ctx *ctx, float target_coverage, float background_coverage, float oligo_overlap, double degeneracy,
	int64_t global_trial, uint32_t *owner_rank, float *best_target, float *best_background, float *best_overlap, double *best_degeneracy,
	int64_t *best_trial)
{
	using namespace pcr::xchg;
	if (!ctx) return 1;
	pcramp_gpu_xchg *x = ctx->xchg;
	if (!x || !x->connected) return fail(ctx, "pcramp_gpu_reduce_best: exchange not created / connected");
	if (x->failed) return fail(ctx, "pcramp_gpu_reduce_best: a rank timed out earlier; destroy the exchange on all ranks and create it again");
	CK(cudaSetDevice(ctx->device));
	BestRecord mine;
	if (global_trial < 0) {
		mine = BestRecord{-1.0e6f, 1.0e6f, 0.0f, 0u, degeneracy, -1};   // Score() (pcramp.h:176-179)
	} else {
		mine = BestRecord{target_coverage, background_coverage, oligo_overlap, 1u, degeneracy, (long long)global_trial};
	}
	x->rstep += 1;
	reduce_best_kernel<<<1, 32, 0, ctx->stream>>>(mine, x->peers, x->world, x->rank, x->rstep, x->rec_off_words, x->rflag_off_words,
		x->timeout_cycles, x->d_err.as<unsigned int>(), x->d_best.as<uint32_t>());
	CK(cudaGetLastError());
	ctx->stats.kernel_launches += 1;
	uint32_t out[2 + REC_WORDS];
	CK(cudaMemcpyAsync(out, x->d_best.p, sizeof(out), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (out[1]) {
		x->failed = true;
		char b[160];
		snprintf(b, sizeof(b), "pcramp_gpu_reduce_best: timed out waiting for rank mask 0x%x", out[1]);
		return fail(ctx, b);
	}
	BestRecord best;
	memcpy(&best, out + 2, sizeof(best));
	if (owner_rank) *owner_rank = out[0];
	if (best_target) *best_target = best.target;
	if (best_background) *best_background = best.background;
	if (best_overlap) *best_overlap = best.overlap;
	if (best_degeneracy) *best_degeneracy = best.degeneracy;
	if (best_trial) *best_trial = best.valid ? (int64_t)best.trial : -1;
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
