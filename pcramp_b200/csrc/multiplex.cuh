// multiplex.cuh -- the multiplex terms of optimize() and its moves (included at the end of pcramp_gpu.cu):
//
//   multiplex background database   main.cpp:989-1003: every amplicon of the assays already chosen is pack()ed WHOLE (no
//                                   select_words, no GC filter) into multiplex_background_db; keys() of it
//   collect_multiplex_background_candidates   pcr_assay.cpp:71-104: per oligo, the key indices with
//                                   (oligo & key) >= unsigned(size * background_threshold)  (match_words, optimize.cpp:291-301)
//   update_multiplex_background_candidates    assay.h:449-453 -> update_identity (optimize.cpp:209-261)
//   compute_multiplex_background_coverage     pcr_assay.cpp:304-336: number of keys whose F- or R-identity reaches the threshold
//   compute_oligo_overlap / the overlap terms of the moves     pcr_assay.cpp:736-754, optimize_pcr.cpp, Word::max_overlap word.h:38-92
//
// As for the other databases, a move evaluates the TRIAL oligo's identities on the key lists collected for the UNMOVED
// assay, so the work is split the same way: one brute-force pass (unique base oligos x keys, integer-issue bound, the
// same 4 LOP3 + POPC compare as Word::operator&) leaves a sorted item list (oligo, key); then one warp per variant
// walks the two lists of its base assay.
#pragma once
#include "ctx.cuh"
#include "score.cuh"

#include <thrust/execution_policy.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include <unordered_map>

namespace pcr {
namespace mpx {

struct WKey {
	uint64_t hi, lo;
};
struct WKeyLess { // __word::operator< (word.h:197-211): lexicographic on (buffer[0], buffer[1])
	__host__ __device__ bool operator()(const WKey &a, const WKey &b) const { return a.hi != b.hi ? a.hi < b.hi : a.lo < b.lo; }
};
struct WKeyEq {
	__host__ __device__ bool operator()(const WKey &a, const WKey &b) const { return a.hi == b.hi && a.lo == b.lo; }
};

// Sequence::pack of EVERY sequence of the set (both strands), words only
__global__ void pack_words_all_kernel(SeqDev sd, PackParams pp, WKey *out, unsigned long long *n_out)
{
	for (uint32_t seq = blockIdx.y; seq < sd.n; seq += gridDim.y) {
		const uint32_t Lc = sd.clen[seq];
		const uint32_t n_full = Lc >= 32u ? Lc - 31u : 0u;
		const EdgeCounts ec = edge_counts(sd, seq, pp);
		const uint32_t total = n_full + ec.n_fill + ec.n_eos + ec.n_tail;
		for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
			uint32_t type, pos;
			if (e < n_full) { type = ENT_FULL; pos = 31u + e; }
			else {
				const uint32_t d = e - n_full;
				if (d < ec.n_fill) { type = ENT_FILL; pos = d; }
				else if (d < ec.n_fill + ec.n_eos) { type = ENT_EOSEVT; pos = sd.eos_pos[sd.eos_off[seq] + (d - ec.n_fill)]; }
				else { type = ENT_TAIL; pos = d - ec.n_fill - ec.n_eos + 1u; }
			}
			W128 wp, wm;
			int lp, lm;
			if (!pack_entry(sd, seq, type, pos, pp, wp, wm, lp, lm)) continue;
			const unsigned long long o = atomicAdd(n_out, 2ull);
			out[o].hi = wp.hi; out[o].lo = wp.lo;
			out[o + 1].hi = wm.hi; out[o + 1].lo = wm.lo;
		}
	}
}

__global__ void key_planes_kernel(const WKey *__restrict__ k, uint64_t n, uint4 *planes)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	W128 w;
	w.hi = k[i].hi; w.lo = k[i].lo;
	const Planes4 p = w_planes(w);
	planes[i] = make_uint4(p.a, p.c, p.g, p.t);
}

// per-oligo constants of a plain word list (score.cuh prep_oligos_kernel takes F / R arrays)
__global__ void prep_words_kernel(const uint64_t *__restrict__ words, uint32_t n, float thr, OligoDev *out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	W128 w;
	w.hi = words[2 * i];
	w.lo = words[2 * i + 1];
	out[i] = make_oligo(w, thr);
}

constexpr int MPX_CHUNK = 512; // oligos per shared-memory chunk (12 KB)

// match_words of every unique base oligo against every key: item = oligo << 32 | key
__global__ void __launch_bounds__(256) collect_kernel(const uint4 *__restrict__ key_planes, uint32_t n_keys, const OligoDev *__restrict__ base,
	uint32_t n_base, unsigned long long *items, unsigned long long *count, unsigned long long cap)
{
	__shared__ OligoDev s_o[MPX_CHUNK];
	for (uint32_t c0 = 0; c0 < n_base; c0 += MPX_CHUNK) {
		const uint32_t nc = min((uint32_t)MPX_CHUNK, n_base - c0);
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) s_o[i] = base[c0 + i];
		__syncthreads();
		for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_keys; k += gridDim.x * blockDim.x) {
			const uint4 kp = key_planes[k];
			for (uint32_t u = 0; u < nc; ++u) {
				const OligoDev &o = s_o[u];
				const uint32_t cnt = (uint32_t)__popc((o.a & kp.x) | (o.c & kp.y) | (o.g & kp.z) | (o.t & kp.w));
				if (cnt >= (o.packed & 255u)) {
					const unsigned long long at = atomicAdd(count, 1ull);
					if (at < cap) items[at] = ((unsigned long long)(c0 + u) << 32) | k;
				}
			}
		}
	}
}

__global__ void item_offsets_kernel(const unsigned long long *__restrict__ items, uint64_t n_items, uint32_t n_base, uint32_t *off)
{
	const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
	if (u > n_base) return;
	const unsigned long long want = (unsigned long long)u << 32;
	uint64_t lo = 0, hi = n_items;
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		if (items[mid] < want) lo = mid + 1; else hi = mid;
	}
	off[u] = (uint32_t)lo;
}

__device__ __forceinline__ bool key_passes(const OligoDev &o, const uint4 kp, float thr, int taq)
{
	ScoreEntry e;
	e.a = kp.x; e.c = kp.y; e.g = kp.z; e.t = kp.w;
	e.loc = 0; e.strand = 0;
	return oligo_identity(o, oligo_count(o, e), e, taq) >= thr;
}

// compute_multiplex_background_coverage after update_identity with the trial oligos: one warp per variant.
// bidx[2v], bidx[2v+1] = the unique-base index of the unmoved F and R; var[2v], var[2v+1] = the trial oligos.
__global__ void __launch_bounds__(256) variant_coverage_kernel(const uint4 *__restrict__ key_planes, const unsigned long long *__restrict__ items,
	const uint32_t *__restrict__ off, const uint32_t *__restrict__ bidx, const OligoDev *__restrict__ var, uint32_t n, float thr, int taq,
	float *cov)
{
	const uint32_t v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
	if (v >= n) return;
	const OligoDev F = var[2 * v], R = var[2 * v + 1];
	const uint32_t f0 = off[bidx[2 * v]], f1 = off[bidx[2 * v] + 1], r0 = off[bidx[2 * v + 1]], r1 = off[bidx[2 * v + 1] + 1];
	uint32_t total = 0;
	for (uint32_t i = f0 + lane; i < f1; i += 32u) {
		const uint32_t k = (uint32_t)items[i];
		total += key_passes(F, key_planes[k], thr, taq) ? 1u : 0u;
	}
	for (uint32_t i = r0 + lane; i < r1; i += 32u) {
		const uint32_t k = (uint32_t)items[i];
		const uint4 kp = key_planes[k];
		if (!key_passes(R, kp, thr, taq)) continue;
		// already counted through the forward list?  (SET<unsigned int> valid, pcr_assay.cpp:313-333)
		uint32_t lo = f0, hi = f1;
		while (lo < hi) {
			const uint32_t mid = (lo + hi) >> 1;
			if ((uint32_t)items[mid] < k) lo = mid + 1; else hi = mid;
		}
		if (lo < f1 && (uint32_t)items[lo] == k && key_passes(F, kp, thr, taq)) continue;
		total += 1u;
	}
	for (int s = 16; s > 0; s >>= 1) total += __shfl_xor_sync(0xffffffffu, total, s);
	if (lane == 0) cov[v] = (float)(double)total; // double ret += 1.0 per key, returned as float
}

// best[i] = max over the pool oligos of words[i].max_overlap(pool oligo)   (0 for an empty pool)
__global__ void pool_overlap_kernel(const uint64_t *__restrict__ words, uint32_t n, const uint64_t *__restrict__ pool, uint32_t n_pool_oligos,
	float *best)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	W128 w;
	w.hi = words[2 * i];
	w.lo = words[2 * i + 1];
	float b = 0.0f;
	for (uint32_t p = 0; p < n_pool_oligos; ++p) {
		W128 s;
		s.hi = pool[2 * p];
		s.lo = pool[2 * p + 1];
		b = fmaxf(b, w_max_overlap(w, s)); // std::max(best, x): a NaN x is never taken, as with fmaxf
	}
	best[i] = b;
}

struct W128Hash {
	size_t operator()(const std::pair<uint64_t, uint64_t> &k) const { return (size_t)(k.first * 0x9E3779B97F4A7C15ull ^ (k.second + (k.first >> 29))); }
};

// coverage[i] of n variants; host arrays of n x 2 uint64
inline int coverage_run(pcramp_gpu_ctx *ctx, const uint64_t *base_f, const uint64_t *base_r, const uint64_t *var_f, const uint64_t *var_r,
	uint32_t n, float threshold, int taq, float *coverage, uint64_t &launches)
{
	if (!ctx->mpx_valid) return fail(ctx, "multiplex coverage: no multiplex key list (call pcramp_gpu_multiplex_keys first)");
	if (n == 0) return 0;
	if (ctx->mpx_n_keys == 0) { // collect_multiplex_background_candidates returns at once on an empty key list
		for (uint32_t i = 0; i < n; ++i) coverage[i] = 0.0f;
		return 0;
	}
	cudaStream_t st = ctx->stream;
	// unique base oligos
	std::unordered_map<std::pair<uint64_t, uint64_t>, uint32_t, W128Hash> ids;
	std::vector<uint64_t> uniq;
	std::vector<uint32_t> bidx(2ull * n);
	auto id_of = [&](const uint64_t *w) {
		const std::pair<uint64_t, uint64_t> k(w[0], w[1]);
		auto it = ids.find(k);
		if (it != ids.end()) return it->second;
		const uint32_t id = (uint32_t)ids.size();
		ids.emplace(k, id);
		uniq.push_back(w[0]);
		uniq.push_back(w[1]);
		return id;
	};
	std::vector<uint64_t> var(4ull * n);
	for (uint32_t i = 0; i < n; ++i) {
		bidx[2 * i] = id_of(base_f + 2 * i);
		bidx[2 * i + 1] = id_of(base_r + 2 * i);
		var[4 * i] = var_f[2 * i]; var[4 * i + 1] = var_f[2 * i + 1];
		var[4 * i + 2] = var_r[2 * i]; var[4 * i + 3] = var_r[2 * i + 1];
	}
	const uint32_t nb = (uint32_t)ids.size();
	const uint32_t nk = (uint32_t)ctx->mpx_n_keys;
	DevBuf d_uniq, d_var;
	CK(d_uniq.ensure(uniq.size() * 8));
	CK(d_var.ensure(var.size() * 8));
	CK(ctx->mpx_base.ensure((size_t)nb * sizeof(OligoDev)));
	CK(ctx->mpx_var.ensure(2ull * n * sizeof(OligoDev)));
	CK(ctx->mpx_bidx.ensure(2ull * n * 4));
	CK(ctx->mpx_item_off.ensure(((size_t)nb + 1) * 4));
	CK(ctx->mpx_cov.ensure((size_t)n * 4));
	CK(ctx->d_counters.ensure(8 * sizeof(unsigned long long)));
	CK(cudaMemcpyAsync(d_uniq.p, uniq.data(), uniq.size() * 8, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_var.p, var.data(), var.size() * 8, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(ctx->mpx_bidx.p, bidx.data(), bidx.size() * 4, cudaMemcpyHostToDevice, st));
	prep_words_kernel<<<grid_for(nb, 256), 256, 0, st>>>(d_uniq.as<uint64_t>(), nb, threshold, ctx->mpx_base.as<OligoDev>());
	prep_words_kernel<<<grid_for(2ull * n, 256), 256, 0, st>>>(d_var.as<uint64_t>(), 2u * n, threshold, ctx->mpx_var.as<OligoDev>());
	CK(cudaGetLastError());
	launches += 2;
	unsigned long long n_items = 0;
	CK(ctx->mpx_items.ensure(std::max<size_t>(1 << 20, ctx->mpx_items.cap)));
	for (int attempt = 0;; ++attempt) {
		const unsigned long long cap = ctx->mpx_items.cap / 8 / 2; // second half = sort buffer
		CK(cudaMemsetAsync(ctx->d_counters.p, 0, 8, st));
		const unsigned grid = std::min<unsigned>(grid_for(nk, 256), (unsigned)ctx->sm_count * 8u);
		collect_kernel<<<grid, 256, 0, st>>>(ctx->mpx_planes.as<uint4>(), nk, ctx->mpx_base.as<OligoDev>(), nb,
			ctx->mpx_items.as<unsigned long long>(), ctx->d_counters.as<unsigned long long>(), cap);
		CK(cudaGetLastError());
		launches += 1;
		CK(cudaMemcpyAsync(&n_items, ctx->d_counters.p, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		if (n_items <= cap) break;
		if (attempt >= 2) return fail(ctx, "multiplex coverage: item buffer kept overflowing");
		CK(ctx->mpx_items.ensure((size_t)(n_items + n_items / 8 + 1024) * 16));
	}
	if (n_items >= (1ull << 32)) return fail(ctx, "multiplex coverage: more than 2^32 (oligo, key) candidates");
	unsigned long long *d_items = ctx->mpx_items.as<unsigned long long>();
	if (n_items) {
		unsigned long long *d_alt = d_items + ctx->mpx_items.cap / 8 / 2;
		cub::DoubleBuffer<unsigned long long> db(d_items, d_alt);
		size_t tb = 0;
		const int end_bit = 32 + (int)bits_for((uint64_t)nb + 1);
		CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, (int64_t)n_items, 0, end_bit, st));
		CK(ctx->cub_tmp.ensure(tb));
		CK(cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tb, db, (int64_t)n_items, 0, end_bit, st));
		launches += 8;
		d_items = db.Current();
	}
	item_offsets_kernel<<<grid_for((uint64_t)nb + 1, 256), 256, 0, st>>>(d_items, n_items, nb, ctx->mpx_item_off.as<uint32_t>());
	variant_coverage_kernel<<<grid_for(32ull * n, 256), 256, 0, st>>>(ctx->mpx_planes.as<uint4>(), d_items, ctx->mpx_item_off.as<uint32_t>(),
		ctx->mpx_bidx.as<uint32_t>(), ctx->mpx_var.as<OligoDev>(), n, threshold, taq, ctx->mpx_cov.as<float>());
	CK(cudaGetLastError());
	launches += 2;
	CK(cudaMemcpyAsync(coverage, ctx->mpx_cov.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	return 0;
}

// best[i] = max over the pool of max_overlap(words[i], .)
inline int overlap_run(pcramp_gpu_ctx *ctx, const uint64_t *words, uint32_t n, float *best, uint64_t &launches)
{
	if (n == 0) return 0;
	const uint32_t np = (uint32_t)(ctx->pool_words.size() / 2);
	if (np == 0) {
		for (uint32_t i = 0; i < n; ++i) best[i] = 0.0f;
		return 0;
	}
	cudaStream_t st = ctx->stream;
	CK(ctx->mpx_ov_words.ensure((size_t)n * 16));
	CK(ctx->mpx_ov.ensure((size_t)n * 4));
	CK(cudaMemcpyAsync(ctx->mpx_ov_words.p, words, (size_t)n * 16, cudaMemcpyHostToDevice, st));
	pool_overlap_kernel<<<grid_for(n, 128), 128, 0, st>>>(ctx->mpx_ov_words.as<uint64_t>(), n, ctx->mpx_pool.as<uint64_t>(), np, ctx->mpx_ov.as<float>());
	CK(cudaGetLastError());
	launches += 1;
	CK(cudaMemcpyAsync(best, ctx->mpx_ov.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	return 0;
}

} // namespace mpx
} // namespace pcr

extern "C" {

int pcramp_gpu_multiplex_keys(pcramp_gpu_ctx *ctx, uint32_t pack_max_degen, uint32_t min_len, uint64_t *n_keys)
{
	using namespace pcr::mpx;
	if (!ctx) return 1;
	if (min_len == 0) return fail(ctx, "pcramp_gpu_multiplex_keys: min_oligo_length must be >= 1");
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[PCRAMP_MULTIPLEX];
	cudaStream_t st = ctx->stream;
	ctx->mpx_valid = false;
	ctx->mpx_n_keys = 0;
	if (s.n) {
		PackParams pp;
		pp.max_degen = pack_max_degen;
		pp.min_gc = 0.0f; // main.cpp:993: "Don't G+C filter the multiplex background sequences"
		pp.max_gc = 1.0f;
		pp.min_len = min_len;
		pp.gc_filter = false;
		uint64_t worst = 0;
		uint32_t longest = 0;
		for (uint32_t i = 0; i < s.n; ++i) {
			worst += 2ull * ((uint64_t)s.plen[i] + 64ull + s.eos[i].size());
			longest = std::max(longest, s.plen[i]);
		}
		DevBuf raw;
		CK(raw.ensure(worst * sizeof(WKey)));
		CK(ctx->d_counters.ensure(8 * sizeof(unsigned long long)));
		CK(cudaMemsetAsync(ctx->d_counters.p, 0, 8, st));
		const dim3 grid(std::max(1u, std::min(grid_for((uint64_t)longest + 64, 256), 64u)), std::min(s.n, 65535u));
		pack_words_all_kernel<<<grid, 256, 0, st>>>(s.dev(), pp, raw.as<WKey>(), ctx->d_counters.as<unsigned long long>());
		CK(cudaGetLastError());
		unsigned long long n = 0;
		CK(cudaMemcpyAsync(&n, ctx->d_counters.p, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		if (n >= (1ull << 32)) return fail(ctx, "pcramp_gpu_multiplex_keys: more than 2^32 words");
		if (n) {
			WKey *p = raw.as<WKey>();
			CachedScratch scratch;
			thrust::sort(thrust::cuda::par(scratch).on(st), p, p + n, WKeyLess());
			WKey *e = thrust::unique(thrust::cuda::par(scratch).on(st), p, p + n, WKeyEq());
			CK(cudaStreamSynchronize(st));
			const uint64_t nk = (uint64_t)(e - p);
			CK(ctx->mpx_words.ensure(nk * sizeof(WKey)));
			CK(ctx->mpx_planes.ensure(nk * sizeof(uint4)));
			CK(cudaMemcpyAsync(ctx->mpx_words.p, p, nk * sizeof(WKey), cudaMemcpyDeviceToDevice, st));
			key_planes_kernel<<<grid_for(nk, 256), 256, 0, st>>>(ctx->mpx_words.as<WKey>(), nk, ctx->mpx_planes.as<uint4>());
			CK(cudaGetLastError());
			CK(cudaStreamSynchronize(st));
			ctx->mpx_n_keys = nk;
		}
	}
	ctx->mpx_valid = true;
	if (n_keys) *n_keys = ctx->mpx_n_keys;
	return 0;
}

int pcramp_gpu_multiplex_keys_copy(pcramp_gpu_ctx *ctx, uint64_t *words)
{
	if (!ctx) return 1;
	if (!ctx->mpx_valid) return fail(ctx, "pcramp_gpu_multiplex_keys_copy: no key list");
	CK(cudaSetDevice(ctx->device));
	if (ctx->mpx_n_keys) {
		CK(cudaMemcpyAsync(words, ctx->mpx_words.p, ctx->mpx_n_keys * 16, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	return 0;
}

int pcramp_gpu_set_pool(pcramp_gpu_ctx *ctx, const uint64_t *pool_f, const uint64_t *pool_r, uint32_t n_pool)
{
	if (!ctx) return 1;
	if (n_pool && (!pool_f || !pool_r)) return fail(ctx, "pcramp_gpu_set_pool: null argument");
	CK(cudaSetDevice(ctx->device));
	ctx->pool_words.clear();
	for (uint32_t i = 0; i < n_pool; ++i) {
		ctx->pool_words.push_back(pool_f[2 * i]);
		ctx->pool_words.push_back(pool_f[2 * i + 1]);
		ctx->pool_words.push_back(pool_r[2 * i]);
		ctx->pool_words.push_back(pool_r[2 * i + 1]);
	}
	if (n_pool) {
		CK(ctx->mpx_pool.ensure(ctx->pool_words.size() * 8));
		CK(cudaMemcpyAsync(ctx->mpx_pool.p, ctx->pool_words.data(), ctx->pool_words.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	return 0;
}

int pcramp_gpu_multiplex_coverage(pcramp_gpu_ctx *ctx, const uint64_t *base_f, const uint64_t *base_r, const uint64_t *var_f,
	const uint64_t *var_r, uint32_t n, float threshold, int use_taq_mama, float *coverage)
{
	if (!ctx) return 1;
	if (n && (!base_f || !base_r || !var_f || !var_r || !coverage)) return fail(ctx, "pcramp_gpu_multiplex_coverage: null argument");
	CK(cudaSetDevice(ctx->device));
	uint64_t launches = 0;
	const int rc = pcr::mpx::coverage_run(ctx, base_f, base_r, var_f, var_r, n, threshold, use_taq_mama, coverage, launches);
	ctx->stats.kernel_launches = launches;
	return rc;
}

int pcramp_gpu_oligo_overlap(pcramp_gpu_ctx *ctx, const uint64_t *f, const uint64_t *r, uint32_t n_pairs, float *overlap)
{
	if (!ctx) return 1;
	if (n_pairs && (!f || !r || !overlap)) return fail(ctx, "pcramp_gpu_oligo_overlap: null argument");
	CK(cudaSetDevice(ctx->device));
	std::vector<uint64_t> words(4ull * n_pairs);
	for (uint32_t i = 0; i < n_pairs; ++i) {
		words[4 * i] = f[2 * i]; words[4 * i + 1] = f[2 * i + 1];
		words[4 * i + 2] = r[2 * i]; words[4 * i + 3] = r[2 * i + 1];
	}
	std::vector<float> best(2ull * n_pairs);
	uint64_t launches = 0;
	if (pcr::mpx::overlap_run(ctx, words.data(), 2u * n_pairs, best.data(), launches)) return 1;
	for (uint32_t i = 0; i < n_pairs; ++i) { // pcr_assay.cpp:752-753
		const float bf = best[2 * i], br = best[2 * i + 1];
		overlap[i] = ((bf == 1.0f) ? 10.0f : bf) + ((br == 1.0f) ? 10.0f : br);
	}
	ctx->stats.kernel_launches = launches;
	return 0;
}

} // extern "C"
