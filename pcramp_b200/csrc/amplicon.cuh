// amplicon.cuh -- multiplex bookkeeping kept on the device (included at the end of pcramp_gpu.cu):
//
//   PCR::collect_unique_amplicons   pcr_assay.cpp:756-813   match_words / find_oligo_match / sort / extract_amplicon_seq for
//                                   {F+,R-} then {R+,F-}; the amplicon strings sorted and made unique; AmpliconBounds
//   PCR::extract_amplicon_seq       pcr_assay.cpp:443-542   the non-primer region between the two binding sites, padded
//   main.cpp:783-803                the amplicons of a trial assay against every primer of the assay pool
//                                   (find_multiplex_background_match on the amplicons, weighted_coverage)
//   main.cpp:989-1017               the accepted assay's amplicons appended to the multiplex background, keys() of its
//                                   database rebuilt, the targets split at begin / centre / end of every amplicon
//
// The candidate records come from amplicon_list_kernel<true> (sw_abi.cuh) in the reference's push order.  An amplicon is never
// copied out as characters on the device: a record is (sequence, first base, length) into the collection's own nibbles; the
// distinct strings of a pair are found by sorting record numbers with a comparator that walks the two regions in the
// collection (std::string order on the IUPAC letters bits_to_base gives, base_table.h:78-117), so that the unique list has the
// reference's order -- the order in which the amplicons become multiplex background sequences.
#pragma once
#include "ctx.cuh"
#include "score.cuh"
#include "sw.cuh"
#include "sw_abi.cuh"

#include <thrust/execution_policy.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

namespace pcr {
namespace amp {

// bits_to_base (base_table.h:78-117) by nibble value; index 0 (EOS) never occurs inside an amplicon
__constant__ char c_bits_to_base[16] = {'?', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N'};

__device__ __forceinline__ uint32_t region_nibble(const uint8_t *__restrict__ raw, uint64_t byte_off, uint32_t i)
{
	const uint8_t v = __ldg(raw + byte_off + (i >> 1));
	return (i & 1u) ? (v & 15u) : (uint32_t)(v >> 4);
}

// records in sorted (reference) order -> region and bounds
__global__ void region_kernel(uint64_t n, const uint64_t *__restrict__ key1_sorted, const uint32_t *__restrict__ perm,
	const uint64_t *__restrict__ key2, const OligoDev *__restrict__ oligos, uint32_t *seq, uint32_t *start, uint32_t *len, uint32_t *pair,
	uint32_t *bounds, uint32_t *flags)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t k1 = key1_sorted[i], k2 = key2[perm[i]];
	const uint32_t p = (uint32_t)(k1 >> 33), pass = (uint32_t)(k1 >> 32) & 1u;
	const int ploc = (int)(uint32_t)(k2 >> 32) - 0x40000000, mloc = (int)(uint32_t)k2 - 0x40000000;
	const OligoDev P = oligos[2 * p + pass], M = oligos[2 * p + (pass ^ 1u)]; // pass 1: R on the plus strand, F on the minus strand
	const int p_start = (int)((P.packed >> 8) & 255u), p_stop = (int)((P.packed >> 16) & 255u);
	const int m_start = (int)((M.packed >> 8) & 255u), m_stop = (int)((M.packed >> 16) & 255u);
	const int a0 = ploc + p_stop + 1 - MPX_PAD;              // pcr_assay.cpp:494-495
	const int m = (mloc - m_stop) - a0 + 2 * MPX_PAD;        // :497-499
	seq[i] = (uint32_t)k1;
	start[i] = (uint32_t)a0;
	len[i] = (uint32_t)m;
	pair[i] = p;
	const int begin = ploc + p_start, end = mloc - m_start;  // :536-538
	bounds[3 * i] = (uint32_t)k1;
	bounds[3 * i + 1] = (uint32_t)begin;
	bounds[3 * i + 2] = (uint32_t)end;
	if (begin < 0) atomicOr(flags, 1u); // AmpliconBounds() throws on begin > end as unsigned (assay.h:82-89)
}

struct Regions {
	const uint8_t *raw;
	const uint64_t *raw_off;
	const uint32_t *seq, *start, *len, *pair;
};

// three-way comparison of two amplicon strings: char by char, a proper prefix sorts first (std::string::compare)
__device__ inline int region_compare(const Regions &R, uint32_t a, uint32_t b)
{
	const uint32_t sa = R.seq[a], sb = R.seq[b], ia = R.start[a], ib = R.start[b], la = R.len[a], lb = R.len[b];
	if (sa == sb && ia == ib) return la < lb ? -1 : (la > lb ? 1 : 0);
	const uint64_t oa = R.raw_off[sa], ob = R.raw_off[sb];
	const uint32_t n = min(la, lb);
	for (uint32_t k = 0; k < n; ++k) {
		const uint32_t x = region_nibble(R.raw, oa, ia + k), y = region_nibble(R.raw, ob, ib + k);
		if (x != y) return c_bits_to_base[x] < c_bits_to_base[y] ? -1 : 1;
	}
	return la < lb ? -1 : (la > lb ? 1 : 0);
}

struct RegionLess {
	Regions R;
	__device__ bool operator()(uint32_t a, uint32_t b) const
	{
		if (R.pair[a] != R.pair[b]) return R.pair[a] < R.pair[b];
		return region_compare(R, a, b) < 0;
	}
};
struct RegionEq {
	Regions R;
	__device__ bool operator()(uint32_t a, uint32_t b) const { return R.pair[a] == R.pair[b] && region_compare(R, a, b) == 0; }
};

// first unique record of every pair (pair_off[n_pairs] = n_uniq)
__global__ void pair_offsets_kernel(const uint32_t *__restrict__ uniq, const uint32_t *__restrict__ pair, uint32_t n_uniq, uint32_t n_pairs,
	uint32_t *pair_off)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p > n_pairs) return;
	uint32_t lo = 0, hi = n_uniq;
	while (lo < hi) {
		const uint32_t mid = (lo + hi) >> 1;
		if (pair[uniq[mid]] < p) lo = mid + 1; else hi = mid;
	}
	pair_off[p] = lo;
}

__global__ void gather_len_kernel(const uint32_t *__restrict__ uniq, const uint32_t *__restrict__ len, uint32_t n, uint64_t *out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = len[uniq[i]];
}

// the amplicon strings as IUPAC text, one CTA per string
__global__ void text_kernel(Regions R, const uint32_t *__restrict__ uniq, const uint64_t *__restrict__ text_off, char *text)
{
	const uint32_t rec = uniq[blockIdx.x];
	const uint64_t o = R.raw_off[R.seq[rec]], dst = text_off[blockIdx.x];
	const uint32_t first = R.start[rec], n = R.len[rec];
	for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) text[dst + k] = c_bits_to_base[region_nibble(R.raw, o, first + k)];
}

// the strings as packed Sequence nibbles (Sequence::operator=(const string&), sequence.cpp:12-41): thread = output byte.
// byte_off[j] = first byte of new sequence j relative to dst
__global__ void pack_new_sequences_kernel(Regions R, const uint32_t *__restrict__ uniq, const uint64_t *__restrict__ byte_off, uint32_t n_new,
	uint64_t total_bytes, uint8_t *dst, uint32_t *flags)
{
	const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= total_bytes) return;
	uint32_t lo = 0, hi = n_new; // last j with byte_off[j] <= b
	while (hi - lo > 1u) {
		const uint32_t mid = (lo + hi) >> 1;
		if (byte_off[mid] <= b) lo = mid; else hi = mid;
	}
	const uint32_t rec = uniq[lo];
	const uint64_t o = R.raw_off[R.seq[rec]];
	const uint32_t first = R.start[rec], n = R.len[rec], k = 2u * (uint32_t)(b - byte_off[lo]);
	const uint32_t x = region_nibble(R.raw, o, first + k), y = (k + 1u < n) ? region_nibble(R.raw, o, first + k + 1u) : 0u;
	dst[b] = (uint8_t)((x << 4) | y);
	if ((x & (x - 1u)) | (y & (y - 1u))) atomicOr(flags + 1, 1u); // a degenerate base
}

struct RegionTarget { // pack_target_slots(Sequence) (seq_overlap.h:1071-1100) of an amplicon that lives inside another sequence
	const uint8_t *raw;
	uint32_t first;
	int len;
	__device__ RegionTarget(const uint8_t *r, uint32_t f, int l) : raw(r), first(f), len(l) {}
	__device__ int length() const { return len; }
	__device__ unsigned at(int j) const
	{
		const uint32_t i = first + (uint32_t)j;
		const unsigned v = __ldg(raw + (i >> 1));
		return (i & 1u) ? (v & 15u) : (v >> 4);
	}
};

// find_multiplex_background_match (background_match.cpp:168-295) of every pool assay against every distinct amplicon:
// thread = (amplicon, pool primer, strand); the same score as multiplex_sw_kernel
__global__ void __launch_bounds__(128) pool_sw_kernel(Regions R, const uint32_t *__restrict__ uniq, uint32_t n_uniq, const uint64_t *__restrict__ pool,
	uint32_t n_pool_words, float threshold, int taq, uint32_t *matched)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t total = (uint64_t)n_uniq * 2ull * n_pool_words;
	if (tid >= total) return;
	const uint32_t slot = (uint32_t)(tid % (2ull * n_pool_words)), u = (uint32_t)(tid / (2ull * n_pool_words));
	W128 w;
	w.hi = pool[2 * (slot >> 1)];
	w.lo = pool[2 * (slot >> 1) + 1];
	const int size = w_size(w);
	if (slot & 1u) w = w_complement(w);
	sw::Query q;
	sw::query_from_word(w, q);
	const uint32_t rec = uniq[u];
	const RegionTarget t(R.raw + R.raw_off[R.seq[rec]], R.start[rec], (int)R.len[rec]);
	const sw::Result s = sw::align_warp<false>(q, t);
	float norm = __fmul_rn(2.0f, (float)size);
	if (norm > 0.0f) norm = __fdiv_rn(1.0f, norm);
	float score = __fmul_rn((float)s.score, norm);
	if (taq) {
		unsigned p0, p1, t0, t1;
		word_last_two(w, p0, p1);
		sw::last_two(s, t, t0, t1);
		score = __fmul_rn(score, taq_correction(p0, p1, t0, t1));
	}
	if (score >= threshold) matched[u] = 1u;
}

__global__ void pool_count_kernel(const uint32_t *__restrict__ uniq, const uint32_t *__restrict__ pair, const uint32_t *__restrict__ matched,
	uint32_t n_uniq, uint32_t *count)
{
	const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
	if (u < n_uniq && matched[u]) atomicAdd(count + pair[uniq[u]], 1u);
}

inline Regions regions_of(pcramp_gpu_ctx *ctx)
{
	const SeqSet &s = ctx->sets[ctx->amp_kind];
	Regions R;
	R.raw = s.d_raw.as<uint8_t>();
	R.raw_off = s.d_raw_off.as<uint64_t>();
	R.seq = ctx->amp_seq.as<uint32_t>();
	R.start = ctx->amp_start.as<uint32_t>();
	R.len = ctx->amp_len.as<uint32_t>();
	R.pair = ctx->amp_pair.as<uint32_t>();
	return R;
}

} // namespace amp
} // namespace pcr

extern "C" {

int pcramp_gpu_unique_amplicons(pcramp_gpu_ctx *ctx, int kind, const uint64_t *f, const uint64_t *r, uint32_t n_pairs, float threshold,
	int amp_min, int amp_max, int want_bounds, uint64_t *n_amplicons, uint64_t *n_bases, uint64_t *n_bounds)
{
	using namespace pcr::amp;
	if (check_kind2(ctx, kind)) return 1;
	if (n_pairs && (!f || !r)) return fail(ctx, "pcramp_gpu_unique_amplicons: null argument");
	if (n_pairs >= (1u << 30)) return fail(ctx, "pcramp_gpu_unique_amplicons: too many pairs in one batch");
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[kind];
	cudaStream_t st = ctx->stream;
	if (fast_resolve(ctx)) return 1;
	if (!s.db_valid) return fail(ctx, "pcramp_gpu_unique_amplicons: no database (call pcramp_gpu_select_words first)");
	ctx->amp_kind = -1;
	ctx->amp_n_rec = ctx->amp_n_uniq = ctx->amp_n_bases = 0;
	ctx->amp_n_pairs = n_pairs;
	ctx->amp_bounds_ok = false;
	ctx->amp_words.clear();
	for (uint32_t i = 0; i < n_pairs; ++i) {
		ctx->amp_words.push_back(f[2 * i]); ctx->amp_words.push_back(f[2 * i + 1]);
		ctx->amp_words.push_back(r[2 * i]); ctx->amp_words.push_back(r[2 * i + 1]);
	}
	if (n_amplicons) *n_amplicons = 0;
	if (n_bases) *n_bases = 0;
	if (n_bounds) *n_bounds = 0;
	ctx->stats.kernel_launches = 0;
	CK(ctx->amp_pair_off.ensure(((size_t)n_pairs + 1) * 4));
	CK(cudaMemsetAsync(ctx->amp_pair_off.p, 0, ((size_t)n_pairs + 1) * 4, st));
	CK(ctx->amp_text_off.ensure(8));
	CK(cudaMemsetAsync(ctx->amp_text_off.p, 0, 8, st));
	CK(cudaStreamSynchronize(st));
	ctx->amp_kind = kind;
	if (!n_pairs || !s.n || !s.n_entries) { ctx->amp_bounds_ok = want_bounds != 0; return 0; }
	DevBuf d_f, d_r, d_ol;
	CK(d_f.ensure((size_t)n_pairs * 16));
	CK(d_r.ensure((size_t)n_pairs * 16));
	CK(d_ol.ensure((size_t)n_pairs * 2 * sizeof(OligoDev)));
	CK(cudaMemcpyAsync(d_f.p, f, (size_t)n_pairs * 16, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_r.p, r, (size_t)n_pairs * 16, cudaMemcpyHostToDevice, st));
	const float thr2 = threshold * threshold; // pcr_assay.cpp:776-777
	prep_oligos_kernel<<<grid_for(2ull * n_pairs, 256), 256, 0, st>>>(d_f.as<uint64_t>(), d_r.as<uint64_t>(), n_pairs, thr2, d_ol.as<OligoDev>());
	CK(cudaGetLastError());
	ctx->stats.kernel_launches++;
	AmpList L;
	if (build_amplicon_list<true>(ctx, s, d_ol.as<OligoDev>(), n_pairs, amp_min, amp_max, L, "pcramp_gpu_unique_amplicons")) return 1;
	const uint64_t n = L.n;
	ctx->amp_n_rec = n;
	if (!n) { ctx->amp_bounds_ok = want_bounds != 0; return 0; }
	CK(ctx->amp_seq.ensure(n * 4));
	CK(ctx->amp_start.ensure(n * 4));
	CK(ctx->amp_len.ensure(n * 4));
	CK(ctx->amp_pair.ensure(n * 4));
	CK(ctx->amp_bounds.ensure(n * 12));
	CK(ctx->amp_uniq.ensure(n * 4));
	CK(ctx->amp_flags.ensure(16));
	CK(cudaMemsetAsync(ctx->amp_flags.p, 0, 16, st));
	region_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, L.k1[0].as<uint64_t>(), L.perm[0].as<uint32_t>(), L.k2[0].as<uint64_t>(), d_ol.as<OligoDev>(),
		ctx->amp_seq.as<uint32_t>(), ctx->amp_start.as<uint32_t>(), ctx->amp_len.as<uint32_t>(), ctx->amp_pair.as<uint32_t>(),
		ctx->amp_bounds.as<uint32_t>(), ctx->amp_flags.as<uint32_t>());
	iota32_kernel<<<grid_for(n, 256), 256, 0, st>>>(ctx->amp_uniq.as<uint32_t>(), n);
	CK(cudaGetLastError());
	ctx->stats.kernel_launches += 2;
	uint32_t h_flags = 0;
	CK(cudaMemcpyAsync(&h_flags, ctx->amp_flags.p, 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	if (want_bounds && (h_flags & 1u)) {
		ctx->amp_kind = -1;
		return fail(ctx, ":AmpliconBounds(): Amplicon begin > amplicon end");
	}
	const Regions R = regions_of(ctx);
	uint32_t *u = ctx->amp_uniq.as<uint32_t>();
	RegionLess less;
	less.R = R;
	RegionEq eq;
	eq.R = R;
	CachedScratch scratch;
	thrust::sort(thrust::cuda::par(scratch).on(st), u, u + n, less); // sort(amplicons) (pcr_assay.cpp:806)
	uint32_t *e = thrust::unique(thrust::cuda::par(scratch).on(st), u, u + n, eq); // :807
	CK(cudaStreamSynchronize(st));
	const uint64_t nu = (uint64_t)(e - u);
	ctx->amp_n_uniq = nu;
	pair_offsets_kernel<<<grid_for((uint64_t)n_pairs + 1, 256), 256, 0, st>>>(u, R.pair, (uint32_t)nu, n_pairs, ctx->amp_pair_off.as<uint32_t>());
	// offsets of the strings in the text form
	CK(ctx->amp_text_off.ensure((nu + 1) * 8));
	DevBuf d_len;
	CK(d_len.ensure((nu + 1) * 8));
	CK(cudaMemsetAsync(d_len.p, 0, (nu + 1) * 8, st));
	gather_len_kernel<<<grid_for(nu, 256), 256, 0, st>>>(u, R.len, (uint32_t)nu, d_len.as<uint64_t>());
	size_t tmp_bytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len.as<uint64_t>(), ctx->amp_text_off.as<uint64_t>(), (int)(nu + 1), st);
	CK(L.tmp.ensure(tmp_bytes));
	CK(cub::DeviceScan::ExclusiveSum(L.tmp.p, tmp_bytes, d_len.as<uint64_t>(), ctx->amp_text_off.as<uint64_t>(), (int)(nu + 1), st));
	CK(cudaGetLastError());
	ctx->stats.kernel_launches += 5;
	uint64_t total = 0;
	CK(cudaMemcpyAsync(&total, ctx->amp_text_off.as<uint64_t>() + nu, 8, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	ctx->amp_n_bases = total;
	ctx->amp_bounds_ok = want_bounds != 0;
	if (n_amplicons) *n_amplicons = nu;
	if (n_bases) *n_bases = total;
	if (n_bounds) *n_bounds = n;
	return 0;
}

int pcramp_gpu_unique_amplicons_copy(pcramp_gpu_ctx *ctx, uint32_t *pair_off, uint64_t *text_off, char *text, uint32_t *bounds_pair, uint32_t *bounds)
{
	using namespace pcr::amp;
	if (!ctx) return 1;
	if (ctx->amp_kind < 0) return fail(ctx, "pcramp_gpu_unique_amplicons_copy: no amplicon list (call pcramp_gpu_unique_amplicons first)");
	CK(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const uint64_t nu = ctx->amp_n_uniq, n = ctx->amp_n_rec;
	if (pair_off) CK(cudaMemcpyAsync(pair_off, ctx->amp_pair_off.p, ((size_t)ctx->amp_n_pairs + 1) * 4, cudaMemcpyDeviceToHost, st));
	if (text_off) CK(cudaMemcpyAsync(text_off, ctx->amp_text_off.p, (nu + 1) * 8, cudaMemcpyDeviceToHost, st));
	DevBuf d_text;
	if (text && nu) {
		CK(d_text.ensure(std::max<uint64_t>(1, ctx->amp_n_bases)));
		text_kernel<<<(unsigned)nu, 128, 0, st>>>(regions_of(ctx), ctx->amp_uniq.as<uint32_t>(), ctx->amp_text_off.as<uint64_t>(), d_text.as<char>());
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(text, d_text.p, ctx->amp_n_bases, cudaMemcpyDeviceToHost, st));
	}
	if (bounds_pair && n) CK(cudaMemcpyAsync(bounds_pair, ctx->amp_pair.p, n * 4, cudaMemcpyDeviceToHost, st));
	if (bounds && n) {
		if (!ctx->amp_bounds_ok) return fail(ctx, "pcramp_gpu_unique_amplicons_copy: the list was built without bounds");
		CK(cudaMemcpyAsync(bounds, ctx->amp_bounds.p, n * 12, cudaMemcpyDeviceToHost, st));
	}
	CK(cudaStreamSynchronize(st));
	return 0;
}

int pcramp_gpu_pool_amplicon_coverage(pcramp_gpu_ctx *ctx, int kind, const uint64_t *f, const uint64_t *r, uint32_t n_pairs, float target_threshold,
	int amp_min, int amp_max, float background_threshold, int taq, float *coverage)
{
	using namespace pcr::amp;
	if (check_kind2(ctx, kind)) return 1;
	if (n_pairs && !coverage) return fail(ctx, "pcramp_gpu_pool_amplicon_coverage: null argument");
	uint64_t nu = 0;
	if (pcramp_gpu_unique_amplicons(ctx, kind, f, r, n_pairs, target_threshold, amp_min, amp_max, 0, &nu, nullptr, nullptr)) return 1;
	for (uint32_t i = 0; i < n_pairs; ++i) coverage[i] = 0.0f;
	const uint32_t n_pool_words = (uint32_t)(ctx->pool_words.size() / 2); // F0 R0 F1 R1 ...
	if (!nu || !n_pool_words) return 0;
	for (size_t i = 0; i < ctx->pool_words.size(); i += 2) // pack_query_slots throws on an empty query (seq_overlap.h:832-834)
		if ((ctx->pool_words[i] | ctx->pool_words[i + 1]) == 0) return fail(ctx, ":SeqOverlap::pack_query_slots: len == 0");
	cudaStream_t st = ctx->stream;
	DevBuf d_matched, d_count;
	CK(d_matched.ensure(nu * 4));
	CK(d_count.ensure((size_t)n_pairs * 4));
	CK(cudaMemsetAsync(d_matched.p, 0, nu * 4, st));
	CK(cudaMemsetAsync(d_count.p, 0, (size_t)n_pairs * 4, st));
	const Regions R = regions_of(ctx);
	const uint64_t total = nu * 2ull * n_pool_words;
	pool_sw_kernel<<<grid_for(total, 128), 128, 0, st>>>(R, ctx->amp_uniq.as<uint32_t>(), (uint32_t)nu, ctx->mpx_pool.as<uint64_t>(), n_pool_words,
		background_threshold, taq, d_matched.as<uint32_t>());
	pool_count_kernel<<<grid_for(nu, 256), 256, 0, st>>>(ctx->amp_uniq.as<uint32_t>(), R.pair, d_matched.as<uint32_t>(), (uint32_t)nu, d_count.as<uint32_t>());
	CK(cudaGetLastError());
	ctx->stats.kernel_launches += 2;
	std::vector<uint32_t> h(n_pairs);
	CK(cudaMemcpyAsync(h.data(), d_count.p, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	// weighted_coverage (main.cpp:1402-1418): the amplicon Sequences carry the default weight 1.0, summed in double
	for (uint32_t i = 0; i < n_pairs; ++i) coverage[i] = (float)(double)h[i];
	return 0;
}

int pcramp_gpu_accept_assay(pcramp_gpu_ctx *ctx, uint32_t pair, uint32_t pack_max_degen, uint32_t min_oligo_length, uint64_t *n_added,
	uint64_t *n_multiplex_keys)
{
	using namespace pcr::amp;
	if (!ctx) return 1;
	if (ctx->amp_kind < 0 || !ctx->amp_bounds_ok)
		return fail(ctx, "pcramp_gpu_accept_assay: no amplicon list with bounds (call pcramp_gpu_unique_amplicons with want_bounds first)");
	if (ctx->amp_kind == PCRAMP_MULTIPLEX) return fail(ctx, "pcramp_gpu_accept_assay: the amplicons were cut from the multiplex collection itself");
	if (pair >= ctx->amp_n_pairs) return fail(ctx, "pcramp_gpu_accept_assay: pair out of range");
	if (ctx->parent) return fail(ctx, "pcramp_gpu_accept_assay: a worker context shares its parent's sequences and cannot change them");
	ctx->text_gen++;
	CK(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	SeqSet &src = ctx->sets[ctx->amp_kind];
	SeqSet &m = ctx->sets[PCRAMP_MULTIPLEX];
	if (n_added) *n_added = 0;
	// the pair's unique amplicons and bounds
	std::vector<uint32_t> pair_off((size_t)ctx->amp_n_pairs + 1);
	CK(cudaMemcpyAsync(pair_off.data(), ctx->amp_pair_off.p, pair_off.size() * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	const uint32_t u0 = pair_off[pair], n_new = pair_off[pair + 1] - u0;
	std::vector<uint64_t> text_off((size_t)n_new + 1, 0);
	if (n_new) {
		CK(cudaMemcpyAsync(text_off.data(), ctx->amp_text_off.as<uint64_t>() + u0, ((size_t)n_new + 1) * 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
	}
	std::vector<uint32_t> b_pair(ctx->amp_n_rec), b(3 * ctx->amp_n_rec);
	if (ctx->amp_n_rec) {
		CK(cudaMemcpyAsync(b_pair.data(), ctx->amp_pair.p, ctx->amp_n_rec * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(b.data(), ctx->amp_bounds.p, ctx->amp_n_rec * 12, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
	}
	Trace tr("accept_assay", st);
	std::vector<uint32_t> sp_seq, sp_pos;
	for (uint64_t i = 0; i < ctx->amp_n_rec; ++i) {
		if (b_pair[i] != pair) continue;
		const uint32_t q = b[3 * i], begin = b[3 * i + 1], end = b[3 * i + 2];
		if (end >= src.len[q]) // split_sequence has no range check (sequence.h:231-243): undefined in the reference
			return fail(ctx, "pcramp_gpu_accept_assay: an amplicon ends beyond its sequence (primer bound to a partial word)");
		sp_seq.push_back(q); sp_pos.push_back(begin);                 // main.cpp:1010
		sp_seq.push_back(q); sp_pos.push_back((begin + end) / 2);     // :1014
		sp_seq.push_back(q); sp_pos.push_back(end);                   // :1016
	}
	// ---- append the amplicons to the multiplex background (main.cpp:989-999) -------------------------------------------
	if (n_new) {
		if ((uint64_t)m.n + n_new >= (1u << 24)) return fail(ctx, "pcramp_gpu_accept_assay: at most 2^24 - 1 sequences per collection");
		std::vector<uint64_t> byte_off((size_t)n_new + 1, 0);
		for (uint32_t j = 0; j < n_new; ++j) byte_off[j + 1] = byte_off[j] + (text_off[j + 1] - text_off[j] + 1) / 2;
		const uint64_t add_bytes = byte_off[n_new], old_bytes = m.n ? m.raw_bytes : 0;
		DevBuf grown, d_off;
		CK(grown.ensure(std::max<uint64_t>(16, old_bytes + add_bytes)));
		CK(d_off.ensure(byte_off.size() * 8));
		if (old_bytes) CK(cudaMemcpyAsync(grown.p, m.d_raw.p, old_bytes, cudaMemcpyDeviceToDevice, st));
		CK(cudaMemcpyAsync(d_off.p, byte_off.data(), byte_off.size() * 8, cudaMemcpyHostToDevice, st));
		CK(ctx->amp_flags.ensure(16));
		CK(cudaMemsetAsync(ctx->amp_flags.p, 0, 16, st));
		pack_new_sequences_kernel<<<grid_for(add_bytes, 256), 256, 0, st>>>(regions_of(ctx), ctx->amp_uniq.as<uint32_t>() + u0, d_off.as<uint64_t>(), n_new,
			add_bytes, grown.as<uint8_t>() + old_bytes, ctx->amp_flags.as<uint32_t>());
		CK(cudaGetLastError());
		uint32_t h_flags[2] = {0, 0};
		CK(cudaMemcpyAsync(h_flags, ctx->amp_flags.p, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		std::swap(m.d_raw.p, grown.p);
		std::swap(m.d_raw.cap, grown.cap);
		if (!m.n) { // nothing uploaded into this collection yet
			m.len.clear(); m.plen.clear(); m.clen.clear(); m.weight.clear(); m.active.clear(); m.raw_off.clear(); m.eos.clear();
			m.grp_off.assign(1, 0);
			m.any_degenerate = false;
			m.unit_weights = true;
		}
		for (uint32_t j = 0; j < n_new; ++j) {
			const uint32_t len = (uint32_t)(text_off[j + 1] - text_off[j]);
			m.len.push_back(len);
			m.clen.push_back(len);
			m.plen.push_back(len + (len & 1u));
			m.weight.push_back(1.0f); // DEFAULT_SCORE_WEIGHT (sequence.h:22)
			m.active.push_back(1);
			m.raw_off.push_back(old_bytes + byte_off[j]);
			m.eos.push_back(std::vector<uint32_t>());
			if (len & 1u) m.eos.back().push_back(len); // the pad nibble pack() also pushes (seqdev.cuh)
			m.grp_off.push_back(m.grp_off.back() + ((uint64_t)len + 31) / 32 + 1);
		}
		m.n += n_new;
		m.raw_bytes = old_bytes + add_bytes;
		m.any_degenerate = m.any_degenerate || (h_flags[1] != 0);
		m.db_valid = false;
		m.idx_drop();
		m.n_entries = m.n_keys = 0;
		ctx->mpx_valid = false;
		std::vector<uint32_t> with_eos;
		for (uint32_t i = 0; i < m.n; ++i)
			if (m.clen[i] != m.len[i]) with_eos.push_back(i);
		if (upload_finish(ctx, m, nullptr, with_eos)) return 1;
		ctx->stats.kernel_launches += 2;
	}
	if (n_added) *n_added = n_new;
	tr.mark("append amplicons");
	// ---- keys() of the multiplex background database (main.cpp:1002) ---------------------------------------------------
	if (pcramp_gpu_multiplex_keys(ctx, pack_max_degen, min_oligo_length, n_multiplex_keys)) return 1;
	tr.mark("multiplex keys");
	// ---- the assay joins the pool (main.cpp:1123) -----------------------------------------------------------------------
	for (int k = 0; k < 4; ++k) ctx->pool_words.push_back(ctx->amp_words[4ull * pair + k]);
	CK(ctx->mpx_pool.ensure(ctx->pool_words.size() * 8));
	CK(cudaMemcpyAsync(ctx->mpx_pool.p, ctx->pool_words.data(), ctx->pool_words.size() * 8, cudaMemcpyHostToDevice, st));
	CK(cudaStreamSynchronize(st));
	// ---- split the targets at begin / centre / end of every amplicon (main.cpp:1008-1017) -------------------------------
	const int kind = ctx->amp_kind;
	ctx->amp_kind = -1; // the records point into text that is about to change
	ctx->amp_bounds_ok = false;
	if (!sp_seq.empty() && pcramp_gpu_split_sequences(ctx, kind, (uint32_t)sp_seq.size(), sp_seq.data(), sp_pos.data())) return 1;
	tr.mark("splits");
	(void)src;
	return 0;
}

} // extern "C"

// ---- the best assay of a batch ----------------------------------------------------------------------------
// main.cpp:829-858 keeps, over the trials in order, the assay whose Score is larger (Score::operator<, pcramp.h:180-187:
// accuracy = target - background coverage, then oligo_overlap) and, on an equal Score, the one with the smaller
// PCR::total_degeneracy() (assay.h:536-539); only trials with background coverage <= max_background_cover compete.  A fold with
// strict comparisons = the FIRST maximum of the total order (accuracy, overlap, -degeneracy), which a parallel reduction finds
// exactly.  reduce_best_assay (main.cpp:1421-1601) applies the same two comparisons to the per-rank winners.
namespace pcr {
namespace amp {

struct BestKey {
	float accuracy, overlap;
	double degeneracy;
	uint32_t index; // 0xFFFFFFFF = nothing competes
};

__host__ __device__ inline bool best_beats(const BestKey &a, const BestKey &b)
{ // would the fold replace b by a, a coming later?  (equal keys: the earlier index stays)
	if (a.index == 0xFFFFFFFFu) return false;
	if (b.index == 0xFFFFFFFFu) return true;
	if (a.accuracy != b.accuracy) return a.accuracy > b.accuracy;
	if (a.overlap != b.overlap) return a.overlap > b.overlap;
	if (a.degeneracy != b.degeneracy) return a.degeneracy < b.degeneracy;
	return a.index < b.index;
}

__global__ void __launch_bounds__(256) best_assay_kernel(uint32_t n, const float *__restrict__ target, const float *__restrict__ background,
	const float *__restrict__ overlap, const uint64_t *__restrict__ f, const uint64_t *__restrict__ r, float max_background, BestKey *out)
{
	__shared__ BestKey s_best[256];
	BestKey best;
	best.accuracy = best.overlap = 0.0f;
	best.degeneracy = 0.0;
	best.index = 0xFFFFFFFFu;
	for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
		if (!(background[i] <= max_background)) continue; // main.cpp:829,846
		W128 wf, wr;
		wf.hi = f[2 * i]; wf.lo = f[2 * i + 1];
		wr.hi = r[2 * i]; wr.lo = r[2 * i + 1];
		double df = 1.0, dr = 1.0; // Word::degeneracy (word.h:97-138)
		for (int k = 0; k < WORD_LEN; ++k) {
			const int a = __popc(w_get(wf, k)), b = __popc(w_get(wr, k));
			if (a) df *= (double)a;
			if (b) dr *= (double)b;
		}
		BestKey c;
		c.accuracy = __fsub_rn(target[i], background[i]); // Score::accuracy (pcramp.h:204-207)
		c.overlap = overlap[i];
		c.degeneracy = df + dr;
		c.index = i;
		if (best_beats(c, best)) best = c;
	}
	s_best[threadIdx.x] = best;
	__syncthreads();
	for (uint32_t s = blockDim.x / 2; s > 0; s >>= 1) {
		if (threadIdx.x < s && best_beats(s_best[threadIdx.x + s], s_best[threadIdx.x])) s_best[threadIdx.x] = s_best[threadIdx.x + s];
		__syncthreads();
	}
	if (threadIdx.x == 0) *out = s_best[0];
}

} // namespace amp
} // namespace pcr

extern "C" {

int pcramp_gpu_best_assay(pcramp_gpu_ctx *ctx, uint32_t n, const float *target_coverage, const float *background_coverage, const float *oligo_overlap,
	const uint64_t *f, const uint64_t *r, float max_background_cover, int64_t *best_index, float *best_accuracy, float *best_overlap,
	double *best_degeneracy)
{
	using namespace pcr::amp;
	if (!ctx) return 1;
	if (n && (!target_coverage || !background_coverage || !f || !r)) return fail(ctx, "pcramp_gpu_best_assay: null argument");
	CK(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	if (best_index) *best_index = -1;
	if (!n) return 0;
	DevBuf d_t, d_b, d_o, d_f, d_r, d_out;
	CK(d_t.ensure((size_t)n * 4)); CK(d_b.ensure((size_t)n * 4)); CK(d_o.ensure((size_t)n * 4));
	CK(d_f.ensure((size_t)n * 16)); CK(d_r.ensure((size_t)n * 16)); CK(d_out.ensure(sizeof(BestKey)));
	CK(cudaMemcpyAsync(d_t.p, target_coverage, (size_t)n * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_b.p, background_coverage, (size_t)n * 4, cudaMemcpyHostToDevice, st));
	if (oligo_overlap) CK(cudaMemcpyAsync(d_o.p, oligo_overlap, (size_t)n * 4, cudaMemcpyHostToDevice, st));
	else CK(cudaMemsetAsync(d_o.p, 0, (size_t)n * 4, st));
	CK(cudaMemcpyAsync(d_f.p, f, (size_t)n * 16, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_r.p, r, (size_t)n * 16, cudaMemcpyHostToDevice, st));
	best_assay_kernel<<<1, 256, 0, st>>>(n, d_t.as<float>(), d_b.as<float>(), d_o.as<float>(), d_f.as<uint64_t>(), d_r.as<uint64_t>(), max_background_cover,
		d_out.as<BestKey>());
	CK(cudaGetLastError());
	BestKey k;
	CK(cudaMemcpyAsync(&k, d_out.p, sizeof(BestKey), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	ctx->stats.kernel_launches = 1;
	if (k.index == 0xFFFFFFFFu) return 0;
	if (best_index) *best_index = (int64_t)k.index;
	if (best_accuracy) *best_accuracy = k.accuracy;
	if (best_overlap) *best_overlap = k.overlap;
	if (best_degeneracy) *best_degeneracy = k.degeneracy;
	return 0;
}

} // extern "C"
