// pcramp_gpu.cu -- context, HBM residency and the C ABI of include/pcramp_gpu.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false (see build.py).
// -fmad=false because identity/coverage arithmetic must round exactly like the reference's scalar
// x86 code (SURVEY.md section 7 "Float bit-exactness"); the hot scan kernel is integer-only.
#include "../../include/pcramp_gpu.h"

#include "ctx.cuh"
#include "score.cuh"
#include "index.cuh"
#include "edge.cuh"

#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace pcr;

namespace {

// ---------------------------------------------------------------------------------------------
// K0: nibbles -> bit-planes
// ---------------------------------------------------------------------------------------------
// sequences without EOS: one thread per 32-base group, 16 raw bytes -> one uint4 of planes
__global__ void planes_direct_kernel(const uint8_t *__restrict__ raw, const uint64_t *__restrict__ raw_off, const uint64_t *__restrict__ grp_off,
	const uint32_t *__restrict__ len, const uint32_t *__restrict__ clen, uint32_t n_seq, uint64_t n_groups, uint4 *planes)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= n_groups) return;
	uint32_t lo = 0, hi = n_seq; // sequence owning group g: last s with grp_off[s] <= g
	while (hi - lo > 1u) {
		const uint32_t mid = (lo + hi) >> 1;
		if (grp_off[mid] <= g) lo = mid; else hi = mid;
	}
	const uint32_t seq = lo;
	if (len[seq] != clen[seq]) return; // has EOS: planes_compact_kernel writes it
	const uint32_t L = len[seq];
	const uint64_t first = (g - grp_off[seq]) * 32ull;
	uint4 out = make_uint4(0, 0, 0, 0);
	const uint8_t *src = raw + raw_off[seq];
	for (uint32_t b = 0; b < 32u; b += 2u) {
		const uint64_t pos = first + b;
		if (pos >= L) break;
		const uint32_t v = src[pos >> 1];
		const uint32_t n0 = v >> 4, n1 = (pos + 1 < L) ? (v & 15u) : 0u;
		out.x |= ((n0 & 1u) << b) | ((n1 & 1u) << (b + 1));
		out.y |= (((n0 >> 1) & 1u) << b) | (((n1 >> 1) & 1u) << (b + 1));
		out.z |= (((n0 >> 2) & 1u) << b) | (((n1 >> 2) & 1u) << (b + 1));
		out.w |= (((n0 >> 3) & 1u) << b) | (((n1 >> 3) & 1u) << (b + 1));
	}
	planes[g] = out;
}

// sequences with EOS: one warp per listed sequence compacts the non-EOS bases in order
__global__ void planes_compact_kernel(const uint8_t *__restrict__ raw, const uint64_t *__restrict__ raw_off, const uint64_t *__restrict__ grp_off,
	const uint32_t *__restrict__ len, const uint32_t *__restrict__ list, uint32_t n_list, uint32_t *planes_u32)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (warp >= n_list) return;
	const uint32_t seq = list[warp];
	const uint32_t L = len[seq];
	const uint64_t g0 = grp_off[seq], g1 = grp_off[seq + 1];
	for (uint64_t w = g0 * 4ull + lane; w < g1 * 4ull; w += 32ull) planes_u32[w] = 0u;
	__syncwarp();
	const uint8_t *src = raw + raw_off[seq];
	uint32_t base = 0;
	for (uint32_t i0 = 0; i0 < L; i0 += 32u) {
		const uint32_t i = i0 + lane;
		uint32_t nib = 0;
		if (i < L) {
			const uint32_t v = src[i >> 1];
			nib = (i & 1u) ? (v & 15u) : (v >> 4);
		}
		const uint32_t m = __ballot_sync(0xffffffffu, nib != 0u);
		if (nib != 0u) {
			const uint32_t j = base + __popc(m & ((1u << lane) - 1u));
			uint32_t *grp = planes_u32 + (g0 + (j >> 5)) * 4ull;
			const uint32_t bit = 1u << (j & 31u);
			if (nib & 1u) atomicOr(grp + 0, bit);
			if (nib & 2u) atomicOr(grp + 1, bit);
			if (nib & 4u) atomicOr(grp + 2, bit);
			if (nib & 8u) atomicOr(grp + 3, bit);
		}
		base += __popc(m);
	}
}

__global__ void clear_raw_nibbles_kernel(uint32_t *raw, const uint64_t *__restrict__ nib, uint32_t n)
{ // nibble q of the collection (even = high half of byte q / 2) becomes EOS (0); several splits may share a 32-bit word
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t byte = nib[i] >> 1;
	const uint32_t shift = 8u * (uint32_t)(byte & 3ull) + ((nib[i] & 1ull) ? 0u : 4u); // little-endian bytes inside the word
	atomicAnd(raw + (byte >> 2), ~(15u << shift));
}

// ---------------------------------------------------------------------------------------------
// candidates and patterns, built on the device from the staged pairs (select_words.cpp:25-84)
// ---------------------------------------------------------------------------------------------
__global__ void cand_count_kernel(const uint64_t *__restrict__ f, const uint64_t *__restrict__ r, uint32_t n_pairs, int opt5, int opt3,
	uint32_t *cnt)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 2u * n_pairs) return;
	const uint64_t *src = (i & 1u) ? r : f;
	W128 w;
	w.hi = src[2 * (i >> 1)];
	w.lo = src[2 * (i >> 1) + 1];
	const int start = w_start(w), stop = w_stop(w);
	uint32_t c = 1;
	if (opt5 && start > 0 && start < 32) c += (uint32_t)start;       // :50-58
	if (opt3 && stop >= 0 && stop < 31) c += (uint32_t)(31 - stop);   // :61-70
	cnt[i] = c;
}

// one oligo -> its candidate family (cand_words / cand_thr, used by the partial-word kernel and for the hit
// keys) and ONE pattern per strand (masks + meta), classified for the seed filter
__global__ void cand_build_kernel(const uint64_t *__restrict__ f, const uint64_t *__restrict__ r, uint32_t n_pairs, int opt5, int opt3,
	float threshold, const uint32_t *__restrict__ off, uint4 *cand_planes, uint32_t *cand_thr, uint4 *pat_mask, uint32_t *pat_meta,
	uint32_t *pat_meta2, uint32_t *pat_seeded)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 2u * n_pairs) return;
	const uint64_t *src = (i & 1u) ? r : f;
	W128 w;
	w.hi = src[2 * (i >> 1)];
	w.lo = src[2 * (i >> 1) + 1];
	const int start = w_start(w), stop = w_stop(w), size = w_size(w);
	const uint32_t thr = (uint32_t)__fmul_rn((float)size, threshold); // select_words.cpp:83 (same for the whole family)
	const uint32_t base = off[i];
	uint32_t c = base;
	const Planes4 fp = w_planes(w); // frame planes; a copy shifted j slots toward 5' is the planes shifted right by j bits
	cand_planes[c] = make_uint4(fp.a, fp.c, fp.g, fp.t); cand_thr[c] = thr; ++c;
	uint32_t nl = 0, nr = 0;
	if (opt5 && start > 0 && start < 32) { // select_words.cpp:50-58
		nl = (uint32_t)start;
		for (int j = 1; j <= start; ++j, ++c) { cand_planes[c] = make_uint4(fp.a >> j, fp.c >> j, fp.g >> j, fp.t >> j); cand_thr[c] = thr; }
	}
	if (opt3 && stop >= 0 && stop < 31) { // :61-70
		nr = (uint32_t)(31 - stop);
		for (int j = 1; j <= 31 - stop; ++j, ++c) { cand_planes[c] = make_uint4(fp.a << j, fp.c << j, fp.g << j, fp.t << j); cand_thr[c] = thr; }
	}
	uint4 mp = make_uint4(0, 0, 0, 0), mm = make_uint4(0, 0, 0, 0);
	for (int k = 0; start + k <= stop; ++k) {
		const uint32_t a = w_get(w, start + k); // plus strand: primer[k]
		mp.x |= (a & 1u) << k;
		mp.y |= ((a >> 1) & 1u) << k;
		mp.z |= ((a >> 2) & 1u) << k;
		mp.w |= ((a >> 3) & 1u) << k;
		const uint32_t b = w_get(w, stop - k);  // minus strand: complement of the primer read backwards
		mm.x |= ((b >> 3) & 1u) << k;           // T -> A
		mm.y |= ((b >> 2) & 1u) << k;           // G -> C
		mm.z |= ((b >> 1) & 1u) << k;           // C -> G
		mm.w |= (b & 1u) << k;                  // A -> T
	}
	// seed class: count >= thr leaves e = size - thr mismatches; e + 1 pieces of >= 5 (6) bases each
	const uint32_t n = stop >= start ? (uint32_t)(stop - start + 1) : 0u;
	uint32_t e = 0, cls = 0;
	if (thr >= 1u && thr <= (uint32_t)size && (uint32_t)size == n) {
		e = (uint32_t)size - thr;
		cls = (n / (e + 1u) >= SEED_QMIN) ? 1u : 0u; // every piece of the even split has at least 5 bases
		if (cls && (seed_entries_needed(mp, n, e + 1u) == 0u || seed_entries_needed(mm, n, e + 1u) == 0u)) cls = 0u;
	}
	pat_mask[2 * i] = mp;
	pat_meta[2 * i] = pat_meta_pack(thr, (uint32_t)start, 0u, base);
	pat_meta2[2 * i] = pat_meta2_pack(nl, nr, n, e, cls);
	pat_seeded[2 * i] = cls ? 1u : 0u;
	pat_mask[2 * i + 1] = mm;
	pat_meta[2 * i + 1] = pat_meta_pack(thr, (uint32_t)(31 - stop), 1u, base);
	pat_meta2[2 * i + 1] = pat_meta2_pack(nl, nr, n, e, cls);
	pat_seeded[2 * i + 1] = cls ? 1u : 0u;
}

// The same for a batch without shift families (the default: no --optimize.5 / .3): candidate i = oligo i, so nothing has to be
// counted or scanned first, and when every pattern is seedable (what the fast form of select_words assumes and then verifies) the
// seeded-first partition is the identity.  flags: bit 0 = a candidate has threshold 0, bit 1 = a pattern is not seedable.
__global__ void cand_build_direct_kernel(const uint64_t *__restrict__ f, const uint64_t *__restrict__ r, uint32_t n_pairs, float threshold,
	uint4 *cand_planes, uint32_t *cand_thr, uint4 *pat_mask, uint32_t *pat_meta, uint32_t *pat_meta2, unsigned long long *flags)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 2u * n_pairs) return;
	const uint64_t *src = (i & 1u) ? r : f;
	W128 w;
	w.hi = src[2 * (i >> 1)];
	w.lo = src[2 * (i >> 1) + 1];
	const int start = w_start(w), stop = w_stop(w), size = w_size(w);
	const uint32_t thr = (uint32_t)__fmul_rn((float)size, threshold); // select_words.cpp:83
	const Planes4 fp = w_planes(w);
	cand_planes[i] = make_uint4(fp.a, fp.c, fp.g, fp.t);
	cand_thr[i] = thr;
	uint4 mp = make_uint4(0, 0, 0, 0), mm = make_uint4(0, 0, 0, 0);
	for (int k = 0; start + k <= stop; ++k) {
		const uint32_t a = w_get(w, start + k); // plus strand: primer[k]
		mp.x |= (a & 1u) << k;
		mp.y |= ((a >> 1) & 1u) << k;
		mp.z |= ((a >> 2) & 1u) << k;
		mp.w |= ((a >> 3) & 1u) << k;
		const uint32_t b = w_get(w, stop - k);  // minus strand: complement of the primer read backwards
		mm.x |= ((b >> 3) & 1u) << k;
		mm.y |= ((b >> 2) & 1u) << k;
		mm.z |= ((b >> 1) & 1u) << k;
		mm.w |= (b & 1u) << k;
	}
	const uint32_t n = stop >= start ? (uint32_t)(stop - start + 1) : 0u;
	uint32_t e = 0, cls = 0;
	if (thr >= 1u && thr <= (uint32_t)size && (uint32_t)size == n) {
		e = (uint32_t)size - thr;
		cls = (n / (e + 1u) >= SEED_QMIN) ? 1u : 0u;
		if (cls && (seed_entries_needed(mp, n, e + 1u) == 0u || seed_entries_needed(mm, n, e + 1u) == 0u)) cls = 0u;
	}
	if (thr == 0u) atomicOr(flags, 1ull);
	if (!cls) atomicOr(flags, 2ull);
	pat_mask[2 * i] = mp;
	pat_meta[2 * i] = pat_meta_pack(thr, (uint32_t)start, 0u, i);
	pat_meta2[2 * i] = pat_meta2_pack(0u, 0u, n, e, cls);
	pat_mask[2 * i + 1] = mm;
	pat_meta[2 * i + 1] = pat_meta_pack(thr, (uint32_t)(31 - stop), 1u, i);
	pat_meta2[2 * i + 1] = pat_meta2_pack(0u, 0u, n, e, cls);
}

// everything the host has to see of a fast batch, in one place: {flags, n_hits, n_queries, n_indexed, n_index_entries, n_candidates, n_entries}
__global__ void fast_gather_kernel(unsigned long long *out, const unsigned long long *hit_count, const unsigned int *idx_counters,
	const unsigned int *fst_flags, const uint32_t *seq_ent_off, uint32_t n_seg)
{
	out[0] |= (unsigned long long)fst_flags[1];
	out[1] = hit_count[0];
	out[2] = idx_counters[0];
	out[3] = idx_counters[1];
	out[4] = (unsigned long long)idx_counters[2] | ((unsigned long long)idx_counters[3] << 32);
	out[5] = idx_counters[4];
	out[6] = seq_ent_off[n_seg];
}

// seeded patterns first (in order), brute-force patterns after them
__global__ void pat_partition_kernel(const uint4 *__restrict__ mask, const uint32_t *__restrict__ meta, const uint32_t *__restrict__ meta2,
	const uint32_t *__restrict__ seeded, const uint32_t *__restrict__ seeded_before, uint32_t n_pat, uint32_t n_seeded, uint4 *o_mask,
	uint32_t *o_meta, uint32_t *o_meta2)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n_pat) return;
	const uint32_t d = seeded[p] ? seeded_before[p] : n_seeded + (p - seeded_before[p]);
	o_mask[d] = mask[p];
	o_meta[d] = meta[p];
	o_meta2[d] = meta2[p];
}

__global__ void seed_overflow_kernel(const uint32_t *__restrict__ cnt, unsigned int *flag)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < SEED_BUCKETS && cnt[i] > 4095u) atomicExch(flag, 1u);
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
int rebuild_tiles(pcramp_gpu_ctx *ctx, SeqSet &s)
{
	std::vector<uint32_t> tseq, tx0;
	s.total_positions = 0;
	for (uint32_t i = 0; i < s.n; ++i) {
		s.total_positions += s.clen[i];
		for (uint64_t x = 0; x < s.clen[i]; x += SCAN_TILE) {
			tseq.push_back(i);
			tx0.push_back((uint32_t)x);
		}
	}
	s.n_tiles = tseq.size();
	CK(s.d_tile_seq.ensure(std::max<size_t>(1, tseq.size()) * 4));
	CK(s.d_tile_x0.ensure(std::max<size_t>(1, tx0.size()) * 4));
	if (!tseq.empty()) {
		CK(cudaMemcpyAsync(s.d_tile_seq.p, tseq.data(), tseq.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(s.d_tile_x0.p, tx0.data(), tx0.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
	}
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

int upload_eos(pcramp_gpu_ctx *ctx, SeqSet &s)
{
	std::vector<uint32_t> off(s.n + 1, 0), pos;
	for (uint32_t i = 0; i < s.n; ++i) {
		off[i] = (uint32_t)pos.size();
		pos.insert(pos.end(), s.eos[i].begin(), s.eos[i].end());
	}
	off[s.n] = (uint32_t)pos.size();
	CK(s.d_eos_off.ensure(off.size() * 4));
	CK(s.d_eos_pos.ensure(std::max<size_t>(1, pos.size()) * 4));
	CK(cudaMemcpyAsync(s.d_eos_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (!pos.empty()) CK(cudaMemcpyAsync(s.d_eos_pos.p, pos.data(), pos.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

int rebuild_dirty(pcramp_gpu_ctx *ctx, SeqSet &s)
{
	s.n_dirty = 0;
	if (!s.any_degenerate || s.n_groups == 0) return 0;
	const uint64_t n_words = (s.n_groups + 31) / 32;
	CK(s.d_dirty_bits.ensure(n_words * 4));
	CK(ctx->d_counters.ensure(8 * sizeof(unsigned long long)));
	unsigned int *d_n = (unsigned int *)ctx->d_counters.p;
	for (int pass = 0; pass < 2; ++pass) { // count, then fill
		CK(cudaMemsetAsync(d_n, 0, 8 * sizeof(unsigned long long), ctx->stream));
		dirty_bits_kernel<<<grid_for(n_words, 128), 128, 0, ctx->stream>>>(s.dev(), s.n_groups, s.d_dirty_bits.as<uint32_t>(),
			s.d_dirty_seq.as<uint32_t>(), s.d_dirty_grp.as<uint32_t>(), d_n, pass == 0 ? 0u : s.n_dirty);
		CK(cudaGetLastError());
		unsigned int n = 0;
		CK(cudaMemcpyAsync(&n, d_n, 4, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		if (pass == 0) {
			s.n_dirty = n;
			if (n == 0) break;
			CK(s.d_dirty_seq.ensure((size_t)n * 4));
			CK(s.d_dirty_grp.ensure((size_t)n * 4));
		}
	}
	return 0;
}

int compact_sequences(pcramp_gpu_ctx *ctx, SeqSet &s, const std::vector<uint32_t> &list)
{
	if (list.empty()) return 0;
	DevBuf d_list;
	CK(d_list.ensure(list.size() * 4));
	CK(cudaMemcpyAsync(d_list.p, list.data(), list.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
	planes_compact_kernel<<<grid_for(list.size() * 32ull, 256), 256, 0, ctx->stream>>>(s.d_raw.as<uint8_t>(), s.d_raw_off.as<uint64_t>(),
		s.d_grp_off.as<uint64_t>(), s.d_len.as<uint32_t>(), d_list.as<uint32_t>(), (uint32_t)list.size(), s.d_planes.as<uint32_t>());
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

inline float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
	float ms = 0.0f;
	cudaEventElapsedTime(&ms, a, b);
	return ms;
}

// The text index of the indexed seed scan (index.cuh): every position of the collection sorted by its 12-mer code, with the
// 48-base context planes, in parts of consecutive sequences that hold fewer than ctx->idx_part_cap (2^31) positions each -- a
// collection of 5 x 10^9 bases is three parts, each queried on its own.  Cost per part: one radix sort of its positions; ~32 bytes
// per base while it is built (the scratch is kept between builds while it is small), 16 afterwards.  A failed allocation leaves the
// collection on scan_seed_kernel.  Splits do not come here (see SeqSet::idx_stale).
// scan_index_async_kernel (a ring of chunks in shared memory filled by cp.async, index.cuh) or, with option "use_async_scan" = 0,
// scan_index_kernel (the entries of one range in registers)
static int launch_scan_index(pcramp_gpu_ctx *ctx, cudaStream_t st, const TextIndex &ix, const IdxQuery *queries, const unsigned int *n_queries,
	uint32_t q_cap, const uint4 *mask, const uint32_t *meta, const IdxCandSink &cs)
{
	const unsigned grid = (unsigned)ctx->sm_count * IDX_BLOCKS_PER_SM;
	if (ctx->use_async_scan) {
		if (!ctx->async_scan_ready) {
			CK(cudaFuncSetAttribute(scan_index_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IDXA_SMEM));
			CK(cudaFuncSetAttribute(scan_index_async_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
			if (getenv("PCRAMP_TRACE")) {
				int nb = 0;
				cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, scan_index_async_kernel, IDX_THREADS, IDXA_SMEM);
				fprintf(stderr, "[trace] scan_index_async_kernel: %d resident CTAs per SM, %zu bytes of shared memory each\n", nb, (size_t)IDXA_SMEM);
			}
			ctx->async_scan_ready = true;
		}
		scan_index_async_kernel<<<grid, IDX_THREADS, IDXA_SMEM, st>>>(ix, queries, n_queries, q_cap, mask, meta, cs);
	} else {
		scan_index_kernel<<<grid, IDX_THREADS, 0, st>>>(ix, queries, n_queries, q_cap, mask, meta, cs);
	}
	CK(cudaGetLastError());
	return 0;
}

int build_index(pcramp_gpu_ctx *ctx, SeqSet &s)
{
	s.idx_drop();
	const uint64_t N = s.total_positions;
	if (N == 0) {
		s.idx_failed = true;
		return 0;
	}
	cudaStream_t st = ctx->stream;
	const uint64_t cap = std::max<uint64_t>(1024, std::min<uint64_t>(ctx->idx_part_cap, 1ull << 31));
	for (uint32_t i = 0; i < s.n; ++i)
		if (s.clen[i] >= cap) { // one sequence longer than a part: not indexable
			s.idx_failed = true;
			return 0;
		}
	const double t0 = Trace::now();
	const SeqDev sd = s.dev();
	bool ok = true;
	uint32_t lo = 0;
	while (lo < s.n && ok) {
		uint32_t hi = lo;
		uint64_t n = 0;
		while (hi < s.n && n + s.clen[hi] < cap) n += s.clen[hi++];
		std::unique_ptr<SeqSet::IndexPart> part(new SeqSet::IndexPart());
		part->seq_lo = lo;
		part->seq_hi = hi;
		part->n = (uint32_t)n;
		const uint32_t ns = hi - lo;
		lo = hi;
		if (n == 0) continue;
		std::vector<uint32_t> cum(ns + 1, 0);
		for (uint32_t i = 0; i < ns; ++i) cum[i + 1] = cum[i] + s.clen[part->seq_lo + i];
		size_t tmp_bytes = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr,
			(int64_t)n, 0, (int)(2 * IDX_K), st);
		ok = part->cum.ensure((size_t)(ns + 1) * 4) == cudaSuccess && ctx->idx_key[0].ensure(n * 4) == cudaSuccess &&
		     ctx->idx_key[1].ensure(n * 4) == cudaSuccess && ctx->idx_val[0].ensure(n * 4) == cudaSuccess && ctx->idx_val[1].ensure(n * 4) == cudaSuccess &&
		     ctx->idx_tmp.ensure(tmp_bytes) == cudaSuccess && part->off.ensure((size_t)(IDX_CODES + 1) * 4) == cudaSuccess &&
		     part->entries.ensure(n * 16) == cudaSuccess;
		if (!ok) break;
		CK(cudaMemcpyAsync(part->cum.p, cum.data(), (size_t)(ns + 1) * 4, cudaMemcpyHostToDevice, st));
		{ // block table: last sequence (part-relative) whose first position is <= b << IDX_BLK_SHIFT (idx_seq_of_fast)
			std::vector<uint32_t> blk((n >> IDX_BLK_SHIFT) + 2, 0);
			uint32_t q = 0;
			for (size_t b = 0; b < blk.size(); ++b) {
				const uint64_t g = (uint64_t)b << IDX_BLK_SHIFT;
				while (q + 1 < ns && cum[q + 1] <= g) ++q;
				blk[b] = q;
			}
			CK(part->blk.ensure(blk.size() * 4));
			CK(cudaMemcpyAsync(part->blk.p, blk.data(), blk.size() * 4, cudaMemcpyHostToDevice, st));
			CK(cudaStreamSynchronize(st));
		}
		uint32_t *k0 = ctx->idx_key[0].as<uint32_t>(), *k1 = ctx->idx_key[1].as<uint32_t>(), *v0 = ctx->idx_val[0].as<uint32_t>(),
		         *v1 = ctx->idx_val[1].as<uint32_t>();
		index_key_kernel<<<grid_for(n, 256), 256, 0, st>>>(sd, part->cum.as<uint32_t>(), part->seq_lo, ns, (uint32_t)n, k0, v0);
		CK(cudaGetLastError());
		CK(cub::DeviceRadixSort::SortPairs(ctx->idx_tmp.p, tmp_bytes, k0, k1, v0, v1, (int64_t)n, 0, (int)(2 * IDX_K), st));
		index_offsets_kernel<<<grid_for(n + 1, 256), 256, 0, st>>>(k1, (uint32_t)n, part->off.as<uint32_t>());
		CK(cudaGetLastError());
		index_entry_kernel<<<grid_for(n, 256), 256, 0, st>>>(sd, part->cum.as<uint32_t>(), part->seq_lo, ns, v1, (uint32_t)n, part->entries.as<uint4>());
		CK(cudaGetLastError());
		CK(cudaStreamSynchronize(st));
		s.idx_bytes += part->entries.cap + part->off.cap + part->cum.cap + part->blk.cap;
		s.idx_parts.push_back(std::move(part));
	}
	// the sort scratch of a large collection (4 x 4 bytes per position) is given back; a small one stays for the next build
	if (ctx->idx_key[0].cap > (256ull << 20)) {
		for (int k = 0; k < 2; ++k) { ctx->idx_key[k].release(); ctx->idx_val[k].release(); }
		ctx->idx_tmp.release();
	}
	if (!ok) { // not enough device memory: keep the table-based scan
		cudaGetLastError();
		s.idx_drop();
		s.idx_failed = true;
		return 0;
	}
	s.idx_stale.assign(s.n, 0);
	s.n_idx_stale = 0;
	CK(s.d_idx_stale.ensure(std::max<size_t>(1, s.n)));
	CK(cudaMemsetAsync(s.d_idx_stale.p, 0, std::max<size_t>(1, s.n), st));
	CK(cudaStreamSynchronize(st));
	s.idx_valid = true;
	s.idx_builds += 1;
	s.idx_build_ms = (float)(Trace::now() - t0);
	return 0;
}

// The partial words of a collection as a table the candidates are looked up in (edge.cuh).  One-time per upload / split and set of
// pack() parameters; synchronous (two small counts come back to size the buffers).  Leaves `failed` set when the table is not
// worth its memory (the callers then keep scan_edge_fst_kernel).
int edge_table_build(pcramp_gpu_ctx *ctx, SeqSet &s, const PackParams &pp)
{
	SeqSet::EdgeTab &t = s.edge;
	t.valid = false;
	t.failed = true;
	t.pp = pp;
	if (s.n == 0) return 0;
	const auto t0 = std::chrono::steady_clock::now();
	cudaStream_t st = ctx->stream;
	Trace tr("edge_table_build", st);
	const SeqDev sd = s.dev();
	DevBuf counters, cnt;
	CK(counters.ensure(16));
	CK(cnt.ensure((size_t)(EDGE_BUCKETS + 1) * 4));
	CK(t.start.ensure((size_t)(EDGE_BUCKETS + 1) * 4));
	CK(cudaMemsetAsync(counters.p, 0, 16, st));
	CK(cudaMemsetAsync(cnt.p, 0, (size_t)(EDGE_BUCKETS + 1) * 4, st));
	const unsigned grid = (unsigned)std::min<uint64_t>(((uint64_t)s.n + 3) / 4, (uint64_t)ctx->sm_count * 16);
	edge_words_kernel<false><<<grid, 128, 0, st>>>(sd, pp, counters.as<unsigned int>(), cnt.as<uint32_t>(), nullptr, nullptr, nullptr, 0u, 0u);
	CK(cudaGetLastError());
	size_t tmp_bytes = 0;
	CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.as<uint32_t>(), t.start.as<uint32_t>(), (int)(EDGE_BUCKETS + 1), st));
	CK(ctx->cub_tmp.ensure(tmp_bytes));
	CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, cnt.as<uint32_t>(), t.start.as<uint32_t>(), (int)(EDGE_BUCKETS + 1), st));
	unsigned int h_cnt[2] = {0, 0};
	uint32_t total = 0;
	CK(cudaMemcpyAsync(h_cnt, counters.p, 8, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(&total, t.start.as<uint32_t>() + EDGE_BUCKETS, 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	tr.mark("count + prefix sum");
	const uint64_t bytes = (uint64_t)h_cnt[0] * 32ull + (uint64_t)total * 4ull + (uint64_t)h_cnt[1] * 4ull + (uint64_t)(EDGE_BUCKETS + 1) * 4ull;
	// (tens of millions of sequence ends: the scan kernel handles them.  The first test also keeps the 32-bit bucket offsets exact:
	// a word sits in at most EDGE_POS buckets)
	if ((uint64_t)h_cnt[0] * EDGE_POS >= (1ull << 31) || bytes > (8ull << 30)) return 0;
	CK(t.planes.ensure(std::max<size_t>(1, h_cnt[0]) * 16));
	CK(t.meta.ensure(std::max<size_t>(1, h_cnt[0]) * 16));
	CK(t.degen.ensure(std::max<size_t>(1, h_cnt[1]) * 4));
	CK(t.ids.ensure(std::max<size_t>(1, total) * 4));
	CK(cudaMemsetAsync(counters.p, 0, 16, st));
	edge_words_kernel<true><<<grid, 128, 0, st>>>(sd, pp, counters.as<unsigned int>(), nullptr, t.planes.as<uint4>(), t.meta.as<uint4>(), t.degen.as<uint32_t>(),
		h_cnt[0], h_cnt[1]);
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(cnt.p, t.start.p, (size_t)(EDGE_BUCKETS + 1) * 4, cudaMemcpyDeviceToDevice, st));
	if (h_cnt[0]) {
		edge_fill_kernel<<<grid_for(h_cnt[0], 256), 256, 0, st>>>(t.planes.as<uint4>(), h_cnt[0], cnt.as<uint32_t>(), t.ids.as<uint32_t>(), total);
		CK(cudaGetLastError());
	}
	CK(cudaStreamSynchronize(st));
	tr.mark("words + ids");
	t.n_words = h_cnt[0];
	t.n_degen = h_cnt[1];
	t.bytes = bytes;
	t.builds++;
	t.build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
	t.valid = true;
	t.failed = false;
	return 0;
}

int fast_resolve(pcramp_gpu_ctx *ctx); // below: verify (and if need be re-run) a batch that select_words_fast left unverified

int check_kind(pcramp_gpu_ctx *ctx, int kind)
{
	if (!ctx) return 1;
	if (kind < 0 || kind >= PCRAMP_NUM_KINDS) return fail(ctx, "pcramp_gpu: bad sequence kind");
	if (ctx->parent && ctx->parent->text_gen != ctx->seen_gen)
		return fail(ctx, "pcramp_gpu: the parent's sequences changed after this worker was created (destroy it and create a new one)");
	return fast_resolve(ctx);
}

// calls that change a collection: not on a worker (its text is the parent's); every change is counted for the workers' guard
int text_change(pcramp_gpu_ctx *ctx, const char *who)
{
	if (ctx->parent) return fail(ctx, std::string(who) + ": a worker context shares its parent's sequences and cannot change them");
	ctx->text_gen++;
	return 0;
}

} // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int pcramp_gpu_create(pcramp_gpu_ctx **out, int device)
{
	if (!out) return 1;
	*out = nullptr;
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return 2; // no CPU fallback
	if (device < 0 || device >= n_dev) return 3;
	pcramp_gpu_ctx *ctx = new pcramp_gpu_ctx();
	ctx->device = device;
	if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
		delete ctx;
		return 4;
	}
	cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
	cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
	ctx->max_smem_optin -= 2048; // room for the kernel's static shared memory and the runtime's reservation
	if (cudaFuncSetAttribute(scan_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin) != cudaSuccess) {
		cudaGetLastError();
		cudaStreamDestroy(ctx->stream);
		delete ctx;
		return 5;
	}
	for (auto &e : ctx->ev) cudaEventCreate(&e);
	cudaMallocHost((void **)&ctx->h_counters, 8 * sizeof(unsigned long long));
	cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
	cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
	cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming);
	cudaMallocHost((void **)&ctx->h_fast, 8 * sizeof(unsigned long long));
	// PCRAMP_OPTIONS="name=value,name=value": pcramp_gpu_set_option for every context of the process (A/B runs of a host that has no
	// switch of its own for an option; an unknown name fails the create -- a typo must not pass for a measurement)
	if (const char *env = getenv("PCRAMP_OPTIONS")) {
		std::string all(env);
		size_t at = 0;
		while (at < all.size()) {
			size_t end = all.find(',', at);
			if (end == std::string::npos) end = all.size();
			const std::string item = all.substr(at, end - at);
			at = end + 1;
			if (item.empty()) continue;
			const size_t eq = item.find('=');
			if (eq == std::string::npos || pcramp_gpu_set_option(ctx, item.substr(0, eq).c_str(), atoi(item.c_str() + eq + 1))) {
				fprintf(stderr, "pcramp_gpu_create: bad PCRAMP_OPTIONS item '%s'\n", item.c_str());
				pcramp_gpu_destroy(ctx);
				return 6;
			}
		}
	}
	*out = ctx;
	return 0;
}

// A second context on the same device that reads the parent's collections and text index in place and owns everything a batch
// writes (stream, candidates, hit lists, word database, results): independent batches of a sweep then run concurrently on the
// GPU, one host thread per context, and the short latency-bound stages of one batch (sorts, list builds, host round trips that
// size the next stage) fill behind the bandwidth-bound scan of another.
int pcramp_gpu_create_worker(pcramp_gpu_ctx *parent, pcramp_gpu_ctx **out)
{
	if (!parent || !out) return 1;
	*out = nullptr;
	if (parent->parent) return fail(parent, "pcramp_gpu_create_worker: the parent is itself a worker");
	pcramp_gpu_ctx *ctx = parent;
	CK(cudaSetDevice(parent->device));
	for (int kind = 0; kind < PCRAMP_NUM_KINDS; ++kind) { // the index is shared: build it once, here
		SeqSet &s = parent->sets[kind];
		if (kind != PCRAMP_MULTIPLEX && parent->use_index && s.n_tiles && !s.idx_valid && !s.idx_failed && build_index(parent, s)) return 1;
	}
	CK(cudaStreamSynchronize(parent->stream));
	pcramp_gpu_ctx *w = nullptr;
	const int rc = pcramp_gpu_create(&w, parent->device);
	if (rc) return fail(parent, "pcramp_gpu_create_worker: pcramp_gpu_create failed");
	w->parent = parent;
	w->seen_gen = parent->text_gen.load();
	parent->n_workers.fetch_add(1);
	w->use_fst = parent->use_fst; w->force_brute = parent->force_brute; w->use_index = parent->use_index; w->idx_part_cap = parent->idx_part_cap;
	w->use_neigh = parent->use_neigh; w->use_tier_table = parent->use_tier_table; w->use_fused_score = parent->use_fused_score;
	w->use_entry_score = parent->use_entry_score; w->use_seg_db = parent->use_seg_db; w->use_fast = parent->use_fast;
	w->use_async_scan = parent->use_async_scan; w->use_unit_score = parent->use_unit_score; w->use_background_units = parent->use_background_units;
	w->use_variant_groups = parent->use_variant_groups; w->use_edge_table = parent->use_edge_table;
	for (int kind = 0; kind < PCRAMP_NUM_KINDS; ++kind) {
		const SeqSet &p = parent->sets[kind];
		SeqSet &s = w->sets[kind];
		s.n = p.n; s.any_degenerate = p.any_degenerate; s.unit_weights = p.unit_weights; s.total_positions = p.total_positions;
		s.len = p.len; s.plen = p.plen; s.clen = p.clen; s.weight = p.weight; s.active = p.active; s.raw_off = p.raw_off; s.grp_off = p.grp_off;
		s.eos = p.eos;
		s.raw_bytes = p.raw_bytes; s.n_groups = p.n_groups; s.n_tiles = p.n_tiles; s.n_dirty = p.n_dirty;
		s.d_raw.alias(p.d_raw); s.d_raw_off.alias(p.d_raw_off); s.d_len.alias(p.d_len); s.d_plen.alias(p.d_plen); s.d_clen.alias(p.d_clen);
		s.d_planes.alias(p.d_planes); s.d_grp_off.alias(p.d_grp_off); s.d_eos_pos.alias(p.d_eos_pos); s.d_eos_off.alias(p.d_eos_off);
		s.d_weight.alias(p.d_weight); s.d_active.alias(p.d_active); s.d_tile_seq.alias(p.d_tile_seq); s.d_tile_x0.alias(p.d_tile_x0);
		s.d_dirty_bits.alias(p.d_dirty_bits); s.d_dirty_seq.alias(p.d_dirty_seq); s.d_dirty_grp.alias(p.d_dirty_grp);
		s.idx_parts.clear();
		for (const auto &pp : p.idx_parts) {
			std::unique_ptr<SeqSet::IndexPart> q(new SeqSet::IndexPart());
			q->seq_lo = pp->seq_lo; q->seq_hi = pp->seq_hi; q->n = pp->n;
			q->entries.alias(pp->entries); q->off.alias(pp->off); q->cum.alias(pp->cum); q->blk.alias(pp->blk);
			s.idx_parts.push_back(std::move(q));
		}
		s.idx_stale = p.idx_stale; s.n_idx_stale = p.n_idx_stale; s.d_idx_stale.alias(p.d_idx_stale);
		s.idx_bytes = p.idx_bytes; s.idx_builds = p.idx_builds; s.idx_build_ms = p.idx_build_ms;
		s.idx_valid = p.idx_valid; s.idx_failed = p.idx_failed || !p.idx_valid; // never build a private copy of the index
	}
	*out = w;
	return 0;
}

void pcramp_gpu_destroy(pcramp_gpu_ctx *ctx)
{
	if (!ctx) return;
	if (ctx->n_workers.load() > 0) { // its workers read its sequences and text index in place: freeing them would leave dangling pointers
		fprintf(stderr, "pcramp_gpu_destroy: %d worker context(s) still alive; destroy the workers first (context kept)\n", ctx->n_workers.load());
		ctx->err = "pcramp_gpu_destroy: worker contexts still alive";
		return;
	}
	if (ctx->parent) ctx->parent->n_workers.fetch_sub(1);
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	for (auto &e : ctx->ev) cudaEventDestroy(e);
	if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
	if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
	if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
	if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
	if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
	if (ctx->h_fast) cudaFreeHost(ctx->h_fast);
	nc::thermo_state_free(ctx->thermo);
	ctx->thermo = nullptr;
	pcramp_gpu_exchange_destroy(ctx);
	pcramp_gpu_fasta_free(ctx);
	cudaStream_t s = ctx->stream;
	delete ctx;
	cudaStreamDestroy(s);
}

const char *pcramp_gpu_last_error(const pcramp_gpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "pcramp_gpu: null context"; }
void *pcramp_gpu_stream(pcramp_gpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int pcramp_gpu_synchronize(pcramp_gpu_ctx *ctx)
{
	if (!ctx) return 1;
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

int upload_finish(pcramp_gpu_ctx *ctx, pcr::SeqSet &s, const uint8_t *nibbles, const std::vector<uint32_t> &with_eos);

int pcramp_gpu_upload_sequences(pcramp_gpu_ctx *ctx, int kind, uint32_t n, const uint8_t *nibbles, const uint64_t *byte_off,
	const uint32_t *len, const float *weight)
{
	if (check_kind(ctx, kind) || text_change(ctx, "pcramp_gpu_upload_sequences")) return 1;
	if (n >= (1u << 24)) return fail(ctx, "pcramp_gpu_upload_sequences: at most 2^24 - 1 sequences per collection");
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[kind];
	s.n = n;
	s.db_valid = false;
	if (kind == PCRAMP_MULTIPLEX) ctx->mpx_valid = false;
	s.idx_drop(); // the text index (index.cuh) is rebuilt on the next seeded scan
	s.n_entries = s.n_keys = 0;
	s.len.assign(len, len + n);
	s.plen.assign(n, 0);
	s.clen.assign(n, 0);
	s.weight.assign(n, 1.0f);
	if (weight) s.weight.assign(weight, weight + n);
	s.unit_weights = true;
	for (float w : s.weight) s.unit_weights = s.unit_weights && (w == 1.0f);
	s.active.assign(n, 1);
	s.raw_off.assign(byte_off, byte_off + n);
	s.eos.assign(n, std::vector<uint32_t>());
	s.grp_off.assign(n + 1, 0);
	s.any_degenerate = false;
	s.raw_bytes = 0;
	std::vector<uint32_t> with_eos;
	for (uint32_t i = 0; i < n; ++i) {
		const uint64_t bytes = ((uint64_t)len[i] + 1) / 2;
		s.raw_bytes = std::max(s.raw_bytes, byte_off[i] + bytes);
		const uint8_t *src = nibbles + byte_off[i];
		// host pass: EOS positions and "any degenerate base" (SWAR over 8 bytes = 16 nibbles)
		uint64_t p = 0;
		for (; p + 8 <= len[i] / 2; p += 8) {
			uint64_t v;
			memcpy(&v, src + p, 8);
			const uint64_t b0 = v & 0x1111111111111111ull, b1 = (v >> 1) & 0x1111111111111111ull, b2 = (v >> 2) & 0x1111111111111111ull,
			               b3 = (v >> 3) & 0x1111111111111111ull;
			if ((b0 & b1) | (b0 & b2) | (b0 & b3) | (b1 & b2) | (b1 & b3) | (b2 & b3)) s.any_degenerate = true;
			if ((b0 | b1 | b2 | b3) != 0x1111111111111111ull) {
				for (uint32_t q = 0; q < 16; ++q) {
					const uint32_t pos = (uint32_t)(2 * p + q);
					const uint32_t nib = (q & 1u) ? (src[pos >> 1] & 15u) : (src[pos >> 1] >> 4);
					if (nib == 0u) s.eos[i].push_back(pos);
				}
			}
		}
		for (uint64_t pos = 2 * p; pos < len[i]; ++pos) {
			const uint32_t nib = (pos & 1u) ? (src[pos >> 1] & 15u) : (src[pos >> 1] >> 4);
			if (nib == 0u) s.eos[i].push_back((uint32_t)pos);
			else if (nib & (nib - 1u)) s.any_degenerate = true;
		}
		s.clen[i] = len[i] - (uint32_t)s.eos[i].size();
		if (!s.eos[i].empty()) with_eos.push_back(i);
		s.plen[i] = len[i] + (len[i] & 1u);
		if (len[i] & 1u) s.eos[i].push_back(len[i]); // the pad nibble pack() also pushes (seqdev.cuh)
		s.grp_off[i + 1] = s.grp_off[i] + ((uint64_t)s.clen[i] + 31) / 32 + 1;
	}
	return upload_finish(ctx, s, nibbles, with_eos);
}

// device side of an upload: metadata to HBM, nibbles -> bit-planes, EOS compaction, dirty groups, tiles.  host_nibbles == NULL:
// d_raw already holds the packed nibbles (the FASTA ingest path packs them on the device, fasta.cuh).
int upload_finish(pcramp_gpu_ctx *ctx, SeqSet &s, const uint8_t *nibbles, const std::vector<uint32_t> &with_eos)
{
	const uint32_t n = s.n;
	s.n_groups = s.grp_off[n];
	CK(s.d_raw.ensure(std::max<uint64_t>(16, s.raw_bytes)));
	CK(s.d_raw_off.ensure(std::max<size_t>(1, n) * 8));
	CK(s.d_len.ensure(std::max<size_t>(1, n) * 4));
	CK(s.d_clen.ensure(std::max<size_t>(1, n) * 4));
	CK(s.d_plen.ensure(std::max<size_t>(1, n) * 4));
	CK(s.d_weight.ensure(std::max<size_t>(1, n) * 4));
	CK(s.d_active.ensure(std::max<size_t>(1, n)));
	CK(s.d_grp_off.ensure((size_t)(n + 1) * 8));
	CK(s.d_planes.ensure(std::max<uint64_t>(1, s.n_groups) * 16));
	if (s.raw_bytes && nibbles) CK(cudaMemcpyAsync(s.d_raw.p, nibbles, s.raw_bytes, cudaMemcpyHostToDevice, ctx->stream));
	if (n) {
		CK(cudaMemcpyAsync(s.d_raw_off.p, s.raw_off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(s.d_len.p, s.len.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(s.d_clen.p, s.clen.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(s.d_plen.p, s.plen.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(s.d_weight.p, s.weight.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(s.d_active.p, s.active.data(), (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
	}
	CK(cudaMemcpyAsync(s.d_grp_off.p, s.grp_off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemsetAsync(s.d_planes.p, 0, std::max<uint64_t>(1, s.n_groups) * 16, ctx->stream));
	if (s.n_groups) {
		planes_direct_kernel<<<grid_for(s.n_groups, 256), 256, 0, ctx->stream>>>(s.d_raw.as<uint8_t>(), s.d_raw_off.as<uint64_t>(),
			s.d_grp_off.as<uint64_t>(), s.d_len.as<uint32_t>(), s.d_clen.as<uint32_t>(), n, s.n_groups, s.d_planes.as<uint4>());
		CK(cudaGetLastError());
	}
	CK(cudaStreamSynchronize(ctx->stream));
	if (upload_eos(ctx, s)) return 1;
	if (compact_sequences(ctx, s, with_eos)) return 1;
	if (rebuild_dirty(ctx, s)) return 1;
	return rebuild_tiles(ctx, s);
}

int pcramp_gpu_set_weights(pcramp_gpu_ctx *ctx, int kind, const float *weight)
{ // Sequence::weight(w) (sequence.h), e.g. the per-file normalisation of main.cpp:268-278 after a FASTA upload
	if (check_kind(ctx, kind) || text_change(ctx, "pcramp_gpu_set_weights")) return 1;
	SeqSet &s = ctx->sets[kind];
	if (s.n && !weight) return fail(ctx, "pcramp_gpu_set_weights: null argument");
	CK(cudaSetDevice(ctx->device));
	s.weight.assign(weight, weight + s.n);
	s.unit_weights = true;
	for (float w : s.weight) s.unit_weights = s.unit_weights && (w == 1.0f);
	if (s.n) CK(cudaMemcpyAsync(s.d_weight.p, s.weight.data(), (size_t)s.n * 4, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

int pcramp_gpu_set_active(pcramp_gpu_ctx *ctx, int kind, const uint8_t *active)
{
	if (check_kind(ctx, kind) || text_change(ctx, "pcramp_gpu_set_active")) return 1;
	if (!active && ctx->sets[kind].n) return fail(ctx, "pcramp_gpu_set_active: null argument");
	SeqSet &s = ctx->sets[kind];
	for (uint32_t i = 0; i < s.n; ++i) s.active[i] = active[i] ? 1 : 0;
	if (s.n) CK(cudaMemcpyAsync(s.d_active.p, s.active.data(), s.n, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

int pcramp_gpu_split_sequences(pcramp_gpu_ctx *ctx, int kind, uint32_t n, const uint32_t *seq, const uint32_t *pos)
{ // Sequence::split_sequence (sequence.h:231-243) for a list of positions, e.g. the three splits per amplicon of main.cpp:1008-1017
	if (check_kind(ctx, kind) || text_change(ctx, "pcramp_gpu_split_sequence")) return 1;
	SeqSet &s = ctx->sets[kind];
	if (n && (!seq || !pos)) return fail(ctx, "pcramp_gpu_split_sequences: null argument");
	for (uint32_t i = 0; i < n; ++i)
		if (seq[i] >= s.n || pos[i] >= s.len[seq[i]]) return fail(ctx, "pcramp_gpu_split_sequence: out of range");
	CK(cudaSetDevice(ctx->device));
	std::vector<uint64_t> nib; // global nibble index of every position that becomes EOS now
	std::vector<uint32_t> touched;
	for (uint32_t i = 0; i < n; ++i) {
		std::vector<uint32_t> &e = s.eos[seq[i]];
		auto it = std::lower_bound(e.begin(), e.end(), pos[i]);
		if (it != e.end() && *it == pos[i]) continue; // already EOS
		e.insert(it, pos[i]);
		s.clen[seq[i]] -= 1;
		nib.push_back(2 * s.raw_off[seq[i]] + pos[i]);
		touched.push_back(seq[i]);
	}
	if (nib.empty()) return 0;
	std::sort(touched.begin(), touched.end());
	touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
	s.db_valid = false;
	s.edge.drop(); // the split sequences have new partial words (EOS events): the table is rebuilt by the next fast batch
	if (kind == PCRAMP_MULTIPLEX) ctx->mpx_valid = false;
	if (s.idx_valid) { // the index stays: the split sequences' entries are ignored from now on (SeqSet::idx_stale)
		for (uint32_t q : touched)
			if (!s.idx_stale[q]) {
				s.idx_stale[q] = 1;
				s.n_idx_stale++;
			}
		CK(cudaMemcpyAsync(s.d_idx_stale.p, s.idx_stale.data(), s.n, cudaMemcpyHostToDevice, ctx->stream));
	} else {
		s.idx_failed = false;
	}
	DevBuf d_nib;
	CK(d_nib.ensure(nib.size() * 8));
	CK(cudaMemcpyAsync(d_nib.p, nib.data(), nib.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
	clear_raw_nibbles_kernel<<<grid_for(nib.size(), 128), 128, 0, ctx->stream>>>(s.d_raw.as<uint32_t>(), d_nib.as<uint64_t>(), (uint32_t)nib.size());
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(s.d_clen.p, s.clen.data(), (size_t)s.n * 4, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (upload_eos(ctx, s)) return 1;
	if (compact_sequences(ctx, s, touched)) return 1;
	if (rebuild_dirty(ctx, s)) return 1; // the compressed text moved
	return rebuild_tiles(ctx, s);
}

int pcramp_gpu_split_sequence(pcramp_gpu_ctx *ctx, int kind, uint32_t seq, uint32_t pos)
{
	return pcramp_gpu_split_sequences(ctx, kind, 1, &seq, &pos);
}

int pcramp_gpu_pack(pcramp_gpu_ctx *ctx, int kind, uint32_t seq, uint32_t pack_max_degen, float min_gc, float max_gc, uint32_t min_len,
	uint64_t cap, uint64_t *words, int32_t *loc, uint32_t *strand, uint64_t *n_out)
{
	if (check_kind(ctx, kind)) return 1;
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[kind];
	if (seq >= s.n) return fail(ctx, "pcramp_gpu_pack: sequence out of range");
	if (min_len == 0) return fail(ctx, "pcramp_gpu_pack: min_oligo_length must be >= 1");
	PackParams pp;
	pp.max_degen = pack_max_degen;
	pp.min_gc = min_gc;
	pp.max_gc = max_gc;
	pp.min_len = min_len;
	pp.gc_filter = (min_gc > 0.0f) || (max_gc < 1.0f);
	const uint64_t worst = 2ull * ((uint64_t)s.plen[seq] + 64ull + s.eos[seq].size());
	DevBuf o_words, o_loc, o_strand;
	CK(o_words.ensure(worst * 16));
	CK(o_loc.ensure(worst * 4));
	CK(o_strand.ensure(worst * 4));
	CK(ctx->d_counters.ensure(8 * sizeof(unsigned long long)));
	CK(cudaMemsetAsync(ctx->d_counters.p, 0, 8 * sizeof(unsigned long long), ctx->stream));
	pack_dump_kernel<<<grid_for(worst / 2 + 1, 256), 256, 0, ctx->stream>>>(s.dev(), seq, pp, o_words.as<uint64_t>(), o_loc.as<int32_t>(),
		o_strand.as<uint32_t>(), ctx->d_counters.as<unsigned long long>());
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	const uint64_t n = ctx->h_counters[0];
	if (n_out) *n_out = n;
	if (n > cap || !words) return 0; // caller sizes with cap = 0 first
	if (n) {
		CK(cudaMemcpyAsync(words, o_words.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaMemcpyAsync(loc, o_loc.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaMemcpyAsync(strand, o_strand.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	return 0;
}

int pcramp_gpu_stage_pairs(pcramp_gpu_ctx *ctx, const uint64_t *f, const uint64_t *r, uint32_t n_pairs)
{
	if (!ctx) return 1;
	if (fast_resolve(ctx)) return 1; // a re-run of the previous batch would need the pairs that are about to be replaced
	CK(cudaSetDevice(ctx->device));
	ctx->n_pairs = ctx->n_staged = n_pairs;
	ctx->batch_first = 0;
	CK(ctx->d_f.ensure(std::max<size_t>(1, n_pairs) * 16));
	CK(ctx->d_r.ensure(std::max<size_t>(1, n_pairs) * 16));
	if (n_pairs) {
		CK(cudaMemcpyAsync(ctx->d_f.p, f, (size_t)n_pairs * 16, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(ctx->d_r.p, r, (size_t)n_pairs * 16, cudaMemcpyHostToDevice, ctx->stream));
	}
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

int pcramp_gpu_set_batch(pcramp_gpu_ctx *ctx, uint32_t first, uint32_t count)
{
	if (!ctx) return 1;
	if (fast_resolve(ctx)) return 1;
	if ((uint64_t)first + count > ctx->n_staged) return fail(ctx, "pcramp_gpu_set_batch: window exceeds the staged pairs");
	ctx->batch_first = first;
	ctx->n_pairs = count;
	return 0;
}

static int fst_build(pcramp_gpu_ctx *ctx, const uint4 *d_planes, const uint32_t *d_thr, uint32_t n, Fst &t, uint32_t &n_brute_out);

// The canonical (word, index, loc, strand) order of the database and its keys() numbering (read_only_multimap::sort,
// pcramp.h:231-256).  Pair scoring on the GPU does not need it, so it is built on demand: by db_copy / keys_copy, by
// select_words when the caller asks for the key count, and by the key-matrix fallback of pair scoring.
// The words of the entries (and the (index, loc, strand) sort key): the segmented build does not write them -- pair scoring reads
// letter planes -- so whoever needs them (canonical order, keys(), db_copy, the Smith-Waterman background test) asks here first.
static int db_words(pcramp_gpu_ctx *ctx, SeqSet &s)
{
	if (s.words_valid || !s.db_valid) return 0;
	const uint64_t n = s.n_entries;
	CK(s.e_hi.ensure(std::max<uint64_t>(1, n) * 8));
	CK(s.e_lo.ensure(std::max<uint64_t>(1, n) * 8));
	CK(s.e_order.ensure(std::max<uint64_t>(1, n) * 8));
	if (n) {
		words_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(s.dev(), s.db_pp, s.e_id.as<uint32_t>(), s.e_seq.as<uint32_t>(), s.e_strand.as<uint32_t>(), n,
			s.db_pb, s.e_hi.as<uint64_t>(), s.e_lo.as<uint64_t>(), s.e_order.as<uint64_t>());
		CK(cudaGetLastError());
		ctx->stats.kernel_launches++;
	}
	s.words_valid = true;
	return 0;
}

static int db_finalize_keys(pcramp_gpu_ctx *ctx, SeqSet &s)
{
	if (s.keys_valid || !s.db_valid) return 0;
	if (db_words(ctx, s)) return 1;
	cudaStream_t st = ctx->stream;
	const uint64_t n_ent = s.n_entries;
	if (n_ent == 0) {
		s.n_keys = 0;
		s.keys_valid = true;
		return 0;
	}
	pcramp_gpu_stats &stat = ctx->stats;
	const uint32_t seq_bits = s.seq_bits;
	size_t tmp_bytes = 0;
	CK(s.e_perm.ensure(n_ent * 4));
	CK(s.e_keyrank.ensure(n_ent * 4));
	CK(ctx->order_key[0].ensure(n_ent * 8));
	CK(ctx->order_key[1].ensure(n_ent * 8));
	CK(ctx->perm[0].ensure(n_ent * 4));
	CK(ctx->perm[1].ensure(n_ent * 4));
	CK(ctx->head.ensure(n_ent * 4));
	const unsigned ge = grid_for(n_ent, 256);
	iota_kernel<<<ge, 256, 0, st>>>(ctx->perm[0].as<uint32_t>(), n_ent);
	CK(cudaGetLastError());
	stat.kernel_launches += 1;
	// three stable LSD passes: (index, loc, strand), then word.lo, then word.hi
	auto sort_pass = [&](const uint64_t *key_in, int end_bit, int src, int dst) -> int {
		size_t tb = 0;
		CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, key_in, ctx->order_key[1].as<uint64_t>(), ctx->perm[src].as<uint32_t>(),
			ctx->perm[dst].as<uint32_t>(), (int64_t)n_ent, 0, end_bit, st));
		CK(ctx->cub_tmp.ensure(tb));
		CK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tb, key_in, ctx->order_key[1].as<uint64_t>(), ctx->perm[src].as<uint32_t>(),
			ctx->perm[dst].as<uint32_t>(), (int64_t)n_ent, 0, end_bit, st));
		stat.kernel_launches += 10;
		return 0;
	};
	if (sort_pass(s.e_order.as<uint64_t>(), (int)(34 + seq_bits), 0, 1)) return 1;
	gather_u64_kernel<<<ge, 256, 0, st>>>(s.e_lo.as<uint64_t>(), ctx->perm[1].as<uint32_t>(), ctx->order_key[0].as<uint64_t>(), n_ent);
	CK(cudaGetLastError());
	if (sort_pass(ctx->order_key[0].as<uint64_t>(), 64, 1, 0)) return 1;
	gather_u64_kernel<<<ge, 256, 0, st>>>(s.e_hi.as<uint64_t>(), ctx->perm[0].as<uint32_t>(), ctx->order_key[0].as<uint64_t>(), n_ent);
	CK(cudaGetLastError());
	if (sort_pass(ctx->order_key[0].as<uint64_t>(), 64, 0, 1)) return 1;
	CK(cudaMemcpyAsync(s.e_perm.p, ctx->perm[1].p, n_ent * 4, cudaMemcpyDeviceToDevice, st));
	key_heads_kernel<<<ge, 256, 0, st>>>(s.e_hi.as<uint64_t>(), s.e_lo.as<uint64_t>(), s.e_perm.as<uint32_t>(), ctx->head.as<uint32_t>(), n_ent);
	CK(cudaGetLastError());
	CK(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, ctx->head.as<uint32_t>(), s.e_keyrank.as<uint32_t>(), (int)n_ent, st));
	CK(ctx->cub_tmp.ensure(tmp_bytes));
	CK(cub::DeviceScan::InclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->head.as<uint32_t>(), s.e_keyrank.as<uint32_t>(), (int)n_ent, st));
	stat.kernel_launches += 5;
	uint32_t n_keys = 0;
	CK(cudaMemcpyAsync(&n_keys, s.e_keyrank.as<uint32_t>() + (n_ent - 1), 4, cudaMemcpyDeviceToHost, st));
	CK(s.e_key.ensure(n_ent * 4));
	CK(s.key_planes.ensure(n_ent * 16)); // at most one key per entry
	key_index_kernel<<<ge, 256, 0, st>>>(s.e_perm.as<uint32_t>(), ctx->head.as<uint32_t>(), s.e_keyrank.as<uint32_t>(), s.e_planes.as<uint4>(), n_ent,
		s.e_key.as<uint32_t>(), s.key_planes.as<uint4>());
	CK(cudaGetLastError());
	stat.kernel_launches++;
	CK(cudaStreamSynchronize(st));
	s.n_keys = n_keys;
	s.keys_valid = true;
	return 0;
}

static int select_words_general(pcramp_gpu_ctx *ctx, int kind, int opt5, int opt3, float threshold, uint32_t pack_max_degen,
	float min_gc, float max_gc, uint32_t min_len, uint64_t *n_entries_out, uint64_t *n_keys_out)
{
	if (check_kind(ctx, kind)) return 1;
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[kind];
	cudaStream_t st = ctx->stream;
	pcramp_gpu_stats &stat = ctx->stats;
	stat = pcramp_gpu_stats();
	bool used_edge_fst = false;
	ctx->fast_hint[kind].ok = false;
	ctx->pend_ms_db = ctx->pend_ms_score = false;
	s.db_valid = false;
	s.n_entries = s.n_keys = 0;
	if (n_entries_out) *n_entries_out = 0;
	if (n_keys_out) *n_keys_out = 0;
	const uint32_t n_pairs = ctx->n_pairs;
	if (min_len == 0) return fail(ctx, "pcramp_gpu_select_words: min_oligo_length must be >= 1");
	PackParams pp;
	pp.max_degen = pack_max_degen;
	pp.min_gc = min_gc;
	pp.max_gc = max_gc;
	pp.min_len = min_len;
	pp.gc_filter = (min_gc > 0.0f) || (max_gc < 1.0f); // sequence.cpp:102
	CK(s.seq_ent_off.ensure((size_t)(2 * s.n + 1) * 4));
	if (s.n == 0 || n_pairs == 0) { // select_words.cpp:13-15
		CK(cudaMemsetAsync(s.seq_ent_off.p, 0, (size_t)(2 * s.n + 1) * 4, st));
		CK(cudaStreamSynchronize(st));
		s.db_valid = true;
		return 0;
	}
	const SeqDev sd = s.dev();
	Trace tr("select_words", st);
	unsigned long long *d_cnt = nullptr;
	CK(ctx->d_counters.ensure(8 * sizeof(unsigned long long)));
	d_cnt = ctx->d_counters.as<unsigned long long>();

	// ---- candidates and patterns --------------------------------------------------------------
	const uint32_t n_oligo = 2u * n_pairs;
	CK(ctx->d_cand_cnt.ensure((size_t)n_oligo * 4));
	CK(ctx->d_cand_off.ensure((size_t)n_oligo * 4));
	cand_count_kernel<<<grid_for(n_oligo, 256), 256, 0, st>>>(ctx->pf(), ctx->pr(), n_pairs, opt5, opt3,
		ctx->d_cand_cnt.as<uint32_t>());
	CK(cudaGetLastError());
	stat.kernel_launches++;
	size_t tmp_bytes = 0;
	CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->d_cand_cnt.as<uint32_t>(), ctx->d_cand_off.as<uint32_t>(), (int)n_oligo, st));
	CK(ctx->cub_tmp.ensure(tmp_bytes));
	CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->d_cand_cnt.as<uint32_t>(), ctx->d_cand_off.as<uint32_t>(), (int)n_oligo, st));
	uint32_t last_off = 0, last_cnt = 0;
	CK(cudaMemcpyAsync(&last_off, ctx->d_cand_off.as<uint32_t>() + (n_oligo - 1), 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(&last_cnt, ctx->d_cand_cnt.as<uint32_t>() + (n_oligo - 1), 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	const uint32_t n_cand = last_off + last_cnt;
	if (n_cand > PAT_MAX_CAND) return fail(ctx, "pcramp_gpu_select_words: more than 2^20 candidate words in one batch");
	const uint32_t n_pat = 2u * n_oligo; // one pattern per (oligo, strand); shift families ride on it
	CK(ctx->d_cand_words.ensure((size_t)n_cand * 16));
	CK(ctx->d_cand_thr.ensure((size_t)n_cand * 4));
	CK(ctx->d_pat_mask.ensure((size_t)n_pat * 16));
	CK(ctx->d_pat_meta.ensure((size_t)n_pat * 4));
	CK(ctx->d_pat_meta2.ensure((size_t)n_pat * 4));
	CK(ctx->d_pat_seeded.ensure((size_t)n_pat * 4));
	CK(ctx->d_pat_sbefore.ensure((size_t)n_pat * 4));
	CK(ctx->d_part_mask.ensure((size_t)n_pat * 16));
	CK(ctx->d_part_meta.ensure((size_t)n_pat * 4));
	CK(ctx->d_part_meta2.ensure((size_t)n_pat * 4));
	cand_build_kernel<<<grid_for(n_oligo, 256), 256, 0, st>>>(ctx->pf(), ctx->pr(), n_pairs, opt5, opt3, threshold,
		ctx->d_cand_off.as<uint32_t>(), ctx->d_cand_words.as<uint4>(), ctx->d_cand_thr.as<uint32_t>(), ctx->d_pat_mask.as<uint4>(),
		ctx->d_pat_meta.as<uint32_t>(), ctx->d_pat_meta2.as<uint32_t>(), ctx->d_pat_seeded.as<uint32_t>());
	CK(cudaGetLastError());
	stat.kernel_launches++;
	CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->d_pat_seeded.as<uint32_t>(), ctx->d_pat_sbefore.as<uint32_t>(), (int)n_pat, st));
	CK(ctx->cub_tmp.ensure(tmp_bytes));
	CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->d_pat_seeded.as<uint32_t>(), ctx->d_pat_sbefore.as<uint32_t>(), (int)n_pat, st));
	uint32_t n_seeded = 0;
	{ // a zero threshold would select every word of every sequence (select_words.cpp:99-117): refuse
		std::vector<uint32_t> thr(n_cand);
		uint32_t sb = 0, sl = 0;
		CK(cudaMemcpyAsync(thr.data(), ctx->d_cand_thr.p, (size_t)n_cand * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(&sb, ctx->d_pat_sbefore.as<uint32_t>() + (n_pat - 1), 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(&sl, ctx->d_pat_seeded.as<uint32_t>() + (n_pat - 1), 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		for (uint32_t t : thr)
			if (t == 0) return fail(ctx, "pcramp_gpu_select_words: a candidate has match threshold 0 (empty oligo or threshold too low)");
		n_seeded = sb + sl;
	}
	if (ctx->force_brute) n_seeded = 0; // testing hook: everything through the brute-force kernel
	const uint32_t n_brute = n_pat - n_seeded;
	if (ctx->force_brute) {
		CK(cudaMemcpyAsync(ctx->d_part_mask.p, ctx->d_pat_mask.p, (size_t)n_pat * 16, cudaMemcpyDeviceToDevice, st));
		CK(cudaMemcpyAsync(ctx->d_part_meta.p, ctx->d_pat_meta.p, (size_t)n_pat * 4, cudaMemcpyDeviceToDevice, st));
		CK(cudaMemcpyAsync(ctx->d_part_meta2.p, ctx->d_pat_meta2.p, (size_t)n_pat * 4, cudaMemcpyDeviceToDevice, st));
	} else {
		pat_partition_kernel<<<grid_for(n_pat, 256), 256, 0, st>>>(ctx->d_pat_mask.as<uint4>(), ctx->d_pat_meta.as<uint32_t>(),
			ctx->d_pat_meta2.as<uint32_t>(), ctx->d_pat_seeded.as<uint32_t>(), ctx->d_pat_sbefore.as<uint32_t>(), n_pat, n_seeded,
			ctx->d_part_mask.as<uint4>(), ctx->d_part_meta.as<uint32_t>(), ctx->d_part_meta2.as<uint32_t>());
		CK(cudaGetLastError());
		stat.kernel_launches++;
	}
	const uint32_t cand_bits = bits_for(n_cand), seq_bits = bits_for(s.n);
	stat.n_patterns = 2ull * n_cand; // brute-force-equivalent pattern count (every family member, both strands)
	stat.n_seeded = n_seeded;
	stat.n_positions = 0;
	for (uint32_t i = 0; i < s.n; ++i)
		if (s.active[i]) stat.n_positions += s.clen[i];

	// ---- scan (re-run with a larger hit buffer if it overflowed) -------------------------------
	CK(ctx->d_seed_cnt.ensure(SEED_BUCKETS * 4));
	CK(ctx->d_seed_start.ensure(SEED_BUCKETS * 4));
	CK(ctx->d_seed_bucket.ensure(SEED_BUCKETS * 4));
	CK(ctx->d_tile_counter.ensure(16));
	uint64_t n_hits = 0;
	const bool tiny = ctx->tiny_buffers != 0; // testing hook: every growable buffer starts far too small, so the overflow re-runs engage
	uint64_t cap = std::max<uint64_t>(ctx->hit_key[0].cap / 8, tiny ? 2048ull : 1ull << 20);
	for (int attempt = 0;; ++attempt) {
		CK(ctx->hit_key[0].ensure(cap * 8));
		CK(ctx->hit_val[0].ensure(cap * 4));
		cap = std::min<uint64_t>(ctx->hit_key[0].cap / 8, ctx->hit_val[0].cap / 4);
		HitSink hs;
		hs.key = ctx->hit_key[0].as<uint64_t>();
		hs.val = ctx->hit_val[0].as<uint32_t>();
		hs.count = d_cnt;
		hs.cap = cap;
		CK(cudaMemsetAsync(d_cnt, 0, 8 * sizeof(unsigned long long), st));
		CK(cudaEventRecord(ctx->ev[0], st));
		// (a0) seeded patterns whose segment prefixes are plain k-mers: the indexed scan (index.cuh)
		bool use_idx = ctx->use_index && !ctx->force_brute && s.n_tiles && n_seeded;
		// sequences split since the index was built and active again: below a few per cent of the text the table scan covers
		// them (their index entries are ignored); above, the index is rebuilt
		std::vector<uint32_t> stale_seq;
		if (use_idx && s.idx_valid && s.n_idx_stale) {
			uint64_t stale_pos = 0;
			for (uint32_t i = 0; i < s.n; ++i)
				if (s.idx_stale[i] && s.active[i]) {
					stale_seq.push_back(i);
					stale_pos += s.clen[i];
				}
			if (stale_pos * 16ull > stat.n_positions && !ctx->parent) { // > 1/16 of the active text (workers never rebuild: they keep the table scan)
				s.idx_valid = false;
				stale_seq.clear();
			}
		}
		if (use_idx && !s.idx_valid && !s.idx_failed) {
			tr.mark("patterns");
			if (build_index(ctx, s)) return 1;
			tr.mark("build_index");
			CK(cudaEventRecord(ctx->ev[0], st)); // the one-time build is not part of the scan's time
		}
		use_idx = use_idx && s.idx_valid;
		stat.ms_index_build = s.idx_build_ms;
		stat.index_bytes = s.idx_bytes;
		stat.n_index_builds = s.idx_builds;
		stat.n_index_stale = (uint64_t)stale_seq.size();
		if (use_idx) {
			CK(ctx->d_idx_counters.ensure(32));
			unsigned int *d_nq = ctx->d_idx_counters.as<unsigned int>();
			stat.ms_index_kernel = 0.0f;
			for (const auto &part : s.idx_parts) {
				TextIndex ix;
				ix.entries = part->entries.as<uint4>();
				ix.off = part->off.as<uint32_t>();
				ix.cum = part->cum.as<uint32_t>();
				ix.blk = part->blk.as<uint32_t>();
				ix.n = part->n;
				ix.seq_lo = part->seq_lo;
				ix.n_seq = part->seq_hi - part->seq_lo;
				IdxCandSink cs;
				unsigned int h_idx[4] = {0, 0, 0, 0};
				for (int grow = 0;; ++grow) { // queries and candidates awaiting resolution: grown if they overflow (sizes repeat from batch to batch)
					// extended seeds expand a neighbour into up to IDX_EXT_MAX single-bucket queries: ~600 per 18-mer, fewer for longer primers
					const uint64_t qcap = std::min<uint64_t>(std::max<uint64_t>(ctx->d_idx_queries.cap / sizeof(IdxQuery), tiny ? 1024ull : (uint64_t)n_seeded * 640u), 0xFFFFFFF0ull);
					CK(ctx->d_idx_queries.ensure(qcap * sizeof(IdxQuery)));
					CK(cudaMemsetAsync(ctx->d_idx_counters.p, 0, 32, st));
					index_query_kernel<<<grid_for((uint64_t)n_seeded * IDX_SLOTS, 256), 256, 0, st>>>(ctx->d_part_mask.as<uint4>(),
						ctx->d_part_meta2.as<uint32_t>(), n_seeded, ix.off, ctx->d_idx_queries.as<IdxQuery>(), (uint32_t)qcap, d_nq, d_nq + 1,
						(unsigned long long *)(d_nq + 2));
					CK(cudaGetLastError());
					const uint64_t ccap = std::min<uint64_t>(std::max<uint64_t>(ctx->d_idx_cand.cap / sizeof(IdxCand), tiny ? 1024ull : cap), 0xFFFFFFF0ull);
					CK(ctx->d_idx_cand.ensure(ccap * sizeof(IdxCand)));
					cs.buf = ctx->d_idx_cand.as<IdxCand>();
					cs.count = d_nq + 4;
					cs.cap = (uint32_t)ccap;
					CK(cudaEventRecord(ctx->ev[8], st));
					if (launch_scan_index(ctx, st, ix, ctx->d_idx_queries.as<IdxQuery>(), d_nq, (uint32_t)qcap, ctx->d_part_mask.as<uint4>(),
							ctx->d_part_meta.as<uint32_t>(), cs)) return 1;
					CK(cudaGetLastError());
					CK(cudaEventRecord(ctx->ev[9], st));
					stat.kernel_launches++;
					unsigned int n_c = 0;
					CK(cudaMemcpyAsync(&n_c, d_nq + 4, 4, cudaMemcpyDeviceToHost, st));
					CK(cudaMemcpyAsync(h_idx, d_nq, 16, cudaMemcpyDeviceToHost, st)); // query / entry counters: same round trip
					CK(cudaStreamSynchronize(st));
					if (n_c <= cs.cap && h_idx[0] <= qcap) break;
					if (grow >= 3) return fail(ctx, "pcramp_gpu_select_words: index query / candidate buffers kept overflowing");
					if (h_idx[0] > qcap) CK(ctx->d_idx_queries.ensure(((size_t)h_idx[0] + h_idx[0] / 8 + 1024) * sizeof(IdxQuery)));
					if (n_c > cs.cap) CK(ctx->d_idx_cand.ensure(((size_t)n_c + n_c / 8 + 1024) * sizeof(IdxCand)));
				}
				stat.ms_index_kernel += ev_ms(ctx->ev[8], ctx->ev[9]);
				index_hits_kernel<<<(unsigned)ctx->sm_count * 8u, 256, 0, st>>>(sd, ix, cs, ctx->d_part_meta.as<uint32_t>(), ctx->d_part_meta2.as<uint32_t>(),
					s.n_dirty ? s.d_dirty_bits.as<uint32_t>() : nullptr, s.n_idx_stale ? s.d_idx_stale.as<uint8_t>() : nullptr, cand_bits, hs);
				CK(cudaGetLastError());
				stat.kernel_launches += 2;
				stat.n_index_queries += h_idx[0];
				stat.n_indexed = h_idx[1];
				stat.n_index_entries += (uint64_t)h_idx[2] | ((uint64_t)h_idx[3] << 32);
			}
		}
		tr.mark("indexed scan");
		// (a) the table-based seeded scan, in chunks of patterns that fit shared memory: the patterns the index did not take over
		// every sequence, and -- when split sequences are active again -- EVERY seeded pattern over the tiles of those sequences
		auto table_scan = [&](const uint32_t *d_tseq, const uint32_t *d_tx0, uint32_t n_tiles, int skip_idx, uint32_t done) -> int {
			uint32_t chunk = std::min<uint32_t>(n_seeded, 4096u);
			while (done < n_seeded) {
				const uint32_t cn = std::min<uint32_t>(chunk, n_seeded - done);
				const uint4 *c_mask = ctx->d_part_mask.as<uint4>() + done;
				const uint32_t *c_meta = ctx->d_part_meta.as<uint32_t>() + done, *c_meta2 = ctx->d_part_meta2.as<uint32_t>() + done;
				// count -> exclusive scan -> fill
				CK(cudaMemsetAsync(ctx->d_seed_cnt.p, 0, SEED_BUCKETS * 4, st));
				seed_count_kernel<<<grid_for(cn, 128), 128, 0, st>>>(c_mask, c_meta2, cn, ctx->d_seed_cnt.as<uint32_t>(), skip_idx);
				CK(cudaGetLastError());
				CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->d_seed_cnt.as<uint32_t>(), ctx->d_seed_start.as<uint32_t>(), (int)SEED_BUCKETS, st));
				CK(ctx->cub_tmp.ensure(tmp_bytes));
				CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->d_seed_cnt.as<uint32_t>(), ctx->d_seed_start.as<uint32_t>(), (int)SEED_BUCKETS, st));
				unsigned int *d_flag = ctx->d_tile_counter.as<unsigned int>() + 1;
				CK(cudaMemsetAsync(ctx->d_tile_counter.p, 0, 16, st));
				seed_overflow_kernel<<<grid_for(SEED_BUCKETS, 256), 256, 0, st>>>(ctx->d_seed_cnt.as<uint32_t>(), d_flag);
				CK(cudaGetLastError());
				uint32_t last_start = 0, last_cnt2 = 0, overflow = 0;
				CK(cudaMemcpyAsync(&last_start, ctx->d_seed_start.as<uint32_t>() + (SEED_BUCKETS - 1), 4, cudaMemcpyDeviceToHost, st));
				CK(cudaMemcpyAsync(&last_cnt2, ctx->d_seed_cnt.as<uint32_t>() + (SEED_BUCKETS - 1), 4, cudaMemcpyDeviceToHost, st));
				CK(cudaMemcpyAsync(&overflow, d_flag, 4, cudaMemcpyDeviceToHost, st));
				CK(cudaStreamSynchronize(st));
				stat.kernel_launches += 4;
				const uint32_t n_ent = last_start + last_cnt2;
				if (n_ent == 0u) { // every pattern of this chunk went through the index
					done += cn;
					continue;
				}
				const uint32_t ecap = (n_ent + 3u) & ~3u, pcap = (cn + 3u) & ~3u;
				const size_t smem = seed_smem_bytes(ecap, pcap);
				if (overflow || smem > (size_t)ctx->max_smem_optin || n_ent >= (1u << 20)) {
					if (cn <= 16) return fail(ctx, "pcramp_gpu_select_words: seed table does not fit shared memory");
					chunk = cn / 2; // fewer patterns per pass
					continue;
				}
				CK(ctx->d_seed_entries.ensure(std::max<size_t>(1, ecap) * 4));
				seed_bucket_pack_kernel<<<grid_for(SEED_BUCKETS, 256), 256, 0, st>>>(ctx->d_seed_start.as<uint32_t>(), ctx->d_seed_cnt.as<uint32_t>(),
					ctx->d_seed_bucket.as<uint32_t>());
				CK(cudaGetLastError());
				seed_fill_kernel<<<grid_for(cn, 128), 128, 0, st>>>(c_mask, c_meta, c_meta2, cn, ctx->d_seed_start.as<uint32_t>(),
					ctx->d_seed_entries.as<uint32_t>(), ecap, skip_idx);
				CK(cudaGetLastError());
				SeedChunk ch;
				ch.bucket = ctx->d_seed_bucket.as<uint32_t>();
				ch.entries = ctx->d_seed_entries.as<uint32_t>();
				ch.mask = c_mask;
				ch.meta = c_meta;
				ch.meta2 = c_meta2;
				ch.n_entries = n_ent;
				ch.n_pat = cn;
				ch.ecap = ecap;
				ch.pcap = pcap;
				const unsigned grid = (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count);
				scan_seed_kernel<<<grid, SEED_THREADS, smem, st>>>(sd, d_tseq, d_tx0, n_tiles, ctx->d_tile_counter.as<unsigned int>(), ch,
					s.n_dirty ? s.d_dirty_bits.as<uint32_t>() : nullptr, cand_bits, hs);
				CK(cudaGetLastError());
				stat.kernel_launches += 3;
				stat.n_seed_entries += n_ent;
				done += cn;
			}
			return 0;
		};
		if (s.n_tiles && n_seeded) {
			// nothing left for the table-based scan when the index took every seeded pattern
			if (table_scan(s.d_tile_seq.as<uint32_t>(), s.d_tile_x0.as<uint32_t>(), (uint32_t)s.n_tiles, use_idx ? 1 : 0,
					(use_idx && stat.n_indexed == n_seeded) ? n_seeded : 0u)) return 1;
			if (use_idx && !stale_seq.empty()) {
				std::vector<uint32_t> tseq, tx0;
				for (uint32_t q : stale_seq)
					for (uint64_t x = 0; x < s.clen[q]; x += SCAN_TILE) {
						tseq.push_back(q);
						tx0.push_back((uint32_t)x);
					}
				if (!tseq.empty()) {
					CK(ctx->d_stale_tile_seq.ensure(tseq.size() * 4));
					CK(ctx->d_stale_tile_x0.ensure(tx0.size() * 4));
					CK(cudaMemcpyAsync(ctx->d_stale_tile_seq.p, tseq.data(), tseq.size() * 4, cudaMemcpyHostToDevice, st));
					CK(cudaMemcpyAsync(ctx->d_stale_tile_x0.p, tx0.data(), tx0.size() * 4, cudaMemcpyHostToDevice, st));
					CK(cudaStreamSynchronize(st)); // (the host vectors go out of scope)
					// only the patterns the index took: the others were just scanned over every tile, these sequences included
					if (table_scan(ctx->d_stale_tile_seq.as<uint32_t>(), ctx->d_stale_tile_x0.as<uint32_t>(), (uint32_t)tseq.size(), 2, 0u)) return 1;
				}
			}
		// (b) the seeded patterns, brute force, on the groups whose text holds a degenerate base
			if (s.n_dirty) {
				scan_groups_kernel<<<(unsigned)std::min<uint64_t>(((uint64_t)s.n_dirty + 7) / 8, (uint64_t)ctx->sm_count * 8), 256, 0, st>>>(sd,
					s.d_dirty_seq.as<uint32_t>(), s.d_dirty_grp.as<uint32_t>(), s.n_dirty, ctx->d_part_mask.as<uint4>(),
					ctx->d_part_meta.as<uint32_t>(), ctx->d_part_meta2.as<uint32_t>(), n_seeded, cand_bits, hs);
				CK(cudaGetLastError());
				stat.kernel_launches++;
			}
		}
		tr.mark("table scan + dirty groups");
		CK(cudaEventRecord(ctx->ev[7], st));
		// (c) patterns that cannot be seeded: brute force over every alignment
		if (s.n_tiles && n_brute) {
			const unsigned grid = (unsigned)std::min<uint64_t>(s.n_tiles, (uint64_t)ctx->sm_count * 2);
			scan_full_kernel<<<grid, SCAN_THREADS, 0, st>>>(sd, s.d_tile_seq.as<uint32_t>(), s.d_tile_x0.as<uint32_t>(), (uint32_t)s.n_tiles,
				ctx->d_part_mask.as<uint4>() + n_seeded, ctx->d_part_meta.as<uint32_t>() + n_seeded, ctx->d_part_meta2.as<uint32_t>() + n_seeded,
				n_brute, cand_bits, hs);
			CK(cudaGetLastError());
			stat.kernel_launches++;
		}
		CK(cudaEventRecord(ctx->ev[1], st));
		tr.mark("brute-force scan");
		{
			bool edge_fst = ctx->use_fst != 0;
			Fst fst;
			if (edge_fst) { // the candidate words as a frame-aligned seed table; mostly unseedable candidates: compare with all
				uint32_t n_brute = 0;
				if (fst_build(ctx, ctx->d_cand_words.as<uint4>(), ctx->d_cand_thr.as<uint32_t>(), n_cand, fst, n_brute)) return 1;
				if ((uint64_t)n_brute * 4u > n_cand) edge_fst = false;
			}
			used_edge_fst = edge_fst;
			if (edge_fst)
				scan_edge_fst_kernel<<<(unsigned)std::min<uint64_t>(((uint64_t)s.n + 3) / 4, (uint64_t)ctx->sm_count * 16), 128, 0, st>>>(sd, pp, fst,
					cand_bits, hs);
			else
				scan_edge_kernel<<<(unsigned)std::min<uint64_t>(((uint64_t)s.n + 7) / 8, (uint64_t)ctx->sm_count * 8), 256, 0, st>>>(sd, pp,
					ctx->d_cand_words.as<uint4>(), ctx->d_cand_thr.as<uint32_t>(), n_cand, cand_bits, hs);
			CK(cudaGetLastError());
			stat.kernel_launches++;
		}
		CK(cudaEventRecord(ctx->ev[2], st));
		CK(cudaMemcpyAsync(ctx->h_counters, d_cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		n_hits = ctx->h_counters[0];
		stat.ms_seed += ev_ms(ctx->ev[0], ctx->ev[7]);
		stat.ms_scan += ev_ms(ctx->ev[7], ctx->ev[1]);
		stat.ms_edge += ev_ms(ctx->ev[1], ctx->ev[2]);
		if (n_hits <= cap) break;
		if (attempt >= 2) return fail(ctx, "pcramp_gpu_select_words: hit buffer kept overflowing");
		cap = n_hits + n_hits / 8 + 1024;
	}
	stat.n_hits = n_hits;
	tr.mark("partial words");
	CK(cudaEventRecord(ctx->ev[3], st));
	if (n_hits == 0) {
		CK(cudaMemsetAsync(s.seq_ent_off.p, 0, (size_t)(2 * s.n + 1) * 4, st));
		CK(cudaStreamSynchronize(st));
		s.db_valid = true;
		return 0;
	}

	// ---- validate, sort, tier ------------------------------------------------------------------
	if (pp.gc_filter || s.any_degenerate) {
		validate_hits_kernel<<<grid_for(n_hits, 256), 256, 0, st>>>(sd, pp, ctx->hit_key[0].as<uint64_t>(), ctx->hit_val[0].as<uint32_t>(), n_hits,
			cand_bits, nullptr);
		CK(cudaGetLastError());
		stat.kernel_launches++;
	}
	uint32_t longest = 0;
	for (uint32_t L : s.plen) longest = std::max(longest, L);
	const uint32_t pos_bits = std::min<uint32_t>(32u, bits_for((uint64_t)longest + 64ull)); // hit positions are below plen + 32
	const uint64_t tier_cells = (uint64_t)s.n * n_cand;
	uint32_t tier_span = 0; // largest (size - threshold count) over the oligo sizes: the tiers a candidate can have
	for (uint32_t sz = 1; sz <= 32u; ++sz) tier_span = std::max(tier_span, sz - std::min(sz, (uint32_t)((float)sz * threshold)));
	s.db_pp = pp;
	s.db_pb = pos_bits;
	if (ctx->use_seg_db && ctx->use_tier_table && tier_span <= 7u && tier_cells <= (1ull << 31) && pos_bits + 2u <= 32u && n_hits < (1ull << 32)) {
		// ---- segmented build (db.cuh): tiers, unique entries grouped by (sequence, strand) and sorted by position, planes -- no library
		//      sort, every size on the device
		const uint32_t n_seg = 2u * s.n;
		const uint64_t words = tier_cells / 4 + 1;
		CK(ctx->d_tier_best.ensure(words * 4));
		CK(ctx->seg_cnt.ensure((size_t)(n_seg + 1) * 4));
		CK(ctx->seg_off.ensure((size_t)(n_seg + 1) * 4));
		CK(ctx->seg_cursor.ensure((size_t)(n_seg + 1) * 4));
		CK(ctx->seg_uniq.ensure((size_t)(n_seg + 1) * 4));
		CK(ctx->seg_full.ensure((size_t)(n_seg + 1) * 4));
		CK(ctx->ent_id[0].ensure(n_hits * 8));
		CK(ctx->ent_cand[0].ensure(n_hits * 4));
		CK(s.e_planes.ensure(n_hits * 16)); // sized by the hits: the number of unique entries stays on the device until somebody asks
		CK(s.e_seq.ensure(n_hits * 4));
		CK(s.e_loc.ensure(n_hits * 4));
		CK(s.e_strand.ensure(n_hits * 4));
		CK(s.e_cand.ensure(n_hits * 4));
		CK(s.e_id.ensure(n_hits * 4));
		CK(s.seq_full_end.ensure((size_t)(n_seg + 1) * 4));
		CK(s.c_planes.ensure(std::max<size_t>(1, n_cand) * 16));
		CK(s.c_thr.ensure(std::max<size_t>(1, n_cand) * 4));
		CK(cudaMemsetAsync(ctx->d_tier_best.p, 0, words * 4, st));
		CK(cudaMemsetAsync(ctx->seg_cnt.p, 0, (size_t)(n_seg + 1) * 4, st));
		CK(cudaMemsetAsync(ctx->seg_cursor.p, 0, (size_t)(n_seg + 1) * 4, st));
		const unsigned gh = (unsigned)std::min<uint64_t>(grid_for(n_hits, 256), (uint64_t)ctx->sm_count * 32u);
		const unsigned gs = (unsigned)std::min<uint64_t>(n_seg, (uint64_t)ctx->sm_count * 32u);
		tier_mask_kernel<<<grid_for(n_hits, 256), 256, 0, st>>>(ctx->hit_key[0].as<uint64_t>(), n_hits, cand_bits, n_cand, ctx->d_cand_thr.as<uint32_t>(),
			ctx->d_tier_best.as<uint32_t>(), nullptr);
		seg_count_kernel<<<gh, 256, 0, st>>>(ctx->hit_key[0].as<uint64_t>(), d_cnt, n_hits, cand_bits, n_cand, ctx->d_cand_thr.as<uint32_t>(),
			ctx->d_tier_best.as<uint32_t>(), ctx->seg_cnt.as<uint32_t>());
		CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ctx->seg_cnt.as<uint32_t>(), ctx->seg_off.as<uint32_t>(), (int)(n_seg + 1), st));
		CK(ctx->cub_tmp.ensure(tmp_bytes));
		CK(ctx->seg_big.ensure((size_t)(n_seg + 2) * 4));
		CK(cudaMemsetAsync(ctx->seg_big.p, 0, 4, st)); // word 0: number of long segments; the list follows
		CK(cudaMemsetAsync(ctx->seg_uniq.as<uint32_t>() + n_seg, 0, 4, st));
		// (seg_cnt[n_seg] = 0: the scan's last output is the total)
		CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->seg_cnt.as<uint32_t>(), ctx->seg_off.as<uint32_t>(), (int)(n_seg + 1), st));
		seg_scatter_kernel<<<gh, 256, 0, st>>>(ctx->hit_key[0].as<uint64_t>(), ctx->hit_val[0].as<uint32_t>(), d_cnt, n_hits, cand_bits, n_cand,
			ctx->d_cand_thr.as<uint32_t>(), ctx->d_tier_best.as<uint32_t>(), pos_bits, ctx->seg_off.as<uint32_t>(), ctx->seg_cursor.as<uint32_t>(),
			ctx->ent_id[0].as<uint64_t>(), ctx->ent_cand[0].as<uint32_t>());
		seg_sort_small_kernel<<<(unsigned)std::min<uint64_t>(grid_for(n_seg, SEG_WARPS), (uint64_t)ctx->sm_count * 16u), SEG_WARPS * 32u, 0, st>>>(
			ctx->seg_off.as<uint32_t>(), n_seg, ctx->ent_id[0].as<uint64_t>(), pos_bits, ctx->seg_uniq.as<uint32_t>(), ctx->seg_full.as<uint32_t>(),
			ctx->seg_big.as<uint32_t>() + 1, ctx->seg_big.as<unsigned int>());
		seg_sort_big_kernel<<<(unsigned)ctx->sm_count * 2u, SEG_BIG_THREADS, 0, st>>>(ctx->seg_off.as<uint32_t>(), ctx->seg_big.as<uint32_t>() + 1,
			ctx->seg_big.as<unsigned int>(), ctx->ent_id[0].as<uint64_t>(), pos_bits, ctx->seg_uniq.as<uint32_t>(), ctx->seg_full.as<uint32_t>());
		CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->seg_uniq.as<uint32_t>(), s.seq_ent_off.as<uint32_t>(), (int)(n_seg + 1), st));
		seg_materialise_kernel<<<gh, 256, 0, st>>>(sd, pp, ctx->seg_off.as<uint32_t>(), s.seq_ent_off.as<uint32_t>(), ctx->seg_full.as<uint32_t>(), n_seg,
			ctx->ent_id[0].as<uint64_t>(), ctx->ent_cand[0].as<uint32_t>(), pos_bits, s.e_planes.as<uint4>(), s.e_seq.as<uint32_t>(), s.e_loc.as<int32_t>(),
			s.e_strand.as<uint32_t>(), s.e_cand.as<uint32_t>(), s.e_id.as<uint32_t>(), s.seq_full_end.as<uint32_t>());
		CK(cudaGetLastError());
		stat.kernel_launches += 10;
		CK(cudaMemcpyAsync(s.c_planes.p, ctx->d_cand_words.p, (size_t)n_cand * 16, cudaMemcpyDeviceToDevice, st));
		CK(cudaMemcpyAsync(s.c_thr.p, ctx->d_cand_thr.p, (size_t)n_cand * 4, cudaMemcpyDeviceToDevice, st));
		s.n_cand = n_cand;
		uint32_t n_ent32 = 0;
		CK(cudaMemcpyAsync(&n_ent32, s.seq_ent_off.as<uint32_t>() + n_seg, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		tr.mark("segmented database");
		CK(cudaEventRecord(ctx->ev[4], st));
		ctx->pend_ms_db = true;
		s.n_entries = n_ent32;
		s.n_keys = 0;
		s.keys_valid = false;
		s.words_valid = false; // e_hi / e_lo / e_order on demand (db_words)
		s.seq_bits = seq_bits;
		s.db_valid = true;
		stat.n_entries = n_ent32;
		// the next batch of this shape may run without the host in the loop: every pattern went through a one-part index, the
		// partial words through the seed table, nothing was brute-forced -- and its buffers are known to be large enough
		if (!opt5 && !opt3 && used_edge_fst && n_brute == 0u && stat.n_indexed == n_seeded && n_seeded == n_pat && s.idx_parts.size() == 1 &&
		    stat.n_index_stale == 0u) {
			pcramp_gpu_ctx::FastHint &h = ctx->fast_hint[kind];
			h.ok = true;
			h.n_pairs = n_pairs;
			h.threshold = threshold;
			h.pp = pp;
			h.hit_cap = std::min<uint64_t>(ctx->hit_key[0].cap / 8, ctx->hit_val[0].cap / 4);
			h.q_cap = ctx->d_idx_queries.cap / sizeof(IdxQuery);
			h.c_cap = ctx->d_idx_cand.cap / sizeof(IdxCand);
			h.fst_cap = ctx->d_fst_ids.cap / 4;
		}
		if (n_entries_out) *n_entries_out = n_ent32;
		if (n_keys_out) {
			if (db_finalize_keys(ctx, s)) return 1;
			stat.n_keys = s.n_keys;
			*n_keys_out = s.n_keys;
		}
		return 0;
	}
	CK(ctx->ent_id[0].ensure(n_hits * 8));
	CK(ctx->ent_id[1].ensure(n_hits * 8));
	CK(ctx->ent_cand[0].ensure(n_hits * 4));
	CK(ctx->ent_cand[1].ensure(n_hits * 4));
	CK(cudaMemsetAsync(d_cnt, 0, 8 * sizeof(unsigned long long), st));
	if (ctx->use_tier_table && tier_span <= 7u && tier_cells <= (1ull << 31)) {
		// best tier per (sequence, candidate) through a byte-per-cell table of tier bits (db.cuh): no sort of the hit list
		const uint64_t words = tier_cells / 4 + 1;
		CK(ctx->d_tier_best.ensure(words * 4));
		CK(cudaMemsetAsync(ctx->d_tier_best.p, 0, words * 4, st));
		tier_mask_kernel<<<grid_for(n_hits, 256), 256, 0, st>>>(ctx->hit_key[0].as<uint64_t>(), n_hits, cand_bits, n_cand, ctx->d_cand_thr.as<uint32_t>(),
			ctx->d_tier_best.as<uint32_t>(), nullptr);
		tier_mask_select_kernel<<<grid_for(n_hits, 256), 256, 0, st>>>(ctx->hit_key[0].as<uint64_t>(), ctx->hit_val[0].as<uint32_t>(), n_hits, cand_bits,
			n_cand, ctx->d_cand_thr.as<uint32_t>(), ctx->d_tier_best.as<uint32_t>(), pos_bits, ctx->ent_id[0].as<uint64_t>(),
			ctx->ent_cand[0].as<uint32_t>(), d_cnt);
		CK(cudaGetLastError());
		stat.kernel_launches += 2;
	} else if (ctx->use_tier_table && tier_cells <= (1ull << 27)) {
		// best tier per (sequence, candidate) through a table of maxima: no sort of the hit list
		CK(ctx->d_tier_best.ensure(std::max<uint64_t>(1, tier_cells) * 4));
		CK(cudaMemsetAsync(ctx->d_tier_best.p, 0, tier_cells * 4, st));
		tier_best_kernel<<<grid_for(n_hits, 256), 256, 0, st>>>(ctx->hit_key[0].as<uint64_t>(), n_hits, cand_bits, n_cand, ctx->d_tier_best.as<uint32_t>());
		tier_table_kernel<<<grid_for(n_hits, 256), 256, 0, st>>>(ctx->hit_key[0].as<uint64_t>(), ctx->hit_val[0].as<uint32_t>(), n_hits, cand_bits, n_cand,
			ctx->d_tier_best.as<uint32_t>(), pos_bits, ctx->ent_id[0].as<uint64_t>(), ctx->ent_cand[0].as<uint32_t>(), d_cnt);
		CK(cudaGetLastError());
		stat.kernel_launches += 2;
	} else {
	CK(ctx->hit_key[1].ensure(n_hits * 8));
	CK(ctx->hit_val[1].ensure(n_hits * 4));
	// all 64 bits when hits may have been invalidated to ~0 (they must sort last), else only the used field
	const int sort_end = (pp.gc_filter || s.any_degenerate) ? 64 : (int)(HIT_GROUP_SHIFT + cand_bits + seq_bits);
	CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx->hit_key[0].as<uint64_t>(), ctx->hit_key[1].as<uint64_t>(),
		ctx->hit_val[0].as<uint32_t>(), ctx->hit_val[1].as<uint32_t>(), (int64_t)n_hits, 3, sort_end, st));
	CK(ctx->cub_tmp.ensure(tmp_bytes));
	CK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp_bytes, ctx->hit_key[0].as<uint64_t>(), ctx->hit_key[1].as<uint64_t>(),
		ctx->hit_val[0].as<uint32_t>(), ctx->hit_val[1].as<uint32_t>(), (int64_t)n_hits, 3, sort_end, st));
	stat.kernel_launches += 8;
	tier_kernel<<<grid_for(n_hits, 256), 256, 0, st>>>(ctx->hit_key[1].as<uint64_t>(), ctx->hit_val[1].as<uint32_t>(), n_hits, cand_bits, pos_bits,
		ctx->ent_id[0].as<uint64_t>(), ctx->ent_cand[0].as<uint32_t>(), d_cnt);
	CK(cudaGetLastError());
	stat.kernel_launches++;
	}
	CK(cudaMemcpyAsync(ctx->h_counters, d_cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	const uint64_t n_flag = ctx->h_counters[0];
	tr.mark("tiers");
	if (n_flag == 0) {
		CK(cudaMemsetAsync(s.seq_ent_off.p, 0, (size_t)(2 * s.n + 1) * 4, st));
		CK(cudaStreamSynchronize(st));
		s.db_valid = true;
		return 0;
	}

	// ---- unique entries -------------------------------------------------------------------------
	// (the candidate of each flagged hit rides along; any one of an entry's candidates will do, so the first is kept)
	CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx->ent_id[0].as<uint64_t>(), ctx->ent_id[1].as<uint64_t>(), ctx->ent_cand[0].as<uint32_t>(),
		ctx->ent_cand[1].as<uint32_t>(), (int64_t)n_flag, 0, (int)(pos_bits + 3 + seq_bits), st));
	CK(ctx->cub_tmp.ensure(tmp_bytes));
	CK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp_bytes, ctx->ent_id[0].as<uint64_t>(), ctx->ent_id[1].as<uint64_t>(),
		ctx->ent_cand[0].as<uint32_t>(), ctx->ent_cand[1].as<uint32_t>(), (int64_t)n_flag, 0, (int)(pos_bits + 3 + seq_bits), st));
	CK(cub::DeviceSelect::UniqueByKey(nullptr, tmp_bytes, ctx->ent_id[1].as<uint64_t>(), ctx->ent_cand[1].as<uint32_t>(), ctx->ent_id[0].as<uint64_t>(),
		ctx->ent_cand[0].as<uint32_t>(), d_cnt + 1, (int64_t)n_flag, st));
	CK(ctx->cub_tmp.ensure(tmp_bytes));
	CK(cub::DeviceSelect::UniqueByKey(ctx->cub_tmp.p, tmp_bytes, ctx->ent_id[1].as<uint64_t>(), ctx->ent_cand[1].as<uint32_t>(),
		ctx->ent_id[0].as<uint64_t>(), ctx->ent_cand[0].as<uint32_t>(), d_cnt + 1, (int64_t)n_flag, st));
	stat.kernel_launches += 10;
	CK(cudaMemcpyAsync(ctx->h_counters, d_cnt, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	const uint64_t n_ent = ctx->h_counters[1];
	if (n_ent >= (1ull << 32)) return fail(ctx, "pcramp_gpu_select_words: database exceeds 2^32 entries");

	// ---- materialise + canonical order + keys ------------------------------------------------
	CK(s.e_hi.ensure(n_ent * 8));
	CK(s.e_lo.ensure(n_ent * 8));
	CK(s.e_planes.ensure(n_ent * 16));
	CK(s.e_seq.ensure(n_ent * 4));
	CK(s.e_loc.ensure(n_ent * 4));
	CK(s.e_strand.ensure(n_ent * 4));
	CK(s.e_order.ensure(n_ent * 8));
	CK(s.e_cand.ensure(n_ent * 4));
	CK(s.c_planes.ensure(std::max<size_t>(1, n_cand) * 16));
	CK(s.c_thr.ensure(std::max<size_t>(1, n_cand) * 4));
	CK(cudaMemcpyAsync(s.e_cand.p, ctx->ent_cand[0].p, n_ent * 4, cudaMemcpyDeviceToDevice, st));
	CK(cudaMemcpyAsync(s.c_planes.p, ctx->d_cand_words.p, (size_t)n_cand * 16, cudaMemcpyDeviceToDevice, st));
	CK(cudaMemcpyAsync(s.c_thr.p, ctx->d_cand_thr.p, (size_t)n_cand * 4, cudaMemcpyDeviceToDevice, st));
	s.n_cand = n_cand;
	const unsigned ge = grid_for(n_ent, 256);
	materialise_kernel<<<ge, 256, 0, st>>>(sd, pp, ctx->ent_id[0].as<uint64_t>(), n_ent, pos_bits, s.e_hi.as<uint64_t>(), s.e_lo.as<uint64_t>(),
		s.e_planes.as<uint4>(), s.e_seq.as<uint32_t>(), s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(), s.e_order.as<uint64_t>());
	CK(cudaGetLastError());
	seq_offsets_kernel<<<grid_for(2ull * s.n + 1, 256), 256, 0, st>>>(s.e_seq.as<uint32_t>(), s.e_strand.as<uint32_t>(), n_ent, s.n,
		s.seq_ent_off.as<uint32_t>());
	CK(cudaGetLastError());
	CK(s.seq_full_end.ensure((size_t)(2 * s.n + 1) * 4));
	seq_full_end_kernel<<<grid_for(2ull * s.n, 256), 256, 0, st>>>(ctx->ent_id[0].as<uint64_t>(), n_ent, s.n, pos_bits, s.seq_full_end.as<uint32_t>());
	CK(cudaGetLastError());
	stat.kernel_launches += 3;
	tr.mark("unique + materialise");
	CK(cudaEventRecord(ctx->ev[4], st));
	ctx->pend_ms_db = true; // no host round trip here: pcramp_gpu_get_stats waits for the event when somebody asks
	s.n_entries = n_ent;
	s.n_keys = 0;
	s.keys_valid = false;
	s.words_valid = true;
	s.seq_bits = seq_bits;
	s.db_valid = true;
	stat.n_entries = n_ent;
	if (n_entries_out) *n_entries_out = n_ent;
	if (n_keys_out) { // keys() is only materialised for callers that ask for it (db_copy / keys_copy / this count)
		if (db_finalize_keys(ctx, s)) return 1;
		stat.n_keys = s.n_keys;
		*n_keys_out = s.n_keys;
	}
	return 0;
}

// ---- the fast form of select_words ------------------------------------------------------------------------------------------------
// A design run, a sweep, a target-sharded job all send batch after batch of the same shape.  Once a batch of a shape has gone
// through the general form above -- every pattern through the (one-part) text index, nothing brute-forced, the partial words through
// the seed table -- the next one is launched WITHOUT the host reading anything back: candidate i = oligo i (no shift families), the
// buffers keep the sizes that were enough last time, every kernel takes its counts from device memory, the partial-word scan runs
// on a second stream beside the indexed scan, and one small kernel gathers the counters and flags that say whether the
// assumptions held.  Whoever needs the database next (pair scoring, a copy, the statistics) waits for that one read-back
// (fast_resolve); if an assumption failed -- a buffer overflowed, a pattern was not indexable -- the batch is run again in the
// general form, so the result is the same either way.
static int select_words_fast(pcramp_gpu_ctx *ctx, int kind, float threshold, const PackParams &pp)
{
	SeqSet &s = ctx->sets[kind];
	cudaStream_t st = ctx->stream, st2 = ctx->stream2;
	pcramp_gpu_stats &stat = ctx->stats;
	const pcramp_gpu_ctx::FastHint &h = ctx->fast_hint[kind];
	stat = pcramp_gpu_stats();
	ctx->pend_ms_db = ctx->pend_ms_score = false;
	const uint32_t n_pairs = ctx->n_pairs, n_oligo = 2u * n_pairs, n_cand = n_oligo, n_pat = 2u * n_oligo, n_seg = 2u * s.n;
	const uint32_t cand_bits = bits_for(n_cand), seq_bits = bits_for(s.n);
	const SeqDev sd = s.dev();
	uint32_t longest = 0;
	for (uint32_t L : s.plen) longest = std::max(longest, L);
	const uint32_t pos_bits = std::min<uint32_t>(32u, bits_for((uint64_t)longest + 64ull));
	const uint64_t tier_cells = (uint64_t)s.n * n_cand, tier_words = tier_cells / 4 + 1;
	const uint64_t cap = h.hit_cap;
	// buffers: everything sized by capacities the host already knows
	CK(ctx->d_counters.ensure(8 * sizeof(unsigned long long)));
	CK(ctx->d_fast.ensure(8 * sizeof(unsigned long long)));
	CK(ctx->d_idx_counters.ensure(32));
	CK(ctx->d_part_mask.ensure((size_t)n_pat * 16));
	CK(ctx->d_part_meta.ensure((size_t)n_pat * 4));
	CK(ctx->d_part_meta2.ensure((size_t)n_pat * 4));
	CK(s.c_planes.ensure((size_t)n_cand * 16));
	CK(s.c_thr.ensure((size_t)n_cand * 4));
	CK(ctx->hit_key[0].ensure(cap * 8));
	CK(ctx->hit_val[0].ensure(cap * 4));
	CK(ctx->d_idx_queries.ensure(h.q_cap * sizeof(IdxQuery)));
	CK(ctx->d_idx_cand.ensure(h.c_cap * sizeof(IdxCand)));
	CK(ctx->d_fst_cnt.ensure((size_t)(FST_BUCKETS + 1) * 4));
	CK(ctx->d_fst_start.ensure((size_t)(FST_BUCKETS + 1) * 4));
	CK(ctx->d_fst_cursor.ensure((size_t)(FST_BUCKETS + 1) * 4));
	CK(ctx->d_fst_combo.ensure((2 * FST_COMBOS + 1) * 4));
	CK(ctx->d_fst_brute.ensure(std::max<size_t>(1, n_cand) * 4));
	CK(ctx->d_fst_nbrute.ensure(16));
	CK(ctx->d_fst_ids.ensure(std::max<uint64_t>(1, h.fst_cap) * 4));
	CK(ctx->d_tier_best.ensure(tier_words * 4));
	CK(ctx->seg_cnt.ensure((size_t)(n_seg + 1) * 4));
	CK(ctx->seg_off.ensure((size_t)(n_seg + 1) * 4));
	CK(ctx->seg_cursor.ensure((size_t)(n_seg + 1) * 4));
	CK(ctx->seg_uniq.ensure((size_t)(n_seg + 1) * 4));
	CK(ctx->seg_full.ensure((size_t)(n_seg + 1) * 4));
	CK(ctx->seg_big.ensure((size_t)(n_seg + 2) * 4));
	CK(ctx->ent_id[0].ensure(cap * 8));
	CK(ctx->ent_cand[0].ensure(cap * 4));
	CK(s.e_planes.ensure(cap * 16));
	CK(s.e_seq.ensure(cap * 4));
	CK(s.e_loc.ensure(cap * 4));
	CK(s.e_strand.ensure(cap * 4));
	CK(s.e_cand.ensure(cap * 4));
	CK(s.e_id.ensure(cap * 4));
	CK(s.seq_ent_off.ensure((size_t)(n_seg + 1) * 4));
	CK(s.seq_full_end.ensure((size_t)(n_seg + 1) * 4));
	size_t tmp_bytes = 0;
	CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)std::max<uint32_t>(n_seg + 1, FST_BUCKETS + 1), st));
	CK(ctx->cub_tmp.ensure(tmp_bytes));
	unsigned long long *d_cnt = ctx->d_counters.as<unsigned long long>();
	unsigned long long *d_flags = ctx->d_fast.as<unsigned long long>();
	unsigned int *d_nq = ctx->d_idx_counters.as<unsigned int>();
	HitSink hs;
	hs.key = ctx->hit_key[0].as<uint64_t>();
	hs.val = ctx->hit_val[0].as<uint32_t>();
	hs.count = d_cnt;
	hs.cap = cap;
	// ---- the collection's partial words as a table (edge.cuh): built on the first fast batch after an upload / split ---------------
	bool use_et = ctx->use_edge_table && ctx->edge_off_thr[kind] != threshold;
	if (use_et) {
		SeqSet::EdgeTab &t = s.edge;
		if (t.pp.max_degen != pp.max_degen || t.pp.min_gc != pp.min_gc || t.pp.max_gc != pp.max_gc || t.pp.min_len != pp.min_len ||
		    t.pp.gc_filter != pp.gc_filter)
			t.drop();
		if (!t.valid && !t.failed && edge_table_build(ctx, s, pp)) return 1;
		use_et = t.valid;
	}
	// ---- candidates + patterns ------------------------------------------------------------------------------------------
	CK(cudaEventRecord(ctx->ev[0], st));
	CK(cudaMemsetAsync(d_cnt, 0, 8 * sizeof(unsigned long long), st));
	CK(cudaMemsetAsync(d_flags, 0, 8 * sizeof(unsigned long long), st));
	CK(cudaMemsetAsync(d_nq, 0, 32, st));
	CK(cudaMemsetAsync(ctx->d_fst_nbrute.p, 0, 16, st));
	cand_build_direct_kernel<<<grid_for(n_oligo, 256), 256, 0, st>>>(ctx->pf(), ctx->pr(), n_pairs, threshold, s.c_planes.as<uint4>(), s.c_thr.as<uint32_t>(),
		ctx->d_part_mask.as<uint4>(), ctx->d_part_meta.as<uint32_t>(), ctx->d_part_meta2.as<uint32_t>(), d_flags);
	CK(cudaGetLastError());
	CK(cudaEventRecord(ctx->ev_fork, st));
	// ---- second stream: the partial words through the seed table over the candidates ---------------------------------------
	CK(cudaStreamWaitEvent(st2, ctx->ev_fork, 0));
	if (use_et) { // the candidates look themselves up in the table of partial words
		EdgeTable et;
		et.planes = s.edge.planes.as<uint4>();
		et.meta = s.edge.meta.as<uint4>();
		et.start = s.edge.start.as<uint32_t>();
		et.ids = s.edge.ids.as<uint32_t>();
		et.degen = s.edge.degen.as<uint32_t>();
		et.n_words = s.edge.n_words;
		et.n_degen = s.edge.n_degen;
		const uint64_t tasks = std::max<uint64_t>((uint64_t)n_cand * EDGE_MAX_PIECES, std::min<uint64_t>(et.n_degen, (uint64_t)ctx->sm_count * 16));
		edge_lookup_kernel<<<(unsigned)std::max<uint64_t>(1, tasks), EDGE_THREADS, 0, st2>>>(sd, et,
			s.c_planes.as<uint4>(), s.c_thr.as<uint32_t>(), n_cand, cand_bits, hs, ctx->d_fst_nbrute.as<unsigned int>() + 1);
		CK(cudaGetLastError());
		CK(cudaEventRecord(ctx->ev_join, st2));
	} else {
		Fst fst;
		CK(cudaMemsetAsync(ctx->d_fst_cnt.p, 0, (size_t)(FST_BUCKETS + 1) * 4, st2));
		CK(cudaMemsetAsync(ctx->d_fst_combo.p, 0, FST_COMBOS * 4, st2));
		fst_build_kernel<<<grid_for(n_cand, 128), 128, 0, st2>>>(s.c_planes.as<uint4>(), s.c_thr.as<uint32_t>(), n_cand, ctx->d_fst_cnt.as<uint32_t>(),
			ctx->d_fst_combo.as<uint32_t>(), ctx->d_fst_brute.as<uint32_t>(), ctx->d_fst_nbrute.as<uint32_t>(), nullptr, 0u, nullptr);
		fst_combo_list_kernel<<<1, 32, 0, st2>>>(ctx->d_fst_combo.as<uint32_t>());
		// (a second scratch for the scan: the main stream scans too)
		CK(ctx->cub_tmp2.ensure(tmp_bytes));
		size_t tb = tmp_bytes;
		CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp2.p, tb, ctx->d_fst_cnt.as<uint32_t>(), ctx->d_fst_start.as<uint32_t>(), (int)(FST_BUCKETS + 1), st2));
		CK(cudaMemcpyAsync(ctx->d_fst_cursor.p, ctx->d_fst_start.p, (size_t)(FST_BUCKETS + 1) * 4, cudaMemcpyDeviceToDevice, st2));
		fst_build_kernel<<<grid_for(n_cand, 128), 128, 0, st2>>>(s.c_planes.as<uint4>(), s.c_thr.as<uint32_t>(), n_cand, ctx->d_fst_cursor.as<uint32_t>(),
			ctx->d_fst_combo.as<uint32_t>(), ctx->d_fst_brute.as<uint32_t>(), ctx->d_fst_nbrute.as<uint32_t>(), ctx->d_fst_ids.as<uint32_t>(),
			(uint32_t)std::min<uint64_t>(h.fst_cap, 0xFFFFFFFFull), ctx->d_fst_nbrute.as<unsigned int>() + 1);
		fst.planes = s.c_planes.as<uint4>();
		fst.thr = s.c_thr.as<uint32_t>();
		fst.start = ctx->d_fst_start.as<uint32_t>();
		fst.ids = ctx->d_fst_ids.as<uint32_t>();
		fst.combo = ctx->d_fst_combo.as<uint32_t>();
		fst.brute = ctx->d_fst_brute.as<uint32_t>();
		fst.n_brute = ctx->d_fst_nbrute.as<uint32_t>();
		fst.n = n_cand;
		scan_edge_fst_kernel<<<(unsigned)std::min<uint64_t>(((uint64_t)s.n + 3) / 4, (uint64_t)ctx->sm_count * 16), 128, 0, st2>>>(sd, pp, fst, cand_bits, hs);
		CK(cudaGetLastError());
		CK(cudaEventRecord(ctx->ev_join, st2));
	}
	// ---- main stream: the indexed scan --------------------------------------------------------------------------------------
	{
		const SeqSet::IndexPart &part = *s.idx_parts[0];
		TextIndex ix;
		ix.entries = part.entries.as<uint4>();
		ix.off = part.off.as<uint32_t>();
		ix.cum = part.cum.as<uint32_t>();
		ix.blk = part.blk.as<uint32_t>();
		ix.n = part.n;
		ix.seq_lo = part.seq_lo;
		ix.n_seq = part.seq_hi - part.seq_lo;
		IdxCandSink cs;
		cs.buf = ctx->d_idx_cand.as<IdxCand>();
		cs.count = d_nq + 4;
		cs.cap = (uint32_t)std::min<uint64_t>(h.c_cap, 0xFFFFFFF0ull);
		const uint32_t qcap = (uint32_t)std::min<uint64_t>(h.q_cap, 0xFFFFFFF0ull);
		index_query_kernel<<<grid_for((uint64_t)n_pat * IDX_SLOTS, 256), 256, 0, st>>>(ctx->d_part_mask.as<uint4>(), ctx->d_part_meta2.as<uint32_t>(), n_pat,
			ix.off, ctx->d_idx_queries.as<IdxQuery>(), qcap, d_nq, d_nq + 1, (unsigned long long *)(d_nq + 2));
		CK(cudaEventRecord(ctx->ev[8], st));
		if (launch_scan_index(ctx, st, ix, ctx->d_idx_queries.as<IdxQuery>(), d_nq, qcap, ctx->d_part_mask.as<uint4>(), ctx->d_part_meta.as<uint32_t>(), cs))
			return 1;
		CK(cudaEventRecord(ctx->ev[9], st));
		index_hits_kernel<<<(unsigned)ctx->sm_count * 8u, 256, 0, st>>>(sd, ix, cs, ctx->d_part_meta.as<uint32_t>(), ctx->d_part_meta2.as<uint32_t>(),
			s.n_dirty ? s.d_dirty_bits.as<uint32_t>() : nullptr, nullptr, cand_bits, hs);
		CK(cudaGetLastError());
		stat.kernel_launches += 4;
		if (s.n_dirty) { // the seeded patterns, brute force, on the groups whose text holds a degenerate base
			scan_groups_kernel<<<(unsigned)std::min<uint64_t>(((uint64_t)s.n_dirty + 7) / 8, (uint64_t)ctx->sm_count * 8), 256, 0, st>>>(sd,
				s.d_dirty_seq.as<uint32_t>(), s.d_dirty_grp.as<uint32_t>(), s.n_dirty, ctx->d_part_mask.as<uint4>(), ctx->d_part_meta.as<uint32_t>(),
				ctx->d_part_meta2.as<uint32_t>(), n_pat, cand_bits, hs);
			CK(cudaGetLastError());
			stat.kernel_launches++;
		}
	}
	CK(cudaEventRecord(ctx->ev[7], st));
	CK(cudaEventRecord(ctx->ev[1], st));
	CK(cudaStreamWaitEvent(st, ctx->ev_join, 0)); // every hit is in the list from here on
	CK(cudaEventRecord(ctx->ev[2], st));
	CK(cudaEventRecord(ctx->ev[3], st));
	stat.kernel_launches += use_et ? 1 : 6;
	stat.edge_table_used = use_et ? 1u : 0u;
	if (use_et) {
		stat.n_edge_words = s.edge.n_words;
		stat.edge_table_bytes = s.edge.bytes;
		stat.ms_edge_table_build = s.edge.build_ms;
	}
	// ---- database (db.cuh, segmented form; the hit count stays on the device) -----------------------------------------------
	if (pp.gc_filter || s.any_degenerate) {
		validate_hits_kernel<<<grid_for(cap, 256), 256, 0, st>>>(sd, pp, hs.key, hs.val, cap, cand_bits, d_cnt);
		stat.kernel_launches++;
	}
	CK(cudaMemsetAsync(ctx->d_tier_best.p, 0, tier_words * 4, st));
	CK(cudaMemsetAsync(ctx->seg_cnt.p, 0, (size_t)(n_seg + 1) * 4, st));
	CK(cudaMemsetAsync(ctx->seg_cursor.p, 0, (size_t)(n_seg + 1) * 4, st));
	CK(cudaMemsetAsync(ctx->seg_big.p, 0, 4, st));
	CK(cudaMemsetAsync(ctx->seg_uniq.as<uint32_t>() + n_seg, 0, 4, st));
	const unsigned gh = (unsigned)std::min<uint64_t>(grid_for(cap, 256), (uint64_t)ctx->sm_count * 32u);
	tier_mask_kernel<<<grid_for(cap, 256), 256, 0, st>>>(hs.key, cap, cand_bits, n_cand, s.c_thr.as<uint32_t>(), ctx->d_tier_best.as<uint32_t>(), d_cnt);
	seg_count_kernel<<<gh, 256, 0, st>>>(hs.key, d_cnt, cap, cand_bits, n_cand, s.c_thr.as<uint32_t>(), ctx->d_tier_best.as<uint32_t>(), ctx->seg_cnt.as<uint32_t>());
	CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->seg_cnt.as<uint32_t>(), ctx->seg_off.as<uint32_t>(), (int)(n_seg + 1), st));
	seg_scatter_kernel<<<gh, 256, 0, st>>>(hs.key, hs.val, d_cnt, cap, cand_bits, n_cand, s.c_thr.as<uint32_t>(), ctx->d_tier_best.as<uint32_t>(), pos_bits,
		ctx->seg_off.as<uint32_t>(), ctx->seg_cursor.as<uint32_t>(), ctx->ent_id[0].as<uint64_t>(), ctx->ent_cand[0].as<uint32_t>());
	seg_sort_small_kernel<<<(unsigned)std::min<uint64_t>(grid_for(n_seg, SEG_WARPS), (uint64_t)ctx->sm_count * 16u), SEG_WARPS * 32u, 0, st>>>(
		ctx->seg_off.as<uint32_t>(), n_seg, ctx->ent_id[0].as<uint64_t>(), pos_bits, ctx->seg_uniq.as<uint32_t>(), ctx->seg_full.as<uint32_t>(),
		ctx->seg_big.as<uint32_t>() + 1, ctx->seg_big.as<unsigned int>());
	seg_sort_big_kernel<<<(unsigned)ctx->sm_count * 2u, SEG_BIG_THREADS, 0, st>>>(ctx->seg_off.as<uint32_t>(), ctx->seg_big.as<uint32_t>() + 1,
		ctx->seg_big.as<unsigned int>(), ctx->ent_id[0].as<uint64_t>(), pos_bits, ctx->seg_uniq.as<uint32_t>(), ctx->seg_full.as<uint32_t>());
	CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp_bytes, ctx->seg_uniq.as<uint32_t>(), s.seq_ent_off.as<uint32_t>(), (int)(n_seg + 1), st));
	seg_materialise_kernel<<<gh, 256, 0, st>>>(sd, pp, ctx->seg_off.as<uint32_t>(), s.seq_ent_off.as<uint32_t>(), ctx->seg_full.as<uint32_t>(), n_seg,
		ctx->ent_id[0].as<uint64_t>(), ctx->ent_cand[0].as<uint32_t>(), pos_bits, s.e_planes.as<uint4>(), s.e_seq.as<uint32_t>(), s.e_loc.as<int32_t>(),
		s.e_strand.as<uint32_t>(), s.e_cand.as<uint32_t>(), s.e_id.as<uint32_t>(), s.seq_full_end.as<uint32_t>());
	fast_gather_kernel<<<1, 1, 0, st>>>(d_flags, d_cnt, d_nq, ctx->d_fst_nbrute.as<unsigned int>(), s.seq_ent_off.as<uint32_t>(), n_seg);
	CK(cudaGetLastError());
	stat.kernel_launches += 11;
	CK(cudaEventRecord(ctx->ev[4], st));
	CK(cudaMemcpyAsync(ctx->h_fast, d_flags, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	CK(cudaEventRecord(ctx->ev_done, st));
	// host-side state of the database: valid, its size pending
	s.n_cand = n_cand;
	s.db_pp = pp;
	s.db_pb = pos_bits;
	s.n_entries = 0;
	s.n_keys = 0;
	s.keys_valid = false;
	s.words_valid = false;
	s.seq_bits = seq_bits;
	s.db_valid = true;
	stat.n_patterns = 2ull * n_cand;
	stat.n_seeded = n_pat;
	stat.n_positions = 0;
	for (uint32_t i = 0; i < s.n; ++i)
		if (s.active[i]) stat.n_positions += s.clen[i];
	stat.ms_index_build = s.idx_build_ms;
	stat.index_bytes = s.idx_bytes;
	stat.n_index_builds = s.idx_builds;
	pcramp_gpu_ctx::FastPending &p = ctx->fast_pending;
	p.active = true;
	p.kind = kind;
	p.threshold = threshold;
	p.pp = pp;
	p.n_pat = n_pat;
	p.n_seg = n_seg;
	p.hit_cap = cap;
	p.q_cap = std::min<uint64_t>(h.q_cap, 0xFFFFFFF0ull);
	p.c_cap = std::min<uint64_t>(h.c_cap, 0xFFFFFFF0ull);
	ctx->n_fast++;
	return 0;
}

extern "C++" {
namespace {
int fast_resolve(pcramp_gpu_ctx *ctx)
{
	if (!ctx->fast_pending.active) return 0;
	const pcramp_gpu_ctx::FastPending p = ctx->fast_pending;
	ctx->fast_pending.active = false;
	CK(cudaSetDevice(ctx->device));
	CK(cudaEventSynchronize(ctx->ev_done));
	const unsigned long long *h = ctx->h_fast;
	SeqSet &s = ctx->sets[p.kind];
	pcramp_gpu_stats &stat = ctx->stats;
	const bool ok = h[0] == 0ull && h[1] <= p.hit_cap && h[2] <= p.q_cap && h[5] <= p.c_cap && h[3] == p.n_pat && h[6] < (1ull << 32);
	if (ok) {
		s.n_entries = h[6];
		stat.n_hits = h[1];
		stat.n_entries = h[6];
		stat.n_index_queries = h[2];
		stat.n_indexed = h[3];
		stat.n_index_entries = h[4];
		stat.ms_seed = ev_ms(ctx->ev[0], ctx->ev[7]);
		stat.ms_scan = 0.0f;
		stat.ms_edge = ev_ms(ctx->ev[1], ctx->ev[2]); // what is left of the partial-word scan after the indexed scan it runs beside
		stat.ms_index_kernel = ev_ms(ctx->ev[8], ctx->ev[9]);
		ctx->pend_ms_db = true;
		return 0;
	}
	// an assumption did not hold (a buffer was too small, a pattern not indexable, a zero threshold): the general form decides
	if (h[0] & (unsigned long long)EDGE_FLAG_UNSUITABLE) ctx->edge_off_thr[p.kind] = p.threshold; // pieces shorter than the table's runs
	ctx->fast_hint[p.kind].ok = false;
	ctx->n_fast_redo++;
	return select_words_general(ctx, p.kind, 0, 0, p.threshold, p.pp.max_degen, p.pp.min_gc, p.pp.max_gc, p.pp.min_len, nullptr, nullptr);
}
} // namespace
} // extern "C++"

int pcramp_gpu_select_words_staged(pcramp_gpu_ctx *ctx, int kind, int opt5, int opt3, float threshold, uint32_t pack_max_degen,
	float min_gc, float max_gc, uint32_t min_len, uint64_t *n_entries_out, uint64_t *n_keys_out)
{
	if (check_kind(ctx, kind)) return 1; // (also settles a batch the fast form left unverified)
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[kind];
	const pcramp_gpu_ctx::FastHint &h = ctx->fast_hint[kind];
	PackParams pp;
	pp.max_degen = pack_max_degen;
	pp.min_gc = min_gc;
	pp.max_gc = max_gc;
	pp.min_len = min_len;
	pp.gc_filter = (min_gc > 0.0f) || (max_gc < 1.0f);
	bool fast = ctx->use_fast && h.ok && !opt5 && !opt3 && h.n_pairs == ctx->n_pairs && h.threshold == threshold && h.pp.max_degen == pp.max_degen &&
	            h.pp.min_gc == pp.min_gc && h.pp.max_gc == pp.max_gc && h.pp.min_len == pp.min_len && ctx->use_index && !ctx->force_brute &&
	            ctx->use_fst && ctx->use_seg_db && ctx->use_tier_table && s.idx_valid && s.idx_parts.size() == 1 && s.n > 0 &&
	            ctx->n_pairs > 0;
	if (fast && s.n_idx_stale) // split sequences that are active again need the table scan: the general form
		for (uint32_t i = 0; i < s.n && fast; ++i)
			if (s.idx_stale[i] && s.active[i]) fast = false;
	if (!fast) return select_words_general(ctx, kind, opt5, opt3, threshold, pack_max_degen, min_gc, max_gc, min_len, n_entries_out, n_keys_out);
	if (n_entries_out) *n_entries_out = 0;
	if (n_keys_out) *n_keys_out = 0;
	if (select_words_fast(ctx, kind, threshold, pp)) return 1;
	if (n_entries_out || n_keys_out) { // the caller wants sizes now: settle the batch
		if (fast_resolve(ctx)) return 1;
		if (n_entries_out) *n_entries_out = s.n_entries;
		if (n_keys_out) {
			if (db_finalize_keys(ctx, s)) return 1;
			ctx->stats.n_keys = s.n_keys;
			*n_keys_out = s.n_keys;
		}
	}
	return 0;
}

int pcramp_gpu_select_words(pcramp_gpu_ctx *ctx, int kind, const uint64_t *f, const uint64_t *r, uint32_t n_pairs, int opt5, int opt3,
	float threshold, uint32_t pack_max_degen, float min_gc, float max_gc, uint32_t min_len, uint64_t *n_entries, uint64_t *n_keys)
{
	if (pcramp_gpu_stage_pairs(ctx, f, r, n_pairs)) return 1;
	return pcramp_gpu_select_words_staged(ctx, kind, opt5, opt3, threshold, pack_max_degen, min_gc, max_gc, min_len, n_entries, n_keys);
}

int pcramp_gpu_db_copy(pcramp_gpu_ctx *ctx, int kind, uint64_t *words, uint32_t *index, int32_t *loc, uint32_t *strand, uint32_t *key_index)
{
	if (check_kind(ctx, kind)) return 1;
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[kind];
	if (!s.db_valid) return fail(ctx, "pcramp_gpu_db_copy: no database (call pcramp_gpu_select_words first)");
	if (db_finalize_keys(ctx, s)) return 1;
	const uint64_t n = s.n_entries;
	if (n == 0) return 0;
	DevBuf o_words, o_index, o_loc, o_strand, o_key, o_keys;
	CK(o_words.ensure(n * 16));
	CK(o_index.ensure(n * 4));
	CK(o_loc.ensure(n * 4));
	CK(o_strand.ensure(n * 4));
	CK(o_key.ensure(n * 4));
	CK(o_keys.ensure(std::max<uint64_t>(1, s.n_keys) * 16));
	db_export_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(s.e_hi.as<uint64_t>(), s.e_lo.as<uint64_t>(), s.e_seq.as<uint32_t>(),
		s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(), s.e_perm.as<uint32_t>(), s.e_keyrank.as<uint32_t>(), n, o_words.as<uint64_t>(),
		o_index.as<uint32_t>(), o_loc.as<int32_t>(), o_strand.as<uint32_t>(), o_key.as<uint32_t>(), o_keys.as<uint64_t>());
	CK(cudaGetLastError());
	if (words) CK(cudaMemcpyAsync(words, o_words.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
	if (index) CK(cudaMemcpyAsync(index, o_index.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (loc) CK(cudaMemcpyAsync(loc, o_loc.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (strand) CK(cudaMemcpyAsync(strand, o_strand.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (key_index) CK(cudaMemcpyAsync(key_index, o_key.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

int pcramp_gpu_db_size(pcramp_gpu_ctx *ctx, int kind, uint64_t *n_entries, uint64_t *n_keys)
{
	if (check_kind(ctx, kind)) return 1;
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[kind];
	if (!s.db_valid) return fail(ctx, "pcramp_gpu_db_size: no database (call pcramp_gpu_select_words first)");
	if (n_entries) *n_entries = s.n_entries;
	if (n_keys) {
		if (db_finalize_keys(ctx, s)) return 1;
		*n_keys = s.n_keys;
	}
	return 0;
}

int pcramp_gpu_keys_copy(pcramp_gpu_ctx *ctx, int kind, uint64_t *keys)
{
	if (check_kind(ctx, kind)) return 1;
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[kind];
	if (!s.db_valid) return fail(ctx, "pcramp_gpu_keys_copy: no database (call pcramp_gpu_select_words first)");
	if (db_finalize_keys(ctx, s)) return 1;
	const uint64_t n = s.n_entries;
	if (n == 0 || !keys) return 0;
	DevBuf o_words, o_index, o_loc, o_strand, o_key, o_keys;
	CK(o_words.ensure(n * 16));
	CK(o_index.ensure(n * 4));
	CK(o_loc.ensure(n * 4));
	CK(o_strand.ensure(n * 4));
	CK(o_key.ensure(n * 4));
	CK(o_keys.ensure(std::max<uint64_t>(1, s.n_keys) * 16));
	db_export_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(s.e_hi.as<uint64_t>(), s.e_lo.as<uint64_t>(), s.e_seq.as<uint32_t>(),
		s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(), s.e_perm.as<uint32_t>(), s.e_keyrank.as<uint32_t>(), n, o_words.as<uint64_t>(),
		o_index.as<uint32_t>(), o_loc.as<int32_t>(), o_strand.as<uint32_t>(), o_key.as<uint32_t>(), o_keys.as<uint64_t>());
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(keys, o_keys.p, s.n_keys * 16, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

uint32_t pcramp_gpu_bitset_words(pcramp_gpu_ctx *ctx, int kind) { return (ctx->sets[kind].n + 31u) / 32u; }

// frame-aligned seed table (fst.cuh) over n oligos given as letter planes + thresholds (device arrays)
static int fst_build(pcramp_gpu_ctx *ctx, const uint4 *d_planes, const uint32_t *d_thr, uint32_t n, Fst &t, uint32_t &n_brute_out)
{
	cudaStream_t st = ctx->stream;
	CK(ctx->d_fst_cnt.ensure((size_t)(FST_BUCKETS + 1) * 4));
	CK(ctx->d_fst_start.ensure((size_t)(FST_BUCKETS + 1) * 4));
	CK(ctx->d_fst_cursor.ensure((size_t)(FST_BUCKETS + 1) * 4));
	CK(ctx->d_fst_combo.ensure((2 * FST_COMBOS + 1) * 4));
	CK(ctx->d_fst_brute.ensure(std::max<size_t>(1, n) * 4));
	CK(ctx->d_fst_nbrute.ensure(16));
	CK(cudaMemsetAsync(ctx->d_fst_cnt.p, 0, (size_t)(FST_BUCKETS + 1) * 4, st));
	CK(cudaMemsetAsync(ctx->d_fst_combo.p, 0, FST_COMBOS * 4, st));
	CK(cudaMemsetAsync(ctx->d_fst_nbrute.p, 0, 16, st));
	fst_build_kernel<<<grid_for(n, 128), 128, 0, st>>>(d_planes, d_thr, n, ctx->d_fst_cnt.as<uint32_t>(), ctx->d_fst_combo.as<uint32_t>(),
		ctx->d_fst_brute.as<uint32_t>(), ctx->d_fst_nbrute.as<uint32_t>(), nullptr, 0u, nullptr);
	CK(cudaGetLastError());
	fst_combo_list_kernel<<<1, 32, 0, st>>>(ctx->d_fst_combo.as<uint32_t>());
	CK(cudaGetLastError());
	size_t tb = 0;
	CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, ctx->d_fst_cnt.as<uint32_t>(), ctx->d_fst_start.as<uint32_t>(), (int)(FST_BUCKETS + 1), st));
	CK(ctx->cub_tmp.ensure(tb));
	CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tb, ctx->d_fst_cnt.as<uint32_t>(), ctx->d_fst_start.as<uint32_t>(), (int)(FST_BUCKETS + 1), st));
	uint32_t total = 0, nb = 0;
	CK(cudaMemcpyAsync(&total, ctx->d_fst_start.as<uint32_t>() + FST_BUCKETS, 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(&nb, ctx->d_fst_nbrute.p, 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(ctx->d_fst_cursor.p, ctx->d_fst_start.p, (size_t)(FST_BUCKETS + 1) * 4, cudaMemcpyDeviceToDevice, st));
	CK(cudaStreamSynchronize(st));
	CK(ctx->d_fst_ids.ensure(std::max<size_t>(1, total) * 4));
	fst_build_kernel<<<grid_for(n, 128), 128, 0, st>>>(d_planes, d_thr, n, ctx->d_fst_cursor.as<uint32_t>(), ctx->d_fst_combo.as<uint32_t>(),
		ctx->d_fst_brute.as<uint32_t>(), ctx->d_fst_nbrute.as<uint32_t>(), ctx->d_fst_ids.as<uint32_t>(), 0xFFFFFFFFu, nullptr);
	CK(cudaGetLastError());
	ctx->stats.kernel_launches += 5;
	t.planes = d_planes;
	t.thr = d_thr;
	t.start = ctx->d_fst_start.as<uint32_t>();
	t.ids = ctx->d_fst_ids.as<uint32_t>();
	t.combo = ctx->d_fst_combo.as<uint32_t>();
	t.brute = ctx->d_fst_brute.as<uint32_t>();
	t.n_brute = ctx->d_fst_nbrute.as<uint32_t>();
	t.n = n;
	n_brute_out = nb;
	return 0;
}

// K2 launch sequence shared by pair scoring and move-variant scoring.  d_f / d_r: the oligos whose identities are
// computed; d_bf / d_br (variant scoring only, else null): the base assays whose candidate amplicon lists are used.
// pair scoring by (sequence, pair) units for thresholds at which the filters exclude nobody (sw_abi.cuh)
static int score_by_units(pcramp_gpu_ctx *ctx, SeqSet &s, const OligoDev *d_member, const OligoDev *d_oligos, uint32_t n_pairs, float detect_threshold,
	int amp_min, int amp_max, int taq, uint32_t *d_bits_any, uint32_t *d_bits_pass1, uint32_t n_words);

static int score_launch(pcramp_gpu_ctx *ctx, int kind, const uint64_t *d_f, const uint64_t *d_r, const uint64_t *d_bf, const uint64_t *d_br,
	uint32_t n_pairs, float search_threshold, float detect_threshold, int amp_min, int amp_max, int taq)
{
	SeqSet &s = ctx->sets[kind];
	cudaStream_t st = ctx->stream;
	if (!s.db_valid) return fail(ctx, "pcramp_gpu_score_pairs: no database (call pcramp_gpu_select_words first)");
	const uint32_t n_words = (s.n + 31u) / 32u;
	const bool variant = d_bf != nullptr;
	ctx->res_words = n_words;
	const size_t bits_bytes = std::max<size_t>(1, (size_t)n_pairs * n_words) * 4;
	CK(ctx->d_cov.ensure(std::max<size_t>(1, n_pairs) * 4));
	CK(ctx->d_bits.ensure(bits_bytes));
	CK(ctx->d_bits1.ensure(bits_bytes));
	CK(ctx->d_oligos.ensure(std::max<size_t>(1, n_pairs) * 2 * sizeof(OligoDev)));
	if (variant) CK(ctx->d_oligos_base.ensure(std::max<size_t>(1, n_pairs) * 2 * sizeof(OligoDev)));
	CK(cudaEventRecord(ctx->ev[5], st));
	CK(cudaMemsetAsync(ctx->d_bits.p, 0, bits_bytes, st));
	CK(cudaMemsetAsync(ctx->d_bits1.p, 0, bits_bytes, st));
	CK(cudaMemsetAsync(ctx->d_cov.p, 0, std::max<size_t>(1, n_pairs) * 4, st));
	ctx->stats.ms_score = 0.0f;
	ctx->pend_ms_score = false;
	if (n_pairs && s.n) {
		const float thr2 = search_threshold * search_threshold; // pcr_assay.cpp:31-32 (float product)
		prep_oligos_kernel<<<grid_for(2ull * n_pairs, 256), 256, 0, st>>>(d_f, d_r, n_pairs, thr2, ctx->d_oligos.as<OligoDev>());
		CK(cudaGetLastError());
		ctx->stats.kernel_launches++;
		if (variant) {
			prep_oligos_kernel<<<grid_for(2ull * n_pairs, 256), 256, 0, st>>>(d_bf, d_br, n_pairs, thr2, ctx->d_oligos_base.as<OligoDev>());
			CK(cudaGetLastError());
			ctx->stats.kernel_launches++;
		}
		if (s.n_entries) {
			// key-matrix filter (score.cuh), in chunks of pairs so that the (key x oligo) bit matrix stays below ~1 GB
			const OligoDev *d_member = variant ? ctx->d_oligos_base.as<OligoDev>() : ctx->d_oligos.as<OligoDev>();
			// chunks of pairs (whole 32-oligo words) such that `rows` bit rows of the filter matrix stay below ~1 GB
			auto chunk_for = [&](uint64_t rows) -> uint32_t {
				const uint64_t fit = std::max<uint64_t>(16, ((1ull << 30) / (std::max<uint64_t>(1, rows) * 4)) * 16);
				return fit >= n_pairs ? n_pairs : (uint32_t)(fit & ~15ull);
			};
			uint32_t chunk_pairs = chunk_for(s.n_entries); // key matrix: at most one key per entry
			CK(ctx->d_item_count.ensure(16));
			uint64_t item_cap = std::max<uint64_t>(ctx->d_items.cap / sizeof(ScoreItem), ctx->tiny_buffers ? 256ull : 1ull << 20);
			// seed-table filter (fst.cuh) unless too many oligos cannot be seeded (low thresholds: backgrounds at 0.72^2)
			Fst fst;
			// neighbour filter (score.cuh) when the candidates x oligos comparison is small next to the table walk it replaces
			// ... and when the neighbour bound can exclude anybody: candidate K (found at the seed threshold t, e1 = (1 - t) n misses) and oligo O
			// (matched at t^2, e2 = (1 - t^2) n misses) are neighbours when they differ in at most e1 + e2 of their ~n common slots, and two
			// unrelated words differ in 0.75 n -+ a few: below t^2 + t = 1.5 (t < 0.823: the background thresholds, 0.72) nearly every oligo
			// is every candidate's neighbour, the slots overflow and every entry would be compared with every oligo
			const bool neigh_selective = search_threshold * search_threshold + search_threshold >= 1.5f;
			const bool use_neigh = ctx->use_neigh != 0 && neigh_selective && s.n_cand > 0 && (uint64_t)s.n_cand * 2ull * n_pairs <= (1ull << 28);
			bool units_done = false;
			if (!neigh_selective) { // nothing to filter with: membership lists per (sequence, pair) and one thread per unit
				const int rc = score_by_units(ctx, s, d_member, ctx->d_oligos.as<OligoDev>(), n_pairs, detect_threshold, amp_min, amp_max, taq,
					ctx->d_bits.as<uint32_t>(), ctx->d_bits1.as<uint32_t>(), n_words);
				if (rc == 1) return 1;
				units_done = rc == 0;
			}
			bool use_fst = ctx->use_fst != 0 && !use_neigh && !units_done;
			if (use_fst) {
				CK(ctx->d_fst_planes.ensure((size_t)n_pairs * 2 * 16));
				CK(ctx->d_fst_thr.ensure((size_t)n_pairs * 2 * 4));
				oligo_split_kernel<<<grid_for(2ull * n_pairs, 256), 256, 0, st>>>(d_member, 2u * n_pairs, ctx->d_fst_planes.as<uint4>(),
					ctx->d_fst_thr.as<uint32_t>());
				CK(cudaGetLastError());
				ctx->stats.kernel_launches++;
			}
			if (use_neigh && ctx->use_entry_score) {
				// entry-driven scoring (score.cuh): neighbour slots per candidate, then one thread per plus-strand entry -- three launches,
				// no list whose size the host would have to learn
				const uint32_t n_olig = 2u * n_pairs;
				CK(ctx->d_neigh_off.ensure(((size_t)s.n_cand + 1) * 4));
				CK(ctx->d_neigh.ensure((size_t)s.n_cand * NEIGH_SLOTS * 4));
				CK(cudaMemsetAsync(ctx->d_neigh_off.p, 0, ((size_t)s.n_cand + 1) * 4, st));
				neigh_slots_kernel<<<dim3(grid_for(s.n_cand, 256), grid_for(n_olig, 256)), 256, 0, st>>>(s.c_planes.as<uint4>(), s.c_thr.as<uint32_t>(),
					s.n_cand, d_member, n_olig, ctx->d_neigh.as<uint32_t>(), ctx->d_neigh_off.as<uint32_t>());
				CK(cudaGetLastError());
				if (variant)
					score_entries_kernel<true><<<grid_for(s.n_entries, 128), 128, 0, st>>>(s.dev(), s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(),
						s.e_strand.as<uint32_t>(), s.e_seq.as<uint32_t>(), s.e_cand.as<uint32_t>(), s.n_entries, s.seq_ent_off.as<uint32_t>(),
						s.seq_full_end.as<uint32_t>(), ctx->d_neigh.as<uint32_t>(), ctx->d_neigh_off.as<uint32_t>(), d_member, ctx->d_oligos.as<OligoDev>(),
						n_olig, detect_threshold, amp_min, amp_max, taq, ctx->d_bits.as<uint32_t>(), ctx->d_bits1.as<uint32_t>(), n_words);
				else
					score_entries_kernel<false><<<grid_for(s.n_entries, 128), 128, 0, st>>>(s.dev(), s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(),
						s.e_strand.as<uint32_t>(), s.e_seq.as<uint32_t>(), s.e_cand.as<uint32_t>(), s.n_entries, s.seq_ent_off.as<uint32_t>(),
						s.seq_full_end.as<uint32_t>(), ctx->d_neigh.as<uint32_t>(), ctx->d_neigh_off.as<uint32_t>(), d_member, ctx->d_oligos.as<OligoDev>(),
						n_olig, detect_threshold, amp_min, amp_max, taq, ctx->d_bits.as<uint32_t>(), ctx->d_bits1.as<uint32_t>(), n_words);
				CK(cudaGetLastError());
				ctx->stats.kernel_launches += 2;
				chunk_pairs = n_pairs;
			}
			if (use_fst || use_neigh) chunk_pairs = std::min(chunk_pairs, chunk_for(2ull * s.n)); // two bit rows per sequence
			for (uint32_t p0 = ((use_neigh && ctx->use_entry_score) || units_done) ? n_pairs : 0u; p0 < n_pairs; p0 += chunk_pairs) {
				const uint32_t pc = std::min<uint32_t>(chunk_pairs, n_pairs - p0), nw = (2u * pc + 31u) / 32u;
				bool chunk_fst = use_fst;
				if (use_neigh) {
					const uint32_t n_olig = 2u * pc;
					const OligoDev *d_ol = d_member + 2ull * p0;
					CK(ctx->d_neigh_off.ensure(((size_t)s.n_cand + 1) * 4));
					unsigned int n_np = 0;
					for (int attempt = 0;; ++attempt) { // (candidate, oligo) pairs, then sorted by candidate
						CK(ctx->d_neigh.ensure(std::max<size_t>(ctx->d_neigh.cap, ctx->tiny_buffers ? (size_t)16 * 64 : (size_t)16 * (8ull * n_olig + 4096))));
						const uint32_t cap = (uint32_t)std::min<size_t>(ctx->d_neigh.cap / 16, 0x7FFFFFF0u); // second half: sort buffer
						CK(cudaMemsetAsync(ctx->d_item_count.p, 0, 16, st));
						neigh_pairs_kernel<<<dim3(grid_for(s.n_cand, 256), grid_for(n_olig, 256)), 256, 0, st>>>(s.c_planes.as<uint4>(), s.c_thr.as<uint32_t>(),
							s.n_cand, d_ol, n_olig, ctx->d_neigh.as<unsigned long long>(), ctx->d_item_count.as<unsigned int>(), cap);
						CK(cudaGetLastError());
						ctx->stats.kernel_launches++;
						CK(cudaMemcpyAsync(&n_np, ctx->d_item_count.p, 4, cudaMemcpyDeviceToHost, st));
						CK(cudaStreamSynchronize(st));
						if (n_np <= cap) break;
						if (attempt >= 2) return fail(ctx, "pcramp_gpu_score_pairs: neighbour list kept overflowing");
						CK(ctx->d_neigh.ensure((size_t)16 * ((size_t)n_np + n_np / 8 + 4096)));
					}
					unsigned long long *d_np = ctx->d_neigh.as<unsigned long long>();
					if (n_np > 1) {
						cub::DoubleBuffer<unsigned long long> dbuf(d_np, d_np + ctx->d_neigh.cap / 16);
						size_t tb = 0;
						const int end_bit = 32 + (int)bits_for((uint64_t)s.n_cand + 1);
						// grouping by candidate is all neigh_offsets_kernel / entry_neigh_kernel need (the oligos of a candidate in any order):
						// only the candidate bits are sorted -- two radix passes instead of six
						CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, dbuf, (int)n_np, 32, end_bit, st));
						CK(ctx->cub_tmp.ensure(tb));
						CK(cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tb, dbuf, (int)n_np, 32, end_bit, st));
						d_np = dbuf.Current();
						ctx->stats.kernel_launches += 4;
					}
					neigh_offsets_kernel<<<grid_for((uint64_t)s.n_cand + 1, 256), 256, 0, st>>>(d_np, n_np, s.n_cand, ctx->d_neigh_off.as<uint32_t>());
					const size_t row_bytes = (size_t)2 * s.n * nw * 4;
					CK(ctx->d_seqbits.ensure(std::max<size_t>(4, row_bytes)));
					CK(cudaMemsetAsync(ctx->d_seqbits.p, 0, row_bytes, st));
					entry_neigh_kernel<<<grid_for(s.n_entries, 128), 128, 0, st>>>(d_np, ctx->d_neigh_off.as<uint32_t>(), d_ol, n_olig, s.e_planes.as<uint4>(),
						s.e_seq.as<uint32_t>(), s.e_strand.as<uint32_t>(), s.e_cand.as<uint32_t>(), s.n_entries, nw, ctx->d_seqbits.as<uint32_t>());
					CK(cudaGetLastError());
					ctx->stats.kernel_launches += 2;
				}
				if (chunk_fst) {
					uint32_t n_brute = 0;
					if (fst_build(ctx, ctx->d_fst_planes.as<uint4>() + 2ull * p0, ctx->d_fst_thr.as<uint32_t>() + 2ull * p0, 2u * pc, fst, n_brute)) return 1;
					if ((uint64_t)n_brute * 4u > 2ull * pc) chunk_fst = false; // mostly unseedable: the key matrix is the better brute force
				}
				if (use_neigh) {
				} else if (chunk_fst) {
					const size_t row_bytes = (size_t)2 * s.n * nw * 4;
					CK(ctx->d_seqbits.ensure(std::max<size_t>(4, row_bytes)));
					CK(cudaMemsetAsync(ctx->d_seqbits.p, 0, row_bytes, st));
					entry_match_kernel<<<grid_for(s.n_entries, 128), 128, 0, st>>>(fst, s.e_planes.as<uint4>(), s.e_seq.as<uint32_t>(),
						s.e_strand.as<uint32_t>(), s.n_entries, nw, ctx->d_seqbits.as<uint32_t>());
					CK(cudaGetLastError());
					ctx->stats.kernel_launches++;
				} else {
					if (db_finalize_keys(ctx, s)) return 1;
					const uint64_t n_keys = std::max<uint64_t>(1, s.n_keys);
					CK(ctx->d_keybits.ensure((size_t)n_keys * nw * 4));
					key_match_kernel<<<dim3(grid_for(s.n_keys, KEYM_THREADS), nw), KEYM_THREADS, 0, st>>>(s.key_planes.as<uint4>(), (uint32_t)s.n_keys,
						d_member + 2ull * p0, 2u * pc, nw, ctx->d_keybits.as<uint32_t>());
					CK(cudaGetLastError());
					ctx->stats.kernel_launches++;
				}
				if ((chunk_fst || use_neigh) && ctx->use_fused_score) {
					// bit rows -> exact amplicon test, one CTA per sequence with its entries in shared memory (no item list, no round trip)
					const unsigned grid = (unsigned)std::min<uint64_t>(s.n, (uint64_t)ctx->sm_count * 8u);
					if (variant)
						score_seqbits_kernel<true><<<grid, SCORE_THREADS, 0, st>>>(s.dev(), s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(),
							s.seq_ent_off.as<uint32_t>(), ctx->d_oligos.as<OligoDev>() + 2ull * p0, ctx->d_oligos_base.as<OligoDev>() + 2ull * p0,
							ctx->d_seqbits.as<uint32_t>(), nw, pc, detect_threshold, amp_min, amp_max, taq,
							ctx->d_bits.as<uint32_t>() + (size_t)p0 * n_words, ctx->d_bits1.as<uint32_t>() + (size_t)p0 * n_words, n_words);
					else
						score_seqbits_kernel<false><<<grid, SCORE_THREADS, 0, st>>>(s.dev(), s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(),
							s.seq_ent_off.as<uint32_t>(), ctx->d_oligos.as<OligoDev>() + 2ull * p0, nullptr, ctx->d_seqbits.as<uint32_t>(), nw, pc,
							detect_threshold, amp_min, amp_max, taq,
							ctx->d_bits.as<uint32_t>() + (size_t)p0 * n_words, ctx->d_bits1.as<uint32_t>() + (size_t)p0 * n_words, n_words);
					CK(cudaGetLastError());
					ctx->stats.kernel_launches++;
					continue;
				}
				for (int attempt = 0;; ++attempt) {
					CK(ctx->d_items.ensure(item_cap * sizeof(ScoreItem)));
					CK(cudaMemsetAsync(ctx->d_item_count.p, 0, 16, st));
					if (chunk_fst || use_neigh) {
						seq_pairs_kernel<<<grid_for((uint64_t)s.n * nw, 256), 256, 0, st>>>(s.dev(), s.seq_ent_off.as<uint32_t>(),
							ctx->d_seqbits.as<uint32_t>(), nw, pc, ctx->d_items.as<ScoreItem>(), ctx->d_item_count.as<unsigned int>(),
							(uint32_t)std::min<uint64_t>(item_cap, 0xFFFFFFF0ull));
					} else {
						const unsigned fthreads = std::min<uint32_t>(256u, (nw + 31u) & ~31u);
						seq_filter_kernel<<<(unsigned)std::min<uint64_t>(s.n, (uint64_t)ctx->sm_count * 16), fthreads, 0, st>>>(s.dev(),
							s.seq_ent_off.as<uint32_t>(), s.e_key.as<uint32_t>(), ctx->d_keybits.as<uint32_t>(), nw, pc, ctx->d_items.as<ScoreItem>(),
							ctx->d_item_count.as<unsigned int>(), (uint32_t)std::min<uint64_t>(item_cap, 0xFFFFFFF0ull));
					}
					CK(cudaGetLastError());
					ctx->stats.kernel_launches++;
					unsigned int n_items = 0;
					CK(cudaMemcpyAsync(&n_items, ctx->d_item_count.p, 4, cudaMemcpyDeviceToHost, st));
					CK(cudaStreamSynchronize(st));
					if (n_items <= item_cap) break;
					if (attempt >= 2) return fail(ctx, "pcramp_gpu_score_pairs: work list kept overflowing");
					item_cap = (uint64_t)n_items + n_items / 8 + 1024;
				}
				const unsigned grid = (unsigned)ctx->sm_count * 8u;
				const uint32_t icap = (uint32_t)std::min<uint64_t>(item_cap, 0xFFFFFFF0ull);
				if (variant)
					score_items_kernel<true><<<grid, 256, 0, st>>>(s.dev(), s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(),
						s.seq_ent_off.as<uint32_t>(), ctx->d_oligos.as<OligoDev>() + 2ull * p0, ctx->d_oligos_base.as<OligoDev>() + 2ull * p0,
						ctx->d_items.as<ScoreItem>(), ctx->d_item_count.as<unsigned int>(), icap, detect_threshold, amp_min, amp_max, taq,
						ctx->d_bits.as<uint32_t>() + (size_t)p0 * n_words, ctx->d_bits1.as<uint32_t>() + (size_t)p0 * n_words, n_words);
				else
					score_items_kernel<false><<<grid, 256, 0, st>>>(s.dev(), s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(),
						s.seq_ent_off.as<uint32_t>(), ctx->d_oligos.as<OligoDev>() + 2ull * p0, nullptr, ctx->d_items.as<ScoreItem>(),
						ctx->d_item_count.as<unsigned int>(), icap, detect_threshold, amp_min, amp_max, taq,
						ctx->d_bits.as<uint32_t>() + (size_t)p0 * n_words, ctx->d_bits1.as<uint32_t>() + (size_t)p0 * n_words, n_words);
				CK(cudaGetLastError());
				ctx->stats.kernel_launches++;
			}
			if (s.unit_weights)
				coverage_count_kernel<<<grid_for(32ull * n_pairs, 256), 256, 0, st>>>(ctx->d_bits.as<uint32_t>(), n_pairs, n_words, ctx->d_cov.as<float>());
			else
				coverage_kernel<<<grid_for(n_pairs, 128), 128, 0, st>>>(ctx->d_bits.as<uint32_t>(), ctx->d_bits1.as<uint32_t>(), s.d_weight.as<float>(),
					n_pairs, n_words, s.n, ctx->d_cov.as<float>());
			CK(cudaGetLastError());
			ctx->stats.kernel_launches += 1;
		}
	}
	CK(cudaEventRecord(ctx->ev[6], st));
	ctx->pend_ms_score = true; // the results stay on the stream; fetch / exchange / get_stats synchronise when they need to
	return 0;
}

int pcramp_gpu_score_pairs_staged(pcramp_gpu_ctx *ctx, int kind, float search_threshold, float detect_threshold, int amp_min, int amp_max,
	int taq)
{
	if (check_kind(ctx, kind)) return 1;
	CK(cudaSetDevice(ctx->device));
	return score_launch(ctx, kind, ctx->pf(), ctx->pr(), nullptr, nullptr, ctx->n_pairs, search_threshold, detect_threshold, amp_min, amp_max, taq);
}

// The scoring step of an optimisation move for a batch of trial oligos (optimize_pcr.cpp:8-989: every move is
// "mutate one oligo -> is_valid -> update_identity -> compute_coverage" against the candidate amplicons that
// collect_candidates built for the UNMOVED assay): variant i is scored against the candidate list of base assay i.
// score_variants by groups of variants that share their base assay (score_entries_groups_kernel): d_words = {variant F, variant R,
// group base F, group base R}; -> 0 done, 1 error, 2 not applicable (the caller takes the general path)
static int score_variants_grouped(pcramp_gpu_ctx *ctx, int kind, const uint64_t *d_vf, const uint64_t *d_vr, const uint64_t *d_gf, const uint64_t *d_gr,
	const uint32_t *d_goff, uint32_t n, uint32_t G, float search_threshold, float detect_threshold, int amp_min, int amp_max, int taq)
{
	SeqSet &s = ctx->sets[kind];
	cudaStream_t st = ctx->stream;
	if (!s.db_valid) return fail(ctx, "pcramp_gpu_score_pairs: no database (call pcramp_gpu_select_words first)");
	if (!(ctx->use_neigh && ctx->use_entry_score && s.n_cand > 0 && s.n_entries > 0 && s.n && n && (uint64_t)s.n_cand * 2ull * G <= (1ull << 28))) return 2;
	if (search_threshold * search_threshold + search_threshold < 1.5f) return 2; // the neighbour bound excludes nobody (score_launch)
	const uint32_t n_words = (s.n + 31u) / 32u;
	ctx->res_words = n_words;
	const size_t bits_bytes = (size_t)n * n_words * 4;
	CK(ctx->d_cov.ensure((size_t)n * 4));
	CK(ctx->d_bits.ensure(bits_bytes));
	CK(ctx->d_bits1.ensure(bits_bytes));
	CK(ctx->d_oligos.ensure((size_t)n * 2 * sizeof(OligoDev)));
	CK(ctx->d_oligos_base.ensure((size_t)G * 2 * sizeof(OligoDev)));
	CK(ctx->d_neigh_off.ensure(((size_t)s.n_cand + 1) * 4));
	CK(ctx->d_neigh.ensure((size_t)s.n_cand * NEIGH_SLOTS * 4));
	CK(cudaEventRecord(ctx->ev[5], st));
	CK(cudaMemsetAsync(ctx->d_bits.p, 0, bits_bytes, st));
	CK(cudaMemsetAsync(ctx->d_bits1.p, 0, bits_bytes, st));
	CK(cudaMemsetAsync(ctx->d_neigh_off.p, 0, ((size_t)s.n_cand + 1) * 4, st));
	ctx->stats.ms_score = 0.0f;
	ctx->pend_ms_score = false;
	const float thr2 = search_threshold * search_threshold; // pcr_assay.cpp:31-32 (float product)
	prep_oligos_kernel<<<grid_for(2ull * n, 256), 256, 0, st>>>(d_vf, d_vr, n, thr2, ctx->d_oligos.as<OligoDev>());
	prep_oligos_kernel<<<grid_for(2ull * G, 256), 256, 0, st>>>(d_gf, d_gr, G, thr2, ctx->d_oligos_base.as<OligoDev>());
	neigh_slots_kernel<<<dim3(grid_for(s.n_cand, 256), grid_for(2u * G, 256)), 256, 0, st>>>(s.c_planes.as<uint4>(), s.c_thr.as<uint32_t>(), s.n_cand,
		ctx->d_oligos_base.as<OligoDev>(), 2u * G, ctx->d_neigh.as<uint32_t>(), ctx->d_neigh_off.as<uint32_t>());
	score_entries_groups_kernel<<<grid_for(s.n_entries, 128), 128, 0, st>>>(s.dev(), s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(),
		s.e_seq.as<uint32_t>(), s.e_cand.as<uint32_t>(), s.n_entries, s.seq_ent_off.as<uint32_t>(), s.seq_full_end.as<uint32_t>(),
		ctx->d_neigh.as<uint32_t>(), ctx->d_neigh_off.as<uint32_t>(), ctx->d_oligos_base.as<OligoDev>(), 2u * G, d_goff, ctx->d_oligos.as<OligoDev>(),
		detect_threshold, amp_min, amp_max, taq, ctx->d_bits.as<uint32_t>(), ctx->d_bits1.as<uint32_t>(), n_words);
	if (s.unit_weights)
		coverage_count_kernel<<<grid_for(32ull * n, 256), 256, 0, st>>>(ctx->d_bits.as<uint32_t>(), n, n_words, ctx->d_cov.as<float>());
	else
		coverage_kernel<<<grid_for(n, 128), 128, 0, st>>>(ctx->d_bits.as<uint32_t>(), ctx->d_bits1.as<uint32_t>(), s.d_weight.as<float>(), n, n_words, s.n,
			ctx->d_cov.as<float>());
	CK(cudaGetLastError());
	ctx->stats.kernel_launches += 5;
	CK(cudaEventRecord(ctx->ev[6], st));
	ctx->pend_ms_score = true;
	return 0;
}

int pcramp_gpu_score_variants(pcramp_gpu_ctx *ctx, int kind, const uint64_t *base_f, const uint64_t *base_r, const uint64_t *var_f,
	const uint64_t *var_r, uint32_t n, float search_threshold, float detect_threshold, int amp_min, int amp_max, int taq, float *coverage,
	uint32_t *bitsets)
{
	if (check_kind(ctx, kind)) return 1;
	if (n && (!base_f || !base_r || !var_f || !var_r)) return fail(ctx, "pcramp_gpu_score_variants: null argument");
	CK(cudaSetDevice(ctx->device));
	DevBuf &d = ctx->d_variants;
	CK(d.ensure(std::max<size_t>(1, n) * 64));
	uint64_t *p = d.as<uint64_t>();
	if (n) {
		CK(cudaMemcpyAsync(p, base_f, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(p + 2ull * n, base_r, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(p + 4ull * n, var_f, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(p + 6ull * n, var_r, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
	}
	ctx->stats.kernel_launches = 0;
	// variants that follow each other with the same base assay (the moves of one trial, optimize_abi.cuh) form a group
	int grouped = 2;
	if (n && ctx->use_variant_groups) {
		std::vector<uint64_t> &gw = ctx->h_group_words;
		std::vector<uint32_t> &go = ctx->h_group_off;
		gw.clear();
		go.clear();
		std::vector<uint64_t> gr;
		for (uint32_t i = 0; i < n; ++i) {
			if (i == 0 || base_f[2 * i] != base_f[2 * i - 2] || base_f[2 * i + 1] != base_f[2 * i - 1] || base_r[2 * i] != base_r[2 * i - 2] ||
			    base_r[2 * i + 1] != base_r[2 * i - 1]) {
				go.push_back(i);
				gw.push_back(base_f[2 * i]); gw.push_back(base_f[2 * i + 1]);
				gr.push_back(base_r[2 * i]); gr.push_back(base_r[2 * i + 1]);
			}
		}
		const uint32_t G = (uint32_t)go.size();
		go.push_back(n);
		gw.insert(gw.end(), gr.begin(), gr.end());
		DevBuf &dg = ctx->d_variant_groups;
		CK(dg.ensure((size_t)G * 32 + ((size_t)G + 1) * 4));
		CK(cudaMemcpyAsync(dg.p, gw.data(), (size_t)G * 32, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync((char *)dg.p + (size_t)G * 32, go.data(), ((size_t)G + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
		grouped = score_variants_grouped(ctx, kind, p + 4ull * n, p + 6ull * n, dg.as<uint64_t>(), dg.as<uint64_t>() + 2ull * G,
			(const uint32_t *)((char *)dg.p + (size_t)G * 32), n, G, search_threshold, detect_threshold, amp_min, amp_max, taq);
		if (grouped == 1) return 1;
		if (grouped == 0) CK(cudaStreamSynchronize(ctx->stream)); // the group arrays are host vectors of the context: done with them
	}
	if (grouped == 2 &&
	    score_launch(ctx, kind, p + 4ull * n, p + 6ull * n, p, p + 2ull * n, n, search_threshold, detect_threshold, amp_min, amp_max, taq)) return 1;
	const uint32_t save = ctx->n_pairs;
	ctx->n_pairs = n;
	const int rc = pcramp_gpu_fetch_results(ctx, coverage, bitsets);
	ctx->n_pairs = save;
	return rc;
}

void *pcramp_gpu_device_coverage(pcramp_gpu_ctx *ctx) { return ctx->d_cov.p; }
void *pcramp_gpu_device_bitsets(pcramp_gpu_ctx *ctx) { return ctx->d_bits.p; }
void *pcramp_gpu_device_bitsets_pass1(pcramp_gpu_ctx *ctx) { return ctx->d_bits1.p; }

// ---- multi-GPU: combine per-shard results -----------------------------------
// Each rank scored the same pairs against its own contiguous shard of the sequences.  After the
// all-gather of the shards' bitsets (rank-major: shard s holds n_pairs x words_s words), splice them
// into one global LSB-first bitset per pair and recompute the coverage over ALL sequences in the
// reference's order (pass-1 detections ascending, then pass-2-only ascending) from the global weights.
__global__ void merge_shards_kernel(const uint32_t *__restrict__ any, const uint32_t *__restrict__ p1, uint32_t n_shards,
	const uint32_t *__restrict__ shard_nseq, const uint64_t *__restrict__ shard_word_off, const uint32_t *__restrict__ shard_seq_off,
	const float *__restrict__ weight, uint32_t n_pairs, uint32_t out_words, uint32_t *out_bits, float *out_cov)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n_pairs) return;
	uint32_t *ob = out_bits + (size_t)p * out_words;
	for (uint32_t w = 0; w < out_words; ++w) ob[w] = 0u;
	double acc = 0.0;
	for (int pass = 0; pass < 2; ++pass) {
		for (uint32_t s = 0; s < n_shards; ++s) {
			const uint32_t ns = shard_nseq[s], ws = (ns + 31u) / 32u;
			const uint32_t *a = any + shard_word_off[s] + (size_t)p * ws, *b = p1 + shard_word_off[s] + (size_t)p * ws;
			for (uint32_t w = 0; w < ws; ++w) {
				uint32_t m = pass == 0 ? b[w] : (a[w] & ~b[w]);
				while (m) {
					const uint32_t local = w * 32u + (uint32_t)(__ffs(m) - 1);
					m &= m - 1u;
					if (local >= ns) continue;
					const uint32_t g = shard_seq_off[s] + local;
					ob[g >> 5] |= 1u << (g & 31u);
					acc = __dadd_rn(acc, (double)weight[g]);
				}
			}
		}
	}
	out_cov[p] = (float)acc;
}

int pcramp_gpu_merge_shards(pcramp_gpu_ctx *ctx, const void *d_any_gathered, const void *d_pass1_gathered, uint32_t n_shards,
	const uint32_t *shard_nseq, const float *weight_all, uint32_t n_pairs, void *d_out_bits, void *d_out_cov)
{
	if (!ctx) return 1;
	CK(cudaSetDevice(ctx->device));
	std::vector<uint64_t> woff(n_shards);
	std::vector<uint32_t> soff(n_shards);
	uint64_t wo = 0;
	uint32_t so = 0;
	for (uint32_t s = 0; s < n_shards; ++s) {
		woff[s] = wo;
		soff[s] = so;
		wo += (uint64_t)n_pairs * ((shard_nseq[s] + 31u) / 32u);
		so += shard_nseq[s];
	}
	DevBuf d_ns, d_wo, d_so, d_w;
	CK(d_ns.ensure(n_shards * 4));
	CK(d_wo.ensure(n_shards * 8));
	CK(d_so.ensure(n_shards * 4));
	CK(d_w.ensure(std::max<size_t>(1, so) * 4));
	CK(cudaMemcpyAsync(d_ns.p, shard_nseq, n_shards * 4, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(d_wo.p, woff.data(), n_shards * 8, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(d_so.p, soff.data(), n_shards * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (weight_all) CK(cudaMemcpyAsync(d_w.p, weight_all, (size_t)so * 4, cudaMemcpyHostToDevice, ctx->stream));
	else {
		std::vector<float> ones(so, 1.0f);
		CK(cudaMemcpyAsync(d_w.p, ones.data(), (size_t)so * 4, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	merge_shards_kernel<<<grid_for(n_pairs, 128), 128, 0, ctx->stream>>>((const uint32_t *)d_any_gathered, (const uint32_t *)d_pass1_gathered,
		n_shards, d_ns.as<uint32_t>(), d_wo.as<uint64_t>(), d_so.as<uint32_t>(), d_w.as<float>(), n_pairs, (so + 31u) / 32u,
		(uint32_t *)d_out_bits, (float *)d_out_cov);
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(ctx->stream));
	ctx->stats.kernel_launches++;
	return 0;
}

// ---- measured integer-pipe peak: the denominator of the scan kernel's issue roofline ----------------
// Each thread runs 8 independent chains of the scan's own instruction mix (4 LOP3 + POPC + compare per
// alignment) on register-resident data, so the figure is "alignments/s if nothing but issue limits it".
__global__ void __launch_bounds__(256, 2) int_peak_kernel(uint32_t iters, uint32_t seed, uint32_t *sink)
{
	uint32_t wa[8], wc[8], wg[8], wt[8];
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	#pragma unroll
	for (int r = 0; r < 8; ++r) {
		wa[r] = (t * 2654435761u) ^ (seed + r);
		wc[r] = ~wa[r] & (t * 40503u + r);
		wg[r] = (wa[r] >> 3) ^ 0x5a5a5a5au;
		wt[r] = ~(wa[r] | wc[r] | wg[r]);
	}
	uint4 b = make_uint4(seed * 3u + 1u, seed * 5u + 2u, seed * 7u + 3u, seed * 11u + 4u);
	uint32_t hits = 0;
	for (uint32_t i = 0; i < iters; ++i) {
		bool any = false;
		#pragma unroll
		for (int r = 0; r < 8; ++r) {
			const uint32_t m = (b.x & wa[r]) | (b.y & wc[r]) | (b.z & wg[r]) | (b.w & wt[r]);
			any |= (__popc(m) >= 31);
		}
		if (any) ++hits;
		b.x = b.x * 1664525u + 1013904223u; // a new pattern every iteration (cheap, on the FMA pipe)
		b.y += b.x;
		b.z ^= b.y;
		b.w -= b.z;
	}
	if (hits == 0xffffffffu) sink[t] = hits;
}

// The dynamic-programming kernels' instruction mix (K3 fill, K4 cells): three-input maximum, add-maximum, logic op, add on
// eight independent register chains per thread: "INT32 operations/s if nothing but issue limits them" -- the denominator of the
// DP rooflines (SURVEY.md 8d costs a cell in INT32 operations).
__global__ void __launch_bounds__(256, 2) int32_peak_kernel(uint32_t iters, int seed, int *sink)
{
	int x[8], y[8];
	const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
	#pragma unroll
	for (int r = 0; r < 8; ++r) {
		x[r] = t * 7919 + seed + r;
		y[r] = (t ^ (r * 0x9e37)) - seed;
	}
	for (uint32_t i = 0; i < iters; ++i) {
		#pragma unroll
		for (int r = 0; r < 8; ++r) {
			const int a = __vimax3_s32(x[r], y[r], seed);
			const int b = __viaddmax_s32(a, -5, y[r]);
			const int c = (b & 0x7fff0fff) ^ x[r];
			x[r] = c + y[r];
			y[r] = a;
		}
	}
	int acc = 0;
	#pragma unroll
	for (int r = 0; r < 8; ++r) acc ^= x[r] ^ y[r];
	if (acc == 0x7fffffff) *sink = acc;
}

int pcramp_gpu_measure_int_peak(pcramp_gpu_ctx *ctx, double *alignments_per_s)
{
	if (!ctx || !alignments_per_s) return 1;
	CK(cudaSetDevice(ctx->device));
	DevBuf sink;
	CK(sink.ensure(4));
	const uint32_t iters = 200000;
	const unsigned grid = (unsigned)ctx->sm_count * 2;
	int_peak_kernel<<<grid, 256, 0, ctx->stream>>>(1000, 1u, sink.as<uint32_t>()); // warm-up
	CK(cudaEventRecord(ctx->ev[0], ctx->stream));
	int_peak_kernel<<<grid, 256, 0, ctx->stream>>>(iters, 2u, sink.as<uint32_t>());
	CK(cudaEventRecord(ctx->ev[1], ctx->stream));
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(ctx->stream));
	const double s = ev_ms(ctx->ev[0], ctx->ev[1]) * 1e-3;
	*alignments_per_s = (double)grid * 256.0 * 8.0 * (double)iters / s;
	return 0;
}

int pcramp_gpu_measure_int32_peak(pcramp_gpu_ctx *ctx, double *ops_per_s)
{
	if (!ctx || !ops_per_s) return 1;
	CK(cudaSetDevice(ctx->device));
	DevBuf sink;
	CK(sink.ensure(4));
	const uint32_t iters = 100000;
	const unsigned grid = (unsigned)ctx->sm_count * 4;
	int32_peak_kernel<<<grid, 256, 0, ctx->stream>>>(1000, 1, sink.as<int>()); // warm-up
	CK(cudaEventRecord(ctx->ev[0], ctx->stream));
	int32_peak_kernel<<<grid, 256, 0, ctx->stream>>>(iters, 2, sink.as<int>());
	CK(cudaEventRecord(ctx->ev[1], ctx->stream));
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(ctx->stream));
	const double s = ev_ms(ctx->ev[0], ctx->ev[1]) * 1e-3;
	*ops_per_s = (double)grid * 256.0 * 8.0 * 4.0 * (double)iters / s;
	return 0;
}

int pcramp_gpu_fetch_results(pcramp_gpu_ctx *ctx, float *coverage, uint32_t *bitsets)
{
	if (!ctx) return 1;
	CK(cudaSetDevice(ctx->device));
	if (coverage && ctx->n_pairs) CK(cudaMemcpyAsync(coverage, ctx->d_cov.p, (size_t)ctx->n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (bitsets && ctx->n_pairs && ctx->res_words)
		CK(cudaMemcpyAsync(bitsets, ctx->d_bits.p, (size_t)ctx->n_pairs * ctx->res_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return 0;
}

int pcramp_gpu_score_pairs(pcramp_gpu_ctx *ctx, int kind, const uint64_t *f, const uint64_t *r, uint32_t n_pairs, float search_threshold,
	float detect_threshold, int amp_min, int amp_max, int taq, float *coverage, uint32_t *bitsets)
{
	if (pcramp_gpu_stage_pairs(ctx, f, r, n_pairs)) return 1;
	if (pcramp_gpu_score_pairs_staged(ctx, kind, search_threshold, detect_threshold, amp_min, amp_max, taq)) return 1;
	return pcramp_gpu_fetch_results(ctx, coverage, bitsets);
}

int pcramp_gpu_set_option(pcramp_gpu_ctx *ctx, const char *name, int value)
{
	if (!ctx || !name) return 1;
	if (strcmp(name, "force_brute_scan") == 0) { ctx->force_brute = value; return 0; }
	if (strcmp(name, "use_index") == 0) { ctx->use_index = value; return 0; }
	if (strcmp(name, "use_seed_table") == 0) { ctx->use_fst = value; return 0; }
	if (strcmp(name, "use_neighbours") == 0) { ctx->use_neigh = value; return 0; }
	if (strcmp(name, "use_tier_table") == 0) { ctx->use_tier_table = value; return 0; }
	if (strcmp(name, "use_fused_score") == 0) { ctx->use_fused_score = value; return 0; }
	if (strcmp(name, "use_entry_score") == 0) { ctx->use_entry_score = value; return 0; }
	if (strcmp(name, "use_variant_groups") == 0) { ctx->use_variant_groups = value; return 0; }
	if (strcmp(name, "use_background_units") == 0) { ctx->use_background_units = value; return 0; }
	if (strcmp(name, "use_async_scan") == 0) { ctx->use_async_scan = value; return 0; }
	if (strcmp(name, "use_unit_score") == 0) { ctx->use_unit_score = value; return 0; }
	if (strcmp(name, "use_edge_table") == 0) { ctx->use_edge_table = value; return 0; }
	if (strcmp(name, "use_segmented_db") == 0) { ctx->use_seg_db = value; return 0; }
	if (strcmp(name, "use_fast_path") == 0) { ctx->use_fast = value; return 0; }
	if (strcmp(name, "tiny_buffers") == 0) { ctx->tiny_buffers = value; return 0; }
	if (strcmp(name, "index_part_positions") == 0) {
		ctx->idx_part_cap = value > 0 ? (uint64_t)value : (1ull << 31);
		for (auto &s : ctx->sets) s.idx_drop(); // rebuilt with the new part size on the next seeded scan
		return 0;
	}
	return fail(ctx, std::string("pcramp_gpu_set_option: unknown option ") + name);
}

int pcramp_gpu_get_stats(pcramp_gpu_ctx *ctx, pcramp_gpu_stats *out)
{
	if (!ctx || !out) return 1;
	if (fast_resolve(ctx)) return 1;
	if (ctx->pend_ms_db) {
		CK(cudaEventSynchronize(ctx->ev[4]));
		ctx->stats.ms_db = ev_ms(ctx->ev[3], ctx->ev[4]);
		ctx->pend_ms_db = false;
	}
	if (ctx->pend_ms_score) {
		CK(cudaEventSynchronize(ctx->ev[6]));
		ctx->stats.ms_score = ev_ms(ctx->ev[5], ctx->ev[6]);
		ctx->pend_ms_score = false;
	}
	ctx->stats.n_fast = ctx->n_fast;
	ctx->stats.n_fast_redo = ctx->n_fast_redo;
	*out = ctx->stats;
	return 0;
}

// ---- host-side word helpers -----------------------------------------------------------------
static inline uint32_t iupac_to_bits(char c)
{ // base_table.h:31-76
	switch (c) {
	case 'A': case 'a': return 1;
	case 'C': case 'c': return 2;
	case 'G': case 'g': return 4;
	case 'T': case 't': case 'U': case 'u': return 8;
	case 'M': case 'm': return 3;
	case 'R': case 'r': return 5;
	case 'S': case 's': return 6;
	case 'V': case 'v': return 7;
	case 'W': case 'w': return 9;
	case 'Y': case 'y': return 10;
	case 'H': case 'h': return 11;
	case 'K': case 'k': return 12;
	case 'D': case 'd': return 13;
	case 'B': case 'b': return 14;
	case 'N': case 'n': case 'I': case 'i': case 'X': case 'x': return 15;
	default: return 0;
	}
}

void pcramp_word_from_string(const char *iupac, int centre, uint64_t out[2])
{
	W128 w;
	w.hi = w.lo = 0;
	for (int i = 0; iupac[i] && i < WORD_LEN; ++i) w_set(w, i, iupac_to_bits(iupac[i]));
	if (centre) w = w_center(w);
	out[0] = w.hi;
	out[1] = w.lo;
}

int pcramp_word_to_string(const uint64_t in[2], char out[33])
{ // word.h:649-657
	static const char sym[] = "-ACMGRSVTWYHKDBN";
	W128 w;
	w.hi = in[0];
	w.lo = in[1];
	int n = 0;
	for (int i = w_start(w); i <= w_stop(w); ++i) out[n++] = sym[w_get(w, i)];
	out[n] = 0;
	return n;
}

uint32_t pcramp_word_and(const uint64_t a[2], const uint64_t b[2])
{
	W128 x, y;
	x.hi = a[0]; x.lo = a[1];
	y.hi = b[0]; y.lo = b[1];
	return (uint32_t)w_and_count(x, y);
}
uint32_t pcramp_word_size(const uint64_t a[2])
{
	W128 x;
	x.hi = a[0]; x.lo = a[1];
	return (uint32_t)w_size(x);
}
int pcramp_word_start(const uint64_t a[2])
{
	W128 x;
	x.hi = a[0]; x.lo = a[1];
	return w_start(x);
}
int pcramp_word_stop(const uint64_t a[2])
{
	W128 x;
	x.hi = a[0]; x.lo = a[1];
	return w_stop(x);
}
void pcramp_word_complement(const uint64_t a[2], uint64_t out[2])
{
	W128 x;
	x.hi = a[0]; x.lo = a[1];
	const W128 r = w_complement(x);
	out[0] = r.hi;
	out[1] = r.lo;
}
void pcramp_word_center(const uint64_t a[2], uint64_t out[2])
{
	W128 x;
	x.hi = a[0]; x.lo = a[1];
	const W128 r = w_center(x);
	out[0] = r.hi;
	out[1] = r.lo;
}
float pcramp_word_max_overlap(const uint64_t a[2], const uint64_t b[2])
{
	W128 x, y;
	x.hi = a[0]; x.lo = a[1];
	y.hi = b[0]; y.lo = b[1];
	return w_max_overlap(x, y);
}

} // extern "C"

#include "sw_abi.cuh" // K4: Smith-Waterman batches, find_background_match, find_multiplex_background_match
#include "fasta.cuh" // FASTA text -> device-resident collection (parse_fasta + Sequence packing)
#include "xchg.cuh" // multi-GPU: peer-memory exchange of the shards' bitsets fused into the tail of pair scoring
#include "multiplex.cuh" // the multiplex terms of optimize(): multiplex background keys / coverage, pool overlap
#include "optimize_abi.cuh" // optimize() and its moves for a batch of trials
#include "amplicon.cuh" // multiplex bookkeeping: unique amplicons of an assay, pool x amplicon coverage, accept (append / keys / splits)
