// db.cuh -- from scan hits to the word database the reference calls target_db / background_db.
//
//   validate   drop full-window hits whose window pack() would not have emitted (degeneracy / GC
//              filters, sequence.cpp:127-153) -- done on hits, not on every window
//   tier       select_words.cpp:99-117: per (candidate, sequence) only the best-scoring words at or
//              above the threshold survive.  Hits are sorted by (seq, cand, 63-count), so the
//              survivors of a group are the leading run that shares the first element's count
//   unique     a window chosen by several candidates is stored once per (word, index, loc, strand)
//              occurrence (matched_words is a set and equal_range copies each occurrence once,
//              select_words.cpp:131-138): entry ids (seq, type, strand, pos) are sorted + uniqued
//   materialise word / loc of each entry (pack_entry, seqdev.cuh)
//   order      (word, index, loc, strand) permutation + keys() numbering (pcramp.h:231-256)
#pragma once
#include "scan.cuh"

namespace pcr {

// entry id, low to high: pos[pb-1:0] | type[pb+1:pb] | minus[pb+2] | seq[pb+3 ...] with pb = bits of the longest sequence: sorted ids
// group a sequence's plus-strand entries before its minus-strand entries (what pair scoring iterates over), and the radix sort of
// the ids only has pb + 3 + seq_bits key bits to go through (33 for 20 000 x 30 kb: 5 passes, not 7)
__host__ __device__ __forceinline__ uint64_t entry_id_pack(uint32_t seq, uint32_t type, uint32_t minus, uint32_t pos, uint32_t pb)
{
	return ((uint64_t)seq << (pb + 3u)) | ((uint64_t)minus << (pb + 2u)) | ((uint64_t)type << pb) | pos;
}

// n_dev (may be NULL): the number of hits as the device counted it, when the host has not read it (n_hits is then the capacity)
__global__ void validate_hits_kernel(SeqDev sd, PackParams pp, uint64_t *hit_key, const uint32_t *hit_val, uint64_t n_hits,
	uint32_t cand_bits, const unsigned long long *__restrict__ n_dev)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (n_dev ? min((uint64_t)*n_dev, n_hits) : n_hits)) return;
	const uint64_t k = hit_key[i];
	if (((k >> 1) & 3u) != ENT_FULL) return; // partial words were filtered when they were built
	const uint32_t seq = (uint32_t)(k >> (HIT_GROUP_SHIFT + cand_bits));
	W128 wp, wm;
	int lp, lm;
	if (!pack_entry(sd, seq, ENT_FULL, hit_val[i], pp, wp, wm, lp, lm)) hit_key[i] = ~0ull;
}

// hits sorted ascending by key; invalidated hits (key == ~0) sort last
__global__ void tier_kernel(const uint64_t *__restrict__ hit_key, const uint32_t *__restrict__ hit_val, uint64_t n_hits,
	uint32_t cand_bits, uint32_t pb, uint64_t *entry_id, uint32_t *entry_cand, unsigned long long *n_out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_hits) return;
	const uint64_t k = hit_key[i];
	if (k == ~0ull) return;
	const uint64_t group_first = (k >> HIT_GROUP_SHIFT) << HIT_GROUP_SHIFT;
	uint64_t lo = 0, hi = i; // lower_bound of the group's first possible key in [0, i]
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		if (hit_key[mid] < group_first) lo = mid + 1; else hi = mid;
	}
	if ((hit_key[lo] >> 3) != (k >> 3)) return; // not in the best tier of its (seq, cand) group
	const uint32_t seq = (uint32_t)(k >> (HIT_GROUP_SHIFT + cand_bits));
	const unsigned long long o = warp_slot(n_out);
	entry_id[o] = entry_id_pack(seq, (uint32_t)(k >> 1) & 3u, (uint32_t)k & 1u, hit_val[i], pb);
	// one candidate word that reaches its seed threshold on this window: the anchor of the neighbour filter of pair scoring (score.cuh)
	entry_cand[o] = (uint32_t)(k >> HIT_GROUP_SHIFT) & ((1u << cand_bits) - 1u);
}

// The same best-tier rule without sorting the hits: a (sequence, candidate) table of the best match count seen
// (select_words.cpp:93-117 keeps, per candidate and sequence, only the words of the highest tier), filled with atomicMax and
// read back by every hit.  Used when the table (4 bytes per cell) is small enough; the sorted variant above otherwise.
__global__ void tier_best_kernel(const uint64_t *__restrict__ hit_key, uint64_t n_hits, uint32_t cand_bits, uint32_t n_cand, uint32_t *best)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_hits) return;
	const uint64_t k = hit_key[i];
	if (k == ~0ull) return;
	const uint32_t seq = (uint32_t)(k >> (HIT_GROUP_SHIFT + cand_bits)), cand = (uint32_t)(k >> HIT_GROUP_SHIFT) & ((1u << cand_bits) - 1u);
	atomicMax(best + (size_t)seq * n_cand + cand, 64u - ((uint32_t)(k >> 3) & 63u)); // count + 1
}

__global__ void tier_table_kernel(const uint64_t *__restrict__ hit_key, const uint32_t *__restrict__ hit_val, uint64_t n_hits, uint32_t cand_bits,
	uint32_t n_cand, const uint32_t *__restrict__ best, uint32_t pb, uint64_t *entry_id, uint32_t *entry_cand, unsigned long long *n_out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_hits) return;
	const uint64_t k = hit_key[i];
	if (k == ~0ull) return;
	const uint32_t seq = (uint32_t)(k >> (HIT_GROUP_SHIFT + cand_bits)), cand = (uint32_t)(k >> HIT_GROUP_SHIFT) & ((1u << cand_bits) - 1u);
	if (best[(size_t)seq * n_cand + cand] != 64u - ((uint32_t)(k >> 3) & 63u)) return; // not in the best tier of its (seq, cand) group
	const unsigned long long o = warp_slot(n_out);
	entry_id[o] = entry_id_pack(seq, (uint32_t)(k >> 1) & 3u, (uint32_t)k & 1u, hit_val[i], pb);
	entry_cand[o] = cand;
}

// The same table with ONE BYTE per cell: bit t of a cell = "a hit with count = threshold + t was seen" (every hit of a candidate
// has count >= its threshold, so t <= size - threshold; the host uses this form when that is at most 7 for every oligo size).
// 20 000 sequences x 2 000 candidates are then 40 MB instead of 160 MB -- resident in L2, where the 160 MB table turned every
// atomic into a DRAM round trip (tier_best + tier_table were 0.20 ms of the step) -- and the best tier is the highest bit set.
__global__ void tier_mask_kernel(const uint64_t *__restrict__ hit_key, uint64_t n_hits, uint32_t cand_bits, uint32_t n_cand,
	const uint32_t *__restrict__ cand_thr, uint32_t *mask, const unsigned long long *__restrict__ n_dev)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (n_dev ? min((uint64_t)*n_dev, n_hits) : n_hits)) return;
	const uint64_t k = hit_key[i];
	if (k == ~0ull) return;
	const uint32_t seq = (uint32_t)(k >> (HIT_GROUP_SHIFT + cand_bits)), cand = (uint32_t)(k >> HIT_GROUP_SHIFT) & ((1u << cand_bits) - 1u);
	const uint32_t count = 63u - ((uint32_t)(k >> 3) & 63u), t = min(count - __ldg(cand_thr + cand), 7u);
	const uint64_t cell = (uint64_t)seq * n_cand + cand;
	atomicOr(mask + (cell >> 2), 1u << (8u * (uint32_t)(cell & 3ull) + t));
}

__global__ void tier_mask_select_kernel(const uint64_t *__restrict__ hit_key, const uint32_t *__restrict__ hit_val, uint64_t n_hits, uint32_t cand_bits,
	uint32_t n_cand, const uint32_t *__restrict__ cand_thr, const uint32_t *__restrict__ mask, uint32_t pb, uint64_t *entry_id, uint32_t *entry_cand,
	unsigned long long *n_out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_hits) return;
	const uint64_t k = hit_key[i];
	if (k == ~0ull) return;
	const uint32_t seq = (uint32_t)(k >> (HIT_GROUP_SHIFT + cand_bits)), cand = (uint32_t)(k >> HIT_GROUP_SHIFT) & ((1u << cand_bits) - 1u);
	const uint32_t count = 63u - ((uint32_t)(k >> 3) & 63u), t = min(count - __ldg(cand_thr + cand), 7u);
	const uint64_t cell = (uint64_t)seq * n_cand + cand;
	const uint32_t byte = (mask[cell >> 2] >> (8u * (uint32_t)(cell & 3ull))) & 255u;
	if ((byte >> (t + 1u)) != 0u) return; // a higher tier exists for this (sequence, candidate)
	const unsigned long long o = warp_slot(n_out);
	entry_id[o] = entry_id_pack(seq, (uint32_t)(k >> 1) & 3u, (uint32_t)k & 1u, hit_val[i], pb);
	entry_cand[o] = cand;
}

// ---- the segmented form of "tier -> unique -> grouped by (sequence, strand), sorted by position" (the default) ---------------------
// The surviving hits are not sorted as one list.  Every (sequence, strand) run is a SEGMENT: the hits of the best tiers are counted
// per segment, an exclusive scan gives the segments their slots, a second pass over the hits scatters (type | position, candidate)
// into them, and one CTA per segment sorts its slots in shared memory (a few dozen values on average), drops the duplicates
// (a window chosen by several candidates is one entry; the smallest candidate rides along) and counts what is left; a second
// scan places the unique entries, and the materialise kernel walks the segments.  Seven small launches, no library sort (the
// radix sort of 33-bit ids it replaces: five onesweep passes + unique-by-key = 19 launches), and nothing the host has to read.
__device__ __forceinline__ bool tier_mask_keep(uint64_t k, uint32_t cand_bits, uint32_t n_cand, const uint32_t *__restrict__ cand_thr,
	const uint32_t *__restrict__ mask, uint32_t &seq, uint32_t &cand)
{
	if (k == ~0ull) return false;
	seq = (uint32_t)(k >> (HIT_GROUP_SHIFT + cand_bits));
	cand = (uint32_t)(k >> HIT_GROUP_SHIFT) & ((1u << cand_bits) - 1u);
	const uint32_t count = 63u - ((uint32_t)(k >> 3) & 63u), t = min(count - __ldg(cand_thr + cand), 7u);
	const uint64_t cell = (uint64_t)seq * n_cand + cand;
	const uint32_t byte = (mask[cell >> 2] >> (8u * (uint32_t)(cell & 3ull))) & 255u;
	return (byte >> (t + 1u)) == 0u; // no higher tier exists for this (sequence, candidate)
}

__global__ void seg_count_kernel(const uint64_t *__restrict__ hit_key, const unsigned long long *__restrict__ n_hits_ptr, uint64_t cap, uint32_t cand_bits,
	uint32_t n_cand, const uint32_t *__restrict__ cand_thr, const uint32_t *__restrict__ mask, uint32_t *seg_cnt)
{
	const uint64_t n_hits = min((uint64_t)*n_hits_ptr, cap);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_hits; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t seq, cand;
		const uint64_t k = hit_key[i];
		if (tier_mask_keep(k, cand_bits, n_cand, cand_thr, mask, seq, cand)) atomicAdd(seg_cnt + 2u * seq + ((uint32_t)k & 1u), 1u);
	}
}

// exclusive scan of n values by ONE CTA (n = 2 x sequences: 40 000 for the bench) in chunks of 4096 (one uint4 per thread, coalesced;
// warp shuffles, 32 warp totals through shared memory, a running carry); out[n] = the total.  in == out allowed.
__global__ void __launch_bounds__(1024) seg_scan_kernel(const uint32_t *in, uint32_t n, uint32_t *out)
{
	__shared__ uint32_t s_warp[32];
	__shared__ uint32_t s_carry;
	const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
	if (t == 0u) s_carry = 0u;
	__syncthreads();
	for (uint32_t base = 0; base < n; base += 4096u) {
		const uint32_t i = base + 4u * t;
		uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
		if (i + 3u < n && (((uintptr_t)(in + i)) & 15u) == 0u) {
			const uint4 v = *reinterpret_cast<const uint4 *>(in + i);
			x0 = v.x; x1 = v.y; x2 = v.z; x3 = v.w;
		} else {
			if (i < n) x0 = in[i];
			if (i + 1u < n) x1 = in[i + 1u];
			if (i + 2u < n) x2 = in[i + 2u];
			if (i + 3u < n) x3 = in[i + 3u];
		}
		const uint32_t mine = x0 + x1 + x2 + x3;
		uint32_t incl = mine;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
			if ((int)lane >= d) incl += y;
		}
		if (lane == 31u) s_warp[warp] = incl;
		__syncthreads();
		if (warp == 0u) {
			uint32_t w = s_warp[lane];
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
				if ((int)lane >= d) w += y;
			}
			s_warp[lane] = w; // inclusive over the warps
		}
		__syncthreads();
		const uint32_t carry = s_carry;
		uint32_t run = carry + (warp ? s_warp[warp - 1u] : 0u) + incl - mine;
		if (i < n) out[i] = run;
		run += x0;
		if (i + 1u < n) out[i + 1u] = run;
		run += x1;
		if (i + 2u < n) out[i + 2u] = run;
		run += x2;
		if (i + 3u < n) out[i + 3u] = run;
		__syncthreads();
		if (t == 1023u) s_carry = carry + s_warp[31];
		__syncthreads();
	}
	if (t == 0u) out[n] = s_carry;
}

__global__ void seg_scatter_kernel(const uint64_t *__restrict__ hit_key, const uint32_t *__restrict__ hit_val, const unsigned long long *__restrict__ n_hits_ptr,
	uint64_t cap, uint32_t cand_bits, uint32_t n_cand, const uint32_t *__restrict__ cand_thr, const uint32_t *__restrict__ mask, uint32_t pb,
	const uint32_t *__restrict__ seg_off, uint32_t *seg_cursor, uint64_t *slot, uint32_t *slot_seg)
{
	const uint64_t n_hits = min((uint64_t)*n_hits_ptr, cap);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_hits; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t seq, cand;
		const uint64_t k = hit_key[i];
		if (!tier_mask_keep(k, cand_bits, n_cand, cand_thr, mask, seq, cand)) continue;
		const uint32_t seg = 2u * seq + ((uint32_t)k & 1u);
		const uint32_t at = __ldg(seg_off + seg) + atomicAdd(seg_cursor + seg, 1u);
		// (type | position) above the candidate: sorting the 64-bit values groups equal windows, smallest candidate first
		slot[at] = ((uint64_t)((((uint32_t)(k >> 1) & 3u) << pb) | hit_val[i]) << 32) | cand;
		slot_seg[at] = seg; // (the materialise kernel runs one thread per slot)
	}
}

constexpr uint32_t SEG_WARPS = 4u;         // small segments: one warp each, SEG_WARPS per CTA
constexpr uint32_t SEG_WARP_SLOTS = 256u;  // ... sorted in the warp's own slice of shared memory
constexpr uint32_t SEG_BIG_THREADS = 256u; // large segments: one CTA each
constexpr uint32_t SEG_BIG_SLOTS = 4096u;  // ... in shared memory up to this many slots, in place in global memory beyond

// the bitonic network in its ascending-only form (the first stage of every merge compares mirror images, i with i ^ (k - 1)), so
// the padding up to a power of two stays virtual: a slot at or past n is +infinity and never has to move.  `sync` separates stages.
template <class Sync>
__device__ __forceinline__ void seg_bitonic(uint64_t *v, uint32_t n, uint32_t t, uint32_t nt, Sync sync)
{
	uint32_t m = 1u;
	while (m < n) m <<= 1;
	for (uint32_t k = 2u; k <= m; k <<= 1) {
		for (uint32_t i = t; i < n; i += nt) {
			const uint32_t l = i ^ (k - 1u);
			if (l > i && l < n) {
				const uint64_t a = v[i], c = v[l];
				if (a > c) { v[i] = c; v[l] = a; }
			}
		}
		sync();
		for (uint32_t j = k >> 2; j > 0u; j >>= 1) {
			for (uint32_t i = t; i < n; i += nt) {
				const uint32_t l = i ^ j;
				if (l > i && l < n) {
					const uint64_t a = v[i], c = v[l];
					if (a > c) { v[i] = c; v[l] = a; }
				}
			}
			sync();
		}
	}
}

// segments of at most SEG_WARP_SLOTS slots: one WARP sorts, drops the duplicates (the first of each run of equal (type | position)
// stays, i.e. the smallest candidate) and writes the survivors to the front of the segment's slots; only warp-level barriers
__global__ void __launch_bounds__(SEG_WARPS * 32u) seg_sort_small_kernel(const uint32_t *__restrict__ seg_off, uint32_t n_seg, uint64_t *slot, uint32_t pb,
	uint32_t *uniq_cnt, uint32_t *full_cnt, uint32_t *big_list, unsigned int *n_big)
{
	__shared__ uint64_t s_all[SEG_WARPS * SEG_WARP_SLOTS];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint64_t *v = s_all + warp * SEG_WARP_SLOTS;
	for (uint32_t seg = blockIdx.x * SEG_WARPS + warp; seg < n_seg; seg += gridDim.x * SEG_WARPS) {
		const uint32_t b = seg_off[seg], n = seg_off[seg + 1] - b;
		if (n > SEG_WARP_SLOTS) { // left to seg_sort_big_kernel
			if (lane == 0u) big_list[atomicAdd(n_big, 1u)] = seg;
			continue;
		}
		if (n == 0u) {
			if (lane == 0u) { uniq_cnt[seg] = 0u; full_cnt[seg] = 0u; }
			continue;
		}
		__syncwarp();
		for (uint32_t i = lane; i < n; i += 32u) v[i] = slot[b + i];
		__syncwarp();
		if (n > 1u) seg_bitonic(v, n, lane, 32u, [] { __syncwarp(); });
		uint32_t run = 0u, fulls = 0u;
		for (uint32_t base = 0u; base < n; base += 32u) {
			const uint32_t i = base + lane;
			uint64_t x = 0;
			bool head = false;
			if (i < n) {
				x = v[i];
				head = i == 0u || (uint32_t)(v[i - 1u] >> 32) != (uint32_t)(x >> 32);
			}
			const uint32_t hm = __ballot_sync(0xffffffffu, head);
			const uint32_t fm = __ballot_sync(0xffffffffu, head && (((uint32_t)(x >> 32)) >> pb) == 0u); // ENT_FULL == 0
			if (head) slot[b + run + (uint32_t)__popc(hm & ((1u << lane) - 1u))] = x;
			run += (uint32_t)__popc(hm);
			fulls += (uint32_t)__popc(fm);
		}
		if (lane == 0u) { uniq_cnt[seg] = run; full_cnt[seg] = fulls; }
	}
}

// inclusive scan over the threads of a SEG_BIG_THREADS CTA; total = the sum
__device__ __forceinline__ uint32_t seg_scan_cta(uint32_t x, uint32_t *s_w, uint32_t &total)
{
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	uint32_t incl = x;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
		if ((int)lane >= d) incl += y;
	}
	__syncthreads(); // s_w may still be read by the previous call
	if (lane == 31u) s_w[warp] = incl;
	__syncthreads();
	uint32_t before = 0u, all = 0u;
	#pragma unroll
	for (uint32_t w = 0; w < SEG_BIG_THREADS / 32u; ++w) {
		const uint32_t y = s_w[w];
		if (w < warp) before += y;
		all += y;
	}
	total = all;
	return incl + before;
}

// segments longer than that: one CTA each (a collection of few, long sequences: hundreds of hits per segment)
__global__ void __launch_bounds__(SEG_BIG_THREADS) seg_sort_big_kernel(const uint32_t *__restrict__ seg_off, const uint32_t *__restrict__ big_list,
	const unsigned int *__restrict__ n_big, uint64_t *slot, uint32_t pb, uint32_t *uniq_cnt, uint32_t *full_cnt)
{
	__shared__ uint64_t s_v[SEG_BIG_SLOTS];
	__shared__ uint32_t s_w[SEG_BIG_THREADS / 32u];
	__shared__ uint64_t s_last;
	const uint32_t t = threadIdx.x;
	const uint32_t total_big = *n_big;
	for (uint32_t q = blockIdx.x; q < total_big; q += gridDim.x) {
		const uint32_t seg = big_list[q];
		const uint32_t b = seg_off[seg], n = seg_off[seg + 1] - b;
		const bool in_smem = n <= SEG_BIG_SLOTS;
		uint64_t *v = in_smem ? s_v : slot + b;
		__syncthreads(); // the previous segment is done with s_v
		if (in_smem) {
			for (uint32_t i = t; i < n; i += SEG_BIG_THREADS) s_v[i] = slot[b + i];
			__syncthreads();
		}
		seg_bitonic(v, n, t, SEG_BIG_THREADS, [] { __syncthreads(); });
		uint32_t run = 0u, fulls = 0u;
		for (uint32_t base = 0u; base < n; base += SEG_BIG_THREADS) {
			const uint32_t i = base + t;
			uint64_t x = 0;
			bool head = false;
			if (i < n) {
				x = v[i];
				const uint64_t prev = (i == 0u) ? ~x : (t == 0u ? s_last : v[i - 1u]);
				head = (uint32_t)(prev >> 32) != (uint32_t)(x >> 32);
			}
			uint32_t total, n_full;
			const uint32_t incl = seg_scan_cta(head ? 1u : 0u, s_w, total); // (its barriers: every value of the chunk has been read)
			(void)seg_scan_cta((head && (((uint32_t)(x >> 32)) >> pb) == 0u) ? 1u : 0u, s_w, n_full);
			if (i < n && (t == SEG_BIG_THREADS - 1u || i == n - 1u)) s_last = x;
			if (head) slot[b + run + incl - 1u] = x; // rank <= i: behind every value still to be read
			run += total;
			fulls += n_full;
			__syncthreads();
		}
		if (t == 0u) { uniq_cnt[seg] = run; full_cnt[seg] = fulls; }
	}
}

// word planes, loc, strand, seq, candidate of each unique entry: one thread per slot (slot_seg names its segment; the unique values
// sit at the front of the segment's slots); ent_off = exclusive scan of uniq_cnt.  Thread s < n_seg also writes full_end[s].
__global__ void __launch_bounds__(256) seg_materialise_kernel(SeqDev sd, PackParams pp, const uint32_t *__restrict__ seg_off,
	const uint32_t *__restrict__ ent_off, const uint32_t *__restrict__ full_cnt, uint32_t n_seg, const uint64_t *__restrict__ slot,
	const uint32_t *__restrict__ slot_seg, uint32_t pb, uint4 *e_planes, uint32_t *e_seq, int32_t *e_loc, uint32_t *e_strand, uint32_t *e_cand,
	uint32_t *e_id, uint32_t *full_end)
{
	const uint32_t n_slots = seg_off[n_seg];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < max(n_slots, n_seg); i += gridDim.x * blockDim.x) {
		if (i < n_seg) full_end[i] = ent_off[i] + full_cnt[i];
		if (i >= n_slots) continue;
		const uint32_t seg = slot_seg[i];
		const uint32_t r = i - __ldg(seg_off + seg), dst0 = __ldg(ent_off + seg);
		if (r >= __ldg(ent_off + seg + 1u) - dst0) continue; // a duplicate's slot
		const uint64_t x = slot[i];
		const uint32_t seq = seg >> 1, minus = seg & 1u;
		const uint32_t id = (uint32_t)(x >> 32), type = id >> pb, pos = id & ((1u << pb) - 1u);
		Planes4 pl;
		int loc;
		pack_entry_planes(sd, seq, type, pos, pp, minus != 0u, pl, loc); // validated earlier: the event emits
		const uint32_t o = dst0 + r;
		e_planes[o] = make_uint4(pl.a, pl.c, pl.g, pl.t);
		e_seq[o] = seq;
		e_loc[o] = loc;
		e_strand[o] = minus ? STRAND_MINUS : STRAND_PLUS;
		e_cand[o] = (uint32_t)x;
		e_id[o] = id;
	}
}

// the words themselves and the (index, loc, strand) sort key: only the canonical order / keys() / db_copy / the Smith-Waterman
// background test read them, so they are materialised on demand from the stored (type | position) of every entry
__global__ void words_kernel(SeqDev sd, PackParams pp, const uint32_t *__restrict__ e_id, const uint32_t *__restrict__ e_seq,
	const uint32_t *__restrict__ e_strand, uint64_t n, uint32_t pb, uint64_t *w_hi, uint64_t *w_lo, uint64_t *order_key)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t id = e_id[i], type = id >> pb, pos = id & ((1u << pb) - 1u), seq = e_seq[i];
	const bool minus = e_strand[i] == STRAND_MINUS;
	W128 wp, wm;
	int lp = 0, lm = 0;
	wp.hi = wp.lo = wm.hi = wm.lo = 0;
	pack_entry(sd, seq, type, pos, pp, wp, wm, lp, lm);
	const W128 w = minus ? wm : wp;
	w_hi[i] = w.hi;
	w_lo[i] = w.lo;
	order_key[i] = ((uint64_t)seq << 34) | ((uint64_t)((uint32_t)(minus ? lm : lp) ^ 0x80000000u) << 2) | (minus ? 2u : 1u);
}

// word, loc, strand, seq of each unique entry (entry ids sorted => grouped by sequence)
__global__ void materialise_kernel(SeqDev sd, PackParams pp, const uint64_t *__restrict__ entry_id, uint64_t n, uint32_t pb, uint64_t *w_hi,
	uint64_t *w_lo, uint4 *e_planes, uint32_t *e_seq, int32_t *e_loc, uint32_t *e_strand, uint64_t *order_key)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t id = entry_id[i];
	const uint32_t seq = (uint32_t)(id >> (pb + 3u)), type = (uint32_t)(id >> pb) & 3u, minus = (uint32_t)(id >> (pb + 2u)) & 1u;
	const uint32_t pos = (uint32_t)(id & ((1ull << pb) - 1ull));
	W128 wp, wm;
	int lp = 0, lm = 0;
	wp.hi = wp.lo = wm.hi = wm.lo = 0;
	pack_entry(sd, seq, type, pos, pp, wp, wm, lp, lm); // validated earlier, always true here
	const W128 w = minus ? wm : wp;
	const int loc = minus ? lm : lp;
	w_hi[i] = w.hi;
	w_lo[i] = w.lo;
	const Planes4 pl = w_planes(w); // what pair scoring compares against (score.cuh)
	e_planes[i] = make_uint4(pl.a, pl.c, pl.g, pl.t);
	e_seq[i] = seq;
	e_loc[i] = loc;
	e_strand[i] = minus ? STRAND_MINUS : STRAND_PLUS;
	// least significant sort key of the canonical order: (index, loc, strand)
	order_key[i] = ((uint64_t)seq << 34) | ((uint64_t)((uint32_t)loc ^ 0x80000000u) << 2) | (minus ? 2u : 1u);
}

// Sequence::pack of ONE sequence, entry by entry (used by pcramp_gpu_pack; the scan never materialises this)
__global__ void pack_dump_kernel(SeqDev sd, uint32_t seq, PackParams pp, uint64_t *o_words, int32_t *o_loc, uint32_t *o_strand,
	unsigned long long *n_out)
{
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
		o_words[2 * o] = wp.hi; o_words[2 * o + 1] = wp.lo; o_loc[o] = lp; o_strand[o] = STRAND_PLUS;
		o_words[2 * o + 2] = wm.hi; o_words[2 * o + 3] = wm.lo; o_loc[o + 1] = lm; o_strand[o + 1] = STRAND_MINUS;
	}
}

__global__ void iota_kernel(uint32_t *p, uint64_t n)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = (uint32_t)i;
}

__global__ void gather_u64_kernel(const uint64_t *__restrict__ src, const uint32_t *__restrict__ perm, uint64_t *dst, uint64_t n)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[i] = src[perm[i]];
}

// perm = canonical order; head[i] = 1 where entry perm[i] starts a new word
__global__ void key_heads_kernel(const uint64_t *__restrict__ w_hi, const uint64_t *__restrict__ w_lo, const uint32_t *__restrict__ perm,
	uint32_t *head, uint64_t n)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (i == 0) { head[0] = 1; return; }
	const uint32_t a = perm[i], b = perm[i - 1];
	head[i] = (w_hi[a] != w_hi[b] || w_lo[a] != w_lo[b]) ? 1u : 0u;
}

// key index of every entry (in entry-id order) and the letter planes of every unique word, for the key-matrix form of K2
__global__ void key_index_kernel(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ head, const uint32_t *__restrict__ keyrank,
	const uint4 *__restrict__ e_planes, uint64_t n, uint32_t *e_key, uint4 *key_planes)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t e = perm[i], k = keyrank[i] - 1u;
	e_key[e] = k;
	if (head[i]) key_planes[k] = e_planes[e];
}

// first entry of each (sequence, strand) run in the grouped arrays: off[2*s + minus] = lower_bound; off[2*n_seq] = n
__global__ void seq_offsets_kernel(const uint32_t *__restrict__ e_seq, const uint32_t *__restrict__ e_strand, uint64_t n, uint32_t n_seq,
	uint32_t *off)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k > 2u * n_seq) return;
	uint64_t lo = 0, hi = n;
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		const uint32_t key = 2u * e_seq[mid] + (e_strand[mid] == STRAND_MINUS ? 1u : 0u);
		if (key < k) lo = mid + 1; else hi = mid;
	}
	off[k] = (uint32_t)lo;
}

__global__ void db_export_kernel(const uint64_t *__restrict__ w_hi, const uint64_t *__restrict__ w_lo, const uint32_t *__restrict__ e_seq,
	const int32_t *__restrict__ e_loc, const uint32_t *__restrict__ e_strand, const uint32_t *__restrict__ perm,
	const uint32_t *__restrict__ key_rank_incl, uint64_t n, uint64_t *o_words, uint32_t *o_index, int32_t *o_loc,
	uint32_t *o_strand, uint32_t *o_key, uint64_t *o_keys)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t e = perm[i];
	const uint32_t k = key_rank_incl[i] - 1u;
	o_words[2 * i] = w_hi[e];
	o_words[2 * i + 1] = w_lo[e];
	o_index[i] = e_seq[e];
	o_loc[i] = e_loc[e];
	o_strand[i] = e_strand[e];
	o_key[i] = k;
	if (i == 0 || key_rank_incl[i - 1] != key_rank_incl[i]) {
		o_keys[2 * (uint64_t)k] = w_hi[e];
		o_keys[2 * (uint64_t)k + 1] = w_lo[e];
	}
}

} // namespace pcr
