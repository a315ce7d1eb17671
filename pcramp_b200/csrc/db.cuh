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

__global__ void validate_hits_kernel(SeqDev sd, PackParams pp, uint64_t *hit_key, const uint32_t *hit_val, uint64_t n_hits,
	uint32_t cand_bits)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_hits) return;
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
	const uint32_t *__restrict__ cand_thr, uint32_t *mask)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_hits) return;
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
