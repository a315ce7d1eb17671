// index.cuh -- K1, indexed seed scan: the text side of the pigeonhole filter turned around.
//
// scan_seed_kernel (scan.cuh) walks every text position and asks "which primers have a seed here?" -- ~5 candidate
// verifications per position, 3 x 10^9 per batch of 4000 patterns on 6 x 10^8 positions, all of them integer work on an
// SM-resident table.  This path asks the opposite question.  The collection is indexed ONCE (per upload / split): every
// text position sorted by the 2-bit code of the 12 bases that start there, each entry carrying its position and the code
// bit-planes of the 48 bases around it (16 before, 32 from the position on).  A pattern with e allowed mismatches is cut
// into s = e/2 + 1 segments; one of them has at most ONE mismatch (pigeonhole), hence so has its k-base prefix
// (k = min(segment length, 12)).  For every segment the 1 + 3k k-mers within Hamming distance 1 of the prefix are looked
// up as prefix ranges of the sorted index, and every entry in range is verified from its own 16 bytes: shift the context
// planes to the alignment, one LOP3 per letter plane, POPC, compare -- no table in shared memory, no text access, a
// coalesced 16-byte stream.  k = 9..12 makes the ranges 10^2..10^3 x more selective than the 5/6/7-base exact seeds
// (4^-10 against 4^-6 per lookup), so a batch touches a few GB of index instead of issuing 10^10 warp instructions, and
// the kernel is bound by HBM bandwidth, not by integer issue.
//
// Exactness: (i) pigeonhole as above; (ii) the text k-mer of a true alignment is itself one of the enumerated
// neighbours, exactly one, so an alignment is found once per qualifying segment and reported through the leftmost
// qualifying segment only; (iii) alignments that start in a "dirty" 32-base group (text with a degenerate base nearby)
// are left to scan_groups_kernel, exactly as in scan_seed_kernel; (iv) patterns whose segment prefixes contain a
// degenerate base with more than 16 letter combinations (or an empty position) or are shorter than 8 stay on scan_seed_kernel /
// scan_full_kernel; the degenerate positions of a prefix are enumerated letter by letter.  Hits leave through the same
// emit_family / HitSink as the other scan kernels.
#pragma once
#include "scan.cuh"

namespace pcr {

constexpr uint32_t IDX_K = 12u;                     // bases per index key
constexpr uint32_t IDX_KMIN = 8u;                   // shortest usable segment prefix
constexpr uint32_t IDX_CODES = 1u << (2u * IDX_K);  // 16 M prefix offsets
constexpr uint32_t IDX_MAX_SEG = 4u;                // e <= 7
constexpr uint32_t IDX_SLOTS = IDX_MAX_SEG * (1u + 3u * IDX_K);
constexpr uint32_t IDX_CTX_BEFORE = 16u;

struct TextIndex { // one part of the index: the sequences [seq_lo, seq_lo + n_seq) of the collection, positions counted from the part's start
	const uint4 *entries;   // sorted by 12-mer code: {position in the part, plane0[31:0], plane1[31:0], plane0[47:32] | plane1[47:32] << 16}
	const uint32_t *off;    // IDX_CODES + 1: first entry whose code is >= c
	const uint32_t *cum;    // n_seq + 1: position of each sequence's first base (as of the build: splits do not move it)
	const uint32_t *blk;    // (n >> IDX_BLK_SHIFT) + 2: sequence (part-relative) that holds position b << IDX_BLK_SHIFT
	uint32_t n;             // entries
	uint32_t seq_lo, n_seq;
};
constexpr uint32_t IDX_BLK_SHIFT = 10u;

// segments of a pattern of n bases with e allowed mismatches
__host__ __device__ __forceinline__ uint32_t idx_segments(uint32_t e) { return e / 2u + 1u; }
__host__ __device__ __forceinline__ void idx_segment(uint32_t n, uint32_t segs, uint32_t i, uint32_t &o, uint32_t &k)
{
	o = (i * n) / segs;
	const uint32_t len = ((i + 1u) * n) / segs - o;
	k = len < IDX_K ? len : IDX_K;
}

// letter code (A=0 C=1 G=2 T=3) of pattern position j, or 4 when the position is degenerate / empty
__device__ __forceinline__ uint32_t idx_letter(const uint4 &m, uint32_t j)
{
	const uint32_t a = (m.x >> j) & 1u, c = (m.y >> j) & 1u, g = (m.z >> j) & 1u, t = (m.w >> j) & 1u;
	if (a + c + g + t != 1u) return 4u;
	return c | (g << 1) | (t * 3u);
}

// the letters pattern position j admits: bit 0 = A, 1 = C, 2 = G, 3 = T (the letter codes of the index key)
__device__ __forceinline__ uint32_t idx_set(const uint4 &m, uint32_t j)
{
	return ((m.x >> j) & 1u) | (((m.y >> j) & 1u) << 1) | (((m.z >> j) & 1u) << 2) | (((m.w >> j) & 1u) << 3);
}
__device__ __forceinline__ uint32_t idx_nth_letter(uint32_t set, uint32_t n)
{ // the n-th letter (ascending) of a non-empty set
	for (uint32_t l = 0; l < 4u; ++l)
		if ((set >> l) & 1u) {
			if (n == 0u) return l;
			--n;
		}
	return 0u;
}
constexpr uint32_t IDX_MAX_COMBOS = 16u; // letter combinations of the degenerate positions of one segment prefix (primers of degeneracy <= 16)

// can this (seedable) pattern go through the index?  every segment prefix >= IDX_KMIN bases, no empty position, and at most
// IDX_MAX_COMBOS letter combinations over its degenerate positions (each combination is queried on its own, see index_query_kernel)
__device__ __forceinline__ bool idx_indexable(const uint4 &m, uint32_t meta2)
{
	const uint32_t n = (meta2 >> 10) & 63u, e = (meta2 >> 16) & 63u, cls = (meta2 >> 22) & 7u;
	if (cls == 0u || n == 0u || n > 32u) return false;
	const uint32_t segs = idx_segments(e);
	if (segs > IDX_MAX_SEG) return false;
	for (uint32_t i = 0; i < segs; ++i) {
		uint32_t o, k;
		idx_segment(n, segs, i, o, k);
		if (k < IDX_KMIN || o > IDX_CTX_BEFORE) return false; // the entry's context reaches 16 bases back
		uint32_t combos = 1u;
		for (uint32_t j = 0; j < k; ++j) {
			const uint32_t c = (uint32_t)__popc(idx_set(m, o + j));
			if (c == 0u) return false;
			combos *= c;
			if (combos > IDX_MAX_COMBOS) return false;
		}
	}
	return true;
}

__device__ __forceinline__ uint32_t idx_spread12(uint32_t v)
{ // 12 bits -> even bit positions
	v = (v | (v << 8)) & 0x00FF00FFu;
	v = (v | (v << 4)) & 0x0F0F0F0Fu;
	v = (v | (v << 2)) & 0x33333333u;
	v = (v | (v << 1)) & 0x55555555u;
	return v;
}

// code bit-planes b0 = C|T, b1 = G|T of bases [x - 16, x + 32) of one sequence (zero outside the sequence)
__device__ __forceinline__ void idx_context(const SeqDev &sd, uint32_t seq, uint32_t x, uint64_t &c0, uint64_t &c1)
{
	const uint64_t gbase = sd.grp_off[seq];
	const uint32_t ngrp = (uint32_t)(sd.grp_off[seq + 1] - gbase);
	const uint32_t g = x >> 5;
	const uint4 z = make_uint4(0, 0, 0, 0);
	const uint4 pm = g > 0u ? __ldg(sd.planes + gbase + g - 1u) : z;
	const uint4 pc = __ldg(sd.planes + gbase + g);
	const uint4 pp = (g + 1u < ngrp) ? __ldg(sd.planes + gbase + g + 1u) : z;
	const uint32_t start = (x & 31u) + 32u - IDX_CTX_BEFORE; // bit of base x - 16 in the 96-bit string pm:pc:pp
	uint32_t lo, hi;
	take64(pm.y | pm.w, pc.y | pc.w, pp.y | pp.w, start, lo, hi);
	c0 = ((uint64_t)(hi & 0xFFFFu) << 32) | lo;
	take64(pm.z | pm.w, pc.z | pc.w, pp.z | pp.w, start, lo, hi);
	c1 = ((uint64_t)(hi & 0xFFFFu) << 32) | lo;
}

__device__ __forceinline__ uint32_t idx_seq_of(const uint32_t *__restrict__ cum, uint32_t n_seq, uint32_t gpos)
{ // last sequence with cum[s] <= gpos (empty sequences share an offset with their successor: take the last)
	uint32_t lo = 0, hi = n_seq;
	while (hi - lo > 1u) {
		const uint32_t mid = (lo + hi) >> 1;
		if (__ldg(cum + mid) <= gpos) lo = mid; else hi = mid;
	}
	return lo;
}

// the same through the block table: the answer lies between the sequences that hold the ends of the candidate's 1024-position
// block -- one or two probes instead of log2(n_seq) dependent loads per candidate (index_hits_kernel was 0.14 ms of the step)
__device__ __forceinline__ uint32_t idx_seq_of_fast(const TextIndex &ix, uint32_t n_seq, uint32_t gpos)
{
	const uint32_t b = gpos >> IDX_BLK_SHIFT;
	uint32_t lo = __ldg(ix.blk + b), hi = min(__ldg(ix.blk + b + 1u) + 1u, n_seq);
	while (hi - lo > 1u) {
		const uint32_t mid = (lo + hi) >> 1;
		if (__ldg(ix.cum + mid) <= gpos) lo = mid; else hi = mid;
	}
	return lo;
}

// ---- build ---------------------------------------------------------------------------------------------------
__global__ void index_key_kernel(SeqDev sd, const uint32_t *__restrict__ cum, uint32_t seq_lo, uint32_t n_seq, uint32_t n_pos, uint32_t *key,
	uint32_t *val)
{
	const uint32_t gpos = blockIdx.x * blockDim.x + threadIdx.x;
	if (gpos >= n_pos) return;
	const uint32_t rel = idx_seq_of(cum, n_seq, gpos), seq = seq_lo + rel;
	uint64_t c0, c1;
	idx_context(sd, seq, gpos - __ldg(cum + rel), c0, c1);
	const uint32_t b0 = (uint32_t)(c0 >> IDX_CTX_BEFORE) & 0xFFFu, b1 = (uint32_t)(c1 >> IDX_CTX_BEFORE) & 0xFFFu;
	// first base most significant: reverse the 12 bits, then interleave (b1 above b0)
	const uint32_t r0 = __brev(b0) >> 20, r1 = __brev(b1) >> 20;
	key[gpos] = idx_spread12(r0) | (idx_spread12(r1) << 1);
	val[gpos] = gpos;
}

__global__ void index_offsets_kernel(const uint32_t *__restrict__ key_sorted, uint32_t n, uint32_t *off)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n) return;
	const int64_t prev = (i == 0u) ? -1 : (int64_t)key_sorted[i - 1u];
	const int64_t cur = (i == n) ? (int64_t)IDX_CODES : (int64_t)key_sorted[i];
	for (int64_t c = prev + 1; c <= cur; ++c) off[c] = i;
}

__global__ void index_entry_kernel(SeqDev sd, const uint32_t *__restrict__ cum, uint32_t seq_lo, uint32_t n_seq,
	const uint32_t *__restrict__ pos_sorted, uint32_t n, uint4 *entries)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t gpos = pos_sorted[i];
	const uint32_t rel = idx_seq_of(cum, n_seq, gpos), seq = seq_lo + rel;
	uint64_t c0, c1;
	idx_context(sd, seq, gpos - __ldg(cum + rel), c0, c1);
	entries[i] = make_uint4(gpos, (uint32_t)c0, (uint32_t)c1, (uint32_t)(c0 >> 32) | ((uint32_t)(c1 >> 32) << 16));
}

// ---- per batch: the neighbour queries of every indexable pattern -----------------------------------------------------
struct IdxQuery {
	uint32_t lo, hi;   // entry range
	uint32_t pid;      // pattern (index into the seeded part of the partitioned pattern arrays)
	uint32_t seg;      // segment offset o[7:0] | prefix length k[15:8] | segment number[23:16] | window offset wo[31:24]
};

// Extended seeds.  A segment prefix shorter than the 12-base index key (k = 9..11: 18..23-mers) used to be looked up as a
// PREFIX RANGE -- 4^(12-k) buckets per neighbour, 73 % of all entries streamed came from the k = 9 queries of 18/19-mers.
// But the x = 12 - k bases next to the segment are pattern bases too, and the whole alignment has at most e mismatches:
// a neighbour that already spends m (0 or 1) of them inside the segment leaves e - m for those x bases.  When that is
// fewer than x, only the 12-mers whose extension is within e - m mismatches of the pattern can belong to a hit, and they
// are queried one bucket each (e = 2, k = 9: 37 + 27 * 10 = 307 buckets instead of 28 * 64).  The window is the segment
// plus the next x pattern bases (window offset wo = o), or, for a segment that ends with the pattern, the x bases before
// it (wo = o - x; the bucket of text position X then belongs to the alignment whose primer base 0 is at X - wo).  The
// Allowed codes that are consecutive (the extension sits in the low digits when it follows the segment) are queried as one range.  The
// pigeonhole argument and the leftmost-segment rule are untouched: a true hit is still found through every segment
// whose k-prefix is within one mismatch, exactly once per segment (the text 12-mer at the window is one code).
constexpr uint32_t IDX_EXT_MAX = 64u;  // most single-bucket queries one neighbour may expand into

__device__ __forceinline__ uint32_t idx_ext_count(uint32_t x, uint32_t b)
{ // 12-mers with at most b substitutions in x positions
	uint32_t total = 0, c = 1; // c = C(x, j) 3^j
	for (uint32_t j = 0; j <= b && j <= x; ++j) {
		total += c;
		c = c * (x - j) * 3u / (j + 1u);
	}
	return total;
}

__global__ void index_query_kernel(const uint4 *__restrict__ mask, const uint32_t *__restrict__ meta2, uint32_t n_pat, const uint32_t *__restrict__ off,
	IdxQuery *queries, uint32_t q_cap, unsigned int *n_queries, unsigned int *n_indexed, unsigned long long *n_entries)
{
	// every lane stays to the end: the three counters are bumped once per WARP (a quarter of a million same-address atomics
	// serialise in L2 -- they, not the work, were this kernel's 0.23 ms)
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
	const uint32_t p = t / IDX_SLOTS, slot = t % IDX_SLOTS;
	bool counts = false;
	uint32_t n_emit = 0;        // queries this lane writes
	uint32_t k = 0, o = 0, si = 0, wo = 0, x = 0, budget = 0, ext_pat = 0, n_combo = 1u, sub_letter = 0u, sub_at = 0xFFFFFFFFu;
	uint4 seg_m = make_uint4(0u, 0u, 0u, 0u);
	bool ext_left = false, ext = false;
	uint64_t okm = 0ull; // extended seeds with x <= 3: bit c = extension code c is within the budget
	if (p < n_pat) {
		const uint32_t m2 = meta2[p];
		const uint32_t n = (m2 >> 10) & 63u, e = (m2 >> 16) & 63u, segs = idx_segments(e);
		const uint32_t per = 1u + 3u * IDX_K;
		const uint32_t j = slot % per;
		si = slot / per;
		// (half of a pattern's slots belong to segments it does not have: they leave before the indexability test, which walks
		// every prefix base; slot 0 always takes it -- it counts the pattern)
		const uint4 m = (si < segs) ? mask[p] : make_uint4(0u, 0u, 0u, 0u);
		if (si < segs && idx_indexable(m, m2)) {
			counts = (slot == 0u);
			{
				idx_segment(n, segs, si, o, k);
				uint32_t sub_pos = 0xFFFFFFFFu, sub_alt = 0u;
				bool ok = true;
				if (j > 0u) {
					sub_pos = (j - 1u) / 3u;
					sub_alt = (j - 1u) % 3u;
					ok = sub_pos < k;
				}
				if (ok) {
					// Degenerate prefix positions: the text k-mers within one SET mismatch of the prefix are enumerated without
					// repeats as (which position misses, with which letter outside its set) x (one letter of its set at every other
					// position): n_combo codes per neighbour, each queried like the single code of a plain prefix.
					n_combo = 1u;
					for (uint32_t q = 0; q < k; ++q) {
						const uint32_t set = idx_set(m, o + q);
						if (q == sub_pos) {
							const uint32_t out = ~set & 15u; // the substituted letter is one the pattern does NOT admit here
							if (sub_alt >= (uint32_t)__popc(out)) ok = false;
							else sub_letter = idx_nth_letter(out, sub_alt);
						} else {
							n_combo *= (uint32_t)__popc(set);
						}
					}
				}
				if (ok) {
					seg_m = m;
					sub_at = sub_pos;
					n_emit = 1u;
					wo = o;
					x = IDX_K - k;
					budget = e - (j > 0u ? 1u : 0u); // e >= 1 whenever a substituted neighbour can be a hit; e == 0 -> j > 0 finds nothing
					if (j > 0u && e == 0u) n_emit = 0u;
					if (n_emit && x > 0u && budget < x && idx_ext_count(x, budget) <= IDX_EXT_MAX) {
						// the x pattern bases after the segment, or before it when the pattern ends with the segment
						uint32_t first = 0;
						bool can = false;
						if (o + IDX_K <= n) { first = o + k; can = true; ext_left = false; }
						else if (o >= x && o - x <= IDX_CTX_BEFORE) { first = o - x; can = true; ext_left = true; }
						for (uint32_t q = 0; can && q < x; ++q) {
							const uint32_t l = idx_letter(m, first + q);
							if (l > 3u) can = false;
							ext_pat = (ext_pat << 2) | l;
						}
						if (can) {
							ext = true;
							if (ext_left) wo = o - x;
							n_emit = idx_ext_count(x, budget);
							if (x <= 3u) { // the allowed codes as one 64-bit mask: counted and emitted from its set bits, not by two walks over the codes
								for (uint32_t c = 0; c < (1u << (2u * x)); ++c) {
									const uint32_t d = c ^ ext_pat;
									okm |= (uint64_t)((uint32_t)__popc((d | (d >> 1)) & 0x55u) <= budget ? 1u : 0u) << c;
								}
								// extension in the low digits of the code: runs of consecutive allowed codes are ONE range
								if (!ext_left) n_emit = (uint32_t)__popcll(okm & ~(okm << 1));
							} else if (!ext_left) {
								n_emit = 0u;
								bool prev = false;
								for (uint32_t c = 0; c < (1u << (2u * x)); ++c) {
									const uint32_t d = c ^ ext_pat;
									const bool okc = (uint32_t)__popc((d | (d >> 1)) & 0x55u) <= budget;
									n_emit += (okc && !prev) ? 1u : 0u;
									prev = okc;
								}
							}
						}
					}
				}
			}
		}
	}
	n_emit *= n_combo;
	// reserve: one atomic per warp
	uint32_t incl = n_emit;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
		if ((int)lane >= d) incl += v;
	}
	const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
	const uint32_t cnt = __ballot_sync(0xffffffffu, counts);
	uint32_t base = 0u;
	if (lane == 0u) {
		if (total) base = atomicAdd(n_queries, total);
		if (cnt) atomicAdd(n_indexed, (unsigned int)__popc(cnt));
	}
	base = __shfl_sync(0xffffffffu, base, 0);
	uint32_t at = base + incl - n_emit;
	unsigned long long span = 0;
	if (n_emit) {
		IdxQuery qy;
		qy.pid = p;
		qy.seg = o | (k << 8) | (si << 16) | (wo << 24);
		for (uint32_t combo = 0; combo < n_combo; ++combo) {
			uint32_t seg_code = 0u, rest = combo;
			for (uint32_t q = 0; q < k; ++q) {
				uint32_t l;
				if (q == sub_at) {
					l = sub_letter;
				} else {
					const uint32_t set = idx_set(seg_m, o + q), c = (uint32_t)__popc(set);
					l = idx_nth_letter(set, rest % c);
					rest /= c;
				}
				seg_code = (seg_code << 2) | l;
			}
			if (!ext) { // the k-prefix as a range of 4^x buckets
				const uint32_t sh = 2u * x;
				qy.lo = __ldg(off + (seg_code << sh));
				qy.hi = __ldg(off + ((seg_code + 1u) << sh));
				span += qy.hi - qy.lo;
				if (at < q_cap) queries[at] = qy;
				++at;
			} else if (x <= 3u && ext_left) {
				for (uint64_t rem = okm; rem; rem &= rem - 1ull) {
					const uint32_t code = ((uint32_t)(__ffsll((long long)rem) - 1) << (2u * k)) | seg_code;
					qy.lo = __ldg(off + code);
					qy.hi = __ldg(off + code + 1u);
					span += qy.hi - qy.lo;
					if (at < q_cap) queries[at] = qy;
					++at;
				}
			} else if (x <= 3u) {
				const uint32_t code0 = seg_code << (2u * x);
				for (uint64_t starts = okm & ~(okm << 1); starts; starts &= starts - 1ull) {
					const uint32_t s0 = (uint32_t)(__ffsll((long long)starts) - 1);
					const uint32_t len = (uint32_t)(__ffsll((long long)~(okm >> s0)) - 1); // the run's allowed codes (bit 63 set: >> leaves zeros above)
					qy.lo = __ldg(off + code0 + s0);
					qy.hi = __ldg(off + code0 + s0 + len);
					span += qy.hi - qy.lo;
					if (at < q_cap) queries[at] = qy;
					++at;
				}
			} else if (ext_left) {
				for (uint32_t c = 0; c < (1u << (2u * x)); ++c) {
					const uint32_t d = c ^ ext_pat, bad = (d | (d >> 1)) & 0x55u;
					if ((uint32_t)__popc(bad) > budget) continue;
					const uint32_t code = (c << (2u * k)) | seg_code;
					qy.lo = __ldg(off + code);
					qy.hi = __ldg(off + code + 1u);
					span += qy.hi - qy.lo;
					if (at < q_cap) queries[at] = qy;
					++at;
				}
			} else {
				const uint32_t n_codes = 1u << (2u * x), code0 = seg_code << (2u * x);
				uint32_t run_start = 0;
				bool prev = false;
				for (uint32_t c = 0; c <= n_codes; ++c) { // one step past the end closes the last run
					const uint32_t d = c ^ ext_pat;
					const bool okc = c < n_codes && (uint32_t)__popc((d | (d >> 1)) & 0x55u) <= budget;
					if (okc && !prev) run_start = c;
					if (!okc && prev) {
						qy.lo = __ldg(off + code0 + run_start);
						qy.hi = __ldg(off + code0 + c);
						span += qy.hi - qy.lo;
						if (at < q_cap) queries[at] = qy;
						++at;
					}
					prev = okc;
				}
			}
		}
	}
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1) span += __shfl_xor_sync(0xffffffffu, span, d);
	if (lane == 0u && span) atomicAdd(n_entries, span);
}

// A verified candidate.  Resolving it (which sequence, is that one active, is the alignment somebody else's, which family
// members see it) costs a chain of dependent loads, and ~2 % of the streamed entries are candidates -- done inline it
// stalls the streaming warps.  So the scan only appends 16 bytes per candidate and index_hits_kernel resolves them
// afterwards, one thread each.
struct IdxCand {
	uint32_t gpos, m, pid, seg;
};

struct IdxCandSink {
	IdxCand *buf;
	unsigned int *count; // total produced (may exceed cap: the host grows the buffer and re-runs)
	uint32_t cap;
};

constexpr int IDX_THREADS = 256;
#ifndef IDX_BLOCKS
#define IDX_BLOCKS 4
#endif
constexpr int IDX_BLOCKS_PER_SM = IDX_BLOCKS; // one resident wave of persistent warps, eight 16-byte loads in flight per lane

__device__ __forceinline__ uint32_t index_mask(const uint4 &en, const uint4 &B, uint32_t sh)
{
	const uint64_t c0 = ((uint64_t)(en.w & 0xFFFFu) << 32) | en.y, c1 = ((uint64_t)(en.w >> 16) << 32) | en.z;
	const uint32_t t0 = (uint32_t)(c0 >> sh), t1 = (uint32_t)(c1 >> sh);
	return (B.x & ~t1 & ~t0) | (B.y & ~t1 & t0) | (B.z & t1 & ~t0) | (B.w & t1 & t0);
}

// Append the candidates among N entries per lane.  Candidates are sparse (~2 % of the entries) but nearly every batch of 256
// entries holds one, so even ONE atomic per warp per batch is 6 x 10^5 read-modify-writes of a single counter per launch, and
// those serialise in L2 at ~1 ns each -- that, not HBM, was the kernel's limit (0.64 ms).  A warp therefore reserves
// IDX_BLOCK slots at a time and fills them over many batches; what it leaves unused is marked invalid (gpos = ~0) for
// index_hits_kernel to skip.  `cs.count` counts reserved slots.  All 32 lanes must call this together.
constexpr uint32_t IDX_BLOCK = 512;
constexpr uint32_t IDX_INVALID = 0xFFFFFFFFu;

struct IdxWarpBlock { // uniform over the warp
	uint32_t base = 0, used = IDX_BLOCK; // used == IDX_BLOCK: nothing reserved yet
};

__device__ __forceinline__ void index_block_pad(const IdxCandSink &cs, const IdxWarpBlock &wb, uint32_t lane)
{ // mark [base + used, base + IDX_BLOCK) invalid
	for (uint32_t k = wb.used + lane; k < IDX_BLOCK; k += 32u)
		if (wb.base + k < cs.cap) cs.buf[wb.base + k].gpos = IDX_INVALID;
}

template <int N>
__device__ __forceinline__ void index_append(const uint4 (&e)[N], uint32_t valid, const uint4 &B, uint32_t thr, uint32_t sh, const IdxQuery &qy,
	const IdxCandSink &cs, uint32_t lane, IdxWarpBlock &wb)
{
	uint32_t hits = 0u; // bit u: entry u of this lane is a candidate
	#pragma unroll
	for (int u = 0; u < N; ++u)
		if (((valid >> u) & 1u) && (uint32_t)__popc(index_mask(e[u], B, sh)) >= thr) hits |= 1u << u;
	if (!__any_sync(0xffffffffu, hits != 0u)) return;
	const uint32_t mine = (uint32_t)__popc(hits);
	uint32_t incl = mine; // inclusive prefix sum over the lanes
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
		if ((int)lane >= o) incl += v;
	}
	const uint32_t total = __shfl_sync(0xffffffffu, incl, 31); // <= 32 N <= 256 < IDX_BLOCK
	if (wb.used + total > IDX_BLOCK) {
		if (wb.used < IDX_BLOCK) index_block_pad(cs, wb, lane);
		unsigned int b = 0;
		if (lane == 0u) b = atomicAdd(cs.count, IDX_BLOCK);
		wb.base = __shfl_sync(0xffffffffu, b, 0);
		wb.used = 0u;
	}
	uint32_t at = wb.base + wb.used + (incl - mine);
	wb.used += total;
	#pragma unroll
	for (int u = 0; u < N; ++u)
		if ((hits >> u) & 1u) {
			if (at < cs.cap) {
				IdxCand c;
				c.gpos = e[u].x;
				c.m = index_mask(e[u], B, sh);
				c.pid = qy.pid;
				c.seg = qy.seg;
				cs.buf[at] = c;
			}
			++at;
		}
}

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{ // read-once stream: do not keep it in L1
	uint4 v;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

// one warp per query; lanes stride over the entry range (coalesced 16-byte loads, IDX_ROWS in flight per lane), the pattern
// sits in registers.  Measured alternatives, all slower than this loop's 0.31 ms on the bench batch (6 x 10^7 entries in 9.5 x 10^5
// ranges): issuing the next range's loads before verifying the current chunk 0.42 ms; 8 or 16 lanes per range with 4 / 2 ranges
// side by side in a warp 0.71 / 0.65 ms (per-lane pattern registers push the kernel into spills at 64 registers); 5, 6 or 8
// resident CTAs per SM 0.40 / 0.54 / 0.83 ms (spills again).  What is left is latency: ~60 entries per range, one range at a time
// per warp, 32 warps per SM -- 3.1 TB/s of index stream, 0.47 of the HBM peak.
#ifndef IDX_ROWS
#define IDX_ROWS 8
#endif
__global__ void __launch_bounds__(IDX_THREADS, IDX_BLOCKS_PER_SM)
scan_index_kernel(TextIndex ix, const IdxQuery *__restrict__ queries, const unsigned int *__restrict__ n_queries, uint32_t q_cap,
	const uint4 *__restrict__ mask, const uint32_t *__restrict__ meta, IdxCandSink cs)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
	const uint32_t nq = min(*n_queries, q_cap); // more queries than slots: the host grows the buffer and re-runs
	if (warp >= nq) return;
	IdxQuery qy = queries[warp];
	IdxWarpBlock wb;
	for (uint32_t q = warp; q < nq; q += n_warps) {
		const IdxQuery cur = qy;
		if (q + n_warps < nq) qy = queries[q + n_warps]; // next descriptor in flight while this range streams
		const uint4 B = __ldg(mask + cur.pid);
		const uint32_t thr = __ldg(meta + cur.pid) & 63u;
		const uint32_t sh = IDX_CTX_BEFORE - (cur.seg >> 24); // context bit of primer base 0 (window offset <= 16)
		// whole rows of 32 entries, IDX_ROWS rows (then up to four) at a time; lanes past the end of the range carry no entry
		uint32_t i = cur.lo;
#if IDX_ROWS > 4
		for (; i + 32u * IDX_ROWS <= cur.hi; i += 32u * IDX_ROWS) {
			uint4 e[IDX_ROWS];
			#pragma unroll
			for (int u = 0; u < IDX_ROWS; ++u) e[u] = ldg_stream(ix.entries + i + 32u * u + lane);
			index_append<IDX_ROWS>(e, (1u << IDX_ROWS) - 1u, B, thr, sh, cur, cs, lane, wb);
		}
#endif
		for (; i < cur.hi; i += 128u) {
			uint4 e[4];
			uint32_t valid = 0u;
			#pragma unroll
			for (int u = 0; u < 4; ++u) {
				const uint32_t j = i + 32u * u + lane;
				e[u] = make_uint4(0, 0, 0, 0);
				if (j < cur.hi) {
					e[u] = ldg_stream(ix.entries + j);
					valid |= 1u << u;
				}
			}
			index_append<4>(e, valid, B, thr, sh, cur, cs, lane, wb);
		}
	}
	if (wb.used < IDX_BLOCK) index_block_pad(cs, wb, lane);
}

// The same scan with the entry loads taken out of the registers (round 2; option "use_async_scan" = 1, NOT the default: measured
// slower).  ncu's source page of the kernel above puts 49 % of its stall samples on the first use of the loaded entries: a warp has
// one range (~75 entries, three 16-byte loads per lane) in flight, waits a DRAM round trip for it, verifies, and only then asks for
// the next; keeping a second range in flight in registers costs 16 more of them and spills (measured, above).  Here every warp owns
// a ring of IDXA_STAGES chunks of 128 entries in SHARED memory and fills it with cp.async (16 bytes per lane, no destination
// register): while chunk k is verified, chunks k+1 .. k+STAGES-1 -- the next ranges, or the next 128 entries of a long one -- are
// already on their way, together with their pattern mask and threshold.  Result on the bench batch (64 registers, no spills, 4
// resident CTAs per SM): 0.371 / 0.367 / 0.439 ms with 2 / 3 / 4 stages against 0.293 ms for the register version -- two or three
// times the bytes in flight per warp buy nothing, so the limit is not the latency a warp sees but what the memory system delivers
// for this access pattern: 7.9 x 10^5 independent ~1.2 KB reads scattered over 9.6 GB (every range opens its own DRAM page), 3.3 TB/s
// = 0.5 of the streaming peak.  Kept for the record and as a cross-check of the scan (tests/test_gpu_parity.py).
#ifndef IDXA_STAGES
#define IDXA_STAGES 3
#endif
constexpr int IDXA_CHUNK = 128;                                   // entries per chunk: four rows of 32
constexpr int IDXA_WARPS = IDX_THREADS / 32;
struct IdxaMeta {                                                 // per (warp, stage), written by lane 0 / cp.async at issue time
	uint4 mask;                                                   // the pattern's letter planes (cp.async)
	uint32_t thr_word;                                            // meta[pid] (cp.async)
	uint32_t pid, seg, n;                                         // n = entries in the chunk, 0 = the warp has run out of work
};
constexpr size_t IDXA_SMEM = (size_t)IDXA_WARPS * IDXA_STAGES * (IDXA_CHUNK * sizeof(uint4) + sizeof(IdxaMeta));

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
	const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
	const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(IDX_THREADS, IDX_BLOCKS_PER_SM)
scan_index_async_kernel(TextIndex ix, const IdxQuery *__restrict__ queries, const unsigned int *__restrict__ n_queries, uint32_t q_cap,
	const uint4 *__restrict__ mask, const uint32_t *__restrict__ meta, IdxCandSink cs)
{
	extern __shared__ uint4 s_idxa[];
	const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
	uint4 *ring = s_idxa + (size_t)wib * IDXA_STAGES * IDXA_CHUNK;
	IdxaMeta *smeta = reinterpret_cast<IdxaMeta *>(s_idxa + (size_t)IDXA_WARPS * IDXA_STAGES * IDXA_CHUNK) + wib * IDXA_STAGES;
	const uint32_t nq = min(*n_queries, q_cap); // more queries than slots: the host grows the buffer and re-runs
	// producer: the range being cut into chunks, and the next descriptor already loaded
	uint32_t q = warp;
	IdxQuery cur, nxt;
	cur.lo = cur.hi = cur.pid = cur.seg = 0u;
	nxt = cur;
	bool have_cur = q < nq, have_nxt = false;
	if (have_cur) cur = queries[q];
	if (q + n_warps < nq) { nxt = queries[q + n_warps]; have_nxt = true; }
	uint32_t off = 0u;
	auto issue = [&](uint32_t stage) {
		while (have_cur && cur.lo + off >= cur.hi) { // this range is finished (or empty): on to the next one
			cur = nxt;
			have_cur = have_nxt;
			off = 0u;
			q += n_warps;
			have_nxt = q + n_warps < nq;
			if (have_nxt) nxt = queries[q + n_warps];
		}
		uint32_t n = 0u;
		if (have_cur) {
			const uint32_t base = cur.lo + off;
			n = min((uint32_t)IDXA_CHUNK, cur.hi - base);
			uint4 *dst = ring + (size_t)stage * IDXA_CHUNK;
			#pragma unroll
			for (int u = 0; u < IDXA_CHUNK / 32; ++u) {
				const uint32_t j = 32u * u + lane;
				if (j < n) cp_async16(dst + j, ix.entries + base + j);
			}
			if (lane == 0u) {
				cp_async16(&smeta[stage].mask, mask + cur.pid);
				cp_async4(&smeta[stage].thr_word, meta + cur.pid);
				smeta[stage].pid = cur.pid;
				smeta[stage].seg = cur.seg;
			}
			off += n;
		}
		if (lane == 0u) smeta[stage].n = n;
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
	#pragma unroll
	for (int s = 0; s < IDXA_STAGES; ++s) issue((uint32_t)s);
	IdxWarpBlock wb;
	uint32_t head = 0u;
	for (;;) {
		asm volatile("cp.async.wait_group %0;" ::"n"(IDXA_STAGES - 1) : "memory");
		__syncwarp();
		const IdxaMeta m = smeta[head];
		if (m.n == 0u) break; // chunks are issued in order: an empty one means nothing follows
		const uint32_t thr = m.thr_word & 63u;
		const uint32_t sh = IDX_CTX_BEFORE - (m.seg >> 24); // context bit of primer base 0 (window offset <= 16)
		IdxQuery qy;
		qy.lo = qy.hi = 0u;
		qy.pid = m.pid;
		qy.seg = m.seg;
		const uint4 *src = ring + (size_t)head * IDXA_CHUNK;
		uint4 e[IDXA_CHUNK / 32];
		uint32_t valid = 0u;
		#pragma unroll
		for (int u = 0; u < IDXA_CHUNK / 32; ++u) {
			const uint32_t j = 32u * u + lane;
			e[u] = make_uint4(0, 0, 0, 0);
			if (j < m.n) {
				e[u] = src[j];
				valid |= 1u << u;
			}
		}
		index_append<IDXA_CHUNK / 32>(e, valid, m.mask, thr, sh, qy, cs, lane, wb);
		__syncwarp(); // every lane has read this stage before it is refilled
		issue(head);
		head = head + 1u == (uint32_t)IDXA_STAGES ? 0u : head + 1u;
	}
	asm volatile("cp.async.wait_group 0;" ::: "memory");
	if (wb.used < IDX_BLOCK) index_block_pad(cs, wb, lane);
}

// place each candidate in its sequence, drop what other kernels own, report once.
// A candidate is a chain of dependent loads (record -> block table -> sequence starts -> the sequence's flags and length) with a few
// instructions in between: ncu had 13 warps per issue waiting on them.  The kernel can walk IDX_HITS_ROWS candidates per thread side
// by side (the loads of a step issued for all of them before any is used); measured on the bench batch, 4 batches in flight / one at
// a time: 1 row 0.850 / 1.145 ms per step, 2 rows 0.852 / 1.141, 4 rows 0.857 / 1.137 -- the latency a lone batch gains is lost again
// when other batches' kernels fill the stalls anyway, so the default stays 1.  The lanes of a warp meet again before the hit list is
// touched, so a row of 32 candidates costs one atomic on the list counter.
#ifndef IDX_HITS_BLOCKS
#define IDX_HITS_BLOCKS 1
#endif
#ifndef IDX_HITS_ROWS
#define IDX_HITS_ROWS 1
#endif
__global__ void __launch_bounds__(256, IDX_HITS_BLOCKS)
index_hits_kernel(SeqDev sd, TextIndex ix, IdxCandSink cs, const uint32_t *__restrict__ g_meta, const uint32_t *__restrict__ g_meta2,
	const uint32_t *__restrict__ dirty_bits, const uint8_t *__restrict__ stale, uint32_t cand_bits, HitSink hs)
{
	constexpr int R = IDX_HITS_ROWS;
	const uint32_t total = min(*cs.count, cs.cap);
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t r0 = warp * R; r0 * 32u < total; r0 += n_warps * R) { // rows r0 .. r0 + R - 1 of 32 candidates: warp-uniform trip count
		IdxCand c[R];
		bool live[R];
		uint32_t meta[R], meta2[R], lo[R], hi[R];
		#pragma unroll
		for (int u = 0; u < R; ++u) {
			const uint32_t i = (r0 + u) * 32u + lane;
			live[u] = i < total;
			c[u].gpos = IDX_INVALID;
			c[u].m = c[u].pid = c[u].seg = 0u;
			if (live[u]) c[u] = cs.buf[i];
			live[u] = live[u] && c[u].gpos != IDX_INVALID; // (else: the unused tail of a warp's block)
		}
		#pragma unroll
		for (int u = 0; u < R; ++u) { // independent of the chain that places the candidate: issued beside it
			meta[u] = meta2[u] = lo[u] = 0u;
			hi[u] = 1u;
			if (live[u]) {
				meta[u] = __ldg(g_meta + c[u].pid);
				meta2[u] = __ldg(g_meta2 + c[u].pid);
				const uint32_t b = c[u].gpos >> IDX_BLK_SHIFT; // idx_seq_of_fast, its probes taken in step with the other rows'
				lo[u] = __ldg(ix.blk + b);
				hi[u] = min(__ldg(ix.blk + b + 1u) + 1u, ix.n_seq);
			}
		}
		for (;;) {
			bool any = false;
			uint32_t mid[R], at[R];
			#pragma unroll
			for (int u = 0; u < R; ++u) {
				mid[u] = (lo[u] + hi[u]) >> 1;
				at[u] = (live[u] && hi[u] - lo[u] > 1u) ? __ldg(ix.cum + mid[u]) : 0u;
			}
			#pragma unroll
			for (int u = 0; u < R; ++u)
				if (live[u] && hi[u] - lo[u] > 1u) {
					if (at[u] <= c[u].gpos) lo[u] = mid[u]; else hi[u] = mid[u];
					any = any || hi[u] - lo[u] > 1u;
				}
			if (!any) break;
		}
		uint32_t seq[R], first[R], clen[R];
		bool ok[R];
		#pragma unroll
		for (int u = 0; u < R; ++u) {
			seq[u] = ix.seq_lo + lo[u];
			first[u] = clen[u] = 0u;
			ok[u] = false;
			if (live[u]) {
				first[u] = __ldg(ix.cum + lo[u]);
				clen[u] = sd.clen[seq[u]];
				// a sequence split since the build has moved under its entries: the table scan covers it (pcramp_gpu.cu)
				ok[u] = sd.active[seq[u]] && !(stale && stale[seq[u]]);
			}
		}
		#pragma unroll
		for (int u = 0; u < R; ++u) {
			const uint32_t wo = c[u].seg >> 24, si = (c[u].seg >> 16) & 255u;
			const int64_t x = (int64_t)(c[u].gpos - first[u]) - (int64_t)wo; // text index of primer base 0
			bool good = ok[u] && x >= 0;
			if (good && dirty_bits) { // alignments touching a degenerate text base belong to scan_groups_kernel
				const uint64_t G = sd.grp_off[seq[u]] + (uint64_t)(x >> 5);
				if ((dirty_bits[G >> 5] >> (G & 31u)) & 1u) good = false;
			}
			uint32_t cnt = 0;
			if (good) {
				const uint32_t n = (meta2[u] >> 10) & 63u, segs = idx_segments((meta2[u] >> 16) & 63u);
				for (uint32_t k = 0; k < si; ++k) { // an earlier segment whose prefix is within one mismatch reports this alignment
					uint32_t oo, kk;
					idx_segment(n, segs, k, oo, kk);
					if ((uint32_t)__popc((c[u].m >> oo) & ((1u << kk) - 1u)) + 1u >= kk) good = false;
				}
				cnt = (uint32_t)__popc(c[u].m);
			}
			const bool family = good && (meta2[u] & 1023u) != 0u; // 5'/3' shift families fan out to several candidates: the general path
			if (family) emit_family(hs, seq[u], clen[u], cand_bits, meta[u], meta2[u], x, cnt);
			// the plain case (no shift family): at most one hit per candidate, appended with one atomic per warp
			bool one = good && !family;
			uint32_t wstart = 0;
			if (one) {
				const int64_t ws = x - (int64_t)((meta[u] >> 6) & 31u);
				one = ws >= 0 && ws + 32 <= (int64_t)clen[u];
				wstart = (uint32_t)ws;
			}
			const uint32_t bal = __ballot_sync(0xffffffffu, one);
			if (bal) {
				unsigned long long base = 0;
				if (lane == 0u) base = atomicAdd(hs.count, (unsigned long long)__popc(bal));
				base = __shfl_sync(0xffffffffu, base, 0);
				if (one) {
					const unsigned long long at = base + (unsigned long long)__popc(bal & ((1u << lane) - 1u));
					if (at < hs.cap) {
						hs.key[at] = hit_key_pack(seq[u], meta[u] >> 12, cand_bits, cnt, ENT_FULL, (meta[u] >> 11) & 1u);
						hs.val[at] = wstart + 31u;
					}
				}
			}
		}
	}
}

} // namespace pcr
