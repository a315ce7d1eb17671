// fst.cuh -- frame-aligned seed table: "which of these N oligos reach their threshold on THIS 32-slot word?" without
// comparing the word with every oligo.
//
// Word::operator& compares two words slot by slot in the same 32-slot frame (word.cpp:111-154) -- there is no sliding.
// An oligo of n bases that must reach `thr` matching slots may miss e = n - thr of them, so of e + 1 disjoint pieces of
// the oligo one matches in every slot (pigeonhole).  The table holds, for every oligo piece, its first q <= 6 bases
// (every IUPAC expansion) under the key (frame position, q, 2-bit code); a word looks up the codes it carries at the
// (position, q) combinations in use and verifies only the oligos it finds there.  Used by pair scoring (match_words at
// thr^2, optimize.cpp:291-301: the database words against the trial oligos) and by the partial-word scan (the words
// pack() emits at sequence ends against the candidate words, select_words.cpp:103-117).  Oligos whose pieces are shorter
// than 3 bases or expand too far are listed separately and compared with every word; words holding a degenerate base
// are compared with every oligo.
#pragma once
#include "word128.cuh"

#include <cuda_runtime.h>

namespace pcr {

constexpr uint32_t FST_QMIN = 3u, FST_QMAX = 6u;
constexpr uint32_t FST_COMBOS = 32u * 4u;                 // (frame position, q - 3)
constexpr uint32_t FST_BUCKETS = FST_COMBOS << 12;        // x 12-bit code (two 6-bit planes)
constexpr uint32_t FST_MAX_EXPANSIONS = 16u;              // per piece

struct Fst {
	const uint4 *planes;       // per oligo: frame letter planes A, C, G, T
	const uint32_t *thr;       // per oligo: slots that must match
	const uint32_t *start;     // FST_BUCKETS + 1
	const uint32_t *ids;       // oligo ids grouped by bucket
	const uint32_t *combo;     // FST_COMBOS flags: is any seed stored under (position, q)?  then (fst_combo_list_kernel)
	                           // combo[FST_COMBOS] = number of combinations in use, combo[FST_COMBOS + 1 ...] = their list
	const uint32_t *brute;     // oligos that are compared with every word
	const uint32_t *n_brute;
	uint32_t n;
};

__device__ __forceinline__ void fst_shape(const uint4 &p, uint32_t thr, uint32_t &first, uint32_t &n, uint32_t &pieces, bool &usable)
{
	const uint32_t occ = p.x | p.y | p.z | p.w;
	first = occ ? (uint32_t)__ffs(occ) - 1u : 0u;
	n = (uint32_t)__popc(occ);
	const bool contiguous = occ != 0u && ((occ >> first) == ((n >= 32u) ? 0xFFFFFFFFu : ((1u << n) - 1u)));
	usable = contiguous && thr >= 1u && thr <= n;
	pieces = usable ? (n - thr) + 1u : 1u;
}
__device__ __forceinline__ void fst_piece(uint32_t n, uint32_t pieces, uint32_t i, uint32_t &o, uint32_t &q)
{
	o = (i * n) / pieces;
	const uint32_t len = ((i + 1u) * n) / pieces - o;
	q = len < FST_QMAX ? len : FST_QMAX;
}

// every expansion of the q bases at frame position p0 of an oligo, as bucket indices; false if it cannot be seeded
template <class F>
__device__ __forceinline__ bool fst_piece_seeds(const uint4 &p, uint32_t p0, uint32_t q, F f)
{
	if (q < FST_QMIN) return false;
	uint32_t total = 1u;
	for (uint32_t j = 0; j < q; ++j) {
		const uint32_t k = ((p.x >> (p0 + j)) & 1u) + ((p.y >> (p0 + j)) & 1u) + ((p.z >> (p0 + j)) & 1u) + ((p.w >> (p0 + j)) & 1u);
		total *= k;
		if (k == 0u || total > FST_MAX_EXPANSIONS) return false;
	}
	for (uint32_t x = 0; x < total; ++x) {
		uint32_t rest = x, b0 = 0u, b1 = 0u;
		for (uint32_t j = 0; j < q; ++j) {
			const uint32_t s = ((p.x >> (p0 + j)) & 1u) | (((p.y >> (p0 + j)) & 1u) << 1) | (((p.z >> (p0 + j)) & 1u) << 2) | (((p.w >> (p0 + j)) & 1u) << 3);
			const uint32_t k = (uint32_t)__popc(s);
			uint32_t pick = rest % k, letter = 0u;
			rest /= k;
			for (uint32_t l = 0; l < 4u; ++l)
				if ((s >> l) & 1u) {
					if (pick == 0u) { letter = l; break; }
					--pick;
				}
			b0 |= (letter & 1u) << j;          // A=0 C=1 G=2 T=3: low code bit = C|T, high = G|T
			b1 |= (letter >> 1) << j;
		}
		f((((p0 << 2) | (q - FST_QMIN)) << 12) | b0 | (b1 << 6));
	}
	return true;
}

// pass 0 (ids == nullptr): bucket counts, combo flags, brute list; pass 1: fill ids (cursor = bucket starts, advanced)
// ids_cap / overflow: the fill pass of a table whose size the host has not read (capacity from the previous batch)
__global__ void fst_build_kernel(const uint4 *__restrict__ planes, const uint32_t *__restrict__ thr, uint32_t n, uint32_t *bucket, uint32_t *combo,
	uint32_t *brute, uint32_t *n_brute, uint32_t *ids, uint32_t ids_cap, unsigned int *overflow)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint4 p = planes[i];
	uint32_t first, len, pieces;
	bool usable;
	fst_shape(p, thr[i], first, len, pieces, usable);
	bool ok = usable;
	if (ok) { // dry run: every piece must be seedable
		for (uint32_t k = 0; k < pieces && ok; ++k) {
			uint32_t o, q;
			fst_piece(len, pieces, k, o, q);
			ok = fst_piece_seeds(p, first + o, q, [](uint32_t) {});
		}
	}
	if (!ok) {
		if (!ids && (p.x | p.y | p.z | p.w) != 0u && thr[i] <= len) brute[atomicAdd(n_brute, 1u)] = i; // thr > n can never match
		return;
	}
	for (uint32_t k = 0; k < pieces; ++k) {
		uint32_t o, q;
		fst_piece(len, pieces, k, o, q);
		fst_piece_seeds(p, first + o, q, [&](uint32_t b) {
			const uint32_t slot = atomicAdd(bucket + b, 1u);
			if (ids) {
				if (slot < ids_cap) ids[slot] = i;
				else if (overflow) atomicOr(overflow, 4u);
			} else combo[b >> 12] = 1u;
		});
	}
}

// one warp: compact the combo flags into a list behind them
__global__ void fst_combo_list_kernel(uint32_t *combo)
{
	const uint32_t lane = threadIdx.x & 31u;
	uint32_t n = 0;
	for (uint32_t base = 0; base < FST_COMBOS; base += 32u) {
		const bool used = combo[base + lane] != 0u;
		const uint32_t m = __ballot_sync(0xffffffffu, used);
		if (used) combo[FST_COMBOS + 1u + n + (uint32_t)__popc(m & ((1u << lane) - 1u))] = base + lane;
		n += (uint32_t)__popc(m);
	}
	if (lane == 0u) combo[FST_COMBOS] = n;
}

// letter planes of a word -> (single-letter slots, code planes)
struct FstWord {
	uint32_t a, c, g, t, single, b0, b1;
	bool degenerate; // some occupied slot holds more than one letter: the seeds cannot see it
};
__device__ __forceinline__ FstWord fst_word(uint32_t a, uint32_t c, uint32_t g, uint32_t t)
{
	FstWord w;
	w.a = a; w.c = c; w.g = g; w.t = t;
	const uint32_t multi = (a & c) | (a & g) | (a & t) | (c & g) | (c & t) | (g & t);
	w.single = (a | c | g | t) & ~multi;
	w.degenerate = multi != 0u;
	w.b0 = (c | t) & w.single;
	w.b1 = (g | t) & w.single;
	return w;
}

// call hit(oligo id, match mask) for every oligo that reaches its threshold on the word; with `once` each (word, oligo)
// pair is reported through its leftmost fully matching seed only (callers whose sink is not idempotent need that)
template <bool ONCE, class F>
__device__ __forceinline__ void fst_match(const Fst &t, const FstWord &w, F hit)
{
	auto verify = [&](uint32_t id, uint32_t p0, bool seeded) {
		const uint4 p = __ldg(t.planes + id);
		const uint32_t m = (p.x & w.a) | (p.y & w.c) | (p.z & w.g) | (p.w & w.t);
		const uint32_t need = __ldg(t.thr + id);
		if ((uint32_t)__popc(m) < need) return;
		if (ONCE && seeded) {
			uint32_t first, len, pieces;
			bool usable;
			fst_shape(p, need, first, len, pieces, usable);
			for (uint32_t k = 0; k < pieces; ++k) {
				uint32_t o, q;
				fst_piece(len, pieces, k, o, q);
				if (first + o >= p0) break;
				const uint32_t qm = (1u << q) - 1u;
				if (((m >> (first + o)) & qm) == qm) return; // an earlier seed of this oligo finds the word too
			}
		}
		hit(id, m);
	};
	if (w.degenerate) { // IUPAC text: no seed code for it; compare with everybody
		for (uint32_t id = 0; id < t.n; ++id) verify(id, 0u, false);
		return;
	}
	const uint32_t nb = __ldg(t.n_brute);
	for (uint32_t k = 0; k < nb; ++k) verify(__ldg(t.brute + k), 0u, false);
	const uint32_t n_combo = __ldg(t.combo + FST_COMBOS);
	for (uint32_t ci = 0; ci < n_combo; ++ci) {
		const uint32_t cb = __ldg(t.combo + FST_COMBOS + 1u + ci);
		const uint32_t p0 = cb >> 2, q = (cb & 3u) + FST_QMIN, qm = (1u << q) - 1u;
		if (p0 + q > 32u || ((w.single >> p0) & qm) != qm) continue;
		const uint32_t b = (cb << 12) | ((w.b0 >> p0) & qm) | (((w.b1 >> p0) & qm) << 6);
		const uint32_t lo = __ldg(t.start + b), hi = __ldg(t.start + b + 1u);
		for (uint32_t e = lo; e < hi; ++e) verify(__ldg(t.ids + e), p0, true);
	}
}

} // namespace pcr
