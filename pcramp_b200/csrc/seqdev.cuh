// seqdev.cuh -- how a sequence collection lives in HBM, and the closed-form model of
// Sequence::pack (sequence.cpp:92-267) that every kernel shares.
//
// HBM layout per collection (all arrays are struct-of-arrays indexed by sequence):
//   raw      the reference's packed nibbles as uploaded (2 bases/byte, sequence.h:223-228); only
//            the GC filter and the slow paths read it.
//   planes   the COMPRESSED text (EOS nibbles removed -- pack() absorbs them, Appendix A.3 of
//            SURVEY.md) as four interleaved bit-planes: one uint4 {A,C,G,T} per 32 bases, bit b of
//            a plane word = base 32*g+b carries that letter.  Same 4 bits/base as the nibbles, but
//            a 32-base template window for ANY alignment is two 16-byte loads and four funnel
//            shifts, and "does primer position k match" is a plain AND.  Each sequence owns
//            ceil(clen/32)+1 groups (one zero halo group), bits past clen are zero.
//   eos_pos  sorted raw positions of the EOS nibbles of each sequence (CSR via eos_off); empty for
//            almost every sequence.  Gives raw<->compressed index maps and has_split() by search.
//
// Closed form of pack().  Let b_0..b_{clen-1} be the non-EOS bases, c_i the number of bases in
// raw[0..i].  Entries are emitted at these events (both strands each), subject to the degeneracy
// and GC filters:
//   FULL    every compressed index j >= 31: word = b_{j-31..j}; with i = raw index of b_j:
//           loc+ = i - 31, loc- = i                                   (sequence.cpp:182-194)
//   FILL    every raw index i with 1 <= c_i < 32 and c_i >= min_len (EOS positions included, which
//           re-emit the unchanged word): word = b_0..b_{n-1}, n = c_i, centred at st = (33-n)/2;
//           loc+ = i + 1 - n - st, loc- = i + st                      (:155-181)
//   EOSEVT  every raw index i holding EOS with c_i >= 32 (31 >= min_len): word = the last 31 bases,
//           st = 1; loc+ = i - 31, loc- = i + 1                       (:155-181 with size 31)
//   TAIL    q = 1..s0: the last n = n0 - q bases, counter s = s0 - q >= min_len, st = (33-n)/2;
//           loc+ = L - s - st, loc- = L - 1 + st                      (:198-263)
//           where (n0, s0) = (clen, clen) if clen < 32, (31, 31) if the last raw nibble is EOS,
//           else (32, 31).
// Odd lengths: pack() iterates BYTES (`iter != seq_buffer.end()`, sequence.cpp:111), so a sequence of
// odd length also pushes the unused low nibble of its last byte, which Sequence::operator= left zero
// (sequence.cpp:23): one trailing EOS at raw index len.  plen = len + (len & 1) is the length the
// model above uses for L, and that virtual EOS is part of eos_pos (has_split never looks that far).
// tests/test_gpu_parity.py checks this model entry by entry against the reference's state machine.
#pragma once
#include "word128.cuh"

namespace pcr {

struct SeqDev {
	uint32_t n;
	const uint4 *planes;
	const uint64_t *grp_off; // n + 1, in groups
	const uint32_t *clen;    // compressed length (bases)
	const uint32_t *len;     // raw length (nibbles incl. EOS)
	const uint32_t *plen;    // length pack() walks: len rounded up to even (see below)
	const uint8_t *raw;
	const uint64_t *raw_off; // byte offset of each sequence in raw
	const uint32_t *eos_pos;
	const uint32_t *eos_off; // n + 1
	const float *weight;
	const uint8_t *active;
};

enum : uint32_t { ENT_FULL = 0, ENT_FILL = 1, ENT_EOSEVT = 2, ENT_TAIL = 3 };

struct PackParams {
	uint32_t max_degen;
	float min_gc, max_gc;
	uint32_t min_len;
	int gc_filter;
};

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t comp_nibble_at(const SeqDev &sd, uint32_t seq, uint32_t j)
{
	const uint4 g = __ldg(sd.planes + sd.grp_off[seq] + (j >> 5));
	const uint32_t b = j & 31u;
	return ((g.x >> b) & 1u) | (((g.y >> b) & 1u) << 1) | (((g.z >> b) & 1u) << 2) | (((g.w >> b) & 1u) << 3);
}

__device__ __forceinline__ uint32_t raw_nibble_at(const SeqDev &sd, uint32_t seq, uint32_t i)
{
	if (i >= sd.len[seq]) return 0u; // the virtual pad nibble of an odd-length sequence
	const uint8_t v = __ldg(sd.raw + sd.raw_off[seq] + (i >> 1));
	return (i & 1u) ? (v & 0xFu) : (v >> 4);
}

// number of EOS k (0-based, raw position e_k) with e_k - k <= j, i.e. that precede base j
__device__ __forceinline__ uint32_t eos_before_base(const SeqDev &sd, uint32_t seq, uint32_t j)
{
	const uint32_t lo0 = sd.eos_off[seq], hi0 = sd.eos_off[seq + 1];
	uint32_t lo = 0, hi = hi0 - lo0;
	while (lo < hi) {
		const uint32_t mid = (lo + hi) >> 1;
		if (sd.eos_pos[lo0 + mid] - mid <= j) lo = mid + 1; else hi = mid;
	}
	return lo;
}
__device__ __forceinline__ uint32_t raw_of_comp(const SeqDev &sd, uint32_t seq, uint32_t j) { return j + eos_before_base(sd, seq, j); }

// number of EOS with raw position <= i
__device__ __forceinline__ uint32_t eos_upto_raw(const SeqDev &sd, uint32_t seq, int64_t i)
{
	const uint32_t lo0 = sd.eos_off[seq], hi0 = sd.eos_off[seq + 1];
	uint32_t lo = 0, hi = hi0 - lo0;
	while (lo < hi) {
		const uint32_t mid = (lo + hi) >> 1;
		if ((int64_t)sd.eos_pos[lo0 + mid] <= i) lo = mid + 1; else hi = mid;
	}
	return lo;
}

// Sequence::has_split (sequence.cpp:304-330) for an in-range [loc, loc+len)
__device__ __forceinline__ bool has_split_dev(const SeqDev &sd, uint32_t seq, int loc, int len)
{
	if (sd.eos_off[seq + 1] == sd.eos_off[seq] || len <= 0) return false;
	return eos_upto_raw(sd, seq, (int64_t)loc + len - 1) != eos_upto_raw(sd, seq, (int64_t)loc - 1);
}

// 16 plane bits -> 16 nibble slots of a 64-bit limb: bit j -> bit 4 j
__device__ __forceinline__ uint64_t spread_to_nibbles(uint32_t v)
{
	uint64_t x = v & 0xFFFFu;
	x = (x ^ (x << 24)) & 0x000000FF000000FFull;
	x = (x ^ (x << 12)) & 0x000F000F000F000Full;
	x = (x ^ (x << 6)) & 0x0303030303030303ull;
	x = (x ^ (x << 3)) & 0x1111111111111111ull;
	return x;
}

// n <= 32 bases b_{first..first+n-1} of the compressed text, left-justified in a word.  Two 16-byte plane loads and four funnel
// shifts give the window as four 32-bit letter planes (bit k = base first + k); word position k is nibble 15 - k of its limb
// (word128.cuh), so each 16-bit half is bit-reversed and spread to every fourth bit.  (One plane load per base -- 32 dependent
// extractions per entry -- made materialise_kernel 0.22 ms of the step.)
__device__ inline W128 gather_bases(const SeqDev &sd, uint32_t seq, uint32_t first, uint32_t n)
{
	const uint64_t gb = sd.grp_off[seq];
	const uint32_t ngrp = (uint32_t)(sd.grp_off[seq + 1] - gb), g = first >> 5, sh = first & 31u;
	const uint4 z = make_uint4(0, 0, 0, 0);
	const uint4 p0 = g < ngrp ? __ldg(sd.planes + gb + g) : z;
	const uint4 p1 = g + 1u < ngrp ? __ldg(sd.planes + gb + g + 1u) : z;
	const uint32_t keep = n >= 32u ? 0xFFFFFFFFu : ((1u << n) - 1u);
	const uint32_t a = __brev(__funnelshift_r(p0.x, p1.x, sh) & keep), c = __brev(__funnelshift_r(p0.y, p1.y, sh) & keep);
	const uint32_t gg = __brev(__funnelshift_r(p0.z, p1.z, sh) & keep), t = __brev(__funnelshift_r(p0.w, p1.w, sh) & keep);
	// after the reversal base k sits at bit 31 - k: bases 0..15 are the high half-word, in limb order already
	W128 w;
	w.hi = spread_to_nibbles(a >> 16) | (spread_to_nibbles(c >> 16) << 1) | (spread_to_nibbles(gg >> 16) << 2) | (spread_to_nibbles(t >> 16) << 3);
	w.lo = spread_to_nibbles(a) | (spread_to_nibbles(c) << 1) | (spread_to_nibbles(gg) << 2) | (spread_to_nibbles(t) << 3);
	return w;
}

// GC count over raw[i-31..i] clipped at 0 (sequence.cpp:127-141: the last 32 pushed nibbles, EOS included)
__device__ inline uint32_t gc_window_raw(const SeqDev &sd, uint32_t seq, int64_t i)
{
	uint32_t c = 0;
	for (int64_t p = (i >= 31 ? i - 31 : 0); p <= i; ++p) c += ((raw_nibble_at(sd, seq, (uint32_t)p) & 6u) != 0u);
	return c;
}
__device__ __forceinline__ bool gc_pass(uint32_t num_gc, const PackParams &pp)
{
	const float fraction = __fmul_rn((float)num_gc, 1.0f / 32.0f);
	return !(fraction < pp.min_gc || fraction > pp.max_gc);
}

// Geometry of one pack() event, identified by (type, pos): the n bases starting at compressed position `first`, the frame shift
// st of the word, the two WordMatch.loc values and where the G+C window ends.  false: the event emits nothing (too short).
__device__ inline bool pack_geom(const SeqDev &sd, uint32_t seq, uint32_t type, uint32_t pos, const PackParams &pp, uint32_t &n, uint32_t &first,
	int &st, int &loc_p, int &loc_m, int64_t &gc_end)
{
	const uint32_t L = sd.plen[seq], Lc = sd.clen[seq];
	gc_end = -1; // raw index the GC window ends at; -2 = tail rule
	if (type == ENT_FULL) {
		if (pos < 31u || pos >= Lc) return false;
		n = 32; first = pos - 31u; st = 0;
		const uint32_t i = raw_of_comp(sd, seq, pos);
		loc_p = (int)i - 31;
		loc_m = (int)i;
		gc_end = i;
	} else if (type == ENT_FILL) {
		if (pos >= L) return false;
		const uint32_t c = pos + 1u - eos_upto_raw(sd, seq, pos);
		if (c >= 32u || c == 0u || c < pp.min_len) return false;
		n = c; first = 0; st = (33 - (int)n) / 2;
		loc_p = (int)pos + 1 - (int)n - st;
		loc_m = (int)pos + st;
		gc_end = pos;
	} else if (type == ENT_EOSEVT) {
		if (pos >= L || raw_nibble_at(sd, seq, pos) != 0u) return false;
		const uint32_t c = pos + 1u - eos_upto_raw(sd, seq, pos);
		if (c < 32u || 31u < pp.min_len) return false;
		n = 31; first = c - 31u; st = 1;
		loc_p = (int)pos - 31;
		loc_m = (int)pos + 1;
		gc_end = pos;
	} else {
		uint32_t n0, s0;
		if (Lc < 32u) { n0 = Lc; s0 = Lc; }
		else if (raw_nibble_at(sd, seq, L - 1u) == 0u) { n0 = 31; s0 = 31; }
		else { n0 = 32; s0 = 31; }
		if (pos == 0u || pos > s0) return false;
		const uint32_t s = s0 - pos;
		n = n0 - pos;
		if (s < pp.min_len || n == 0u) return false;
		first = Lc - n; st = (33 - (int)n) / 2;
		loc_p = (int)L - (int)s - st;
		loc_m = (int)L - 1 + st;
		gc_end = -2;
	}
	return true;
}

// One pack() entry pair, identified by (type, pos).  Returns false when the event emits nothing
// (too short, filtered).  plus/minus are the two database words, loc_p/loc_m their WordMatch.loc.
__device__ inline bool pack_entry(const SeqDev &sd, uint32_t seq, uint32_t type, uint32_t pos, const PackParams &pp,
	W128 &plus, W128 &minus, int &loc_p, int &loc_m)
{
	const uint32_t L = sd.plen[seq];
	uint32_t n, first;
	int st;
	int64_t gc_end;
	if (!pack_geom(sd, seq, type, pos, pp, n, first, st, loc_p, loc_m, gc_end)) return false;
	if (pp.gc_filter) {
		uint32_t num_gc;
		if (gc_end == -2) { // tail: one pop when 32 nibbles were buffered, then frozen (sequence.cpp:204-222)
			num_gc = (L >= 32u) ? gc_window_raw(sd, seq, (int64_t)L - 1) - ((raw_nibble_at(sd, seq, L - 32u) & 6u) != 0u)
			                    : gc_window_raw(sd, seq, (int64_t)L - 1);
		} else {
			num_gc = gc_window_raw(sd, seq, gc_end);
		}
		if (!gc_pass(num_gc, pp)) return false;
	}
	const W128 left = gather_bases(sd, seq, first, n);
	if (w_degeneracy_sat(left) > (uint64_t)pp.max_degen) return false; // :149-153 (shift-invariant)
	plus = w_shr(left, st);
	minus = w_shr(w_complement(left), st); // complement() left-justifies; centring gives the same st
	return true;
}

// The letter planes (word128.cuh w_planes: bit k = word position k) and loc of ONE strand's word of an entry that is known to
// pass the filters -- what the database build needs.  The window's planes come straight off the collection's bit-planes (two
// 16-byte loads, four funnel shifts); the reverse complement is a bit reversal with A <-> T, C <-> G; no nibble form in between.
__device__ inline void pack_entry_planes(const SeqDev &sd, uint32_t seq, uint32_t type, uint32_t pos, const PackParams &pp, bool minus, Planes4 &out,
	int &loc)
{
	uint32_t n = 0, first = 0;
	int st = 0, lp = 0, lm = 0;
	int64_t gc_end;
	out.a = out.c = out.g = out.t = 0u;
	loc = 0;
	if (!pack_geom(sd, seq, type, pos, pp, n, first, st, lp, lm, gc_end)) return;
	const uint64_t gb = sd.grp_off[seq];
	const uint32_t ngrp = (uint32_t)(sd.grp_off[seq + 1] - gb), g = first >> 5, sh = first & 31u;
	const uint4 z = make_uint4(0, 0, 0, 0);
	const uint4 p0 = g < ngrp ? __ldg(sd.planes + gb + g) : z;
	const uint4 p1 = g + 1u < ngrp ? __ldg(sd.planes + gb + g + 1u) : z;
	const uint32_t keep = n >= 32u ? 0xFFFFFFFFu : ((1u << n) - 1u);
	const uint32_t a = __funnelshift_r(p0.x, p1.x, sh) & keep, c = __funnelshift_r(p0.y, p1.y, sh) & keep;
	const uint32_t gg = __funnelshift_r(p0.z, p1.z, sh) & keep, t = __funnelshift_r(p0.w, p1.w, sh) & keep;
	if (!minus) {
		out.a = a << st; out.c = c << st; out.g = gg << st; out.t = t << st;
		loc = lp;
	} else { // word position k of the complement = complement of base n - 1 - k
		const uint32_t r = 32u - n;
		out.a = (__brev(t) >> r) << st; out.c = (__brev(gg) >> r) << st; out.g = (__brev(c) >> r) << st; out.t = (__brev(a) >> r) << st;
		loc = lm;
	}
}

#endif // __CUDACC__

} // namespace pcr
