// word128.cuh -- the 32-nibble IUPAC word of PCRamp as two 64-bit limbs, host + device.
//
// Semantics follow the reference's __word<unsigned long,2> (word.h:12-690, word.cpp:24-231):
// nibble i lives in limb i/16 at bit (15 - i%16)*4, so limb 0's top nibble is the 5' end and the
// numeric order of (hi, lo) is the lexicographic order of the nibbles (word.h:197-211).
// A = 1, C = 2, G = 4, T = 8, degenerate = OR, 0 = EOS (base_table.h:9-28).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PCR_HD __host__ __device__ __forceinline__
#else
#define PCR_HD inline
#endif

namespace pcr {

struct W128 {
	uint64_t hi, lo;
};

enum : uint32_t { STRAND_PLUS = 1u, STRAND_MINUS = 2u }; // sequence.h:27-32
constexpr int WORD_LEN = 32;                                // word.h:692

PCR_HD int popc64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
	return __popcll(x);
#else
	return __builtin_popcountll(x);
#endif
}
PCR_HD int clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
	return __clzll((long long)x);
#else
	return x ? __builtin_clzll(x) : 64;
#endif
}
PCR_HD int ctz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
	return x ? (__ffsll((long long)x) - 1) : 64;
#else
	return x ? __builtin_ctzll(x) : 64;
#endif
}

// One bit per nibble (at the nibble's top bit) set iff the nibble is non-zero.
PCR_HD uint64_t nibble_nonzero(uint64_t v)
{
	return (v | (v << 1) | (v << 2) | (v << 3)) & 0x8888888888888888ull;
}

PCR_HD uint32_t w_get(const W128 &w, int i) { return (uint32_t)(((i < 16) ? w.hi : w.lo) >> ((15 - (i & 15)) * 4)) & 0xFu; }

PCR_HD void w_set(W128 &w, int i, uint32_t b)
{
	const int sh = (15 - (i & 15)) * 4;
	uint64_t &limb = (i < 16) ? w.hi : w.lo;
	limb = (limb & ~(0xFull << sh)) | ((uint64_t)(b & 0xFu) << sh);
}

// Word::operator& (word.cpp:111-154): number of positions whose nibble sets intersect.
PCR_HD int w_and_count(const W128 &a, const W128 &b) { return popc64(nibble_nonzero(a.hi & b.hi)) + popc64(nibble_nonzero(a.lo & b.lo)); }

// Word::size (word.cpp:198-213): number of non-EOS positions.
PCR_HD int w_size(const W128 &a) { return popc64(nibble_nonzero(a.hi)) + popc64(nibble_nonzero(a.lo)); }

// Word::start / stop (word.h:256-288): first / last non-EOS index, 32 / -1 for the empty word.
PCR_HD int w_start(const W128 &a)
{
	const uint64_t h = nibble_nonzero(a.hi), l = nibble_nonzero(a.lo);
	if (h) return clz64(h) >> 2;
	if (l) return 16 + (clz64(l) >> 2);
	return WORD_LEN;
}
PCR_HD int w_stop(const W128 &a)
{
	const uint64_t h = nibble_nonzero(a.hi), l = nibble_nonzero(a.lo);
	if (l) return 31 - (ctz64(l) >> 2);
	if (h) return 15 - (ctz64(h) >> 2);
	return -1;
}

// shift by k nibbles toward the 5' end (shift_left, word.cpp:215-222) or the 3' end (:224-231)
PCR_HD W128 w_shl(const W128 &a, int k)
{
	W128 r;
	const int s = 4 * k;
	if (s == 0) return a;
	if (s >= 128) { r.hi = r.lo = 0; return r; }
	if (s >= 64) { r.hi = a.lo << (s - 64); r.lo = 0; return r; }
	r.hi = (a.hi << s) | (a.lo >> (64 - s));
	r.lo = a.lo << s;
	return r;
}
PCR_HD W128 w_shr(const W128 &a, int k)
{
	W128 r;
	const int s = 4 * k;
	if (s == 0) return a;
	if (s >= 128) { r.hi = r.lo = 0; return r; }
	if (s >= 64) { r.lo = a.hi >> (s - 64); r.hi = 0; return r; }
	r.lo = (a.lo >> s) | (a.hi << (64 - s));
	r.hi = a.hi >> s;
	return r;
}

// Word::center (word.h:392-418): left = start, right = 32 - stop (sic), delta = (right-left)/2
// truncating toward zero; shift toward 3' when positive, toward 5' when negative.
PCR_HD W128 w_center(const W128 &a)
{
	const int left = w_start(a);
	int right = w_stop(a);
	if (left > right) return a;
	right = WORD_LEN - right;
	const int delta = (right - left) / 2;
	return delta > 0 ? w_shr(a, delta) : w_shl(a, -delta);
}

// Word::max_overlap (word.h:38-92): dp[i][j] = dp[i-1][j-1] + (q[i] == s[j]) never resets, so the result is the largest
// number of EQUAL nibbles on one diagonal of [start, stop] x [start, stop], over max(size, size).
PCR_HD float w_max_overlap(const W128 &q, const W128 &s)
{
	const int qs = w_start(q), qe = w_stop(q), ss = w_start(s), se = w_stop(s);
	if (qs > qe || ss > se) { // an empty word: no cell is visited; 0 / max(size, size) (0/0 = NaN when both are empty)
		const int d = w_size(q) > w_size(s) ? w_size(q) : w_size(s);
		return 0.0f / (float)d;
	}
	const int nq = qe - qs + 1, ns = se - ss + 1;
	const W128 a = w_shl(q, qs), b = w_shl(s, ss);
	W128 ones;
	ones.hi = ones.lo = 0x8888888888888888ull;
	int best = 0;
	for (int d = -(ns - 1); d <= nq - 1; ++d) { // q position i faces s position i - d
		const W128 y = d >= 0 ? w_shr(b, d) : w_shl(b, -d);
		const int lo = d > 0 ? d : 0;
		int hi = ns + d < nq ? ns + d : nq; // exclusive
		if (hi <= lo) continue;
		const W128 m = w_shr(w_shl(ones, 32 - (hi - lo)), lo);
		const uint64_t eh = ~nibble_nonzero(a.hi ^ y.hi) & m.hi, el = ~nibble_nonzero(a.lo ^ y.lo) & m.lo;
		const int c = popc64(eh) + popc64(el);
		best = c > best ? c : best;
	}
	const int den = w_size(q) > w_size(s) ? w_size(q) : w_size(s);
#ifdef __CUDA_ARCH__
	return __fdiv_rn((float)best, (float)den);
#else
	return (float)best / (float)den;
#endif
}

// complement of every nibble: A<->T (bits 0,3), C<->G (bits 1,2) -- i.e. reverse the 4 bits.
PCR_HD uint64_t comp_nibbles(uint64_t v)
{
	v = ((v & 0x5555555555555555ull) << 1) | ((v >> 1) & 0x5555555555555555ull);
	v = ((v & 0x3333333333333333ull) << 2) | ((v >> 2) & 0x3333333333333333ull);
	return v;
}
// reverse the order of the 16 nibbles of a limb
PCR_HD uint64_t rev_nibbles(uint64_t v)
{
	v = ((v & 0x0F0F0F0F0F0F0F0Full) << 4) | ((v >> 4) & 0x0F0F0F0F0F0F0F0Full);
	v = ((v & 0x00FF00FF00FF00FFull) << 8) | ((v >> 8) & 0x00FF00FF00FF00FFull);
	v = ((v & 0x0000FFFF0000FFFFull) << 16) | ((v >> 16) & 0x0000FFFF0000FFFFull);
	return (v << 32) | (v >> 32);
}
// Word::complement (word.h:140-183): reverse complement of [start, stop], written LEFT-justified.
PCR_HD W128 w_complement(const W128 &a)
{
	const int last = w_stop(a);
	if (last < 0) return a;
	W128 r;
	r.hi = comp_nibbles(rev_nibbles(a.lo)); // position i -> 31 - i
	r.lo = comp_nibbles(rev_nibbles(a.hi));
	return w_shl(r, 31 - last);
}

// Word::degeneracy (word.h:97-138) as an exact integer where it fits; saturates at 2^62 (any
// threshold the reference can express is an unsigned int, pcramp.h:120).
PCR_HD uint64_t w_degeneracy_sat(const W128 &a)
{
	uint64_t d = 1;
	for (int i = 0; i < WORD_LEN; ++i) {
		const uint32_t b = w_get(a, i);
		const uint32_t c = (b & 1u) + ((b >> 1) & 1u) + ((b >> 2) & 1u) + (b >> 3);
		if (c > 1u) {
			if (d > (1ull << 60)) return 1ull << 62;
			d *= c;
		}
	}
	return d;
}

// The word as four 32-bit letter planes over the FRAME: bit k of plane l set iff nibble k admits letter l.
// With both operands in this form Word::operator& is popc((a.A&b.A)|(a.C&b.C)|(a.G&b.G)|(a.T&b.T)).
struct Planes4 {
	uint32_t a, c, g, t;
};
PCR_HD Planes4 w_planes(const W128 &w)
{
	Planes4 p;
	p.a = p.c = p.g = p.t = 0u;
	for (int k = 0; k < WORD_LEN; ++k) {
		const uint32_t b = w_get(w, k);
		p.a |= (b & 1u) << k;
		p.c |= ((b >> 1) & 1u) << k;
		p.g |= ((b >> 2) & 1u) << k;
		p.t |= ((b >> 3) & 1u) << k;
	}
	return p;
}
PCR_HD int planes_and_count(const Planes4 &x, const Planes4 &y)
{
	const uint32_t m = (x.a & y.a) | (x.c & y.c) | (x.g & y.g) | (x.t & y.t);
#if defined(__CUDA_ARCH__)
	return __popc(m);
#else
	return __builtin_popcount(m);
#endif
}
PCR_HD uint32_t planes_nibble(const Planes4 &p, int k)
{
	return ((p.a >> k) & 1u) | (((p.c >> k) & 1u) << 1) | (((p.g >> k) & 1u) << 2) | (((p.t >> k) & 1u) << 3);
}

PCR_HD bool is_degen_nibble(uint32_t b) { return !(b == 1u || b == 2u || b == 4u || b == 8u); } // base_table.h:124-137

} // namespace pcr
