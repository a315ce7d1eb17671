// scan.cuh -- K1, the seed scan: every candidate oligo against every 32-base window of every
// active sequence (the reference's select_words inner loop, select_words.cpp:93-117, which is
// >= 90 % of its run time -- SURVEY.md section 6).
//
// Reformulation (exact for every IUPAC code on either side).  A candidate word is a primer of n
// bases in a 32-slot frame at [start, stop] with EOS flanks (pcr_assay.cpp:719, word.h:392-418), so
//     cand & window_at(p)   ==   #{k < n : primer[k] n text[p + start + k] != {} }
// and against the minus-strand word of the same window (its reverse complement, sequence.cpp:188)
//     cand & rc(window_at(p)) == #{k < n : rc(primer)[k] n text[p + 31 - stop + k] != {} }.
// Both are "count matching positions of an n-letter PATTERN laid on the text at alignment x", so a
// pattern is four 32-bit masks B_A,B_C,B_G,B_T (bit k: pattern[k] admits that letter) and with the
// text held as bit-planes (seqdev.cuh)
//     m = (B_A & W_A(x)) | (B_C & W_C(x)) | (B_G & W_G(x)) | (B_T & W_T(x));  count = popc(m)
// = 4 LOP3 + 1 POPC + 1 ISETP, against ~24 integer ops for the reference's 128-bit nibble fold.
//
// The 5'/3' shift family select_words adds per oligo (select_words.cpp:38-72) lays the SAME primer on
// the SAME text alignments, seen through neighbouring windows: one pattern per (oligo, strand) is
// scanned and a hit is fanned out to every family member whose window exists (emit_family).
//
// Three kernels produce the same flat hit list (candidate, strand, sequence, window, count >= thr):
//   scan_seed_kernel   the fast path.  count >= thr allows e = n - thr mismatches, so of e+1 disjoint
//                      pieces of the primer one must match exactly (pigeonhole).  Piece seeds of q = 5,
//                      6 or 7 bases (all IUPAC expansions) are indexed by their 2-bit code in shared memory;
//                      each text position looks up the primers having that seed and only those
//                      alignments are counted.  Exact, ~10^2-10^3 x fewer alignments than brute force.
//   scan_full_kernel   brute force over all alignments, for patterns whose pieces would be < 5 bases
//                      (low thresholds, e.g. backgrounds at 0.72) or too degenerate to expand.
//   scan_groups_kernel brute force of the seeded patterns over the few 32-base groups whose text
//                      contains a degenerate base (a 2-bit code cannot represent it); the seed kernel
//                      skips exactly those alignments.
// The per-(candidate, sequence) "best tier only" rule of select_words (select_words.cpp:99-117) is a
// max over ALL words of a sequence, so it is applied afterwards on the hit list (db.cuh).
#pragma once
#include "seqdev.cuh"
#include "fst.cuh"

namespace pcr {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_R = 8;                         // alignments per thread held in registers
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_R;  // 2048 template positions per tile
constexpr int SCAN_TILE_GROUPS = SCAN_TILE / 32;  // 64 plane groups (+1 halo)
constexpr int SCAN_PAT_CHUNK = 1024;              // patterns staged in shared memory at a time

// pattern meta word: thr[5:0] | frame_offset[10:6] | minus[11] | base candidate id[31:12]
constexpr uint32_t PAT_MAX_CAND = 1u << 20;
__host__ __device__ __forceinline__ uint32_t pat_meta_pack(uint32_t thr, uint32_t off, uint32_t minus, uint32_t cand)
{
	return (thr & 63u) | ((off & 31u) << 6) | ((minus & 1u) << 11) | (cand << 12);
}
// second meta word: n_left[4:0] | n_right[9:5] | n[15:10] | e[21:16] | seed class q[24:22] (0 = brute force)
__host__ __device__ __forceinline__ uint32_t pat_meta2_pack(uint32_t nl, uint32_t nr, uint32_t n, uint32_t e, uint32_t cls)
{
	return (nl & 31u) | ((nr & 31u) << 5) | ((n & 63u) << 10) | ((e & 63u) << 16) | ((cls & 7u) << 22);
}

// hit key, low to high: minus[0] | type[2:1] | 63-count[8:3] | cand[9 .. 9+cand_bits) | seq[...]
// so that an ascending sort groups by (seq, cand) and puts the best count first inside a group.
constexpr int HIT_GROUP_SHIFT = 9;
__device__ __forceinline__ uint64_t hit_key_pack(uint32_t seq, uint32_t cand, uint32_t cand_bits, uint32_t count, uint32_t type,
	uint32_t minus)
{
	return ((((uint64_t)seq << cand_bits) | cand) << HIT_GROUP_SHIFT) | ((uint64_t)(63u - count) << 3) | (type << 1) | minus;
}

struct HitSink {
	uint64_t *key;
	uint32_t *val;
	unsigned long long *count; // total produced (may exceed cap: the host then grows and re-runs)
	uint64_t cap;
};

// One slot of a list behind a 64-bit counter, one atomic per group of converged lanes (the lanes of a warp that reach this
// call together): millions of appends to ONE counter otherwise serialise in L2 (64-bit adds are not merged by the compiler).
__device__ __forceinline__ unsigned long long warp_slot(unsigned long long *counter)
{
	const unsigned m = __activemask();
	const unsigned lane = threadIdx.x & 31u;
	const int leader = __ffs(m) - 1;
	unsigned long long base = 0;
	if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
	base = __shfl_sync(m, base, leader);
	return base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
}

__device__ __forceinline__ void hit_append(const HitSink &hs, uint64_t key, uint32_t val)
{
	const unsigned long long i = warp_slot(hs.count);
	if (i < hs.cap) {
		hs.key[i] = key;
		hs.val[i] = val;
	}
}

// A pattern matched `cnt` positions at alignment x (text index of primer base 0).  Report it for the
// base candidate and for every 5'/3'-shifted copy whose 32-base window lies inside the sequence:
// the copy shifted j slots toward 5' sees the same alignment through the window starting j later.
__device__ __forceinline__ void emit_family(const HitSink &hs, uint32_t seq, uint32_t clen, uint32_t cand_bits, uint32_t meta, uint32_t meta2,
	int64_t x, uint32_t cnt)
{
	const int off = (int)((meta >> 6) & 31u);
	const uint32_t minus = (meta >> 11) & 1u, base = meta >> 12;
	const int nl = (int)(meta2 & 31u), nr = (int)((meta2 >> 5) & 31u);
	for (int v = 0; v <= nl + nr; ++v) {
		int o;
		if (v == 0) o = off;
		else if (v <= nl) o = minus ? off + v : off - v;               // shift_left v times (select_words.cpp:50-58)
		else o = minus ? off - (v - nl) : off + (v - nl);              // shift_right (v - nl) times (:61-70)
		const int64_t wstart = x - o;
		if (wstart >= 0 && wstart + 32 <= (int64_t)clen)
			hit_append(hs, hit_key_pack(seq, base + (uint32_t)v, cand_bits, cnt, ENT_FULL, minus), (uint32_t)wstart + 31u);
	}
}

// ---------------------------------------------------------------------------------------------
// K1 brute force: one CTA per tile of 2048 alignments (persistent, strided), all given patterns
// per tile.  Each thread funnel-shifts its 8 alignments' plane windows out of shared memory once
// per tile and keeps them in registers for every pattern.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS, 2)
scan_full_kernel(SeqDev sd, const uint32_t *__restrict__ tile_seq, const uint32_t *__restrict__ tile_x0, uint32_t n_tiles,
	const uint4 *__restrict__ pat_mask, const uint32_t *__restrict__ pat_meta, const uint32_t *__restrict__ pat_meta2, uint32_t n_pat,
	uint32_t cand_bits, HitSink hs)
{
	__shared__ uint4 s_grp[SCAN_TILE_GROUPS + 1];
	__shared__ uint4 s_mask[SCAN_PAT_CHUNK];
	__shared__ uint32_t s_meta[SCAN_PAT_CHUNK];

	const uint32_t tid = threadIdx.x;
	for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		const uint32_t seq = tile_seq[tile];
		if (!sd.active[seq]) continue; // main.cpp:581,650: inactive sequences are not indexed
		const uint32_t x0 = tile_x0[tile];
		const uint64_t gbase = sd.grp_off[seq];
		const uint32_t ngrp = (uint32_t)(sd.grp_off[seq + 1] - gbase);
		const uint32_t clen = sd.clen[seq];

		__syncthreads(); // everyone is done with the previous tile's planes
		if (tid <= SCAN_TILE_GROUPS) {
			const uint32_t g = (x0 >> 5) + tid;
			s_grp[tid] = (g < ngrp) ? __ldg(sd.planes + gbase + g) : make_uint4(0, 0, 0, 0);
		}
		__syncthreads();

		uint32_t wa[SCAN_R], wc[SCAN_R], wg[SCAN_R], wt[SCAN_R];
		#pragma unroll
		for (int r = 0; r < SCAN_R; ++r) {
			const uint32_t idx = r * SCAN_THREADS + tid; // a warp reads one group pair: broadcast, no conflicts
			const uint4 lo = s_grp[idx >> 5], hi = s_grp[(idx >> 5) + 1];
			const uint32_t sh = idx & 31u;
			wa[r] = __funnelshift_r(lo.x, hi.x, sh);
			wc[r] = __funnelshift_r(lo.y, hi.y, sh);
			wg[r] = __funnelshift_r(lo.z, hi.z, sh);
			wt[r] = __funnelshift_r(lo.w, hi.w, sh);
		}

		for (uint32_t c0 = 0; c0 < n_pat; c0 += SCAN_PAT_CHUNK) {
			const uint32_t cn = min((uint32_t)SCAN_PAT_CHUNK, n_pat - c0);
			__syncthreads();
			for (uint32_t i = tid; i < cn; i += SCAN_THREADS) {
				s_mask[i] = __ldg(pat_mask + c0 + i);
				s_meta[i] = __ldg(pat_meta + c0 + i);
			}
			__syncthreads();
			#pragma unroll 2
			for (uint32_t p = 0; p < cn; ++p) {
				const uint4 b = s_mask[p];
				const uint32_t meta = s_meta[p];
				const int thr = (int)(meta & 63u);
				bool any = false;
				#pragma unroll
				for (int r = 0; r < SCAN_R; ++r) {
					const uint32_t m = (b.x & wa[r]) | (b.y & wc[r]) | (b.z & wg[r]) | (b.w & wt[r]);
					any |= (__popc(m) >= thr);
				}
				if (any) { // rare: a real binding site (or a chance hit at a low threshold)
					const uint32_t meta2 = __ldg(pat_meta2 + c0 + p);
					#pragma unroll
					for (int r = 0; r < SCAN_R; ++r) {
						const uint32_t m = (b.x & wa[r]) | (b.y & wc[r]) | (b.z & wg[r]) | (b.w & wt[r]);
						const int cnt = __popc(m);
						if (cnt >= thr) emit_family(hs, seq, clen, cand_bits, meta, meta2, (int64_t)x0 + r * SCAN_THREADS + tid, (uint32_t)cnt);
					}
				}
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// K1 brute force over listed 32-alignment groups (the "dirty" groups whose windows touch a
// degenerate text base).  One warp per group, lane = alignment, patterns streamed through smem.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scan_groups_kernel(SeqDev sd, const uint32_t *__restrict__ grp_seq, const uint32_t *__restrict__ grp_idx, uint32_t n_groups,
	const uint4 *__restrict__ pat_mask, const uint32_t *__restrict__ pat_meta, const uint32_t *__restrict__ pat_meta2, uint32_t n_pat,
	uint32_t cand_bits, HitSink hs)
{
	__shared__ uint4 s_mask[SCAN_PAT_CHUNK];
	__shared__ uint32_t s_meta[SCAN_PAT_CHUNK];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	for (uint32_t g0 = blockIdx.x * 8u; g0 < n_groups; g0 += gridDim.x * 8u) {
		const uint32_t gi = g0 + warp;
		bool live = gi < n_groups;
		uint32_t seq = 0, grp = 0, clen = 0;
		uint32_t wa = 0, wc = 0, wg = 0, wt = 0;
		if (live) {
			seq = grp_seq[gi];
			grp = grp_idx[gi];
			live = sd.active[seq] != 0;
		}
		if (live) {
			clen = sd.clen[seq];
			const uint64_t gbase = sd.grp_off[seq];
			const uint32_t ngrp = (uint32_t)(sd.grp_off[seq + 1] - gbase);
			const uint4 lo = __ldg(sd.planes + gbase + grp);
			const uint4 hi = (grp + 1u < ngrp) ? __ldg(sd.planes + gbase + grp + 1u) : make_uint4(0, 0, 0, 0);
			wa = __funnelshift_r(lo.x, hi.x, lane);
			wc = __funnelshift_r(lo.y, hi.y, lane);
			wg = __funnelshift_r(lo.z, hi.z, lane);
			wt = __funnelshift_r(lo.w, hi.w, lane);
		}
		for (uint32_t c0 = 0; c0 < n_pat; c0 += SCAN_PAT_CHUNK) {
			const uint32_t cn = min((uint32_t)SCAN_PAT_CHUNK, n_pat - c0);
			__syncthreads();
			for (uint32_t i = threadIdx.x; i < cn; i += blockDim.x) {
				s_mask[i] = __ldg(pat_mask + c0 + i);
				s_meta[i] = __ldg(pat_meta + c0 + i);
			}
			__syncthreads();
			if (!live) continue;
			for (uint32_t p = 0; p < cn; ++p) {
				const uint4 b = s_mask[p];
				const uint32_t meta = s_meta[p];
				const uint32_t m = (b.x & wa) | (b.y & wc) | (b.z & wg) | (b.w & wt);
				const int cnt = __popc(m);
				if (cnt >= (int)(meta & 63u))
					emit_family(hs, seq, clen, cand_bits, meta, __ldg(pat_meta2 + c0 + p), (int64_t)grp * 32 + lane, (uint32_t)cnt);
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// K1 fast path: pigeonhole seed filter.
//
// Shared memory (one CTA of 1024 threads per SM, everything hot is on chip):
//   bucket[1024 + 4096 + 8192]  per 2-bit seed code (q = 5, 6, 7): first entry | count << 20
//   entries[]            seed -> pattern << 11 | threshold << 5 | 27 - offset of the seed inside the pattern
//   mask[]               the chunk's pattern masks
// The seed code at text position x is the q low bits of the code planes b0 = C|T, b1 = G|T (A=00 C=01 G=10
// T=11).  For every (pattern, offset) in the bucket the alignment x - offset is counted with the same
// 4 LOP3 + POPC as the brute-force kernel.  A primer's e+1 pieces are the even split of its n bases; a
// piece's seed is its first min(len, 7) bases.  An alignment found through several of its pieces is
// reported by the leftmost matching one only.
// ---------------------------------------------------------------------------------------------
constexpr int SEED_THREADS = 1024;
constexpr uint32_t SEED_B5 = 0u, SEED_B6 = 1024u, SEED_B7 = 1024u + 4096u; // bucket table offsets per seed length
constexpr uint32_t SEED_BUCKETS = 1024u + 4096u + 8192u;                    // q = 7 codes (14 bits) are folded to 13
constexpr uint32_t SEED_MAX_EXPANSIONS = 64u;          // per pattern, over all its pieces
constexpr uint32_t SEED_MAX_PATTERNS = 1u << 21;       // entry word: pattern[31:11] | thr[10:5] | 27 - offset[4:0]
constexpr uint32_t SEED_QMIN = 5u, SEED_QMAX = 7u;

__host__ __device__ __forceinline__ uint32_t seed_bucket_index(uint32_t q, uint32_t b0, uint32_t b1)
{ // b0 / b1 = the q low bits of the two code planes
	if (q == 5u) return SEED_B5 + (b0 | (b1 << 5));
	if (q == 6u) return SEED_B6 + (b0 | (b1 << 6));
	const uint32_t c = b0 | (b1 << 7);
	return SEED_B7 + ((c ^ (c >> 13)) & 8191u); // two codes per bucket; verification discards the stranger
}

struct SeedChunk {
	const uint32_t *bucket;  // SEED_BUCKETS words
	const uint32_t *entries; // n_entries words
	const uint4 *mask;       // n_pat
	const uint32_t *meta;    // n_pat (read on hits only, stays in global memory)
	const uint32_t *meta2;   // n_pat (ditto)
	uint32_t n_entries, n_pat;
	uint32_t ecap, pcap;     // shared-memory capacities the launch was sized for (multiples of 4)
};

__host__ __device__ __forceinline__ size_t seed_smem_bytes(uint32_t ecap, uint32_t pcap)
{
	return (size_t)SEED_BUCKETS * 4 + (size_t)ecap * 4 + (size_t)pcap * 16;
}

// piece i of a pattern of n bases cut into `pieces`: [o, o + len); its seed is the first q = min(len, 7) bases
__host__ __device__ __forceinline__ void seed_piece(uint32_t n, uint32_t pieces, uint32_t i, uint32_t &o, uint32_t &q)
{
	o = (i * n) / pieces;
	const uint32_t len = ((i + 1u) * n) / pieces - o;
	q = len < SEED_QMAX ? len : SEED_QMAX;
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
	uint32_t v;
	asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
	uint4 v;
	asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}

// bits [start, start + 64) of the 96-bit string lo:mid:hi (bit 0 = LSB of lo), zero beyond; start in [0, 63]
__device__ __forceinline__ void take64(uint32_t lo, uint32_t mid, uint32_t hi, uint32_t start, uint32_t &o_lo, uint32_t &o_hi)
{
	const bool up = start >= 32u;
	const uint32_t sh = start & 31u;
	o_lo = __funnelshift_r(up ? mid : lo, up ? hi : mid, sh);
	o_hi = __funnelshift_r(up ? hi : mid, up ? 0u : hi, sh);
}

__device__ __forceinline__ void seed_hit(uint32_t en, uint32_t m, uint32_t xpos, uint64_t gbase, const uint32_t *__restrict__ dirty_bits,
	const uint32_t *__restrict__ g_meta, const uint32_t *__restrict__ g_meta2, uint32_t seq, uint32_t clen, uint32_t cand_bits, const HitSink &hs)
{
	const uint32_t off = 27u - (en & 31u), pid = en >> 11;
	const int64_t x = (int64_t)xpos - off;
	if (x < 0) return;
	if (dirty_bits) { // alignments touching a degenerate text base belong to scan_groups_kernel
		const uint64_t G = gbase + (uint64_t)(x >> 5);
		if ((dirty_bits[G >> 5] >> (G & 31u)) & 1u) return;
	}
	const uint32_t meta = __ldg(g_meta + pid), meta2 = __ldg(g_meta2 + pid);
	const uint32_t n = (meta2 >> 10) & 63u, pieces = ((meta2 >> 16) & 63u) + 1u;
	// Report through the leftmost exactly-matching piece only -- and that piece must be THIS one.  A q = 7 bucket holds two seed
	// codes, so an entry is also proposed at texts that carry the bucket's other code (base 0 and base 6 of the seed off by one
	// code bit each).  With a degenerate base 0 (M, K) one of those two is no mismatch at all, the alignment can pass the count, and
	// it used to be reported here AND through a later exact piece: the same hit twice (deduplicated downstream, but n_hits and the
	// hit buffers carried it).  An alignment at or above threshold always has an exact piece (pigeonhole), so nothing is lost.
	for (uint32_t i = 0; i < pieces; ++i) {
		uint32_t o, pq;
		seed_piece(n, pieces, i, o, pq);
		const uint32_t pm = (1u << pq) - 1u;
		const bool exact = ((m >> o) & pm) == pm;
		if (o >= off) {
			if (!exact) return;
			break;
		}
		if (exact) return;
	}
	emit_family(hs, seq, clen, cand_bits, meta, meta2, x, (uint32_t)__popc(m));
}

// one table entry: count the alignment it proposes (1 shared load + 4 SHF + 4 LOP3 + POPC + compare)
#define seed_verify(EN, A_MASK, CA0, CA1, CC0, CC1, CG0, CG1, CT0, CT1, XPOS, GBASE, DIRTY, CH, SEQ, CLEN, CBITS, HS)                     \
	do {                                                                                                                              \
		const uint32_t en__ = (EN);                                                                                                   \
		const uint4 b__ = lds128((A_MASK) + ((en__ >> 7) & 0x1FFFFF0u));                                                            \
		const uint32_t sh__ = en__ & 31u;                                                                                             \
		const uint32_t m__ = (b__.x & __funnelshift_r(CA0, CA1, sh__)) | (b__.y & __funnelshift_r(CC0, CC1, sh__)) |                  \
		                     (b__.z & __funnelshift_r(CG0, CG1, sh__)) | (b__.w & __funnelshift_r(CT0, CT1, sh__));                  \
		if ((uint32_t)__popc(m__) << 5 >= (en__ & 0x7E0u)) /* count >= thr: a binding site (rare) */                                  \
			seed_hit(en__, m__, XPOS, GBASE, DIRTY, (CH).meta, (CH).meta2, SEQ, CLEN, CBITS, HS);                                    \
	} while (0)

// The seed table and the pattern masks are staged once per CTA in shared memory (the only barrier).  After
// that every WARP works alone: it pulls a tile (64 groups = 2048 text positions of one sequence) from a
// global counter and slides over it one group at a time -- lane = text position, the group's planes and its
// two neighbours arrive as warp-uniform 16-byte loads straight from L2 (each base is read once; there is no
// shared-memory tile and no block barrier to wait at).  Per step a lane cuts, for each letter plane, the 64
// bits that every alignment it may have to verify falls into (x - offset, offset <= 27), so a verification is
// ONE 16-byte shared load (pattern masks) + 4 funnel shifts + 4 LOP3 + POPC.
__global__ void __launch_bounds__(SEED_THREADS, 1)
scan_seed_kernel(SeqDev sd, const uint32_t *__restrict__ tile_seq, const uint32_t *__restrict__ tile_x0, uint32_t n_tiles,
	unsigned int *tile_counter, SeedChunk ch, const uint32_t *__restrict__ dirty_bits, uint32_t cand_bits, HitSink hs)
{
	extern __shared__ __align__(16) uint32_t smem[];
	{
		uint32_t *s_bucket = smem;
		uint32_t *s_entries = s_bucket + SEED_BUCKETS;
		uint4 *s_mask = (uint4 *)(s_entries + ch.ecap);
		for (uint32_t i = threadIdx.x; i < SEED_BUCKETS; i += SEED_THREADS) s_bucket[i] = __ldg(ch.bucket + i);
		for (uint32_t i = threadIdx.x; i < ch.n_entries; i += SEED_THREADS) s_entries[i] = __ldg(ch.entries + i);
		for (uint32_t i = threadIdx.x; i < ch.n_pat; i += SEED_THREADS) s_mask[i] = __ldg(ch.mask + i);
	}
	__syncthreads();
	const uint32_t a_bucket = (uint32_t)__cvta_generic_to_shared(smem);
	const uint32_t a_entries = a_bucket + SEED_BUCKETS * 4u;
	const uint32_t a_mask = a_entries + ch.ecap * 4u;
	const uint32_t lane = threadIdx.x & 31u;

	for (;;) {
		uint32_t tile = 0;
		if (lane == 0u) tile = atomicAdd(tile_counter, 1u);
		tile = __shfl_sync(0xffffffffu, tile, 0);
		if (tile >= n_tiles) break;
		const uint32_t seq = __ldg(tile_seq + tile);
		if (!sd.active[seq]) continue;
		const uint32_t x0 = __ldg(tile_x0 + tile);
		const uint64_t gbase = sd.grp_off[seq];
		const uint32_t ngrp = (uint32_t)(sd.grp_off[seq + 1] - gbase); // includes the zero halo group
		const uint32_t clen = sd.clen[seq];
		const uint32_t g0 = x0 >> 5;
		const uint32_t g_end = min(g0 + (uint32_t)SCAN_TILE_GROUPS, (clen + 31u) >> 5);
		const uint4 zero4 = make_uint4(0, 0, 0, 0);
		uint4 gm = g0 > 0u ? __ldg(sd.planes + gbase + g0 - 1u) : zero4;
		uint4 gc = __ldg(sd.planes + gbase + g0);
		uint4 gp = (g0 + 1u < ngrp) ? __ldg(sd.planes + gbase + g0 + 1u) : zero4;
		for (uint32_t g = g0; g < g_end; ++g) {
			__syncwarp(); // lanes leave the entry loops at different times: re-form the warp for the uniform part
			const uint4 gn = (g + 2u < ngrp) ? __ldg(sd.planes + gbase + g + 2u) : zero4; // prefetch for the next step
			const uint32_t xpos = (g << 5) + lane;
			if (xpos < clen) {
				// 64-bit cuts of the four letter planes starting 27 bases left of this lane's position
				uint32_t ca0, ca1, cc0, cc1, cg0, cg1, ct0, ct1;
				take64(gm.x, gc.x, gp.x, lane + 5u, ca0, ca1);
				take64(gm.y, gc.y, gp.y, lane + 5u, cc0, cc1);
				take64(gm.z, gc.z, gp.z, lane + 5u, cg0, cg1);
				take64(gm.w, gc.w, gp.w, lane + 5u, ct0, ct1);
				// seed codes at x: low q bits of the code planes b0 = C|T, b1 = G|T (A=00 C=01 G=10 T=11)
				const uint32_t w0 = __funnelshift_r(gc.y | gc.w, gp.y | gp.w, lane);
				const uint32_t w1 = __funnelshift_r(gc.z | gc.w, gp.z | gp.w, lane);
				// one loop over the position's three buckets: a lane's trip count is the SUM of its three bucket
				// sizes, whose spread across the warp is smaller than the three spreads added up
				const uint32_t b5 = lds32(a_bucket + 4u * seed_bucket_index(5u, w0 & 31u, w1 & 31u));
				const uint32_t b6 = lds32(a_bucket + 4u * seed_bucket_index(6u, w0 & 63u, w1 & 63u));
				const uint32_t b7 = lds32(a_bucket + 4u * seed_bucket_index(7u, w0 & 127u, w1 & 127u));
				const uint32_t n5 = b5 >> 20, n56 = n5 + (b6 >> 20), n567 = n56 + (b7 >> 20);
				const uint32_t e5 = a_entries + 4u * (b5 & 0xFFFFFu), e6 = a_entries + 4u * ((b6 & 0xFFFFFu) - n5),
				               e7 = a_entries + 4u * ((b7 & 0xFFFFFu) - n56);
				for (uint32_t j = 0; j < n567; ++j)
					seed_verify(lds32((j < n5 ? e5 : (j < n56 ? e6 : e7)) + 4u * j), a_mask, ca0, ca1, cc0, cc1, cg0, cg1, ct0, ct1, xpos, gbase,
						dirty_bits, ch, seq, clen, cand_bits, hs);
			}
			gm = gc;
			gc = gp;
			gp = gn;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// K1b: the partial words pack() emits at sequence starts, ends and EOS events (FILL / EOSEVT /
// TAIL, seqdev.cuh).  A few dozen per sequence; they are built explicitly and compared with the
// same letter-plane count as everywhere else.  One warp per sequence, lanes = events, loop over candidates.
// ---------------------------------------------------------------------------------------------
struct EdgeCounts {
	uint32_t n_fill, n_eos, n_tail;
};

__device__ inline EdgeCounts edge_counts(const SeqDev &sd, uint32_t seq, const PackParams &pp)
{
	EdgeCounts ec;
	const uint32_t L = sd.plen[seq], Lc = sd.clen[seq];
	ec.n_fill = (Lc >= 32u) ? raw_of_comp(sd, seq, 31u) : L; // raw indices [0, n_fill) have c_i < 32
	ec.n_eos = sd.eos_off[seq + 1] - sd.eos_off[seq];         // every EOS is tried; pack_entry() keeps those with c_i >= 32
	uint32_t s0 = (Lc < 32u) ? Lc : 31u;
	ec.n_tail = (s0 > pp.min_len) ? s0 - pp.min_len : 0u;     // q = 1 .. s0 - min_len
	return ec;
}

// event d of a sequence's partial-word list: FILL events, then EOS events, then TAIL events
__device__ __forceinline__ void edge_event(const SeqDev &sd, uint32_t seq, const EdgeCounts &ec, uint32_t d, uint32_t &type, uint32_t &pos)
{
	if (d < ec.n_fill) { type = ENT_FILL; pos = d; }
	else if (d < ec.n_fill + ec.n_eos) { type = ENT_EOSEVT; pos = sd.eos_pos[sd.eos_off[seq] + (d - ec.n_fill)]; }
	else { type = ENT_TAIL; pos = d - ec.n_fill - ec.n_eos + 1u; }
}

__global__ void __launch_bounds__(256)
scan_edge_kernel(SeqDev sd, PackParams pp, const uint4 *__restrict__ cand_planes, const uint32_t *__restrict__ cand_thr,
	uint32_t n_cand, uint32_t cand_bits, HitSink hs)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t seq = warp; seq < sd.n; seq += n_warps) {
		if (!sd.active[seq]) continue;
		const EdgeCounts ec = edge_counts(sd, seq, pp);
		const uint32_t total = ec.n_fill + ec.n_eos + ec.n_tail;
		for (uint32_t d0 = 0; d0 < total; d0 += 32u) {
			const uint32_t d = d0 + lane;
			uint32_t type = 0, pos = 0;
			bool ok = d < total;
			if (ok) edge_event(sd, seq, ec, d, type, pos);
			W128 wp, wm;
			int lp, lm;
			wp.hi = wp.lo = wm.hi = wm.lo = 0;
			ok = ok && pack_entry(sd, seq, type, pos, pp, wp, wm, lp, lm);
			if (!__any_sync(0xffffffffu, ok)) continue;
			const Planes4 pp4 = w_planes(wp), pm4 = w_planes(wm);
			#pragma unroll 4
			for (uint32_t c = 0; c < n_cand; ++c) {
				const uint4 cv = __ldg(cand_planes + c); // uniform address: one broadcast load per warp
				const int thr = (int)__ldg(cand_thr + c);
				const int np = __popc((cv.x & pp4.a) | (cv.y & pp4.c) | (cv.z & pp4.g) | (cv.w & pp4.t));
				const int nm = __popc((cv.x & pm4.a) | (cv.y & pm4.c) | (cv.z & pm4.g) | (cv.w & pm4.t));
				if (ok && (np >= thr || nm >= thr)) {
					if (np >= thr) hit_append(hs, hit_key_pack(seq, c, cand_bits, (uint32_t)np, type, 0u), pos);
					if (nm >= thr) hit_append(hs, hit_key_pack(seq, c, cand_bits, (uint32_t)nm, type, 1u), pos);
				}
			}
		}
	}
}

// the same scan through the frame-aligned seed table (fst.cuh): one thread per partial word, which only meets the
// candidates that share a seed with it
__global__ void __launch_bounds__(128)
scan_edge_fst_kernel(SeqDev sd, PackParams pp, Fst t, uint32_t cand_bits, HitSink hs)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t seq = warp; seq < sd.n; seq += n_warps) {
		if (!sd.active[seq]) continue;
		const EdgeCounts ec = edge_counts(sd, seq, pp);
		const uint32_t total = ec.n_fill + ec.n_eos + ec.n_tail;
		for (uint32_t d0 = 0; d0 < total; d0 += 32u) {
			const uint32_t d = d0 + lane;
			uint32_t type = 0, pos = 0;
			bool ok = d < total;
			if (ok) edge_event(sd, seq, ec, d, type, pos);
			W128 wp, wm;
			int lp, lm;
			wp.hi = wp.lo = wm.hi = wm.lo = 0;
			ok = ok && pack_entry(sd, seq, type, pos, pp, wp, wm, lp, lm);
			if (!ok) continue;
			const Planes4 pp4 = w_planes(wp), pm4 = w_planes(wm);
			fst_match<true>(t, fst_word(pp4.a, pp4.c, pp4.g, pp4.t),
				[&](uint32_t c, uint32_t m) { hit_append(hs, hit_key_pack(seq, c, cand_bits, (uint32_t)__popc(m), type, 0u), pos); });
			fst_match<true>(t, fst_word(pm4.a, pm4.c, pm4.g, pm4.t),
				[&](uint32_t c, uint32_t m) { hit_append(hs, hit_key_pack(seq, c, cand_bits, (uint32_t)__popc(m), type, 1u), pos); });
		}
	}
}

// ---------------------------------------------------------------------------------------------
// dirty groups: a 32-base group is "dirty" when it or its right neighbour holds a degenerate base,
// i.e. when some alignment starting in it reads a base that is not exactly one letter.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool group_degenerate(const uint4 g)
{
	return ((g.x & g.y) | (g.x & g.z) | (g.x & g.w) | (g.y & g.z) | (g.y & g.w) | (g.z & g.w)) != 0u;
}

// one thread per 32 consecutive GLOBAL groups -> one word of dirty_bits; sequences are found by search
__global__ void dirty_bits_kernel(SeqDev sd, uint64_t n_groups, uint32_t *dirty_bits, uint32_t *list_seq, uint32_t *list_grp,
	unsigned int *n_list, uint32_t list_cap)
{
	const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (w * 32ull >= n_groups) return;
	uint32_t lo = 0, hi = sd.n; // sequence owning the first group of this word
	while (hi - lo > 1u) {
		const uint32_t mid = (lo + hi) >> 1;
		if (sd.grp_off[mid] <= w * 32ull) lo = mid; else hi = mid;
	}
	uint32_t seq = lo, bits = 0;
	for (uint32_t b = 0; b < 32u; ++b) {
		const uint64_t G = w * 32ull + b;
		if (G >= n_groups) break;
		while (seq + 1u < sd.n && sd.grp_off[seq + 1] <= G) ++seq;
		const bool last = (G + 1 == sd.grp_off[seq + 1]); // the zero halo group of this sequence
		const bool d = group_degenerate(__ldg(sd.planes + G)) || (!last && group_degenerate(__ldg(sd.planes + G + 1)));
		if (d) {
			bits |= 1u << b;
			const unsigned int o = atomicAdd(n_list, 1u);
			if (o < list_cap) {
				list_seq[o] = seq;
				list_grp[o] = (uint32_t)(G - sd.grp_off[seq]);
			}
		}
	}
	dirty_bits[w] = bits;
}

// ---------------------------------------------------------------------------------------------
// seed tables (per chunk of seeded patterns), built on the device
// ---------------------------------------------------------------------------------------------
// Enumerate the IUPAC expansions of the seed of piece `i` of a pattern and call f(bucket, offset).
template <class F>
__device__ __forceinline__ void for_each_seed(const uint4 &mask, uint32_t n, uint32_t pieces, uint32_t i, F f)
{
	uint32_t o, q;
	seed_piece(n, pieces, i, o, q);
	// per-position letter sets as 4-bit {A,C,G,T}; letter codes A=0 C=1 G=2 T=3 -> b0 = code & 1, b1 = code >> 1
	uint32_t sets[SEED_QMAX];
	uint32_t total = 1;
	for (uint32_t k = 0; k < q; ++k) {
		const uint32_t s = ((mask.x >> (o + k)) & 1u) | (((mask.y >> (o + k)) & 1u) << 1) | (((mask.z >> (o + k)) & 1u) << 2) |
		                   (((mask.w >> (o + k)) & 1u) << 3);
		sets[k] = s;
		total *= (uint32_t)__popc(s);
	}
	for (uint32_t t = 0; t < total; ++t) { // mixed-radix counter over the sets
		uint32_t rem = t, b0 = 0, b1 = 0;
		for (uint32_t k = 0; k < q; ++k) {
			const uint32_t cnt = (uint32_t)__popc(sets[k]);
			uint32_t pick = rem % cnt;
			rem /= cnt;
			uint32_t s = sets[k], letter = 0;
			for (;;) { // pick-th set bit
				letter = (uint32_t)(__ffs(s) - 1);
				if (pick == 0) break;
				--pick;
				s &= s - 1u;
			}
			b0 |= (letter & 1u) << k;
			b1 |= (letter >> 1) << k;
		}
		f(seed_bucket_index(q, b0, b1), o);
	}
}

// number of seed-table entries a pattern needs, or 0 when it cannot be seeded
__device__ __forceinline__ uint32_t seed_entries_needed(const uint4 &mask, uint32_t n, uint32_t pieces)
{
	uint32_t sum = 0;
	for (uint32_t i = 0; i < pieces; ++i) {
		uint32_t o, q;
		seed_piece(n, pieces, i, o, q);
		if (q < SEED_QMIN) return 0u;
		uint32_t total = 1;
		for (uint32_t k = 0; k < q; ++k) {
			const uint32_t c = ((mask.x >> (o + k)) & 1u) + ((mask.y >> (o + k)) & 1u) + ((mask.z >> (o + k)) & 1u) + ((mask.w >> (o + k)) & 1u);
			if (c == 0u) return 0u; // an EOS inside the primer: leave it to brute force
			total *= c;
			if (total > SEED_MAX_EXPANSIONS) return 0u;
		}
		sum += total;
		if (sum > SEED_MAX_EXPANSIONS) return 0u;
	}
	return sum;
}

__device__ __forceinline__ bool idx_indexable(const uint4 &m, uint32_t meta2); // index.cuh

// skip_indexable = 1: patterns the indexed scan (index.cuh) takes are left out of the table; 2: ONLY those go in (the pass over
// sequences whose index entries are stale); 0: every pattern
__global__ void seed_count_kernel(const uint4 *__restrict__ mask, const uint32_t *__restrict__ meta2, uint32_t n_pat, uint32_t *bucket_cnt,
	int skip_indexable)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n_pat) return;
	const uint4 m = mask[p];
	if (skip_indexable && (idx_indexable(m, meta2[p]) == (skip_indexable == 1))) return;
	const uint32_t m2 = meta2[p], n = (m2 >> 10) & 63u, pieces = ((m2 >> 16) & 63u) + 1u;
	for (uint32_t i = 0; i < pieces; ++i) for_each_seed(m, n, pieces, i, [&](uint32_t code, uint32_t) { atomicAdd(bucket_cnt + code, 1u); });
}

// bucket_cnt holds the exclusive start of every bucket on entry and is advanced as a cursor
__global__ void seed_fill_kernel(const uint4 *__restrict__ mask, const uint32_t *__restrict__ meta, const uint32_t *__restrict__ meta2,
	uint32_t n_pat, uint32_t *cursor, uint32_t *entries, uint32_t ecap, int skip_indexable)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n_pat) return;
	const uint4 m = mask[p];
	if (skip_indexable && (idx_indexable(m, meta2[p]) == (skip_indexable == 1))) return;
	const uint32_t thr = meta[p] & 63u;
	const uint32_t m2 = meta2[p], n = (m2 >> 10) & 63u, pieces = ((m2 >> 16) & 63u) + 1u;
	for (uint32_t i = 0; i < pieces; ++i)
		for_each_seed(m, n, pieces, i, [&](uint32_t code, uint32_t o) {
			const uint32_t slot = atomicAdd(cursor + code, 1u);
			if (slot < ecap) entries[slot] = (p << 11) | (thr << 5) | (27u - o);
		});
}

__global__ void seed_bucket_pack_kernel(const uint32_t *__restrict__ start, const uint32_t *__restrict__ cnt, uint32_t *bucket)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < SEED_BUCKETS) bucket[i] = (start[i] & 0xFFFFFu) | (min(cnt[i], 4095u) << 20);
}

} // namespace pcr
