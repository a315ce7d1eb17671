// scan.cuh -- K1, the seed scan: every candidate oligo against every 32-base window of every
// active sequence (the reference's select_words inner loop, select_words.cpp:93-117, which is
// >= 90 % of its run time -- SURVEY.md section 6).
//
// Reformulation (exact for every IUPAC code on either side).  A candidate word is a primer of n
// bases in a 32-slot frame at [start, stop] with EOS flanks (pcr_assay.cpp:719, word.h:392-418), so
//     cand & window_at(p)   ==   #{k < n : primer[k] n text[p + start + k] != {} }
// and against the minus-strand word of the same window (its reverse complement, sequence.cpp:188)
//     cand & rc(window_at(p)) == #{k < n : rc(primer)[k] n text[p + 31 - stop + k] != {} }.
// Both are "count matching positions of an n-letter pattern laid on the text at x", so a pattern is
// four 32-bit masks B_A,B_C,B_G,B_T (bit k: pattern[k] admits that letter) and with the text held
// as bit-planes (seqdev.cuh) the per-alignment work is
//     m = (B_A & W_A(x)) | (B_C & W_C(x)) | (B_G & W_G(x)) | (B_T & W_T(x));  count = popc(m)
// = 4 LOP3 + 1 POPC + 1 ISETP, against ~24 integer ops for the reference's 128-bit nibble fold.
// W_l(x) (the 32 plane bits from x) is pattern-independent: each thread funnel-shifts its R
// alignments out of shared memory once per tile and keeps them in registers for ALL patterns, so
// HBM is read once per tile and the kernel is bound by the integer pipes, not by memory.
//
// Output: a flat list of hits (candidate, strand, sequence, window, count) with count >= the
// candidate's threshold.  The per-(candidate, sequence) "best tier only" rule of select_words
// (select_words.cpp:99-117) is a max over ALL words of a sequence, so it is applied afterwards on
// the (tiny) hit list, not inside the scan (db.cuh).
#pragma once
#include "seqdev.cuh"

namespace pcr {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_R = 8;                         // alignments per thread held in registers
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_R;  // 2048 template positions per tile
constexpr int SCAN_TILE_GROUPS = SCAN_TILE / 32;  // 64 plane groups (+1 halo)
constexpr int SCAN_PAT_CHUNK = 1024;              // patterns staged in shared memory at a time

// pattern meta word: thr[5:0] | frame_offset[10:6] | minus[11] | cand[31:12]
constexpr uint32_t PAT_MAX_CAND = 1u << 20;
__host__ __device__ __forceinline__ uint32_t pat_meta_pack(uint32_t thr, uint32_t off, uint32_t minus, uint32_t cand)
{
	return (thr & 63u) | ((off & 31u) << 6) | ((minus & 1u) << 11) | (cand << 12);
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

__device__ __forceinline__ void hit_append(const HitSink &hs, uint64_t key, uint32_t val)
{
	const unsigned long long i = atomicAdd(hs.count, 1ull);
	if (i < hs.cap) {
		hs.key[i] = key;
		hs.val[i] = val;
	}
}

// ---------------------------------------------------------------------------------------------
// K1a: full windows.  One CTA per tile of 2048 template positions (persistent, strided), all
// patterns per tile.  __launch_bounds__(256, 2): ~64 registers, 16 warps/SM -- the loop is pure
// register arithmetic with 8 independent chains per thread, so occupancy only has to cover the
// 4-cycle ALU latency and the shared-memory pattern broadcast.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS, 2)
scan_full_kernel(SeqDev sd, const uint32_t *__restrict__ tile_seq, const uint32_t *__restrict__ tile_x0, uint32_t n_tiles,
	const uint4 *__restrict__ pat_mask, const uint32_t *__restrict__ pat_meta, uint32_t n_pat, uint32_t cand_bits, HitSink hs)
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
					const int off = (int)((meta >> 6) & 31u);
					const uint32_t minus = (meta >> 11) & 1u, cand = meta >> 12;
					#pragma unroll
					for (int r = 0; r < SCAN_R; ++r) {
						const uint32_t m = (b.x & wa[r]) | (b.y & wc[r]) | (b.z & wg[r]) | (b.w & wt[r]);
						const int cnt = __popc(m);
						if (cnt >= thr) {
							const int64_t wstart = (int64_t)x0 + r * SCAN_THREADS + tid - off; // window start p
							if (wstart >= 0 && wstart + 32 <= (int64_t)clen)
								hit_append(hs, hit_key_pack(seq, cand, cand_bits, (uint32_t)cnt, ENT_FULL, minus), (uint32_t)wstart + 31u);
						}
					}
				}
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// K1b: the partial words pack() emits at sequence starts, ends and EOS events (FILL / EOSEVT /
// TAIL, seqdev.cuh).  A few dozen per sequence; they are built explicitly and compared with the
// reference's own 128-bit formulation.  One warp per sequence, lanes = events, loop over candidates.
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

__global__ void __launch_bounds__(256)
scan_edge_kernel(SeqDev sd, PackParams pp, const uint64_t *__restrict__ cand_words, const uint32_t *__restrict__ cand_thr,
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
			if (ok) {
				if (d < ec.n_fill) { type = ENT_FILL; pos = d; }
				else if (d < ec.n_fill + ec.n_eos) { type = ENT_EOSEVT; pos = sd.eos_pos[sd.eos_off[seq] + (d - ec.n_fill)]; }
				else { type = ENT_TAIL; pos = d - ec.n_fill - ec.n_eos + 1u; }
			}
			W128 wp, wm;
			int lp, lm;
			wp.hi = wp.lo = wm.hi = wm.lo = 0;
			ok = ok && pack_entry(sd, seq, type, pos, pp, wp, wm, lp, lm);
			if (!__any_sync(0xffffffffu, ok)) continue;
			for (uint32_t c = 0; c < n_cand; ++c) {
				W128 cw;
				cw.hi = __ldg(cand_words + 2 * c);
				cw.lo = __ldg(cand_words + 2 * c + 1);
				const int thr = (int)__ldg(cand_thr + c);
				if (ok) {
					const int np = w_and_count(cw, wp), nm = w_and_count(cw, wm);
					if (np >= thr) hit_append(hs, hit_key_pack(seq, c, cand_bits, (uint32_t)np, type, 0u), pos);
					if (nm >= thr) hit_append(hs, hit_key_pack(seq, c, cand_bits, (uint32_t)nm, type, 1u), pos);
				}
			}
		}
	}
}

} // namespace pcr
