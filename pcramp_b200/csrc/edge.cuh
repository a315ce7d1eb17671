// edge.cuh -- K1b turned around: the partial words of a collection as a table the candidates look themselves up in.
//
// scan_edge_fst_kernel (scan.cuh) rebuilds, in EVERY batch, the few dozen partial words pack() emits per sequence (FILL / EOSEVT / TAIL
// events, seqdev.cuh) and walks each of them through a seed table over the batch's candidates: ~40 (frame position, q) lookups per
// word, 1.8 x 10^6 words for 20 000 sequences -- 1.2 x 10^8 warp instructions, 0.17 ms of every 0.98 ms step (measured by leaving the
// kernel out), for a part of the text that holds 0.2 % of the windows.  But the partial words belong to the COLLECTION, not to the
// batch: they only change when sequences are uploaded or split.  So they are built once (both strands, frame letter planes + event),
// and indexed by every run of EDGE_Q = 5 single-letter slots they hold: bucket (frame position p0, code of slots p0 .. p0 + 4).
// A batch then works from the candidates' side: a candidate that must reach thr of its n slots is cut into n - thr + 1 pieces, one of
// which matches in every slot (pigeonhole, the same pieces as fst.cuh); every (candidate, piece, IUPAC expansion of the piece's first
// five bases) is ONE bucket -- ~10^3 words, verified with the usual 4 AND + POPC -- and a (word, candidate) pair found through
// several pieces is reported by the leftmost one whose seed matches in every slot only, the rule of fst_match<true>.  Words that hold a degenerate
// base have no code: they are listed apart and compared with every candidate (as before).  Work per batch: 2000 candidates x ~3.5
// pieces x ~1500 words = 10^7 verifications instead of 1.8 x 10^6 words x 40 lookups.
//
// The table serves candidates whose pieces are all at least five slots long (0.9 and up for 18..32-mers: at 0.9 a 21-mer may miss
// three slots and is cut into four pieces of five); a batch
// with any other candidate raises a flag and is run again in the general form (pcramp_gpu.cu: select_words_fast / fast_resolve),
// and that threshold keeps the scan kernel from then on.
#pragma once
#include "scan.cuh"

namespace pcr {

constexpr uint32_t EDGE_Q = 5u;                       // slots per run: the shortest piece the table serves
constexpr uint32_t EDGE_POS = 32u - EDGE_Q + 1u;     // frame positions a run of EDGE_Q slots can start at
constexpr uint32_t EDGE_BUCKETS = EDGE_POS << 12;    // x code (two planes in 6-bit fields, the layout of fst.cuh; EDGE_Q bits of each in use)
constexpr uint32_t EDGE_MAX_PIECES = 8u;             // n - thr + 1 <= 8
constexpr unsigned int EDGE_FLAG_UNSUITABLE = 8u;    // a candidate the table cannot serve (joins the fast form's flag word)

struct EdgeTable {
	const uint4 *planes;    // per word: frame letter planes A, C, G, T (bit k = slot k)
	const uint4 *meta;      // per word: {sequence, event position, event type << 1 | minus strand, 0}
	const uint32_t *start;  // EDGE_BUCKETS + 1
	const uint32_t *ids;    // word ids grouped by bucket
	const uint32_t *degen;  // words holding a degenerate base
	uint32_t n_words, n_degen;
};

template <class F>
__device__ __forceinline__ void edge_word_buckets(const FstWord &w, F f)
{
	const uint32_t qm = (1u << EDGE_Q) - 1u;
	for (uint32_t p0 = 0; p0 < EDGE_POS; ++p0)
		if (((w.single >> p0) & qm) == qm) f((p0 << 12) | ((w.b0 >> p0) & qm) | (((w.b1 >> p0) & qm) << 6));
}

// One warp per sequence, lanes = events (the loop of scan_edge_fst_kernel; inactive sequences included -- `active` is looked at when
// a word is matched).  WRITE = false: count the words, the degenerate ones and the bucket sizes; WRITE = true: write the words down.
// counters: [0] words, [1] degenerate words.
template <bool WRITE>
__global__ void __launch_bounds__(128)
edge_words_kernel(SeqDev sd, PackParams pp, unsigned int *counters, uint32_t *bucket_cnt, uint4 *planes, uint4 *meta, uint32_t *degen, uint32_t cap_words,
	uint32_t cap_degen)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t seq = warp; seq < sd.n; seq += n_warps) {
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
			for (uint32_t minus = 0; minus < 2u; ++minus) {
				const Planes4 P = w_planes(minus ? wm : wp);
				const FstWord w = fst_word(P.a, P.c, P.g, P.t);
				const uint32_t idx = atomicAdd(counters, 1u);
				if (WRITE && idx < cap_words) {
					planes[idx] = make_uint4(P.a, P.c, P.g, P.t);
					meta[idx] = make_uint4(seq, pos, (type << 1) | minus, 0u);
				}
				if (w.degenerate) {
					const uint32_t k = atomicAdd(counters + 1, 1u);
					if (WRITE && k < cap_degen) degen[k] = idx;
				} else if (!WRITE) {
					edge_word_buckets(w, [&](uint32_t b) { atomicAdd(bucket_cnt + b, 1u); });
				}
			}
		}
	}
}

// one thread per word: its id into every bucket it belongs to (cursor = bucket starts, advanced)
__global__ void edge_fill_kernel(const uint4 *__restrict__ planes, uint32_t n_words, uint32_t *cursor, uint32_t *ids, uint32_t cap_ids)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_words) return;
	const uint4 p = planes[i];
	const FstWord w = fst_word(p.x, p.y, p.z, p.w);
	if (w.degenerate) return;
	edge_word_buckets(w, [&](uint32_t b) {
		const uint32_t slot = atomicAdd(cursor + b, 1u);
		if (slot < cap_ids) ids[slot] = i;
	});
}

// can the table serve this candidate?  (every piece at least EDGE_Q slots long, its first EDGE_Q slots seedable: fst_piece_seeds' rules)
__device__ __forceinline__ bool edge_candidate_ok(const uint4 &p, uint32_t first, uint32_t len, uint32_t pieces, bool usable)
{
	if (!usable || pieces > EDGE_MAX_PIECES) return false;
	for (uint32_t k = 0; k < pieces; ++k) {
		uint32_t o, q;
		fst_piece(len, pieces, k, o, q);
		if (q < EDGE_Q || !fst_piece_seeds(p, first + o, EDGE_Q, [](uint32_t) {})) return false;
	}
	return true;
}

// Per batch.  CTA t handles (candidate t / EDGE_MAX_PIECES, piece t % EDGE_MAX_PIECES): its threads stride over the words of the
// piece's bucket(s), four words in flight per thread.  (A bucket is not a few hundred words when the collection is a handful of
// clades: the same 5-mer sits at the same slot of the same event in every member -- tens of thousands of words; one warp per bucket
// made those the kernel's tail, 1.1 ms.)  Afterwards the CTAs share out the degenerate words (threads = candidates).  Hits leave
// through the same HitSink as the other scans, with the key scan_edge_fst_kernel gives them.
constexpr uint32_t EDGE_THREADS = 128u, EDGE_UNROLL = 4u;

__global__ void __launch_bounds__(EDGE_THREADS)
edge_lookup_kernel(SeqDev sd, EdgeTable et, const uint4 *__restrict__ cand_planes, const uint32_t *__restrict__ cand_thr, uint32_t n_cand,
	uint32_t cand_bits, HitSink hs, unsigned int *flags)
{
	const uint32_t qm = (1u << EDGE_Q) - 1u;
	for (uint32_t task = blockIdx.x; task < n_cand * EDGE_MAX_PIECES; task += gridDim.x) {
		const uint32_t c = task / EDGE_MAX_PIECES, k = task % EDGE_MAX_PIECES;
		const uint4 p = __ldg(cand_planes + c);
		const uint32_t need = __ldg(cand_thr + c);
		uint32_t first, len, pieces;
		bool usable;
		fst_shape(p, need, first, len, pieces, usable);
		if ((p.x | p.y | p.z | p.w) == 0u || need > len) continue; // can never match (fst_build_kernel leaves it out as well)
		// (five of a candidate's eight CTAs have no piece: they leave here.  The CTA of piece 0 decides whether the table can serve the
		// candidate at all; the others only look at their own piece -- if another piece is the problem the batch is re-run anyway)
		if (k != 0u && (!usable || k >= pieces)) continue;
		if (k == 0u && !edge_candidate_ok(p, first, len, pieces, usable)) {
			if (threadIdx.x == 0u) atomicOr(flags, EDGE_FLAG_UNSUITABLE);
			continue;
		}
		uint32_t o, q;
		fst_piece(len, pieces, k, o, q);
		if (q < EDGE_Q) continue;
		const uint32_t p0 = first + o;
		uint32_t seed_earlier[EDGE_MAX_PIECES]; // seed masks of the pieces before this one
		for (uint32_t j = 0; j < EDGE_MAX_PIECES; ++j) {
			uint32_t oj = 0, qj = 0;
			if (j < k) fst_piece(len, pieces, j, oj, qj);
			seed_earlier[j] = j < k ? qm << (first + oj) : 0u;
		}
		fst_piece_seeds(p, p0, EDGE_Q, [&](uint32_t fb) {
			const uint32_t b = (p0 << 12) | (fb & 4095u);
			const uint32_t lo = __ldg(et.start + b), hi = __ldg(et.start + b + 1u);
			for (uint32_t e0 = lo + threadIdx.x; e0 < hi; e0 += EDGE_THREADS * EDGE_UNROLL) {
				uint32_t id[EDGE_UNROLL];
				uint4 w[EDGE_UNROLL];
				#pragma unroll
				for (uint32_t u = 0; u < EDGE_UNROLL; ++u) {
					const uint32_t e = e0 + u * EDGE_THREADS;
					id[u] = e < hi ? __ldg(et.ids + e) : 0xFFFFFFFFu;
				}
				#pragma unroll
				for (uint32_t u = 0; u < EDGE_UNROLL; ++u) w[u] = id[u] != 0xFFFFFFFFu ? __ldg(et.planes + id[u]) : make_uint4(0u, 0u, 0u, 0u);
				#pragma unroll
				for (uint32_t u = 0; u < EDGE_UNROLL; ++u) {
					const uint32_t m = (p.x & w[u].x) | (p.y & w[u].y) | (p.z & w[u].z) | (p.w & w[u].w);
					if (id[u] == 0xFFFFFFFFu || (uint32_t)__popc(m) < need) continue;
					bool earlier = false; // an earlier piece whose seed matches in every slot finds the word too, and reports it
					#pragma unroll
					for (uint32_t j = 0; j < EDGE_MAX_PIECES; ++j) earlier = earlier || (seed_earlier[j] != 0u && (m & seed_earlier[j]) == seed_earlier[j]);
					if (earlier) continue;
					const uint4 mt = __ldg(et.meta + id[u]);
					if (!sd.active[mt.x]) continue;
					hit_append(hs, hit_key_pack(mt.x, c, cand_bits, (uint32_t)__popc(m), mt.z >> 1, mt.z & 1u), mt.y);
				}
			}
		});
	}
	// words with a degenerate base: no seed code, compared with every candidate
	for (uint32_t i = blockIdx.x; i < et.n_degen; i += gridDim.x) {
		const uint32_t id = __ldg(et.degen + i);
		const uint4 mt = __ldg(et.meta + id);
		if (!sd.active[mt.x]) continue;
		const uint4 w = __ldg(et.planes + id);
		for (uint32_t c = threadIdx.x; c < n_cand; c += EDGE_THREADS) {
			const uint4 p = __ldg(cand_planes + c);
			const uint32_t need = __ldg(cand_thr + c);
			const uint32_t m = (p.x & w.x) | (p.y & w.y) | (p.z & w.z) | (p.w & w.w);
			if ((uint32_t)__popc(m) >= need) hit_append(hs, hit_key_pack(mt.x, c, cand_bits, (uint32_t)__popc(m), mt.z >> 1, mt.z & 1u), mt.y);
		}
	}
}

} // namespace pcr
