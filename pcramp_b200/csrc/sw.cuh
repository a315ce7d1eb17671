// sw.cuh -- K4 core: the nucleic-acid Smith-Waterman of SO::SeqOverlap (seq_overlap.cpp:347-609) for ONE problem.
//
// The reference runs 8 independent problems per SSE register (int16 lanes, seq_overlap.h:58-83); slots never
// interact (cells past a slot's own lengths are masked out of the maximum, :571-573), so a problem is defined
// by its query and target alone:
//   score(q, t) = +2 if the 4-bit IUPAC sets intersect, else -3 (:419-424; mask_N_na is off)
//   M(i,j)  = max(max(M, Iq, It)(i-1,j-1), 0) + score              (:431-434)
//   Iq(i,j) = max(max(M(i,j-1), 0) - 5, max(Iq(i,j-1), 0) - 2)     gap in the query   (:518-522)
//   It(i,j) = max(max(M(i-1,j), 0) - 5, max(It(i-1,j), 0) - 2)     gap in the target  (:541-545)
//   borders: M = 0, Iq = It = -5                                    (:381-389, :400-408)
//   result  = the LAST cell in (i outer, j inner) order with M >= every earlier maximum, starting from a
//             maximum of 0 (:575-604: "not less than" replaces), its coordinates (stop_i, stop_j), and the start
//             of the path that reaches it (M_start_*: propagated through the three states, :457-512).
// Scores stay within [-69, 64] for 32-base operands, so the reference's int16 never wraps; int here.
//
// The DP is walked column by column (target outer, query inner) so that the per-problem state is indexed by
// the query (<= 32 bases, Word length) whatever the target length, and kept in registers (see align_score / align_start);
// the reference's row-major tie rule is kept by comparing (score, i) keys.
#pragma once
#include "word128.cuh"

namespace pcr {
namespace sw {

constexpr int SW_MATCH = 2, SW_MISMATCH = -3, SW_GAP_OPEN = -5, SW_GAP_EXTEND = -2; // seq_overlap.cpp:60-66 (blastn defaults)
constexpr int SW_MAX_QUERY = 32;

struct Result {
	int score;                     // SeqOverlap::score
	int q_start, q_stop;           // alignment_range_query  (seq_overlap.h:1290-1301)
	int t_start, t_stop;           // alignment_range_target (:1304-1314)
	bool any;                      // false: no cell reached the initial maximum (0); the reference then reports stale coordinates
};

struct Query { // the query as four bit-planes over its positions: bit i of plane[k] = bit k of the 4-bit code of base i
	uint32_t plane[4];
	int len;
};

// pack_query_slots(Word) (seq_overlap.h:828-869): size() nibbles starting at start()
PCR_HD void query_from_word(const W128 &w, Query &q)
{
	q.len = w_size(w);
	const int s = w_start(w);
	q.plane[0] = q.plane[1] = q.plane[2] = q.plane[3] = 0u;
	for (int i = 0; i < SW_MAX_QUERY; ++i) {
		const uint32_t b = (i < q.len && s >= 0 && s + i < WORD_LEN) ? w_get(w, s + i) : 0u;
		q.plane[0] |= (b & 1u) << i;
		q.plane[1] |= ((b >> 1) & 1u) << i;
		q.plane[2] |= ((b >> 2) & 1u) << i;
		q.plane[3] |= ((b >> 3) & 1u) << i;
	}
}

// bit i = "query base i and the target code tb share a letter" (the +2 / -3 choice of :419-424 for a whole column)
PCR_HD uint32_t match_mask(const Query &q, unsigned tb)
{
	return ((tb & 1u) ? q.plane[0] : 0u) | ((tb & 2u) ? q.plane[1] : 0u) | ((tb & 4u) ? q.plane[2] : 0u) | ((tb & 8u) ? q.plane[3] : 0u);
}

struct WordTarget { // pack_target_slots(Word) (seq_overlap.h:1102-1136)
	W128 w;
	int first, len;
	PCR_HD explicit WordTarget(const W128 &x) : w(x), first(w_start(x)), len(w_size(x)) {}
	PCR_HD int length() const { return len; }
	PCR_HD unsigned at(int j) const { return (first + j < WORD_LEN) ? w_get(w, first + j) : 0u; }
};

struct NibbleTarget { // pack_target_slots(Sequence) (seq_overlap.h:1071-1100): every nibble of the sequence, EOS (0) included
	const unsigned char *raw;
	int len;
	PCR_HD NibbleTarget(const unsigned char *r, int l) : raw(r), len(l) {}
	PCR_HD int length() const { return len; }
	PCR_HD unsigned at(int j) const
	{
		const unsigned v = raw[j >> 1];
		return (j & 1) ? (v & 15u) : (v >> 4);
	}
};

// three-input integer maximum / fused add-maximum: one instruction each on sm_100a (VIMNMX3 / VIADDMNMX)
PCR_HD int max3i(int a, int b, int c)
{
#ifdef __CUDA_ARCH__
	return __vimax3_s32(a, b, c);
#else
	return max(max(a, b), c);
#endif
}
PCR_HD int max3i_relu(int a, int b, int c)
{
#ifdef __CUDA_ARCH__
	return __vimax3_s32_relu(a, b, c);
#else
	return max(max(max(a, b), c), 0);
#endif
}
PCR_HD int addmaxi(int a, int b, int c) // max(a + b, c)
{
#ifdef __CUDA_ARCH__
	return __viaddmax_s32(a, b, c);
#else
	return max(a + b, c);
#endif
}

// The state of column j-1 lives in registers: the loop over the query is fully unrolled (32 statically indexed slots,
// left early at the query's end) and only the loop over the target is a run-time loop.
//
// Score only (any target length):
//   M  = max(aM, aIq, aIt, 0) + s                 one VIMNMX3.RELU + the +2 / -3 choice
//   Iq = max(cM - 5, cIq - 2, -2)                 (max(x, 0) - 5 = max(x - 5, -5) and -5 < -2): two VIADDMNMX
//   It = max(bM - 5, bIt - 2, -2)
// The winner -- the largest (i, j) among the cells equal to the maximum -- is kept as one key (score, i): columns are
// visited in increasing j, so among equal scores a later cell replaces the kept one exactly when its i is not smaller.
//
// ROWS = the number of query slots computed (a compile-time count >= the query's length; kernels pick the smallest
// instantiation that covers the longest query of the warp).  Slots past the end of the query need no mask: their plane
// bits are 0, so they always score a mismatch, and a cell there stays at least 3 below the best real cell seen so far.
template <int ROWS, class Target>
PCR_HD Result align_score(const Query &q, const Target &t)
{
	Result r;
	r.score = 0;
	r.q_start = r.q_stop = r.t_start = r.t_stop = 0;
	r.any = false;
	const int tlen = t.length();
	int M[ROWS], Iq[ROWS], It[ROWS];
#pragma unroll
	for (int i = 0; i < ROWS; ++i) { // column -1: the border (curr_row[0], :400-408)
		M[i] = 0;
		Iq[i] = It[i] = SW_GAP_OPEN;
	}
	int best_key = 0, best_j = -1; // key = score * 64 + i
	for (int j = 0; j < tlen; ++j) {
		const uint32_t m = match_mask(q, t.at(j));
		int aM = 0, aIq = SW_GAP_OPEN, aIt = SW_GAP_OPEN; // (i-1, j-1) for i = 0: the top border (last_row, :381-389)
		int bM = 0, bIt = SW_GAP_OPEN;                    // (i-1, j)
#pragma unroll
		for (int i = 0; i < ROWS; ++i) {
			const int cM = M[i], cIq = Iq[i], cIt = It[i]; // (i, j-1)
			const int xM = max3i_relu(aM, aIq, aIt) + (((m >> i) & 1u) ? SW_MATCH : SW_MISMATCH);
			const int xIq = addmaxi(cM, SW_GAP_OPEN, addmaxi(cIq, SW_GAP_EXTEND, SW_GAP_EXTEND));
			const int xIt = addmaxi(bM, SW_GAP_OPEN, addmaxi(bIt, SW_GAP_EXTEND, SW_GAP_EXTEND));
			const int key = xM * 64 + i;
			if (key >= best_key) { best_key = key; best_j = j; }
			aM = cM; aIq = cIq; aIt = cIt;
			bM = xM; bIt = xIt;
			M[i] = xM; Iq[i] = xIq; It[i] = xIt;
		}
	}
	if (best_j >= 0) {
		r.any = true;
		r.score = best_key >> 6;
		r.q_stop = best_key & 63;
		r.t_stop = best_j;
	}
	return r;
}

// With start coordinates (operands of at most 32 bases: Word against Word).  Every state is ONE 32-bit word
//   [31:22] score   [21:16] row + 1   [13:12] state (M = 3, Iq = 1, It = 0: re-labelling a maximum is one OR / AND)   [11:6] 63 - start_i   [5:0] 63 - start_j
// (the words that meet in a maximum belong to the same row, so the row field never decides there; it rides along -- the constants that
// move a word one row down add 1 to it -- so that the winner, the largest (score, row) and among equals the latest, is one compare)
// and the reference's propagation rules (:457-512) are what a plain signed maximum of such words does:
//   * M takes the start of M when M is not beaten, else of Iq when Iq >= It, else of It: ties on the score fall to the
//     larger state code;
//   * a gap state takes the start of its opening M unless the extension scores strictly more: M (3) wins ties against
//     Iq (1) and It (0) (the label left on a negative word is irrelevant: it only orders words nobody reports);
//   * max(x, 0) ahead of a gap: a gap state that only exists through the clamp has a negative score, and a negative
//     state can never hand its start to a cell that is reported (it loses to the clamp or to the fresh start below),
//     so the clamp is the constant word "-2" in the maximum;
//   * the fresh start (i, j) is taken when the incoming maximum is negative (`0 > amax` strictly, :502-512): a path
//     that comes in with score 0 started at some i' < i, so its inverted start field is larger than the fresh one and a
//     maximum against the word (0, M, i, j) keeps it.  The borders carry score 0 with the start the fresh rule would give.
constexpr int SWP_ONE = 1 << 22, SWP_ROW = 1 << 16, SWP_STATE_M = 3 << 12, SWP_STATE_IQ = 1 << 12;
PCR_HD int swp_pack(int score, int row_field, int state, int si, int sj)
{
	return score * SWP_ONE + row_field * SWP_ROW + state + ((63 - si) << 6) + (63 - sj);
}

template <int ROWS, class Target>
PCR_HD Result align_start(const Query &q, const Target &t)
{
	Result r;
	r.score = 0;
	r.q_start = r.q_stop = r.t_start = r.t_stop = 0;
	r.any = false;
	const int tlen = t.length();
	int M[ROWS], Iq[ROWS], It[ROWS];
	const int border_gap = swp_pack(SW_GAP_OPEN, 0, 0, 63, 63), floor_gap = swp_pack(SW_GAP_EXTEND, 0, 0, 63, 63);
#pragma unroll
	for (int i = 0; i < ROWS; ++i) { // column -1 (:400-408): M = 0 with start (i + 1, 0)
		M[i] = swp_pack(0, i + 1, SWP_STATE_M, i + 1, 0);
		Iq[i] = It[i] = border_gap + (i + 1) * SWP_ROW;
	}
	int best_w = 0, best_j = -1; // best_w: the winning M word -- its (score, row) decide, its start is the answer
	for (int j = 0; j < tlen; ++j) {
		const uint32_t m = match_mask(q, t.at(j));
		// the top border (:381-389), row field 0: M = 0 with start (0, j) on the diagonal side and (0, j + 1) above the column
		int aM = swp_pack(0, 0, SWP_STATE_M, 0, j), aIq = border_gap, aIt = border_gap;
		int bM = swp_pack(0, 0, SWP_STATE_M, 0, j + 1), bIt = border_gap;
		const int fresh0 = swp_pack(0, 0, SWP_STATE_M, 0, j);
#pragma unroll
		for (int i = 0; i < ROWS; ++i) {
			const int cM = M[i], cIq = Iq[i], cIt = It[i];
			const int in = max3i(aM, aIq, aIt) | SWP_STATE_M; // row field i (the cell above-left), as the fresh word's
			const int xM = max(in, fresh0 + i * SWP_ROW - (i << 6)) +
			               (((m >> i) & 1u) ? SW_MATCH * SWP_ONE + SWP_ROW : SW_MISMATCH * SWP_ONE + SWP_ROW);
			const int floor_i = floor_gap + (i + 1) * SWP_ROW;
			const int xIq = (addmaxi(cM, SW_GAP_OPEN * SWP_ONE, addmaxi(cIq, SW_GAP_EXTEND * SWP_ONE, floor_i)) & ~(SWP_STATE_M ^ SWP_STATE_IQ));
			const int xIt = addmaxi(bM, SW_GAP_OPEN * SWP_ONE + SWP_ROW, addmaxi(bIt, SW_GAP_EXTEND * SWP_ONE + SWP_ROW, floor_i)) & ~SWP_STATE_M;
			if ((xM | (SWP_ROW - 1)) >= best_w) { best_w = xM; best_j = j; } // (score, row) not smaller: a later cell replaces
			aM = cM; aIq = cIq; aIt = cIt;
			bM = xM; bIt = xIt;
			M[i] = xM; Iq[i] = xIq; It[i] = xIt;
		}
	}
	if (best_j >= 0) {
		r.any = true;
		r.score = best_w >> 22;
		r.q_stop = ((best_w >> 16) & 63) - 1;
		r.t_stop = best_j;
		r.q_start = 63 - ((best_w >> 6) & 63);
		r.t_start = 63 - (best_w & 63);
	}
	return r;
}

template <bool WITH_START, class Target, int ROWS = SW_MAX_QUERY>
PCR_HD Result align(const Query &q, const Target &t)
{
	if constexpr (WITH_START) return align_start<ROWS>(q, t);
	else return align_score<ROWS>(q, t);
}

// the same with ROWS chosen at run time (a warp-uniform value keeps the warp on one instantiation)
template <bool WITH_START, class Target>
PCR_HD Result align_rows(const Query &q, const Target &t, int rows)
{
	if (rows <= 20) return align<WITH_START, Target, 20>(q, t);
	if (rows <= 24) return align<WITH_START, Target, 24>(q, t);
	if (rows <= 28) return align<WITH_START, Target, 28>(q, t);
	return align<WITH_START, Target, 32>(q, t);
}

#ifdef __CUDACC__
// align_rows with the row count of the longest query among the lanes that are here together
template <bool WITH_START, class Target>
__device__ __forceinline__ Result align_warp(const Query &q, const Target &t)
{
	const int rows = __reduce_max_sync(__activemask(), q.len);
	return align_rows<WITH_START>(q, t, rows);
}
#endif

#ifdef __CUDACC__
// Two score-only alignments against the SAME target in one pass: the two problems ride in the 16-bit halves of one word
// (a primer and its reverse complement against a database word -- find_background_match's slot pairs, background_match.cpp:66-118).
// The recurrences are align_score's, every operation the two-lane form of the same instruction (VIMNMX3.S16X2.RELU, VIADDMNMX.S16X2;
// an add is "add-maximum against -32768"): 11 instructions per PAIR of cells.  Scores only: no coordinates, so no TaqMAMA bases.
template <int ROWS, class Target>
__device__ __forceinline__ void align_score_pair_rows(const Query &qa, const Query &qb, const Target &t, int &score_a, int &score_b)
{
	constexpr unsigned GO2 = 0xFFFBFFFBu, GE2 = 0xFFFEFFFEu, LOW2 = 0x80008000u; // (-5, -5), (-2, -2), (-32768, -32768)
	const int tlen = t.length();
	unsigned M[ROWS], Iq[ROWS], It[ROWS];
#pragma unroll
	for (int i = 0; i < ROWS; ++i) {
		M[i] = 0u;
		Iq[i] = It[i] = GO2;
	}
	unsigned best = 0u;
	for (int j = 0; j < tlen; ++j) {
		const unsigned tb = t.at(j);
		const uint32_t ma = match_mask(qa, tb), mb = match_mask(qb, tb);
		unsigned aM = 0u, aIq = GO2, aIt = GO2, bM = 0u, bIt = GO2;
#pragma unroll
		for (int i = 0; i < ROWS; ++i) {
			const unsigned cM = M[i], cIq = Iq[i], cIt = It[i];
			const unsigned s2 = (((ma >> i) & 1u) ? 0x00000002u : 0x0000FFFDu) | (((mb >> i) & 1u) ? 0x00020000u : 0xFFFD0000u);
			const unsigned xM = __viaddmax_s16x2(__vimax3_s16x2_relu(aM, aIq, aIt), s2, LOW2);
			const unsigned xIq = __viaddmax_s16x2(cM, GO2, __viaddmax_s16x2(cIq, GE2, GE2));
			const unsigned xIt = __viaddmax_s16x2(bM, GO2, __viaddmax_s16x2(bIt, GE2, GE2));
			best = __vimax_s16x2_relu(best, xM);
			aM = cM; aIq = cIq; aIt = cIt;
			bM = xM; bIt = xIt;
			M[i] = xM; Iq[i] = xIq; It[i] = xIt;
		}
	}
	score_a = (int)(short)(best & 0xFFFFu);
	score_b = (int)(short)(best >> 16);
}

template <class Target>
__device__ __forceinline__ void align_score_pair_warp(const Query &qa, const Query &qb, const Target &t, int &score_a, int &score_b)
{
	const int rows = __reduce_max_sync(__activemask(), max(qa.len, qb.len));
	if (rows <= 20) align_score_pair_rows<20>(qa, qb, t, score_a, score_b);
	else if (rows <= 24) align_score_pair_rows<24>(qa, qb, t, score_a, score_b);
	else if (rows <= 28) align_score_pair_rows<28>(qa, qb, t, score_a, score_b);
	else align_score_pair_rows<32>(qa, qb, t, score_a, score_b);
}
#endif

// target_last_two_aligned (seq_overlap.h:1265-1283): {N, N} unless 1 <= stop_j < target length
template <class Target>
PCR_HD void last_two(const Result &r, const Target &t, unsigned &first, unsigned &second)
{
	first = second = 15u;
	if (!r.any || r.t_stop < 1 || r.t_stop >= t.length()) return;
	first = t.at(r.t_stop - 1);
	second = t.at(r.t_stop);
}

} // namespace sw
} // namespace pcr
