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
// the query (<= 32 bases, Word length) whatever the target length; the reference's row-major tie rule is kept
// by comparing coordinates explicitly.
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

struct Query {
	unsigned char b[SW_MAX_QUERY]; // 4-bit codes
	int len;
};

// pack_query_slots(Word) (seq_overlap.h:828-869): size() nibbles starting at start()
PCR_HD void query_from_word(const W128 &w, Query &q)
{
	q.len = w_size(w);
	const int s = w_start(w);
	for (int i = 0; i < SW_MAX_QUERY; ++i) q.b[i] = (i < q.len && s >= 0 && s + i < WORD_LEN) ? (unsigned char)w_get(w, s + i) : 0;
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

template <bool WITH_START, class Target>
PCR_HD Result align(const Query &q, const Target &t)
{
	Result r;
	r.score = 0;
	r.q_start = r.q_stop = r.t_start = r.t_stop = 0;
	r.any = false;
	const int qlen = q.len, tlen = t.length();
	// state of column j-1, indexed by query position
	int M[SW_MAX_QUERY], Iq[SW_MAX_QUERY], It[SW_MAX_QUERY];
	int Msi[WITH_START ? SW_MAX_QUERY : 1], Msj[WITH_START ? SW_MAX_QUERY : 1];   // start of the path into M
	int Qsi[WITH_START ? SW_MAX_QUERY : 1], Qsj[WITH_START ? SW_MAX_QUERY : 1];   // ... into Iq
	int Tsi[WITH_START ? SW_MAX_QUERY : 1], Tsj[WITH_START ? SW_MAX_QUERY : 1];   // ... into It
	for (int i = 0; i < qlen; ++i) { // column -1: the border (curr_row[0], :400-408)
		M[i] = 0;
		Iq[i] = It[i] = SW_GAP_OPEN;
		if (WITH_START) {
			Msi[i] = i + 1;
			Msj[i] = 0;
			Qsi[i] = Qsj[i] = Tsi[i] = Tsj[i] = 0; // never observable: only ever carried by negative scores
		}
	}
	int best = 0, best_i = -1, best_j = -1, best_si = 0, best_sj = 0;
	for (int j = 0; j < tlen; ++j) {
		const unsigned tb = t.at(j);
		// cell (i-1, j-1) and (i-1, j) for i = 0: the top border (last_row, :381-389): M = 0, gaps = -5, start = (0, j)
		int aM = 0, aIq = SW_GAP_OPEN, aIt = SW_GAP_OPEN, aMsi = 0, aMsj = j, aQsi = 0, aQsj = 0, aTsi = 0, aTsj = 0;
		int bM = 0, bIt = SW_GAP_OPEN, bMsi = 0, bMsj = j + 1, bTsi = 0, bTsj = 0;
		for (int i = 0; i < qlen; ++i) {
			// C = (i, j-1) is the stored state; it becomes A for the next i
			const int cM = M[i], cIq = Iq[i], cIt = It[i];
			int cMsi = 0, cMsj = 0, cQsi = 0, cQsj = 0, cTsi = 0, cTsj = 0;
			if (WITH_START) {
				cMsi = Msi[i]; cMsj = Msj[i]; cQsi = Qsi[i]; cQsj = Qsj[i]; cTsi = Tsi[i]; cTsj = Tsj[i];
			}
			const int amax = max(max(aM, aIq), aIt);
			const int xM = max(amax, 0) + ((q.b[i] & tb) ? SW_MATCH : SW_MISMATCH);
			int xMsi = 0, xMsj = 0;
			if (WITH_START) { // :457-512, in the reference's order of overrides
				const bool gap_beats_m = (aM < aIq) || (aM < aIt);
				xMsi = gap_beats_m ? 0 : aMsi;
				xMsj = gap_beats_m ? 0 : aMsj;
				if (!(aIq < aIt) && (aIq > aM)) { xMsi = aQsi; xMsj = aQsj; }
				if ((aIt > aM) && (aIt > aIq)) { xMsi = aTsi; xMsj = aTsj; }
				if (0 > amax) { xMsi = i; xMsj = j; }
			}
			const int q_open = max(cM, 0) + SW_GAP_OPEN, q_ext = max(cIq, 0) + SW_GAP_EXTEND;
			const int xIq = max(q_open, q_ext);
			const int t_open = max(bM, 0) + SW_GAP_OPEN, t_ext = max(bIt, 0) + SW_GAP_EXTEND;
			const int xIt = max(t_open, t_ext);
			int xQsi = 0, xQsj = 0, xTsi = 0, xTsj = 0;
			if (WITH_START) {
				const bool qe = q_open < q_ext;
				xQsi = qe ? cQsi : cMsi;
				xQsj = qe ? cQsj : cMsj;
				const bool te = t_open < t_ext;
				xTsi = te ? bTsi : bMsi;
				xTsj = te ? bTsj : bMsj;
			}
			// the reference replaces the maximum whenever M is not less than it, scanning i outer / j inner:
			// the winner is the cell with the largest (i, j) among those equal to the final maximum
			if (xM > best || (xM == best && (i > best_i || (i == best_i && j > best_j)))) {
				best = xM;
				best_i = i;
				best_j = j;
				if (WITH_START) { best_si = xMsi; best_sj = xMsj; }
			}
			// slide: this column's (i) becomes B for i+1; the stored previous column's (i) becomes A for i+1
			aM = cM; aIq = cIq; aIt = cIt;
			if (WITH_START) { aMsi = cMsi; aMsj = cMsj; aQsi = cQsi; aQsj = cQsj; aTsi = cTsi; aTsj = cTsj; }
			bM = xM; bIt = xIt;
			if (WITH_START) { bMsi = xMsi; bMsj = xMsj; bTsi = xTsi; bTsj = xTsj; }
			M[i] = xM; Iq[i] = xIq; It[i] = xIt;
			if (WITH_START) { Msi[i] = xMsi; Msj[i] = xMsj; Qsi[i] = xQsi; Qsj[i] = xQsj; Tsi[i] = xTsi; Tsj[i] = xTsj; }
		}
	}
	if (best_i >= 0) {
		r.any = true;
		r.score = best;
		r.q_stop = best_i;
		r.t_stop = best_j;
		r.q_start = best_si;
		r.t_start = best_sj;
	}
	return r;
}

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
