// nuccruc.cuh -- K3: the SantaLucia nearest-neighbour engine ("NucCruc v5.3") that PCR::is_valid,
// PCR::max_dimer_tm and PCR::multiplex_compatible call (valid_pcr.cpp:5-45, pcr_assay.cpp:232-269,815-852).
//
// One problem = one oligo (hairpin, homodimer, perfect-match duplex) or one oligo pair (heterodimer).
// The whole evaluation of a problem runs in ONE thread: an integer dynamic-programming fill of
// dG x 10^4 scores with co-optimal trace bits (nuc_cruc.cpp:347-816), the enumeration of up to 17
// co-optimal paths and their zero-score truncations (:818-1471) and the floating-point evaluation of each
// candidate alignment (:1473-2232); the alignment with the lowest dG = dH - T dS wins.  The enumeration is
// sequential and branchy by nature (a trace stack that persists across paths), so the parallelism is
// across problems, 32 per warp; the DP matrix (trace bits + match score) lives in per-thread local memory.
//
// Everything is __host__ __device__ so that tests/host_thermo_harness.cpp can run the same code on the CPU
// against the compiled reference during development; the product only ever calls the CUDA kernels.
//
// Exactness: parameter tables are the reference's own floats (santalucia_tables.inc, generated); the two
// logarithms (ln[Na+], ln Ct) are taken on the host with the same libm call the reference makes and passed
// in; all float arithmetic is in the reference's order with FMA contraction off.
//
// One reference defect is pinned rather than reproduced at random: trace_back() pushes the pair of the cell
// that ENDS a path, and when that cell is in row 0 it reads query[query_len] -- one element past the query in
// an unchecked ring buffer (circle_buffer.h:136-139), i.e. whatever an earlier, longer query left there.  We
// define that slot (and the one after it) as base A, the content of a zero-initialised buffer; oracle/ref_driver.cpp
// fills the reference's buffers the same way before every call so that goldens are reproducible.
#pragma once
#include <math.h>
#include <string.h>
#include "word128.cuh"

namespace pcr {
namespace nc {

enum : int { bA = 0, bC = 1, bG = 2, bT = 3, bI = 4, bE = 5, bGAP = 6, NBASE = 7, NPAIR = 49 };
enum : int { TR_DIAG = 1, TR_UP = 2, TR_LEFT = 4, TR_INVALID = 8 }; // im1_jm1, im1_j, i_jm1 (nuc_cruc.h:120-123)
enum : int { MODE_HOMO = 0, MODE_HETERO = 1, MODE_HAIRPIN = 2 };
enum : int { OP_PM_DUPLEX = 0, OP_HAIRPIN = 1, OP_HOMODIMER = 2, OP_HETERODIMER = 3, OP_HETERODIMER_DIAG = 4, OP_HOMODIMER_DIAG = 5, OP_COUNT = 6 };
enum : int { P_AT = 3, P_TA = 21, P_CG = 9, P_GC = 15, P_GT = 17, P_TG = 23, P_NONE = 48 };
enum : int { SUPP_LOOP_H = 0, SUPP_LOOP_S, SUPP_BULGE_H, SUPP_BULGE_S, SUPP_TM_AT_H, SUPP_TM_AT_S, SUPP_TM_GC_H, SUPP_TM_GC_S, SUPP_TM_I_H,
	SUPP_TM_I_S, SUPP_TMM_H, SUPP_TMM_S };
constexpr int NC_MAX_LEN = 32;              // WORD_LENGTH: the longest oligo pcramp can hold (options.cpp:854-860)
constexpr int NC_STRIDE = 40;               // row pitch of the DP matrix (cell id = i * NC_STRIDE + j, 0 <= i, j <= 32): a multiple of 8 so
                                            // that the eight cells a strip writes per row are one aligned 16-byte store
constexpr int NC_INFO_PAD = 7;              // Ctx::info = storage + NC_INFO_PAD: column 1 of every row lands on a 16-byte boundary
constexpr int NC_CELLS = (NC_MAX_LEN + 1) * NC_STRIDE + 8; // storage, in 16-bit words (16-byte aligned)
#ifndef NC_STRIP_W
#define NC_STRIP_W 8
#endif
constexpr int NC_STRIP = NC_STRIP_W;        // columns per register strip of the DP fill (8 or 16)
constexpr int NC_SEQ_CAP = NC_MAX_LEN + 4;  // room for the stale slots past the end
constexpr int NC_ALN_CAP = 96;
constexpr int NC_ALN_HEAD = 8;
constexpr int NC_STACK_CAP = 80;
constexpr int NC_MAX_PATH_ENUM = 16;        // max_dp_path_enum (nuc_cruc.cpp:130)

PCR_HD int bpair(int x, int y) { return x * NBASE + y; }

struct Tables { // the reference's parameter set, as floats (santalucia_tables.inc)
	float H[NPAIR * NPAIR], S[NPAIR * NPAIR];
	float loop_term_H[NPAIR * NPAIR], loop_term_S[NPAIR * NPAIR];
	float hairpin_term_H[NPAIR * NPAIR], hairpin_term_S[NPAIR * NPAIR];
	float loop_S[129], bulge_S[129], hairpin_S[129];
	float special_H[131], special_S[131];
	float supp[12], supp_salt[4];
	float init_H, init_S, asym_loop_dS, bulge_AT_closing_S, AT_closing_H, AT_closing_S, symmetry_S, SALT;
	unsigned char wc[NPAIR];
	char special_loop[131][7];
};

struct DpTable { // update_dp_param (nuc_cruc.cpp:191-342) for one salt concentration at 37 C
	int dg[NPAIR * NPAIR];
	float target_T; // 310.15
	float log_na;   // logf([Na+]), host libm
};

struct Aln { // struct alignment (nuc_cruc.h:317-412): two parallel base strings with deque semantics
	unsigned char q[NC_ALN_CAP], t[NC_ALN_CAP];
	int head, n;
	int fm_first, fm_second, lm_first, lm_second;
	float dH, dS, tm;
	bool valid;
};

PCR_HD void aln_clear(Aln &a)
{
	a.head = NC_ALN_HEAD;
	a.n = 0;
	a.fm_first = a.fm_second = a.lm_first = a.lm_second = 0;
	a.dH = a.dS = a.tm = 0.0f;
	a.valid = false;
}
PCR_HD void aln_push_back(Aln &a, int qb, int tb)
{
	if (a.head + a.n < NC_ALN_CAP) {
		a.q[a.head + a.n] = (unsigned char)qb;
		a.t[a.head + a.n] = (unsigned char)tb;
		++a.n;
	}
}
PCR_HD int aln_q(const Aln &a, int i) { return a.q[a.head + i]; }
PCR_HD int aln_t(const Aln &a, int i) { return a.t[a.head + i]; }

struct Ctx {
	const Tables *T;
	const DpTable *D;
	const unsigned char *q, *t; // 5'-3' bases, NC_SEQ_CAP entries each, slots past the length hold bA
	int qlen, tlen;
	float log_strand; // logf(strand concentration), host libm
	// DP storage: one 16-bit word per cell is all the enumeration needs --
	//   M_trace[3:0] | Iq_trace[7:4] | It_trace[11:8] | Iq<0 [12] | It<0 [13] | M<0 [14] | M==0 [15]
	// (the scores themselves only matter for "is this a maximal cell", kept as a short list, see dp_fill)
	unsigned short *info;   // storage (NC_CELLS words, 16-byte aligned) + NC_INFO_PAD
	int max_cell[8];        // the first NC_MAX_CELLS cells whose M equals the maximum, in row-major order
	int n_max_cell;         // how many cells equal the maximum (may exceed NC_MAX_CELLS)
};
constexpr int NC_MAX_CELLS = 8;

PCR_HD int seq_at(const unsigned char *s, int i) { return (i >= 0 && i < NC_SEQ_CAP) ? s[i] : bA; }

// ---------------------------------------------------------------------------------------------
// DP fill (nuc_cruc.cpp:347-541 dimer, :546-612 diagonal, :616-816 hairpin)
// ---------------------------------------------------------------------------------------------
PCR_HD int dp_step(int prev_score, int dg) { return (0 < prev_score) ? prev_score - dg : -dg; }

// one interior cell; A = (i-1, j-1), B = (i-1, j), C = (i, j-1)
PCR_HD void dp_cell(const DpTable *D, int tb, int ptb, int qb, int pqb, int aM, int aIq, int aIt, int bM, int bIt, int cM, int cIq, int &xM,
	int &xIq, int &xIt, unsigned short &info)
{
	int cur = bpair(tb, qb);
	const int dg1 = dp_step(aM, D->dg[bpair(ptb, pqb) * NPAIR + cur]);
	const int dg2 = dp_step(aIq, D->dg[bpair(ptb, bGAP) * NPAIR + cur]);
	const int dg3 = dp_step(aIt, D->dg[bpair(bGAP, pqb) * NPAIR + cur]);
	int mtr;
	if (dg1 >= dg2) {
		if (dg1 >= dg3) {
			xM = dg1;
			mtr = TR_DIAG;
			if (dg1 == dg2) mtr |= TR_LEFT;
			if (dg1 == dg3) mtr |= TR_UP;
		} else {
			xM = dg3;
			mtr = TR_UP;
		}
	} else {
		if (dg2 >= dg3) {
			xM = dg2;
			mtr = TR_LEFT;
			if (dg2 == dg3) mtr |= TR_UP;
		} else {
			xM = dg3;
			mtr = TR_UP;
		}
	}
	cur = bpair(tb, bGAP);
	int ins = dp_step(cM, D->dg[bpair(ptb, qb) * NPAIR + cur]);
	int ext = dp_step(cIq, D->dg[bpair(ptb, bGAP) * NPAIR + cur]);
	int qtr;
	if (ins >= ext) {
		xIq = ins;
		qtr = TR_DIAG;
		if (ins == ext) qtr |= TR_LEFT;
	} else {
		xIq = ext;
		qtr = TR_LEFT;
	}
	cur = bpair(bGAP, qb);
	ins = dp_step(bM, D->dg[bpair(tb, pqb) * NPAIR + cur]);
	ext = dp_step(bIt, D->dg[bpair(bGAP, pqb) * NPAIR + cur]);
	int ttr;
	if (ins >= ext) {
		xIt = ins;
		ttr = TR_DIAG;
		if (ins == ext) ttr |= TR_UP;
	} else {
		xIt = ext;
		ttr = TR_UP;
	}
	info = (unsigned short)(mtr | (qtr << 4) | (ttr << 8) | ((xIq < 0) ? 0x1000 : 0) | ((xIt < 0) ? 0x2000 : 0) | ((xM < 0) ? 0x4000 : 0) |
	                        ((xM == 0) ? 0x8000 : 0));
}

PCR_HD void dp_border(Ctx &c)
{ // NC_Elem() : scores -1, traces invalid (nuc_cruc.h:427-433); row 0 and column 0 are never written
	const unsigned short b = (unsigned short)(TR_INVALID | (TR_INVALID << 4) | (TR_INVALID << 8) | 0x7000);
	for (int k = 0; k <= NC_MAX_LEN; ++k) {
		c.info[k] = b;
		c.info[k * NC_STRIDE] = b;
	}
}

// max_ptr bookkeeping (nuc_cruc.cpp:517-537): cells that equal the running maximum, reset when it rises
PCR_HD void note_cell(Ctx &c, int cell, int xM, int &max_score)
{
	if (xM > max_score) {
		max_score = xM;
		c.n_max_cell = 1;
		c.max_cell[0] = cell;
	} else if (xM == max_score) {
		if (c.n_max_cell < NC_MAX_CELLS) c.max_cell[c.n_max_cell] = cell;
		++c.n_max_cell;
	}
}

struct Aln;
PCR_HD void enumerate_dimer(Ctx &c, int cell, Aln &best, int mode);
PCR_HD void enumerate_hairpin(Ctx &c, int cell, Aln &best);

// gapped dimer (hairpin = false) or hairpin triangle (hairpin = true, target = query); returns the maximum score.
// One row of (M, Iq, It) is kept: cell (i-1, j) is read from it and overwritten by (i, j); (i-1, j-1) and (i, j-1) ride
// along in registers.  REPLAY = true is the rare second pass for problems with more than NC_MAX_CELLS maximal cells: the
// trace bits are complete, the scores are recomputed and every cell equal to `replay_max` is enumerated in row-major order.
template <bool REPLAY>
PCR_HD int dp_fill_t(Ctx &c, bool hairpin, long long *cells_out, int replay_max, Aln *best, int mode)
{
	if (!REPLAY) {
		dp_border(c);
		c.n_max_cell = 0;
	}
	const int qlen = c.qlen, tlen = hairpin ? c.qlen : c.tlen;
	const unsigned char *tq = hairpin ? c.q : c.t;
	const int max_stem = qlen - 4; // steric limit 3 + 1 (nuc_cruc.cpp:627-635)
	const int rows = hairpin ? max_stem : qlen;
	int rM[NC_STRIDE], rIq[NC_STRIDE], rIt[NC_STRIDE];
	for (int j = 0; j < NC_STRIDE; ++j) rM[j] = rIq[j] = rIt[j] = -1;
	int max_score = -1;
	long long cells = 0;
	for (int i = 1; i <= rows; ++i) {
		const int qb = seq_at(c.q, qlen - i);
		const int pqb = (i == 1) ? bGAP : seq_at(c.q, qlen - (i - 1));
		const int cols = hairpin ? (max_stem - (i - 1)) : tlen;
		int aM = -1, aIq = -1, aIt = -1; // (i-1, j-1): column 0 is the border
		int cM = -1, cIq = -1;           // (i, j-1)
		int ptb = bGAP;
		for (int j = 1; j <= cols; ++j) {
			const int tb = seq_at(tq, j - 1);
			const int bM = rM[j], bIq = rIq[j], bIt = rIt[j]; // (i-1, j)
			int xM, xIq, xIt;
			unsigned short inf;
			dp_cell(c.D, tb, ptb, qb, pqb, aM, aIq, aIt, bM, bIt, cM, cIq, xM, xIq, xIt, inf);
			rM[j] = xM;
			rIq[j] = xIq;
			rIt[j] = xIt;
			const int cell = i * NC_STRIDE + j;
			if (!REPLAY) {
				c.info[cell] = inf;
				note_cell(c, cell, xM, max_score);
			} else if (xM == replay_max) {
				if (hairpin) enumerate_hairpin(c, cell, *best);
				else enumerate_dimer(c, cell, *best, mode);
			}
			aM = bM; aIq = bIq; aIt = bIt;
			cM = xM; cIq = xIq;
			ptb = tb;
		}
		cells += cols;
	}
	if (cells_out) *cells_out = cells;
	return REPLAY ? replay_max : max_score;
}

// The same fill (first pass only), strip-mined for a thread that owns the whole problem: NC_STRIP columns at a time, all
// rows.  The previous row of the strip (clamped M, Iq, It per column), the column constants and the running (i, j-1) cell
// stay in registers with static indices; only the strip's right-hand boundary column goes through a small per-thread
// array (3 loads + 3 stores per row per STRIP instead of per cell), and the eight 16-bit info words of a row leave as one
// 16-byte store.  Table addresses split into a column term and a row term (index = prev * 49 + cur with prev / cur =
// 7 * target base + query base), so a lookup is one add + one shared-memory load; the two lookups that depend on the
// column alone / the row alone are hoisted.  dp_step(prev, dg) = max(prev, 0) - dg, so the state is kept clamped.
// Cells are visited strip-major; the list of maximal cells is put back into row-major order at the end (it is only used
// when it is complete, i.e. at most NC_MAX_CELLS entries; otherwise the row-major replay pass takes over).
PCR_HD int imax(int a, int b) { return a > b ? a : b; }

struct alignas(16) InfoRow8 {
	unsigned int w[4];
};

PCR_HD int dp_fill_strips(Ctx &c, bool hairpin, long long *cells_out)
{
	constexpr int W = NC_STRIP;
	constexpr int NEG = -0x7fffffff;
	dp_border(c);
	c.n_max_cell = 0;
	const int qlen = c.qlen, tlen = hairpin ? c.qlen : c.tlen;
	const unsigned char *tq = hairpin ? c.q : c.t;
	const int max_stem = qlen - 4; // steric limit 3 + 1 (nuc_cruc.cpp:627-635)
	const int rows = hairpin ? max_stem : qlen;
	const int max_cols = hairpin ? max_stem : tlen;
	const int *__restrict__ dg = c.D->dg;
	int eM[NC_MAX_LEN + 2], eIq[NC_MAX_LEN + 2], eIt[NC_MAX_LEN + 2]; // clamped state of column j0 - 1, per row (column 0: the border)
	for (int i = 0; i < NC_MAX_LEN + 2; ++i) eM[i] = eIq[i] = eIt[i] = 0;
	// max_ptr bookkeeping (nuc_cruc.cpp:517-537) per (strip, row), not per cell: the row's largest M within the strip and which cells
	// reach it are compared with the running maximum -- a larger one restarts the list, an equal one appends to it.  (Round 1 stored
	// both per (strip, row) and read them back after the fill: a chain of dependent local-memory loads that was 20 % of the kernel's
	// stall samples, ncu source page.)  Cells arrive strip-major; the list is put into row-major order at the end.
	int max_score = -1;
	long long cells = 0;
	int n_strips = 0;
	for (int j0 = 1; j0 <= max_cols; j0 += W, ++n_strips) {
		int cA[W], cB[W], cC[W];              // column terms: ptb * 343 + tb * 7, tb * 7, tb * 343
		int pM[W], pIq[W], pIt[W];            // row i - 1 of the strip, clamped (row 0: the border)
		{
			int ptb = (j0 == 1) ? (int)bGAP : seq_at(tq, j0 - 2);
#pragma unroll
			for (int k = 0; k < W; ++k) {
				const int tb = seq_at(tq, j0 + k - 1); // past the end: padding (a valid code); those columns are computed and ignored
				cB[k] = tb * NBASE;
				cC[k] = tb * (NBASE * NPAIR);
				cA[k] = ptb * (NBASE * NPAIR) + cB[k];
				ptb = tb;
				pM[k] = pIq[k] = pIt[k] = 0;
			}
		}
		int a0M = 0, a0Iq = 0, a0It = 0; // (i - 1, j0 - 1)
		int pqb = bGAP;
		// the loads of a row (query base, boundary column) are issued one row ahead of their use
		int n_qb = seq_at(c.q, qlen - 1), n_eM = eM[1], n_eIq = eIq[1], n_eIt = eIt[1];
		for (int i = 1; i <= rows; ++i) {
			const int cols = hairpin ? (max_stem - (i - 1)) : tlen;
			if (j0 > cols) break; // hairpin triangle: the rows only get shorter
			const int nv = cols - j0 + 1; // columns of this strip that exist in this row (the rest is computed and ignored)
			const int qb = n_qb;
			int cM = n_eM, cIq = n_eIq;                      // (i, j - 1)
			const int cIt0 = n_eIt;
			n_qb = seq_at(c.q, qlen - (i + 1));
			n_eM = eM[i + 1]; n_eIq = eIq[i + 1]; n_eIt = eIt[i + 1];
			const int rA = pqb * NPAIR + qb;                 // L1 = dg[cA + rA]
			const int rB = bGAP * NPAIR + qb;                // L2 = dg[cA + rB]
			const int rC = bGAP * NBASE * NPAIR + rA;        // L3 = dg[cB + rC]
			const int rD = qb * NPAIR + bGAP;                // L4 = dg[cA + rD]
			const int rE = rA + bGAP * NBASE;                // L6 = dg[cC + rE]
			const int dg7 = dg[rC + bGAP * NBASE];           // (GAP, pqb) -> (GAP, qb): the row alone
			int aM = a0M, aIq = a0Iq, aIt = a0It;            // (i - 1, j - 1)
			a0M = cM; a0Iq = cIq; a0It = cIt0;
			InfoRow8 out[W / 8 > 0 ? W / 8 : 1];
			int xm[W];
			unsigned int pack = 0u;
#pragma unroll
			for (int k = 0; k < W; ++k) {
				const int d1 = aM - dg[cA[k] + rA];
				const int d2 = aIq - dg[cA[k] + rB];
				const int d3 = aIt - dg[cB[k] + rC];
				const int xM = imax(d1, imax(d2, d3));
				const unsigned int mtr = (d1 == xM ? (unsigned)TR_DIAG : 0u) | (d2 == xM ? (unsigned)TR_LEFT : 0u) | (d3 == xM ? (unsigned)TR_UP : 0u);
				const int qi = cM - dg[cA[k] + rD];
				const int qe = cIq - dg[cA[k] + bGAP * NPAIR + bGAP]; // (ptb, GAP) -> (tb, GAP): the column alone
				const int xIq = imax(qi, qe);
				const unsigned int qtr = (qi == xIq ? (unsigned)TR_DIAG : 0u) | (qe == xIq ? (unsigned)TR_LEFT : 0u);
				const int ti = pM[k] - dg[cC[k] + rE];
				const int te = pIt[k] - dg7;
				const int xIt = imax(ti, te);
				const unsigned int ttr = (ti == xIt ? (unsigned)TR_DIAG : 0u) | (te == xIt ? (unsigned)TR_UP : 0u);
				const unsigned int inf = mtr | (qtr << 4) | (ttr << 8) | (((unsigned)xIq >> 31) << 12) | (((unsigned)xIt >> 31) << 13) |
				                         (((unsigned)xM >> 31) << 14) | (xM == 0 ? 0x8000u : 0u);
				if (k & 1) out[k >> 3].w[(k >> 1) & 3] = pack | (inf << 16);
				else pack = inf;
				xm[k] = xM;
				aM = pM[k]; aIq = pIq[k]; aIt = pIt[k];
				cM = imax(xM, 0); cIq = imax(xIq, 0);
				pM[k] = cM; pIq[k] = cIq; pIt[k] = imax(xIt, 0);
			}
#pragma unroll
			for (int v = 0; v < W / 8; ++v) *(InfoRow8 *)(c.info + i * NC_STRIDE + j0 + 8 * v) = out[v]; // cells past `cols` are never read
			eM[i] = pM[W - 1]; eIq[i] = pIq[W - 1]; eIt[i] = pIt[W - 1]; // only read by the next strip when column j0 + W - 1 exists in this row
			int row_max = NEG;
			if (nv < W) {
#pragma unroll
				for (int k = 0; k < W; ++k) xm[k] = (k < nv) ? xm[k] : NEG;
			}
#pragma unroll
			for (int k = 0; k < W; ++k) row_max = imax(row_max, xm[k]);
			unsigned int eq = 0u;
#pragma unroll
			for (int k = 0; k < W; ++k) eq |= (xm[k] == row_max ? 1u : 0u) << k;
			if (row_max >= max_score) {
				if (row_max > max_score) {
					max_score = row_max;
					c.n_max_cell = 0;
				}
				while (eq) {
					const int k = ctz64((uint64_t)eq);
					eq &= eq - 1u;
					if (c.n_max_cell < NC_MAX_CELLS) c.max_cell[c.n_max_cell] = i * NC_STRIDE + j0 + k;
					++c.n_max_cell;
				}
			}
			cells += nv < W ? nv : W;
			pqb = qb;
		}
	}
	// row-major order (cell ids ascend with (row, column)); a list that overflowed is not used (the replay pass takes over)
	if (c.n_max_cell <= NC_MAX_CELLS)
		for (int a = 1; a < c.n_max_cell; ++a) {
			const int x = c.max_cell[a];
			int b = a;
			while (b > 0 && c.max_cell[b - 1] > x) {
				c.max_cell[b] = c.max_cell[b - 1];
				--b;
			}
			c.max_cell[b] = x;
		}
	if (cells_out) *cells_out = cells;
	return max_score;
}

// gap-free main diagonal only (fast_alignment(true))
template <bool REPLAY>
PCR_HD int dp_fill_diagonal_t(Ctx &c, long long *cells_out, int replay_max, Aln *best, int mode)
{
	if (!REPLAY) {
		dp_border(c);
		c.n_max_cell = 0;
	}
	const int len = c.qlen < c.tlen ? c.qlen : c.tlen;
	int max_score = -1, prev_pair = bpair(bGAP, bGAP), aM = -1;
	for (int i = 1; i <= len; ++i) {
		const int cur_pair = bpair(seq_at(c.t, i - 1), seq_at(c.q, c.qlen - i));
		const int xM = dp_step(aM, c.D->dg[prev_pair * NPAIR + cur_pair]);
		const int cell = i * NC_STRIDE + i;
		if (!REPLAY) {
			c.info[cell] = (unsigned short)(TR_DIAG | (TR_INVALID << 4) | (TR_INVALID << 8) | 0x3000 | ((xM < 0) ? 0x4000 : 0) | ((xM == 0) ? 0x8000 : 0));
			note_cell(c, cell, xM, max_score);
		} else if (xM == replay_max) {
			enumerate_dimer(c, cell, *best, mode);
		}
		aM = xM;
		prev_pair = cur_pair;
	}
	if (cells_out) *cells_out = len;
	return REPLAY ? replay_max : max_score;
}

// ---------------------------------------------------------------------------------------------
// trace back (nuc_cruc.cpp:1262-1471) with the persistent split stack (nuc_cruc.h:256-315)
// ---------------------------------------------------------------------------------------------
struct Branch {
	int id;            // cell * 3 + state: the identity of the trace byte (mask_ptr)
	unsigned char mask, cur;
};
struct TraceStack {
	Branch b[NC_STACK_CAP];
	int n;
};

PCR_HD bool path_split(int m) { return ((m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1)) > 1; }
PCR_HD int branch_first(int mask) { return (mask & TR_DIAG) ? TR_DIAG : ((mask & TR_UP) ? TR_UP : TR_LEFT); }
PCR_HD bool branch_next(Branch &b)
{
	int cur = b.cur;
	while ((cur = (cur << 1) & 0xFF) < TR_INVALID && cur != 0) {
		if (cur & b.mask) {
			b.cur = (unsigned char)cur;
			return true;
		}
	}
	b.cur = (unsigned char)cur;
	return false;
}

PCR_HD void trace_back(Ctx &c, int cell, TraceStack &st, int &zero_count, Aln &a, bool hairpin)
{
	const unsigned char *tq = hairpin ? c.q : c.t;
	const int query_len = c.qlen;
	int last_i = cell / NC_STRIDE, last_j = cell % NC_STRIDE;
	a.fm_first = query_len - last_i;
	a.fm_second = last_j - 1;
	int truncate_at_zero = 0;
	bool count_zeros = false;
	if (zero_count < 0) {
		zero_count = 0;
		count_zeros = true;
	} else {
		truncate_at_zero = zero_count--;
	}
	int match_id = -1;          // -1 = the static `first_match` byte (always a plain diagonal step)
	int match_val = TR_DIAG;
	for (;;) {
		bool valid = true;
		int local_match;
		if (path_split(match_val)) {
			int found = -1;
			for (int k = 0; k < st.n; ++k)
				if (st.b[k].id == match_id) { found = k; break; }
			if (found < 0) {
				if (st.n < NC_STACK_CAP) {
					st.b[st.n].id = match_id;
					st.b[st.n].mask = (unsigned char)match_val;
					st.b[st.n].cur = (unsigned char)branch_first(match_val);
					found = st.n++;
				}
				local_match = found >= 0 ? st.b[found].cur : branch_first(match_val);
			} else {
				local_match = st.b[found].cur;
			}
		} else {
			local_match = match_val;
		}
		const int inf = c.info[cell];
		if (local_match == TR_DIAG) { // query_target
			if (last_i > query_len || last_j < 1) {
				valid = false;
			} else {
				if (inf & 0x4000) valid = false; // M < 0
				else if (inf & 0x8000) {         // M == 0
					if (count_zeros) zero_count++;
					else {
						truncate_at_zero--;
						if (truncate_at_zero == 0) valid = false;
					}
				}
				aln_push_back(a, seq_at(c.q, query_len - last_i), seq_at(tq, last_j - 1));
				a.lm_first = query_len - last_i;
				a.lm_second = last_j - 1;
				match_id = cell * 3;
				match_val = inf & 15;
				last_i--;
				last_j--;
			}
		} else if (local_match == TR_LEFT) { // gap_target: gap the query
			if (last_j < 1) {
				valid = false;
			} else {
				if (inf & 0x1000) valid = false;
				aln_push_back(a, bGAP, seq_at(tq, last_j - 1));
				a.lm_first = query_len - last_i + 1;
				a.lm_second = last_j - 1;
				match_id = cell * 3 + 1;
				match_val = (inf >> 4) & 15;
				last_j--;
			}
		} else if (local_match == TR_UP) { // query_gap: gap the target
			if (last_i > query_len) {
				valid = false;
			} else {
				if (inf & 0x2000) valid = false;
				aln_push_back(a, seq_at(c.q, query_len - last_i), bGAP);
				a.lm_first = query_len - last_i;
				a.lm_second = last_j;
				match_id = cell * 3 + 2;
				match_val = (inf >> 8) & 15;
				last_i--;
			}
		} else {
			break; // "invalid_match in trace back": the reference throws; unreachable from computed cells
		}
		if (!valid) break;
		if (last_i < 0 || last_j < 0) break;
		cell = last_i * NC_STRIDE + last_j;
	}
}

// ---------------------------------------------------------------------------------------------
// evaluate_alignment (nuc_cruc.cpp:1473-2137)
// ---------------------------------------------------------------------------------------------
PCR_HD bool non_virtual_pair(int p) { return (p % NBASE < bE) && (p / NBASE < bE); }
PCR_HD bool pair_has_gap(int p) { return (p % NBASE == bGAP) || (p / NBASE >= bGAP); }

PCR_HD bool has_AT_initiation(const Aln &a, int idx)
{ // nuc_cruc.cpp:2747-2763
	int k = idx;
	do {
		--k;
	} while (k != 0 && (aln_q(a, k) == bGAP || aln_t(a, k) == bGAP));
	if (k < 0) k = 0;
	const int p = bpair(aln_q(a, k), aln_t(a, k));
	return p == P_AT || p == P_TA;
}

PCR_HD bool evaluate_alignment(const Ctx &c, Aln &a, int mode)
{
	const Tables *T = c.T;
	int terminal_pair = P_NONE, last_last = P_NONE, last = P_NONE, cur = P_NONE;
	if (mode != MODE_HAIRPIN) {
		a.dH = T->init_H;
		a.dS = T->init_S + ((mode == MODE_HOMO) ? T->symmetry_S : 0.0f);
	}
	unsigned num_query_gap = 0, num_target_gap = 0, num_mismatch = 0, num_base = 0;
	bool terminal_5 = false;
	const int n = a.n;
	if (n <= 0) return false;
	cur = bpair(aln_q(a, 0), aln_t(a, 0));
	if (T->wc[cur]) {
		terminal_5 = true;
		if (cur == P_AT || cur == P_TA) {
			a.dH += T->AT_closing_H;
			a.dS += T->AT_closing_S;
		}
	}
	num_base += (aln_q(a, 0) < bE) ? 1 : 0;
	num_base += (aln_t(a, 0) < bE) ? 1 : 0;
	for (int idx = 1; idx < n; ++idx) {
		const int qb = aln_q(a, idx), tb = aln_t(a, idx);
		last_last = last;
		last = cur;
		cur = bpair(qb, tb);
		const bool align_start = (idx == 1), align_stop = (idx == n - 1);
		const bool in_loop_or_bulge = (qb == bGAP) || (tb == bGAP) || (!T->wc[last] && !T->wc[cur]);
		if (!in_loop_or_bulge) {
			if (align_start && !T->wc[last] && non_virtual_pair(last)) { // frayed start = the two dangling ends
				const int tq = last / NBASE, tt = last % NBASE;
				int tmp = bpair(tq, bE);
				a.dH += T->H[tmp * NPAIR + cur];
				a.dS += T->S[tmp * NPAIR + cur];
				tmp = bpair(bE, tt);
				a.dH += T->H[tmp * NPAIR + cur];
				a.dS += T->S[tmp * NPAIR + cur];
			} else if (align_stop && !T->wc[cur] && non_virtual_pair(cur)) { // frayed end
				int tmp = bpair(qb, bE);
				a.dH += T->H[last * NPAIR + tmp];
				a.dS += T->S[last * NPAIR + tmp];
				tmp = bpair(bE, tb);
				a.dH += T->H[last * NPAIR + tmp];
				a.dS += T->S[last * NPAIR + tmp];
			} else {
				a.dH += T->H[last * NPAIR + cur];
				a.dS += T->S[last * NPAIR + cur];
			}
			num_base += (qb < bE) ? 1 : 0;
			num_base += (tb < bE) ? 1 : 0;
		}
		if (T->wc[cur]) {
			terminal_pair = cur;
			if (!terminal_5) {
				terminal_5 = true;
				if (cur == P_AT || cur == P_TA) {
					a.dH += T->AT_closing_H;
					a.dS += T->AT_closing_S;
				}
			}
			const unsigned max_gap = num_query_gap > num_target_gap ? num_query_gap : num_target_gap;
			if ((num_mismatch > 1) || ((max_gap > 0) && (num_mismatch == 1))) { // closing an internal loop
				const unsigned gap_difference = (num_query_gap > num_target_gap) ? num_query_gap - num_target_gap : num_target_gap - num_query_gap;
				const unsigned loop_size = num_mismatch * 2 + gap_difference;
				if ((loop_size == 2) && (last == P_GT || last == P_TG) && (last_last == P_GT || last_last == P_TG)) {
					a.dH += T->H[last_last * NPAIR + last];
					a.dS += T->S[last_last * NPAIR + last];
					num_base += 2;
				} else {
					a.dS += T->loop_S[loop_size < 129u ? loop_size : 128u];
					a.dS += gap_difference * T->asym_loop_dS;
					int rhs_q = idx - 1, rhs_t = idx - 1;
					a.dH -= T->H[last * NPAIR + cur];
					a.dS -= T->S[last * NPAIR + cur];
					if (!pair_has_gap(last)) { // right terminal mismatch
						a.dH += T->loop_term_H[last * NPAIR + cur];
						a.dS += T->loop_term_S[last * NPAIR + cur];
					} else {
						int mm = P_NONE;
						if (last / NBASE == bGAP) {
							for (;;) {
								if (aln_q(a, rhs_q) < bE) { mm = bpair(aln_q(a, rhs_q), last % NBASE); break; }
								if (rhs_q == 0) break;
								--rhs_q;
							}
						} else {
							for (;;) {
								if (aln_t(a, rhs_t) < bE) { mm = bpair(last / NBASE, aln_t(a, rhs_t)); break; }
								if (rhs_t == 0) break;
								--rhs_t;
							}
						}
						a.dH += T->loop_term_H[mm * NPAIR + cur];
						a.dS += T->loop_term_S[mm * NPAIR + cur];
					}
					// left terminal mismatch: walk back to the Watson-Crick pair that opened the loop
					int lhs_q = idx - 1, lhs_t = idx - 1;
					for (;;) {
						const int pm = bpair(aln_q(a, lhs_q), aln_t(a, lhs_t));
						if (T->wc[pm]) {
							++lhs_q;
							++lhs_t;
							if (aln_q(a, lhs_q) != bGAP && aln_t(a, lhs_t) != bGAP) {
								const int mm = bpair(aln_q(a, lhs_q), aln_t(a, lhs_t));
								a.dH -= T->H[pm * NPAIR + mm];
								a.dS -= T->S[pm * NPAIR + mm];
							}
							num_base += 2;
							while (lhs_q < n && aln_q(a, lhs_q) == bGAP) ++lhs_q;
							while (lhs_t < n && aln_t(a, lhs_t) == bGAP) ++lhs_t;
							const int mm = bpair(aln_q(a, lhs_q < n ? lhs_q : n - 1), aln_t(a, lhs_t < n ? lhs_t : n - 1));
							a.dH += T->loop_term_H[pm * NPAIR + mm];
							a.dS += T->loop_term_S[pm * NPAIR + mm];
							break;
						}
						if (lhs_q == 0) break;
						--lhs_q;
						--lhs_t;
					}
					if (rhs_q != lhs_q) num_base++;
					if (rhs_t != lhs_t) num_base++;
				}
			} else if (num_query_gap || num_target_gap) { // a bulge
				const unsigned bulge_size = (num_query_gap > num_target_gap) ? num_query_gap : num_target_gap;
				if (bulge_size == 1) {
					a.dH += T->H[last_last * NPAIR + cur];
					a.dS += T->S[last_last * NPAIR + cur];
				}
				a.dS += T->bulge_S[bulge_size < 129u ? bulge_size : 128u];
				if ((bulge_size != 1) && (qb == bA || qb == bT)) a.dS += T->bulge_AT_closing_S; // UNAFOLD_COMPATIBILITY (nuc_cruc.h:100)
				if ((bulge_size != 1) && has_AT_initiation(a, idx)) a.dS += T->bulge_AT_closing_S;
			}
			num_query_gap = 0;
			num_target_gap = 0;
			num_mismatch = 0;
		} else {
			num_mismatch += ((qb < bE) && (tb < bE)) ? 1 : 0;
		}
		num_query_gap += (qb == bGAP) ? 1 : 0;
		num_target_gap += (tb == bGAP) ? 1 : 0;
	}
	if (terminal_pair == P_AT || terminal_pair == P_TA) {
		a.dH += T->AT_closing_H;
		a.dS += T->AT_closing_S;
	}
	if (a.dH >= 0.0f) return false; // binding must be enthalpically driven
	a.dS += T->SALT * (0.5f * num_base - 1) * c.D->log_na;
	float tm;
	if (mode == MODE_HAIRPIN) tm = a.dH / a.dS - 273.15f;
	else tm = a.dH / (1.9872e-3f * c.log_strand + a.dS) - 273.15f;
	a.tm = tm > 0.0f ? tm : 0.0f;
	return true;
}

// find_loop_index (nuc_cruc.cpp:2478-2728): index of the special tri/tetra loop equal to the len bases at start
PCR_HD int find_loop_index(const Ctx &c, int start, int len)
{
	const char name[] = "ACGTE";
	char b[6];
	for (int k = 0; k < len; ++k) {
		const int v = seq_at(c.q, start + k);
		b[k] = v <= bE ? name[v] : '\0';
	}
	for (int i = 0; i < 131; ++i) {
		const char *s = c.T->special_loop[i];
		int k = 0;
		while (k < len && s[k] == b[k]) ++k;
		if (k == len && s[len] == '\0') return i;
	}
	return -1;
}

// evaluate_hairpin_alignment (nuc_cruc.cpp:2139-2232)
PCR_HD bool evaluate_hairpin_alignment(const Ctx &c, Aln &a)
{
	const Tables *T = c.T;
	const int last_3 = a.fm_first, last_5 = a.fm_second;
	const unsigned loop_len = (unsigned)(last_3 - last_5 - 1);
	a.dH = 0.0f;
	a.dS = 0.0f;
	a.dS += T->hairpin_S[loop_len < 129u ? loop_len : 128u];
	const int last_pair = bpair(seq_at(c.q, last_5), seq_at(c.q, last_3));
	if (loop_len == 3u) {
		const int li = find_loop_index(c, last_5, 5);
		if (li >= 0) {
			a.dH += T->special_H[li];
			a.dS += T->special_S[li];
		}
		if (last_pair == P_AT || last_pair == P_TA) a.dS += T->bulge_AT_closing_S;
	} else {
		if (loop_len == 4u) {
			const int li = find_loop_index(c, last_5, 6);
			if (li >= 0) {
				a.dH += T->special_H[li];
				a.dS += T->special_S[li];
			}
		}
		const int cur_pair = bpair(seq_at(c.q, last_5 + 1), seq_at(c.q, last_3 - 1));
		a.dH += T->hairpin_term_H[last_pair * NPAIR + cur_pair];
		a.dS += T->hairpin_term_S[last_pair * NPAIR + cur_pair];
	}
	return evaluate_alignment(c, a, MODE_HAIRPIN);
}

// ---------------------------------------------------------------------------------------------
// enumeration of co-optimal alignments from one maximal cell
// ---------------------------------------------------------------------------------------------
PCR_HD void trim_frayed(const Tables *T, Aln &a)
{ // nuc_cruc.cpp:866-902
	while (a.n > 0 && !T->wc[bpair(aln_q(a, a.n - 1), aln_t(a, a.n - 1))]) {
		if (aln_q(a, a.n - 1) < bE) a.lm_first--;
		if (aln_t(a, a.n - 1) < bE) a.lm_second++;
		--a.n;
	}
	while (a.n > 0 && !T->wc[bpair(aln_q(a, 0), aln_t(a, 0))]) {
		if (aln_q(a, 0) < bE) a.fm_first++;
		if (aln_t(a, 0) < bE) a.fm_second--;
		++a.head;
		--a.n;
	}
}
PCR_HD void advance_stack(TraceStack &st, int &zero_count)
{ // nuc_cruc.cpp:904-914
	if (zero_count == 0 && st.n > 0) {
		while (st.n > 0 && !branch_next(st.b[st.n - 1])) --st.n;
		zero_count = -1;
	}
}
PCR_HD void keep_if_better(const Ctx &c, Aln &best, float &best_dg, const Aln &local)
{
	const float local_dg = local.dH - c.D->target_T * local.dS;
	if (!best.valid || local_dg < best_dg) {
		best = local;
		best.valid = true;
		best_dg = local_dg;
	}
}

PCR_HD void enumerate_dimer(Ctx &c, int cell, Aln &best, int mode)
{ // nuc_cruc.cpp:818-1019
	bool first_time = true;
	TraceStack st;
	st.n = 0;
	int zero_count = -1;
	unsigned trace_count = 0;
	float best_dg = best.dH - c.D->target_T * best.dS;
	const int query_len = c.qlen, target_len = c.tlen;
	for (;;) {
		if (!first_time && st.n == 0 && zero_count <= 0) break;
		if (NC_MAX_PATH_ENUM < (int)trace_count) break;
		trace_count++;
		first_time = false;
		Aln a;
		aln_clear(a);
		trace_back(c, cell, st, zero_count, a, false);
		trim_frayed(c.T, a);
		advance_stack(st, zero_count);
		// dangling ends / frayed ends on both sides (enable_dangle = true, true)
		if (a.fm_first != 0 || a.fm_second != target_len - 1) {
			int qb, tb;
			if (a.fm_first == 0) qb = bE;
			else { a.fm_first--; qb = seq_at(c.q, a.fm_first); }
			if (a.fm_second == target_len - 1) tb = bE;
			else { a.fm_second++; tb = seq_at(c.t, a.fm_second); }
			if (a.head > 0) {
				--a.head;
				a.q[a.head] = (unsigned char)qb;
				a.t[a.head] = (unsigned char)tb;
				++a.n;
			}
		}
		if (a.lm_first != query_len - 1 || a.lm_second != 0) {
			int qb, tb;
			if (a.lm_first == query_len - 1) qb = bE;
			else { a.lm_first++; qb = seq_at(c.q, a.lm_first); }
			if (a.lm_second == 0) tb = bE;
			else { a.lm_second--; tb = seq_at(c.t, a.lm_second); }
			aln_push_back(a, qb, tb);
		}
		if (a.n < 3) continue;
		if (evaluate_alignment(c, a, mode)) keep_if_better(c, best, best_dg, a);
	}
}

PCR_HD void enumerate_hairpin(Ctx &c, int cell, Aln &best)
{ // nuc_cruc.cpp:1021-1260
	bool first_time = true;
	TraceStack st;
	st.n = 0;
	int zero_count = -1;
	unsigned trace_count = 0;
	float best_dg = best.dH - c.D->target_T * best.dS;
	const int query_len = c.qlen;
	for (;;) {
		if (!first_time && st.n == 0 && zero_count <= 0) break;
		if (NC_MAX_PATH_ENUM < (int)trace_count) break;
		trace_count++;
		first_time = false;
		Aln a;
		aln_clear(a);
		trace_back(c, cell, st, zero_count, a, true);
		trim_frayed(c.T, a);
		advance_stack(st, zero_count);
		if (a.n >= 3 && evaluate_hairpin_alignment(c, a)) keep_if_better(c, best, best_dg, a);
		if (a.lm_second != 0 || a.lm_first != query_len - 1) { // the open end of the stem
			int qb, tb;
			if (a.lm_second == 0) tb = bE;
			else { a.lm_second--; tb = seq_at(c.q, a.lm_second); }
			if (a.lm_first == query_len - 1) qb = bE;
			else { a.lm_first++; qb = seq_at(c.q, a.lm_first); }
			aln_push_back(a, qb, tb);
		}
		const int align_size = a.n;
		if (align_size < 3) continue;
		if (evaluate_hairpin_alignment(c, a)) keep_if_better(c, best, best_dg, a);
		if (align_size <= 3) continue;
		// an A-T closing pair carries a penalty: try the stem without it
		const int last_3 = a.fm_first, last_5 = a.fm_second;
		const int last_pair = bpair(seq_at(c.q, last_5), seq_at(c.q, last_3));
		if (last_pair == P_GC || last_pair == P_CG) continue;
		a.fm_first++;
		a.fm_second--;
		++a.head;
		--a.n;
		if (evaluate_hairpin_alignment(c, a)) keep_if_better(c, best, best_dg, a);
	}
}

// ---------------------------------------------------------------------------------------------
// front ends (nuc_cruc.h:723-759, nuc_cruc.cpp:2236-2455).  out = {tm, dH, dS, dp_dg}
// ---------------------------------------------------------------------------------------------
struct Result {
	float tm, dH, dS, dp_dg;
	long long cells;
};

PCR_HD Result run_problem(Ctx &c, int op)
{
	Result r;
	r.tm = r.dH = r.dS = r.dp_dg = 0.0f;
	r.cells = 0;
	Aln best;
	aln_clear(best);
	if (op == OP_PM_DUPLEX) { // tm_pm_duplex: the oligo against its exact complement
		for (int i = 0; i < c.qlen; ++i) {
			const int b = seq_at(c.q, i);
			aln_push_back(best, b, bT - b); // A<->T, C<->G
		}
		if (best.n > 0) evaluate_alignment(c, best, MODE_HETERO);
		r.tm = best.tm;
		r.dH = best.dH;
		r.dS = best.dS;
		r.dp_dg = 0.0f + c.T->init_H - c.D->target_T * c.T->init_S;
		return r;
	}
	int max_score;
	const bool hairpin = (op == OP_HAIRPIN);
	const bool diagonal = (op == OP_HETERODIMER_DIAG || op == OP_HOMODIMER_DIAG);
	const int mode = hairpin ? MODE_HAIRPIN : ((op == OP_HOMODIMER || op == OP_HOMODIMER_DIAG) ? MODE_HOMO : MODE_HETERO);
	if (diagonal) max_score = dp_fill_diagonal_t<false>(c, &r.cells, 0, nullptr, mode);
	else max_score = dp_fill_strips(c, hairpin, &r.cells);
	if (c.n_max_cell <= NC_MAX_CELLS) { // tm_dimer / approximate_tm_hairpin: every maximal cell, in row-major order
		for (int k = 0; k < c.n_max_cell; ++k) {
			if (hairpin) enumerate_hairpin(c, c.max_cell[k], best);
			else enumerate_dimer(c, c.max_cell[k], best, mode);
		}
	} else if (diagonal) {
		dp_fill_diagonal_t<true>(c, nullptr, max_score, &best, mode);
	} else {
		dp_fill_t<true>(c, hairpin, nullptr, max_score, &best, mode);
	}
	r.tm = best.tm;
	r.dH = best.dH;
	r.dS = best.dS;
	// delta_G_dp() = dp_dg + the initiation free energy (nuc_cruc.cpp:3067-3072)
	r.dp_dg = -((float)max_score / 10000.0f) + c.T->init_H - c.D->target_T * c.T->init_S;
	return r;
}


// ---------------------------------------------------------------------------------------------
// host side: the parameter set and the integer DP table for one salt concentration
// ---------------------------------------------------------------------------------------------
namespace host {
#include "santalucia_tables.inc"
}

inline void build_tables(Tables &t)
{
	memset(&t, 0, sizeof(t));
	memcpy(t.H, host::NC_PARAM_H, sizeof(t.H));
	memcpy(t.S, host::NC_PARAM_S, sizeof(t.S));
	memcpy(t.loop_term_H, host::NC_LOOP_TERMINAL_H, sizeof(t.loop_term_H));
	memcpy(t.loop_term_S, host::NC_LOOP_TERMINAL_S, sizeof(t.loop_term_S));
	memcpy(t.hairpin_term_H, host::NC_HAIRPIN_TERMINAL_H, sizeof(t.hairpin_term_H));
	memcpy(t.hairpin_term_S, host::NC_HAIRPIN_TERMINAL_S, sizeof(t.hairpin_term_S));
	memcpy(t.loop_S, host::NC_LOOP_S, sizeof(t.loop_S));
	memcpy(t.bulge_S, host::NC_BULGE_S, sizeof(t.bulge_S));
	memcpy(t.hairpin_S, host::NC_HAIRPIN_S, sizeof(t.hairpin_S));
	memcpy(t.special_H, host::NC_HAIRPIN_SPECIAL_H, sizeof(t.special_H));
	memcpy(t.special_S, host::NC_HAIRPIN_SPECIAL_S, sizeof(t.special_S));
	memcpy(t.supp, host::NC_SUPP, sizeof(t.supp));
	memcpy(t.supp_salt, host::NC_SUPP_SALT, sizeof(t.supp_salt));
	t.init_H = host::NC_SCALARS[0];
	t.init_S = host::NC_SCALARS[1];
	t.asym_loop_dS = host::NC_SCALARS[2];
	t.bulge_AT_closing_S = host::NC_SCALARS[3];
	t.AT_closing_H = host::NC_SCALARS[4];
	t.AT_closing_S = host::NC_SCALARS[5];
	t.symmetry_S = host::NC_SCALARS[6];
	t.SALT = host::NC_SCALARS[7];
	memcpy(t.wc, host::NC_WATSON_CRICK, sizeof(t.wc));
	memcpy(t.special_loop, host::NC_SPECIAL_LOOP, sizeof(t.special_loop));
}

// update_dp_param (nuc_cruc.cpp:191-342): dG x 10^4 at target_T for every (previous pair, current pair), with
// the supplementary (fitted) terms for gaps, terminal matches next to gaps and double mismatches floored at 0.
// Float arithmetic in the reference's order; ln[Na+] is the float overload of log, i.e. logf.
inline void build_dp(const Tables &t, float na, float target_T, DpTable &d)
{
	const float log_na = logf(na);
	d.target_T = target_T;
	d.log_na = log_na;
	const float salt_correction = t.SALT * log_na;
	const float loop_sc = salt_correction * t.supp_salt[0];
	const float bulge_sc = salt_correction * t.supp_salt[1];
	const float term_match_sc = salt_correction * t.supp_salt[2];
	const float term_mismatch_sc = salt_correction * t.supp_salt[3];
	auto scale = [](float x) { return (int)(x * 10000.0f); };
	auto floor0 = [](int x) { return x > 0 ? x : 0; };
	for (int i = 0; i < NPAIR; ++i)
		for (int j = 0; j < NPAIR; ++j) d.dg[i * NPAIR + j] = scale(t.H[i * NPAIR + j] - target_T * (t.S[i * NPAIR + j] + salt_correction));
	const int loop_dg = floor0(scale(t.supp[SUPP_LOOP_H] - target_T * (t.supp[SUPP_LOOP_S] + loop_sc)));
	const int bulge_dg = floor0(scale(t.supp[SUPP_BULGE_H] - target_T * (t.supp[SUPP_BULGE_S] + bulge_sc)));
	const int tm_at = floor0(scale(t.supp[SUPP_TM_AT_H] - target_T * (t.supp[SUPP_TM_AT_S] + term_match_sc)));
	const int tm_gc = floor0(scale(t.supp[SUPP_TM_GC_H] - target_T * (t.supp[SUPP_TM_GC_S] + term_match_sc)));
	const int tm_i = floor0(scale(t.supp[SUPP_TM_I_H] - target_T * (t.supp[SUPP_TM_I_S] + term_match_sc)));
	const int tmm = floor0(scale(t.supp[SUPP_TMM_H] - target_T * (t.supp[SUPP_TMM_S] + term_mismatch_sc)));
	for (int i = bA; i <= bI; ++i)
		for (int j = bA; j <= bI; ++j) {
			const int cur = bpair(i, j);
			int v;
			if (t.wc[cur]) v = (cur == P_AT || cur == P_TA) ? tm_at : ((cur == P_GC || cur == P_CG) ? tm_gc : tm_i);
			else v = tmm;
			for (int k = bA; k <= bI; ++k) {
				const int p1 = bpair(k, bGAP), p2 = bpair(bGAP, k);
				d.dg[cur * NPAIR + p1] = d.dg[p1 * NPAIR + cur] = d.dg[cur * NPAIR + p2] = d.dg[p2 * NPAIR + cur] = v;
			}
			for (int k = bA; k <= bI; ++k)
				for (int l = bA; l <= bI; ++l) {
					const int prev = bpair(k, l);
					if (!t.wc[cur] && !t.wc[prev]) d.dg[cur * NPAIR + prev] = loop_dg;
				}
		}
	for (int i = bA; i <= bI; ++i)
		for (int j = bA; j <= bI; ++j) {
			d.dg[bpair(i, bGAP) * NPAIR + bpair(j, bGAP)] = bulge_dg;
			d.dg[bpair(bGAP, i) * NPAIR + bpair(bGAP, j)] = bulge_dg;
		}
}

// 'A','C','G','T','I' (either case) -> base code; anything else -> -1 (set_query throws, nuc_cruc.h:875-913)
inline int base_code(char ch)
{
	switch (ch) {
	case 'A': case 'a': return bA;
	case 'C': case 'c': return bC;
	case 'G': case 'g': return bG;
	case 'T': case 't': return bT;
	case 'I': case 'i': return bI;
	default: return -1;
	}
}

} // namespace nc
} // namespace pcr
