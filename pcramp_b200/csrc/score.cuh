// score.cuh -- K2, pair scoring against the word database: PCR::collect_candidates,
// find_amplicon_match, update_identity, compute_coverage and find_target_match
// (pcr_assay.cpp:12-69,271-302,338-441,544-578; optimize.cpp:209-301).
//
// The reference scans ALL database keys per oligo (match_words), expands them to occurrences
// (find_oligo_match), sorts by (sequence, loc) and pairs plus/minus hits per sequence.  Whether a
// sequence is detected is a pure set predicate over that sequence's entries -- the loop order and the
// two `break`s in find_amplicon_match only prune (amplicon length and has_split are monotone in the
// minus hit's loc, and a minus hit sorted before the plus hit always overlaps it) -- so the database
// is kept grouped by sequence and one CTA scores one sequence against every pair: its entries sit in
// shared memory, a warp takes a pair, lanes take entries.
#pragma once
#include "db.cuh"

namespace pcr {

struct OligoDev {
	uint64_t hi, lo;
	float norm;      // float(1.0 / len), optimize.cpp:221
	uint32_t packed; // thr[7:0] | start[15:8] | stop[23:16] | penultimate nibble[27:24] | last nibble[31:28]
};

// Table 2 of Li et al., Genomics 83 (2004) 311-320 as used by taq_mama_correction (word.cpp:249-294):
// rows = template pair, columns = primer pair, order {CC,GC,AC,TC,CG,GG,AG,TG,CA,GA,AA,TA,CT,GT,AT,TT}.
__constant__ float c_taq_mama[256] = {
	1.000f, 0.968f, 0.947f, 1.034f, 0.547f, 0.253f, 0.230f, 0.359f, 0.606f, 0.282f, 0.372f, 0.347f, 0.957f, 0.382f, 0.399f, 0.687f,
	0.989f, 1.000f, 1.023f, 1.000f, 0.420f, 0.662f, 0.445f, 0.367f, 0.870f, 0.512f, 0.492f, 0.508f, 0.372f, 1.000f, 0.492f, 0.714f,
	1.011f, 1.000f, 1.000f, 1.000f, 0.459f, 0.277f, 0.570f, 0.343f, 0.927f, 0.362f, 0.590f, 0.542f, 0.439f, 0.488f, 0.978f, 0.662f,
	1.000f, 0.907f, 1.000f, 1.000f, 0.382f, 0.234f, 0.228f, 0.542f, 0.763f, 0.309f, 0.410f, 0.473f, 0.426f, 0.347f, 0.423f, 0.947f,
	0.590f, 0.334f, 0.445f, 0.323f, 1.000f, 0.978f, 0.927f, 0.989f, 0.907f, 0.645f, 0.525f, 0.455f, 0.927f, 0.408f, 0.408f, 0.707f,
	0.327f, 0.595f, 0.319f, 0.396f, 0.947f, 1.000f, 0.978f, 0.989f, 0.405f, 0.861f, 0.681f, 0.512f, 0.410f, 0.968f, 0.452f, 0.714f,
	0.410f, 0.420f, 0.590f, 0.311f, 1.023f, 1.000f, 1.000f, 1.000f, 0.488f, 0.898f, 0.907f, 0.566f, 0.442f, 0.449f, 0.989f, 0.707f,
	0.423f, 0.343f, 0.305f, 0.585f, 1.034f, 0.879f, 0.927f, 1.000f, 0.473f, 0.720f, 0.547f, 0.957f, 0.459f, 0.374f, 0.459f, 1.023f,
	1.023f, 0.429f, 0.473f, 0.477f, 1.023f, 0.466f, 0.420f, 0.477f, 1.000f, 0.978f, 0.907f, 0.978f, 0.907f, 0.380f, 0.525f, 0.669f,
	0.442f, 1.046f, 0.455f, 0.470f, 0.432f, 1.058f, 0.481f, 0.485f, 0.917f, 1.000f, 1.023f, 1.023f, 0.336f, 0.968f, 0.534f, 0.639f,
	0.617f, 0.452f, 1.011f, 0.439f, 0.492f, 0.504f, 0.978f, 0.462f, 0.989f, 0.947f, 1.000f, 0.978f, 0.405f, 0.405f, 0.888f, 0.606f,
	0.601f, 0.377f, 0.377f, 1.046f, 0.500f, 0.399f, 0.408f, 1.034f, 0.978f, 0.720f, 0.870f, 1.000f, 0.402f, 0.313f, 0.651f, 0.927f,
	0.978f, 0.462f, 0.466f, 0.488f, 0.420f, 0.239f, 0.225f, 0.336f, 0.504f, 0.269f, 0.319f, 0.656f, 1.000f, 0.835f, 0.907f, 1.034f,
	0.429f, 1.011f, 0.473f, 0.477f, 0.340f, 0.413f, 0.357f, 0.354f, 0.352f, 0.538f, 0.413f, 0.794f, 0.927f, 1.000f, 1.058f, 1.000f,
	0.595f, 0.492f, 0.968f, 0.485f, 0.367f, 0.282f, 0.388f, 0.439f, 0.413f, 0.309f, 0.566f, 0.917f, 0.957f, 0.957f, 1.000f, 0.989f,
	0.590f, 0.380f, 0.410f, 0.968f, 0.364f, 0.223f, 0.230f, 0.416f, 0.321f, 0.239f, 0.301f, 0.645f, 0.978f, 0.714f, 0.947f, 1.000f};

__device__ __forceinline__ int taq_index(uint32_t b)
{ // word.cpp:233-247
	return b == 2u ? 0 : b == 4u ? 1 : b == 1u ? 2 : b == 8u ? 3 : -1;
}

// per-oligo constants: match_words threshold unsigned(size * thr^2) (optimize.cpp:293), identity norm
__global__ void prep_oligos_kernel(const uint64_t *__restrict__ f, const uint64_t *__restrict__ r, uint32_t n_pairs, float thr2,
	OligoDev *out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 2u * n_pairs) return;
	const uint64_t *src = (i & 1u) ? r : f;
	W128 w;
	w.hi = src[2 * (i >> 1)];
	w.lo = src[2 * (i >> 1) + 1];
	const int size = w_size(w), start = w_start(w), stop = w_stop(w);
	OligoDev o;
	o.hi = w.hi;
	o.lo = w.lo;
	o.norm = size > 0 ? (float)(1.0 / (double)size) : 0.0f;
	const uint32_t thr = (uint32_t)__fmul_rn((float)size, thr2);
	const uint32_t pen = stop >= 1 ? w_get(w, stop - 1) : 0u, last = stop >= 0 ? w_get(w, stop) : 0u;
	o.packed = (thr & 255u) | ((uint32_t)(start & 255) << 8) | ((uint32_t)(stop & 255) << 16) | (pen << 24) | (last << 28);
	out[i] = o;
}

struct ScoreEntry {
	uint64_t hi, lo;
	int32_t loc;
	uint32_t strand;
};

// identity of an oligo against one database word (optimize.cpp:221-259)
__device__ __forceinline__ float oligo_identity(const OligoDev &o, int count, uint64_t e_hi, uint64_t e_lo, int taq)
{
	float v = __fmul_rn((float)count, o.norm);
	if (taq) {
		const uint32_t p0 = (o.packed >> 24) & 15u, p1 = o.packed >> 28;
		const int a = taq_index(p0), b = taq_index(p1);
		if (a >= 0 && b >= 0) { // neither primer base degenerate
			const int stop = (int)((o.packed >> 16) & 255u);
			W128 e;
			e.hi = e_hi;
			e.lo = e_lo;
			const int c = stop >= 1 ? taq_index(w_get(e, stop - 1)) : -1, d = taq_index(w_get(e, stop));
			if (c >= 0 && d >= 0) v = __fmul_rn(v, fminf(1.0f, c_taq_mama[16 * (4 * d + c) + (4 * b + a)]));
		}
	}
	return v;
}

constexpr int SCORE_THREADS = 256;
constexpr int SCORE_SMEM_ENTRIES = 1024; // 24 KB; longer per-sequence lists spill to L2 reads

// one pass of collect_candidates: oligo P binds the plus strand, oligo M the minus strand
__device__ inline bool amplicon_pass(const SeqDev &sd, uint32_t seq, const ScoreEntry *s_ent, const uint64_t *__restrict__ g_hi,
	const uint64_t *__restrict__ g_lo, const int32_t *__restrict__ g_loc, const uint32_t *__restrict__ g_strand, uint32_t e0, uint32_t E,
	const OligoDev &P, const OligoDev &M, float detect, int amp_min, int amp_max, int taq, uint32_t lane)
{
	const int p_thr = (int)(P.packed & 255u), p_start = (int)((P.packed >> 8) & 255u), p_stop = (int)((P.packed >> 16) & 255u);
	const int m_thr = (int)(M.packed & 255u), m_start = (int)((M.packed >> 8) & 255u), m_stop = (int)((M.packed >> 16) & 255u);
	const int L = (int)sd.len[seq];
	W128 pw, mw;
	pw.hi = P.hi; pw.lo = P.lo;
	mw.hi = M.hi; mw.lo = M.lo;
	for (uint32_t base = 0; base < E; base += 32u) {
		const uint32_t e = base + lane;
		ScoreEntry en;
		en.hi = en.lo = 0; en.loc = 0; en.strand = 0;
		if (e < E) {
			if (e < (uint32_t)SCORE_SMEM_ENTRIES) en = s_ent[e];
			else { en.hi = g_hi[e0 + e]; en.lo = g_lo[e0 + e]; en.loc = g_loc[e0 + e]; en.strand = g_strand[e0 + e]; }
		}
		W128 ew;
		ew.hi = en.hi; ew.lo = en.lo;
		const int cp = w_and_count(pw, ew);
		uint32_t plus_mask = __ballot_sync(0xffffffffu, e < E && en.strand == STRAND_PLUS && cp >= p_thr);
		while (plus_mask) {
			const int src = __ffs(plus_mask) - 1;
			plus_mask &= plus_mask - 1u;
			const int loc_p = __shfl_sync(0xffffffffu, en.loc, src);
			const int cnt_p = __shfl_sync(0xffffffffu, cp, src);
			const uint64_t hi_p = __shfl_sync(0xffffffffu, en.hi, src), lo_p = __shfl_sync(0xffffffffu, en.lo, src);
			const float ident_p = oligo_identity(P, cnt_p, hi_p, lo_p, taq);
			const int plus_loc3 = loc_p + p_stop; // sequence.h:67-75, plus strand
			for (uint32_t base2 = 0; base2 < E; base2 += 32u) {
				const uint32_t e2 = base2 + lane;
				bool ok = false;
				if (e2 < E) {
					ScoreEntry m2;
					if (e2 < (uint32_t)SCORE_SMEM_ENTRIES) m2 = s_ent[e2];
					else { m2.hi = g_hi[e0 + e2]; m2.lo = g_lo[e0 + e2]; m2.loc = g_loc[e0 + e2]; m2.strand = g_strand[e0 + e2]; }
					if (m2.strand == STRAND_MINUS) {
						W128 w2;
						w2.hi = m2.hi; w2.lo = m2.lo;
						const int cm = w_and_count(mw, w2);
						if (cm >= m_thr) {
							const int minus_loc5 = m2.loc - m_stop; // sequence.h:57-65, minus strand
							if (plus_loc3 < minus_loc5) {             // pcr_assay.cpp:368-371
								int amp_start = loc_p + p_start;
								const int amp_stop = min(m2.loc - m_start, L - 1);
								int amp_len = amp_stop - amp_start + 1;
								if (amp_len >= amp_min && amp_len <= amp_max) { // :383-392
									if (amp_start < 0) { amp_len += amp_start; amp_start = 0; } // :412-416
									if (amp_len >= 0 && !has_split_dev(sd, seq, amp_start, amp_len)) { // :418
										const float ident_m = oligo_identity(M, cm, m2.hi, m2.lo, taq);
										ok = __fsqrt_rn(__fmul_rn(ident_p, ident_m)) >= detect; // :292-294
									}
								}
							}
						}
					}
				}
				if (__any_sync(0xffffffffu, ok)) return true;
			}
		}
	}
	return false;
}

__global__ void __launch_bounds__(SCORE_THREADS)
score_kernel(SeqDev sd, const uint64_t *__restrict__ g_hi, const uint64_t *__restrict__ g_lo, const int32_t *__restrict__ g_loc,
	const uint32_t *__restrict__ g_strand, const uint32_t *__restrict__ seq_off, const OligoDev *__restrict__ oligos, uint32_t n_pairs,
	float detect, int amp_min, int amp_max, int taq, uint32_t *bits_any, uint32_t *bits_pass1, uint32_t n_words)
{
	__shared__ ScoreEntry s_ent[SCORE_SMEM_ENTRIES];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = SCORE_THREADS / 32;
	for (uint32_t seq = blockIdx.x; seq < sd.n; seq += gridDim.x) {
		const uint32_t e0 = seq_off[seq], E = seq_off[seq + 1] - e0;
		if (E == 0u || !sd.active[seq]) continue; // optimize.cpp:280-283
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < min(E, (uint32_t)SCORE_SMEM_ENTRIES); i += SCORE_THREADS) {
			ScoreEntry en;
			en.hi = g_hi[e0 + i]; en.lo = g_lo[e0 + i]; en.loc = g_loc[e0 + i]; en.strand = g_strand[e0 + i];
			s_ent[i] = en;
		}
		__syncthreads();
		for (uint32_t p = warp; p < n_pairs; p += n_warps) {
			const OligoDev F = oligos[2 * p], R = oligos[2 * p + 1];
			// {F(+), R(-)} then {R(+), F(-)}  (pcr_assay.cpp:37-59)
			const bool d1 = amplicon_pass(sd, seq, s_ent, g_hi, g_lo, g_loc, g_strand, e0, E, F, R, detect, amp_min, amp_max, taq, lane);
			const bool d2 = d1 ? false : amplicon_pass(sd, seq, s_ent, g_hi, g_lo, g_loc, g_strand, e0, E, R, F, detect, amp_min, amp_max, taq, lane);
			if (lane == 0u && (d1 || d2)) {
				const uint32_t bit = 1u << (seq & 31u);
				atomicOr(bits_any + (size_t)p * n_words + (seq >> 5), bit);
				if (d1) atomicOr(bits_pass1 + (size_t)p * n_words + (seq >> 5), bit);
			}
		}
	}
}

// compute_coverage (pcr_assay.cpp:271-302): weights of the detected sequences summed in double in the
// order the reference meets them -- pass-1 amplicons by ascending sequence, then the sequences only
// pass 2 finds, ascending -- and narrowed to float on return.
__global__ void coverage_kernel(const uint32_t *__restrict__ bits_any, const uint32_t *__restrict__ bits_pass1, const float *__restrict__ weight,
	uint32_t n_pairs, uint32_t n_words, uint32_t n_seq, float *coverage)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n_pairs) return;
	double acc = 0.0;
	for (int pass = 0; pass < 2; ++pass) {
		for (uint32_t w = 0; w < n_words; ++w) {
			const uint32_t a = bits_any[(size_t)p * n_words + w], b = bits_pass1[(size_t)p * n_words + w];
			uint32_t m = pass == 0 ? b : (a & ~b);
			while (m) {
				const uint32_t s = w * 32u + (uint32_t)(__ffs(m) - 1);
				m &= m - 1u;
				if (s < n_seq) acc = __dadd_rn(acc, (double)weight[s]);
			}
		}
	}
	coverage[p] = (float)acc;
}

} // namespace pcr
