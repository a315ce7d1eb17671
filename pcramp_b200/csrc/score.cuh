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
#include "fst.cuh"

namespace pcr {

struct OligoDev {
	uint32_t a, c, g, t; // letter planes over the frame (word128.cuh)
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
__device__ __forceinline__ OligoDev make_oligo(const W128 &w, float thr2)
{
	const int size = w_size(w), start = w_start(w), stop = w_stop(w);
	OligoDev o;
	const Planes4 pl = w_planes(w);
	o.a = pl.a; o.c = pl.c; o.g = pl.g; o.t = pl.t;
	o.norm = size > 0 ? (float)(1.0 / (double)size) : 0.0f;
	const uint32_t thr = (uint32_t)__fmul_rn((float)size, thr2);
	const uint32_t pen = stop >= 1 ? w_get(w, stop - 1) : 0u, last = stop >= 0 ? w_get(w, stop) : 0u;
	o.packed = (thr & 255u) | ((uint32_t)(start & 255) << 8) | ((uint32_t)(stop & 255) << 16) | (pen << 24) | (last << 28);
	return o;
}

__global__ void prep_oligos_kernel(const uint64_t *__restrict__ f, const uint64_t *__restrict__ r, uint32_t n_pairs, float thr2,
	OligoDev *out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 2u * n_pairs) return;
	const uint64_t *src = (i & 1u) ? r : f;
	W128 w;
	w.hi = src[2 * (i >> 1)];
	w.lo = src[2 * (i >> 1) + 1];
	out[i] = make_oligo(w, thr2);
}

struct ScoreEntry {
	uint32_t a, c, g, t; // letter planes of the database word
	int32_t loc;
	uint32_t strand;
};

__device__ __forceinline__ int oligo_count(const OligoDev &o, const ScoreEntry &e)
{ // Word::operator& (word.cpp:111-154) on letter planes: 4 LOP3 + POPC
	return __popc((o.a & e.a) | (o.c & e.c) | (o.g & e.g) | (o.t & e.t));
}

// identity of an oligo against one database word (optimize.cpp:221-259)
__device__ __forceinline__ float oligo_identity(const OligoDev &o, int count, const ScoreEntry &e, int taq)
{
	float v = __fmul_rn((float)count, o.norm);
	if (taq) {
		const uint32_t p0 = (o.packed >> 24) & 15u, p1 = o.packed >> 28;
		const int a = taq_index(p0), b = taq_index(p1);
		if (a >= 0 && b >= 0) { // neither primer base degenerate
			const int stop = (int)((o.packed >> 16) & 255u);
			Planes4 ep;
			ep.a = e.a; ep.c = e.c; ep.g = e.g; ep.t = e.t;
			const int c = stop >= 1 ? taq_index(planes_nibble(ep, stop - 1)) : -1, d = taq_index(planes_nibble(ep, stop));
			if (c >= 0 && d >= 0) v = __fmul_rn(v, fminf(1.0f, c_taq_mama[16 * (4 * d + c) + (4 * b + a)]));
		}
	}
	return v;
}

constexpr int SCORE_THREADS = 256;
constexpr int SCORE_SMEM_ENTRIES = 1024; // 24 KB; longer per-sequence lists spill to L2 reads

__device__ __forceinline__ ScoreEntry load_entry(const ScoreEntry *s_ent, const uint4 *__restrict__ g_pl, const int32_t *__restrict__ g_loc,
	const uint32_t *__restrict__ g_strand, uint32_t e0, uint32_t e)
{
	if (s_ent && e < (uint32_t)SCORE_SMEM_ENTRIES) return s_ent[e]; // s_ent == nullptr: no staged copy, read the database in place
	ScoreEntry en;
	const uint4 v = g_pl[e0 + e];
	en.a = v.x; en.c = v.y; en.g = v.z; en.t = v.w;
	en.loc = g_loc[e0 + e];
	en.strand = g_strand[e0 + e];
	return en;
}

// one pass of collect_candidates: oligo P binds the plus strand (entries [0, Ep)), oligo M the minus strand
// (entries [Ep, E)).  Warp-cooperative: lanes take entries.
//
// VARIANT = true is the scoring an optimisation move does (optimize_pcr.cpp: is_valid -> update_identity -> compute_coverage):
// the candidate amplicons -- which database words take part, and the amplicon geometry -- are those of the BASE assay
// (Pb, Mb: collect_candidates ran on it, pcr_assay.cpp:12-69), while the identities are recomputed for the trial oligos
// (P, M: update_identity, optimize.cpp:209-261).
template <bool VARIANT>
__device__ inline bool amplicon_pass(const SeqDev &sd, uint32_t seq, const ScoreEntry *s_ent, const uint4 *__restrict__ g_pl,
	const int32_t *__restrict__ g_loc, const uint32_t *__restrict__ g_strand, uint32_t e0, uint32_t Ep, uint32_t E, const OligoDev &P,
	const OligoDev &M, const OligoDev &Pb, const OligoDev &Mb, float detect, int amp_min, int amp_max, int taq, uint32_t lane)
{
	const int p_thr = (int)(Pb.packed & 255u), p_start = (int)((Pb.packed >> 8) & 255u), p_stop = (int)((Pb.packed >> 16) & 255u);
	const int m_thr = (int)(Mb.packed & 255u), m_start = (int)((Mb.packed >> 8) & 255u), m_stop = (int)((Mb.packed >> 16) & 255u);
	const int L = (int)sd.len[seq];
	for (uint32_t base = 0; base < Ep; base += 32u) {
		const uint32_t e = base + lane;
		ScoreEntry en;
		en.a = en.c = en.g = en.t = 0u; en.loc = 0; en.strand = 0u;
		if (e < Ep) en = load_entry(s_ent, g_pl, g_loc, g_strand, e0, e);
		const int cp = oligo_count(P, en);
		uint32_t plus_mask = __ballot_sync(0xffffffffu, e < Ep && (VARIANT ? oligo_count(Pb, en) : cp) >= p_thr);
		while (plus_mask) {
			const int src = __ffs(plus_mask) - 1;
			plus_mask &= plus_mask - 1u;
			ScoreEntry pe;
			pe.a = __shfl_sync(0xffffffffu, en.a, src); pe.c = __shfl_sync(0xffffffffu, en.c, src);
			pe.g = __shfl_sync(0xffffffffu, en.g, src); pe.t = __shfl_sync(0xffffffffu, en.t, src);
			pe.loc = __shfl_sync(0xffffffffu, en.loc, src);
			pe.strand = STRAND_PLUS;
			const int cnt_p = __shfl_sync(0xffffffffu, cp, src);
			const float ident_p = oligo_identity(P, cnt_p, pe, taq);
			const int plus_loc3 = pe.loc + p_stop; // sequence.h:67-75, plus strand
			for (uint32_t base2 = Ep; base2 < E; base2 += 32u) {
				const uint32_t e2 = base2 + lane;
				bool ok = false;
				if (e2 < E) {
					const ScoreEntry m2 = load_entry(s_ent, g_pl, g_loc, g_strand, e0, e2);
					const int cm = oligo_count(M, m2);
					if ((VARIANT ? oligo_count(Mb, m2) : cm) >= m_thr) {
						const int minus_loc5 = m2.loc - m_stop; // sequence.h:57-65, minus strand
						if (plus_loc3 < minus_loc5) {             // pcr_assay.cpp:368-371
							int amp_start = pe.loc + p_start;
							const int amp_stop = min(m2.loc - m_start, L - 1);
							int amp_len = amp_stop - amp_start + 1;
							if (amp_len >= amp_min && amp_len <= amp_max) { // :383-392
								if (amp_start < 0) { amp_len += amp_start; amp_start = 0; } // :412-416
								if (amp_len >= 0 && !has_split_dev(sd, seq, amp_start, amp_len)) { // :418
									const float ident_m = oligo_identity(M, cm, m2, taq);
									ok = __fsqrt_rn(__fmul_rn(ident_p, ident_m)) >= detect; // :292-294
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

// One CTA per sequence; its database entries (plus strand first) sit in shared memory.
//   filter: a warp takes 32 pairs, ONE PAIR PER LANE with both oligos' planes in registers, and sweeps the
//           entries (broadcast loads): does F (R) reach its match_words threshold on any plus entry, on any
//           minus entry?  A pair can only amplify this sequence if F+ & R- or R+ & F- (pcr_assay.cpp:37-59).
//           ~14 instructions per (32 pairs x entry); all but ~1 in 10^3 (pair, sequence) combinations end here.
//   exact:  the surviving pairs of the warp are scored one after the other by the whole warp
//           (amplicon_pass: geometry, has_split, identities).
//   VARIANT: `base` holds the assays whose candidate lists are used (filter + membership), `oligos` the trial oligos.
template <bool VARIANT>
__global__ void __launch_bounds__(SCORE_THREADS)
score_kernel(SeqDev sd, const uint4 *__restrict__ g_pl, const int32_t *__restrict__ g_loc, const uint32_t *__restrict__ g_strand,
	const uint32_t *__restrict__ seq_off2, const OligoDev *__restrict__ oligos, const OligoDev *__restrict__ base, uint32_t n_pairs,
	float detect, int amp_min, int amp_max, int taq, uint32_t *bits_any, uint32_t *bits_pass1, uint32_t n_words)
{
	__shared__ ScoreEntry s_ent[SCORE_SMEM_ENTRIES];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = SCORE_THREADS / 32;
	const uint32_t n_chunks = (n_pairs + 31u) / 32u;
	for (uint32_t seq = blockIdx.x; seq < sd.n; seq += gridDim.x) {
		const uint32_t e0 = seq_off2[2 * seq], Ep = seq_off2[2 * seq + 1] - e0, E = seq_off2[2 * seq + 2] - e0;
		if (Ep == 0u || Ep == E || !sd.active[seq]) continue; // needs both strands; inactive: optimize.cpp:280-283
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < min(E, (uint32_t)SCORE_SMEM_ENTRIES); i += SCORE_THREADS) {
			ScoreEntry en;
			const uint4 v = g_pl[e0 + i];
			en.a = v.x; en.c = v.y; en.g = v.z; en.t = v.w;
			en.loc = g_loc[e0 + i]; en.strand = g_strand[e0 + i];
			s_ent[i] = en;
		}
		__syncthreads();
		for (uint32_t chunk = warp; chunk < n_chunks; chunk += n_warps) {
			const uint32_t p = chunk * 32u + lane;
			OligoDev F, R;
			F.a = F.c = F.g = F.t = R.a = R.c = R.g = R.t = 0u;
			F.norm = R.norm = 0.0f;
			F.packed = R.packed = 255u; // threshold nothing reaches
			if (p < n_pairs) { F = (VARIANT ? base : oligos)[2 * p]; R = (VARIANT ? base : oligos)[2 * p + 1]; }
			const int f_thr = (int)(F.packed & 255u), r_thr = (int)(R.packed & 255u);
			bool fp = false, rp = false, fm = false, rm = false;
			for (uint32_t e = 0; e < Ep; ++e) { // plus-strand entries
				const ScoreEntry en = load_entry(s_ent, g_pl, g_loc, g_strand, e0, e);
				fp |= oligo_count(F, en) >= f_thr;
				rp |= oligo_count(R, en) >= r_thr;
			}
			for (uint32_t e = Ep; e < E; ++e) { // minus-strand entries
				const ScoreEntry en = load_entry(s_ent, g_pl, g_loc, g_strand, e0, e);
				fm |= oligo_count(F, en) >= f_thr;
				rm |= oligo_count(R, en) >= r_thr;
			}
			uint32_t todo = __ballot_sync(0xffffffffu, (fp && rm) || (rp && fm));
			while (todo) {
				const uint32_t src = (uint32_t)__ffs(todo) - 1u;
				todo &= todo - 1u;
				const uint32_t q = chunk * 32u + src;
				const OligoDev Fq = oligos[2 * q], Rq = oligos[2 * q + 1];
				const OligoDev Fb = VARIANT ? base[2 * q] : Fq, Rb = VARIANT ? base[2 * q + 1] : Rq;
				// {F(+), R(-)} then {R(+), F(-)}  (pcr_assay.cpp:37-59)
				const bool d1 = amplicon_pass<VARIANT>(sd, seq, s_ent, g_pl, g_loc, g_strand, e0, Ep, E, Fq, Rq, Fb, Rb, detect, amp_min, amp_max, taq, lane);
				const bool d2 = d1 ? false
				                   : amplicon_pass<VARIANT>(sd, seq, s_ent, g_pl, g_loc, g_strand, e0, Ep, E, Rq, Fq, Rb, Fb, detect, amp_min, amp_max, taq, lane);
				if (lane == 0u && (d1 || d2)) {
					const uint32_t bit = 1u << (seq & 31u);
					atomicOr(bits_any + (size_t)q * n_words + (seq >> 5), bit);
					if (d1) atomicOr(bits_pass1 + (size_t)q * n_words + (seq >> 5), bit);
				}
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------------
// K2, key-matrix form (the default): the filter question "does oligo o reach its match_words threshold on some plus
// (minus) entry of sequence s?" is answered per UNIQUE database word instead of per entry.  Sequences of a collection
// share most of their words (the database of 2.7 M entries of the bench holds ~10x fewer keys), and match_words itself
// runs over keys (optimize.cpp:291-301):
//   key_match_kernel   bit (key, oligo) = (oligo & key) >= unsigned(size * thr^2)         one compare per (key, oligo)
//   seq_filter_kernel  per sequence, OR the key rows of its plus entries and of its minus entries; a pair can only amplify
//                      the sequence if F(+) & R(-) or R(+) & F(-) (pcr_assay.cpp:37-59): those (sequence, pair) items
//                      go to a work list (~1 in 10^3 of all combinations)
//   score_items_kernel one warp per item: the exact amplicon test of amplicon_pass (geometry, has_split, identities)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int KEYM_THREADS = 128;

// grid: (ceil(n_keys / KEYM_THREADS), n_words); block: one 32-oligo word against KEYM_THREADS keys
__global__ void __launch_bounds__(KEYM_THREADS)
key_match_kernel(const uint4 *__restrict__ key_planes, uint32_t n_keys, const OligoDev *__restrict__ member, uint32_t n_oligos, uint32_t n_words,
	uint32_t *keybits)
{
	__shared__ OligoDev s_ol[32];
	const uint32_t w = blockIdx.y;
	if (threadIdx.x < 32u) {
		const uint32_t o = w * 32u + threadIdx.x;
		OligoDev d;
		d.a = d.c = d.g = d.t = 0u;
		d.norm = 0.0f;
		d.packed = 255u; // a threshold nothing reaches
		if (o < n_oligos) d = member[o];
		s_ol[threadIdx.x] = d;
	}
	__syncthreads();
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_keys) return;
	const uint4 kp = __ldg(key_planes + k);
	ScoreEntry en;
	en.a = kp.x; en.c = kp.y; en.g = kp.z; en.t = kp.w;
	en.loc = 0; en.strand = 0u;
	uint32_t bits = 0u;
	#pragma unroll 8
	for (uint32_t b = 0; b < 32u; ++b) {
		const OligoDev &o = s_ol[b];
		bits |= (uint32_t)(oligo_count(o, en) >= (int)(o.packed & 255u)) << b;
	}
	keybits[(size_t)k * n_words + w] = bits;
}

// --- seed-table form of the same filter (fst.cuh): one thread per database entry finds the few oligos that reach their
//     threshold on its word and sets (sequence, strand, oligo) bits; pairs are then read off those rows directly
__global__ void oligo_split_kernel(const OligoDev *__restrict__ o, uint32_t n, uint4 *planes, uint32_t *thr)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const OligoDev d = o[i];
	planes[i] = make_uint4(d.a, d.c, d.g, d.t);
	thr[i] = d.packed & 255u;
}

__global__ void __launch_bounds__(128)
entry_match_kernel(Fst t, const uint4 *__restrict__ e_planes, const uint32_t *__restrict__ e_seq, const uint32_t *__restrict__ e_strand, uint64_t n_ent,
	uint32_t n_words, uint32_t *seqbits)
{
	const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n_ent) return;
	const uint4 kp = __ldg(e_planes + e);
	const FstWord w = fst_word(kp.x, kp.y, kp.z, kp.w);
	uint32_t *row = seqbits + ((size_t)__ldg(e_seq + e) * 2u + (__ldg(e_strand + e) == STRAND_MINUS ? 1u : 0u)) * n_words;
	fst_match<false>(t, w, [&](uint32_t id, uint32_t) { atomicOr(row + (id >> 5), 1u << (id & 31u)); });
}

// ---- neighbour filter: match_words without a table walk ---------------------------------------------------------------
// Every database entry W exists because some candidate word K of the select_words call reached its seed threshold on W
// in-frame (at most e1 = size(K) - thr(K) slots of K miss).  If an oligo O also reaches ITS threshold on W (at most
// e2 = size(O) - thr(O) misses), then on the slots K and O share, the letter of W is in both sets wherever neither misses, so
// K and O have an empty intersection on at most e1 + e2 of their common slots.  (W must carry single letters for this: entries
// holding a degenerate text base are compared with every oligo, as in the table walk.)  So the oligos that can match W are
// among the "neighbours" of ONE candidate recorded for W (SeqSet::e_cand): a candidates x oligos comparison per batch
// (2000 x 2000 set-intersection counts) leaves a short list per candidate -- the oligo itself, its shifted family, oligos cut
// from the same place of a related target -- and every entry verifies only those.
__global__ void __launch_bounds__(256) neigh_pairs_kernel(const uint4 *__restrict__ c_planes, const uint32_t *__restrict__ c_thr, uint32_t n_cand,
	const OligoDev *__restrict__ olig, uint32_t n_olig, unsigned long long *pairs, unsigned int *count, uint32_t cap)
{
	__shared__ OligoDev s_o[256];
	const uint32_t o0 = blockIdx.y * 256u;
	const uint32_t no = min(256u, n_olig - o0);
	if (threadIdx.x < no) s_o[threadIdx.x] = olig[o0 + threadIdx.x];
	__syncthreads();
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_cand) return;
	const uint4 kp = c_planes[k];
	const uint32_t kocc = kp.x | kp.y | kp.z | kp.w;
	const uint32_t kn = (uint32_t)__popc(kocc), kt = c_thr[k];
	if (kt > kn) return; // a candidate that cannot reach its threshold produced no entry
	const uint32_t e1 = kn - kt;
	for (uint32_t j = 0; j < no; ++j) {
		const OligoDev &o = s_o[j];
		const uint32_t oocc = o.a | o.c | o.g | o.t;
		const uint32_t on = (uint32_t)__popc(oocc), ot = o.packed & 255u;
		if (ot > on) continue; // can never match (match_words compares against unsigned(size * thr^2))
		const uint32_t inter = (kp.x & o.a) | (kp.y & o.c) | (kp.z & o.g) | (kp.w & o.t);
		if ((uint32_t)__popc(kocc & oocc & ~inter) <= e1 + (on - ot)) {
			const unsigned int at = atomicAdd(count, 1u);
			if (at < cap) pairs[at] = ((unsigned long long)k << 32) | (o0 + j);
		}
	}
}

__global__ void neigh_offsets_kernel(const unsigned long long *__restrict__ pairs, uint32_t n_pairs, uint32_t n_cand, uint32_t *off)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k > n_cand) return;
	const unsigned long long want = (unsigned long long)k << 32;
	uint32_t lo = 0, hi = n_pairs;
	while (lo < hi) {
		const uint32_t mid = (lo + hi) >> 1;
		if (pairs[mid] < want) lo = mid + 1; else hi = mid;
	}
	off[k] = lo;
}

__global__ void __launch_bounds__(128)
entry_neigh_kernel(const unsigned long long *__restrict__ pairs, const uint32_t *__restrict__ off, const OligoDev *__restrict__ olig, uint32_t n_olig,
	const uint4 *__restrict__ e_planes, const uint32_t *__restrict__ e_seq, const uint32_t *__restrict__ e_strand, const uint32_t *__restrict__ e_cand,
	uint64_t n_ent, uint32_t n_words, uint32_t *seqbits)
{
	const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n_ent) return;
	const uint4 w = __ldg(e_planes + e);
	uint32_t *row = seqbits + ((size_t)__ldg(e_seq + e) * 2u + (__ldg(e_strand + e) == STRAND_MINUS ? 1u : 0u)) * n_words;
	auto verify = [&](uint32_t id) {
		const OligoDev o = olig[id];
		if ((uint32_t)__popc((o.a & w.x) | (o.c & w.y) | (o.g & w.z) | (o.t & w.w)) >= (o.packed & 255u)) atomicOr(row + (id >> 5), 1u << (id & 31u));
	};
	const uint32_t multi = (w.x & w.y) | (w.x & w.z) | (w.x & w.w) | (w.y & w.z) | (w.y & w.w) | (w.z & w.w);
	if (multi) { // a degenerate text base: the bound above does not hold, compare with everybody
		for (uint32_t id = 0; id < n_olig; ++id) verify(id);
		return;
	}
	const uint32_t k = __ldg(e_cand + e);
	for (uint32_t i = __ldg(off + k), i1 = __ldg(off + k + 1u); i < i1; ++i) verify((uint32_t)pairs[i]);
}

// ---- entry-driven scoring (the default): no bit rows, no item list, no host round trip --------------------------------------------
// A (pair, sequence) is amplified iff some plus-strand entry p and some minus-strand entry m of that sequence pass the test of
// find_amplicon_match (pcr_assay.cpp:338-441) with {F on p, R on m} (pass 1) or {R on p, F on m} (pass 2) -- a predicate over
// PAIRS OF ENTRIES, whatever order the reference meets them in.  So one thread takes one plus-strand entry, finds the oligos that
// reach their match_words threshold on it among the neighbours of the entry's candidate (the filter above; a fixed number of
// slots per candidate, filled with one atomic each -- a candidate with more neighbours, or an entry holding a degenerate text
// base, is compared with every oligo), and for each such oligo looks for its partner among the minus-strand entries of the SAME
// sequence that can close an amplicon: the full-window entries of a (sequence, strand) run are sorted by loc (entry ids sort by
// position, db.cuh), and amplicon_max bounds m.loc - p.loc, so that is a binary search plus a handful of entries; the few
// partial-word entries at the end of the run are all tried.  The exact test itself is amplicon_pass's, line by line.
constexpr uint32_t NEIGH_SLOTS = 24u;

__global__ void __launch_bounds__(256) neigh_slots_kernel(const uint4 *__restrict__ c_planes, const uint32_t *__restrict__ c_thr, uint32_t n_cand,
	const OligoDev *__restrict__ olig, uint32_t n_olig, uint32_t *slots, uint32_t *cnt)
{
	__shared__ OligoDev s_o[256];
	const uint32_t o0 = blockIdx.y * 256u;
	const uint32_t no = min(256u, n_olig - o0);
	if (threadIdx.x < no) s_o[threadIdx.x] = olig[o0 + threadIdx.x];
	__syncthreads();
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_cand) return;
	const uint4 kp = c_planes[k];
	const uint32_t kocc = kp.x | kp.y | kp.z | kp.w;
	const uint32_t kn = (uint32_t)__popc(kocc), kt = c_thr[k];
	if (kt > kn) return; // a candidate that cannot reach its threshold produced no entry
	const uint32_t e1 = kn - kt;
	for (uint32_t j = 0; j < no; ++j) {
		const OligoDev &o = s_o[j];
		const uint32_t oocc = o.a | o.c | o.g | o.t;
		const uint32_t on = (uint32_t)__popc(oocc), ot = o.packed & 255u;
		if (ot > on) continue; // can never match (match_words compares against unsigned(size * thr^2))
		const uint32_t inter = (kp.x & o.a) | (kp.y & o.c) | (kp.z & o.g) | (kp.w & o.t);
		if ((uint32_t)__popc(kocc & oocc & ~inter) <= e1 + (on - ot)) {
			const uint32_t at = atomicAdd(cnt + k, 1u);
			if (at < NEIGH_SLOTS) slots[(size_t)k * NEIGH_SLOTS + at] = o0 + j;
		}
	}
}

// first entry of each (sequence, strand) run that is not a full-window entry (entry ids: seq | minus | type | pos, db.cuh)
__global__ void seq_full_end_kernel(const uint64_t *__restrict__ entry_id, uint64_t n, uint32_t n_seq, uint32_t pb, uint32_t *full_end)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= 2u * n_seq) return;
	const uint64_t want = ((uint64_t)k << 2) | 1ull; // (seq, minus) then type >= 1
	uint64_t lo = 0, hi = n;
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		if ((entry_id[mid] >> pb) < want) lo = mid + 1; else hi = mid;
	}
	full_end[k] = (uint32_t)lo;
}

template <bool VARIANT>
__global__ void __launch_bounds__(128)
score_entries_kernel(SeqDev sd, const uint4 *__restrict__ e_planes, const int32_t *__restrict__ e_loc, const uint32_t *__restrict__ e_strand,
	const uint32_t *__restrict__ e_seq, const uint32_t *__restrict__ e_cand, uint64_t n_ent, const uint32_t *__restrict__ seq_off2,
	const uint32_t *__restrict__ full_end, const uint32_t *__restrict__ slots, const uint32_t *__restrict__ slot_cnt,
	const OligoDev *__restrict__ member, const OligoDev *__restrict__ oligos, uint32_t n_olig, float detect, int amp_min, int amp_max, int taq,
	uint32_t *bits_any, uint32_t *bits_pass1, uint32_t n_words_seq)
{
	const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n_ent) return;
	if (__ldg(e_strand + e) != STRAND_PLUS) return;
	const uint32_t seq = __ldg(e_seq + e);
	if (!sd.active[seq]) return; // optimize.cpp:280-283
	const uint32_t m0 = __ldg(seq_off2 + 2u * seq + 1u), m1 = __ldg(seq_off2 + 2u * seq + 2u);
	if (m0 == m1) return; // no minus-strand entry: nothing can pair
	const uint32_t mf = __ldg(full_end + 2u * seq + 1u);
	const uint4 w = __ldg(e_planes + e);
	ScoreEntry pe;
	pe.a = w.x; pe.c = w.y; pe.g = w.z; pe.t = w.w;
	pe.loc = __ldg(e_loc + e);
	pe.strand = STRAND_PLUS;
	const int L = (int)sd.len[seq];
	auto try_oligo = [&](uint32_t id) {
		const OligoDev Pb = member[id];
		const int cpb = oligo_count(Pb, pe);
		if (cpb < (int)(Pb.packed & 255u)) return;
		const uint32_t partner = id ^ 1u;
		const OligoDev Mb = member[partner];
		const OligoDev P = VARIANT ? oligos[id] : Pb, M = VARIANT ? oligos[partner] : Mb;
		const int cnt_p = VARIANT ? oligo_count(P, pe) : cpb;
		const int p_start = (int)((Pb.packed >> 8) & 255u), p_stop = (int)((Pb.packed >> 16) & 255u);
		const int m_thr = (int)(Mb.packed & 255u), m_start = (int)((Mb.packed >> 8) & 255u), m_stop = (int)((Mb.packed >> 16) & 255u);
		const int plus_loc3 = pe.loc + p_stop; // sequence.h:67-75, plus strand
		const float ident_p = oligo_identity(P, cnt_p, pe, taq);
		bool found = false;
		auto test = [&](uint32_t j) {
			ScoreEntry m2;
			const uint4 v = __ldg(e_planes + j);
			m2.a = v.x; m2.c = v.y; m2.g = v.z; m2.t = v.w;
			m2.loc = __ldg(e_loc + j);
			m2.strand = STRAND_MINUS;
			const int cm = oligo_count(M, m2);
			if ((VARIANT ? oligo_count(Mb, m2) : cm) < m_thr) return;
			const int minus_loc5 = m2.loc - m_stop; // sequence.h:57-65, minus strand
			if (!(plus_loc3 < minus_loc5)) return;   // pcr_assay.cpp:368-371
			int amp_start = pe.loc + p_start;
			const int amp_stop = min(m2.loc - m_start, L - 1);
			int amp_len = amp_stop - amp_start + 1;
			if (amp_len < amp_min || amp_len > amp_max) return; // :383-392
			if (amp_start < 0) { amp_len += amp_start; amp_start = 0; } // :412-416
			if (amp_len >= 0 && !has_split_dev(sd, seq, amp_start, amp_len)) { // :418
				const float ident_m = oligo_identity(M, cm, m2, taq);
				if (__fsqrt_rn(__fmul_rn(ident_p, ident_m)) >= detect) found = true; // :292-294
			}
		};
		// full-window entries: m.loc > plus_loc3 + m_stop, and amp_len <= amp_max bounds m.loc - p.loc by amp_max + 62 (both frame offsets
		// are below 32; a clipped amplicon ends at L - 1 >= m.loc - 31)
		{
			const int lo_loc = plus_loc3 + m_stop, hi_loc = pe.loc + max(amp_max, 0) + 96;
			uint32_t lo = m0, hi = mf;
			while (lo < hi) {
				const uint32_t mid = (lo + hi) >> 1;
				if (__ldg(e_loc + mid) <= lo_loc) lo = mid + 1; else hi = mid;
			}
			for (uint32_t j = lo; j < mf && !found; ++j) {
				if (__ldg(e_loc + j) > hi_loc) break;
				test(j);
			}
		}
		for (uint32_t j = mf; j < m1 && !found; ++j) test(j); // the partial words of the run
		if (found) {
			const uint32_t q = id >> 1, bit = 1u << (seq & 31u);
			atomicOr(bits_any + (size_t)q * n_words_seq + (seq >> 5), bit);
			if (!(id & 1u)) atomicOr(bits_pass1 + (size_t)q * n_words_seq + (seq >> 5), bit); // {F(+), R(-)}: pass 1 (pcr_assay.cpp:37-47)
		}
	};
	const uint32_t multi = (w.x & w.y) | (w.x & w.z) | (w.x & w.w) | (w.y & w.z) | (w.y & w.w) | (w.z & w.w);
	const uint32_t k = __ldg(e_cand + e);
	const uint32_t nk = __ldg(slot_cnt + k);
	if (multi || nk > NEIGH_SLOTS) { // a degenerate text base (the neighbour bound does not hold) or more neighbours than slots: everybody
		for (uint32_t id = 0; id < n_olig; ++id) try_oligo(id);
		return;
	}
	for (uint32_t i = 0; i < nk; ++i) try_oligo(__ldg(slots + (size_t)k * NEIGH_SLOTS + i));
}

// Move scoring by groups (optimize_pcr.cpp: every trial oligo of a move is scored against the candidates of the UNMOVED assay): the
// variants [goff[g], goff[g + 1]) share base assay g = (member[2g], member[2g + 1]).  Membership (which entries an oligo of the base
// assay matches), the partner entries and the amplicon geometry belong to the base assay and are found once per (entry, base oligo);
// only the two identities and the detection test are the variant's.  The neighbour slots are built over the 2G base oligos, so their
// number does not grow with the variants.
__global__ void __launch_bounds__(128)
score_entries_groups_kernel(SeqDev sd, const uint4 *__restrict__ e_planes, const int32_t *__restrict__ e_loc, const uint32_t *__restrict__ e_strand,
	const uint32_t *__restrict__ e_seq, const uint32_t *__restrict__ e_cand, uint64_t n_ent, const uint32_t *__restrict__ seq_off2,
	const uint32_t *__restrict__ full_end, const uint32_t *__restrict__ slots, const uint32_t *__restrict__ slot_cnt,
	const OligoDev *__restrict__ member, uint32_t n_member, const uint32_t *__restrict__ goff, const OligoDev *__restrict__ oligos, float detect,
	int amp_min, int amp_max, int taq, uint32_t *bits_any, uint32_t *bits_pass1, uint32_t n_words_seq)
{
	const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n_ent) return;
	if (__ldg(e_strand + e) != STRAND_PLUS) return;
	const uint32_t seq = __ldg(e_seq + e);
	if (!sd.active[seq]) return; // optimize.cpp:280-283
	const uint32_t m0 = __ldg(seq_off2 + 2u * seq + 1u), m1 = __ldg(seq_off2 + 2u * seq + 2u);
	if (m0 == m1) return;
	const uint32_t mf = __ldg(full_end + 2u * seq + 1u);
	const uint4 w = __ldg(e_planes + e);
	ScoreEntry pe;
	pe.a = w.x; pe.c = w.y; pe.g = w.z; pe.t = w.w;
	pe.loc = __ldg(e_loc + e);
	pe.strand = STRAND_PLUS;
	const int L = (int)sd.len[seq];
	const uint32_t bit = 1u << (seq & 31u), word = seq >> 5;
	auto try_base = [&](uint32_t id) {
		const OligoDev Pb = member[id];
		if (oligo_count(Pb, pe) < (int)(Pb.packed & 255u)) return;
		const uint32_t side = id & 1u, g = id >> 1;
		const OligoDev Mb = member[id ^ 1u];
		const uint32_t v0 = __ldg(goff + g), v1 = __ldg(goff + g + 1u);
		if (v0 == v1) return;
		const int p_start = (int)((Pb.packed >> 8) & 255u), p_stop = (int)((Pb.packed >> 16) & 255u);
		const int m_thr = (int)(Mb.packed & 255u), m_start = (int)((Mb.packed >> 8) & 255u), m_stop = (int)((Mb.packed >> 16) & 255u);
		const int plus_loc3 = pe.loc + p_stop;
		auto test = [&](uint32_t j) {
			ScoreEntry m2;
			const uint4 v = __ldg(e_planes + j);
			m2.a = v.x; m2.c = v.y; m2.g = v.z; m2.t = v.w;
			m2.loc = __ldg(e_loc + j);
			m2.strand = STRAND_MINUS;
			if (oligo_count(Mb, m2) < m_thr) return;
			const int minus_loc5 = m2.loc - m_stop;
			if (!(plus_loc3 < minus_loc5)) return;   // pcr_assay.cpp:368-371
			int amp_start = pe.loc + p_start;
			const int amp_stop = min(m2.loc - m_start, L - 1);
			int amp_len = amp_stop - amp_start + 1;
			if (amp_len < amp_min || amp_len > amp_max) return; // :383-392
			if (amp_start < 0) { amp_len += amp_start; amp_start = 0; } // :412-416
			if (amp_len < 0 || has_split_dev(sd, seq, amp_start, amp_len)) return; // :418
			for (uint32_t q = v0; q < v1; ++q) { // the variants of this base assay
				const OligoDev P = oligos[2u * q + side], M = oligos[2u * q + (side ^ 1u)];
				const float ident_p = oligo_identity(P, oligo_count(P, pe), pe, taq);
				const float ident_m = oligo_identity(M, oligo_count(M, m2), m2, taq);
				if (__fsqrt_rn(__fmul_rn(ident_p, ident_m)) >= detect) { // :292-294
					uint32_t *a = bits_any + (size_t)q * n_words_seq + word;
					if (!(*a & bit)) atomicOr(a, bit);
					if (!side) { // {F(+), R(-)}: pass 1 (pcr_assay.cpp:37-47)
						uint32_t *b = bits_pass1 + (size_t)q * n_words_seq + word;
						if (!(*b & bit)) atomicOr(b, bit);
					}
				}
			}
		};
		{
			const int lo_loc = plus_loc3 + m_stop, hi_loc = pe.loc + max(amp_max, 0) + 96;
			uint32_t lo = m0, hi = mf;
			while (lo < hi) {
				const uint32_t mid = (lo + hi) >> 1;
				if (__ldg(e_loc + mid) <= lo_loc) lo = mid + 1; else hi = mid;
			}
			for (uint32_t j = lo; j < mf; ++j) {
				if (__ldg(e_loc + j) > hi_loc) break;
				test(j);
			}
		}
		for (uint32_t j = mf; j < m1; ++j) test(j); // the partial words of the run
	};
	const uint32_t multi = (w.x & w.y) | (w.x & w.z) | (w.x & w.w) | (w.y & w.z) | (w.y & w.w) | (w.z & w.w);
	const uint32_t k = __ldg(e_cand + e);
	const uint32_t nk = __ldg(slot_cnt + k);
	if (multi || nk > NEIGH_SLOTS) {
		for (uint32_t id = 0; id < n_member; ++id) try_base(id);
		return;
	}
	for (uint32_t i = 0; i < nk; ++i) try_base(__ldg(slots + (size_t)k * NEIGH_SLOTS + i));
}

struct ScoreItem;
__global__ void seq_pairs_kernel(SeqDev sd, const uint32_t *__restrict__ seq_off2, const uint32_t *__restrict__ seqbits, uint32_t n_words, uint32_t n_pairs,
	ScoreItem *items, unsigned int *n_items, uint32_t cap);

struct ScoreItem {
	uint32_t seq, pair; // pair[29:0] | try pass 1 [30] | try pass 2 [31]
};

// grid: sequences (strided); block: n_words threads (rounded up to a warp), thread w ORs word w over the sequence's entries
__global__ void seq_filter_kernel(SeqDev sd, const uint32_t *__restrict__ seq_off2, const uint32_t *__restrict__ e_key, const uint32_t *__restrict__ keybits,
	uint32_t n_words, uint32_t n_pairs, ScoreItem *items, unsigned int *n_items, uint32_t cap)
{
	for (uint32_t seq = blockIdx.x; seq < sd.n; seq += gridDim.x) {
		const uint32_t e0 = seq_off2[2 * seq], ep = seq_off2[2 * seq + 1], e1 = seq_off2[2 * seq + 2];
		if (ep == e0 || ep == e1 || !sd.active[seq]) continue; // needs both strands; inactive: optimize.cpp:280-283
		for (uint32_t w = threadIdx.x; w < n_words; w += blockDim.x) {
			uint32_t plus = 0u, minus = 0u;
			for (uint32_t e = e0; e < ep; ++e) plus |= __ldg(keybits + (size_t)__ldg(e_key + e) * n_words + w);
			for (uint32_t e = ep; e < e1; ++e) minus |= __ldg(keybits + (size_t)__ldg(e_key + e) * n_words + w);
			// oligo 2p = F of pair p, 2p + 1 = R: even bits F, odd bits R
			const uint32_t pass1 = (plus & 0x55555555u) & ((minus >> 1) & 0x55555555u);   // F(+) & R(-)
			const uint32_t pass2 = ((plus >> 1) & 0x55555555u) & (minus & 0x55555555u);   // R(+) & F(-)
			uint32_t any = pass1 | pass2;
			while (any) {
				const uint32_t b = (uint32_t)__ffs(any) - 1u;
				any &= any - 1u;
				const uint32_t pair = (w * 32u + b) >> 1;
				if (pair < n_pairs) {
					const unsigned int i = atomicAdd(n_items, 1u);
					if (i < cap) {
						ScoreItem it;
						it.seq = seq;
						it.pair = pair | (((pass1 >> b) & 1u) << 30) | (((pass2 >> b) & 1u) << 31);
						items[i] = it;
					}
				}
			}
		}
	}
}

__global__ void seq_pairs_kernel(SeqDev sd, const uint32_t *__restrict__ seq_off2, const uint32_t *__restrict__ seqbits, uint32_t n_words, uint32_t n_pairs,
	ScoreItem *items, unsigned int *n_items, uint32_t cap)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (tid >= (uint64_t)sd.n * n_words) return;
	const uint32_t seq = (uint32_t)(tid / n_words), w = (uint32_t)(tid % n_words);
	const uint32_t e0 = seq_off2[2 * seq], ep = seq_off2[2 * seq + 1], e1 = seq_off2[2 * seq + 2];
	if (ep == e0 || ep == e1 || !sd.active[seq]) return; // needs both strands; inactive: optimize.cpp:280-283
	const uint32_t plus = seqbits[(size_t)(2u * seq) * n_words + w], minus = seqbits[(size_t)(2u * seq + 1u) * n_words + w];
	const uint32_t pass1 = (plus & 0x55555555u) & ((minus >> 1) & 0x55555555u); // F(+) & R(-): even bits F, odd bits R
	const uint32_t pass2 = ((plus >> 1) & 0x55555555u) & (minus & 0x55555555u); // R(+) & F(-)
	uint32_t any = pass1 | pass2;
	while (any) {
		const uint32_t b = (uint32_t)__ffs(any) - 1u;
		any &= any - 1u;
		const uint32_t pair = (w * 32u + b) >> 1;
		if (pair < n_pairs) {
			// one atomic per group of lanes that are here together (~10^6 items per batch on ONE counter otherwise)
			const unsigned am = __activemask(), ln = threadIdx.x & 31u;
			const int leader = __ffs(am) - 1;
			unsigned int i = 0;
			if ((int)ln == leader) i = atomicAdd(n_items, (unsigned int)__popc(am));
			i = __shfl_sync(am, i, leader) + (unsigned int)__popc(am & ((1u << ln) - 1u));
			if (i < cap) {
				ScoreItem it;
				it.seq = seq;
				it.pair = pair | (((pass1 >> b) & 1u) << 30) | (((pass2 >> b) & 1u) << 31);
				items[i] = it;
			}
		}
	}
}

template <bool VARIANT>
__global__ void __launch_bounds__(256)
score_items_kernel(SeqDev sd, const uint4 *__restrict__ g_pl, const int32_t *__restrict__ g_loc, const uint32_t *__restrict__ g_strand,
	const uint32_t *__restrict__ seq_off2, const OligoDev *__restrict__ oligos, const OligoDev *__restrict__ base, const ScoreItem *__restrict__ items,
	const unsigned int *__restrict__ n_items, uint32_t cap, float detect, int amp_min, int amp_max, int taq, uint32_t *bits_any, uint32_t *bits_pass1,
	uint32_t n_words_seq)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
	const uint32_t total = min(*n_items, cap);
	for (uint32_t i = warp; i < total; i += n_warps) {
		const ScoreItem it = items[i];
		const uint32_t seq = it.seq, q = it.pair & 0x3FFFFFFFu;
		const uint32_t e0 = seq_off2[2 * seq], Ep = seq_off2[2 * seq + 1] - e0, E = seq_off2[2 * seq + 2] - e0;
		const OligoDev Fq = oligos[2 * q], Rq = oligos[2 * q + 1];
		const OligoDev Fb = VARIANT ? base[2 * q] : Fq, Rb = VARIANT ? base[2 * q + 1] : Rq;
		const bool d1 = ((it.pair >> 30) & 1u)
		                    ? amplicon_pass<VARIANT>(sd, seq, nullptr, g_pl, g_loc, g_strand, e0, Ep, E, Fq, Rq, Fb, Rb, detect, amp_min, amp_max, taq, lane)
		                    : false;
		const bool d2 = (d1 || !((it.pair >> 31) & 1u))
		                    ? false
		                    : amplicon_pass<VARIANT>(sd, seq, nullptr, g_pl, g_loc, g_strand, e0, Ep, E, Rq, Fq, Rb, Fb, detect, amp_min, amp_max, taq, lane);
		if (lane == 0u && (d1 || d2)) {
			const uint32_t bit = 1u << (seq & 31u);
			atomicOr(bits_any + (size_t)q * n_words_seq + (seq >> 5), bit);
			if (d1) atomicOr(bits_pass1 + (size_t)q * n_words_seq + (seq >> 5), bit);
		}
	}
}

// seq_pairs_kernel + score_items_kernel in one: a CTA takes a sequence, stages its database entries in shared memory once
// (~130 entries, 24 bytes each), and its warps walk the sequence's two (strand, oligo) bit rows; every pair with F(+) & R(-) or
// R(+) & F(-) gets the exact amplicon test straight away, against the staged entries.  The item list, the host round trip that
// sized it, and the per-item re-reads of the entries through L2 (score_items_kernel: 0.34 ms, long-scoreboard bound) go away.
// MEASURED SLOWER on the bench batch (pair scoring 0.65 ms against 0.60 ms with the item list): ~50 items per sequence over 8
// warps and two barriers per sequence leave the SMs emptier than one warp per item over 10^6 items does.  Option
// "use_fused_score" (default off); tests/test_gpu_parity.py::test_fused_sequence_scoring_equals_item_list keeps it honest.
template <bool VARIANT>
__global__ void __launch_bounds__(SCORE_THREADS)
score_seqbits_kernel(SeqDev sd, const uint4 *__restrict__ g_pl, const int32_t *__restrict__ g_loc, const uint32_t *__restrict__ g_strand,
	const uint32_t *__restrict__ seq_off2, const OligoDev *__restrict__ oligos, const OligoDev *__restrict__ base, const uint32_t *__restrict__ seqbits,
	uint32_t n_words, uint32_t n_pairs, float detect, int amp_min, int amp_max, int taq, uint32_t *bits_any, uint32_t *bits_pass1, uint32_t n_words_seq)
{
	__shared__ ScoreEntry s_ent[SCORE_SMEM_ENTRIES];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = SCORE_THREADS / 32;
	for (uint32_t seq = blockIdx.x; seq < sd.n; seq += gridDim.x) {
		const uint32_t e0 = seq_off2[2 * seq], Ep = seq_off2[2 * seq + 1] - e0, E = seq_off2[2 * seq + 2] - e0;
		if (Ep == 0u || Ep == E || !sd.active[seq]) continue; // needs both strands; inactive: optimize.cpp:280-283
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < min(E, (uint32_t)SCORE_SMEM_ENTRIES); i += SCORE_THREADS) {
			ScoreEntry en;
			const uint4 v = g_pl[e0 + i];
			en.a = v.x; en.c = v.y; en.g = v.z; en.t = v.w;
			en.loc = g_loc[e0 + i]; en.strand = g_strand[e0 + i];
			s_ent[i] = en;
		}
		__syncthreads();
		for (uint32_t w = warp; w < n_words; w += n_warps) {
			const uint32_t plus = __ldg(seqbits + (size_t)(2u * seq) * n_words + w), minus = __ldg(seqbits + (size_t)(2u * seq + 1u) * n_words + w);
			const uint32_t pass1 = (plus & 0x55555555u) & ((minus >> 1) & 0x55555555u); // F(+) & R(-): even bits F, odd bits R
			const uint32_t pass2 = ((plus >> 1) & 0x55555555u) & (minus & 0x55555555u); // R(+) & F(-)
			uint32_t any = pass1 | pass2; // uniform over the warp
			while (any) {
				const uint32_t b = (uint32_t)__ffs(any) - 1u;
				any &= any - 1u;
				const uint32_t q = (w * 32u + b) >> 1;
				if (q >= n_pairs) continue;
				const OligoDev Fq = oligos[2 * q], Rq = oligos[2 * q + 1];
				const OligoDev Fb = VARIANT ? base[2 * q] : Fq, Rb = VARIANT ? base[2 * q + 1] : Rq;
				const bool d1 = ((pass1 >> b) & 1u)
				                    ? amplicon_pass<VARIANT>(sd, seq, s_ent, g_pl, g_loc, g_strand, e0, Ep, E, Fq, Rq, Fb, Rb, detect, amp_min, amp_max, taq, lane)
				                    : false;
				const bool d2 = (d1 || !((pass2 >> b) & 1u))
				                    ? false
				                    : amplicon_pass<VARIANT>(sd, seq, s_ent, g_pl, g_loc, g_strand, e0, Ep, E, Rq, Fq, Rb, Fb, detect, amp_min, amp_max, taq, lane);
				if (lane == 0u && (d1 || d2)) {
					const uint32_t bit = 1u << (seq & 31u);
					atomicOr(bits_any + (size_t)q * n_words_seq + (seq >> 5), bit);
					if (d1) atomicOr(bits_pass1 + (size_t)q * n_words_seq + (seq >> 5), bit);
				}
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

// every weight is 1.0f (the default, sequence.h:22): the double sum of k ones is k whatever the order, so the coverage is
// the population count of the pair's bitset -- one warp per pair instead of one thread walking the words in order
__global__ void __launch_bounds__(256) coverage_count_kernel(const uint32_t *__restrict__ bits_any, uint32_t n_pairs, uint32_t n_words, float *coverage)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (p >= n_pairs) return;
	uint32_t c = 0;
	for (uint32_t w = lane; w < n_words; w += 32u) c += (uint32_t)__popc(bits_any[(size_t)p * n_words + w]);
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if (lane == 0u) coverage[p] = (float)(double)c;
}

} // namespace pcr
