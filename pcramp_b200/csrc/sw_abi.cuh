// sw_abi.cuh -- K4 (included at the end of pcramp_gpu.cu: it shares score.cuh's kernels with K2): SO::SeqOverlap Smith-Waterman batches and the two background tests built on it,
// PCR::find_background_match and PCR::find_multiplex_background_match (background_match.cpp:7-295).
//
//   sw_words_kernel         one thread = one (query word, target word) alignment, with start coordinates
//   amplicon_list_kernel    PCR::collect_candidates + find_amplicon_match for the background thresholds
//                           (pcr_assay.cpp:12-69,338-441): every geometric candidate amplicon of every pair, as
//                           (pair, pass, sequence, plus loc, minus loc) records; CTA per sequence, the same
//                           lane-per-pair filter as score_kernel
//   two stable radix sorts  -> the reference's background_amplicons order: pass {F+,R-} then {R+,F-}, each by
//                           (sequence, plus loc, minus loc)
//   background_sw_kernel    one thread = one candidate amplicon: the four alignments F / rc(F) vs the forward key,
//                           R / rc(R) vs the reverse key, the normalised product score and the sequence's bit
//   multiplex_sw_kernel     one thread = one (pair, sequence, primer strand) alignment against a whole sequence
#pragma once
#include "ctx.cuh"
#include "score.cuh"
#include "sw.cuh"

#include <cub/cub.cuh>

#include <algorithm>

namespace {

int check_kind2(pcramp_gpu_ctx *ctx, int kind)
{
	if (!ctx) return 1;
	if (kind < 0 || kind >= PCRAMP_NUM_KINDS) return fail(ctx, "pcramp_gpu: bad sequence kind");
	return fast_resolve(ctx);
}

__device__ __forceinline__ float taq_correction(unsigned p0, unsigned p1, unsigned t0, unsigned t1)
{ // taq_mama_correction (word.cpp:249-294)
	const int a = taq_index(p0), b = taq_index(p1), c = taq_index(t0), d = taq_index(t1);
	if (a < 0 || b < 0 || c < 0 || d < 0) return 1.0f;
	return fminf(1.0f, c_taq_mama[16 * (4 * d + c) + (4 * b + a)]);
}

__device__ __forceinline__ void word_last_two(const W128 &w, unsigned &p0, unsigned &p1)
{ // Word::get_last_two (word.h:299-305)
	const int last = w_stop(w);
	p0 = last >= 1 ? w_get(w, last - 1) : 0u;
	p1 = last >= 0 ? w_get(w, last) : 0u;
}

// ---- raw SeqOverlap batch -------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) sw_words_kernel(uint32_t n, const uint64_t *__restrict__ query, const uint64_t *__restrict__ target, int *out)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = p < n;
	W128 qw, tw;
	qw.hi = qw.lo = tw.hi = tw.lo = 0ull;
	if (live) {
		qw.hi = query[2 * p]; qw.lo = query[2 * p + 1];
		tw.hi = target[2 * p]; tw.lo = target[2 * p + 1];
	}
	sw::Query q;
	sw::query_from_word(qw, q);
	const sw::WordTarget t(tw);
	const int rows = __reduce_max_sync(0xffffffffu, q.len); // the warp runs the instantiation that covers its longest query
	const sw::Result r = sw::align_rows<true>(q, t, rows);
	if (!live) return;
	unsigned a, b;
	sw::last_two(r, t, a, b);
	// one array per field (what the caller's arrays are): each leaves the device with one copy, nothing is re-packed on the host
	out[p] = r.score;
	out[(size_t)n + p] = r.any ? r.q_start : -1;
	out[2 * (size_t)n + p] = r.any ? r.q_stop : -1;
	out[3 * (size_t)n + p] = r.any ? r.t_start : -1;
	out[4 * (size_t)n + p] = r.any ? r.t_stop : -1;
	reinterpret_cast<uchar2 *>(out + 5 * (size_t)n)[p] = make_uchar2((unsigned char)a, (unsigned char)b);
}

// ---- candidate amplicons --------------------------------------------------------------------------------
struct AmpliconOut {
	uint64_t *key1;   // pair[63:33] | pass[32] | sequence[31:0]
	uint64_t *key2;   // (plus loc + 2^31)[63:32] | (minus loc + 2^31)[31:0]
	uint32_t *plus_id, *minus_id; // database entry ids
	unsigned long long *count;
	uint64_t cap;
};

// EXTRACT = false: find_amplicon_match's rules (pcr_assay.cpp:338-441); EXTRACT = true: extract_amplicon_seq's
// (pcr_assay.cpp:443-542): no clipping to the sequence end, and the test for a split runs over the padded non-primer
// region that becomes the amplicon sequence (amplicon.cuh)
constexpr int MPX_PAD = 4; // MULTIPLEX_AMPLICON_PADDING (pcramp.h:57)
template <bool EXTRACT>
__device__ inline void amplicon_emit_pass(const SeqDev &sd, uint32_t seq, const ScoreEntry *s_ent, const uint4 *__restrict__ g_pl,
	const int32_t *__restrict__ g_loc, const uint32_t *__restrict__ g_strand, uint32_t e0, uint32_t Ep, uint32_t E, const OligoDev &P,
	const OligoDev &M, int amp_min, int amp_max, uint32_t lane, uint32_t pair, uint32_t pass, const AmpliconOut &out)
{
	const int p_thr = (int)(P.packed & 255u), p_start = (int)((P.packed >> 8) & 255u), p_stop = (int)((P.packed >> 16) & 255u);
	const int m_thr = (int)(M.packed & 255u), m_start = (int)((M.packed >> 8) & 255u), m_stop = (int)((M.packed >> 16) & 255u);
	const int L = (int)sd.len[seq];
	for (uint32_t base = 0; base < Ep; base += 32u) {
		const uint32_t e = base + lane;
		ScoreEntry en;
		en.a = en.c = en.g = en.t = 0u; en.loc = 0; en.strand = 0u;
		if (e < Ep) en = load_entry(s_ent, g_pl, g_loc, g_strand, e0, e);
		uint32_t plus_mask = __ballot_sync(0xffffffffu, e < Ep && oligo_count(P, en) >= p_thr);
		while (plus_mask) {
			const int src = __ffs(plus_mask) - 1;
			plus_mask &= plus_mask - 1u;
			const int ploc = __shfl_sync(0xffffffffu, en.loc, src);
			const int plus_loc3 = ploc + p_stop;
			for (uint32_t base2 = Ep; base2 < E; base2 += 32u) {
				const uint32_t e2 = base2 + lane;
				bool ok = false;
				int mloc = 0;
				if (e2 < E) {
					const ScoreEntry m2 = load_entry(s_ent, g_pl, g_loc, g_strand, e0, e2);
					mloc = m2.loc;
					if (EXTRACT) {
						if (oligo_count(M, m2) >= m_thr && plus_loc3 < m2.loc - m_stop) { // pcr_assay.cpp:474-477
							const int amp_len = (m2.loc - m_start) - (ploc + p_start) + 1; // :479-490 (the `break` only prunes)
							if (amp_len >= amp_min && amp_len <= amp_max) {
								// :494-499: the region starts MPX_PAD bases inside the plus primer and its length adds 2 * MPX_PAD
								// to (minus 5' end - region start), i.e. it ends MPX_PAD + 3 bases inside the minus primer
								const int a0 = plus_loc3 + 1 - MPX_PAD, n = (m2.loc - m_stop) - a0 + 2 * MPX_PAD;
								// :505-523 (an EOS also lies inside every longer region: `break` only prunes).  A region that
								// leaves [0, L) is undefined in the reference (unchecked deque index); rejected here
								ok = a0 >= 0 && a0 + n <= L && !has_split_dev(sd, seq, a0, n);
							}
						}
					} else if (oligo_count(M, m2) >= m_thr && plus_loc3 < m2.loc - m_stop) { // pcr_assay.cpp:368-371
						int amp_start = ploc + p_start;
						const int amp_stop = min(m2.loc - m_start, L - 1);
						int amp_len = amp_stop - amp_start + 1;
						if (amp_len >= amp_min && amp_len <= amp_max) { // :383-392 (the `break` only prunes: length grows with the minus loc)
							if (amp_start < 0) { amp_len += amp_start; amp_start = 0; }
							// :418 (a split also stays inside every longer amplicon; a negative length throws in the reference)
							ok = amp_len >= 0 && !has_split_dev(sd, seq, amp_start, amp_len);
						}
					}
				}
				const uint32_t okm = __ballot_sync(0xffffffffu, ok);
				if (okm) {
					unsigned long long pos = 0;
					if (lane == 0u) pos = atomicAdd(out.count, (unsigned long long)__popc(okm));
					pos = __shfl_sync(0xffffffffu, pos, 0);
					if (ok) {
						const uint64_t o = pos + (uint64_t)__popc(okm & ((1u << lane) - 1u));
						if (o < out.cap) {
							out.key1[o] = ((uint64_t)pair << 33) | ((uint64_t)pass << 32) | seq;
							out.key2[o] = ((uint64_t)(uint32_t)(ploc + 0x40000000) << 32) | (uint32_t)(mloc + 0x40000000);
							out.plus_id[o] = e0 + (base + (uint32_t)src);
							out.minus_id[o] = e0 + e2;
						}
					}
				}
			}
		}
	}
}

template <bool EXTRACT>
__global__ void __launch_bounds__(SCORE_THREADS)
amplicon_list_kernel(SeqDev sd, const uint4 *__restrict__ g_pl, const int32_t *__restrict__ g_loc, const uint32_t *__restrict__ g_strand,
	const uint32_t *__restrict__ seq_off2, const OligoDev *__restrict__ oligos, uint32_t n_pairs, int amp_min, int amp_max, AmpliconOut out)
{
	__shared__ ScoreEntry s_ent[SCORE_SMEM_ENTRIES];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = SCORE_THREADS / 32;
	const uint32_t n_chunks = (n_pairs + 31u) / 32u;
	for (uint32_t seq = blockIdx.x; seq < sd.n; seq += gridDim.x) {
		const uint32_t e0 = seq_off2[2 * seq], Ep = seq_off2[2 * seq + 1] - e0, E = seq_off2[2 * seq + 2] - e0;
		if (Ep == 0u || Ep == E || !sd.active[seq]) continue;
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
			F.packed = R.packed = 255u;
			if (p < n_pairs) { F = oligos[2 * p]; R = oligos[2 * p + 1]; }
			const int f_thr = (int)(F.packed & 255u), r_thr = (int)(R.packed & 255u);
			bool fp = false, rp = false, fm = false, rm = false;
			for (uint32_t e = 0; e < Ep; ++e) {
				const ScoreEntry en = load_entry(s_ent, g_pl, g_loc, g_strand, e0, e);
				fp |= oligo_count(F, en) >= f_thr;
				rp |= oligo_count(R, en) >= r_thr;
			}
			for (uint32_t e = Ep; e < E; ++e) {
				const ScoreEntry en = load_entry(s_ent, g_pl, g_loc, g_strand, e0, e);
				fm |= oligo_count(F, en) >= f_thr;
				rm |= oligo_count(R, en) >= r_thr;
			}
			uint32_t todo1 = __ballot_sync(0xffffffffu, fp && rm), todo2 = __ballot_sync(0xffffffffu, rp && fm);
			uint32_t todo = todo1 | todo2;
			while (todo) {
				const uint32_t src = (uint32_t)__ffs(todo) - 1u;
				todo &= todo - 1u;
				const uint32_t q = chunk * 32u + src;
				const OligoDev Fq = oligos[2 * q], Rq = oligos[2 * q + 1];
				if ((todo1 >> src) & 1u)
					amplicon_emit_pass<EXTRACT>(sd, seq, s_ent, g_pl, g_loc, g_strand, e0, Ep, E, Fq, Rq, amp_min, amp_max, lane, q, 0u, out);
				if ((todo2 >> src) & 1u)
					amplicon_emit_pass<EXTRACT>(sd, seq, s_ent, g_pl, g_loc, g_strand, e0, Ep, E, Rq, Fq, amp_min, amp_max, lane, q, 1u, out);
			}
		}
	}
}

__global__ void iota32_kernel(uint32_t *p, uint64_t n)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = (uint32_t)i;
}
__global__ void gather64_kernel(const uint64_t *__restrict__ src, const uint32_t *__restrict__ perm, uint64_t *dst, uint64_t n)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dst[i] = src[perm[i]];
}

// Candidate amplicons of every pair in the reference's order -- pass {F+,R-} then {R+,F-}, each by (sequence, plus loc, minus loc):
// the list kernel (re-run once if the first capacity guess was short; the list is deterministic) and two stable radix sorts
// carrying a permutation.  After the call: k1[0] = sorted key1, perm[0][i] = record of sorted position i; k2[0], pid, mid are
// indexed by record.
struct AmpList {
	DevBuf k1[2], k2[2], pid, mid, perm[2], tmp, cnt;
	uint64_t n = 0;
};

template <bool EXTRACT>
int build_amplicon_list(pcramp_gpu_ctx *ctx, SeqSet &s, const OligoDev *d_ol, uint32_t n_pairs, int amp_min, int amp_max, AmpList &L, const char *who)
{
	cudaStream_t st = ctx->stream;
	CK(L.cnt.ensure(8));
	uint64_t cap = std::max<uint64_t>(1u << 16, (uint64_t)n_pairs * 64), n = 0;
	const unsigned grid = (unsigned)std::min<uint64_t>(s.n, (uint64_t)ctx->sm_count * 8);
	for (int attempt = 0; attempt < 2; ++attempt) {
		CK(L.k1[0].ensure(cap * 8));
		CK(L.k2[0].ensure(cap * 8));
		CK(L.pid.ensure(cap * 4));
		CK(L.mid.ensure(cap * 4));
		CK(cudaMemsetAsync(L.cnt.p, 0, 8, st));
		AmpliconOut out;
		out.key1 = L.k1[0].as<uint64_t>();
		out.key2 = L.k2[0].as<uint64_t>();
		out.plus_id = L.pid.as<uint32_t>();
		out.minus_id = L.mid.as<uint32_t>();
		out.count = L.cnt.as<unsigned long long>();
		out.cap = cap;
		amplicon_list_kernel<EXTRACT><<<grid, SCORE_THREADS, 0, st>>>(s.dev(), s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(),
			s.seq_ent_off.as<uint32_t>(), d_ol, n_pairs, amp_min, amp_max, out);
		CK(cudaGetLastError());
		ctx->stats.kernel_launches++;
		unsigned long long h_n = 0;
		CK(cudaMemcpyAsync(&h_n, L.cnt.p, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		n = h_n;
		if (n <= cap) break;
		if (attempt == 1) return fail(ctx, std::string(who) + ": candidate list kept growing");
		cap = n; // the list is deterministic: the second run fits exactly
	}
	L.n = n;
	if (n >= (1ull << 31)) return fail(ctx, std::string(who) + ": more than 2^31 candidate amplicons in one batch");
	if (!n) return 0;
	CK(L.k1[1].ensure(n * 8));
	CK(L.k2[1].ensure(n * 8));
	CK(L.perm[0].ensure(n * 4));
	CK(L.perm[1].ensure(n * 4));
	iota32_kernel<<<grid_for(n, 256), 256, 0, st>>>(L.perm[0].as<uint32_t>(), n);
	size_t tmp_bytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, L.k2[0].as<uint64_t>(), L.k2[1].as<uint64_t>(), L.perm[0].as<uint32_t>(), L.perm[1].as<uint32_t>(),
		(int)n, 0, 64, st);
	CK(L.tmp.ensure(tmp_bytes));
	CK(cub::DeviceRadixSort::SortPairs(L.tmp.p, tmp_bytes, L.k2[0].as<uint64_t>(), L.k2[1].as<uint64_t>(), L.perm[0].as<uint32_t>(),
		L.perm[1].as<uint32_t>(), (int)n, 0, 64, st));
	gather64_kernel<<<grid_for(n, 256), 256, 0, st>>>(L.k1[0].as<uint64_t>(), L.perm[1].as<uint32_t>(), L.k1[1].as<uint64_t>(), n);
	CK(cub::DeviceRadixSort::SortPairs(L.tmp.p, tmp_bytes, L.k1[1].as<uint64_t>(), L.k1[0].as<uint64_t>(), L.perm[1].as<uint32_t>(),
		L.perm[0].as<uint32_t>(), (int)n, 0, 64, st));
	CK(cudaGetLastError());
	ctx->stats.kernel_launches += 4;
	return 0;
}

// find_background_match's scoring loop (background_match.cpp:66-165) for the amplicon at sorted position i.
// The reference walks the list two at a time and guards the second of each pair with `(i + 1) >= num_seq`
// (:122): an odd-indexed candidate is scored only while its index is below the number of
// SEQUENCES.  Reproduced as written.  (Its other effect -- reading one candidate past the end of an odd-length
// list -- is undefined behaviour in the reference and has no counterpart here.)
__global__ void __launch_bounds__(128) background_sw_kernel(uint64_t n, const uint64_t *__restrict__ key1_sorted, const uint32_t *__restrict__ perm,
	const uint32_t *__restrict__ plus_id, const uint32_t *__restrict__ minus_id, const uint64_t *__restrict__ e_hi, const uint64_t *__restrict__ e_lo,
	const uint64_t *__restrict__ f, const uint64_t *__restrict__ r, uint32_t n_seq, float threshold, int taq, uint32_t *bits, uint32_t n_words)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t k1 = key1_sorted[i];
	const uint32_t pair = (uint32_t)(k1 >> 33), pass = (uint32_t)(k1 >> 32) & 1u, seq = (uint32_t)k1;
	const uint64_t first_key = (uint64_t)pair << 33;
	uint64_t lo = 0, hi = i; // first record of this pair
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		if (key1_sorted[mid] < first_key) lo = mid + 1; else hi = mid;
	}
	const uint64_t idx = i - lo;
	if ((idx & 1ull) && idx >= (uint64_t)n_seq) return;
	const uint32_t rec = perm[i];
	const uint32_t pe = plus_id[rec], me = minus_id[rec];
	// pass 0: F on the plus entry, R on the minus entry; pass 1: R on plus, F on minus (pcr_assay.cpp:421-435)
	const uint32_t fe = pass ? me : pe, re = pass ? pe : me;
	W128 fw, rw, fk, rk;
	fw.hi = f[2 * pair]; fw.lo = f[2 * pair + 1];
	rw.hi = r[2 * pair]; rw.lo = r[2 * pair + 1];
	fk.hi = e_hi[fe]; fk.lo = e_lo[fe];
	rk.hi = e_hi[re]; rk.lo = e_lo[re];
	const W128 fc = w_complement(fw), rc = w_complement(rw);
	const sw::WordTarget tf(fk), tr(rk);
	sw::Query q;
	sw::query_from_word(fw, q);
	const sw::Result s0 = sw::align_warp<false>(q, tf); // slot 0: F   + f
	sw::query_from_word(fc, q);
	const sw::Result s1 = sw::align_warp<false>(q, tf); // slot 1: (F) + f
	sw::query_from_word(rw, q);
	const sw::Result s2 = sw::align_warp<false>(q, tr); // slot 2: R   + r
	sw::query_from_word(rc, q);
	const sw::Result s3 = sw::align_warp<false>(q, tr); // slot 3: (R) + r
	float f_norm = __fmul_rn(2.0f, (float)w_size(fw)), r_norm = __fmul_rn(2.0f, (float)w_size(rw));
	if (f_norm > 0.0f) f_norm = __fdiv_rn(1.0f, f_norm);
	if (r_norm > 0.0f) r_norm = __fdiv_rn(1.0f, r_norm);
	float FpRm = __fmul_rn(__fmul_rn((float)(s0.score * s3.score), f_norm), r_norm);
	float RpFm = __fmul_rn(__fmul_rn((float)(s1.score * s2.score), f_norm), r_norm);
	if (taq) {
		unsigned p0, p1, t0, t1, u0, u1, v0, v1;
		word_last_two(fw, p0, p1); // Fp
		sw::last_two(s0, tf, t0, t1);
		word_last_two(rc, u0, u1); // Rm
		sw::last_two(s3, tr, v0, v1);
		FpRm = __fmul_rn(FpRm, __fmul_rn(taq_correction(p0, p1, t0, t1), taq_correction(u0, u1, v0, v1)));
		word_last_two(rw, p0, p1); // Rp
		sw::last_two(s2, tr, t0, t1);
		word_last_two(fc, u0, u1); // Fm
		sw::last_two(s1, tf, v0, v1);
		RpFm = __fmul_rn(RpFm, __fmul_rn(taq_correction(p0, p1, t0, t1), taq_correction(u0, u1, v0, v1)));
	}
	const float score = (FpRm > RpFm) ? __fsqrt_rn(FpRm) : __fsqrt_rn(RpFm);
	if (score >= threshold) atomicOr(bits + (size_t)pair * n_words + (seq >> 5), 1u << (seq & 31u));
}

// find_multiplex_background_match (background_match.cpp:168-295): thread = (pair, sequence, slot); slots 0..3 =
// F, rc(F), R, rc(R) against the whole sequence.  Consecutive threads share the sequence, so the loop length is
// uniform in a warp and the target nibbles are broadcast loads.
__global__ void __launch_bounds__(128) multiplex_sw_kernel(SeqDev sd, const uint64_t *__restrict__ f, const uint64_t *__restrict__ r, uint32_t n_pairs,
	float threshold, int taq, uint32_t *bits, uint32_t n_words)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t total = (uint64_t)n_pairs * 4ull * sd.n;
	if (tid >= total) return;
	const uint32_t pair = (uint32_t)(tid % n_pairs);
	const uint64_t rest = tid / n_pairs;
	const uint32_t slot = (uint32_t)(rest & 3ull), seq = (uint32_t)(rest >> 2);
	W128 w;
	const uint64_t *src = (slot & 2u) ? r : f;
	w.hi = src[2 * pair]; w.lo = src[2 * pair + 1];
	const int size = w_size(w);
	if (slot & 1u) w = w_complement(w);
	sw::Query q;
	sw::query_from_word(w, q);
	const sw::NibbleTarget t(sd.raw + sd.raw_off[seq], (int)sd.len[seq]);
	const sw::Result s = sw::align_warp<false>(q, t);
	float norm = __fmul_rn(2.0f, (float)size);
	if (norm > 0.0f) norm = __fdiv_rn(1.0f, norm);
	float score = __fmul_rn((float)s.score, norm);
	if (taq) {
		unsigned p0, p1, t0, t1;
		word_last_two(w, p0, p1);
		sw::last_two(s, t, t0, t1);
		score = __fmul_rn(score, taq_correction(p0, p1, t0, t1));
	}
	if (score >= threshold) atomicOr(bits + (size_t)pair * n_words + (seq >> 5), 1u << (seq & 31u));
}


// ---- find_background_match by units ----------------------------------------------------------------------------------
// The record list above holds one entry per candidate amplicon and background_sw_kernel aligns four times per record.  At the
// background thresholds a pair has ~10^5 candidate amplicons on 10^3 sequences, but an amplicon is a (plus entry, minus entry)
// combination and its four alignments are two per entry: F and rc(F) against the word the forward primer matched, R and rc(R)
// against the word the reverse primer matched (background_match.cpp:66-118).  So the alignments are done once per
// (pair, matching entry) and the amplicons are never written down:
//   unit = (sequence, pair); its four lists = the plus / minus entries its F / R match (PF, PR, MF, MR), by loc
//   bg_match_kernel<false / true>   sizes of the lists, then the entry ids (CTA per sequence, lane = pair, as amplicon_list_kernel)
//   bg_sw_kernel                    one thread per list element: the two alignments of its oligo against the entry's word
//   bg_amp_kernel<false>            valid (plus, minus) combinations per (pair, pass, sequence): find_amplicon_match's geometry
//   prefix sum per pair             = the index each unit's first amplicon has in the reference's list of the pair
//                                   (pass {F+,R-} then {R+,F-}, each by (sequence, plus loc, minus loc))
//   bg_amp_kernel<true>             walks the combinations in that order, applies the odd-index guard of :122 and the score
struct BgLists {
	const uint32_t *off4;    // exclusive prefix of the list sizes: list (unit, l) = [off4[4 unit + l], off4[4 unit + l + 1])
	const uint32_t *entry;   // database entry ids, each list by (loc, id)
};

template <bool FILL>
__global__ void __launch_bounds__(SCORE_THREADS)
bg_match_kernel(SeqDev sd, const uint4 *__restrict__ g_pl, const int32_t *__restrict__ g_loc, const uint32_t *__restrict__ g_strand,
	const uint32_t *__restrict__ seq_off2, const OligoDev *__restrict__ oligos, uint32_t n_pairs, uint32_t *cnt4, const uint32_t *__restrict__ off4,
	uint32_t *m_entry)
{
	__shared__ ScoreEntry s_ent[SCORE_SMEM_ENTRIES];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = SCORE_THREADS / 32;
	const uint32_t n_chunks = (n_pairs + 31u) / 32u;
	for (uint32_t seq = blockIdx.x; seq < sd.n; seq += gridDim.x) {
		const uint32_t e0 = seq_off2[2 * seq], Ep = seq_off2[2 * seq + 1] - e0, E = seq_off2[2 * seq + 2] - e0;
		if (Ep == 0u || Ep == E || !sd.active[seq]) continue;
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
			if (p >= n_pairs) continue;
			const OligoDev F = oligos[2 * p], R = oligos[2 * p + 1];
			const int f_thr = (int)(F.packed & 255u), r_thr = (int)(R.packed & 255u);
			const size_t u4 = ((size_t)seq * n_pairs + p) * 4u;
			uint32_t at[4] = {0u, 0u, 0u, 0u}, want[4] = {0u, 0u, 0u, 0u};
			if (FILL) {
#pragma unroll
				for (int l = 0; l < 4; ++l) {
					at[l] = off4[u4 + l];
					want[l] = off4[u4 + l + 1] - at[l];
				}
				if (!(want[0] | want[1] | want[2] | want[3])) continue;
			}
			uint32_t n[4] = {0u, 0u, 0u, 0u};
			for (uint32_t e = 0; e < E; ++e) {
				const ScoreEntry en = load_entry(s_ent, g_pl, g_loc, g_strand, e0, e);
				const int base = e < Ep ? 0 : 2;
				if (oligo_count(F, en) >= f_thr) {
					if (FILL && want[base]) m_entry[at[base] + n[base]] = e0 + e;
					++n[base];
				}
				if (oligo_count(R, en) >= r_thr) {
					if (FILL && want[base + 1]) m_entry[at[base + 1] + n[base + 1]] = e0 + e;
					++n[base + 1];
				}
			}
			if (!FILL) { // pass {F+, R-} needs PF and MR, pass {R+, F-} needs PR and MF: a list without its partner is dropped
				const bool p0 = n[0] && n[3], p1 = n[1] && n[2];
				cnt4[u4 + 0] = p0 ? n[0] : 0u;
				cnt4[u4 + 3] = p0 ? n[3] : 0u;
				cnt4[u4 + 1] = p1 ? n[1] : 0u;
				cnt4[u4 + 2] = p1 ? n[2] : 0u;
			} else { // by (loc, id): the full-window entries of a run already are, the few partial words at its end are moved into place
#pragma unroll
				for (int l = 0; l < 4; ++l) {
					uint32_t *L = m_entry + at[l];
					for (uint32_t i = 1; i < want[l]; ++i) {
						const uint32_t x = L[i];
						const int xl = g_loc[x];
						uint32_t j = i;
						while (j > 0 && g_loc[L[j - 1]] > xl) {
							L[j] = L[j - 1];
							--j;
						}
						L[j] = x;
					}
				}
			}
		}
	}
}

// one thread per list element: {score(oligo, word), score(rc(oligo), word)} and, for TaqMAMA, the two aligned target bases of each.
// Without TaqMAMA only the scores are needed: the two alignments run as the 16-bit halves of one pass (sw.cuh).
template <bool TAQ>
__global__ void __launch_bounds__(128) bg_sw_kernel(uint32_t n_match, uint32_t n_lists, BgLists B, const uint64_t *__restrict__ e_hi,
	const uint64_t *__restrict__ e_lo, const uint64_t *__restrict__ f, const uint64_t *__restrict__ r, uint32_t n_pairs, uint2 *res)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = i < n_match;
	W128 ow, tw;
	ow.hi = ow.lo = tw.hi = tw.lo = 0ull;
	if (live) {
		uint32_t lo = 0, hi = n_lists; // the list that holds element i: the last one that starts at or before it
		while (lo < hi) {
			const uint32_t mid = (lo + hi) >> 1;
			if (__ldg(B.off4 + mid + 1) <= i) lo = mid + 1; else hi = mid;
		}
		const uint32_t pair = (lo >> 2) % n_pairs, l = lo & 3u;
		const uint64_t *src = (l == 0u || l == 2u) ? f : r; // PF, MF: the forward primer's words
		ow.hi = src[2 * pair]; ow.lo = src[2 * pair + 1];
		const uint32_t e = B.entry[i];
		tw.hi = e_hi[e]; tw.lo = e_lo[e];
	}
	const sw::WordTarget t(tw);
	sw::Query q, qc;
	sw::query_from_word(ow, q);
	sw::query_from_word(w_complement(ow), qc);
	if (!TAQ) {
		int sa, sb;
		sw::align_score_pair_warp(q, qc, t, sa, sb);
		if (live) res[i] = make_uint2((uint32_t)(uint16_t)(int16_t)sa | ((uint32_t)(uint16_t)(int16_t)sb << 16), 0u);
		return;
	}
	const sw::Result a = sw::align_warp<false>(q, t);
	const sw::Result b = sw::align_warp<false>(qc, t);
	if (!live) return;
	unsigned a0, a1, b0, b1;
	sw::last_two(a, t, a0, a1);
	sw::last_two(b, t, b0, b1);
	res[i] = make_uint2((uint32_t)(uint16_t)(int16_t)a.score | ((uint32_t)(uint16_t)(int16_t)b.score << 16), (a0 << 4) | a1 | (b0 << 12) | (b1 << 8));
}

template <bool SCORE>
__global__ void __launch_bounds__(128) bg_amp_kernel(SeqDev sd, BgLists B, const int32_t *__restrict__ g_loc, const OligoDev *__restrict__ oligos,
	uint32_t n_pairs, int amp_min, int amp_max, unsigned long long *cnt2, const unsigned long long *__restrict__ off2, const uint2 *__restrict__ res,
	const uint64_t *__restrict__ f, const uint64_t *__restrict__ r, float threshold, int taq, uint32_t *bits, uint32_t n_words)
{
	const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= (uint64_t)sd.n * n_pairs) return;
	const uint32_t seq = (uint32_t)(u / n_pairs), pair = (uint32_t)(u % n_pairs);
	uint32_t o[5];
#pragma unroll
	for (int l = 0; l < 5; ++l) o[l] = B.off4[4 * u + l];
	if (o[4] == o[0]) return;
	const int L = (int)sd.len[seq];
	const OligoDev F = oligos[2 * pair], R = oligos[2 * pair + 1];
	float f_norm = 0.0f, r_norm = 0.0f;
	W128 fw, rw, fc, rc;
	fw.hi = fw.lo = rw.hi = rw.lo = 0ull;
	fc = fw; rc = fw;
	if (SCORE) {
		fw.hi = f[2 * pair]; fw.lo = f[2 * pair + 1];
		rw.hi = r[2 * pair]; rw.lo = r[2 * pair + 1];
		fc = w_complement(fw); rc = w_complement(rw);
		f_norm = __fmul_rn(2.0f, (float)w_size(fw)); r_norm = __fmul_rn(2.0f, (float)w_size(rw));
		if (f_norm > 0.0f) f_norm = __fdiv_rn(1.0f, f_norm);
		if (r_norm > 0.0f) r_norm = __fdiv_rn(1.0f, r_norm);
	}
	bool found = false;
	for (uint32_t pass = 0; pass < 2u && !found; ++pass) {
		// pass 0: F on the plus entry (PF), R on the minus entry (MR); pass 1: R on plus (PR), F on minus (MF) (pcr_assay.cpp:421-435)
		const uint32_t p0 = pass ? o[1] : o[0], p1 = pass ? o[2] : o[1], m0 = pass ? o[2] : o[3], m1 = pass ? o[3] : o[4];
		if (p0 == p1 || m0 == m1) continue;
		const OligoDev &P = pass ? R : F, &M = pass ? F : R;
		const int p_start = (int)((P.packed >> 8) & 255u), p_stop = (int)((P.packed >> 16) & 255u);
		const int m_start = (int)((M.packed >> 8) & 255u), m_stop = (int)((M.packed >> 16) & 255u);
		const size_t slot = ((size_t)pair * 2u + pass) * sd.n + seq;
		unsigned long long rank = 0;
		if (SCORE) rank = off2[slot] - off2[(size_t)pair * 2u * sd.n]; // position of this unit's first amplicon in the pair's list
		unsigned long long count = 0;
		for (uint32_t a = p0; a < p1 && !found; ++a) {
			const int ploc = g_loc[B.entry[a]];
			const int plus_loc3 = ploc + p_stop;
			for (uint32_t b = m0; b < m1; ++b) {
				const int mloc = g_loc[B.entry[b]];
				if (!(plus_loc3 < mloc - m_stop)) continue; // pcr_assay.cpp:368-371
				int amp_start = ploc + p_start;
				const int amp_stop = min(mloc - m_start, L - 1);
				int amp_len = amp_stop - amp_start + 1;
				if (amp_len < amp_min || amp_len > amp_max) continue; // :383-392
				if (amp_start < 0) { amp_len += amp_start; amp_start = 0; }
				if (amp_len < 0 || has_split_dev(sd, seq, amp_start, amp_len)) continue; // :418
				if (!SCORE) { ++count; continue; }
				const unsigned long long idx = rank++;
				if ((idx & 1ull) && idx >= (unsigned long long)sd.n) continue; // background_match.cpp:122, as written
				const uint2 vf = res[pass ? b : a], vr = res[pass ? a : b]; // the entry the forward / the reverse primer matched
				const int s0 = (int)(int16_t)(vf.x & 0xffffu), s1 = (int)(int16_t)(vf.x >> 16); // F, rc(F) against the f-key
				const int s2 = (int)(int16_t)(vr.x & 0xffffu), s3 = (int)(int16_t)(vr.x >> 16); // R, rc(R) against the r-key
				float FpRm = __fmul_rn(__fmul_rn((float)(s0 * s3), f_norm), r_norm);
				float RpFm = __fmul_rn(__fmul_rn((float)(s1 * s2), f_norm), r_norm);
				if (taq) {
					unsigned q0, q1, u0, u1;
					word_last_two(fw, q0, q1); // Fp against slot 0's target bases
					word_last_two(rc, u0, u1); // Rm against slot 3's
					FpRm = __fmul_rn(FpRm, __fmul_rn(taq_correction(q0, q1, (vf.y >> 4) & 15u, vf.y & 15u), taq_correction(u0, u1, (vr.y >> 12) & 15u, (vr.y >> 8) & 15u)));
					word_last_two(rw, q0, q1); // Rp against slot 2's
					word_last_two(fc, u0, u1); // Fm against slot 1's
					RpFm = __fmul_rn(RpFm, __fmul_rn(taq_correction(q0, q1, (vr.y >> 4) & 15u, vr.y & 15u), taq_correction(u0, u1, (vf.y >> 12) & 15u, (vf.y >> 8) & 15u)));
				}
				const float score = (FpRm > RpFm) ? __fsqrt_rn(FpRm) : __fsqrt_rn(RpFm);
				if (score >= threshold) { found = true; break; }
			}
		}
		if (!SCORE) cnt2[slot] = count;
	}
	if (SCORE && found) atomicOr(bits + (size_t)pair * n_words + (seq >> 5), 1u << (seq & 31u));
}

// 64-bit total of the list sizes: the prefix sum runs in 32 bits and must not wrap unnoticed
__global__ void bg_total_kernel(const uint32_t *__restrict__ cnt, uint32_t n, unsigned long long *total)
{
	unsigned long long v = 0;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v += cnt[i];
	for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
	if ((threadIdx.x & 31u) == 0u && v) atomicAdd(total, v);
}

int background_match_units(pcramp_gpu_ctx *ctx, SeqSet &s, const uint64_t *d_f, const uint64_t *d_r, const OligoDev *d_ol, uint32_t n_pairs,
	float detect_threshold, int amp_min, int amp_max, int taq, uint32_t *d_bits, uint32_t n_words, uint64_t *n_amplicons)
{
	cudaStream_t st = ctx->stream;
	Trace tr("background_match", st);
	const uint64_t U = (uint64_t)s.n * n_pairs;
	if (U * 4 + 1 >= (1ull << 32)) return fail(ctx, "pcramp_gpu_background_match: too many (sequence, pair) units in one batch");
	const uint32_t n_lists = (uint32_t)(U * 4);
	DevBuf &d_cnt4 = ctx->bg_cnt4, &d_off4 = ctx->bg_off4, &d_entry = ctx->bg_entry, &d_res = ctx->bg_res, &d_cnt2 = ctx->bg_cnt2, &d_off2 = ctx->bg_off2;
	CK(d_cnt4.ensure(((size_t)n_lists + 1) * 4));
	CK(d_off4.ensure(((size_t)n_lists + 1) * 4));
	CK(cudaMemsetAsync(d_cnt4.p, 0, ((size_t)n_lists + 1) * 4, st));
	const unsigned grid = (unsigned)std::min<uint64_t>(s.n, (uint64_t)ctx->sm_count * 8);
	bg_match_kernel<false><<<grid, SCORE_THREADS, 0, st>>>(s.dev(), s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(),
		s.seq_ent_off.as<uint32_t>(), d_ol, n_pairs, d_cnt4.as<uint32_t>(), nullptr, nullptr);
	CK(cudaGetLastError());
	size_t tb = 0, tb2 = 0;
	const size_t n2 = (size_t)n_pairs * 2 * s.n + 1;
	CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_cnt4.as<uint32_t>(), d_off4.as<uint32_t>(), (int)(n_lists + 1), st));
	CK(cub::DeviceScan::ExclusiveSum(nullptr, tb2, d_cnt2.as<unsigned long long>(), d_off2.as<unsigned long long>(), (int)n2, st));
	CK(ctx->cub_tmp.ensure(std::max(tb, tb2)));
	tb = tb2 = ctx->cub_tmp.cap;
	CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tb, d_cnt4.as<uint32_t>(), d_off4.as<uint32_t>(), (int)(n_lists + 1), st));
	CK(ctx->d_item_count.ensure(16));
	CK(cudaMemsetAsync(ctx->d_item_count.p, 0, 8, st));
	bg_total_kernel<<<(unsigned)ctx->sm_count * 4u, 256, 0, st>>>(d_cnt4.as<uint32_t>(), n_lists, ctx->d_item_count.as<unsigned long long>());
	uint32_t n_match = 0;
	unsigned long long n_match64 = 0;
	CK(cudaMemcpyAsync(&n_match, d_off4.as<uint32_t>() + n_lists, 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(&n_match64, ctx->d_item_count.p, 8, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	ctx->stats.kernel_launches += 4;
	if (n_match64 >= (1ull << 32)) return fail(ctx, "pcramp_gpu_background_match: more than 2^32 (pair, matching entry) combinations in one batch (use fewer pairs per call)");
	tr.mark("list sizes");
	if (n_amplicons) *n_amplicons = 0;
	if (!n_match) return 0;
	CK(d_entry.ensure((size_t)n_match * 4));
	CK(d_res.ensure((size_t)n_match * 8));
	CK(d_cnt2.ensure(n2 * 8));
	CK(d_off2.ensure(n2 * 8));
	BgLists B;
	B.off4 = d_off4.as<uint32_t>();
	B.entry = d_entry.as<uint32_t>();
	bg_match_kernel<true><<<grid, SCORE_THREADS, 0, st>>>(s.dev(), s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(),
		s.seq_ent_off.as<uint32_t>(), d_ol, n_pairs, nullptr, d_off4.as<uint32_t>(), d_entry.as<uint32_t>());
	tr.mark("lists");
	if (taq)
		bg_sw_kernel<true><<<grid_for(n_match, 128), 128, 0, st>>>(n_match, n_lists, B, s.e_hi.as<uint64_t>(), s.e_lo.as<uint64_t>(), d_f, d_r, n_pairs, d_res.as<uint2>());
	else
		bg_sw_kernel<false><<<grid_for(n_match, 128), 128, 0, st>>>(n_match, n_lists, B, s.e_hi.as<uint64_t>(), s.e_lo.as<uint64_t>(), d_f, d_r, n_pairs, d_res.as<uint2>());
	tr.mark("alignments");
	CK(cudaMemsetAsync(d_cnt2.p, 0, n2 * 8, st));
	bg_amp_kernel<false><<<grid_for(U, 128), 128, 0, st>>>(s.dev(), B, s.e_loc.as<int32_t>(), d_ol, n_pairs, amp_min, amp_max, d_cnt2.as<unsigned long long>(),
		nullptr, nullptr, nullptr, nullptr, 0.0f, 0, nullptr, 0u);
	CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tb2, d_cnt2.as<unsigned long long>(), d_off2.as<unsigned long long>(), (int)n2, st));
	tr.mark("amplicon counts");
	bg_amp_kernel<true><<<grid_for(U, 128), 128, 0, st>>>(s.dev(), B, s.e_loc.as<int32_t>(), d_ol, n_pairs, amp_min, amp_max, nullptr,
		d_off2.as<unsigned long long>(), d_res.as<uint2>(), d_f, d_r, detect_threshold, taq, d_bits, n_words);
	CK(cudaGetLastError());
	ctx->stats.kernel_launches += 6;
	tr.mark("amplicon scores");
	if (tr.on) fprintf(stderr, "[trace] background_match: %u list elements (2 alignments each), %llu units\n", n_match, (unsigned long long)U);
	if (n_amplicons) {
		unsigned long long total = 0;
		CK(cudaMemcpyAsync(&total, d_off2.as<unsigned long long>() + (n2 - 1), 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		*n_amplicons = total;
	}
	return 0;
}

// Pair scoring by units, for thresholds at which neither the neighbour bound nor a seed table can exclude anybody (the background
// thresholds: search 0.72, oligos matched at 0.52).  The lists of find_background_match above are exactly collect_candidates'
// membership -- the plus / minus entries of a sequence that F / R match at thr^2 -- so find_amplicon_match (pcr_assay.cpp:338-441) is
// one thread per (sequence, pair) walking PF x MR and PR x MF: geometry, split test, the two identities, sqrtf(f * r) >= detect.
// `member` decides the lists (the unmoved assay in variant mode), `oligos` gives the identities.
__global__ void __launch_bounds__(128) unit_score_kernel(SeqDev sd, BgLists B, const uint4 *__restrict__ e_planes, const int32_t *__restrict__ e_loc,
	const OligoDev *__restrict__ member, const OligoDev *__restrict__ oligos, uint32_t n_pairs, float detect, int amp_min, int amp_max, int taq,
	uint32_t *bits_any, uint32_t *bits_pass1, uint32_t n_words)
{
	const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= (uint64_t)sd.n * n_pairs) return;
	const uint32_t seq = (uint32_t)(u / n_pairs), pair = (uint32_t)(u % n_pairs);
	uint32_t o[5];
#pragma unroll
	for (int l = 0; l < 5; ++l) o[l] = B.off4[4 * u + l];
	if (o[4] == o[0]) return;
	const int L = (int)sd.len[seq];
	bool any = false, first = false;
	for (uint32_t pass = 0; pass < 2u; ++pass) {
		// pass 0: F on the plus entry (PF), R on the minus entry (MR); pass 1: R on plus (PR), F on minus (MF) (pcr_assay.cpp:37-59)
		const uint32_t p0 = pass ? o[1] : o[0], p1 = pass ? o[2] : o[1], m0 = pass ? o[2] : o[3], m1 = pass ? o[3] : o[4];
		if (p0 == p1 || m0 == m1) continue;
		const OligoDev Pb = member[2 * pair + pass], Mb = member[2 * pair + (pass ^ 1u)];
		const OligoDev P = oligos[2 * pair + pass], M = oligos[2 * pair + (pass ^ 1u)];
		const int p_start = (int)((Pb.packed >> 8) & 255u), p_stop = (int)((Pb.packed >> 16) & 255u);
		const int m_start = (int)((Mb.packed >> 8) & 255u), m_stop = (int)((Mb.packed >> 16) & 255u);
		bool found = false;
		for (uint32_t a = p0; a < p1 && !found; ++a) {
			const uint32_t ea = B.entry[a];
			ScoreEntry pe;
			const uint4 w = __ldg(e_planes + ea);
			pe.a = w.x; pe.c = w.y; pe.g = w.z; pe.t = w.w;
			pe.loc = __ldg(e_loc + ea);
			pe.strand = STRAND_PLUS;
			const float ident_p = oligo_identity(P, oligo_count(P, pe), pe, taq);
			const int plus_loc3 = pe.loc + p_stop;
			for (uint32_t b = m0; b < m1; ++b) {
				const uint32_t eb = B.entry[b];
				const int mloc = __ldg(e_loc + eb);
				if (!(plus_loc3 < mloc - m_stop)) continue; // pcr_assay.cpp:368-371
				int amp_start = pe.loc + p_start;
				const int amp_stop = min(mloc - m_start, L - 1);
				int amp_len = amp_stop - amp_start + 1;
				if (amp_len < amp_min || amp_len > amp_max) continue; // :383-392
				if (amp_start < 0) { amp_len += amp_start; amp_start = 0; } // :412-416
				if (amp_len < 0 || has_split_dev(sd, seq, amp_start, amp_len)) continue; // :418
				ScoreEntry m2;
				const uint4 v = __ldg(e_planes + eb);
				m2.a = v.x; m2.c = v.y; m2.g = v.z; m2.t = v.w;
				m2.loc = mloc;
				m2.strand = STRAND_MINUS;
				const float ident_m = oligo_identity(M, oligo_count(M, m2), m2, taq);
				if (__fsqrt_rn(__fmul_rn(ident_p, ident_m)) >= detect) { found = true; break; } // :292-294
			}
		}
		if (found) {
			any = true;
			if (pass == 0u) first = true;
		}
	}
	if (any) {
		const uint32_t bit = 1u << (seq & 31u);
		atomicOr(bits_any + (size_t)pair * n_words + (seq >> 5), bit);
		if (first) atomicOr(bits_pass1 + (size_t)pair * n_words + (seq >> 5), bit); // {F(+), R(-)}: pass 1
	}
}

} // namespace

// -> 0 done (bits_any / bits_pass1 hold the result), 1 error, 2 not applicable
static int score_by_units(pcramp_gpu_ctx *ctx, SeqSet &s, const OligoDev *d_member, const OligoDev *d_oligos, uint32_t n_pairs, float detect_threshold,
	int amp_min, int amp_max, int taq, uint32_t *d_bits_any, uint32_t *d_bits_pass1, uint32_t n_words)
{
	cudaStream_t st = ctx->stream;
	const uint64_t U = (uint64_t)s.n * n_pairs;
	if (!ctx->use_unit_score || U * 4 + 1 >= (1ull << 32)) return 2;
	const uint32_t n_lists = (uint32_t)(U * 4);
	DevBuf &d_cnt4 = ctx->bg_cnt4, &d_off4 = ctx->bg_off4, &d_entry = ctx->bg_entry;
	CK(d_cnt4.ensure(((size_t)n_lists + 1) * 4));
	CK(d_off4.ensure(((size_t)n_lists + 1) * 4));
	CK(cudaMemsetAsync(d_cnt4.p, 0, ((size_t)n_lists + 1) * 4, st));
	const unsigned grid = (unsigned)std::min<uint64_t>(s.n, (uint64_t)ctx->sm_count * 8);
	bg_match_kernel<false><<<grid, SCORE_THREADS, 0, st>>>(s.dev(), s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(),
		s.seq_ent_off.as<uint32_t>(), d_member, n_pairs, d_cnt4.as<uint32_t>(), nullptr, nullptr);
	CK(cudaGetLastError());
	size_t tb = 0;
	CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_cnt4.as<uint32_t>(), d_off4.as<uint32_t>(), (int)(n_lists + 1), st));
	CK(ctx->cub_tmp.ensure(tb));
	tb = ctx->cub_tmp.cap;
	CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tb, d_cnt4.as<uint32_t>(), d_off4.as<uint32_t>(), (int)(n_lists + 1), st));
	CK(ctx->d_item_count.ensure(16));
	CK(cudaMemsetAsync(ctx->d_item_count.p, 0, 8, st));
	bg_total_kernel<<<(unsigned)ctx->sm_count * 4u, 256, 0, st>>>(d_cnt4.as<uint32_t>(), n_lists, ctx->d_item_count.as<unsigned long long>());
	uint32_t n_match = 0;
	unsigned long long n_match64 = 0;
	CK(cudaMemcpyAsync(&n_match, d_off4.as<uint32_t>() + n_lists, 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(&n_match64, ctx->d_item_count.p, 8, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	ctx->stats.kernel_launches += 4;
	if (n_match64 >= (1ull << 32)) return 2; // more list elements than 32-bit offsets hold: the general path
	if (!n_match) return 0;
	CK(d_entry.ensure((size_t)n_match * 4));
	BgLists B;
	B.off4 = d_off4.as<uint32_t>();
	B.entry = d_entry.as<uint32_t>();
	bg_match_kernel<true><<<grid, SCORE_THREADS, 0, st>>>(s.dev(), s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(), s.e_strand.as<uint32_t>(),
		s.seq_ent_off.as<uint32_t>(), d_member, n_pairs, nullptr, d_off4.as<uint32_t>(), d_entry.as<uint32_t>());
	unit_score_kernel<<<grid_for(U, 128), 128, 0, st>>>(s.dev(), B, s.e_planes.as<uint4>(), s.e_loc.as<int32_t>(), d_member, d_oligos, n_pairs,
		detect_threshold, amp_min, amp_max, taq, d_bits_any, d_bits_pass1, n_words);
	CK(cudaGetLastError());
	ctx->stats.kernel_launches += 2;
	return 0;
}

namespace {
} // namespace

extern "C" {

int pcramp_gpu_sw_batch(pcramp_gpu_ctx *ctx, uint32_t n, const uint64_t *query, const uint64_t *target, int32_t *score, int32_t *q_start,
	int32_t *q_stop, int32_t *t_start, int32_t *t_stop, uint8_t *last_two)
{
	if (!ctx) return 1;
	if (n && (!query || !target)) return fail(ctx, "pcramp_gpu_sw_batch: null argument");
	CK(cudaSetDevice(ctx->device));
	if (!n) return 0;
	for (uint32_t p = 0; p < n; ++p) { // pack_query_slots throws on an empty query (seq_overlap.h:832-834)
		if ((query[2 * p] | query[2 * p + 1]) == 0) return fail(ctx, ":SeqOverlap::pack_query_slots: len == 0");
	}
	DevBuf &dq = ctx->sw_q, &dt = ctx->sw_t, &dout = ctx->sw_out;
	CK(dq.ensure((size_t)n * 16));
	CK(dt.ensure((size_t)n * 16));
	CK(dout.ensure((size_t)n * 22));
	// the caller's arrays are the staging: page-locked arrays move at link speed, pageable ones through the driver's bounce buffers
	CK(cudaMemcpyAsync(dq.p, query, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(dt.p, target, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaEventRecord(ctx->ev[0], ctx->stream));
	sw_words_kernel<<<grid_for(n, 128), 128, 0, ctx->stream>>>(n, dq.as<uint64_t>(), dt.as<uint64_t>(), dout.as<int>());
	CK(cudaGetLastError());
	CK(cudaEventRecord(ctx->ev[1], ctx->stream));
	int32_t *dst[5] = {score, q_start, q_stop, t_start, t_stop};
	for (int k = 0; k < 5; ++k)
		if (dst[k]) CK(cudaMemcpyAsync(dst[k], dout.as<int>() + (size_t)k * n, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (last_two) CK(cudaMemcpyAsync(last_two, dout.as<int>() + 5 * (size_t)n, (size_t)n * 2, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	cudaEventElapsedTime(&ctx->sw_ms_kernel, ctx->ev[0], ctx->ev[1]);
	ctx->stats.kernel_launches = 1;
	return 0;
}

int pcramp_gpu_sw_timing(pcramp_gpu_ctx *ctx, float *ms_kernel)
{ // CUDA-event time of sw_words_kernel in the last pcramp_gpu_sw_batch
	if (!ctx) return 1;
	if (ms_kernel) *ms_kernel = ctx->sw_ms_kernel;
	return 0;
}

int pcramp_gpu_background_match(pcramp_gpu_ctx *ctx, int kind, const uint64_t *f, const uint64_t *r, uint32_t n_pairs, float search_threshold,
	float detect_threshold, int amp_min, int amp_max, int taq, uint32_t *bitsets, uint64_t *n_amplicons)
{
	if (check_kind2(ctx, kind)) return 1;
	if (n_pairs && (!f || !r || !bitsets)) return fail(ctx, "pcramp_gpu_background_match: null argument");
	if (n_pairs >= (1u << 30)) return fail(ctx, "pcramp_gpu_background_match: too many pairs in one batch");
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[kind];
	cudaStream_t st = ctx->stream;
	if (!s.db_valid) return fail(ctx, "pcramp_gpu_background_match: no database (call pcramp_gpu_select_words first)");
	if (n_amplicons) *n_amplicons = 0;
	const uint32_t n_words = (s.n + 31u) / 32u;
	const size_t bits_bytes = std::max<size_t>(1, (size_t)n_pairs * n_words) * 4;
	ctx->stats.kernel_launches = 0;
	if (!n_pairs || !s.n) return 0;
	memset(bitsets, 0, (size_t)n_pairs * n_words * 4);
	if (!s.n_entries) return 0; // collect_background_candidates does nothing on an empty database (assay.h:411-421)
	if (db_words(ctx, s)) return 1; // the alignments read the words themselves
	DevBuf d_f, d_r, d_ol, d_bits;
	CK(d_f.ensure((size_t)n_pairs * 16));
	CK(d_r.ensure((size_t)n_pairs * 16));
	CK(d_ol.ensure((size_t)n_pairs * 2 * sizeof(OligoDev)));
	CK(d_bits.ensure(bits_bytes));
	CK(cudaMemcpyAsync(d_f.p, f, (size_t)n_pairs * 16, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_r.p, r, (size_t)n_pairs * 16, cudaMemcpyHostToDevice, st));
	CK(cudaMemsetAsync(d_bits.p, 0, bits_bytes, st));
	const float thr2 = search_threshold * search_threshold; // pcr_assay.cpp:31-32
	prep_oligos_kernel<<<grid_for(2ull * n_pairs, 256), 256, 0, st>>>(d_f.as<uint64_t>(), d_r.as<uint64_t>(), n_pairs, thr2, d_ol.as<OligoDev>());
	CK(cudaGetLastError());
	ctx->stats.kernel_launches++;
	if (ctx->use_background_units && (uint64_t)s.n * n_pairs * 4ull + 1ull < (1ull << 32)) { // (more units than 32-bit list offsets hold: the record form)
		if (background_match_units(ctx, s, d_f.as<uint64_t>(), d_r.as<uint64_t>(), d_ol.as<OligoDev>(), n_pairs, detect_threshold, amp_min, amp_max, taq,
				d_bits.as<uint32_t>(), n_words, n_amplicons)) return 1;
		CK(cudaMemcpyAsync(bitsets, d_bits.p, (size_t)n_pairs * n_words * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		return 0;
	}
	AmpList L;
	if (build_amplicon_list<false>(ctx, s, d_ol.as<OligoDev>(), n_pairs, amp_min, amp_max, L, "pcramp_gpu_background_match")) return 1;
	const uint64_t n = L.n;
	if (n_amplicons) *n_amplicons = n;
	if (n) {
		background_sw_kernel<<<grid_for(n, 128), 128, 0, st>>>(n, L.k1[0].as<uint64_t>(), L.perm[0].as<uint32_t>(), L.pid.as<uint32_t>(), L.mid.as<uint32_t>(),
			s.e_hi.as<uint64_t>(), s.e_lo.as<uint64_t>(), d_f.as<uint64_t>(), d_r.as<uint64_t>(), s.n, detect_threshold, taq, d_bits.as<uint32_t>(), n_words);
		CK(cudaGetLastError());
		ctx->stats.kernel_launches += 1;
	}
	CK(cudaMemcpyAsync(bitsets, d_bits.p, (size_t)n_pairs * n_words * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	return 0;
}

int pcramp_gpu_multiplex_background_match(pcramp_gpu_ctx *ctx, int kind, const uint64_t *f, const uint64_t *r, uint32_t n_pairs, float threshold,
	int taq, uint32_t *bitsets)
{
	if (check_kind2(ctx, kind)) return 1;
	if (n_pairs && (!f || !r || !bitsets)) return fail(ctx, "pcramp_gpu_multiplex_background_match: null argument");
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[kind];
	cudaStream_t st = ctx->stream;
	const uint32_t n_words = (s.n + 31u) / 32u;
	ctx->stats.kernel_launches = 0;
	if (!n_pairs || !s.n) return 0;
	for (uint32_t p = 0; p < n_pairs; ++p)
		if ((f[2 * p] | f[2 * p + 1]) == 0 || (r[2 * p] | r[2 * p + 1]) == 0) return fail(ctx, ":SeqOverlap::pack_query_slots: len == 0");
	const size_t bits_bytes = (size_t)n_pairs * n_words * 4;
	DevBuf d_f, d_r, d_bits;
	CK(d_f.ensure((size_t)n_pairs * 16));
	CK(d_r.ensure((size_t)n_pairs * 16));
	CK(d_bits.ensure(bits_bytes));
	CK(cudaMemcpyAsync(d_f.p, f, (size_t)n_pairs * 16, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_r.p, r, (size_t)n_pairs * 16, cudaMemcpyHostToDevice, st));
	CK(cudaMemsetAsync(d_bits.p, 0, bits_bytes, st));
	const uint64_t total = (uint64_t)n_pairs * 4ull * s.n;
	multiplex_sw_kernel<<<grid_for(total, 128), 128, 0, st>>>(s.dev(), d_f.as<uint64_t>(), d_r.as<uint64_t>(), n_pairs, threshold, taq,
		d_bits.as<uint32_t>(), n_words);
	CK(cudaGetLastError());
	ctx->stats.kernel_launches = 1;
	CK(cudaMemcpyAsync(bitsets, d_bits.p, bits_bytes, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	return 0;
}

} // extern "C"
