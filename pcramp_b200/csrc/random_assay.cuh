// random_assay.cuh -- candidate generation on the device (included at the end of thermo_abi.cu):
// PCR::random_assay (pcr_assay.cpp:580-734) with the seeding protocol of main.cpp:527-548.
//
// The reference draws its trial assays inside `#pragma omp parallel`: every OpenMP thread owns one NucCruc object and one
// seed (local_seed = rand_r(&global_seed)), and works through a contiguous chunk of the trials (static schedule).  Inside a
// thread everything is a serial chain: each rejected candidate consumes rand_r draws, and which candidates are rejected
// depends on the thermodynamic filters (PCR::is_valid, PCR::max_dimer_tm), whose hairpin / dimer evaluation can read two
// slots past the end of the oligo last loaded into the thread's NucCruc object (nuccruc.cuh header) -- so the result of a
// chain depends on its seed, on the order of its loads, and on nothing outside it.  That is the unit of parallelism here:
// one GPU thread per chain ("stream" = one OpenMP thread of the reference), carrying the seed and the two ring buffers of
// its NucCruc object in registers / local memory and calling the same device functions as thermo_kernel for every Tm.
// A run with as many streams as trials (the reference with --thread num_trial) is the GPU-friendly setting; a single stream
// reproduces `--thread 1`.
//
// rand_r is glibc's (stdlib/rand_r.c: three steps of the 1103515245 / 12345 LCG giving 11 + 10 + 10 bits); the reference
// links it from libc, so it is restated here and pinned by the tests against the compiled reference.
#pragma once
#include "ctx.cuh"
#include "thermo.cuh"

namespace pcr {
namespace nc {

struct RandomAssayParams {
	int primer_min, primer_max, amp_min, amp_max;
	double max_degen;     // m_opt.degen (unsigned int) as the double it is compared in
	float tm_min, tm_max, max_hairpin, max_dimer;
	uint32_t n_degen;     // achievable oligo degeneracies <= max_degen, ascending
	const uint32_t *degen_value;  // [n_degen]
	const float *log_strand;      // [n_degen]: logf(float(primer_strand / d)), host libm
	const float *log_hetero;      // [n_degen x n_degen]: logf(strand(c_f, c_r)) (nuc_cruc.h:818-838)
	const uint32_t *active;       // indices of the active sequences (pcr_assay.cpp:593-601)
	uint32_t n_active;
};

enum : uint32_t { RA_OK = 0, RA_NO_ASSAY = 1, RA_SHORT_SEQUENCE = 2 };

__device__ __forceinline__ int rand_r_dev(unsigned int &seed)
{ // glibc stdlib/rand_r.c
	unsigned int next = seed;
	next = next * 1103515245u + 12345u;
	int result = (int)((next / 65536u) % 2048u);
	next = next * 1103515245u + 12345u;
	result <<= 10;
	result ^= (int)((next / 65536u) % 1024u);
	next = next * 1103515245u + 12345u;
	result <<= 10;
	result ^= (int)((next / 65536u) % 1024u);
	seed = next;
	return result;
}

struct Oligo5 { // an oligo 5'->3' as nibbles
	unsigned char nib[32];
	int len;
	double degen;
	int degen_idx;
};

__device__ inline int degen_index(const RandomAssayParams &P, double d)
{
	for (uint32_t i = 0; i < P.n_degen; ++i)
		if ((double)P.degen_value[i] == d) return (int)i;
	return -1;
}

// the idx-th concrete expansion (Word::begin / Word::next, word.h:525-647: position 0 fastest, letters in bit order A C G T)
__device__ inline void expansion_codes_dev(const Oligo5 &o, uint32_t idx, unsigned char *dst)
{
	for (int i = 0; i < o.len; ++i) {
		const uint32_t nib = o.nib[i];
		const uint32_t k = (uint32_t)__popc(nib), pick = idx % k;
		idx /= k;
		uint32_t rest = nib;
		for (uint32_t s = 0; s < pick; ++s) rest &= rest - 1u;
		dst[i] = (unsigned char)(__ffs((int)rest) - 1); // bit 0..3 = A C G T = base codes 0..3
	}
}

// PCR::is_valid (valid_pcr.cpp:5-45) with m_check_homo_dimer = true, on the stream's NucCruc object (qring)
__device__ inline bool is_valid_dev(Ctx &c, const RandomAssayParams &P, const Oligo5 &o, unsigned char *qring)
{
	const uint32_t count = (uint32_t)o.degen;
	c.log_strand = P.log_strand[o.degen_idx];
	unsigned char e[NC_SEQ_CAP];
	for (int k = 0; k < NC_SEQ_CAP; ++k) e[k] = 0;
	for (uint32_t x = 0; x < count; ++x) {
		expansion_codes_dev(o, x, e);
		c.q = e; c.t = e; c.qlen = c.tlen = o.len;
		float tm = run_problem(c, OP_PM_DUPLEX).tm;
		if ((tm < P.tm_min) || (tm > P.tm_max)) return false;
		for (int k = 0; k < o.len; ++k) qring[k] = e[k]; // set_query (nuc_cruc.h:875-913): the slots past the end keep their content
		c.q = qring; c.t = qring;
		tm = run_problem(c, OP_HAIRPIN).tm;
		if (tm > P.max_hairpin) return false;
		tm = run_problem(c, OP_HOMODIMER).tm;
		if (tm > P.max_dimer) return false;
	}
	return true;
}

// PCR::max_dimer_tm (pcr_assay.cpp:232-269)
__device__ inline float max_dimer_dev(Ctx &c, const RandomAssayParams &P, const Oligo5 &f, const Oligo5 &r, unsigned char *qring, unsigned char *tring)
{
	float ret = 0.0f;
	c.log_strand = P.log_hetero[(size_t)f.degen_idx * P.n_degen + r.degen_idx];
	unsigned char e[32];
	const uint32_t nf = (uint32_t)f.degen, nr = (uint32_t)r.degen;
	for (uint32_t x = 0; x < nf; ++x) {
		expansion_codes_dev(f, x, e);
		for (int k = 0; k < f.len; ++k) qring[k] = e[k]; // set_query
		for (uint32_t y = 0; y < nr; ++y) {
			expansion_codes_dev(r, y, e);
			for (int k = 0; k < r.len; ++k) tring[k] = e[k]; // set_target
			c.q = qring; c.t = tring; c.qlen = f.len; c.tlen = r.len;
			const float tm = run_problem(c, OP_HETERODIMER).tm;
			ret = fmaxf(ret, tm);
		}
	}
	return ret;
}

// Sequence::subword (sequence.cpp:269-302) as 5'->3' nibbles; false when the word holds an EOS (size() != length,
// pcr_assay.cpp:649,691) or is too degenerate (:653-657,695-699)
__device__ inline bool cut_oligo(const SeqDev &sd, uint32_t seq, int start, int len, bool revcomp, const RandomAssayParams &P, Oligo5 &o)
{
	o.len = len;
	double d = 1.0;
	bool whole = true;
	for (int k = 0; k < len; ++k) {
		uint32_t nib = raw_nibble_at(sd, seq, (uint32_t)(start + k));
		if (nib == 0u) whole = false;
		else d *= (double)__popc(nib);
		if (revcomp) {
			nib = ((nib & 1u) << 3) | ((nib & 2u) << 1) | ((nib & 4u) >> 1) | ((nib & 8u) >> 3); // A<->T, C<->G
			o.nib[len - 1 - k] = (unsigned char)nib;
		} else {
			o.nib[k] = (unsigned char)nib;
		}
	}
	o.degen = d;
	if (!whole || d > P.max_degen) return false;
	o.degen_idx = degen_index(P, d);
	return o.degen_idx >= 0;
}

__device__ inline W128 oligo_word_centred(const Oligo5 &o)
{ // subword() fills an empty Word from position 0 (Word::push_back, word.cpp:31-48) and complement() writes left-justified
	// (word.h:140-183); then PCR::center() (pcr_assay.cpp:720, assay.h:395-399)
	W128 w;
	w.hi = w.lo = 0;
	for (int k = 0; k < o.len; ++k) w_set(w, k, o.nib[k]);
	return w_center(w);
}

// lanes_per_stream = 32: one stream per WARP (lane 0 works).  The streams walk their own accept / reject paths, so the lanes of a warp
// that hold different streams run one after the other; with fewer streams than the GPU has warp slots a warp of its own costs nothing
// and takes that serialisation away (16 streams: 130 -> ~35 ms per design iteration).  lanes_per_stream = 1 packs them for large counts.
__global__ void __launch_bounds__(64) random_assay_kernel(SeqDev sd, RandomAssayParams P, uint32_t n_streams, uint32_t lanes_per_stream, uint32_t *seeds,
	const uint32_t *__restrict__ first_trial, const Tables *__restrict__ tables, const DpTable *__restrict__ dp, uint64_t *f_out, uint64_t *r_out,
	uint32_t *attempts, uint32_t *status)
{
	__shared__ DpTable s_dp;
	{
		const int *src = (const int *)dp;
		int *dst = (int *)&s_dp;
		for (int k = threadIdx.x; k < (int)(sizeof(DpTable) / sizeof(int)); k += blockDim.x) dst[k] = src[k];
	}
	__syncthreads();
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot % lanes_per_stream) return;
	const uint32_t stream = slot / lanes_per_stream;
	if (stream >= n_streams) return;
	unsigned int seed = seeds[stream];
	__align__(16) unsigned char qring[NC_SEQ_CAP], tring[NC_SEQ_CAP]; // NucCruc::query / target of a fresh object
	for (int k = 0; k < NC_SEQ_CAP; ++k) qring[k] = tring[k] = 0;
	__align__(16) unsigned short info[NC_CELLS];
	Ctx c;
	c.T = tables;
	c.D = &s_dp;
	c.info = info + NC_INFO_PAD;
	uint32_t st = RA_OK;
	for (uint32_t t = first_trial[stream]; t < first_trial[stream + 1] && st == RA_OK; ++t) {
		uint32_t sequence_iteration = 0, total_attempts = 0;
		bool done = false;
		while (!done) {
			if (++sequence_iteration > 100u) { st = RA_NO_ASSAY; break; }                    // pcr_assay.cpp:612-614
			const uint32_t seq = P.active[(uint32_t)rand_r_dev(seed) % P.n_active];           // :619
			const int len = (int)sd.len[seq];
			if (len < P.amp_min) { st = RA_SHORT_SEQUENCE; break; }                           // :623-625
			for (uint32_t assay_iteration = 1; assay_iteration <= 100u; ++assay_iteration) { // :631-637
				++total_attempts;
				const int span = P.primer_max - P.primer_min + 1;
				const int f_len = P.primer_min + rand_r_dev(seed) % span;
				const int r_len = P.primer_min + rand_r_dev(seed) % span;
				if (f_len + r_len > len) continue;
				const int f_start = rand_r_dev(seed) % ((len + 1) - P.amp_min);              // random_location (sample.cpp:6-12)
				Oligo5 F, R;
				if (!cut_oligo(sd, seq, f_start, f_len, false, P, F)) continue;
				if (!is_valid_dev(c, P, F, qring)) continue;
				const int r_lo = f_start + P.amp_min - r_len, r_hi = min((len + 1) - r_len, (f_start + P.amp_max + 1) - r_len);
				const int r_start = r_lo + rand_r_dev(seed) % (r_hi - r_lo);
				const int amp_len = r_start - f_start + r_len;
				if (amp_len > P.amp_max || amp_len < P.amp_min) continue;
				if (!cut_oligo(sd, seq, r_start, r_len, true, P, R)) continue;
				if (has_split_dev(sd, seq, f_start, amp_len)) continue;                        // :703
				if (!is_valid_dev(c, P, R, qring)) continue;
				if (max_dimer_dev(c, P, F, R, qring, tring) > P.max_dimer) continue;           // :715
				const W128 fw = oligo_word_centred(F), rw = oligo_word_centred(R);
				f_out[2 * (size_t)t] = fw.hi; f_out[2 * (size_t)t + 1] = fw.lo;
				r_out[2 * (size_t)t] = rw.hi; r_out[2 * (size_t)t + 1] = rw.lo;
				done = true;
				break;
			}
		}
		if (attempts) attempts[t] = total_attempts;
	}
	seeds[stream] = seed;
	status[stream] = st;
}

} // namespace nc
} // namespace pcr

extern "C" {

int pcramp_gpu_random_assays(pcramp_gpu_ctx *ctx, int kind, uint32_t n_streams, uint32_t *seeds, const uint32_t *trials_per_stream,
	const pcramp_gpu_random_assay_options *o, uint64_t *f, uint64_t *r, uint32_t *attempts)
{
	if (!ctx) return 1;
	if (kind < 0 || kind >= PCRAMP_NUM_KINDS) return fail(ctx, "pcramp_gpu: bad sequence kind");
	if (!o || (n_streams && (!seeds || !trials_per_stream || !f || !r))) return fail(ctx, "pcramp_gpu_random_assays: null argument");
	if (o->primer_min < 1 || o->primer_max > NC_MAX_LEN || o->primer_min > o->primer_max)
		return fail(ctx, "pcramp_gpu_random_assays: primer range must lie in [1, 32]");
	if (o->amplicon_min < 1 || o->amplicon_min > o->amplicon_max) return fail(ctx, "pcramp_gpu_random_assays: bad amplicon range");
	if (o->degen < 1 || o->degen > 4096) return fail(ctx, "pcramp_gpu_random_assays: degen must lie in [1, 4096]");
	CK(cudaSetDevice(ctx->device));
	ThermoState *t = nullptr;
	if (thermo_get(ctx, &t)) return 1;
	if (thermo_set_salt(ctx, t, o->salt)) return 1;
	pcr::SeqSet &s = ctx->sets[kind];
	cudaStream_t st = ctx->stream;
	std::vector<uint32_t> active;
	for (uint32_t i = 0; i < s.n; ++i)
		if (s.active[i]) active.push_back(i);
	std::vector<uint32_t> first(n_streams + 1, 0);
	for (uint32_t i = 0; i < n_streams; ++i) first[i + 1] = first[i] + trials_per_stream[i];
	const uint32_t n_trials = first[n_streams];
	if (!n_trials) return 0;
	if (active.empty()) return fail(ctx, ":PCR::random_assay: No active sequences found"); // pcr_assay.cpp:605-607
	// achievable degeneracies (products of 2, 3 and 4 = 2^a 3^b) and the logarithms the reference takes with libm
	std::vector<uint32_t> dv;
	for (uint64_t a = 1; a <= o->degen; a *= 2)
		for (uint64_t b = a; b <= o->degen; b *= 3) dv.push_back((uint32_t)b);
	std::sort(dv.begin(), dv.end());
	const uint32_t nd = (uint32_t)dv.size();
	std::vector<float> ls(nd), lh((size_t)nd * nd);
	for (uint32_t i = 0; i < nd; ++i) {
		const float ca = (float)((double)o->primer_strand / (double)dv[i]); // valid_pcr.cpp:13
		if (ca < 0.0f) return fail(ctx, ":strand: strand_concentration < 0.0f");
		if (!(ca > 0.0f)) return fail(ctx, ":NucCruc::tm_dimer: Invalid strand_concentration");
		ls[i] = logf(ca);
		for (uint32_t j = 0; j < nd; ++j) {
			const float cb = (float)((double)o->primer_strand / (double)dv[j]); // pcr_assay.cpp:244
			lh[(size_t)i * nd + j] = logf(hetero_strand(ca, cb));
		}
	}
	DevBuf d_dv, d_ls, d_lh, d_act, d_seed, d_first, d_f, d_r, d_att, d_status;
	CK(d_dv.ensure(nd * 4)); CK(d_ls.ensure(nd * 4)); CK(d_lh.ensure((size_t)nd * nd * 4));
	CK(d_act.ensure(active.size() * 4)); CK(d_seed.ensure((size_t)n_streams * 4)); CK(d_first.ensure(first.size() * 4));
	CK(d_f.ensure((size_t)n_trials * 16)); CK(d_r.ensure((size_t)n_trials * 16)); CK(d_att.ensure((size_t)n_trials * 4));
	CK(d_status.ensure((size_t)n_streams * 4));
	CK(cudaMemcpyAsync(d_dv.p, dv.data(), nd * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_ls.p, ls.data(), nd * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_lh.p, lh.data(), (size_t)nd * nd * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_act.p, active.data(), active.size() * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_seed.p, seeds, (size_t)n_streams * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(d_first.p, first.data(), first.size() * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemsetAsync(d_f.p, 0, (size_t)n_trials * 16, st));
	CK(cudaMemsetAsync(d_r.p, 0, (size_t)n_trials * 16, st));
	CK(cudaMemsetAsync(d_att.p, 0, (size_t)n_trials * 4, st));
	RandomAssayParams P;
	P.primer_min = o->primer_min; P.primer_max = o->primer_max; P.amp_min = o->amplicon_min; P.amp_max = o->amplicon_max;
	P.max_degen = (double)o->degen;
	P.tm_min = o->primer_tm_min; P.tm_max = o->primer_tm_max; P.max_hairpin = o->max_hairpin; P.max_dimer = o->max_dimer;
	P.n_degen = nd;
	P.degen_value = d_dv.as<uint32_t>(); P.log_strand = d_ls.as<float>(); P.log_hetero = d_lh.as<float>();
	P.active = d_act.as<uint32_t>(); P.n_active = (uint32_t)active.size();
	CK(cudaEventRecord(t->ev0, st));
	// a warp per stream while the streams are fewer than the warps the GPU holds at this kernel's occupancy
	const uint32_t lanes = (uint64_t)n_streams <= (uint64_t)ctx->sm_count * 16u ? 32u : 1u;
	random_assay_kernel<<<grid_for((uint64_t)n_streams * lanes, 64), 64, 0, st>>>(s.dev(), P, n_streams, lanes, d_seed.as<uint32_t>(), d_first.as<uint32_t>(),
		t->d_tables.as<Tables>(), t->d_dp.as<DpTable>(), d_f.as<uint64_t>(), d_r.as<uint64_t>(), d_att.as<uint32_t>(), d_status.as<uint32_t>());
	CK(cudaGetLastError());
	CK(cudaEventRecord(t->ev1, st));
	std::vector<uint32_t> status(n_streams);
	CK(cudaMemcpyAsync(status.data(), d_status.p, (size_t)n_streams * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(seeds, d_seed.p, (size_t)n_streams * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(f, d_f.p, (size_t)n_trials * 16, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(r, d_r.p, (size_t)n_trials * 16, cudaMemcpyDeviceToHost, st));
	if (attempts) CK(cudaMemcpyAsync(attempts, d_att.p, (size_t)n_trials * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	float ms = 0.0f;
	cudaEventElapsedTime(&ms, t->ev0, t->ev1);
	t->stats.kernel_launches = 1;
	t->stats.n_problems = n_trials;
	t->stats.dp_cells = 0;
	t->stats.ms_kernel = ms;
	for (uint32_t i = 0; i < n_streams; ++i) {
		if (status[i] == RA_NO_ASSAY) return fail(ctx, ":PCR::random_assay: Unable to generate a valid initial assay to test!");
		if (status[i] == RA_SHORT_SEQUENCE) return fail(ctx, ":PCR::random_assay: sequence length is too small!");
	}
	return 0;
}

} // extern "C"
