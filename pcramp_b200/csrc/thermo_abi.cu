// thermo_abi.cu -- K3 host side: batches NucCruc problems onto the GPU and mirrors the thermodynamic filters
// PCR::is_valid / max_dimer_tm / multiplex_compatible (valid_pcr.cpp:5-45, pcr_assay.cpp:232-269,815-852).
// Part of libpcramp_gpu.so; no CPU evaluation path exists here: every Tm comes from thermo_kernel.
#include "ctx.cuh"
#include "thermo.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

using namespace pcr;
using namespace pcr::nc;

namespace pcr {
namespace nc {

struct PinnedBuf {
	void *p = nullptr;
	size_t cap = 0;
	~PinnedBuf() { release(); }
	void release()
	{
		if (p) cudaFreeHost(p);
		p = nullptr;
		cap = 0;
	}
	cudaError_t ensure(size_t bytes)
	{
		if (bytes <= cap) return cudaSuccess;
		release();
		const size_t want = bytes + bytes / 4 + 256;
		cudaError_t e = cudaMallocHost(&p, want);
		if (e != cudaSuccess) { p = nullptr; return e; }
		cap = want;
		return cudaSuccess;
	}
	template <class T> T *as() const { return (T *)p; }
};

struct ThermoState {
	bool tables_ready = false;
	Tables h_tables;
	DpTable h_dp;
	float dp_salt = -1.0f;
	DevBuf d_tables, d_dp, d_a, d_b, d_la, d_lb, d_ls, d_out;
	DevBuf d_text_a, d_text_b, d_note, d_fields; // string batches: the caller's text as it is, {first error, cell count}, one array per result field
	PinnedBuf h_note;
	// pipelined batches (thermo_batch_pipelined): one stream, one sort scratch and one join event per chunk
	static constexpr int PIPE = 4;
	cudaStream_t pipe[PIPE] = {};
	DevBuf d_pipe_tmp[PIPE];
	cudaEvent_t ev_fork = nullptr, ev_join[PIPE] = {}, ev_in[PIPE] = {};
	DevBuf d_key[2], d_ord[2], d_sort_tmp; // size-binned launch order (thermo_order)
	const uint32_t *order = nullptr;
	PinnedBuf h_a, h_b, h_la, h_lb, h_ls, h_out;
	int op = -1;
	uint32_t n = 0;
	uint64_t cells = 0;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	pcramp_gpu_thermo_stats stats = {};
};

void thermo_state_free(ThermoState *t)
{
	if (!t) return;
	if (t->ev0) cudaEventDestroy(t->ev0);
	if (t->ev1) cudaEventDestroy(t->ev1);
	if (t->ev_fork) cudaEventDestroy(t->ev_fork);
	for (int k = 0; k < ThermoState::PIPE; ++k) {
		if (t->ev_join[k]) cudaEventDestroy(t->ev_join[k]);
		if (t->ev_in[k]) cudaEventDestroy(t->ev_in[k]);
		if (t->pipe[k]) { cudaStreamSynchronize(t->pipe[k]); cudaStreamDestroy(t->pipe[k]); }
	}
	delete t;
}

} // namespace nc
} // namespace pcr

namespace {

int thermo_get(pcramp_gpu_ctx *ctx, ThermoState **out)
{
	if (!ctx->thermo) {
		ThermoState *t = new ThermoState();
		build_tables(t->h_tables);
		ctx->thermo = t;
		CK(cudaEventCreate(&t->ev0));
		CK(cudaEventCreate(&t->ev1));
		CK(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
	}
	ThermoState *t = ctx->thermo;
	if (!t->tables_ready) {
		CK(t->d_tables.ensure(sizeof(Tables)));
		CK(cudaMemcpyAsync(t->d_tables.p, &t->h_tables, sizeof(Tables), cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		t->tables_ready = true;
	}
	*out = t;
	return 0;
}

int thermo_set_salt(pcramp_gpu_ctx *ctx, ThermoState *t, float salt)
{ // NucCruc::salt (nuc_cruc.h:779-794)
	if (!(salt >= 1.0e-6f)) return fail(ctx, ":salt: [Na+] < 1.0e-6f");
	if (salt > 1.0f) return fail(ctx, ":salt: [Na+] > 1.0f");
	if (salt == t->dp_salt) return 0;
	CK(cudaStreamSynchronize(ctx->stream)); // a previous launch may still read the table
	build_dp(t->h_tables, salt, 310.15f, t->h_dp);
	CK(t->d_dp.ensure(sizeof(DpTable)));
	CK(cudaMemcpyAsync(t->d_dp.p, &t->h_dp, sizeof(DpTable), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	t->dp_salt = salt;
	return 0;
}

inline bool two_sequences(int op) { return op == OP_HETERODIMER || op == OP_HETERODIMER_DIAG; }
inline bool needs_strand(int op) { return op != OP_HAIRPIN; }

// reserve pinned staging for n problems
int thermo_reserve(pcramp_gpu_ctx *ctx, ThermoState *t, uint32_t n)
{
	// single staging buffers: a batch staged earlier may still be copying out of them (stage -> stage without a fetch in between)
	CK(cudaStreamSynchronize(ctx->stream));
	const size_t m = n ? n : 1;
	CK(t->h_a.ensure(m * THERMO_SEQ_STRIDE));
	CK(t->h_b.ensure(m * THERMO_SEQ_STRIDE));
	CK(t->h_la.ensure(m));
	CK(t->h_lb.ensure(m));
	CK(t->h_ls.ensure(m * sizeof(float)));
	CK(t->h_out.ensure(m * sizeof(float4)));
	CK(t->d_a.ensure(m * THERMO_SEQ_STRIDE));
	CK(t->d_b.ensure(m * THERMO_SEQ_STRIDE));
	CK(t->d_la.ensure(m));
	CK(t->d_lb.ensure(m));
	CK(t->d_ls.ensure(m * sizeof(float)));
	CK(t->d_out.ensure(m * sizeof(float4)));
	return 0;
}

int thermo_order(pcramp_gpu_ctx *ctx, ThermoState *t);

// problems already encoded in the pinned staging buffers -> HBM
int thermo_upload(pcramp_gpu_ctx *ctx, ThermoState *t, int op, uint32_t n)
{
	t->op = op;
	t->n = n;
	t->order = nullptr;
	uint64_t cells = 0;
	const uint8_t *la = t->h_la.as<uint8_t>(), *lb = t->h_lb.as<uint8_t>();
	for (uint32_t p = 0; p < n; ++p) cells += (uint64_t)problem_cells(op, la[p], two_sequences(op) ? lb[p] : la[p]);
	t->cells = cells;
	if (!n) return 0;
	CK(cudaMemcpyAsync(t->d_a.p, t->h_a.p, (size_t)n * THERMO_SEQ_STRIDE, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(t->d_la.p, t->h_la.p, n, cudaMemcpyHostToDevice, ctx->stream));
	if (two_sequences(op)) {
		CK(cudaMemcpyAsync(t->d_b.p, t->h_b.p, (size_t)n * THERMO_SEQ_STRIDE, cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaMemcpyAsync(t->d_lb.p, t->h_lb.p, n, cudaMemcpyHostToDevice, ctx->stream));
	}
	CK(cudaMemcpyAsync(t->d_ls.p, t->h_ls.p, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
	return thermo_order(ctx, t);
}

// Launch order: a warp pays for its longest problem (rows x column strips of the DP fill), so problems of equal size are
// put next to each other: slot s of the launch works on problem order[s] (results still land in out[problem]).
constexpr uint32_t THERMO_ORDER_MIN = 2048; // below this the launch is a handful of warps anyway
constexpr uint32_t THERMO_CHUNK = 1u << 20;  // problems per launch of the library's own large batches (thermo_run_codes, multiplex_compatible)

__global__ void thermo_key_kernel(int op, uint32_t n, const uint8_t *__restrict__ len_a, const uint8_t *__restrict__ len_b, uint16_t *key, uint32_t *ord)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	const uint32_t q = len_a[p], t = (op == OP_HETERODIMER || op == OP_HETERODIMER_DIAG) ? len_b[p] : q;
	key[p] = (uint16_t)(0xfffu - ((t << 6) | q)); // columns (strips) first, then rows; largest first, so that the last wave of CTAs is the cheapest
	ord[p] = p;
}

int thermo_order(pcramp_gpu_ctx *ctx, ThermoState *t)
{
	t->order = nullptr;
	const uint32_t n = t->n;
	if (n < THERMO_ORDER_MIN || t->op == OP_PM_DUPLEX) return 0;
	for (int k = 0; k < 2; ++k) {
		CK(t->d_key[k].ensure((size_t)n * 2));
		CK(t->d_ord[k].ensure((size_t)n * 4));
	}
	thermo_key_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(t->op, n, t->d_la.as<uint8_t>(), t->d_lb.as<uint8_t>(), t->d_key[0].as<uint16_t>(),
		t->d_ord[0].as<uint32_t>());
	CK(cudaGetLastError());
	size_t tb = 0;
	CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, t->d_key[0].as<uint16_t>(), t->d_key[1].as<uint16_t>(), t->d_ord[0].as<uint32_t>(),
		t->d_ord[1].as<uint32_t>(), (int)n, 0, 12, ctx->stream));
	CK(t->d_sort_tmp.ensure(tb));
	CK(cub::DeviceRadixSort::SortPairs(t->d_sort_tmp.p, tb, t->d_key[0].as<uint16_t>(), t->d_key[1].as<uint16_t>(), t->d_ord[0].as<uint32_t>(),
		t->d_ord[1].as<uint32_t>(), (int)n, 0, 12, ctx->stream));
	t->order = t->d_ord[1].as<uint32_t>();
	return 0;
}

int thermo_launch(pcramp_gpu_ctx *ctx, ThermoState *t)
{
	t->stats.n_problems = t->n;
	t->stats.dp_cells = t->cells;
	t->stats.kernel_launches = 0;
	t->stats.ms_kernel = 0.0f;
	if (!t->n) return 0;
	CK(cudaEventRecord(t->ev0, ctx->stream));
	thermo_kernel<<<grid_for(t->n, THERMO_BLOCK), THERMO_BLOCK, 0, ctx->stream>>>(t->op, t->n, t->order, t->d_a.as<uint8_t>(), t->d_b.as<uint8_t>(),
		t->d_la.as<uint8_t>(), t->d_lb.as<uint8_t>(), t->d_ls.as<float>(), t->d_tables.as<Tables>(), t->d_dp.as<DpTable>(), t->d_out.as<float4>());
	CK(cudaGetLastError());
	CK(cudaEventRecord(t->ev1, ctx->stream));
	t->stats.kernel_launches = 1;
	return 0;
}

int thermo_download(pcramp_gpu_ctx *ctx, ThermoState *t)
{
	if (t->n) CK(cudaMemcpyAsync(t->h_out.p, t->d_out.p, (size_t)t->n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (t->n && t->stats.kernel_launches) {
		float ms = 0.0f;
		cudaEventElapsedTime(&ms, t->ev0, t->ev1);
		t->stats.ms_kernel = ms;
	}
	return 0;
}

// ASCII -> base codes into slot p of a staging buffer; mirrors set_query's / tm_pm_duplex's throws (message returned)
const char *encode_seq(const char *s, uint32_t stride, bool allow_inosine, uint8_t *dst, uint8_t *len_out)
{
	static const struct Lut {
		int8_t v[256];
		Lut()
		{
			for (int i = 0; i < 256; ++i) v[i] = (int8_t)base_code((char)i);
		}
	} lut;
	uint64_t *d64 = (uint64_t *)dst;
	d64[0] = d64[1] = d64[2] = d64[3] = 0;
	uint32_t len = 0;
	while (len < stride && s[len]) {
		if (len >= (uint32_t)NC_MAX_LEN) return "pcramp_gpu_thermo: sequence longer than 32 bases (Word length)";
		const int c = lut.v[(unsigned char)s[len]];
		if (c < 0 || (c == bI && !allow_inosine)) return allow_inosine ? ":set_query: Illegal base" : "Unknown base in tm_pm_duplex";
		dst[len] = (uint8_t)c;
		++len;
	}
	*len_out = (uint8_t)len;
	return nullptr;
}

inline float hetero_strand(float c_a, float c_b)
{ // NucCruc::strand(c_a, c_b), nuc_cruc.h:818-838
	return (c_a > c_b) ? c_a - 0.5f * c_b : c_b - 0.5f * c_a;
}

// problems [lo, hi): encode, validate, take logf of the strand concentration; returns the first error message or null
const char *stage_range(ThermoState *t, int op, uint32_t lo, uint32_t hi, const char *seq_a, const char *seq_b, uint32_t stride,
	const float *strand_a, const float *strand_b)
{
	uint8_t *ha = t->h_a.as<uint8_t>(), *hb = t->h_b.as<uint8_t>(), *la = t->h_la.as<uint8_t>(), *lb = t->h_lb.as<uint8_t>();
	float *ls = t->h_ls.as<float>();
	float last_strand = -1.0f, last_log = 0.0f;
	for (uint32_t p = lo; p < hi; ++p) {
		if (const char *e = encode_seq(seq_a + (size_t)p * stride, stride, op != OP_PM_DUPLEX, ha + (size_t)p * THERMO_SEQ_STRIDE, la + p)) return e;
		if (op == OP_HAIRPIN && la[p] == 0) return ":NucCruc::align_hairpin: Empty query sequence";
		float strand = 1.0f;
		if (two_sequences(op)) {
			if (const char *e = encode_seq(seq_b + (size_t)p * stride, stride, true, hb + (size_t)p * THERMO_SEQ_STRIDE, lb + p)) return e;
			if (strand_a[p] < 0.0f) return ":strand: m_c_a < 0.0f";
			if (strand_b[p] < 0.0f) return ":strand: m_c_b < 0.0f";
			strand = hetero_strand(strand_a[p], strand_b[p]);
		} else {
			lb[p] = 0;
			if (needs_strand(op)) {
				if (strand_a[p] < 0.0f) return ":strand: strand_concentration < 0.0f";
				strand = strand_a[p];
			}
		}
		if (op != OP_PM_DUPLEX && op != OP_HAIRPIN && !(strand > 0.0f)) return ":NucCruc::tm_dimer: Invalid strand_concentration";
		if (strand != last_strand) { // the float overload the reference's log() resolves to (nuc_cruc.cpp:2129)
			last_strand = strand;
			last_log = logf(strand);
		}
		ls[p] = last_log;
	}
	return nullptr;
}

// ---- string batches: the text is encoded on the device ---------------------------------------------------------
// Error of problem p, ordered as stage_range meets them: key = p << 8 | stage << 4 | detail (stage 0: first sequence, 1: empty
// hairpin query, 2: second sequence, 3: strand concentration -- the host's); the smallest key is the error the caller sees.
constexpr unsigned long long NOTE_NONE = ~0ull;

__device__ inline unsigned long long encode_text(const char *__restrict__ s, uint32_t stride, bool allow_inosine, uint8_t *dst, uint8_t *len_out,
	unsigned long long p, unsigned stage)
{ // encode_seq on the device
	uint4 *d = (uint4 *)dst;
	d[0] = d[1] = make_uint4(0u, 0u, 0u, 0u);
	uint32_t len = 0;
	unsigned long long err = NOTE_NONE;
	while (len < stride) {
		const char ch = s[len];
		if (!ch) break;
		if (len >= (uint32_t)NC_MAX_LEN) { err = (p << 8) | (stage << 4) | 0u; break; }
		int c;
		switch (ch) {
		case 'A': case 'a': c = bA; break;
		case 'C': case 'c': c = bC; break;
		case 'G': case 'g': c = bG; break;
		case 'T': case 't': c = bT; break;
		case 'I': case 'i': c = bI; break;
		default: c = -1;
		}
		if (c < 0 || (c == bI && !allow_inosine)) { err = (p << 8) | (stage << 4) | 1u; break; }
		dst[len] = (uint8_t)c;
		++len;
	}
	*len_out = (uint8_t)len;
	return err;
}

// problems [p0, p0 + n) of a batch; every pointer is the range's own
__global__ void __launch_bounds__(256) thermo_encode_kernel(int op, uint32_t n, uint32_t p0, const char *__restrict__ text_a, const char *__restrict__ text_b,
	uint32_t stride, uint8_t *seq_a, uint8_t *seq_b, uint8_t *len_a, uint8_t *len_b, unsigned long long *note)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned long long gp = (unsigned long long)p0 + p;
	unsigned long long err = NOTE_NONE, cells = 0;
	if (p < n) {
		const bool two = (op == OP_HETERODIMER || op == OP_HETERODIMER_DIAG);
		uint8_t la = 0, lb = 0;
		err = encode_text(text_a + (size_t)p * stride, stride, op != OP_PM_DUPLEX, seq_a + (size_t)p * THERMO_SEQ_STRIDE, &la, gp, 0u);
		if (err == NOTE_NONE && op == OP_HAIRPIN && la == 0) err = (gp << 8) | (1u << 4);
		if (err == NOTE_NONE && two) err = encode_text(text_b + (size_t)p * stride, stride, true, seq_b + (size_t)p * THERMO_SEQ_STRIDE, &lb, gp, 2u);
		len_a[p] = la;
		len_b[p] = lb;
		cells = (unsigned long long)problem_cells(op, la, two ? lb : la);
	}
	for (int d = 16; d; d >>= 1) {
		cells += __shfl_xor_sync(0xffffffffu, cells, d);
		const unsigned long long o = __shfl_xor_sync(0xffffffffu, err, d);
		err = o < err ? o : err;
	}
	if ((threadIdx.x & 31u) == 0u) {
		if (cells) atomicAdd(note + 1, cells);
		if (err != NOTE_NONE) atomicMin(note, err);
	}
}

// the same from words (word.h: 4-bit codes, what PCR::is_valid / max_dimer_tm hold): Word::str()'s bases from start() to stop(),
// every one a single letter (a degenerate base or an EOS inside the oligo is an error, as its text would be)
__device__ inline unsigned long long encode_word(const uint64_t *__restrict__ w, uint8_t *dst, uint8_t *len_out, unsigned long long p, unsigned stage)
{
	uint4 *d = (uint4 *)dst;
	d[0] = d[1] = make_uint4(0u, 0u, 0u, 0u);
	W128 x;
	x.hi = w[0];
	x.lo = w[1];
	const int first = w_start(x), last = w_stop(x);
	uint32_t len = 0;
	unsigned long long err = NOTE_NONE;
	for (int i = first; i >= 0 && i <= last; ++i) {
		const uint32_t nib = w_get(x, i);
		int c;
		switch (nib) {
		case 1u: c = bA; break;
		case 2u: c = bC; break;
		case 4u: c = bG; break;
		case 8u: c = bT; break;
		default: c = -1;
		}
		if (c < 0) { err = (p << 8) | (stage << 4) | 1u; break; }
		dst[len++] = (uint8_t)c;
	}
	*len_out = (uint8_t)len;
	return err;
}

__global__ void __launch_bounds__(256) thermo_encode_words_kernel(int op, uint32_t n, uint32_t p0, const uint64_t *__restrict__ words_a,
	const uint64_t *__restrict__ words_b, uint8_t *seq_a, uint8_t *seq_b, uint8_t *len_a, uint8_t *len_b, unsigned long long *note)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	const unsigned long long gp = (unsigned long long)p0 + p;
	unsigned long long err = NOTE_NONE, cells = 0;
	if (p < n) {
		const bool two = (op == OP_HETERODIMER || op == OP_HETERODIMER_DIAG);
		uint8_t la = 0, lb = 0;
		err = encode_word(words_a + 2ull * p, seq_a + (size_t)p * THERMO_SEQ_STRIDE, &la, gp, 0u);
		if (err == NOTE_NONE && op == OP_HAIRPIN && la == 0) err = (gp << 8) | (1u << 4);
		if (err == NOTE_NONE && two) err = encode_word(words_b + 2ull * p, seq_b + (size_t)p * THERMO_SEQ_STRIDE, &lb, gp, 2u);
		len_a[p] = la;
		len_b[p] = lb;
		cells = (unsigned long long)problem_cells(op, la, two ? lb : la);
	}
	for (int d = 16; d; d >>= 1) {
		cells += __shfl_xor_sync(0xffffffffu, cells, d);
		const unsigned long long o = __shfl_xor_sync(0xffffffffu, err, d);
		err = o < err ? o : err;
	}
	if ((threadIdx.x & 31u) == 0u) {
		if (cells) atomicAdd(note + 1, cells);
		if (err != NOTE_NONE) atomicMin(note, err);
	}
}

__global__ void thermo_fields_kernel(uint32_t n, const float4 *__restrict__ out, float *fields)
{ // {Tm, dH, dS, dG_dp} records -> one array per field: each leaves with one copy into the caller's array
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	const float4 v = out[p];
	fields[p] = v.x;
	fields[(size_t)n + p] = v.y;
	fields[2 * (size_t)n + p] = v.z;
	fields[3 * (size_t)n + p] = v.w;
}

// the strand half of stage_range: logf of the strand concentration of problems [lo, hi) -> *err_p = the first problem with a bad concentration
const char *stage_strands(ThermoState *t, int op, uint32_t lo, uint32_t hi, const float *strand_a, const float *strand_b, uint32_t *err_p)
{
	float *ls = t->h_ls.as<float>();
	float last_strand = -1.0f, last_log = 0.0f;
	for (uint32_t p = lo; p < hi; ++p) {
		float strand = 1.0f;
		*err_p = p;
		if (two_sequences(op)) {
			if (strand_a[p] < 0.0f) return ":strand: m_c_a < 0.0f";
			if (strand_b[p] < 0.0f) return ":strand: m_c_b < 0.0f";
			strand = hetero_strand(strand_a[p], strand_b[p]);
		} else if (needs_strand(op)) {
			if (strand_a[p] < 0.0f) return ":strand: strand_concentration < 0.0f";
			strand = strand_a[p];
		}
		if (op != OP_PM_DUPLEX && op != OP_HAIRPIN && !(strand > 0.0f)) return ":NucCruc::tm_dimer: Invalid strand_concentration";
		if (strand != last_strand) { // the float overload the reference's log() resolves to (nuc_cruc.cpp:2129)
			last_strand = strand;
			last_log = logf(strand);
		}
		ls[p] = last_log;
	}
	*err_p = 0xffffffffu;
	return nullptr;
}

// Large string batches: the caller's text goes to the device as it is (page-locked arrays move at link speed) and is encoded
// there, while the host takes the logarithms of the strand concentrations (the reference's libm).  Same errors, in the same
// order, as the host loop (stage_range) that small batches use.
constexpr uint32_t THERMO_DEVICE_ENCODE_MIN = 4096;
const char *stage_strands_all(ThermoState *t, int op, uint32_t n, const float *strand_a, const float *strand_b, uint32_t *err_p);
int thermo_note_error(pcramp_gpu_ctx *ctx, int op, unsigned long long note, const char *host_msg, uint32_t host_p);

int thermo_stage_strings_device(pcramp_gpu_ctx *ctx, ThermoState *t, int op, uint32_t n, const char *seq_a, const char *seq_b, uint32_t stride,
	const float *strand_a, const float *strand_b)
{
	cudaStream_t st = ctx->stream;
	const bool two = two_sequences(op);
	if (thermo_reserve(ctx, t, n)) return 1;
	CK(t->d_text_a.ensure((size_t)n * stride));
	if (two) CK(t->d_text_b.ensure((size_t)n * stride));
	CK(t->d_note.ensure(16));
	CK(t->h_note.ensure(16));
	CK(cudaStreamSynchronize(st)); // the page-locked staging below may still be the source of an earlier batch's copies
	unsigned long long *h_note = t->h_note.as<unsigned long long>();
	h_note[0] = NOTE_NONE;
	h_note[1] = 0;
	CK(cudaMemcpyAsync(t->d_note.p, h_note, 16, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(t->d_text_a.p, seq_a, (size_t)n * stride, cudaMemcpyHostToDevice, st));
	if (two) CK(cudaMemcpyAsync(t->d_text_b.p, seq_b, (size_t)n * stride, cudaMemcpyHostToDevice, st));
	thermo_encode_kernel<<<grid_for(n, 256), 256, 0, st>>>(op, n, 0u, t->d_text_a.as<char>(), t->d_text_b.as<char>(), stride, t->d_a.as<uint8_t>(),
		t->d_b.as<uint8_t>(), t->d_la.as<uint8_t>(), t->d_lb.as<uint8_t>(), t->d_note.as<unsigned long long>());
	CK(cudaGetLastError());
	// the host's half, while the copies and the encoder run
	uint32_t host_p = 0xffffffffu;
	const char *host_msg = stage_strands_all(t, op, n, strand_a, strand_b, &host_p);
	CK(cudaMemcpyAsync(t->d_ls.p, t->h_ls.p, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(h_note, t->d_note.p, 16, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	t->op = -1;
	if (thermo_note_error(ctx, op, h_note[0], host_msg, host_p)) return 1;
	t->op = op;
	t->n = n;
	t->cells = h_note[1];
	return thermo_order(ctx, t);
}

// every problem with the same concentrations (the usual batch): one validation, one logf, one fill
bool strands_uniform(int op, uint32_t n, const float *strand_a, const float *strand_b)
{
	if (!n) return false;
	const bool two = two_sequences(op);
	if (!two && !needs_strand(op)) return true;
	const float a0 = strand_a[0], b0 = two ? strand_b[0] : 0.0f;
	uint32_t diff = 0;
	for (uint32_t p = 0; p < n; ++p) diff |= (uint32_t)(strand_a[p] != a0);
	if (two)
		for (uint32_t p = 0; p < n; ++p) diff |= (uint32_t)(strand_b[p] != b0);
	return diff == 0;
}

const char *stage_strands_all(ThermoState *t, int op, uint32_t n, const float *strand_a, const float *strand_b, uint32_t *err_p)
{
	if (strands_uniform(op, n, strand_a, strand_b)) {
		const char *e = stage_strands(t, op, 0, 1, strand_a, strand_b, err_p);
		if (e) return e; // problem 0 is the first to fail
		float *ls = t->h_ls.as<float>();
		std::fill(ls + 1, ls + n, ls[0]);
		return nullptr;
	}
	const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
	const uint32_t n_thr = n < 65536u ? 1u : std::min<uint32_t>(8u, hw);
	std::vector<const char *> err(n_thr, nullptr);
	std::vector<uint32_t> where(n_thr, 0xffffffffu);
	if (n_thr == 1) {
		err[0] = stage_strands(t, op, 0, n, strand_a, strand_b, &where[0]);
	} else {
		std::vector<std::thread> pool;
		for (uint32_t k = 0; k < n_thr; ++k) {
			const uint32_t lo = (uint32_t)((uint64_t)n * k / n_thr), hi = (uint32_t)((uint64_t)n * (k + 1) / n_thr);
			pool.emplace_back([&, k, lo, hi]() { err[k] = stage_strands(t, op, lo, hi, strand_a, strand_b, &where[k]); });
		}
		for (std::thread &th : pool) th.join();
	}
	for (uint32_t k = 0; k < n_thr; ++k)
		if (err[k]) { *err_p = where[k]; return err[k]; } // ranges ascend: the first is the smallest
	*err_p = 0xffffffffu;
	return nullptr;
}

int thermo_note_error(pcramp_gpu_ctx *ctx, int op, unsigned long long note, const char *host_msg, uint32_t host_p)
{ // the smallest of the device's and the host's first errors (NOTE_NONE / null: none); 0 = no error
	const unsigned long long host_key = host_msg ? (((unsigned long long)host_p << 8) | (3u << 4)) : NOTE_NONE;
	if (note < host_key) {
		const unsigned stage = (unsigned)(note >> 4) & 15u, detail = (unsigned)note & 15u;
		if (stage == 1u) return fail(ctx, ":NucCruc::align_hairpin: Empty query sequence");
		if (detail == 0u) return fail(ctx, "pcramp_gpu_thermo: sequence longer than 32 bases (Word length)");
		return fail(ctx, (stage == 2u || op != OP_PM_DUPLEX) ? ":set_query: Illegal base" : "Unknown base in tm_pm_duplex");
	}
	if (host_msg) return fail(ctx, host_msg);
	return 0;
}

int thermo_order_range(pcramp_gpu_ctx *ctx, ThermoState *t, cudaStream_t st, DevBuf &tmp, size_t tmp_bytes, uint32_t lo, uint32_t m)
{ // thermo_order for problems [lo, lo + m): a launch order local to the range
	thermo_key_kernel<<<grid_for(m, 256), 256, 0, st>>>(t->op, m, t->d_la.as<uint8_t>() + lo, t->d_lb.as<uint8_t>() + lo, t->d_key[0].as<uint16_t>() + lo,
		t->d_ord[0].as<uint32_t>() + lo);
	CK(cudaGetLastError());
	CK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, t->d_key[0].as<uint16_t>() + lo, t->d_key[1].as<uint16_t>() + lo, t->d_ord[0].as<uint32_t>() + lo,
		t->d_ord[1].as<uint32_t>() + lo, (int)m, 0, 12, st));
	return 0;
}

// One call = one batch, host arrays in and out: the batch is cut into chunks, each on its own stream, so the copies of one
// chunk (text in, fields out) run beside the kernels of the others: the copy engine streams the text in back to back and the last
// kernel ends soon after the last byte arrived.  Each chunk: text -> device, encode, launch order, thermo_kernel, fields, one copy
// per result array the caller wants.  The host takes the logarithms while the first chunk's text moves.  The chunks shrink
// towards the end (a chunk's kernel cannot take less than one problem's latency, whatever its size).
constexpr uint32_t THERMO_PIPELINE_MIN = 65536;
constexpr int THERMO_PIPELINE_CHUNKS = ThermoState::PIPE;

int thermo_batch_pipelined(pcramp_gpu_ctx *ctx, ThermoState *t, int op, uint32_t n, const char *seq_a, const char *seq_b, uint32_t stride,
	const float *strand_a, const float *strand_b, float *tm, float *dH, float *dS, float *dG_dp, bool words = false)
{ // words: seq_a / seq_b are arrays of 16-byte words (stride 16) instead of text
	if (op < 0 || op >= OP_COUNT) return fail(ctx, "pcramp_gpu_thermo: unknown op");
	if (!seq_a || stride == 0) return fail(ctx, "pcramp_gpu_thermo: null sequences");
	const bool two = two_sequences(op);
	if (two && (!seq_b || !strand_b)) return fail(ctx, "pcramp_gpu_thermo: heterodimer ops need seq_b and strand_b");
	if (needs_strand(op) && !strand_a) return fail(ctx, "pcramp_gpu_thermo: null strand concentration");
	for (int k = 0; k < THERMO_PIPELINE_CHUNKS; ++k) {
		if (!t->pipe[k]) CK(cudaStreamCreateWithFlags(&t->pipe[k], cudaStreamNonBlocking));
		if (!t->ev_join[k]) CK(cudaEventCreateWithFlags(&t->ev_join[k], cudaEventDisableTiming));
		if (!t->ev_in[k]) CK(cudaEventCreateWithFlags(&t->ev_in[k], cudaEventDisableTiming));
	}
	cudaStream_t main_st = ctx->stream;
	if (thermo_reserve(ctx, t, n)) return 1;
	CK(t->d_text_a.ensure((size_t)n * stride));
	if (two) CK(t->d_text_b.ensure((size_t)n * stride));
	CK(t->d_note.ensure(16));
	CK(t->h_note.ensure(16));
	CK(t->d_fields.ensure((size_t)n * 16));
	uint32_t bound[THERMO_PIPELINE_CHUNKS + 1];
	{
		static const double upto[THERMO_PIPELINE_CHUNKS] = {0.35, 0.65, 0.85, 1.0};
		bound[0] = 0;
		for (int c = 0; c < THERMO_PIPELINE_CHUNKS; ++c)
			bound[c + 1] = c + 1 == THERMO_PIPELINE_CHUNKS ? n : std::min<uint32_t>(n, (uint32_t)((double)n * upto[c]) / THERMO_BLOCK * THERMO_BLOCK);
	}
	const uint32_t chunk = bound[1]; // the largest
	const bool ordered = op != OP_PM_DUPLEX;
	size_t tmp_bytes = 0;
	if (ordered) {
		for (int k = 0; k < 2; ++k) {
			CK(t->d_key[k].ensure((size_t)n * 2));
			CK(t->d_ord[k].ensure((size_t)n * 4));
		}
		CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, t->d_key[0].as<uint16_t>(), t->d_key[1].as<uint16_t>(), t->d_ord[0].as<uint32_t>(),
			t->d_ord[1].as<uint32_t>(), (int)chunk, 0, 12, main_st));
		for (int k = 0; k < THERMO_PIPELINE_CHUNKS; ++k) CK(t->d_pipe_tmp[k].ensure(tmp_bytes));
	}
	CK(cudaStreamSynchronize(main_st)); // the page-locked staging below may still be the source of an earlier batch's copies
	unsigned long long *h_note = t->h_note.as<unsigned long long>();
	h_note[0] = NOTE_NONE;
	h_note[1] = 0;
	CK(cudaMemcpyAsync(t->d_note.p, h_note, 16, cudaMemcpyHostToDevice, main_st));
	CK(cudaEventRecord(t->ev0, main_st));
	CK(cudaEventRecord(t->ev_fork, main_st));
	for (int k = 0; k < THERMO_PIPELINE_CHUNKS; ++k) CK(cudaStreamWaitEvent(t->pipe[k], t->ev_fork, 0));
	t->op = op;
	t->n = n;
	t->order = nullptr;
	const char *host_msg = nullptr;
	uint32_t host_p = 0xffffffffu;
	float *dst[4] = {tm, dH, dS, dG_dp};
	const bool trace = getenv("PCRAMP_TRACE") != nullptr;
	cudaEvent_t tev[THERMO_PIPELINE_CHUNKS][4] = {};
	if (trace)
		for (auto &row : tev)
			for (cudaEvent_t &e : row) cudaEventCreate(&e);
	int n_chunks = 0;
	bool ls_uniform = false;
	float ls_all = 0.0f;
	for (int c = 0; c < THERMO_PIPELINE_CHUNKS; ++c) {
		const uint32_t lo = bound[c], m = bound[c + 1] - bound[c];
		if (!m) continue;
		++n_chunks;
		cudaStream_t st = t->pipe[c];
		// the text moves on the context's stream, chunk after chunk in order (one queue: the first chunk is complete as early as the
		// link allows); the chunk's own stream takes over from there
		if (trace) cudaEventRecord(tev[c][0], main_st);
		CK(cudaMemcpyAsync(t->d_text_a.as<char>() + (size_t)lo * stride, seq_a + (size_t)lo * stride, (size_t)m * stride, cudaMemcpyHostToDevice, main_st));
		if (two) CK(cudaMemcpyAsync(t->d_text_b.as<char>() + (size_t)lo * stride, seq_b + (size_t)lo * stride, (size_t)m * stride, cudaMemcpyHostToDevice, main_st));
		if (c == 0) { // the host's half, while the first chunk's text moves
			if (strands_uniform(op, n, strand_a, strand_b)) { // one concentration for the whole batch: its logarithm rides in the launch
				host_msg = stage_strands(t, op, 0, 1, strand_a, strand_b, &host_p);
				ls_all = t->h_ls.as<float>()[0];
				ls_uniform = true;
			} else {
				host_msg = stage_strands_all(t, op, n, strand_a, strand_b, &host_p);
				CK(cudaMemcpyAsync(t->d_ls.p, t->h_ls.p, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, main_st));
			}
		}
		CK(cudaEventRecord(t->ev_in[c], main_st));
		CK(cudaStreamWaitEvent(st, t->ev_in[c], 0));
		if (words)
			thermo_encode_words_kernel<<<grid_for(m, 256), 256, 0, st>>>(op, m, lo, (const uint64_t *)(t->d_text_a.as<char>() + (size_t)lo * stride),
				(const uint64_t *)(t->d_text_b.as<char>() + (size_t)lo * stride), t->d_a.as<uint8_t>() + (size_t)lo * THERMO_SEQ_STRIDE,
				t->d_b.as<uint8_t>() + (size_t)lo * THERMO_SEQ_STRIDE, t->d_la.as<uint8_t>() + lo, t->d_lb.as<uint8_t>() + lo,
				t->d_note.as<unsigned long long>());
		else
			thermo_encode_kernel<<<grid_for(m, 256), 256, 0, st>>>(op, m, lo, t->d_text_a.as<char>() + (size_t)lo * stride, t->d_text_b.as<char>() + (size_t)lo * stride,
				stride, t->d_a.as<uint8_t>() + (size_t)lo * THERMO_SEQ_STRIDE, t->d_b.as<uint8_t>() + (size_t)lo * THERMO_SEQ_STRIDE, t->d_la.as<uint8_t>() + lo,
				t->d_lb.as<uint8_t>() + lo, t->d_note.as<unsigned long long>());
		CK(cudaGetLastError());
		const uint32_t *order = nullptr;
		if (ordered) {
			if (thermo_order_range(ctx, t, st, t->d_pipe_tmp[c], tmp_bytes, lo, m)) return 1;
			order = t->d_ord[1].as<uint32_t>() + lo;
		}
		if (trace) cudaEventRecord(tev[c][1], st);
		thermo_kernel<<<grid_for(m, THERMO_BLOCK), THERMO_BLOCK, 0, st>>>(op, m, order, t->d_a.as<uint8_t>() + (size_t)lo * THERMO_SEQ_STRIDE,
			t->d_b.as<uint8_t>() + (size_t)lo * THERMO_SEQ_STRIDE, t->d_la.as<uint8_t>() + lo, t->d_lb.as<uint8_t>() + lo,
			ls_uniform ? nullptr : t->d_ls.as<float>() + lo, t->d_tables.as<Tables>(), t->d_dp.as<DpTable>(), t->d_out.as<float4>() + lo, ls_all);
		CK(cudaGetLastError());
		float *fields = t->d_fields.as<float>() + 4 * (size_t)lo;
		if (trace) cudaEventRecord(tev[c][2], st);
		thermo_fields_kernel<<<grid_for(m, 256), 256, 0, st>>>(m, t->d_out.as<float4>() + lo, fields);
		CK(cudaGetLastError());
		for (int k = 0; k < 4; ++k)
			if (dst[k]) CK(cudaMemcpyAsync(dst[k] + lo, fields + (size_t)k * m, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
		if (trace) cudaEventRecord(tev[c][3], st);
	}
	for (int k = 0; k < THERMO_PIPELINE_CHUNKS; ++k) {
		CK(cudaEventRecord(t->ev_join[k], t->pipe[k]));
		CK(cudaStreamWaitEvent(main_st, t->ev_join[k], 0));
	}
	CK(cudaEventRecord(t->ev1, main_st));
	CK(cudaMemcpyAsync(h_note, t->d_note.p, 16, cudaMemcpyDeviceToHost, main_st));
	CK(cudaStreamSynchronize(main_st));
	t->cells = h_note[1];
	t->stats.n_problems = n;
	t->stats.dp_cells = t->cells;
	t->stats.kernel_launches = n_chunks;
	float ms = 0.0f;
	cudaEventElapsedTime(&ms, t->ev0, t->ev1);
	t->stats.ms_kernel = ms; // the whole pipeline here (copies included): the kernel's own time is a staged run's
	t->op = -1;              // nothing stays staged: the chunks were ordered one by one
	if (trace) {
		for (int c = 0; c < THERMO_PIPELINE_CHUNKS; ++c) {
			float v[4] = {0, 0, 0, 0};
			for (int k = 0; k < 4; ++k) {
				if (bound[c + 1] > bound[c]) cudaEventElapsedTime(&v[k], t->ev0, tev[c][k]);
				cudaEventDestroy(tev[c][k]);
			}
			fprintf(stderr, "[pcramp thermo] chunk %d: copy in from %.3f ms, kernel %.3f .. %.3f ms, results out by %.3f ms (of %.3f)\n", c, v[0], v[1], v[2], v[3], ms);
		}
	}
	return thermo_note_error(ctx, op, h_note[0], host_msg, host_p);
}

int thermo_stage_strings(pcramp_gpu_ctx *ctx, ThermoState *t, int op, uint32_t n, const char *seq_a, const char *seq_b, uint32_t stride,
	const float *strand_a, const float *strand_b)
{
	if (op < 0 || op >= OP_COUNT) return fail(ctx, "pcramp_gpu_thermo: unknown op");
	if (n && (!seq_a || stride == 0)) return fail(ctx, "pcramp_gpu_thermo: null sequences");
	if (n && two_sequences(op) && (!seq_b || !strand_b)) return fail(ctx, "pcramp_gpu_thermo: heterodimer ops need seq_b and strand_b");
	if (n && needs_strand(op) && !strand_a) return fail(ctx, "pcramp_gpu_thermo: null strand concentration");
	if (n >= THERMO_DEVICE_ENCODE_MIN) return thermo_stage_strings_device(ctx, t, op, n, seq_a, seq_b, stride, strand_a, strand_b);
	if (thermo_reserve(ctx, t, n)) return 1;

	// host staging is a plain byte loop: split it over a few threads for large batches
	const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
	const uint32_t n_thr = n < 65536u ? 1u : std::min<uint32_t>(8u, hw);
	std::vector<const char *> err(n_thr, nullptr);
	if (n_thr == 1) {
		err[0] = stage_range(t, op, 0, n, seq_a, seq_b, stride, strand_a, strand_b);
	} else {
		std::vector<std::thread> pool;
		for (uint32_t k = 0; k < n_thr; ++k) {
			const uint32_t lo = (uint32_t)((uint64_t)n * k / n_thr), hi = (uint32_t)((uint64_t)n * (k + 1) / n_thr);
			pool.emplace_back([&, k, lo, hi]() { err[k] = stage_range(t, op, lo, hi, seq_a, seq_b, stride, strand_a, strand_b); });
		}
		for (std::thread &th : pool) th.join();
	}
	for (const char *e : err)
		if (e) return fail(ctx, e);
	return thermo_upload(ctx, t, op, n);
}

// ---- degenerate oligo expansion (Word::begin / Word::next, word.h:525-647; Word::str, :649-666) -----------------
struct Expansion {
	int len = 0;
	int n_letters[32];
	uint8_t letters[32][4];
	uint64_t count = 1;
};

int expand_word(pcramp_gpu_ctx *ctx, const uint64_t w[2], Expansion &e)
{
	W128 x;
	x.hi = w[0];
	x.lo = w[1];
	const int first = w_start(x), last = w_stop(x);
	e.len = 0;
	e.count = 1;
	if (last < first || last < 0) return fail(ctx, "pcramp_gpu_thermo: empty oligo");
	static const uint8_t code_of_bit[4] = {bA, bC, bG, bT}; // nibble bits A=1 C=2 G=4 T=8 (base_table.h:6-29)
	for (int i = first; i <= last; ++i) {
		const uint32_t nib = w_get(x, i);
		if (nib == 0) return fail(ctx, "Unknown base in tm_pm_duplex"); // str() writes '-' for an internal EOS
		int k = 0;
		for (int b = 0; b < 4; ++b)
			if (nib & (1u << b)) e.letters[e.len][k++] = code_of_bit[b];
		e.n_letters[e.len] = k;
		e.count *= (uint64_t)k;
		if (e.count > (1ull << 24)) return fail(ctx, "pcramp_gpu_thermo: oligo degeneracy above 2^24");
		++e.len;
	}
	return 0;
}

// the idx-th expansion (mixed radix, position 0 fastest) as base codes
inline void expansion_at(const Expansion &e, uint64_t idx, uint8_t *dst)
{
	memset(dst, 0, THERMO_SEQ_STRIDE);
	for (int i = 0; i < e.len; ++i) {
		const int k = e.n_letters[i];
		dst[i] = e.letters[i][idx % k];
		idx /= k;
	}
}

int run_and_fetch(pcramp_gpu_ctx *ctx, ThermoState *t, int op, uint32_t n)
{
	if (thermo_upload(ctx, t, op, n)) return 1;
	if (thermo_launch(ctx, t)) return 1;
	return thermo_download(ctx, t);
}

} // namespace

namespace pcr {
namespace nc {
// Internal entry for the other translation units of the library (the optimisation loop): problems already in base
// codes.  codes: n x 32 bytes, the bytes past a sequence's length are what the caller wants the reference's ring
// buffer to hold there (nuccruc.cuh header) -- zero (= A) unless a longer query was loaded before.
int thermo_run_codes(pcramp_gpu_ctx *ctx, int op, uint32_t n, const uint8_t *codes, const uint8_t *len, const float *strand, float salt,
	float *tm_out)
{
	ThermoState *t = nullptr;
	if (thermo_get(ctx, &t)) return 1;
	if (thermo_set_salt(ctx, t, salt)) return 1;
	if (two_sequences(op)) return fail(ctx, "thermo_run_codes: single-sequence ops only");
	// in chunks: the page-locked staging stays at one chunk whatever the batch (growing it is slow: hundreds of ms for 10^7 problems)
	uint64_t launches = 0, cells = 0;
	float ms = 0.0f;
	for (uint32_t lo = 0; lo < n || lo == 0; lo += THERMO_CHUNK) {
		const uint32_t m = std::min<uint32_t>(THERMO_CHUNK, n - lo);
		if (thermo_reserve(ctx, t, m)) return 1;
		if (m) {
			memcpy(t->h_a.p, codes + (size_t)lo * THERMO_SEQ_STRIDE, (size_t)m * THERMO_SEQ_STRIDE);
			memcpy(t->h_la.p, len + lo, m);
			memset(t->h_lb.p, 0, m);
			float *ls = t->h_ls.as<float>();
			float last_strand = -1.0f, last_log = 0.0f;
			for (uint32_t p = 0; p < m; ++p) { // runs of equal concentrations (the expansions of one oligo) share one logf
				if (strand[lo + p] != last_strand) {
					last_strand = strand[lo + p];
					last_log = logf(last_strand);
				}
				ls[p] = last_log;
			}
		}
		if (run_and_fetch(ctx, t, op, m)) return 1;
		const float4 *o = t->h_out.as<float4>();
		for (uint32_t p = 0; p < m; ++p) tm_out[lo + p] = o[p].x;
		launches += t->stats.kernel_launches;
		cells += t->stats.dp_cells;
		ms += t->stats.ms_kernel;
		if (!n) break;
	}
	t->stats.kernel_launches = launches;
	t->stats.dp_cells = cells;
	t->stats.n_problems = n;
	t->stats.ms_kernel = ms;
	return 0;
}
} // namespace nc
} // namespace pcr

extern "C" {

int pcramp_gpu_thermo_stage(pcramp_gpu_ctx *ctx, int op, uint32_t n, const char *seq_a, const char *seq_b, uint32_t stride, float salt,
	const float *strand_a, const float *strand_b)
{
	if (!ctx) return 1;
	CK(cudaSetDevice(ctx->device));
	ThermoState *t = nullptr;
	if (thermo_get(ctx, &t)) return 1;
	if (thermo_set_salt(ctx, t, salt)) return 1;
	return thermo_stage_strings(ctx, t, op, n, seq_a, seq_b, stride, strand_a, strand_b);
}

int pcramp_gpu_thermo_run_staged(pcramp_gpu_ctx *ctx)
{
	if (!ctx) return 1;
	CK(cudaSetDevice(ctx->device));
	if (!ctx->thermo || ctx->thermo->op < 0) return fail(ctx, "pcramp_gpu_thermo_run_staged: nothing staged");
	return thermo_launch(ctx, ctx->thermo);
}

int pcramp_gpu_thermo_fetch(pcramp_gpu_ctx *ctx, float *tm, float *dH, float *dS, float *dG_dp)
{
	if (!ctx) return 1;
	CK(cudaSetDevice(ctx->device));
	ThermoState *t = ctx->thermo;
	if (!t || t->op < 0) return fail(ctx, "pcramp_gpu_thermo_fetch: nothing staged");
	if (t->n >= THERMO_DEVICE_ENCODE_MIN) { // one array per field on the device, one copy per array the caller wants
		const uint32_t n = t->n;
		CK(t->d_fields.ensure((size_t)n * 16));
		thermo_fields_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(n, t->d_out.as<float4>(), t->d_fields.as<float>());
		CK(cudaGetLastError());
		float *dst[4] = {tm, dH, dS, dG_dp};
		for (int k = 0; k < 4; ++k)
			if (dst[k]) CK(cudaMemcpyAsync(dst[k], t->d_fields.as<float>() + (size_t)k * n, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		if (t->stats.kernel_launches) {
			float ms = 0.0f;
			cudaEventElapsedTime(&ms, t->ev0, t->ev1);
			t->stats.ms_kernel = ms;
		}
		return 0;
	}
	if (thermo_download(ctx, t)) return 1;
	const float4 *o = t->h_out.as<float4>();
	for (uint32_t p = 0; p < t->n; ++p) {
		if (tm) tm[p] = o[p].x;
		if (dH) dH[p] = o[p].y;
		if (dS) dS[p] = o[p].z;
		if (dG_dp) dG_dp[p] = o[p].w;
	}
	return 0;
}

int pcramp_gpu_thermo_batch(pcramp_gpu_ctx *ctx, int op, uint32_t n, const char *seq_a, const char *seq_b, uint32_t stride, float salt,
	const float *strand_a, const float *strand_b, float *tm, float *dH, float *dS, float *dG_dp)
{
	if (ctx && n >= THERMO_PIPELINE_MIN) {
		CK(cudaSetDevice(ctx->device));
		ThermoState *t = nullptr;
		if (thermo_get(ctx, &t)) return 1;
		if (thermo_set_salt(ctx, t, salt)) return 1;
		return thermo_batch_pipelined(ctx, t, op, n, seq_a, seq_b, stride, strand_a, strand_b, tm, dH, dS, dG_dp);
	}
	if (pcramp_gpu_thermo_stage(ctx, op, n, seq_a, seq_b, stride, salt, strand_a, strand_b)) return 1;
	if (pcramp_gpu_thermo_run_staged(ctx)) return 1;
	return pcramp_gpu_thermo_fetch(ctx, tm, dH, dS, dG_dp);
}

int pcramp_gpu_thermo_words(pcramp_gpu_ctx *ctx, int op, uint32_t n, const uint64_t *words_a, const uint64_t *words_b, float salt,
	const float *strand_a, const float *strand_b, float *tm, float *dH, float *dS, float *dG_dp)
{
	if (!ctx) return 1;
	if (n && !words_a) return fail(ctx, "pcramp_gpu_thermo_words: null argument");
	if (n && two_sequences(op) && !words_b) return fail(ctx, "pcramp_gpu_thermo: heterodimer ops need seq_b and strand_b");
	if (n >= THERMO_PIPELINE_MIN) {
		CK(cudaSetDevice(ctx->device));
		ThermoState *t = nullptr;
		if (thermo_get(ctx, &t)) return 1;
		if (thermo_set_salt(ctx, t, salt)) return 1;
		return thermo_batch_pipelined(ctx, t, op, n, (const char *)words_a, (const char *)words_b, 16u, strand_a, strand_b, tm, dH, dS, dG_dp, true);
	}
	// small batches: Word::str() on the host, then the text path
	static const char letter[16] = {'-', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N'}; // bits_to_base (base_table.h)
	const uint32_t stride = 33;
	std::vector<char> ta((size_t)n * stride, 0), tb(two_sequences(op) ? (size_t)n * stride : 0, 0);
	auto text = [&](const uint64_t *w, char *dst) {
		W128 x;
		x.hi = w[0];
		x.lo = w[1];
		const int first = w_start(x), last = w_stop(x);
		int k = 0;
		for (int i = first; i >= 0 && i <= last; ++i) dst[k++] = letter[w_get(x, i)];
	};
	for (uint32_t p = 0; p < n; ++p) {
		text(words_a + 2ull * p, ta.data() + (size_t)p * stride);
		if (!tb.empty()) text(words_b + 2ull * p, tb.data() + (size_t)p * stride);
	}
	return pcramp_gpu_thermo_batch(ctx, op, n, ta.data(), tb.empty() ? nullptr : tb.data(), stride, salt, strand_a, strand_b, tm, dH, dS, dG_dp);
}

int pcramp_gpu_get_thermo_stats(pcramp_gpu_ctx *ctx, pcramp_gpu_thermo_stats *out)
{
	if (!ctx || !out) return 1;
	if (ctx->thermo) *out = ctx->thermo->stats;
	else memset(out, 0, sizeof(*out));
	return 0;
}

// PCR::is_valid for a batch.  The reference stops at the first failing expansion / test; the verdict is the AND over all
// of them, so the batch runs in three rounds -- duplex Tm of every expansion, then hairpins of the oligos still
// alive, then homodimers -- each round one kernel launch.
int pcramp_gpu_is_valid(pcramp_gpu_ctx *ctx, const uint64_t *words, uint32_t n, float salt, float primer_strand, float tm_min, float tm_max,
	float max_hairpin, float max_dimer, int check_homo_dimer, int fast_alignment, uint8_t *valid)
{
	if (!ctx) return 1;
	if (n && (!words || !valid)) return fail(ctx, "pcramp_gpu_is_valid: null argument");
	CK(cudaSetDevice(ctx->device));
	ThermoState *t = nullptr;
	if (thermo_get(ctx, &t)) return 1;
	if (thermo_set_salt(ctx, t, salt)) return 1;
	std::vector<Expansion> ex(n);
	std::vector<float> log_strand(n);
	uint64_t total = 0;
	for (uint32_t i = 0; i < n; ++i) {
		if (expand_word(ctx, words + 2 * (size_t)i, ex[i])) return 1;
		const double degen = (double)ex[i].count; // Word::degeneracy (word.h:97-138)
		const float strand = (float)((double)primer_strand / degen); // valid_pcr.cpp:13 -> NucCruc::strand(const float &)
		if (strand < 0.0f) return fail(ctx, ":strand: strand_concentration < 0.0f");
		if (check_homo_dimer && !(strand > 0.0f)) return fail(ctx, ":NucCruc::tm_dimer: Invalid strand_concentration");
		log_strand[i] = logf(strand);
		total += ex[i].count;
		valid[i] = 1;
	}
	if (total > 0xffffffffull) return fail(ctx, "pcramp_gpu_is_valid: too many expansions in one batch");
	uint64_t launches = 0, cells = 0, problems = 0;
	float ms = 0.0f;
	const int ops[3] = {OP_PM_DUPLEX, OP_HAIRPIN, fast_alignment ? OP_HOMODIMER_DIAG : OP_HOMODIMER};
	std::vector<uint32_t> owner;
	for (int round = 0; round < (check_homo_dimer ? 3 : 2); ++round) {
		uint64_t m = 0;
		for (uint32_t i = 0; i < n; ++i)
			if (valid[i]) m += ex[i].count;
		if (!m) break;
		if (thermo_reserve(ctx, t, (uint32_t)m)) return 1;
		owner.resize(m);
		uint8_t *ha = t->h_a.as<uint8_t>(), *la = t->h_la.as<uint8_t>(), *lb = t->h_lb.as<uint8_t>();
		float *ls = t->h_ls.as<float>();
		uint32_t p = 0;
		for (uint32_t i = 0; i < n; ++i) {
			if (!valid[i]) continue;
			for (uint64_t k = 0; k < ex[i].count; ++k, ++p) {
				expansion_at(ex[i], k, ha + (size_t)p * THERMO_SEQ_STRIDE);
				la[p] = (uint8_t)ex[i].len;
				lb[p] = 0;
				ls[p] = log_strand[i];
				owner[p] = i;
			}
		}
		if (run_and_fetch(ctx, t, ops[round], (uint32_t)m)) return 1;
		launches += t->stats.kernel_launches;
		cells += t->stats.dp_cells;
		problems += m;
		ms += t->stats.ms_kernel;
		const float4 *o = t->h_out.as<float4>();
		for (uint32_t q = 0; q < (uint32_t)m; ++q) {
			const float tm = o[q].x;
			bool ok;
			if (round == 0) ok = !((tm < tm_min) || (tm > tm_max)); // valid_pcr.cpp:20
			else if (round == 1) ok = !(tm > max_hairpin);           // :28
			else ok = !(tm > max_dimer);                             // :37
			if (!ok) valid[owner[q]] = 0;
		}
	}
	t->stats.kernel_launches = launches;
	t->stats.dp_cells = cells;
	t->stats.n_problems = problems;
	t->stats.ms_kernel = ms;
	return 0;
}

namespace {
// heterodimer problems for every expansion pair of (a x b); appends to the staging buffers starting at slot p
void stage_pair_products(ThermoState *t, const Expansion &a, const Expansion &b, float log_strand, uint32_t owner_id, std::vector<uint32_t> &owner,
	uint32_t &p)
{
	uint8_t *ha = t->h_a.as<uint8_t>(), *hb = t->h_b.as<uint8_t>(), *la = t->h_la.as<uint8_t>(), *lb = t->h_lb.as<uint8_t>();
	float *ls = t->h_ls.as<float>();
	for (uint64_t i = 0; i < a.count; ++i)
		for (uint64_t j = 0; j < b.count; ++j, ++p) {
			expansion_at(a, i, ha + (size_t)p * THERMO_SEQ_STRIDE);
			expansion_at(b, j, hb + (size_t)p * THERMO_SEQ_STRIDE);
			la[p] = (uint8_t)a.len;
			lb[p] = (uint8_t)b.len;
			ls[p] = log_strand;
			owner[p] = owner_id;
		}
}
} // namespace

int pcramp_gpu_max_dimer_tm(pcramp_gpu_ctx *ctx, const uint64_t *f, const uint64_t *r, uint32_t n_pairs, float salt, float primer_strand,
	int fast_alignment, float *tm)
{
	if (!ctx) return 1;
	if (n_pairs && (!f || !r || !tm)) return fail(ctx, "pcramp_gpu_max_dimer_tm: null argument");
	CK(cudaSetDevice(ctx->device));
	ThermoState *t = nullptr;
	if (thermo_get(ctx, &t)) return 1;
	if (thermo_set_salt(ctx, t, salt)) return 1;
	std::vector<Expansion> ef(n_pairs), er(n_pairs);
	uint64_t total = 0;
	for (uint32_t i = 0; i < n_pairs; ++i) {
		if (expand_word(ctx, f + 2 * (size_t)i, ef[i]) || expand_word(ctx, r + 2 * (size_t)i, er[i])) return 1;
		total += ef[i].count * er[i].count;
	}
	if (total > 0xffffffffull) return fail(ctx, "pcramp_gpu_max_dimer_tm: too many expansions in one batch");
	if (thermo_reserve(ctx, t, (uint32_t)total)) return 1;
	std::vector<uint32_t> owner(total);
	uint32_t p = 0;
	for (uint32_t i = 0; i < n_pairs; ++i) {
		// pcr_assay.cpp:244: strand(primer_strand / degen_f, primer_strand / degen_r), each narrowed to float
		const float ca = (float)((double)primer_strand / (double)ef[i].count), cb = (float)((double)primer_strand / (double)er[i].count);
		if (ca < 0.0f || cb < 0.0f) return fail(ctx, ":strand: m_c_a < 0.0f");
		const float strand = hetero_strand(ca, cb);
		if (!(strand > 0.0f)) return fail(ctx, ":NucCruc::tm_dimer: Invalid strand_concentration");
		stage_pair_products(t, ef[i], er[i], logf(strand), i, owner, p);
		tm[i] = 0.0f;
	}
	if (run_and_fetch(ctx, t, fast_alignment ? OP_HETERODIMER_DIAG : OP_HETERODIMER, (uint32_t)total)) return 1;
	const float4 *o = t->h_out.as<float4>();
	for (uint32_t q = 0; q < (uint32_t)total; ++q) tm[owner[q]] = std::max(tm[owner[q]], o[q].x);
	return 0;
}

int pcramp_gpu_multiplex_compatible(pcramp_gpu_ctx *ctx, const uint64_t *f, const uint64_t *r, uint32_t n_pairs, const uint64_t *pool_f,
	const uint64_t *pool_r, uint32_t n_pool, float salt, float primer_strand, float max_dimer, int fast_alignment, uint8_t *ok)
{
	if (!ctx) return 1;
	if (n_pairs && (!f || !r || !ok)) return fail(ctx, "pcramp_gpu_multiplex_compatible: null argument");
	if (n_pool && (!pool_f || !pool_r)) return fail(ctx, "pcramp_gpu_multiplex_compatible: null pool");
	CK(cudaSetDevice(ctx->device));
	ThermoState *t = nullptr;
	if (thermo_get(ctx, &t)) return 1;
	if (thermo_set_salt(ctx, t, salt)) return 1;
	for (uint32_t i = 0; i < n_pairs; ++i) ok[i] = 1;
	if (!n_pairs || !n_pool) return 0;
	if (primer_strand < 0.0f) return fail(ctx, ":strand: strand_concentration < 0.0f");
	if (!(primer_strand > 0.0f)) return fail(ctx, ":NucCruc::tm_dimer: Invalid strand_concentration");
	const float log_strand = logf(primer_strand); // pcr_assay.cpp:819-821: no degeneracy correction
	std::vector<Expansion> trial(2 * (size_t)n_pairs), pool(2 * (size_t)n_pool);
	for (uint32_t i = 0; i < n_pairs; ++i)
		if (expand_word(ctx, f + 2 * (size_t)i, trial[2 * i]) || expand_word(ctx, r + 2 * (size_t)i, trial[2 * i + 1])) return 1;
	for (uint32_t i = 0; i < n_pool; ++i)
		if (expand_word(ctx, pool_f + 2 * (size_t)i, pool[2 * i]) || expand_word(ctx, pool_r + 2 * (size_t)i, pool[2 * i + 1])) return 1;
	uint64_t pool_count = 0, total = 0;
	for (const Expansion &e : pool) pool_count += e.count;
	for (const Expansion &e : trial) total += e.count * pool_count;
	// groups of trials of about THERMO_CHUNK problems: the page-locked staging stays at one group whatever the pool has grown to
	std::vector<uint64_t> first((size_t)n_pairs + 1, 0);
	for (uint32_t i = 0; i < n_pairs; ++i) first[i + 1] = first[i] + (trial[2 * i].count + trial[2 * i + 1].count) * pool_count;
	(void)total;
	std::vector<uint32_t> owner;
	uint64_t launches = 0, cells = 0;
	float ms = 0.0f;
	for (uint32_t g0 = 0; g0 < n_pairs;) {
		uint32_t g1 = g0 + 1;
		while (g1 < n_pairs && first[g1 + 1] - first[g0] <= THERMO_CHUNK) ++g1;
		const uint64_t m = first[g1] - first[g0];
		if (m > 0xffffffffull) return fail(ctx, "pcramp_gpu_multiplex_compatible: too many expansions of one trial against the pool");
		if (thermo_reserve(ctx, t, (uint32_t)m)) return 1;
		owner.resize(m);
		auto stage = [&](uint32_t lo, uint32_t hi) { // every trial's problems start at a known slot
			for (uint32_t i = lo; i < hi; ++i) {
				uint32_t p = (uint32_t)(first[i] - first[g0]);
				for (const Expansion &q : pool) // main.cpp:748-752: pool_assay.multiplex_compatible(melt, opt, trial) -> the pool oligo is the query
					for (int so = 0; so < 2; ++so) stage_pair_products(t, q, trial[2 * i + so], log_strand, i, owner, p);
			}
		};
		const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
		const uint32_t n_thr = m < 65536ull ? 1u : std::min<uint32_t>(std::min<uint32_t>(8u, hw), g1 - g0);
		if (n_thr <= 1) {
			stage(g0, g1);
		} else { // equal shares of the problems, cut at trial boundaries
			std::vector<std::thread> workers;
			uint32_t lo = g0;
			for (uint32_t k = 0; k < n_thr; ++k) {
				uint32_t hi = lo;
				const uint64_t want = first[g0] + m * (k + 1) / n_thr;
				while (hi < g1 && (first[hi + 1] <= want || k + 1 == n_thr)) ++hi;
				if (hi > lo) workers.emplace_back(stage, lo, hi);
				lo = hi;
			}
			for (std::thread &th : workers) th.join();
		}
		if (run_and_fetch(ctx, t, fast_alignment ? OP_HETERODIMER_DIAG : OP_HETERODIMER, (uint32_t)m)) return 1;
		const float4 *o = t->h_out.as<float4>();
		for (uint32_t q = 0; q < (uint32_t)m; ++q)
			if (o[q].x >= max_dimer) ok[owner[q]] = 0; // pcr_assay.cpp:842
		launches += t->stats.kernel_launches;
		cells += t->stats.dp_cells;
		ms += t->stats.ms_kernel;
		g0 = g1;
	}
	t->stats.kernel_launches = launches;
	t->stats.dp_cells = cells;
	t->stats.n_problems = first[n_pairs];
	t->stats.ms_kernel = ms;
	return 0;
}

} // extern "C"

#include "random_assay.cuh" // candidate generation: PCR::random_assay, one GPU thread per seed stream
