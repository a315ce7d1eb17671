// fasta.cuh -- FASTA text -> a device-resident sequence collection (included at the end of pcramp_gpu.cu).
//
// Replaces parse_fasta (parse_fasta.cpp:9-89) + Sequence::operator=(deque<char>) (sequence.cpp:43-71; base_to_bits,
// base_table.h:30-76) + Sequence::defline / extract_weight (sequence.h:190-200, sequence.cpp:332-493) for already-inflated
// text (zlib stays on the host).  Division of labour:
//
//   host    one memchr-speed walk over the text that reproduces the reference's reader: gzgets chunks (at most 2047 bytes,
//           ending after a newline), a chunk holding '>' ANYWHERE is a defline chunk and closes the record before it, the
//           defline is the chunk up to its first CR / LF; weights and the ignore list are evaluated on the defline; the
//           final record is closed without the "non-empty" test the reference applies between records.  A record's
//           residues are the non-space bytes of one contiguous span of the text.
//   device  fasta_count_kernel: non-space bytes per (record, 4 KB block) work item -> record lengths (the length window
//           needs them before anything is packed); fasta_pack_kernel: every residue character -> its 4-bit IUPAC set,
//           white space dropped, two bases per byte at the record's place in the collection's nibble array, plus the EOS
//           ('-') positions, the "some base is degenerate" flag and the first illegal symbol.  HBM-bound: the text is read
//           twice (1 B per character each) and 0.5 B per base is written.
// From there on the collection is what pcramp_gpu_upload_sequences would have produced (same upload_finish).
#pragma once
#include "ctx.cuh"

#include <cub/device/device_scan.cuh>

#include <chrono>
#include <cstdlib>
#include <string>
#include <vector>

namespace pcr {
namespace fasta {

constexpr uint32_t BLOCK_BYTES = 4096; // one warp per work item: 128 bytes per lane
constexpr int READ_CHUNK = 2047;       // gzgets(fin, buffer, 2048)

struct Record {
	uint32_t file;
	uint64_t def_off;   // in the file's text
	uint32_t def_len;
	uint64_t begin, end; // residue span, offsets into the concatenated device text
	float weight;
	bool ignored;
};

struct Item { // a record's share of one aligned 4 KB block of the text
	uint64_t begin, end;
	uint32_t rec;
};

inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); } // isspace() in the C locale

// Sequence::extract_weight (sequence.cpp:332-493): the first "[w=<number>]" (spaces allowed around w, = and the number; a
// '[' restarts the match); the number is what atof makes of the characters from the first value character on
inline float extract_weight(const std::string &d)
{
	enum { NO_MATCH, LEFT, WEIGHT, EQUAL, VALUE, RIGHT } st = NO_MATCH;
	auto value_char = [](char c) { return c == '-' || c == '+' || c == '.' || c == 'e' || (c >= '0' && c <= '9'); };
	size_t start = std::string::npos, stop = std::string::npos;
	auto result = [&]() {
		const std::string v = (stop == std::string::npos) ? d.substr(start) : d.substr(start, stop - start + 1);
		return (float)atof(v.c_str());
	};
	for (size_t i = 0; i < d.size(); ++i) {
		const char c = d[i];
		switch (st) {
		case NO_MATCH:
			if (c == '[') st = LEFT;
			break;
		case LEFT:
			if (c == 'W' || c == 'w') st = WEIGHT;
			else if (c == ' ' || c == '\t' || c == '[') {}
			else st = NO_MATCH;
			break;
		case WEIGHT:
			if (c == '=') st = EQUAL;
			else if (c == ' ' || c == '\t') {}
			else if (c == '[') st = LEFT;
			else st = NO_MATCH;
			break;
		case EQUAL:
			if (c == ' ' || c == '\t') {}
			else if (value_char(c)) { start = i; st = VALUE; }
			else if (c == '[') st = LEFT;
			else st = NO_MATCH;
			break;
		case VALUE:
			if (c == ' ' || c == '\t') st = RIGHT;
			else if (c == ']') return result();
			else if (value_char(c)) stop = i;
			else if (c == '[') st = LEFT;
			else st = NO_MATCH;
			break;
		case RIGHT:
			if (c == ' ' || c == '\t') {}
			else if (c == ']') return result();
			else if (c == '[') st = LEFT;
			else st = NO_MATCH;
			break;
		}
	}
	return 1.0f; // DEFAULT_SCORE_WEIGHT
}

// the reader of parse_fasta over one file; `base` = the file's offset in the concatenated text.
// gzgets hands the reference chunks of at most READ_CHUNK bytes that end after a newline; only the chunks holding a '>' matter
// (they are deflines and close the record before them), so the walk jumps from '>' to '>' (one memchr sweep over the text) and
// reconstructs the chunk around each: chunks of a line start at the line's first byte and every READ_CHUNK bytes after it.
inline void split_records(const char *text, uint64_t n, uint32_t file, uint64_t base, std::vector<Record> &out)
{
	bool have_def = false;
	uint64_t def_off = 0, span_begin = 0;
	uint32_t def_len = 0;
	auto any_data = [&](uint64_t b, uint64_t e) { // "!seq.empty()": a non-space byte in the data chunks
		for (uint64_t k = b; k < e; ++k)
			if (!is_space((unsigned char)text[k])) return true;
		return false;
	};
	auto close = [&](uint64_t span_end, bool need_data) {
		if (need_data && !any_data(span_begin, span_end)) return; // the record after the last defline is closed without that test
		Record r;
		r.file = file;
		r.def_off = def_off;
		r.def_len = have_def ? def_len : 0u;
		r.begin = base + span_begin;
		r.end = base + span_end;
		r.weight = 1.0f;
		r.ignored = false;
		out.push_back(r);
	};
	uint64_t pos = 0; // always the first byte of a chunk
	while (pos < n) {
		const char *gt = (const char *)memchr(text + pos, '>', n - pos);
		if (!gt) break;
		const uint64_t g = (uint64_t)(gt - text);
		// the line holding g starts after the last newline in [pos, g)
		uint64_t line = pos;
		if (g > pos) {
			const char *nl = (const char *)memrchr(text + pos, '\n', g - pos);
			if (nl) line = (uint64_t)(nl - text) + 1;
		}
		const uint64_t chunk = line + ((g - line) / READ_CHUNK) * READ_CHUNK;
		const uint64_t lim = std::min<uint64_t>(n, chunk + READ_CHUNK);
		const char *nl = (const char *)memchr(text + chunk, '\n', lim - chunk);
		const uint64_t end = nl ? (uint64_t)(nl - text) + 1 : lim;
		close(chunk, true);
		uint64_t e = chunk; // the defline: the chunk up to its first CR / LF
		while (e < end && text[e] != '\n' && text[e] != '\r') ++e;
		have_def = true;
		def_off = chunk;
		def_len = (uint32_t)(e - chunk);
		span_begin = end;
		pos = end;
	}
	close(n, false);
}

// base_to_bits (base_table.h:30-76); 0xFF = the symbol the reference throws on, 0xFE = white space
__device__ __forceinline__ uint32_t base_bits(uint32_t c)
{
	switch (c | 0x20u) { // letters: fold to lower case ('-' and white space are not letters: handled first)
	case 'a': return 1u;
	case 'c': return 2u;
	case 'g': return 4u;
	case 't': case 'u': return 8u;
	case 'm': return 3u;
	case 'r': return 5u;
	case 's': return 6u;
	case 'v': return 7u;
	case 'w': return 9u;
	case 'y': return 10u;
	case 'h': return 11u;
	case 'k': return 12u;
	case 'd': return 13u;
	case 'b': return 14u;
	case 'n': case 'i': case 'x': return 15u;
	default: return 0xFFu;
	}
}
__device__ __forceinline__ uint32_t classify(uint32_t c)
{
	if (c == ' ' || (c >= 9u && c <= 13u)) return 0xFEu;
	if (c == '-') return 0u;
	if ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z')) return base_bits(c);
	return 0xFFu;
}

// bytes of w that are white space (' ' or 9..13), one 0xFF per such byte
__device__ __forceinline__ uint32_t space_mask4(uint32_t w)
{
	return __vcmpeq4(w, 0x20202020u) | (__vcmpgeu4(w, 0x09090909u) & __vcmpleu4(w, 0x0D0D0D0Du));
}
// 0xFF for the bytes of the 4-byte word at text offset `at` that lie inside [begin, end)
__device__ __forceinline__ uint32_t inside_mask4(uint64_t at, uint64_t begin, uint64_t end)
{
	if (at >= begin && at + 4 <= end) return 0xFFFFFFFFu;
	uint32_t m = 0u;
	#pragma unroll
	for (int k = 0; k < 4; ++k)
		if (at + k >= begin && at + k < end) m |= 255u << (8 * k);
	return m;
}

__device__ __forceinline__ uint32_t warp_excl_sum(uint32_t v, uint32_t lane, uint32_t &total)
{
	uint32_t incl = v;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if ((int)lane >= o) incl += t;
	}
	total = __shfl_sync(0xffffffffu, incl, 31);
	return incl - v;
}

// residues (non-space bytes) per work item; one warp per item, lane l owns bytes [128 l, 128 l + 128) of the aligned block
__global__ void __launch_bounds__(256) fasta_count_kernel(const uint8_t *__restrict__ text, const Item *__restrict__ items, uint32_t n_items,
	unsigned long long *count)
{
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (it >= n_items) return;
	const Item w = items[it];
	const uint64_t blk = w.begin & ~(uint64_t)(BLOCK_BYTES - 1u);
	uint32_t c = 0;
	#pragma unroll
	for (int v = 0; v < 8; ++v) {
		const uint64_t at = blk + 128ull * lane + 16ull * v;
		if (at + 16 <= w.begin || at >= w.end) continue;
		const uint4 q = *(const uint4 *)(text + at); // the text buffer is padded to a multiple of BLOCK_BYTES
		const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
		#pragma unroll
		for (int k = 0; k < 4; ++k) c += (uint32_t)__popc(~space_mask4(wd[k]) & inside_mask4(at + 4u * k, w.begin, w.end)) >> 3;
	}
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if (lane == 0u) count[it] = c;
}

struct PackOut {
	uint8_t *raw;                 // the collection's nibble array (zeroed)
	const uint64_t *rec_raw_off;  // per record: byte offset of its sequence's first nibble, ~0 = record not kept
	const uint32_t *rec_nib_base; // per record: residue index of its first base inside that sequence (0 unless records are grouped)
	unsigned long long *bad;      // smallest text offset of an illegal symbol (~0 = none)
	unsigned int *flags;          // bit 0: some base is degenerate
	uint2 *eos;                   // (record, residue index) of every '-'
	unsigned int *n_eos;
	uint32_t eos_cap;
};

// class table entry of a text byte: nibble[3:0] | residue [4] | degenerate [5] | EOS '-' [6] | unknown symbol [7]; white space = 0
__device__ __forceinline__ uint32_t class_entry(uint32_t c)
{
	const uint32_t b = classify(c);
	if (b == 0xFEu) return 0u;
	if (b == 0xFFu) return 0x80u;
	return b | 0x10u | ((b & (b - 1u)) ? 0x20u : 0u) | (b == 0u ? 0x40u : 0u);
}

// item_off[it] = residues of the record before this item.  One warp per item, lane l owns bytes [128 l, 128 l + 128) of the block.
// Pass 1 counts the lane's residues (SIMD byte compares), a warp scan gives the lane its first residue index; pass 2 walks the
// same 128 bytes again (L1 hits) in a rolled loop of 16-byte vectors -- table look-up per character, nibbles OR-ed into a 32-bit
// word that is stored when full (plain store when this lane produced all 8 nibbles, atomicOr for the shared words at the lane's
// ends).  '-' and unknown symbols only raise a flag in the hot loop; their positions come from a rare third walk.
__global__ void __launch_bounds__(256) fasta_pack_kernel(const uint8_t *__restrict__ text, const Item *__restrict__ items, uint32_t n_items,
	const uint32_t *__restrict__ item_off, PackOut o)
{
	__shared__ uint8_t s_cls[256];
	s_cls[threadIdx.x & 255u] = (uint8_t)class_entry(threadIdx.x & 255u);
	__syncthreads();
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (it >= n_items) return;
	const Item w = items[it];
	const uint64_t roff = o.rec_raw_off[w.rec];
	if (roff == ~0ull) return;
	const uint64_t blk = w.begin & ~(uint64_t)(BLOCK_BYTES - 1u);
	const uint64_t mine = blk + 128ull * lane;
	uint32_t c = 0;
	#pragma unroll
	for (int v = 0; v < 8; ++v) {
		const uint64_t at = mine + 16ull * v;
		if (at + 16 <= w.begin || at >= w.end) continue;
		const uint4 x = *(const uint4 *)(text + at);
		const uint32_t wd[4] = {x.x, x.y, x.z, x.w};
		#pragma unroll
		for (int k = 0; k < 4; ++k) c += (uint32_t)__popc(~space_mask4(wd[k]) & inside_mask4(at + 4u * k, w.begin, w.end)) >> 3;
	}
	uint32_t total;
	const uint64_t r0 = (uint64_t)o.rec_nib_base[w.rec] + item_off[it] + warp_excl_sum(c, lane, total); // residue index of this lane's first base
	uint32_t *out32 = (uint32_t *)(o.raw + roff); // records start 16-byte aligned
	uint64_t r = r0;
	uint32_t acc = 0, have = 0, flags = 0;
	#pragma unroll 1
	for (int v = 0; v < 8; ++v) {
		const uint64_t at = mine + 16ull * v;
		if (at + 16 <= w.begin || at >= w.end) continue;
		const uint4 x = *(const uint4 *)(text + at);
		const uint32_t wd[4] = {x.x, x.y, x.z, x.w};
		#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const uint32_t y = wd[k] & inside_mask4(at + 4u * k, w.begin, w.end); // bytes outside the item become NUL ...
			#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const uint32_t ch = (y >> (8 * j)) & 255u;
				const uint32_t e = ch ? s_cls[ch] : 0u;                 // ... which counts as nothing at all
				flags |= e;
				const uint32_t is = (e >> 4) & 1u;
				// residue r: byte r >> 1, high nibble for even r; little-endian word r >> 3
				acc |= (is ? (e & 15u) : 0u) << ((((uint32_t)r & 7u) << 2) ^ 4u);
				have += is;
				r += is;
				if (is && ((uint32_t)r & 7u) == 0u) {
					if (have == 8u) out32[(r - 1u) >> 3] = acc;
					else atomicOr(out32 + ((r - 1u) >> 3), acc);
					acc = 0u;
					have = 0u;
				}
			}
		}
	}
	if (have) atomicOr(out32 + (r >> 3), acc);
	if (flags & 0xC0u) { // a '-' or an unknown symbol among this lane's bytes: find where
		uint64_t rr = r0;
		for (uint64_t p = (mine > w.begin ? mine : w.begin); p < mine + 128ull && p < w.end; ++p) {
			const uint32_t e = s_cls[text[p]];
			if (e & 0x80u) atomicMin(o.bad, (unsigned long long)p);
			if (e & 0x40u) {
				const unsigned int k = atomicAdd(o.n_eos, 1u);
				if (k < o.eos_cap) o.eos[k] = make_uint2(w.rec, (uint32_t)rr);
			}
			rr += (e >> 4) & 1u;
		}
	}
	if (__any_sync(0xffffffffu, (flags & 0x20u) != 0u) && lane == 0u) atomicOr(o.flags, 1u);
}

__global__ void fasta_lengths_kernel(const unsigned long long *__restrict__ excl, const unsigned long long *__restrict__ count,
	const uint32_t *__restrict__ rec_first, uint32_t n_rec, uint32_t n_items, unsigned long long *len, uint32_t *item_off)
{ // rec_first[r] = first item of record r (rec_first[n_rec] = n_items); items of a record are consecutive
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_rec) return;
	const uint32_t a = rec_first[r], b = rec_first[r + 1];
	const unsigned long long base = a < n_items ? excl[a] : 0ull;
	const unsigned long long end = b < n_items ? excl[b] : (n_items ? excl[n_items - 1] + count[n_items - 1] : 0ull);
	len[r] = (a < b) ? end - base : 0ull;
	for (uint32_t i = a; i < b; ++i) item_off[i] = (uint32_t)(excl[i] - base); // records longer than 2^32 - 1 bases are refused by the caller
}

struct Table { // the record table of the last upload_fasta, for pcramp_gpu_fasta_records
	float ms_count = 0.0f, ms_pack = 0.0f; // CUDA-event times of the two kernels of the last upload
	uint64_t text_bytes = 0, n_bases = 0;
	std::vector<uint32_t> file, def_len, length;
	std::vector<uint64_t> def_off;
	std::vector<float> weight;
};

} // namespace fasta
} // namespace pcr

struct pcramp_gpu_fasta : pcr::fasta::Table {};

extern "C" {

// file_group == NULL: parse_fasta, one sequence per kept record.  Otherwise append_fasta_group (parse_fasta.cpp:91-169) driven as
// main.cpp:296-341 does: the files of a group (file_group[f], non-decreasing) form ONE sequence -- the kept records in order, num_pad
// EOS between them (Sequence::pad, sequence.h:258-270) -- and a group that keeps nothing gives no sequence.
static int upload_fasta_impl(pcramp_gpu_ctx *ctx, int kind, uint32_t n_files, const char *const *text, const uint64_t *bytes, uint64_t min_length,
	uint64_t max_length, uint32_t n_ignore, const char *const *ignore, uint32_t *n_records, const uint32_t *file_group, uint32_t num_pad)
{
	using namespace pcr::fasta;
	if (check_kind(ctx, kind) || text_change(ctx, "pcramp_gpu_upload_fasta")) return 1;
	for (uint32_t f = 1; file_group && f < n_files; ++f)
		if (file_group[f] < file_group[f - 1]) return fail(ctx, "pcramp_gpu_upload_fasta_groups: the files of a group must be consecutive");
	if (n_files && (!text || !bytes)) return fail(ctx, "pcramp_gpu_upload_fasta: null argument");
	CK(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const bool dbg = getenv("PCRAMP_FASTA_TIMING") != nullptr;
	auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const double t0 = now();
	// ---- host: records ---------------------------------------------------------------------------------------
	std::vector<Record> recs;
	std::vector<uint64_t> base(n_files + 1, 0);
	for (uint32_t f = 0; f < n_files; ++f) base[f + 1] = (base[f] + bytes[f] + 15u) & ~15ull; // files start 16-byte aligned
	for (uint32_t f = 0; f < n_files; ++f) split_records(text[f], bytes[f], f, base[f], recs);
	for (Record &r : recs) {
		const std::string def(text[r.file] + r.def_off, r.def_len);
		r.weight = extract_weight(def);
		if (n_ignore) { // ignore_record (parse_fasta.cpp:171-188): the ignore strings are already lower case
			std::string low(def);
			for (char &c : low) c = (char)tolower((unsigned char)c);
			for (uint32_t k = 0; k < n_ignore && !r.ignored; ++k) r.ignored = low.find(ignore[k]) != std::string::npos;
		}
	}
	const uint32_t n_rec = (uint32_t)recs.size();
	// ---- work items ------------------------------------------------------------------------------------------
	std::vector<Item> items;
	std::vector<uint32_t> rec_first(n_rec + 1, 0);
	for (uint32_t r = 0; r < n_rec; ++r) {
		rec_first[r] = (uint32_t)items.size();
		for (uint64_t b = recs[r].begin & ~(uint64_t)(BLOCK_BYTES - 1u); b < recs[r].end; b += BLOCK_BYTES) {
			Item it;
			it.begin = std::max<uint64_t>(b, recs[r].begin);
			it.end = std::min<uint64_t>(b + BLOCK_BYTES, recs[r].end);
			it.rec = r;
			items.push_back(it);
		}
	}
	rec_first[n_rec] = (uint32_t)items.size();
	if (items.size() >= (1ull << 31)) return fail(ctx, "pcramp_gpu_upload_fasta: text too large for one call");
	const uint32_t n_items = (uint32_t)items.size();
	const double t1 = now();
	// ---- text to HBM -----------------------------------------------------------------------------------------
	const uint64_t text_bytes = (base[n_files] + BLOCK_BYTES) & ~(uint64_t)(BLOCK_BYTES - 1u);
	DevBuf d_text, d_items, d_count, d_excl, d_first, d_len, d_item_off, d_rec_off, d_nib_base, d_eos, d_misc, d_tmp;
	CK(d_text.ensure(text_bytes));
	for (uint32_t f = 0; f < n_files; ++f)
		if (bytes[f]) CK(cudaMemcpyAsync((char *)d_text.p + base[f], text[f], bytes[f], cudaMemcpyHostToDevice, st));
	CK(d_items.ensure(std::max<size_t>(1, n_items) * sizeof(Item)));
	CK(d_count.ensure(std::max<size_t>(1, n_items) * 8));
	CK(d_excl.ensure(std::max<size_t>(1, n_items) * 8));
	CK(d_item_off.ensure(std::max<size_t>(1, n_items) * 4));
	CK(d_first.ensure((size_t)(n_rec + 1) * 4));
	CK(d_len.ensure(std::max<size_t>(1, n_rec) * 8));
	CK(d_rec_off.ensure(std::max<size_t>(1, n_rec) * 8));
	CK(d_misc.ensure(64));
	std::vector<unsigned long long> rec_len(n_rec, 0);
	if (n_items) {
		CK(cudaMemcpyAsync(d_items.p, items.data(), (size_t)n_items * sizeof(Item), cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(d_first.p, rec_first.data(), (size_t)(n_rec + 1) * 4, cudaMemcpyHostToDevice, st));
		CK(cudaEventRecord(ctx->ev[0], st));
		fasta_count_kernel<<<grid_for(32ull * n_items, 256), 256, 0, st>>>(d_text.as<uint8_t>(), d_items.as<Item>(), n_items, d_count.as<unsigned long long>());
		CK(cudaGetLastError());
		CK(cudaEventRecord(ctx->ev[1], st));
		size_t tb = 0;
		CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_count.as<unsigned long long>(), d_excl.as<unsigned long long>(), (int)n_items, st));
		CK(d_tmp.ensure(tb));
		CK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tb, d_count.as<unsigned long long>(), d_excl.as<unsigned long long>(), (int)n_items, st));
		fasta_lengths_kernel<<<grid_for(n_rec, 128), 128, 0, st>>>(d_excl.as<unsigned long long>(), d_count.as<unsigned long long>(),
			d_first.as<uint32_t>(), n_rec, n_items, d_len.as<unsigned long long>(), d_item_off.as<uint32_t>());
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(rec_len.data(), d_len.p, (size_t)n_rec * 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
	}
	const double t2 = now();
	// ---- which records become sequences (parse_fasta.cpp:35-47,76-86) ---------------------------------------------
	SeqSet &s = ctx->sets[kind];
	if (!ctx->fasta[kind]) ctx->fasta[kind] = new pcramp_gpu_fasta();
	pcramp_gpu_fasta &tab = *ctx->fasta[kind];
	tab = pcramp_gpu_fasta();
	if (n_items) tab.ms_count = ev_ms(ctx->ev[0], ctx->ev[1]);
	tab.text_bytes = base[n_files];
	std::vector<uint64_t> rec_off(n_rec, ~0ull);
	std::vector<uint32_t> rec_seq(n_rec, 0xFFFFFFFFu), rec_nib_base(n_rec, 0u);
	std::vector<std::pair<uint32_t, uint32_t>> pads; // (sequence, position) of the EOS between the records of a group
	std::vector<uint64_t> raw_off;
	uint64_t off = 0;
	if (!file_group) {
		for (uint32_t r = 0; r < n_rec; ++r) {
			const uint64_t L = rec_len[r];
			if (L < min_length || L > max_length || recs[r].ignored) continue;
			if (L >= (1ull << 32)) return fail(ctx, "pcramp_gpu_upload_fasta: a record of 2^32 bases or more");
			if (recs[r].weight < 0.0f) return fail(ctx, ":Sequence::defline: Negative weights are not allowed!"); // sequence.h:197-199
			rec_seq[r] = (uint32_t)tab.length.size();
			rec_off[r] = off;
			raw_off.push_back(off);
			off += (((L + 1) / 2) + 15u) & ~15ull;
			tab.file.push_back(recs[r].file);
			tab.def_off.push_back(recs[r].def_off);
			tab.def_len.push_back(recs[r].def_len);
			tab.length.push_back((uint32_t)L);
			tab.weight.push_back(recs[r].weight);
		}
	} else {
		std::vector<uint64_t> seq_len;
		uint32_t cur_group = 0;
		bool open = false;
		for (uint32_t r = 0; r < n_rec; ++r) {
			const uint64_t L = rec_len[r];
			if (L < min_length || L > max_length || recs[r].ignored) continue; // parse_fasta.cpp:116-118,155-156
			const uint32_t g = file_group[recs[r].file];
			if (!open || g != cur_group) { // the next group's (so far empty) Sequence, main.cpp:300
				open = false;
				cur_group = g;
			}
			if (!open && L == 0) continue; // m_seq is empty: no pad, and nothing is appended
			if (!open) {
				open = true;
				seq_len.push_back(0);
				tab.file.push_back(recs[r].file);
				tab.def_off.push_back(recs[r].def_off);
				tab.def_len.push_back(recs[r].def_len);
				tab.weight.push_back(1.0f); // DEFAULT_SCORE_WEIGHT; a weight in the group's name is the caller's (pcramp_gpu_set_weights)
			}
			const uint32_t q = (uint32_t)seq_len.size() - 1u;
			if (seq_len[q] > 0) { // :120-125: m_num_pad EOS between records
				for (uint32_t k = 0; k < num_pad; ++k) pads.push_back(std::make_pair(q, (uint32_t)(seq_len[q] + k)));
				seq_len[q] += num_pad;
			}
			rec_seq[r] = q;
			rec_nib_base[r] = (uint32_t)seq_len[q];
			seq_len[q] += L;
			if (seq_len[q] >= (1ull << 32)) return fail(ctx, "pcramp_gpu_upload_fasta_groups: a group of 2^32 bases or more");
		}
		for (uint64_t L : seq_len) {
			raw_off.push_back(off);
			off += (((L + 1) / 2) + 15u) & ~15ull;
			tab.length.push_back((uint32_t)L);
		}
		for (uint32_t r = 0; r < n_rec; ++r)
			if (rec_seq[r] != 0xFFFFFFFFu) rec_off[r] = raw_off[rec_seq[r]];
	}
	const uint32_t n = (uint32_t)tab.length.size();
	if (n >= (1u << 24)) return fail(ctx, "pcramp_gpu_upload_fasta: at most 2^24 - 1 sequences per collection");
	// ---- pack -------------------------------------------------------------------------------------------------
	s.n = n;
	s.db_valid = false;
	if (kind == PCRAMP_MULTIPLEX) ctx->mpx_valid = false;
	s.idx_drop();
	s.n_entries = s.n_keys = 0;
	s.len = tab.length;
	s.weight = tab.weight;
	s.unit_weights = true;
	for (float w : s.weight) s.unit_weights = s.unit_weights && (w == 1.0f);
	s.active.assign(n, 1);
	s.raw_off = raw_off;
	s.raw_bytes = off;
	s.eos.assign(n, std::vector<uint32_t>());
	s.any_degenerate = false;
	CK(s.d_raw.ensure(std::max<uint64_t>(16, s.raw_bytes)));
	CK(cudaMemsetAsync(s.d_raw.p, 0, std::max<uint64_t>(16, s.raw_bytes), st));
	uint32_t eos_cap = 1u << 16;
	std::vector<uint2> eos_list;
	for (int attempt = 0; n_items && n; ++attempt) {
		CK(d_eos.ensure((size_t)eos_cap * sizeof(uint2)));
		unsigned long long init[2] = {~0ull, 0ull}; // bad offset | flags (lo), n_eos (hi)
		CK(cudaMemcpyAsync(d_misc.p, init, 16, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(d_rec_off.p, rec_off.data(), (size_t)n_rec * 8, cudaMemcpyHostToDevice, st));
		CK(d_nib_base.ensure(std::max<size_t>(1, n_rec) * 4));
		CK(cudaMemcpyAsync(d_nib_base.p, rec_nib_base.data(), (size_t)n_rec * 4, cudaMemcpyHostToDevice, st));
		if (attempt) CK(cudaMemsetAsync(s.d_raw.p, 0, std::max<uint64_t>(16, s.raw_bytes), st));
		PackOut po;
		po.raw = s.d_raw.as<uint8_t>();
		po.rec_raw_off = d_rec_off.as<uint64_t>();
		po.rec_nib_base = d_nib_base.as<uint32_t>();
		po.bad = (unsigned long long *)d_misc.p;
		po.flags = (unsigned int *)d_misc.p + 2;
		po.n_eos = (unsigned int *)d_misc.p + 3;
		po.eos = d_eos.as<uint2>();
		po.eos_cap = eos_cap;
		CK(cudaEventRecord(ctx->ev[0], st));
		fasta_pack_kernel<<<grid_for(32ull * n_items, 256), 256, 0, st>>>(d_text.as<uint8_t>(), d_items.as<Item>(), n_items, d_item_off.as<uint32_t>(), po);
		CK(cudaGetLastError());
		CK(cudaEventRecord(ctx->ev[1], st));
		unsigned long long res[2];
		CK(cudaMemcpyAsync(res, d_misc.p, 16, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		if (res[0] != ~0ull) { // base_to_bits throws on the first symbol it does not know (base_table.h:68-73)
			char b[160];
			uint32_t f = 0;
			while (f + 1 < n_files && res[0] >= base[f + 1]) ++f;
			snprintf(b, sizeof(b), ":base_to_bits: Illegal base (symbol %d at byte %llu of file %u)", (int)(unsigned char)text[f][res[0] - base[f]],
				(unsigned long long)(res[0] - base[f]), f);
			s.n = 0;
			return fail(ctx, b);
		}
		tab.ms_pack = ev_ms(ctx->ev[0], ctx->ev[1]);
		s.any_degenerate = ((uint32_t)res[1] & 1u) != 0u;
		const uint32_t n_eos = (uint32_t)(res[1] >> 32);
		if (n_eos > eos_cap) {
			if (attempt >= 1) return fail(ctx, "pcramp_gpu_upload_fasta: EOS list kept overflowing");
			eos_cap = n_eos + 1024u;
			continue;
		}
		eos_list.resize(n_eos);
		if (n_eos) {
			CK(cudaMemcpyAsync(eos_list.data(), d_eos.p, (size_t)n_eos * sizeof(uint2), cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
		}
		break;
	}
	for (const uint2 &e : eos_list) s.eos[rec_seq[e.x]].push_back(e.y);
	for (const std::pair<uint32_t, uint32_t> &pd : pads) s.eos[pd.first].push_back(pd.second);
	std::vector<uint32_t> with_eos;
	s.plen.assign(n, 0);
	s.clen.assign(n, 0);
	s.grp_off.assign(n + 1, 0);
	for (uint32_t i = 0; i < n; ++i) {
		std::sort(s.eos[i].begin(), s.eos[i].end());
		const uint32_t L = s.len[i];
		s.clen[i] = L - (uint32_t)s.eos[i].size();
		if (!s.eos[i].empty()) with_eos.push_back(i);
		s.plen[i] = L + (L & 1u);
		if (L & 1u) s.eos[i].push_back(L); // the pad nibble pack() also pushes (seqdev.cuh)
		s.grp_off[i + 1] = s.grp_off[i] + ((uint64_t)s.clen[i] + 31) / 32 + 1;
	}
	if (n_records) *n_records = n;
	for (uint32_t L : tab.length) tab.n_bases += L;
	ctx->stats.kernel_launches = n_items ? 4 : 0;
	const double t3 = now();
	const int rc = upload_finish(ctx, s, nullptr, with_eos);
	if (dbg) fprintf(stderr, "upload_fasta: split %.1f ms, H2D + count %.1f ms, pack %.1f ms, finish %.1f ms\n", t1 - t0, t2 - t1, t3 - t2, now() - t3);
	return rc;
}

int pcramp_gpu_upload_fasta(pcramp_gpu_ctx *ctx, int kind, uint32_t n_files, const char *const *text, const uint64_t *bytes, uint64_t min_length,
	uint64_t max_length, uint32_t n_ignore, const char *const *ignore, uint32_t *n_records)
{
	return upload_fasta_impl(ctx, kind, n_files, text, bytes, min_length, max_length, n_ignore, ignore, n_records, nullptr, 0);
}

int pcramp_gpu_upload_fasta_groups(pcramp_gpu_ctx *ctx, int kind, uint32_t n_files, const char *const *text, const uint64_t *bytes,
	const uint32_t *file_group, uint64_t min_length, uint64_t max_length, uint32_t num_pad, uint32_t n_ignore, const char *const *ignore,
	uint32_t *n_sequences)
{
	if (n_files && !file_group) return ctx ? fail(ctx, "pcramp_gpu_upload_fasta_groups: null argument") : 1;
	static const uint32_t none = 0;
	return upload_fasta_impl(ctx, kind, n_files, text, bytes, min_length, max_length, n_ignore, ignore, n_sequences, n_files ? file_group : &none, num_pad);
}

/* host-only: the record split of one file's text as pcramp_gpu_upload_fasta sees it (before the length window and the ignore
 * list): per record the defline (offset, length), the residue span [begin, end) and the weight.  Returns the number of records;
 * fills at most cap entries. */
uint32_t pcramp_fasta_scan(const char *text, uint64_t bytes, uint32_t cap, uint64_t *defline_off, uint32_t *defline_len, uint64_t *begin,
	uint64_t *end, float *weight)
{
	std::vector<pcr::fasta::Record> recs;
	pcr::fasta::split_records(text, bytes, 0, 0, recs);
	for (size_t i = 0; i < recs.size() && i < cap; ++i) {
		if (defline_off) defline_off[i] = recs[i].def_off;
		if (defline_len) defline_len[i] = recs[i].def_len;
		if (begin) begin[i] = recs[i].begin;
		if (end) end[i] = recs[i].end;
		if (weight) weight[i] = pcr::fasta::extract_weight(std::string(text + recs[i].def_off, recs[i].def_len));
	}
	return (uint32_t)recs.size();
}

/* CUDA-event times (ms) of the two device passes of the last pcramp_gpu_upload_fasta on `kind`, its text bytes and kept bases */
int pcramp_gpu_fasta_timing(pcramp_gpu_ctx *ctx, int kind, float *ms_count, float *ms_pack, uint64_t *text_bytes, uint64_t *n_bases)
{
	if (check_kind(ctx, kind)) return 1;
	if (!ctx->fasta[kind]) return fail(ctx, "pcramp_gpu_fasta_timing: no FASTA upload on this collection");
	const pcramp_gpu_fasta &t = *ctx->fasta[kind];
	if (ms_count) *ms_count = t.ms_count;
	if (ms_pack) *ms_pack = t.ms_pack;
	if (text_bytes) *text_bytes = t.text_bytes;
	if (n_bases) *n_bases = t.n_bases;
	return 0;
}

void pcramp_gpu_fasta_free(pcramp_gpu_ctx *ctx)
{
	if (!ctx) return;
	for (int k = 0; k < PCRAMP_NUM_KINDS; ++k) {
		delete ctx->fasta[k];
		ctx->fasta[k] = nullptr;
	}
}

/* the record table of the last pcramp_gpu_upload_fasta on `kind`: per sequence the file it came from, its defline as
 * (offset, length) into that file's text, its length and weight; any output may be NULL */
int pcramp_gpu_fasta_records(pcramp_gpu_ctx *ctx, int kind, uint32_t *file, uint64_t *defline_off, uint32_t *defline_len, uint32_t *length,
	float *weight)
{
	if (check_kind(ctx, kind)) return 1;
	if (!ctx->fasta[kind]) return fail(ctx, "pcramp_gpu_fasta_records: no FASTA upload on this collection");
	const pcramp_gpu_fasta &t = *ctx->fasta[kind];
	const size_t n = t.length.size();
	if (file) memcpy(file, t.file.data(), n * 4);
	if (defline_off) memcpy(defline_off, t.def_off.data(), n * 8);
	if (defline_len) memcpy(defline_len, t.def_len.data(), n * 4);
	if (length) memcpy(length, t.length.data(), n * 4);
	if (weight) memcpy(weight, t.weight.data(), n * 4);
	return 0;
}

/* the collection as the reference stores it (sequence.h:85,223-228): per sequence byte offset (16-byte aligned) and length, and
 * the packed nibbles; any output may be NULL.  total_bytes = size of the nibble array. */
int pcramp_gpu_sequences_copy(pcramp_gpu_ctx *ctx, int kind, uint32_t *n, uint64_t *total_bytes, uint64_t *byte_off, uint32_t *length,
	uint8_t *nibbles)
{
	if (check_kind(ctx, kind)) return 1;
	CK(cudaSetDevice(ctx->device));
	SeqSet &s = ctx->sets[kind];
	if (n) *n = s.n;
	if (total_bytes) *total_bytes = s.raw_bytes;
	if (byte_off && s.n) memcpy(byte_off, s.raw_off.data(), (size_t)s.n * 8);
	if (length && s.n) memcpy(length, s.len.data(), (size_t)s.n * 4);
	if (nibbles && s.raw_bytes) {
		CK(cudaMemcpyAsync(nibbles, s.d_raw.p, s.raw_bytes, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	return 0;
}

} // extern "C"
