// ctx.cuh -- the context object behind the C ABI (shared by the translation units of libpcramp_gpu.so).
#pragma once
#include "../../include/pcramp_gpu.h"
#include "seqdev.cuh"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <memory>
#include <string>
#include <atomic>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

struct pcramp_gpu_ctx;
struct pcramp_gpu_xchg; // xchg.cuh
struct pcramp_gpu_fasta; // fasta.cuh
namespace pcr {
namespace nc {
struct ThermoState; // thermo_abi.cu
void thermo_state_free(ThermoState *);
int thermo_run_codes(pcramp_gpu_ctx *ctx, int op, uint32_t n, const uint8_t *codes, const uint8_t *len, const float *strand, float salt,
	float *tm_out);
}
}

namespace pcr {

// device allocation bookkeeping for the stage trace (PCRAMP_TRACE): calls and host milliseconds spent in cudaMalloc / cudaFree
struct AllocTrace {
	static std::atomic<uint64_t> &calls() { static std::atomic<uint64_t> v{0}; return v; }
	static std::atomic<uint64_t> &micros() { static std::atomic<uint64_t> v{0}; return v; }
	static bool on() { static const bool v = getenv("PCRAMP_TRACE") != nullptr; return v; }
	static uint64_t now_us()
	{
		struct timespec ts;
		clock_gettime(CLOCK_MONOTONIC, &ts);
		return (uint64_t)ts.tv_sec * 1000000ull + (uint64_t)ts.tv_nsec / 1000ull;
	}
};

// A process-wide cache of released device blocks.  cudaMalloc / cudaFree are not cheap and not steady: under the stage trace single
// calls of 100-400 ms show up now and then (a 1 MB buffer in pcramp_gpu_multiplex_keys: 380 ms), which is several design iterations.
// Per-call scratch (candidate lists, sort buffers, bit rows) is therefore handed back here and re-used by the next request of a
// similar size; blocks above 1 GB (index build scratch) go back to the driver, a full cache (8 GB) gives up its oldest blocks, and a
// failed cudaMalloc empties the cache and tries again.  A release synchronises the device first, as cudaFree does, so a block never
// changes hands under a kernel that still reads it.  PCRAMP_NO_ALLOC_CACHE=1 switches the cache off.
struct AllocCache {
	struct Block {
		void *p;
		size_t cap;
		int dev;
	};
	static constexpr size_t MAX_BLOCK = 1ull << 30, MAX_TOTAL = 8ull << 30;
	static std::mutex &mu() { static std::mutex m; return m; }
	static std::vector<Block> &blocks() { static std::vector<Block> *v = new std::vector<Block>(); return *v; } // never destroyed: no CUDA call at exit
	static size_t &bytes() { static size_t b = 0; return b; }
	static bool enabled() { static const bool v = getenv("PCRAMP_NO_ALLOC_CACHE") == nullptr; return v; }
	static void *take(size_t want, size_t &cap_out)
	{
		if (!enabled()) return nullptr;
		int dev = 0;
		cudaGetDevice(&dev);
		std::lock_guard<std::mutex> lock(mu());
		std::vector<Block> &b = blocks();
		size_t best = b.size();
		for (size_t i = 0; i < b.size(); ++i)
			if (b[i].dev == dev && b[i].cap >= want && b[i].cap <= 2 * want + 65536 && (best == b.size() || b[i].cap < b[best].cap)) best = i;
		if (best == b.size()) return nullptr;
		void *p = b[best].p;
		cap_out = b[best].cap;
		bytes() -= b[best].cap;
		b.erase(b.begin() + (std::ptrdiff_t)best); // keeps the age order put() evicts by
		return p;
	}
	static bool put(void *p, size_t cap)
	{
		if (!enabled() || cap > MAX_BLOCK) return false;
		int dev = 0;
		cudaGetDevice(&dev);
		std::lock_guard<std::mutex> lock(mu());
		std::vector<Block> &b = blocks();
		while (bytes() + cap > MAX_TOTAL && !b.empty()) { // full: the oldest blocks go back to the driver (sizes nobody asked for again)
			cudaFree(b.front().p);
			bytes() -= b.front().cap;
			b.erase(b.begin());
		}
		blocks().push_back(Block{p, cap, dev});
		bytes() += cap;
		return true;
	}
	static void flush()
	{
		std::lock_guard<std::mutex> lock(mu());
		for (const Block &k : blocks()) cudaFree(k.p);
		blocks().clear();
		bytes() = 0;
	}
};

struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	bool owned = true; // false: a view of another context's buffer (worker contexts, pcramp_gpu_create_worker)
	DevBuf() = default;
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	~DevBuf() { release(); }
	void release()
	{
		if (p && owned) {
			const uint64_t t0 = AllocTrace::on() ? AllocTrace::now_us() : 0;
			bool cached = false;
			if (AllocCache::enabled() && cap <= AllocCache::MAX_BLOCK) {
				cudaDeviceSynchronize(); // what cudaFree does implicitly: nobody still reads the block
				cached = AllocCache::put(p, cap);
			}
			if (!cached) {
				cudaFree(p);
				if (AllocTrace::on()) { AllocTrace::calls()++; AllocTrace::micros() += AllocTrace::now_us() - t0; }
			}
		}
		p = nullptr;
		cap = 0;
		owned = true;
	}
	void alias(const DevBuf &o)
	{
		release();
		p = o.p;
		cap = o.cap;
		owned = false;
	}
	cudaError_t ensure(size_t bytes)
	{
		if (bytes <= cap) return cudaSuccess;
		release();
		size_t want = bytes + bytes / 4 + 256;
		size_t got = 0;
		if (void *q = AllocCache::take(want, got)) {
			p = q;
			cap = got;
			return cudaSuccess;
		}
		const uint64_t t0 = AllocTrace::on() ? AllocTrace::now_us() : 0;
		cudaError_t e = cudaMalloc(&p, want);
		if (e != cudaSuccess) { // out of memory with blocks parked in the cache: give them back and try once more
			(void)cudaGetLastError();
			AllocCache::flush();
			e = cudaMalloc(&p, want);
		}
		if (AllocTrace::on()) { AllocTrace::calls()++; AllocTrace::micros() += AllocTrace::now_us() - t0; }
		if (e != cudaSuccess) { p = nullptr; return e; }
		cap = want;
		return cudaSuccess;
	}
	template <class T> T *as() const { return (T *)p; }
};

// scratch allocator for the library algorithms that take one (thrust::cuda::par(alloc)): blocks come from and go back to the cache
// above instead of one cudaMalloc + cudaFree per call
struct CachedScratch {
	typedef char value_type;
	std::vector<std::pair<char *, size_t>> live;
	char *allocate(std::ptrdiff_t n)
	{
		DevBuf b;
		if (b.ensure((size_t)std::max<std::ptrdiff_t>(n, 1)) != cudaSuccess) throw std::bad_alloc();
		char *p = (char *)b.p;
		live.push_back(std::make_pair(p, b.cap));
		b.p = nullptr; // ownership moves to `live`
		b.cap = 0;
		return p;
	}
	void deallocate(char *p, size_t)
	{
		for (size_t i = 0; i < live.size(); ++i)
			if (live[i].first == p) {
				DevBuf b;
				b.p = p;
				b.cap = live[i].second;
				live[i] = live.back();
				live.pop_back();
				return; // b's destructor hands the block to the cache
			}
	}
};

struct SeqSet {
	uint32_t n = 0;
	bool any_degenerate = false;
	bool unit_weights = true; // every weight == 1.0f: coverage = population count
	uint64_t total_positions = 0; // sum of clen over all sequences
	std::vector<uint32_t> len, plen, clen;
	std::vector<float> weight;
	std::vector<uint8_t> active;
	std::vector<uint64_t> raw_off, grp_off;
	std::vector<std::vector<uint32_t>> eos; // per sequence, sorted raw positions
	uint64_t raw_bytes = 0, n_groups = 0, n_tiles = 0;
	DevBuf d_raw, d_raw_off, d_len, d_plen, d_clen, d_planes, d_grp_off, d_eos_pos, d_eos_off, d_weight, d_active, d_tile_seq, d_tile_x0;
	DevBuf d_dirty_bits, d_dirty_seq, d_dirty_grp; // groups whose alignments read a degenerate base (scan.cuh)
	uint32_t n_dirty = 0;
	// text index for the indexed seed scan (index.cuh): built lazily on the first seeded scan, in PARTS of consecutive sequences that
	// hold fewer than idx_part_cap positions each (32-bit positions inside a part; a 5 x 10^9-base collection is three parts).
	// A split does not rebuild it: the split sequences are flagged stale -- their entries are ignored, and while they are active
	// (normally they are retired by the assay that split them, main.cpp:1116-1121) the table scan covers them; the index is rebuilt
	// only when the stale sequences that are active again hold a sizeable share of the text.
	struct IndexPart {
		uint32_t seq_lo = 0, seq_hi = 0; // sequences [seq_lo, seq_hi)
		uint32_t n = 0;                  // positions = entries
		DevBuf entries, off, cum, blk;
	};
	std::vector<std::unique_ptr<IndexPart>> idx_parts;
	bool idx_valid = false, idx_failed = false;
	std::vector<uint8_t> idx_stale;      // per sequence: split since the index was built
	uint32_t n_idx_stale = 0;
	DevBuf d_idx_stale;
	uint64_t idx_bytes = 0, idx_builds = 0;
	float idx_build_ms = 0.0f;
	// the partial words of the collection as a table (edge.cuh): built on the first fast batch after an upload / split, for one set
	// of pack() parameters; `failed` = not worth building (too large) until the text changes
	struct EdgeTab {
		DevBuf planes, meta, start, ids, degen;
		uint32_t n_words = 0, n_degen = 0;
		bool valid = false, failed = false;
		PackParams pp = {};
		uint64_t bytes = 0, builds = 0;
		float build_ms = 0.0f;
		void drop() { valid = failed = false; }
	} edge;
	void idx_drop()
	{
		idx_parts.clear();
		idx_valid = idx_failed = false;
		idx_stale.clear();
		n_idx_stale = 0;
		idx_bytes = 0;
		edge.drop(); // (every caller has just changed the text)
	}
	// database (seq-grouped order = entry-id order) + canonical permutation
	uint64_t n_entries = 0, n_keys = 0;
	bool db_valid = false;
	DevBuf e_hi, e_lo, e_planes, e_seq, e_loc, e_strand, e_perm, e_keyrank, seq_ent_off;
	DevBuf e_id;         // (type | position) of every entry: what the words are re-derived from on demand (db.cuh words_kernel)
	bool words_valid = false;
	PackParams db_pp = {};
	uint32_t db_pb = 0;
	DevBuf seq_full_end; // per (sequence, strand) run: first entry that is not a full-window entry (score.cuh, entry-driven scoring)
	DevBuf e_key, key_planes; // key index per entry (entry-id order), letter planes per unique word
	// neighbour filter of pair scoring (score.cuh): per entry one candidate word that produced it, and the candidate words of
	// the select_words call that built this database (letter planes + seed threshold)
	DevBuf e_cand, c_planes, c_thr;
	uint32_t n_cand = 0;
	DevBuf e_order;           // (index, loc, strand) sort key per entry: input of the canonical order, built on demand
	bool keys_valid = false;
	uint32_t seq_bits = 1;

	SeqDev dev() const
	{
		SeqDev sd;
		sd.n = n;
		sd.planes = d_planes.as<uint4>();
		sd.grp_off = d_grp_off.as<uint64_t>();
		sd.clen = d_clen.as<uint32_t>();
		sd.len = d_len.as<uint32_t>();
		sd.plen = d_plen.as<uint32_t>();
		sd.raw = d_raw.as<uint8_t>();
		sd.raw_off = d_raw_off.as<uint64_t>();
		sd.eos_pos = d_eos_pos.as<uint32_t>();
		sd.eos_off = d_eos_off.as<uint32_t>();
		sd.weight = d_weight.as<float>();
		sd.active = d_active.as<uint8_t>();
		return sd;
	}
};

} // namespace pcr

struct pcramp_gpu_ctx {
	using DevBuf = pcr::DevBuf;
	using SeqSet = pcr::SeqSet;
	// worker contexts (pcramp_gpu_create_worker): own stream / scratch / database, the parent's sequences and text index by reference.
	// text_gen counts changes of the parent's collections; a worker refuses to run once it is behind.
	pcramp_gpu_ctx *parent = nullptr;
	std::atomic<uint64_t> text_gen{0};   // read by the workers' host threads
	uint64_t seen_gen = 0;
	std::atomic<int> n_workers{0};       // live workers: the parent is not destroyed under them
	int device = 0;
	int sm_count = 148;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[10] = {};
	std::string err;
	SeqSet sets[PCRAMP_NUM_KINDS];
	// staged pairs + results
	uint32_t n_pairs = 0, res_words = 0; // n_pairs = size of the current batch window
	uint32_t n_staged = 0, batch_first = 0;
	const uint64_t *pf() const { return d_f.as<uint64_t>() + 2ull * batch_first; }
	const uint64_t *pr() const { return d_r.as<uint64_t>() + 2ull * batch_first; }
	DevBuf d_f, d_r, d_oligos, d_oligos_base, d_variants, d_cov, d_bits, d_bits1;
	DevBuf d_keybits, d_items, d_item_count; // K2 filter: key matrix / work list
	DevBuf d_fst_planes, d_fst_thr, d_fst_cnt, d_fst_start, d_fst_cursor, d_fst_ids, d_fst_combo, d_fst_brute, d_fst_nbrute, d_seqbits; // fst.cuh
	int use_fst = 1;
	// candidates / patterns
	DevBuf d_cand_cnt, d_cand_off, d_cand_words, d_cand_thr, d_pat_mask, d_pat_meta, d_pat_meta2, d_pat_seeded, d_pat_sbefore;
	DevBuf d_part_mask, d_part_meta, d_part_meta2; // seeded patterns first, brute-force patterns after
	DevBuf d_seed_cnt, d_seed_start, d_seed_bucket, d_seed_entries, d_tile_counter;
	int max_smem_optin = 0;
	int force_brute = 0;
	int use_index = 1;
	// the speculative, host-round-trip-free form of select_words (pcramp_gpu.cu select_words_fast): sizes come from the previous
	// batch of the same shape, everything is verified from ONE read-back when somebody needs the result
	struct FastHint {
		bool ok = false;
		uint32_t n_pairs = 0;
		float threshold = 0.0f;
		pcr::PackParams pp = {};
		uint64_t hit_cap = 0, q_cap = 0, c_cap = 0, fst_cap = 0;
	} fast_hint[PCRAMP_NUM_KINDS];
	struct FastPending {
		bool active = false;
		int kind = 0;
		float threshold = 0.0f;
		pcr::PackParams pp = {};
		uint32_t n_pat = 0, n_seg = 0;
		uint64_t hit_cap = 0, q_cap = 0, c_cap = 0;
	} fast_pending;
	int use_fast = 1;
	uint64_t n_fast = 0, n_fast_redo = 0; // batches that took the fast form; of those, batches whose verification failed (re-run in the general form)
	cudaStream_t stream2 = nullptr;       // the partial-word scan runs beside the indexed scan
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_done = nullptr;
	unsigned long long *h_fast = nullptr; // pinned: {flags, n_hits, n_queries, n_indexed, n_index_entries, n_candidates, n_entries}
	DevBuf d_fast;
	int tiny_buffers = 0; // testing hook (pcramp_gpu_set_option): growable buffers start far too small
	uint64_t idx_part_cap = 1ull << 31; // positions per part of the text index (option "index_part_positions": small values for the tests)
	DevBuf idx_key[2], idx_val[2], idx_tmp, d_stale_tile_seq, d_stale_tile_x0; // build scratch (kept while small), tiles of stale sequences
	DevBuf d_idx_queries, d_idx_counters, d_idx_cand;
	// scratch
	DevBuf ent_cand[2], d_neigh, d_neigh_off, d_tier_best;
	int use_seg_db = 1;      // segmented database build (db.cuh); 0 = radix sort + unique-by-key of the entry ids
	DevBuf seg_cnt, seg_off, seg_cursor, seg_uniq, seg_full, seg_big;
	int use_entry_score = 1; // pair scoring driven by the plus-strand entries (score.cuh); 0 = bit rows + item list
	int use_neigh = 1, use_tier_table = 1, use_fused_score = 0; // score_seqbits_kernel: measured slower than the item list (0.65 vs 0.60 ms), kept as an option
	DevBuf hit_key[2], hit_val[2], ent_id[2], d_counters, cub_tmp, cub_tmp2, order_key[2], perm[2], head;
	unsigned long long *h_counters = nullptr; // pinned
	pcramp_gpu_stats stats = {};
	bool pend_ms_db = false, pend_ms_score = false; // event times of the last calls not read back yet (pcramp_gpu_get_stats)
	pcr::nc::ThermoState *thermo = nullptr; // K3 state, created on first use (thermo_abi.cu)
	pcramp_gpu_xchg *xchg = nullptr;        // multi-GPU exchange state (xchg.cuh)
	pcramp_gpu_fasta *fasta[PCRAMP_NUM_KINDS] = {}; // record table of the last FASTA upload per collection (fasta.cuh)
	// multiplex terms of optimize() (multiplex.cuh): unique words of the multiplex background, the assay pool
	DevBuf mpx_words, mpx_planes, mpx_items, mpx_item_off, mpx_base, mpx_var, mpx_bidx, mpx_cov, mpx_pool, mpx_ov_words, mpx_ov;
	uint64_t mpx_n_keys = 0;
	bool mpx_valid = false;
	std::vector<uint64_t> pool_words; // F0 R0 F1 R1 ... (2 x uint64 each)
	// unique amplicons of the last pcramp_gpu_unique_amplicons call (amplicon.cuh): candidate records in the reference's order
	// (region = sequence, first base, length; bounds), and the records that stand for the distinct amplicon strings of every pair
	DevBuf amp_seq, amp_start, amp_len, amp_pair, amp_bounds, amp_uniq, amp_pair_off, amp_text_off, amp_flags;
	uint64_t amp_n_rec = 0, amp_n_uniq = 0, amp_n_bases = 0;
	uint32_t amp_n_pairs = 0;
	int amp_kind = -1;
	bool amp_bounds_ok = false;
	std::vector<uint64_t> amp_words; // F0 R0 F1 R1 ... of that call
	DevBuf sw_q, sw_t, sw_out;       // pcramp_gpu_sw_batch staging (kept between calls)
	DevBuf d_variant_groups;         // score_variants: base assays of the groups + group offsets
	std::vector<uint64_t> h_group_words;
	std::vector<uint32_t> h_group_off;
	int use_variant_groups = 1;      // option "use_variant_groups"
	DevBuf bg_cnt4, bg_off4, bg_entry, bg_res, bg_cnt2, bg_off2; // find_background_match by units (sw_abi.cuh)
	int use_background_units = 1;    // option "use_background_units"
	int use_unit_score = 1;          // option "use_unit_score": pair scoring by (sequence, pair) units at unselective thresholds
	int use_async_scan = 0;          // option "use_async_scan" = 1: scan_index_async_kernel instead of scan_index_kernel (measured slower)
	int use_edge_table = 1;          // option "use_edge_table": fast batches look the candidates up in the collection's partial-word table (edge.cuh)
	float edge_off_thr[PCRAMP_NUM_KINDS] = {-1.0f, -1.0f, -1.0f}; // a threshold whose candidates the table could not serve: the scan kernel keeps it
	bool async_scan_ready = false;   // its dynamic shared memory size has been set on this context's device
	float sw_ms_kernel = 0.0f;
};

namespace pcr {

#define CK(call)                                                                                      \
	do {                                                                                              \
		cudaError_t e__ = (call);                                                                     \
		if (e__ != cudaSuccess) {                                                                     \
			char b__[512];                                                                            \
			snprintf(b__, sizeof(b__), "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
			ctx->err = b__;                                                                           \
			return 1;                                                                                 \
		}                                                                                             \
	} while (0)

inline int fail(pcramp_gpu_ctx *ctx, const std::string &m)
{
	ctx->err = m;
	return 1;
}

inline uint32_t bits_for(uint64_t n)
{
	uint32_t b = 1;
	while ((1ull << b) < n) ++b;
	return b;
}

inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// host-side stage trace (environment PCRAMP_TRACE=1): wall-clock between marks, printed to stderr; synchronises the stream at
// every mark, so it is a debugging aid and never on in measurements
struct Trace {
	bool on;
	cudaStream_t st;
	const char *who;
	double t0, last;
	static double now()
	{
		struct timespec ts;
		clock_gettime(CLOCK_MONOTONIC, &ts);
		return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
	}
	Trace(const char *w, cudaStream_t s) : on(getenv("PCRAMP_TRACE") != nullptr), st(s), who(w), t0(0), last(0)
	{
		if (on) t0 = last = now();
	}
	void mark(const char *what)
	{
		if (!on) return;
		cudaStreamSynchronize(st);
		const double t = now();
		fprintf(stderr, "[trace] %s: %-28s %9.3f ms (at %9.3f)\n", who, what, t - last, t - t0);
		last = t;
	}
};

} // namespace pcr
