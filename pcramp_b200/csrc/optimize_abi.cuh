// optimize_abi.cuh -- the local search optimize() (optimize.cpp:14-207) and its six moves (optimize_pcr.cpp:8-989) for a
// BATCH of trial assays in lock step (included at the end of pcramp_gpu.cu).
//
// The reference runs optimize() once per trial inside an OpenMP loop (main.cpp:697-729); every iteration scores the
// assay (collect_candidates / update_identity / compute_coverage on targets and backgrounds), then tries each move on
// each oligo -- a move is a list of trial oligos, each filtered by PCR::is_valid and scored against the candidate
// amplicons of the unmoved assay -- and applies the best one.  Here one iteration is a handful of batched GPU calls
// over ALL live trials: K2 for the assays, K3 (duplex Tm, then hairpins) for every trial oligo of every move, K2 in
// variant mode for the survivors; the selection logic (strict / tie comparisons, the running score threshold that
// gates the background evaluation, the revisited-assay set) then replays on the host in the reference's order.
//
// NucCruc history.  optimize() owns ONE NucCruc object per trial (optimize.cpp:49-52) and PCR::is_valid loads each
// trial oligo whose duplex Tm is in range into it (valid_pcr.cpp:27).  The hairpin evaluation reads up to two slots past
// the end of that query (nuccruc.cuh header), i.e. bases of the last LONGER oligo loaded before -- e.g. the untrimmed
// primer when a trim move is evaluated.  The loop below replays the loads per trial and passes those bases along
// (stale slots start as A, what a zero-filled object holds).
#pragma once
#include "ctx.cuh"
#include "nuccruc.cuh"

#include <algorithm>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

namespace pcr {
namespace opt {

enum : int { MV_INC_DEGEN = 0, MV_DEC_DEGEN = 1, MV_TRIM5 = 2, MV_TRIM3 = 3, MV_GROW5 = 4, MV_GROW3 = 5, MV_COUNT = 6 }; // assay.h:21-29

struct ScoreH { // struct Score (pcramp.h:158-215)
	float target = -1.0e6f, background = 1.0e6f, overlap = 0.0f;
	float accuracy() const { return target - background; }
	bool lt(const ScoreH &o) const { return accuracy() == o.accuracy() ? overlap < o.overlap : accuracy() < o.accuracy(); }
	bool gt(const ScoreH &o) const { return accuracy() == o.accuracy() ? overlap > o.overlap : accuracy() > o.accuracy(); }
	bool eq(const ScoreH &o) const { return accuracy() == o.accuracy() && overlap == o.overlap; }
};

inline W128 word_of(const uint64_t *p)
{
	W128 w;
	w.hi = p[0];
	w.lo = p[1];
	return w;
}
inline double degeneracy(const W128 &w)
{ // Word::degeneracy (word.h:97-138)
	double d = 1.0;
	for (int i = 0; i < WORD_LEN; ++i) {
		const uint32_t b = w_get(w, i);
		const int c = (int)((b & 1u) + ((b >> 1) & 1u) + ((b >> 2) & 1u) + (b >> 3));
		if (c) d *= c;
	}
	return d;
}
inline std::string packed(const W128 &w)
{ // Word::write (word.h:649-657): the nibbles between start() and stop(); any injective code will do
	std::string s;
	const int a = w_start(w), b = w_stop(w);
	for (int i = a; i <= b; ++i) s.push_back((char)('a' + w_get(w, i)));
	return s;
}

struct Variant {
	uint32_t trial;
	int oligo; // 0 = FORWARD, 1 = REVERSE
	int move;
	W128 w;
	// filled in as the batch proceeds
	bool valid = false;
	float cov_t = 0.0f, cov_b = 0.0f;
	float cov_m = 0.0f, ov = 0.0f; // multiplex background coverage; max over the pool of max_overlap(trial oligo, .)
	uint32_t first_exp = 0, n_exp = 0; // its expansions in the thermo batch
};

// host loops over independent trials / variants, split over a few threads: fn(part, lo, hi) on [lo, hi) of n items
template <class F>
inline void parallel_ranges(size_t n, size_t min_per_part, const std::vector<size_t> *cuts, F fn)
{
	const size_t hw = std::max<size_t>(1, std::thread::hardware_concurrency());
	size_t parts = std::min<size_t>(std::min<size_t>(8, hw), n / std::max<size_t>(1, min_per_part));
	if (parts <= 1) {
		fn(0, 0, n);
		return;
	}
	std::vector<size_t> edge(parts + 1, n);
	edge[0] = 0;
	for (size_t k = 1; k < parts; ++k) {
		size_t e = n * k / parts;
		if (cuts) { // move the edge up to the next allowed cut (a sorted list of item indices)
			const auto it = std::lower_bound(cuts->begin(), cuts->end(), e);
			e = it == cuts->end() ? n : *it;
		}
		edge[k] = std::max(e, edge[k - 1]);
	}
	std::vector<std::thread> pool;
	for (size_t k = 0; k < parts; ++k)
		if (edge[k] < edge[k + 1]) pool.emplace_back([&, k]() { fn(k, edge[k], edge[k + 1]); });
	for (std::thread &t : pool) t.join();
}
constexpr size_t PARALLEL_PARTS = 8;

// the trial oligos a move evaluates, in the reference's order (optimize_pcr.cpp)
inline void move_variants(int move, const W128 &o, uint32_t degen_max, int primer_min, int primer_max, std::vector<W128> &out)
{
	out.clear();
	const int first = w_start(o), last = w_stop(o), len = w_size(o);
	switch (move) {
	case MV_INC_DEGEN: // :8-196
		if (degeneracy(o) >= (double)degen_max) return;
		for (int i = first; i <= last; ++i)
			for (uint32_t b = 1u; b <= 8u; b <<= 1) {
				if (w_get(o, i) & b) continue;
				W128 t = o;
				w_set(t, i, w_get(o, i) | b);
				if (degeneracy(t) > (double)degen_max) continue; // tested before is_valid: never reaches the NucCruc object
				out.push_back(t);
			}
		break;
	case MV_DEC_DEGEN: // :198-381
		for (int i = first; i <= last; ++i) {
			const uint32_t cur = w_get(o, i);
			for (uint32_t b = 1u; b <= 8u; b <<= 1) {
				const uint32_t d = cur & ~b;
				if (!d || d == cur) continue;
				W128 t = o;
				w_set(t, i, d);
				out.push_back(t);
			}
		}
		break;
	case MV_TRIM5: // :383-522
		if (len == primer_min) return;
		{
			W128 t = o;
			if (first < WORD_LEN) w_set(t, first, 0u); // shrink_front (word.h:358-365)
			out.push_back(t);
		}
		break;
	case MV_TRIM3: // :524-662
		if (len == primer_min) return;
		{
			W128 t = o;
			if (last >= 0) w_set(t, last, 0u); // shrink_back
			out.push_back(t);
		}
		break;
	case MV_GROW5: // :664-827
		if (len == primer_max) return;
		for (uint32_t b = 1u; b <= 8u; b <<= 1) {
			W128 t = o;
			if (first - 1 >= 0) w_set(t, first - 1, b); // grow_front: only if there is room (word.h:376-383)
			out.push_back(t);
		}
		break;
	case MV_GROW3: // :829-989
		if (len == primer_max) return;
		for (uint32_t b = 1u; b <= 8u; b <<= 1) {
			W128 t = o;
			if (last + 1 < WORD_LEN) w_set(t, last + 1, b); // grow_back
			out.push_back(t);
		}
		break;
	default: break;
	}
}

// every concrete expansion of a (possibly degenerate) oligo as base codes, position 0 fastest (Word::begin / next order
// differs, but is_valid is an AND over all of them)
inline uint64_t expansion_count(const W128 &w)
{
	uint64_t n = 1;
	const int a = w_start(w), b = w_stop(w);
	for (int i = a; i <= b; ++i) {
		const uint32_t nb = w_get(w, i);
		n *= (uint64_t)((nb & 1u) + ((nb >> 1) & 1u) + ((nb >> 2) & 1u) + (nb >> 3));
	}
	return n;
}
inline void expansion_codes(const W128 &w, uint64_t idx, uint8_t *dst, int &len)
{
	static const uint8_t code_of_bit[4] = {nc::bA, nc::bC, nc::bG, nc::bT};
	const int a = w_start(w), b = w_stop(w);
	len = 0;
	for (int i = a; i <= b; ++i) {
		const uint32_t nb = w_get(w, i);
		uint8_t letters[4];
		int k = 0;
		for (int bit = 0; bit < 4; ++bit)
			if (nb & (1u << bit)) letters[k++] = code_of_bit[bit];
		dst[len++] = letters[idx % (uint64_t)k];
		idx /= (uint64_t)k;
	}
}

// all n_exp expansions of an oligo, in expansion_codes' order, 32 bytes each: the single-letter positions are written once and the few
// degenerate ones (at most four for degeneracy <= 16) are stepped like an odometer, position 0 fastest
inline void expansion_codes_all(const W128 &w, uint32_t n_exp, uint8_t *dst, uint8_t *lens)
{
	static const uint8_t code_of_bit[4] = {nc::bA, nc::bC, nc::bG, nc::bT};
	const int a = w_start(w), b = w_stop(w);
	uint8_t base[32] = {0};
	uint8_t deg_pos[32], deg_k[32], deg_letters[32][4], digit[32];
	int len = 0, nd = 0;
	for (int i = a; i >= 0 && i <= b; ++i) {
		const uint32_t nb = w_get(w, i);
		uint8_t letters[4];
		int k = 0;
		for (int bit = 0; bit < 4; ++bit)
			if (nb & (1u << bit)) letters[k++] = code_of_bit[bit];
		base[len] = k ? letters[0] : 0;
		if (k > 1) {
			deg_pos[nd] = (uint8_t)len;
			deg_k[nd] = (uint8_t)k;
			memcpy(deg_letters[nd], letters, 4);
			digit[nd] = 0;
			++nd;
		}
		++len;
	}
	for (uint32_t e = 0; e < n_exp; ++e) {
		uint8_t *out = dst + (size_t)e * 32;
		memcpy(out, base, 32);
		for (int d = 0; d < nd; ++d) out[deg_pos[d]] = deg_letters[d][digit[d]];
		lens[e] = (uint8_t)len;
		for (int d = 0; d < nd; ++d) { // next combination
			if (++digit[d] < deg_k[d]) break;
			digit[d] = 0;
		}
	}
}

} // namespace opt
} // namespace pcr

extern "C" int pcramp_gpu_optimize(pcramp_gpu_ctx *ctx, uint64_t *f, uint64_t *r, uint32_t n_trials, const int *moves, uint32_t n_moves,
	const pcramp_gpu_optimize_options *o, float *target_coverage, float *background_coverage, float *oligo_overlap, uint32_t *iterations)
{
	using namespace pcr::opt;
	if (!ctx) return 1;
	if (!o || (n_trials && (!f || !r)) || (n_moves && !moves)) return fail(ctx, "pcramp_gpu_optimize: null argument");
	if (fast_resolve(ctx)) return 1;
	CK(cudaSetDevice(ctx->device));
	SeqSet &T = ctx->sets[PCRAMP_TARGET], &B = ctx->sets[PCRAMP_BACKGROUND], &M = ctx->sets[PCRAMP_MULTIPLEX];
	if (!T.db_valid) return fail(ctx, "pcramp_gpu_optimize: no target database (call pcramp_gpu_select_words first)");
	// the multiplex terms (optimize.cpp:76-96): the key list of the multiplex background (pcramp_gpu_multiplex_keys) and the
	// assay pool (pcramp_gpu_set_pool); both may be empty, which is the first assay of a run
	const bool mpx = o->use_multiplex != 0;
	if (mpx && M.n && !ctx->mpx_valid)
		return fail(ctx, "pcramp_gpu_optimize: multiplex background sequences are loaded but their key list was not built "
		                 "(call pcramp_gpu_multiplex_keys)");
	const bool have_mkeys = mpx && ctx->mpx_valid && ctx->mpx_n_keys > 0;
	const bool have_pool = mpx && !ctx->pool_words.empty();
	auto bonus = [](float v) { return v == 1.0f ? 10.0f : v; }; // MULTIPLEX_OLIGO_REUSE_BONUS (assay.h:19)
	std::vector<float> best_f(n_trials, 0.0f), best_r(n_trials, 0.0f); // max over the pool of max_overlap(F / R of the trial's assay, .)
	for (uint32_t m = 0; m < n_moves; ++m)
		if (moves[m] < 0 || moves[m] >= MV_COUNT) return fail(ctx, ":optimization_move: Unknown move");
	const bool have_bg = B.db_valid && B.n_entries > 0; // collect_background_candidates skips an empty key list (assay.h:415)
	const float t_search = o->target_threshold * o->target_search_multiplier;       // assay.h:407 (float product)
	const float b_search = o->background_threshold * o->background_search_multiplier; // assay.h:418
	const uint32_t n = n_trials;
	std::vector<W128> af(n), ar(n), bf(n), br(n); // approx / best oligos
	std::vector<ScoreH> best(n), approx(n);
	std::vector<uint8_t> live(n, 1);
	std::vector<std::unordered_set<std::string>> previous(n);
	std::vector<uint32_t> iters(n, 0);
	std::vector<std::vector<uint8_t>> ring(n, std::vector<uint8_t>(nc::NC_MAX_LEN + 2, 0)); // the trial's NucCruc query buffer
	for (uint32_t t = 0; t < n; ++t) {
		af[t] = bf[t] = word_of(f + 2 * t);
		ar[t] = br[t] = word_of(r + 2 * t);
		previous[t].insert(packed(af[t]) + "|" + packed(ar[t]));
	}
	std::vector<uint64_t> pf, pr, vf, vr;
	std::vector<float> cov_t, cov_b, cov_m, ovl;
	std::vector<uint32_t> idx;
	std::vector<Variant> vars;
	std::vector<W128> tmp;
	std::vector<uint8_t> codes, lens;
	std::vector<float> strand, tm;
	uint64_t launches = 0;
	// PCRAMP_TRACE=1: host wall clock per stage, summed over the rounds
	const bool trace_on = getenv("PCRAMP_TRACE") != nullptr;
	double tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tr_last = trace_on ? Trace::now() : 0.0;
	auto lap = [&](int k) {
		if (!trace_on) return;
		const double t = Trace::now();
		tr_acc[k] += t - tr_last;
		tr_last = t;
	};
	uint32_t rounds = 0;
	uint64_t tr_vars = 0, tr_exp = 0, tr_hp = 0, tr_scored = 0;
	for (;;) {
		idx.clear();
		for (uint32_t t = 0; t < n; ++t)
			if (live[t]) idx.push_back(t);
		if (idx.empty()) break;
		++rounds;
		const uint32_t na = (uint32_t)idx.size();
		// ---- score the current assays (optimize.cpp:62-75) -------------------------------------------------
		pf.resize(2ull * na);
		pr.resize(2ull * na);
		for (uint32_t k = 0; k < na; ++k) {
			pf[2 * k] = af[idx[k]].hi; pf[2 * k + 1] = af[idx[k]].lo;
			pr[2 * k] = ar[idx[k]].hi; pr[2 * k + 1] = ar[idx[k]].lo;
		}
		cov_t.assign(na, 0.0f);
		cov_b.assign(na, 0.0f);
		if (pcramp_gpu_score_variants(ctx, PCRAMP_TARGET, pf.data(), pr.data(), pf.data(), pr.data(), na, t_search, o->target_threshold,
				o->target_amplicon_min, o->target_amplicon_max, o->use_taq_mama, cov_t.data(), nullptr)) return 1;
		launches += ctx->stats.kernel_launches;
		lap(7);
		if (have_bg) {
			if (pcramp_gpu_score_variants(ctx, PCRAMP_BACKGROUND, pf.data(), pr.data(), pf.data(), pr.data(), na, b_search, o->background_threshold,
					o->background_amplicon_min, o->background_amplicon_max, o->use_taq_mama, cov_b.data(), nullptr)) return 1;
			launches += ctx->stats.kernel_launches;
		}
		cov_m.assign(na, 0.0f);
		if (have_mkeys && pcr::mpx::coverage_run(ctx, pf.data(), pr.data(), pf.data(), pr.data(), na, o->background_threshold, o->use_taq_mama,
				cov_m.data(), launches)) return 1;
		if (have_pool) { // compute_oligo_overlap (pcr_assay.cpp:736-754)
			vf.resize(4ull * na);
			for (uint32_t k = 0; k < na; ++k) {
				vf[4 * k] = pf[2 * k]; vf[4 * k + 1] = pf[2 * k + 1];
				vf[4 * k + 2] = pr[2 * k]; vf[4 * k + 3] = pr[2 * k + 1];
			}
			ovl.assign(2ull * na, 0.0f);
			if (pcr::mpx::overlap_run(ctx, vf.data(), 2u * na, ovl.data(), launches)) return 1;
			for (uint32_t k = 0; k < na; ++k) {
				best_f[idx[k]] = ovl[2 * k];
				best_r[idx[k]] = ovl[2 * k + 1];
			}
		}
		lap(0);
		vars.clear();
		std::vector<uint32_t> gen; // the trials that go on to try their moves
		for (uint32_t k = 0; k < na; ++k) {
			const uint32_t t = idx[k];
			++iters[t];
			approx[t].target = cov_t[k];
			approx[t].background = cov_b[k];
			approx[t].overlap = 0.0f;
			if (mpx) { // optimize.cpp:76-96
				approx[t].background += cov_m[k];
				approx[t].overlap = bonus(best_f[t]) + bonus(best_r[t]);
			}
			if (approx[t].lt(best[t])) { // optimize.cpp:97-103
				live[t] = 0;
				continue;
			}
			best[t] = approx[t];
			bf[t] = af[t];
			br[t] = ar[t];
			gen.push_back(t);
		}
		// ---- the trial oligos of every move, FORWARD then REVERSE, in move-list order (optimize.cpp:117-139) -----
		{
			std::vector<Variant> part[PARALLEL_PARTS];
			bool bad[PARALLEL_PARTS] = {};
			parallel_ranges(gen.size(), 16, nullptr, [&](size_t which, size_t lo, size_t hi) {
				std::vector<W128> local;
				std::vector<Variant> &out = part[which];
				for (size_t g = lo; g < hi; ++g) {
					const uint32_t t = gen[g];
					for (int og = 0; og < 2; ++og)
						for (uint32_t m = 0; m < n_moves; ++m) {
							move_variants(moves[m], og ? ar[t] : af[t], o->degen, o->primer_min, o->primer_max, local);
							for (const W128 &w : local) {
								Variant v;
								v.trial = t;
								v.oligo = og;
								v.move = (int)m;
								v.w = w;
								const uint64_t c = expansion_count(w);
								if (c == 0 || c > (1ull << 20)) bad[which] = true;
								v.n_exp = (uint32_t)c;
								out.push_back(v);
							}
						}
				}
			});
			for (size_t k = 0; k < PARALLEL_PARTS; ++k) {
				if (bad[k]) return fail(ctx, "pcramp_gpu_optimize: empty or too degenerate trial oligo");
				vars.insert(vars.end(), part[k].begin(), part[k].end());
			}
		}
		// ---- PCR::is_valid without the dimer check (valid_pcr.cpp:5-45): duplex Tm of every expansion ... ---------------
		uint64_t n_exp = 0;
		for (Variant &v : vars) {
			v.first_exp = (uint32_t)n_exp;
			n_exp += v.n_exp;
		}
		if (n_exp > 0xffffffffull) return fail(ctx, "pcramp_gpu_optimize: too many trial oligos in one iteration");
		codes.resize((size_t)n_exp * 32); // every byte is written below
		lens.resize(n_exp);
		strand.resize(n_exp);
		tm.resize(n_exp);
		parallel_ranges(vars.size(), 2048, nullptr, [&](size_t, size_t lo, size_t hi) {
			for (size_t i = lo; i < hi; ++i) {
				const Variant &v = vars[i];
				const double dg = degeneracy(v.w);
				const float st = (float)((double)o->primer_strand / dg); // valid_pcr.cpp:13
				expansion_codes_all(v.w, v.n_exp, codes.data() + (size_t)v.first_exp * 32, lens.data() + v.first_exp);
				for (uint32_t e = 0; e < v.n_exp; ++e) strand[v.first_exp + e] = st;
			}
		});
		lap(1);
		tr_vars += vars.size();
		tr_exp += n_exp;
		if (nc::thermo_run_codes(ctx, nc::OP_PM_DUPLEX, (uint32_t)n_exp, codes.data(), lens.data(), strand.data(), o->salt, tm.data())) return 1;
		launches += 1;
		lap(2);
		// ... then the hairpin of the expansions whose duplex passed, with the stale slots of the trial's NucCruc object.
		// Replay per trial, in call order: an expansion is loaded (set_query) when its duplex Tm is in range.
		std::vector<uint32_t> hp_src; // expansion index of each hairpin problem
		std::vector<uint8_t> hp_codes, hp_lens;
		std::vector<float> hp_strand;
		{
			// how many hairpin problems a variant contributes: its expansions up to the first whose duplex Tm is out of range
			std::vector<uint32_t> hp_at(vars.size() + 1, 0);
			parallel_ranges(vars.size(), 4096, nullptr, [&](size_t, size_t lo, size_t hi) {
				for (size_t i = lo; i < hi; ++i) {
					Variant &v = vars[i];
					v.valid = true;
					uint32_t k = 0;
					for (; k < v.n_exp; ++k) {
						const float t_pm = tm[v.first_exp + k];
						if ((t_pm < o->primer_tm_min) || (t_pm > o->primer_tm_max)) { // valid_pcr.cpp:20-22: return false
							v.valid = false;
							break;
						}
					}
					hp_at[i + 1] = k;
				}
			});
			for (size_t i = 0; i < vars.size(); ++i) hp_at[i + 1] += hp_at[i];
			const size_t n_hp = hp_at[vars.size()];
			hp_src.resize(n_hp);
			hp_codes.resize(n_hp * 32);
			hp_lens.resize(n_hp);
			hp_strand.resize(n_hp);
			std::vector<size_t> cuts; // first variant of every trial: a trial's variants share its ring buffer and stay on one thread
			for (size_t i = 0; i < vars.size(); ++i)
				if (i == 0 || vars[i].trial != vars[i - 1].trial) cuts.push_back(i);
			parallel_ranges(vars.size(), 4096, &cuts, [&](size_t, size_t lo, size_t hi) {
				for (size_t i = lo; i < hi; ++i) {
					const Variant &v = vars[i];
					std::vector<uint8_t> &rg = ring[v.trial];
					const uint32_t n_load = hp_at[i + 1] - hp_at[i];
					for (uint32_t e = 0; e < n_load; ++e) {
						const uint32_t x = v.first_exp + e;
						const size_t at = (size_t)hp_at[i] + e;
						const int len = lens[x];
						uint8_t *slot = hp_codes.data() + at * 32;
						memcpy(slot, codes.data() + (size_t)x * 32, 32);
						for (int k = len; k < 32 && k < len + 2; ++k) slot[k] = rg[k]; // what set_query leaves behind past the new end
						for (int k = 0; k < len; ++k) rg[k] = slot[k];                  // set_query (nuc_cruc.h:875-913)
						hp_src[at] = x;
						hp_lens[at] = (uint8_t)len;
						hp_strand[at] = strand[x];
					}
				}
			});
		}
		lap(3);
		tr_hp += hp_src.size();
		std::vector<float> hp_tm(hp_src.size(), 0.0f);
		if (!hp_src.empty()) {
			if (nc::thermo_run_codes(ctx, nc::OP_HAIRPIN, (uint32_t)hp_src.size(), hp_codes.data(), hp_lens.data(), hp_strand.data(), o->salt,
					hp_tm.data())) return 1;
			launches += 1;
		}
		{
			std::vector<float> hp_of(n_exp, -1.0f);
			for (size_t k = 0; k < hp_src.size(); ++k) hp_of[hp_src[k]] = hp_tm[k];
			for (Variant &v : vars) {
				if (!v.valid) continue;
				for (uint32_t e = 0; e < v.n_exp; ++e)
					if (hp_of[v.first_exp + e] > o->max_hairpin) { v.valid = false; break; } // valid_pcr.cpp:28-30
			}
		}
		lap(4);
		// ---- coverage of the valid trial oligos against the unmoved assay's candidates ---------------------------------
		std::vector<uint32_t> vi;
		for (uint32_t k = 0; k < vars.size(); ++k)
			if (vars[k].valid) vi.push_back(k);
		const uint32_t nv = (uint32_t)vi.size();
		pf.resize(2ull * nv); pr.resize(2ull * nv); vf.resize(2ull * nv); vr.resize(2ull * nv);
		for (uint32_t k = 0; k < nv; ++k) {
			const Variant &v = vars[vi[k]];
			const W128 &F = af[v.trial], &R = ar[v.trial];
			pf[2 * k] = F.hi; pf[2 * k + 1] = F.lo;
			pr[2 * k] = R.hi; pr[2 * k + 1] = R.lo;
			const W128 &VF = v.oligo == 0 ? v.w : F, &VR = v.oligo == 1 ? v.w : R;
			vf[2 * k] = VF.hi; vf[2 * k + 1] = VF.lo;
			vr[2 * k] = VR.hi; vr[2 * k + 1] = VR.lo;
		}
		cov_t.assign(nv, 0.0f);
		cov_b.assign(nv, 0.0f);
		if (nv) {
			if (pcramp_gpu_score_variants(ctx, PCRAMP_TARGET, pf.data(), pr.data(), vf.data(), vr.data(), nv, t_search, o->target_threshold,
					o->target_amplicon_min, o->target_amplicon_max, o->use_taq_mama, cov_t.data(), nullptr)) return 1;
			launches += ctx->stats.kernel_launches;
			if (have_bg) {
				if (pcramp_gpu_score_variants(ctx, PCRAMP_BACKGROUND, pf.data(), pr.data(), vf.data(), vr.data(), nv, b_search,
						o->background_threshold, o->background_amplicon_min, o->background_amplicon_max, o->use_taq_mama, cov_b.data(), nullptr))
					return 1;
				launches += ctx->stats.kernel_launches;
			}
		}
		cov_m.assign(nv, 0.0f);
		ovl.assign(nv, 0.0f);
		if (nv && have_mkeys && pcr::mpx::coverage_run(ctx, pf.data(), pr.data(), vf.data(), vr.data(), nv, o->background_threshold,
				o->use_taq_mama, cov_m.data(), launches)) return 1;
		if (nv && have_pool) {
			std::vector<uint64_t> vw(2ull * nv);
			for (uint32_t k = 0; k < nv; ++k) {
				vw[2 * k] = vars[vi[k]].w.hi;
				vw[2 * k + 1] = vars[vi[k]].w.lo;
			}
			if (pcr::mpx::overlap_run(ctx, vw.data(), nv, ovl.data(), launches)) return 1;
		}
		for (uint32_t k = 0; k < nv; ++k) {
			vars[vi[k]].cov_t = cov_t[k];
			vars[vi[k]].cov_b = cov_b[k];
			vars[vi[k]].cov_m = cov_m[k];
			vars[vi[k]].ov = ovl[k];
		}
		lap(5);
		tr_scored += nv;
		// ---- selection, per trial, in the reference's order ------------------------------------------------------------
		size_t pos = 0;
		for (uint32_t k = 0; k < na; ++k) {
			const uint32_t t = idx[k];
			if (!live[t]) continue;
			bool improved = false;
			W128 local_seq;
			local_seq.hi = local_seq.lo = 0;
			int local_oligo = -1;
			ScoreH local = approx[t];
			for (int og = 0; og < 2; ++og)
				for (uint32_t m = 0; m < n_moves; ++m) {
					W128 ret_w;
					ret_w.hi = ret_w.lo = 0;
					ScoreH ret, trial; // Score(): -1e6 / 1e6 / 0
					for (; pos < vars.size() && vars[pos].trial == t && vars[pos].oligo == og && vars[pos].move == (int)m; ++pos) {
						const Variant &v = vars[pos];
						if (!v.valid) continue;
						trial.target = v.cov_t;
						const float bound = trial.target + local.background - local.target;
						if ((o->use_multiplex && bound < 0.0f) || (!o->use_multiplex && bound <= 0.0f)) continue; // background not evaluated
						trial.background = v.cov_b;
						if (mpx) {
							trial.background += v.cov_m; // compute_multiplex_background_coverage (pcr_assay.cpp:304-336)
							// the oligo that is not being moved contributes its own best overlap with the pool
							const float other = bonus(og == 0 ? best_r[t] : best_f[t]);
							// increase_degeneracy never resets trial_score.oligo_overlap (optimize_pcr.cpp:135-144; the other moves
							// do, :315,765,931, or start from fresh best_f / best_r, :466-491): the maximum carries over from the
							// previous trial oligo's TOTAL
							const float own = moves[m] == MV_INC_DEGEN ? std::max(trial.overlap, v.ov) : v.ov;
							trial.overlap = bonus(own) + other;
						}
						if (trial.gt(ret)) {
							ret = trial;
							ret_w = v.w;
						}
					}
					if (ret.gt(local) || (ret.eq(local) && degeneracy(ret_w) < degeneracy(local_seq))) { // optimize.cpp:127-137
						local = ret;
						local_seq = ret_w;
						local_oligo = og;
						improved = true;
					}
				}
			if (!improved) {
				live[t] = 0;
				continue;
			}
			approx[t] = local;
			local_seq = w_center(local_seq);
			if (local_oligo == 0) af[t] = local_seq; else ar[t] = local_seq;
			const std::string key = packed(af[t]) + "|" + packed(ar[t]);
			if (previous[t].count(key)) { // optimize.cpp:195-199
				live[t] = 0;
				continue;
			}
			previous[t].insert(key);
		}
		lap(6);
	}
	if (trace_on)
		fprintf(stderr, "[trace] optimize: %u rounds, %llu variants, %llu duplex + %llu hairpin problems, %llu variants scored; ms: score current %.1f, "
			"variants + expansions %.1f, duplex Tm %.1f, hairpin replay %.1f, hairpin Tm %.1f, score variants %.1f, selection %.1f (of score current: targets %.1f)\n", rounds,
			(unsigned long long)tr_vars, (unsigned long long)tr_exp, (unsigned long long)tr_hp, (unsigned long long)tr_scored, tr_acc[0] + tr_acc[7], tr_acc[1],
			tr_acc[2], tr_acc[3], tr_acc[4], tr_acc[5], tr_acc[6], tr_acc[7]);
	for (uint32_t t = 0; t < n; ++t) {
		f[2 * t] = bf[t].hi; f[2 * t + 1] = bf[t].lo;
		r[2 * t] = br[t].hi; r[2 * t + 1] = br[t].lo;
		if (target_coverage) target_coverage[t] = best[t].target;
		if (background_coverage) background_coverage[t] = best[t].background;
		if (oligo_overlap) oligo_overlap[t] = best[t].overlap;
		if (iterations) iterations[t] = iters[t];
	}
	ctx->stats.kernel_launches = launches;
	return 0;
}
