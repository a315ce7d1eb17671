// design_loop.cu -- one design iteration of pcramp (main.cpp:471-1130) driven through the C ABI of this library.
//
// The reference's main() loops "draw num_trial random assays -> index the backgrounds and targets against them -> optimize()
// every trial -> screen the survivors (multiplex compatibility, Smith-Waterman background matches) -> keep the best -> record
// its target matches / amplicons -> split the targets, grow the multiplex background, retire the detected targets".  Here the
// same iteration is a short sequence of BATCH calls over all trials (every one of them an entry point of include/pcramp_gpu.h,
// so this file adds no arithmetic of its own); what the reference decides trial by trial inside `#pragma omp critical`
// (the running best score gates which trials are screened at all) is replayed on the host in trial order from the batched
// results, which is exactly the reference's order at `--thread 1`.
//
// Host-only code; compiled into libpcramp_gpu.so next to the kernels it drives.
#include "ctx.cuh"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct ScoreH { // struct Score (pcramp.h:158-215)
	float target = -1.0e6f, background = 1.0e6f, overlap = 0.0f;
	float accuracy() const { return target - background; }
	bool lt(const ScoreH &o) const { return accuracy() == o.accuracy() ? overlap < o.overlap : accuracy() < o.accuracy(); }
	bool eq(const ScoreH &o) const { return accuracy() == o.accuracy() && overlap == o.overlap; }
};

double word_degeneracy(const uint64_t *w)
{ // Word::degeneracy (word.h:97-138)
	double d = 1.0;
	for (int limb = 0; limb < 2; ++limb)
		for (int k = 0; k < 16; ++k) {
			const unsigned b = (unsigned)((w[limb] >> (4 * k)) & 15ull);
			const int c = (int)((b & 1u) + ((b >> 1) & 1u) + ((b >> 2) & 1u) + (b >> 3));
			if (c) d *= c;
		}
	return d;
}

double now_ms()
{
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

struct pcramp_gpu_design {
	pcramp_gpu_ctx *ctx = nullptr;
	pcramp_gpu_design_options opt = {};
	unsigned int global_seed = 0;
	uint32_t iteration = 0, major_id = 1, minor_id = 1;
	std::vector<int> moves;
	std::vector<uint64_t> pool;                           // F0 R0 F1 R1 ... (2 x uint64 each): assay_pool (main.cpp:447)
	std::vector<std::vector<uint32_t>> pool_background;   // best_background_match of every accepted assay (main.cpp:448)
	std::vector<uint32_t> target_match, background_match; // LSB-first bitsets of the last iteration's best assay
	std::vector<uint64_t> trial_f, trial_r;               // scratch
	std::string err;
};

namespace {

int dfail(pcramp_gpu_design *d, const std::string &m)
{
	d->err = m;
	return 1;
}

// rc of a library call -> this object's error
#define DCALL(call)                                           \
	do {                                                      \
		if ((call) != 0) return dfail(d, pcramp_gpu_last_error(d->ctx)); \
	} while (0)

} // namespace

extern "C" {

void pcramp_gpu_design_default_options(pcramp_gpu_design_options *o)
{ // Options::Options (options.cpp:40-92), pcramp.h:14-51
	if (!o) return;
	memset(o, 0, sizeof(*o));
	o->num_trial = 1000;
	o->n_streams = 1;
	o->degen = 1;
	o->optimize_5 = o->optimize_3 = 0;
	o->primer_min = 18;
	o->primer_max = 25;
	o->primer_tm_min = 50.0f;
	o->primer_tm_max = 75.0f;
	o->primer_strand = 900.0e-9f;
	o->salt = 0.05f;
	o->max_hairpin = 40.0f;
	o->max_dimer = 40.0f;
	o->target_amplicon_min = 80;
	o->target_amplicon_max = 200;
	o->background_amplicon_min = 0;
	o->background_amplicon_max = 2000;
	o->target_threshold = 1.0f;
	o->background_threshold = 0.8f;
	o->target_search_multiplier = 0.9f;
	o->background_search_multiplier = 0.9f;
	o->min_target_cover = 0.0f;
	o->max_background_cover = 0.0f;
	o->pack_max_degen = 256;
	o->pack_min_gc = 0.0f;
	o->pack_max_gc = 1.0f;
	o->use_taq_mama = 0;
	o->use_multiplex = 1;
}

int pcramp_gpu_design_create(pcramp_gpu_ctx *ctx, const pcramp_gpu_design_options *options, uint32_t seed, pcramp_gpu_design **out)
{
	if (!ctx) return 1;
	if (!options || !out) return pcr::fail(ctx, "pcramp_gpu_design_create: null argument");
	if (options->num_trial == 0) return pcr::fail(ctx, "pcramp_gpu_design_create: num_trial must be >= 1");
	if (options->n_streams == 0) return pcr::fail(ctx, "pcramp_gpu_design_create: n_streams must be >= 1");
	if (ctx->parent) return pcr::fail(ctx, "pcramp_gpu_design_create: a worker context cannot change its parent's collections");
	pcramp_gpu_design *d = new pcramp_gpu_design();
	d->ctx = ctx;
	d->opt = *options;
	d->global_seed = seed; // main.cpp:112: unsigned int global_seed = opt.seed
	// the allowed moves in the reference's order (main.cpp:77-96; values of assay.h:21-29)
	if (options->degen > 1) { d->moves.push_back(0); d->moves.push_back(1); }
	if (options->optimize_5) { d->moves.push_back(2); d->moves.push_back(4); }
	if (options->optimize_3) { d->moves.push_back(3); d->moves.push_back(5); }
	*out = d;
	return 0;
}

void pcramp_gpu_design_destroy(pcramp_gpu_design *d) { delete d; }

const char *pcramp_gpu_design_last_error(const pcramp_gpu_design *d) { return d ? d->err.c_str() : "pcramp_gpu_design: null object"; }

int pcramp_gpu_design_iteration(pcramp_gpu_design *d, pcramp_gpu_design_result *res)
{
	if (!d || !res) return 1;
	pcramp_gpu_ctx *ctx = d->ctx;
	const pcramp_gpu_design_options &o = d->opt;
	memset(res, 0, sizeof(*res));
	pcr::SeqSet &T = ctx->sets[PCRAMP_TARGET], &B = ctx->sets[PCRAMP_BACKGROUND], &M = ctx->sets[PCRAMP_MULTIPLEX];
	const uint32_t n_target = T.n, n_background = B.n, n_trial = o.num_trial;
	const uint32_t t_words = (n_target + 31u) / 32u, b_words = (n_background + 31u) / 32u;
	const double t_start = now_ms();
	double t_mark = t_start;
	auto lap = [&](float &slot) {
		const double t = now_ms();
		slot += (float)(t - t_mark);
		t_mark = t;
	};
	++d->iteration;

	// ---- targets still to detect (main.cpp:475-502) ------------------------------------------------------------------------
	std::vector<uint8_t> active(T.active.begin(), T.active.end());
	uint32_t remaining = 0;
	for (uint32_t i = 0; i < n_target; ++i) remaining += active[i] ? 1u : 0u;
	if (remaining == 0) { // all targets detected: start over on all of them, next major id
		active.assign(n_target, 1);
		if (n_target) DCALL(pcramp_gpu_set_active(ctx, PCRAMP_TARGET, active.data()));
		remaining = n_target;
		++d->major_id;
		d->minor_id = 1;
	}
	res->iteration = d->iteration;
	res->major_id = d->major_id;
	res->minor_id = d->minor_id;
	res->targets_remaining = remaining;

	// ---- trial assays (main.cpp:523-558): one seed stream per OpenMP thread, static schedule ----------------------------------
	const uint32_t n_streams = o.n_streams;
	std::vector<uint32_t> seeds(n_streams), per(n_streams);
	for (uint32_t s = 0; s < n_streams; ++s) {
		seeds[s] = (uint32_t)rand_r(&d->global_seed); // :541, in thread order
		per[s] = n_trial / n_streams + (s < n_trial % n_streams ? 1u : 0u);
	}
	pcramp_gpu_random_assay_options ra;
	ra.primer_min = o.primer_min;
	ra.primer_max = o.primer_max;
	ra.amplicon_min = o.target_amplicon_min;
	ra.amplicon_max = o.target_amplicon_max;
	ra.degen = o.degen;
	ra.salt = o.salt;
	ra.primer_strand = o.primer_strand;
	ra.primer_tm_min = o.primer_tm_min;
	ra.primer_tm_max = o.primer_tm_max;
	ra.max_hairpin = o.max_hairpin;
	ra.max_dimer = o.max_dimer;
	std::vector<uint64_t> &f = d->trial_f, &r = d->trial_r;
	f.assign(2ull * n_trial, 0);
	r.assign(2ull * n_trial, 0);
	DCALL(pcramp_gpu_random_assays(ctx, PCRAMP_TARGET, n_streams, seeds.data(), per.data(), &ra, f.data(), r.data(), nullptr));
	lap(res->ms_candidates);

	// ---- word databases (main.cpp:560-691) ------------------------------------------------------------------------------------
	const uint32_t min_oligo = (uint32_t)std::max(0, o.primer_min); // Options::min_oligo_length (pcramp.h:136-139)
	uint32_t num_active_background = 0;
	float active_background_norm = 0.0f;
	for (uint32_t i = 0; i < n_background; ++i)
		if (B.active[i]) {
			++num_active_background;
			active_background_norm += B.weight[i];
		}
	if (n_background > 0) {
		uint64_t ne = 0;
		// :592-601: no G+C filter, words down to 90 % of the shortest primer, threshold * multiplier as a float product
		DCALL(pcramp_gpu_select_words(ctx, PCRAMP_BACKGROUND, f.data(), r.data(), n_trial, o.optimize_5, o.optimize_3,
			o.background_threshold * o.background_search_multiplier, o.pack_max_degen, 0.0f, 1.0f, (uint32_t)(min_oligo * 0.9), &ne, nullptr));
		res->n_background_entries = ne;
	}
	lap(res->ms_select_background);
	uint32_t num_active_target = 0;
	float active_target_norm = 0.0f;
	for (uint32_t i = 0; i < n_target; ++i)
		if (active[i]) {
			++num_active_target;
			active_target_norm += T.weight[i];
		}
	{
		uint64_t ne = 0;
		DCALL(pcramp_gpu_select_words(ctx, PCRAMP_TARGET, f.data(), r.data(), n_trial, o.optimize_5, o.optimize_3,
			o.target_threshold * o.target_search_multiplier, o.pack_max_degen, o.pack_min_gc, o.pack_max_gc, min_oligo, &ne, nullptr));
		res->n_target_entries = ne;
	}
	lap(res->ms_select_target);
	res->num_active_target = num_active_target;
	res->num_active_background = num_active_background;
	res->active_target_norm = active_target_norm;
	res->active_background_norm = active_background_norm;

	// ---- optimize() every trial (main.cpp:697-729) -------------------------------------------------------------------------
	pcramp_gpu_optimize_options oo;
	oo.target_threshold = o.target_threshold;
	oo.target_search_multiplier = o.target_search_multiplier;
	oo.target_amplicon_min = o.target_amplicon_min;
	oo.target_amplicon_max = o.target_amplicon_max;
	oo.background_threshold = o.background_threshold;
	oo.background_search_multiplier = o.background_search_multiplier;
	oo.background_amplicon_min = o.background_amplicon_min;
	oo.background_amplicon_max = o.background_amplicon_max;
	oo.use_taq_mama = o.use_taq_mama;
	oo.use_multiplex = o.use_multiplex;
	oo.degen = o.degen;
	oo.primer_min = o.primer_min;
	oo.primer_max = o.primer_max;
	oo.salt = o.salt;
	oo.primer_strand = o.primer_strand;
	oo.primer_tm_min = o.primer_tm_min;
	oo.primer_tm_max = o.primer_tm_max;
	oo.max_hairpin = o.max_hairpin;
	std::vector<float> tc(n_trial), bc(n_trial), ov(n_trial);
	DCALL(pcramp_gpu_optimize(ctx, f.data(), r.data(), n_trial, d->moves.empty() ? nullptr : d->moves.data(), (uint32_t)d->moves.size(), &oo,
		tc.data(), bc.data(), ov.data(), nullptr));
	lap(res->ms_optimize);

	// ---- screening (main.cpp:731-887): batched for every trial that survives the first filter, then replayed in trial order ---
	std::vector<uint32_t> cand; // :735-739
	for (uint32_t t = 0; t < n_trial; ++t)
		if (!((bc[t] > o.max_background_cover) || (tc[t] < o.min_target_cover))) cand.push_back(t);
	const uint32_t nc = (uint32_t)cand.size();
	std::vector<uint64_t> cf(2ull * nc), cr(2ull * nc);
	for (uint32_t k = 0; k < nc; ++k) {
		memcpy(&cf[2ull * k], &f[2ull * cand[k]], 16);
		memcpy(&cr[2ull * k], &r[2ull * cand[k]], 16);
	}
	const uint32_t n_pool = (uint32_t)(d->pool.size() / 4);
	std::vector<uint8_t> compat(nc, 1);
	std::vector<float> mbg_cov(nc, 0.0f), pool_amp_cov(nc, 0.0f), bg_cov(nc, 0.0f);
	std::vector<uint32_t> bg_bits;
	pcr::Trace trs("design screen", ctx->stream);
	if (nc && o.use_multiplex) {
		if (n_pool) { // :746-752
			std::vector<uint64_t> pf(2ull * n_pool), pr(2ull * n_pool);
			for (uint32_t i = 0; i < n_pool; ++i) {
				memcpy(&pf[2ull * i], &d->pool[4ull * i], 16);
				memcpy(&pr[2ull * i], &d->pool[4ull * i + 2], 16);
			}
			DCALL(pcramp_gpu_multiplex_compatible(ctx, cf.data(), cr.data(), nc, pf.data(), pr.data(), n_pool, o.salt, o.primer_strand, o.max_dimer,
				0, compat.data()));
			trs.mark("multiplex_compatible");
		}
		if (M.n) { // :760-771: weighted_coverage over the multiplex background (weights 1: a count)
			const uint32_t mw = (M.n + 31u) / 32u;
			std::vector<uint32_t> mb((size_t)nc * mw);
			DCALL(pcramp_gpu_multiplex_background_match(ctx, PCRAMP_MULTIPLEX, cf.data(), cr.data(), nc, o.background_threshold, o.use_taq_mama,
				mb.data()));
			for (uint32_t k = 0; k < nc; ++k) {
				double sum = 0.0;
				for (uint32_t i = 0; i < M.n; ++i)
					if ((mb[(size_t)k * mw + (i >> 5)] >> (i & 31u)) & 1u) sum += M.weight[i];
				mbg_cov[k] = (float)sum;
			}
			trs.mark("multiplex background");
		}
		if (n_pool) { // :783-803
			DCALL(pcramp_gpu_pool_amplicon_coverage(ctx, PCRAMP_TARGET, cf.data(), cr.data(), nc, o.target_threshold, o.target_amplicon_min,
				o.target_amplicon_max, o.background_threshold, o.use_taq_mama, pool_amp_cov.data()));
			trs.mark("pool x amplicons");
		}
	}
	if (nc && num_active_background > 0) { // :814-834
		bg_bits.assign((size_t)nc * b_words, 0);
		DCALL(pcramp_gpu_background_match(ctx, PCRAMP_BACKGROUND, cf.data(), cr.data(), nc, o.background_threshold * o.background_search_multiplier,
			o.background_threshold, o.background_amplicon_min, o.background_amplicon_max, o.use_taq_mama, bg_bits.data(), nullptr));
		for (uint32_t k = 0; k < nc; ++k) {
			double sum = 0.0;
			for (uint32_t i = 0; i < n_background; ++i)
				if ((bg_bits[(size_t)k * b_words + (i >> 5)] >> (i & 31u)) & 1u) sum += B.weight[i];
			bg_cov[k] = (float)sum;
		}
		trs.mark("find_background_match");
	}
	ScoreH best;
	int64_t best_trial = -1;
	double best_degeneracy = 2.0; // PCR().total_degeneracy(): two empty words
	for (uint32_t k = 0; k < nc; ++k) {
		const uint32_t t = cand[k];
		ScoreH s;
		s.target = tc[t];
		s.background = 0.0f; // :742
		s.overlap = ov[t];
		if (o.use_multiplex) {
			if (!compat[k]) continue; // :754-757
			if (best.lt(s)) {         // :759
				s.background += mbg_cov[k];
				if (s.background <= o.max_background_cover) s.background += pool_amp_cov[k]; // :782-803
			}
		}
		const double deg = word_degeneracy(&f[2ull * t]) + word_degeneracy(&r[2ull * t]);
		bool update = false;
		if (num_active_background > 0) { // :812-840
			if (best.lt(s) && s.background <= o.max_background_cover) {
				s.background += bg_cov[k];
				update = (s.background <= o.max_background_cover) && (best.lt(s) || (best.eq(s) && best_degeneracy > deg));
			}
		} else { // :842-856
			update = (s.background <= o.max_background_cover) && (best.lt(s) || (best.eq(s) && best_degeneracy > deg));
		}
		if (update) {
			best = s;
			best_trial = t;
			best_degeneracy = deg;
		}
	}
	lap(res->ms_screen);

	// ---- the best assay: target matches, amplicons (main.cpp:889-927) ----------------------------------------------------------
	d->target_match.assign(t_words, 0);
	d->background_match.assign(b_words, 0);
	res->target_coverage = best.target;
	res->background_coverage = best.background;
	res->oligo_overlap = best.overlap;
	if (!(best.target > 0.0f) || best_trial < 0) { // :928-932: nothing detected a single target -> the run ends
		res->found = 0;
		res->ms_total = (float)(now_ms() - t_start);
		return 0;
	}
	uint64_t bf[2], br[2];
	memcpy(bf, &f[2ull * best_trial], 16);
	memcpy(br, &r[2ull * best_trial], 16);
	memcpy(res->f, bf, 16);
	memcpy(res->r, br, 16);
	res->trial = (uint32_t)best_trial;
	res->degeneracy_f = word_degeneracy(bf);
	res->degeneracy_r = word_degeneracy(br);
	if (num_active_background > 0)
		for (uint32_t k = 0; k < nc; ++k)
			if (cand[k] == (uint32_t)best_trial) memcpy(d->background_match.data(), &bg_bits[(size_t)k * b_words], (size_t)b_words * 4);
	// find_target_match (pcr_assay.cpp:544-578): search = detect = opt.target_threshold
	pcr::Trace tr("design accept", ctx->stream);
	DCALL(pcramp_gpu_score_pairs(ctx, PCRAMP_TARGET, bf, br, 1, o.target_threshold, o.target_threshold, o.target_amplicon_min,
		o.target_amplicon_max, o.use_taq_mama, nullptr, d->target_match.data()));
	tr.mark("find_target_match");
	// PCR::write(fout, assay_pool) (assay.h:305-343): an oligo that is re-used from the pool is written in lower case
	for (uint32_t i = 0; i < 2 * n_pool; ++i) {
		if (pcramp_word_max_overlap(bf, &d->pool[2ull * i]) == 1.0f) res->reused_f = 1;
		if (pcramp_word_max_overlap(br, &d->pool[2ull * i]) == 1.0f) res->reused_r = 1;
	}
	if (o.use_multiplex) { // :918-921, :989-1017
		uint64_t n_amp = 0, n_bases = 0, n_bounds = 0, n_added = 0, n_keys = 0;
		DCALL(pcramp_gpu_unique_amplicons(ctx, PCRAMP_TARGET, bf, br, 1, o.target_threshold, o.target_amplicon_min, o.target_amplicon_max, 1, &n_amp,
			&n_bases, &n_bounds));
		tr.mark("unique amplicons");
		DCALL(pcramp_gpu_accept_assay(ctx, 0, o.pack_max_degen, min_oligo, &n_added, &n_keys)); // also: the assay joins the library's pool
		tr.mark("accept_assay");
		res->n_amplicons_added = n_added;
		res->n_splits = 3 * n_bounds;
		res->n_multiplex_keys = n_keys;
	}
	// :1116-1121: the detected targets are retired
	for (uint32_t i = 0; i < n_target; ++i)
		if ((d->target_match[i >> 5] >> (i & 31u)) & 1u) active[i] = 0;
	if (n_target) DCALL(pcramp_gpu_set_active(ctx, PCRAMP_TARGET, active.data()));
	tr.mark("set_active");
	d->pool.insert(d->pool.end(), bf, bf + 2); // :1123-1124
	d->pool.insert(d->pool.end(), br, br + 2);
	d->pool_background.push_back(d->background_match);
	res->found = 1;
	lap(res->ms_accept);
	if (pcr::AllocTrace::on())
		fprintf(stderr, "[trace] design iteration %u: %llu cudaMalloc / cudaFree calls so far, %.1f ms in them\n", res->iteration,
			(unsigned long long)pcr::AllocTrace::calls().load(), pcr::AllocTrace::micros().load() * 1e-3);
	res->ms_total = (float)(now_ms() - t_start);
	return 0;
}

int pcramp_gpu_design_matches(pcramp_gpu_design *d, uint32_t *target_bits, uint32_t *background_bits)
{
	if (!d) return 1;
	if (target_bits && !d->target_match.empty()) memcpy(target_bits, d->target_match.data(), d->target_match.size() * 4);
	if (background_bits && !d->background_match.empty()) memcpy(background_bits, d->background_match.data(), d->background_match.size() * 4);
	return 0;
}

int pcramp_gpu_design_active(pcramp_gpu_design *d, uint8_t *target_active, uint32_t *background_union)
{
	if (!d) return 1;
	const pcr::SeqSet &T = d->ctx->sets[PCRAMP_TARGET], &B = d->ctx->sets[PCRAMP_BACKGROUND];
	if (target_active && T.n) memcpy(target_active, T.active.data(), T.n);
	if (background_union) { // main.cpp:1148-1153: the union of every accepted assay's background matches
		const uint32_t bw = (B.n + 31u) / 32u;
		for (uint32_t w = 0; w < bw; ++w) background_union[w] = 0;
		for (const std::vector<uint32_t> &m : d->pool_background)
			for (uint32_t w = 0; w < bw && w < m.size(); ++w) background_union[w] |= m[w];
	}
	return 0;
}

} // extern "C"
