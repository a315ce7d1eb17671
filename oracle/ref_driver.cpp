// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
//
// A thin extern "C" surface over the UNMODIFIED reference sources in /root/reference, which
// are compiled where they lie by oracle/Makefile into oracle/_ref/libpcramp_ref.so together
// with this file.  Nothing here re-implements reference arithmetic: every entry point below
// only marshals plain arrays into the reference's own types and calls the reference's own
// functions, so that tests/ can pin both the C restatement (oracle/pcramp_oracle.cpp) and the
// CUDA path against outputs of the reference itself.
//
//   Sequence::pack               sequence.cpp:92-267
//   select_words                 select_words.cpp:8-139   (driven as main.cpp:644-691 does)
//   keys()                       pcramp.h:231-256
//   PCR::collect_target_candidates / update_target_candidates / compute_target_coverage
//                                assay.h:401-457, pcr_assay.cpp:12-69,271-302, optimize.cpp:209-261
//   PCR::find_target_match       pcr_assay.cpp:544-578
//   PCR::random_assay            pcr_assay.cpp:580-734
//   NucCruc front-ends           nuc_cruc.h:696-763, nuc_cruc.cpp:2236-2455
//   Word helpers                 word.h / word.cpp
#include "assay.h"
#include "seq_overlap.h"

#include <omp.h>
#include <stdint.h>
#include <sstream>
#include <string>
#include <vector>
#include <deque>

// Globals the reference expects main.cpp to define (main.cpp:33-34).
int mpi_numtasks = 1;
int mpi_rank = 0;

using namespace std;

namespace {

struct RefCtx {
	deque<Sequence> seq;
	MULTIMAP<Word, WordMatch> db;
	vector<Word> db_keys;
	string err;
};

inline Word make_word(const uint64_t *p)
{
	Word w;
	unsigned char tmp[16];
	memcpy(tmp, p, 16);
	w.mpi_unpack(tmp); // plain memcpy into buffer[0], buffer[1] (word.h:669-675)
	return w;
}

inline void put_word(uint64_t *p, const Word &w)
{
	unsigned char tmp[16];
	w.mpi_pack(tmp);
	memcpy(p, tmp, 16);
}

template <class F>
int guarded(RefCtx *c, F fn)
{
	try {
		fn();
		return 0;
	} catch (const char *e) {
		if (c) c->err = e;
		return 1;
	} catch (const string &e) {
		if (c) c->err = e;
		return 1;
	} catch (...) {
		if (c) c->err = "unknown exception";
		return 1;
	}
}

void dump_db(const MULTIMAP<Word, WordMatch> &db, uint64_t *words, uint32_t *index, int32_t *loc, uint32_t *strand)
{
	size_t n = 0;
	if (db.empty()) return;
	for (MULTIMAP<Word, WordMatch>::const_iterator i = db.begin(); i != db.end(); ++i, ++n) {
		put_word(words + 2 * n, i->first);
		index[n] = i->second.index;
		loc[n] = i->second.loc;
		strand[n] = (uint32_t)i->second.s;
	}
}

} // namespace

extern "C" {

void *ref_create() { return new RefCtx(); }
void ref_destroy(void *h) { delete (RefCtx *)h; }
const char *ref_last_error(void *h) { return ((RefCtx *)h)->err.c_str(); }
void ref_set_threads(int n) { omp_set_num_threads(n > 0 ? n : omp_get_num_procs()); }
int ref_max_threads() { return omp_get_max_threads(); }

// Sequences arrive as IUPAC text ('-' = EOS); Sequence::operator=(string) does the nibble packing
// (sequence.cpp:12-41).
int ref_set_sequences(void *h, uint32_t n, const char *text, const uint64_t *off, const uint32_t *len,
	const float *weight, const uint8_t *active)
{
	RefCtx *c = (RefCtx *)h;
	return guarded(c, [&]() {
		c->seq.clear();
		for (uint32_t i = 0; i < n; ++i) {
			c->seq.push_back(Sequence(string(text + off[i], len[i]), weight ? weight[i] : 1.0f));
			c->seq.back().active(active ? active[i] != 0 : true);
		}
	});
}

int ref_set_active(void *h, const uint8_t *active)
{
	RefCtx *c = (RefCtx *)h;
	for (size_t i = 0; i < c->seq.size(); ++i) c->seq[i].active(active[i] != 0);
	return 0;
}

int ref_split_sequence(void *h, uint32_t seq, uint32_t pos)
{
	RefCtx *c = (RefCtx *)h;
	c->seq[seq].split_sequence(pos);
	return 0;
}

// Sequence::pack of ONE sequence; returns the number of entries (call with NULL outputs to size).
long ref_pack(void *h, uint32_t seq, uint32_t pack_max_degen, float min_gc, float max_gc, uint32_t min_len,
	uint64_t *words, uint32_t *index, int32_t *loc, uint32_t *strand)
{
	RefCtx *c = (RefCtx *)h;
	long n = -1;
	guarded(c, [&]() {
		MULTIMAP<Word, WordMatch> local;
		c->seq[seq].pack(local, seq, pack_max_degen, min_gc, max_gc, min_len);
		n = (long)local.size();
		if (words) dump_db(local, words, index, loc, strand);
	});
	return n;
}

// The indexing loop of main.cpp:644-691 (identical in shape to :579-631 for backgrounds):
// for every active sequence pack -> select_words, then sort and keys().  Returns |db|.
long ref_select_words(void *h, uint32_t n_pairs, const uint64_t *f, const uint64_t *r, int opt5, int opt3,
	float threshold, uint32_t pack_max_degen, float min_gc, float max_gc, uint32_t min_len)
{
	RefCtx *c = (RefCtx *)h;
	long n = -1;
	guarded(c, [&]() {
		vector<PCR> trial(n_pairs);
		for (uint32_t t = 0; t < n_pairs; ++t) {
			trial[t].oligo(FORWARD, make_word(f + 2 * t));
			trial[t].oligo(REVERSE, make_word(r + 2 * t));
		}
		c->db = MULTIMAP<Word, WordMatch>();
		for (uint32_t i = 0; i < c->seq.size(); ++i) {
			if (!c->seq[i].active()) continue;
			MULTIMAP<Word, WordMatch> local;
			c->seq[i].pack(local, i, pack_max_degen, min_gc, max_gc, min_len);
			select_words(c->db, local, trial, opt5 != 0, opt3 != 0, threshold);
		}
		c->db.sort();
		c->db_keys = keys(c->db);
		n = (long)c->db.size();
	});
	return n;
}

long ref_db_size(void *h) { return (long)((RefCtx *)h)->db.size(); }
long ref_num_keys(void *h) { return (long)((RefCtx *)h)->db_keys.size(); }

void ref_db_copy(void *h, uint64_t *words, uint32_t *index, int32_t *loc, uint32_t *strand)
{
	dump_db(((RefCtx *)h)->db, words, index, loc, strand);
}

void ref_keys_copy(void *h, uint64_t *words)
{
	RefCtx *c = (RefCtx *)h;
	for (size_t i = 0; i < c->db_keys.size(); ++i) put_word(words + 2 * i, c->db_keys[i]);
}

// Install a DB directly (entries in any order) instead of running select_words.
int ref_db_set(void *h, long n, const uint64_t *words, const uint32_t *index, const int32_t *loc, const uint32_t *strand)
{
	RefCtx *c = (RefCtx *)h;
	return guarded(c, [&]() {
		c->db = MULTIMAP<Word, WordMatch>();
		for (long i = 0; i < n; ++i) {
			c->db.insert(make_pair(make_word(words + 2 * i), WordMatch(index[i], loc[i], (Strand)strand[i])));
		}
		c->db.sort();
		c->db_keys = keys(c->db);
	});
}

// Per pair, against the ctx DB:
//   coverage[t]  = what optimize() computes first (optimize.cpp:62-75): collect_target_candidates at
//                  target_threshold*target_search_multiplier, update_target_candidates,
//                  compute_target_coverage(target_threshold).
//   bits[t][i]   = PCR::find_target_match (pcr_assay.cpp:544-578) into a fresh BitSet, 1 byte per sequence.
// Parallel over pairs exactly as main.cpp:697-706 parallelises over trials.
int ref_score_pairs(void *h, uint32_t n_pairs, const uint64_t *f, const uint64_t *r, float target_threshold,
	float search_multiplier, int amp_min, int amp_max, int taq_mama, float *coverage, uint8_t *bits)
{
	RefCtx *c = (RefCtx *)h;
	Options opt;
	opt.target_threshold = target_threshold;
	opt.target_search_multiplier = search_multiplier;
	opt.target_amplicon_range = make_pair(amp_min, amp_max);
	opt.use_taq_mama = (taq_mama != 0);
	const size_t n_seq = c->seq.size();
	int fail = 0;
	#pragma omp parallel for schedule(dynamic)
	for (uint32_t t = 0; t < n_pairs; ++t) {
		try {
			PCR p;
			p.oligo(FORWARD, make_word(f + 2 * t));
			p.oligo(REVERSE, make_word(r + 2 * t));
			if (coverage) {
				p.collect_target_candidates(c->db_keys, c->db, c->seq, opt);
				p.update_target_candidates(c->db_keys, opt.use_taq_mama);
				coverage[t] = p.compute_target_coverage(opt.target_threshold);
			}
			if (bits) {
				BitSet m;
				p.find_target_match(m, c->db_keys, c->db, c->seq, opt);
				for (size_t i = 0; i < n_seq; ++i) bits[(size_t)t * n_seq + i] = m[i] ? 1 : 0;
			}
		} catch (...) {
			#pragma omp critical
			fail = 1;
		}
	}
	if (fail) c->err = "exception inside ref_score_pairs";
	return fail;
}

// PCR::random_assay with the thread-local seed protocol of main.cpp:537-548 collapsed to one
// thread (seed advances across trials).  Returns centred F/R words.
int ref_random_assays(void *h, uint32_t n_pairs, uint32_t seed, int primer_min, int primer_max, int amp_min,
	int amp_max, uint32_t degen, float salt, uint64_t *f, uint64_t *r)
{
	RefCtx *c = (RefCtx *)h;
	return guarded(c, [&]() {
		Options opt;
		opt.primer_range = make_pair(primer_min, primer_max);
		opt.target_amplicon_range = make_pair(amp_min, amp_max);
		opt.degen = degen;
		opt.salt = salt;
		opt.output_filter = Options::SILENT;
		NucCruc melt;
		melt.salt(opt.salt);
		unsigned int local_seed = seed;
		ostringstream sink;
		for (uint32_t t = 0; t < n_pairs; ++t) {
			PCR p;
			p.random_assay(c->seq, melt, opt, local_seed, sink);
			put_word(f + 2 * t, p.oligo(FORWARD));
			put_word(r + 2 * t, p.oligo(REVERSE));
		}
	});
}

// ---- Word helpers (word.h / word.cpp) -------------------------------------------------------
void ref_word_from_string(const char *s, int centre, uint64_t *out)
{
	Word w(s);
	if (centre) w.center();
	put_word(out, w);
}
uint32_t ref_word_and(const uint64_t *a, const uint64_t *b) { return make_word(a) & make_word(b); }
uint32_t ref_word_size(const uint64_t *a) { return make_word(a).size(); }
int ref_word_start(const uint64_t *a) { return make_word(a).start(); }
int ref_word_stop(const uint64_t *a) { return make_word(a).stop(); }
double ref_word_degeneracy(const uint64_t *a) { return make_word(a).degeneracy(); }
void ref_word_complement(const uint64_t *a, uint64_t *out) { put_word(out, make_word(a).complement()); }
void ref_word_center(const uint64_t *a, uint64_t *out) { Word w = make_word(a); w.center(); put_word(out, w); }
void ref_word_shift(const uint64_t *a, int left, uint64_t *out)
{
	Word w = make_word(a);
	if (left) w.shift_left(); else w.shift_right();
	put_word(out, w);
}
void ref_word_push_back(const uint64_t *a, uint8_t b, uint64_t *out) { Word w = make_word(a); w.push_back(b); put_word(out, w); }
float ref_word_max_overlap(const uint64_t *a, const uint64_t *b) { return make_word(a).max_overlap(make_word(b)); }
float ref_taq_mama(uint8_t p0, uint8_t p1, uint8_t t0, uint8_t t1)
{
	return taq_mama_correction(make_pair(p0, p1), make_pair(t0, t1));
}
// Word::begin()/next() expansion order (word.h:525-647); returns count, writes up to cap words.
long ref_word_expand(const uint64_t *a, long cap, uint64_t *out)
{
	const Word w = make_word(a);
	Word it = w.begin();
	long n = 0;
	do {
		if (n < cap) put_word(out + 2 * n, it);
		++n;
	} while (w.next(it));
	return n;
}
int ref_has_split(void *h, uint32_t seq, int loc, int len)
{
	RefCtx *c = (RefCtx *)h;
	int ret = -1;
	guarded(c, [&]() { ret = c->seq[seq].has_split(loc, len) ? 1 : 0; });
	return ret;
}

// ---- NucCruc front-ends (nuc_cruc.h:696-763) ------------------------------------------------
// op: 0 tm_pm_duplex(a), 1 approximate_tm_hairpin(a), 2 approximate_tm_homodimer(a),
//     3 approximate_tm_heterodimer(query=a, target=b) gapped, 4 same with fast_alignment(true).
// out = {tm, dH, dS, dG, dG_dp}
int ref_thermo(int op, const char *a, const char *b, float salt, float strand_a, float strand_b, float *out)
{
	return guarded(NULL, [&]() {
		NucCruc melt;
		melt.salt(salt);
		float tm = 0.0f;
		switch (op) {
		case 0:
			melt.strand(strand_a);
			tm = melt.tm_pm_duplex(a);
			break;
		case 1:
			melt.strand(strand_a);
			melt.set_query(a);
			tm = melt.approximate_tm_hairpin();
			break;
		case 2:
			melt.strand(strand_a);
			melt.set_query(a);
			tm = melt.approximate_tm_homodimer();
			break;
		case 3:
		case 4:
			melt.fast_alignment(op == 4);
			melt.strand(strand_a, strand_b);
			melt.set_query(a);
			melt.set_target(b);
			tm = melt.approximate_tm_heterodimer();
			break;
		default:
			throw "bad op";
		}
		out[0] = tm;
		out[1] = melt.delta_H();
		out[2] = melt.delta_S();
		out[3] = melt.delta_G();
		out[4] = melt.delta_G_dp();
	});
}

} // extern "C"
