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
//   PCR::is_valid / max_dimer_tm / multiplex_compatible    valid_pcr.cpp:5-45, pcr_assay.cpp:232-269,815-852
//   Word helpers                 word.h / word.cpp
#include <omp.h>
#include <stdint.h>
#include <algorithm>
#include <deque>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

// PCR::is_valid is a private member (assay.h:237); the driver needs to call it as optimize_pcr.cpp does.  Access
// specifiers do not change the class layout, so the objects compiled from the unmodified sources stay compatible.
#define private public
#include "assay.h"
#undef private
#include "seq_overlap.h"

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

// The multiplex background database of main.cpp:989-1003: every sequence pack()ed WHOLE into the database (running index, no
// select_words, no G+C filter), keys() recomputed.  Returns |db|.
long ref_pack_all(void *h, uint32_t pack_max_degen, uint32_t min_len)
{
	RefCtx *c = (RefCtx *)h;
	long n = -1;
	guarded(c, [&]() {
		c->db = MULTIMAP<Word, WordMatch>();
		for (uint32_t i = 0; i < c->seq.size(); ++i) c->seq[i].pack(c->db, i, pack_max_degen, 0.0, 1.0, min_len);
		c->db_keys = keys(c->db);
		n = (long)c->db.size();
	});
	return n;
}

// parse_fasta (parse_fasta.cpp:9-89) of the listed files into the context's sequences (appending, as main.cpp:255-268 does file
// after file).  `ignore`: n_ignore NUL-terminated lower-case strings, back to back.  Returns the number of sequences or -1.
long ref_parse_fasta(void *h, int n_files, const char *const *paths, uint64_t min_len, uint64_t max_len, int n_ignore, const char *ignore)
{
	RefCtx *c = (RefCtx *)h;
	long n = -1;
	guarded(c, [&]() {
		deque<string> ig;
		const char *p = ignore;
		for (int i = 0; i < n_ignore; ++i) {
			ig.push_back(string(p));
			p += ig.back().size() + 1;
		}
		c->seq.clear();
		for (int f = 0; f < n_files; ++f) parse_fasta(paths[f], c->seq, min_len, max_len, ig);
		n = (long)c->seq.size();
	});
	return n;
}

// length, weight and nibbles (one per byte, Sequence::operator[]) of sequence i; nibbles may be NULL
long ref_sequence_get(void *h, uint32_t i, float *weight, uint8_t *nibbles)
{
	RefCtx *c = (RefCtx *)h;
	if (i >= c->seq.size()) return -1;
	const Sequence &q = c->seq[i];
	if (weight) *weight = q.weight();
	if (nibbles)
		for (size_t k = 0; k < q.length(); ++k) nibbles[k] = q[(unsigned int)k];
	return (long)q.length();
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
//     3 approximate_tm_heterodimer(query=a, target=b) gapped, 4 same with fast_alignment(true),
//     5 approximate_tm_homodimer(a) with fast_alignment(true).
// out = {tm, dH, dS, dG, dG_dp}
// The reference reads one or two slots past the end of its query ring buffer when a trace reaches row 0
// (nuc_cruc.cpp:1375 with last_i == 0; circle_buffer.h:136-139 has no bounds check), i.e. whatever an earlier,
// longer query left there (or uninitialised memory).  To make its answers reproducible the driver first loads
// 256 x 'A' through the public set_query / set_target, so those slots hold base A.
static void thermo_one(NucCruc &melt, int op, const char *a, const char *b, float strand_a, float strand_b, float *out)
{
	static const std::string fill(MAX_SEQUENCE_LENGTH, 'A');
	melt.set_query(fill);
	melt.set_target(fill);
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
	case 5:
		melt.fast_alignment(op == 5);
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
}

int ref_thermo(int op, const char *a, const char *b, float salt, float strand_a, float strand_b, float *out)
{
	return guarded(NULL, [&]() {
		NucCruc melt;
		melt.salt(salt);
		thermo_one(melt, op, a, b, strand_a, strand_b, out);
	});
}

// n problems; a / b: strings of stride 33 bytes (NUL padded); strand: n x 2 floats; out: n x 5 floats.
// One NucCruc per OpenMP thread, as main.cpp:528-535 does.
int ref_thermo_batch(int op, int n, const char *a, const char *b, float salt, const float *strand, float *out)
{
	int bad = 0;
#pragma omp parallel
	{
		try {
			NucCruc melt;
			melt.salt(salt);
#pragma omp for schedule(dynamic, 64)
			for (int p = 0; p < n; ++p)
				thermo_one(melt, op, a + (size_t)p * 33, b ? b + (size_t)p * 33 : a + (size_t)p * 33, strand[2 * p], strand[2 * p + 1],
					out + 5 * (size_t)p);
		} catch (...) {
#pragma omp atomic write
			bad = 1;
		}
	}
	return bad ? -1 : 0;
}

// ---- the thermodynamic filters of PCR (valid_pcr.cpp:5-45, pcr_assay.cpp:232-269,815-852) ---------------
static void fill_melt(NucCruc &melt)
{
	static const std::string fill(MAX_SEQUENCE_LENGTH, 'A');
	melt.set_query(fill);
	melt.set_target(fill);
}

static Options thermo_options(float salt, float primer_strand, float tm_min, float tm_max, float max_hairpin, float max_dimer)
{
	Options opt;
	opt.salt = salt;
	opt.primer_strand = primer_strand;
	opt.primer_tm_range = make_pair(tm_min, tm_max);
	opt.max_hairpin = max_hairpin;
	opt.max_dimer = max_dimer;
	return opt;
}

int ref_is_valid(uint32_t n, const uint64_t *words, float salt, float primer_strand, float tm_min, float tm_max, float max_hairpin,
	float max_dimer, int check_homo_dimer, int fast_alignment, uint8_t *valid)
{
	return guarded(NULL, [&]() {
		const Options opt = thermo_options(salt, primer_strand, tm_min, tm_max, max_hairpin, max_dimer);
		NucCruc melt;
		melt.fast_alignment(fast_alignment != 0);
		melt.salt(salt);
		PCR p;
		for (uint32_t i = 0; i < n; ++i) {
			fill_melt(melt);
			valid[i] = p.is_valid(FORWARD, make_word(words + 2 * i), melt, opt, check_homo_dimer != 0) ? 1 : 0;
		}
	});
}

int ref_max_dimer_tm(uint32_t n, const uint64_t *f, const uint64_t *r, float salt, float primer_strand, int fast_alignment, float *tm)
{
	return guarded(NULL, [&]() {
		const Options opt = thermo_options(salt, primer_strand, 0.0f, 0.0f, 0.0f, 0.0f);
		NucCruc melt;
		melt.fast_alignment(fast_alignment != 0);
		melt.salt(salt);
		for (uint32_t i = 0; i < n; ++i) {
			PCR p;
			p.oligo(FORWARD, make_word(f + 2 * i));
			p.oligo(REVERSE, make_word(r + 2 * i));
			fill_melt(melt);
			tm[i] = p.max_dimer_tm(melt, opt);
		}
	});
}

// NOTE: multiplex_compatible loads FORWARD then REVERSE of the POOL assay as the query; when FORWARD is the longer
// one the reference's off-the-end read sees FORWARD's bases, not the 'A' fill -- callers keep len(F) <= len(R) in the pool.
int ref_multiplex_compatible(uint32_t n, const uint64_t *f, const uint64_t *r, uint32_t n_pool, const uint64_t *pf, const uint64_t *pr,
	float salt, float primer_strand, float max_dimer, int fast_alignment, uint8_t *ok)
{
	return guarded(NULL, [&]() {
		const Options opt = thermo_options(salt, primer_strand, 0.0f, 0.0f, 0.0f, max_dimer);
		NucCruc melt;
		melt.fast_alignment(fast_alignment != 0);
		melt.salt(salt);
		for (uint32_t i = 0; i < n; ++i) {
			PCR p;
			p.oligo(FORWARD, make_word(f + 2 * i));
			p.oligo(REVERSE, make_word(r + 2 * i));
			bool all = true;
			for (uint32_t j = 0; j < n_pool && all; ++j) {
				PCR q;
				q.oligo(FORWARD, make_word(pf + 2 * j));
				q.oligo(REVERSE, make_word(pr + 2 * j));
				fill_melt(melt);
				all = q.multiplex_compatible(melt, opt, p); // main.cpp:748-752: the pool assay is `this`
			}
			ok[i] = all ? 1 : 0;
		}
	});
}

// ---- SO::SeqOverlap (seq_overlap.h / seq_overlap.cpp:347-609), eight problems per align() as the reference runs it ----
// out: n x 6 int32 {score, query start, query stop, target start, target stop, (last_two.first << 4) | last_two.second}
int ref_sw_batch(uint32_t n, const uint64_t *query, const uint64_t *target, int32_t *out)
{
	return guarded(NULL, [&]() {
		SO::SeqOverlap align(SO::SeqOverlap::SmithWaterman, true);
		for (uint32_t base = 0; base < n; base += SO_LEN) {
			const uint32_t m = std::min<uint32_t>(SO_LEN, n - base);
			for (uint32_t s = 0; s < SO_LEN; ++s) { // unused slots repeat the last problem so that every slot holds valid data
				const uint32_t p = base + std::min(s, m - 1);
				align.pack_query_slots((unsigned char)(1u << s), make_word(query + 2 * p));
				align.pack_target_slots((unsigned char)(1u << s), make_word(target + 2 * p));
			}
			align.align();
			for (uint32_t s = 0; s < m; ++s) {
				int32_t *o = out + 6 * (size_t)(base + s);
				o[0] = align.score(s);
				const pair<int, int> q = align.alignment_range_query(s), t = align.alignment_range_target(s);
				o[1] = q.first; o[2] = q.second; o[3] = t.first; o[4] = t.second;
				const pair<unsigned char, unsigned char> l = align.target_last_two_aligned(s);
				o[5] = ((int)l.first << 4) | (int)l.second;
			}
		}
	});
}

// PCR::find_background_match (background_match.cpp:7-166) of every pair against the context's database and sequences
// n_candidates[t] (may be NULL) = |background_amplicons| of pair t.  When that count is odd and below the number of
// sequences the reference reads one element past the end of the list (background_match.cpp:122):
// the driver does not call the reference for such pairs (it was seen to crash) and reports their bits as 255.
int ref_background_match(void *h, uint32_t n_pairs, const uint64_t *f, const uint64_t *r, float background_threshold, float search_multiplier,
	int amp_min, int amp_max, int taq_mama, uint8_t *bits, uint32_t *n_candidates)
{
	RefCtx *c = (RefCtx *)h;
	Options opt;
	opt.background_threshold = background_threshold;
	opt.background_search_multiplier = search_multiplier;
	opt.background_amplicon_range = make_pair(amp_min, amp_max);
	opt.use_taq_mama = (taq_mama != 0);
	const size_t n_seq = c->seq.size();
	int fail = 0;
	#pragma omp parallel for schedule(dynamic)
	for (uint32_t t = 0; t < n_pairs; ++t) {
		try {
			PCR p;
			p.oligo(FORWARD, make_word(f + 2 * t));
			p.oligo(REVERSE, make_word(r + 2 * t));
			BitSet m(n_seq, false);
			ostringstream sink;
			p.collect_background_candidates(c->db_keys, c->db, c->seq, opt);
			const size_t n_amp = p.background_amplicons.size();
			if (n_candidates) n_candidates[t] = (uint32_t)n_amp;
			if ((n_amp & 1) && n_amp < n_seq) { // the reference would index background_amplicons[n_amp] (it crashes on some inputs)
				for (size_t i = 0; i < n_seq; ++i) bits[(size_t)t * n_seq + i] = 255;
				continue;
			}
			p.find_background_match(m, c->db_keys, c->db, c->seq, opt, sink);
			for (size_t i = 0; i < n_seq; ++i) bits[(size_t)t * n_seq + i] = m[i] ? 1 : 0;
		} catch (...) {
			#pragma omp critical
			fail = 1;
		}
	}
	if (fail) c->err = "exception inside ref_background_match";
	return fail;
}

// PCR::find_multiplex_background_match (background_match.cpp:168-295) against the context's sequences
int ref_multiplex_background_match(void *h, uint32_t n_pairs, const uint64_t *f, const uint64_t *r, float background_threshold, int taq_mama,
	uint8_t *bits)
{
	RefCtx *c = (RefCtx *)h;
	Options opt;
	opt.background_threshold = background_threshold;
	opt.use_taq_mama = (taq_mama != 0);
	const size_t n_seq = c->seq.size();
	int fail = 0;
	#pragma omp parallel for schedule(dynamic)
	for (uint32_t t = 0; t < n_pairs; ++t) {
		try {
			PCR p;
			p.oligo(FORWARD, make_word(f + 2 * t));
			p.oligo(REVERSE, make_word(r + 2 * t));
			BitSet m(n_seq, false);
			ostringstream sink;
			p.find_multiplex_background_match(m, c->seq, opt, sink);
			for (size_t i = 0; i < n_seq; ++i) bits[(size_t)t * n_seq + i] = m[i] ? 1 : 0;
		} catch (...) {
			#pragma omp critical
			fail = 1;
		}
	}
	if (fail) c->err = "exception inside ref_multiplex_background_match";
	return fail;
}

// ---- move scoring and the local search (optimize.cpp, optimize_pcr.cpp) -----------------------------------------------------
// What every move does to score a trial oligo (optimize_pcr.cpp: update_identity -> compute_coverage) against the candidate
// amplicons collect_candidates built for the unmoved assay.
int ref_score_variants(void *h, uint32_t n, const uint64_t *base_f, const uint64_t *base_r, const uint64_t *var_f, const uint64_t *var_r,
	float target_threshold, float search_multiplier, int amp_min, int amp_max, int taq_mama, float *coverage)
{
	RefCtx *c = (RefCtx *)h;
	Options opt;
	opt.target_threshold = target_threshold;
	opt.target_search_multiplier = search_multiplier;
	opt.target_amplicon_range = make_pair(amp_min, amp_max);
	opt.use_taq_mama = (taq_mama != 0);
	int fail = 0;
	#pragma omp parallel for schedule(dynamic)
	for (uint32_t t = 0; t < n; ++t) {
		try {
			PCR p;
			p.oligo(FORWARD, make_word(base_f + 2 * t));
			p.oligo(REVERSE, make_word(base_r + 2 * t));
			p.collect_target_candidates(c->db_keys, c->db, c->seq, opt);
			update_identity(p.target_f_identity, make_word(var_f + 2 * t), c->db_keys, opt.use_taq_mama);
			update_identity(p.target_r_identity, make_word(var_r + 2 * t), c->db_keys, opt.use_taq_mama);
			coverage[t] = p.compute_target_coverage(opt.target_threshold);
		} catch (...) {
			#pragma omp critical
			fail = 1;
		}
	}
	if (fail) c->err = "exception inside ref_score_variants";
	return fail;
}

// optimize() creates its own NucCruc on the stack (optimize.cpp:49) and the hairpin evaluation reads past the end of its
// query buffer (see thermo_one above), so the answer depends on what the stack held.  Zero the stack region below us first:
// the object then starts as if zero-filled (stale slots = base A), which is the convention the goldens are pinned to.
static void __attribute__((noinline)) paint_stack()
{
	volatile char buf[1 << 21];
	for (size_t i = 0; i < sizeof(buf); ++i) buf[i] = 0;
}

struct RefOptimizeOptions { // mirrors pcramp_gpu_optimize_options (include/pcramp_gpu.h)
	float target_threshold, target_search_multiplier;
	int target_amplicon_min, target_amplicon_max;
	float background_threshold, background_search_multiplier;
	int background_amplicon_min, background_amplicon_max;
	int use_taq_mama, use_multiplex;
	uint32_t degen;
	int primer_min, primer_max;
	float salt, primer_strand, primer_tm_min, primer_tm_max, max_hairpin;
};

// optimize() (optimize.cpp:14-207) of each trial, serially, against the target context h_t, (if not NULL) the background
// context h_b, (if not NULL) the multiplex background context h_m (ref_pack_all) and the assay pool.  f / r are updated in
// place; score: n x 3 floats.
int ref_optimize_multiplex(void *h_t, void *h_b, void *h_m, uint32_t n, uint64_t *f, uint64_t *r, const int *moves, uint32_t n_moves,
	const RefOptimizeOptions *o, uint32_t n_pool, const uint64_t *pool_f, const uint64_t *pool_r, float *score)
{
	RefCtx *ct = (RefCtx *)h_t, *cb = (RefCtx *)h_b, *cm = (RefCtx *)h_m;
	return guarded(ct, [&]() {
		Options opt;
		opt.target_threshold = o->target_threshold;
		opt.target_search_multiplier = o->target_search_multiplier;
		opt.target_amplicon_range = make_pair(o->target_amplicon_min, o->target_amplicon_max);
		opt.background_threshold = o->background_threshold;
		opt.background_search_multiplier = o->background_search_multiplier;
		opt.background_amplicon_range = make_pair(o->background_amplicon_min, o->background_amplicon_max);
		opt.use_taq_mama = (o->use_taq_mama != 0);
		opt.use_multiplex = (o->use_multiplex != 0);
		opt.degen = o->degen;
		opt.primer_range = make_pair(o->primer_min, o->primer_max);
		opt.salt = o->salt;
		opt.primer_strand = o->primer_strand;
		opt.primer_tm_range = make_pair(o->primer_tm_min, o->primer_tm_max);
		opt.max_hairpin = o->max_hairpin;
		opt.output_filter = Options::SILENT;
		vector<Move> mv;
		for (uint32_t i = 0; i < n_moves; ++i) mv.push_back((Move)moves[i]);
		const vector<Word> no_keys;
		const MULTIMAP<Word, WordMatch> no_db;
		const deque<Sequence> no_seq;
		deque<PCR> pool;
		for (uint32_t i = 0; i < n_pool; ++i) {
			PCR q;
			q.oligo(FORWARD, make_word(pool_f + 2 * i));
			q.oligo(REVERSE, make_word(pool_r + 2 * i));
			pool.push_back(q);
		}
		ostringstream sink;
		for (uint32_t t = 0; t < n; ++t) {
			PCR p;
			p.oligo(FORWARD, make_word(f + 2 * t));
			p.oligo(REVERSE, make_word(r + 2 * t));
			paint_stack();
			const Score s = optimize(p, mv, ct->db_keys, ct->db, ct->seq, cb ? cb->db_keys : no_keys, cb ? cb->db : no_db, cb ? cb->seq : no_seq,
				cm ? cm->db_keys : no_keys, cm ? cm->db : no_db, cm ? cm->seq : no_seq, pool, opt, sink);
			put_word(f + 2 * t, p.oligo(FORWARD));
			put_word(r + 2 * t, p.oligo(REVERSE));
			score[3 * t] = s.target_coverage;
			score[3 * t + 1] = s.background_coverage;
			score[3 * t + 2] = s.oligo_overlap;
		}
	});
}

int ref_optimize(void *h_t, void *h_b, uint32_t n, uint64_t *f, uint64_t *r, const int *moves, uint32_t n_moves, const RefOptimizeOptions *o,
	float *score)
{
	return ref_optimize_multiplex(h_t, h_b, NULL, n, f, r, moves, n_moves, o, 0, NULL, NULL, score);
}

// collect_multiplex_background_candidates (pcr_assay.cpp:71-104) for the base assay, update_multiplex_background_candidates
// (assay.h:449-453) with the trial oligos, compute_multiplex_background_coverage (pcr_assay.cpp:304-336); h = a ref_pack_all context
int ref_multiplex_coverage(void *h, uint32_t n, const uint64_t *base_f, const uint64_t *base_r, const uint64_t *var_f, const uint64_t *var_r,
	float background_threshold, int taq_mama, float *coverage)
{
	RefCtx *c = (RefCtx *)h;
	return guarded(c, [&]() {
		Options opt;
		opt.background_threshold = background_threshold;
		opt.use_taq_mama = (taq_mama != 0);
		for (uint32_t t = 0; t < n; ++t) {
			PCR p;
			p.oligo(FORWARD, make_word(base_f + 2 * t));
			p.oligo(REVERSE, make_word(base_r + 2 * t));
			p.collect_multiplex_background_candidates(c->db_keys, c->db, c->seq, opt);
			p.oligo(FORWARD, make_word(var_f + 2 * t));
			p.oligo(REVERSE, make_word(var_r + 2 * t));
			p.update_multiplex_background_candidates(c->db_keys, opt.use_taq_mama);
			coverage[t] = p.compute_multiplex_background_coverage(opt.background_threshold);
		}
	});
}

// PCR::compute_oligo_overlap (pcr_assay.cpp:736-754)
int ref_oligo_overlap(uint32_t n, const uint64_t *f, const uint64_t *r, uint32_t n_pool, const uint64_t *pool_f, const uint64_t *pool_r,
	float *overlap)
{
	deque<PCR> pool;
	for (uint32_t i = 0; i < n_pool; ++i) {
		PCR q;
		q.oligo(FORWARD, make_word(pool_f + 2 * i));
		q.oligo(REVERSE, make_word(pool_r + 2 * i));
		pool.push_back(q);
	}
	for (uint32_t t = 0; t < n; ++t) {
		PCR p;
		p.oligo(FORWARD, make_word(f + 2 * t));
		p.oligo(REVERSE, make_word(r + 2 * t));
		overlap[t] = p.compute_oligo_overlap(pool);
	}
	return 0;
}

// ---- multiplex bookkeeping ---------------------------------------------------------------------------------
// PCR::collect_unique_amplicons (pcr_assay.cpp:756-813) of one assay against the context's database and sequences.
// counts[0] = amplicons returned, counts[1] = their total length, counts[2] = AmpliconBounds pushed.  text_off / text / bounds
// (may be NULL: sizing call) receive the returned Sequences spelled with bits_to_base, and {index, begin, end} per bound.
int ref_unique_amplicons(void *h, const uint64_t *f, const uint64_t *r, float threshold, int amp_min, int amp_max, int want_bounds,
	uint64_t *counts, uint64_t *text_off, char *text, uint32_t *bounds)
{
	RefCtx *c = (RefCtx *)h;
	return guarded(c, [&]() {
		PCR p;
		p.oligo(FORWARD, make_word(f));
		p.oligo(REVERSE, make_word(r));
		deque<AmpliconBounds> b;
		const deque<Sequence> amp = p.collect_unique_amplicons(c->db_keys, c->db, c->seq, threshold, make_pair(amp_min, amp_max),
			want_bounds ? &b : NULL);
		uint64_t total = 0;
		for (size_t i = 0; i < amp.size(); ++i) {
			if (text_off) text_off[i] = total;
			if (text)
				for (unsigned int k = 0; k < amp[i].length(); ++k) text[total + k] = bits_to_base(amp[i][k]);
			total += amp[i].length();
		}
		if (text_off) text_off[amp.size()] = total;
		if (bounds)
			for (size_t i = 0; i < b.size(); ++i) {
				bounds[3 * i] = b[i].index;
				bounds[3 * i + 1] = b[i].begin;
				bounds[3 * i + 2] = b[i].end;
			}
		counts[0] = amp.size();
		counts[1] = total;
		counts[2] = b.size();
	});
}

// main.cpp:783-803 for every trial assay: its unique amplicons against every assay of the pool through
// PCR::find_multiplex_background_match (background_match.cpp:168-295), one bitset accumulated over the pool, then the sum
// of weighted_coverage (main.cpp:1402-1418; main.o is not linked: its five lines -- a double sum of Sequence::weight over the
// set bits, returned as float -- are spelled out here).
int ref_pool_amplicon_coverage(void *h, uint32_t n_pairs, const uint64_t *f, const uint64_t *r, uint32_t n_pool, const uint64_t *pool_f,
	const uint64_t *pool_r, float target_threshold, int amp_min, int amp_max, float background_threshold, int taq_mama, float *coverage)
{
	RefCtx *c = (RefCtx *)h;
	Options opt;
	opt.background_threshold = background_threshold;
	opt.use_taq_mama = (taq_mama != 0);
	deque<PCR> pool;
	for (uint32_t i = 0; i < n_pool; ++i) {
		PCR q;
		q.oligo(FORWARD, make_word(pool_f + 2 * i));
		q.oligo(REVERSE, make_word(pool_r + 2 * i));
		pool.push_back(q);
	}
	int fail = 0;
	#pragma omp parallel for schedule(dynamic)
	for (uint32_t t = 0; t < n_pairs; ++t) {
		try {
			PCR p;
			p.oligo(FORWARD, make_word(f + 2 * t));
			p.oligo(REVERSE, make_word(r + 2 * t));
			const deque<Sequence> amplicons = p.collect_unique_amplicons(c->db_keys, c->db, c->seq, target_threshold, make_pair(amp_min, amp_max));
			BitSet local(amplicons.size(), false);
			ostringstream sink;
			for (deque<PCR>::const_iterator i = pool.begin(); i != pool.end(); ++i) i->find_multiplex_background_match(local, amplicons, opt, sink);
			double ret = 0.0;
			for (size_t i = 0; i < amplicons.size(); ++i)
				if (local[i]) ret += amplicons[i].weight();
			coverage[t] = ret;
		} catch (...) {
			#pragma omp critical
			fail = 1;
		}
	}
	if (fail) c->err = "exception inside ref_pool_amplicon_coverage";
	return fail;
}

// main.cpp:989-1017 with the reference's own calls: the amplicons of the assay are appended to the multiplex context `h_m`
// (whose database / keys are rebuilt by pack() of every sequence, as ref_pack_all) and the sequences of `h` are split at begin,
// centre and end of every bound.  Returns the number of amplicons appended or -1.
long ref_accept_assay(void *h, void *h_m, const uint64_t *f, const uint64_t *r, float threshold, int amp_min, int amp_max, uint32_t pack_max_degen,
	uint32_t min_len)
{
	RefCtx *c = (RefCtx *)h, *m = (RefCtx *)h_m;
	long n = -1;
	guarded(c, [&]() {
		PCR p;
		p.oligo(FORWARD, make_word(f));
		p.oligo(REVERSE, make_word(r));
		deque<AmpliconBounds> bounds;
		const deque<Sequence> amplicons = p.collect_unique_amplicons(c->db_keys, c->db, c->seq, threshold, make_pair(amp_min, amp_max), &bounds);
		for (deque<Sequence>::const_iterator i = amplicons.begin(); i != amplicons.end(); ++i) {
			i->pack(m->db, m->seq.size(), pack_max_degen, 0.0, 1.0, min_len);
			m->seq.push_back(*i);
		}
		m->db_keys = keys(m->db);
		for (deque<AmpliconBounds>::const_iterator i = bounds.begin(); i != bounds.end(); ++i) {
			c->seq[i->index].split_sequence(i->begin);
			c->seq[i->index].split_sequence((i->begin + i->end) / 2);
			c->seq[i->index].split_sequence(i->end);
		}
		n = (long)amplicons.size();
	});
	return n;
}

// PCR::random_assay (pcr_assay.cpp:580-734) as ONE OpenMP thread of main.cpp:527-548 runs it: a fresh NucCruc object, the
// thread's local seed (advanced in place), n_trials consecutive trials.  attempts is not available from the reference.
int ref_random_assay_stream(void *h, uint32_t n_trials, uint32_t *seed, int primer_min, int primer_max, int amp_min, int amp_max, uint32_t degen,
	float salt, float primer_strand, float tm_min, float tm_max, float max_hairpin, float max_dimer, uint64_t *f, uint64_t *r)
{
	RefCtx *c = (RefCtx *)h;
	return guarded(c, [&]() {
		Options opt;
		opt.primer_range = make_pair(primer_min, primer_max);
		opt.target_amplicon_range = make_pair(amp_min, amp_max);
		opt.degen = degen;
		opt.salt = salt;
		opt.primer_strand = primer_strand;
		opt.primer_tm_range = make_pair(tm_min, tm_max);
		opt.max_hairpin = max_hairpin;
		opt.max_dimer = max_dimer;
		opt.output_filter = Options::SILENT;
		NucCruc melt;
		melt.salt(opt.salt);
		unsigned int local_seed = *seed;
		ostringstream sink;
		for (uint32_t t = 0; t < n_trials; ++t) {
			PCR p;
			p.random_assay(c->seq, melt, opt, local_seed, sink);
			put_word(f + 2 * t, p.oligo(FORWARD));
			put_word(r + 2 * t, p.oligo(REVERSE));
		}
		*seed = local_seed;
	});
}

// append_fasta_group (parse_fasta.cpp:91-169) driven as main.cpp:296-341 drives it: one (initially empty) Sequence per group, every
// file of the group appended to it, the Sequence dropped again when nothing was kept.  Replaces the context's sequences; returns
// their number or -1.
long ref_append_fasta_groups(void *h, int n_files, const char *const *paths, const uint32_t *file_group, uint64_t min_len, uint64_t max_len,
	uint64_t num_pad, int n_ignore, const char *ignore)
{
	RefCtx *c = (RefCtx *)h;
	long n = -1;
	guarded(c, [&]() {
		deque<string> ig;
		const char *p = ignore;
		for (int i = 0; i < n_ignore; ++i) {
			ig.push_back(string(p));
			p += ig.back().size() + 1;
		}
		c->seq.clear();
		int f = 0;
		while (f < n_files) {
			c->seq.push_back(Sequence());
			Sequence &ref = c->seq.back();
			ref.active(true);
			const uint32_t g = file_group[f];
			for (; f < n_files && file_group[f] == g; ++f) append_fasta_group(paths[f], ref, min_len, max_len, num_pad, ig);
			if (ref.empty()) c->seq.pop_back();
		}
		n = (long)c->seq.size();
	});
	return n;
}

// The best-assay update rule of main.cpp:829-858 (the branch without background sequences; the one with them, :812-827, applies the
// same test) folded over n scored trials in order, with the reference's own Score::operator< / == (pcramp.h:180-201) and
// PCR::total_degeneracy (assay.h:536-539) on the reference's own objects.  best = -1 when no trial was taken.
int ref_best_assay(uint32_t n, const float *target_cov, const float *background_cov, const float *overlap, const uint64_t *f, const uint64_t *r,
	float max_background_cover, int64_t *best, float *best_accuracy, float *best_overlap, double *best_degeneracy)
{
	return guarded(NULL, [&]() {
		PCR best_assay;
		Score best_score;
		int64_t bi = -1;
		for (uint32_t t = 0; t < n; ++t) {
			Score s;
			s.target_coverage = target_cov[t];
			s.background_coverage = background_cov[t];
			s.oligo_overlap = overlap ? overlap[t] : 0.0f;
			PCR trial;
			trial.oligo(FORWARD, make_word(f + 2 * t));
			trial.oligo(REVERSE, make_word(r + 2 * t));
			const bool update_best = (s.background_coverage <= max_background_cover) &&
				((best_score < s) || ((best_score == s) && (best_assay.total_degeneracy() > trial.total_degeneracy())));
			if (update_best) {
				best_score = s;
				best_assay.copy_oligos(trial);
				bi = t;
			}
		}
		*best = bi;
		if (bi >= 0) {
			*best_accuracy = best_score.accuracy();
			*best_overlap = best_score.oligo_overlap;
			*best_degeneracy = best_assay.total_degeneracy();
		}
	});
}

// The root's receive loop of reduce_best_assay (main.cpp:1455-1480) over the records of n_ranks ranks taken in rank order (rank 0 =
// the root's own best; MPI_ANY_SOURCE makes the real arrival order unspecified, rank order is one of them): a record replaces the
// running best unless `trial_score < m_score`, or `trial_score == m_score && m_assay.total_degeneracy() <= trial_degeneracy`.
// score: n_ranks x {target, background, overlap}.  -> index of the rank whose record is kept.
int ref_reduce_best(uint32_t n_ranks, const float *score, const double *degeneracy)
{
	Score m_score;
	m_score.target_coverage = score[0];
	m_score.background_coverage = score[1];
	m_score.oligo_overlap = score[2];
	double m_degeneracy = degeneracy[0];
	int owner = 0;
	for (uint32_t k = 1; k < n_ranks; ++k) {
		Score trial_score;
		trial_score.target_coverage = score[3 * k];
		trial_score.background_coverage = score[3 * k + 1];
		trial_score.oligo_overlap = score[3 * k + 2];
		if (trial_score < m_score) continue;
		if ((trial_score == m_score) && (m_degeneracy <= degeneracy[k])) continue;
		m_score = trial_score;
		m_degeneracy = degeneracy[k];
		owner = (int)k;
	}
	return owner;
}

} // extern "C"
