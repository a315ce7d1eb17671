// oracle/pcramp_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see pcramp_oracle.h).
//
// A deliberately plain CPU restatement of PCRamp's primer-pair scoring path.  Words are held as
// 32 explicit nibbles and every operation is a loop over positions; there are no bit tricks, no
// SIMD and no GPU-style reformulations here, so that this file is an independent check of the
// CUDA path in pcramp_b200/csrc (which works on bit-planes and never materialises most words).
// Each function cites the reference file:line whose behaviour it restates.  The restatement is
// pinned against the reference itself (oracle/_ref) by tests/test_oracle_vs_ref.py and by the
// vectors committed under tests/golden/.
//
// Floating point: the reference is built for x86-64 (SSE scalar float, no x87 excess precision,
// no FMA contraction at -O3 without -march), so plain float/double expressions compiled with
// -ffp-contract=off reproduce it bit for bit.
#include "pcramp_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <map>
#include <string>
#include <vector>

namespace {

enum { EOS = 0, BA = 1, BC = 2, BG = 4, BT = 8 };
enum { PLUS = 1, MINUS = 2 }; // sequence.h:27-32
const int WLEN = 32;          // word.h:692

// ---------------------------------------------------------------------------------------------
// Word: 32 nibbles, index 0 = most significant nibble of buffer[0] (word.h:290-297, word.cpp:13-19)
// ---------------------------------------------------------------------------------------------
struct OWord {
	uint8_t n[WLEN];
	OWord() { memset(n, 0, sizeof(n)); }
};

OWord from_u64(const uint64_t *p)
{
	OWord w;
	for (int i = 0; i < WLEN; ++i) w.n[i] = (uint8_t)((p[i / 16] >> ((15 - i % 16) * 4)) & 0xF);
	return w;
}

void to_u64(const OWord &w, uint64_t *p)
{
	p[0] = p[1] = 0;
	for (int i = 0; i < WLEN; ++i) p[i / 16] |= (uint64_t)w.n[i] << ((15 - i % 16) * 4);
}

// word.h:256-288
int w_start(const OWord &w)
{
	for (int i = 0; i < WLEN; ++i)
		if (w.n[i] != EOS) return i;
	return WLEN;
}
int w_stop(const OWord &w)
{
	for (int i = WLEN - 1; i >= 0; --i)
		if (w.n[i] != EOS) return i;
	return -1;
}
// word.cpp:198-213
unsigned w_size(const OWord &w)
{
	unsigned c = 0;
	for (int i = 0; i < WLEN; ++i) c += (w.n[i] != EOS);
	return c;
}
// word.cpp:67-196 (live body :151-154): positions whose nibble sets intersect
unsigned w_and(const OWord &a, const OWord &b)
{
	unsigned c = 0;
	for (int i = 0; i < WLEN; ++i) c += ((a.n[i] & b.n[i]) != 0);
	return c;
}
// word.cpp:215-231
void w_shift_left(OWord &w)
{
	for (int i = 0; i + 1 < WLEN; ++i) w.n[i] = w.n[i + 1];
	w.n[WLEN - 1] = EOS;
}
void w_shift_right(OWord &w)
{
	for (int i = WLEN - 1; i > 0; --i) w.n[i] = w.n[i - 1];
	w.n[0] = EOS;
}
// word.cpp:31-48
void w_push_back(OWord &w, uint8_t b)
{
	const int last = w_stop(w) + 1;
	if (last < WLEN) {
		w.n[last] = b;
		return;
	}
	w_shift_left(w);
	w.n[WLEN - 1] = b;
}
// word.h:392-418 -- note "right = max_size() - stop()" (32, not 31) and truncating division
void w_center(OWord &w)
{
	const int left = w_start(w);
	int right = w_stop(w);
	if (left > right) return;
	right = WLEN - right;
	const int delta = (right - left) / 2;
	if (delta > 0)
		for (int i = 0; i < delta; ++i) w_shift_right(w);
	else
		for (int i = 0; i > delta; --i) w_shift_left(w);
}
uint8_t comp_nibble(uint8_t b)
{
	uint8_t c = 0;
	if (b & BA) c |= BT;
	if (b & BT) c |= BA;
	if (b & BG) c |= BC;
	if (b & BC) c |= BG;
	return c;
}
// word.h:140-183: reverse complement written LEFT-justified from index 0
OWord w_complement(const OWord &w)
{
	OWord r;
	const int first = w_start(w), last = w_stop(w);
	int dest = 0;
	for (int src = last; src >= first; --src, ++dest) r.n[dest] = comp_nibble(w.n[src]);
	return r;
}
int popcount4(uint8_t b) { return (b & 1) + ((b >> 1) & 1) + ((b >> 2) & 1) + ((b >> 3) & 1); }
// word.h:97-138
double w_degeneracy(const OWord &w)
{
	double d = 1.0;
	// the reference multiplies limb by limb, byte by byte from the LSB, low nibble first; the
	// factors are small integers so the product is exact in any order.
	for (int i = 0; i < WLEN; ++i) {
		const int c = popcount4(w.n[i]);
		if (c != 0) d *= c;
	}
	return d;
}
bool w_less(const OWord &a, const OWord &b)
{ // word.h:197-211: lexicographic on (buffer[0], buffer[1]) == lexicographic on nibbles
	for (int i = 0; i < WLEN; ++i) {
		if (a.n[i] < b.n[i]) return true;
		if (a.n[i] > b.n[i]) return false;
	}
	return false;
}
bool w_equal(const OWord &a, const OWord &b) { return memcmp(a.n, b.n, WLEN) == 0; }
bool is_degen(uint8_t b) { return !(b == BA || b == BC || b == BG || b == BT); } // base_table.h:124-137

uint8_t base_to_bits(char c)
{ // base_table.h:31-76
	switch (c) {
	case 'A': case 'a': return 1;
	case 'C': case 'c': return 2;
	case 'G': case 'g': return 4;
	case 'T': case 't': case 'U': case 'u': return 8;
	case 'M': case 'm': return 1 | 2;
	case 'R': case 'r': return 4 | 1;
	case 'S': case 's': return 4 | 2;
	case 'V': case 'v': return 4 | 2 | 1;
	case 'W': case 'w': return 1 | 8;
	case 'Y': case 'y': return 8 | 2;
	case 'H': case 'h': return 1 | 2 | 8;
	case 'K': case 'k': return 4 | 8;
	case 'D': case 'd': return 4 | 1 | 8;
	case 'B': case 'b': return 4 | 8 | 2;
	case 'N': case 'n': case 'I': case 'i': case 'X': case 'x': return 15;
	default: return 0; // '-' (the reference throws on anything else)
	}
}

// word.cpp:233-294; Table 2 of Li et al., Genomics 83 (2004) 311-320.  Rows = template pair,
// columns = primer pair, both ordered {CC,GC,AC,TC,CG,GG,AG,TG,CA,GA,AA,TA,CT,GT,AT,TT}.
const float TAQ_MAMA[16][16] = {
	{1.000f, 0.968f, 0.947f, 1.034f, 0.547f, 0.253f, 0.230f, 0.359f, 0.606f, 0.282f, 0.372f, 0.347f, 0.957f, 0.382f, 0.399f, 0.687f},
	{0.989f, 1.000f, 1.023f, 1.000f, 0.420f, 0.662f, 0.445f, 0.367f, 0.870f, 0.512f, 0.492f, 0.508f, 0.372f, 1.000f, 0.492f, 0.714f},
	{1.011f, 1.000f, 1.000f, 1.000f, 0.459f, 0.277f, 0.570f, 0.343f, 0.927f, 0.362f, 0.590f, 0.542f, 0.439f, 0.488f, 0.978f, 0.662f},
	{1.000f, 0.907f, 1.000f, 1.000f, 0.382f, 0.234f, 0.228f, 0.542f, 0.763f, 0.309f, 0.410f, 0.473f, 0.426f, 0.347f, 0.423f, 0.947f},
	{0.590f, 0.334f, 0.445f, 0.323f, 1.000f, 0.978f, 0.927f, 0.989f, 0.907f, 0.645f, 0.525f, 0.455f, 0.927f, 0.408f, 0.408f, 0.707f},
	{0.327f, 0.595f, 0.319f, 0.396f, 0.947f, 1.000f, 0.978f, 0.989f, 0.405f, 0.861f, 0.681f, 0.512f, 0.410f, 0.968f, 0.452f, 0.714f},
	{0.410f, 0.420f, 0.590f, 0.311f, 1.023f, 1.000f, 1.000f, 1.000f, 0.488f, 0.898f, 0.907f, 0.566f, 0.442f, 0.449f, 0.989f, 0.707f},
	{0.423f, 0.343f, 0.305f, 0.585f, 1.034f, 0.879f, 0.927f, 1.000f, 0.473f, 0.720f, 0.547f, 0.957f, 0.459f, 0.374f, 0.459f, 1.023f},
	{1.023f, 0.429f, 0.473f, 0.477f, 1.023f, 0.466f, 0.420f, 0.477f, 1.000f, 0.978f, 0.907f, 0.978f, 0.907f, 0.380f, 0.525f, 0.669f},
	{0.442f, 1.046f, 0.455f, 0.470f, 0.432f, 1.058f, 0.481f, 0.485f, 0.917f, 1.000f, 1.023f, 1.023f, 0.336f, 0.968f, 0.534f, 0.639f},
	{0.617f, 0.452f, 1.011f, 0.439f, 0.492f, 0.504f, 0.978f, 0.462f, 0.989f, 0.947f, 1.000f, 0.978f, 0.405f, 0.405f, 0.888f, 0.606f},
	{0.601f, 0.377f, 0.377f, 1.046f, 0.500f, 0.399f, 0.408f, 1.034f, 0.978f, 0.720f, 0.870f, 1.000f, 0.402f, 0.313f, 0.651f, 0.927f},
	{0.978f, 0.462f, 0.466f, 0.488f, 0.420f, 0.239f, 0.225f, 0.336f, 0.504f, 0.269f, 0.319f, 0.656f, 1.000f, 0.835f, 0.907f, 1.034f},
	{0.429f, 1.011f, 0.473f, 0.477f, 0.340f, 0.413f, 0.357f, 0.354f, 0.352f, 0.538f, 0.413f, 0.794f, 0.927f, 1.000f, 1.058f, 1.000f},
	{0.595f, 0.492f, 0.968f, 0.485f, 0.367f, 0.282f, 0.388f, 0.439f, 0.413f, 0.309f, 0.566f, 0.917f, 0.957f, 0.957f, 1.000f, 0.989f},
	{0.590f, 0.380f, 0.410f, 0.968f, 0.364f, 0.223f, 0.230f, 0.416f, 0.321f, 0.239f, 0.301f, 0.645f, 0.978f, 0.714f, 0.947f, 1.000f}};

int taq_index(uint8_t b)
{ // word.cpp:233-247: order C,G,A,T
	switch (b) {
	case BC: return 0;
	case BG: return 1;
	case BA: return 2;
	case BT: return 3;
	}
	return -1;
}
float taq_mama(uint8_t p_pen, uint8_t p_last, uint8_t t_pen, uint8_t t_last)
{ // word.cpp:249-294
	const int a = taq_index(p_pen), b = taq_index(p_last), c = taq_index(t_pen), d = taq_index(t_last);
	if (a < 0 || b < 0 || c < 0 || d < 0) return 1.0f;
	const float v = TAQ_MAMA[4 * d + c][4 * b + a];
	return v < 1.0f ? v : 1.0f;
}

// ---------------------------------------------------------------------------------------------
// Sequences (sequence.h:85,223-243; sequence.cpp:304-330)
// ---------------------------------------------------------------------------------------------
struct OSeq {
	std::vector<uint8_t> nib; // one nibble per element (unpacked for clarity)
	float weight;
	bool active;
};

struct OEntry {
	OWord w;
	uint32_t index;
	int32_t loc;
	uint32_t strand;
};

bool entry_less(const OEntry &a, const OEntry &b)
{ // canonical total order used for comparisons (the reference's multimap order among equal keys is unspecified)
	if (w_less(a.w, b.w)) return true;
	if (w_less(b.w, a.w)) return false;
	if (a.index != b.index) return a.index < b.index;
	if (a.loc != b.loc) return a.loc < b.loc;
	return a.strand < b.strand;
}

// sequence.cpp:92-267 -- the sliding-window emitter, restated event by event.
void pack(const OSeq &s, uint32_t index, uint32_t degen_thr, float min_gc, float max_gc, uint32_t min_len,
	std::vector<OEntry> &out)
{
	OWord w;
	size_t size = 0; // curr_word_size: counts non-EOS pushes, NOT the number of bases in w
	const bool gc_filter = (min_gc > 0.0f) || (max_gc < 1.0f);
	std::deque<uint8_t> gc;
	unsigned num_gc = 0;
	const float norm = 1.0f / WLEN;
	// The reference walks BYTES (`iter != seq_buffer.end()`, sequence.cpp:111): a sequence of odd length
	// therefore also pushes the unused low nibble of its last byte, which Sequence::operator= left at
	// zero (sequence.cpp:23) -- i.e. one trailing EOS at raw index L.
	const size_t L = s.nib.size() + (s.nib.size() & 1);
	int loc = 1;

	auto emit = [&](const OWord &word, int l, uint32_t strand) {
		OEntry e;
		e.w = word;
		e.index = index;
		e.loc = l;
		e.strand = strand;
		out.push_back(e);
	};

	for (size_t i = 0; i < L; ++i, ++loc) {
		const uint8_t b = i < s.nib.size() ? s.nib[i] : (uint8_t)EOS;
		w_push_back(w, b);
		size += (b != EOS);
		if (gc_filter) { // :127-147 -- window of the last 32 RAW nibbles, EOS included, always /32
			if (gc.size() == (size_t)WLEN) {
				num_gc -= ((gc.front() & (BG | BC)) != 0);
				gc.pop_front();
			}
			gc.push_back(b);
			num_gc += ((b & (BG | BC)) != 0);
			const float fraction_gc = num_gc * norm;
			if (fraction_gc < min_gc || fraction_gc > max_gc) {
				size = std::min(size, (size_t)WLEN - 1);
				continue;
			}
		}
		if (w_degeneracy(w) > degen_thr) { // :149-153
			size = std::min(size, (size_t)WLEN - 1);
			continue;
		}
		if (size < (size_t)WLEN) { // :155-181 partial word, centred
			if (size >= min_len) {
				OWord tmp = w;
				w_center(tmp);
				emit(tmp, loc - (int)size - w_start(tmp), PLUS);
				tmp = w_complement(tmp);
				w_center(tmp);
				emit(tmp, loc - 1 + w_start(tmp), MINUS);
			}
		} else { // :182-194 full word
			emit(w, loc - (int)size, PLUS);
			emit(w_complement(w), loc - 1, MINUS);
			--size;
		}
	}
	// :198-263 tail: keep shifting left; loc is NOT advanced
	while (size > 0) {
		w_shift_left(w);
		--size;
		if (gc_filter) {
			if (gc.size() == (size_t)WLEN) {
				num_gc -= ((gc.front() & (BG | BC)) != 0);
				gc.pop_front();
			}
			const float fraction_gc = num_gc * norm;
			if (fraction_gc < min_gc || fraction_gc > max_gc) continue;
		}
		if (w_degeneracy(w) > degen_thr) continue;
		if (size >= min_len) {
			OWord tmp = w;
			w_center(tmp);
			emit(tmp, loc - 1 - (int)size - w_start(tmp), PLUS);
			tmp = w_complement(tmp);
			w_center(tmp);
			emit(tmp, loc - 2 + w_start(tmp), MINUS);
		}
	}
}

// sequence.cpp:304-330; returns -1 where the reference throws
int has_split(const OSeq &s, int loc, int len)
{
	if (loc < 0 || len < 0) return -1;
	if ((size_t)loc + (size_t)len > s.nib.size()) return -1;
	for (int i = 0; i < len; ++i)
		if (s.nib[loc + i] == EOS) return 1;
	return 0;
}

} // namespace

struct oracle_ctx {
	std::vector<OSeq> seq;
	std::vector<OEntry> db;   // canonical order
	std::vector<OWord> keys;  // sorted unique words of db (pcramp.h:231-256)
	std::vector<uint32_t> key_of_entry;
};

namespace {

void rebuild_keys(oracle_ctx *c)
{
	std::sort(c->db.begin(), c->db.end(), entry_less);
	c->keys.clear();
	c->key_of_entry.assign(c->db.size(), 0);
	for (size_t i = 0; i < c->db.size(); ++i) {
		if (c->keys.empty() || !w_equal(c->keys.back(), c->db[i].w)) c->keys.push_back(c->db[i].w);
		c->key_of_entry[i] = (uint32_t)(c->keys.size() - 1);
	}
}

// select_words.cpp:25-75: F and R of every trial, plus the 5'/3' shift family when enabled
void candidate_words(uint32_t n_pairs, const uint64_t *f, const uint64_t *r, bool opt5, bool opt3, std::vector<OWord> &out)
{
	for (uint32_t t = 0; t < n_pairs; ++t) {
		for (int o = 0; o < 2; ++o) {
			const OWord w = from_u64((o == 0 ? f : r) + 2 * t);
			out.push_back(w);
			if (opt5 || opt3) {
				const int start = w_start(w), stop = w_stop(w);
				if (opt5 && start > 0) {
					OWord tmp = w;
					for (int j = 0; j < start; ++j) {
						w_shift_left(tmp);
						out.push_back(tmp);
					}
				}
				if (opt3 && stop < WLEN - 1) {
					OWord tmp = w;
					for (int j = stop; j < WLEN - 1; ++j) {
						w_shift_right(tmp);
						out.push_back(tmp);
					}
				}
			}
		}
	}
}

// select_words.cpp:77-138 on ONE sequence's packed entries: for each candidate keep the best-scoring
// tier (>= threshold); the destination receives every entry of every surviving word.
void select_words_one(const std::vector<OEntry> &src, const std::vector<OWord> &cand, const std::vector<unsigned> &thr,
	std::vector<OEntry> &dst)
{
	if (src.empty() || cand.empty()) return;
	std::vector<uint8_t> keep(src.size(), 0);
	std::vector<unsigned> score(src.size());
	for (size_t i = 0; i < cand.size(); ++i) {
		unsigned best = thr[i];
		bool any = false;
		for (size_t j = 0; j < src.size(); ++j) {
			score[j] = w_and(cand[i], src[j].w);
			if (score[j] >= best) {
				best = score[j];
				any = true;
			}
		}
		if (!any) continue;
		for (size_t j = 0; j < src.size(); ++j)
			if (score[j] == best) keep[j] = 1;
	}
	for (size_t j = 0; j < src.size(); ++j)
		if (keep[j]) dst.push_back(src[j]);
}

struct OHit { // assay.h:34-62 OligoMatch
	uint32_t index;
	int32_t loc;
	uint32_t strand;
	uint32_t key;
	int oligo;
};

struct OAmp { // assay.h:120-164 PCROligos
	uint32_t index;
	float weight;
	uint32_t f, r;
};

int loc5(const OHit &h, int start, int stop) { return h.strand == PLUS ? h.loc + start : h.loc - stop; } // sequence.h:57-65
int loc3(const OHit &h, int start, int stop) { return h.strand == PLUS ? h.loc + stop : h.loc - start; } // sequence.h:67-75

// optimize.cpp:291-301
void match_words(std::vector<uint32_t> &m, const OWord &oligo, const std::vector<OWord> &keys, float threshold)
{
	const unsigned scaled = (unsigned)(w_size(oligo) * threshold);
	for (size_t k = 0; k < keys.size(); ++k)
		if (w_and(oligo, keys[k]) >= scaled) m.push_back((uint32_t)k);
}

// optimize.cpp:263-289
void find_oligo_match(std::vector<OHit> &out, const std::vector<uint32_t> &word_matches, int oligo, uint32_t strand,
	const oracle_ctx *c, const std::vector<std::pair<size_t, size_t>> &range_of_key)
{
	for (uint32_t k : word_matches) {
		for (size_t e = range_of_key[k].first; e < range_of_key[k].second; ++e) {
			const OEntry &en = c->db[e];
			if (!(en.strand & strand)) continue;
			if (!c->seq[en.index].active) continue;
			OHit h;
			h.index = en.index;
			h.loc = en.loc;
			h.strand = en.strand;
			h.key = k;
			h.oligo = oligo;
			out.push_back(h);
		}
	}
}

// pcr_assay.cpp:338-441
void find_amplicon_match(std::vector<OAmp> &amps, const std::vector<OHit> &m, int plus_oligo, int minus_oligo,
	const OWord oligo[2], const oracle_ctx *c, int amp_min, int amp_max)
{
	const int plus_start = w_start(oligo[plus_oligo]), plus_stop = w_stop(oligo[plus_oligo]);
	const int minus_start = w_start(oligo[minus_oligo]), minus_stop = w_stop(oligo[minus_oligo]);
	for (size_t p = 0; p < m.size(); ++p) {
		if (m[p].oligo != plus_oligo) continue;
		for (size_t q = p; q < m.size(); ++q) {
			if (m[p].index != m[q].index) break;
			if (m[q].oligo != minus_oligo) continue;
			if (loc3(m[p], plus_start, plus_stop) >= loc5(m[q], minus_start, minus_stop)) continue;
			int amp_start = loc5(m[p], plus_start, plus_stop);
			const int amp_stop = std::min(loc3(m[q], minus_start, minus_stop), (int)c->seq[m[p].index].nib.size() - 1);
			int amp_len = amp_stop - amp_start + 1;
			if (amp_len < amp_min) continue;
			if (amp_len > amp_max) break;
			if (amp_start < 0) {
				amp_len += amp_start;
				amp_start = 0;
			}
			if (has_split(c->seq[m[p].index], amp_start, amp_len) != 0) break; // (an exception in the reference would abort the run)
			OAmp a;
			a.index = m[p].index;
			a.weight = c->seq[m[p].index].weight;
			if (m[p].oligo == 0) {
				a.f = m[p].key;
				a.r = m[q].key;
			} else {
				a.f = m[q].key;
				a.r = m[p].key;
			}
			amps.push_back(a);
		}
	}
}

bool hit_less(const OHit &a, const OHit &b)
{ // assay.h:48-61
	if (a.index != b.index) return a.index < b.index;
	return a.loc < b.loc;
}

// optimize.cpp:209-261 for one key
float identity(const OWord &oligo, const OWord &key, bool use_taq)
{
	const unsigned len = w_size(oligo);
	const float norm = 1.0 / len; // double division narrowed to float, as in the reference
	float v = w_and(oligo, key) * norm;
	if (use_taq) {
		const int last = w_stop(oligo), pen = last - 1;
		const uint8_t p0 = oligo.n[pen], p1 = oligo.n[last];
		if (!is_degen(p0) && !is_degen(p1)) {
			const uint8_t t0 = key.n[pen], t1 = key.n[last];
			if (!is_degen(t0) && !is_degen(t1)) v *= taq_mama(p0, p1, t0, t1);
		}
	}
	return v;
}

} // namespace

extern "C" {

oracle_ctx *oracle_create(void) { return new oracle_ctx(); }
void oracle_destroy(oracle_ctx *c) { delete c; }

int oracle_set_sequences(oracle_ctx *c, uint32_t n, const uint8_t *nibbles, const uint64_t *byte_off, const uint32_t *len,
	const float *weight, const uint8_t *active)
{
	c->seq.assign(n, OSeq());
	for (uint32_t i = 0; i < n; ++i) {
		OSeq &s = c->seq[i];
		s.nib.resize(len[i]);
		for (uint32_t p = 0; p < len[i]; ++p) {
			const uint8_t b = nibbles[byte_off[i] + p / 2];
			s.nib[p] = (p % 2 == 1) ? (b & 0xF) : (b >> 4); // sequence.h:223-228
		}
		s.weight = weight ? weight[i] : 1.0f;
		s.active = active ? active[i] != 0 : true;
	}
	return 0;
}

int oracle_set_active(oracle_ctx *c, const uint8_t *active)
{
	for (size_t i = 0; i < c->seq.size(); ++i) c->seq[i].active = active[i] != 0;
	return 0;
}

int oracle_split_sequence(oracle_ctx *c, uint32_t seq, uint32_t pos)
{ // sequence.h:231-243
	c->seq[seq].nib[pos] = EOS;
	return 0;
}

long oracle_pack(oracle_ctx *c, uint32_t seq, uint32_t pack_max_degen, float min_gc, float max_gc, uint32_t min_len,
	uint64_t *words, uint32_t *index, int32_t *loc, uint32_t *strand)
{
	std::vector<OEntry> e;
	pack(c->seq[seq], seq, pack_max_degen, min_gc, max_gc, min_len, e);
	if (words) {
		for (size_t i = 0; i < e.size(); ++i) {
			to_u64(e[i].w, words + 2 * i);
			index[i] = e[i].index;
			loc[i] = e[i].loc;
			strand[i] = e[i].strand;
		}
	}
	return (long)e.size();
}

long oracle_select_words(oracle_ctx *c, uint32_t n_pairs, const uint64_t *f, const uint64_t *r, int opt5, int opt3,
	float threshold, uint32_t pack_max_degen, float min_gc, float max_gc, uint32_t min_len)
{
	std::vector<OWord> cand;
	candidate_words(n_pairs, f, r, opt5 != 0, opt3 != 0, cand);
	std::vector<unsigned> thr(cand.size());
	for (size_t i = 0; i < cand.size(); ++i) thr[i] = (unsigned)(w_size(cand[i]) * threshold); // select_words.cpp:83
	c->db.clear();
	for (uint32_t i = 0; i < c->seq.size(); ++i) {
		if (!c->seq[i].active) continue; // main.cpp:581,650
		std::vector<OEntry> local;
		pack(c->seq[i], i, pack_max_degen, min_gc, max_gc, min_len, local);
		select_words_one(local, cand, thr, c->db);
	}
	rebuild_keys(c);
	return (long)c->db.size();
}

long oracle_db_size(oracle_ctx *c) { return (long)c->db.size(); }
long oracle_num_keys(oracle_ctx *c) { return (long)c->keys.size(); }

void oracle_db_copy(oracle_ctx *c, uint64_t *words, uint32_t *index, int32_t *loc, uint32_t *strand)
{
	for (size_t i = 0; i < c->db.size(); ++i) {
		to_u64(c->db[i].w, words + 2 * i);
		index[i] = c->db[i].index;
		loc[i] = c->db[i].loc;
		strand[i] = c->db[i].strand;
	}
}

void oracle_keys_copy(oracle_ctx *c, uint64_t *words)
{
	for (size_t i = 0; i < c->keys.size(); ++i) to_u64(c->keys[i], words + 2 * i);
}

int oracle_db_set(oracle_ctx *c, long n, const uint64_t *words, const uint32_t *index, const int32_t *loc, const uint32_t *strand)
{
	c->db.resize(n);
	for (long i = 0; i < n; ++i) {
		c->db[i].w = from_u64(words + 2 * i);
		c->db[i].index = index[i];
		c->db[i].loc = loc[i];
		c->db[i].strand = strand[i];
	}
	rebuild_keys(c);
	return 0;
}

int oracle_score_pairs(oracle_ctx *c, uint32_t n_pairs, const uint64_t *f, const uint64_t *r, float search_threshold,
	float detect_threshold, int amp_min, int amp_max, int taq_mama_on, float *coverage, uint8_t *bits)
{
	const size_t n_seq = c->seq.size();
	// equal_range per key (read_only_multimap.h:104-117)
	std::vector<std::pair<size_t, size_t>> range(c->keys.size(), std::make_pair((size_t)0, (size_t)0));
	for (size_t e = 0; e < c->db.size(); ++e) {
		const uint32_t k = c->key_of_entry[e];
		if (range[k].second == 0) range[k].first = e;
		range[k].second = e + 1;
	}
	#pragma omp parallel for schedule(dynamic)
	for (uint32_t t = 0; t < n_pairs; ++t) {
		OWord oligo[2] = {from_u64(f + 2 * t), from_u64(r + 2 * t)};
		// pcr_assay.cpp:12-69
		std::vector<uint32_t> fm, rm;
		const float thr2 = search_threshold * search_threshold;
		match_words(fm, oligo[0], c->keys, thr2);
		match_words(rm, oligo[1], c->keys, thr2);
		std::vector<OAmp> amps;
		std::vector<OHit> hits;
		find_oligo_match(hits, fm, 0, PLUS, c, range);
		find_oligo_match(hits, rm, 1, MINUS, c, range);
		std::sort(hits.begin(), hits.end(), hit_less);
		find_amplicon_match(amps, hits, 0, 1, oligo, c, amp_min, amp_max);
		hits.clear();
		find_oligo_match(hits, fm, 0, MINUS, c, range);
		find_oligo_match(hits, rm, 1, PLUS, c, range);
		std::sort(hits.begin(), hits.end(), hit_less);
		find_amplicon_match(amps, hits, 1, 0, oligo, c, amp_min, amp_max);
		// pcr_assay.cpp:271-302 (+ :560-575 for the bitset)
		double cov = 0.0;
		std::vector<uint8_t> valid(n_seq, 0);
		for (const OAmp &a : amps) {
			const float fi = identity(oligo[0], c->keys[a.f], taq_mama_on != 0);
			const float ri = identity(oligo[1], c->keys[a.r], taq_mama_on != 0);
			const float local = sqrtf(fi * ri);
			if (local >= detect_threshold && !valid[a.index]) {
				valid[a.index] = 1;
				cov += a.weight;
			}
		}
		if (coverage) coverage[t] = (float)cov;
		if (bits) memcpy(bits + (size_t)t * n_seq, valid.data(), n_seq);
	}
	return 0;
}

void oracle_word_from_string(const char *s, int centre, uint64_t *out)
{ // word.h:229-245
	OWord w;
	const size_t len = strlen(s);
	for (size_t i = 0; i < len && i < (size_t)WLEN; ++i) w.n[i] = base_to_bits(s[i]);
	if (centre) w_center(w);
	to_u64(w, out);
}
uint32_t oracle_word_and(const uint64_t *a, const uint64_t *b) { return w_and(from_u64(a), from_u64(b)); }
uint32_t oracle_word_size(const uint64_t *a) { return w_size(from_u64(a)); }
int oracle_word_start(const uint64_t *a) { return w_start(from_u64(a)); }
int oracle_word_stop(const uint64_t *a) { return w_stop(from_u64(a)); }
double oracle_word_degeneracy(const uint64_t *a) { return w_degeneracy(from_u64(a)); }
void oracle_word_complement(const uint64_t *a, uint64_t *out) { to_u64(w_complement(from_u64(a)), out); }
void oracle_word_center(const uint64_t *a, uint64_t *out)
{
	OWord w = from_u64(a);
	w_center(w);
	to_u64(w, out);
}
void oracle_word_shift(const uint64_t *a, int left, uint64_t *out)
{
	OWord w = from_u64(a);
	if (left) w_shift_left(w); else w_shift_right(w);
	to_u64(w, out);
}
void oracle_word_push_back(const uint64_t *a, uint8_t b, uint64_t *out)
{
	OWord w = from_u64(a);
	w_push_back(w, b);
	to_u64(w, out);
}
float oracle_taq_mama(uint8_t p0, uint8_t p1, uint8_t t0, uint8_t t1) { return taq_mama(p0, p1, t0, t1); }

// word.h:525-647: begin() = lowest letter of every position; next() = odometer that runs over limb 0
// from its least significant nibble (position 15) up to position 0, then limb 1 from position 31 to 16.
long oracle_word_expand(const uint64_t *a, long cap, uint64_t *out)
{
	const OWord src = from_u64(a);
	OWord it;
	for (int i = 0; i < WLEN; ++i) {
		it.n[i] = 0;
		for (int k = 0; k < 4; ++k)
			if (src.n[i] & (1 << k)) {
				it.n[i] = (uint8_t)(1 << k);
				break;
			}
	}
	long n = 0;
	for (;;) {
		if (n < cap) to_u64(it, out + 2 * n);
		++n;
		bool advanced = false;
		for (int limb = 0; limb < 2 && !advanced; ++limb) {
			for (int p = limb * 16 + 15; p >= limb * 16; --p) {
				uint8_t nib = it.n[p];
				if (nib == EOS) continue;
				bool wrapped = false;
				do {
					if (nib == BT) {
						nib = BA;
						wrapped = true;
					} else {
						nib = (uint8_t)(nib << 1);
					}
				} while (!(nib & src.n[p]));
				it.n[p] = nib;
				if (!wrapped) {
					advanced = true;
					break;
				}
			}
		}
		if (!advanced) break;
	}
	return n;
}

int oracle_has_split(oracle_ctx *c, uint32_t seq, int loc, int len) { return has_split(c->seq[seq], loc, len); }

} // extern "C"
