/* oracle/pcramp_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C ABI of the CPU restatement (oracle/pcramp_oracle.cpp) of PCRamp's primer-pair scoring
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it;
 * the product (pcramp_b200/) never does.  Parity status: PINNED -- every function is checked
 * against outputs of the unmodified reference compiled into oracle/_ref (tests/test_oracle_vs_ref.py,
 * run in the dev container) and against the committed vectors in tests/golden/ generated from
 * that same build (tests/golden/make_golden.py).
 *
 * Words cross the ABI as two uint64 (buffer[0], buffer[1] of the reference's
 * __word<unsigned long,2>, word.h:12-17): nibble i lives in limb i/16 at bit (15 - i%16)*4.
 * Sequences cross as the reference's packed nibbles: two bases per byte, even index in the
 * high nibble (sequence.h:223-228), A=1 C=2 G=4 T=8, 0 = EOS (base_table.h:9-28).
 */
#ifndef PCRAMP_ORACLE_H
#define PCRAMP_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_ctx oracle_ctx;

oracle_ctx *oracle_create(void);
void oracle_destroy(oracle_ctx *);

/* nibbles: concatenated packed sequences; byte_off[i] = first byte of sequence i; len[i] in bases. */
int oracle_set_sequences(oracle_ctx *, uint32_t n, const uint8_t *nibbles, const uint64_t *byte_off,
	const uint32_t *len, const float *weight, const uint8_t *active);
int oracle_set_active(oracle_ctx *, const uint8_t *active);
int oracle_split_sequence(oracle_ctx *, uint32_t seq, uint32_t pos);

/* Sequence::pack (sequence.cpp:92-267) of one sequence, in emission order (unsorted). */
long oracle_pack(oracle_ctx *, uint32_t seq, uint32_t pack_max_degen, float min_gc, float max_gc,
	uint32_t min_len, uint64_t *words, uint32_t *index, int32_t *loc, uint32_t *strand);

/* main.cpp:644-691: per active sequence pack -> select_words (select_words.cpp:8-139); DB kept in ctx,
 * canonically ordered by (word, index, loc, strand).  Returns |DB|. */
long oracle_select_words(oracle_ctx *, uint32_t n_pairs, const uint64_t *f, const uint64_t *r, int opt5,
	int opt3, float threshold, uint32_t pack_max_degen, float min_gc, float max_gc, uint32_t min_len);
long oracle_db_size(oracle_ctx *);
long oracle_num_keys(oracle_ctx *);
void oracle_db_copy(oracle_ctx *, uint64_t *words, uint32_t *index, int32_t *loc, uint32_t *strand);
void oracle_keys_copy(oracle_ctx *, uint64_t *words);
int oracle_db_set(oracle_ctx *, long n, const uint64_t *words, const uint32_t *index, const int32_t *loc,
	const uint32_t *strand);

/* PCR::collect_candidates + update_identity + compute_coverage (pcr_assay.cpp:12-69,271-302,
 * optimize.cpp:209-301) with explicit thresholds.  coverage[t] float; bits[t*n_seq+i] in {0,1}. */
int oracle_score_pairs(oracle_ctx *, uint32_t n_pairs, const uint64_t *f, const uint64_t *r,
	float search_threshold, float detect_threshold, int amp_min, int amp_max, int taq_mama,
	float *coverage, uint8_t *bits);

/* Word helpers (word.h / word.cpp). */
void oracle_word_from_string(const char *s, int centre, uint64_t *out);
uint32_t oracle_word_and(const uint64_t *a, const uint64_t *b);
uint32_t oracle_word_size(const uint64_t *a);
int oracle_word_start(const uint64_t *a);
int oracle_word_stop(const uint64_t *a);
double oracle_word_degeneracy(const uint64_t *a);
void oracle_word_complement(const uint64_t *a, uint64_t *out);
void oracle_word_center(const uint64_t *a, uint64_t *out);
void oracle_word_shift(const uint64_t *a, int left, uint64_t *out);
void oracle_word_push_back(const uint64_t *a, uint8_t b, uint64_t *out);
float oracle_taq_mama(uint8_t p0, uint8_t p1, uint8_t t0, uint8_t t1);
long oracle_word_expand(const uint64_t *a, long cap, uint64_t *out);
int oracle_has_split(oracle_ctx *, uint32_t seq, int loc, int len);

#ifdef __cplusplus
}
#endif
#endif
