/* pcramp_gpu.h -- C ABI of the B200 (sm_100a) primer-pair scoring path for PCRamp.
 *
 * PCRamp (the reference) has no plugin or FFI layer: its hot path is a set of C++ free functions
 * and PCR methods called from main().  Each entry point below replaces one
 * of those internal seams with a batch call over plain host pointers; the citation beside it is the
 * reference code a maintainer would swap for the call (INTEGRATION.md shows the patch).
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; pcramp_gpu_last_error(ctx) gives the
 *    message (the reference throws const char* / std::string, main.cpp:1269-1292 -- the C++ wrapper
 *    in INTEGRATION.md rethrows it).
 *  - words are two uint64 {buffer[0], buffer[1]} of the reference's __word<unsigned long,2>
 *    (word.h:12-17): nibble i sits in limb i/16 at bit (15 - i%16)*4; A=1 C=2 G=4 T=8, 0 = EOS.
 *  - sequences are the reference's packed nibbles: two bases per byte, even index in the high
 *    nibble (sequence.h:223-228).
 *  - strand values: 1 = plus, 2 = minus (sequence.h:27-32).
 *  - bitsets are uint32 words, bit (i & 31) of word (i >> 5) = sequence i (LSB first).  The
 *    reference's MPI wire format is MSB-first bytes (mpi_util.cpp:152-233); convert at that edge only.
 *  - a ctx is bound to ONE GPU and is not thread-safe.  Multi-GPU = one process (rank) per GPU, each
 *    with its own ctx over its shard of the sequences.
 *  - there is no CPU fallback: without a CUDA device pcramp_gpu_create fails.
 */
#ifndef PCRAMP_GPU_H
#define PCRAMP_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pcramp_gpu_ctx pcramp_gpu_ctx;

/* Which sequence collection a call addresses (main.cpp:244-434 keeps three deques). */
enum { PCRAMP_TARGET = 0, PCRAMP_BACKGROUND = 1, PCRAMP_MULTIPLEX = 2, PCRAMP_NUM_KINDS = 3 };

/* ---- lifetime ------------------------------------------------------------------------------ */
int pcramp_gpu_create(pcramp_gpu_ctx **ctx, int device);
/* A context that still has live workers (below) is NOT destroyed (they read its device buffers in place): the call prints a message,
 * records it as the context's last error and returns; destroy the workers first. */
void pcramp_gpu_destroy(pcramp_gpu_ctx *ctx);
/* A worker: a second context on the parent's device that reads the parent's collections and text index in place (no copy:
 * 300 MB of planes + 9.6 GB of index for 20 000 x 30 kb stay single) and owns its stream, scratch, word database and results.
 * Independent batches -- the trial batches of a sweep, main.cpp:697-887 run for several batches at once -- are then driven from
 * one host thread per context and overlap on the GPU.  Calls that change a collection (upload, split, set_active, set_weights,
 * accept_assay) are refused on a worker; after such a call on the parent its workers fail every call ("sequences changed")
 * and must be re-created.  Destroy the workers before the parent. */
int pcramp_gpu_create_worker(pcramp_gpu_ctx *parent, pcramp_gpu_ctx **worker);
const char *pcramp_gpu_last_error(const pcramp_gpu_ctx *ctx);
/* cudaStream_t every kernel of this ctx is launched on (for CUDA-event timing by the caller). */
void *pcramp_gpu_stream(pcramp_gpu_ctx *ctx);
int pcramp_gpu_synchronize(pcramp_gpu_ctx *ctx);

/* ---- sequences: replaces the resident std::deque<Sequence> (main.cpp:244-434) ------------------- */
/* nibbles: concatenated packed sequences; byte_off[i] = first byte of sequence i; len[i] in bases.
 * weight may be NULL (1.0f each, sequence.h:22).  All sequences start active (sequence.h:143). */
int pcramp_gpu_upload_sequences(pcramp_gpu_ctx *ctx, int kind, uint32_t n, const uint8_t *nibbles,
	const uint64_t *byte_off, const uint32_t *len, const float *weight);
/* Sequence weights after the fact (Sequence::weight, e.g. opt.normalize_target_weight_per_file, main.cpp:268-278, applied to the
 * weights pcramp_gpu_fasta_records returned): n floats for the n sequences of `kind`. */
int pcramp_gpu_set_weights(pcramp_gpu_ctx *ctx, int kind, const float *weight);
/* Sequence::active(bool) for the whole collection (main.cpp:493-495,1116-1121). */
int pcramp_gpu_set_active(pcramp_gpu_ctx *ctx, int kind, const uint8_t *active);
/* Sequence::split_sequence (sequence.h:231-243; called at main.cpp:1010-1016). */
int pcramp_gpu_split_sequence(pcramp_gpu_ctx *ctx, int kind, uint32_t seq, uint32_t pos);
/* The same for n (sequence, position) pairs in one call: one rebuild of the device-side text for the whole list (the loop of
 * main.cpp:1008-1017 issues three splits per amplicon). */
int pcramp_gpu_split_sequences(pcramp_gpu_ctx *ctx, int kind, uint32_t n, const uint32_t *seq, const uint32_t *pos);

/* Sequence::pack (sequence.cpp:92-267) of ONE sequence: every (word, loc, strand) it would insert, in no
 * particular order.  The scan below never materialises this list; the call exists so that the device-side
 * model of pack() can be checked entry by entry.  Call with cap = 0 to size, then again with buffers. */
int pcramp_gpu_pack(pcramp_gpu_ctx *ctx, int kind, uint32_t seq, uint32_t pack_max_degen, float pack_min_gc,
	float pack_max_gc, uint32_t min_oligo_length, uint64_t cap, uint64_t *words, int32_t *loc, uint32_t *strand,
	uint64_t *n_out);

/* ---- seed scan: replaces the per-sequence Sequence::pack + select_words loops, the final sort()
 *      and keys() (main.cpp:579-631 and :644-691; sequence.cpp:92-267; select_words.cpp:8-139;
 *      pcramp.h:231-256).  f/r: n_pairs x 2 uint64 trial oligos (PCR::oligo(FORWARD/REVERSE)).
 *      threshold = opt.{target,background}_threshold * opt.*_search_multiplier (main.cpp:601,669).
 *      The resulting word database stays on the GPU for pcramp_gpu_score_pairs. ------------------- */
int pcramp_gpu_select_words(pcramp_gpu_ctx *ctx, int kind, const uint64_t *f, const uint64_t *r,
	uint32_t n_pairs, int optimize_5, int optimize_3, float threshold, uint32_t pack_max_degen,
	float pack_min_gc, float pack_max_gc, uint32_t min_oligo_length, uint64_t *n_entries, uint64_t *n_keys);
/* The fast form.  Batch after batch of one shape (same number of pairs, thresholds and pack parameters, no --optimize.5 / .3
 * families) is the normal load: a sweep, a design run, a sharded job.  After one batch of a shape has gone through the general
 * form with every pattern resolved by the one-part text index, the following ones are launched without the host reading anything
 * back -- sizes from the previous batch, counts in device memory, the partial-word scan on a second stream beside the indexed scan
 * -- and verified from ONE small read-back when the database is next needed (pair scoring, a copy, the statistics, n_entries /
 * n_keys != NULL here).  A batch whose assumptions did not hold (a buffer too small, a pattern the index cannot take) is run again
 * in the general form, so results are the same either way; option "use_fast_path" = 0 switches the fast form off.
 * n_keys may be NULL in the two calls above: the canonical order of the database and its keys() numbering are only
 * needed by a host that still walks the database itself; pair scoring on the GPU does not use them, and they are then
 * built on demand (by the calls below). */
int pcramp_gpu_db_size(pcramp_gpu_ctx *ctx, int kind, uint64_t *n_entries, uint64_t *n_keys);
/* Copy the database out, ordered by (word, index, loc, strand) -- the order read_only_multimap::sort()
 * (read_only_multimap.h:93-101) leaves, with its unspecified ties made canonical.  key_index[i] is the
 * position of entry i's word in keys() order.  Any output pointer may be NULL. */
int pcramp_gpu_db_copy(pcramp_gpu_ctx *ctx, int kind, uint64_t *words, uint32_t *index, int32_t *loc,
	uint32_t *strand, uint32_t *key_index);
int pcramp_gpu_keys_copy(pcramp_gpu_ctx *ctx, int kind, uint64_t *keys);

/* ---- pair scoring: replaces PCR::collect_candidates + update_identity + compute_coverage
 *      (pcr_assay.cpp:12-69,271-302,338-441; optimize.cpp:209-301) and PCR::find_target_match
 *      (pcr_assay.cpp:544-578) for a batch of pairs against the database built above.
 *        optimize()'s first score (optimize.cpp:62-75): search = threshold*multiplier, detect = threshold
 *        find_target_match:                              search = detect = opt.target_threshold
 *      coverage[t] (float, may be NULL) = sum of weights of detected sequences, accumulated in double
 *      in the reference's amplicon order; bitsets (may be NULL) = n_pairs x ceil(n_seq/32) words. --- */
int pcramp_gpu_score_pairs(pcramp_gpu_ctx *ctx, int kind, const uint64_t *f, const uint64_t *r,
	uint32_t n_pairs, float search_threshold, float detect_threshold, int amplicon_min, int amplicon_max,
	int use_taq_mama, float *coverage, uint32_t *bitsets);

/* The scoring step of an optimisation move (optimize_pcr.cpp:8-989; every one of the six moves is "mutate one oligo ->
 * is_valid -> update_identity -> compute_coverage", optimize.cpp:209-261, pcr_assay.cpp:271-302): trial assay
 * (var_f[i], var_r[i]) is scored against the candidate amplicons collect_candidates found for the UNMOVED assay
 * (base_f[i], base_r[i]) -- same database words, same amplicon geometry, identities of the trial oligos.  With
 * var == base this equals pcramp_gpu_score_pairs.  coverage / bitsets as there (either may be NULL). */
int pcramp_gpu_score_variants(pcramp_gpu_ctx *ctx, int kind, const uint64_t *base_f, const uint64_t *base_r,
	const uint64_t *var_f, const uint64_t *var_r, uint32_t n, float search_threshold, float detect_threshold,
	int amplicon_min, int amplicon_max, int use_taq_mama, float *coverage, uint32_t *bitsets);

/* optimize() (optimize.cpp:14-207) with its moves (optimize_pcr.cpp:8-989) for n_trials assays at once: what the
 * OpenMP trial loop of main.cpp:697-729 does one assay at a time.  f / r are updated in place to the best assay found
 * (m_assay.copy_oligos(best)); the three score arrays (any may be NULL) receive the returned Score (pcramp.h:158-215).
 * moves: the reference's list in its order (main.cpp:77-96): values 0..5 = IncreaseDegeneracy, DecreaseDegeneracy,
 * Trim5, Trim3, Grow5, Grow3 (assay.h:21-29).  Needs the TARGET database and uses the BACKGROUND database when one was
 * built.  With options->use_multiplex the multiplex terms (optimize.cpp:76-96 and the same terms inside every move) are
 * included: the key list of the multiplex background (pcramp_gpu_multiplex_keys) adds
 * compute_multiplex_background_coverage to the background coverage, and the assay pool (pcramp_gpu_set_pool) gives
 * Score::oligo_overlap; either may be empty (the first assay of a run). */
typedef struct pcramp_gpu_optimize_options {
	float target_threshold, target_search_multiplier;         /* opt.target_threshold, opt.target_search_multiplier */
	int target_amplicon_min, target_amplicon_max;             /* opt.target_amplicon_range */
	float background_threshold, background_search_multiplier; /* opt.background_threshold, opt.background_search_multiplier */
	int background_amplicon_min, background_amplicon_max;     /* opt.background_amplicon_range */
	int use_taq_mama;                                         /* opt.use_taq_mama */
	int use_multiplex;                                        /* opt.use_multiplex (hard-wired true, options.cpp:72) */
	uint32_t degen;                                           /* opt.degen */
	int primer_min, primer_max;                               /* opt.primer_range */
	float salt, primer_strand;                                /* opt.salt, opt.primer_strand */
	float primer_tm_min, primer_tm_max, max_hairpin;          /* opt.primer_tm_range, opt.max_hairpin */
} pcramp_gpu_optimize_options;
int pcramp_gpu_optimize(pcramp_gpu_ctx *ctx, uint64_t *f, uint64_t *r, uint32_t n_trials, const int *moves, uint32_t n_moves,
	const pcramp_gpu_optimize_options *options, float *target_coverage, float *background_coverage, float *oligo_overlap,
	uint32_t *iterations);

/* ---- FASTA ingest on the device (fasta.cuh): parse_fasta (parse_fasta.cpp:9-89) + Sequence::operator=(deque<char>)
 *      (sequence.cpp:43-71, base_to_bits base_table.h:30-76) + Sequence::defline / extract_weight (sequence.h:190-200,
 *      sequence.cpp:332-493) for the INFLATED text of n_files FASTA files (zlib stays with the host).  The host part splits the
 *      text into the reference's gzgets chunks and records (a chunk holding '>' is a defline) and evaluates deflines, weights
 *      ("[w=...]") and the ignore list (lower-case substrings, ignore_record parse_fasta.cpp:171-188); the device maps the residue
 *      characters to nibbles, drops white space and packs two bases per byte.  Records outside [min_length, max_length] or
 *      ignored are dropped as the reference does; an unknown symbol in a kept record is the reference's "Illegal base" error.
 *      The collection `kind` is replaced, exactly as if pcramp_gpu_upload_sequences had been called with the result. ------- */
int pcramp_gpu_upload_fasta(pcramp_gpu_ctx *ctx, int kind, uint32_t n_files, const char *const *text, const uint64_t *bytes,
	uint64_t min_length, uint64_t max_length, uint32_t n_ignore, const char *const *ignore, uint32_t *n_records);
/* The same for GROUPS of files -- append_fasta_group (parse_fasta.cpp:91-169) driven as main.cpp:296-341 / :386-431 drive it: the
 * files with the same file_group[f] (consecutive) form ONE sequence, the kept records in order with num_pad EOS between them
 * (target_group_padding = 1, main.cpp:212); the length window (the caller passes max(amplicon_min, length_min) as main.cpp:328 does)
 * and the ignore list apply per record; a group that keeps nothing gives no sequence.  Weights are 1 (a "[w=...]" in a group's name
 * is the caller's: pcramp_gpu_set_weights); pcramp_gpu_fasta_records reports, per sequence, the file and defline of its first
 * kept record, from which the caller recovers the group. */
int pcramp_gpu_upload_fasta_groups(pcramp_gpu_ctx *ctx, int kind, uint32_t n_files, const char *const *text, const uint64_t *bytes,
	const uint32_t *file_group, uint64_t min_length, uint64_t max_length, uint32_t num_pad, uint32_t n_ignore, const char *const *ignore,
	uint32_t *n_sequences);
/* per sequence of that upload: source file, defline as (offset, length) into the file's text, length, weight (any may be NULL) */
int pcramp_gpu_fasta_records(pcramp_gpu_ctx *ctx, int kind, uint32_t *file, uint64_t *defline_off, uint32_t *defline_len,
	uint32_t *length, float *weight);
/* CUDA-event times (ms) of the two device passes (count, pack) of the last FASTA upload, its text bytes and kept bases */
int pcramp_gpu_fasta_timing(pcramp_gpu_ctx *ctx, int kind, float *ms_count, float *ms_pack, uint64_t *text_bytes, uint64_t *n_bases);
void pcramp_gpu_fasta_free(pcramp_gpu_ctx *ctx);
/* host-only view of the record split of ONE file (before the length window / ignore list): returns the number of records and
 * fills at most cap entries of each non-NULL array */
uint32_t pcramp_fasta_scan(const char *text, uint64_t bytes, uint32_t cap, uint64_t *defline_off, uint32_t *defline_len,
	uint64_t *begin, uint64_t *end, float *weight);
/* a collection as the reference stores it (sequence.h:85,223-228): count, size of the nibble array, per sequence byte offset and
 * length, the packed nibbles (any output may be NULL) */
int pcramp_gpu_sequences_copy(pcramp_gpu_ctx *ctx, int kind, uint32_t *n, uint64_t *total_bytes, uint64_t *byte_off, uint32_t *length,
	uint8_t *nibbles);

/* ---- the multiplex terms of optimize() ---------------------------------------------------------------------
 * pcramp_gpu_multiplex_keys: main.cpp:989-1003 -- every sequence of the PCRAMP_MULTIPLEX collection (the amplicons of
 * the assays chosen so far) is pack()ed whole (no select_words, no G+C filter) and keys() (pcramp.h:231-256) of the
 * result is kept in HBM; n_keys (may be NULL) receives their number, keys_copy their words (n_keys x 2 uint64) in the
 * reference's key order.  Rebuild after every upload to PCRAMP_MULTIPLEX. */
int pcramp_gpu_multiplex_keys(pcramp_gpu_ctx *ctx, uint32_t pack_max_degen, uint32_t min_oligo_length, uint64_t *n_keys);
int pcramp_gpu_multiplex_keys_copy(pcramp_gpu_ctx *ctx, uint64_t *words);
/* The assay pool (`assay_pool`, main.cpp:744-752,958-985; m_pool of optimize()): the oligos of the assays already chosen. */
int pcramp_gpu_set_pool(pcramp_gpu_ctx *ctx, const uint64_t *pool_f, const uint64_t *pool_r, uint32_t n_pool);
/* collect_multiplex_background_candidates (pcr_assay.cpp:71-104) for the assay (base_f[i], base_r[i]), then
 * update_multiplex_background_candidates (assay.h:449-453) with the trial oligos (var_f[i], var_r[i]) and
 * compute_multiplex_background_coverage(threshold) (pcr_assay.cpp:304-336): the number of multiplex keys either oligo
 * reaches the threshold on.  threshold = opt.background_threshold. */
int pcramp_gpu_multiplex_coverage(pcramp_gpu_ctx *ctx, const uint64_t *base_f, const uint64_t *base_r, const uint64_t *var_f,
	const uint64_t *var_r, uint32_t n, float threshold, int use_taq_mama, float *coverage);
/* PCR::compute_oligo_overlap (pcr_assay.cpp:736-754) of n_pairs assays against the pool: Word::max_overlap
 * (word.h:38-92) of each oligo with every pool oligo, MULTIPLEX_OLIGO_REUSE_BONUS for an exact reuse. */
int pcramp_gpu_oligo_overlap(pcramp_gpu_ctx *ctx, const uint64_t *f, const uint64_t *r, uint32_t n_pairs, float *overlap);

/* ---- multiplex bookkeeping on the device ----------------------------------------------------------
 * pcramp_gpu_unique_amplicons: PCR::collect_unique_amplicons (pcr_assay.cpp:756-813) with extract_amplicon_seq
 * (:443-542) for n_pairs assays against the database pcramp_gpu_select_words built on `kind`: threshold =
 * opt.target_threshold (match_words at threshold^2), amplicon range = opt.target_amplicon_range.  The amplicons stay
 * in HBM as regions of the collection; n_amplicons = distinct amplicon strings summed over the pairs, n_bases = their
 * total length, n_bounds = candidate amplicons before the strings are made unique (one AmpliconBounds each, assay.h:64-90).
 * want_bounds != 0 is the call with m_bounds_ptr (main.cpp:920): a negative begin then fails with the reference's
 * ":AmpliconBounds(): Amplicon begin > amplicon end".  A region that leaves its sequence is undefined in the reference
 * (unchecked deque index, sequence.h:223-228) and is dropped here like a region holding an EOS.
 * pcramp_gpu_unique_amplicons_copy (any output may be NULL): pair_off[n_pairs + 1] = first amplicon of every pair;
 * text_off[n_amplicons + 1] / text = the strings (bits_to_base letters, no terminators) in the order of the returned
 * deque<Sequence> (sorted, unique) pair after pair; bounds_pair[n_bounds] / bounds[3 * n_bounds] = pair and
 * {index, begin, end} of every AmpliconBounds in the reference's push order, pair after pair. */
int pcramp_gpu_unique_amplicons(pcramp_gpu_ctx *ctx, int kind, const uint64_t *f, const uint64_t *r, uint32_t n_pairs, float threshold,
	int amplicon_min, int amplicon_max, int want_bounds, uint64_t *n_amplicons, uint64_t *n_bases, uint64_t *n_bounds);
int pcramp_gpu_unique_amplicons_copy(pcramp_gpu_ctx *ctx, uint32_t *pair_off, uint64_t *text_off, char *text, uint32_t *bounds_pair,
	uint32_t *bounds);
/* main.cpp:783-803 for n_pairs trial assays: collect_unique_amplicons (no bounds), then find_multiplex_background_match
 * (background_match.cpp:168-295) of EVERY assay of the pool (pcramp_gpu_set_pool / pcramp_gpu_accept_assay) against those
 * amplicons, accumulated in one bitset, and weighted_coverage (main.cpp:1402-1418) of it: coverage[i] = the number of
 * distinct amplicons of assay i that some pool primer (either strand) aligns to with a normalised Smith-Waterman score >=
 * background_threshold. */
int pcramp_gpu_pool_amplicon_coverage(pcramp_gpu_ctx *ctx, int kind, const uint64_t *f, const uint64_t *r, uint32_t n_pairs,
	float target_threshold, int amplicon_min, int amplicon_max, float background_threshold, int use_taq_mama, float *coverage);
/* main.cpp:989-1017 + :1123 for assay `pair` of the last pcramp_gpu_unique_amplicons(want_bounds = 1) call, nothing
 * leaving the device: its amplicons are appended to the PCRAMP_MULTIPLEX collection (index = its size so far, weight 1,
 * active), keys() of the multiplex background database is rebuilt (as pcramp_gpu_multiplex_keys: n_multiplex_keys), the
 * assay joins the pool, and the collection the amplicons were cut from is split at begin, (begin + end) / 2 and end of
 * every AmpliconBounds.  n_added = amplicons appended.  The word database of the split collection is dropped (the next
 * design iteration rebuilds it, main.cpp:574-631); marking the amplified targets inactive stays with the caller
 * (pcramp_gpu_set_active). */
int pcramp_gpu_accept_assay(pcramp_gpu_ctx *ctx, uint32_t pair, uint32_t pack_max_degen, uint32_t min_oligo_length, uint64_t *n_added,
	uint64_t *n_multiplex_keys);

/* The best assay of a batch of scored trials: the update rule of main.cpp:829-858 (Score::operator< / ==, pcramp.h:180-201: accuracy =
 * target - background coverage, then oligo_overlap; on an equal Score the smaller PCR::total_degeneracy(), assay.h:536-539; only
 * trials with background coverage <= max_background_cover compete) folded over the trials in order = the first maximum, found by a
 * device reduction.  best_index = -1 when nothing competes.  oligo_overlap may be NULL (no multiplex: 0).  Across ranks
 * (reduce_best_assay, main.cpp:1421-1601) the same rule is applied to the gathered per-rank winners (pcramp_b200/sharding.py
 * reduce_best in the Python mirror: all-gather of {accuracy, overlap, degeneracy, global trial index}, lowest index on ties). */
int pcramp_gpu_best_assay(pcramp_gpu_ctx *ctx, uint32_t n, const float *target_coverage, const float *background_coverage,
	const float *oligo_overlap, const uint64_t *f, const uint64_t *r, float max_background_cover, int64_t *best_index, float *best_accuracy,
	float *best_overlap, double *best_degeneracy);

/* ---- resident variants: the same two steps with the pairs already staged in HBM and the results
 *      left in HBM (what a multi-batch driver, the NCCL exchange and bench.py's `value` use). -------- */
int pcramp_gpu_stage_pairs(pcramp_gpu_ctx *ctx, const uint64_t *f, const uint64_t *r, uint32_t n_pairs);
/* Select the window [first, first+count) of the staged pairs as the current batch (a "design iteration" worth
 * of trials, main.cpp:523-558) without touching the host. */
int pcramp_gpu_set_batch(pcramp_gpu_ctx *ctx, uint32_t first, uint32_t count);
int pcramp_gpu_select_words_staged(pcramp_gpu_ctx *ctx, int kind, int optimize_5, int optimize_3, float threshold,
	uint32_t pack_max_degen, float pack_min_gc, float pack_max_gc, uint32_t min_oligo_length,
	uint64_t *n_entries, uint64_t *n_keys);
int pcramp_gpu_score_pairs_staged(pcramp_gpu_ctx *ctx, int kind, float search_threshold, float detect_threshold,
	int amplicon_min, int amplicon_max, int use_taq_mama);
/* device pointers of the last staged results: coverage float[n_pairs], bitsets uint32[n_pairs*words] */
void *pcramp_gpu_device_coverage(pcramp_gpu_ctx *ctx);
void *pcramp_gpu_device_bitsets(pcramp_gpu_ctx *ctx);
/* bitsets of the sequences detected by the {F(+), R(-)} pass alone (pcr_assay.cpp:37-47); with the bitsets above
 * they fix the order compute_coverage sums weights in, which the multi-GPU merge needs to stay bit-exact */
void *pcramp_gpu_device_bitsets_pass1(pcramp_gpu_ctx *ctx);
uint32_t pcramp_gpu_bitset_words(pcramp_gpu_ctx *ctx, int kind);
int pcramp_gpu_fetch_results(pcramp_gpu_ctx *ctx, float *coverage, uint32_t *bitsets);

/* ---- multi-GPU: replaces reduce_best_assay's gather of BitSets (main.cpp:1421-1601) --------------------
 * Sequences are sharded across ranks in contiguous index ranges; every rank scores the same pairs on its
 * shard.  After the caller has all-gathered the shards' device bitsets (NCCL, rank-major), this splices
 * them into global bitsets (n_pairs x ceil(sum(shard_nseq)/32)) and re-sums the coverage over all
 * sequences in the reference's order.  All d_* are device pointers; shard_nseq/weight_all are host. */
int pcramp_gpu_merge_shards(pcramp_gpu_ctx *ctx, const void *d_any_gathered, const void *d_pass1_gathered,
	uint32_t n_shards, const uint32_t *shard_nseq, const float *weight_all, uint32_t n_pairs, void *d_out_bits,
	void *d_out_cov);

/* The same exchange without NCCL and without leaving the library's stream (xchg.cuh): every rank owns a double-buffered
 * global result (n_pairs x ceil(N/32) words) that its peers address over NVLink.  After pair scoring, one kernel stores the
 * shard's bitset words into every rank's buffer and publishes a step flag, one waits for all ranks' flags, and the coverage
 * kernels run on the merged bitsets.  Shards: contiguous, every shard but the last a multiple of 32 sequences.
 *   create   once per run (rank, world <= 16, shard_nseq[world], largest batch, weights of ALL sequences or NULL = 1.0)
 *   buffer / ipc_handle   this rank's buffer as a device pointer (contexts of one process) or a 64-byte cudaIpcMemHandle_t
 *   connect  peers = world pointers (from_ipc = 0) or world x 64 bytes of handles gathered from all ranks (from_ipc = 1)
 *   step     after pcramp_gpu_score_pairs_staged on this rank's shard; asynchronous
 *   coverage / bitsets / words   device pointers + row pitch of the merged result;  fetch = host copies (synchronises, and
 *            reports a rank that did not arrive within the timeout instead of hanging) */
int pcramp_gpu_exchange_create(pcramp_gpu_ctx *ctx, uint32_t rank, uint32_t world, const uint32_t *shard_nseq, uint32_t max_pairs,
	const float *weight_all);
void *pcramp_gpu_exchange_buffer(pcramp_gpu_ctx *ctx);
int pcramp_gpu_exchange_ipc_handle(pcramp_gpu_ctx *ctx, void *handle64);
int pcramp_gpu_exchange_connect(pcramp_gpu_ctx *ctx, const void *peers, int from_ipc);
int pcramp_gpu_exchange_step(pcramp_gpu_ctx *ctx, int kind);
void *pcramp_gpu_exchange_coverage(pcramp_gpu_ctx *ctx);
void *pcramp_gpu_exchange_bitsets(pcramp_gpu_ctx *ctx);
uint32_t pcramp_gpu_exchange_words(pcramp_gpu_ctx *ctx);
int pcramp_gpu_exchange_fetch(pcramp_gpu_ctx *ctx, float *coverage, uint32_t *bitsets);
/* rows (pairs) of the last step: what pcramp_gpu_exchange_fetch copies (size the host arrays from this, not from memory) */
uint32_t pcramp_gpu_exchange_pairs(pcramp_gpu_ctx *ctx);
/* Failure model: a rank that never arrives is detected after a bound (default 120 s of GPU time; ranks legitimately skew by
 * seconds -- first-step index build, allocation growth, rank-0 I/O).  The step's coverage is then poisoned (NaN) for consumers of
 * the device pointers, _status / _fetch report the missing ranks, and every later _step / _reduce_best fails until the exchange
 * has been destroyed on ALL ranks and created again (_create refuses while one exists; _connect refuses a second call and
 * undoes a partial one).  Ranks must not share a device. */
int pcramp_gpu_exchange_set_timeout_ms(pcramp_gpu_ctx *ctx, uint32_t ms);
int pcramp_gpu_exchange_status(pcramp_gpu_ctx *ctx, uint32_t *timed_out_mask);
int pcramp_gpu_exchange_destroy(pcramp_gpu_ctx *ctx);
/* reduce_best_assay (main.cpp:1421-1601) over the ranks of the exchange, without MPI / NCCL / host staging: every rank passes the
 * winner of ITS trials (pcramp_gpu_best_assay: Score members, PCR::total_degeneracy, GLOBAL trial index; index < 0 = no assay, the
 * rank then competes with the default Score, pcramp.h:176-179) and receives the overall winner and the rank that owns it (which
 * then broadcasts oligos / amplicons, as the root does at main.cpp:1503-1601).  Rule of main.cpp:1455-1480 folded in rank order:
 * a later record wins unless its Score is lower (Score::operator<, pcramp.h:180-187), or equal (operator==) with a total degeneracy
 * that is not smaller.  One single-CTA kernel pushes the 32-byte record into every peer's buffer, waits for all ranks' flags and
 * folds.  Synchronises the stream. */
int pcramp_gpu_reduce_best(pcramp_gpu_ctx *ctx, float target_coverage, float background_coverage, float oligo_overlap, double degeneracy,
	int64_t global_trial, uint32_t *owner_rank, float *best_target, float *best_background, float *best_overlap, double *best_degeneracy,
	int64_t *best_trial);

/* ---- K3: SantaLucia nearest-neighbour thermodynamics: replaces the NucCruc call surface pcramp uses
 *      (nuc_cruc.h:696-763 tm_pm_duplex / approximate_tm_hairpin, :775-838 salt / strand, :875-994 set_query /
 *      set_target; nuc_cruc.cpp:2236-2455 approximate_tm_{heterodimer,homodimer,hairpin}) for a BATCH of
 *      oligos / oligo pairs.  One NucCruc object per OpenMP thread (main.cpp:528-535, optimize.cpp:49-52)
 *      becomes one call per batch.
 *        seq_a / seq_b : n strings of bases (ACGT, and I except for PM_DUPLEX), NUL terminated, at a fixed
 *                        stride (bytes); at most 32 bases (Word length, options.cpp:854-860).  seq_b is read by
 *                        the heterodimer ops only (query = a, target = b).
 *        salt          : NucCruc::salt(), [Na+] in mol/l, 1e-6 .. 1
 *        strand_a      : n total strand concentrations as NucCruc::strand(c) takes them; for the heterodimer
 *                        ops strand_b (n) is required and the pair goes through NucCruc::strand(c_a, c_b)
 *        tm, dH, dS, dG_dp : n floats each, any may be NULL: the return value of the Tm call, delta_H(),
 *                        delta_S(), delta_G_dp() (nuc_cruc.h:1361-1383).
 *      Errors mirror the reference's throws (illegal base, empty hairpin query, sequence too long). ------- */
enum {
	PCRAMP_TM_PM_DUPLEX = 0,        /* tm_pm_duplex(a)                                   nuc_cruc.h:723-759 */
	PCRAMP_TM_HAIRPIN = 1,          /* set_query(a); approximate_tm_hairpin()            nuc_cruc.cpp:2381-2455 */
	PCRAMP_TM_HOMODIMER = 2,        /* set_query(a); approximate_tm_homodimer()          nuc_cruc.cpp:2296-2354 */
	PCRAMP_TM_HETERODIMER = 3,      /* set_query(a); set_target(b); approximate_tm_heterodimer()  :2236-2294 */
	PCRAMP_TM_HETERODIMER_DIAG = 4, /* the same with fast_alignment(true) (optimize.cpp:51)         :546-612 */
	PCRAMP_TM_HOMODIMER_DIAG = 5,   /* approximate_tm_homodimer() with fast_alignment(true) */
	PCRAMP_TM_NUM_OPS = 6
};
int pcramp_gpu_thermo_batch(pcramp_gpu_ctx *ctx, int op, uint32_t n, const char *seq_a, const char *seq_b, uint32_t stride,
	float salt, const float *strand_a, const float *strand_b, float *tm, float *dH, float *dS, float *dG_dp);
/* Resident variant: stage a batch in HBM once, run the kernel on it any number of times, fetch when wanted. */
int pcramp_gpu_thermo_stage(pcramp_gpu_ctx *ctx, int op, uint32_t n, const char *seq_a, const char *seq_b, uint32_t stride,
	float salt, const float *strand_a, const float *strand_b);
int pcramp_gpu_thermo_run_staged(pcramp_gpu_ctx *ctx);
int pcramp_gpu_thermo_fetch(pcramp_gpu_ctx *ctx, float *tm, float *dH, float *dS, float *dG_dp);

/* PCR::is_valid (valid_pcr.cpp:5-45) for n trial oligos (words, n x 2 uint64): every concrete expansion of a
 * degenerate oligo (Word::begin/next, word.h:525-647) must have tm_pm_duplex in [tm_min, tm_max],
 * approximate_tm_hairpin <= max_hairpin and, if check_homo_dimer, approximate_tm_homodimer <= max_dimer, at strand
 * concentration primer_strand / degeneracy.  fast_alignment = the NucCruc::fast_alignment() flag of the caller's
 * object (true inside optimize(), optimize.cpp:51).  valid[i] = 1 / 0. */
int pcramp_gpu_is_valid(pcramp_gpu_ctx *ctx, const uint64_t *words, uint32_t n, float salt, float primer_strand, float tm_min,
	float tm_max, float max_hairpin, float max_dimer, int check_homo_dimer, int fast_alignment, uint8_t *valid);
/* PCR::max_dimer_tm (pcr_assay.cpp:232-269): the highest heterodimer Tm over all expansions of F x R. */
int pcramp_gpu_max_dimer_tm(pcramp_gpu_ctx *ctx, const uint64_t *f, const uint64_t *r, uint32_t n_pairs, float salt,
	float primer_strand, int fast_alignment, float *tm);
/* PCR::multiplex_compatible (pcr_assay.cpp:815-852) as main.cpp:748-752 calls it: for trial assay i (f[i], r[i]),
 * ok[i] = AND over the pool of pool_assay.multiplex_compatible(melt, opt, trial_i), i.e. no heterodimer
 * (query = an expansion of a pool oligo, target = an expansion of a trial oligo) reaches max_dimer. */
int pcramp_gpu_multiplex_compatible(pcramp_gpu_ctx *ctx, const uint64_t *f, const uint64_t *r, uint32_t n_pairs,
	const uint64_t *pool_f, const uint64_t *pool_r, uint32_t n_pool, float salt, float primer_strand, float max_dimer,
	int fast_alignment, uint8_t *ok);
/* Counters of the last K3 call: problems, DP cells (q*t gapped dimer, (n-4)(n-3)/2 hairpin, min(q,t) diagonal;
 * SURVEY.md section 8d), kernel launches, CUDA-event time of the kernels. */
/* The same batch with the oligos as WORDS (two uint64 each: what PCR::is_valid / max_dimer_tm hold and Word::str() turns into the text
 * the calls above take, word.h:649-666): 16 bytes per oligo across the host link instead of its text.  Every base must be a single
 * letter (a degenerate base or an EOS inside an oligo fails like its text would: "Unknown base" / "Illegal base"). */
int pcramp_gpu_thermo_words(pcramp_gpu_ctx *ctx, int op, uint32_t n, const uint64_t *words_a, const uint64_t *words_b, float salt,
	const float *strand_a, const float *strand_b, float *tm, float *dH, float *dS, float *dG_dp);
typedef struct pcramp_gpu_thermo_stats {
	uint64_t n_problems;
	uint64_t dp_cells;
	uint64_t kernel_launches;
	float ms_kernel;
} pcramp_gpu_thermo_stats;
int pcramp_gpu_get_thermo_stats(pcramp_gpu_ctx *ctx, pcramp_gpu_thermo_stats *out);

/* ---- K4: SO::SeqOverlap Smith-Waterman (seq_overlap.h, seq_overlap.cpp:347-609; nucleic-acid SmithWaterman mode,
 *      +2 / -3 / -5 / -2) and the background tests built on it (background_match.cpp:7-295). ------------------ */
/* Raw alignments of n (query word, target word) pairs: what pack_query_slots(Word) + pack_target_slots(Word) +
 * align() + score() / alignment_range_query() / alignment_range_target() / target_last_two_aligned() return for one
 * slot (seq_overlap.h:828-869,1102-1136,1265-1330).  Coordinates are -1 when no cell reaches the initial maximum
 * (0): the reference then reports stale values of an earlier alignment.  last_two: 2 bytes per problem (4-bit codes,
 * 15 = N).  Any output may be NULL. */
int pcramp_gpu_sw_batch(pcramp_gpu_ctx *ctx, uint32_t n, const uint64_t *query, const uint64_t *target, int32_t *score,
	int32_t *q_start, int32_t *q_stop, int32_t *t_start, int32_t *t_stop, uint8_t *last_two);
/* CUDA-event time (ms) of the alignment kernel of the last pcramp_gpu_sw_batch (instrumentation, bench.py's sw_gcups leg). */
int pcramp_gpu_sw_timing(pcramp_gpu_ctx *ctx, float *ms_kernel);
/* PCR::find_background_match (background_match.cpp:7-166) for n_pairs assays against the database built by
 * pcramp_gpu_select_words on `kind` (normally PCRAMP_BACKGROUND): collect_background_candidates with
 * search_threshold = opt.background_threshold * opt.background_search_multiplier and the background amplicon range
 * (assay.h:411-421), then four alignments per candidate amplicon and sqrt(S_F S_R / (2|F| 2|R|)) >= detect_threshold
 * (= opt.background_threshold).  bitsets: n_pairs x ceil(n_seq/32), written (not OR-ed).  n_amplicons (may be NULL)
 * receives the number of candidate amplicons.  The reference's guard `(i + 1) >= num_seq` (:122) is reproduced: a
 * candidate at an odd position of a pair's list is scored only while that position is below the number of sequences. */
int pcramp_gpu_background_match(pcramp_gpu_ctx *ctx, int kind, const uint64_t *f, const uint64_t *r, uint32_t n_pairs,
	float search_threshold, float detect_threshold, int amplicon_min, int amplicon_max, int use_taq_mama,
	uint32_t *bitsets, uint64_t *n_amplicons);
/* PCR::find_multiplex_background_match (background_match.cpp:168-295): every sequence of `kind` (normally
 * PCRAMP_MULTIPLEX, the amplicons of the assays already chosen, main.cpp:989-1008) against F, rc(F), R, rc(R) of
 * every pair; a sequence is matched when any of the four normalised scores reaches threshold. */
int pcramp_gpu_multiplex_background_match(pcramp_gpu_ctx *ctx, int kind, const uint64_t *f, const uint64_t *r,
	uint32_t n_pairs, float threshold, int use_taq_mama, uint32_t *bitsets);

/* ---- candidate generation ----------------------------------------------------------------------
 * PCR::random_assay (pcr_assay.cpp:580-734) under the seeding protocol of main.cpp:527-548.  A "stream" is what one
 * OpenMP thread of the reference runs: its own NucCruc object, its own seed (seeds[i] = the thread's
 * `local_seed = rand_r(&global_seed)`; advanced in place, glibc rand_r), and trials_per_stream[i] consecutive trials of
 * the static schedule.  Stream i's assays follow stream i-1's in f / r (2 uint64 per oligo, centred like
 * PCR::center()).  One GPU thread per stream: n_streams = number of trials gives the reference's result for
 * `--thread <num_trial>`, n_streams = 1 for `--thread 1`.  attempts (may be NULL) = candidates tried per trial.
 * Errors carry the reference's messages (":PCR::random_assay: No active sequences found", "... Unable to generate a
 * valid initial assay to test!", "... sequence length is too small!"). */
typedef struct pcramp_gpu_random_assay_options {
	int primer_min, primer_max;       /* opt.primer_range */
	int amplicon_min, amplicon_max;   /* opt.target_amplicon_range */
	uint32_t degen;                   /* opt.degen */
	float salt, primer_strand;        /* opt.salt, opt.primer_strand */
	float primer_tm_min, primer_tm_max, max_hairpin, max_dimer;
} pcramp_gpu_random_assay_options;
int pcramp_gpu_random_assays(pcramp_gpu_ctx *ctx, int kind, uint32_t n_streams, uint32_t *seeds, const uint32_t *trials_per_stream,
	const pcramp_gpu_random_assay_options *options, uint64_t *f, uint64_t *r, uint32_t *attempts);

/* ---- one design iteration: the body of the `while(true)` loop of main.cpp:471-1130 ---------------------------------------------
 * With the targets (PCRAMP_TARGET) and backgrounds (PCRAMP_BACKGROUND) resident on `ctx`, one call does what one pass of that loop
 * does: draws opt.num_trial trial assays (PCR::random_assay under the seeding of :527-548), indexes the backgrounds and the active
 * targets against them (:560-691), runs optimize() on every trial (:697-729), screens the survivors -- multiplex_compatible with
 * the pool, find_multiplex_background_match, the pool x amplicon test, find_background_match (:731-840) -- keeps the best under the
 * update rule of :829-858, records its target matches (find_target_match, :898) and, for multiplex designs, appends its amplicons
 * to the multiplex background and splits the targets at their bounds (:989-1017), and retires the detected targets (:1116-1121).
 * Every step is a batch call of this header over all trials; the decisions the reference takes trial by trial (the running best
 * score gates which trials are screened) are replayed in trial order from the batched results = the reference at `--thread 1`.
 *   n_streams   seed streams of candidate generation: 1 = `--thread 1` (one stream draws all trials, bit-exact with the stock
 *               program); k > 1 = the static schedule of `--thread k` with the threads' seeds drawn in thread order (the reference
 *               draws them in whatever order its threads reach the critical section).
 *   seed        opt.seed (main.cpp:112: the global seed every iteration's local seeds are drawn from with rand_r).
 * --optimize.top-down (make_degenerate) is not offered. */
typedef struct pcramp_gpu_design_options {
	uint32_t num_trial, n_streams;                             /* opt.num_trial; OpenMP threads of the trial loop */
	uint32_t degen;                                            /* opt.degen (-d) */
	int optimize_5, optimize_3;                                /* opt.optimize_5 / optimize_3 */
	int primer_min, primer_max;                                /* opt.primer_range */
	float primer_tm_min, primer_tm_max, primer_strand, salt, max_hairpin, max_dimer;
	int target_amplicon_min, target_amplicon_max, background_amplicon_min, background_amplicon_max;
	float target_threshold, target_search_multiplier, background_threshold, background_search_multiplier;
	float min_target_cover, max_background_cover;              /* opt.min_target_cover, opt.max_background_cover */
	uint32_t pack_max_degen;
	float pack_min_gc, pack_max_gc;
	int use_taq_mama, use_multiplex;
} pcramp_gpu_design_options;
typedef struct pcramp_gpu_design_result {
	int found;                                /* best_score.target_coverage > 0 (main.cpp:928): 0 ends the run */
	uint32_t iteration, major_id, minor_id;   /* assay_iteration, major_assay_id, minor_assay_id (:458-502) */
	uint32_t targets_remaining, num_active_target, num_active_background, trial;
	float active_target_norm, active_background_norm;
	uint64_t f[2], r[2];                      /* best_assay */
	double degeneracy_f, degeneracy_r;
	int reused_f, reused_r;                   /* PCR::write(out, pool) writes a re-used oligo in lower case (assay.h:305-343) */
	float target_coverage, background_coverage, oligo_overlap; /* best_score */
	uint64_t n_target_entries, n_background_entries, n_amplicons_added, n_splits, n_multiplex_keys;
	float ms_total, ms_candidates, ms_select_background, ms_select_target, ms_optimize, ms_screen, ms_accept; /* host wall clock */
} pcramp_gpu_design_result;
typedef struct pcramp_gpu_design pcramp_gpu_design;
void pcramp_gpu_design_default_options(pcramp_gpu_design_options *options); /* Options::Options(), options.cpp:40-92 */
int pcramp_gpu_design_create(pcramp_gpu_ctx *ctx, const pcramp_gpu_design_options *options, uint32_t seed, pcramp_gpu_design **out);
void pcramp_gpu_design_destroy(pcramp_gpu_design *d);
const char *pcramp_gpu_design_last_error(const pcramp_gpu_design *d);
int pcramp_gpu_design_iteration(pcramp_gpu_design *d, pcramp_gpu_design_result *result);
/* LSB-first bitsets of the last iteration's best assay: best_target_match, best_background_match (either may be NULL) */
int pcramp_gpu_design_matches(pcramp_gpu_design *d, uint32_t *target_bits, uint32_t *background_bits);
/* Sequence::active() of every target, and the union of the accepted assays' background matches (main.cpp:1131-1153) */
int pcramp_gpu_design_active(pcramp_gpu_design *d, uint8_t *target_active, uint32_t *background_union);

/* ---- instrumentation ----------------------------------------------------------------------------- */
/* Counters of the last select_words / score_pairs call on this ctx. */
typedef struct pcramp_gpu_stats {
	uint64_t n_patterns;      /* candidate words x 2 strands scanned */
	uint64_t n_positions;     /* template positions streamed (active sequences) */
	uint64_t n_hits;          /* (candidate, window) hits at or above threshold */
	uint64_t n_entries;       /* database entries */
	uint64_t n_keys;          /* unique words */
	uint64_t kernel_launches; /* kernels of this library launched by the call */
	uint64_t n_seeded;        /* patterns (oligo x strand) that went through the seed filter; the rest were brute-forced */
	uint64_t n_seed_entries;  /* seed-table entries built for them */
	uint64_t n_indexed;       /* of those, patterns taken by the indexed scan (index.cuh) */
	uint64_t n_index_queries; /* k-mer neighbour queries issued against the text index */
	uint64_t n_index_entries; /* index entries those queries cover (16 bytes each: the scan's HBM traffic) */
	float ms_seed;            /* CUDA-event time of the seeded scan: index queries + scan_index_kernel, seed-table build + scan_seed_kernel, dirty groups */
	float ms_scan;            /* ... of the brute-force scan kernel (patterns that cannot be seeded) */
	float ms_edge;            /* ... of the partial-window kernel */
	float ms_db;              /* ... of hit filtering, sorting, materialisation */
	float ms_score;           /* ... of the pair-scoring kernels */
	float ms_index_kernel;    /* ... of scan_index_kernel alone (summed over the parts of the index) */
	float ms_index_build;     /* host wall clock of the last build of the collection's text index (one-time per upload; splits do not rebuild) */
	uint64_t index_bytes;     /* device memory the text index holds */
	uint64_t n_index_builds;  /* builds of this collection's index so far */
	uint64_t n_index_stale;   /* sequences split since the build and active again: covered by the table scan in this call */
	uint64_t n_fast;          /* select_words batches of this ctx that ran in the fast form (no host round trip inside) so far */
	uint64_t n_fast_redo;     /* ... of those, batches whose verification failed and that were run again in the general form */
	uint64_t n_edge_words;    /* partial words (both strands) in the collection's table (edge.cuh) when this call used it, else 0 */
	uint64_t edge_table_bytes;/* device memory of that table */
	float ms_edge_table_build;/* host wall clock of its last build (one-time per upload / split and set of pack() parameters) */
	uint32_t edge_table_used; /* 1: the partial words were matched through the table; 0: by the per-batch scan kernel */
} pcramp_gpu_stats;
int pcramp_gpu_get_stats(pcramp_gpu_ctx *ctx, pcramp_gpu_stats *out);
/* Tuning / testing switches.  "force_brute_scan" = 1 sends every pattern through the brute-force scan kernel;
 * "use_index" = 0 keeps the seeded patterns on scan_seed_kernel instead of the indexed scan (default 1: the text index
 * is built on first use, ~32 bytes per base of device memory); "use_seed_table" = 0 makes pair scoring and the
 * partial-word scan compare every word with every oligo instead of going through the frame-aligned seed table (fst.cuh);
 * "use_neighbours" = 0 makes pair scoring find the oligos of a database word by that table walk instead of through the
 * neighbour list of the candidate that produced the word (score.cuh); "use_entry_score" = 0 goes back from the entry-driven scoring
 * kernel (one thread per plus-strand entry looks for the partner among the minus-strand entries in amplicon range) to the
 * (sequence, strand, oligo) bit rows + work list + one warp per item; "use_unit_score" = 0 sends pair scoring at thresholds where neither the neighbour bound nor a
 * seed table excludes anybody (search below ~0.82: the background thresholds) through the key-matrix path instead of one thread per (sequence, pair)
 * unit over the membership lists (sw_abi.cuh); "use_async_scan" = 1 makes the indexed scan keep a ring of chunks in flight in shared memory
 * through cp.async (scan_index_async_kernel, measured slower) instead of holding the entries of one range in registers (scan_index_kernel); "use_background_units" = 0 makes pcramp_gpu_background_match write
 * down every candidate amplicon and align four times per amplicon instead of aligning once per (pair, matching database entry) and
 * walking the (plus, minus) combinations of every (sequence, pair) in the reference's order (sw_abi.cuh); "use_variant_groups" = 0 scores the variants of pcramp_gpu_score_variants one
 * by one as pairs with their own copy of the base assay instead of by groups that share it (score_entries_groups_kernel); "use_segmented_db" = 0 builds the database through a radix
 * sort + unique-by-key of the entry ids instead of per-(sequence, strand) segments sorted in shared memory (db.cuh); "use_tier_table" = 0 applies select_words' best-tier rule
 * by sorting the hit list instead of through a (sequence, candidate) table of maxima (db.cuh); "tiny_buffers" = 1 makes every
 * growable device buffer (hits, index queries / candidates, neighbour list, work list) start far too small on a fresh context,
 * so that the overflow -> grow -> re-run paths are exercised; "index_part_positions" = the most positions one part of the text index
 * may hold (default 2^31; a small value cuts a test collection into several parts, as a collection above 2^31 bases is);
 * "use_edge_table" = 0 makes the fast form of select_words match the partial words pack() emits at sequence ends by the per-batch scan
 * kernel (scan_edge_fst_kernel) instead of looking the candidates up in the collection's table of partial words (edge.cuh).
 * All paths are CUDA and give identical results; the tests compare them.
 * The environment variable PCRAMP_OPTIONS="name=value,name=value" applies pcramp_gpu_set_option to every context the process
 * creates (A/B runs of a host that has no switch of its own); an unknown name makes pcramp_gpu_create fail with 6. */
int pcramp_gpu_set_option(pcramp_gpu_ctx *ctx, const char *name, int value);
/* Issue-bound ceiling of the scan's own instruction mix on this GPU (alignments/s), measured live. */
int pcramp_gpu_measure_int_peak(pcramp_gpu_ctx *ctx, double *alignments_per_s);
/* Issue-bound ceiling of the dynamic-programming kernels' instruction mix (three-input maximum, add-maximum, logic op, add) in
 * INT32 operations/s, measured live: the denominator of the K3 / K4 rooflines (SURVEY.md 8d costs a DP cell in INT32 operations). */
int pcramp_gpu_measure_int32_peak(pcramp_gpu_ctx *ctx, double *ops_per_s);

/* ---- host-side word helpers (word.h / word.cpp), for building candidate lists ------------------- */
void pcramp_word_from_string(const char *iupac, int centre, uint64_t out[2]);
int pcramp_word_to_string(const uint64_t w[2], char out[33]);
uint32_t pcramp_word_and(const uint64_t a[2], const uint64_t b[2]);
uint32_t pcramp_word_size(const uint64_t a[2]);
int pcramp_word_start(const uint64_t a[2]);
int pcramp_word_stop(const uint64_t a[2]);
void pcramp_word_complement(const uint64_t a[2], uint64_t out[2]);
void pcramp_word_center(const uint64_t a[2], uint64_t out[2]);
/* Word::max_overlap (word.h:38-92): the largest number of equal nibbles on one diagonal over max(size, size) */
float pcramp_word_max_overlap(const uint64_t a[2], const uint64_t b[2]);

#ifdef __cplusplus
}
#endif
#endif /* PCRAMP_GPU_H */
