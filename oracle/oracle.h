/*
 * oracle.h — C interface of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  This library is a CPU restatement of the
 * read->guide matching and counting path of noamteyssier/sgcount v0.1.35 (Rust).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  Nothing under sgcount_b200/ links or calls it.
 *
 * The reference cannot be compiled here (no cargo/rustc; crates not vendored), so
 * the oracle is pinned against the reference's own unit-test vectors
 * (tests/test_oracle_reference_vectors.py, one test per row of SURVEY.md §4) and
 * against matcher-independent labels carried by the example fixtures' read headers
 * (tests/golden/, tests/test_oracle_fixtures.py).
 *
 * Third-party behaviour that is NOT in /root/reference and is restated from the
 * crates' published behaviour: fxread ^0.2.5 (record framing, seq_rev_comp),
 * ndarray-stats ^0.6 (entropy, mean_sq_err, argmin), hashbrown ^0.15 (map semantics).
 * No lockfile pins them.  Each is isolated behind one function here.
 */
#ifndef SGCOUNT_ORACLE_H
#define SGCOUNT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes (negative = the reference would panic, positive = it would return Err) */
#define ORC_OK 0
#define ORC_ERR_INCONSISTENT_SIZE 1   /* library.rs:83 */
#define ORC_ERR_READ_TOO_SHORT 2      /* offsetter.rs:154-156 */
#define ORC_ERR_IO 3
#define ORC_PANIC_DUPLICATE_SEQ (-1)  /* library.rs:92 */
#define ORC_PANIC_NAN (-2)            /* offsetter.rs:123-141 */
#define ORC_PANIC_EMPTY_READER (-3)   /* offsetter.rs:38 */
#define ORC_PANIC_GENEMAP (-4)        /* genemap.rs:58,60 */
#define ORC_PANIC_MALFORMED (-5)      /* fxread iterator on malformed record */

/* how Record::seq_rev_comp treats non-ACGT bytes */
#define ORC_RC_BITTRICK 0 /* c&2 ? c^4 : c^21  (N -> J, J -> N) : recalled fxread 0.2 behaviour */
#define ORC_RC_KEEP_N 1   /* A<->T C<->G, everything else unchanged */

typedef struct orc_library orc_library;
typedef struct orc_permuter orc_permuter;
typedef struct orc_counter orc_counter;
typedef struct orc_records orc_records;

const char* orc_last_error(void);

/* ---- records (fxread restatement) ------------------------------------------------ */
/* Parse FASTA (2 lines/record) or FASTQ (4 lines/record) from memory; format sniffed
 * from the first byte.  gz != 0: buffer is (multi-member) gzip. */
int orc_records_from_memory(const uint8_t* buf, size_t len, int gz, orc_records** out);
/* Open a path; gzip iff it ends with ".gz". */
int orc_records_from_path(const char* path, orc_records** out);
/* Build records directly from concatenated sequences (ids become "r<idx>"). */
int orc_records_from_seqs(const uint8_t* seqs, const uint64_t* off, uint64_t n, orc_records** out);
uint64_t orc_records_len(const orc_records*);
const uint8_t* orc_records_seq(const orc_records*, uint64_t i, uint64_t* len);
const uint8_t* orc_records_id(const orc_records*, uint64_t i, uint64_t* len);
/* total bytes of all sequences, and a copy-out of them as newline-terminated lines
 * plus offsets[n+1] (the format the CUDA path consumes) */
uint64_t orc_records_seq_bytes(const orc_records*);
void orc_records_export_lines(const orc_records*, uint8_t* lines, uint64_t* off);
void orc_records_free(orc_records*);
/* fxread::Record::seq_rev_comp on one buffer */
void orc_seq_rev_comp(const uint8_t* seq, size_t len, int rc_mode, uint8_t* out);

/* ---- Library (library.rs:17-99) --------------------------------------------------- */
int orc_library_from_records(const orc_records*, orc_library** out);
uint64_t orc_library_len(const orc_library*);
uint64_t orc_library_size(const orc_library*); /* common sequence length */
/* index (insertion order) of the record whose sequence == token, or -1 */
int64_t orc_library_contains(const orc_library*, const uint8_t* token, size_t len);
const uint8_t* orc_library_seq(const orc_library*, uint64_t idx, uint64_t* len);
const uint8_t* orc_library_alias(const orc_library*, uint64_t idx, uint64_t* len);
void orc_library_free(orc_library*);

/* ---- Permuter (permutes.rs:47-158) ----------------------------------------------- */
/* order: permutation of 0..n-1 giving the key iteration order (NULL = insertion order);
 * the reference iterates a randomly seeded HashMap, so every order is legal. */
int orc_permuter_new(const orc_library*, const uint64_t* order, orc_permuter** out);
/* parent library index, or -1 */
int64_t orc_permuter_contains(const orc_permuter*, const uint8_t* token, size_t len);
uint64_t orc_permuter_map_len(const orc_permuter*);
uint64_t orc_permuter_null_len(const orc_permuter*);
int orc_permuter_null_contains(const orc_permuter*, const uint8_t* token, size_t len);
void orc_permuter_free(orc_permuter*);

/* ---- Counter (counter.rs:36-252) ------------------------------------------------- */
/* position: 0 Plus, 1 Minus, 2 Centered, 3 Null (counter.rs:7-12) */
int orc_bounds(uint64_t seq_len, uint64_t offset, uint64_t size, int position, uint64_t* min,
               uint64_t* max); /* returns 1 = Some, 0 = None */
/* one read through Counter::assign; returns library index or -1 */
int64_t orc_assign(const orc_library*, const orc_permuter* /* NULL = --exact */, const uint8_t* read,
                   size_t len, int is_reverse, uint64_t offset, int position_recursion, int rc_mode);
/* Counter::new over a record set.  n_threads <= 1 is the reference's literal
 * one-thread-per-sample loop; > 1 shards records over threads and merges the maps
 * (a benchmarking convenience the reference does not have).
 * assign_out (optional, n records): library index or -1 per record. */
int orc_counter_new(const orc_records*, const orc_library*, const orc_permuter*, int is_reverse,
                    uint64_t offset, int position_recursion, int rc_mode, int n_threads,
                    int32_t* assign_out, orc_counter** out);
uint64_t orc_counter_get_value(const orc_counter*, const uint8_t* alias, size_t len);
uint64_t orc_counter_total_reads(const orc_counter*);
uint64_t orc_counter_matched_reads(const orc_counter*);
/* counts by library index (sequences sharing an alias each report the combined count,
 * exactly like results.rs:65-67 would print them) */
void orc_counter_counts_by_index(const orc_counter*, const orc_library*, uint64_t* out);
void orc_counter_free(orc_counter*);

/* ---- Offsetter (offsetter.rs:37-210) --------------------------------------------- */
/* position_counts over records[0..take): first record gives `size` and is not counted.
 * out must hold 4*size doubles where size = length of the first record;
 * call with out == NULL to query size. */
int orc_position_counts(const orc_records*, uint64_t take, double* out, uint64_t* size);
int orc_positional_entropy(const orc_records*, uint64_t take, double* out, uint64_t* size);
int orc_entropy_from_counts(const double* counts, uint64_t size, double* out);
int orc_minimize_mse(const double* reference, uint64_t n_ref, const double* comparison,
                     uint64_t n_cmp, int* is_reverse, uint64_t* index);
int orc_entropy_offset(const orc_records* library, const orc_records* sample, uint64_t subsample,
                       int* is_reverse, uint64_t* index);

/* ---- GeneMap + results (genemap.rs:53-86, results.rs:32-99, utils.rs:18-49) -------- */
/* Renders the count table.  Rows are emitted in library insertion order (the reference's
 * order is hash-iteration order; compare as sets of rows).  genemap_buf may be NULL.
 * Returns a malloc'ed NUL-terminated string in *out (free with orc_free). */
int orc_render_results(const orc_counter* const* counters, const char* const* names, uint64_t n_samples,
                       const orc_library*, const uint8_t* genemap_buf, size_t genemap_len,
                       int include_zero, char** out);
/* newline-joined default sample names (utils.rs:18-49) */
int orc_generate_sample_names(const char* const* paths, uint64_t n, char** out);
void orc_free(void*);

#ifdef __cplusplus
}
#endif
#endif
