/*
 * synth.h — deterministic synthetic libraries and reads in the shapes BASELINE.json names
 * (SURVEY.md §8d).  Benchmark / test input generator, not part of the counting path.
 *
 * Every random draw is a counter-based hash of (seed, stream, index), so any read of any
 * sample can be produced on the host or on the device, in any order, with identical bytes.
 *
 * Library: base j of a guide is G with P = 0.55 - 0.30 j/(k-1), else A/C/T evenly (gives the
 * entropy detector a monotone profile to lock onto); 0.5 % of guides are planted Hamming-1
 * neighbours of an earlier guide and 0.5 % Hamming-2 neighbours; duplicates are redrawn.
 *
 * Reads (fixed length L, newline-terminated lines, stride L+1):
 *   prefix[offset] + window[k] + scaffold, cut to L; guide index ~ log-normal(sigma = 1)
 *   abundance; class mix: 80 % exact, 8 % one ACGT substitution, 2 % one N, 2 % two
 *   substitutions, 2 % shifted +1, 2 % shifted -1, 1 % truncated (bases from a cut point
 *   <= offset+k-3 on are N: the fixed-length stand-in for a short read), 3 % random bases.
 *   Reverse samples hold the reverse complement of the whole read and replace the N class
 *   by an ACGT substitution (keeps them independent of how fxread maps N).
 */
#ifndef SGCOUNT_SYNTH_H
#define SGCOUNT_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgs_sample sgs_sample;

const char* sgs_last_error(void);

/* n*k ASCII bytes, row-major */
int sgs_make_library(uint64_t seed, uint32_t n, uint32_t k, uint8_t* out);

int sgs_sample_create(uint64_t seed, uint32_t sample_idx, const uint8_t* library, uint32_t n, uint32_t k,
                      uint32_t read_len, uint32_t offset, int reverse, sgs_sample** out);
void sgs_sample_destroy(sgs_sample*);

/* reads [first_read, first_read + n_reads) as (read_len+1)-byte lines */
int sgs_sample_fill_host(const sgs_sample*, uint64_t first_read, uint64_t n_reads, uint8_t* out, int n_threads);
int sgs_sample_fill_device(sgs_sample*, int device, uint64_t first_read, uint64_t n_reads, uint8_t* d_out,
                           void* stream);

/* 4-line FASTQ (@r<idx>, constant quality), gzip with one member per `reads_per_member`
 * reads so a reader can inflate members in parallel; gz_level 0 writes plain text. */
int sgs_sample_write_fastq(const sgs_sample*, uint64_t first_read, uint64_t n_reads, const char* path,
                           uint64_t reads_per_member, int gz_level, int n_threads);

/* The same records as BGZF (bgzip's blocked gzip: members of <= 64 KB with the 'BC' block-size extra
 * field, cut every block_bytes of text wherever that falls, and the empty end-of-file block). */
int sgs_sample_write_fastq_bgzf(const sgs_sample*, uint64_t first_read, uint64_t n_reads, const char* path, int gz_level,
                                int n_threads, uint32_t block_bytes);

#ifdef __cplusplus
}
#endif
#endif
