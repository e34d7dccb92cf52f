/*
 * quack_oracle.h -- CPU restatement of quack's per-read statistics accumulation.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle for the sm_100a CUDA path.
 * Nothing under quack_b200/ (the product) may include, link or call it; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do, and only as the checker or as the timed CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function
 * here against the unmodified reference compiled from /root/reference into
 * oracle/_ref/ (see oracle/Makefile) and against the known-answer vectors of
 * SURVEY.md Appendix B committed under tests/golden/.
 *
 * Each function cites the reference lines it follows (quack.c / klib/kseq.h at
 * reference commit cf39419, klib de09fb7).
 */
#ifndef QUACK_ORACLE_H
#define QUACK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Row layout of one read position: identical to base_information, quack.c:134-139.
 * u64 index: 0..90 scores[q-33], 91..94 content[A,T,C,G], 95 length_count, 96 kmer_count. */
#define QO_ROW_U64 97
#define QO_SCORES 91
#define QO_COL_CONTENT 91
#define QO_COL_LENGTH 95
#define QO_COL_KMER 96
#define QO_KMER_SIZE 10
#define QO_TABLE_SIZE (1u << 20) /* 4^10, quack.c:156 */

typedef struct {
  uint64_t *rows;         /* [cap][97], zero-initialised */
  uint64_t cap;           /* rows allocated */
  uint64_t max_length;    /* longest read seen, quack.c:194-198 */
  uint64_t n_reads;       /* number_of_sequences, quack.c:220 */
  uint64_t n_invalid_qual; /* quality bytes outside [33,123]: UB in the reference, ignored here */
} qo_stats;

/* lookup[(c-65) & ~32], quack.c:148-150, 200-201.  Exact on the reference's defined
 * domain [A-Ta-t]: C/c->2, G/g->3, T/t->1, everything else (incl. N) -> 0.  Bytes outside
 * that domain index lookup[] out of bounds in the reference (UB); here they are defined as
 *   2*[(b&0x5B)==0x43] + [(b&0x1F)==0x07 || (b&0x1F)==0x14]
 * which is the same rule the CUDA kernel uses, so both agree on all 256 byte values. */
int qo_base_code(int c);

/* ---- adapter set: read_adapters(), quack.c:154-178 ---- */
uint8_t *qo_table_new(void);                 /* QO_TABLE_SIZE bytes of 0 */
void qo_table_free(uint8_t *table);
/* one FASTA record: windows ending at i = 10..l-1 are inserted, the first window is not
 * (quack.c:165-172).  Keys: first base most significant, 2 bits per base. */
void qo_table_add_record(uint8_t *table, const char *seq, size_t l);
/* whole (gzipped or plain) FASTA file; returns number of records, <0 if it cannot be opened */
long qo_read_adapters(const char *path, uint8_t *table);
/* sorted list of the distinct keys in a table; returns count (out may be NULL to count) */
uint32_t qo_table_keys(const uint8_t *table, uint32_t *out);

/* ---- accumulation: body of the while loop in read_fastq(), quack.c:193-221 ---- */
int qo_stats_init(qo_stats *st);
void qo_stats_free(qo_stats *st);
/* table == NULL reproduces the no -a behaviour incl. kmer_count[10]++ for l > 10 */
int qo_accumulate_read(qo_stats *st, const uint8_t *seq, const uint8_t *qual, size_t l,
                       const uint8_t *table);
/* packed batch exactly as handed to the C-ABI (concatenated bytes + offsets + lengths) */
int qo_accumulate_batch(qo_stats *st, const uint8_t *seq, const uint8_t *qual,
                        const uint32_t *offset, const uint32_t *length, uint64_t n_reads,
                        const uint8_t *table);

/* ---- record framing: kseq_read(), klib/kseq.h:177-218, over zlib gzread ---- */
typedef struct qo_reader qo_reader;
qo_reader *qo_reader_open(const char *path);
/* returns seq length >= 0, -1 EOF, -2 truncated quality, -3 stream error (kseq.h:171-176);
 * *seq / *qual point into reader-owned buffers valid until the next call; *qual_len is 0
 * for a FASTA record */
long qo_reader_next(qo_reader *r, const uint8_t **seq, const uint8_t **qual, size_t *qual_len);
void qo_reader_close(qo_reader *r);

/* whole file: read_fastq(), quack.c:180-228.  Returns 0, <0 if the file cannot be opened. */
int qo_read_fastq(const char *path, const uint8_t *table, qo_stats *st);

/* transform(), quack.c:230-293 (in place; max_length is updated when binning applies).
 * Kept so the host renderer's arithmetic (single-precision ceil, integer percent) can be
 * checked against KAT-F of SURVEY.md Appendix B. */
void qo_transform(qo_stats *st, uint64_t *original_max_length);

/* Side outputs the reference does not compute (PARITY UNPINNED: no reference oracle exists). */
void qo_extras(const uint8_t *seq, const uint8_t *qual, const uint32_t *offset, const uint32_t *length, uint64_t n_reads,
               uint64_t *n_count, uint64_t *qual_sum, uint64_t rows, uint64_t *mean_hist);

#ifdef __cplusplus
}
#endif
#endif
