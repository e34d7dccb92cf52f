/* fq_reader.h -- batching FASTA/Q record reader over zlib, with the framing semantics of
 * kseq_read() (reference klib/kseq.h:177-218) but packing records straight into the batch layout the
 * C-ABI takes (include/quack_b200.h) instead of handing out one record at a time. */
#ifndef QB_FQ_READER_H
#define QB_FQ_READER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fqr_reader fqr_reader;

/* Opens path (gzip, multi-member gzip, BGZF or plain text).  NULL if it cannot be opened.
 * BGZF input (bgzip / htslib blocked gzip) is inflated by `threads` pool threads (0: QUACK_DECODE_THREADS or
 * half of the online cores, at most 8; 1: no pool); every other input goes through gzread() like the
 * reference (quack.c:187).  The bytes handed to the framing code are the same either way. */
fqr_reader *fqr_open(const char *path);
fqr_reader *fqr_open_mt(const char *path, int threads);
int fqr_default_threads(void);
int fqr_decode_threads(const fqr_reader *r); /* pool threads of this reader (1: gzread path) */
void fqr_close(fqr_reader *r);

/* Appends whole records to seq[]/qual[]/offset[]/length[] until the stream ends, cap_reads records are
 * in, or the next record would not fit cap_bytes.  Returns 1 if the stream may hold more records, 0 if
 * it ended: clean EOF (-1), truncated quality (-2), stream error (-3) or a FASTA record -- exactly the
 * points where the reference's `while ((l = kseq_read(seq)) >= 0)` loop stops (quack.c:193); the
 * reason is available from fqr_status().  *max_len receives the longest read appended.
 * A record longer than cap_bytes on its own makes fqr_fill return -1 (status -4). */
int fqr_fill(fqr_reader *r, uint8_t *seq, uint8_t *qual, uint32_t *offset, uint32_t *length, uint64_t cap_bytes,
             uint32_t cap_reads, uint32_t *n_reads, uint64_t *n_bytes, uint32_t *max_len);
int fqr_status(const fqr_reader *r); /* 0 while records keep coming, else -1/-2/-3 as kseq_read, -4 too long, -5 FASTA */

/* One record at a time (adapter FASTA files, tests): returns the sequence length >= 0 or the negative
 * kseq_read() code; pointers stay valid until the next call; *qual_len == 0 for FASTA records. */
long fqr_next(fqr_reader *r, const uint8_t **seq, const uint8_t **qual, size_t *qual_len);

/* Up to cap raw decompressed bytes of the stream, no framing (for the device-side framing, qb_text_submit).  Returns
 * the number of bytes written; fewer than cap: the stream ended (fqr_status() says how: -1 clean, -3 damaged). */
long fqr_read_raw(fqr_reader *r, uint8_t *dst, size_t cap);

/* bytes of decompressed input consumed so far, and seconds spent inside gzread() -- or, with a BGZF pool,
 * waiting for the pool's next block (host gzip decode, reported separately from the statistics path as
 * BASELINE.json asks) */
uint64_t fqr_bytes_in(const fqr_reader *r);
double fqr_inflate_seconds(const fqr_reader *r);

/* read_adapters(), quack.c:154-178: keys (reference order) of every record of a FASTA file.  *keys is
 * malloc'd (caller frees).  Returns the number of keys, or -1 if the file cannot be opened. */
long fqr_read_adapter_keys(const char *path, uint32_t **keys);

#ifdef __cplusplus
}
#endif
#endif
