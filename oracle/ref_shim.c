/*
 * ref_shim.c -- exposes the UNMODIFIED reference functions to the tests.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile straight from the sources where
 * they lie under $(REF) (default /root/reference); no reference source is copied into
 * this repository.  The reference translation unit is pulled in with its main() renamed
 * so read_adapters()/read_fastq()/transform()/draw() (quack.c:154,180,230,295 -- all
 * non-static) can be called from ctypes.  Output: oracle/_ref/libquack_ref.so.
 */
#define main quack_reference_main
#include QB_REF_QUACK_C /* e.g. "/root/reference/quack.c", set by the Makefile */
#undef main

#include <string.h>

/* Runs the reference read_fastq() (optionally after read_adapters()) and copies the raw,
 * pre-transform accumulator out.  Returns 0, or -1 if cap_rows is too small. */
int qref_read_fastq(const char *fastq, const char *adapters, uint64_t *rows_out,
                    uint64_t cap_rows, uint64_t *max_length, uint64_t *n_reads) {
  int *kmers = adapters ? read_adapters((char *)adapters) : NULL;
  sequence_data *d = read_fastq((char *)fastq, kmers);
  int rc = 0;
  *max_length = d->max_length;
  *n_reads = d->number_of_sequences;
  if (d->max_length > cap_rows)
    rc = -1;
  else if (d->max_length)
    memcpy(rows_out, d->bases, d->max_length * sizeof(base_information));
  free(d->bases);
  free(d);
  free(kmers);
  return rc;
}

/* The reference adapter table as bytes (1 << 20 entries of 0/1). */
int qref_read_adapters(const char *adapters, uint8_t *table_out) {
  int *kmers = read_adapters((char *)adapters);
  for (int i = 0; i < (1 << 20); i++) table_out[i] = kmers[i] != 0;
  free(kmers);
  return 0;
}

/* The reference transform() on caller-provided rows (in place). */
int qref_transform(uint64_t *rows, uint64_t *max_length, uint64_t n_reads,
                   uint64_t *original_max_length) {
  sequence_data d;
  d.bases = (base_information *)rows;
  d.max_length = *max_length;
  d.original_max_length = 0;
  d.number_of_sequences = n_reads;
  transform(&d);
  *max_length = d.max_length;
  *original_max_length = d.original_max_length;
  return 0;
}

/* The reference lookup[] expression on one byte of its defined domain. */
int qref_base_code(int c) { return lookup[c - 65 & ~32]; }

int qref_row_bytes(void) { return (int)sizeof(base_information); }

/* The reference record reader (kseq_read, klib/kseq.h:177-218) over one file: writes
 * "seq\tqual\n" per record into out (up to cap bytes) and returns the final negative
 * return code through *last_rc.  Returns bytes written, or -1 if out is too small. */
long qref_parse(const char *path, char *out, long cap, int *last_rc) {
  gzFile fp = gzopen(path, "r");
  kseq_t *seq = kseq_init(fp);
  long n = 0;
  int l;
  while ((l = kseq_read(seq)) >= 0) {
    long need = (long)seq->seq.l + (long)seq->qual.l + 2;
    if (n + need > cap) { n = -1; break; }
    memcpy(out + n, seq->seq.s, seq->seq.l); n += seq->seq.l;
    out[n++] = '\t';
    if (seq->qual.l) { memcpy(out + n, seq->qual.s, seq->qual.l); n += seq->qual.l; }
    out[n++] = '\n';
  }
  *last_rc = l;
  kseq_destroy(seq);
  gzclose(fp);
  return n;
}
