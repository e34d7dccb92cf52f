/* fq_reader.c -- see fq_reader.h.  Written against the behaviour of kseq_read() (klib/kseq.h:177-218):
 *   - a record starts at the next '@' or '>' (anything before the first header is skipped);
 *   - name = up to the first white space, rest of the line ignored;
 *   - sequence = following lines up to a line starting with '+', '>' or '@'; empty lines skipped; a
 *     trailing '\r' is dropped when the accumulated string is longer than one byte (kseq.h:141);
 *   - the '+' line is skipped; quality = following lines until at least as many bytes as the sequence;
 *   - quality length != sequence length -> -2 and the stream is over for quack (quack.c:193).
 * Lines are located with memchr over a 1 MiB inflate buffer instead of kseq's per-byte loops. */
#include "fq_reader.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

#include "../../include/quack_b200.h"

#define FQR_BUF (1u << 20)

typedef struct {
  uint8_t *s;
  size_t l, m;
} fqr_str;

struct fqr_reader {
  gzFile f;
  uint8_t *buf;
  size_t begin, end;
  int is_eof, err;
  int last_char; /* header character already consumed, or 0 (kseq.h:183-186, 199) */
  int status;
  fqr_str seq, qual;
  int pending; /* a parsed record is waiting in seq/qual because the previous batch was full */
  uint64_t bytes_in;
  double inflate_s;
};

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

fqr_reader *fqr_open(const char *path) {
  gzFile f = gzopen(path, "r");
  if (!f) return NULL;
  gzbuffer(f, 1u << 18);
  fqr_reader *r = (fqr_reader *)calloc(1, sizeof *r);
  r->f = f;
  r->buf = (uint8_t *)malloc(FQR_BUF);
  return r;
}

void fqr_close(fqr_reader *r) {
  if (!r) return;
  gzclose(r->f);
  free(r->buf);
  free(r->seq.s);
  free(r->qual.s);
  free(r);
}

int fqr_status(const fqr_reader *r) { return r->status; }
uint64_t fqr_bytes_in(const fqr_reader *r) { return r->bytes_in; }
double fqr_inflate_seconds(const fqr_reader *r) { return r->inflate_s; }

/* refill; returns 0 when bytes are available, -1 at end of file, -3 on a stream error */
static int refill(fqr_reader *r) {
  if (r->err) return -3;
  if (r->is_eof) return -1;
  const double t0 = now_s();
  const int n = gzread(r->f, r->buf, FQR_BUF);
  r->inflate_s += now_s() - t0;
  r->begin = 0;
  if (n <= 0) {
    r->end = 0;
    r->is_eof = 1;
    if (n < 0) {
      r->err = 1;
      return -3;
    }
    return -1;
  }
  r->end = (size_t)n;
  r->bytes_in += (uint64_t)n;
  return 0;
}

static inline int get_byte(fqr_reader *r) { /* ks_getc */
  if (r->begin >= r->end) {
    const int e = refill(r);
    if (e) return e;
  }
  return r->buf[r->begin++];
}

static void reserve(fqr_str *s, size_t extra) {
  if (s->m - s->l < extra + 1) {
    size_t m = s->m ? s->m : 512;
    while (m - s->l < extra + 1) m *= 2;
    s->s = (uint8_t *)realloc(s->s, m);
    s->m = m;
  }
}

/* ks_getuntil2(KS_SEP_LINE, append): returns >= 0, or -1 when nothing at all could be read (EOF), -3 */
static int append_line(fqr_reader *r, fqr_str *str) {
  int got_any = 0;
  for (;;) {
    if (r->begin >= r->end) {
      const int e = refill(r);
      if (e == -3) return -3;
      if (e) break;
    }
    got_any = 1;
    const uint8_t *p = r->buf + r->begin;
    const size_t avail = r->end - r->begin;
    const uint8_t *nl = (const uint8_t *)memchr(p, '\n', avail);
    const size_t n = nl ? (size_t)(nl - p) : avail;
    reserve(str, n);
    memcpy(str->s + str->l, p, n);
    str->l += n;
    r->begin += n + (nl ? 1 : 0);
    if (nl) break;
  }
  if (!got_any) return -1;
  if (str->l > 1 && str->s[str->l - 1] == '\r') str->l--;
  return 0;
}

/* consumes the rest of the current line; returns the terminating '\n' or a negative code */
static int skip_line(fqr_reader *r) {
  for (;;) {
    if (r->begin >= r->end) {
      const int e = refill(r);
      if (e) return e;
    }
    const uint8_t *p = r->buf + r->begin;
    const uint8_t *nl = (const uint8_t *)memchr(p, '\n', r->end - r->begin);
    if (nl) {
      r->begin = (size_t)(nl - r->buf) + 1;
      return '\n';
    }
    r->begin = r->end;
  }
}

/* one record into r->seq / r->qual; returns its length or a negative kseq_read() code; *is_fasta set
 * when the record had no '+' line */
static long parse_record(fqr_reader *r, int *is_fasta) {
  int c;
  *is_fasta = 0;
  if (r->last_char == 0) {
    while ((c = get_byte(r)) >= 0 && c != '>' && c != '@') {
    }
    if (c < 0) return c;
    r->last_char = c;
  }
  r->seq.l = r->qual.l = 0;
  { /* name: up to the first white-space byte; -1 only if the header byte is the last byte of the stream */
    int got_any = 0;
    c = -1;
    for (;;) {
      if (r->begin >= r->end) {
        const int e = refill(r);
        if (e == -3) return -3;
        if (e) break;
      }
      got_any = 1;
      size_t i = r->begin;
      while (i < r->end && !isspace(r->buf[i])) i++;
      if (i < r->end) {
        c = r->buf[i];
        r->begin = i + 1;
        break;
      }
      r->begin = r->end;
    }
    if (!got_any) return -1;
    if (c >= 0 && c != '\n') {
      const int e = skip_line(r); /* comment */
      if (e == -3) return -3;
    }
  }
  while ((c = get_byte(r)) >= 0 && c != '>' && c != '+' && c != '@') {
    if (c == '\n') continue;
    reserve(&r->seq, 1);
    r->seq.s[r->seq.l++] = (uint8_t)c;
    append_line(r, &r->seq);
  }
  if (c == '>' || c == '@') r->last_char = c;
  if (c != '+') { /* FASTA record (or the stream ended inside the sequence) */
    *is_fasta = 1;
    return (long)r->seq.l;
  }
  c = skip_line(r); /* rest of the '+' line */
  if (c == -1) return -2;
  if (c == -3) return -3;
  while (append_line(r, &r->qual) >= 0 && r->qual.l < r->seq.l) {
  }
  r->last_char = 0;
  if (r->seq.l != r->qual.l) return -2;
  return (long)r->seq.l;
}

long fqr_next(fqr_reader *r, const uint8_t **seq, const uint8_t **qual, size_t *qual_len) {
  int fasta;
  const long l = parse_record(r, &fasta);
  if (l < 0) return l;
  reserve(&r->seq, 0);
  reserve(&r->qual, 0);
  *seq = r->seq.s;
  *qual = fasta ? NULL : r->qual.s;
  *qual_len = fasta ? 0 : r->qual.l;
  return l;
}

int fqr_fill(fqr_reader *r, uint8_t *seq, uint8_t *qual, uint32_t *offset, uint32_t *length, uint64_t cap_bytes,
             uint32_t cap_reads, uint32_t *n_reads, uint64_t *n_bytes, uint32_t *max_len) {
  uint32_t n = 0, longest = 0;
  uint64_t bytes = 0;
  int more = 1;
  while (r->status == 0 && n < cap_reads) {
    if (!r->pending) {
      int fasta;
      const long l = parse_record(r, &fasta);
      if (l < 0) {
        r->status = (int)l;
        break;
      }
      if (fasta) { /* the reference would read stale quality bytes here (quack.c:203): stop instead */
        r->status = -5;
        break;
      }
    }
    const size_t l = r->seq.l;
    if (l > cap_bytes) {
      r->status = -4;
      *n_reads = n, *n_bytes = bytes, *max_len = longest;
      return -1;
    }
    if (bytes + l > cap_bytes) { /* does not fit: keep it for the next batch */
      r->pending = 1;
      break;
    }
    r->pending = 0;
    memcpy(seq + bytes, r->seq.s, l);
    memcpy(qual + bytes, r->qual.s, l);
    offset[n] = (uint32_t)bytes;
    length[n] = (uint32_t)l;
    if (l > longest) longest = (uint32_t)l;
    bytes += l;
    n++;
  }
  if (r->status != 0) more = 0;
  *n_reads = n;
  *n_bytes = bytes;
  *max_len = longest;
  return more;
}

long fqr_read_adapter_keys(const char *path, uint32_t **keys_out) {
  fqr_reader *r = fqr_open(path);
  if (!r) return -1;
  uint32_t *keys = NULL;
  size_t n = 0, cap = 0;
  const uint8_t *s, *q;
  size_t ql;
  long l;
  while ((l = fqr_next(r, &s, &q, &ql)) >= 0) { /* quack.c:164 */
    if ((size_t)l > QB_KMER_SIZE) {
      const size_t add = (size_t)l - QB_KMER_SIZE;
      if (n + add > cap) {
        cap = (n + add) * 2 + 64;
        keys = (uint32_t *)realloc(keys, cap * sizeof *keys);
      }
      const int got = qb_adapter_record_keys((const char *)s, (size_t)l, keys + n, cap - n);
      if (got > 0) n += (size_t)got;
    }
  }
  fqr_close(r);
  if (!keys) keys = (uint32_t *)malloc(sizeof *keys);
  *keys_out = keys;
  return (long)n;
}
