/*
 * quack_oracle.c -- plain-C restatement of the reference hot path.  TEST INFRASTRUCTURE
 * ONLY (see quack_oracle.h).  Written for clarity, not speed: one read at a time, one
 * counter increment at a time, exactly the order of operations of the reference loop.
 */
#include "quack_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* ------------------------------------------------------------------ base code */

int qo_base_code(int c) {
  unsigned b = (unsigned)c & 0xFFu;
  /* reference domain: lookup[20] indexed by (c-65)&~32 (quack.c:150, 201): index 2 (C) -> 2,
   * index 6 (G) -> 3, index 19 (T) -> 1, all other indices 0..19 -> 0.  The bit tests below
   * give exactly that on [A-Ta-t] and define the rest of the byte range. */
  int cg = (b & 0x5Bu) == 0x43u;                                /* C,c,G,g */
  int lo = (b & 0x1Fu) == 0x07u || (b & 0x1Fu) == 0x14u;        /* G,g / T,t */
  return 2 * cg + lo;
}

/* ------------------------------------------------------------------ adapter table */

uint8_t *qo_table_new(void) { return (uint8_t *)calloc(QO_TABLE_SIZE, 1); }
void qo_table_free(uint8_t *table) { free(table); }

void qo_table_add_record(uint8_t *table, const char *seq, size_t l) {
  /* quack.c:165-172: index is first packed from seq[0..9] (no insertion), then every
   * further base is rolled in and the resulting index is inserted. */
  uint32_t index = 0;
  size_t i;
  if (l <= QO_KMER_SIZE) return; /* the insertion loop never runs */
  for (i = 0; i < QO_KMER_SIZE; i++)
    index = ((index << 2) + (uint32_t)qo_base_code((unsigned char)seq[i])) & (QO_TABLE_SIZE - 1);
  for (; i < l; i++) {
    index = ((index << 2) + (uint32_t)qo_base_code((unsigned char)seq[i])) & (QO_TABLE_SIZE - 1);
    table[index] = 1;
  }
}

uint32_t qo_table_keys(const uint8_t *table, uint32_t *out) {
  uint32_t n = 0, k;
  for (k = 0; k < QO_TABLE_SIZE; k++)
    if (table[k]) {
      if (out) out[n] = k;
      n++;
    }
  return n;
}

/* ------------------------------------------------------------------ accumulation */

int qo_stats_init(qo_stats *st) {
  memset(st, 0, sizeof *st);
  return 0;
}

void qo_stats_free(qo_stats *st) {
  free(st->rows);
  memset(st, 0, sizeof *st);
}

static int qo_grow(qo_stats *st, uint64_t need) {
  if (need <= st->cap) return 0;
  uint64_t cap = st->cap ? st->cap : 64;
  while (cap < need) cap *= 2;
  uint64_t *p = (uint64_t *)realloc(st->rows, cap * QO_ROW_U64 * sizeof(uint64_t));
  if (!p) return -1;
  memset(p + st->cap * QO_ROW_U64, 0, (cap - st->cap) * QO_ROW_U64 * sizeof(uint64_t));
  st->rows = p;
  st->cap = cap;
  return 0;
}

int qo_accumulate_read(qo_stats *st, const uint8_t *seq, const uint8_t *qual, size_t l,
                       const uint8_t *table) {
  size_t i;
  /* quack.c:194-198: rows exist (zeroed) up to the longest read seen */
  if (qo_grow(st, l > 11 ? l : 11)) return -1;
  if (l > st->max_length) st->max_length = l;

  /* quack.c:199-205 */
  for (i = 0; i < l; i++) {
    uint64_t *row = st->rows + i * QO_ROW_U64;
    row[QO_COL_CONTENT + qo_base_code(seq[i])]++;
    int q = (int)qual[i] - 33;
    if (q >= 0 && q < QO_SCORES)
      row[q]++;
    else
      st->n_invalid_qual++; /* reference: out-of-bounds write (UB) */
  }

  /* quack.c:206-217 */
  if (l > QO_KMER_SIZE) {
    uint32_t index = 0;
    for (i = 0; i < QO_KMER_SIZE; i++)
      index = ((index << 2) + (uint32_t)qo_base_code(seq[i])) & (QO_TABLE_SIZE - 1);
    if (table) {
      for (; table[index] == 0 && i < l; i++)
        index = ((index << 2) + (uint32_t)qo_base_code(seq[i])) & (QO_TABLE_SIZE - 1);
    }
    if (i < l) st->rows[i * QO_ROW_U64 + QO_COL_KMER]++;
  }
  /* l <= 10: the reference packs stale buffer bytes, then i == 10 >= l so nothing is counted */

  /* quack.c:219-220.  l == 0 indexes bases[-1] in the reference (UB); here the read is
   * counted in n_reads only. */
  if (l > 0) st->rows[(l - 1) * QO_ROW_U64 + QO_COL_LENGTH]++;
  st->n_reads++;
  return 0;
}

int qo_accumulate_batch(qo_stats *st, const uint8_t *seq, const uint8_t *qual,
                        const uint32_t *offset, const uint32_t *length, uint64_t n_reads,
                        const uint8_t *table) {
  uint64_t r;
  for (r = 0; r < n_reads; r++)
    if (qo_accumulate_read(st, seq + offset[r], qual + offset[r], length[r], table)) return -1;
  return 0;
}

/* ------------------------------------------------------------------ record framing */

typedef struct {
  uint8_t *s;
  size_t l, m;
} qo_str;

struct qo_reader {
  gzFile f;
  uint8_t *buf;
  int begin, end, is_eof, err;
  int last_char;
  qo_str name, seq, qual;
};

#define QO_BUFSZ 65536
enum { QO_DELIM_SPACE, QO_DELIM_LINE };

static int qo_fill(qo_reader *r) { /* kseq.h:72-76 / 103-107 */
  r->begin = 0;
  r->end = gzread(r->f, r->buf, QO_BUFSZ);
  if (r->end == 0) {
    r->is_eof = 1;
    return -1;
  }
  if (r->end < 0) {
    r->is_eof = 1;
    r->err = 1;
    r->end = 0;
    return -3;
  }
  return 0;
}

static int qo_getc(qo_reader *r) { /* ks_getc, kseq.h:67-79 */
  if (r->err) return -3;
  if (r->is_eof && r->begin >= r->end) return -1;
  if (r->begin >= r->end) {
    int e = qo_fill(r);
    if (e) return e;
  }
  return r->buf[r->begin++];
}

static void qo_str_reserve(qo_str *s, size_t extra) {
  if (s->m - s->l < extra + 1) {
    size_t m = s->m ? s->m : 256;
    while (m - s->l < extra + 1) m *= 2;
    s->s = (uint8_t *)realloc(s->s, m);
    s->m = m;
  }
}

/* ks_getuntil2, kseq.h:93-144: append bytes up to (not including) the delimiter */
static long qo_getuntil(qo_reader *r, int mode, qo_str *str, int *dret, int append) {
  int gotany = 0;
  if (dret) *dret = 0;
  if (!append) str->l = 0;
  for (;;) {
    int i;
    if (r->err) return -3;
    if (r->begin >= r->end) {
      if (r->is_eof) break;
      int e = qo_fill(r);
      if (e == -1) break;
      if (e == -3) return -3;
    }
    if (mode == QO_DELIM_LINE) {
      for (i = r->begin; i < r->end; i++)
        if (r->buf[i] == '\n') break;
    } else {
      for (i = r->begin; i < r->end; i++)
        if (isspace(r->buf[i])) break;
    }
    qo_str_reserve(str, (size_t)(i - r->begin));
    gotany = 1;
    memcpy(str->s + str->l, r->buf + r->begin, (size_t)(i - r->begin));
    str->l += (size_t)(i - r->begin);
    r->begin = i + 1;
    if (i < r->end) {
      if (dret) *dret = r->buf[i];
      break;
    }
  }
  if (!gotany && r->is_eof && r->begin >= r->end) return -1;
  qo_str_reserve(str, 0);
  if (mode == QO_DELIM_LINE && str->l > 1 && str->s[str->l - 1] == '\r') str->l--; /* kseq.h:141 */
  str->s[str->l] = 0;
  return (long)str->l;
}

qo_reader *qo_reader_open(const char *path) {
  gzFile f = gzopen(path, "r");
  if (!f) return NULL;
  qo_reader *r = (qo_reader *)calloc(1, sizeof *r);
  r->f = f;
  r->buf = (uint8_t *)malloc(QO_BUFSZ);
  return r;
}

void qo_reader_close(qo_reader *r) {
  if (!r) return;
  gzclose(r->f);
  free(r->buf);
  free(r->name.s);
  free(r->seq.s);
  free(r->qual.s);
  free(r);
}

long qo_reader_next(qo_reader *r, const uint8_t **seq, const uint8_t **qual, size_t *qual_len) {
  int c;
  long rc;
  if (r->last_char == 0) { /* kseq.h:182-186: skip to the next header character */
    while ((c = qo_getc(r)) >= 0 && c != '>' && c != '@') {
    }
    if (c < 0) return c;
    r->last_char = c;
  }
  r->seq.l = r->qual.l = 0;
  if ((rc = qo_getuntil(r, QO_DELIM_SPACE, &r->name, &c, 0)) < 0) return rc; /* kseq.h:188 */
  if (c != '\n') qo_getuntil(r, QO_DELIM_LINE, &r->name, NULL, 0);           /* comment, ignored */
  /* kseq.h:194-198: sequence lines until a line starting with '+', '>' or '@' */
  while ((c = qo_getc(r)) >= 0 && c != '>' && c != '+' && c != '@') {
    if (c == '\n') continue;
    qo_str_reserve(&r->seq, 1);
    r->seq.s[r->seq.l++] = (uint8_t)c;
    qo_getuntil(r, QO_DELIM_LINE, &r->seq, NULL, 1);
  }
  if (c == '>' || c == '@') r->last_char = c;
  qo_str_reserve(&r->seq, 0);
  r->seq.s[r->seq.l] = 0;
  *seq = r->seq.s;
  *qual = NULL;
  *qual_len = 0;
  if (c != '+') return (long)r->seq.l; /* FASTA record (also at EOF), kseq.h:206 */
  while ((c = qo_getc(r)) >= 0 && c != '\n') { /* rest of the '+' line, kseq.h:211 */
  }
  if (c == -1) return -2;
  /* kseq.h:213: quality lines until qual.l >= seq.l or nothing more can be read */
  while (qo_getuntil(r, QO_DELIM_LINE, &r->qual, NULL, 1) >= 0 && r->qual.l < r->seq.l) {
  }
  r->last_char = 0;
  qo_str_reserve(&r->qual, 0);
  *qual = r->qual.s;
  *qual_len = r->qual.l;
  if (r->seq.l != r->qual.l) return -2;
  return (long)r->seq.l;
}

long qo_read_adapters(const char *path, uint8_t *table) {
  qo_reader *r = qo_reader_open(path);
  const uint8_t *s, *q;
  size_t ql;
  long l, n = 0;
  if (!r) return -1;
  while ((l = qo_reader_next(r, &s, &q, &ql)) >= 0) { /* quack.c:164 */
    qo_table_add_record(table, (const char *)s, (size_t)l);
    n++;
  }
  qo_reader_close(r);
  return n;
}

int qo_read_fastq(const char *path, const uint8_t *table, qo_stats *st) {
  qo_reader *r = qo_reader_open(path);
  const uint8_t *s, *q;
  size_t ql;
  long l;
  if (!r) return -1;
  while ((l = qo_reader_next(r, &s, &q, &ql)) >= 0) { /* quack.c:193 */
    if (!q || ql != (size_t)l) break; /* FASTA record in a FASTQ stream: reference reads qual.s (UB) */
    if (qo_accumulate_read(st, s, q, (size_t)l, table)) break;
  }
  qo_reader_close(r);
  return 0;
}

/* ------------------------------------------------------------------ transform */

void qo_transform(qo_stats *st, uint64_t *original_max_length) {
  uint64_t i;
  int j;
  uint64_t *rows = st->rows;
  if (original_max_length) *original_max_length = st->max_length;
  if (st->max_length > 3000) { /* quack.c:234-262 */
    uint64_t unbinned, binned = 0;
    for (unbinned = 1; unbinned < st->max_length; unbinned++) {
      uint64_t *src = rows + unbinned * QO_ROW_U64;
      if (unbinned % 100 == 0) {
        binned++;
        uint64_t *z = rows + binned * QO_ROW_U64;
        for (j = 0; j < QO_COL_KMER; j++) z[j] = 0; /* scores, content, length -- not kmer */
      }
      uint64_t *dst = rows + binned * QO_ROW_U64;
      for (j = 0; j < QO_ROW_U64; j++) dst[j] = dst[j] + src[j];
    }
    st->max_length = binned;
  }
  for (i = 1; i < st->max_length; i++) /* quack.c:264-266 */
    rows[i * QO_ROW_U64 + QO_COL_KMER] += rows[(i - 1) * QO_ROW_U64 + QO_COL_KMER];
  for (i = 0; i < st->max_length; i++) { /* quack.c:269-291 */
    uint64_t *row = rows + i * QO_ROW_U64;
    int score_sum = 0;
    for (j = 0; j < QO_SCORES; j++) score_sum = (int)((uint64_t)score_sum + row[j]);
    if (score_sum != 0)
      for (j = 0; j < QO_SCORES; j++) row[j] = 100 * row[j] / (uint64_t)(int64_t)score_sum;
    /* single-precision product and quotient, then ceil in double (quack.c:288-289) */
    row[QO_COL_LENGTH] = (uint64_t)ceil(100 * (float)row[QO_COL_LENGTH] / st->n_reads);
    row[QO_COL_KMER] = (uint64_t)ceil(100 * (float)row[QO_COL_KMER] / (float)st->n_reads);
  }
}

/* ------------------------------------------------------------------ extras (PARITY UNPINNED)
 * The side outputs north_star names and the reference does not compute: there is no
 * reference line, golden vector or fixture for them, so this restatement only pins the CUDA path against an
 * independent scalar definition:
 *   n_count[p]   reads whose base p is 'N' or 'n'
 *   qual_sum[p]  sum over the reads of (q - 33) at position p, for quality bytes in the heatmap's range [33, 123]
 *                (what sum_s s * scores[p][s] over the rows of quack.c:203-204 gives)
 *   mean_hist[m] reads with floor(sum_p clamp(q, 33, 126) - 33) / l) == m, m = 0..93 */
void qo_extras(const uint8_t *seq, const uint8_t *qual, const uint32_t *offset, const uint32_t *length, uint64_t n_reads,
               uint64_t *n_count, uint64_t *qual_sum, uint64_t rows, uint64_t *mean_hist) {
  for (uint64_t r = 0; r < n_reads; r++) {
    const uint8_t *s = seq + offset[r], *q = qual + offset[r];
    const uint32_t l = length[r];
    if (l == 0 || l > rows) continue;
    uint64_t sum = 0;
    for (uint32_t p = 0; p < l; p++) {
      if (s[p] == 'N' || s[p] == 'n') n_count[p]++;
      if (q[p] >= 33 && q[p] <= 123) qual_sum[p] += (uint64_t)(q[p] - 33);
      const unsigned c = q[p] < 33 ? 33u : q[p] > 126 ? 126u : q[p];
      sum += c - 33u;
    }
    mean_hist[sum / l]++;
  }
}
