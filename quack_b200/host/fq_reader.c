/* fq_reader.c -- see fq_reader.h.  Written against the behaviour of kseq_read() (klib/kseq.h:177-218):
 *   - a record starts at the next '@' or '>' (anything before the first header is skipped);
 *   - name = up to the first white space, rest of the line ignored;
 *   - sequence = following lines up to a line starting with '+', '>' or '@'; empty lines skipped; a
 *     trailing '\r' is dropped when the accumulated string is longer than one byte (kseq.h:141);
 *   - the '+' line is skipped; quality = following lines until at least as many bytes as the sequence;
 *   - quality length != sequence length -> -2 and the stream is over for quack (quack.c:193).
 * Lines are located with memchr over a 1 MiB inflate buffer instead of kseq's per-byte loops.
 *
 * Input decode: a plain / gzip / multi-member gzip file goes through gzread()
 * like the reference (quack.c:187, kseq.h:74).  A BGZF file (the blocked gzip of bgzip / htslib, reference
 * klib/bgzf.c:63-71: every member carries its compressed size in a 'BC' extra subfield and inflates to at
 * most 64 KiB on its own) is inflated by a pool of threads, one block per thread at a time, and handed to
 * the framing code in file order -- the same bytes gzread() would deliver, so everything downstream is
 * unchanged. */
#include "fq_reader.h"

#include <ctype.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>

#include "../../include/quack_b200.h"

#define FQR_BUF (1u << 20)

typedef struct {
  uint8_t *s;
  size_t l, m;
} fqr_str;

/* ---- BGZF block pool ---- */
#define BGZF_MAX_BLOCK 65536u
typedef struct {
  uint8_t *cbuf; /* deflate payload + 8-byte trailer of one block */
  uint8_t *ubuf; /* inflated bytes */
  uint32_t csize, usize;
  int state; /* 0 free, 1 being inflated, 2 ready for the consumer */
  int err;   /* 0, -1 end of file in front of this block, -3 broken block, -6 a gzip member that is not BGZF */
} bgzf_slot;

typedef struct {
  FILE *fp;
  int n_threads, n_slots;
  bgzf_slot *slot;
  pthread_t *th;
  pthread_mutex_t mu;
  pthread_cond_t cv_ready, cv_free;
  uint64_t next_ticket;  /* next block to read from the file */
  uint64_t next_consume; /* next block the framing code takes */
  int input_done;        /* a reader saw the end of the file or a broken block: no more tickets */
  int stop;
  double inflate_cpu_s;  /* summed over the workers */
  long long switch_off;  /* err -6: file offset of the member the serial zlib path continues from */
} bgzf_pool;

/* ---- multi-member gzip pool ----
 * An ordinary .gz file may be a concatenation of gzip members (RFC 1952 section 2.2; what `cat a.gz b.gz`, pigz -i,
 * and the benchmark generator write; gzread() -- the reference, quack.c:187 -- decodes them back to back).  Members
 * are independent deflate streams, so they can be inflated in parallel once their starts are known.  They are not
 * marked in any index: the pool SPECULATES.  A scanner looks for byte strings that look like a member header
 * (1f 8b 08, reserved flag bits clear, a plausible XFL / OS byte); every candidate is inflated by a worker with
 * zlib's own gzip wrapper (header, CRC32 and ISIZE checked).  The consumer walks the chain in file order: the
 * member that starts where the previous one ENDED is the next one, whatever the scanner guessed -- a candidate
 * inside a member's data is simply never asked for (its worker usually fails on the first block), a member whose
 * header the scanner did not recognise is inflated on demand.  What reaches the framing code is byte for byte
 * what gzread() would deliver: trailing bytes that start no member end the stream silently, a damaged member
 * yields the bytes in front of the damage and then -3.
 * A single-member file has one candidate: one worker streams it through 1 MiB chunks (no parallelism, but the
 * inflate overlaps the framing and packing of the reader thread). */
#define MGZ_CHUNK FQR_BUF
typedef struct mgz_chunk {
  struct mgz_chunk *next;
  uint32_t len;
  uint8_t *data; /* MGZ_CHUNK bytes */
} mgz_chunk;

typedef struct mgz_pool mgz_pool;
typedef struct mgz_job {
  mgz_pool *pool;
  uint64_t start, end; /* compressed offsets; end valid when done == 1 */
  int claimed, done;   /* done: 0 running, 1 stream end reached, 2 the file ended inside the member, -3 error */
  int cancel;          /* the consumer went past this candidate: stop */
  mgz_chunk *head, *tail;
} mgz_job;

struct mgz_pool {
  const uint8_t *map;
  uint64_t size;
  int fd;
  int n_threads;
  pthread_t *th;
  pthread_t *extra; /* one thread per member that had to be inflated on demand (rare) */
  size_t n_extra;
  pthread_mutex_t mu;
  pthread_cond_t cv_data, cv_room;
  mgz_job **job; /* candidates in file order (the nodes never move) */
  size_t n_jobs, cap_jobs;
  size_t next_claim;  /* first job no worker has taken yet */
  uint64_t scan_pos;  /* the scanner has looked at every offset below this */
  uint64_t expect;    /* compressed offset where the consumer's next member starts */
  mgz_job *cur;       /* consumer: the job that starts at `expect`, NULL until it is looked up */
  int stop;
  uint64_t budget;    /* bytes of inflated data the pool may hold ahead of the consumer */
  uint64_t held;
  double inflate_cpu_s;
};

struct fqr_reader {
  gzFile f;
  bgzf_pool *pool; /* != NULL: BGZF input decoded by the pool, f unused */
  mgz_pool *mgz;   /* != NULL: gzip members inflated by the pool, f unused */
  char *path;      /* to reopen the file when a BGZF file continues with ordinary gzip members */
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

/* Reads the next BGZF block of fp into s (payload + trailer).  0 ok; -1 end of the stream: clean end of file, or
 * bytes that do not start a gzip member (gzread ignores such trailing garbage the same way); -3 truncated block;
 * -6 a gzip member that is not a BGZF block (*member_off = where it starts: the caller hands the rest of the file
 * to zlib, which is what gzread would have done all along).  Layout (RFC 1952 + SAM spec 4.1; reference
 * klib/bgzf.c:63-71, 330-355): 10-byte gzip header with FLG = FEXTRA, XLEN, extra subfields of which one is
 * 'B' 'C' 2 BSIZE (block size - 1), deflate data, CRC32, ISIZE. */
static int bgzf_read_block(FILE *fp, bgzf_slot *s, long long *member_off) {
  uint8_t h[12];
  *member_off = (long long)ftello(fp);
  const size_t got = fread(h, 1, 12, fp);
  if (got < 2 || h[0] != 0x1f || h[1] != 0x8b) return -1;
  if (got != 12 || h[2] != 8 || h[3] != 4) return -6;
  const uint32_t xlen = (uint32_t)h[10] | (uint32_t)h[11] << 8;
  uint8_t extra[512];
  if (xlen < 6 || xlen > sizeof extra) return -6;
  if (fread(extra, 1, xlen, fp) != xlen) return -3;
  uint32_t bsize = 0;
  for (uint32_t i = 0; i + 4 <= xlen;) {
    const uint32_t slen = (uint32_t)extra[i + 2] | (uint32_t)extra[i + 3] << 8;
    if (extra[i] == 'B' && extra[i + 1] == 'C' && slen == 2 && i + 6 <= xlen) bsize = ((uint32_t)extra[i + 4] | (uint32_t)extra[i + 5] << 8) + 1u;
    i += 4 + slen;
  }
  if (bsize == 0) return -6;
  if (bsize < 12 + xlen + 8 || bsize > BGZF_MAX_BLOCK) return -3;
  s->csize = bsize - 12 - xlen;
  if (fread(s->cbuf, 1, s->csize, fp) != s->csize) return -3;
  return 0;
}

static void *bgzf_worker(void *arg) {
  bgzf_pool *p = (bgzf_pool *)arg;
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, -15) != Z_OK) return NULL;
  double cpu = 0;
  for (;;) {
    pthread_mutex_lock(&p->mu);
    while (!p->stop && !p->input_done && p->slot[p->next_ticket % (uint64_t)p->n_slots].state != 0)
      pthread_cond_wait(&p->cv_free, &p->mu);
    if (p->stop || p->input_done) {
      pthread_mutex_unlock(&p->mu);
      break;
    }
    bgzf_slot *s = &p->slot[p->next_ticket++ % (uint64_t)p->n_slots];
    long long member_off = 0;
    const int rc = bgzf_read_block(p->fp, s, &member_off); /* sequential file order: under the lock */
    s->err = rc;
    s->usize = 0;
    if (rc == -6) p->switch_off = member_off;
    if (rc) {
      p->input_done = 1;
      s->state = 2;
      pthread_cond_broadcast(&p->cv_ready);
      pthread_cond_broadcast(&p->cv_free);
      pthread_mutex_unlock(&p->mu);
      break;
    }
    s->state = 1;
    pthread_mutex_unlock(&p->mu);

    const double t0 = now_s();
    int err = 0;
    uint32_t out = 0;
    inflateReset(&zs);
    zs.next_in = s->cbuf;
    zs.avail_in = s->csize - 8;
    zs.next_out = s->ubuf;
    zs.avail_out = BGZF_MAX_BLOCK;
    if (inflate(&zs, Z_FINISH) != Z_STREAM_END) {
      err = -3;
    } else {
      out = BGZF_MAX_BLOCK - zs.avail_out;
      const uint8_t *t = s->cbuf + s->csize - 8;
      const uint32_t crc = (uint32_t)t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
      const uint32_t isize = (uint32_t)t[4] | (uint32_t)t[5] << 8 | (uint32_t)t[6] << 16 | (uint32_t)t[7] << 24;
      if (isize != out || (uint32_t)crc32(crc32(0L, Z_NULL, 0), s->ubuf, out) != crc) err = -3; /* as gzread checks */
    }
    cpu += now_s() - t0;

    pthread_mutex_lock(&p->mu);
    s->usize = err ? 0 : out;
    s->err = err;
    s->state = 2;
    pthread_cond_broadcast(&p->cv_ready);
    pthread_mutex_unlock(&p->mu);
  }
  inflateEnd(&zs);
  pthread_mutex_lock(&p->mu);
  p->inflate_cpu_s += cpu;
  pthread_mutex_unlock(&p->mu);
  return NULL;
}

static bgzf_pool *bgzf_pool_open(FILE *fp, int n_threads) {
  bgzf_pool *p = (bgzf_pool *)calloc(1, sizeof *p);
  p->fp = fp;
  p->n_threads = n_threads;
  p->n_slots = 64 * n_threads > 1024 ? 1024 : 64 * n_threads; /* up to 64 MiB of text ahead of the framing code */
  p->slot = (bgzf_slot *)calloc((size_t)p->n_slots, sizeof *p->slot);
  for (int i = 0; i < p->n_slots; i++) {
    p->slot[i].cbuf = (uint8_t *)malloc(BGZF_MAX_BLOCK);
    p->slot[i].ubuf = (uint8_t *)malloc(BGZF_MAX_BLOCK);
  }
  pthread_mutex_init(&p->mu, NULL);
  pthread_cond_init(&p->cv_ready, NULL);
  pthread_cond_init(&p->cv_free, NULL);
  p->th = (pthread_t *)calloc((size_t)n_threads, sizeof *p->th);
  int started = 0;
  for (int i = 0; i < n_threads; i++)
    if (pthread_create(&p->th[started], NULL, bgzf_worker, p) == 0) started++;
  p->n_threads = started; /* fewer threads than asked for still decode the file; none: the caller falls back */
  return p;
}

static void bgzf_pool_close(bgzf_pool *p) {
  pthread_mutex_lock(&p->mu);
  p->stop = 1;
  pthread_cond_broadcast(&p->cv_free);
  pthread_mutex_unlock(&p->mu);
  for (int i = 0; i < p->n_threads; i++) pthread_join(p->th[i], NULL);
  for (int i = 0; i < p->n_slots; i++) {
    free(p->slot[i].cbuf);
    free(p->slot[i].ubuf);
  }
  free(p->slot);
  free(p->th);
  pthread_mutex_destroy(&p->mu);
  pthread_cond_destroy(&p->cv_ready);
  pthread_cond_destroy(&p->cv_free);
  fclose(p->fp);
  free(p);
}

/* ---- multi-member gzip pool: implementation ---- */

/* does a gzip member header plausibly start at map[o]?  (RFC 1952: ID1 ID2 CM FLG MTIME[4] XFL OS) */
static int mgz_header_at(const mgz_pool *p, uint64_t o) {
  if (o + 18 > p->size) return 0;
  const uint8_t *h = p->map + o;
  return h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 0xE0) == 0 && (h[8] == 0 || h[8] == 2 || h[8] == 4) &&
         (h[9] <= 13 || h[9] == 255);
}

static mgz_job *mgz_insert_job(mgz_pool *p, size_t at, uint64_t start) { /* mu held */
  if (p->n_jobs == p->cap_jobs) {
    p->cap_jobs = p->cap_jobs ? 2 * p->cap_jobs : 64;
    p->job = (mgz_job **)realloc(p->job, p->cap_jobs * sizeof *p->job);
  }
  memmove(&p->job[at + 1], &p->job[at], (p->n_jobs - at) * sizeof *p->job);
  p->n_jobs++;
  mgz_job *j = (mgz_job *)calloc(1, sizeof *j);
  j->pool = p;
  j->start = start;
  p->job[at] = j;
  return j;
}

/* looks for the next candidate at or behind scan_pos and appends it; 0 if the file holds no more (mu held) */
static int mgz_scan_next(mgz_pool *p) {
  uint64_t o = p->scan_pos;
  while (o + 18 <= p->size) {
    const uint8_t *hit = (const uint8_t *)memchr(p->map + o, 0x1f, (size_t)(p->size - 17 - o));
    if (!hit) break;
    o = (uint64_t)(hit - p->map);
    if (mgz_header_at(p, o)) {
      mgz_insert_job(p, p->n_jobs, o);
      p->scan_pos = o + 18;
      return 1;
    }
    o++;
  }
  p->scan_pos = p->size;
  return 0;
}

/* inflates one member (zlib's gzip wrapper checks header, CRC32 and ISIZE) into 1 MiB chunks on the job's list */
static void mgz_inflate_job(mgz_job *j, z_stream *zs) {
  mgz_pool *p = j->pool;
  inflateReset2(zs, 15 + 16);
  zs->avail_in = 0;
  uint64_t in_pos = j->start;
  int status = 0;
  while (status == 0) {
    mgz_chunk *c = (mgz_chunk *)malloc(sizeof *c);
    c->data = (uint8_t *)malloc(MGZ_CHUNK);
    c->next = NULL;
    zs->next_out = c->data;
    zs->avail_out = MGZ_CHUNK;
    while (zs->avail_out && status == 0) {
      if (zs->avail_in == 0) {
        const uint64_t left = p->size - in_pos;
        if (left == 0) { /* the file ends inside the member: gzread() hands out what it has and stops without an
                            error (Z_BUF_ERROR, "unexpected end of file", is not fatal to it) */
          status = 2;
          break;
        }
        const uInt n = left > (1u << 30) ? (1u << 30) : (uInt)left;
        zs->next_in = (Bytef *)(p->map + in_pos);
        zs->avail_in = n;
        in_pos += n;
      }
      const int rc = inflate(zs, Z_NO_FLUSH);
      if (rc == Z_STREAM_END)
        status = 1;
      else if (rc != Z_OK && rc != Z_BUF_ERROR)
        status = -3;
    }
    c->len = MGZ_CHUNK - zs->avail_out;
    pthread_mutex_lock(&p->mu);
    if (c->len) {
      if (j->tail)
        j->tail->next = c;
      else
        j->head = c;
      j->tail = c;
      p->held += c->len;
    } else {
      free(c->data);
      free(c);
    }
    if (status) {
      j->end = in_pos - zs->avail_in;
      j->done = status;
    }
    pthread_cond_broadcast(&p->cv_data);
    /* a member ahead of the consumer waits here while the pool holds too much; the consumer's own never does */
    while (!status && p->held >= p->budget && p->cur != j && !p->stop && !j->cancel) pthread_cond_wait(&p->cv_room, &p->mu);
    if (!status && (p->stop || j->cancel)) {
      j->done = -3;
      status = -3;
      pthread_cond_broadcast(&p->cv_data);
    }
    pthread_mutex_unlock(&p->mu);
  }
}

static void *mgz_worker(void *arg) {
  mgz_pool *p = (mgz_pool *)arg;
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, 15 + 16) != Z_OK) return NULL;
  double cpu = 0;
  for (;;) {
    pthread_mutex_lock(&p->mu);
    mgz_job *j = NULL;
    while (!p->stop && !j) {
      /* candidates the consumer already passed are never claimed */
      while (p->next_claim < p->n_jobs && (p->job[p->next_claim]->claimed || p->job[p->next_claim]->start < p->expect)) p->next_claim++;
      if (p->cur && !p->cur->claimed)
        j = p->cur;
      else if (p->held >= p->budget)
        pthread_cond_wait(&p->cv_room, &p->mu); /* speculation is bounded */
      else if (p->next_claim < p->n_jobs)
        j = p->job[p->next_claim];
      else if (!(p->scan_pos < p->size && mgz_scan_next(p)))
        pthread_cond_wait(&p->cv_room, &p->mu); /* nothing left to look for */
    }
    if (p->stop) {
      pthread_mutex_unlock(&p->mu);
      break;
    }
    j->claimed = 1;
    pthread_mutex_unlock(&p->mu);
    const double t0 = now_s();
    mgz_inflate_job(j, &zs);
    cpu += now_s() - t0;
  }
  inflateEnd(&zs);
  pthread_mutex_lock(&p->mu);
  p->inflate_cpu_s += cpu;
  pthread_mutex_unlock(&p->mu);
  return NULL;
}

static void *mgz_one_job(void *arg) { /* a member nobody had scheduled: its own thread */
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, 15 + 16) == Z_OK) {
    mgz_inflate_job((mgz_job *)arg, &zs);
    inflateEnd(&zs);
  } else {
    mgz_job *j = (mgz_job *)arg;
    pthread_mutex_lock(&j->pool->mu);
    j->done = -3;
    pthread_cond_broadcast(&j->pool->cv_data);
    pthread_mutex_unlock(&j->pool->mu);
  }
  return NULL;
}

static void mgz_free_job_chunks(mgz_pool *p, mgz_job *j) { /* mu held */
  while (j->head) {
    mgz_chunk *c = j->head;
    j->head = c->next;
    p->held -= c->len;
    free(c->data);
    free(c);
  }
  j->tail = NULL;
}

static mgz_pool *mgz_pool_open(const char *path, int n_threads) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return NULL;
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 18) {
    close(fd);
    return NULL;
  }
  void *map = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  if (map == MAP_FAILED) {
    close(fd);
    return NULL;
  }
  madvise(map, (size_t)st.st_size, MADV_SEQUENTIAL);
  mgz_pool *p = (mgz_pool *)calloc(1, sizeof *p);
  p->map = (const uint8_t *)map;
  p->size = (uint64_t)st.st_size;
  p->fd = fd;
  p->budget = (uint64_t)(n_threads + 2) * (96ull << 20);
  pthread_mutex_init(&p->mu, NULL);
  pthread_cond_init(&p->cv_data, NULL);
  pthread_cond_init(&p->cv_room, NULL);
  p->cur = mgz_insert_job(p, 0, 0);
  p->scan_pos = 18;
  p->th = (pthread_t *)calloc((size_t)n_threads, sizeof *p->th);
  int started = 0;
  for (int i = 0; i < n_threads; i++)
    if (pthread_create(&p->th[started], NULL, mgz_worker, p) == 0) started++;
  p->n_threads = started;
  return p;
}

static void mgz_pool_close(mgz_pool *p) {
  pthread_mutex_lock(&p->mu);
  p->stop = 1;
  pthread_cond_broadcast(&p->cv_room);
  pthread_cond_broadcast(&p->cv_data);
  pthread_mutex_unlock(&p->mu);
  for (int i = 0; i < p->n_threads; i++) pthread_join(p->th[i], NULL);
  for (size_t i = 0; i < p->n_extra; i++) pthread_join(p->extra[i], NULL);
  for (size_t i = 0; i < p->n_jobs; i++) {
    mgz_free_job_chunks(p, p->job[i]);
    free(p->job[i]);
  }
  free(p->job);
  free(p->th);
  free(p->extra);
  pthread_mutex_destroy(&p->mu);
  pthread_cond_destroy(&p->cv_data);
  pthread_cond_destroy(&p->cv_room);
  munmap((void *)p->map, (size_t)p->size);
  close(p->fd);
  free(p);
}

/* next chunk of inflated bytes in file order into *buf (buffers are swapped, not copied); 0 ok, -1 end, -3 error */
static int mgz_next_chunk(mgz_pool *p, uint8_t **buf, size_t *len) {
  pthread_mutex_lock(&p->mu);
  for (;;) {
    if (!p->cur) {
      /* the member that starts exactly where the last one ended */
      if (p->expect + 18 > p->size || p->map[p->expect] != 0x1f || p->map[p->expect + 1] != 0x8b) {
        pthread_mutex_unlock(&p->mu); /* end of file, or trailing bytes that start no member: gzread stops silently */
        return -1;
      }
      size_t i = 0;
      while (i < p->n_jobs && p->job[i]->start < p->expect) {
        if (!p->job[i]->cancel) { /* a candidate inside the previous member's data (or a member already delivered) */
          p->job[i]->cancel = 1;
          mgz_free_job_chunks(p, p->job[i]);
        }
        i++;
      }
      if (i < p->n_jobs && p->job[i]->start == p->expect) {
        p->cur = p->job[i];
      } else {
        /* the scanner has not got here yet, or did not take these bytes for a header: the member gets a thread of
         * its own (the pool's workers may all be waiting for room behind members further down the file) */
        p->cur = mgz_insert_job(p, i, p->expect);
        p->cur->claimed = 1;
        if (p->next_claim > i) p->next_claim = i;
        if (p->scan_pos < p->expect + 18) p->scan_pos = p->expect + 18;
        p->extra = (pthread_t *)realloc(p->extra, (p->n_extra + 1) * sizeof *p->extra);
        if (pthread_create(&p->extra[p->n_extra], NULL, mgz_one_job, p->cur) == 0)
          p->n_extra++;
        else
          p->cur->done = -3;
      }
      pthread_cond_broadcast(&p->cv_room);
    }
    mgz_job *j = p->cur;
    if (j->head) {
      mgz_chunk *c = j->head;
      j->head = c->next;
      if (!j->head) j->tail = NULL;
      p->held -= c->len;
      uint8_t *old = *buf;
      *buf = c->data;
      *len = c->len;
      pthread_cond_broadcast(&p->cv_room);
      pthread_mutex_unlock(&p->mu);
      free(old);
      free(c);
      return 0;
    }
    if (j->done == 1) {
      p->expect = j->end;
      p->cur = NULL;
      continue;
    }
    if (j->done == 2) {
      pthread_mutex_unlock(&p->mu);
      return -1;
    }
    if (j->done < 0) {
      pthread_mutex_unlock(&p->mu);
      return -3;
    }
    pthread_cond_wait(&p->cv_data, &p->mu);
  }
}

/* 1 if the file starts with a BGZF block header */
static int is_bgzf(FILE *fp) {
  uint8_t h[18];
  const size_t got = fread(h, 1, sizeof h, fp);
  if (fseeko(fp, 0, SEEK_SET) != 0) return 0;
  return got == sizeof h && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && h[3] == 4 && h[10] == 6 && h[11] == 0 &&
         h[12] == 'B' && h[13] == 'C' && h[14] == 2 && h[15] == 0;
}

int fqr_default_threads(void) {
  const char *e = getenv("QUACK_DECODE_THREADS");
  if (e && atoi(e) >= 1) return atoi(e) > 64 ? 64 : atoi(e);
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  if (n < 2) return 1;
  n /= 2; /* two mates are decoded concurrently */
  return n > 8 ? 8 : (int)n;
}

fqr_reader *fqr_open_mt(const char *path, int threads) {
  if (threads < 1) threads = fqr_default_threads();
  /* The BGZF probe reads the head of the file and seeks back: only on regular files.  A pipe, a FIFO, /dev/stdin or
   * a process substitution is opened exactly once, by gzopen(), like the reference does (quack.c:187). */
  struct stat st;
  if (threads > 1 && stat(path, &st) == 0 && S_ISREG(st.st_mode)) {
    FILE *fp = fopen(path, "rb");
    if (!fp) return NULL;
    if (is_bgzf(fp)) {
      bgzf_pool *pool = bgzf_pool_open(fp, threads);
      if (pool->n_threads > 0) {
        fqr_reader *r = (fqr_reader *)calloc(1, sizeof *r);
        r->pool = pool;
        r->buf = (uint8_t *)malloc(BGZF_MAX_BLOCK);
        r->path = strdup(path);
        return r;
      }
      bgzf_pool_close(pool); /* no thread could be started: gzread path below (closes fp) */
    } else {
      uint8_t h[3] = {0, 0, 0};
      const size_t got = fread(h, 1, 3, fp);
      fclose(fp);
      if (got == 3 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && !getenv("QUACK_NO_MGZ")) {
        /* ordinary gzip: members inflated in parallel where the file has several (and ahead of the framing code
         * on a thread of its own where it has one) */
        mgz_pool *mp = mgz_pool_open(path, threads);
        if (mp && mp->n_threads > 0) {
          fqr_reader *r = (fqr_reader *)calloc(1, sizeof *r);
          r->mgz = mp;
          r->buf = (uint8_t *)malloc(MGZ_CHUNK);
          return r;
        }
        if (mp) mgz_pool_close(mp);
      }
    }
  }
  gzFile f = gzopen(path, "r");
  if (!f) return NULL;
  fqr_reader *r = (fqr_reader *)calloc(1, sizeof *r);
  r->f = f;
  r->buf = (uint8_t *)malloc(FQR_BUF);
  return r;
}

fqr_reader *fqr_open(const char *path) { return fqr_open_mt(path, 0); }

int fqr_decode_threads(const fqr_reader *r) { return r->pool ? r->pool->n_threads : r->mgz ? r->mgz->n_threads : 1; }

void fqr_close(fqr_reader *r) {
  if (!r) return;
  if (r->pool)
    bgzf_pool_close(r->pool);
  else if (r->mgz)
    mgz_pool_close(r->mgz);
  else if (r->f)
    gzclose(r->f);
  free(r->path);
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
  if (r->err) {
    r->is_eof = 1;
    return -3;
  }
  if (r->is_eof) return -1;
  if (r->mgz) { /* next chunk of the current gzip member, the members in file order */
    const double t0 = now_s();
    for (;;) {
      size_t n = 0;
      const int rc = mgz_next_chunk(r->mgz, &r->buf, &n);
      r->begin = 0;
      r->end = rc ? 0 : n;
      if (rc) {
        r->is_eof = 1;
        r->inflate_s += now_s() - t0;
        if (rc == -3) {
          r->err = 1;
          return -3;
        }
        return -1;
      }
      if (n) break;
    }
    r->bytes_in += r->end;
    r->inflate_s += now_s() - t0; /* time the framing code waited for the pool */
    return 0;
  }
  if (r->pool) { /* next inflated block, in file order; empty blocks (the BGZF end marker) are skipped */
    bgzf_pool *p = r->pool;
    const double t0 = now_s();
    for (;;) {
      pthread_mutex_lock(&p->mu);
      bgzf_slot *s = &p->slot[p->next_consume % (uint64_t)p->n_slots];
      while (s->state != 2) pthread_cond_wait(&p->cv_ready, &p->mu);
      const int err = s->err;
      uint32_t n = 0;
      if (!err) {
        uint8_t *t = r->buf; /* swap buffers instead of copying */
        r->buf = s->ubuf;
        s->ubuf = t;
        n = s->usize;
        s->state = 0;
        p->next_consume++;
        pthread_cond_broadcast(&p->cv_free);
      }
      pthread_mutex_unlock(&p->mu);
      r->begin = 0;
      r->end = n;
      if (err == -6) {
        /* the file goes on with an ordinary gzip member: the blocks in front of it were delivered, zlib takes
         * the rest from that member on -- the bytes gzread() would have produced for the whole file */
        const long long off = p->switch_off;
        bgzf_pool_close(p);
        r->pool = NULL;
        const int fd = open(r->path, O_RDONLY);
        if (fd < 0 || lseek(fd, (off_t)off, SEEK_SET) < 0 || !(r->f = gzdopen(fd, "r"))) {
          if (fd >= 0) close(fd);
          r->is_eof = r->err = 1;
          return -3;
        }
        gzbuffer(r->f, 1u << 18);
        free(r->buf); /* a 64 KiB block buffer so far (fully consumed): the gzread path fills FQR_BUF at a time */
        r->buf = (uint8_t *)malloc(FQR_BUF);
        r->begin = r->end = 0;
        r->inflate_s += now_s() - t0;
        return refill(r);
      }
      if (err) {
        r->is_eof = 1;
        r->inflate_s += now_s() - t0;
        if (err == -3) {
          r->err = 1;
          return -3;
        }
        return -1;
      }
      if (n) break;
    }
    r->bytes_in += r->end;
    r->inflate_s += now_s() - t0; /* time the framing code waited for the pool */
    return 0;
  }
  /* gzread() in requests of 16 KiB, the size kseq asks for (klib/kseq.h:74,228): a request that runs into damaged
   * data fails as a whole, so with the reference's request size the bytes in front of the damage that reach the
   * framing code are the reference's too */
  const double t0 = now_s();
  size_t filled = 0;
  int n = 0;
  while (filled + 16384 <= FQR_BUF) {
    n = gzread(r->f, r->buf + filled, 16384);
    if (n <= 0) break;
    filled += (size_t)n;
    if (n < 16384) break;
  }
  r->inflate_s += now_s() - t0;
  r->begin = 0;
  r->end = filled;
  r->bytes_in += (uint64_t)filled;
  if (n < 0) r->err = 1; /* reported once the bytes in front of it are consumed */
  if (filled == 0) {
    r->is_eof = 1;
    return n < 0 ? -3 : -1;
  }
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

/* Fast path of fqr_fill for the canonical record -- '@' header line, ONE non-empty sequence line, a '+' line,
 * ONE quality line of the same length, no '\r', all four newlines inside the current buffer: the bytes go
 * straight from the inflate buffer into the batch, with exactly the result parse_record() would give (every
 * other shape -- multi-line records, '\r', a record cut by the buffer end, FASTA, garbage -- returns 0 and
 * takes parse_record()).  Returns 1 and sets s/q/l without consuming anything; the caller commits with
 * r->begin = *next. */
static inline int fast_record(const fqr_reader *r, const uint8_t **s, const uint8_t **q, size_t *l, size_t *next) {
  if (r->last_char != 0 || r->pending || r->begin >= r->end) return 0;
  const uint8_t *p = r->buf + r->begin, *end = r->buf + r->end;
  if (*p != '@') return 0;
  const uint8_t *n1 = (const uint8_t *)memchr(p, '\n', (size_t)(end - p));
  if (!n1 || n1 + 1 >= end) return 0;
  const uint8_t *sq = n1 + 1;
  if (*sq == '\n' || *sq == '>' || *sq == '+' || *sq == '@') return 0;
  const uint8_t *n2 = (const uint8_t *)memchr(sq, '\n', (size_t)(end - sq));
  if (!n2 || n2 + 1 >= end || n2[-1] == '\r' || n2[1] != '+') return 0;
  const uint8_t *n3 = (const uint8_t *)memchr(n2 + 1, '\n', (size_t)(end - (n2 + 1)));
  if (!n3) return 0;
  const size_t len = (size_t)(n2 - sq);
  const uint8_t *ql = n3 + 1;
  if ((size_t)(end - ql) < len + 1 || ql[len] != '\n' || ql[len - 1] == '\r') return 0;
  if (memchr(ql, '\n', len)) return 0; /* a shorter first quality line: parse_record() joins lines */
  *s = sq, *q = ql, *l = len, *next = (size_t)(ql + len + 1 - r->buf);
  return 1;
}

int fqr_fill(fqr_reader *r, uint8_t *seq, uint8_t *qual, uint32_t *offset, uint32_t *length, uint64_t cap_bytes,
             uint32_t cap_reads, uint32_t *n_reads, uint64_t *n_bytes, uint32_t *max_len) {
  uint32_t n = 0, longest = 0;
  uint64_t bytes = 0;
  int more = 1;
  while (r->status == 0 && n < cap_reads) {
    const uint8_t *fs, *fq;
    size_t fl, fnext;
    if (fast_record(r, &fs, &fq, &fl, &fnext)) {
      if (fl > cap_bytes) { /* same outcome as below, through the general path */
      } else if (bytes + fl > cap_bytes) {
        break; /* does not fit: nothing was consumed, the next batch starts with it */
      } else {
        memcpy(seq + bytes, fs, fl);
        memcpy(qual + bytes, fq, fl);
        offset[n] = (uint32_t)bytes;
        length[n] = (uint32_t)fl;
        if (fl > longest) longest = (uint32_t)fl;
        bytes += fl;
        n++;
        r->begin = fnext;
        continue;
      }
    }
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

/* Raw decompressed bytes (no framing): what the device-side framing of the text path takes (qb_text_submit). */
long fqr_read_raw(fqr_reader *r, uint8_t *dst, size_t cap) {
  size_t n = 0;
  while (n < cap) {
    if (r->begin >= r->end) {
      const int e = refill(r);
      if (e) {
        if (e == -3) r->status = -3;
        else if (r->status == 0) r->status = -1;
        break;
      }
    }
    size_t k = r->end - r->begin;
    if (k > cap - n) k = cap - n;
    memcpy(dst + n, r->buf + r->begin, k);
    r->begin += k;
    n += k;
  }
  return (long)n;
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
