/* quack_main.c -- the `quack` command line on top of the B200 statistics library.
 *
 * Same options, same usage texts, same exit codes and the same SVG on stdout as the reference main()
 * (quack.c:54-132, 858-928).  What changed is the middle: instead of calling read_fastq() once per
 * mate, one host thread per mate runs the batching reader (fq_reader.c), packs records into the pinned
 * slots handed out by qb_acquire() and submits them (qb_submit), so both mates are decoded
 * concurrently while copies and kernels overlap on the GPU(s); qb_finish() then returns the
 * accumulator in base_information layout for the unchanged transform/draw stage (render.c).
 *
 * Extra knobs live in the environment so that the five reference options stay untouched:
 *   QB_DEVICES=N        number of GPUs to spread batches over (default 1)
 *   QB_BATCH_MB=M       pinned slot size in MiB for seq[] and for qual[] (default 16)
 *   QB_LEN_CAP=L        longest read accepted (default and maximum 1048576; the accumulators start at 512 rows
 *                       and grow with the longest read seen, so the limit costs nothing until it is needed)
 *   QB_KERNEL=0|1|2|3   auto | simple | fused | wtile
 *   QB_CLEAN_EXIT=1     free everything before exit (default: _exit after the SVG is flushed)
 *   QB_DEVICE_FRAMING=1 the DEVICE frames the records (qb_text_submit): the reader threads only inflate.  Canonical
 *                       4-line FASTQ only; on anything else the run starts over with the host reader.  Regular files.
 *   QB_DEVICE_INFLATE=1 BGZF files: the device also inflates (qb_bgzf_submit): the host only reads the file.  Falls back
 *                       the same way (not BGZF -> device framing; damaged or odd input -> host reader).  Unset: on
 *                       for BGZF files of >= QB_DEVICE_INFLATE_MIN_MB (1024) MiB each; 0: never.
 *   QB_DEVICE_TRANSFORM=1 transform() (binning, percentages; quack.c:230-293) runs on the device
 *   QB_EXTRAS_JSON=path side outputs the reference does not have (N count and quality sum per position, per-read mean
 *                       quality distribution; qb_extras_*): written there, never part of the SVG
 *   QB_STATS_JSON=path  write reads/s, bases/s and stage times there (stdout stays the SVG)
 *   QUACK_DECODE_THREADS=n  inflate threads per BGZF input file (default: half of the cores, at most 8)
 */
#include <pthread.h>
#include <unistd.h>
#include <stdio.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/quack_b200.h"
#include "fq_reader.h"
#include "render.h"

static const char *program_version = "quack 1.1.1";

static const char *usage_head =
    "Usage: quack [OPTION...]\n"
    "quack -- A FASTQ quality assessment tool\n\n"
    "  -1, --forward file.1.fq.gz      Forward strand\n"
    "  -2, --reverse file.2.fq.gz      Reverse strand\n";

struct options {
  const char *name, *forward, *reverse, *unpaired, *adapters;
};

/* quack.c:59-132, behaviour preserved case by case ("Config / flags") */
static struct options parse_options(int argc, char **argv) {
  struct options o = {NULL, NULL, NULL, NULL, NULL};
  if (argc == 1 || argc == 2) {
    if (argc == 1 || !strcmp(argv[1], "--help") || !strcmp(argv[1], "--usage") || !strcmp(argv[1], "-?"))
      printf("%s"
             "  -a, --adapters adapters.fa.gz   (Optional) Adapters file\n"
             "  -n, --name NAME                 (Optional) Display in output\n"
             "  -u, --unpaired unpaired.fq.gz   Data (only use with -u)\n"
             "  -?, --help                      Give this help list\n"
             "      --usage                     (use alone)\n"
             "  -V, --version                   Print program version (use alone)\n"
             "Report bugs to <thrash@igbb.msstate.edu>.\n",
             usage_head);
    if (argc == 2 && (!strcmp(argv[1], "-V") || !strcmp(argv[1], "--version"))) {
      printf("%s\n", program_version);
      exit(0);
    }
  }
  if (argc > 2 && argc % 2 != 0) {
    for (int i = 1; i < argc; i += 2) {
      const char *f = argv[i];
      if (!strcmp(f, "--forward") || !strcmp(f, "-1"))
        o.forward = argv[i + 1];
      else if (!strcmp(f, "--reverse") || !strcmp(f, "-2"))
        o.reverse = argv[i + 1];
      else if (!strcmp(f, "--adapters") || !strcmp(f, "-a"))
        o.adapters = argv[i + 1];
      else if (!strcmp(f, "--unpaired") || !strcmp(f, "-u"))
        o.unpaired = argv[i + 1];
      else if (!strcmp(f, "--name") || !strcmp(f, "-n"))
        o.name = argv[i + 1];
      else {
        fprintf(stderr,
                "%s"
                "  -a, --adapters adapters.fa.gz    Adapters file\n"
                "  -n, --name NAME            Display in output\n"
                "  -u, --unpaired unpaired.fq.gz        Data (only use with -u)\n"
                "  -?, --help                 Give this help list\n"
                "      --usage                (use alone)\n"
                "  -V, --version              Print program version (use alone)\n"
                "Report bugs to <thrash@igbb.msstate.edu>.\n",
                usage_head);
        exit(EXIT_FAILURE);
      }
    }
  }
  return o;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

struct mate_job {
  qb_ctx *ctx;
  int mate;
  const char *path;
  int rc;                /* 0 ok, else a qb_status / reader failure */
  char err[256];
  uint64_t reads, bases, text_bytes;
  double inflate_s, wall_s;
  int stream_status;
  int decode_threads;
  uint32_t len_cap;
  fqr_reader *reader; /* opened before the CUDA context exists: its inflate pool works through the start-up */
  int device_framing; /* 1: the device frames the text; 2: it also inflates the BGZF blocks */
  uint64_t text_bytes_sent;
};

/* BGZF file, block bytes straight to the device: no inflate, no framing on the host */
static void stream_bgzf_blocks(struct mate_job *j) {
  const int fd = open(j->path, O_RDONLY);
  struct stat sb;
  if (fd < 0 || fstat(fd, &sb) != 0) {
    j->rc = QB_ERR_TEXT;
    if (fd >= 0) close(fd);
    return;
  }
  off_t off = 0;
  double t_acq = 0, t_read = 0, t_sub = 0, t0 = now_s(), t1;
  for (;;) {
    qb_text t;
    if ((j->rc = qb_text_acquire(j->ctx, &t))) break;
    t1 = now_s(), t_acq += t1 - t0, t0 = t1;
    ssize_t got = 0;
    while ((uint64_t)got < t.cap_bytes) { /* (pread returns at most 2 GiB - 4 KiB per call) */
      const ssize_t k = pread(fd, t.text + got, (size_t)(t.cap_bytes - (uint64_t)got), off + got);
      if (k <= 0) break;
      got += k;
    }
    uint64_t whole = 0, text = 0;
    const int frc = qb_bgzf_fit(t.text, (uint64_t)got, t.cap_bytes, &whole, &text);
    const int at_end = off + (off_t)whole >= sb.st_size;
    if (frc || (whole == 0 && !at_end)) { /* not BGZF, or a block cut short: the host reader decides what that means */
      qb_bgzf_submit(j->ctx, &t, j->mate, 0, 1);
      j->rc = QB_ERR_TEXT;
      break;
    }
    j->text_bytes += text;
    j->text_bytes_sent += whole;
    t1 = now_s(), t_read += t1 - t0, t0 = t1;
    if ((j->rc = qb_bgzf_submit(j->ctx, &t, j->mate, whole, at_end))) break;
    t1 = now_s(), t_sub += t1 - t0, t0 = t1;
    off += (off_t)whole;
    if (at_end) break;
  }
  close(fd);
  if (getenv("QB_VERBOSE") && atoi(getenv("QB_VERBOSE")) > 1)
    fprintf(stderr, "quack: mate %d: waiting for a slot %.3f s, reading the file %.3f s, submitting %.3f s\n", j->mate, t_acq, t_read, t_sub);
  j->stream_status = -1;
}

/* reader thread of one mate: inflate + frame + pack + submit, until the stream ends */
static void *mate_thread(void *arg) {
  struct mate_job *j = (struct mate_job *)arg;
  const double t0 = now_s();
  fqr_reader *r = j->reader;
  int more = 1;
  if (j->device_framing == 2) {
    stream_bgzf_blocks(j);
    if (j->rc > -100 && j->rc != 0 && !j->err[0]) snprintf(j->err, sizeof j->err, "%s", qb_last_error(j->ctx));
    j->wall_s = now_s() - t0;
    return NULL;
  }
  if (j->device_framing) { /* raw text to the device, which frames it */
    int last_byte = '\n';
    while (more) {
      qb_text t;
      if ((j->rc = qb_text_acquire(j->ctx, &t))) break;
      long n = fqr_read_raw(r, t.text, (size_t)t.cap_bytes - 1);
      more = fqr_status(r) == 0;
      if (n > 0) last_byte = t.text[n - 1];
      if (!more && last_byte != '\n') t.text[n++] = '\n'; /* kseq takes the end of the stream for a line end */
      j->text_bytes_sent += (uint64_t)n;
      if ((j->rc = qb_text_submit(j->ctx, &t, j->mate, (uint64_t)n, !more))) break;
    }
    more = 0;
  }
  while (more) {
    qb_batch b;
    if ((j->rc = qb_acquire(j->ctx, &b))) break;
    uint32_t n = 0, max_len = 0;
    uint64_t nb = 0;
    more = fqr_fill(r, b.seq, b.qual, b.offset, b.length, b.cap_bytes, b.cap_reads, &n, &nb, &max_len);
    if (more < 0) {
      j->rc = QB_ERR_CAPACITY;
      snprintf(j->err, sizeof j->err, "%s: a record is longer than a batch slot (raise QB_BATCH_MB)", j->path);
      qb_submit(j->ctx, &b, j->mate, 0, 0, 0);
      break;
    }
    if (max_len > j->len_cap) {
      j->rc = QB_ERR_CAPACITY;
      snprintf(j->err, sizeof j->err, "%s: a read of %u bp is longer than QB_LEN_CAP=%u (the limit is 1048576)", j->path,
               max_len, j->len_cap);
      qb_submit(j->ctx, &b, j->mate, 0, 0, 0);
      break;
    }
    if ((j->rc = qb_submit(j->ctx, &b, j->mate, n, nb, max_len))) break;
    j->reads += n;
    j->bases += nb;
  }
  if (j->rc > -100 && j->rc != 0 && !j->err[0]) snprintf(j->err, sizeof j->err, "%s", qb_last_error(j->ctx));
  j->stream_status = fqr_status(r);
  j->inflate_s = fqr_inflate_seconds(r);
  j->text_bytes = fqr_bytes_in(r);
  j->decode_threads = fqr_decode_threads(r);
  fqr_close(r);
  j->wall_s = now_s() - t0;
  return NULL;
}

static long env_long(const char *name, long dflt) {
  const char *v = getenv(name);
  return v && *v ? atol(v) : dflt;
}

int main(int argc, char **argv) {
  const struct options o = parse_options(argc, argv);
  const int paired = o.forward != NULL && o.reverse != NULL;
  const int unpaired = o.unpaired != NULL;
  const int adapters = o.adapters != NULL;
  if (paired == unpaired) { /* quack.c:872-875 */
    printf("%s\n", "Usage: quack [OPTION...]\nTry `quack --help' or `quack --usage' for more information.");
    exit(1);
  }
  const double t_start = now_s();
  /* one GPU is used unless QB_DEVICES says otherwise: on a multi-GPU node the driver then initialises only that one
   * (cuInit time grows with the number of visible devices) */
  if (env_long("QB_DEVICES", 1) == 1) setenv("CUDA_VISIBLE_DEVICES", "0", 0);

  uint32_t *keys = NULL;
  long n_keys = 0;
  if (adapters) { /* read_adapters(), quack.c:877 */
    n_keys = fqr_read_adapter_keys(o.adapters, &keys);
    if (n_keys < 0) {
      fprintf(stderr, "quack: cannot open adapters file %s\n", o.adapters);
      return 2;
    }
  }

  /* The inputs are opened first: a compressed file's inflate pool (fq_reader.c) starts to decode at once and
   * runs ahead of the framing code by up to a few hundred MiB, so the ~0.5 s the CUDA context, the pinned ring and
   * the kernel images take to come up are spent inflating instead of waiting. */
  struct mate_job jobs[2];
  const int n_mates = paired ? 2 : 1;
  qb_config cfg;
  qb_ctx *ctx = NULL;
  double t_created = 0;
  /* QB_DEVICE_FRAMING=1: first pass with the device framing the text; anything it does not take (not canonical
   * 4-line FASTQ, a partial record at the end) starts the run over with the host reader, so only for regular files */
  /* QB_DEVICE_INFLATE: 1 = BGZF inputs are inflated and framed on the device, 0 = never; unset = only when every input
   * is a BGZF file of at least QB_DEVICE_INFLATE_MIN_MB (1024) MiB: there the device path streams 2.4x faster than the
   * 16-core host pool, while for small inputs the host pool hides behind the CUDA start-up anyway. */
  const char *die = getenv("QB_DEVICE_INFLATE");
  const int inflate_auto = !(die && *die);
  const int framing_only = env_long("QB_DEVICE_FRAMING", 0) != 0;
  int device_framing = (inflate_auto || atol(die) != 0) ? 2 : framing_only;
  if (inflate_auto && env_long("QB_DEVICES", 1) > 1) device_framing = framing_only; /* (the text path drives one GPU) */
  const long long min_bytes = inflate_auto ? (long long)env_long("QB_DEVICE_INFLATE_MIN_MB", 1024) << 20 : 0;
  for (int m = 0; m < n_mates && device_framing; m++) {
    struct stat sb;
    const char *p = paired ? (m == 0 ? o.forward : o.reverse) : o.unpaired;
    if (stat(p, &sb) != 0 || !S_ISREG(sb.st_mode)) device_framing = 0;
    if (device_framing == 2) { /* BGZF? (klib/bgzf.c:63-71: gzip member with a 6-byte 'BC' extra field) */
      unsigned char h[16] = {0};
      FILE *f = fopen(p, "rb");
      const size_t n = f ? fread(h, 1, sizeof h, f) : 0;
      if (f) fclose(f);
      if (n < 16 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || h[3] != 4 || h[12] != 'B' || h[13] != 'C' ||
          (long long)sb.st_size < min_bytes)
        device_framing = inflate_auto ? framing_only : 1;
    }
  }
  if (device_framing == 2 && env_long("QB_VERBOSE", 0)) fprintf(stderr, "quack: BGZF input: inflate and framing on the device\n");
  for (;;) {
    memset(jobs, 0, sizeof jobs);
    for (int m = 0; m < n_mates; m++) {
      jobs[m].path = paired ? (m == 0 ? o.forward : o.reverse) : o.unpaired;
      jobs[m].reader = device_framing == 2 ? NULL : fqr_open(jobs[m].path);
      if (!jobs[m].reader && device_framing != 2) {
        fprintf(stderr, "quack: cannot open %s\n", jobs[m].path);
        return 2;
      }
    }

    memset(&cfg, 0, sizeof cfg);
    cfg.n_devices = device_framing ? 1 : (int)env_long("QB_DEVICES", 1);
    cfg.len_cap = (uint32_t)env_long("QB_LEN_CAP", 1 << 20);
    cfg.n_mates = paired ? 2 : 1;
    cfg.adapters_enabled = adapters;
    cfg.adapter_keys = keys;
    cfg.n_adapter_keys = (uint32_t)n_keys;
    cfg.batch_bytes = (uint64_t)env_long("QB_BATCH_MB", 16) << 20; /* the program is decode-bound: small pinned ring, short start-up */
    /* device inflate: the reader threads only copy file bytes, so more chunks in flight per mate keep the GPU busy */
    cfg.ring_depth = (int)env_long("QB_RING", device_framing == 2 ? 6 : 3);
    cfg.kernel = (int)env_long("QB_KERNEL", QB_KERNEL_AUTO);
    ctx = NULL;
    if (qb_create(&cfg, &ctx)) {
      fprintf(stderr, "quack: %s\n", qb_last_error(NULL));
      return 2;
    }
    if (getenv("QB_EXTRAS_JSON") && *getenv("QB_EXTRAS_JSON") && qb_extras_enable(ctx)) {
      fprintf(stderr, "quack: %s\n", qb_last_error(ctx));
      return 2;
    }
    t_created = now_s(); /* CUDA start-up, pinned ring, accumulators: a fixed cost per process */

    pthread_t th[2];
    for (int m = 0; m < cfg.n_mates; m++) {
      jobs[m].ctx = ctx;
      jobs[m].mate = m;
      jobs[m].len_cap = cfg.len_cap;
      jobs[m].device_framing = device_framing;
      pthread_create(&th[m], NULL, mate_thread, &jobs[m]);
    }
    for (int m = 0; m < cfg.n_mates; m++) pthread_join(th[m], NULL);
    if (device_framing) {
      int again = 0;
      for (int m = 0; m < cfg.n_mates; m++) {
        uint64_t tail = 0;
        if (jobs[m].rc == QB_ERR_TEXT) again = 1;
        if (!jobs[m].rc) {
          const int trc = qb_text_status(ctx, m, &jobs[m].reads, &tail);
          if (trc == QB_ERR_TEXT || trc == QB_ERR_ARG || (trc == QB_OK && tail != 0) || jobs[m].stream_status != -1) again = 1;
          else if (trc) jobs[m].rc = trc, snprintf(jobs[m].err, sizeof jobs[m].err, "%s", qb_last_error(ctx));
        }
      }
      if (again) {
        if (env_long("QB_VERBOSE", 0)) fprintf(stderr, "quack: device framing declined the input; host reader takes it\n");
        qb_destroy(ctx);
        device_framing = 0;
        continue;
      }
    }
    break;
  }
  for (int m = 0; m < cfg.n_mates; m++)
    if (jobs[m].rc) {
      fprintf(stderr, "quack: %s\n", jobs[m].err[0] ? jobs[m].err : "statistics path failed");
      qb_destroy(ctx);
      return 2;
    }
  for (int m = 0; m < cfg.n_mates; m++) {
    /* -1 is the clean end of the stream.  Anything else ended it early: like the reference (quack.c:193: any
     * negative kseq_read() leaves the loop) the records in front of the damage are reported, but not silently */
    const int st = jobs[m].stream_status;
    if (st != -1 && st != 0)
      fprintf(stderr, "quack: warning: %s: %s; the report covers the %llu records in front of it\n", jobs[m].path,
              st == -2   ? "truncated record (quality string shorter than the sequence)"
              : st == -3 ? "the compressed stream is damaged or truncated"
              : st == -5 ? "a record without quality line (FASTA) in a FASTQ input"
                         : "the stream ended with an error",
              (unsigned long long)jobs[m].reads);
  }
  const double t_stream = now_s();

  qr_data data[2];
  memset(data, 0, sizeof data);
  for (int m = 0; m < cfg.n_mates; m++) {
    /* longest read first (the reduce runs once, for both mates), then exactly that many rows */
    int frc = qb_finish(ctx, m, NULL, 0, &data[m].max_length, &data[m].n_reads);
    if (!frc) {
      data[m].rows = (uint64_t *)malloc(sizeof(uint64_t) * QB_ROW_U64 * (size_t)(data[m].max_length ? data[m].max_length : 1));
      frc = qb_finish(ctx, m, data[m].rows, data[m].max_length, &data[m].max_length, &data[m].n_reads);
    }
    if (frc) {
      fprintf(stderr, "quack: %s\n", qb_last_error(ctx));
      qb_destroy(ctx);
      return 2;
    }
    if (data[m].n_reads == 0 || data[m].max_length == 0) {
      /* the reference dereferences a NULL accumulator here (quack.c:450) */
      fprintf(stderr, "quack: no reads in %s\n", jobs[m].path);
      qb_destroy(ctx);
      return 2;
    }
  }
  const char *xjs = getenv("QB_EXTRAS_JSON");
  if (xjs && *xjs) { /* side outputs, straight to their own file */
    FILE *f = fopen(xjs, "w");
    if (f) {
      fprintf(f, "{\"note\": \"not computed by the reference: no reference oracle\", \"mates\": [");
      for (int m = 0; m < cfg.n_mates; m++) {
        const uint64_t ml = data[m].max_length;
        uint64_t *nc = (uint64_t *)calloc(2 * ml + 94, sizeof(uint64_t)), *qs = nc + ml, *mh = qs + ml;
        if (qb_extras_finish(ctx, m, nc, qs, ml, mh)) {
          fprintf(stderr, "quack: %s\n", qb_last_error(ctx));
          return 2;
        }
        fprintf(f, "%s{\"n_count\": [", m ? ", " : "");
        for (uint64_t p = 0; p < ml; p++) fprintf(f, "%s%llu", p ? ", " : "", (unsigned long long)nc[p]);
        fprintf(f, "], \"qual_sum\": [");
        for (uint64_t p = 0; p < ml; p++) fprintf(f, "%s%llu", p ? ", " : "", (unsigned long long)qs[p]);
        fprintf(f, "], \"mean_quality_hist\": [");
        for (int b = 0; b < 94; b++) fprintf(f, "%s%llu", b ? ", " : "", (unsigned long long)mh[b]);
        fprintf(f, "]}");
        free(nc);
      }
      fprintf(f, "]}\n");
      fclose(f);
    }
  }
  uint64_t framed_reads = 0, framed_bases = 0; /* (before qr_transform() turns the counts into fractions) */
  for (int m = 0; m < cfg.n_mates; m++) {
    framed_reads += data[m].n_reads;
    for (uint64_t p = 0; p < data[m].max_length; p++) /* length_count sits in the row of the last base */
      framed_bases += (p + 1) * data[m].rows[p * QB_ROW_U64 + QB_COL_LENGTH];
  }
  const double t_finish = now_s();

  qr_begin_document(paired, adapters, o.name, stdout);
  const int device_transform = env_long("QB_DEVICE_TRANSFORM", 0) != 0 && !(xjs && *xjs);
  for (int m = 0; m < cfg.n_mates; m++) {
    if (device_transform) { /* transform() on the device (qb_finish_transformed): only the binned rows cross the link */
      uint64_t ml = 0, nr = 0, orig = 0;
      if (qb_finish_transformed(ctx, m, data[m].rows, data[m].max_length, &ml, &nr, &orig)) {
        fprintf(stderr, "quack: %s\n", qb_last_error(ctx));
        return 2;
      }
      if (orig > 3000) fprintf(stderr, "Binning...\n");
      data[m].max_length = ml, data[m].original_max_length = orig;
    } else {
      qr_transform(&data[m], stderr);
    }
    qr_draw(&data[m], m, adapters, stdout);
  }
  qr_end_document(o.name, stdout);
  fflush(stdout);
  const double t_end = now_s();

  const char *js = getenv("QB_STATS_JSON");
  if (js && *js) {
    FILE *f = fopen(js, "w");
    if (f) {
      uint64_t reads = 0, bases = 0, text = 0;
      double inflate = 0;
      for (int m = 0; m < cfg.n_mates; m++) {
        reads += jobs[m].reads, bases += jobs[m].bases, text += jobs[m].text_bytes;
        if (jobs[m].inflate_s > inflate) inflate = jobs[m].inflate_s;
      }
      if (device_framing) reads = framed_reads, bases = framed_bases; /* the device framed the records */
      const double stream_s = t_stream - t_start;
      fprintf(f,
              "{\"reads\": %llu, \"bases\": %llu, \"text_bytes\": %llu, \"devices\": %d, \"launches\": %llu, "
              "\"stream_s\": %.6f, \"finish_s\": %.6f, \"render_s\": %.6f, \"total_s\": %.6f, "
              "\"host_gzip_decode_s_max_over_mates\": %.6f, \"host_gzip_decode_MBps\": %.2f, "
              "\"reads_per_s\": %.1f, \"bases_per_s\": %.1f, \"decode_threads\": %d, \"create_s\": %.6f}\n",
              (unsigned long long)reads, (unsigned long long)bases, (unsigned long long)text, cfg.n_devices,
              (unsigned long long)qb_launch_count(ctx), stream_s, t_finish - t_stream, t_end - t_finish, t_end - t_start,
              inflate, inflate > 0 ? (double)text / cfg.n_mates / inflate / 1e6 : 0.0, (double)reads / stream_s,
              (double)bases / stream_s, jobs[0].decode_threads, t_created - t_start);
      fclose(f);
    }
  }
  if (env_long("QB_CLEAN_EXIT", 0)) { /* tests under memory checkers; otherwise the process is over: like the reference,
                                        which frees nothing (quack.c:914-927), leave the teardown to the OS */
    for (int m = 0; m < cfg.n_mates; m++) free(data[m].rows);
    free(keys);
    qb_destroy(ctx);
    return 0;
  }
  fflush(stdout);
  fflush(stderr);
  _exit(0);
}
