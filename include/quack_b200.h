/*
 * quack_b200.h -- C-ABI of the B200-native replacement for quack's per-read statistics
 * accumulation (the body of read_fastq(), reference quack.c:180-228).
 *
 * The reference has no plugin/FFI layer: the seam is the plain C call
 *     sequence_data *read_fastq(char *fastq_file, int *kmers);          (quack.c:180)
 * made once per mate from main() (quack.c:911, 917), fed by kseq_read() (klib/kseq.h:177)
 * and by the table read_adapters() builds (quack.c:154-178), and consumed by transform()
 * and draw() (quack.c:230, 295).  This header is what a host program binds instead of that
 * loop: the host keeps the record reader and the renderer, packs records into batches
 *     seq[]    concatenated base bytes   (kseq_t.seq.s,  l bytes per read, no padding)
 *     qual[]   concatenated quality bytes(kseq_t.qual.s, same offsets as seq[])
 *     offset[] u32 start of each read in seq[]/qual[]   (ascending)
 *     length[] u32 read length                          (kseq_read() return value)
 * and gets back the accumulator in the reference's own memory layout:
 *     uint64_t rows[max_length][97]  ==  base_information bases[max_length]  (quack.c:134-139)
 *       [0..90] scores[q-33]   [91..94] content[A,T,C,G]   [95] length_count   [96] kmer_count
 * so it can be handed to transform()/draw() (or their restatement) unchanged.
 *
 * Everything is plain pointers and sizes; no CUDA or torch types cross this boundary.
 * All entry points return 0 on success or a negative qb_status; qb_last_error() gives the
 * message.  There is no CPU fallback: without a usable CUDA device qb_create() fails.
 */
#ifndef QUACK_B200_H
#define QUACK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB_ROW_U64 97        /* sizeof(base_information)/8, quack.c:134-139 */
#define QB_N_SCORES 91       /* scores[91] */
#define QB_COL_CONTENT 91    /* content[4]: A,T,C,G (lookup[] order, quack.c:148-150) */
#define QB_COL_LENGTH 95     /* length_count */
#define QB_COL_KMER 96       /* kmer_count */
#define QB_KMER_SIZE 10      /* quack.c:155, 184 */
#define QB_KEY_SPACE (1u << 20)

typedef enum {
  QB_OK = 0,
  QB_ERR_CUDA = -1,        /* a CUDA runtime call failed (message has the CUDA error string) */
  QB_ERR_NCCL = -2,        /* an NCCL call failed */
  QB_ERR_ARG = -3,         /* bad argument */
  QB_ERR_CAPACITY = -4,    /* batch exceeds the ring slot, or a read exceeds len_cap */
  QB_ERR_NOMEM = -5,
  QB_ERR_LAYOUT = -6,      /* offsets not ascending / reads overlap */
  QB_ERR_TEXT = -7         /* text path: the FASTQ text is not canonical 4-line records (or a record is longer than the
                              carry buffer): nothing of that mate's text is reliable, use the host reader instead */
} qb_status;

typedef enum {
  QB_KERNEL_AUTO = 0,      /* period kernel for batches of back-to-back reads of one length in [32, 256]; the flat
                              kernel for other batches of back-to-back reads of 16..320 bp; else the fused kernel up
                              to 320 bp, the warp-tile kernel, or simple */
  QB_KERNEL_SIMPLE = 1,    /* one warp per read, global atomics: any len_cap, slow */
  QB_KERNEL_FUSED = 2,     /* CTA-wide TMA-staged tiles, joint (base,score) shared-memory histogram (v3) */
  QB_KERNEL_WTILE = 3,     /* (v4 warp-tile kernel, removed in round 2: slower than v3 wherever both ran; the value is
                              still accepted and takes the fused kernel) */
  QB_KERNEL_PERIOD = 4,    /* v5: lanes own fixed positions of a k-read period, one aligned load + PRMT + RED per base;
                              uniform-length batches only (an error otherwise), the reads that do not fill a tile
                              go to the AUTO choice among the others */
  QB_KERNEL_FLAT = 5       /* v6: ragged batches of back-to-back reads of 16..320 bp, lane <-> 16-byte unit, read
                              boundaries from a bit set per chunk (an error for other batches) */
} qb_kernel;

typedef struct qb_ctx qb_ctx;       /* one per process; owns devices, streams, accumulators */
typedef struct qb_dbatch qb_dbatch; /* a batch resident in device memory */

typedef struct {
  int n_devices;              /* GPUs driven by this process (>= 1) */
  const int *device_ids;      /* NULL -> 0..n_devices-1 */
  uint32_t len_cap;           /* longest read accepted; rows allocated per mate */
  int n_mates;                /* 1 (-u) or 2 (-1/-2): independent accumulators */
  int adapters_enabled;       /* 0: no -a (reference then counts kmer_count[10] per read with
                                 l > 10, quack.c:210-217); 1: use adapter_keys */
  const uint32_t *adapter_keys; /* distinct 20-bit keys, first base most significant: exactly the
                                   indices read_adapters() sets in kmers[] (quack.c:165-172) */
  uint32_t n_adapter_keys;
  uint64_t batch_bytes;       /* capacity of one ring slot for seq[] (and for qual[]); 0 -> 64 MiB */
  uint32_t batch_reads;       /* capacity of one ring slot in reads; 0 -> batch_bytes/32 */
  int ring_depth;             /* slots per device (>= 2 to overlap H2D with compute); 0 -> 3 */
  int kernel;                 /* qb_kernel */
} qb_config;

/* One pinned host slot of the ring (filled by the record reader, then submitted). */
typedef struct {
  uint8_t *seq;
  uint8_t *qual;
  uint32_t *offset;
  uint32_t *length;
  uint64_t cap_bytes;
  uint32_t cap_reads;
  int device_index;           /* which of the ctx's devices this slot feeds */
  int slot;
} qb_batch;

/* ---- lifetime ---- */
int qb_create(const qb_config *cfg, qb_ctx **out);
void qb_destroy(qb_ctx *ctx);
const char *qb_last_error(const qb_ctx *ctx); /* ctx may be NULL: error of a failed qb_create */
int qb_device_count(void);                    /* CUDA devices visible, <0 on error */

/* ---- streaming path (replaces the while loop of read_fastq, quack.c:193-221) ---- */
/* Blocks until a slot of the ring is free (its previous kernel finished) and hands its pinned
 * host buffers to the caller.  Slots rotate over the devices round-robin.  Thread-safe. */
int qb_acquire(qb_ctx *ctx, qb_batch *out);
/* Queues H2D copies of the first n_bytes / n_reads of the slot and the statistics kernel on the
 * slot's stream, and returns without waiting.  max_len = longest read in the batch (0 = unknown). */
int qb_submit(qb_ctx *ctx, const qb_batch *b, int mate, uint32_t n_reads, uint64_t n_bytes,
              uint32_t max_len);
/* Same, from caller-owned host memory (pinned memory copies asynchronously): takes the next free
 * device slot itself.  Must respect the slot capacities. */
int qb_submit_from(qb_ctx *ctx, int mate, const uint8_t *seq, const uint8_t *qual,
                   const uint32_t *offset, const uint32_t *length, uint32_t n_reads,
                   uint64_t n_bytes, uint32_t max_len);
/* Convenience for callers with one big host array: splits at read boundaries into slot-sized
 * batches (rebasing offsets) and submits them all.  Offsets must be ascending. */
int qb_accumulate_host(qb_ctx *ctx, int mate, const uint8_t *seq, const uint8_t *qual,
                       const uint32_t *offset, const uint32_t *length, uint64_t n_reads);
/* Waits for all queued work of every device. */
int qb_sync(qb_ctx *ctx);

/* ---- text path: the DEVICE frames the records (replaces kseq_read(), klib/kseq.h:177-218, for canonical FASTQ) ----
 * The host hands over decompressed FASTQ text in chunks cut ANYWHERE (no framing on the host, one host-to-device copy
 * per chunk); the device finds the line ends, checks that every four lines are a canonical record ('@' line, one
 * sequence line, '+' line, one quality line of the same length, no '\r'), packs bases and quality bytes and runs
 * the statistics kernels on them.  A record that straddles two chunks is completed with the next chunk (the chunks
 * of a mate must be submitted in stream order, from one thread).  Text that is not canonical -- multi-line records,
 * blank lines, FASTA, garbage in front of the first header, CR LF -- makes qb_finish() of that mate fail with
 * QB_ERR_TEXT: the caller then counts the stream through qb_acquire/qb_submit and a reader with the full kseq
 * semantics (host/fq_reader.c).  One device per context. */
typedef struct {
  uint8_t *text;              /* pinned host buffer for the chunk */
  uint64_t cap_bytes;
  int device_index;
  int slot;
} qb_text;
int qb_text_acquire(qb_ctx *ctx, qb_text *out);
/* last != 0: the stream ends with this chunk (the caller appends a '\n' if the file does not end with one). */
int qb_text_submit(qb_ctx *ctx, const qb_text *t, int mate, uint64_t n_bytes, int last);
/* After qb_sync()/qb_finish(): records framed so far, and the bytes that were left behind the last complete record
 * when the stream ended (!= 0: truncated record -- kseq_read() returns -2 there and quack stops, quack.c:193). */
int qb_text_status(qb_ctx *ctx, int mate, uint64_t *n_reads, uint64_t *tail_bytes);

/* ---- compressed text path: the DEVICE inflates BGZF blocks (replaces the inflate inside gzread(), quack.c:160,187 ->
 * klib/kseq.h:74,105, for block-gzipped input; block layout klib/bgzf.c:63-71, trailer 261-266) ----
 * Same flow as qb_text_submit(), but the buffer from qb_text_acquire() is filled with WHOLE BGZF blocks as they are
 * in the file: one warp per block inflates them on the device (stored, fixed and dynamic DEFLATE blocks; CRC-32 and
 * ISIZE of every block are checked there), the text goes straight into the framing kernels, and the host-to-device
 * copy carries the compressed bytes only.  The blocks of a chunk must hold at most cap_bytes of text: qb_bgzf_fit()
 * tells how many bytes of a buffer are whole blocks that fit.  A block that does not inflate, a wrong CRC-32 or bytes
 * that are not BGZF blocks make the mate fail with QB_ERR_TEXT like text that is not canonical: the caller then reads
 * the file through the host reader, which reports damaged input the way the reference does. */
int qb_bgzf_submit(qb_ctx *ctx, const qb_text *t, int mate, uint64_t n_bytes, int last);
/* Both submits from caller-owned host memory (pinned memory copies asynchronously and must stay valid until the next
 * qb_sync()/qb_finish()); they take the next free slot themselves.  Same capacity rules. */
int qb_bgzf_submit_from(qb_ctx *ctx, int mate, const uint8_t *blocks, uint64_t n_bytes, int last);
int qb_text_submit_from(qb_ctx *ctx, int mate, const uint8_t *text, uint64_t n_bytes, int last);
/* Bytes of text one chunk may hold (qb_text.cap_bytes of every slot). */
int qb_text_capacity(qb_ctx *ctx, uint64_t *cap_bytes);
/* n_whole = bytes of buf[0..n_bytes) that are whole BGZF blocks holding at most text_cap_bytes of text together
 * (n_text).  QB_ERR_TEXT: a block header is not BGZF.  Pure host code. */
int qb_bgzf_fit(const uint8_t *buf, uint64_t n_bytes, uint64_t text_cap_bytes, uint64_t *n_whole, uint64_t *n_text);
/* The decoder on its own: inflates whole BGZF blocks (host memory) on device 0 and copies the text back. */
int qb_bgzf_inflate(qb_ctx *ctx, const uint8_t *comp, uint64_t n_bytes, uint8_t *text_out, uint64_t text_cap_bytes,
                    uint64_t *n_text_out);
/* Diagnostics: the inflate kernel alone over all blocks of a buffer (< 3.7 GB of text), device-resident, CUDA-event
 * time per launch. */
int qb_bgzf_inflate_bench(qb_ctx *ctx, const uint8_t *comp, uint64_t n_bytes, int iters, float *ms_per_launch,
                          uint64_t *n_text_out, uint32_t *n_blocks_out);

/* ---- result (what read_fastq() returns, quack.c:222-227) ---- */
/* Synchronises, sums the per-device accumulators (one NCCL reduce to the first device / rank 0
 * when more than one GPU took part), and copies rows[0..max_length) to rows_out in the
 * base_information layout.  max_length = longest read, n_reads = number_of_sequences.
 * With qb_comm_init_rank() only rank 0 receives the summed result; other ranks get their own
 * partial rows.  rows_cap is in rows (positions). */
int qb_finish(qb_ctx *ctx, int mate, uint64_t *rows_out, uint64_t rows_cap, uint64_t *max_length,
              uint64_t *n_reads);
/* qb_finish() followed by the reference's transform() (quack.c:230-293) ON THE DEVICE: reads longer than 3000 bp binned
 * by 100 positions (with the reference's in-place quirks), kmer_count as a running sum, scores as integer
 * percentages, length / kmer counts as ceil(100 * (float) x / nseq).  rows_out receives *max_length transformed rows
 * (the binned length when original_max_length > 3000) -- exactly what draw() (quack.c:295) takes; only those rows
 * cross the link (10 k instead of 1 M for million-base reads).  rows_out == NULL queries the sizes. */
int qb_finish_transformed(qb_ctx *ctx, int mate, uint64_t *rows_out, uint64_t rows_cap, uint64_t *max_length,
                          uint64_t *n_reads, uint64_t *original_max_length);
/* Zeroes the accumulators of one mate (all devices) so a context can be reused. */
int qb_reset(qb_ctx *ctx, int mate);
/* ---- opt-in side outputs that the reference does NOT compute (no reference oracle; never part of the SVG) ----
 * qb_extras_enable(): before the first batch; every batch then takes a second, small kernel pass.
 * qb_extras_finish(): n_count[p] = reads whose base p is 'N' or 'n' (the main result keeps folding N into A like
 * lookup[] does, quack.c:150); qual_sum[p] = sum of (q - 33) over the reads at position p (derived from the heatmap
 * rows); mean_hist[m], m = 0..93 = reads whose mean quality floor(sum(q - 33) / l) is m (q clamped to [33, 126]).
 * n_count and qual_sum hold rows_cap entries (>= max_length); any of the three pointers may be NULL.  With
 * qb_comm_init_rank() rank 0 receives the sums. */
int qb_extras_enable(qb_ctx *ctx);
int qb_extras_finish(qb_ctx *ctx, int mate, uint64_t *n_count, uint64_t *qual_sum, uint64_t rows_cap, uint64_t *mean_hist);
/* Diagnostics: quality bytes outside [33,123] seen (undefined behaviour in the reference). */
int qb_invalid_quality_count(qb_ctx *ctx, int mate, uint64_t *out);

/* ---- multi-process multi-GPU (one process per GPU, e.g. under torchrun) ---- */
#define QB_NCCL_ID_BYTES 128
int qb_nccl_unique_id(uint8_t id_out[QB_NCCL_ID_BYTES]);             /* call on rank 0, broadcast */
/* Every rank's context must have been created with the same len_cap, n_mates and adapters_enabled. */
int qb_comm_init_rank(qb_ctx *ctx, int n_ranks, int rank, const uint8_t id[QB_NCCL_ID_BYTES]);

/* ---- device-resident batches (config 5: kernel-only sweep; bench.py `value`) ---- */
int qb_dbatch_upload(qb_ctx *ctx, int device_index, const uint8_t *seq, const uint8_t *qual,
                     const uint32_t *offset, const uint32_t *length, uint32_t n_reads,
                     uint64_t n_bytes, uint32_t max_len, qb_dbatch **out);
/* Allocates an empty device batch and fills it on the device's own host-side generator output
 * (see qb_gen_reads) without keeping a host copy. */
int qb_dbatch_generate(qb_ctx *ctx, int device_index, uint64_t seed, int mate, uint64_t first_read,
                       uint32_t n_reads, uint32_t len_min, uint32_t len_max, double adapter_rate,
                       qb_dbatch **out);
int qb_dbatch_run(qb_ctx *ctx, qb_dbatch *b, int mate);              /* async launch */
/* iters launches timed with CUDA events on the launching stream; average ms per launch.
 * flush_l2 != 0 writes a buffer larger than L2 between launches (outside the timed events). */
int qb_dbatch_time(qb_ctx *ctx, qb_dbatch *b, int mate, int warmup, int iters, int flush_l2,
                   float *ms_avg, float *ms_min);
int qb_dbatch_info(const qb_dbatch *b, uint32_t *n_reads, uint64_t *n_bytes);
void qb_dbatch_free(qb_ctx *ctx, qb_dbatch *b);
/* Number of kernel launches issued by this context so far (bench.py gpu_launches). */
uint64_t qb_launch_count(const qb_ctx *ctx);
/* How many of those launches took the simple kernel / one of the shared-memory kernels (warp-tile or
 * fused; chosen per batch whenever the batch's longest read fits the shared-memory histogram,
 * whatever len_cap is). */
int qb_kernel_counts(const qb_ctx *ctx, uint64_t *n_simple, uint64_t *n_fused);
/* How many launches took the period kernel (they are also counted in n_fused above). */
uint64_t qb_period_launch_count(const qb_ctx *ctx);
/* How many launches took the flat kernel (also counted in n_fused above). */
uint64_t qb_flat_launch_count(const qb_ctx *ctx);
/* Bytes queued for host-to-device copy by qb_submit() / qb_submit_from() / qb_accumulate_host() so far: bases +
 * quality bytes of every batch, plus offsets / lengths of the reads that do not take the period kernel (it needs
 * none: the host verified the batch shape).  bench.py reports e2e.h2d_bytes_per_step from this counter. */
uint64_t qb_h2d_bytes(const qb_ctx *ctx);
/* Live kernel timing: after qb_profile_enable(ctx, n) every statistics-kernel launch is bracketed by
 * a CUDA event pair on the stream it is launched on (up to n launches, then recording stops).
 * qb_profile_collect() synchronises and returns the per-launch durations in ms and the algorithmic
 * bytes (2*bases + 8*reads) of each launch; it returns the number of launches
 * written (<= cap) and resets the recorder.  qb_profile_enable(ctx, 0) switches it off. */
int qb_profile_enable(qb_ctx *ctx, int max_launches);
int qb_profile_collect(qb_ctx *ctx, float *ms_out, uint64_t *bytes_out, int cap);
/* CUDA-event stopwatch on the main stream of one device (the stream qb_dbatch_run launches on). */
int qb_timer_start(qb_ctx *ctx, int device_index);
int qb_timer_stop(qb_ctx *ctx, int device_index, float *ms);
/* Measured pinned H2D bandwidth of device_index in GB/s (best of `iters` copies of `bytes`). */
int qb_measure_h2d(qb_ctx *ctx, int device_index, uint64_t bytes, int iters, double *gbs);
/* The same for ONE pass over n pinned host buffers (each byte leaves host memory once: the streaming ceiling). */
int qb_measure_h2d_list(qb_ctx *ctx, int device_index, const void *const *ptrs, const uint64_t *sizes, uint32_t n, double *gbs);
/* Geometry and shared-memory counter layout of the period kernel (QB_KERNEL_PERIOD) for reads of one length.
 * Needs no GPU.  0 if the kernel takes such batches, -1 otherwise.  info = reads per period, words per period,
 * warp steps per period, periods per tile, reads per tile, stages, warps per CTA; slot[p] = histogram block << 7 | 32-bit
 * column of position p (its bank is the column modulo 32). */
int qb_period_layout(uint32_t read_len, int adapters, uint32_t info[7], uint8_t slot[256]);
/* Tools: the rest of that plan.  out = reads per period, words per period, steps, periods per tile, reads per tile,
 * stages, warps, histogram blocks, bytes per warp block, header bytes of it, tiles in the -a packed-code ring, ring
 * row stride, shared address of the ring when it lives in the unused columns of histogram block 1 (else 0), ring
 * bytes per warp there, ring offset inside the warp block otherwise, bytes of one staged buffer. */
int qb_period_plan_info(uint32_t read_len, int adapters, uint32_t out[16]);

/* ---- host helpers that define the kernel's inputs ---- */
/* lookup[(c-65)&~32] of quack.c:150,201 extended to all byte values. */
int qb_base_code(int c);
/* Keys of one adapter record exactly as read_adapters() inserts them (quack.c:165-172): windows
 * ending at i = 10..l-1.  Appends to keys[] (capacity cap), returns the number appended or
 * QB_ERR_CAPACITY.  Duplicates are allowed in adapter_keys. */
int qb_adapter_record_keys(const char *seq, size_t l, uint32_t *keys, size_t cap);
/* Deterministic synthetic reads: read i of a file depends only on
 * (seed, mate, i), so shards can be generated independently.  Fills n_reads reads starting at
 * index first_read; lengths uniform in [len_min,len_max]; returns bytes written via *n_bytes.
 * Buffers: seq/qual >= n_reads*len_max bytes; offset/length >= n_reads. */
int qb_gen_reads(uint64_t seed, int mate, uint64_t first_read, uint32_t n_reads, uint32_t len_min,
                 uint32_t len_max, double adapter_rate, uint8_t *seq, uint8_t *qual,
                 uint32_t *offset, uint32_t *length, uint64_t *n_bytes);
/* Pinned host memory for callers that stage their own batches (qb_submit_from). */
void *qb_host_alloc(size_t bytes);
void qb_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* QUACK_B200_H */
