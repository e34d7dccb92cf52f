// qb_kernels.cuh -- internal interface between the C-ABI layer (qb_api.cu) and the sm_100a
// kernels (qb_kernels.cu).  Not installed; the public boundary is include/quack_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace qb {

constexpr int kRow = 97;             // base_information as u64[97], quack.c:134-139
constexpr int kColContent = 91;
constexpr int kColLength = 95;
constexpr int kColKmer = 96;
constexpr uint32_t kNoHit = 0xFFFFFFFFu;

// counters[] slots (u64, per mate)
constexpr int kCntReads = 0;         // number_of_sequences, quack.c:220
constexpr int kCntInvalidQual = 1;   // quality bytes outside [33,123] (UB in the reference)
constexpr int kCntError = 2;         // != 0: a tile/read exceeded capacity (results invalid)
constexpr int kNumCounters = 4;

// ---- adapter set as the kernels see it -------------------------------------------------
// Internal key order: first base of the window LEAST significant (2 bits per base, codes
// A0 T1 C2 G3).  The public API takes the reference order (first base most significant);
// qb_api.cu converts.
constexpr uint32_t kAnchorBases = 7;                       // anchor = 7-mer, one probed per 4 bases
constexpr uint32_t kAnchorSpace = 1u << (2 * kAnchorBases);  // 2^14 anchors
constexpr uint32_t kAnchorWords = kAnchorSpace / 32;       // 512 words: exact bitmap
constexpr uint32_t kAnchorCopies = 8;                      // copies in shared memory (row = 8 words = 32 B)
constexpr uint32_t kAnchorSmemBytes = kAnchorWords * kAnchorCopies * 4;  // 16 KiB

// exact key set in shared memory: a cuckoo table of two halves, so membership is two loads, no loop
constexpr uint32_t kExactSlots = 2048;         // 2 x 1024 slots (8 KiB)
constexpr uint32_t kExactHalf = kExactSlots / 2;
constexpr uint32_t kExactEmpty = 0xFFFFFFFFu;
constexpr uint32_t kExactMul1 = 0x9E3779B1u, kExactMul2 = 0x85EBCA6Bu;
// byte offsets of the two candidate slots of a key
__host__ __device__ inline uint32_t exact_off1(uint32_t key) { return ((key * kExactMul1) >> 20) & 0xFFCu; }
__host__ __device__ inline uint32_t exact_off2(uint32_t key) { return (((key * kExactMul2) >> 20) & 0xFFCu) + kExactHalf * 4u; }

struct AdapterSet {
  const uint32_t *bitmap;  // exact membership, 2^20 bits, device global (stays in L2)
  const uint32_t *anchor;  // [kAnchorWords] exact bitmap of the 7-mer anchors, device global
  const uint32_t *exact;   // [kExactSlots] cuckoo table of the keys, or nullptr if they do not fit
  int enabled;             // 0: no -a (kmer_count handled at finish)
};

// ---- one batch in device memory ---------------------------------------------------------
struct BatchView {
  const uint8_t *seq;      // padded: readable up to round_up(n_bytes,16)+16
  const uint8_t *qual;
  const uint32_t *offset;  // padded to a multiple of 4 entries (+4)
  const uint32_t *length;
  uint32_t n_reads;
  uint64_t n_bytes;
  uint32_t max_len;        // longest read in the batch (upper bound)
  uint2 *tiles;            // scratch of n_reads entries (the flat kernel keeps its chunk index there)
  uint32_t uniform_len;    // != 0: the HOST verified that every read has this length and that the reads lie back to
                           // back (offset[r] = offset[0] + r * uniform_len); 0: unknown / ragged
  uint32_t first_offset;   // offset[0] of a uniform / contiguous batch
  uint32_t contig_min_len; // != 0: the HOST verified that the reads lie back to back (offset[r + 1] = offset[r] +
                           // length[r]); the value is the shortest read's length.  0: unknown / gaps between reads
};

struct Accum {
  unsigned long long *rows;      // [len_cap][97]
  unsigned long long *counters;  // [kNumCounters]
  uint32_t len_cap;
};

// ---- fused kernel geometry (computed on the host, see fused_plan) ------------------------
struct FusedPlan {
  uint32_t half_len;        // Lh: positions [0,Lh) count in the low u16, [Lh,2Lh) in the high u16
  uint32_t tile_bytes;      // capacity of one stage buffer (seq or qual), multiple of 16
  uint32_t reads_per_tile;  // multiple of 4, <= kMaxTileReads
  uint32_t stages;
  uint32_t qbase;           // score field s = q - qbase, s in [0,46] counted in shared memory
  uint32_t smem_bytes;
  uint32_t grid;
  int ok;                   // 0: len_cap does not fit -> use the simple kernel
};

constexpr uint32_t kMaxTileReads = 248;   // reads per tile (8 per consumer warp at 31 warps)
#ifndef QB_CW
#define QB_CW 15
#endif
constexpr int kFusedConsumerWarps = QB_CW;  // + 1 producer warp = 512 threads (up to 128 registers each), 1 CTA per SM

// ---- period kernel geometry (qb_period.cu; computed on the host, see period_plan) --------
constexpr uint32_t kPeriodMaxLen = 256;  // two histogram blocks of 128 positions

struct PeriodPlan {
  uint32_t len;              // the common read length l (32..256; odd lengths up to 159)
  uint32_t k;                // reads per period: k * l is a multiple of 4
  uint32_t wp;               // 32-bit words per period = k * l / 4
  uint32_t steps;            // warp steps per period = ceil(wp / 32): 3, 4 or 5 (template parameter)
  uint32_t ppt;              // periods per tile
  uint32_t tile_bytes;       // ppt * k * l, a multiple of 4
  uint32_t buf_bytes;        // one staged buffer (bases or quality bytes of a tile + alignment slack + padding)
  uint32_t reads_per_tile;   // ppt * k, a multiple of 4
  uint32_t stages;           // staged tiles per warp (2..4)
  uint32_t warps;            // autonomous warps per CTA (one CTA per SM): 24, 20 or 16 (template parameter)
  uint32_t nblocks;          // histogram blocks (128 positions each)
  uint32_t wblock;           // bytes of one warp block (barriers, first hits, queue, stages x (seq + qual))
  uint32_t smem_base;        // shared address the dynamic shared memory must start at (checked by the kernel)
  uint32_t smem_bytes;
  uint32_t afilt_s, exact_s, kmerhist_s, slot_s;  // shared addresses of the CTA-wide arrays
  uint32_t region_s[3], region_n[3];      // warp blocks: region_n[i] blocks from region_s[i]
  // -a: every warp keeps the 2-bit codes of its last `rt` tiles (one byte per 32-bit word of bases, one row per
  // period) so that anchor hits can wait in a queue until 32 of them fill a confirmation pass
  uint32_t rt;               // tiles in the packed-code ring (2..4)
  uint32_t prow_stride;      // bytes from one period row to the next (word i of a period at row + 4 + i)
  uint32_t pring_hole;       // != 0: warp 0's ring, in the unused columns of histogram block 1; 0: inside the warp block
  uint32_t pring_wstride;    // hole rings: bytes from one warp's ring to the next
  uint32_t pring_off;        // in-block rings: offset inside the warp block
  uint32_t hdr_bytes;        // warp block header (barriers, queue, first hits[, ring]); the staged tiles follow
  uint32_t qbase;
  uint32_t grid;
  int ok;                    // 0: not a batch for this kernel
};

PeriodPlan period_plan(uint32_t uniform_len, uint32_t first_offset, int adapters, int sm_count, uint32_t smem_optin,
                       uint32_t smem_reserved, uint32_t qbase);
// counts the first n_main = floor(n_reads / reads_per_tile) * reads_per_tile reads of a uniform batch; the caller
// hands the rest to another kernel.  *n_main_out = 0: nothing launched.
cudaError_t launch_period(const BatchView &b, const Accum &a, const AdapterSet &ad, const PeriodPlan &plan,
                          cudaStream_t stream, uint32_t *n_main_out);
cudaError_t period_configure();

// ---- flat kernel geometry (qb_flat.cu; ragged batches of back-to-back reads of 16..320 bp) --------
struct FlatPlan {
  uint32_t max_len;          // longest read of the batch (histogram positions 0 .. max_len + 2)
  uint32_t stride;           // 32-bit columns per histogram row, a multiple of 32
  uint32_t chunk_bytes;      // byte window that defines a chunk of reads (one warp at a time)
  uint32_t epoch;            // rounds between two flushes of the u16 counters
  uint32_t hist_o, lenhist_o, kmerhist_o, afilt_o, exact_o, wblock_o, wblock;  // offsets in dynamic shared memory
  uint32_t qbase;
  uint32_t smem_base;        // shared address the dynamic shared memory must start at (checked by the kernel)
  uint32_t smem_bytes;
  uint32_t grid;
  uint32_t warps, max_reads; // warps per CTA; reads per chunk the warp blocks are sized for (128 or 256)
  uint32_t max_steps;        // 32-word steps per chunk at most (32; 64 without -a for reads long enough)
  int ok;                    // 0: not a batch for this kernel
};
FlatPlan flat_plan(uint32_t batch_max_len, uint32_t batch_min_len, int adapters, int sm_count, uint32_t smem_optin,
                   uint32_t smem_reserved, uint32_t qbase);
cudaError_t launch_flat(const BatchView &b, const Accum &a, const AdapterSet &ad, const FlatPlan &plan, cudaStream_t stream);
cudaError_t flat_configure();

FusedPlan fused_plan(uint32_t len_cap, uint32_t batch_max_len, int adapters, int sm_count,
                     uint32_t smem_optin, uint32_t qbase);

cudaError_t launch_simple(const BatchView &b, const Accum &a, const AdapterSet &ad, int sm_count,
                          cudaStream_t stream);
cudaError_t launch_fused(const BatchView &b, const Accum &a, const AdapterSet &ad, const FusedPlan &plan,
                         cudaStream_t stream);
cudaError_t fused_configure();  // opt in to the large dynamic shared memory once per device

// ---- on-device FASTQ framing (qb_text.cu) ------------------------------------------------
struct TextState {
  uint32_t tail_len, broken;
};
struct TextSummary {
  uint32_t n_reads, n_bytes, min_len, max_len, valid, tail_len, n_lines, pad;
};
cudaError_t launch_text_frame(const uint8_t *chunk, uint32_t n_chunk, const uint8_t *carry_in, uint8_t *carry_out,
                              uint32_t carry_cap, TextState *state, uint32_t *scratch, uint32_t nl_cap, uint32_t rec_cap,
                              uint8_t *seq, uint8_t *qual, uint32_t *offset, uint32_t *length, uint32_t out_cap,
                              TextSummary *sum_dev, cudaStream_t stream);
size_t text_scratch_words(uint32_t chunk_cap, uint32_t carry_cap, uint32_t nl_cap, uint32_t rec_cap);

// ---- device-side transform() (qb_transform.cu; quack.c:230-293) ----
uint32_t transformed_rows(uint32_t max_length);
cudaError_t launch_transform(const unsigned long long *rows, uint32_t max_length, unsigned long long n_reads, bool noad_quirk,
                             unsigned long long *out, unsigned long long *scratch1, cudaStream_t stream);

// ---- opt-in side outputs (qb_extras.cu): N count per position, per-read mean quality distribution ----
constexpr uint32_t kExtrasMeanBins = 94;
cudaError_t launch_extras(const BatchView &b, uint32_t len_cap, unsigned long long *n_count, unsigned long long *mean_hist,
                          int sm_count, cudaStream_t stream);

// ---- on-device inflate of BGZF blocks (qb_inflate.cu) --------------------------------------
struct BgzfBlock {
  uint32_t in_off, in_len;    // the raw DEFLATE stream of the block inside the chunk's compressed bytes
  uint32_t out_off, out_len;  // where its text goes (prefix sums of ISIZE) and ISIZE
  uint32_t crc;               // CRC-32 of the text, from the block trailer
};
cudaError_t launch_inflate_bgzf(const uint8_t *comp, const BgzfBlock *blocks, uint32_t n_blocks, uint8_t *text, uint32_t *bad,
                                uint32_t *status, cudaStream_t stream);
cudaError_t launch_inflate_merge(const uint32_t *bad, TextState *state, cudaStream_t stream);

// L2 flush helper for timing (writes `bytes` of scratch)
cudaError_t launch_l2_flush(uint32_t *scratch, size_t words, cudaStream_t stream);

// Microbenchmarks of the shared-memory pipes the fused kernel leans on (tools/microbench).
cudaError_t run_microbench(int sm_count, char *report, size_t report_cap);

}  // namespace qb
