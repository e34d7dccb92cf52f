// qb_kernels.cu -- sm_100a kernels for quack's per-read statistics accumulation
// (reference: the while loop of read_fastq(), quack.c:193-221).
//
// Two kernels compute the same thing:
//   simple_kernel : one warp per read, 64-bit global atomics.  Any read length.  The fallback
//                   for len_cap beyond what the shared-memory histogram holds (still CUDA).
//   fused_kernel  : persistent, one CTA per SM.  A producer warp streams tiles of whole reads
//                   (seq bytes, qual bytes, offsets, lengths) into shared memory with 1-D TMA bulk
//                   copies (cp.async.bulk + mbarrier, 3 stages).  Consumer warps then make two
//                   passes over the staged tile:
//                     phase A (flat, one aligned 4-byte word per lane): SWAR base -> 2-bit code,
//                       score byte -> 6-bit bin, fused into one key byte K = code<<6 | bin per
//                       base, written to a K buffer; with -a, the 4 adapter 10-mer windows ending
//                       in the lane's word are tested against a bank-replicated blocked Bloom
//                       filter in shared memory (one conflict-free LDS per window) and confirmed
//                       against the exact 2^20-bit set in L2 only on a filter hit.
//                     phase H (one warp per read, lane <-> position): one shared-memory atomic
//                       per base into a JOINT (code,score) x position histogram of packed u16
//                       counters whose bank is the position mod 32, so a warp's 32 updates never
//                       conflict.  content[] and scores[] are its marginals, taken at flush time.
//                   Counters are flushed to the u64 global accumulator with one RED per non-zero
//                   bin, at the end of the launch or every 65535 reads per CTA.
//
// No tensor cores: the path is an integer histogram.
#include "qb_kernels.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace qb {

// ------------------------------------------------------------------------------------------
// shared helpers
// ------------------------------------------------------------------------------------------

// lookup[(c-65)&~32] of quack.c:150,201 on [A-Ta-t]; the rest of the byte range is defined by
// the same bit tests (oracle: qo_base_code).  A=0 T=1 C=2 G=3, N and everything else 0.
__host__ __device__ __forceinline__ uint32_t base_code(uint32_t b) {
  uint32_t cg = ((b & 0x5Bu) == 0x43u) ? 1u : 0u;
  uint32_t lo = (((b & 0x1Fu) == 0x07u) || ((b & 0x1Fu) == 0x14u)) ? 1u : 0u;
  return 2u * cg + lo;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// simple kernel
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) simple_kernel(BatchView b, Accum a, AdapterSet ad) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned long long n_reads = 0, n_invalid = 0;

  for (uint32_t r = warp; r < b.n_reads; r += nwarps) {
    const uint32_t off = b.offset[r], len = b.length[r];
    if (len > a.len_cap) {
      if (lane == 0) atomicAdd(&a.counters[kCntError], 1ull);
      continue;
    }
    const uint8_t *s = b.seq + off;
    const uint8_t *q = b.qual + off;
    for (uint32_t i = lane; i < len; i += 32) {  // quack.c:199-205
      atomicAdd(&a.rows[(size_t)i * kRow + kColContent + base_code(s[i])], 1ull);
      const int sc = (int)q[i] - 33;
      if (sc >= 0 && sc < 91)
        atomicAdd(&a.rows[(size_t)i * kRow + sc], 1ull);
      else
        n_invalid++;
    }
    if (ad.enabled && len > 10) {  // quack.c:206-217: first window (end p >= 9) in the set
      uint32_t hit_pos = kNoHit;
      for (uint32_t base = 9; base < len && hit_pos == kNoHit; base += 32) {
        const uint32_t p = base + lane;
        bool hit = false;
        if (p < len) {
          uint32_t key = 0;
#pragma unroll
          for (int j = 0; j < 10; j++) key |= base_code(s[p - 9 + j]) << (2 * j);
          hit = (ad.bitmap[key >> 5] >> (key & 31u)) & 1u;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) hit_pos = base + __ffs(m) - 1;
      }
      if (lane == 0 && hit_pos != kNoHit && hit_pos + 1 < len)
        atomicAdd(&a.rows[(size_t)(hit_pos + 1) * kRow + kColKmer], 1ull);
    }
    if (lane == 0) {  // quack.c:219-220
      if (len) atomicAdd(&a.rows[(size_t)(len - 1) * kRow + kColLength], 1ull);
      n_reads++;
    }
  }
  n_invalid = warp_sum(n_invalid);
  if (lane == 0) {
    if (n_reads) atomicAdd(&a.counters[kCntReads], n_reads);
    if (n_invalid) atomicAdd(&a.counters[kCntInvalidQual], n_invalid);
  }
}

cudaError_t launch_simple(const BatchView &b, const Accum &a, const AdapterSet &ad, int sm_count,
                          cudaStream_t stream) {
  if (b.n_reads == 0) return cudaSuccess;
  simple_kernel<<<sm_count * 8, 256, 0, stream>>>(b, a, ad);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// fused kernel
// ------------------------------------------------------------------------------------------

constexpr int kCW = kFusedConsumerWarps;       // consumer warps
constexpr int kCThreads = kCW * 32;
constexpr int kThreads = kCThreads + 32;       // + producer warp
constexpr int kMaxStages = 3;

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "QB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra QB_DONE;\n"
      "bra QB_WAIT;\n"
      "QB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void consumer_bar() {
  asm volatile("bar.sync 1, %0;" ::"n"(kCThreads) : "memory");
}

struct FusedArgs {
  BatchView b;
  Accum a;
  AdapterSet ad;
  FusedPlan plan;
  uint32_t n_tiles;
};

// shared-memory carve-up, all offsets multiples of 16 bytes
struct SmemLayout {
  uint32_t hist, bloom, lenhist, kmerhist, kbuf, stage0, stage_stride, seq_off, qual_off, soff_off,
      slen_off, fhit_off, meta, bars, total;
};

__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

__host__ __device__ inline SmemLayout smem_layout(uint32_t half_len, uint32_t len_cap, uint32_t tile_bytes,
                                                  uint32_t stages, int adapters) {
  SmemLayout L;
  uint32_t o = 0;
  L.hist = o;
  o += 256u * half_len * 4u;
  L.bloom = o;
  o += adapters ? kBloomBytes : 0u;
  L.lenhist = o;
  o += align16(len_cap * 4u);
  L.kmerhist = o;
  o += align16(len_cap * 4u);
  L.kbuf = o;
  o += tile_bytes + 16u;
  L.stage0 = o;
  L.seq_off = 0;
  L.qual_off = tile_bytes + 16u;
  L.soff_off = 2u * (tile_bytes + 16u);
  L.slen_off = L.soff_off + kMaxTileReads * 4u;
  L.fhit_off = L.slen_off + kMaxTileReads * 4u;
  L.stage_stride = L.fhit_off + kMaxTileReads * 4u;
  o += stages * L.stage_stride;
  L.meta = o;
  o += 16u * kMaxStages;  // lo_al, n_reads, span, pad per stage
  L.bars = o;
  o += 16u * kMaxStages;  // full, empty per stage
  L.total = o;
  return L;
}

// phase A arithmetic for one aligned word: returns the 4 key bytes K (code<<6 | s') and, through
// cc, the 2-bit codes alone in bits 7:6 of each byte.
//   s' = q - qbase in [1,63] is counted in shared memory; any other quality byte in the word sends
//   the whole word to the exact slow path of phase H (K = code<<6 | 0, the dummy rows).
__device__ __forceinline__ uint32_t key_bytes(uint32_t sw, uint32_t qw, uint32_t qsub, uint32_t &cc) {
  // per byte: bit7 of n_cg is 0 iff (b & 0x5B) == 0x43; bit6 of n_g / n_t is 0 iff (b & 0x1F) == 7 / 0x14
  const uint32_t n_cg = ((sw & 0x5B5B5B5Bu) ^ 0x43434343u) + 0x7F7F7F7Fu;
  const uint32_t n_g = ((sw & 0x1F1F1F1Fu) ^ 0x07070707u) + 0x3F3F3F3Fu;
  const uint32_t n_t = ((sw & 0x1F1F1F1Fu) ^ 0x14141414u) + 0x3F3F3F3Fu;
  const uint32_t nc = (n_cg & 0x80808080u) | (n_g & n_t & 0x40404040u);
  cc = nc ^ 0xC0C0C0C0u;
  const uint32_t qs = qw - qsub;  // a borrow can only start at a byte that is itself out of range
  return (qs & 0xC0C0C0C0u) ? cc : (qs | cc);
}

template <bool kAdapters>
__global__ void __launch_bounds__(kThreads, 1) fused_kernel(const FusedArgs args) {
  extern __shared__ __align__(128) uint8_t smem[];
  const FusedPlan &plan = args.plan;
  const uint32_t Lh = plan.half_len;
  const uint32_t len_cap = args.a.len_cap;
  const uint32_t S = plan.stages;
  const SmemLayout L = smem_layout(Lh, len_cap, plan.tile_bytes, S, kAdapters);

  uint32_t *hist = reinterpret_cast<uint32_t *>(smem + L.hist);
  const uint32_t *bloom_s = reinterpret_cast<const uint32_t *>(smem + L.bloom);
  uint32_t *lenhist = reinterpret_cast<uint32_t *>(smem + L.lenhist);
  uint32_t *kmerhist = reinterpret_cast<uint32_t *>(smem + L.kmerhist);
  uint8_t *kbuf = smem + L.kbuf;
  uint32_t *meta = reinterpret_cast<uint32_t *>(smem + L.meta);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);  // [2*s] full, [2*s+1] empty

  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  const uint32_t warp = tid >> 5;

  // ---- prologue: zero histograms, load the Bloom filter, init barriers ----
  for (uint32_t i = tid; i < 256u * Lh; i += kThreads) hist[i] = 0;
  for (uint32_t i = tid; i < len_cap; i += kThreads) {
    lenhist[i] = 0;
    kmerhist[i] = 0;
  }
  if (kAdapters) {
    uint32_t *bw = reinterpret_cast<uint32_t *>(smem + L.bloom);
    for (uint32_t i = tid; i < kBloomWords * 32u; i += kThreads) bw[i] = args.ad.bloom[i];
  }
  for (uint32_t s = 0; s < S; s++) {
    uint32_t *fh = reinterpret_cast<uint32_t *>(smem + L.stage0 + s * L.stage_stride + L.fhit_off);
    for (uint32_t i = tid; i < kMaxTileReads; i += kThreads) fh[i] = kNoHit;
  }
  if (tid == 0) {
    for (uint32_t s = 0; s < S; s++) {
      mbar_init(&bars[2 * s], 1);
      mbar_init(&bars[2 * s + 1], kCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint32_t RT = plan.reads_per_tile;
  const uint32_t n_reads = args.b.n_reads;

  if (warp == kCW) {
    // =============================== producer warp ===============================
    if (lane == 0) {
      unsigned long long reads_seen = 0;
      uint32_t it = 0;
      for (uint32_t tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % S, use = it / S;
        if (use > 0) mbar_wait(&bars[2 * s + 1], (use & 1u) ^ 1u);
        const uint32_t r0 = tile * RT;
        const uint32_t r1 = min(r0 + RT, n_reads);
        uint32_t nr = r1 - r0;
        const uint32_t lo = args.b.offset[r0];
        const uint32_t hi = args.b.offset[r1 - 1] + args.b.length[r1 - 1];
        const uint32_t lo_al = lo & ~15u;
        uint32_t span = hi > lo_al ? ((hi - lo_al + 15u) & ~15u) : 0u;
        if (span > plan.tile_bytes || hi < lo) {  // capacity / layout violation: never corrupt silently
          atomicAdd(&args.a.counters[kCntError], 1ull);
          span = 0;
          nr = 0;
        }
        uint8_t *st = smem + L.stage0 + s * L.stage_stride;
        meta[4 * s + 0] = lo_al;
        meta[4 * s + 1] = nr;
        meta[4 * s + 2] = span;
        const uint32_t idx_bytes = ((nr + 3u) & ~3u) * 4u;
        mbar_arrive_expect_tx(&bars[2 * s], 2u * span + 2u * idx_bytes);
        if (span) {
          bulk_g2s(st + L.seq_off, args.b.seq + lo_al, span, &bars[2 * s]);
          bulk_g2s(st + L.qual_off, args.b.qual + lo_al, span, &bars[2 * s]);
        }
        if (idx_bytes) {
          bulk_g2s(st + L.soff_off, args.b.offset + r0, idx_bytes, &bars[2 * s]);
          bulk_g2s(st + L.slen_off, args.b.length + r0, idx_bytes, &bars[2 * s]);
        }
        reads_seen += nr;
      }
      if (reads_seen) atomicAdd(&args.a.counters[kCntReads], reads_seen);
    }
    return;
  }

  // ================================= consumer warps =================================
  const uint32_t ctid = tid;  // 0 .. kCThreads-1
  const uint32_t qsub = plan.qbase * 0x01010101u;
  unsigned long long n_invalid = 0;
  uint32_t reads_since_flush = 0;

  auto flush = [&]() {
    // all consumer warps have passed the barrier that ends phase H
    const uint32_t npos = min(2u * Lh, len_cap);
    for (uint32_t pos = ctid; pos < npos; pos += kCThreads) {
      const bool hi = pos >= Lh;
      const uint32_t col = hi ? pos - Lh : pos;
      const uint32_t sh = hi ? 16u : 0u;
      unsigned long long *row = args.a.rows + (size_t)pos * kRow;
      uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      for (uint32_t sp = 1; sp < 64; sp++) {
        const uint32_t v0 = (hist[(sp)*Lh + col] >> sh) & 0xFFFFu;
        const uint32_t v1 = (hist[(64u + sp) * Lh + col] >> sh) & 0xFFFFu;
        const uint32_t v2 = (hist[(128u + sp) * Lh + col] >> sh) & 0xFFFFu;
        const uint32_t v3 = (hist[(192u + sp) * Lh + col] >> sh) & 0xFFFFu;
        const uint32_t tot = v0 + v1 + v2 + v3;
        c0 += v0, c1 += v1, c2 += v2, c3 += v3;
        if (tot) {
          const int sc = (int)(sp + plan.qbase) - 33;
          if (sc >= 0 && sc < 91)
            atomicAdd(&row[sc], (unsigned long long)tot);
          else
            n_invalid += tot;
        }
      }
      if (c0) atomicAdd(&row[kColContent + 0], (unsigned long long)c0);
      if (c1) atomicAdd(&row[kColContent + 1], (unsigned long long)c1);
      if (c2) atomicAdd(&row[kColContent + 2], (unsigned long long)c2);
      if (c3) atomicAdd(&row[kColContent + 3], (unsigned long long)c3);
      if (pos < len_cap) {
        const uint32_t lc = lenhist[pos], kc = kmerhist[pos];
        if (lc) atomicAdd(&row[kColLength], (unsigned long long)lc);
        if (kc) atomicAdd(&row[kColKmer], (unsigned long long)kc);
        lenhist[pos] = 0;
        kmerhist[pos] = 0;
      }
    }
    consumer_bar();
    uint4 *h4 = reinterpret_cast<uint4 *>(hist);
    for (uint32_t i = ctid; i < 64u * Lh; i += kCThreads) h4[i] = make_uint4(0, 0, 0, 0);
    consumer_bar();
  };

  uint32_t it = 0;
  for (uint32_t tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x, ++it) {
    const uint32_t s = it % S, use = it / S;
    mbar_wait(&bars[2 * s], use & 1u);
    const uint32_t lo_al = meta[4 * s + 0];
    const uint32_t nr = meta[4 * s + 1];
    const uint32_t nwords = meta[4 * s + 2] >> 2;
    uint8_t *st = smem + L.stage0 + s * L.stage_stride;
    const uint32_t *seqw = reinterpret_cast<const uint32_t *>(st + L.seq_off);
    const uint32_t *qualw = reinterpret_cast<const uint32_t *>(st + L.qual_off);
    const uint8_t *qualb = st + L.qual_off;
    const uint32_t *soff = reinterpret_cast<const uint32_t *>(st + L.soff_off);
    const uint32_t *slen = reinterpret_cast<const uint32_t *>(st + L.slen_off);
    uint32_t *fhit = reinterpret_cast<uint32_t *>(st + L.fhit_off);
    uint32_t *kw = reinterpret_cast<uint32_t *>(kbuf);

    // ------------------------------ phase A: flat over the tile's words ------------------------------
    if (!kAdapters) {
      for (uint32_t w = warp * 32u + lane; w < nwords; w += kCThreads) {
        uint32_t cc;
        kw[w] = key_bytes(seqw[w], qualw[w], qsub, cc);
      }
    } else {
      // each warp step covers 29 new words; lanes 0..2 re-read the 3 words before them so that
      // every 10-mer window ending in lanes 3..31 finds its 9 earlier bases inside the warp
      const uint32_t nsteps = (nwords + 28u) / 29u;
      const uint32_t M = args.ad.bloom_mul;
      const uint32_t lane4 = lane * 4u;
      for (uint32_t step = warp; step < nsteps; step += kCW) {
        const int w = (int)(step * 29u + lane) - 3;
        const bool inr = (w >= 0) && ((uint32_t)w < nwords);
        const uint32_t sw = inr ? seqw[w] : 0u;
        const uint32_t qw = inr ? qualw[w] : 0u;
        uint32_t cc;
        const uint32_t K = key_bytes(sw, qw, qsub, cc);
        const bool own = inr && lane >= 3u;
        if (own) kw[w] = K;
        // 4 bases -> 8 bits, first base least significant
        const uint32_t p8 = (cc * 0x41041u) >> 24;
        const uint32_t a16 = __shfl_up_sync(0xffffffffu, p8, 1) | (p8 << 8);
        const uint32_t R = __shfl_up_sync(0xffffffffu, a16, 2) | (a16 << 16);  // 16 bases, mine on top
        uint32_t acc = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint32_t wj = R >> (6 + 2 * j);  // low 20 bits: window ending at my byte j
          const uint32_t p = wj * M;             // low 20 bits depend on the window only
          const uint32_t word =
              *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(bloom_s) + ((p & 0x7F80u) | lane4));
          acc |= __funnelshift_r(word, 0u, wj) & __funnelshift_r(word, 0u, p >> 15);
        }
        const bool maybe = own && (acc & 1u);
        if (__any_sync(0xffffffffu, maybe)) {
          if (maybe) {
#pragma unroll 1
            for (int j = 0; j < 4; j++) {
              const uint32_t key = (R >> (6 + 2 * j)) & 0xFFFFFu;
              if ((args.ad.bitmap[key >> 5] >> (key & 31u)) & 1u) {
                const uint32_t end_abs = lo_al + 4u * (uint32_t)w + (uint32_t)j;  // byte where the window ends
                // read containing that byte: largest r with soff[r] <= end_abs
                int a = 0, b = (int)nr - 1, r = -1;
                while (a <= b) {
                  const int m = (a + b) >> 1;
                  if (soff[m] <= end_abs) {
                    r = m;
                    a = m + 1;
                  } else
                    b = m - 1;
                }
                if (r >= 0) {
                  const uint32_t p = end_abs - soff[r];
                  if (p >= 9u && p < slen[r]) atomicMin(&fhit[r], p);
                }
              }
            }
          }
        }
      }
    }
    consumer_bar();

    // ------------------------------ phase H: one warp per read ------------------------------
    for (uint32_t r = warp; r < nr; r += kCW) {
      const uint32_t len = slen[r];
      const uint32_t off = soff[r] - lo_al;
      if (len > len_cap) {
        if (lane == 0) atomicAdd(&args.a.counters[kCntError], 1ull);
        continue;
      }
      const uint8_t *kb = kbuf + off;
      for (uint32_t pos0 = 0; pos0 < len; pos0 += 32) {
        const uint32_t pos = pos0 + lane;
        if (pos < len) {
          const uint32_t k = kb[pos];
          const bool hi = pos0 >= Lh;  // warp-uniform: Lh is a multiple of 32
          const uint32_t col = hi ? pos - Lh : pos;
          atomicAdd(&hist[k * Lh + col], hi ? 0x10000u : 1u);
          if ((k & 63u) == 0u) {
            // quality outside the shared-memory window: count this base exactly in global memory
            unsigned long long *row = args.a.rows + (size_t)pos * kRow;
            atomicAdd(&row[kColContent + (k >> 6)], 1ull);
            const int sc = (int)qualb[off + pos] - 33;
            if (sc >= 0 && sc < 91)
              atomicAdd(&row[sc], 1ull);
            else
              n_invalid++;
          }
        }
      }
      if (lane == 0) {
        if (len) atomicAdd(&lenhist[len - 1], 1u);
        if (kAdapters) {
          const uint32_t fh = fhit[r];
          if (fh != kNoHit) {
            if (fh + 1u < len) atomicAdd(&kmerhist[fh + 1u], 1u);
            fhit[r] = kNoHit;
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[2 * s + 1]);  // stage buffers free for the producer
    consumer_bar();                                  // K buffer free for the next phase A

    reads_since_flush += nr;
    if (reads_since_flush + RT > 65535u) {  // u16 counters: flush before any bin can wrap
      flush();
      reads_since_flush = 0;
    }
  }
  flush();
  n_invalid = warp_sum(n_invalid);
  if (lane == 0 && n_invalid) atomicAdd(&args.a.counters[kCntInvalidQual], n_invalid);
}

FusedPlan fused_plan(uint32_t len_cap, uint32_t batch_max_len, int adapters, int sm_count,
                     uint32_t smem_optin, uint32_t qbase) {
  FusedPlan p;
  memset(&p, 0, sizeof p);
  if (len_cap == 0) return p;
  uint32_t half = ((len_cap + 1u) / 2u + 31u) & ~31u;
  if (half < 32u) half = 32u;
  p.half_len = half;
  p.qbase = qbase;
  if (batch_max_len == 0 || batch_max_len > len_cap) batch_max_len = len_cap;
  for (uint32_t stages = kMaxStages; stages >= 2; stages--) {
    const SmemLayout L0 = smem_layout(half, len_cap, 0, stages, adapters);
    if (L0.total + 128u >= smem_optin) continue;
    uint32_t avail = smem_optin - L0.total - 128u;  // L0 already holds every +16 pad
    uint32_t tile = (avail / (2u * stages + 1u)) & ~15u;
    if (tile > 32768u) tile = 32768u;
    if (tile < 64u) continue;
    uint32_t rt = ((tile - 32u) / batch_max_len) & ~3u;
    if (rt > kMaxTileReads) rt = kMaxTileReads;
    if (rt < 4u) continue;
    if (stages == kMaxStages && rt < 32u) continue;  // prefer fewer, larger stages for long reads
    p.stages = stages;
    p.tile_bytes = tile;
    p.reads_per_tile = rt;
    p.smem_bytes = smem_layout(half, len_cap, tile, stages, adapters).total;
    p.grid = (uint32_t)sm_count;
    p.ok = 1;
    return p;
  }
  return p;
}

cudaError_t fused_configure() {
  cudaError_t e = cudaFuncSetAttribute(fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
}

cudaError_t launch_fused(const BatchView &b, const Accum &a, const AdapterSet &ad, const FusedPlan &plan,
                         cudaStream_t stream) {
  if (b.n_reads == 0) return cudaSuccess;
  FusedArgs args;
  args.b = b;
  args.a = a;
  args.ad = ad;
  args.plan = plan;
  args.n_tiles = (b.n_reads + plan.reads_per_tile - 1u) / plan.reads_per_tile;
  uint32_t grid = args.n_tiles < plan.grid ? args.n_tiles : plan.grid;
  if (const char *g = getenv("QB_FUSED_GRID")) {  // test hook: few CTAs exercise the u16 flush path
    const uint32_t v = (uint32_t)atoi(g);
    if (v >= 1 && v < grid) grid = v;
  }
  if (ad.enabled)
    fused_kernel<true><<<grid, kThreads, plan.smem_bytes, stream>>>(args);
  else
    fused_kernel<false><<<grid, kThreads, plan.smem_bytes, stream>>>(args);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// L2 flush (timing hygiene: write a buffer larger than the 126 MB L2 between timed launches)
// ------------------------------------------------------------------------------------------

__global__ void l2_flush_kernel(uint32_t *p, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = (uint32_t)i;
}

cudaError_t launch_l2_flush(uint32_t *scratch, size_t words, cudaStream_t stream) {
  l2_flush_kernel<<<148 * 8, 256, 0, stream>>>(scratch, words);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// microbenchmarks: cycles per warp instruction per SM for the shared-memory operations the fused
// kernel is built from (B200_PROFILING.md: measure, don't guess)
// ------------------------------------------------------------------------------------------

enum { MB_ATOMS_FREE = 0, MB_ATOMS_22LANES, MB_ATOMS_2WAY, MB_ATOMS_SAMEADDR, MB_LDS_U8, MB_LDS_32, MB_LDS_128,
       MB_SHFL, MB_IMAD, MB_ATOMS_LDS_MIX, MB_COUNT };

template <int MODE>
__global__ void __launch_bounds__(1024, 1) microbench_kernel(unsigned long long *cycles, uint32_t *sink, int iters) {
  extern __shared__ __align__(16) uint32_t sm[];
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = i;
  __syncthreads();
  uint32_t x = threadIdx.x * 2654435761u, acc = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x = x * 1664525u + 1013904223u;
      const uint32_t row = (x >> 20) & 255u;  // 256 rows x 32 words = 32 KiB window
      if (MODE == MB_ATOMS_FREE) atomicAdd(&sm[row * 32u + lane], 1u);
      if (MODE == MB_ATOMS_22LANES) { if (lane < 22u) atomicAdd(&sm[row * 32u + lane], 1u); }
      if (MODE == MB_ATOMS_2WAY) atomicAdd(&sm[row * 32u + (lane & 15u) + ((lane >> 4) * 32u * 256u)], 1u);
      if (MODE == MB_ATOMS_SAMEADDR) atomicAdd(&sm[(row & ~1u) * 32u + (lane >> 1)], 1u);
      if (MODE == MB_LDS_U8) acc += reinterpret_cast<const uint8_t *>(sm)[(row * 32u + lane) & 0xFFFFu];
      if (MODE == MB_LDS_32) acc += sm[row * 32u + lane];
      if (MODE == MB_LDS_128) { const uint4 v = reinterpret_cast<const uint4 *>(sm)[(row * 32u + lane) & 4095u]; acc += v.x ^ v.y ^ v.z ^ v.w; }
      if (MODE == MB_SHFL) acc += __shfl_up_sync(0xffffffffu, x, 1);
      if (MODE == MB_IMAD) acc = acc * x + row;
      if (MODE == MB_ATOMS_LDS_MIX) { const uint32_t k = reinterpret_cast<const uint8_t *>(sm)[(x >> 12) & 0x3FFFu]; atomicAdd(&sm[4096u + (k & 255u) * 32u + lane], 1u); }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc + x;
}

template <int MODE>
static double microbench_one(int sm_count, unsigned long long *d_cycles, uint32_t *d_sink, int threads) {
  const int iters = 2000;
  cudaFuncSetAttribute(microbench_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  microbench_kernel<MODE><<<sm_count, threads, 131072>>>(d_cycles, d_sink, 200);
  microbench_kernel<MODE><<<sm_count, threads, 131072>>>(d_cycles, d_sink, iters);
  unsigned long long h[256];
  if (cudaMemcpy(h, d_cycles, sizeof(unsigned long long) * sm_count, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  double avg = 0;
  for (int i = 0; i < sm_count; i++) avg += (double)h[i];
  avg /= sm_count;
  const double warp_instr = (double)iters * 8.0 * (threads / 32);
  return avg / warp_instr;  // SM cycles per warp-level instruction (incl. the LCG/address ALU work)
}

cudaError_t run_microbench(int sm_count, char *report, size_t cap) {
  unsigned long long *d_cycles;
  uint32_t *d_sink;
  cudaError_t e = cudaMalloc(&d_cycles, 256 * sizeof(unsigned long long));
  if (e != cudaSuccess) return e;
  e = cudaMalloc(&d_sink, 16);
  if (e != cudaSuccess) return e;
  size_t n = 0;
  const int T[2] = {1024, 512};
  for (int t = 0; t < 2; t++) {
    const int th = T[t];
    n += snprintf(report + n, cap - n, "threads/SM=%d  (SM cycles per warp instruction)\n", th);
#define QB_MB(mode) n += snprintf(report + n, cap - n, "  %-20s %.3f\n", #mode, microbench_one<mode>(sm_count, d_cycles, d_sink, th));
    QB_MB(MB_ATOMS_FREE)
    QB_MB(MB_ATOMS_22LANES)
    QB_MB(MB_ATOMS_2WAY)
    QB_MB(MB_ATOMS_SAMEADDR)
    QB_MB(MB_LDS_U8)
    QB_MB(MB_LDS_32)
    QB_MB(MB_LDS_128)
    QB_MB(MB_SHFL)
    QB_MB(MB_IMAD)
    QB_MB(MB_ATOMS_LDS_MIX)
#undef QB_MB
  }
  e = cudaDeviceSynchronize();
  cudaFree(d_cycles);
  cudaFree(d_sink);
  return e;
}

}  // namespace qb
