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
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
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

constexpr uint32_t kPadBefore = 16;   // readable bytes in front of a staged buffer (halo lanes, windows)
constexpr uint32_t kPadAfter = 256;   // readable bytes behind it (last partial warp step)
constexpr uint32_t kCandCap = 512;    // Bloom-positive candidates queued per tile

struct Candidate {
  int32_t unit;      // 8-byte unit index in the staged tile
  uint32_t lo, hi;   // 2-bit codes of the 24 bases ending with this unit's 8 bases (see phase A)
};

// shared-memory carve-up, all offsets multiples of 16 bytes
struct SmemLayout {
  uint32_t hist, bloom, exact, lenhist, kmerhist, cand, stage0, stage_stride, seq_off, qual_off, soff_off, slen_off,
      fhit_off, meta, bars, total;
};

__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

__host__ __device__ inline SmemLayout smem_layout(uint32_t half_len, uint32_t len_cap, uint32_t tile_bytes,
                                                  uint32_t stages, int adapters) {
  SmemLayout L;
  const uint32_t buf = kPadBefore + tile_bytes + kPadAfter;
  uint32_t o = 0;
  L.hist = o;
  o += 256u * half_len * 4u;
  L.bloom = o;
  o += adapters ? kBloomBytes : 0u;
  L.exact = o;
  o += adapters ? kExactSlots * 4u : 0u;
  L.lenhist = o;
  o += align16(len_cap * 4u);
  L.kmerhist = o;
  o += align16(len_cap * 4u);
  L.cand = o;
  o += adapters ? align16(kCandCap * (uint32_t)sizeof(Candidate)) : 0u;
  L.stage0 = o;
  L.seq_off = kPadBefore;
  L.qual_off = buf + kPadBefore;
  L.soff_off = 2u * buf;
  L.slen_off = L.soff_off + kMaxTileReads * 4u;
  L.fhit_off = L.slen_off + kMaxTileReads * 4u;
  L.stage_stride = L.fhit_off + kMaxTileReads * 4u;
  o += stages * L.stage_stride;
  L.meta = o;
  o += 16u * kMaxStages + 16u;  // per stage: lo_al, n_reads, span, pad; then 2 candidate counters
  L.bars = o;
  o += 16u * kMaxStages;  // full, empty per stage
  L.total = o;
  return L;
}

// three-input logic op with an explicit truth table (a = 0xF0, b = 0xCC, c = 0xAA); constants passed
// as operands stay in registers, so e.g. (x & A) ^ B is ONE LOP3 instead of two
template <int kLut>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(kLut));
  return d;
}

// loop-invariant SWAR constants, kept in registers
struct KeyConsts {
  uint32_t m5b, x43, m1f, x07, x14, a7f, a3f, m80, mc0, one, qsub;
  __device__ __forceinline__ explicit KeyConsts(uint32_t qbase)
      : m5b(0x5B5B5B5Bu), x43(0x43434343u), m1f(0x1F1F1F1Fu), x07(0x07070707u), x14(0x14141414u),
        a7f(0x7F7F7F7Fu), a3f(0x3F3F3F3Fu), m80(0x80808080u), mc0(0xC0C0C0C0u), one(0x01010101u),
        qsub(qbase * 0x01010101u) {}
};

// Phase A arithmetic for one aligned word of 4 bases + 4 quality bytes.  Returns the 4 key bytes
//   K = code << 6 | s,   code = A0 T1 C2 G3 (quack.c:150),   s = q - qbase in [0,62]
// `nc` gets the inverted codes in bits 7:6 of each byte (other bits undefined); `bad` accumulates
// qs | (qs + 1), whose bits 7:6 are non-zero iff some quality byte is outside the window: then the
// caller re-keys the word to s = 63 (rows nobody reads) and counts its 4 bases exactly.
__device__ __forceinline__ uint32_t key_bytes(uint32_t sw, uint32_t qw, const KeyConsts &c, uint32_t &nc,
                                              uint32_t &bad) {
  // per byte: bit7 of n_cg is 0 iff (b & 0x5B) == 0x43; bit6 of n_g / n_t is 0 iff (b & 0x1F) == 7 / 0x14
  const uint32_t n_cg = lop3<0x6A>(sw, c.m5b, c.x43) + c.a7f;
  const uint32_t n_g = lop3<0x6A>(sw, c.m1f, c.x07) + c.a3f;
  const uint32_t n_t = lop3<0x6A>(sw, c.m1f, c.x14) + c.a3f;
  nc = lop3<0xE4>(n_cg, n_g & n_t, c.m80);  // bit 7 from n_cg, the rest from n_g & n_t
  // a borrow can only start at a byte that is itself out of range, and that byte ends >= 0xC0
  const uint32_t qs = qw - c.qsub;
  bad = lop3<0xFE>(bad, qs, qs + c.one);
  return lop3<0xF2>(qs, nc, c.mc0);  // qs | (~nc & 0xC0C0C0C0)
}
// key bytes of a word that has an out-of-window quality byte: code << 6 | 63
__device__ __forceinline__ uint32_t key_bytes_bad(uint32_t nc) { return (~nc & 0xC0C0C0C0u) | 0x3F3F3F3Fu; }

__device__ __forceinline__ bool word_bad(uint32_t qw, uint32_t qsub) {
  const uint32_t qs = qw - qsub;
  return ((qs | (qs + 0x01010101u)) & 0xC0C0C0C0u) != 0u;
}

// largest r with soff[r] <= abs (reads of a tile are in ascending offset order), -1 if none
__device__ __forceinline__ int find_read(const uint32_t *soff, uint32_t nr, uint32_t abs) {
  int a = 0, b = (int)nr - 1, r = -1;
  while (a <= b) {
    const int m = (a + b) >> 1;
    if (soff[m] <= abs) {
      r = m;
      a = m + 1;
    } else
      b = m - 1;
  }
  return r;
}

// rare path: the 4 bases of a word with an out-of-window quality byte, counted one by one
__device__ __noinline__ uint32_t exact_word(uint32_t sw, uint32_t qw, uint32_t abs0, const uint32_t *soff,
                                            const uint32_t *slen, uint32_t nr, const Accum a) {
  uint32_t n_invalid = 0;
  for (uint32_t j = 0; j < 4; j++) {
    const uint32_t abs = abs0 + j;
    const int r = find_read(soff, nr, abs);
    if (r < 0) continue;
    const uint32_t p = abs - soff[r], len = slen[r];
    if (p >= len || len > a.len_cap) continue;  // alignment slack, or a read the launch rejects anyway
    unsigned long long *row = a.rows + (size_t)p * kRow;
    atomicAdd(&row[kColContent + base_code((sw >> (8 * j)) & 0xFFu)], 1ull);
    const int sc = (int)((qw >> (8 * j)) & 0xFFu) - 33;
    if (sc >= 0 && sc < 91)
      atomicAdd(&row[sc], 1ull);
    else
      n_invalid++;
  }
  return n_invalid;
}

// rare path: window t (0..7) of a Bloom-positive unit.  Re-test it against the filter, then against the
// exact key set (shared-memory hash table; the 2^20-bit map in L2 if the set was too big for it).
__device__ __forceinline__ void confirm_window(const Candidate c, int t, const AdapterSet ad,
                                               const uint8_t *bloom_lane, const uint32_t *exact_s, uint32_t lo_al,
                                               const uint32_t *soff, const uint32_t *slen, uint32_t nr,
                                               uint32_t *fhit) {
  const uint32_t key = __funnelshift_r(c.lo, c.hi, 14 + 2 * t) & 0xFFFFFu;
  const uint32_t p = key * ad.bloom_mul;
  const uint32_t word = *reinterpret_cast<const uint32_t *>(bloom_lane + (p & 0x7F80u));
  if (!((word >> (key & 31u)) & (word >> ((p >> 15) & 31u)) & 1u)) return;
  bool member = false;
  if (ad.exact) {
    for (uint32_t i = exact_slot(key);; i = (i + 1u) & (kExactSlots - 1u)) {
      const uint32_t v = exact_s[i];
      if (v == key) member = true;
      if (v == key || v == kExactEmpty) break;
    }
  } else {
    member = (ad.bitmap[key >> 5] >> (key & 31u)) & 1u;
  }
  if (!member) return;
  const uint32_t abs = lo_al + (uint32_t)(c.unit * 8 + t);  // byte on which the window ends
  const int r = find_read(soff, nr, abs);
  if (r < 0) return;
  const uint32_t pos = abs - soff[r];
  if (pos >= 9u && pos < slen[r]) atomicMin(&fhit[r], pos);  // whole window inside the read
}

__device__ __noinline__ void confirm_candidate(const Candidate c, const AdapterSet ad, const uint8_t *bloom_lane,
                                               const uint32_t *exact_s, uint32_t lo_al, const uint32_t *soff,
                                               const uint32_t *slen, uint32_t nr, uint32_t *fhit) {
  for (int t = 0; t < 8; t++) confirm_window(c, t, ad, bloom_lane, exact_s, lo_al, soff, slen, nr, fhit);
}

template <bool kAdapters, uint32_t kLh>
__global__ void __launch_bounds__(kThreads, 1) fused_kernel(const FusedArgs args) {
  extern __shared__ __align__(128) uint8_t smem[];
  const FusedPlan &plan = args.plan;
  constexpr uint32_t Lh = kLh;  // == plan.half_len
  const uint32_t len_cap = args.a.len_cap;
  const uint32_t S = plan.stages;
  const SmemLayout L = smem_layout(Lh, len_cap, plan.tile_bytes, S, kAdapters);

  uint32_t *hist = reinterpret_cast<uint32_t *>(smem + L.hist);
  uint32_t *lenhist = reinterpret_cast<uint32_t *>(smem + L.lenhist);
  uint32_t *kmerhist = reinterpret_cast<uint32_t *>(smem + L.kmerhist);
  Candidate *cand = reinterpret_cast<Candidate *>(smem + L.cand);
  const uint32_t *exact_s = reinterpret_cast<const uint32_t *>(smem + L.exact);
  uint32_t *meta = reinterpret_cast<uint32_t *>(smem + L.meta);
  uint32_t *cand_count = meta + 4 * kMaxStages;  // [2], alternating per tile
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
    uint32_t *ex = reinterpret_cast<uint32_t *>(smem + L.exact);
    if (args.ad.exact)
      for (uint32_t i = tid; i < kExactSlots; i += kThreads) ex[i] = args.ad.exact[i];
  }
  for (uint32_t s = 0; s < S; s++) {
    uint8_t *st = smem + L.stage0 + s * L.stage_stride;
    uint32_t *fh = reinterpret_cast<uint32_t *>(st + L.fhit_off);
    for (uint32_t i = tid; i < kMaxTileReads; i += kThreads) fh[i] = kNoHit;
    // pads are read (never consumed) by halo lanes: keep them defined
    for (uint32_t i = tid; i < kPadBefore / 4; i += kThreads) {
      reinterpret_cast<uint32_t *>(st)[i] = 0;
      reinterpret_cast<uint32_t *>(st + L.qual_off - kPadBefore)[i] = 0;
    }
  }
  if (tid == 0) {
    cand_count[0] = cand_count[1] = 0;
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
  const KeyConsts kc(plan.qbase);
  const uint32_t qsub = kc.qsub;
  constexpr uint32_t Lh4 = Lh * 4u;
  uint8_t *const hist_lane = smem + L.hist + lane * 4u;  // byte address of this lane's bank column
  unsigned long long n_invalid = 0;
  uint32_t reads_since_flush = 0;

  auto flush = [&]() {
    // all consumer warps have passed the barrier that ends phase H
    const uint32_t npos = min(2u * Lh, len_cap);
    for (uint32_t pos = tid; pos < npos; pos += kCThreads) {
      const bool hi = pos >= Lh;
      const uint32_t col = hi ? pos - Lh : pos;
      const uint32_t sh = hi ? 16u : 0u;
      unsigned long long *row = args.a.rows + (size_t)pos * kRow;
      uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      for (uint32_t sp = 0; sp < 63; sp++) {  // s = 63 rows are the dummies
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
      const uint32_t lc = lenhist[pos], kcnt = kmerhist[pos];
      if (lc) atomicAdd(&row[kColLength], (unsigned long long)lc);
      if (kcnt) atomicAdd(&row[kColKmer], (unsigned long long)kcnt);
      lenhist[pos] = 0;
      kmerhist[pos] = 0;
    }
    consumer_bar();
    uint4 *h4 = reinterpret_cast<uint4 *>(hist);
    for (uint32_t i = tid; i < 64u * Lh; i += kCThreads) h4[i] = make_uint4(0, 0, 0, 0);
    consumer_bar();
  };

  uint32_t it = 0;
  for (uint32_t tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x, ++it) {
    const uint32_t s = it % S, use = it / S;
    mbar_wait(&bars[2 * s], use & 1u);
    const uint32_t lo_al = meta[4 * s + 0];
    const uint32_t nr = meta[4 * s + 1];
    const uint32_t span = meta[4 * s + 2];
    uint8_t *st = smem + L.stage0 + s * L.stage_stride;
    const uint32_t *soff = reinterpret_cast<const uint32_t *>(st + L.soff_off);
    const uint32_t *slen = reinterpret_cast<const uint32_t *>(st + L.slen_off);
    uint32_t *fhit = reinterpret_cast<uint32_t *>(st + L.fhit_off);
    uint8_t *kbuf = st + L.qual_off;  // phase A overwrites the quality bytes with the key bytes K
    uint32_t *ccount = &cand_count[it & 1u];

    // ---------------- phase A: flat over the tile, K written in place of the quality bytes ----------------
    if (!kAdapters) {
      const uint4 *s4 = reinterpret_cast<const uint4 *>(st + L.seq_off);
      uint4 *q4 = reinterpret_cast<uint4 *>(st + L.qual_off);
      const uint32_t nvec = span >> 4;
      for (uint32_t v = tid; v < nvec; v += kCThreads) {
        const uint4 sv = s4[v], qv = q4[v];
        uint32_t n0, n1, n2, n3, bad = 0;
        uint4 K;
        K.x = key_bytes(sv.x, qv.x, kc, n0, bad);
        K.y = key_bytes(sv.y, qv.y, kc, n1, bad);
        K.z = key_bytes(sv.z, qv.z, kc, n2, bad);
        K.w = key_bytes(sv.w, qv.w, kc, n3, bad);
        if (bad & 0xC0C0C0C0u) {  // rare: re-key the offending words to the dummy rows, count them exactly
          const uint32_t abs0 = lo_al + v * 16u;
          if (word_bad(qv.x, qsub)) K.x = key_bytes_bad(n0), n_invalid += exact_word(sv.x, qv.x, abs0, soff, slen, nr, args.a);
          if (word_bad(qv.y, qsub)) K.y = key_bytes_bad(n1), n_invalid += exact_word(sv.y, qv.y, abs0 + 4u, soff, slen, nr, args.a);
          if (word_bad(qv.z, qsub)) K.z = key_bytes_bad(n2), n_invalid += exact_word(sv.z, qv.z, abs0 + 8u, soff, slen, nr, args.a);
          if (word_bad(qv.w, qsub)) K.w = key_bytes_bad(n3), n_invalid += exact_word(sv.w, qv.w, abs0 + 12u, soff, slen, nr, args.a);
        }
        q4[v] = K;
      }
    } else {
      // One 8-byte unit (8 bases) per lane.  A warp step covers 30 new units; lanes 0 and 1 re-read the
      // two units in front so that every 10-mer window ending in lanes 2..31 finds its 9 earlier bases
      // inside the warp (two shuffles).  Halo lanes compute but never store or report.
      const uint2 *s8 = reinterpret_cast<const uint2 *>(st + L.seq_off);
      uint2 *q8 = reinterpret_cast<uint2 *>(st + L.qual_off);
      const uint8_t *bloom_b = smem + L.bloom;
      const uint32_t lane4 = lane * 4u;
      const int n8 = (int)(span >> 3);
      const uint32_t nsteps = ((uint32_t)n8 + 29u) / 30u;
      const uint32_t M = args.ad.bloom_mul;
      for (uint32_t step = warp; step < nsteps; step += kCW) {
        const int u = (int)(step * 30u + lane) - 2;  // -2, -1 land in the pad in front; > n8 in the pad behind
        const uint2 sv = s8[u], qv = q8[u];
        uint32_t n0, n1, bad = 0;
        uint2 K;
        K.x = key_bytes(sv.x, qv.x, kc, n0, bad);
        K.y = key_bytes(sv.y, qv.y, kc, n1, bad);
        const bool own = lane >= 2u && u < n8;
        if (own) {
          if (bad & 0xC0C0C0C0u) {
            const uint32_t abs0 = lo_al + (uint32_t)u * 8u;
            if (word_bad(qv.x, qsub)) K.x = key_bytes_bad(n0), n_invalid += exact_word(sv.x, qv.x, abs0, soff, slen, nr, args.a);
            if (word_bad(qv.y, qsub)) K.y = key_bytes_bad(n1), n_invalid += exact_word(sv.y, qv.y, abs0 + 4u, soff, slen, nr, args.a);
          }
          q8[u] = K;
        }
        // 8 bases -> 16 bits, first base least significant (codes are the inverted bits 7:6 of n0/n1)
        const uint32_t c0 = ~n0 & 0xC0C0C0C0u, c1 = ~n1 & 0xC0C0C0C0u;
        const uint32_t p16 = ((c0 * 0x41041u) >> 24) | (((c1 * 0x41041u) >> 16) & 0xFF00u);
        const uint32_t prev = __shfl_up_sync(0xffffffffu, p16, 1);
        const uint32_t pp = __shfl_up_sync(0xffffffffu, p16, 2);
        const uint32_t lo = (pp & 0xFFFFu) | (prev << 16);  // 16 earlier bases
        const uint32_t hi = p16;                             // my 8 bases
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
          const uint32_t wj = __funnelshift_r(lo, hi, 14 + 2 * t);  // low 20 bits: window ending at my base t
          const uint32_t p = wj * M;                                // low 20 bits depend on the window only
          const uint32_t word = *reinterpret_cast<const uint32_t *>(bloom_b + ((p & 0x7F80u) | lane4));
          acc |= __funnelshift_r(word, 0u, wj) & __funnelshift_r(word, 0u, p >> 15);
        }
        if (own && (acc & 1u)) {
          const Candidate c{u, lo, hi};
          const uint32_t idx = atomicAdd(ccount, 1u);
          if (idx < kCandCap)
            cand[idx] = c;
          else
            confirm_candidate(c, args.ad, bloom_b + lane4, exact_s, lo_al, soff, slen, nr, fhit);
        }
      }
    }
    consumer_bar();

    // ---------------- phase A2: confirm queued Bloom positives, spread over all threads ----------------
    if (kAdapters) {
      const uint32_t nc = min(*ccount, kCandCap);
      for (uint32_t i = tid; i < nc * 8u; i += kCThreads)  // one (unit, window) pair per thread: all warps share it
        confirm_window(cand[i >> 3], (int)(i & 7u), args.ad, smem + L.bloom + lane * 4u, exact_s, lo_al, soff, slen,
                       nr, fhit);
    }

    // ---------------- phase H: one warp per read, lane <-> position, one shared atomic per base ----------------
    // Every 32-position step is straight-line code: which half-word it counts in and its column
    // offset are compile-time constants (kLh template); only the lane predicate pos < len is dynamic.
    // Explicit shared-space PTX keeps a step at ISETP + LDS.U8 + IMAD + ATOMS.
#define QB_STEP(sidx)                                                                               \
  if (rem > (sidx) * 32u) {                                                                         \
    constexpr uint32_t pos0 = (sidx) * 32u;                                                         \
    constexpr bool hi_half = pos0 >= Lh;                                                            \
    const uint32_t k = lds_u8(kb_s + pos0);                                                         \
    red_shared_add(k * Lh4 + hist_s + (hi_half ? pos0 - Lh : pos0) * 4u, hi_half ? 0x10000u : 1u);  \
  }
    {
      const uint32_t kbuf_s = smem_u32(kbuf) + lane - lo_al;
      const uint32_t hist_s = smem_u32(hist_lane);
      for (uint32_t r = warp; r < nr; r += kCW) {
        const uint32_t len = slen[r];
        if (len > len_cap) continue;  // reported in the bookkeeping pass below
        const uint32_t kb_s = kbuf_s + soff[r];
        const uint32_t rem = len > lane ? len - lane : 0u;  // lane handles positions lane, lane+32, ... < len
        QB_STEP(0) QB_STEP(1) QB_STEP(2)
        if (len > 96u) {
          QB_STEP(3) QB_STEP(4)
          if constexpr (Lh == 96u) {
            QB_STEP(5)
          } else if (len > 160u) {
            QB_STEP(5) QB_STEP(6) QB_STEP(7)
            if (len > 256u) { QB_STEP(8) QB_STEP(9) }
          }
        }
      }
    }
#undef QB_STEP
    consumer_bar();  // K bytes, candidates and fhit of this tile are final / no longer needed

    // ---------------- bookkeeping: per-read counters, one thread per read ----------------
    for (uint32_t r = tid; r < nr; r += kCThreads) {
      const uint32_t len = slen[r];
      if (len > len_cap) {
        atomicAdd(&args.a.counters[kCntError], 1ull);
        continue;
      }
      if (len) atomicAdd(&lenhist[len - 1u], 1u);  // quack.c:219
      if (kAdapters) {
        const uint32_t fh = fhit[r];
        if (fh != kNoHit) {
          if (fh + 1u < len) atomicAdd(&kmerhist[fh + 1u], 1u);  // quack.c:215-217
          fhit[r] = kNoHit;
        }
      }
    }
    if (kAdapters && tid == 0) *ccount = 0;
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[2 * s + 1]);  // stage buffers free for the producer

    reads_since_flush += nr;
    if (reads_since_flush + RT > 65535u) {  // u16 counters: flush before any bin can wrap
      consumer_bar();
      flush();
      reads_since_flush = 0;
    }
  }
  consumer_bar();
  flush();
  n_invalid = warp_sum(n_invalid);
  if (lane == 0 && n_invalid) atomicAdd(&args.a.counters[kCntInvalidQual], n_invalid);
}

FusedPlan fused_plan(uint32_t len_cap, uint32_t batch_max_len, int adapters, int sm_count,
                     uint32_t smem_optin, uint32_t qbase) {
  FusedPlan p;
  memset(&p, 0, sizeof p);
  if (len_cap == 0) return p;
  if (len_cap > 320u) return p;  // beyond the shared-memory histogram: simple kernel
  const uint32_t half = len_cap <= 192u ? 96u : 160u;  // the two instantiations of fused_kernel
  p.half_len = half;
  p.qbase = qbase;
  if (batch_max_len == 0 || batch_max_len > len_cap) batch_max_len = len_cap;
  for (uint32_t stages = kMaxStages; stages >= 2; stages--) {
    const SmemLayout L0 = smem_layout(half, len_cap, 0, stages, adapters);
    if (L0.total + 128u >= smem_optin) continue;
    uint32_t avail = smem_optin - L0.total - 128u;  // L0 already holds every pad
    uint32_t tile = (avail / (2u * stages)) & ~15u;
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
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(fused_kernel<true, 96>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
  if ((e = cudaFuncSetAttribute(fused_kernel<false, 96>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
  if ((e = cudaFuncSetAttribute(fused_kernel<true, 160>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
  return cudaFuncSetAttribute(fused_kernel<false, 160>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
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
  if (plan.half_len == 96u) {
    if (ad.enabled)
      fused_kernel<true, 96><<<grid, kThreads, plan.smem_bytes, stream>>>(args);
    else
      fused_kernel<false, 96><<<grid, kThreads, plan.smem_bytes, stream>>>(args);
  } else {
    if (ad.enabled)
      fused_kernel<true, 160><<<grid, kThreads, plan.smem_bytes, stream>>>(args);
    else
      fused_kernel<false, 160><<<grid, kThreads, plan.smem_bytes, stream>>>(args);
  }
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
