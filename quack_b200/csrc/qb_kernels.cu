// qb_kernels.cu -- sm_100a kernels for quack's per-read statistics accumulation
// (reference: the while loop of read_fastq(), quack.c:193-221).
//
// Two kernels compute the same thing:
//   simple_kernel : one warp per read, 64-bit global atomics.  Any read length.  The fallback
//                   for len_cap beyond what the shared-memory histogram holds (still CUDA).
//   fused_kernel  : persistent, one CTA per SM.  A producer warp streams tiles of whole reads
//                   (seq bytes, qual bytes, offsets, lengths) into shared memory with 1-D TMA bulk
//                   copies (cp.async.bulk + mbarrier, 2-4 stages).  31 consumer warps then make two
//                   passes over the staged tile:
//                     phase A (flat, one aligned 16-byte unit per lane): SWAR base -> 2-bit code,
//                       score byte -> 6-bit bin, fused into one key byte K = code<<6 | bin per
//                       base, written over the quality bytes; with -a, one 7-mer "anchor" per 4
//                       bases is tested against an exact 2^14-bit map in shared memory (every 10-mer
//                       window contains exactly one anchor), hits are queued and the 4 windows around
//                       each are confirmed against the exact key set.
//                     phase H (one warp per read, lane <-> position): one shared-memory atomic
//                       per base into a JOINT (code,score) x position histogram of packed u16
//                       counters whose bank is the position mod 32, so a warp's 32 updates never
//                       conflict.  content[] and scores[] are its marginals, taken at flush time.
//                   Counters are flushed to the u64 global accumulator with one RED per non-zero
//                   bin, at the end of the launch or every 65535 reads per CTA.
//
// No tensor cores: the path is an integer histogram.
#include "qb_dev.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace qb {

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
// fused kernel (v3)
//
// The kernel is bound by instruction issue (4 warp instructions per cycle per SM) and, second, by
// shared-memory wavefronts, not by HBM: at 2 bytes per base the HBM roofline leaves 2.9 SM cycles per 32
// bases.  Everything below is organised to spend as few warp instructions per base as possible:
//   * one CTA-wide barrier per tile (two with -a); stages are released through the mbarrier, the
//     per-read counters of a tile are settled while the next tile is in flight;
//   * phase A is flat SWAR over aligned 16-byte units (4 bases per 32-bit lane operation);
//   * phase H is straight-line code per read: N byte loads, N multiply-adds, N shared atomics, selected
//     once per tile when all reads of the tile have the same length (reduction barrier), else per read;
//   * the adapter scan probes ONE 7-mer per 4 bases (the anchor that every 10-mer window must contain)
//     instead of one 10-mer per base; anchor hits are queued and confirmed against the exact key set.
// ------------------------------------------------------------------------------------------

constexpr int kCW = kFusedConsumerWarps;       // consumer warps
constexpr int kCThreads = kCW * 32;
constexpr int kThreads = kCThreads + 32;       // + producer warp
constexpr int kMaxStages = 4;
constexpr uint32_t kPadAfter = 48;             // readable bytes behind a staged buffer (look-ahead unit)
constexpr uint32_t kCandCap = 768;             // anchor hits queued per tile (one entry per anchor)
constexpr uint32_t kIdxCap = kMaxTileReads + 8;  // staged offsets/lengths: tile reads + alignment slack

__device__ __forceinline__ void consumer_bar() {
  asm volatile("bar.sync 1, %0;" ::"n"(kCThreads) : "memory");
}
// barrier of the consumer warps that also ANDs a predicate over all of them
__device__ __forceinline__ uint32_t consumer_bar_and(uint32_t pred) {
  uint32_t out;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.u32 q, %1, 0;\n"
      "bar.red.and.pred p, 1, %2, q;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(out)
      : "r"(pred), "n"(kCThreads)
      : "memory");
  return out;
}

struct FusedArgs {
  BatchView b;
  Accum a;
  AdapterSet ad;
  FusedPlan plan;
  uint32_t n_tiles;
};


// shared-memory carve-up, all offsets multiples of 16 bytes.  Everything a tile needs (bytes, index
// slices, its description and its two mbarriers) sits in one stage block at fixed offsets, so the tile
// loop carries a single stage address.
struct SmemLayout {
  uint32_t hist, afilt, exact, lenhist, kmerhist, cand, fhit, ccount, stage0, stage_stride;
  uint32_t buf;                                  // bytes of one staged byte buffer (seq or qual)
  uint32_t o_qual, o_soff, o_slen, o_meta, o_full, o_empty;  // offsets inside a stage block
  uint32_t total;
};

__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

__host__ __device__ inline SmemLayout smem_layout(uint32_t half_len, uint32_t len_cap, uint32_t tile_bytes,
                                                  uint32_t stages, int adapters) {
  SmemLayout L;
  L.buf = tile_bytes + kPadAfter;
  uint32_t o = 0;
  L.hist = o;
  o += kHistRows * half_len * 4u;
  L.afilt = o;
  o += adapters ? kAnchorSmemBytes : 0u;
  L.exact = o;
  o += adapters ? kExactSlots * 4u : 0u;
  L.lenhist = o;
  o += align16(len_cap * 4u);
  L.kmerhist = o;
  o += adapters ? align16(len_cap * 4u) : 0u;
  L.cand = o;
  o += adapters ? 2u * kCandCap * 8u : 0u;
  L.fhit = o;
  o += adapters ? 2u * kMaxTileReads * 4u : 0u;
  L.ccount = o;
  o += 16u;  // 3 candidate counters, rotating per tile
  L.stage0 = o;
  L.o_qual = L.buf;
  L.o_soff = 2u * L.buf;
  L.o_slen = L.o_soff + kIdxCap * 4u;
  L.o_meta = L.o_slen + kIdxCap * 4u;  // uint4: lo_al, n_reads, span, first index slot
  L.o_full = L.o_meta + 16u;
  L.o_empty = L.o_full + 8u;
  L.stage_stride = L.o_meta + 32u;
  o += stages * L.stage_stride;
  L.total = o;
  return L;
}



// ---- phase H building blocks ----
// Shared-memory instruction throughput (the MIO queue), not issue, bounds this phase: a byte load costs
// about as much as two word loads and a shared atomic ~1.3 cycles whatever the number of active lanes.
// So a read is walked in WORD steps -- lane l loads the 4 key bytes of positions 128 s + 4 l .. + 3 (two
// aligned 32-bit loads and a funnel shift, since reads start at any byte) and issues 4 atomics -- and its
// last <= 32 positions in one BYTE step (lane <-> position).
//
// Column of a position inside a histogram row.  Positions [0, Lh) count in the low half-word of a u32
// counter, [Lh, 2 Lh) in the high half-word.  Within a half, position 4 q + t sits in column
// t * Lh/4 + q, so the t-th atomic of a word step touches consecutive banks across the lanes; the high
// half is rotated by kRot columns so that a word step which straddles Lh still puts every lane in its own
// bank.  Any 32 consecutive positions of one half also fall into 32 different banks (byte steps).
template <uint32_t Lh>
struct HLayout {
  static constexpr uint32_t kRot = (Lh == 96u) ? 24u : 8u;  // Lh = 96: 128 = Lh + 32; Lh = 160: 256 = Lh + 96
  __host__ __device__ static inline uint32_t column(uint32_t pos, uint32_t &unit) {
    const bool hi = pos >= Lh;
    const uint32_t pp = hi ? pos - Lh : pos;
    uint32_t col = (pp & 3u) * (Lh / 4u) + (pp >> 2);
    if (hi) {
      col += kRot;
      if (col >= Lh) col -= Lh;
    }
    unit = hi ? 0x10000u : 1u;
    return col;
  }
};

// per-lane constants of the word steps: shared address of the lane's 4 columns (row 0) and the counter unit
template <uint32_t Lh>
struct HLane {
  static constexpr int kWordSteps = (int)((2u * Lh + 127u) / 128u);
  uint32_t hb[kWordSteps][4];
  uint32_t unit[kWordSteps];
  __device__ __forceinline__ void init(uint32_t hist_s, uint32_t lane) {
#pragma unroll
    for (int s = 0; s < kWordSteps; s++) {
      const uint32_t p = 128u * s + 4u * lane;
      uint32_t u = 0;
#pragma unroll
      for (int t = 0; t < 4; t++) hb[s][t] = hist_s + 4u * HLayout<Lh>::column(min(p + t, 2u * Lh - 1u), u);
      unit[s] = p < 2u * Lh ? u : 0u;
    }
  }
};


// shape of a read of length L: full word steps, then a partial word step (kind 2), a byte step (kind 1) or nothing
__device__ __forceinline__ uint32_t h_shape(uint32_t L) {
  const uint32_t rem = L & 127u;
  return (L >> 7) * 3u + (rem == 0u ? 0u : (rem <= 32u ? 1u : 2u));
}

// the 4 key bytes of positions 128 s + 4 lane .. + 3 of the read whose lane-th word starts at kw (any alignment)
__device__ __forceinline__ uint32_t h_load_word(uint32_t kw) {
  const uint32_t al = kw & ~3u;
  const uint32_t w0 = lds_u32(al), w1 = lds_u32(al + 4u);
  return __funnelshift_r(w0, w1, kw << 3);  // shift = 8 * (kw & 3): the funnel shift wraps at 32
}
__device__ __forceinline__ void h_red_word(uint32_t k4, const uint32_t (&hb)[4], uint32_t i0, uint32_t i1, uint32_t i2,
                                           uint32_t i3, uint32_t row_bytes) {
  red_shared_add<0>(byte_of(k4, 0) * row_bytes + hb[0], i0);
  red_shared_add<0>(byte_of(k4, 1) * row_bytes + hb[1], i1);
  red_shared_add<0>(byte_of(k4, 2) * row_bytes + hb[2], i2);
  red_shared_add<0>(byte_of(k4, 3) * row_bytes + hb[3], i3);
}

// One read.  kb = shared address of the read's first key byte; NWF full word steps; KIND as in h_shape.
// pinc[] = increments of the partial word step, binc / bhb = increment and column address of the byte step.
template <uint32_t Lh, int NWF, int KIND>
struct HRead {
  static constexpr int kWords = NWF + (KIND == 2 ? 1 : 0);
  uint32_t k4[kWords > 0 ? kWords : 1];
  uint32_t kbyte;
  __device__ __forceinline__ void load(uint32_t kb, uint32_t lane) {
#pragma unroll
    for (int s = 0; s < kWords; s++) k4[s] = h_load_word(kb + 128u * s + 4u * lane);
    if (KIND == 1) kbyte = lds_u8(kb + 128u * NWF + lane);
  }
  __device__ __forceinline__ void red(const HLane<Lh> &hl, const uint32_t (&pinc)[4], uint32_t bhb, uint32_t binc) const {
#pragma unroll
    for (int s = 0; s < NWF; s++) h_red_word(k4[s], hl.hb[s], hl.unit[s], hl.unit[s], hl.unit[s], hl.unit[s], Lh * 4u);
    if (KIND == 2) h_red_word(k4[NWF], hl.hb[NWF], pinc[0], pinc[1], pinc[2], pinc[3], Lh * 4u);
    if (KIND == 1) red_shared_add<0>(kbyte * (Lh * 4u) + bhb, binc);
  }
};

// increments / addresses of the tail of a read of length L (lanes behind the read's end add 0 to whatever
// bin the stray byte selects, which keeps the steps free of predicates and branches)
template <uint32_t Lh, int NWF, int KIND>
__device__ __forceinline__ void h_tail(const HLane<Lh> &hl, uint32_t hist_s, uint32_t L, uint32_t lane, uint32_t (&pinc)[4],
                                       uint32_t &bhb, uint32_t &binc) {
  pinc[0] = pinc[1] = pinc[2] = pinc[3] = 0;
  bhb = hist_s;
  binc = 0;
  if (KIND == 2) {
    const uint32_t p = 128u * NWF + 4u * lane;
#pragma unroll
    for (int t = 0; t < 4; t++) pinc[t] = p + t < L ? hl.unit[NWF < HLane<Lh>::kWordSteps ? NWF : 0] : 0u;
  }
  if (KIND == 1) {
    const uint32_t p = 128u * NWF + lane;
    uint32_t u;
    bhb = hist_s + 4u * HLayout<Lh>::column(min(p, 2u * Lh - 1u), u);
    binc = p < L ? u : 0u;
  }
}

// all reads of the tile have length L and lie back to back from kb0: two reads in flight per warp
template <uint32_t Lh, int NWF, int KIND>
__device__ __forceinline__ void h_uniform(const HLane<Lh> &hl, uint32_t hist_s, uint32_t kb0, uint32_t L, uint32_t nr,
                                          uint32_t warp, uint32_t lane) {
  uint32_t pinc[4], bhb, binc;
  h_tail<Lh, NWF, KIND>(hl, hist_s, L, lane, pinc, bhb, binc);
  uint32_t kb = kb0 + warp * L;
  const uint32_t stride = (uint32_t)kCW * L;
  uint32_t r = warp;
  for (; r + kCW < nr; r += 2 * kCW) {
    HRead<Lh, NWF, KIND> a, b;
    a.load(kb, lane);
    b.load(kb + stride, lane);
    a.red(hl, pinc, bhb, binc);
    b.red(hl, pinc, bhb, binc);
    kb += 2u * stride;
  }
  if (r < nr) {
    HRead<Lh, NWF, KIND> a;
    a.load(kb, lane);
    a.red(hl, pinc, bhb, binc);
  }
}
template <uint32_t Lh, int NWF, int KIND>
__device__ __forceinline__ void h_one(const HLane<Lh> &hl, uint32_t hist_s, uint32_t kb, uint32_t L, uint32_t lane) {
  uint32_t pinc[4], bhb, binc;
  h_tail<Lh, NWF, KIND>(hl, hist_s, L, lane, pinc, bhb, binc);
  HRead<Lh, NWF, KIND> a;
  a.load(kb, lane);
  a.red(hl, pinc, bhb, binc);
}


// rare path of phase A: re-key the words of a 16-byte unit whose quality bytes fall outside the window,
// counting their bases exactly
__device__ __noinline__ uint32_t fix_bad_unit(uint4 sv, uint4 qv, uint4 &K, uint32_t n0, uint32_t n1, uint32_t n2,
                                              uint32_t n3, uint32_t qsub, uint32_t abs0, uint32_t idx_s, uint32_t nr,
                                              const Accum a) {
  const uint32_t *soff = shared_ptr<const uint32_t>(idx_s);
  const uint32_t *slen = soff + kIdxCap;
  uint32_t n_invalid = 0;
  if (word_bad(qv.x, qsub)) K.x = key_bytes_bad(n0), n_invalid += exact_word(sv.x, qv.x, abs0, soff, slen, nr, a);
  if (word_bad(qv.y, qsub)) K.y = key_bytes_bad(n1), n_invalid += exact_word(sv.y, qv.y, abs0 + 4u, soff, slen, nr, a);
  if (word_bad(qv.z, qsub)) K.z = key_bytes_bad(n2), n_invalid += exact_word(sv.z, qv.z, abs0 + 8u, soff, slen, nr, a);
  if (word_bad(qv.w, qsub)) K.w = key_bytes_bad(n3), n_invalid += exact_word(sv.w, qv.w, abs0 + 12u, soff, slen, nr, a);
  return n_invalid;
}

// One window of a queued anchor hit.  `lo`/`hi` hold the 25 bases from the start of 16-byte unit `unit`
// (2 bits per base, first base least significant); `w` is the window start in bases from the unit start.
// A window found in the exact key set (shared-memory hash table; the 2^20-bit map in L2 if the set was
// too big for it) whose 10 bases lie inside one read lowers that read's first-hit position.  A hit that
// ends on the last base of its read is dropped: it can only be the first hit if there is no other, and
// then the reference counts nothing (quack.c:215).
__device__ __forceinline__ void confirm_window(uint32_t lo, uint32_t hi, uint32_t unit, uint32_t w, const AdapterSet ad,
                                               uint32_t exact_s, uint32_t lo_al, uint32_t idx_s, uint32_t nr,
                                               uint32_t uniform_len, uint32_t fhit_s) {
  const uint32_t key = __funnelshift_r(lo, hi, 2u * w) & 0xFFFFFu;
  bool member;
  if (ad.exact)
    member = lds_u32(exact_s + exact_off1(key)) == key || lds_u32(exact_s + exact_off2(key)) == key;
  else
    member = (ad.bitmap[key >> 5] >> (key & 31u)) & 1u;
  if (!member) return;
  const uint32_t abs = lo_al + unit * 16u + w + 9u;  // byte on which the window ends
  uint32_t r, pos, len;
  if (uniform_len) {  // reads of one length, back to back: divide instead of searching
    const uint32_t d = abs - lds_u32(idx_s);
    if ((int32_t)d < 0) return;
    // d < 2^16 and (d + 0.5) / len is never closer than 1/(2 len) to an integer: the float quotient is exact
    r = (uint32_t)(((float)d + 0.5f) * __frcp_rn((float)uniform_len));
    pos = d - r * uniform_len;
    len = uniform_len;
    if (r >= nr) return;
  } else {
    const uint32_t *soff = shared_ptr<const uint32_t>(idx_s);
    const int rr = find_read(soff, nr, abs);
    if (rr < 0) return;
    r = (uint32_t)rr;
    pos = abs - soff[r];
    len = soff[kIdxCap + r];
  }
  if (pos >= 9u && pos + 1u < len) atomicMin(shared_ptr<uint32_t>(fhit_s) + r, pos);  // whole window inside the read
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
  uint32_t *fhit_all = reinterpret_cast<uint32_t *>(smem + L.fhit);  // [2][kMaxTileReads]

  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  const uint32_t warp = tid >> 5;
  const uint32_t smem_s = smem_u32(smem);

  // ---- prologue: zero histograms, load the adapter filter, init barriers ----
  {
    uint4 *h4 = reinterpret_cast<uint4 *>(hist);
    for (uint32_t i = tid; i < kHistRows / 4u * Lh; i += kThreads) h4[i] = make_uint4(0, 0, 0, 0);
  }
  for (uint32_t i = tid; i < len_cap; i += kThreads) {
    lenhist[i] = 0;
    if (kAdapters) kmerhist[i] = 0;
  }
  if (kAdapters) {
    uint32_t *af = reinterpret_cast<uint32_t *>(smem + L.afilt);
    for (uint32_t i = tid; i < kAnchorWords * kAnchorCopies; i += kThreads) af[i] = args.ad.anchor[i / kAnchorCopies];
    uint32_t *ex = reinterpret_cast<uint32_t *>(smem + L.exact);
    if (args.ad.exact)
      for (uint32_t i = tid; i < kExactSlots; i += kThreads) ex[i] = args.ad.exact[i];
    for (uint32_t i = tid; i < 2u * kMaxTileReads; i += kThreads) fhit_all[i] = kNoHit;
  }
  if (tid == 0) {
    uint32_t *cc = reinterpret_cast<uint32_t *>(smem + L.ccount);
    cc[0] = cc[1] = cc[2] = cc[3] = 0;
    for (uint32_t s = 0; s < S; s++) {
      uint8_t *st = smem + L.stage0 + s * L.stage_stride;
      mbar_init(reinterpret_cast<uint64_t *>(st + L.o_full), 1);
      mbar_init(reinterpret_cast<uint64_t *>(st + L.o_empty), kCW);
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
      uint32_t s = 0, ph = 1;  // a fresh "empty" barrier passes a wait on parity 1
      for (uint32_t tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x) {
        uint8_t *st = smem + L.stage0 + s * L.stage_stride;
        uint64_t *full = reinterpret_cast<uint64_t *>(st + L.o_full);
        mbar_wait_relaxed(smem_u32(st + L.o_empty), ph);
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
        const uint32_t r0_al = r0 & ~3u;  // TMA sources are 16-byte aligned
        const uint32_t idx_bytes = nr ? ((r1 - r0_al + 3u) & ~3u) * 4u : 0u;
        *reinterpret_cast<uint4 *>(st + L.o_meta) = make_uint4(lo_al, nr, span, r0 - r0_al);
        mbar_arrive_expect_tx(full, 2u * span + 2u * idx_bytes);
        if (span) {
          bulk_g2s(st, args.b.seq + lo_al, span, full);
          bulk_g2s(st + L.o_qual, args.b.qual + lo_al, span, full);
        }
        if (idx_bytes) {
          bulk_g2s(st + L.o_soff, args.b.offset + r0_al, idx_bytes, full);
          bulk_g2s(st + L.o_slen, args.b.length + r0_al, idx_bytes, full);
        }
        reads_seen += nr;
        if (++s == S) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (reads_seen) atomicAdd(&args.a.counters[kCntReads], reads_seen);
    }
    return;
  }

  // ================================= consumer warps =================================
  // launch constants the tile loop needs, pinned in registers; everything else is an immediate
  const uint32_t qsub = pin(plan.qbase * 0x01010101u);
  const KeyConsts kc(plan.qbase);
  const uint32_t buf = pin(L.buf);                         // stage + buf = quality / key bytes
  const uint32_t o_meta = pin(L.o_meta);                   // soff = o_meta - 2*kIdxCap*4, slen = o_meta - kIdxCap*4
  const uint32_t stride = pin(L.stage_stride);
  const uint32_t stage0_s = pin(smem_s + L.stage0);
  const uint32_t stage_end_s = pin(smem_s + L.stage0 + S * L.stage_stride);
  const uint32_t hist_s = smem_s + L.hist;
  HLane<Lh> hl;  // this lane's histogram columns for the word steps
  hl.init(hist_s, lane);
  const uint32_t tid16 = tid * 16u;
  const uint32_t afilt_s = smem_s + L.afilt;
  const uint32_t afilt_copy = (lane >> 2) * 4u;  // this lane's copy of the anchor map (8 copies, 32-byte rows)
  const uint32_t exact_s = smem_s + L.exact;
  const uint32_t ccount_s = smem_s + L.ccount;
  const uint32_t cand_s = smem_s + L.cand;
  const uint32_t fhit_s0 = smem_s + L.fhit;
  const uint32_t lenhist_s = smem_s + L.lenhist;
  const uint32_t kmerhist_s = smem_s + L.kmerhist;
  unsigned long long n_invalid = 0;
  uint32_t reads_since_flush = 0;

  auto flush = [&]() {
    // all consumer warps have passed the barrier that ends phase H
    const uint32_t npos = min(2u * Lh, len_cap);
    for (uint32_t pos = tid; pos < npos; pos += kCThreads) {
      uint32_t unit;
      const uint32_t col = HLayout<Lh>::column(pos, unit);
      const uint32_t sh = unit == 1u ? 0u : 16u;
      unsigned long long *row = args.a.rows + (size_t)pos * kRow;
      uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      for (uint32_t sp = 0; sp < kScoreBins; sp++) {  // the rows of s = kScoreBins are the dummies
        const uint32_t v0 = (hist[(4u * sp + 0u) * Lh + col] >> sh) & 0xFFFFu;
        const uint32_t v1 = (hist[(4u * sp + 1u) * Lh + col] >> sh) & 0xFFFFu;
        const uint32_t v2 = (hist[(4u * sp + 2u) * Lh + col] >> sh) & 0xFFFFu;
        const uint32_t v3 = (hist[(4u * sp + 3u) * Lh + col] >> sh) & 0xFFFFu;
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
      const uint32_t lc = lenhist[pos];
      if (lc) atomicAdd(&row[kColLength], (unsigned long long)lc);
      lenhist[pos] = 0;
      if (kAdapters) {
        const uint32_t kcnt = kmerhist[pos];
        if (kcnt) atomicAdd(&row[kColKmer], (unsigned long long)kcnt);
        kmerhist[pos] = 0;
      }
    }
    consumer_bar();
    uint4 *h4 = reinterpret_cast<uint4 *>(hist);
    for (uint32_t i = tid; i < kHistRows / 4u * Lh; i += kCThreads) h4[i] = make_uint4(0, 0, 0, 0);
    consumer_bar();
  };

  // first-hit positions of the previous tile -> adapter histogram (quack.c:215-217), one thread per read
  auto settle_hits = [&](uint32_t buf_index, uint32_t nr_prev) {
    if (tid < nr_prev) {  // a tile holds at most kMaxTileReads <= kCThreads reads
      const uint32_t a = fhit_s0 + (buf_index * kMaxTileReads + tid) * 4u;
      const uint32_t p = lds_u32(a);
      if (p != kNoHit) {
        red_shared_add<4>(kmerhist_s + p * 4u, 1u);  // bin p + 1
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(kNoHit) : "memory");
      }
    }
  };

  uint32_t st = stage0_s;  // shared-space address of the current stage block
  uint32_t ph = 0, it = 0, nr_prev = 0;
  uint32_t cslot = 0;  // it % 3
  for (uint32_t tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x, ++it) {
    mbar_wait(st + o_meta + 16u, ph);
    const uint4 mt = lds_u128(st + o_meta);
    const uint32_t lo_al = mt.x, nr = mt.y, span = mt.z;
    const uint32_t idx_s = st + o_meta - 2u * kIdxCap * 4u + mt.w * 4u;  // soff[]; slen[] is kIdxCap entries behind
    const uint32_t kbuf_s = st + buf;  // phase A overwrites the quality bytes with the key bytes K

    // is the tile uniform (all reads of one length, back to back)?  one compare per read, ANDed by the barrier
    uint32_t len0 = 0, soff0 = 0;
    if (nr) {
      soff0 = lds_u32(idx_s);
      len0 = lds_u32(idx_s + kIdxCap * 4u);
    }
    uint32_t my_ok = 1u;
    if (tid < nr) {
      const uint32_t o = lds_u32(idx_s + tid * 4u), l = lds_u32(idx_s + (kIdxCap + tid) * 4u);
      my_ok = (l == len0 && o == soff0 + tid * len0) ? 1u : 0u;
    }

    // ---------------- phase A: flat over the tile, K written in place of the quality bytes ----------------
    if (!kAdapters) {
      for (uint32_t v16 = tid16; v16 < span; v16 += kCThreads * 16u) {
        const uint32_t a = st + v16;
        const uint4 sv = lds_u128(a), qv = lds_u128(a + buf);
        uint32_t n0, n1, n2, n3, bad = 0;
        uint4 K;
        K.x = key_bytes(sv.x, qv.x, kc, n0, bad);
        K.y = key_bytes(sv.y, qv.y, kc, n1, bad);
        K.z = key_bytes(sv.z, qv.z, kc, n2, bad);
        K.w = key_bytes(sv.w, qv.w, kc, n3, bad);
        if (bad & 0xC0C0C0C0u)  // rare: re-key the offending words to the dummy rows, count them exactly
          n_invalid += fix_bad_unit(sv, qv, K, n0, n1, n2, n3, qsub, lo_al + v16, idx_s, nr, args.a);
        sts_u128(a + buf, K);
      }
    } else {
      // One 16-byte unit per lane.  A warp step covers 31 new units; lane 31 re-reads the unit behind them
      // so that the 7-mer anchors starting at bases 3, 7, 11, 15 of lanes 0..30 find their bases (one
      // shuffle).  The look-ahead lane computes but never stores or reports.
      const uint32_t ccnt_s = ccount_s + cslot * 4u;
      const uint32_t cq_s = cand_s + (it & 1u) * (kCandCap * 8u);
      const uint32_t n16 = span >> 4;
      for (uint32_t u0 = warp * 31u; u0 < n16; u0 += kCW * 31u) {
        const uint32_t u = u0 + lane;
        const uint32_t a = st + min(u, n16) * 16u;  // at most the 16 bytes behind the span are read
        const uint4 sv = lds_u128(a), qv = lds_u128(a + buf);
        uint32_t n0, n1, n2, n3, bad = 0;
        uint4 K;
        K.x = key_bytes(sv.x, qv.x, kc, n0, bad);
        K.y = key_bytes(sv.y, qv.y, kc, n1, bad);
        K.z = key_bytes(sv.z, qv.z, kc, n2, bad);
        K.w = key_bytes(sv.w, qv.w, kc, n3, bad);
        const bool own = lane < 31u && u < n16;
        if (own) {
          if (bad & 0xC0C0C0C0u)
            n_invalid += fix_bad_unit(sv, qv, K, n0, n1, n2, n3, qsub, lo_al + u * 16u, idx_s, nr, args.a);
          sts_u128(a + buf, K);
        }
        // 16 bases -> 32 bits.  The codes come from the base bytes alone (~n): the look-ahead lane may
        // read quality bytes another warp is already replacing with its keys.
        const uint32_t p = pack16(~n0, ~n1, ~n2, ~n3);
        const uint32_t nx = __shfl_down_sync(0xffffffffu, p, 1);
        // anchor j = the 7-mer starting at base 4j+3: row = its bits 13:5 (32-byte rows), bit = its bits 4:0
        const uint32_t e0 = __funnelshift_r(p, nx, 6), e1 = __funnelshift_r(p, nx, 14);
        const uint32_t e2 = __funnelshift_r(p, nx, 22), e3 = __funnelshift_r(p, nx, 30);
        // (e & 0x3FE0) | copy offset in one LOP3; the table base is the load's immediate
        const uint32_t w0 = lds_u32_at(lop3<0xEA>(e0, 0x3FE0u, afilt_copy), afilt_s);
        const uint32_t w1 = lds_u32_at(lop3<0xEA>(e1, 0x3FE0u, afilt_copy), afilt_s);
        const uint32_t w2 = lds_u32_at(lop3<0xEA>(e2, 0x3FE0u, afilt_copy), afilt_s);
        const uint32_t w3 = lds_u32_at(lop3<0xEA>(e3, 0x3FE0u, afilt_copy), afilt_s);
        const uint32_t m0 = __funnelshift_r(w0, 0u, e0), m1 = __funnelshift_r(w1, 0u, e1);
        const uint32_t m2 = __funnelshift_r(w2, 0u, e2), m3 = __funnelshift_r(w3, 0u, e3);
        const bool hit = own && ((m0 | m1 | m2 | m3) & 1u);
        if (__ballot_sync(0xffffffffu, hit)) {
          if (hit) {  // one queue entry per anchor that passed: 25 bases, anchor index, unit
            const uint32_t b0 = m0 & 1u, b1 = m1 & 1u, b2 = m2 & 1u, b3 = m3 & 1u;
            const uint32_t idx = atomicAdd(shared_ptr<uint32_t>(ccnt_s), b0 + b1 + b2 + b3);
            if (idx + 4u <= kCandCap) {  // beyond the cap: the tile is re-scanned below
              const uint32_t hi18 = (nx & 0x3FFFFu) | (u << 20);  // 18 bits of bases, 2 of anchor index, 12 of unit
              uint32_t a = cq_s + idx * 8u;
              if (b0) sts_u64(a, p, hi18);
              a += b0 * 8u;
              if (b1) sts_u64(a, p, hi18 | (1u << 18));
              a += b1 * 8u;
              if (b2) sts_u64(a, p, hi18 | (2u << 18));
              a += b2 * 8u;
              if (b3) sts_u64(a, p, hi18 | (3u << 18));
            }
          }
        }
      }
    }
    const uint32_t uniform = consumer_bar_and(my_ok);  // K bytes and the candidate queue of this tile are complete

    // ---------------- per-read counters, one thread per read (quack.c:219) ----------------
    if (tid < nr) {
      const uint32_t len = lds_u32(idx_s + (kIdxCap + tid) * 4u);
      if (len > len_cap)
        atomicAdd(&args.a.counters[kCntError], 1ull);
      else if (len)
        red_shared_add<0>(lenhist_s + (len - 1u) * 4u, 1u);
    }
    if (kAdapters) {
      if (it) settle_hits((it & 1u) ^ 1u, nr_prev);  // every warp finished the previous tile's confirmations
      if (tid == 0)  // counter of tile it+2: idle since tile it-1
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ccount_s + (cslot == 0 ? 2u : cslot - 1u) * 4u), "r"(0u) : "memory");
      // ---------------- phase A2: confirm queued anchor hits, one (hit, window offset) pair per thread ----------------
      const uint32_t total = lds_u32(ccount_s + cslot * 4u);
      const uint32_t fhit_s = fhit_s0 + (it & 1u) * (kMaxTileReads * 4u);
      const uint32_t ulen = uniform ? len0 : 0u;
      if (total + 4u <= kCandCap) {  // every push found room
        const uint32_t cq_s = cand_s + (it & 1u) * (kCandCap * 8u);
        for (uint32_t i = tid; i < total * 4u; i += kCThreads) {
          const uint2 c = lds_u64(cq_s + (i >> 2) * 8u);
          // anchor j starts at base 4j+3 of its unit; the windows that contain it start at bases 4j .. 4j+3
          confirm_window(c.x, c.y & 0x3FFFFu, c.y >> 20, ((c.y >> 16) & 12u) + (i & 3u), args.ad, exact_s, lo_al, idx_s,
                         nr, ulen, fhit_s);
        }
      } else {  // queue overflow (adapter-dimer-like data): test every window of the tile exactly
        const uint32_t n16 = span >> 4;
        for (uint32_t i = tid; i < n16 * 4u; i += kCThreads) {
          const uint32_t unit = i >> 2, t = i & 3u;
          const uint4 ka = lds_u128(kbuf_s + unit * 16u), kb = lds_u128(kbuf_s + unit * 16u + 16u);
          const uint32_t lo = pack16_keys(ka), hi = pack16_keys(kb);
          for (uint32_t j = 0; j < 4u; j++)
            confirm_window(lo, hi, unit, 4u * j + t, args.ad, exact_s, lo_al, idx_s, nr, ulen, fhit_s);
        }
      }
    }

    // ---------------- phase H: one warp per read, one shared atomic per base ----------------
    {
      const uint32_t k0_s = kbuf_s - lo_al;  // + absolute offset of a read = shared address of its first key byte
#define QB_SHAPES(FN)                                                                      \
  FN(0, 1) FN(0, 2) FN(1, 0) FN(1, 1) FN(1, 2) FN(2, 0) FN(2, 1) FN(2, 2) FN(3, 0)
      if (uniform) {
        if (len0 && len0 <= len_cap) {
          switch (h_shape(len0)) {
#define QB_CASE(nwf, kind)                                                                  \
  case (nwf)*3 + (kind):                                                                    \
    if constexpr (128u * (nwf) + ((kind) ? 1u : 0u) <= 2u * Lh)                              \
      h_uniform<Lh, nwf, kind>(hl, hist_s, k0_s + soff0, len0, nr, warp, lane);              \
    break;
            QB_SHAPES(QB_CASE)
#undef QB_CASE
          }
        }
      } else {
        for (uint32_t r = warp; r < nr; r += kCW) {
          const uint32_t len = lds_u32(idx_s + (kIdxCap + r) * 4u);
          if (len == 0 || len > len_cap) continue;  // reported in the bookkeeping pass above
          const uint32_t kb = k0_s + lds_u32(idx_s + r * 4u);
          switch (h_shape(len)) {
#define QB_CASE(nwf, kind)                                                                  \
  case (nwf)*3 + (kind):                                                                    \
    if constexpr (128u * (nwf) + ((kind) ? 1u : 0u) <= 2u * Lh) h_one<Lh, nwf, kind>(hl, hist_s, kb, len, lane); \
    break;
            QB_SHAPES(QB_CASE)
#undef QB_CASE
          }
        }
      }
#undef QB_SHAPES
    }
    // The stage goes back to the producer.  Its next TMA write must be ordered behind the (generic-proxy)
    // writes of key bytes into the buffer: those all happened before the barrier above, so one proxy
    // fence by the arriving lane covers them.
    __syncwarp();
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(st + o_meta + 24u);
    }

    nr_prev = nr;
    reads_since_flush += nr;
    if (reads_since_flush + RT > 65535u) {  // u16 counters: flush before any bin can wrap
      consumer_bar();
      flush();
      reads_since_flush = 0;
    }
    st += stride;
    if (st == stage_end_s) {
      st = stage0_s;
      ph ^= 1u;
    }
    if (++cslot == 3u) cslot = 0;
  }
  consumer_bar();
  if (kAdapters && it) {
    settle_hits((it & 1u) ^ 1u, nr_prev);
    consumer_bar();
  }
  flush();
  n_invalid = warp_sum(n_invalid);
  if (lane == 0 && n_invalid) atomicAdd(&args.a.counters[kCntInvalidQual], n_invalid);
}

// tile geometry: one phase-A round per tile (every consumer thread takes one 16-byte unit), reads shared
// evenly by the consumer warps in phase H, and as many stages as shared memory allows (the loads of a
// tile need about two tile times to land)
static uint32_t choose_reads_per_tile(uint32_t bytes, uint32_t max_len) {
  if (bytes < max_len + 32u) return 0;
  uint32_t rt_max = (bytes - 32u) / max_len;
  if (rt_max > kMaxTileReads) rt_max = kMaxTileReads;
  // warp instructions per tile and warp: ~130 for phase A and the tile bookkeeping, 4 + 3.6 per 32-base
  // step for every read the warp takes in phase H
  const double per_read = 4.0 + 3.6 * (double)((max_len + 31u) / 32u);
  uint32_t best = rt_max;
  double best_score = 0;
  for (uint32_t rt = rt_max; rt >= 1u && rt + (uint32_t)kCW > rt_max; rt--) {
    const double rounds = (double)((rt + (uint32_t)kCW - 1u) / (uint32_t)kCW);
    const double score = (double)rt / (130.0 + rounds * per_read);
    if (score > best_score * 1.01) {
      best_score = score;
      best = rt;
    }
  }
  return best;
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
  // bytes all consumer threads cover in one phase-A pass (with -a each warp spends a lane on look-ahead);
  // a tile holds as many passes as come closest to 32 KB (the per-tile costs -- barrier, waits, set-up --
  // are what is left to amortise), in as many stages as then fit (two are enough at this size)
  const uint32_t pass_bytes = adapters ? (uint32_t)kCW * 31u * 16u : (uint32_t)kCThreads * 16u;
  uint32_t passes = (32768u + pass_bytes / 2u) / pass_bytes ? (32768u + pass_bytes / 2u) / pass_bytes : 1u;
  uint32_t max_stages = kMaxStages;
  if (const char *e = getenv("QB_TILE_PASSES")) {  // tuning hooks (tools/sweep_tiles.py)
    const int v = atoi(e);
    if (v >= 1 && v <= 16) passes = (uint32_t)v;
  }
  if (const char *e = getenv("QB_STAGES")) {
    const int v = atoi(e);
    if (v >= 2 && v <= (int)kMaxStages) max_stages = (uint32_t)v;
  }
  const uint32_t round_bytes = passes * pass_bytes;
  auto tile_avail = [&](uint32_t stages) -> uint32_t {
    const SmemLayout L0 = smem_layout(half, len_cap, 0, stages, adapters);
    if (L0.total + 128u >= smem_optin) return 0;
    return ((smem_optin - L0.total - 128u) / (2u * stages)) & ~15u;  // L0 already holds every pad
  };
  uint32_t rt = choose_reads_per_tile(round_bytes, batch_max_len);
  uint32_t stages = 0;
  for (uint32_t st = max_stages; st >= 2 && rt; st--)
    if (tile_avail(st) >= align16(rt * batch_max_len + 32u)) {
      stages = st;
      break;
    }
  if (!stages) {  // long reads: whatever two stages hold
    stages = 2;
    rt = choose_reads_per_tile(tile_avail(2), batch_max_len);
    if (!rt) return p;
  }
  if (align16(rt * batch_max_len + 32u) > 65536u) return p;  // queue entries address 4096 16-byte units
  p.stages = stages;
  p.tile_bytes = align16(rt * batch_max_len + 32u);
  p.reads_per_tile = rt;
  p.smem_bytes = smem_layout(half, len_cap, p.tile_bytes, stages, adapters).total;
  p.grid = (uint32_t)sm_count;
  p.ok = 1;
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
