// qb_period.cu -- period kernel (v5) for quack's per-read statistics accumulation
// (reference: the while loop of read_fastq(), quack.c:193-221), for batches whose reads all have ONE length
// l (32..256 bp) and lie back to back -- the shape of untrimmed Illumina data (configs 1-3, 5).
//
// The v3 / v4 kernels are bound by the shared-memory pipe (ncu: 84 % busy once the 1.4-cycle cost of a shared
// atomic is counted): they touch every base three times there (TMA fill, flat pass that writes key bytes,
// per-read pass that re-reads them through two unaligned loads + a byte step).  This kernel touches a base
// twice (TMA fill, ONE aligned 32-bit load) and spends 2 instructions per base on the histogram update:
//   * k reads form a PERIOD of k*l bytes that is a whole number of 32-bit words (k = 1, 2 or a multiple that
//     fills the last warp step: 4 x 150 bp = 150 words = 4.7 warp steps).  A warp walks a period in `steps`
//     steps of 32 aligned words, lane <-> word.  Word 32 s + lane of EVERY period holds the same four
//     positions, so the shared address of each counter column is a per-lane constant kept in a register
//     (steps x 4 of them); there is no per-read work at all, no offsets / lengths are read (the host verified
//     the batch shape), no alignment fix-ups, no tail steps;
//   * key byte K = score << 2 | code as in the v3 kernel; the joint (score, code) x position histogram has
//     256-byte rows at a 64 KiB-aligned shared address, so ONE byte permute builds a counter's address from the
//     key byte and the lane's column register (PRMT + RED per base).  Positions 0..127 live in block 0,
//     128..255 in block 1 (+64 KiB); only 192 of a block's 256 rows exist, the 16 KiB behind them hold the
//     anchor filter.  Which u16 slot of a row a position uses is a TABLE (PArgs::slot) chosen on the host so
//     that the 32 lanes of a warp step fall into different banks as far as the geometry allows;
//   * -a: the 7-mer anchor probe of the v3 / v4 kernels runs in the same loop on the key bytes just computed
//     (codes gathered by one multiply, the neighbour lane's by one shuffle); hits queue the word index and are
//     confirmed per tile against the exact key set from the staged bases, 4 lanes per hit;
//   * autonomous warps with private TMA rings as in v4 (no producer warp, no CTA barrier in steady state).
// Reads that do not fill a tile (< reads_per_tile at the end of a batch) are left to the v4 / v3 kernel.
//
// No tensor cores: the path is an integer histogram.
#include "qb_dev.cuh"

#include <cstdio>
#include <cstdlib>
#include <array>
#include <cstring>
#include <map>
#include <mutex>

namespace qb {

// How a warp stages its tiles.  TMA: one 1-D bulk copy per array (cp.async.bulk, SASS UBLKCP) on an mbarrier, issued
// by one lane -- ~60 warp instructions per tile, the operands go through an elect / R2UR loop each.  cp.async:
// per-lane 16-byte asynchronous copies (SASS LDGSTS), 3 + 3 warp instructions for the 1200 + 1200 bytes of an
// 8 x 150 bp tile plus their predicates, completion through the lane's own copy groups -- ~50 per tile.  Measured
// (profiles/r02/staging_ab.jsonl, 16 M reads of 150 bp): without adapters TMA 0.904 / cp.async 0.889 of the HBM
// roofline, with -a 0.499 / 0.507: each variant takes its winner.  -DQB_PT_TMA=0|1 forces one for both.
#ifdef QB_PT_TMA
template <bool kAd>
constexpr bool kPTmaFor = QB_PT_TMA != 0;
#else
template <bool kAd>
constexpr bool kPTmaFor = !kAd;
#endif
constexpr uint32_t kPHist0 = 0x10000u;             // shared address of histogram block 0
constexpr uint32_t kPBlockStride = 0x10000u;       // block b at kPHist0 + b * 64 KiB (its address has byte 1 == 0)
constexpr uint32_t kPBlockBytes = kHistRows * 256u;  // 192 rows x 256 B = 48 KiB
constexpr uint32_t kPMaxStages = 4;
constexpr uint32_t kPMaxRpt = 64;                  // reads per tile
// offsets inside a warp block (multiples of 16)
constexpr uint32_t kPoBar = 0;                     // kPMaxStages mbarriers
constexpr uint32_t kPoQueue = 32;                  // -a: kPQueue u16 entries: hit bit of the group << 5 | lane
constexpr uint32_t kPQueue = 64;
constexpr uint32_t kPoFhit = kPoQueue + kPQueue * 2u;  // -a: first-hit position of every read of the group's tiles
constexpr uint32_t kPHoleOff = 64;                 // -a: a ring row inside a row of histogram block 1 starts here
constexpr uint32_t kPHoleRowBytes = 256u - kPHoleOff;
// one staged buffer: the tile's bytes from the 16-byte boundary below its first byte to the one above its last
// (tile_bytes is a multiple of 4, not of 16: a tile starts 0, 4, 8 or 12 bytes behind a boundary)
__host__ __device__ inline uint32_t pbuf_bytes(uint32_t tile_bytes) {
  return ((tile_bytes + 15u) & ~15u) + ((tile_bytes & 15u) ? 16u : 0u);
}

struct PArgs {
  const uint8_t *seq, *qual;  // first byte of the first read (16-byte aligned)
  Accum a;
  AdapterSet ad;
  PeriodPlan plan;
  uint32_t n_tiles;
  uint32_t inc_lo, inc_hi;      // 1 and 65536, passed as arguments so that the atomics stay plain ATOMS.ADD (with
                                // a known 1 the compiler emits the warp-aggregating ATOMS.POPC.INC form)
  uint8_t slot[kPeriodMaxLen];  // position -> block << 7 | u32 column of its row (0..63); u16 half = position & 1
};

__device__ __forceinline__ uint32_t p_slot_addr(uint32_t e) {
  return kPHist0 + (e >> 7) * kPBlockStride + ((e & 63u) << 2);
}

// rare path: the 4 bases of a word with an out-of-window quality byte, counted one by one in global memory
// (their key bytes were moved to the dummy row).  b0 = byte of the word inside its period.
static __device__ __noinline__ uint32_t p_exact_word(uint32_t sw, uint32_t qw, uint32_t b0, uint32_t len, const Accum a) {
  uint32_t n_invalid = 0;
  for (uint32_t j = 0; j < 4; j++) {
    const uint32_t p = (b0 + j) % len;
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

__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}

template <bool kAd, int kS, int kPW, bool kOdd, bool kAligned>
__global__ void __launch_bounds__(kPW * 32, 1) period_kernel(const __grid_constant__ PArgs args) {
  constexpr uint32_t kPThreads = kPW * 32;
  constexpr bool kPTma = kPTmaFor<kAd>;
  extern __shared__ __align__(128) uint8_t smem[];
  const PeriodPlan &P = args.plan;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  // (making `warp` provably warp-uniform with a shuffle moves the TMA operands into uniform registers and drops
  // the elect loops around UBLKCP, but measured 1.5-3 % slower: profiles/r01c_v5/notes.md)
  const uint32_t smem_s = smem_u32(smem);
  constexpr uint32_t kFull = 0xffffffffu;

  if (smem_s != P.smem_base) {  // the histogram must sit at its fixed shared address: fail loudly, count nothing
    if (tid == 0) atomicAdd(&args.a.counters[kCntError], 1ull);
    return;
  }
  auto gen = [&](uint32_t shared_addr) -> uint8_t * { return smem + (shared_addr - smem_s); };

  const uint32_t len = P.len, wp = P.wp, pbytes = P.wp * 4u, ppt = P.ppt, rpt = P.reads_per_tile;
  const uint32_t tb = P.tile_bytes, buf = P.buf_bytes, stages = P.stages, nblocks = P.nblocks;
  constexpr bool aligned = kAligned;  // tile_bytes % 16 == 0: every tile starts on a 16-byte boundary (8 x 150 bp)
  const uint32_t last = wp - 32u * (uint32_t)(kS - 1);  // active lanes of the last step (1..32)
  // Lane <-> word map of a period: word 32 s + lane in step s, the lanes of a step side by side.  (Giving a lane kS
  // CONSECUTIVE words would save the -a scan four of its five shuffles per period, but then lanes 75 words apart
  // -- the same positions of two reads -- meet in one atomic: 4 x 150 bp would pay a second wavefront on every
  // update, 6 x 100 bp five more.  Side by side, the 32 lanes of a step always hold 32 different positions.)
  const uint32_t w0 = lane;                                    // this lane's word of step 0
  constexpr uint32_t kWs = 32u;                                // words from one step to the next
  const uint32_t nact = lane < last ? (uint32_t)kS : (uint32_t)kS - 1u;  // steps in which this lane has a word

  // ---- this warp's block ----
  uint32_t wb_s;
  {
    uint32_t w = warp;
    if (w < P.region_n[0])
      wb_s = P.region_s[0] + w * P.wblock;
    else if ((w -= P.region_n[0]) < P.region_n[1])
      wb_s = P.region_s[1] + w * P.wblock;
    else
      wb_s = P.region_s[2] + (w - P.region_n[1]) * P.wblock;
  }
  const uint32_t ring_s = wb_s + P.hdr_bytes;  // stage s: bases at ring_s + 2 s buf, quality bytes + buf
  const uint32_t fhit_s = wb_s + kPoFhit, q_s = wb_s + kPoQueue;
  const uint32_t kmerhist_s = P.kmerhist_s, exact_s = P.exact_s;
  // packed-code ring of this warp: row r (one period) at pring_s + r * prs, the code byte of word i at + 4 + i
  const uint32_t pring_s = P.pring_hole ? P.pring_hole + warp * P.pring_wstride : wb_s + P.pring_off;
  const uint32_t prs = P.prow_stride, rt = P.rt;

  // ---- prologue: zero the histograms, load the adapter tables, init barriers ----
  auto clear_counters = [&]() {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint32_t b = 0; b < nblocks; b++) {
      if (kAd && b == 1u && P.pring_hole) {  // block 1 shares its rows with the packed-code rings: counter columns only
        for (uint32_t i = tid; i < kHistRows * (kPHoleOff / 16u); i += kPThreads) {
          const uint32_t r = i / (kPHoleOff / 16u), c = i - r * (kPHoleOff / 16u);
          *reinterpret_cast<uint4 *>(gen(kPHist0 + kPBlockStride + r * 256u + c * 16u)) = z;
        }
        continue;
      }
      uint4 *h4 = reinterpret_cast<uint4 *>(gen(kPHist0 + b * kPBlockStride));
      for (uint32_t i = tid; i < kPBlockBytes / 16u; i += kPThreads) h4[i] = z;
    }
    if (kAd) {
      uint32_t *kh = reinterpret_cast<uint32_t *>(gen(kmerhist_s));
      for (uint32_t i = tid; i <= len; i += kPThreads) kh[i] = 0;
    }
  };
  clear_counters();
  if (kAd) {
    // the anchor bitmap with the bits of every word reversed: `word << anchor` then leaves the answer in bit 31,
    // from where a funnel shift moves it into the lane's hit mask (2 instructions per probe instead of 4)
    uint32_t *af = reinterpret_cast<uint32_t *>(gen(P.afilt_s));
    for (uint32_t i = tid; i < kAnchorWords * kAnchorCopies; i += kPThreads) af[i] = __brev(args.ad.anchor[i / kAnchorCopies]);
    uint32_t *ex = reinterpret_cast<uint32_t *>(gen(exact_s));
    if (args.ad.exact)
      for (uint32_t i = tid; i < kExactSlots; i += kPThreads) ex[i] = args.ad.exact[i];
    uint32_t *fh = reinterpret_cast<uint32_t *>(gen(fhit_s));
    for (uint32_t i = lane; i < rt * rpt; i += 32u) fh[i] = kNoHit;
  }
  // the slot table into shared memory: indexing the kernel arguments with a per-thread position is a divergent
  // constant load (one transaction per lane): 20 of them per thread cost ~10 us per launch
  if (tid < kPeriodMaxLen / 4u) reinterpret_cast<uint32_t *>(gen(P.slot_s))[tid] = reinterpret_cast<const uint32_t *>(args.slot)[tid];
  if (lane == 0) {
    for (uint32_t s = 0; s < stages; s++) mbar_init(reinterpret_cast<uint64_t *>(gen(wb_s + kPoBar + 8u * s)), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t n_tiles = args.n_tiles;
  if (blockIdx.x == 0 && tid == 0) {  // quack.c:219-220: every read has length l
    const unsigned long long n = (unsigned long long)n_tiles * rpt;
    atomicAdd(&args.a.counters[kCntReads], n);
    atomicAdd(&args.a.rows[(size_t)(len - 1u) * kRow + kColLength], n);
  }
  __syncthreads();

  const uint32_t G = gridDim.x * kPW;
  const uint32_t g0 = blockIdx.x * kPW;
  const uint32_t iters = n_tiles > g0 ? (n_tiles - g0 + G - 1u) / G : 0u;  // the same for every warp of the CTA
  uint32_t epoch = 65535u / ((uint32_t)kPW * rpt);  // iterations between two flushes of the u16 counters
  if (epoch == 0) epoch = 1;

  const KeyConsts kc(P.qbase);
  const uint32_t inc_lo = args.inc_lo, inc_hi = args.inc_hi;
  // this lane's counter columns: word 32 s + lane of a period holds positions (4 (32 s + lane) + j) mod l
  uint32_t col[kS][4];
#pragma unroll
  for (int s = 0; s < kS; s++)
#pragma unroll
    for (int j = 0; j < 4; j++) col[s][j] = p_slot_addr(lds_u8(P.slot_s + (4u * (w0 + kWs * s) + j) % len));
  // Even l: the j-th byte of a word always holds positions of the parity of j, the increments are the two
  // constants.  Odd l: the parity also depends on the read inside the period -> one increment per column.
  uint32_t inc[kOdd ? kS : 1][4];
  if (kOdd) {
#pragma unroll
    for (int s = 0; s < (kOdd ? kS : 1); s++)
#pragma unroll
      for (int j = 0; j < 4; j++) inc[s][j] = (((4u * (w0 + kWs * s) + j) % len) & 1u) ? inc_hi : inc_lo;
  }
  const uint32_t afilt_or = P.afilt_s | ((lane >> 2) * 4u);  // this lane's copy of the anchor map (8 copies, 32-byte rows)
  const uint32_t nxt = (lane + 1u) & 31u;
  const uint32_t len_magic = 0xFFFFFFFFu / len + 1u;  // floor(b / len) = umulhi(b, len_magic) for b < 2^24
  // hit mask of a tile: one bit per (period, step), pushed in from the right -> (pp, s) sits at bit
  // (ppt - 1 - pp) kS + (kS - 1 - s).  Lanes without a word in the last step never report a hit there.
  uint32_t hm_valid = (ppt * (uint32_t)kS >= 32u) ? kFull : (1u << (ppt * (uint32_t)kS)) - 1u;
  for (uint32_t pp = 0; pp < ppt; pp++)
    for (uint32_t s = nact; s < (uint32_t)kS; s++) hm_valid &= ~(1u << (pp * (uint32_t)kS + ((uint32_t)kS - 1u - s)));
  long long n_invalid = 0;

  auto flush = [&]() {  // all warps are behind a barrier
    for (uint32_t pos = tid; pos < len; pos += kPThreads) {
      const uint32_t base = p_slot_addr(lds_u8(P.slot_s + pos));
      const uint32_t sh = (pos & 1u) * 16u;
      unsigned long long *row = args.a.rows + (size_t)pos * kRow;
      uint32_t cc[4] = {0, 0, 0, 0};
      for (uint32_t sp = 0; sp < kScoreBins; sp++) {
        uint32_t tot = 0;
#pragma unroll
        for (uint32_t c = 0; c < 4u; c++) {  // row = score << 2 | code
          const uint32_t v = (lds_u32(base + ((sp << 2 | c) << 8)) >> sh) & 0xFFFFu;
          cc[c] += v;
          tot += v;
        }
        if (tot) {
          const int sc = (int)(sp + P.qbase) - 33;
          if (sc >= 0 && sc < 91)
            atomicAdd(&row[sc], (unsigned long long)tot);
          else
            n_invalid += tot;
        }
      }
#pragma unroll
      for (uint32_t c = 0; c < 4u; c++)
        if (cc[c]) atomicAdd(&row[kColContent + c], (unsigned long long)cc[c]);
      if (kAd) {
        const uint32_t kcnt = lds_u32(kmerhist_s + pos * 4u);
        if (kcnt) atomicAdd(&row[kColKmer], (unsigned long long)kcnt);
      }
    }
  };

  // Start the copies of tile t into stage s (have = false: past the end of the batch -- with cp.async still an
  // (empty) copy group, so that "all but the newest stages - 1 groups are complete" always means "tile t arrived").
  // The whole warp calls this once it is done with the stage.
  auto issue = [&](uint32_t s, uint32_t t, bool have) {
    __syncwarp();  // every lane is done with the buffer's old contents
    if (kPTma) {
      if (have && lane == 0) {
        const uint32_t bar_s = wb_s + kPoBar + 8u * s;
        const uint32_t dst = ring_s + 2u * s * buf;
        const size_t off = (size_t)t * tb;
        // the TMA (async proxy) write must be ordered behind the generic-proxy accesses to the buffer
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (aligned) {
          mbar_arrive_expect_tx_s(bar_s, 2u * tb);
          bulk_g2s_s(dst, args.seq + off, tb, bar_s);
          bulk_g2s_s(dst + buf, args.qual + off, tb, bar_s);
        } else {
          const uint32_t so = (uint32_t)(off & 15u);       // the copy starts at the 16-byte boundary below the tile
          const uint32_t bytes = (so + tb + 15u) & ~15u;   // (the batch buffers are readable 64 bytes past their end)
          mbar_arrive_expect_tx_s(bar_s, 2u * bytes);
          bulk_g2s_s(dst, args.seq + (off - so), bytes, bar_s);
          bulk_g2s_s(dst + buf, args.qual + (off - so), bytes, bar_s);
        }
      }
    } else {
      if (have) {
        size_t off = (size_t)t * tb;
        uint32_t bytes = tb;
        if (!aligned) {  // from the 16-byte boundary below the tile (the batch buffers are readable 64 bytes past their end)
          const uint32_t so = (uint32_t)(off & 15u);
          off -= so;
          bytes = (so + tb + 15u) & ~15u;
        }
        const uint32_t lo = lane * 16u;
        const uint32_t dst = ring_s + 2u * s * buf + lo;
        const uint8_t *gs = args.seq + off + lo, *gq = args.qual + off + lo;
        // round i copies bytes [512 i, 512 i + 512) of both arrays, 16 per lane: one predicate per round, the
        // addresses are the two pointers above plus immediates
        cp_async16_pair<0>(lo < bytes, dst, dst + buf, gs, gq);
        cp_async16_pair<512>(lo + 512u < bytes, dst, dst + buf, gs, gq);
        cp_async16_pair<1024>(lo + 1024u < bytes, dst, dst + buf, gs, gq);
        if (bytes > 1536u) {  // (warp-uniform)
          cp_async16_pair<1536>(lo + 1536u < bytes, dst, dst + buf, gs, gq);
          cp_async16_pair<2048>(lo + 2048u < bytes, dst, dst + buf, gs, gq);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  // tile in stage s arrived: the mbarrier's phase flipped / this lane's copy group is complete, then the warp syncs
  auto arrived = [&](uint32_t s, uint32_t phase) {
    if (kPTma) {
      mbar_wait(wb_s + kPoBar + 8u * s, phase);
    } else {
      if (stages == 2u)
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      else if (stages == 3u)
        asm volatile("cp.async.wait_group 2;" ::: "memory");
      else
        asm volatile("cp.async.wait_group 3;" ::: "memory");
      __syncwarp();
    }
  };

  // ---- -a: anchor hits are confirmed per GROUP of up to rt tiles (quack.c:210-217) ----
  // A lane keeps the hit bits of the group's tiles in one register (hm_acc: the newest tile in the low bits), the
  // 2-bit codes of the tiles stay in the packed-code ring.  When about 32 hits have come together (or the ring is
  // full) the hits are expanded into a queue -- entry = bit index << 5 | lane, placed with a warp prefix sum, no
  // atomics -- and confirmed 32 at a time, one lane per entry: the 16 bases of words w - 1 .. w + 2 come as ONE
  // 32-bit value out of the ring row (two aligned loads and a funnel shift), the four 10-mers that contain the
  // anchor are tested against the exact key set.  A window whose 10 bases lie inside one read lowers that read's
  // first-hit position; a hit that ends on the last base of its read is dropped (it can only be the first hit if
  // there is no other, and then the reference counts nothing).  Bytes of a ring row outside the period are garbage:
  // a window that reaches them does not lie inside one read (a period is whole reads).
  const uint32_t hbits = ppt * (uint32_t)kS;                 // hit bits per tile
  const uint32_t hbits_magic = 0xFFFFFFFFu / hbits + 1u;     // floor(b / hbits) = umulhi(b, magic), b < 2^16
  uint32_t hm_acc = 0;                                       // this lane's hit bits of the group
  uint32_t grp_n = 0, grp_cnt = 0;                           // warp-uniform: tiles in the group, hits in the group
  auto confirm_pass = [&](uint32_t e0, uint32_t n) {         // queue entries e0 .. e0 + n - 1, n <= 32
    if (lane < n) {
      const uint32_t ent = lds_u16(q_s + 2u * (e0 + lane));
      const uint32_t b = ent >> 5, ln = ent & 31u;
      const uint32_t jr = __umulhi(b, hbits_magic), bt = b - jr * hbits;  // tile (newest = 0), bit inside the tile
      const uint32_t pr = bt / (uint32_t)kS, s = (uint32_t)(kS - 1) - (bt - pr * (uint32_t)kS);
      const uint32_t row = (grp_n - 1u - jr) * ppt + (ppt - 1u - pr), w = 32u * s + ln;
      const uint32_t a = pring_s + row * prs + 3u + w;  // code byte of word w - 1
      const uint32_t a4 = a & ~3u;
      const uint32_t ctx = __funnelshift_r(lds_u32(a4), lds_u32(a4 + 4u), (a & 3u) * 8u);  // bases 4 w - 4 .. 4 w + 11
      const uint32_t b0 = w * 4u;
      const uint32_t n_in = __umulhi(b0, len_magic), p0 = b0 - n_in * len;  // the read inside the period, the position
      uint32_t *fh = shared_ptr<uint32_t>(fhit_s) + (row * P.k + n_in);
#pragma unroll
      for (uint32_t o = 1; o <= 4u; o++) {  // window o starts oo = 4 - o bases before the anchor
        const uint32_t oo = 4u - o;
        if (p0 >= oo && p0 - oo + 10u < len) {
          const uint32_t key = (ctx >> (2u * o)) & 0xFFFFFu;
          bool member;
          if (args.ad.exact)
            member = lds_u32(exact_s + exact_off1(key)) == key || lds_u32(exact_s + exact_off2(key)) == key;
          else
            member = (args.ad.bitmap[key >> 5] >> (key & 31u)) & 1u;
          if (member) atomicMin(fh, p0 - oo + 9u);
        }
      }
    }
  };
  auto process_group = [&]() {
    uint32_t m = hm_acc;
    while (__any_sync(kFull, m != 0u)) {
      // where this lane's entries go: exclusive prefix sum of the hit counts over the lanes
      const uint32_t c = (uint32_t)__popc(m);
      uint32_t incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, incl, d);
        if (lane >= (uint32_t)d) incl += v;
      }
      const uint32_t total = __shfl_sync(kFull, incl, 31);
      uint32_t off = incl - c;
      while (m && off < kPQueue) {  // (a lane whose entries do not fit keeps their bits for the next round)
        const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
        m &= m - 1u;
        sts_u16(q_s + 2u * off, b << 5 | lane);
        off++;
      }
      __syncwarp();
      const uint32_t n = min(total, kPQueue);
      for (uint32_t e0 = 0; e0 < n; e0 += 32u) confirm_pass(e0, min(32u, n - e0));
      __syncwarp();
    }
    // first hits of the group's reads: kmer_count[p + 1]++ (quack.c:215-216)
    for (uint32_t h = lane; h < grp_n * rpt; h += 32u) {
      const uint32_t fa = fhit_s + 4u * h;
      const uint32_t f = lds_u32(fa);
      if (f != kNoHit) {
        red_shared_add<0u>(kmerhist_s + (f + 1u) * 4u, 1u);
        sts_u32(fa, kNoHit);
      }
    }
    hm_acc = 0, grp_n = 0, grp_cnt = 0;
    __syncwarp();
  };

  const uint32_t g = g0 + warp;
  for (uint32_t s = 0; s + 1u < stages; s++) issue(s, g + s * G, g + s * G < n_tiles);

  uint32_t to_flush = epoch;
  uint32_t tile = g;
  uint32_t st = 0, phase = 0;  // stage of this iteration, its mbarrier phase parity
  for (uint32_t it = 0; it < iters; ++it, tile += G) {
    if (tile < n_tiles) {
      {  // keep stages - 1 tiles in flight: the stage freed by the previous iteration takes tile + (stages - 1) G
        const uint32_t t2 = tile + (stages - 1u) * G;
        const uint32_t s2 = st == 0 ? stages - 1u : st - 1u;
        issue(s2, t2, t2 < n_tiles);
      }
      arrived(st, phase);
      const uint32_t seq_s = ring_s + 2u * st * buf + (aligned ? 0u : (uint32_t)(((size_t)tile * tb) & 15u));
      uint32_t hm = 0;  // -a: this lane's anchor hits of the tile, one bit per (period, step)
      uint32_t pa = pring_s + grp_n * ppt * prs + 4u + w0;  // -a: this lane's first code byte in the period's ring row
      uint32_t a0 = seq_s + w0 * 4u;                          // this lane's first word of the period

      for (uint32_t pp = 0; pp < ppt; pp++, a0 += pbytes, pa += prs) {
        uint32_t sw[kS], qw[kS], K[kS], nc[kS], cd[kS];
        uint32_t bad = 0;
#pragma unroll
        for (int s = 0; s < kS; s++) {
          const bool act = (uint32_t)s < nact;
          sw[s] = 0x41414141u;  // lanes without a word: 'A' with the lowest score, not counted
          qw[s] = kc.qsub;
          if (act) {
            sw[s] = lds_u32(a0 + 4u * kWs * s);
            qw[s] = lds_u32(a0 + buf + 4u * kWs * s);
          }
        }
#pragma unroll
        for (int s = 0; s < kS; s++) K[s] = kAd ? key_bytes_c(sw[s], qw[s], kc, nc[s], bad, cd[s]) : key_bytes(sw[s], qw[s], kc, nc[s], bad);
        if (bad & 0xC0C0C0C0u) {  // a quality byte outside the counted window: exact path for those words
#pragma unroll
          for (int s = 0; s < kS; s++)
            if (word_bad(qw[s], kc.qsub)) {
              K[s] = key_bytes_bad(nc[s]);
              n_invalid += p_exact_word(sw[s], qw[s], 4u * (w0 + kWs * s), len, args.a);
            }
        }
        if (kAd) {
          uint32_t gc[kS];
#pragma unroll
          for (int s = 0; s < kS; s++) gc[s] = cd[s] * 0x01041040u;  // top byte: the word's 4 codes
          uint32_t nx[kS];  // codes of the word that follows word s
          {
            uint32_t r[kS];
#pragma unroll
            for (int s = 0; s < kS; s++) r[s] = __shfl_sync(kFull, gc[s], nxt);
#pragma unroll
            for (int s = 0; s < kS; s++) nx[s] = lane < 31u ? r[s] : (s + 1 < kS ? r[s + 1] : 0u);
          }
#pragma unroll
          for (int s = 0; s < kS; s++) {
            const bool act = (uint32_t)s < nact;
            const uint32_t an = __byte_perm(gc[s], nx[s], 0x7773);  // bits 13:0 = the 7-mer starting at this word
            if (act) sts_u8(pa + kWs * s, an);                      // low byte: this word's codes, kept for the confirmation
            const uint32_t fw = lds_u32((an & 0x3FE0u) | afilt_or);
            // the filter words are bit-reversed: bit (an & 31) arrives in bit 31 and is shifted into the hit mask
            hm = __funnelshift_l(__funnelshift_l(0u, fw, an), hm, 1);
          }
        }
#pragma unroll
        for (int s = 0; s < kS; s++) {
          const bool act = (uint32_t)s < nact;
          if (act) {
            red_shared_add<0u>(__byte_perm(K[s], col[s][0], 0x7604), kOdd ? inc[kOdd ? s : 0][0] : inc_lo);
            red_shared_add<0u>(__byte_perm(K[s], col[s][1], 0x7614), kOdd ? inc[kOdd ? s : 0][1] : inc_hi);
            red_shared_add<0u>(__byte_perm(K[s], col[s][2], 0x7624), kOdd ? inc[kOdd ? s : 0][2] : inc_lo);
            red_shared_add<0u>(__byte_perm(K[s], col[s][3], 0x7634), kOdd ? inc[kOdd ? s : 0][3] : inc_hi);
          }
        }
      }

      if (kAd) {
        // the tile joins the group; the group is confirmed once it holds about a warp's worth of hits or is full
        hm &= hm_valid;
        hm_acc = (hbits >= 32u ? 0u : hm_acc << hbits) | hm;
        grp_cnt += __reduce_add_sync(kFull, (uint32_t)__popc(hm));
        if (++grp_n == rt || grp_cnt >= 24u) process_group();
      }
      if (++st == stages) st = 0, phase ^= 1u;
    }
    if (--to_flush == 0u && it + 1u < iters) {  // u16 counters: flush before any bin can wrap
      to_flush = epoch;
      if (kAd && grp_n) process_group();
      __syncthreads();
      flush();
      __syncthreads();
      clear_counters();
      __syncthreads();
    }
  }
  if (kAd && grp_n) process_group();
  __syncthreads();
  flush();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_invalid += __shfl_xor_sync(kFull, n_invalid, o);
  if (lane == 0 && n_invalid) atomicAdd(&args.a.counters[kCntInvalidQual], (unsigned long long)n_invalid);
}

// ------------------------------------------------------------------------------------------
// plan: period / tile geometry and shared-memory map
// ------------------------------------------------------------------------------------------

static void period_slots(const PeriodPlan &p, uint8_t *slot);

// -a: can the packed-code rings live in the rows of histogram block 1?  Only the first kPHoleOff bytes of a row may
// hold counters then: the counter layout (period_slots) puts the pairs that overflow block 0 into the lowest banks.
static bool block1_leaves_holes(uint32_t l, uint32_t k, uint32_t wp, uint32_t steps) {
  PeriodPlan q;
  memset(&q, 0, sizeof q);
  q.len = l, q.k = k, q.wp = wp, q.steps = steps, q.nblocks = 2;
  uint8_t slot[kPeriodMaxLen];
  period_slots(q, slot);
  for (uint32_t pos = 0; pos < l; pos++)
    if ((slot[pos] >> 7) && (slot[pos] & 63u) >= kPHoleOff / 4u) return false;
  return true;
}

static PeriodPlan period_plan_w(uint32_t l, uint32_t first_offset, int adapters, int sm_count, uint32_t smem_optin,
                                uint32_t smem_reserved, uint32_t qbase, uint32_t warps, bool full_tiles) {
  PeriodPlan p;
  memset(&p, 0, sizeof p);
  if (l < 32u || l > kPeriodMaxLen || (first_offset & 15u)) return p;
  if ((l & 1u) && warps > 20u) return p;  // odd lengths keep one increment per column in registers: 20 warps at most
  p.len = l;
  p.qbase = qbase;
  p.nblocks = l > 128u ? 2u : 1u;
  p.smem_base = smem_reserved;
  const uint32_t end = smem_reserved + smem_optin;
  // free address ranges around the histogram blocks (only 192 of a block's 256 rows exist)
  struct Gap {
    uint32_t a, b;
  } gap[3];
  gap[0] = {smem_reserved, kPHist0};
  if (p.nblocks == 2u) {
    gap[1] = {kPHist0 + kPBlockBytes, kPHist0 + kPBlockStride};
    gap[2] = {kPHist0 + kPBlockStride + kPBlockBytes, end};
  } else {
    gap[1] = {kPHist0 + kPBlockBytes, kPHist0 + kPBlockBytes + kAnchorSmemBytes};
    gap[2] = {kPHist0 + kPBlockBytes + kAnchorSmemBytes, end};
  }
  for (int g = 0; g < 3; g++)
    if (gap[g].a > gap[g].b || gap[g].b > end) return p;
  auto take = [&](int g, uint32_t bytes) -> uint32_t {  // 0: does not fit
    bytes = (bytes + 15u) & ~15u;
    if (gap[g].b - gap[g].a < bytes) return 0;
    const uint32_t at = gap[g].a;
    gap[g].a += bytes;
    return at;
  };
  if (adapters) {
    if (!(p.afilt_s = take(1, kAnchorSmemBytes)) || (p.afilt_s & (kAnchorSmemBytes - 1u))) return p;  // 16 KiB-aligned
    if (!(p.exact_s = take(2, kExactSlots * 4u))) return p;
    if (!(p.kmerhist_s = take(2, (l + 1u) * 4u))) return p;
  }
  if (!(p.slot_s = take(0, kPeriodMaxLen))) return p;
  uint32_t want = 3, target = 1024u;
  if (const char *e = getenv("QB_PT_STAGES")) want = (uint32_t)atoi(e);  // tuning hooks
  if (const char *e = getenv("QB_PT_BYTES")) target = (uint32_t)atoi(e);
  if (want < 2u) want = 2u;
  if (want > kPMaxStages) want = kPMaxStages;

  // Reads per period: a multiple of k0 (so that the period is a whole number of words) with 3..5 warp steps;
  // candidates in order of how well they fill the last step.  Periods per tile:
  // reads per tile a multiple of 4 (the rest of the batch goes to a kernel that loads 4 offsets at a time),
  // about `target` bytes.  The first candidate whose kPW warp blocks fit the gaps wins.
  const uint32_t k0 = (l & 1u) ? 4u : ((l & 3u) ? 2u : 1u);
  bool used[64] = {false};
  for (;;) {
    uint32_t bk = 0;
    double best = 0;
    for (uint32_t k = k0; k * l / 4u <= 160u && k < 64u; k += k0) {
      const uint32_t wp = k * l / 4u, steps = (wp + 31u) / 32u;
      if (used[k] || steps < 3u || steps > 5u) continue;
      const double eff = (double)wp / (32.0 * steps);
      if (eff > best + 1e-9) best = eff, bk = k;
    }
    if (!bk) return p;
    used[bk] = true;
    const uint32_t wp = bk * l / 4u, pb = wp * 4u;
    uint32_t ppt0 = 1;
    while ((ppt0 * bk) & 3u) ppt0 *= 2u;
    if (ppt0 * bk > kPMaxRpt) continue;
    uint32_t ppt = ppt0;
    while (ppt * pb < target && (ppt + ppt0) * bk <= kPMaxRpt) ppt += ppt0;
    for (; ppt >= ppt0 && !p.ok; ppt -= ppt0) {
      if (ppt * wp > 65535u || ppt * ((wp + 31u) / 32u) > 32u || pbuf_bytes(ppt * pb) > 2560u) continue;
      if (full_tiles && ppt * pb < target && (ppt + ppt0) * bk <= kPMaxRpt) break;  // smaller tiles: only in the second round  // u16 queue entries, one hit bit per (period, step)
      for (uint32_t stages = want; stages >= 2u && !p.ok; stages--) {
        // -a: packed-code ring of rt tiles (one group of tiles whose anchor hits are confirmed together: as many as
        // one register holds hit bits for, 4 at most, fewer if memory is short), in the unused columns of histogram
        // block 1 when they are there (rows of 256 bytes, kPHoleOff of them counters), else inside the warp block
        const uint32_t hbits = ppt * ((wp + 31u) / 32u);
        uint32_t rt_max = adapters ? 32u / hbits : 1u;
        if (rt_max > 4u) rt_max = 4u;
        if (rt_max < 1u) rt_max = 1u;
        for (uint32_t rt = rt_max; rt >= 1u && !p.ok; rt--) {
          for (int holes = (adapters && p.nblocks == 2u) ? 1 : 0; holes >= 0 && !p.ok; holes--) {
            uint32_t hdr = kPoQueue, prs = 0;
            if (adapters) {
              hdr = kPoFhit + ((rt * ppt * bk * 4u + 15u) & ~15u);
              if (holes) {
                prs = 256u;
                if (4u + wp + 8u > kPHoleRowBytes || rt * ppt > kHistRows / warps || l > 128u + 2u * (kPHoleOff / 4u) ||
                    !block1_leaves_holes(l, bk, wp, (wp + 31u) / 32u))
                  continue;
              } else {
                prs = (4u + wp + 8u + 15u) & ~15u;
                p.pring_off = hdr;
                hdr += rt * ppt * prs;
              }
            }
            const uint32_t wblock = hdr + stages * 2u * pbuf_bytes(ppt * pb);
            uint32_t fit = 0;
            for (int g = 0; g < 3; g++) fit += (gap[g].b - gap[g].a) / wblock;
            if (fit < warps) continue;
            p.k = bk, p.wp = wp, p.steps = (wp + 31u) / 32u;
            p.ppt = ppt, p.tile_bytes = ppt * pb, p.reads_per_tile = ppt * bk;
            p.stages = stages, p.wblock = wblock, p.buf_bytes = pbuf_bytes(ppt * pb);
            p.hdr_bytes = hdr, p.rt = rt, p.prow_stride = prs;
            p.pring_hole = holes ? kPHist0 + kPBlockStride + kPHoleOff : 0u;
            p.pring_wstride = holes ? (kHistRows / warps) * 256u : 0u;
            if (holes) p.pring_off = 0;
            p.ok = 1;
          }
        }
      }
    }
    if (p.ok) break;
  }
  p.warps = warps;
  uint32_t left = warps;
  for (int g = 0; g < 3; g++) {
    uint32_t n = (gap[g].b - gap[g].a) / p.wblock;
    if (n > left) n = left;
    p.region_s[g] = gap[g].a;
    p.region_n[g] = n;
    left -= n;
  }
  p.smem_bytes = smem_optin;
  p.grid = (uint32_t)sm_count;
  return p;
}

// More resident warps hide more latency (ncu: the kernel is bound by instruction issue with 4 warps per
// scheduler): 24 warps where their blocks fit, else 20, else 16 (24 warps leave 80 registers per thread; every
// instantiation meets that without spills).
PeriodPlan period_plan(uint32_t l, uint32_t first_offset, int adapters, int sm_count, uint32_t smem_optin,
                       uint32_t smem_reserved, uint32_t qbase) {
  uint32_t forced = 0;
  if (const char *e = getenv("QB_PT_WARPS")) forced = (uint32_t)atoi(e);  // tuning hook
  PeriodPlan p;
  memset(&p, 0, sizeof p);
  // tile size first (the per-tile costs -- TMA issue, barrier, the -a confirmation pass -- are fixed), then warps
  for (int full = 1; full >= 0 && !p.ok; full--)
    for (uint32_t w : {24u, 20u, 16u}) {
      if (forced && w != forced) continue;
      p = period_plan_w(l, first_offset, adapters, sm_count, smem_optin, smem_reserved, qbase, w, full != 0);
      if (p.ok) break;
    }
  return p;
}

// ------------------------------------------------------------------------------------------
// Position -> u16 slot of a histogram row.
// The half-word is position & 1; positions 2 e and 2 e + 1 share a u32 column.  (For even l the j-th byte of a
// word always holds positions of the parity of j, so the increments are two constants; for odd l the kernel
// keeps one increment per column register.)  The column of a pair is free, and it decides the BANK its 188
// counters live in.  The 32 lanes of warp step
// (s, j) update positions (4 (32 s + lane) + j) mod l: a conflict-free step needs 32 different banks.  All
// steps cannot be conflict-free (a bank holds at most 2 * nblocks even positions, and a position occurs in k
// steps), so the solver looks for the smallest set of steps to give up and colours the conflict graph of the
// others (DSATUR with a capacity per bank).  For 4 x 150 bp: 3 of 10 steps cost two wavefronts instead of
// all 10 with the natural layout (column = position / 4).
// ------------------------------------------------------------------------------------------
namespace {

struct SlotSolver {
  uint32_t E = 0, cap = 0, nsets = 0;
  uint8_t set[20][32];
  uint32_t set_n[20];
  uint64_t rng = 0x9E3779B97F4A7C15ull;
  uint32_t rnd() {
    rng ^= rng << 13, rng ^= rng >> 7, rng ^= rng << 17;
    return (uint32_t)(rng >> 32);
  }
  // colours E positions with 32 banks so that the sets in `hard` have no repeated bank
  bool colour(uint32_t hard, uint8_t *bank) {
    static thread_local uint8_t adj[128][128];
    for (uint32_t a = 0; a < E; a++) memset(adj[a], 0, E);
    for (uint32_t si = 0; si < nsets; si++)
      if (hard >> si & 1u)
        for (uint32_t a = 0; a < set_n[si]; a++)
          for (uint32_t b = 0; b < set_n[si]; b++)
            if (set[si][a] != set[si][b]) adj[set[si][a]][set[si][b]] = 1;
    uint32_t cnt[32] = {0};
    uint32_t usedmask[128];
    bool done[128];
    for (uint32_t a = 0; a < E; a++) usedmask[a] = 0, done[a] = false;
    for (uint32_t step = 0; step < E; step++) {
      int bp = -1;
      uint64_t bkey = 0;
      for (uint32_t a = 0; a < E; a++) {
        if (done[a]) continue;
        const uint64_t key = ((uint64_t)__builtin_popcount(usedmask[a]) << 40) | (uint64_t)(rnd() & 0xFFFFFFu);
        if (bp < 0 || key > bkey) bp = (int)a, bkey = key;
      }
      int bc = -1;
      uint32_t bcnt = 0;
      for (uint32_t c = 0; c < 32; c++) {
        if ((usedmask[bp] >> c & 1u) || cnt[c] >= cap) continue;
        if (bc < 0 || cnt[c] < bcnt || (cnt[c] == bcnt && (rnd() & 1u))) bc = (int)c, bcnt = cnt[c];
      }
      if (bc < 0) return false;
      bank[bp] = (uint8_t)bc;
      cnt[bc]++;
      done[bp] = true;
      for (uint32_t a = 0; a < E; a++)
        if (adj[bp][a]) usedmask[a] |= 1u << bc;
    }
    return true;
  }
  uint32_t cost(const uint8_t *bank) const {
    uint32_t tot = 0;
    for (uint32_t si = 0; si < nsets; si++) {
      uint32_t c[32] = {0}, m = 0;
      for (uint32_t a = 0; a < set_n[si]; a++) m = ++c[bank[set[si][a]]] > m ? c[bank[set[si][a]]] : m;
      tot += m;
    }
    return tot;
  }
};

struct SlotCache {  // every table solved so far, by (read length, reads per period)
  std::mutex mu;
  std::map<uint32_t, std::array<uint8_t, kPeriodMaxLen>> solved;
};

}  // namespace

static void period_slots(const PeriodPlan &p, uint8_t *slot) {
  const uint32_t l = p.len, E = (l + 1u) / 2u, cap = 2u * p.nblocks;  // E pairs of positions (one u32 column each)
  // natural layout (also the fallback): positions 4 c .. 4 c + 3 of a block in bank c
  for (uint32_t pos = 0; pos < kPeriodMaxLen; pos++) {
    const uint32_t blk = pos >> 7, q = pos & 127u;
    slot[pos] = (uint8_t)(blk << 7 | ((q >> 2) + 32u * ((q >> 1) & 1u)));
  }
  if (getenv("QB_PT_NATURAL")) return;  // tuning hook
  static SlotCache *cache = new SlotCache();
  std::lock_guard<std::mutex> lk(cache->mu);
  const uint32_t cache_key = l << 8 | p.k;
  if (auto it = cache->solved.find(cache_key); it != cache->solved.end()) {
    memcpy(slot, it->second.data(), kPeriodMaxLen);
    return;
  }
  SlotSolver sv;
  sv.E = E, sv.cap = cap;
  for (uint32_t s = 0; s < p.steps; s++)
    for (uint32_t j = 0; j < 4u; j += (l & 1u) ? 1u : 2u) {  // even l: byte j + 1 touches the same columns as byte j
      uint32_t n = 0;
      for (uint32_t i = 0; i < 32u && 32u * s + i < p.wp; i++) sv.set[sv.nsets][n++] = (uint8_t)(((4u * (32u * s + i) + j) % l) / 2u);
      sv.set_n[sv.nsets++] = n;
    }
  uint8_t bank[128], best_bank[128];
  for (uint32_t e = 0; e < E; e++) best_bank[e] = (uint8_t)(slot[2u * e] & 31u);
  uint32_t best = sv.cost(best_bank);
  uint32_t lower = 0;
  for (uint32_t si = 0; si < sv.nsets; si++) {  // a position that occurs twice in a step costs a wavefront anyway
    uint32_t c[128] = {0}, m = 1;
    for (uint32_t a = 0; a < sv.set_n[si]; a++) m = ++c[sv.set[si][a]] > m ? c[sv.set[si][a]] : m;
    lower += m;
  }
  // subsets of steps to give up, smallest first (Gosper's hack walks the masks of one size); the search is
  // bounded: at most kBudget colourings (a few milliseconds), then the best layout found so far stands
  const uint32_t all = (1u << sv.nsets) - 1u;
  constexpr int kBudget = 400;
  int calls = 0;
  for (uint32_t nsac = 0; nsac <= 4u && nsac <= sv.nsets && best > lower + nsac && calls < kBudget; nsac++) {
    uint32_t sac = nsac ? (1u << nsac) - 1u : 0u;
    for (;;) {
      for (int t = 0; t < 2 && calls < kBudget; t++) {
        calls++;
        if (sv.colour(all & ~sac, bank)) {
          const uint32_t c = sv.cost(bank);
          if (c < best) best = c, memcpy(best_bank, bank, E);
          break;
        }
      }
      if (nsac == 0 || best <= lower + nsac || calls >= kBudget) break;
      const uint32_t c = sac & (0u - sac), r = sac + c;  // next mask with the same number of bits
      sac = (((r ^ sac) >> 2) / c) | r;
      if (sac > all) break;
    }
  }
  // Bank labels are free (a step is conflict-free iff its lanes use DIFFERENT banks): the fullest banks get the
  // lowest labels, so the pairs that overflow into block 1 sit in its first columns and the rest of every
  // block-1 row stays free (the -a kernel keeps its packed-code rings there, see block1_leaves_holes).
  {
    uint32_t load[32] = {0}, order[32], label[32];
    for (uint32_t e = 0; e < E; e++) load[best_bank[e] & 31u]++;
    for (uint32_t b = 0; b < 32u; b++) order[b] = b;
    for (uint32_t i = 1; i < 32u; i++)  // insertion sort by load, descending, stable
      for (uint32_t j = i; j > 0 && load[order[j]] > load[order[j - 1]]; j--) {
        const uint32_t t = order[j];
        order[j] = order[j - 1], order[j - 1] = t;
      }
    for (uint32_t i = 0; i < 32u; i++) label[order[i]] = i;
    for (uint32_t e = 0; e < E; e++) best_bank[e] = (uint8_t)label[best_bank[e] & 31u];
  }
  // banks -> columns: the i-th even position of bank b takes column b (i = 0), b + 32 (1), then block 1
  uint32_t cnt[32] = {0};
  bool ok = true;
  for (uint32_t e = 0; e < E && ok; e++) {
    const uint32_t b = best_bank[e], i = cnt[b]++;
    if (i >= cap) ok = false;
    const uint8_t v = (uint8_t)((i >> 1) << 7 | (b + 32u * (i & 1u)));
    slot[2u * e] = v;
    if (2u * e + 1u < kPeriodMaxLen) slot[2u * e + 1u] = v;
  }
  if (!ok)  // cannot happen (the natural layout respects the capacity); keep the natural layout
    for (uint32_t pos = 0; pos < kPeriodMaxLen; pos++) {
      const uint32_t blk = pos >> 7, q = pos & 127u;
      slot[pos] = (uint8_t)(blk << 7 | ((q >> 2) + 32u * ((q >> 1) & 1u)));
    }
  memcpy(cache->solved[cache_key].data(), slot, kPeriodMaxLen);
}

template <bool kAd, int kPW, bool kOdd, bool kAligned>
static cudaError_t period_launch_steps(const PArgs &args, uint32_t grid, cudaStream_t stream) {
  const uint32_t smem = args.plan.smem_bytes;
  switch (args.plan.steps) {
    case 3: period_kernel<kAd, 3, kPW, kOdd, kAligned><<<grid, kPW * 32, smem, stream>>>(args); break;
    case 4: period_kernel<kAd, 4, kPW, kOdd, kAligned><<<grid, kPW * 32, smem, stream>>>(args); break;
    case 5: period_kernel<kAd, 5, kPW, kOdd, kAligned><<<grid, kPW * 32, smem, stream>>>(args); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
// instantiated: even lengths x {16, 20, 24 warps} x {tiles on 16-byte boundaries or not}; odd lengths x {16, 20 warps}
template <bool kAd>
static cudaError_t period_launch_warps(const PArgs &args, uint32_t grid, cudaStream_t stream) {
  if (args.plan.len & 1u)
    return args.plan.warps == 16u   ? period_launch_steps<kAd, 16, true, false>(args, grid, stream)
           : args.plan.warps == 20u ? period_launch_steps<kAd, 20, true, false>(args, grid, stream)
                                    : cudaErrorInvalidValue;
  const bool al = (args.plan.tile_bytes & 15u) == 0u;
  switch (args.plan.warps) {
    case 16: return al ? period_launch_steps<kAd, 16, false, true>(args, grid, stream) : period_launch_steps<kAd, 16, false, false>(args, grid, stream);
    case 20: return al ? period_launch_steps<kAd, 20, false, true>(args, grid, stream) : period_launch_steps<kAd, 20, false, false>(args, grid, stream);
    case 24: return al ? period_launch_steps<kAd, 24, false, true>(args, grid, stream) : period_launch_steps<kAd, 24, false, false>(args, grid, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t period_configure() {
  cudaError_t e;
#define QB_PCFG(A, S, W, O, L)                                                                                          \
  if ((e = cudaFuncSetAttribute(period_kernel<A, S, W, O, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
#define QB_PCFG3(A, W, O, L) QB_PCFG(A, 3, W, O, L) QB_PCFG(A, 4, W, O, L) QB_PCFG(A, 5, W, O, L)
#define QB_PCFG_EVEN(A, W) QB_PCFG3(A, W, false, true) QB_PCFG3(A, W, false, false)
  QB_PCFG_EVEN(false, 16) QB_PCFG_EVEN(false, 20) QB_PCFG_EVEN(false, 24) QB_PCFG_EVEN(true, 16) QB_PCFG_EVEN(true, 20)
  QB_PCFG_EVEN(true, 24) QB_PCFG3(false, 16, true, false) QB_PCFG3(true, 16, true, false)
  QB_PCFG3(false, 20, true, false) QB_PCFG3(true, 20, true, false)
#undef QB_PCFG_EVEN
#undef QB_PCFG3
#undef QB_PCFG
  return cudaSuccess;
}

cudaError_t launch_period(const BatchView &b, const Accum &a, const AdapterSet &ad, const PeriodPlan &plan,
                          cudaStream_t stream, uint32_t *n_main_out) {
  *n_main_out = 0;
  if (!plan.ok || b.uniform_len != plan.len) return cudaErrorInvalidValue;
  const uint32_t n_tiles = b.n_reads / plan.reads_per_tile;
  if (n_tiles == 0) return cudaSuccess;
  PArgs args;
  memset(&args, 0, sizeof args);
  args.seq = b.seq + b.first_offset;
  args.qual = b.qual + b.first_offset;
  args.a = a;
  args.ad = ad;
  args.plan = plan;
  args.n_tiles = n_tiles;
  args.inc_lo = 1u, args.inc_hi = 0x10000u;
  period_slots(plan, args.slot);
  uint32_t grid = (n_tiles + plan.warps - 1u) / plan.warps;
  if (grid > plan.grid) grid = plan.grid;
  if (const char *g = getenv("QB_FUSED_GRID")) {  // test hook: few CTAs exercise the u16 flush path
    const uint32_t v = (uint32_t)atoi(g);
    if (v >= 1 && v < grid) grid = v;
  }
  const cudaError_t e = ad.enabled ? period_launch_warps<true>(args, grid, stream) : period_launch_warps<false>(args, grid, stream);
  if (e == cudaSuccess) *n_main_out = n_tiles * plan.reads_per_tile;
  return e;
}

}  // namespace qb

// tools / tests: the geometry and the counter layout the period kernel would use for reads of one length on a
// B200 (227 KiB of shared memory per block, 1 KiB reserved).  Needs no GPU.  Returns 0 if the kernel takes
// such batches, -1 otherwise.  info[0..6] = reads per period, words per period, steps, periods per tile,
// reads per tile, stages, warps; slot[p] = block << 7 | u32 column of position p.
extern "C" int qb_period_plan_info(uint32_t read_len, int adapters, uint32_t out[16]) {
  const qb::PeriodPlan p = qb::period_plan(read_len, 0, adapters, 148, 232448u, 1024u, 33u);
  if (!p.ok) return -1;
  const uint32_t v[16] = {p.k, p.wp, p.steps, p.ppt, p.reads_per_tile, p.stages, p.warps, p.nblocks, p.wblock, p.hdr_bytes,
                          p.rt, p.prow_stride, p.pring_hole, p.pring_wstride, p.pring_off, p.buf_bytes};
  for (int i = 0; i < 16; i++) out[i] = v[i];
  return 0;
}

extern "C" int qb_period_layout(uint32_t read_len, int adapters, uint32_t info[7], uint8_t slot[256]) {
  const qb::PeriodPlan p = qb::period_plan(read_len, 0, adapters, 148, 232448u, 1024u, 33u);
  if (!p.ok) return -1;
  if (info) info[0] = p.k, info[1] = p.wp, info[2] = p.steps, info[3] = p.ppt, info[4] = p.reads_per_tile, info[5] = p.stages, info[6] = p.warps;
  if (slot) qb::period_slots(p, slot);
  return 0;
}

