// qb_wtile.cu -- warp-tile kernel (v4) for quack's per-read statistics accumulation
// (reference: the while loop of read_fastq(), quack.c:193-221).
//
// Same arithmetic as fused_kernel (qb_kernels.cu) -- SWAR key bytes, joint (score, code) x position
// histogram of packed u16 counters in shared memory, 7-mer anchor filter + exact confirmation for -a --
// but organised around AUTONOMOUS WARPS instead of a CTA-wide pipeline:
//   * every warp owns a private 2-stage ring of small tiles (R whole reads, ~1-4 KB of seq + of qual).
//     The warp itself issues the 1-D TMA bulk copies of its tile t+2 when it is done with tile t and then
//     waits on the mbarrier of tile t+1: no producer warp, no polling, no CTA barrier in the steady state.
//     Warps drift apart, so the ALU-heavy phase A of one warp overlaps the shared-memory-heavy phase H of
//     another (the v3 kernel ran them in lock step and left both pipes < 45 % busy);
//   * a small first pass (tile_desc_kernel, one warp per tile) validates the reads of every tile and writes an
//     8-byte tile descriptor (first byte, read count, common read length or 0, byte count); the hot kernel
//     reads one descriptor per tile, one tile ahead, and never touches offsets / lengths of tiles whose reads
//     all have one length and lie back to back (the normal case);
//   * candidate queue positions come from a ballot + popc instead of shared atomics, queue entries are one
//     per 16-byte unit (4-bit anchor mask);
//   * histogram rows are 256 bytes at a 64 KiB-aligned shared address, so ONE byte permute builds the
//     address of a base's counter from its key byte (row = key): 2 instructions per base in phase H
//     (PRMT + RED) instead of 3.
// The CTA synchronises only in the prologue, at the final flush and every 65535 reads (u16 counters).
//
// No tensor cores: the path is an integer histogram.
#include "qb_dev.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace qb {

constexpr int kWW = kWtileWarps;                 // warps per CTA, one CTA per SM
constexpr int kWThreads = kWW * 32;
constexpr uint32_t kWQueue = 64;                 // anchor-hit entries per tile (one per 16-byte unit)
constexpr uint32_t kWPad = 32;                   // readable bytes behind a staged buffer (look-ahead unit + word)
constexpr uint32_t kMainBase = 0x10000u;         // shared address of histogram set 0 (set s at (s + 1) << 16)
constexpr uint32_t kMainBytes = kHistRows * 256u;  // 48 KiB: 192 rows of 64 u32 = 128 positions x u16
constexpr uint32_t kTailRow = 128u;              // tail histogram: 192 rows of 32 u32 = 64 positions x u16
constexpr uint32_t kTailBytes = kHistRows * kTailRow;

// offsets inside a warp block (all multiples of 16)
constexpr uint32_t kWoBar = 0;                   // 2 mbarriers
constexpr uint32_t kWoStage = 16;                // per stage: meta uint4, soff[32], slen[32]
constexpr uint32_t kWStageHdr = 16 + 128 + 128;
constexpr uint32_t kWoFhit = kWoStage + 2 * kWStageHdr;  // first-hit position per read of the tile (-a)
constexpr uint32_t kWoQueue = kWoFhit + 128;
__host__ __device__ inline uint32_t wblock_hdr(int adapters) { return adapters ? kWoQueue + kWQueue * 8u : kWoFhit; }

struct WArgs {
  BatchView b;
  Accum a;
  AdapterSet ad;
  WtilePlan plan;
  uint32_t n_tiles;
};

// ---- phase H building blocks ----
// Position p < 128 * kSets lives in set p >> 7; inside a set, position 4 l + t sits in u32 column
// l + 32 (t >> 1), half-word t & 1: the t-th atomic of a word step (lane l <-> positions 4 l .. 4 l + 3)
// touches bank l in every lane.  Positions behind the sets live in the tail histogram, position
// 128 kSets + q in column q & 31, half-word q >> 5 (byte steps, lane <-> position).
// per-lane constants of phase H: c = (shared address of set 0) | lane << 2; lo / hi = the increments of the
// low and the high u16 of a counter word, kept in registers so that the atomics are plain ATOMS.ADD (with
// immediates the compiler emits the warp-aggregating ATOMS.POPC.INC form instead)
struct WInc {
  uint32_t c, lo, hi;
};
template <uint32_t kOff>
__device__ __forceinline__ void w_red_word(uint32_t k4, WInc c, uint32_t i0, uint32_t i1, uint32_t i2, uint32_t i3) {
  // address = c with byte 1 replaced by the key byte: (set base) | key << 8 | lane << 2
  red_shared_add<kOff>(__byte_perm(k4, c.c, 0x7604), i0);
  red_shared_add<kOff>(__byte_perm(k4, c.c, 0x7614), i1);
  red_shared_add<kOff + 0x80u>(__byte_perm(k4, c.c, 0x7624), i2);
  red_shared_add<kOff + 0x80u>(__byte_perm(k4, c.c, 0x7634), i3);
}
// the 4 key bytes of positions 4 lane .. + 3 of a word step whose lane-th word starts at kw (any alignment)
__device__ __forceinline__ uint32_t w_load_word(uint32_t kw) {
  const uint32_t al = kw & ~3u;
  const uint32_t w0 = lds_u32(al), w1 = lds_u32(al + 4u);
  return __funnelshift_r(w0, w1, kw << 3);
}

// shape of a read of length L for kSets main sets: NF full word steps, then
//   KIND 0 nothing, 1 a partial word step in set NF (NF < kSets), 2 one tail byte step, 3 two tail byte steps
template <int kSets>
__device__ __forceinline__ uint32_t w_shape(uint32_t L) {
  const uint32_t nf = min(L >> 7, (uint32_t)kSets);
  const uint32_t rem = L - 128u * nf;
  uint32_t kind = 0;
  if (rem) kind = nf < (uint32_t)kSets ? 1u : (rem <= 32u ? 2u : 3u);
  return nf * 4u + kind;
}

template <int kSets, int NF, int KIND>
struct WRead {
  static constexpr int kWords = NF + (KIND == 1 ? 1 : 0);
  uint32_t k4[kWords > 0 ? kWords : 1];
  uint32_t kb0, kb1;
  // `part`: this lane has a position in the partial word step (the others must not load: they would read up
  // to 130 bytes behind the read's end)
  __device__ __forceinline__ void load(uint32_t kb, uint32_t lane, bool part) {
#pragma unroll
    for (int s = 0; s < NF; s++) k4[s] = w_load_word(kb + 128u * s + 4u * lane);
    if constexpr (KIND == 1) {
      k4[NF] = 0;
      if (part) k4[NF] = w_load_word(kb + 128u * NF + 4u * lane);
    }
    if constexpr (KIND >= 2) kb0 = lds_u8(kb + 128u * kSets + lane);
    if constexpr (KIND == 3) kb1 = lds_u8(kb + 128u * kSets + 32u + lane);
  }
  // pinc: increments of the partial word step; t0 / t1: increments of the tail byte steps; tl: tail column address
  __device__ __forceinline__ void red(WInc c, const uint32_t (&pinc)[4], uint32_t tl, uint32_t t0, uint32_t t1) const {
    if constexpr (NF >= 1) w_red_word<0u>(k4[0], c, c.lo, c.hi, c.lo, c.hi);
    if constexpr (NF >= 2) w_red_word<0x10000u>(k4[1], c, c.lo, c.hi, c.lo, c.hi);
    if constexpr (KIND == 1 && NF == 0) w_red_word<0u>(k4[0], c, pinc[0], pinc[1], pinc[2], pinc[3]);
    if constexpr (KIND == 1 && NF == 1) w_red_word<0x10000u>(k4[1], c, pinc[0], pinc[1], pinc[2], pinc[3]);
    if constexpr (KIND >= 2) red_shared_add<0>(kb0 * kTailRow + tl, t0);
    if constexpr (KIND == 3) red_shared_add<0>(kb1 * kTailRow + tl, t1);
  }
};

// increments of the last step(s) of a read of length L (lanes behind the read's end add 0 to whatever bin the
// stray key byte selects: stage buffers only ever hold key bytes < kHistRows or zeroes behind a tile)
template <int kSets, int NF, int KIND>
__device__ __forceinline__ void w_tail(uint32_t L, uint32_t lane, uint32_t (&pinc)[4], uint32_t &t0, uint32_t &t1) {
  pinc[0] = pinc[1] = pinc[2] = pinc[3] = 0;
  t0 = t1 = 0;
  if (KIND == 1) {
    const uint32_t p = 128u * NF + 4u * lane;
    pinc[0] = p < L ? 1u : 0u;
    pinc[1] = p + 1u < L ? 0x10000u : 0u;
    pinc[2] = p + 2u < L ? 1u : 0u;
    pinc[3] = p + 3u < L ? 0x10000u : 0u;
  }
  if (KIND >= 2) t0 = 128u * kSets + lane < L ? 1u : 0u;
  if (KIND == 3) t1 = 128u * kSets + 32u + lane < L ? 0x10000u : 0u;
}

// nr reads of length L back to back from kb.  Four reads per iteration: 4 L is a multiple of 4, so the
// byte alignment of read j of a group (the funnel-shift amount) is loop invariant and its aligned word
// address just advances by 4 L -- two address instructions per read instead of six.
template <int kSets, int NF, int KIND>
__device__ __forceinline__ void w_uniform(WInc c, uint32_t tl, uint32_t kb, uint32_t L, uint32_t nr, uint32_t lane) {
  uint32_t pinc[4], t0, t1;
  w_tail<kSets, NF, KIND>(L, lane, pinc, t0, t1);
  constexpr int kWords = NF + (KIND == 1 ? 1 : 0);
  uint32_t r = 0;
  if (nr >= 4u) {
    uint32_t al[4], sh[4], kt[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t kw = kb + (uint32_t)j * L + 4u * lane;
      al[j] = kw & ~3u;
      sh[j] = kw << 3;
      kt[j] = kb + (uint32_t)j * L + 128u * kSets + lane;
    }
    const uint32_t L4 = 4u * L;
    for (; r + 4u <= nr; r += 4u) {
      uint32_t k4[4][kWords > 0 ? kWords : 1], b0[4], b1[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
#pragma unroll
        for (int s = 0; s < NF; s++)
          k4[j][s] = __funnelshift_r(lds_u32(al[j] + 128u * s), lds_u32(al[j] + 128u * s + 4u), sh[j]);
        if constexpr (KIND == 1) {
          k4[j][NF] = 0;
          if (pinc[0]) k4[j][NF] = __funnelshift_r(lds_u32(al[j] + 128u * NF), lds_u32(al[j] + 128u * NF + 4u), sh[j]);
        }
        if constexpr (KIND >= 2) b0[j] = lds_u8(kt[j]);
        if constexpr (KIND == 3) b1[j] = lds_u8(kt[j] + 32u);
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if constexpr (NF >= 1) w_red_word<0u>(k4[j][0], c, c.lo, c.hi, c.lo, c.hi);
        if constexpr (NF >= 2) w_red_word<0x10000u>(k4[j][1], c, c.lo, c.hi, c.lo, c.hi);
        if constexpr (KIND == 1 && NF == 0) w_red_word<0u>(k4[j][0], c, pinc[0], pinc[1], pinc[2], pinc[3]);
        if constexpr (KIND == 1 && NF == 1) w_red_word<0x10000u>(k4[j][1], c, pinc[0], pinc[1], pinc[2], pinc[3]);
        if constexpr (KIND >= 2) red_shared_add<0>(b0[j] * kTailRow + tl, t0);
        if constexpr (KIND == 3) red_shared_add<0>(b1[j] * kTailRow + tl, t1);
        al[j] += L4;
        if constexpr (KIND >= 2) kt[j] += L4;
      }
    }
    kb += r * L;
  }
  for (; r < nr; r++) {
    WRead<kSets, NF, KIND> x;
    x.load(kb, lane, pinc[0] != 0u);
    x.red(c, pinc, tl, t0, t1);
    kb += L;
  }
}
template <int kSets, int NF, int KIND>
__device__ __forceinline__ void w_one(WInc c, uint32_t tl, uint32_t kb, uint32_t L, uint32_t lane) {
  uint32_t pinc[4], t0, t1;
  w_tail<kSets, NF, KIND>(L, lane, pinc, t0, t1);
  WRead<kSets, NF, KIND> x;
  x.load(kb, lane, pinc[0] != 0u);
  x.red(c, pinc, tl, t0, t1);
}

// Where absolute byte `abs` of the batch falls: read index inside the tile, position inside the read, read
// length.  Tiles whose reads all have length ulen and lie back to back from byte lo divide; the others
// search the staged offsets.  false: the byte belongs to no read of the tile.
struct WTileMap {
  uint32_t lo, ulen, nr, soff_s;  // soff_s: shared address of soff[32], slen[32] behind it (ragged tiles only)
};
__device__ __forceinline__ bool w_locate(const WTileMap &tm, uint32_t abs, uint32_t &r, uint32_t &pos, uint32_t &len) {
  if (tm.ulen) {
    const uint32_t d = abs - tm.lo;
    if ((int32_t)d < 0) return false;
    // d < 2^16 and (d + 0.5) / len is never closer than 1/(2 len) to an integer: the float quotient is exact
    r = (uint32_t)(((float)d + 0.5f) * __frcp_rn((float)tm.ulen));
    pos = d - r * tm.ulen;
    len = tm.ulen;
    return r < tm.nr;
  }
  const uint32_t *soff = shared_ptr<const uint32_t>(tm.soff_s);
  const int rr = find_read(soff, tm.nr, abs);
  if (rr < 0) return false;
  r = (uint32_t)rr;
  pos = abs - soff[r];
  len = soff[32 + r];
  return pos < len;
}

// rare path of phase A: the 4 bases of a word with an out-of-window quality byte, counted one by one
__device__ __noinline__ uint32_t w_exact_word(uint32_t sw, uint32_t qw, uint32_t abs0, const WTileMap tm, const Accum a) {
  uint32_t n_invalid = 0;
  for (uint32_t j = 0; j < 4; j++) {
    uint32_t r, p, len;
    if (!w_locate(tm, abs0 + j, r, p, len)) continue;  // alignment slack
    if (len > a.len_cap) continue;                      // a read the launch rejects anyway
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
// ... re-keying the words of a 16-byte unit whose quality bytes fall outside the window
__device__ __noinline__ uint32_t w_fix_bad_unit(uint4 sv, uint4 qv, uint4 &K, uint32_t n0, uint32_t n1, uint32_t n2,
                                                uint32_t n3, uint32_t qsub, uint32_t abs0, const WTileMap tm, const Accum a) {
  uint32_t n_invalid = 0;
  if (word_bad(qv.x, qsub)) K.x = key_bytes_bad(n0), n_invalid += w_exact_word(sv.x, qv.x, abs0, tm, a);
  if (word_bad(qv.y, qsub)) K.y = key_bytes_bad(n1), n_invalid += w_exact_word(sv.y, qv.y, abs0 + 4u, tm, a);
  if (word_bad(qv.z, qsub)) K.z = key_bytes_bad(n2), n_invalid += w_exact_word(sv.z, qv.z, abs0 + 8u, tm, a);
  if (word_bad(qv.w, qsub)) K.w = key_bytes_bad(n3), n_invalid += w_exact_word(sv.w, qv.w, abs0 + 12u, tm, a);
  return n_invalid;
}

// One window of an anchor hit.  `lo`/`hi` hold the 25 bases from the start of 16-byte unit `unit` (2 bits
// per base, first base least significant); `w` is the window start in bases from the unit start.  A window
// found in the exact key set whose 10 bases lie inside one read lowers that read's first-hit position.  A
// hit that ends on the last base of its read is dropped: it can only be the first hit if there is no
// other, and then the reference counts nothing (quack.c:215).
// Returns true when a hit was recorded and every later window of the unit lies in the same read (so none
// of them can be that read's first hit).
__device__ __forceinline__ bool w_confirm(uint32_t lo, uint32_t hi, uint32_t unit, uint32_t w, const AdapterSet ad,
                                          uint32_t exact_s, uint32_t lo_al, const WTileMap &tm, uint32_t fhit_s) {
  const uint32_t key = __funnelshift_r(lo, hi, 2u * w) & 0xFFFFFu;
  bool member;
  if (ad.exact)
    member = lds_u32(exact_s + exact_off1(key)) == key || lds_u32(exact_s + exact_off2(key)) == key;
  else
    member = (ad.bitmap[key >> 5] >> (key & 31u)) & 1u;
  if (!member) return false;
  uint32_t r, pos, len;
  if (!w_locate(tm, lo_al + unit * 16u + w + 9u, r, pos, len)) return false;  // byte on which the window ends
  if (pos < 9u || pos + 1u >= len) return false;  // the window spans two reads, or ends on the last base
  atomicMin(shared_ptr<uint32_t>(fhit_s) + r, pos);
  return pos + 16u < len;
}

// ------------------------------------------------------------------------------------------
// first pass: one descriptor per tile of R reads
//   x = first byte of the tile (offset of its first read)
//   y = reads (6 bits) | common read length if the reads all have one length and lie back to back, else 0
//       (10 bits) | bytes from the first byte of the first read to the last byte of the last (16 bits)
// A tile that breaks the batch contract (offsets not ascending, reads overlapping, span beyond the staged
// buffer) gets an empty descriptor and raises the error counter: never corrupt silently.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_desc_kernel(BatchView b, Accum a, uint32_t R, uint32_t tile_bytes, uint32_t n_tiles) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (tile >= n_tiles) return;
  constexpr uint32_t kFull = 0xffffffffu;
  const uint32_t r0 = tile * R;
  uint32_t nr = min(R, b.n_reads - r0);
  uint32_t off = 0, len = 0;
  if (lane < nr) {
    off = __ldg(b.offset + r0 + lane);
    len = __ldg(b.length + r0 + lane);
  }
  const uint32_t end = off + len;
  const uint32_t lo = __shfl_sync(kFull, off, 0);
  const uint32_t hi = __shfl_sync(kFull, end, nr - 1u);
  const uint32_t prev_end = __shfl_up_sync(kFull, end, 1);
  const uint32_t len0 = __shfl_sync(kFull, len, 0);
  const bool mine = lane < nr;
  const bool ok = !mine || (end >= off && (lane == 0 || off >= prev_end) && len <= a.len_cap);
  const bool back_to_back = !mine || (len == len0 && (lane == 0 || off == prev_end));
  const uint32_t span = (hi - (lo & ~15u) + 15u) & ~15u;
  const bool valid = __all_sync(kFull, ok) && hi >= lo && span <= tile_bytes && hi - lo < 65536u;
  uint32_t ulen = __all_sync(kFull, back_to_back) ? len0 : 0u;
  if (ulen > 1023u) ulen = 0;
  if (lane == 0) {
    if (!valid) atomicAdd(&a.counters[kCntError], 1ull);
    b.tiles[tile] = valid ? make_uint2(lo, nr | (ulen << 6) | ((hi - lo) << 16)) : make_uint2(0u, 0u);
  }
}

template <bool kAdapters, int kSets>
__global__ void __launch_bounds__(kWThreads, 1) wtile_kernel(const WArgs args) {
  extern __shared__ __align__(128) uint8_t smem[];
  const WtilePlan &P = args.plan;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t smem_s = smem_u32(smem);
  const uint32_t len_cap = args.a.len_cap;
  constexpr uint32_t kFull = 0xffffffffu;
  constexpr uint32_t kTail0 = 128u * kSets;  // first position of the tail histogram

  if (smem_s != P.smem_base) {  // the histogram must sit at its fixed shared address: fail loudly, count nothing
    if (tid == 0) atomicAdd(&args.a.counters[kCntError], 1ull);
    return;
  }
  auto gen = [&](uint32_t shared_addr) -> uint8_t * { return smem + (shared_addr - smem_s); };

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
  const uint32_t buf = P.buf;
  const uint32_t stage_s0 = wb_s + wblock_hdr(kAdapters);  // stage s: seq at + 2 s buf, qual / keys at + (2 s + 1) buf

  // ---- prologue: zero histograms and stage buffers, load the adapter tables, init barriers ----
  {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < kSets; s++) {
      uint4 *h4 = reinterpret_cast<uint4 *>(gen(kMainBase * (uint32_t)(s + 1)));
      for (uint32_t i = tid; i < kMainBytes / 16u; i += kWThreads) h4[i] = z;
    }
    uint4 *t4 = reinterpret_cast<uint4 *>(gen(P.tail_s));
    for (uint32_t i = tid; i < kTailBytes / 16u; i += kWThreads) t4[i] = z;
    uint32_t *lenhist = reinterpret_cast<uint32_t *>(gen(P.lenhist_s));
    uint32_t *kmerhist = reinterpret_cast<uint32_t *>(gen(P.kmerhist_s));
    for (uint32_t i = tid; i < len_cap; i += kWThreads) {
      lenhist[i] = 0;
      if (kAdapters) kmerhist[i] = 0;
    }
    if (kAdapters) {
      uint32_t *af = reinterpret_cast<uint32_t *>(gen(P.afilt_s));
      for (uint32_t i = tid; i < kAnchorWords * kAnchorCopies; i += kWThreads) af[i] = args.ad.anchor[i / kAnchorCopies];
      uint32_t *ex = reinterpret_cast<uint32_t *>(gen(P.exact_s));
      if (args.ad.exact)
        for (uint32_t i = tid; i < kExactSlots; i += kWThreads) ex[i] = args.ad.exact[i];
    }
    // the warp's own block: zeroed stage buffers (stray key bytes behind a tile then select valid rows)
    uint4 *b4 = reinterpret_cast<uint4 *>(gen(wb_s));
    for (uint32_t i = lane; i < P.wblock / 16u; i += 32u) b4[i] = z;
    __syncwarp();
    if (kAdapters) reinterpret_cast<uint32_t *>(gen(wb_s + kWoFhit))[lane] = kNoHit;
    if (lane == 0) {
      mbar_init(reinterpret_cast<uint64_t *>(gen(wb_s + kWoBar)), 1);
      mbar_init(reinterpret_cast<uint64_t *>(gen(wb_s + kWoBar + 8u)), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (blockIdx.x == 0 && tid == 0) atomicAdd(&args.a.counters[kCntReads], (unsigned long long)args.b.n_reads);
  }
  __syncthreads();

  const uint32_t R = P.reads_per_tile;
  const uint32_t n_tiles = args.n_tiles;
  const uint32_t G = gridDim.x * kWW;
  const uint32_t g0 = blockIdx.x * kWW;
  const uint32_t iters = n_tiles > g0 ? (n_tiles - g0 + G - 1u) / G : 0u;  // the same for every warp of the CTA
  uint32_t epoch = 65535u / ((uint32_t)kWW * R);  // iterations between two flushes of the u16 counters
  if (epoch == 0) epoch = 1;

  const uint32_t qsub = P.qbase * 0x01010101u;
  const KeyConsts kc(P.qbase);
  WInc hc;                                            // this lane's column in histogram set 0, counter units
  hc.c = kMainBase | (lane << 2);
#ifdef QB_WT_CONST_INC
  hc.lo = 1u, hc.hi = 0x10000u;
#else
  hc.lo = pin(1u), hc.hi = pin(0x10000u);
#endif
  const uint32_t tl = P.tail_s + (lane << 2);         // ... and in the tail histogram
  const uint32_t afilt_s = P.afilt_s, afilt_copy = (lane >> 2) * 4u;
  const uint32_t exact_s = P.exact_s;
  const uint32_t fhit_s = wb_s + kWoFhit, q_s = wb_s + kWoQueue;
  const uint32_t lenhist_s = P.lenhist_s, kmerhist_s = P.kmerhist_s;
  const uint32_t lt_mask = (1u << lane) - 1u;
  unsigned long long n_invalid = 0;

  auto flush = [&]() {  // all warps are behind a barrier
    const uint32_t npos = min(kTail0 + 64u, len_cap);
    for (uint32_t pos = tid; pos < npos; pos += kWThreads) {
      uint32_t base, rstride, sh;
      if (pos < kTail0) {
        const uint32_t q = pos & 127u;
        base = kMainBase * ((pos >> 7) + 1u) + 4u * ((q >> 2) + 32u * ((q & 3u) >> 1));
        sh = (q & 1u) * 16u;
        rstride = 256u;
      } else {
        const uint32_t q = pos - kTail0;
        base = P.tail_s + 4u * (q & 31u);
        sh = (q >> 5) * 16u;
        rstride = kTailRow;
      }
      unsigned long long *row = args.a.rows + (size_t)pos * kRow;
      uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      for (uint32_t sp = 0; sp < kScoreBins; sp++) {  // the rows of s = kScoreBins are the dummies
        const uint32_t a = base + 4u * sp * rstride;
        const uint32_t v0 = (lds_u32(a) >> sh) & 0xFFFFu, v1 = (lds_u32(a + rstride) >> sh) & 0xFFFFu;
        const uint32_t v2 = (lds_u32(a + 2u * rstride) >> sh) & 0xFFFFu, v3 = (lds_u32(a + 3u * rstride) >> sh) & 0xFFFFu;
        const uint32_t tot = v0 + v1 + v2 + v3;
        c0 += v0, c1 += v1, c2 += v2, c3 += v3;
        if (tot) {
          const int sc = (int)(sp + P.qbase) - 33;
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
      const uint32_t lc = lds_u32(lenhist_s + pos * 4u);
      if (lc) atomicAdd(&row[kColLength], (unsigned long long)lc);
      if (kAdapters) {
        const uint32_t kcnt = lds_u32(kmerhist_s + pos * 4u);
        if (kcnt) atomicAdd(&row[kColKmer], (unsigned long long)kcnt);
      }
    }
  };
  auto clear_counters = [&]() {  // between two epochs
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < kSets; s++) {
      uint4 *h4 = reinterpret_cast<uint4 *>(gen(kMainBase * (uint32_t)(s + 1)));
      for (uint32_t i = tid; i < kMainBytes / 16u; i += kWThreads) h4[i] = z;
    }
    uint4 *t4 = reinterpret_cast<uint4 *>(gen(P.tail_s));
    for (uint32_t i = tid; i < kTailBytes / 16u; i += kWThreads) t4[i] = z;
    uint32_t *lenhist = reinterpret_cast<uint32_t *>(gen(P.lenhist_s));
    uint32_t *kmerhist = reinterpret_cast<uint32_t *>(gen(P.kmerhist_s));
    for (uint32_t i = tid; i < len_cap; i += kWThreads) {
      lenhist[i] = 0;
      if (kAdapters) kmerhist[i] = 0;
    }
  };

  // Start the bulk copies of a tile (descriptor d) into stage s and stash the descriptor in the stage header.
  // The whole warp calls this once it is done with the stage's previous tile.
  auto issue = [&](uint32_t s, uint2 d) {
    __syncwarp();  // every lane is done with the stage's old contents
    if (lane == 0) {
      const uint32_t lo_al = d.x & ~15u;
      const uint32_t span = ((d.x & 15u) + (d.y >> 16) + 15u) & ~15u;
      sts_u64(wb_s + kWoStage + s * kWStageHdr, d.x, d.y);
      // the TMA (async proxy) write must be ordered behind the generic-proxy key-byte writes into the buffer
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      uint64_t *bar = reinterpret_cast<uint64_t *>(gen(wb_s + kWoBar + 8u * s));
      const bool any = (d.y & 63u) != 0u && span != 0u;
      mbar_arrive_expect_tx(bar, any ? 2u * span : 0u);
      if (any) {
        bulk_g2s(gen(stage_s0 + 2u * s * buf), args.b.seq + lo_al, span, bar);
        bulk_g2s(gen(stage_s0 + (2u * s + 1u) * buf), args.b.qual + lo_al, span, bar);
      }
    }
  };

  const uint32_t g = g0 + warp;
  const uint2 *tiles = args.b.tiles;
  for (uint32_t k = 0; k < 2u; k++) {
    const uint32_t tile = g + k * G;
    if (tile < n_tiles) issue(k, __ldg(tiles + tile));
  }

  uint32_t to_flush = epoch;
  uint32_t tile = g;
  for (uint32_t it = 0; it < iters; ++it, tile += G) {
    if (tile < n_tiles) {
      const uint32_t s = it & 1u;
      // descriptor of the tile after the next one: in flight while this tile is processed
      const bool more = n_tiles - tile > 2u * G;
      uint2 d2 = make_uint2(0u, 0u);
      if (more) d2 = __ldg(tiles + tile + 2u * G);

      const uint32_t hdr_s = wb_s + kWoStage + s * kWStageHdr;
      mbar_wait(wb_s + kWoBar + 8u * s, (it >> 1) & 1u);
      const uint2 mt = lds_u64(hdr_s);
      const uint32_t lo = mt.x, nr = mt.y & 63u, ulen = (mt.y >> 6) & 1023u;
      const uint32_t lo_al = lo & ~15u;
      const uint32_t n16 = ((lo & 15u) + (mt.y >> 16) + 15u) >> 4;  // 16-byte units of the tile
      const uint32_t soff_s = hdr_s + 16u, slen_s = hdr_s + 144u;
      const uint32_t seq_s = stage_s0 + 2u * s * buf;
      const uint32_t key_s = seq_s + buf;  // phase A overwrites the quality bytes with the key bytes
      const WTileMap tm{lo, ulen, nr, soff_s};

      if (!ulen && nr) {  // ragged tile: stage its offsets / lengths, lane <-> read (zero for lanes without a read)
        uint32_t off = 0, len = 0;
        if (lane < nr) {
          off = __ldg(args.b.offset + tile * R + lane);
          len = __ldg(args.b.length + tile * R + lane);
        }
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(soff_s + lane * 4u), "r"(off) : "memory");
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(slen_s + lane * 4u), "r"(len) : "memory");
        if (len) red_shared_add<0>(lenhist_s + (len - 1u) * 4u, 1u);  // quack.c:219 (len <= len_cap: checked by the first pass)
        __syncwarp();
      } else if (ulen && lane == 0) {
        red_shared_add<0>(lenhist_s + (ulen - 1u) * 4u, nr);
      }

      // ---------------- phase A: flat over the tile, key bytes written in place of the quality bytes ----------------
      uint32_t qn = 0;  // queued anchor hits
      if (!kAdapters) {
        for (uint32_t u = lane; u < n16; u += 32u) {
          const uint32_t a = seq_s + u * 16u;
          const uint4 sv = lds_u128(a), qv = lds_u128(a + buf);
          uint32_t n0, n1, n2, n3, bad = 0;
          uint4 K;
          K.x = key_bytes(sv.x, qv.x, kc, n0, bad);
          K.y = key_bytes(sv.y, qv.y, kc, n1, bad);
          K.z = key_bytes(sv.z, qv.z, kc, n2, bad);
          K.w = key_bytes(sv.w, qv.w, kc, n3, bad);
          if (bad & 0xC0C0C0C0u)  // rare: re-key the offending words to the dummy rows, count them exactly
            n_invalid += w_fix_bad_unit(sv, qv, K, n0, n1, n2, n3, qsub, lo_al + u * 16u, tm, args.a);
          sts_u128(a + buf, K);
        }
      } else {
        // One 16-byte unit per lane.  A step covers 31 new units; lane 31 re-reads the unit behind them so
        // that the 7-mer anchors starting at bases 3, 7, 11, 15 of lanes 0..30 find their bases (one
        // shuffle).  The look-ahead lane computes but never stores or reports.
        for (uint32_t u0 = 0; u0 < n16; u0 += 31u) {
          const uint32_t u = u0 + lane;
          const uint32_t a = seq_s + min(u, n16) * 16u;  // at most the 16 bytes behind the span are read
          const uint4 sv = lds_u128(a), qv = lds_u128(a + buf);
          uint32_t n0, n1, n2, n3, bad = 0;
          uint4 K;
          K.x = key_bytes(sv.x, qv.x, kc, n0, bad);
          K.y = key_bytes(sv.y, qv.y, kc, n1, bad);
          K.z = key_bytes(sv.z, qv.z, kc, n2, bad);
          K.w = key_bytes(sv.w, qv.w, kc, n3, bad);
          const bool own = lane < 31u && u < n16;
          // 16 bases -> 32 bits.  The codes come from the base bytes alone (~n), not from the key bytes.
          const uint32_t p = pack16(~n0, ~n1, ~n2, ~n3);
          const uint32_t nx = __shfl_down_sync(kFull, p, 1);
          if (own) {
            if (bad & 0xC0C0C0C0u)
              n_invalid += w_fix_bad_unit(sv, qv, K, n0, n1, n2, n3, qsub, lo_al + u * 16u, tm, args.a);
            sts_u128(a + buf, K);
          }
          // anchor j = the 7-mer starting at base 4j+3: row = its bits 13:5 (32-byte rows), bit = its bits 4:0
          const uint32_t e0 = __funnelshift_r(p, nx, 6), e1 = __funnelshift_r(p, nx, 14);
          const uint32_t e2 = __funnelshift_r(p, nx, 22), e3 = __funnelshift_r(p, nx, 30);
          const uint32_t w0 = lds_u32_at(lop3<0xEA>(e0, 0x3FE0u, afilt_copy), afilt_s);
          const uint32_t w1 = lds_u32_at(lop3<0xEA>(e1, 0x3FE0u, afilt_copy), afilt_s);
          const uint32_t w2 = lds_u32_at(lop3<0xEA>(e2, 0x3FE0u, afilt_copy), afilt_s);
          const uint32_t w3 = lds_u32_at(lop3<0xEA>(e3, 0x3FE0u, afilt_copy), afilt_s);
          // bit 0 of m_j = anchor j passed (shift amounts wrap at 32)
          const uint32_t m0 = __funnelshift_r(w0, 0u, e0), m1 = __funnelshift_r(w1, 0u, e1);
          const uint32_t m2 = __funnelshift_r(w2, 0u, e2), m3 = __funnelshift_r(w3, 0u, e3);
          const bool hit = own && ((m0 | m1 | m2 | m3) & 1u);
          const uint32_t bal = __ballot_sync(kFull, hit);
          if (bal) {  // one queue entry per unit with a hit: 25 bases, anchor mask, unit
            if (hit) {
              const uint32_t idx = qn + __popc(bal & lt_mask);
              const uint32_t m = (m0 & 1u) | ((m1 & 1u) << 1) | ((m2 & 1u) << 2) | ((m3 & 1u) << 3);
              if (idx < kWQueue) sts_u64(q_s + idx * 8u, p, (nx & 0x3FFFFu) | (m << 18) | (u << 22));
            }
            qn += __popc(bal);
          }
        }
      }
      __syncwarp();  // key bytes and the candidate queue of this tile are complete

      // ---------------- phase A2: confirm the queued anchor hits (quack.c:210-217) ----------------
      if (kAdapters && qn) {
        if (qn <= kWQueue) {
          // 8 entries per pass: a quad of lanes per entry, lane t of the quad tests the windows 4j + t of the
          // entry's anchors j in ascending order; the quad stops at the first recorded hit whose read also
          // holds the rest of the unit (no later window can then be that read's first hit)
          for (uint32_t e0 = 0; e0 < qn; e0 += 8u) {
            const uint32_t e = e0 + (lane >> 2);
            uint2 c = make_uint2(0u, 0u);
            if (e < qn) c = lds_u64(q_s + e * 8u);
            uint32_t m = (c.y >> 18) & 15u;
            while (__any_sync(kFull, m != 0u)) {
              bool stop = false;
              if (m) {  // anchor j starts at base 4j+3 of its unit; the windows that contain it start at 4j .. 4j+3
                const uint32_t j = __ffs(m) - 1u;
                m &= m - 1u;
                stop = w_confirm(c.x, c.y & 0x3FFFFu, c.y >> 22, 4u * j + (lane & 3u), args.ad, exact_s, lo_al, tm, fhit_s);
              }
              const uint32_t stops = __ballot_sync(kFull, stop);
              if ((stops >> (lane & 28u)) & 15u) m = 0;
            }
          }
        } else {  // queue overflow (adapter-dimer-like data): test every window of the tile exactly
          for (uint32_t i = lane; i < n16 * 4u; i += 32u) {
            const uint32_t unit = i >> 2, t = i & 3u;
            const uint4 ka = lds_u128(key_s + unit * 16u), kb = lds_u128(key_s + unit * 16u + 16u);
            const uint32_t klo = pack16_keys(ka), khi = pack16_keys(kb);
            for (uint32_t j = 0; j < 4u; j++) w_confirm(klo, khi, unit, 4u * j + t, args.ad, exact_s, lo_al, tm, fhit_s);
          }
        }
        __syncwarp();
        // first hit of a read -> adapter histogram (quack.c:215-217: counted at p + 1, which is inside the read)
        const uint32_t fa = fhit_s + lane * 4u;
        const uint32_t pfirst = lds_u32(fa);
        if (pfirst != kNoHit) {
          if (pfirst + 1u < len_cap) red_shared_add<4>(kmerhist_s + pfirst * 4u, 1u);
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(fa), "r"(kNoHit) : "memory");
        }
      }

      // ---------------- phase H: one shared atomic per base ----------------
      {
        const uint32_t k0_s = key_s - lo_al;  // + absolute offset of a read = shared address of its first key byte
#define QB_WSHAPES(FN) FN(0, 1) FN(1, 0) FN(1, 1) FN(1, 2) FN(1, 3) FN(2, 0) FN(2, 2) FN(2, 3)
#define QB_WVALID(nf, kind) ((nf) <= kSets && ((kind) == 0 || ((kind) == 1 ? (nf) < kSets : (nf) == kSets)))
        if (ulen) {
          switch (w_shape<kSets>(ulen)) {
#define QB_WCASE(nf, kind)                                                                          \
  case (nf)*4 + (kind):                                                                             \
    if constexpr (QB_WVALID(nf, kind)) w_uniform<kSets, nf, kind>(hc, tl, k0_s + lo, ulen, nr, lane); \
    break;
            QB_WSHAPES(QB_WCASE)
#undef QB_WCASE
          }
        } else {
          for (uint32_t r = 0; r < nr; r++) {
            const uint32_t len = lds_u32(slen_s + r * 4u);
            if (len == 0) continue;
            const uint32_t kb = k0_s + lds_u32(soff_s + r * 4u);
            switch (w_shape<kSets>(len)) {
#define QB_WCASE(nf, kind)                                                                  \
  case (nf)*4 + (kind):                                                                     \
    if constexpr (QB_WVALID(nf, kind)) w_one<kSets, nf, kind>(hc, tl, kb, len, lane);        \
    break;
              QB_WSHAPES(QB_WCASE)
#undef QB_WCASE
            }
          }
        }
#undef QB_WVALID
#undef QB_WSHAPES
      }

      // the stage is free: refill it with the tile after the next one
      if (more) issue(s, d2);
    }
    if (--to_flush == 0u && it + 1u < iters) {  // u16 counters: flush before any bin can wrap
      to_flush = epoch;
      __syncthreads();
      flush();
      __syncthreads();
      clear_counters();
      __syncthreads();
    }
  }
  __syncthreads();
  flush();
  n_invalid = warp_sum(n_invalid);
  if (lane == 0 && n_invalid) atomicAdd(&args.a.counters[kCntInvalidQual], n_invalid);
}

// ------------------------------------------------------------------------------------------
// plan: shared-memory map and tile geometry
// ------------------------------------------------------------------------------------------

WtilePlan wtile_plan(uint32_t len_cap, uint32_t batch_max_len, int adapters, int sm_count, uint32_t smem_optin,
                     uint32_t smem_reserved, uint32_t qbase) {
  WtilePlan p;
  memset(&p, 0, sizeof p);
  if (len_cap == 0 || len_cap > 320u) return p;  // beyond the shared-memory histogram
  if (batch_max_len == 0 || batch_max_len > len_cap) batch_max_len = len_cap;
  p.nsets = len_cap <= 192u ? 1u : 2u;
  p.qbase = qbase;
  p.smem_base = smem_reserved;  // dynamic shared memory starts right behind the driver's reserved bytes
  const uint32_t end = smem_reserved + smem_optin;
  // free address ranges around the histogram sets
  struct Gap {
    uint32_t a, b;
  } gap[3] = {{smem_reserved, kMainBase}, {kMainBase + kMainBytes, 2u * kMainBase}, {0, 0}};
  if (p.nsets == 1)
    gap[2] = Gap{2u * kMainBase, end};
  else
    gap[2] = Gap{2u * kMainBase + kMainBytes, end};
  if (gap[0].a > gap[0].b || gap[2].a > gap[2].b) return p;
  auto take = [&](int g, uint32_t bytes) -> uint32_t {  // 0: does not fit
    bytes = (bytes + 15u) & ~15u;
    if (gap[g].b - gap[g].a < bytes) return 0;
    const uint32_t at = gap[g].a;
    gap[g].a += bytes;
    return at;
  };
  if (adapters) {
    if (!(p.afilt_s = take(1, kAnchorSmemBytes))) return p;  // exactly the 16 KiB between set 0 and 0x20000
    if (!(p.exact_s = take(2, kExactSlots * 4u))) return p;
  }
  if (!(p.tail_s = take(2, kTailBytes))) return p;
  if (!(p.lenhist_s = take(2, len_cap * 4u))) return p;
  if (adapters && !(p.kmerhist_s = take(2, len_cap * 4u))) return p;

  // reads per tile: the largest-throughput R whose kWW warp blocks fit the remaining gaps
  const uint32_t hdr = wblock_hdr(adapters);
  const uint32_t U = adapters ? 31u : 32u;                 // units per phase-A step
  const double c_step = adapters ? 115.0 : 72.0, c_tile = 75.0;  // warp instructions per step / per tile (measured)
  uint32_t forced = 0;
  if (const char *e = getenv("QB_WT_READS")) forced = (uint32_t)atoi(e);  // tuning hook (tools/sweep_wtile.py)
  double best_score = 0;
  uint32_t best_r = 0;
  for (uint32_t r = 32; r >= 1; r--) {
    const uint32_t tb = (r * batch_max_len + 15u + 15u) & ~15u;
    if (tb > 1023u * 16u) continue;  // queue entries address 1024 units
    const uint32_t wblock = hdr + 4u * (tb + kWPad);
    uint32_t fit = 0;
    for (int g = 0; g < 3; g++) fit += (gap[g].b - gap[g].a) / wblock;
    if (fit < (uint32_t)kWW) continue;
    const uint32_t units = (r * batch_max_len + 15u + 15u) / 16u;
    const double steps = (double)((units + U - 1u) / U);
    const double score = (double)(r * batch_max_len) / (steps * c_step + c_tile);
    if (forced ? r == forced : score > best_score) {
      best_score = score;
      best_r = r;
      if (forced) break;
    }
  }
  if (!best_r) return p;
  p.reads_per_tile = best_r;
  p.tile_bytes = (best_r * batch_max_len + 15u + 15u) & ~15u;
  p.buf = p.tile_bytes + kWPad;
  p.wblock = hdr + 4u * p.buf;
  uint32_t left = (uint32_t)kWW;
  for (int g = 0; g < 3; g++) {
    uint32_t n = (gap[g].b - gap[g].a) / p.wblock;
    if (n > left) n = left;
    p.region_s[g] = gap[g].a;
    p.region_n[g] = n;
    left -= n;
  }
  p.smem_bytes = smem_optin;
  p.grid = (uint32_t)sm_count;
  p.ok = 1;
  return p;
}

cudaError_t wtile_configure() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(wtile_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
  if ((e = cudaFuncSetAttribute(wtile_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
  if ((e = cudaFuncSetAttribute(wtile_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
  return cudaFuncSetAttribute(wtile_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
}

cudaError_t launch_wtile(const BatchView &b, const Accum &a, const AdapterSet &ad, const WtilePlan &plan,
                         cudaStream_t stream) {
  if (b.n_reads == 0) return cudaSuccess;
  WArgs args;
  args.b = b;
  args.a = a;
  args.ad = ad;
  args.plan = plan;
  args.n_tiles = (b.n_reads + plan.reads_per_tile - 1u) / plan.reads_per_tile;
  if (!b.tiles) return cudaErrorInvalidValue;
  tile_desc_kernel<<<(args.n_tiles + 7u) / 8u, 256, 0, stream>>>(b, a, plan.reads_per_tile, plan.tile_bytes, args.n_tiles);
  uint32_t grid = (args.n_tiles + (uint32_t)kWW - 1u) / (uint32_t)kWW;
  if (grid > plan.grid) grid = plan.grid;
  if (const char *g = getenv("QB_FUSED_GRID")) {  // test hook: few CTAs exercise the u16 flush path
    const uint32_t v = (uint32_t)atoi(g);
    if (v >= 1 && v < grid) grid = v;
  }
  if (plan.nsets == 1u) {
    if (ad.enabled)
      wtile_kernel<true, 1><<<grid, kWThreads, plan.smem_bytes, stream>>>(args);
    else
      wtile_kernel<false, 1><<<grid, kWThreads, plan.smem_bytes, stream>>>(args);
  } else {
    if (ad.enabled)
      wtile_kernel<true, 2><<<grid, kWThreads, plan.smem_bytes, stream>>>(args);
    else
      wtile_kernel<false, 2><<<grid, kWThreads, plan.smem_bytes, stream>>>(args);
  }
  return cudaGetLastError();
}

}  // namespace qb
