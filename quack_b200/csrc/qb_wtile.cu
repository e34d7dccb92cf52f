// qb_wtile.cu -- warp-tile kernel (v4) for quack's per-read statistics accumulation
// (reference: the while loop of read_fastq(), quack.c:193-221).  Reads of up to 192 bp.
//
// Same idea as fused_kernel (qb_kernels.cu) -- SWAR key bytes, joint (code, score) x position histogram of
// packed u16 counters in shared memory, 7-mer anchor filter + exact confirmation for -a -- but organised
// around AUTONOMOUS WARPS instead of a CTA-wide pipeline:
//   * every warp owns a private ring of small tiles (R whole reads, ~1-4 KB of seq + of qual): two quality /
//     key-byte buffers and ONE base buffer (the bases are dead once phase A has turned them into key bytes).
//     The warp itself issues the 1-D TMA bulk copies -- the bases of tile t+1 right after phase A of tile t,
//     the quality bytes of tile t+2 when it is done with tile t -- and then waits on the mbarrier of tile
//     t+1: no producer warp, no polling, no CTA barrier in the steady state.
//     Warps drift apart, so the ALU-heavy phase A of one warp overlaps the shared-memory-heavy phase H of
//     another (the v3 kernel ran them in lock step and left both pipes < 45 % busy);
//   * a small first pass (tile_desc_kernel, one warp per tile) validates the reads of every tile and writes an
//     8-byte tile descriptor (first byte, read count, common read length or 0, byte count); the hot kernel
//     reads one descriptor per tile, one tile ahead, and never touches offsets / lengths of tiles whose reads
//     all have one length and lie back to back (the normal case);
//   * key byte K = code << 6 | (q - qbase): 256 histogram rows of 256 bytes at a 64 KiB-aligned shared
//     address, so ONE byte permute builds the address of a base's counter from its key byte (2 instructions
//     per base in phase H: PRMT + RED), every byte value is a valid row, and a quality byte is inside the
//     counted window iff bits 7:6 of q - qbase are clear (one OR per word in phase A);
//   * candidate queue positions come from a ballot + popc instead of shared atomics, queue entries are one
//     per 16-byte unit (4-bit anchor mask), confirmed by quads of lanes that stop at the first hit.
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
constexpr uint32_t kWRows = 256;                 // histogram rows: key byte = code << 6 | score
constexpr uint32_t kWScores = 64;                // score field s = q - qbase in [0, 64)
constexpr uint32_t kMainBase = 0x10000u;         // shared address of the main histogram (positions 0..127)
constexpr uint32_t kMainBytes = kWRows * 256u;   // 64 KiB: rows of 64 u32 = 128 positions x u16
constexpr uint32_t kTail0 = 128u;                // first position of the tail histogram
constexpr uint32_t kTailRow = 128u;              // tail histogram (positions 128..191): rows of 32 u32 = 64 positions x u16
constexpr uint32_t kTailBytes = kWRows * kTailRow;
constexpr uint32_t kWMaxLen = kTail0 + 64u;      // longest read the histograms hold

// offsets inside a warp block (all multiples of 16)
constexpr uint32_t kWoBar = 0;                   // 2 mbarriers
constexpr uint32_t kWoStage = 16;                // per stage: descriptor (16 B), soff[32], slen[32]
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

// ---- phase A arithmetic ----
// loop-invariant SWAR constants
struct WKeyConsts {
  uint32_t m5b, x43, m1f, x07, x14, a7f, a3f, m80, c0, qsub;
  __device__ __forceinline__ explicit WKeyConsts(uint32_t qbase)
      : m5b(0x5B5B5B5Bu), x43(0x43434343u), m1f(0x1F1F1F1Fu), x07(0x07070707u), x14(0x14141414u), a7f(0x7F7F7F7Fu),
        a3f(0x3F3F3F3Fu), m80(0x80808080u), c0(0xC0C0C0C0u), qsub(qbase * 0x01010101u) {}
};
// One aligned word of 4 bases + 4 quality bytes -> 4 key bytes K = code << 6 | s, code = A0 T1 C2 G3
// (quack.c:150; N and everything else 0), s = q - qbase.  `nc` gets the inverted codes in bits 7:6 of each
// byte (other bits undefined), `qs` the 4 score bytes: a word whose qs has a bit 7 or 6 set holds a quality
// byte outside [qbase, qbase + 64) (a borrow can only start at a byte that is itself out of range) and is
// re-keyed by the caller.
__device__ __forceinline__ uint32_t w_key_bytes(uint32_t sw, uint32_t qw, const WKeyConsts &c, uint32_t &nc, uint32_t &qs) {
  // per byte: bit7 of n_cg is 0 iff (b & 0x5B) == 0x43; bit6 of n_g / n_t is 0 iff (b & 0x1F) == 7 / 0x14
  const uint32_t n_cg = lop3<0x6A>(sw, c.m5b, c.x43) + c.a7f;
  const uint32_t n_g = lop3<0x6A>(sw, c.m1f, c.x07) + c.a3f;
  const uint32_t n_t = lop3<0x6A>(sw, c.m1f, c.x14) + c.a3f;
  nc = lop3<0xE4>(n_cg, n_g & n_t, c.m80);  // bit 7 from n_cg, the rest from n_g & n_t
  qs = qw - c.qsub;
  return lop3<0xCE>(nc, qs, c.c0);  // (~nc & 0xC0C0C0C0) | qs
}
// key bytes of a word with an out-of-window quality byte: code kept, s = 63 (compensated by w_exact_word)
__device__ __forceinline__ uint32_t w_key_bytes_bad(uint32_t nc) { return (~nc & 0xC0C0C0C0u) | 0x3F3F3F3Fu; }
__device__ __forceinline__ bool w_word_bad(uint32_t qw, uint32_t qsub) { return ((qw - qsub) & 0xC0C0C0C0u) != 0u; }

// ---- phase H building blocks ----
// Position 4 l + t < 128 sits in u32 column l + 32 (t >> 1), half-word t & 1 of its row: the t-th atomic of
// a word step (lane l <-> positions 4 l .. 4 l + 3) touches bank l in every lane.  Position 128 + q lives in
// the tail histogram, column q & 31, half-word q >> 5 (byte steps, lane <-> position).
//
// per-lane constants: c = (shared address of the main histogram) | lane << 2; lo / hi = the increments of
// the low and the high u16 of a counter word, kept in registers so that the atomics are plain ATOMS.ADD
// (with immediates the compiler emits the warp-aggregating ATOMS.POPC.INC form instead)
struct WInc {
  uint32_t c, lo, hi;
};
__device__ __forceinline__ void w_red_word(uint32_t k4, WInc c, uint32_t i0, uint32_t i1, uint32_t i2, uint32_t i3) {
  // address = c with byte 1 replaced by the key byte: base | key << 8 | lane << 2
  red_shared_add<0u>(__byte_perm(k4, c.c, 0x7604), i0);
  red_shared_add<0u>(__byte_perm(k4, c.c, 0x7614), i1);
  red_shared_add<0x80u>(__byte_perm(k4, c.c, 0x7624), i2);
  red_shared_add<0x80u>(__byte_perm(k4, c.c, 0x7634), i3);
}
// the 4 key bytes of positions 4 lane .. + 3 of a word step whose lane-th word starts at kw (any alignment)
__device__ __forceinline__ uint32_t w_load_word(uint32_t kw) {
  const uint32_t al = kw & ~3u;
  return __funnelshift_r(lds_u32(al), lds_u32(al + 4u), kw << 3);
}

// shape of a read of length L: 0 nothing; 1 a partial word step (L < 128); 2 a full word step (L == 128);
// 3 / 4 a full word step and one / two tail byte steps
__device__ __forceinline__ uint32_t w_shape(uint32_t L) {
  if (L < 128u) return L ? 1u : 0u;
  return L == 128u ? 2u : (L <= 160u ? 3u : 4u);
}

template <int SHAPE>
struct WRead {
  uint32_t k4, b0, b1;
  // `part`: this lane has a position in the partial word step (the others must not load: they would read up
  // to 130 bytes behind the read's end)
  __device__ __forceinline__ void load(uint32_t kb, uint32_t lane, bool part) {
    k4 = 0;
    if (SHAPE >= 2 || part) k4 = w_load_word(kb + 4u * lane);
    if constexpr (SHAPE >= 3) b0 = lds_u8(kb + kTail0 + lane);
    if constexpr (SHAPE == 4) b1 = lds_u8(kb + kTail0 + 32u + lane);
  }
  // pinc: increments of the word step; t0 / t1: increments of the tail byte steps; tl: tail column address
  __device__ __forceinline__ void red(WInc c, const uint32_t (&pinc)[4], uint32_t tl, uint32_t t0, uint32_t t1) const {
    if constexpr (SHAPE == 1) w_red_word(k4, c, pinc[0], pinc[1], pinc[2], pinc[3]);
    if constexpr (SHAPE >= 2) w_red_word(k4, c, c.lo, c.hi, c.lo, c.hi);
    if constexpr (SHAPE >= 3) red_shared_add<0>(b0 * kTailRow + tl, t0);
    if constexpr (SHAPE == 4) red_shared_add<0>(b1 * kTailRow + tl, t1);
  }
};

// increments of the last step(s) of a read of length L (lanes behind the read's end add 0 to whatever row the
// stray key byte selects: every byte value is a row of the histogram)
template <int SHAPE>
__device__ __forceinline__ void w_tail(uint32_t L, uint32_t lane, uint32_t (&pinc)[4], uint32_t &t0, uint32_t &t1) {
  pinc[0] = pinc[1] = pinc[2] = pinc[3] = 0;
  t0 = t1 = 0;
  if (SHAPE == 1) {
    const uint32_t p = 4u * lane;
    pinc[0] = p < L ? 1u : 0u;
    pinc[1] = p + 1u < L ? 0x10000u : 0u;
    pinc[2] = p + 2u < L ? 1u : 0u;
    pinc[3] = p + 3u < L ? 0x10000u : 0u;
  }
  if (SHAPE >= 3) t0 = kTail0 + lane < L ? 1u : 0u;
  if (SHAPE == 4) t1 = kTail0 + 32u + lane < L ? 0x10000u : 0u;
}

// nr reads of length L back to back from kb.  Four reads per iteration: 4 L is a multiple of 4, so the
// byte alignment of read j of a group (the funnel-shift amount) is loop invariant and its aligned word
// address just advances by 4 L -- two address instructions per read instead of six.
template <int SHAPE>
__device__ __forceinline__ void w_uniform(WInc c, uint32_t tl, uint32_t kb, uint32_t L, uint32_t nr, uint32_t lane) {
  uint32_t pinc[4], t0, t1;
  w_tail<SHAPE>(L, lane, pinc, t0, t1);
  uint32_t r = 0;
  if (nr >= 4u) {
    uint32_t al[4], sh[4], kt[4];
    uint32_t kw = kb + 4u * lane, kq = kb + kTail0 + lane;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      al[j] = kw & ~3u;
      sh[j] = kw << 3;
      kt[j] = kq;
      kw += L;
      kq += L;
    }
    const uint32_t L4 = 4u * L;
#pragma unroll 1
    for (; r + 4u <= nr; r += 4u) {
      uint32_t k4[4], b0[4], b1[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        k4[j] = 0;
        if (SHAPE >= 2 || pinc[0]) k4[j] = __funnelshift_r(lds_u32(al[j]), lds_u32(al[j] + 4u), sh[j]);
        if constexpr (SHAPE >= 3) b0[j] = lds_u8(kt[j]);
        if constexpr (SHAPE == 4) b1[j] = lds_u8(kt[j] + 32u);
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if constexpr (SHAPE == 1) w_red_word(k4[j], c, pinc[0], pinc[1], pinc[2], pinc[3]);
        if constexpr (SHAPE >= 2) w_red_word(k4[j], c, c.lo, c.hi, c.lo, c.hi);
        if constexpr (SHAPE >= 3) red_shared_add<0>(b0[j] * kTailRow + tl, t0);
        if constexpr (SHAPE == 4) red_shared_add<0>(b1[j] * kTailRow + tl, t1);
        al[j] += L4;
        if constexpr (SHAPE >= 3) kt[j] += L4;
      }
    }
    kb += r * L;
  }
  for (; r < nr; r++) {
    WRead<SHAPE> x;
    x.load(kb, lane, pinc[0] != 0u);
    x.red(c, pinc, tl, t0, t1);
    kb += L;
  }
}
template <int SHAPE>
__device__ __forceinline__ void w_one(WInc c, uint32_t tl, uint32_t kb, uint32_t L, uint32_t lane) {
  uint32_t pinc[4], t0, t1;
  w_tail<SHAPE>(L, lane, pinc, t0, t1);
  WRead<SHAPE> x;
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

// Rare path of phase A: the 4 bases of a word with a quality byte outside the counted window.  The word is
// re-keyed to s = 63, so phase H counts its bases (content is right) in score row 63; here every base that
// lies inside a read gets its true score counted in the global accumulator and the row-63 count taken back
// (u64 arithmetic wraps: the sums are exact once the CTA's flush has added its part).  Returns the change of
// the invalid-quality count.
__device__ __noinline__ int w_exact_word(uint32_t qw, uint32_t abs0, const WTileMap tm, const Accum a, uint32_t qbase) {
  int d_invalid = 0;
  const int sc63 = (int)(63u + qbase) - 33;
  for (uint32_t j = 0; j < 4; j++) {
    uint32_t r, p, len;
    if (!w_locate(tm, abs0 + j, r, p, len)) continue;  // alignment slack: phase H does not count it either
    unsigned long long *row = a.rows + (size_t)p * kRow;
    const int sc = (int)((qw >> (8 * j)) & 0xFFu) - 33;
    if (sc >= 0 && sc < 91)
      atomicAdd(&row[sc], 1ull);
    else
      d_invalid++;
    if (sc63 >= 0 && sc63 < 91)
      atomicAdd(&row[sc63], ~0ull);  // - 1
    else
      d_invalid--;
  }
  return d_invalid;
}
// ... re-keying the words of a 16-byte unit that hold such a byte
__device__ __noinline__ int w_fix_bad_unit(uint4 qv, uint4 &K, uint32_t n0, uint32_t n1, uint32_t n2, uint32_t n3,
                                           uint32_t qsub, uint32_t abs0, const WTileMap tm, const Accum a, uint32_t qbase) {
  int d = 0;
  if (w_word_bad(qv.x, qsub)) K.x = w_key_bytes_bad(n0), d += w_exact_word(qv.x, abs0, tm, a, qbase);
  if (w_word_bad(qv.y, qsub)) K.y = w_key_bytes_bad(n1), d += w_exact_word(qv.y, abs0 + 4u, tm, a, qbase);
  if (w_word_bad(qv.z, qsub)) K.z = w_key_bytes_bad(n2), d += w_exact_word(qv.z, abs0 + 8u, tm, a, qbase);
  if (w_word_bad(qv.w, qsub)) K.w = w_key_bytes_bad(n3), d += w_exact_word(qv.w, abs0 + 12u, tm, a, qbase);
  return d;
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
// A tile that breaks the batch contract (offsets not ascending, reads overlapping or longer than len_cap,
// span beyond the staged buffer) gets an empty descriptor and raises the error counter: never corrupt
// silently.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_desc_kernel(BatchView b, Accum a, uint32_t R, uint32_t tile_bytes, uint32_t n_tiles) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (tile >= n_tiles) return;
  constexpr uint32_t kFull = 0xffffffffu;
  const uint32_t r0 = tile * R;
  const uint32_t nr = min(R, b.n_reads - r0);
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

// ------------------------------------------------------------------------------------------
// the hot kernel
// ------------------------------------------------------------------------------------------
template <bool kAdapters>
__global__ void __launch_bounds__(kWThreads, 1) wtile_kernel(const WArgs args) {
  extern __shared__ __align__(128) uint8_t smem[];
  const WtilePlan &P = args.plan;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t smem_s = smem_u32(smem);
  const uint32_t len_cap = args.a.len_cap;
  constexpr uint32_t kFull = 0xffffffffu;

  if (smem_s != P.smem_base) {  // the histogram must sit at its fixed shared address: fail loudly, count nothing
    if (tid == 0) atomicAdd(&args.a.counters[kCntError], 1ull);
    return;
  }
  auto gen = [&](uint32_t shared_addr) -> uint8_t * { return smem + (shared_addr - smem_s); };

  // ---- this warp's block ----
  const uint32_t wb_s = warp < P.region_n[0] ? P.region_s[0] + warp * P.wblock : P.region_s[1] + (warp - P.region_n[0]) * P.wblock;
  const uint32_t buf = P.buf;
  const uint32_t seqbuf_s = wb_s + wblock_hdr(kAdapters);  // bases of the current tile; quality / key bytes of stage s at + (1 + s) buf

  // ---- prologue: zero the histograms, load the adapter tables, init barriers ----
  auto clear_counters = [&]() {
    const uint4 z = make_uint4(0, 0, 0, 0);
    uint4 *h4 = reinterpret_cast<uint4 *>(gen(kMainBase));
    for (uint32_t i = tid; i < kMainBytes / 16u; i += kWThreads) h4[i] = z;
    uint4 *t4 = reinterpret_cast<uint4 *>(gen(P.tail_s));
    for (uint32_t i = tid; i < kTailBytes / 16u; i += kWThreads) t4[i] = z;
    uint32_t *lenhist = reinterpret_cast<uint32_t *>(gen(P.lenhist_s));
    uint32_t *kmerhist = reinterpret_cast<uint32_t *>(gen(P.kmerhist_s));
    for (uint32_t i = tid; i < len_cap; i += kWThreads) {
      lenhist[i] = 0;
      if (kAdapters) kmerhist[i] = 0;
    }
  };
  clear_counters();
  if (kAdapters) {
    uint32_t *af = reinterpret_cast<uint32_t *>(gen(P.afilt_s));
    for (uint32_t i = tid; i < kAnchorWords * kAnchorCopies; i += kWThreads) af[i] = args.ad.anchor[i / kAnchorCopies];
    uint32_t *ex = reinterpret_cast<uint32_t *>(gen(P.exact_s));
    if (args.ad.exact)
      for (uint32_t i = tid; i < kExactSlots; i += kWThreads) ex[i] = args.ad.exact[i];
    reinterpret_cast<uint32_t *>(gen(wb_s + kWoFhit))[lane] = kNoHit;
  }
  if (lane == 0) {
    mbar_init(reinterpret_cast<uint64_t *>(gen(wb_s + kWoBar)), 2);  // one arrival for the bases, one for the quality bytes
    mbar_init(reinterpret_cast<uint64_t *>(gen(wb_s + kWoBar + 8u)), 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&args.a.counters[kCntReads], (unsigned long long)args.b.n_reads);
  __syncthreads();

  const uint32_t R = P.reads_per_tile;
  const uint32_t n_tiles = args.n_tiles;
  const uint32_t G = gridDim.x * kWW;
  const uint32_t g0 = blockIdx.x * kWW;
  const uint32_t iters = n_tiles > g0 ? (n_tiles - g0 + G - 1u) / G : 0u;  // the same for every warp of the CTA
  uint32_t epoch = 65535u / ((uint32_t)kWW * R);  // iterations between two flushes of the u16 counters
  if (epoch == 0) epoch = 1;

  const uint32_t qbase = P.qbase;
  const WKeyConsts kc(qbase);
  WInc hc;  // this lane's column in the main histogram, counter units
  hc.c = kMainBase | (lane << 2);
  hc.lo = pin(1u), hc.hi = pin(0x10000u);
  const uint32_t tl = P.tail_s + (lane << 2);  // ... and in the tail histogram
  // this lane's copy of the anchor map (8 copies, 32-byte rows); the table is 16 KiB-aligned, so its base ORs in
  const uint32_t afilt_or = P.afilt_s | ((lane >> 2) * 4u);
  const uint32_t exact_s = P.exact_s;
  const uint32_t fhit_s = wb_s + kWoFhit, q_s = wb_s + kWoQueue;
  const uint32_t lenhist_s = P.lenhist_s, kmerhist_s = P.kmerhist_s;
  const uint32_t lt_mask = (1u << lane) - 1u;
  long long n_invalid = 0;

  auto flush = [&]() {  // all warps are behind a barrier
    const uint32_t npos = min(kWMaxLen, len_cap);
    for (uint32_t pos = tid; pos < npos; pos += kWThreads) {
      uint32_t base, rstride, sh;
      if (pos < kTail0) {
        base = kMainBase + 4u * ((pos >> 2) + 32u * ((pos & 3u) >> 1));
        sh = (pos & 1u) * 16u;
        rstride = 256u;
      } else {
        const uint32_t q = pos - kTail0;
        base = P.tail_s + 4u * (q & 31u);
        sh = (q >> 5) * 16u;
        rstride = kTailRow;
      }
      unsigned long long *row = args.a.rows + (size_t)pos * kRow;
      uint32_t cc[4] = {0, 0, 0, 0};
      for (uint32_t sp = 0; sp < kWScores; sp++) {
        uint32_t tot = 0;
#pragma unroll
        for (uint32_t c = 0; c < 4u; c++) {  // row = code << 6 | score
          const uint32_t v = (lds_u32(base + (c * kWScores + sp) * rstride) >> sh) & 0xFFFFu;
          cc[c] += v;
          tot += v;
        }
        if (tot) {
          const int sc = (int)(sp + qbase) - 33;
          if (sc >= 0 && sc < 91)
            atomicAdd(&row[sc], (unsigned long long)tot);
          else
            n_invalid += tot;
        }
      }
#pragma unroll
      for (uint32_t c = 0; c < 4u; c++)
        if (cc[c]) atomicAdd(&row[kColContent + c], (unsigned long long)cc[c]);
      const uint32_t lc = lds_u32(lenhist_s + pos * 4u);
      if (lc) atomicAdd(&row[kColLength], (unsigned long long)lc);
      if (kAdapters) {
        const uint32_t kcnt = lds_u32(kmerhist_s + pos * 4u);
        if (kcnt) atomicAdd(&row[kColKmer], (unsigned long long)kcnt);
      }
    }
  };

  // Start the bulk copy of one half of a tile (descriptor d): its quality bytes into stage s (the descriptor is
  // stashed in the stage header), or its bases into the base buffer.  Both arrive on the mbarrier of stage s.
  // The whole warp calls this once it is done with the buffer's previous contents.
  auto issue = [&](uint32_t s, uint2 d, bool bases) {
    __syncwarp();  // every lane is done with the buffer's old contents
    if (lane == 0) {
      const uint32_t lo_al = d.x & ~15u;
      const uint32_t span = ((d.x & 15u) + (d.y >> 16) + 15u) & ~15u;
      const uint32_t bar_s = wb_s + kWoBar + 8u * s;
      if (!bases) sts_u64(wb_s + kWoStage + s * kWStageHdr, d.x, d.y);
      // the TMA (async proxy) write must be ordered behind the generic-proxy accesses to the buffer
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if ((d.y & 63u) != 0u && span != 0u) {
        mbar_arrive_expect_tx_s(bar_s, span);
        if (bases)
          bulk_g2s_s(seqbuf_s, args.b.seq + lo_al, span, bar_s);
        else
          bulk_g2s_s(seqbuf_s + (1u + s) * buf, args.b.qual + lo_al, span, bar_s);
      } else {
        mbar_arrive(bar_s);
      }
    }
  };

  const uint32_t g = g0 + warp;
  const uint2 *tiles = args.b.tiles;
  if (g < n_tiles) {
    const uint2 d = __ldg(tiles + g);
    issue(0, d, false);
    issue(0, d, true);
    if (g + G < n_tiles) issue(1, __ldg(tiles + g + G), false);
  }

  uint32_t to_flush = epoch;
  uint32_t tile = g;
  for (uint32_t it = 0; it < iters; ++it, tile += G) {
    if (tile < n_tiles) {
      const uint32_t s = it & 1u;
      // descriptor of the tile after the next one: in flight while this tile is processed
      const uint32_t tile2 = tile + 2u * G;
      const bool more = tile2 < n_tiles;
      uint2 d2 = make_uint2(0u, 0u);
      if (more) d2 = __ldg(tiles + tile2);

      const uint32_t hdr_s = wb_s + kWoStage + s * kWStageHdr;
      mbar_wait(wb_s + kWoBar + 8u * s, (it >> 1) & 1u);
      const uint2 mt = lds_u64(hdr_s);
      const uint32_t lo = mt.x, nr = mt.y & 63u, ulen = (mt.y >> 6) & 1023u;
      const uint32_t lo_al = lo & ~15u;
      const uint32_t n16 = ((lo & 15u) + (mt.y >> 16) + 15u) >> 4;  // 16-byte units of the tile
      const uint32_t soff_s = hdr_s + 16u, slen_s = hdr_s + 144u;
      const uint32_t seq_s = seqbuf_s;
      const uint32_t key_s = seqbuf_s + (1u + s) * buf;  // phase A overwrites the quality bytes with the key bytes
      const uint32_t kd = key_s - seq_s;                 // quality / key byte of a base = its address + kd
      const WTileMap tm{lo, ulen, nr, soff_s};

      // ---------------- per-read counters (quack.c:219) ----------------
      if (ulen) {
        if (lane == 0) red_shared_add<0>(lenhist_s + (ulen - 1u) * 4u, nr);
      } else if (nr) {  // ragged tile: stage its offsets / lengths, lane <-> read (zero for lanes without a read)
        uint32_t off = 0, len = 0;
        if (lane < nr) {
          off = __ldg(args.b.offset + tile * R + lane);
          len = __ldg(args.b.length + tile * R + lane);
        }
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(soff_s + lane * 4u), "r"(off) : "memory");
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(slen_s + lane * 4u), "r"(len) : "memory");
        if (len) red_shared_add<0>(lenhist_s + (len - 1u) * 4u, 1u);  // len <= len_cap: checked by the first pass
        __syncwarp();
      }

      // ---------------- phase A: flat over the tile, key bytes written in place of the quality bytes ----------------
      uint32_t qn = 0;  // queued anchor hits
      if (!kAdapters) {
        for (uint32_t u = lane; u < n16; u += 32u) {
          const uint32_t a = seq_s + u * 16u;
          const uint4 sv = lds_u128(a), qv = lds_u128(a + kd);
          uint32_t n0, n1, n2, n3, q0, q1, q2, q3;
          uint4 K;
          K.x = w_key_bytes(sv.x, qv.x, kc, n0, q0);
          K.y = w_key_bytes(sv.y, qv.y, kc, n1, q1);
          K.z = w_key_bytes(sv.z, qv.z, kc, n2, q2);
          K.w = w_key_bytes(sv.w, qv.w, kc, n3, q3);
          if ((lop3<0xFE>(q0, q1, q2) | q3) & 0xC0C0C0C0u)  // rare: a quality byte outside the counted window
            n_invalid += w_fix_bad_unit(qv, K, n0, n1, n2, n3, kc.qsub, lo_al + u * 16u, tm, args.a, qbase);
          sts_u128(a + kd, K);
        }
      } else {
        // One 16-byte unit per lane.  A step covers 31 new units; lane 31 re-reads the unit behind them so
        // that the 7-mer anchors starting at bases 3, 7, 11, 15 of lanes 0..30 find their bases (one
        // shuffle).  The look-ahead lane computes but never stores or reports.
        for (uint32_t u0 = 0; u0 < n16; u0 += 31u) {
          const uint32_t u = u0 + lane;
          const uint32_t a = seq_s + min(u, n16) * 16u;  // at most the 16 bytes behind the span are read
          const uint4 sv = lds_u128(a), qv = lds_u128(a + kd);
          uint32_t n0, n1, n2, n3, q0, q1, q2, q3;
          uint4 K;
          K.x = w_key_bytes(sv.x, qv.x, kc, n0, q0);
          K.y = w_key_bytes(sv.y, qv.y, kc, n1, q1);
          K.z = w_key_bytes(sv.z, qv.z, kc, n2, q2);
          K.w = w_key_bytes(sv.w, qv.w, kc, n3, q3);
          const bool own = lane < 31u && u < n16;
          // 16 bases -> 32 bits.  The codes come from the base bytes alone (~n), not from the key bytes.
          const uint32_t p = pack16(~n0, ~n1, ~n2, ~n3);
          const uint32_t nx = __shfl_down_sync(kFull, p, 1);
          if (own) {
            if ((lop3<0xFE>(q0, q1, q2) | q3) & 0xC0C0C0C0u)
              n_invalid += w_fix_bad_unit(qv, K, n0, n1, n2, n3, kc.qsub, lo_al + u * 16u, tm, args.a, qbase);
            sts_u128(a + kd, K);
          }
          // anchor j = the 7-mer starting at base 4j+3: row = its bits 13:5 (32-byte rows), bit = its bits 4:0
          const uint32_t e0 = __funnelshift_r(p, nx, 6), e1 = __funnelshift_r(p, nx, 14);
          const uint32_t e2 = __funnelshift_r(p, nx, 22), e3 = __funnelshift_r(p, nx, 30);
          const uint32_t w0 = lds_u32(lop3<0xEA>(e0, 0x3FE0u, afilt_or)), w1 = lds_u32(lop3<0xEA>(e1, 0x3FE0u, afilt_or));
          const uint32_t w2 = lds_u32(lop3<0xEA>(e2, 0x3FE0u, afilt_or)), w3 = lds_u32(lop3<0xEA>(e3, 0x3FE0u, afilt_or));
          // bit 0 of m_j = anchor j passed (funnel shifts wrap at 32)
          const uint32_t m0 = __funnelshift_r(w0, 0u, e0), m1 = __funnelshift_r(w1, 0u, e1);
          const uint32_t m2 = __funnelshift_r(w2, 0u, e2), m3 = __funnelshift_r(w3, 0u, e3);
          const bool hit = own && ((lop3<0xFE>(m0, m1, m2) | m3) & 1u);
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
      // key bytes and the candidate queue of this tile are complete and its bases are dead: fetch the next tile's
      // (its descriptor sits in the other stage's header since its quality bytes were requested)
      if (tile + G < n_tiles) {
        const uint2 d1 = lds_u64(wb_s + kWoStage + (s ^ 1u) * kWStageHdr);
        issue(s ^ 1u, d1, true);
      } else {
        __syncwarp();
      }

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
            const uint32_t klo = pack16(ka.x, ka.y, ka.z, ka.w), khi = pack16(kb.x, kb.y, kb.z, kb.w);  // codes = bits 7:6
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
        if (ulen) {
          const uint32_t kb = k0_s + lo;
          switch (w_shape(ulen)) {
            case 1: w_uniform<1>(hc, tl, kb, ulen, nr, lane); break;
            case 2: w_uniform<2>(hc, tl, kb, ulen, nr, lane); break;
            case 3: w_uniform<3>(hc, tl, kb, ulen, nr, lane); break;
            case 4: w_uniform<4>(hc, tl, kb, ulen, nr, lane); break;
          }
        } else {
          for (uint32_t r = 0; r < nr; r++) {
            const uint32_t len = lds_u32(slen_s + r * 4u);
            const uint32_t kb = k0_s + lds_u32(soff_s + r * 4u);
            switch (w_shape(len)) {
              case 1: w_one<1>(hc, tl, kb, len, lane); break;
              case 2: w_one<2>(hc, tl, kb, len, lane); break;
              case 3: w_one<3>(hc, tl, kb, len, lane); break;
              case 4: w_one<4>(hc, tl, kb, len, lane); break;
            }
          }
        }
      }

      // the stage is free: refill it with the tile after the next one
      if (more) issue(s, d2, false);
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
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_invalid += __shfl_xor_sync(kFull, n_invalid, o);
  if (lane == 0 && n_invalid) atomicAdd(&args.a.counters[kCntInvalidQual], (unsigned long long)n_invalid);
}

// ------------------------------------------------------------------------------------------
// plan: shared-memory map and tile geometry
// ------------------------------------------------------------------------------------------

WtilePlan wtile_plan(uint32_t len_cap, uint32_t batch_max_len, int adapters, int sm_count, uint32_t smem_optin,
                     uint32_t smem_reserved, uint32_t qbase) {
  WtilePlan p;
  memset(&p, 0, sizeof p);
  if (len_cap == 0 || len_cap > kWMaxLen) return p;  // beyond the shared-memory histogram
  if (batch_max_len == 0 || batch_max_len > len_cap) batch_max_len = len_cap;
  p.qbase = qbase;
  p.smem_base = smem_reserved;  // dynamic shared memory starts right behind the driver's reserved bytes
  const uint32_t end = smem_reserved + smem_optin;
  // free address ranges below and above the main histogram
  struct Gap {
    uint32_t a, b;
  } gap[2] = {{smem_reserved, kMainBase}, {kMainBase + kMainBytes, end}};
  if (gap[0].a > gap[0].b || gap[1].a > gap[1].b) return p;
  auto take = [&](int g, uint32_t bytes) -> uint32_t {  // 0: does not fit
    bytes = (bytes + 15u) & ~15u;
    if (gap[g].b - gap[g].a < bytes) return 0;
    const uint32_t at = gap[g].a;
    gap[g].a += bytes;
    return at;
  };
  if (adapters) {
    if (!(p.afilt_s = take(1, kAnchorSmemBytes)) || (p.afilt_s & (kAnchorSmemBytes - 1u))) return p;  // 16 KiB-aligned
    if (!(p.exact_s = take(1, kExactSlots * 4u))) return p;
  }
  if (!(p.tail_s = take(1, kTailBytes))) return p;
  if (!(p.lenhist_s = take(1, len_cap * 4u))) return p;
  if (adapters && !(p.kmerhist_s = take(1, len_cap * 4u))) return p;

  // reads per tile: the R with the fewest warp instructions per base whose kWW warp blocks fit the gaps
  const uint32_t hdr = wblock_hdr(adapters);
  const uint32_t U = adapters ? 31u : 32u;                       // units per phase-A step
  const double c_step = adapters ? 100.0 : 62.0, c_tile = 140.0;  // warp instructions per step / per tile (ncu)
  const double c_group = 17.0 * 4.0 + 3.0, c_single = 24.0;      // phase H: 4 reads unrolled / one read
  uint32_t forced = 0;
  if (const char *e = getenv("QB_WT_READS")) forced = (uint32_t)atoi(e);  // tuning hook (tools/sweep_wtile.py)
  double best_score = 0;
  uint32_t best_r = 0;
  for (uint32_t r = 32; r >= 1; r--) {
    const uint32_t tb = (r * batch_max_len + 15u + 15u) & ~15u;
    if (tb > 1023u * 16u) continue;  // queue entries address 1024 units
    const uint32_t wblock = hdr + 3u * (tb + kWPad);
    if ((gap[0].b - gap[0].a) / wblock + (gap[1].b - gap[1].a) / wblock < (uint32_t)kWW) continue;
    const uint32_t units = (r * batch_max_len + 15u + 15u) / 16u;
    const double steps = (double)((units + U - 1u) / U);
    const double cost = steps * c_step + c_tile + (double)(r / 4u) * c_group + (double)(r % 4u) * c_single;
    const double score = (double)(r * batch_max_len) / cost;
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
  p.wblock = hdr + 3u * p.buf;
  uint32_t left = (uint32_t)kWW;
  for (int g = 0; g < 2; g++) {
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
  if ((e = cudaFuncSetAttribute(wtile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
  return cudaFuncSetAttribute(wtile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
}

cudaError_t launch_wtile(const BatchView &b, const Accum &a, const AdapterSet &ad, const WtilePlan &plan,
                         cudaStream_t stream) {
  if (b.n_reads == 0) return cudaSuccess;
  if (!b.tiles) return cudaErrorInvalidValue;
  WArgs args;
  args.b = b;
  args.a = a;
  args.ad = ad;
  args.plan = plan;
  args.n_tiles = (b.n_reads + plan.reads_per_tile - 1u) / plan.reads_per_tile;
  tile_desc_kernel<<<(args.n_tiles + 7u) / 8u, 256, 0, stream>>>(b, a, plan.reads_per_tile, plan.tile_bytes, args.n_tiles);
  uint32_t grid = (args.n_tiles + (uint32_t)kWW - 1u) / (uint32_t)kWW;
  if (grid > plan.grid) grid = plan.grid;
  if (const char *g = getenv("QB_FUSED_GRID")) {  // test hook: few CTAs exercise the u16 flush path
    const uint32_t v = (uint32_t)atoi(g);
    if (v >= 1 && v < grid) grid = v;
  }
  if (ad.enabled)
    wtile_kernel<true><<<grid, kWThreads, plan.smem_bytes, stream>>>(args);
  else
    wtile_kernel<false><<<grid, kWThreads, plan.smem_bytes, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace qb
