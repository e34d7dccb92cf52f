// qb_flat.cu -- flat kernel (v6) for quack's per-read statistics accumulation (reference: the while loop of
// read_fastq(), quack.c:193-221) on RAGGED batches: reads of any lengths in [16, 304] lying back to back (what
// the host reader packs; config 4, trimmed reads of 35..300 bp).  The v3 kernel gives such batches one warp per
// read, lane <-> 4 positions: a 167-bp read fills 42 of 64 lane slots and pays its per-read set-up again and again
// (0.31 / 0.15 of the HBM roofline without / with -a).  Here nothing is per read in the hot loop:
//   * the batch is cut into CHUNKS of whole reads by byte windows (chunk c = the reads that START in bytes
//     [c CB, (c + 1) CB); a one-pass index kernel finds each window's first read), one warp per chunk at a time;
//   * a warp walks its chunk flat, lane <-> aligned 32-bit word, 32 consecutive words per step: coalesced loads of
//     128 contiguous bytes of bases and of quality bytes straight into registers, four steps ahead of their use --
//     no staging ring, so the shared memory goes to the histogram;
//   * which read a word belongs to comes from a bit set of read starts per chunk (one bit per word: a popcount of
//     the bits below the lane's gives the read, two loads its start and the next read's): ~10 instructions per
//     step instead of a search;
//   * histogram: the joint (score, code) x position table of the other kernels (key byte K = score << 2 | code,
//     u16 counters, even positions in the low half of a 32-bit column, odd ones in the high half) with rows of
//     `stride` 32-bit columns, a multiple of 32 so that the bank depends on the position only;
//   * a word that straddles two reads is counted as if it belonged to the first one (positions len .. len + 2) and
//     put right per read boundary afterwards (<= 3 bytes: subtract there, add at positions 0 .. 2 of the next read);
//   * -a: one 7-mer anchor per word is probed as in qb_period.cu (the anchor that starts at the PREVIOUS word: its
//     codes come from the lane below with one shuffle), the codes of (previous word, word) are kept as one u16 per
//     word for the confirmation, hits are one bit per step in a register, expanded with a warp prefix sum at the
//     end of the chunk and confirmed 32 at a time against the exact key set, one lane per hit.
// Read counts, the length histogram (shared memory) and the first-hit settling are per chunk, lanes over reads.
//
// No tensor cores: the path is an integer histogram.
#include "qb_dev.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace qb {

constexpr uint32_t kFMaxSteps = 32;    // 32 words per step: 1024 words = 4096 bytes per chunk at most (-a: one hit bit per step)
constexpr uint32_t kFLongSteps = 64;   // without -a a chunk may span 8192 bytes: half as many chunk set-ups per byte
constexpr uint32_t kFQueue = 64;
// Per-warp block (bytes), for chunks of at most kMR reads: 256 covers reads >= 16 bp (a chunk spans <= 4096 bytes),
// 128 reads >= 32 bp -- the smaller block lets the -a variant run 24 warps instead of 20.
template <uint32_t kMR, uint32_t kMS = kFMaxSteps>
struct FLay {
  static constexpr uint32_t sWords = kMS + 2u;                    // words of the bit set S (one bit per word of the chunk)
  static constexpr uint32_t oRoff = 0;                            // u32 roff[-1 .. kMR + 2]: read starts, chunk-relative
  static constexpr uint32_t oS = oRoff + (kMR + 4u) * 4u;         // u32 S[kMS + 2]: bit per word, set where a new read is current
  static constexpr uint32_t oSpre = oS + ((sWords * 4u + 15u) & ~15u);  // u32 Spre[34] (-a): popcount of the words below
  static constexpr uint32_t bytesNoAd = oSpre + 144u;
  static constexpr uint32_t oP = bytesNoAd;                       // -a: u8 P[-16 .. 1040): the four 2-bit codes of every word
  static constexpr uint32_t oFhit = oP + 16u + kFMaxSteps * 32u + 16u;  // -a: u32 fhit[kMR]
  static constexpr uint32_t oQ = oFhit + kMR * 4u;                // -a: u16 queue[kFQueue]
  static constexpr uint32_t bytesAd = oQ + kFQueue * 2u;
};

struct FArgs {
  const uint8_t *seq, *qual;   // byte 0 of the batch buffers (16-byte aligned)
  const uint32_t *offset;      // read starts (ascending, back to back)
  const uint32_t *chunk_first; // [n_chunks + 1] first read of every byte window (flat_index_kernel)
  Accum a;
  AdapterSet ad;
  FlatPlan plan;
  uint32_t n_reads, n_chunks;
  uint32_t base0;              // offset[0] (a multiple of 16)
  uint32_t batch_end;          // base0 + bytes of all reads
  uint32_t inc_lo, inc_hi;     // 1 and 65536 as arguments (plain ATOMS.ADD, see qb_period.cu)
};

// chunk_first[c] = first read whose start lies at or behind byte c * CB of the batch; [n_chunks] = n_reads
__global__ void flat_index_kernel(const uint32_t *offset, uint32_t n_reads, uint32_t base0, uint32_t cb, uint32_t n_chunks,
                                  uint32_t *chunk_first) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const uint32_t cur = (offset[r] - base0) / cb;
  const uint32_t lo = r ? (offset[r - 1] - base0) / cb + 1u : 0u;
  for (uint32_t c = lo; c <= cur && c <= n_chunks; c++) chunk_first[c] = r;
  if (r + 1u == n_reads)
    for (uint32_t c = cur + 1u; c <= n_chunks; c++) chunk_first[c] = n_reads;
}

__device__ __forceinline__ uint4 ldg_u128(const uint8_t *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void f_sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void f_sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void f_sts_u8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t f_lds_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
// shared address of the 32-bit column that holds positions 2 q and 2 q + 1
__device__ __forceinline__ uint32_t f_col(uint32_t hist_s, uint32_t q) { return hist_s + 4u * q; }
__device__ __forceinline__ uint32_t ldg_u32(const uint8_t *p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <bool kAd, int kW, uint32_t kMR, uint32_t kMS>
__global__ void __launch_bounds__(kW * 32, 1) flat_kernel(const __grid_constant__ FArgs args) {
  constexpr uint32_t kThreads = kW * 32;
  constexpr uint32_t kFull = 0xffffffffu;
  extern __shared__ __align__(128) uint8_t smem[];
  const FlatPlan &P = args.plan;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t smem_s = smem_u32(smem);
  auto gen = [&](uint32_t shared_addr) -> uint8_t * { return smem + (shared_addr - smem_s); };

  if (smem_s != P.smem_base) {  // the anchor map must sit at a 16 KiB-aligned shared address: fail loudly, count nothing
    if (tid == 0) atomicAdd(&args.a.counters[kCntError], 1ull);
    return;
  }
  const uint32_t hist_s = smem_s + P.hist_o, lenhist_s = smem_s + P.lenhist_o, kmerhist_s = smem_s + P.kmerhist_o;
  const uint32_t afilt_s = smem_s + P.afilt_o, exact_s = smem_s + P.exact_o;
  const uint32_t wb_s = smem_s + P.wblock_o + warp * P.wblock;
  using L = FLay<kMR, kMS>;
  static_assert(!kAd || kMS == kFMaxSteps, "the -a path keeps one hit bit per step in a 32-bit register");
  const uint32_t roff_s = wb_s + L::oRoff + 4u;  // roff[i] at roff_s + 4 i, i = -1 .. nr + 1
  const uint32_t S_s = wb_s + L::oS, Spre_s = wb_s + L::oSpre, P_s = wb_s + L::oP, fhit_s = wb_s + L::oFhit, q_s = wb_s + L::oQ;
  const uint32_t rowbytes = P.stride * 4u, max_len = P.max_len;

  auto clear_counters = [&]() {
    const uint4 z = make_uint4(0, 0, 0, 0);
    uint4 *h4 = reinterpret_cast<uint4 *>(gen(hist_s));
    for (uint32_t i = tid; i < kHistRows * rowbytes / 16u; i += kThreads) h4[i] = z;
    uint32_t *lh = reinterpret_cast<uint32_t *>(gen(lenhist_s));
    for (uint32_t i = tid; i <= max_len; i += kThreads) lh[i] = 0;
    if (kAd) {
      uint32_t *kh = reinterpret_cast<uint32_t *>(gen(kmerhist_s));
      for (uint32_t i = tid; i <= max_len + 1u; i += kThreads) kh[i] = 0;
    }
  };
  clear_counters();
  if (kAd) {
    uint32_t *af = reinterpret_cast<uint32_t *>(gen(afilt_s));  // bit-reversed words: see qb_period.cu
    for (uint32_t i = tid; i < kAnchorWords * kAnchorCopies; i += kThreads) af[i] = __brev(args.ad.anchor[i / kAnchorCopies]);
    uint32_t *ex = reinterpret_cast<uint32_t *>(gen(exact_s));
    if (args.ad.exact)
      for (uint32_t i = tid; i < kExactSlots; i += kThreads) ex[i] = args.ad.exact[i];
    uint32_t *fh = reinterpret_cast<uint32_t *>(gen(fhit_s));
    for (uint32_t i = lane; i < kMR; i += 32u) fh[i] = kNoHit;
  }
  __syncthreads();

  const KeyConsts kc(P.qbase);
  const uint32_t inc_lo = args.inc_lo, inc_hi = args.inc_hi;
  const uint32_t le_mask = 0xFFFFFFFFu >> (31u - lane);  // bits 0 .. lane
  const uint32_t afilt_or = afilt_s | ((lane >> 2) * 4u);
  const uint32_t G = gridDim.x * kW, g0 = blockIdx.x * kW + warp;
  const uint32_t n_chunks = args.n_chunks;
  const uint32_t rounds = (n_chunks + G - 1u) / G;  // the same for every warp of the grid
  unsigned long long n_reads_done = 0;
  long long n_invalid = 0;
  uint32_t to_flush = P.epoch;

  auto flush = [&]() {  // all warps are behind a barrier
    for (uint32_t pos = tid; pos < max_len; pos += kThreads) {
      const uint32_t base = f_col(hist_s, pos >> 1), sh = (pos & 1u) * 16u;
      unsigned long long *row = args.a.rows + (size_t)pos * kRow;
      uint32_t cc[4] = {0, 0, 0, 0};
      for (uint32_t sp = 0; sp < kScoreBins; sp++) {
        uint32_t tot = 0;
#pragma unroll
        for (uint32_t c = 0; c < 4u; c++) {
          const uint32_t v = (lds_u32(base + (sp << 2 | c) * rowbytes) >> sh) & 0xFFFFu;
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
      const uint32_t lc = lds_u32(lenhist_s + pos * 4u);
      if (lc) atomicAdd(&row[kColLength], (unsigned long long)lc);
      if (kAd) {
        const uint32_t kcnt = lds_u32(kmerhist_s + pos * 4u);
        if (kcnt) atomicAdd(&row[kColKmer], (unsigned long long)kcnt);
      }
    }
  };

  // Software pipeline over the rounds, so that a chunk never starts with a chain of dependent global loads:
  // chunk_first[] is read two rounds ahead, the chunk's first 32 read starts one round ahead (at the start of the round
  // in front), its first group of words at the end of that round's flat pass (behind them: the per-read phases).
  auto load_cf = [&](uint32_t cc, uint32_t &f0, uint32_t &f1) {
    f0 = f1 = 0;
    if (cc < n_chunks) f0 = args.chunk_first[cc], f1 = args.chunk_first[cc + 1u];
  };
  auto load_off = [&](uint32_t f0) -> uint32_t {  // start of read f0 + lane (the end of the batch behind the last read)
    const uint32_t gidx = f0 + lane;
    return gidx < args.n_reads ? args.offset[gidx] : args.batch_end;
  };
  uint32_t cf0, cf1, nf0, nf1;
  load_cf(g0, cf0, cf1);
  load_cf(g0 + G, nf0, nf1);
  uint32_t coff = load_off(cf0);
  uint32_t ns[4], nq[4];  // the group of four steps in flight
  bool have_grp = false;  // ns / nq hold the first group of this round's chunk
  for (uint32_t round = 0; round < rounds; round++) {
    const uint32_t c = round * G + g0;
    const uint32_t first = cf0, nr = cf1 - cf0;
    uint32_t ff0, ff1;
    load_cf(c + 2u * G, ff0, ff1);
    const uint32_t noff = load_off(nf0);
    auto issue_next_group = [&]() {
      have_grp = nf1 != nf0;  // (warp-uniform)
      if (have_grp) {
        const uint32_t ub0n = __shfl_sync(kFull, noff, 0) & ~3u;
        const uint8_t *ps = args.seq + ub0n + lane * 4u, *pq = args.qual + ub0n + lane * 4u;
#pragma unroll
        for (int u = 0; u < 4; u++) ns[u] = ldg_u32(ps + 128 * u), nq[u] = ldg_u32(pq + 128 * u);
      }
    };
    if (nr) {  // (warp-uniform)
      // ---- the chunk's reads: starts relative to the 32-bit boundary below the first one ----
      const uint32_t b_start = __shfl_sync(kFull, coff, 0);
      const uint32_t ub0 = b_start & ~3u;  // absolute byte of word 0
      __syncwarp();
      for (uint32_t i = lane; i <= nr + 1u; i += 32u) {
        const uint32_t gidx = first + i;
        const uint32_t o = i < 32u ? coff : gidx < args.n_reads ? args.offset[gidx] : args.batch_end;
        f_sts_u32(roff_s + 4u * i, gidx <= args.n_reads ? o - ub0 : 0xFFFFFF00u);
      }
      if (lane == 0) f_sts_u32(roff_s - 4u, 0u);
      if (kAd && lane == 0) f_sts_u8(P_s + 15u, 0u);  // codes of the word in front of the chunk (never inside one of its reads)
      for (uint32_t i = lane; i < L::sWords; i += 32u) f_sts_u32(S_s + 4u * i, 0u);
      __syncwarp();
      for (uint32_t i = lane; i <= nr; i += 32u) {  // bit w: from word w on, read i is the current one
        const uint32_t w = (lds_u32(roff_s + 4u * i) + 3u) >> 2;
        if (w < L::sWords * 32u) atomicOr(shared_ptr<uint32_t>(S_s) + (w >> 5), 1u << (w & 31u));
      }
      __syncwarp();
      // the words that straddle the first 32 read boundaries, for the fix-up behind the flat pass
      uint32_t bsw = 0, bqw = 0;
      if (lane < nr) {
        const uint32_t sb = lds_u32(roff_s + 4u * (lane + 1u));
        if (sb & 3u) bsw = ldg_u32(args.seq + ub0 + (sb & ~3u)), bqw = ldg_u32(args.qual + ub0 + (sb & ~3u));
      }
      const uint32_t end_rel = lds_u32(roff_s + 4u * nr);  // start of the read behind the chunk / end of the batch
      const uint32_t Wn = (end_rel + 3u) >> 2;             // words whose first byte lies in front of it
      const uint32_t nst = (Wn + 31u) >> 5;
      const bool last_has_next = first + nr < args.n_reads;  // does the read behind the chunk exist?
      if (kAd) {  // popcounts of the words below (read lookup at confirmation time): a warp prefix sum
        const uint32_t cw = (uint32_t)__popc(lds_u32(S_s + 4u * lane));
        uint32_t incl = cw;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t vv = __shfl_up_sync(kFull, incl, d);
          if (lane >= (uint32_t)d) incl += vv;
        }
        f_sts_u32(Spre_s + 4u * lane, incl - cw);
        if (lane == 31) {
          f_sts_u32(Spre_s + 4u * 32u, incl);
          f_sts_u32(Spre_s + 4u * 33u, incl + (uint32_t)__popc(lds_u32(S_s + 4u * 32u)));
        }
      }
      const uint32_t roff0 = lds_u32(roff_s);

      // ---- flat pass: lane <-> word, groups of four steps whose loads are issued one group ahead ----
      const uint8_t *gs = args.seq + ub0 + lane * 4u, *gq = args.qual + ub0 + lane * 4u;
      uint32_t run = 0;        // reads current in front of this step's words
      uint32_t hm = 0;         // -a: one hit bit per step, the newest in the low bits
      uint32_t carry = 0;      // -a: codes of the previous step's last word
      // (Loads are not predicated: a group reads up to 511 bytes behind the chunk's last word -- the next chunk's
      // bytes, or the padding behind the batch buffers (qb_api.cu: pad_bytes); such words are never counted.)
      if (!have_grp) {
#pragma unroll
        for (int u = 0; u < 4; u++) ns[u] = ldg_u32(gs + 128 * u), nq[u] = ldg_u32(gq + 128 * u);
      }
      for (uint32_t st0 = 0; st0 < nst; st0 += 4u) {
        uint32_t cs[4], cq[4];
#pragma unroll
        for (int u = 0; u < 4; u++) cs[u] = ns[u], cq[u] = nq[u];
        gs += 512, gq += 512;
        if (st0 + 4u < nst) {  // (warp-uniform)
#pragma unroll
          for (int u = 0; u < 4; u++) ns[u] = ldg_u32(gs + 128 * u), nq[u] = ldg_u32(gq + 128 * u);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint32_t st = st0 + (uint32_t)u;
          if (st < nst) {  // (warp-uniform)
            const uint32_t w = 32u * st + lane;
            const bool have = w < Wn;
            const uint32_t Sw = lds_u32(S_s + 4u * st);
            const int r = (int)(run + (uint32_t)__popc(Sw & le_mask)) - 1;  // read of the word's first byte (-1: the one in front)
            run += (uint32_t)__popc(Sw);
            const uint32_t B = 4u * w;
            const uint32_t rs = lds_u32(roff_s + 4u * (uint32_t)r);
            const uint32_t pos0 = B - rs;   // position of the word's first byte in read r
            const bool cnt = have && B >= roff0;  // the word in front of the first own read: scan only
            const uint32_t sw = cs[u], qw = cq[u];
            uint32_t nc, cd = 0, bad = 0;
            uint32_t K = kAd ? key_bytes_c(sw, qw, kc, nc, bad, cd) : key_bytes(sw, qw, kc, nc, bad);
            if ((bad & 0xC0C0C0C0u) && cnt) {  // a quality byte outside the counted window: the word byte by byte, exactly
              K = key_bytes_bad(nc);
              {
                const uint32_t e = lds_u32(roff_s + 4u * (uint32_t)r + 4u) - B;  // bytes of the word that belong to read r (>= 4: all)
                const bool next_ok = (uint32_t)(r + 1) < nr || last_has_next;  // read r + 1 exists
                for (uint32_t j = 0; j < 4u; j++) {
                  const bool in_b = j >= e;
                  if (in_b && !next_ok) continue;
                  const uint32_t p = in_b ? j - e : pos0 + j;
                  unsigned long long *row = args.a.rows + (size_t)p * kRow;
                  atomicAdd(&row[kColContent + base_code((sw >> (8u * j)) & 0xFFu)], 1ull);
                  const int sc = (int)((qw >> (8u * j)) & 0xFFu) - 33;
                  if (sc >= 0 && sc < 91)
                    atomicAdd(&row[sc], 1ull);
                  else
                    n_invalid++;
                }
              }
            }
            if (kAd) {
              const uint32_t gc = cd * 0x01041040u;  // top byte: the word's 4 codes
              uint32_t prev = __shfl_up_sync(kFull, gc, 1);
              if (lane == 0) prev = carry;
              carry = __shfl_sync(kFull, gc, 31);
              const uint32_t an = __byte_perm(prev, gc, 0x7773);  // bits 13:0 = the 7-mer that starts at the word in front
              if (have) f_sts_u8(P_s + 16u + w, gc >> 24);       // the word's codes, kept for the confirmation
              const uint32_t fw = lds_u32((an & 0x3FE0u) | afilt_or);
              hm = __funnelshift_l(__funnelshift_l(0u, fw, an), hm, 1);
              if (!have) hm &= ~1u;
            }
            const uint32_t odd = pos0 & 1u;
            const uint32_t A0 = hist_s + 4u * (pos0 >> 1), Ao = A0 + 4u * odd;
            const uint32_t i0 = odd ? inc_hi : inc_lo, i1 = odd ? inc_lo : inc_hi;
            if (cnt) {
              red_shared_add<0u>(__byte_perm(K, 0u, 0x4440) * rowbytes + A0, i0);
              red_shared_add<0u>(__byte_perm(K, 0u, 0x4441) * rowbytes + Ao, i1);
              red_shared_add<4u>(__byte_perm(K, 0u, 0x4442) * rowbytes + A0, i0);
              red_shared_add<4u>((K >> 24) * rowbytes + Ao, i1);
            }
          }
        }
      }
      __syncwarp();
      issue_next_group();

      // ---- per read boundary: the word that straddles it was counted for the read in front ----
      for (uint32_t i0 = 1; i0 <= nr; i0 += 32u) {
        const uint32_t i = i0 + lane;
        if (i <= nr) {
          const uint32_t s = lds_u32(roff_s + 4u * i), m = s & 3u;
          if (m) {
            const uint32_t wbyte = s - m;
            const uint32_t swd = i0 == 1u ? bsw : ldg_u32(args.seq + ub0 + wbyte);
            const uint32_t qwd = i0 == 1u ? bqw : ldg_u32(args.qual + ub0 + wbyte);
            if (!word_bad(qwd, kc.qsub)) {
              uint32_t ncx, badx = 0;
              const uint32_t Kw = key_bytes(swd, qwd, kc, ncx, badx);
              const uint32_t len_prev = s - lds_u32(roff_s + 4u * i - 4u);
              const bool exists = i < nr || last_has_next;
              for (uint32_t j = m; j < 4u; j++) {
                const uint32_t Kj = (Kw >> (8u * j)) & 0xFFu;
                const uint32_t po = len_prev + (j - m), pn = j - m;  // counted there, belongs here
                red_shared_add<0u>(Kj * rowbytes + f_col(hist_s, po >> 1), (po & 1u) ? 0u - inc_hi : 0u - inc_lo);
                if (exists) red_shared_add<0u>(Kj * rowbytes + f_col(hist_s, pn >> 1), (pn & 1u) ? inc_hi : inc_lo);
              }
            }
          }
        }
      }
      // ---- per read: length histogram (quack.c:219), read count (quack.c:220) ----
      for (uint32_t i0 = 0; i0 < nr; i0 += 32u) {
        const uint32_t i = i0 + lane;
        if (i < nr) {
          const uint32_t l = lds_u32(roff_s + 4u * i + 4u) - lds_u32(roff_s + 4u * i);
          if (l == 0u || l > max_len)
            atomicAdd(&args.a.counters[kCntError], 1ull);
          else
            red_shared_add<0u>(lenhist_s + 4u * (l - 1u), 1u);
        }
      }
      if (lane == 0) n_reads_done += nr;

      if (kAd) {
        // ---- -a: expand the hit bits into entries (bit << 5 | lane), confirm them 32 at a time ----
        uint32_t m = hm;
        while (__any_sync(kFull, m != 0u)) {
          const uint32_t cbits = (uint32_t)__popc(m);
          uint32_t incl = cbits;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t vv = __shfl_up_sync(kFull, incl, d);
            if (lane >= (uint32_t)d) incl += vv;
          }
          const uint32_t total = __shfl_sync(kFull, incl, 31);
          uint32_t off = incl - cbits;
          while (m && off < kFQueue) {
            const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1u;
            f_sts_u16(q_s + 2u * off, b << 5 | lane);
            off++;
          }
          __syncwarp();
          const uint32_t n = min(total, kFQueue);
          for (uint32_t e0 = 0; e0 < n; e0 += 32u) {
            if (e0 + lane < n) {
              const uint32_t ent = f_lds_u16(q_s + 2u * (e0 + lane));
              const uint32_t b = ent >> 5, ln = ent & 31u;
              const int w = (int)(32u * (nst - 1u - b) + ln) - 1;  // chunk-relative word whose 7-mer passed the filter
              // codes of words w - 1 .. w + 2 (bases 4 w - 4 .. 4 w + 11): the u16 of word w holds (w - 1, w), the high
              // bytes of the next two hold w + 1 and w + 2.  (Word w + 1 is the word of the lane that saw the hit and
              // always stored; the u16 of a word behind the chunk was never written -- no window that needs it lies
              // inside a read of the chunk.)
              const uint32_t pw = P_s + 16u + (uint32_t)(w - 1);  // bytes of words w - 1 .. w + 2: one unaligned 32-bit read
              const uint32_t ctx = __funnelshift_r(lds_u32(pw & ~3u), lds_u32((pw & ~3u) + 4u), (pw & 3u) * 8u);
              const int ws0 = 4 * w - 3;  // first byte of the first of the four windows
              // read of the first byte of the word that holds it (of word 0 if it lies in front of the chunk)
              const uint32_t vb = (uint32_t)(ws0 < 0 ? 0 : ws0) >> 2;
              int ra = (int)(lds_u32(Spre_s + 4u * (vb >> 5)) + (uint32_t)__popc(lds_u32(S_s + 4u * (vb >> 5)) & (0xFFFFFFFFu >> (31u - (vb & 31u))))) - 1;
              const uint32_t r1 = lds_u32(roff_s + 4u * (uint32_t)ra + 4u), r2 = lds_u32(roff_s + 4u * (uint32_t)ra + 8u);
              const bool in_next = ws0 >= 0 && (uint32_t)ws0 >= r1;  // (the word's first byte was still in the read in front)
              const uint32_t s_a = in_next ? r1 : lds_u32(roff_s + 4u * (uint32_t)ra);  // start of the read of byte ws0
              const uint32_t e_a = in_next ? r2 : r1;                                    // start of the read behind it
              if (in_next) ra++;
              const uint32_t e_b = lds_u32(roff_s + 4u * (uint32_t)ra + 8u);            // and of the one behind that
#pragma unroll
              for (uint32_t o = 1; o <= 4u; o++) {  // window o starts 4 - o bytes in front of the anchor
                const int ws = 4 * w - (int)(4u - o);
                if (ws < 0) continue;
                // a window lies in read ra or (behind a read start within these 3 bytes) in read ra + 1
                const bool nxt = (uint32_t)ws >= e_a;
                const int rr = ra + (nxt ? 1 : 0);
                const uint32_t rstart = nxt ? e_a : s_a, rend = nxt ? e_b : e_a;
                if (rr < 0 || (uint32_t)rr >= nr || (uint32_t)ws < rstart || (uint32_t)ws + 10u >= rend) continue;
                const uint32_t key = (ctx >> (2u * o)) & 0xFFFFFu;
                bool member;
                if (args.ad.exact)
                  member = lds_u32(exact_s + exact_off1(key)) == key || lds_u32(exact_s + exact_off2(key)) == key;
                else
                  member = (args.ad.bitmap[key >> 5] >> (key & 31u)) & 1u;
                if (member) atomicMin(shared_ptr<uint32_t>(fhit_s) + rr, (uint32_t)ws - rstart + 9u);
              }
            }
          }
          __syncwarp();
        }
        // first hits of the chunk's reads: kmer_count[p + 1]++ (quack.c:215-216)
        for (uint32_t i0 = 0; i0 < nr; i0 += 32u) {
          const uint32_t i = i0 + lane;
          if (i < nr) {
            const uint32_t f = lds_u32(fhit_s + 4u * i);
            if (f != kNoHit) {
              red_shared_add<0u>(kmerhist_s + (f + 1u) * 4u, 1u);
              f_sts_u32(fhit_s + 4u * i, kNoHit);
            }
          }
        }
        __syncwarp();
      }
    } else {
      issue_next_group();
    }
    cf0 = nf0, cf1 = nf1, nf0 = ff0, nf1 = ff1, coff = noff;
    if (--to_flush == 0u && round + 1u < rounds) {  // u16 counters: flush before any bin can wrap
      to_flush = P.epoch;
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
  if (lane == 0) {
    if (n_invalid) atomicAdd(&args.a.counters[kCntInvalidQual], (unsigned long long)n_invalid);
    if (n_reads_done) atomicAdd(&args.a.counters[kCntReads], n_reads_done);
  }
}

// ------------------------------------------------------------------------------------------
// plan and launch
// ------------------------------------------------------------------------------------------

// warps per CTA: without -a a warp block is 1.4 KiB and the kernel needs 68 registers -> 24 warps; with -a the packed
// codes (one byte per word), first hits and queue make it 3.5 KiB -> 20 warps, or 2.5 KiB -> 24 warps when every read
// has >= 32 bp (chunks of <= 128 reads).  The -a variant is latency-bound (issue slots 59 % busy at 16 warps, 70 % at
// 20): warps matter.
#ifndef QB_FW
#define QB_FW 24
#endif
constexpr int kFlatWarpsNoAd = QB_FW, kFlatWarpsAd = 20, kFlatWarpsAdShort = 24;

FlatPlan flat_plan(uint32_t batch_max_len, uint32_t batch_min_len, int adapters, int sm_count, uint32_t smem_optin,
                   uint32_t smem_reserved, uint32_t qbase) {
  FlatPlan p;
  memset(&p, 0, sizeof p);
  if (batch_min_len < 16u || batch_max_len < batch_min_len || batch_max_len > 320u) return p;
  p.max_len = batch_max_len;
  p.qbase = qbase;
  const uint32_t pairs = (batch_max_len + 3u + 1u) / 2u + 1u;  // positions 0 .. max_len + 2 (straddling words)
  p.stride = (pairs + 31u) & ~31u;
  p.smem_base = smem_reserved;
  uint32_t o = 0;
  auto take = [&](uint32_t bytes) {
    const uint32_t at = o;
    o += (bytes + 127u) & ~127u;
    return at;
  };
  p.hist_o = take(kHistRows * p.stride * 4u);
  p.lenhist_o = take((batch_max_len + 1u) * 4u);
  p.kmerhist_o = take(adapters ? (batch_max_len + 2u) * 4u : 0u);
  if (adapters) {  // the probe address is `anchor bits | base`: the SHARED ADDRESS of the map is 16 KiB-aligned
    o = ((smem_reserved + o + kAnchorSmemBytes - 1u) & ~(kAnchorSmemBytes - 1u)) - smem_reserved;
    p.afilt_o = take(kAnchorSmemBytes);
  }
  p.exact_o = take(adapters ? kExactSlots * 4u : 0u);
  // a chunk spans its byte window plus the tail of its last read, in 32 steps of 128 bytes at most
  // (without -a: 64 steps when the reads are long enough for <= 256 of them per chunk)
  p.max_steps = kFMaxSteps;
  if (!adapters && (kFLongSteps * 128u - batch_max_len - 8u) / batch_min_len + 4u <= 256u && !getenv("QB_FLAT_STEPS32"))
    p.max_steps = kFLongSteps;
  uint32_t cb = (p.max_steps * 128u - batch_max_len - 8u) & ~63u;
  if (const char *e = getenv("QB_FLAT_CHUNK")) {  // tuning hook
    const uint32_t v = (uint32_t)atoi(e);
    if (v >= 512u && v < cb) cb = v;
  }
  p.chunk_bytes = cb;
  const uint32_t reads_per_chunk = cb / batch_min_len + 2u;
  if (reads_per_chunk + 2u > 256u) return p;
  p.max_reads = reads_per_chunk + 2u <= 128u && !getenv("QB_FLAT_MR256") ? 128u : 256u;
  uint32_t warps = (uint32_t)kFlatWarpsNoAd;
  if (adapters) warps = p.max_reads == 128u ? (uint32_t)kFlatWarpsAdShort : (uint32_t)kFlatWarpsAd;
  p.warps = warps;
  p.wblock = adapters ? (p.max_reads == 128u ? FLay<128>::bytesAd : FLay<256>::bytesAd)
             : p.max_steps == kFLongSteps ? FLay<256, kFLongSteps>::bytesNoAd
                                          : (p.max_reads == 128u ? FLay<128>::bytesNoAd : FLay<256>::bytesNoAd);
  if (p.max_steps == kFLongSteps) p.max_reads = 256u;
  p.wblock = (p.wblock + 127u) & ~127u;
  p.wblock_o = take(p.wblock * warps);
  p.smem_bytes = o;
  if (p.smem_bytes > smem_optin) {  // (the 24-warp -a variant does not fit every max_len: 20 warps then)
    if (!(adapters && p.max_reads == 128u)) return p;
    o = p.wblock_o;
    p.max_reads = 256u, p.warps = warps = (uint32_t)kFlatWarpsAd;
    p.wblock = (FLay<256>::bytesAd + 127u) & ~127u;
    p.wblock_o = take(p.wblock * warps);
    p.smem_bytes = o;
    if (p.smem_bytes > smem_optin) return p;
  }
  p.epoch = 30000u / (warps * reads_per_chunk);
  if (p.epoch == 0) p.epoch = 1;
  p.grid = (uint32_t)sm_count;
  p.ok = 1;
  return p;
}

cudaError_t flat_configure() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(flat_kernel<false, kFlatWarpsNoAd, 256, kFLongSteps>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
  if ((e = cudaFuncSetAttribute(flat_kernel<false, kFlatWarpsNoAd, 128, kFMaxSteps>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
  if ((e = cudaFuncSetAttribute(flat_kernel<false, kFlatWarpsNoAd, 256, kFMaxSteps>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
  if ((e = cudaFuncSetAttribute(flat_kernel<true, kFlatWarpsAdShort, 128, kFMaxSteps>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
  return cudaFuncSetAttribute(flat_kernel<true, kFlatWarpsAd, 256, kFMaxSteps>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
}

cudaError_t launch_flat(const BatchView &b, const Accum &a, const AdapterSet &ad, const FlatPlan &plan, cudaStream_t stream) {
  if (b.n_reads == 0) return cudaSuccess;
  if (!plan.ok || !b.contig_min_len || (b.first_offset & 15u) || !b.tiles) return cudaErrorInvalidValue;
  FArgs args;
  memset(&args, 0, sizeof args);
  args.seq = b.seq;
  args.qual = b.qual;
  args.offset = b.offset;
  args.a = a;
  args.ad = ad;
  args.plan = plan;
  args.n_reads = b.n_reads;
  args.base0 = b.first_offset;
  args.batch_end = b.first_offset + (uint32_t)b.n_bytes;
  args.n_chunks = ((uint32_t)b.n_bytes + plan.chunk_bytes - 1u) / plan.chunk_bytes;
  args.inc_lo = 1u, args.inc_hi = 0x10000u;
  uint32_t *chunk_first = reinterpret_cast<uint32_t *>(b.tiles);  // scratch of 2 x n_reads words: n_chunks + 1 <= n_reads + 1
  args.chunk_first = chunk_first;
  flat_index_kernel<<<(b.n_reads + 255u) / 256u, 256, 0, stream>>>(b.offset, b.n_reads, args.base0, plan.chunk_bytes,
                                                                  args.n_chunks, chunk_first);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const uint32_t warps = plan.warps;
  uint32_t grid = (args.n_chunks + warps - 1u) / warps;
  if (grid > plan.grid) grid = plan.grid;
  if (grid == 0) grid = 1;
  if (const char *g = getenv("QB_FUSED_GRID")) {  // test hook: few CTAs exercise the u16 flush path
    const uint32_t v = (uint32_t)atoi(g);
    if (v >= 1 && v < grid) grid = v;
  }
  if (ad.enabled && plan.max_reads == 128u)
    flat_kernel<true, kFlatWarpsAdShort, 128, kFMaxSteps><<<grid, kFlatWarpsAdShort * 32, plan.smem_bytes, stream>>>(args);
  else if (ad.enabled)
    flat_kernel<true, kFlatWarpsAd, 256, kFMaxSteps><<<grid, kFlatWarpsAd * 32, plan.smem_bytes, stream>>>(args);
  else if (plan.max_steps == kFLongSteps)
    flat_kernel<false, kFlatWarpsNoAd, 256, kFLongSteps><<<grid, kFlatWarpsNoAd * 32, plan.smem_bytes, stream>>>(args);
  else if (plan.max_reads == 128u)
    flat_kernel<false, kFlatWarpsNoAd, 128, kFMaxSteps><<<grid, kFlatWarpsNoAd * 32, plan.smem_bytes, stream>>>(args);
  else
    flat_kernel<false, kFlatWarpsNoAd, 256, kFMaxSteps><<<grid, kFlatWarpsNoAd * 32, plan.smem_bytes, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace qb
