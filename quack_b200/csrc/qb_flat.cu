// qb_flat.cu -- flat kernel (v6) for quack's per-read statistics accumulation (reference: the while loop of
// read_fastq(), quack.c:193-221) on RAGGED batches: reads of any lengths in [16, 304] lying back to back (what
// the host reader packs; config 4, trimmed reads of 35..300 bp).  The v3 kernel gives such batches one warp per
// read, lane <-> 4 positions: a 167-bp read fills 42 of 64 lane slots and pays its per-read set-up again and again
// (0.31 / 0.15 of the HBM roofline without / with -a).  Here nothing is per read in the hot loop:
//   * the batch is cut into CHUNKS of whole reads by byte windows (chunk c = the reads that START in bytes
//     [c CB, (c + 1) CB); a one-pass index kernel finds each window's first read), one warp per chunk at a time;
//   * a warp walks its chunk flat, lane <-> aligned 16-byte unit: two coalesced 16-byte vector loads per lane
//     (512 contiguous bytes of bases and of quality bytes per warp instruction) straight into registers -- no
//     staging ring, so the shared memory goes to the histogram and 16-24 warps fit;
//   * which read a unit belongs to comes from a bit set of read starts per chunk (one bit per unit: a popcount of
//     the bits below the lane's gives the read, two loads its start and the next read's): ~10 instructions per
//     16 bases instead of a search per word; a word inside the unit switches to the next read by one compare;
//   * histogram: the joint (score, code) x position table of the other kernels (key byte K = score << 2 | code,
//     u16 counters, even positions in the low half of a 32-bit column, odd ones in the high half) with rows of
//     `stride` 32-bit columns, a multiple of 32 so that the bank depends on the position only, and one spare column
//     per 32 (column = pair + pair / 32): lanes are 16 bases = 8 columns apart, which would fall into 4 banks; the
//     skew spreads 32 lanes of one read over 32 banks;
//   * a word that straddles two reads is counted as if it belonged to the first one (positions len .. len + 2) and
//     put right per read boundary afterwards (<= 3 bytes: subtract there, add at positions 0 .. 2 of the next read);
//   * -a: the 2-bit codes of a unit are packed into one 32-bit value (kept in a per-chunk array for the
//     confirmation), four 7-mer anchors per unit are probed (the one that starts in the previous unit's last word
//     with one shuffle), hits are one bit per anchor in a register, expanded with a warp prefix sum at the end of
//     the chunk and confirmed 32 at a time against the exact key set, one lane per hit (as in qb_period.cu).
// Read counts, the length histogram (shared memory) and the first-hit settling are per chunk, lanes over reads.
//
// No tensor cores: the path is an integer histogram.
#include "qb_dev.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace qb {

constexpr uint32_t kFMaxReads = 256;   // reads per chunk (reads >= 16 bp, a chunk spans <= 4096 bytes)
constexpr uint32_t kFMaxIter = 8;      // 32 units of 16 bytes per iteration
constexpr uint32_t kFQueue = 64;
// per-warp block (bytes, multiples of 16)
constexpr uint32_t kFoRoff = 0;                              // u32 roff[-1 .. kFMaxReads + 1]: read starts, chunk-relative
constexpr uint32_t kFoS = kFoRoff + (kFMaxReads + 4u) * 4u;  // u32 S[9]: bit per unit, set where a new read is current
constexpr uint32_t kFoSpre = kFoS + 48u;                     // u32 Spre[9]: popcount of the words below
constexpr uint32_t kFWarpBytesNoAd = kFoSpre + 48u;
constexpr uint32_t kFoP = kFWarpBytesNoAd;                   // -a: u8 P[4 + words of the chunk + 8]: packed 2-bit codes
constexpr uint32_t kFoFhit = kFoP + 4u + kFMaxIter * 32u * 4u + 12u;  // -a: u32 fhit[kFMaxReads]
constexpr uint32_t kFoQ = kFoFhit + kFMaxReads * 4u;         // -a: u16 queue[kFQueue]
constexpr uint32_t kFWarpBytesAd = kFoQ + kFQueue * 2u;

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
__device__ __forceinline__ uint32_t f_lds_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
// shared address of the 32-bit column that holds positions 2 q and 2 q + 1: one spare column per 32
__device__ __forceinline__ uint32_t f_col(uint32_t hist_s, uint32_t q) { return hist_s + 4u * (q + (q >> 5)); }

template <bool kAd, int kW>
__global__ void __launch_bounds__(kW * 32, 1) flat_kernel(const __grid_constant__ FArgs args) {
  constexpr uint32_t kThreads = kW * 32;
  constexpr uint32_t kFull = 0xffffffffu;
  extern __shared__ __align__(128) uint8_t smem[];
  const FlatPlan &P = args.plan;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t smem_s = smem_u32(smem);
  auto gen = [&](uint32_t shared_addr) -> uint8_t * { return smem + (shared_addr - smem_s); };

  const uint32_t hist_s = smem_s + P.hist_o, lenhist_s = smem_s + P.lenhist_o, kmerhist_s = smem_s + P.kmerhist_o;
  const uint32_t afilt_s = smem_s + P.afilt_o, exact_s = smem_s + P.exact_o;
  const uint32_t wb_s = smem_s + P.wblock_o + warp * P.wblock;
  const uint32_t roff_s = wb_s + kFoRoff + 4u;  // roff[i] at roff_s + 4 i, i = -1 .. nr + 1
  const uint32_t S_s = wb_s + kFoS, Spre_s = wb_s + kFoSpre, P_s = wb_s + kFoP, fhit_s = wb_s + kFoFhit, q_s = wb_s + kFoQ;
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
    for (uint32_t i = lane; i < kFMaxReads; i += 32u) fh[i] = kNoHit;
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

  for (uint32_t round = 0; round < rounds; round++) {
    const uint32_t c = round * G + g0;
    uint32_t first = 0, nr = 0;
    if (c < n_chunks) {
      first = args.chunk_first[c];
      nr = args.chunk_first[c + 1u] - first;
    }
    if (nr) {  // (warp-uniform)
      // ---- the chunk's reads: starts relative to the 16-byte boundary below the first one ----
      const uint32_t b_start = args.offset[first];
      const uint32_t ub0 = b_start & ~15u;  // absolute byte of unit 0
      __syncwarp();
      for (uint32_t i = lane; i <= nr + 1u; i += 32u) {
        const uint32_t gidx = first + i;
        const uint32_t o = gidx < args.n_reads ? args.offset[gidx] : (gidx == args.n_reads ? args.batch_end : 0xFFFFFFF0u);
        f_sts_u32(roff_s + 4u * i, gidx <= args.n_reads ? o - ub0 : 0xFFFFFF00u);
      }
      if (lane == 0) f_sts_u32(roff_s - 4u, 0u);
      if (lane < 9u) f_sts_u32(S_s + 4u * lane, 0u);
      __syncwarp();
      for (uint32_t i = lane; i <= nr; i += 32u) {  // bit v: from unit v on, read i is the current one
        const uint32_t v = (lds_u32(roff_s + 4u * i) + 15u) >> 4;
        if (v < 9u * 32u) atomicOr(shared_ptr<uint32_t>(S_s) + (v >> 5), 1u << (v & 31u));
      }
      __syncwarp();
      const uint32_t end_rel = lds_u32(roff_s + 4u * nr);  // start of the read behind the chunk / end of the batch
      const uint32_t V = (end_rel + 15u) >> 4;             // units whose first byte lies in front of it
      const uint32_t nit = (V + 31u) >> 5;
      const bool last_has_next = first + nr < args.n_reads;  // does the read behind the chunk exist?
      if (kAd && lane < 9u) {  // popcounts of the words below (read lookup at confirmation time)
        uint32_t acc = 0;
        for (uint32_t j = 0; j < lane; j++) acc += (uint32_t)__popc(lds_u32(S_s + 4u * j));
        f_sts_u32(Spre_s + 4u * lane, acc);
      }
      const uint32_t roff0 = lds_u32(roff_s);

      // ---- flat pass: lane <-> 16-byte unit ----
      const uint8_t *gs = args.seq + ub0 + lane * 16u, *gq = args.qual + ub0 + lane * 16u;
      uint32_t run = 0;        // reads current in front of this iteration's units
      uint32_t hm = 0;         // -a: 4 hit bits per iteration, the newest in the low bits
      uint32_t carry = 0;      // -a: codes of the previous iteration's last unit
      for (uint32_t it = 0; it < nit; it++, gs += 512, gq += 512) {
        const uint32_t v = 32u * it + lane;
        const bool have = v < V;
        uint4 s4 = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u), q4 = make_uint4(kc.qsub, kc.qsub, kc.qsub, kc.qsub);
        if (have) {
          s4 = ldg_u128(gs);
          q4 = ldg_u128(gq);
        }
        const uint32_t Sw = lds_u32(S_s + 4u * it);
        const int r = (int)(run + (uint32_t)__popc(Sw & le_mask)) - 1;  // read of the unit's first byte (-1: the one in front)
        run += (uint32_t)__popc(Sw);
        const uint32_t B = 16u * v;
        const uint32_t rs = lds_u32(roff_s + 4u * (uint32_t)r), rn = lds_u32(roff_s + 4u * (uint32_t)r + 4u);
        const uint32_t pos0 = B - rs;   // position of the unit's first byte in read r
        const uint32_t e = rn - B;      // bytes of the unit that belong to read r (>= 16: all)
        const bool cnt = have && B >= roff0;                       // units in front of the first own read: scan only
        const bool next_ok = (uint32_t)(r + 1) < nr || last_has_next;  // read r + 1 exists
        const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w}, qw[4] = {q4.x, q4.y, q4.z, q4.w};
        uint32_t K[4], nc[4], cd[4], bad = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) K[k] = kAd ? key_bytes_c(sw[k], qw[k], kc, nc[k], bad, cd[k]) : key_bytes(sw[k], qw[k], kc, nc[k], bad);
        if (bad & 0xC0C0C0C0u) {  // a quality byte outside the counted window: those words byte by byte, exactly
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (word_bad(qw[k], kc.qsub)) {
              K[k] = key_bytes_bad(nc[k]);
              if (cnt)
                for (uint32_t j = 0; j < 4u; j++) {
                  const uint32_t ob = 4u * k + j;
                  const bool in_b = ob >= e;
                  if (in_b && !next_ok) continue;
                  const uint32_t p = in_b ? ob - e : pos0 + ob;
                  unsigned long long *row = args.a.rows + (size_t)p * kRow;
                  atomicAdd(&row[kColContent + base_code((sw[k] >> (8u * j)) & 0xFFu)], 1ull);
                  const int sc = (int)((qw[k] >> (8u * j)) & 0xFFu) - 33;
                  if (sc >= 0 && sc < 91)
                    atomicAdd(&row[sc], 1ull);
                  else
                    n_invalid++;
                }
            }
        }
        if (kAd) {
          // the unit's 16 codes as one value (first base in the low bits), kept for the confirmation
          const uint32_t g0c = cd[0] * 0x01041040u, g1c = cd[1] * 0x01041040u, g2c = cd[2] * 0x01041040u, g3c = cd[3] * 0x01041040u;
          const uint32_t P32 = __byte_perm(__byte_perm(g0c, g1c, 0x0073), __byte_perm(g2c, g3c, 0x0073), 0x5410);
          if (have) f_sts_u32(P_s + 4u + 4u * v, P32);
          uint32_t prev = __shfl_up_sync(kFull, P32, 1);
          if (lane == 0) prev = carry;
          carry = __shfl_sync(kFull, P32, 31);
          // anchors: the 7-mer that starts at the previous unit's last word, then at words 0, 1, 2 of this unit
          const uint32_t an[4] = {__funnelshift_r(prev, P32, 24), P32, P32 >> 8, P32 >> 16};
#pragma unroll
          for (int t = 0; t < 4; t++) {
            const uint32_t fw = lds_u32((an[t] & 0x3FE0u) | afilt_or);
            hm = __funnelshift_l(__funnelshift_l(0u, fw, an[t]), hm, 1);
          }
          if (!have) hm &= ~15u;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const bool use_b = 4u * (uint32_t)k >= e;  // the word starts in read r + 1
          const uint32_t pos = use_b ? 4u * (uint32_t)k - e : pos0 + 4u * (uint32_t)k;
          const uint32_t q = pos >> 1, t = q & 31u;
          const bool odd = pos & 1u;
          const uint32_t A0 = f_col(hist_s, q);
          const uint32_t A1 = A0 + (t == 31u ? 8u : 4u);
          const uint32_t A2 = A1 + (t == 30u ? 8u : 4u);
          const uint32_t X1 = odd ? A1 : A0, X3 = odd ? A2 : A1;
          const uint32_t i0 = odd ? inc_hi : inc_lo, i1 = odd ? inc_lo : inc_hi;
          if (cnt && (!use_b || next_ok)) {
            red_shared_add<0u>(__byte_perm(K[k], 0u, 0x4440) * rowbytes + A0, i0);
            red_shared_add<0u>(__byte_perm(K[k], 0u, 0x4441) * rowbytes + X1, i1);
            red_shared_add<0u>(__byte_perm(K[k], 0u, 0x4442) * rowbytes + A1, i0);
            red_shared_add<0u>((K[k] >> 24) * rowbytes + X3, i1);
          }
        }
      }
      __syncwarp();

      // ---- per read boundary: the word that straddles it was counted for the read in front ----
      for (uint32_t i0 = 1; i0 <= nr; i0 += 32u) {
        const uint32_t i = i0 + lane;
        if (i <= nr) {
          const uint32_t s = lds_u32(roff_s + 4u * i), m = s & 3u;
          if (m) {
            const uint32_t wbyte = s - m;
            const uint32_t swd = __ldg(reinterpret_cast<const uint32_t *>(args.seq + ub0 + wbyte));
            const uint32_t qwd = __ldg(reinterpret_cast<const uint32_t *>(args.qual + ub0 + wbyte));
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
              const uint32_t it = nit - 1u - (b >> 2), sub = 3u - (b & 3u);
              const int w = (int)(4u * (32u * it + ln) + sub) - 1;  // chunk-relative word whose 7-mer passed the filter
              const uint32_t a = P_s + 3u + (uint32_t)w;            // code byte of word w - 1
              const uint32_t a4 = a & ~3u;
              const uint32_t ctx = __funnelshift_r(lds_u32(a4), lds_u32(a4 + 4u), (a & 3u) * 8u);  // bases 4 w - 4 .. 4 w + 11
              const int ws0 = 4 * w - 3;  // first byte of the first of the four windows
              // read of that byte (of byte 0 if it lies in front of the chunk's units)
              const uint32_t vb = (uint32_t)(ws0 < 0 ? 0 : ws0) >> 4;
              int ra = (int)(lds_u32(Spre_s + 4u * (vb >> 5)) + (uint32_t)__popc(lds_u32(S_s + 4u * (vb >> 5)) & (0xFFFFFFFFu >> (31u - (vb & 31u))))) - 1;
              const uint32_t r1 = lds_u32(roff_s + 4u * (uint32_t)ra + 4u), r2 = lds_u32(roff_s + 4u * (uint32_t)ra + 8u);
              if (ws0 >= 0 && (uint32_t)ws0 >= r1) ra++;  // (the unit's first byte was still in the read in front)
              const uint32_t s_a = ws0 >= 0 && (uint32_t)ws0 >= r1 ? r1 : lds_u32(roff_s + 4u * (uint32_t)ra);  // start of read ra
              const uint32_t e_a = ws0 >= 0 && (uint32_t)ws0 >= r1 ? r2 : r1;                                    // start of read ra + 1
              const uint32_t e_b = lds_u32(roff_s + 4u * (uint32_t)ra + 8u);                                     // start of read ra + 2
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
    }
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

#ifndef QB_FW
#define QB_FW 16
#endif
constexpr int kFlatWarps = QB_FW;

FlatPlan flat_plan(uint32_t batch_max_len, uint32_t batch_min_len, int adapters, int sm_count, uint32_t smem_optin,
                   uint32_t qbase) {
  FlatPlan p;
  memset(&p, 0, sizeof p);
  if (batch_min_len < 16u || batch_max_len < batch_min_len || batch_max_len > 320u) return p;
  p.max_len = batch_max_len;
  p.qbase = qbase;
  const uint32_t pairs = (batch_max_len + 3u + 1u) / 2u;       // positions 0 .. max_len + 2 (straddling words)
  const uint32_t ncols = pairs + (pairs >> 5) + 1u;
  p.stride = (ncols + 31u) & ~31u;
  uint32_t o = 0;
  auto take = [&](uint32_t bytes) {
    const uint32_t at = o;
    o += (bytes + 127u) & ~127u;
    return at;
  };
  p.hist_o = take(kHistRows * p.stride * 4u);
  p.lenhist_o = take((batch_max_len + 1u) * 4u);
  p.kmerhist_o = take(adapters ? (batch_max_len + 2u) * 4u : 0u);
  p.afilt_o = take(adapters ? kAnchorSmemBytes : 0u);
  if (adapters && (p.afilt_o & (kAnchorSmemBytes - 1u))) {  // the probe address is `anchor bits | base`: 16 KiB-aligned
    o = (p.afilt_o + kAnchorSmemBytes - 1u) & ~(kAnchorSmemBytes - 1u);
    p.afilt_o = take(kAnchorSmemBytes);
  }
  p.exact_o = take(adapters ? kExactSlots * 4u : 0u);
  p.wblock = adapters ? kFWarpBytesAd : kFWarpBytesNoAd;
  p.wblock = (p.wblock + 127u) & ~127u;
  p.wblock_o = take(p.wblock * (uint32_t)kFlatWarps);
  p.smem_bytes = o;
  if (p.smem_bytes > smem_optin) return p;
  // a chunk spans its byte window plus the tail of its last read, in 8 iterations of 512 bytes at most
  uint32_t cb = kFMaxIter * 512u - ((batch_max_len + 32u + 15u) & ~15u);
  if (const char *e = getenv("QB_FLAT_CHUNK")) {  // tuning hook
    const uint32_t v = (uint32_t)atoi(e);
    if (v >= 512u && v < cb) cb = v;
  }
  p.chunk_bytes = cb;
  const uint32_t reads_per_chunk = cb / batch_min_len + 2u;
  if (reads_per_chunk + 2u > kFMaxReads) return p;
  p.epoch = 30000u / ((uint32_t)kFlatWarps * reads_per_chunk);
  if (p.epoch == 0) p.epoch = 1;
  p.grid = (uint32_t)sm_count;
  p.ok = 1;
  return p;
}

cudaError_t flat_configure() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(flat_kernel<false, kFlatWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448))) return e;
  return cudaFuncSetAttribute(flat_kernel<true, kFlatWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
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
  uint32_t grid = (args.n_chunks + (uint32_t)kFlatWarps - 1u) / (uint32_t)kFlatWarps;
  if (grid > plan.grid) grid = plan.grid;
  if (grid == 0) grid = 1;
  if (const char *g = getenv("QB_FUSED_GRID")) {  // test hook: few CTAs exercise the u16 flush path
    const uint32_t v = (uint32_t)atoi(g);
    if (v >= 1 && v < grid) grid = v;
  }
  if (ad.enabled)
    flat_kernel<true, kFlatWarps><<<grid, kFlatWarps * 32, plan.smem_bytes, stream>>>(args);
  else
    flat_kernel<false, kFlatWarps><<<grid, kFlatWarps * 32, plan.smem_bytes, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace qb
