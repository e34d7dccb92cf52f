// qb_dev.cuh -- device helpers shared by the sm_100a kernels (qb_kernels.cu: simple + fused v3,
// qb_period.cu: period kernel, qb_flat.cu: flat kernel): base LUT, SWAR key bytes, mbarrier / TMA bulk copy / shared-memory
// access wrappers.  Internal; the public boundary is include/quack_b200.h.
#pragma once

#include "qb_kernels.cuh"

namespace qb {

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

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar_s) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar_s, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "QB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra QB_DONE;\n"
      "bra QB_WAIT;\n"
      "QB_DONE:\n"
      "}\n" ::"r"(bar_s),
      "r"(parity)
      : "memory");
}
// Producer-side wait: the producer lane has nothing else to do, so it must not burn issue slots of the
// consumer warps that share its scheduler: long hardware suspend hint plus a sleep between polls.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar_s, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar_s), "r"(parity), "r"(2000u)
        : "memory");
    if (done) break;
    __nanosleep(400);
  }
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the same two with raw shared-space addresses (no generic -> shared conversion at the call site)
__device__ __forceinline__ void mbar_arrive_expect_tx_s(uint32_t bar_s, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_s(uint32_t dst_s, const void *src, uint32_t bytes, uint32_t bar_s) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s),
               "l"(src), "r"(bytes), "r"(bar_s)
               : "memory");
}
// 16-byte asynchronous copy global -> shared that bypasses L1 and the register file (SASS: LDGSTS.E.BYPASS.128);
// completion through cp.async.commit_group / wait_group of the issuing thread
__device__ __forceinline__ void cp_async16(uint32_t dst_s, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_s), "l"(src) : "memory");
}
// two of them (bases and quality bytes of a tile) under one predicate, at a compile-time offset from the operands
template <int kOff>
__device__ __forceinline__ void cp_async16_pair(bool p, uint32_t dst_a, uint32_t dst_b, const void *src_a, const void *src_b) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.u32 q, %0, 0;\n"
      "@q cp.async.cg.shared.global [%1+%5], [%3+%5], 16;\n"
      "@q cp.async.cg.shared.global [%2+%5], [%4+%5], 16;\n"
      "}\n" ::"r"((uint32_t)p),
      "r"(dst_a), "r"(dst_b), "l"(src_a), "l"(src_b), "n"(kOff)
      : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32_at(uint32_t off, uint32_t base) {  // [base + off], base uniform
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + off));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
template <uint32_t kOff>
__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(kOff) : "memory");
}

// Histogram rows: key byte K = s << 2 | code, s = q - qbase.  Scores s in [0, kScoreBins) count in shared
// memory (Phred 0..46 with the default qbase 33: every Illumina/ONT/PacBio scale in use); row s =
// kScoreBins is the dummy that re-keyed words land in; anything above takes the exact global path.  Putting
// the score in the high bits keeps the unused scores at the END of the row space, so they cost no memory.
constexpr uint32_t kScoreBins = 47;
constexpr uint32_t kHistRows = (kScoreBins + 1u) * 4u;  // 192 rows of Lh packed u16 pairs

// three-input logic op with an explicit truth table (a = 0xF0, b = 0xCC, c = 0xAA); constants passed
// as operands stay in registers, so e.g. (x & A) ^ B is ONE LOP3 instead of two
template <int kLut>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(kLut));
  return d;
}

// loop-invariant SWAR constants
struct KeyConsts {
  uint32_t m5b, x43, m1f, x07, x14, a7f, a3f, m80, m03, x11, qsub;
  __device__ __forceinline__ explicit KeyConsts(uint32_t qbase)
      : m5b(0x5B5B5B5Bu), x43(0x43434343u), m1f(0x1F1F1F1Fu), x07(0x07070707u), x14(0x14141414u),
        a7f(0x7F7F7F7Fu), a3f(0x3F3F3F3Fu), m80(0x80808080u), m03(0x03030303u),
        x11((64u - kScoreBins) * 0x01010101u), qsub(qbase * 0x01010101u) {}
};

// Phase A arithmetic for one aligned word of 4 bases + 4 quality bytes.  Returns the 4 key bytes
//   K = s << 2 | code,   code = A0 T1 C2 G3 (quack.c:150),   s = q - qbase in [0, kScoreBins)
// `nc` gets the inverted codes in bits 7:6 of each byte (other bits undefined); `bad` accumulates
// qs | (qs + 64 - kScoreBins), whose bits 7:6 are non-zero iff some quality byte is outside the window:
// then the caller re-keys the word to the dummy row and counts its 4 bases exactly.
__device__ __forceinline__ uint32_t key_bytes(uint32_t sw, uint32_t qw, const KeyConsts &c, uint32_t &nc,
                                              uint32_t &bad) {
  // per byte: bit7 of n_cg is 0 iff (b & 0x5B) == 0x43; bit6 of n_g / n_t is 0 iff (b & 0x1F) == 7 / 0x14
  const uint32_t n_cg = lop3<0x6A>(sw, c.m5b, c.x43) + c.a7f;
  const uint32_t n_g = lop3<0x6A>(sw, c.m1f, c.x07) + c.a3f;
  const uint32_t n_t = lop3<0x6A>(sw, c.m1f, c.x14) + c.a3f;
  nc = lop3<0xE4>(n_cg, n_g & n_t, c.m80);  // bit 7 from n_cg, the rest from n_g & n_t
  // a borrow / carry can only start at a byte that is itself out of range, and that byte is flagged
  const uint32_t qs = qw - c.qsub;
  bad = lop3<0xFE>(bad, qs, qs + c.x11);
  return lop3<0xF2>(qs << 2, nc >> 6, c.m03);  // (qs << 2) | (~(nc >> 6) & 0x03030303)
}
// The same with the 2-bit codes as a by-product (`code`: one clean code per byte, what the adapter scan gathers
// with one multiply) and the key bytes formed by a multiply-add on the other pipe: qs * 4 + code == (qs << 2) | code
// whenever no quality byte is out of the window (then the word is re-keyed anyway).
__device__ __forceinline__ uint32_t key_bytes_c(uint32_t sw, uint32_t qw, const KeyConsts &c, uint32_t &nc,
                                                uint32_t &bad, uint32_t &code) {
  const uint32_t n_cg = lop3<0x6A>(sw, c.m5b, c.x43) + c.a7f;
  const uint32_t n_g = lop3<0x6A>(sw, c.m1f, c.x07) + c.a3f;
  const uint32_t n_t = lop3<0x6A>(sw, c.m1f, c.x14) + c.a3f;
  nc = lop3<0xE4>(n_cg, n_g & n_t, c.m80);
  const uint32_t qs = qw - c.qsub;
  bad = lop3<0xFE>(bad, qs, qs + c.x11);
  code = lop3<0x0C>(nc >> 6, c.m03, c.m03);  // ~(nc >> 6) & 0x03030303
  return qs * 4u + code;
}
// key bytes of a word that has an out-of-window quality byte: dummy row, code kept
__device__ __forceinline__ uint32_t key_bytes_bad(uint32_t nc) {
  return (~(nc >> 6) & 0x03030303u) | ((kScoreBins << 2) * 0x01010101u);
}

__device__ __forceinline__ bool word_bad(uint32_t qw, uint32_t qsub) {
  const uint32_t qs = qw - qsub;
  return ((qs | (qs + (64u - kScoreBins) * 0x01010101u)) & 0xC0C0C0C0u) != 0u;
}

// 4 bases' 2-bit codes gathered from bits 7:6 of the 4 bytes of x into the TOP byte of the result,
// first base least significant (other result bytes are scratch)
__device__ __forceinline__ uint32_t gather_codes(uint32_t x) { return (x & 0xC0C0C0C0u) * 0x00041041u; }
// 16 bases (4 words of key bytes or of ~nc) -> 32 bits, first base least significant
__device__ __forceinline__ uint32_t pack16(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
  const uint32_t lo = __byte_perm(gather_codes(c0), gather_codes(c1), 0x0073);  // bytes: c0.top, c1.top
  const uint32_t hi = __byte_perm(gather_codes(c2), gather_codes(c3), 0x0073);
  return __byte_perm(lo, hi, 0x5410);
}

// the same from 4 words of key bytes, whose codes sit in bits 1:0 (rare path)
__device__ __forceinline__ uint32_t pack16_keys(uint4 k) {
  const uint32_t g0 = ((k.x & 0x03030303u) * 0x00041041u >> 18) & 0xFFu, g1 = ((k.y & 0x03030303u) * 0x00041041u >> 18) & 0xFFu;
  const uint32_t g2 = ((k.z & 0x03030303u) * 0x00041041u >> 18) & 0xFFu, g3 = ((k.w & 0x03030303u) * 0x00041041u >> 18) & 0xFFu;
  return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
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

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int t) {  // one PRMT / shift
  return t == 3 ? w >> 24 : __byte_perm(w, 0u, 0x4440u | (uint32_t)t);
}

// an opaque copy: keeps a launch constant in a register instead of re-deriving it from the constant bank
// and the kernel parameters inside the tile loop
__device__ __forceinline__ uint32_t pin(uint32_t x) {
  asm volatile("mov.u32 %0, %0;" : "+r"(x));
  return x;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts_u64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
template <typename T>
__device__ __forceinline__ T *shared_ptr(uint32_t addr) {
  return reinterpret_cast<T *>(__cvta_shared_to_generic((size_t)addr));
}

// rare path: the 4 bases of a word with an out-of-window quality byte, counted one by one
static __device__ __noinline__ uint32_t exact_word(uint32_t sw, uint32_t qw, uint32_t abs0, const uint32_t *soff,
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

}  // namespace qb
