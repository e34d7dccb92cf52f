// qb_inflate.cu -- on-device DEFLATE for BGZF blocks.
//
// Reference: the inflate the reference runs inside gzread() (quack.c:160,187 -> klib/kseq.h:74,105); BGZF framing as
// in klib/bgzf.c:63-71 (18-byte header with the 'BC' block-size field), 261-266 (CRC32 | ISIZE trailer).  zlib itself is
// a dependency of the reference, not part of it: the decoder below restates RFC 1951 (stored, fixed and dynamic
// blocks, canonical Huffman codes) and RFC 1952's CRC-32; parity is pinned against zlib's output in the tests.
//
// One WARP per BGZF block (<= 64 KiB of text, independent of every other block: no window crosses a block
// boundary), thousands of blocks in flight.  All 32 lanes walk the bit stream with the same registers (no
// divergence; table look-ups broadcast), so a symbol costs one warp instruction stream; what the lanes share is the
// work around it: building the code tables of a dynamic block, copying a match (up to 258 bytes) 32 bytes per step,
// and the CRC-32 of the block's text (32 segments, combined with the x^n mod P operator).  Output positions are
// known before anything is decoded: ISIZE of every block is in its trailer, the host prefix-sums them.
//
// Shared memory per warp (4288 B): primary look-up tables (10 bits literal/length, 9 bits distance; u16 entries =
// symbol << 4 | code length), the symbols in canonical order and the number of codes per length (for the rare longer
// codes, decoded length by length), the code lengths read from the block header.
#include "qb_dev.cuh"

namespace qb {

constexpr int kLitBits = 10, kDistBits = 9;
constexpr int kInflateWarps = 4;  // warps (= BGZF blocks) per thread block

struct InflateTables {
  uint16_t lut_lit[1 << kLitBits];
  uint16_t lut_dist[1 << kDistBits];
  uint16_t sym_lit[288];
  uint16_t sym_dist[32];
  uint16_t cnt_lit[16], cnt_dist[16];
  uint16_t run[16], first_index[16], first_code[16];
  uint8_t lens[320];
};

// LSB-first bit reader over aligned 32-bit words; every lane holds the same state
struct BitReader {
  const uint32_t *wp;
  uint64_t buf;
  uint32_t cnt;
  __device__ __forceinline__ void start(const uint8_t *p) {
    const uintptr_t a = (uintptr_t)p;
    wp = (const uint32_t *)(a & ~(uintptr_t)3);
    const uint32_t skip = (uint32_t)(a & 3u) * 8u;
    buf = (uint64_t)(__ldg(wp++) >> skip);
    cnt = 32u - skip;
  }
  __device__ __forceinline__ void refill() {  // afterwards cnt >= 32
    if (cnt <= 32u) {
      buf |= (uint64_t)__ldg(wp++) << cnt;
      cnt += 32u;
    }
  }
  __device__ __forceinline__ uint32_t peek(uint32_t n) const { return (uint32_t)buf & ((1u << n) - 1u); }
  __device__ __forceinline__ void drop(uint32_t n) {
    buf >>= n;
    cnt -= n;
  }
  __device__ __forceinline__ uint32_t take(uint32_t n) {
    const uint32_t v = peek(n);
    drop(n);
    return v;
  }
  __device__ __forceinline__ const uint8_t *byte_pos() const { return (const uint8_t *)wp - (cnt >> 3); }
};

// Canonical Huffman tables from code lengths lens[0 .. n): cnt[len], the symbols ordered by (length, symbol), and the
// primary look-up table over the next `bits` bits of the stream (bit-reversed codes, replicated).  Warp-cooperative.
// Returns false for an over-subscribed set of lengths.
__device__ bool build_tables(const uint8_t *lens, uint32_t n, uint16_t *cnt, uint16_t *sym, uint16_t *lut, uint32_t bits,
                             InflateTables &t, uint32_t lane) {
  // codes per length: lane L (< 16) counts the symbols of length L
  uint32_t c = 0;
  if (lane < 16u)
    for (uint32_t s = 0; s < n; s++) c += lens[s] == lane;
  if (lane == 0) c = 0;  // length 0 = symbol not used
  if (lane < 16u) cnt[lane] = (uint16_t)c;
  // first index (symbols in canonical order) and first code of every length; over-subscription check
  uint32_t idx = 0, code = 0, my_idx = 0, my_code = 0;
  int left = 1;
  bool ok = true;
  for (uint32_t l = 1; l < 16u; l++) {
    const uint32_t cl = __shfl_sync(0xffffffffu, c, (int)l);
    code <<= 1;
    left = (left << 1) - (int)cl;
    if (left < 0) ok = false;
    if (lane == l) my_idx = idx, my_code = code;
    idx += cl;
    code += cl;
  }
  if (lane < 16u) t.run[lane] = t.first_index[lane] = (uint16_t)my_idx, t.first_code[lane] = (uint16_t)my_code;
  for (uint32_t i = lane; i < (1u << bits) / 2u; i += 32u) ((uint32_t *)lut)[i] = 0u;
  __syncwarp();
  if (!ok) return false;
  // symbols in rounds of 32: rank among the lanes with the same length keeps the symbol order
  for (uint32_t s0 = 0; s0 < n; s0 += 32u) {
    const uint32_t s = s0 + lane;
    const uint32_t l = s < n ? lens[s] : 0u;
    const uint32_t peers = __match_any_sync(0xffffffffu, l);
    if (l) {
      const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
      const uint32_t at = t.run[l] + rank;
      sym[at] = (uint16_t)s;
      if (l <= bits) {
        const uint32_t cd = t.first_code[l] + (at - t.first_index[l]);
        const uint32_t r = __brev(cd) >> (32u - l);
        for (uint32_t k = r; k < (1u << bits); k += 1u << l) lut[k] = (uint16_t)(s << 4 | l);
      }
    }
    __syncwarp();
    if (l && (peers & ((1u << lane) - 1u)) == 0u) t.run[l] += (uint16_t)__popc(peers);  // lowest lane of each group
    __syncwarp();
  }
  return true;
}

// one symbol: primary table, else length by length over the canonical order (codes longer than the table is wide)
__device__ __forceinline__ int decode_symbol(BitReader &br, const uint16_t *lut, uint32_t bits, const uint16_t *cnt,
                                             const uint16_t *sym) {
  const uint32_t e = lut[br.peek(bits)];
  if (e & 15u) {
    br.drop(e & 15u);
    return (int)(e >> 4);
  }
  uint32_t code = 0, first = 0, index = 0;
  uint32_t b = (uint32_t)br.buf;
  for (uint32_t l = 1; l < 16u; l++) {
    code |= b & 1u;
    b >>= 1;
    const uint32_t c = cnt[l];
    if (code < first + c) {
      br.drop(l);
      return (int)sym[index + (code - first)];
    }
    index += c;
    first = (first + c) << 1;
    code <<= 1;
  }
  return -1;
}

__constant__ uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// x^a * x^b mod P over GF(2), reflected representation (bit 31 = x^0), P = the CRC-32 polynomial
__device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
  uint32_t p = 0;
  for (uint32_t m = 0x80000000u; m; m >>= 1) {
    if (a & m) p ^= b;
    b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
  }
  return p;
}
// x^(8 n) mod P
__device__ uint32_t gf2_x8n(uint32_t n) {
  uint32_t p = 0x80000000u;   // x^0
  uint32_t sq = 0x00800000u;  // x^8
  while (n) {
    if (n & 1u) p = gf2_mulmod(sq, p);
    sq = gf2_mulmod(sq, sq);
    n >>= 1;
  }
  return p;
}

struct InflateArgs {
  const uint8_t *comp;       // compressed bytes of the chunk (readable 16 bytes past the end)
  const BgzfBlock *blocks;
  uint32_t n_blocks;
  uint8_t *text;             // inflated text of the chunk
  uint32_t *bad;             // set to 1 when a block does not inflate or its CRC-32 / ISIZE is wrong
  uint32_t *status;          // per block: 0 ok, else the failing check (diagnostics)
};

__global__ void __launch_bounds__(kInflateWarps * 32) inflate_bgzf_kernel(const InflateArgs a) {
  __shared__ InflateTables tables[kInflateWarps];
  __shared__ uint32_t crc_table[256];
  for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1u) ? 0xEDB88320u : 0u);
    crc_table[i] = c;
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t bi = blockIdx.x * kInflateWarps + warp;
  if (bi >= a.n_blocks) return;
  InflateTables &t = tables[warp];
  const BgzfBlock blk = a.blocks[bi];
  const uint8_t *in = a.comp + blk.in_off, *in_end = in + blk.in_len;
  uint8_t *out = a.text + blk.out_off;
  const uint32_t out_len = blk.out_len;
  const uint32_t *wlimit = (const uint32_t *)(((uintptr_t)in_end + 11u) & ~(uintptr_t)3);  // one refill past the end at most
  BitReader br;
  br.start(in);
  uint32_t opos = 0, err = 0;
  bool last = false;
  while (!last && !err) {
    br.refill();
    if (br.byte_pos() > in_end) { err = 1; break; }
    last = br.take(1) != 0u;
    const uint32_t type = br.take(2);
    if (type == 0u) {  // stored: LEN, ~LEN, bytes
      br.drop(br.cnt & 7u);
      br.refill();
      const uint32_t len = br.take(16), nlen = br.take(16);
      if ((len ^ nlen) != 0xFFFFu) { err = 2; break; }
      const uint8_t *src = br.byte_pos();
      if (src + len > in_end || opos + len > out_len) { err = 3; break; }
      for (uint32_t i = lane; i < len; i += 32u) out[opos + i] = __ldg(src + i);
      opos += len;
      br.start(src + len);
      continue;
    }
    if (type == 3u) { err = 4; break; }
    if (type == 1u) {  // fixed codes (RFC 1951 3.2.6)
      for (uint32_t s = lane; s < 288u; s += 32u) t.lens[s] = s < 144u ? 8 : s < 256u ? 9 : s < 280u ? 7 : 8;
      t.lens[288u + lane] = 5;
      __syncwarp();
      if (!build_tables(t.lens, 288, t.cnt_lit, t.sym_lit, t.lut_lit, kLitBits, t, lane) ||
          !build_tables(t.lens + 288, 30, t.cnt_dist, t.sym_dist, t.lut_dist, kDistBits, t, lane)) { err = 5; break; }
    } else {           // dynamic codes (3.2.7)
      br.refill();
      const uint32_t hlit = br.take(5) + 257u, hdist = br.take(5) + 1u, hclen = br.take(4) + 4u;
      if (hlit > 286u || hdist > 30u) { err = 6; break; }
      if (lane < 19u) t.lens[lane] = 0;
      __syncwarp();
      for (uint32_t i = 0; i < hclen; i++) {
        br.refill();
        const uint32_t l = br.take(3);
        if (lane == 0) t.lens[kClOrder[i]] = (uint8_t)l;
      }
      __syncwarp();
      // the code-length code borrows the distance tables (7-bit look-up)
      if (!build_tables(t.lens, 19, t.cnt_dist, t.sym_dist, t.lut_dist, 7, t, lane)) { err = 7; break; }
      __syncwarp();
      uint32_t i = 0, prev = 0;
      while (i < hlit + hdist) {
        br.refill();
        const int s = decode_symbol(br, t.lut_dist, 7, t.cnt_dist, t.sym_dist);
        if (s < 0) { err = 8; break; }
        uint32_t rep = 1, val = (uint32_t)s;
        if (s == 16) {
          if (i == 0) { err = 9; break; }
          val = prev, rep = 3u + br.take(2);
        } else if (s == 17) {
          val = 0, rep = 3u + br.take(3);
        } else if (s == 18) {
          val = 0, rep = 11u + br.take(7);
        }
        if (i + rep > hlit + hdist) { err = 10; break; }
        for (uint32_t k = lane; k < rep; k += 32u) t.lens[i + k] = (uint8_t)val;
        i += rep;
        prev = val;
      }
      if (err) break;
      __syncwarp();
      if (t.lens[256] == 0) { err = 11; break; }
      // distance lengths sit behind the literal/length lengths: build those first (build_tables only reads lens)
      if (!build_tables(t.lens + hlit, hdist, t.cnt_dist, t.sym_dist, t.lut_dist, kDistBits, t, lane) ||
          !build_tables(t.lens, hlit, t.cnt_lit, t.sym_lit, t.lut_lit, kLitBits, t, lane)) { err = 12; break; }
    }
    __syncwarp();
    // symbols of this block
    for (;;) {
      if (br.wp > wlimit) { err = 1; break; }  // (a damaged stream must not walk out of the buffer)
      br.refill();
      int s = decode_symbol(br, t.lut_lit, kLitBits, t.cnt_lit, t.sym_lit);
      if (s < 256) {
        if (s < 0) { err = 13; break; }
        if (opos >= out_len) { err = 14; break; }
        if (lane == 0) out[opos] = (uint8_t)s;
        opos++;
        continue;
      }
      if (s == 256) break;
      s -= 257;
      if (s >= 29) { err = 15; break; }
      uint32_t len;
      if (s < 8) len = 3u + (uint32_t)s;
      else if (s == 28) len = 258u;
      else {
        const uint32_t eb = ((uint32_t)s - 4u) >> 2;
        len = 3u + ((4u + ((uint32_t)s & 3u)) << eb) + br.take(eb);
      }
      br.refill();
      const int ds = decode_symbol(br, t.lut_dist, kDistBits, t.cnt_dist, t.sym_dist);
      if (ds < 0 || ds >= 30) { err = 16; break; }
      uint32_t dist;
      if (ds < 4) dist = 1u + (uint32_t)ds;
      else {
        const uint32_t eb = ((uint32_t)ds >> 1) - 1u;
        dist = 1u + ((2u + ((uint32_t)ds & 1u)) << eb) + br.take(eb);
      }
      if (dist > opos || opos + len > out_len) { err = 17; break; }
      __syncwarp();  // the bytes the match reads were written by other lanes
      const uint8_t *from = out + opos - dist;
      if (dist >= len) {
        for (uint32_t i = lane; i < len; i += 32u) out[opos + i] = from[i];
      } else {       // the match overlaps its own output: the source repeats with period dist
        for (uint32_t i = lane; i < len; i += 32u) out[opos + i] = from[i % dist];
      }
      opos += len;
    }
  }
  if (!err && opos != out_len) err = 18;
  if (!err && br.byte_pos() > in_end) err = 19;
  __syncwarp();
  if (!err && out_len) {  // CRC-32 of the text: 32 segments, folded with x^(8 * segment length)
    const uint32_t seg = (out_len + 31u) / 32u;
    const uint32_t lo = min(lane * seg, out_len), hi = min(lo + seg, out_len);
    uint32_t c = 0xFFFFFFFFu;
    for (uint32_t i = lo; i < hi; i++) c = crc_table[(c ^ out[i]) & 0xFFu] ^ (c >> 8);
    c = ~c;  // CRC-32 of the segment on its own (empty segment: 0)
    const uint32_t xs = gf2_x8n(seg);
    uint32_t crc = 0;
    for (uint32_t l = 0; l < 32u; l++) {
      const uint32_t cl = __shfl_sync(0xffffffffu, c, (int)l);
      const uint32_t n = min((l + 1u) * seg, out_len) - min(l * seg, out_len);
      if (n == 0u) continue;
      crc = gf2_mulmod(n == seg ? xs : gf2_x8n(n), crc) ^ cl;
    }
    if (crc != blk.crc) err = 20;
  } else if (!err && blk.crc != 0u) {
    err = 20;
  }
  if (lane == 0) {
    a.status[bi] = err;
    if (err) *a.bad = 1u;
  }
}

__global__ void inflate_merge_flag(const uint32_t *bad, TextState *state) {
  if (*bad) state->broken = 1u;
}

// the chunk's blocks are inflated on `stream`; *bad (device, zeroed here first) turns 1 if any block fails
cudaError_t launch_inflate_bgzf(const uint8_t *comp, const BgzfBlock *blocks, uint32_t n_blocks, uint8_t *text, uint32_t *bad,
                                uint32_t *status, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(bad, 0, 4, stream);
  if (e != cudaSuccess || n_blocks == 0) return e;
  const InflateArgs a{comp, blocks, n_blocks, text, bad, status};
  inflate_bgzf_kernel<<<(n_blocks + kInflateWarps - 1) / kInflateWarps, kInflateWarps * 32, 0, stream>>>(a);
  return cudaGetLastError();
}

// a failed inflate breaks the mate's text stream (the framing kernels then refuse the chunk and all later ones)
cudaError_t launch_inflate_merge(const uint32_t *bad, TextState *state, cudaStream_t stream) {
  inflate_merge_flag<<<1, 1, 0, stream>>>(bad, state);
  return cudaGetLastError();
}

}  // namespace qb
