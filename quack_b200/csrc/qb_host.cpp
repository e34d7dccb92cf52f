// qb_host.cpp -- host-side helpers behind the C-ABI that need no CUDA: the base code table,
// adapter key extraction (read_adapters(), reference quack.c:154-178), the Bloom/bitmap images the
// kernels consume, and the deterministic synthetic read generator.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/quack_b200.h"
#include "qb_host.h"
#include "qb_kernels.cuh"

extern "C" int qb_base_code(int c) {
  const unsigned b = (unsigned)c & 0xFFu;
  // lookup[(c-65)&~32], quack.c:150: C->2, G->3, T->1, every other letter A..T (incl. N) -> 0;
  // the same bit tests define the bytes the reference leaves undefined.
  const int cg = (b & 0x5Bu) == 0x43u;
  const int lo = (b & 0x1Fu) == 0x07u || (b & 0x1Fu) == 0x14u;
  return 2 * cg + lo;
}

extern "C" int qb_adapter_record_keys(const char *seq, size_t l, uint32_t *keys, size_t cap) {
  // quack.c:165-172: pack bases 0..9 (not inserted), then roll in base i = 10..l-1 and insert
  if (l <= QB_KMER_SIZE) return 0;
  if (l - QB_KMER_SIZE > cap) return QB_ERR_CAPACITY;
  uint32_t index = 0;
  size_t i, n = 0;
  for (i = 0; i < QB_KMER_SIZE; i++)
    index = ((index << 2) + (uint32_t)qb_base_code((unsigned char)seq[i])) & (QB_KEY_SPACE - 1);
  for (; i < l; i++) {
    index = ((index << 2) + (uint32_t)qb_base_code((unsigned char)seq[i])) & (QB_KEY_SPACE - 1);
    keys[n++] = index;
  }
  return (int)n;
}

namespace qb {

uint32_t key_to_internal(uint32_t ref_key) {
  // reference order: first base in bits 19:18; kernel order: first base in bits 1:0
  uint32_t k = 0;
  for (int i = 0; i < QB_KMER_SIZE; i++) k |= ((ref_key >> (2 * (QB_KMER_SIZE - 1 - i))) & 3u) << (2 * i);
  return k;
}

void build_adapter_images(const uint32_t *ref_keys, uint32_t n, std::vector<uint32_t> &bitmap,
                          std::vector<uint32_t> &anchor, std::vector<uint32_t> &exact, uint32_t &n_anchors,
                          double &anchor_density) {
  bitmap.assign(QB_KEY_SPACE / 32, 0);
  std::vector<uint32_t> keys;
  keys.reserve(n);
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t k = key_to_internal(ref_keys[i] & (QB_KEY_SPACE - 1));
    if (!((bitmap[k >> 5] >> (k & 31)) & 1u)) {
      bitmap[k >> 5] |= 1u << (k & 31);
      keys.push_back(k);
    }
  }
  // Anchor filter: every 10-mer window [s, s+9] contains exactly one 7-mer that starts at a position
  // congruent to 3 mod 4 (start s, s+1, s+2 or s+3).  A window can only be in the set if that 7-mer is
  // one of the 7-mers found at offsets 0..3 of a set member, so the kernel probes one 7-mer per 4 bases
  // against this exact 2^14-bit map and confirms the 4 windows around a hit against the exact key set.
  anchor.assign(kAnchorWords, 0);
  n_anchors = 0;
  for (uint32_t k : keys)
    for (uint32_t o = 0; o < 4; o++) {
      const uint32_t a = (k >> (2 * o)) & (kAnchorSpace - 1);
      if (!((anchor[a >> 5] >> (a & 31)) & 1u)) {
        anchor[a >> 5] |= 1u << (a & 31);
        n_anchors++;
      }
    }
  anchor_density = (double)n_anchors / (double)kAnchorSpace;
  exact.clear();
  if (keys.size() * 3 <= kExactSlots) {  // cuckoo insertion; at this load it needs a handful of moves at most
    exact.assign(kExactSlots, kExactEmpty);
    bool ok = true;
    for (uint32_t k : keys) {
      uint32_t cur = k;
      bool second = false, placed = false;
      for (int kick = 0; kick < 1000 && !placed; kick++) {
        const uint32_t slot = (second ? exact_off2(cur) : exact_off1(cur)) / 4u;
        if (exact[slot] == kExactEmpty) {
          exact[slot] = cur;
          placed = true;
        } else {
          std::swap(cur, exact[slot]);
          second = (slot < kExactHalf);  // the evicted key lived in this half: try its other slot
        }
      }
      if (!placed) {
        ok = false;
        break;
      }
    }
    if (!ok) exact.clear();  // kernels then confirm against the 2^20-bit map in L2
  }
}

}  // namespace qb

extern "C" int qb_gen_reads(uint64_t seed, int mate, uint64_t first_read, uint32_t n_reads, uint32_t len_min,
                            uint32_t len_max, double adapter_rate, uint8_t *seq, uint8_t *qual, uint32_t *offset,
                            uint32_t *length, uint64_t *n_bytes) {
  if (!seq || !qual || !offset || !length || len_min == 0 || len_max < len_min) return QB_ERR_ARG;
  uint64_t o = 0;
  for (uint32_t r = 0; r < n_reads; r++) {
    const uint32_t l = qb::gen_length(seed, first_read + r, len_min, len_max);
    if (o + l > 0xFFFFFF00ull) return QB_ERR_CAPACITY;
    offset[r] = (uint32_t)o;
    length[r] = l;
    o += l;
  }
  if (n_bytes) *n_bytes = o;
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 4;
  if (nt > 64) nt = 64;
  if (n_reads < 4096) nt = 1;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++) {
    const uint32_t a = (uint32_t)((uint64_t)n_reads * t / nt), b = (uint32_t)((uint64_t)n_reads * (t + 1) / nt);
    th.emplace_back([=]() {
      for (uint32_t r = a; r < b; r++)
        qb::gen_one(seed, mate, first_read + r, length[r], len_max, adapter_rate, seq + offset[r], qual + offset[r]);
    });
  }
  for (auto &t : th) t.join();
  return QB_OK;
}
