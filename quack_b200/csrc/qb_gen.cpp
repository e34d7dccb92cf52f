// qb_gen.cpp -- the deterministic synthetic read generator: read i of a file depends only on
// (seed, mate, i), so shards, device batches and FASTQ files hold exactly the same reads.  No CUDA, no other part of
// the library: this file is also compiled into the standalone tool quack_b200/bin/qb_gen_fastq (tools/gen_fastq.cpp),
// which writes the benchmark inputs without loading libquack_b200.so.
#include <cmath>
#include <cstring>

#include "qb_host.h"

namespace qb {

// ------------------------------------------------------------------ synthetic reads

static inline uint64_t splitmix64(uint64_t &x) {
  uint64_t z = (x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct Xoshiro {
  uint64_t s[4];
  static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  inline uint64_t next() {  // xoshiro256**
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return r;
  }
};

static inline Xoshiro seed_read(uint64_t seed, uint64_t stream, uint64_t i) {
  uint64_t x = seed * 0xD1342543DE82EF95ull + stream * 0xA24BAED4963EE407ull + i * 0x9FB21C651E98DF25ull + 1;
  Xoshiro g;
  for (int k = 0; k < 4; k++) g.s[k] = splitmix64(x);
  return g;
}

static const char kAdapterR1[] = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT";  // TruSeq2_PE_f
static const char kAdapterR2[] = "AGATCGGAAGAGCGGTTCAGCAGGAATGCCGAG";  // TruSeq2_PE_r

uint32_t gen_length(uint64_t seed, uint64_t i, uint32_t len_min, uint32_t len_max) {
  if (len_max <= len_min) return len_min;
  Xoshiro g = seed_read(seed, 7, i);  // shared by both mates
  return len_min + (uint32_t)(g.next() % (uint64_t)(len_max - len_min + 1));
}

void gen_one(uint64_t seed, int mate, uint64_t i, uint32_t len, uint32_t len_max, double adapter_rate,
             uint8_t *seq, uint8_t *qual) {
  Xoshiro g = seed_read(seed, 100 + (uint64_t)mate, i);
  // bases: uniform ACGT, 1/1024 N
  for (uint32_t p = 0; p < len; p += 4) {
    uint64_t r = g.next();
    for (uint32_t k = 0; k < 4 && p + k < len; k++, r >>= 16)
      seq[p + k] = ((r >> 2) & 1023u) == 0 ? 'N' : "ACGT"[r & 3u];
  }
  // read-through: the pair decision and insert size come from the pair stream (same for both mates)
  if (adapter_rate > 0 && len > 21) {
    Xoshiro pg = seed_read(seed, 9, i);
    const double u = (double)(pg.next() >> 11) * (1.0 / 9007199254740992.0);
    if (u < adapter_rate) {
      const uint32_t s = 20 + (uint32_t)(pg.next() % (uint64_t)(len - 20));
      const char *ad = mate == 2 ? kAdapterR2 : kAdapterR1;
      const uint32_t al = (uint32_t)strlen(ad);
      for (uint32_t p = s; p < len; p++) seq[p] = (p - s < al) ? (uint8_t)ad[p - s] : 'A';
    }
  }
  // quality: mean 38 - 10 (p/L)^2 (2 lower for mate 2), sigma 4 (Irwin-Hall of 4 uniforms), Phred [2,41]
  const double L = (double)len_max, m0 = mate == 2 ? 36.0 : 38.0;
  for (uint32_t p = 0; p < len; p++) {
    const uint64_t r = g.next();
    const double z = ((double)(r & 0xFFFF) + (double)((r >> 16) & 0xFFFF) + (double)((r >> 32) & 0xFFFF) +
                      (double)(r >> 48) - 131070.0) * (1.0 / 37837.23);  // sd of the sum = 65536/sqrt(3)
    const double x = (double)p / L;
    double q = std::floor(m0 - 10.0 * x * x + 4.0 * z + 0.5);
    if (q < 2.0) q = 2.0;
    if (q > 41.0) q = 41.0;
    qual[p] = (uint8_t)(33 + (int)q);
  }
}

}  // namespace qb
