// qb_transform.cu -- device-side transform() (reference: quack.c:230-293).
//
// What the reference does to the raw accumulator before drawing, restated for the device, quirks included:
//   * reads longer than 3000 bp are binned by 100 positions (quack.c:234-262): the loop reuses row `bin` in place and
//     clears only its first 96 columns when it enters the bin, so bin b >= 1 keeps the ORIGINAL kmer_count of position
//     b on top of its own sum; bin 0 is position 0 plus positions 1..99; the new max_length is the INDEX of the last
//     bin, i.e. the last bin is dropped;
//   * kmer_count becomes its running sum over the positions (quack.c:264-266);
//   * scores become integer percentages of the row's score total, which the reference adds up in an `int`
//     (quack.c:269-287); length_count and kmer_count become ceil(100 * (float) count / (float) nseq) in single
//     precision (quack.c:288-289).
// For million-base reads the result is 10 k rows instead of a million: only those cross the link.
#include "qb_dev.cuh"

namespace qb {

// without -a the reference counts kmer_count[10] once per read longer than 10 (quack.c:210-217): that sum lives in the
// length column (qb_finish applies it on the host copy; here it is applied on the fly)
__global__ void transform_noad_kmer10(const unsigned long long *rows, uint32_t ml, unsigned long long *out) {
  __shared__ unsigned long long part[256];
  unsigned long long s = 0;
  for (uint32_t i = 10u + threadIdx.x; i < ml; i += 256u) s += rows[(size_t)i * kRow + kColLength];
  part[threadIdx.x] = s;
  __syncthreads();
  for (uint32_t d = 128; d; d >>= 1) {
    if (threadIdx.x < d) part[threadIdx.x] += part[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = part[0];
}

// one block per output row: the row itself (ml <= 3000) or the bin's sum
__global__ void __launch_bounds__(128) transform_bins(const unsigned long long *rows, uint32_t ml, uint32_t binned,
                                                      const unsigned long long *kmer10, unsigned long long *out) {
  const uint32_t b = blockIdx.x, j = threadIdx.x;
  if (j >= (uint32_t)kRow) return;
  auto src = [&](uint32_t pos) -> unsigned long long {
    if (kmer10 && pos == 10u && j == (uint32_t)kColKmer) return *kmer10;
    return rows[(size_t)pos * kRow + j];
  };
  unsigned long long v;
  if (!binned) {
    v = src(b);
  } else {
    const uint32_t lo = b ? 100u * b : 1u, hi = min(100u * b + 100u, ml);
    v = (b == 0u || j == (uint32_t)kColKmer) ? src(b) : 0ull;  // (row `bin` is reused: see the header)
    for (uint32_t pos = lo; pos < hi; pos++) v += src(pos);
  }
  out[(size_t)b * kRow + j] = v;
}

// running sum of kmer_count, then the percentages; one block
__global__ void __launch_bounds__(1024) transform_finish(unsigned long long *out, uint32_t n_rows, unsigned long long n_reads) {
  if (threadIdx.x == 0)
    for (uint32_t pos = 1; pos < n_rows; pos++) out[(size_t)pos * kRow + kColKmer] += out[(size_t)(pos - 1u) * kRow + kColKmer];
  __syncthreads();
  for (uint32_t pos = threadIdx.x; pos < n_rows; pos += 1024u) {
    unsigned long long *row = out + (size_t)pos * kRow;
    unsigned long long s64 = 0;
    for (int j = 0; j < 91; j++) s64 += row[j];
    const int score_sum = (int)(unsigned int)(s64 & 0xFFFFFFFFull);  // the reference sums into an int
    if (score_sum != 0) {
      const unsigned long long div = (unsigned long long)(long long)score_sum;
      for (int j = 0; j < 91; j++) row[j] = 100ull * row[j] / div;
    }
    const float n = __ull2float_rn(n_reads);
    row[kColLength] = (unsigned long long)ceilf(__fdiv_rn(__fmul_rn(100.0f, __ull2float_rn(row[kColLength])), n));
    row[kColKmer] = (unsigned long long)ceilf(__fdiv_rn(__fmul_rn(100.0f, __ull2float_rn(row[kColKmer])), n));
  }
}

// rows: the raw [ml][97] accumulator (device); out: at least max(transformed_rows(ml), 1) rows (device).
// Returns the number of rows of the transformed result through *n_rows_out (host value, computed here).
uint32_t transformed_rows(uint32_t ml) { return ml > 3000u ? (ml - 1u) / 100u : ml; }

cudaError_t launch_transform(const unsigned long long *rows, uint32_t ml, unsigned long long n_reads, bool noad_quirk,
                             unsigned long long *out, unsigned long long *scratch1, cudaStream_t stream) {
  if (ml == 0) return cudaSuccess;
  const bool binned = ml > 3000u;
  const uint32_t n_rows = transformed_rows(ml);
  const bool quirk = noad_quirk && ml > 10u;
  if (quirk) transform_noad_kmer10<<<1, 256, 0, stream>>>(rows, ml, scratch1);
  // (the dropped last bin is not computed: n_rows blocks)
  if (n_rows) transform_bins<<<n_rows, 128, 0, stream>>>(rows, ml, binned ? 1u : 0u, quirk ? scratch1 : nullptr, out);
  if (n_rows) transform_finish<<<1, 1024, 0, stream>>>(out, n_rows, n_reads);
  return cudaGetLastError();
}

}  // namespace qb
