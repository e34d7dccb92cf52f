// qb_extras.cu -- opt-in side outputs that north_star names and the reference does NOT compute:
// per-position count of 'N' bases (the reference folds N into A, quack.c:150,200-202 -- the main kernels keep doing
// that) and the distribution of the per-read mean quality.  The per-position quality SUM is derived from the heatmap
// rows at finish (it equals sum_s s * scores[p][s]) and costs no kernel work.  None of this feeds the SVG.  There is no
// reference oracle for these arrays: parity is pinned only against the CPU restatement oracle/quack_oracle.c:qo_extras()
// ("parity unpinned" there).  A second pass over the batch, one warp per read; it runs only after qb_extras_enable(),
// so the statistics kernels pay nothing for it.
#include "qb_dev.cuh"

namespace qb {

constexpr uint32_t kMeanBins = 94;  // mean of (q - 33), q clamped to [33, 126]

struct XArgs {
  const uint8_t *seq, *qual;
  const uint32_t *offset, *length;  // unused when uniform_len != 0
  uint32_t n_reads, uniform_len, first_offset, len_cap;
  unsigned long long *n_count;      // [len_cap]
  unsigned long long *mean_hist;    // [kMeanBins]
};

__global__ void __launch_bounds__(256) extras_kernel(const XArgs a) {
  __shared__ uint32_t hist[kMeanBins];
  for (uint32_t i = threadIdx.x; i < kMeanBins; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = warp; r < a.n_reads; r += n_warps) {
    const uint32_t o = a.uniform_len ? a.first_offset + r * a.uniform_len : a.offset[r];
    const uint32_t l = a.uniform_len ? a.uniform_len : a.length[r];
    if (l == 0 || l > a.len_cap) continue;  // (the statistics kernels report such reads)
    // aligned 32-bit words that cover bytes [o, o + l); bytes outside are masked
    const uint32_t w0 = o & ~3u, w1 = (o + l + 3u) & ~3u;
    uint32_t sum = 0;
    for (uint32_t b = w0 + 4u * lane; b < w1; b += 128u) {
      uint32_t qw = __ldg(reinterpret_cast<const uint32_t *>(a.qual + b));
      const uint32_t sw = __ldg(reinterpret_cast<const uint32_t *>(a.seq + b));
      uint32_t mask = 0xFFFFFFFFu;  // bytes of the word that belong to the read
      if (b < o) mask &= 0xFFFFFFFFu << (8u * (o - b));
      if (b + 4u > o + l) mask &= 0xFFFFFFFFu >> (8u * (b + 4u - o - l));
      qw = __vminu4(__vmaxu4(qw, 0x21212121u), 0x7E7E7E7Eu) - 0x21212121u;  // q - 33 per byte, clamped to [0, 93]
      sum = __dp4a(qw & mask, 0x01010101u, sum);
      const uint32_t isn = __vcmpeq4(sw & 0xDFDFDFDFu, 0x4E4E4E4Eu) & mask;  // 'N' or 'n'
      if (isn) {
        for (uint32_t j = 0; j < 4u; j++)
          if (isn >> (8u * j) & 0xFFu) atomicAdd(&a.n_count[b + j - o], 1ull);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) atomicAdd(&hist[sum / l], 1u);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < kMeanBins; i += blockDim.x)
    if (hist[i]) atomicAdd(&a.mean_hist[i], (unsigned long long)hist[i]);
}

cudaError_t launch_extras(const BatchView &b, uint32_t len_cap, unsigned long long *n_count, unsigned long long *mean_hist,
                          int sm_count, cudaStream_t stream) {
  if (b.n_reads == 0) return cudaSuccess;
  const XArgs a{b.seq, b.qual, b.offset, b.length, b.n_reads, b.uniform_len, b.first_offset, len_cap, n_count, mean_hist};
  uint32_t grid = (b.n_reads + 7u) / 8u;
  const uint32_t cap = (uint32_t)sm_count * 8u;
  if (grid > cap) grid = cap;
  extras_kernel<<<grid, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace qb
