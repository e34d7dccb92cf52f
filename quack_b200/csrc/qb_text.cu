// qb_text.cu -- on-device FASTQ record framing (reference: kseq_read(), klib/kseq.h:177-218).
//
// The host ships decompressed FASTQ TEXT in chunks cut anywhere (one host-to-device stream instead of four packed
// arrays); the device finds the line ends, takes every four lines as a record, checks that the record has the
// canonical shape -- '@' header line, ONE non-empty sequence line, '+' line, ONE quality line of the same length, no
// '\r' -- and packs bases and quality bytes into the batch layout the statistics kernels take (concatenated, unpadded,
// u32 offsets and lengths).  For such records the result is exactly what kseq_read() delivers.  Anything else
// (multi-line records, blank lines, FASTA records, garbage in front of the first header, '\r\n') clears the `valid`
// flag of the chunk: nothing of it is counted and the caller falls back to the host reader, which implements the
// full kseq semantics (host/fq_reader.c).
//
// A chunk may end inside a record: the bytes behind the last complete record (the TAIL) stay on the device in a
// small per-mate carry buffer and are put in front of the next chunk's text; the chunks of one mate are framed in
// submission order (the API layer chains them with an event).
//
//   frame_count     newlines per 4 KiB tile of [carry | chunk]
//   frame_scan      exclusive prefix sum of a u32 array (one block; used for tile counts and for read lengths)
//   frame_index     position of every newline, in order
//   frame_records   one thread per record: line ends -> sequence / quality ranges, shape check, length
//   frame_sums      per-tile sums of the read lengths (packed offsets = their prefix sums)
//   frame_pack      one warp per record: copies bases and quality bytes to their packed places, writes offset/length
//   frame_tail      moves the tail into the carry buffer, writes the chunk summary
#include "qb_dev.cuh"

namespace qb {

constexpr uint32_t kTTile = 4096;  // bytes per tile of the newline count
constexpr uint32_t kTBlock = 256;

// text byte i of the framing view: carry bytes first, then the chunk
struct TextView {
  const uint8_t *carry, *chunk;
  uint32_t carry_len, total;  // total = carry_len + chunk bytes
  __device__ __forceinline__ uint8_t at(uint32_t i) const { return i < carry_len ? carry[i] : chunk[i - carry_len]; }
};

__global__ void __launch_bounds__(kTBlock) frame_count(const uint8_t *carry, const uint8_t *chunk, uint32_t n_chunk,
                                                        const TextState *st, uint32_t *tile_count, uint32_t n_tiles_cap) {
  if (st->broken) return;
  const TextView v{carry, chunk, st->tail_len, st->tail_len + n_chunk};
  const uint32_t tile = blockIdx.x;
  const uint32_t t0 = tile * kTTile;
  if (t0 >= v.total) {
    if (threadIdx.x == 0 && tile < n_tiles_cap) tile_count[tile] = 0;
    return;
  }
  uint32_t c = 0;
  for (uint32_t i = t0 + threadIdx.x; i < min(t0 + kTTile, v.total); i += kTBlock) c += v.at(i) == '\n';
  c = (uint32_t)warp_sum(c);
  __shared__ uint32_t ws[kTBlock / 32];
  if ((threadIdx.x & 31u) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
    for (uint32_t w = 0; w < kTBlock / 32; w++) s += ws[w];
    tile_count[tile] = s;
  }
}

// in-place exclusive prefix sum of a[0 .. n), total to *total_out; one block of 1024 threads
// (n_lines_p != nullptr: only the first *n_lines_p / 4 + 1 entries are in use -- the records of this chunk)
__global__ void __launch_bounds__(1024) frame_scan(uint32_t *a, uint32_t n, uint32_t *total_out, const TextState *st,
                                                    const uint32_t *n_lines_p) {
  if (st->broken) return;
  if (n_lines_p) n = min(n, *n_lines_p / 4u + 1u);
  __shared__ uint32_t part[1024];
  const uint32_t t = threadIdx.x;
  const uint32_t per = (n + 1023u) / 1024u;
  const uint32_t lo = min(t * per, n), hi = min(lo + per, n);
  uint32_t s = 0;
  for (uint32_t i = lo; i < hi; i++) s += a[i];
  part[t] = s;
  __syncthreads();
  for (uint32_t d = 1; d < 1024u; d <<= 1) {  // Hillis-Steele over the 1024 partial sums
    const uint32_t v = t >= d ? part[t - d] : 0u;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = t ? part[t - 1] : 0u;
  for (uint32_t i = lo; i < hi; i++) {
    const uint32_t x = a[i];
    a[i] = run;
    run += x;
  }
  if (t == 1023u) *total_out = part[1023];
}

__global__ void __launch_bounds__(kTBlock) frame_index(const uint8_t *carry, const uint8_t *chunk, uint32_t n_chunk,
                                                        const TextState *st, const uint32_t *tile_off, uint32_t *nl,
                                                        uint32_t nl_cap) {
  if (st->broken) return;
  const TextView v{carry, chunk, st->tail_len, st->tail_len + n_chunk};
  const uint32_t t0 = blockIdx.x * kTTile;
  if (t0 >= v.total) return;
  __shared__ uint32_t wbase[kTBlock / 32 + 1];
  uint32_t base = tile_off[blockIdx.x];
  // the tile in rounds of kTBlock bytes: ranks from ballots, so the positions come out in order
  for (uint32_t r0 = t0; r0 < min(t0 + kTTile, v.total); r0 += kTBlock) {
    const uint32_t i = r0 + threadIdx.x;
    const bool is_nl = i < v.total && v.at(i) == '\n';
    const uint32_t m = __ballot_sync(0xffffffffu, is_nl);
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    if (lane == 0) wbase[w + 1] = (uint32_t)__popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
      wbase[0] = 0;
      for (uint32_t k = 1; k <= kTBlock / 32; k++) wbase[k] += wbase[k - 1];
    }
    __syncthreads();
    if (is_nl) {
      const uint32_t idx = base + wbase[w] + (uint32_t)__popc(m & ((1u << lane) - 1u));
      if (idx < nl_cap) nl[idx] = i;
    }
    base += wbase[kTBlock / 32];
    __syncthreads();
  }
}

// one thread per record: the four line ends -> ranges; shape check; length
__global__ void __launch_bounds__(kTBlock) frame_records(const uint8_t *carry, const uint8_t *chunk, uint32_t n_chunk,
                                                          const TextState *st, const uint32_t *nl, const uint32_t *n_lines_p,
                                                          uint32_t nl_cap, uint32_t *rec_seq, uint32_t *rec_qual, uint32_t *rec_len,
                                                          uint32_t rec_cap, TextSummary *sum) {
  if (st->broken) return;
  const TextView v{carry, chunk, st->tail_len, st->tail_len + n_chunk};
  const uint32_t n_lines = *n_lines_p;
  const uint32_t n_rec = n_lines / 4u;
  const uint32_t r = blockIdx.x * kTBlock + threadIdx.x;
  if (n_lines > nl_cap || n_rec > rec_cap) {  // more lines than the index holds: not FASTQ of any sane shape
    if (r == 0) sum->valid = 0;
    return;
  }
  if (r >= n_rec) {
    if (r < rec_cap) rec_len[r] = 0;  // (the prefix sum below runs over the whole array)
    return;
  }
  const uint32_t e0 = nl[4u * r], e1 = nl[4u * r + 1u], e2 = nl[4u * r + 2u], e3 = nl[4u * r + 3u];
  const uint32_t s0 = r ? nl[4u * r - 1u] + 1u : 0u;          // header line [s0, e0)
  const uint32_t s1 = e0 + 1u, s2 = e1 + 1u, s3 = e2 + 1u;    // sequence [s1, e1), '+' line [s2, e2), quality [s3, e3)
  const uint32_t l = e1 - s1;
  bool ok = v.at(s0) == '@' && l >= 1u && e2 > s2 && v.at(s2) == '+' && e3 - s3 == l;
  if (ok) {
    const uint8_t f = v.at(s1);
    ok = f != '+' && f != '@' && f != '>';                    // (kseq would take such a line for a record boundary)
    // a '\r' in front of a line end is stripped by kseq (kseq.h:141): leave those files to the host reader
    ok = ok && v.at(e0 - (e0 > s0 ? 1u : 0u)) != '\r' && v.at(e1 - 1u) != '\r' && v.at(e3 - 1u) != '\r';
  }
  rec_seq[r] = s1;
  rec_qual[r] = s3;
  rec_len[r] = ok ? l : 0u;
  if (!ok) sum->valid = 0;
  if (ok) {
    atomicMin(&sum->min_len, l);
    atomicMax(&sum->max_len, l);
  }
}

// rec_len -> packed offsets happen with frame_scan on a copy; this kernel packs: one warp per record
__global__ void __launch_bounds__(kTBlock) frame_pack(const uint8_t *carry, const uint8_t *chunk, uint32_t n_chunk,
                                                       const TextState *st, const uint32_t *n_lines_p, const uint32_t *rec_seq,
                                                       const uint32_t *rec_qual, const uint32_t *rec_len, const uint32_t *rec_off,
                                                       const TextSummary *sum, uint8_t *seq, uint8_t *qual, uint32_t *offset,
                                                       uint32_t *length, uint32_t out_cap) {
  if (st->broken || !sum->valid) return;
  const TextView v{carry, chunk, st->tail_len, st->tail_len + n_chunk};
  const uint32_t n_rec = *n_lines_p / 4u;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t r = (blockIdx.x * kTBlock + threadIdx.x) >> 5; r < n_rec; r += (gridDim.x * kTBlock) >> 5) {
    const uint32_t l = rec_len[r], o = rec_off[r], s = rec_seq[r], q = rec_qual[r];
    if (o + l > out_cap) continue;  // (cannot happen: the packed bytes are fewer than half of the text)
    for (uint32_t i = lane; i < l; i += 32u) {
      seq[o + i] = v.at(s + i);
      qual[o + i] = v.at(q + i);
    }
    if (lane == 0) {
      offset[r] = o;
      length[r] = l;
    }
  }
}

// the bytes behind the last complete record become the carry of the next chunk; chunk summary
__global__ void __launch_bounds__(kTBlock) frame_tail(const uint8_t *carry, const uint8_t *chunk, uint32_t n_chunk, TextState *st,
                                                       const uint32_t *nl, const uint32_t *n_lines_p, const uint32_t *n_bytes_p,
                                                       uint8_t *carry_out, uint32_t carry_cap, uint32_t out_cap, TextSummary *sum) {
  __shared__ uint32_t tail0_s, tail_len_s;
  if (st->broken) {
    if (threadIdx.x == 0) sum->valid = 0;
    return;
  }
  const TextView v{carry, chunk, st->tail_len, st->tail_len + n_chunk};
  if (threadIdx.x == 0) {
    const uint32_t n_lines = *n_lines_p, n_rec = n_lines / 4u;
    const uint32_t t0 = (sum->valid && n_rec) ? nl[4u * n_rec - 1u] + 1u : 0u;
    tail0_s = t0;
    tail_len_s = v.total - t0;
    sum->n_lines = n_lines;
    sum->n_reads = sum->valid ? n_rec : 0u;
    sum->n_bytes = sum->valid ? *n_bytes_p : 0u;
    sum->tail_len = v.total - t0;
    // the carry must hold the tail.  (At the end of the stream a tail that is left over is a truncated record --
    // kseq_read() returns -2 there and the records in front of it stay counted: the host reads tail_len.)
    if (v.total - t0 > carry_cap) sum->valid = 0;
    if (sum->valid && *n_bytes_p > out_cap) sum->valid = 0, sum->n_reads = sum->n_bytes = 0;  // (the API sizes the chunks so that it fits)
  }
  __syncthreads();
  const uint32_t t0 = tail0_s, tl = tail_len_s;
  if (tl <= carry_cap) {
    // carry_out is a second buffer (the view still reads the old carry): copy, then publish the length
    for (uint32_t i = threadIdx.x; i < tl; i += kTBlock) carry_out[i] = v.at(t0 + i);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (sum->valid == 0u) st->broken = 1u;
    st->tail_len = tl <= carry_cap ? tl : 0u;
  }
}

__global__ void frame_init(TextSummary *sum) {
  sum->n_reads = sum->n_bytes = 0;
  sum->min_len = 0xFFFFFFFFu;
  sum->max_len = 0;
  sum->valid = 1;
  sum->tail_len = sum->n_lines = sum->pad = 0;
}

// Queues the framing of one chunk on `stream`.  Buffers (device): carry_in / carry_out = the mate's two carry buffers
// (they swap roles chunk by chunk), scratch = tile counts [n_tiles + 1] | nl [nl_cap] | rec_seq | rec_qual | rec_len |
// rec_off [rec_cap each] | two u32 totals.  The summary lands in sum_dev; the caller copies it to the host.
cudaError_t launch_text_frame(const uint8_t *chunk, uint32_t n_chunk, const uint8_t *carry_in, uint8_t *carry_out,
                              uint32_t carry_cap, TextState *state, uint32_t *scratch, uint32_t nl_cap, uint32_t rec_cap,
                              uint8_t *seq, uint8_t *qual, uint32_t *offset, uint32_t *length, uint32_t out_cap,
                              TextSummary *sum_dev, cudaStream_t stream) {
  const uint32_t max_total = n_chunk + carry_cap;
  const uint32_t n_tiles = (max_total + kTTile - 1u) / kTTile;
  uint32_t *tile_cnt = scratch;
  uint32_t *nl = tile_cnt + n_tiles + 1u;
  uint32_t *rec_seq = nl + nl_cap, *rec_qual = rec_seq + rec_cap, *rec_len = rec_qual + rec_cap, *rec_off = rec_len + rec_cap;
  uint32_t *totals = rec_off + rec_cap;  // [0] lines, [1] packed bytes
  frame_init<<<1, 1, 0, stream>>>(sum_dev);
  frame_count<<<n_tiles, kTBlock, 0, stream>>>(carry_in, chunk, n_chunk, state, tile_cnt, n_tiles);
  frame_scan<<<1, 1024, 0, stream>>>(tile_cnt, n_tiles, totals, state, nullptr);
  frame_index<<<n_tiles, kTBlock, 0, stream>>>(carry_in, chunk, n_chunk, state, tile_cnt, nl, nl_cap);
  const uint32_t rec_blocks = (rec_cap + kTBlock - 1u) / kTBlock;
  frame_records<<<rec_blocks, kTBlock, 0, stream>>>(carry_in, chunk, n_chunk, state, nl, totals, nl_cap, rec_seq, rec_qual, rec_len,
                                                   rec_cap, sum_dev);
  cudaError_t e = cudaMemcpyAsync(rec_off, rec_len, (size_t)rec_cap * 4u, cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return e;
  frame_scan<<<1, 1024, 0, stream>>>(rec_off, rec_cap, totals + 1, state, totals);
  frame_pack<<<592, kTBlock, 0, stream>>>(carry_in, chunk, n_chunk, state, totals, rec_seq, rec_qual, rec_len, rec_off, sum_dev, seq,
                                         qual, offset, length, out_cap);
  frame_tail<<<1, kTBlock, 0, stream>>>(carry_in, chunk, n_chunk, state, nl, totals, totals + 1, carry_out, carry_cap, out_cap, sum_dev);
  return cudaGetLastError();
}

size_t text_scratch_words(uint32_t chunk_cap, uint32_t carry_cap, uint32_t nl_cap, uint32_t rec_cap) {
  return (size_t)((chunk_cap + carry_cap + kTTile - 1u) / kTTile) + 1u + nl_cap + 4u * (size_t)rec_cap + 4u;
}

}  // namespace qb
