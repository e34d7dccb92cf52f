// gen_fastq.cpp -- standalone writer of the synthetic benchmark inputs: 4-line FASTQ records
// '@r<9-digit index>/<mate>', bare '+' line, reads from the deterministic generator of quack_b200/csrc/qb_gen.cpp
// (the same reads qb_gen_reads() puts into device batches).  Output: plain text, gzip (concatenated members of
// 16 MiB of text each -- SURVEY 8d asks for members of at most 64 MiB --, compressed in parallel: what the
// reference's gzread handles too), or BGZF (<= 65280-byte
// blocks with the 'BC' size field, reference klib/bgzf.c:63-71).  Links nothing of the library: bench.py's reference
// arm and the file -> SVG comparison get their inputs from this tool.
//
//   qb_gen_fastq OUT SEED MATE FIRST N_READS LEN_MIN LEN_MAX ADAPTER_RATE plain|gz|bgzf [LEVEL] [THREADS]
#include <zlib.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../quack_b200/csrc/qb_host.h"

namespace {

void chunk_text(uint64_t seed, int mate, uint64_t first, uint32_t n, uint32_t lmin, uint32_t lmax, double rate,
                std::string &out) {
  std::vector<uint8_t> seq(lmax), qual(lmax);
  out.clear();
  out.reserve((size_t)n * (2 * (size_t)lmax + 20));
  char name[32];
  for (uint32_t r = 0; r < n; r++) {
    const uint64_t i = first + r;
    const uint32_t l = qb::gen_length(seed, i, lmin, lmax);
    qb::gen_one(seed, mate, i, l, lmax, rate, seq.data(), qual.data());
    const int nl = snprintf(name, sizeof name, "@r%09llu/%d\n", (unsigned long long)i, mate);
    out.append(name, (size_t)nl);
    out.append((const char *)seq.data(), l);
    out.append("\n+\n", 3);
    out.append((const char *)qual.data(), l);
    out.push_back('\n');
  }
}

bool gzip_member(const std::string &text, int level, std::string &out) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
  out.resize(deflateBound(&zs, text.size()) + 64);
  zs.next_in = (Bytef *)text.data();
  zs.avail_in = (uInt)text.size();
  zs.next_out = (Bytef *)&out[0];
  zs.avail_out = (uInt)out.size();
  const int rc = deflate(&zs, Z_FINISH);
  out.resize(out.size() - zs.avail_out);
  deflateEnd(&zs);
  return rc == Z_STREAM_END;
}

bool bgzf_blocks(const std::string &text, int level, std::string &out) {
  out.clear();
  out.reserve(text.size() / 2 + 1024);
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
  const size_t block = 65280;
  std::vector<uint8_t> buf(65536 + 1024);
  for (size_t o = 0; o < text.size(); o += block) {
    const size_t n = text.size() - o < block ? text.size() - o : block;
    deflateReset(&zs);
    zs.next_in = (Bytef *)text.data() + o;
    zs.avail_in = (uInt)n;
    zs.next_out = buf.data();
    zs.avail_out = (uInt)buf.size();
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) return false;
    const size_t clen = buf.size() - zs.avail_out;
    if (clen + 26 > 65536) return false;  // generated text always compresses
    const uint32_t bsize = (uint32_t)clen + 25, crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)text.data() + o, (uInt)n);
    const uint8_t hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (uint8_t)bsize, (uint8_t)(bsize >> 8)};
    out.append((const char *)hdr, 18);
    out.append((const char *)buf.data(), clen);
    const uint32_t isize = (uint32_t)n;
    const uint8_t tr[8] = {(uint8_t)crc, (uint8_t)(crc >> 8), (uint8_t)(crc >> 16), (uint8_t)(crc >> 24),
                           (uint8_t)isize, (uint8_t)(isize >> 8), (uint8_t)(isize >> 16), (uint8_t)(isize >> 24)};
    out.append((const char *)tr, 8);
  }
  deflateEnd(&zs);
  return true;
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 10) {
    fprintf(stderr, "usage: %s OUT SEED MATE FIRST N_READS LEN_MIN LEN_MAX ADAPTER_RATE plain|gz|bgzf [LEVEL] [THREADS]\n", argv[0]);
    return 2;
  }
  const char *path = argv[1];
  const uint64_t seed = strtoull(argv[2], nullptr, 10);
  const int mate = atoi(argv[3]);
  const uint64_t first = strtoull(argv[4], nullptr, 10), n_reads = strtoull(argv[5], nullptr, 10);
  const uint32_t lmin = (uint32_t)atoi(argv[6]), lmax = (uint32_t)atoi(argv[7]);
  const double rate = atof(argv[8]);
  const std::string mode = argv[9];
  const int level = argc > 10 ? atoi(argv[10]) : 1;
  unsigned nt = argc > 11 ? (unsigned)atoi(argv[11]) : std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (lmin == 0 || lmax < lmin || (mode != "plain" && mode != "gz" && mode != "bgzf")) return 2;
  FILE *f = fopen(path, "wb");
  if (!f) {
    perror(path);
    return 1;
  }
  // chunks of <= 16 MiB of text (one gzip member each), produced by a pool of threads, written in order
  const uint64_t per_chunk = (16ull << 20) / (2ull * lmax + 20);
  const uint64_t n_chunks = (n_reads + per_chunk - 1) / per_chunk;
  std::vector<std::string> done(n_chunks);
  std::vector<std::atomic<int>> ready(n_chunks);
  for (auto &r : ready) r = 0;
  std::atomic<uint64_t> next{0}, written{0};
  std::atomic<bool> failed{false};
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nt; t++)
    pool.emplace_back([&]() {
      std::string text, packed;
      for (;;) {
        const uint64_t c = next++;
        if (c >= n_chunks || failed) return;
        while (c > written + 2ull * nt) std::this_thread::yield();  // bounded memory: stay close to the writer
        const uint64_t r0 = c * per_chunk, n = n_reads - r0 < per_chunk ? n_reads - r0 : per_chunk;
        chunk_text(seed, mate, first + r0, (uint32_t)n, lmin, lmax, rate, text);
        if (mode == "plain")
          done[c].swap(text);
        else if (!(mode == "gz" ? gzip_member(text, level, packed) : bgzf_blocks(text, level, packed)))
          failed = true;
        else
          done[c].swap(packed);
        ready[c] = 1;
      }
    });
  uint64_t text_bytes = 0;
  for (uint64_t c = 0; c < n_chunks && !failed; c++) {
    while (!ready[c] && !failed) std::this_thread::yield();
    if (failed) break;
    if (fwrite(done[c].data(), 1, done[c].size(), f) != done[c].size()) failed = true;
    text_bytes += done[c].size();
    std::string().swap(done[c]);
    written = c + 1;
  }
  for (auto &t : pool) t.join();
  if (mode == "bgzf" && !failed) {
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, sizeof eof, f);
  }
  if (fclose(f) != 0 || failed) {
    fprintf(stderr, "%s: write failed\n", path);
    return 1;
  }
  return 0;
}
