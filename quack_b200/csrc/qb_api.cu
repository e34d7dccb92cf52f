// qb_api.cu -- implementation of the C-ABI in include/quack_b200.h: devices, the pinned
// double/triple-buffered batch ring (cudaMemcpyAsync overlapped with compute, one stream per
// slot), per-mate u64 accumulators, the NCCL reduce of the accumulators, device-resident batches
// and their CUDA-event timing.  Replaces the loop of read_fastq() (reference quack.c:180-228).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <deque>
#include <vector>

#include "../../include/quack_b200.h"
#include "qb_host.h"
#include "qb_kernels.cuh"

namespace {

thread_local std::string g_create_error;

// NCCL is bound at first use with dlopen instead of at link time: a host process may already carry
// its own libnccl.so.2 (PyTorch bundles a newer one than the system's) and must keep exactly one.
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      api.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
#define QB_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym))
    QB_SYM(GetUniqueId, "ncclGetUniqueId");
    QB_SYM(CommInitRank, "ncclCommInitRank");
    QB_SYM(CommInitAll, "ncclCommInitAll");
    QB_SYM(CommDestroy, "ncclCommDestroy");
    QB_SYM(Reduce, "ncclReduce");
    QB_SYM(AllReduce, "ncclAllReduce");
    QB_SYM(GroupStart, "ncclGroupStart");
    QB_SYM(GroupEnd, "ncclGroupEnd");
    QB_SYM(GetErrorString, "ncclGetErrorString");
#undef QB_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.Reduce && api.AllReduce &&
             api.GroupStart && api.GroupEnd && api.GetErrorString;
  });
  return api.ok ? &api : nullptr;
}

struct Slot {
  uint8_t *h_seq = nullptr, *h_qual = nullptr;
  uint32_t *h_off = nullptr, *h_len = nullptr;
  uint8_t *d_seq = nullptr, *d_qual = nullptr;
  uint32_t *d_off = nullptr, *d_len = nullptr;
  uint2 *d_tiles = nullptr;  // tile descriptors of the warp-tile kernel (one per read at most)
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  // Ownership (guarded by qb_ctx::mu).  FREE: nobody uses the buffers.  HELD: a host thread got the slot from
  // qb_acquire()/qb_submit_from() and may write its pinned buffers; no other thread is ever handed it.  PENDING:
  // submitted; copies and the kernel are queued behind `done`, the next owner waits for that event first.
  // DEFERRED (text path): the device is framing the slot's text; the statistics kernel is launched by the host
  // once the chunk summary (number of reads, read lengths) has arrived -- flush_deferred() -- and only then the slot
  // turns PENDING.
  enum State { FREE = 0, HELD = 1, PENDING = 2, DEFERRED = 3, FLUSHING = 4 };  // FLUSHING: one thread is inside flush_deferred()
  State state = FREE;
  uint64_t seq = 0;  // submit order, so that the oldest pending slot is recycled first
  // text path (qb_text_acquire / qb_text_submit), allocated at its first use
  uint8_t *h_text = nullptr, *d_text = nullptr;
  uint32_t *d_scratch = nullptr;
  qb::TextSummary *h_sum = nullptr, *d_sum = nullptr;
  cudaEvent_t framed = nullptr;
  // compressed input of the text path (qb_bgzf_submit), allocated at its first use
  uint8_t *d_comp = nullptr;
  qb::BgzfBlock *h_blk = nullptr, *d_blk = nullptr;
  uint32_t *d_blk_status = nullptr, *d_bad = nullptr;
  int text_mate = -1;
  int text_last = 0;
};

struct Device {
  int id = 0;
  int sm_count = 0;
  int smem_optin = 0;
  int smem_reserved = 0;  // shared memory the driver keeps per block: dynamic shared memory starts behind it
  unsigned long long *acc_all = nullptr;  // one allocation: n_mates x ([cur_cap*97 rows][kNumCounters])
  std::vector<unsigned long long *> acc;  // per mate: acc_all + mate * acc_u64
  unsigned long long *reduce_buf = nullptr;  // same size as acc_all: NCCL receive buffer
  unsigned long long *d_scalar = nullptr;    // one u64: the ranks agree on the accumulator size through it
  uint32_t *d_bitmap = nullptr, *d_anchor = nullptr, *d_exact = nullptr;
  std::vector<Slot> slots;
  cudaStream_t main_stream = nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr;  // qb_timer_*
  ncclComm_t comm = nullptr;  // in-process communicator (n_devices > 1)
  uint32_t *l2_scratch = nullptr;
  size_t l2_words = 0;
  // text path: per mate the framing state, two carry buffers that take turns, the event of the last framed chunk
  struct TextMate {
    qb::TextState *d_state = nullptr;
    uint8_t *d_carry[2] = {nullptr, nullptr};
    int cur = 0;
    cudaEvent_t last_framed = nullptr;
    std::deque<int> deferred;  // slots whose chunk is framed (or being framed) and waits for its statistics launch (ctx->mu)
    int invalid = 0;        // a chunk was not canonical FASTQ (or a record outgrew the carry): use the host reader
    uint64_t tail_at_end = 0;  // bytes behind the last complete record when the stream ended
    uint64_t reads = 0;
  };
  std::vector<TextMate> text;
  // opt-in side outputs (qb_extras_enable): per mate [len_cap N counts | 94 mean-quality bins]
  std::vector<unsigned long long *> ext;
  unsigned long long *ext_reduce = nullptr;
};

}  // namespace

struct qb_ctx {
  qb_config cfg;
  std::vector<Device> dev;
  std::mutex mu;
  std::condition_variable cv_slot;  // a HELD slot was submitted (or released on an error path)
  uint64_t next_slot = 0, submit_seq = 0;
  std::mutex err_mu;
  std::string err;
  std::atomic<uint64_t> launches{0}, launches_fused{0}, launches_simple{0}, launches_period{0}, launches_flat{0};
  std::atomic<uint64_t> h2d_bytes{0};  // bytes queued for host-to-device copy by qb_submit*() so far
  // The accumulators hold cur_cap rows, not cfg.len_cap: they start small and grow (grow_accumulators) when a
  // batch announces a longer read, so a context opened for 2^20-bp reads costs nothing until one shows up.
  // Launches read the accumulator pointers under a shared lock, growth swaps them under the exclusive lock.
  std::shared_mutex acc_mu;
  uint32_t cur_cap = 0;
  size_t acc_u64 = 0;  // cur_cap*97 + counters, per mate
  std::atomic<bool> result_valid{false};  // h_result holds the reduced accumulators of every mate (qb_finish)
  std::mutex text_mu;                     // text path set-up
  std::atomic<bool> extras{false};        // qb_extras_enable(): every batch also takes the extras pass
  qb::AdapterSet ad_host_template{};
  uint32_t n_anchors = 0;     // distinct 7-mer anchors of the adapter set
  double anchor_density = 0;  // n_anchors / 2^14: filter pass rate per probe on random bases
  uint32_t qbase = 33;  // score bin s = q - qbase, s in [0,46] counted in shared memory
  ncclComm_t rank_comm = nullptr;  // multi-process communicator
  int n_ranks = 1, rank = 0;
  unsigned long long *h_result = nullptr;  // pinned staging for qb_finish
  const unsigned long long *d_result = nullptr;  // where reduce_all() left the summed accumulators on device 0
  // live profiling of kernel launches
  struct ProfRec { cudaEvent_t e0, e1; uint64_t bytes; int dev; };
  std::vector<ProfRec> prof;
  int prof_cap = 0;
};

struct qb_dbatch {
  int device_index;
  uint8_t *d_seq, *d_qual;
  uint32_t *d_off, *d_len;
  uint2 *d_tiles;
  uint32_t n_reads;
  uint64_t n_bytes;
  uint32_t max_len;
  uint32_t uniform_len, first_offset, contig_min_len;  // see qb::BatchView
};

namespace {

int fail(qb_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) {
    std::lock_guard<std::mutex> lk(ctx->err_mu);  // both mate threads may fail at once
    ctx->err = buf;
  } else {
    g_create_error = buf;
  }
  return code;
}

#define QB_CUDA(ctx, call)                                                                      \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return fail(ctx, QB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define QB_NCCL(ctx, call)                                                                      \
  do {                                                                                          \
    ncclResult_t r__ = (call);                                                                  \
    if (r__ != ncclSuccess)                                                                     \
      return fail(ctx, QB_ERR_NCCL, "%s failed: %s (%s:%d)", #call, nccl_api()->GetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

// (the flat kernel's unpredicated loads read up to 511 bytes behind the last read: qb_flat.cu)
inline size_t pad_bytes(uint64_t n) { return (size_t)((n + 15) & ~15ull) + 1088; }
inline size_t pad_reads(uint32_t n) { return ((size_t)n + 3 & ~(size_t)3) + 16; }

qb::AdapterSet adapter_set(const qb_ctx *ctx, const Device &d) {
  qb::AdapterSet a;
  a.bitmap = d.d_bitmap;
  a.anchor = d.d_anchor;
  a.exact = d.d_exact;
  a.enabled = ctx->cfg.adapters_enabled ? 1 : 0;
  return a;
}

qb::Accum accum(const qb_ctx *ctx, const Device &d, int mate) {
  qb::Accum a;
  a.rows = d.acc[mate];
  a.counters = d.acc[mate] + (size_t)ctx->cur_cap * qb::kRow;
  a.len_cap = ctx->cur_cap;
  return a;
}

// Host-side look at the batch shape, one vectorisable pass over the two u32 arrays: do the reads lie back to back
// (then *contig_min = the shortest read's length, else 0; the flat kernel needs that), and do they all have the same
// length (returned; 0: ragged -- the period kernel needs that).  *longest = the longest read.
uint32_t detect_shape(const uint32_t *offset, const uint32_t *length, uint32_t n_reads, uint32_t *first_offset,
                      uint32_t *contig_min, uint32_t *longest) {
  *contig_min = 0, *longest = 0, *first_offset = 0;
  if (n_reads == 0) return 0;
  const uint32_t l = length[0], o0 = offset[0];
  uint32_t diff = 0, gaps = 0, mn = 0xFFFFFFFFu, mx = 0;
  for (uint32_t r = 0; r < n_reads; r++) {
    const uint32_t lr = length[r];
    diff |= lr ^ l;
    mn = lr < mn ? lr : mn;
    mx = lr > mx ? lr : mx;
  }
  for (uint32_t r = 0; r + 1 < n_reads; r++) gaps |= (offset[r] + length[r]) ^ offset[r + 1];
  *first_offset = o0;
  *longest = mx;
  *contig_min = gaps ? 0u : mn;
  return (diff || gaps) ? 0u : l;
}

// one launch of the v4 / v3 / simple kernel on a batch view
int launch_other(qb_ctx *ctx, Device &d, const qb::BatchView &v, qb::Accum ac, const qb::AdapterSet &ad, int kernel,
                 cudaStream_t stream) {
  qb::FusedPlan plan{};
  if (kernel == QB_KERNEL_WTILE) kernel = QB_KERNEL_FUSED;  // (the v4 warp-tile kernel is gone: slower than v3 everywhere)
  if (kernel == QB_KERNEL_AUTO || kernel == QB_KERNEL_FLAT) {
    // ragged batches of back-to-back reads: the flat kernel (lane <-> 16-byte unit) when the batch has its shape
    qb::FlatPlan fp{};
    if (v.contig_min_len >= 16u && !(v.first_offset & 15u) && v.tiles && v.n_bytes < 0xFFFFFF00ull && !getenv("QB_NO_FLAT"))
      fp = qb::flat_plan(v.max_len && v.max_len < ctx->cur_cap ? v.max_len : ctx->cur_cap, v.contig_min_len, ad.enabled,
                         d.sm_count, (uint32_t)d.smem_optin, (uint32_t)d.smem_reserved, ctx->qbase);
    if (fp.ok) {
      ctx->launches++;
      ctx->launches_fused++;
      ctx->launches_flat++;
      const cudaError_t e = qb::launch_flat(v, ac, ad, fp, stream);
      if (e != cudaSuccess) return fail(ctx, QB_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
      return QB_OK;
    }
    if (kernel == QB_KERNEL_FLAT)
      return fail(ctx, QB_ERR_CAPACITY, "the flat kernel needs a batch of back-to-back reads of 16..320 bp that starts on a 16-byte boundary");
  }
  if (kernel != QB_KERNEL_SIMPLE) {
    // The shared-memory histogram is sized by the longest read of THIS batch (the caller's max_len
    // promise; a longer read is counted as an error and fails qb_finish), not by len_cap: a context
    // opened for 65536-bp reads still runs short-read batches on the shared-memory kernels.
    uint32_t eff_cap = ctx->cur_cap;
    if (v.max_len && v.max_len < eff_cap) eff_cap = v.max_len < 11u ? 11u : v.max_len;
    plan = qb::fused_plan(eff_cap, v.max_len, ad.enabled, d.sm_count, (uint32_t)d.smem_optin, ctx->qbase);
    if (plan.ok) {
      kernel = QB_KERNEL_FUSED;
      ac.len_cap = eff_cap;
    } else {
      if (kernel != QB_KERNEL_AUTO)
        return fail(ctx, QB_ERR_CAPACITY, "reads of up to %u bp do not fit the shared-memory histogram of kernel %d",
                    eff_cap, kernel);
      kernel = QB_KERNEL_SIMPLE;
    }
  }
  ctx->launches++;
  (kernel == QB_KERNEL_SIMPLE ? ctx->launches_simple : ctx->launches_fused)++;
  cudaError_t e = kernel == QB_KERNEL_FUSED ? qb::launch_fused(v, ac, ad, plan, stream)
                                            : qb::launch_simple(v, ac, ad, d.sm_count, stream);
  if (e != cudaSuccess) return fail(ctx, QB_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return QB_OK;
}

// the period-kernel plan for a batch view, ok = 0 if the batch goes to the other kernels
qb::PeriodPlan period_plan_for(const qb_ctx *ctx, const Device &d, const qb::BatchView &v, int adapters) {
  const int kernel = ctx->cfg.kernel;
  qb::PeriodPlan p{};
  if ((kernel == QB_KERNEL_AUTO || kernel == QB_KERNEL_PERIOD) && v.uniform_len && v.uniform_len <= ctx->cfg.len_cap &&
      (!v.max_len || v.uniform_len <= v.max_len))
    p = qb::period_plan(v.uniform_len, v.first_offset, adapters, d.sm_count, (uint32_t)d.smem_optin,
                        (uint32_t)d.smem_reserved, ctx->qbase);
  return p;
}

// (Re)allocates the accumulators of every device for `cap` rows per mate, keeping what was counted so far.
// Caller holds acc_mu exclusively (or is qb_create).
int alloc_accumulators(qb_ctx *ctx, uint32_t cap) {
  const size_t new_u64 = (size_t)cap * qb::kRow + qb::kNumCounters;
  const int nm = ctx->cfg.n_mates;
  for (Device &d : ctx->dev) {
    QB_CUDA(ctx, cudaSetDevice(d.id));
    if (d.acc_all) QB_CUDA(ctx, cudaDeviceSynchronize());  // kernels in flight still count into the old rows
    unsigned long long *fresh = nullptr, *rbuf = nullptr;
    QB_CUDA(ctx, cudaMalloc(&fresh, new_u64 * 8 * nm));
    QB_CUDA(ctx, cudaMemset(fresh, 0, new_u64 * 8 * nm));
    QB_CUDA(ctx, cudaMalloc(&rbuf, new_u64 * 8 * nm));
    if (d.acc_all) {
      for (int m = 0; m < nm; m++) {
        QB_CUDA(ctx, cudaMemcpy(fresh + m * new_u64, d.acc[m], (size_t)ctx->cur_cap * qb::kRow * 8, cudaMemcpyDeviceToDevice));
        QB_CUDA(ctx, cudaMemcpy(fresh + m * new_u64 + (size_t)cap * qb::kRow, d.acc[m] + (size_t)ctx->cur_cap * qb::kRow,
                                qb::kNumCounters * 8, cudaMemcpyDeviceToDevice));
      }
      cudaFree(d.acc_all);
      cudaFree(d.reduce_buf);
    }
    d.acc_all = fresh;
    d.reduce_buf = rbuf;
    d.acc.resize(nm);
    for (int m = 0; m < nm; m++) d.acc[m] = fresh + m * new_u64;
    if (!d.d_scalar) QB_CUDA(ctx, cudaMalloc(&d.d_scalar, 8));
  }
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  ctx->h_result = nullptr;
  QB_CUDA(ctx, cudaHostAlloc(&ctx->h_result, new_u64 * 8 * nm, cudaHostAllocDefault));
  ctx->cur_cap = cap;
  ctx->acc_u64 = new_u64;
  ctx->result_valid = false;
  return QB_OK;
}

uint32_t grown_cap(const qb_ctx *ctx, uint32_t need) {
  uint32_t cap = ctx->cur_cap ? ctx->cur_cap : 512u;
  while (cap < need) cap *= 2u;
  return cap > ctx->cfg.len_cap ? ctx->cfg.len_cap : cap;
}

// makes room for reads of up to `need` bp (<= cfg.len_cap); cheap when the rows exist already
int ensure_cap(qb_ctx *ctx, uint32_t need) {
  {
    std::shared_lock<std::shared_mutex> lk(ctx->acc_mu);
    if (need <= ctx->cur_cap) return QB_OK;
  }
  std::unique_lock<std::shared_mutex> lk(ctx->acc_mu);
  if (need <= ctx->cur_cap) return QB_OK;
  return alloc_accumulators(ctx, grown_cap(ctx, need));
}

int launch_batch_locked(qb_ctx *ctx, Device &d, const qb::BatchView &v, int mate, cudaStream_t stream);

// chooses and launches the statistics kernel(s) for one device-resident batch
int launch_batch(qb_ctx *ctx, Device &d, const qb::BatchView &v, int mate, cudaStream_t stream) {
  if (v.n_reads == 0) return QB_OK;
  // rows for the longest read this batch may hold (max_len = 0: unknown, whatever len_cap allows)
  const uint32_t need = v.max_len && v.max_len < ctx->cfg.len_cap ? v.max_len : ctx->cfg.len_cap;
  const int rc = ensure_cap(ctx, need);
  if (rc) return rc;
  ctx->result_valid = false;
  std::shared_lock<std::shared_mutex> lk(ctx->acc_mu);
  QB_CUDA(ctx, cudaSetDevice(d.id));  // growth may have switched this thread's device
  return launch_batch_locked(ctx, d, v, mate, stream);
}

int launch_batch_locked(qb_ctx *ctx, Device &d, const qb::BatchView &v, int mate, cudaStream_t stream) {
  const qb::AdapterSet ad = adapter_set(ctx, d);
  qb::Accum ac = accum(ctx, d, mate);
  int kernel = ctx->cfg.kernel;
  const qb::PeriodPlan pplan = period_plan_for(ctx, d, v, ad.enabled);
  if (kernel == QB_KERNEL_PERIOD) {
    if (!pplan.ok)
      return fail(ctx, QB_ERR_CAPACITY, "the period kernel needs a batch of back-to-back reads of one length in [32, %u]",
                  qb::kPeriodMaxLen);
    kernel = QB_KERNEL_AUTO;  // for the reads that do not fill a tile
  }
  qb_ctx::ProfRec *rec = nullptr;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if ((int)ctx->prof.size() < ctx->prof_cap) {
      qb_ctx::ProfRec r{};
      if (cudaEventCreate(&r.e0) == cudaSuccess && cudaEventCreate(&r.e1) == cudaSuccess) {
        r.bytes = 2 * v.n_bytes + 8ull * v.n_reads;
        r.dev = d.id;
        ctx->prof.push_back(r);
        rec = &ctx->prof.back();
      }
    }
  }
  cudaEvent_t e0 = rec ? rec->e0 : nullptr, e1 = rec ? rec->e1 : nullptr;
  if (e0) cudaEventRecord(e0, stream);
  uint32_t n_main = 0;
  int rc = QB_OK;
  if (pplan.ok) {
    const cudaError_t e = qb::launch_period(v, ac, ad, pplan, stream, &n_main);
    if (e != cudaSuccess) return fail(ctx, QB_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    if (n_main) {
      ctx->launches++;
      ctx->launches_fused++;
      ctx->launches_period++;
    }
  }
  if (n_main < v.n_reads) {  // everything, or the reads that do not fill a tile of the period kernel
    qb::BatchView rest = v;
    rest.offset += n_main;
    rest.length += n_main;
    rest.n_reads -= n_main;
    if (n_main) {  // the reads behind the period kernel's tiles: still back to back, one length
      rest.first_offset = v.first_offset + n_main * v.uniform_len;
      rest.n_bytes = (uint64_t)rest.n_reads * v.uniform_len;
    }
    rc = launch_other(ctx, d, rest, ac, ad, kernel, stream);
  }
  if (e1) cudaEventRecord(e1, stream);
  if (rc == QB_OK && ctx->extras.load()) {
    const cudaError_t e = qb::launch_extras(v, ctx->cfg.len_cap, d.ext[mate], d.ext[mate] + ctx->cfg.len_cap, d.sm_count, stream);
    if (e != cudaSuccess) return fail(ctx, QB_ERR_CUDA, "extras kernel launch failed: %s", cudaGetErrorString(e));
    ctx->launches++;
  }
  return rc;
}

int check_mate(qb_ctx *ctx, int mate) {
  if (!ctx) return QB_ERR_ARG;
  if (mate < 0 || mate >= ctx->cfg.n_mates) return fail(ctx, QB_ERR_ARG, "mate %d out of range", mate);
  return QB_OK;
}

}  // namespace

extern "C" {

int qb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return QB_ERR_CUDA;
  return n;
}

const char *qb_last_error(const qb_ctx *ctx) {
  if (!ctx) return g_create_error.c_str();
  // a copy per calling thread: another thread may replace ctx->err while the caller still reads the text
  thread_local std::string copy;
  {
    std::lock_guard<std::mutex> lk(const_cast<qb_ctx *>(ctx)->err_mu);
    copy = ctx->err;
  }
  return copy.c_str();
}

void *qb_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void qb_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int qb_create(const qb_config *cfg_in, qb_ctx **out) {
  if (!cfg_in || !out) return fail(nullptr, QB_ERR_ARG, "null argument");
  *out = nullptr;
  qb_config cfg = *cfg_in;
  if (cfg.n_devices < 1) return fail(nullptr, QB_ERR_ARG, "n_devices must be >= 1");
  if (cfg.n_mates < 1 || cfg.n_mates > 2) return fail(nullptr, QB_ERR_ARG, "n_mates must be 1 or 2");
  if (cfg.len_cap < 11 || cfg.len_cap > (1u << 20)) return fail(nullptr, QB_ERR_ARG, "len_cap must be in [11, 2^20]");
  if (cfg.batch_bytes == 0) cfg.batch_bytes = 64ull << 20;
  if (cfg.batch_bytes > 0xFFFFFF00ull - 64) return fail(nullptr, QB_ERR_ARG, "batch_bytes must stay below 4 GiB (u32 offsets)");
  if (cfg.batch_reads == 0) cfg.batch_reads = (uint32_t)(cfg.batch_bytes / 32 + 1);
  if (cfg.ring_depth == 0) cfg.ring_depth = 3;
  if (cfg.ring_depth < 1) return fail(nullptr, QB_ERR_ARG, "ring_depth must be >= 1");
  // QB_CREATE_TIMING=1: where the start-up goes (stderr)
  const bool timing = getenv("QB_CREATE_TIMING") && atoi(getenv("QB_CREATE_TIMING")) > 0;
  auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tlast = tnow();
  auto lap = [&](const char *what) {
    if (!timing) return;
    const double t = tnow();
    fprintf(stderr, "qb_create: %-28s %8.1f ms\n", what, (t - tlast) * 1e3);
    tlast = t;
  };
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  lap("cudaGetDeviceCount");
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, QB_ERR_CUDA, "no usable CUDA device (%s); quack_b200 has no CPU fallback",
                ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");

  qb_ctx *ctx = new qb_ctx();
  ctx->cfg = cfg;
  ctx->cfg.device_ids = nullptr;
  ctx->cfg.adapter_keys = nullptr;
  if (const char *qb_env = getenv("QB_QBASE")) {
    const int v = atoi(qb_env);
    if (v >= 33 && v <= 64) ctx->qbase = (uint32_t)v;
  }

  std::vector<uint32_t> bitmap, anchor, exact;
  if (cfg.adapters_enabled) {
    if (cfg.n_adapter_keys && !cfg.adapter_keys) {
      delete ctx;
      return fail(nullptr, QB_ERR_ARG, "adapter_keys is NULL");
    }
    qb::build_adapter_images(cfg.adapter_keys, cfg.n_adapter_keys, bitmap, anchor, exact, ctx->n_anchors,
                             ctx->anchor_density);
  }

#define QB_CREATE_CUDA(call)                                                                          \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) {                                                                         \
      fail(nullptr, QB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      qb_destroy(ctx);                                                                                \
      return QB_ERR_CUDA;                                                                             \
    }                                                                                                 \
  } while (0)

  ctx->dev.resize(cfg.n_devices);
  for (int i = 0; i < cfg.n_devices; i++) {
    Device &d = ctx->dev[i];
    d.id = cfg_in->device_ids ? cfg_in->device_ids[i] : i;
    if (d.id < 0 || d.id >= ndev) {
      fail(nullptr, QB_ERR_ARG, "device id %d not available (%d visible)", d.id, ndev);
      qb_destroy(ctx);
      return QB_ERR_ARG;
    }
    QB_CREATE_CUDA(cudaSetDevice(d.id));
    QB_CREATE_CUDA(cudaFree(nullptr));
    lap("context (cudaSetDevice)");
    QB_CREATE_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.id));
    QB_CREATE_CUDA(cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.id));
    QB_CREATE_CUDA(cudaDeviceGetAttribute(&d.smem_reserved, cudaDevAttrReservedSharedMemoryPerBlock, d.id));
    {  // opt every kernel in to the large dynamic shared memory: once per device and process (42 + 4 kernels)
      static std::mutex cfg_mu;
      static bool configured[256] = {false};
      std::lock_guard<std::mutex> lk(cfg_mu);
      if (d.id >= 256 || !configured[d.id]) {
        QB_CREATE_CUDA(qb::fused_configure());
        lap("fused_configure");
        QB_CREATE_CUDA(qb::period_configure());
        lap("period_configure");
        QB_CREATE_CUDA(qb::flat_configure());
        lap("flat_configure");
        if (d.id < 256) configured[d.id] = true;
      }
    }
    QB_CREATE_CUDA(cudaStreamCreateWithFlags(&d.main_stream, cudaStreamNonBlocking));
    if (cfg.adapters_enabled) {
      QB_CREATE_CUDA(cudaMalloc(&d.d_bitmap, bitmap.size() * 4));
      QB_CREATE_CUDA(cudaMemcpy(d.d_bitmap, bitmap.data(), bitmap.size() * 4, cudaMemcpyHostToDevice));
      QB_CREATE_CUDA(cudaMalloc(&d.d_anchor, anchor.size() * 4));
      QB_CREATE_CUDA(cudaMemcpy(d.d_anchor, anchor.data(), anchor.size() * 4, cudaMemcpyHostToDevice));
      if (!exact.empty()) {
        QB_CREATE_CUDA(cudaMalloc(&d.d_exact, exact.size() * 4));
        QB_CREATE_CUDA(cudaMemcpy(d.d_exact, exact.data(), exact.size() * 4, cudaMemcpyHostToDevice));
      }
    }
    lap("adapter images");
    d.slots.resize(cfg.ring_depth);
    for (Slot &s : d.slots) {
      QB_CREATE_CUDA(cudaHostAlloc(&s.h_seq, pad_bytes(cfg.batch_bytes), cudaHostAllocDefault));
      QB_CREATE_CUDA(cudaHostAlloc(&s.h_qual, pad_bytes(cfg.batch_bytes), cudaHostAllocDefault));
      QB_CREATE_CUDA(cudaHostAlloc(&s.h_off, pad_reads(cfg.batch_reads) * 4, cudaHostAllocDefault));
      QB_CREATE_CUDA(cudaHostAlloc(&s.h_len, pad_reads(cfg.batch_reads) * 4, cudaHostAllocDefault));
      lap("slot: pinned host buffers");
      QB_CREATE_CUDA(cudaMalloc(&s.d_seq, pad_bytes(cfg.batch_bytes)));
      QB_CREATE_CUDA(cudaMalloc(&s.d_qual, pad_bytes(cfg.batch_bytes)));
      QB_CREATE_CUDA(cudaMalloc(&s.d_off, pad_reads(cfg.batch_reads) * 4));
      QB_CREATE_CUDA(cudaMalloc(&s.d_len, pad_reads(cfg.batch_reads) * 4));
      QB_CREATE_CUDA(cudaMalloc(&s.d_tiles, pad_reads(cfg.batch_reads) * sizeof(uint2)));
      QB_CREATE_CUDA(cudaMemset(s.d_seq, 0, pad_bytes(cfg.batch_bytes)));
      QB_CREATE_CUDA(cudaMemset(s.d_qual, 0, pad_bytes(cfg.batch_bytes)));
      QB_CREATE_CUDA(cudaMemset(s.d_off, 0, pad_reads(cfg.batch_reads) * 4));
      QB_CREATE_CUDA(cudaMemset(s.d_len, 0, pad_reads(cfg.batch_reads) * 4));
      QB_CREATE_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
      QB_CREATE_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
      lap("slot: device buffers");
    }
    QB_CREATE_CUDA(cudaDeviceSynchronize());
    lap("synchronize");
  }
  if (alloc_accumulators(ctx, grown_cap(ctx, cfg.len_cap < 512u ? cfg.len_cap : 512u)) != QB_OK) {
    g_create_error = ctx->err;
    qb_destroy(ctx);
    return QB_ERR_CUDA;
  }
  lap("accumulators");
  if (cfg.n_devices > 1) {
    std::vector<ncclComm_t> comms(cfg.n_devices);
    std::vector<int> ids(cfg.n_devices);
    for (int i = 0; i < cfg.n_devices; i++) ids[i] = ctx->dev[i].id;
    NcclApi *nc = nccl_api();
    ncclResult_t r = nc ? nc->CommInitAll(comms.data(), cfg.n_devices, ids.data()) : ncclSystemError;
    if (r != ncclSuccess) {
      fail(nullptr, QB_ERR_NCCL, "ncclCommInitAll failed: %s", nc ? nc->GetErrorString(r) : "libnccl.so.2 not found");
      qb_destroy(ctx);
      return QB_ERR_NCCL;
    }
    for (int i = 0; i < cfg.n_devices; i++) ctx->dev[i].comm = comms[i];
  }
#undef QB_CREATE_CUDA
  *out = ctx;
  return QB_OK;
}

void qb_destroy(qb_ctx *ctx) {
  if (!ctx) return;
  if (ctx->rank_comm) nccl_api()->CommDestroy(ctx->rank_comm);
  for (Device &d : ctx->dev) {
    cudaSetDevice(d.id);
    cudaDeviceSynchronize();
    if (d.comm) nccl_api()->CommDestroy(d.comm);
    for (unsigned long long *x : d.ext) cudaFree(x);
    cudaFree(d.ext_reduce);
    for (Slot &s : d.slots) {
      if (s.h_seq) cudaFreeHost(s.h_seq);
      if (s.h_qual) cudaFreeHost(s.h_qual);
      if (s.h_off) cudaFreeHost(s.h_off);
      if (s.h_len) cudaFreeHost(s.h_len);
      cudaFree(s.d_seq);
      cudaFree(s.d_qual);
      cudaFree(s.d_off);
      cudaFree(s.d_len);
      cudaFree(s.d_tiles);
      if (s.stream) cudaStreamDestroy(s.stream);
      if (s.done) cudaEventDestroy(s.done);
      if (s.h_text) cudaFreeHost(s.h_text);
      cudaFree(s.d_text);
      cudaFree(s.d_scratch);
      if (s.h_sum) cudaFreeHost(s.h_sum);
      if (s.h_blk) cudaFreeHost(s.h_blk);
      cudaFree(s.d_comp);
      cudaFree(s.d_blk);
      cudaFree(s.d_blk_status);
      cudaFree(s.d_sum);
      if (s.framed) cudaEventDestroy(s.framed);
    }
    for (auto &m : d.text) {
      cudaFree(m.d_state);
      cudaFree(m.d_carry[0]);
      cudaFree(m.d_carry[1]);
    }
    cudaFree(d.acc_all);
    cudaFree(d.reduce_buf);
    cudaFree(d.d_scalar);
    cudaFree(d.d_bitmap);
    cudaFree(d.d_anchor);
    cudaFree(d.d_exact);
    cudaFree(d.l2_scratch);
    if (d.t0) cudaEventDestroy(d.t0);
    if (d.t1) cudaEventDestroy(d.t1);
    if (d.main_stream) cudaStreamDestroy(d.main_stream);
  }
  for (auto &r : ctx->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  delete ctx;
}

namespace {
int flush_deferred(qb_ctx *ctx, int si);
}

// Hands the calling thread a slot nobody else holds: the next FREE slot in round-robin order over devices x ring,
// else the PENDING slot that was submitted first (its event is waited for outside the lock), else -- every slot is
// HELD by other threads -- it sleeps until one of them submits.  The slot stays HELD (exclusively the caller's)
// until submit_on() queues its work or release_slot() gives it back.
static int take_slot(qb_ctx *ctx, int *dev_index, int *slot_index) {
  const uint64_t total = (uint64_t)ctx->dev.size() * ctx->cfg.ring_depth;
  int di = -1, si = -1;
  bool wait_event = false;
  {
    std::unique_lock<std::mutex> lk(ctx->mu);
    for (;;) {
      int best = -1;
      uint64_t best_seq = 0;
      for (uint64_t i = 0; i < total && di < 0; i++) {
        const uint64_t k = (ctx->next_slot + i) % total;
        Slot &s = ctx->dev[k % ctx->dev.size()].slots[k / ctx->dev.size()];
        if (s.state == Slot::FREE) {
          di = (int)(k % ctx->dev.size()), si = (int)(k / ctx->dev.size());
          ctx->next_slot = k + 1;
        } else if (s.state == Slot::PENDING && (best < 0 || s.seq < best_seq)) {
          best = (int)k, best_seq = s.seq;
        }
      }
      if (di < 0 && best >= 0) {
        di = (int)((uint64_t)best % ctx->dev.size()), si = (int)((uint64_t)best / ctx->dev.size());
        ctx->next_slot = (uint64_t)best + 1;
        wait_event = true;
      }
      if (di >= 0) break;
      // nothing free and nothing pending: a chunk of the text path may be waiting for its statistics launch
      int deferred = -1;
      for (size_t i = 0; i < ctx->dev[0].slots.size() && !ctx->dev[0].text.empty(); i++)
        if (ctx->dev[0].slots[i].state == Slot::DEFERRED) deferred = (int)i;
      if (deferred >= 0) {
        lk.unlock();
        const int frc = flush_deferred(ctx, deferred);
        lk.lock();
        if (frc) return frc;
        continue;
      }
      ctx->cv_slot.wait(lk);
    }
    ctx->dev[di].slots[si].state = Slot::HELD;
  }
  Device &d = ctx->dev[di];
  Slot &s = d.slots[si];
  cudaError_t e = cudaSetDevice(d.id);
  if (e == cudaSuccess && wait_event) e = cudaEventSynchronize(s.done);
  if (e != cudaSuccess) {
    {
      std::lock_guard<std::mutex> lk(ctx->mu);
      s.state = Slot::FREE;
    }
    ctx->cv_slot.notify_all();
    return fail(ctx, QB_ERR_CUDA, "waiting for a ring slot failed: %s", cudaGetErrorString(e));
  }
  *dev_index = di;
  *slot_index = si;
  return QB_OK;
}

// error paths: a HELD slot goes back to the ring so that other threads do not wait for it forever
static void release_slot(qb_ctx *ctx, int di, int si) {
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->dev[di].slots[si].state = Slot::FREE;
  }
  ctx->cv_slot.notify_all();
}

int qb_acquire(qb_ctx *ctx, qb_batch *out) {
  if (!ctx || !out) return QB_ERR_ARG;
  int di, si;
  int rc = take_slot(ctx, &di, &si);
  if (rc) return rc;
  Slot &s = ctx->dev[di].slots[si];
  out->seq = s.h_seq;
  out->qual = s.h_qual;
  out->offset = s.h_off;
  out->length = s.h_len;
  out->cap_bytes = ctx->cfg.batch_bytes;
  out->cap_reads = ctx->cfg.batch_reads;
  out->device_index = di;
  out->slot = si;
  return QB_OK;
}

static int submit_queue(qb_ctx *ctx, int di, int si, int mate, const uint8_t *seq, const uint8_t *qual,
                        const uint32_t *offset, const uint32_t *length, uint32_t n_reads, uint64_t n_bytes,
                        uint32_t max_len);

// queues the work of a HELD slot and turns it PENDING (FREE again if queuing failed)
static int submit_on(qb_ctx *ctx, int di, int si, int mate, const uint8_t *seq, const uint8_t *qual,
                     const uint32_t *offset, const uint32_t *length, uint32_t n_reads, uint64_t n_bytes,
                     uint32_t max_len) {
  const int rc = submit_queue(ctx, di, si, mate, seq, qual, offset, length, n_reads, n_bytes, max_len);
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    Slot &s = ctx->dev[di].slots[si];
    s.state = rc == QB_OK ? Slot::PENDING : Slot::FREE;
    s.seq = ++ctx->submit_seq;
  }
  ctx->cv_slot.notify_all();
  return rc;
}

static int submit_queue(qb_ctx *ctx, int di, int si, int mate, const uint8_t *seq, const uint8_t *qual,
                        const uint32_t *offset, const uint32_t *length, uint32_t n_reads, uint64_t n_bytes,
                        uint32_t max_len) {
  if (n_reads > ctx->cfg.batch_reads || n_bytes > ctx->cfg.batch_bytes)
    return fail(ctx, QB_ERR_CAPACITY, "batch of %u reads / %llu bytes exceeds the slot (%u / %llu)", n_reads,
                (unsigned long long)n_bytes, ctx->cfg.batch_reads, (unsigned long long)ctx->cfg.batch_bytes);
  Device &d = ctx->dev[di];
  Slot &s = d.slots[si];
  QB_CUDA(ctx, cudaSetDevice(d.id));
  if (n_reads) {
    if (n_bytes) {
      QB_CUDA(ctx, cudaMemcpyAsync(s.d_seq, seq, n_bytes, cudaMemcpyHostToDevice, s.stream));
      QB_CUDA(ctx, cudaMemcpyAsync(s.d_qual, qual, n_bytes, cudaMemcpyHostToDevice, s.stream));
    }
    qb::BatchView v{s.d_seq, s.d_qual, s.d_off, s.d_len, n_reads, n_bytes, max_len ? max_len : ctx->cfg.len_cap, s.d_tiles};
    if (ctx->cfg.kernel == QB_KERNEL_AUTO || ctx->cfg.kernel == QB_KERNEL_PERIOD || ctx->cfg.kernel == QB_KERNEL_FLAT) {
      uint32_t longest = 0;
      v.uniform_len = detect_shape(offset, length, n_reads, &v.first_offset, &v.contig_min_len, &longest);
      if (longest > v.max_len) v.contig_min_len = 0;  // a longer read than promised: the kernels that check every
                                                      // read's length take the batch and fail it loudly
      if (longest && longest < v.max_len) v.max_len = longest;
    }
    // The period kernel needs no offsets / lengths on the device (the host just verified the batch shape): only
    // those of the reads it leaves to the other kernels (< reads_per_tile at the end of the batch) are copied.
    uint32_t r0 = 0;
    const qb::PeriodPlan pp = period_plan_for(ctx, d, v, ctx->cfg.adapters_enabled ? 1 : 0);
    if (pp.ok) r0 = n_reads / pp.reads_per_tile * pp.reads_per_tile;
    if (r0 < n_reads) {
      QB_CUDA(ctx, cudaMemcpyAsync(s.d_off + r0, offset + r0, (size_t)(n_reads - r0) * 4, cudaMemcpyHostToDevice, s.stream));
      QB_CUDA(ctx, cudaMemcpyAsync(s.d_len + r0, length + r0, (size_t)(n_reads - r0) * 4, cudaMemcpyHostToDevice, s.stream));
    }
    ctx->h2d_bytes += 2 * n_bytes + 8ull * (n_reads - r0);
    int rc = launch_batch(ctx, d, v, mate, s.stream);
    if (rc) return rc;
  }
  QB_CUDA(ctx, cudaEventRecord(s.done, s.stream));
  return QB_OK;
}

int qb_submit(qb_ctx *ctx, const qb_batch *b, int mate, uint32_t n_reads, uint64_t n_bytes, uint32_t max_len) {
  int rc = check_mate(ctx, mate);
  if (rc) return rc;
  if (!b || b->device_index < 0 || b->device_index >= (int)ctx->dev.size() || b->slot < 0 ||
      b->slot >= ctx->cfg.ring_depth)
    return fail(ctx, QB_ERR_ARG, "bad batch handle");
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    const Slot &s = ctx->dev[b->device_index].slots[b->slot];
    if (s.state != Slot::HELD || b->seq != s.h_seq)
      return fail(ctx, QB_ERR_ARG, "qb_submit: slot %d of device %d is not held by a qb_acquire() (submitted twice?)",
                  b->slot, b->device_index);
  }
  return submit_on(ctx, b->device_index, b->slot, mate, b->seq, b->qual, b->offset, b->length, n_reads, n_bytes,
                   max_len);
}

int qb_submit_from(qb_ctx *ctx, int mate, const uint8_t *seq, const uint8_t *qual, const uint32_t *offset,
                   const uint32_t *length, uint32_t n_reads, uint64_t n_bytes, uint32_t max_len) {
  int rc = check_mate(ctx, mate);
  if (rc) return rc;
  int di, si;
  rc = take_slot(ctx, &di, &si);
  if (rc) return rc;
  return submit_on(ctx, di, si, mate, seq, qual, offset, length, n_reads, n_bytes, max_len);
}

int qb_accumulate_host(qb_ctx *ctx, int mate, const uint8_t *seq, const uint8_t *qual, const uint32_t *offset,
                       const uint32_t *length, uint64_t n_reads) {
  int rc = check_mate(ctx, mate);
  if (rc) return rc;
  uint64_t r0 = 0;
  while (r0 < n_reads) {
    qb_batch b;
    rc = qb_acquire(ctx, &b);
    if (rc) return rc;
    const uint64_t base = offset[r0];
    uint64_t r1 = r0, end = base;
    uint32_t max_len = 0;
    while (r1 < n_reads && r1 - r0 < b.cap_reads) {
      const uint64_t o = offset[r1], l = length[r1];
      if (o < end) {
        release_slot(ctx, b.device_index, b.slot);
        return fail(ctx, QB_ERR_LAYOUT, "offsets must be ascending and reads must not overlap (read %llu)",
                    (unsigned long long)r1);
      }
      if (l > ctx->cfg.len_cap) {
        release_slot(ctx, b.device_index, b.slot);
        return fail(ctx, QB_ERR_CAPACITY, "read %llu has length %llu > len_cap %u", (unsigned long long)r1,
                    (unsigned long long)l, ctx->cfg.len_cap);
      }
      if (o + l - base > b.cap_bytes) break;
      b.offset[r1 - r0] = (uint32_t)(o - base);
      b.length[r1 - r0] = (uint32_t)l;
      if (l > max_len) max_len = (uint32_t)l;
      end = o + l;
      r1++;
    }
    if (r1 == r0) {
      release_slot(ctx, b.device_index, b.slot);
      return fail(ctx, QB_ERR_CAPACITY, "read %llu does not fit an empty slot", (unsigned long long)r0);
    }
    memcpy(b.seq, seq + base, end - base);
    memcpy(b.qual, qual + base, end - base);
    rc = qb_submit(ctx, &b, mate, (uint32_t)(r1 - r0), end - base, max_len);
    if (rc) return rc;
    r0 = r1;
  }
  return QB_OK;
}

// ---------------------------------------------------------------------- text path (device framing)

namespace {
constexpr uint32_t kTextCarry = 1u << 20;  // a record (4 lines) must fit: reads of up to ~500 kbp

// [carry | chunk] holds at most 2 x batch_bytes of text, so the packed bases (fewer than half of a record's text) fit the slot
uint32_t text_carry_cap(const qb_ctx *ctx) { return (uint32_t)std::min<uint64_t>(kTextCarry, ctx->cfg.batch_bytes / 2); }
uint32_t text_cap(const qb_ctx *ctx) {
  const uint64_t c = 2ull * ctx->cfg.batch_bytes - text_carry_cap(ctx);
  return (uint32_t)std::min<uint64_t>(c, 0xE0000000ull);
}
uint32_t text_nl_cap(const qb_ctx *ctx) { return 4u * ctx->cfg.batch_reads + 16u; }

int text_setup(qb_ctx *ctx) {  // once per context: per-mate state, per-slot text buffers
  std::lock_guard<std::mutex> lk(ctx->text_mu);
  Device &d = ctx->dev[0];
  if (!d.text.empty()) return QB_OK;
  if (ctx->dev.size() != 1) return fail(ctx, QB_ERR_ARG, "the text path needs a one-device context");
  QB_CUDA(ctx, cudaSetDevice(d.id));
  std::vector<Device::TextMate> tm(ctx->cfg.n_mates);
  for (auto &m : tm) {
    QB_CUDA(ctx, cudaMalloc(&m.d_state, sizeof(qb::TextState)));
    QB_CUDA(ctx, cudaMemset(m.d_state, 0, sizeof(qb::TextState)));
    for (int k = 0; k < 2; k++) QB_CUDA(ctx, cudaMalloc(&m.d_carry[k], kTextCarry + 64));
  }
  const size_t words = qb::text_scratch_words(text_cap(ctx), text_carry_cap(ctx), text_nl_cap(ctx), ctx->cfg.batch_reads);
  for (Slot &s : d.slots) {
    QB_CUDA(ctx, cudaHostAlloc(&s.h_text, (size_t)text_cap(ctx) + 64, cudaHostAllocDefault));
    QB_CUDA(ctx, cudaMalloc(&s.d_text, (size_t)text_cap(ctx) + 64));
    QB_CUDA(ctx, cudaMalloc(&s.d_scratch, words * 4));
    QB_CUDA(ctx, cudaHostAlloc(&s.h_sum, sizeof(qb::TextSummary), cudaHostAllocDefault));
    QB_CUDA(ctx, cudaMalloc(&s.d_sum, sizeof(qb::TextSummary)));
    QB_CUDA(ctx, cudaEventCreateWithFlags(&s.framed, cudaEventDisableTiming));
  }
  d.text.swap(tm);
  return QB_OK;
}

// the chunk of a DEFERRED slot is framed: read its summary, launch the statistics kernel, the slot turns PENDING
int flush_deferred(qb_ctx *ctx, int si) {
  Device &d = ctx->dev[0];
  Slot &s = d.slots[si];
  {  // any thread may flush any deferred slot (its owner, or one that needs a slot): exactly one of them does
    std::unique_lock<std::mutex> lk(ctx->mu);
    while (s.state == Slot::FLUSHING) ctx->cv_slot.wait(lk);
    if (s.state != Slot::DEFERRED) return QB_OK;
    s.state = Slot::FLUSHING;
  }
  cudaError_t ce = cudaSetDevice(d.id);
  if (ce == cudaSuccess) ce = cudaEventSynchronize(s.framed);
  if (ce != cudaSuccess) {
    {
      std::lock_guard<std::mutex> lk(ctx->mu);
      s.state = Slot::FREE;
      d.text[s.text_mate].invalid = 1;
    }
    ctx->cv_slot.notify_all();
    return fail(ctx, QB_ERR_CUDA, "framing failed: %s", cudaGetErrorString(ce));
  }
  const qb::TextSummary sum = *s.h_sum;
  Device::TextMate &m = d.text[s.text_mate];
  int rc = QB_OK;
  if (!sum.valid) {
    m.invalid = 1;
  } else {
    m.reads += sum.n_reads;
    if (s.text_last) m.tail_at_end = sum.tail_len;
    if (sum.n_reads) {
      qb::BatchView v{s.d_seq, s.d_qual, s.d_off, s.d_len, sum.n_reads, sum.n_bytes, sum.max_len, s.d_tiles,
                      sum.min_len == sum.max_len ? sum.min_len : 0u, 0u, sum.min_len};
      if (sum.max_len > ctx->cfg.len_cap)
        rc = fail(ctx, QB_ERR_CAPACITY, "a read of %u bp is longer than len_cap %u", sum.max_len, ctx->cfg.len_cap);
      else
        rc = launch_batch(ctx, d, v, s.text_mate, s.stream);
    }
  }
  cudaEventRecord(s.done, s.stream);
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    s.state = Slot::PENDING;
    s.seq = ++ctx->submit_seq;
    for (auto it = m.deferred.begin(); it != m.deferred.end(); ++it)
      if (*it == si) {
        m.deferred.erase(it);
        break;
      }
  }
  ctx->cv_slot.notify_all();
  return rc;
}

int flush_all_deferred(qb_ctx *ctx) {
  if (ctx->dev.empty() || ctx->dev[0].text.empty()) return QB_OK;
  Device &d = ctx->dev[0];
  for (size_t si = 0; si < d.slots.size(); si++) {
    bool def;
    {
      std::lock_guard<std::mutex> lk(ctx->mu);
      def = d.slots[si].state == Slot::DEFERRED || d.slots[si].state == Slot::FLUSHING;
    }
    if (def) {
      const int rc = flush_deferred(ctx, (int)si);
      if (rc) return rc;
    }
  }
  return QB_OK;
}
}  // namespace

extern "C" int qb_text_acquire(qb_ctx *ctx, qb_text *out) {
  if (!ctx || !out) return QB_ERR_ARG;
  int rc = text_setup(ctx);
  if (rc) return rc;
  int di, si;
  if ((rc = take_slot(ctx, &di, &si))) return rc;
  Slot &s = ctx->dev[di].slots[si];
  out->text = s.h_text;
  out->cap_bytes = text_cap(ctx);
  out->device_index = di;
  out->slot = si;
  return QB_OK;
}

namespace {
constexpr uint32_t kMaxBgzfBlocks = 1u << 16;  // per chunk

int bgzf_setup(qb_ctx *ctx) {  // once per context: per-slot buffers for compressed bytes and block tables
  std::lock_guard<std::mutex> lk(ctx->text_mu);
  Device &d = ctx->dev[0];
  if (d.slots[0].d_comp) return QB_OK;
  QB_CUDA(ctx, cudaSetDevice(d.id));
  for (Slot &s : d.slots) {
    QB_CUDA(ctx, cudaHostAlloc(&s.h_blk, sizeof(qb::BgzfBlock) * kMaxBgzfBlocks, cudaHostAllocDefault));
    QB_CUDA(ctx, cudaMalloc(&s.d_blk, sizeof(qb::BgzfBlock) * kMaxBgzfBlocks));
    QB_CUDA(ctx, cudaMalloc(&s.d_blk_status, 4 * kMaxBgzfBlocks + 4));
    s.d_bad = s.d_blk_status + kMaxBgzfBlocks;
  }
  for (size_t i = d.slots.size(); i-- > 0;) {  // (slot 0 last: its d_comp is the "set up" mark)
    QB_CUDA(ctx, cudaMalloc(&d.slots[i].d_comp, (size_t)text_cap(ctx) + 64));
    QB_CUDA(ctx, cudaMemset(d.slots[i].d_comp + text_cap(ctx), 0, 64));
  }
  return QB_OK;
}

// Walks the BGZF blocks at buf[0 .. n) (header: klib/bgzf.c:63-71; trailer CRC32 | ISIZE: 261-266): whole blocks only,
// as many as fit text_cap bytes of text / max_blocks.  Returns false for anything that is not a BGZF block header.
bool bgzf_walk(const uint8_t *buf, uint64_t n, uint64_t text_cap_bytes, uint32_t max_blocks, qb::BgzfBlock *out, uint32_t *n_blocks,
               uint64_t *n_whole, uint64_t *n_text) {
  uint64_t pos = 0, text = 0;
  uint32_t nb = 0;
  while (pos + 18 <= n && nb < max_blocks) {
    const uint8_t *h = buf + pos;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || h[3] != 4) return false;  // gzip member with FEXTRA only
    const uint32_t xlen = h[10] | (uint32_t)h[11] << 8;
    if (pos + 12 + xlen > n) break;
    int bsize = -1;
    for (uint32_t q = 12; q + 4 <= 12 + xlen;) {
      const uint32_t slen = h[q + 2] | (uint32_t)h[q + 3] << 8;
      if (h[q] == 'B' && h[q + 1] == 'C' && slen == 2 && q + 6 <= 12 + xlen) bsize = h[q + 4] | (int)h[q + 5] << 8;
      q += 4 + slen;
    }
    if (bsize < 0) return false;
    const uint32_t total = (uint32_t)bsize + 1u;
    if (total < 12 + xlen + 8) return false;
    if (pos + total > n) break;
    const uint8_t *tr = h + total - 8;
    const uint32_t crc = tr[0] | (uint32_t)tr[1] << 8 | (uint32_t)tr[2] << 16 | (uint32_t)tr[3] << 24;
    const uint32_t isize = tr[4] | (uint32_t)tr[5] << 8 | (uint32_t)tr[6] << 16 | (uint32_t)tr[7] << 24;
    if (isize > 65536u) return false;
    if (text + isize > text_cap_bytes) break;
    if (out) out[nb] = qb::BgzfBlock{(uint32_t)(pos + 12 + xlen), total - 12 - xlen - 8, (uint32_t)text, isize, crc};
    nb++;
    text += isize;
    pos += total;
  }
  *n_blocks = nb;
  *n_whole = pos;
  *n_text = text;
  return true;
}

}  // namespace

// src: the host bytes (the slot's own pinned buffer, or caller-owned memory for the *_from entry points)
// (inside text_submit_common: a CUDA failure hands the held slot back, so that no other thread waits for it forever)
#define QB_CUDA_SLOT(ctx, slot, call)                                                            \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      release_slot(ctx, 0, slot);                                                               \
      return fail(ctx, QB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    }                                                                                           \
  } while (0)

static int text_submit_common(qb_ctx *ctx, const qb_text *t, int mate, uint64_t n_bytes, int last, bool bgzf, const uint8_t *src = nullptr) {
  int rc = check_mate(ctx, mate);
  if (rc) return rc;
  if (!t || t->device_index != 0 || t->slot < 0 || t->slot >= ctx->cfg.ring_depth || ctx->dev[0].text.empty())
    return fail(ctx, QB_ERR_ARG, "bad text handle");
  Device &d = ctx->dev[0];
  Slot &s = d.slots[t->slot];
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (s.state != Slot::HELD || t->text != s.h_text) return fail(ctx, QB_ERR_ARG, "qb_text_submit: the slot is not held by a qb_text_acquire()");
  }
  if (!src) src = s.h_text;
  if (n_bytes > text_cap(ctx)) {
    release_slot(ctx, 0, t->slot);
    return fail(ctx, QB_ERR_CAPACITY, "text chunk of %llu bytes exceeds the slot (%u)", (unsigned long long)n_bytes, text_cap(ctx));
  }
  Device::TextMate &m = d.text[mate];
  QB_CUDA_SLOT(ctx, t->slot, cudaSetDevice(d.id));
  ctx->result_valid = false;
  uint64_t n_text = n_bytes;
  if (bgzf) {
    if ((rc = bgzf_setup(ctx))) {
      release_slot(ctx, 0, t->slot);
      return rc;
    }
    uint32_t nb = 0;
    uint64_t whole = 0;
    const bool ok = bgzf_walk(src, n_bytes, text_cap(ctx), kMaxBgzfBlocks, s.h_blk, &nb, &whole, &n_text);
    if (!ok || whole != n_bytes) {
      release_slot(ctx, 0, t->slot);
      if (!ok) {
        m.invalid = 1;
        return fail(ctx, QB_ERR_TEXT, "mate %d: not a sequence of BGZF blocks", mate);
      }
      return fail(ctx, QB_ERR_ARG, "qb_bgzf_submit: %llu of %llu bytes are whole blocks that fit the slot (see qb_bgzf_fit)",
                  (unsigned long long)whole, (unsigned long long)n_bytes);
    }
    if (n_bytes) QB_CUDA_SLOT(ctx, t->slot, cudaMemcpyAsync(s.d_comp, src, n_bytes, cudaMemcpyHostToDevice, s.stream));
    if (nb) QB_CUDA_SLOT(ctx, t->slot, cudaMemcpyAsync(s.d_blk, s.h_blk, sizeof(qb::BgzfBlock) * nb, cudaMemcpyHostToDevice, s.stream));
    cudaError_t e = qb::launch_inflate_bgzf(s.d_comp, s.d_blk, nb, s.d_text, s.d_bad, s.d_blk_status, s.stream);
    if (e == cudaSuccess && m.last_framed) e = cudaStreamWaitEvent(s.stream, m.last_framed, 0);
    if (e == cudaSuccess) e = qb::launch_inflate_merge(s.d_bad, m.d_state, s.stream);
    if (e != cudaSuccess) {
      release_slot(ctx, 0, t->slot);
      return fail(ctx, QB_ERR_CUDA, "inflate launch failed: %s", cudaGetErrorString(e));
    }
    ctx->launches += nb ? 1 : 0;
  } else {
    if (n_bytes) QB_CUDA_SLOT(ctx, t->slot, cudaMemcpyAsync(s.d_text, src, n_bytes, cudaMemcpyHostToDevice, s.stream));
    if (m.last_framed) QB_CUDA_SLOT(ctx, t->slot, cudaStreamWaitEvent(s.stream, m.last_framed, 0));  // the carry comes from the chunk in front
  }
  ctx->h2d_bytes += n_bytes;
  const cudaError_t e = qb::launch_text_frame(s.d_text, (uint32_t)n_text, m.d_carry[m.cur], m.d_carry[m.cur ^ 1], text_carry_cap(ctx), m.d_state,
                                              s.d_scratch, text_nl_cap(ctx), ctx->cfg.batch_reads, s.d_seq, s.d_qual, s.d_off, s.d_len,
                                              (uint32_t)ctx->cfg.batch_bytes, s.d_sum, s.stream);
  if (e != cudaSuccess) {
    release_slot(ctx, 0, t->slot);
    return fail(ctx, QB_ERR_CUDA, "framing launch failed: %s", cudaGetErrorString(e));
  }
  QB_CUDA_SLOT(ctx, t->slot, cudaMemcpyAsync(s.h_sum, s.d_sum, sizeof(qb::TextSummary), cudaMemcpyDeviceToHost, s.stream));
  QB_CUDA_SLOT(ctx, t->slot, cudaEventRecord(s.framed, s.stream));
  m.cur ^= 1;
  m.last_framed = s.framed;
  s.text_mate = mate;
  s.text_last = last;
  // Up to `lag` chunks of a mate stay deferred: the host queues the copy and the inflate / framing kernels of the next
  // chunks while the older ones are still being inflated (the inflate kernel needs thousands of blocks in flight), and
  // launches a chunk's statistics kernel once `lag` newer chunks are queued.  The ring bounds it: every mate may hold
  // lag deferred slots plus the one it fills.
  const size_t lag = (size_t)std::max(1, (ctx->cfg.ring_depth - 2) / ctx->cfg.n_mates);
  std::vector<int> flush;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    s.state = Slot::DEFERRED;
    m.deferred.push_back(t->slot);
    if (last) flush.assign(m.deferred.begin(), m.deferred.end());
    else if (m.deferred.size() > lag) flush.push_back(m.deferred.front());
  }
  for (int si : flush)
    if ((rc = flush_deferred(ctx, si))) return rc;
  return QB_OK;
}

#undef QB_CUDA_SLOT

extern "C" int qb_text_submit(qb_ctx *ctx, const qb_text *t, int mate, uint64_t n_bytes, int last) {
  return text_submit_common(ctx, t, mate, n_bytes, last, false);
}

extern "C" int qb_bgzf_submit(qb_ctx *ctx, const qb_text *t, int mate, uint64_t n_bytes, int last) {
  return text_submit_common(ctx, t, mate, n_bytes, last, true);
}

// Same from caller-owned host memory (pinned memory copies asynchronously; it must stay valid until the next
// qb_sync / qb_finish): takes the next free slot itself.
extern "C" int qb_bgzf_submit_from(qb_ctx *ctx, int mate, const uint8_t *blocks, uint64_t n_bytes, int last) {
  if (!ctx || (!blocks && n_bytes)) return QB_ERR_ARG;
  qb_text t;
  int rc = qb_text_acquire(ctx, &t);
  if (rc) return rc;
  return text_submit_common(ctx, &t, mate, n_bytes, last, true, blocks ? blocks : t.text);
}

extern "C" int qb_text_submit_from(qb_ctx *ctx, int mate, const uint8_t *text, uint64_t n_bytes, int last) {
  if (!ctx || (!text && n_bytes)) return QB_ERR_ARG;
  qb_text t;
  int rc = qb_text_acquire(ctx, &t);
  if (rc) return rc;
  return text_submit_common(ctx, &t, mate, n_bytes, last, false, text ? text : t.text);
}

extern "C" int qb_text_capacity(qb_ctx *ctx, uint64_t *cap_bytes) {
  if (!ctx || !cap_bytes) return QB_ERR_ARG;
  *cap_bytes = text_cap(ctx);
  return QB_OK;
}

extern "C" int qb_bgzf_fit(const uint8_t *buf, uint64_t n_bytes, uint64_t text_cap_bytes, uint64_t *n_whole, uint64_t *n_text) {
  uint32_t nb = 0;
  uint64_t whole = 0, text = 0;
  if (!buf && n_bytes) return QB_ERR_ARG;
  if (!bgzf_walk(buf, n_bytes, text_cap_bytes, kMaxBgzfBlocks, nullptr, &nb, &whole, &text)) return QB_ERR_TEXT;
  if (n_whole) *n_whole = whole;
  if (n_text) *n_text = text;
  return QB_OK;
}

// Diagnostics: the inflate kernel alone on device-resident blocks, timed with CUDA events (ms per launch).
extern "C" int qb_bgzf_inflate_bench(qb_ctx *ctx, const uint8_t *comp, uint64_t n_bytes, int iters, float *ms_per_launch,
                                     uint64_t *n_text_out, uint32_t *n_blocks_out) {
  if (!ctx || !comp || iters < 1 || !ms_per_launch) return QB_ERR_ARG;
  Device &d = ctx->dev[0];
  QB_CUDA(ctx, cudaSetDevice(d.id));
  std::vector<qb::BgzfBlock> blocks;
  uint64_t pos = 0, n_text = 0;
  while (pos < n_bytes) {  // every block of the buffer, however many
    std::vector<qb::BgzfBlock> part(kMaxBgzfBlocks);
    uint32_t nb = 0;
    uint64_t whole = 0, text = 0;
    if (!bgzf_walk(comp + pos, n_bytes - pos, 0xF0000000ull - n_text, kMaxBgzfBlocks, part.data(), &nb, &whole, &text) || whole == 0)
      return fail(ctx, QB_ERR_TEXT, "not a sequence of whole BGZF blocks");
    for (uint32_t i = 0; i < nb; i++) {
      part[i].in_off += (uint32_t)pos;
      part[i].out_off += (uint32_t)n_text;
      blocks.push_back(part[i]);
    }
    pos += whole;
    n_text += text;
  }
  const uint32_t nb = (uint32_t)blocks.size();
  uint8_t *d_comp = nullptr, *d_text = nullptr;
  qb::BgzfBlock *d_blk = nullptr;
  uint32_t *d_status = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = cudaMalloc(&d_comp, n_bytes + 64);
  if (e == cudaSuccess) e = cudaMalloc(&d_text, n_text + 64);
  if (e == cudaSuccess) e = cudaMalloc(&d_blk, sizeof(qb::BgzfBlock) * (nb + 1));
  if (e == cudaSuccess) e = cudaMalloc(&d_status, 4 * (size_t)nb + 8);
  if (e == cudaSuccess) e = cudaMemset(d_comp + n_bytes, 0, 64);
  if (e == cudaSuccess) e = cudaMemcpy(d_comp, comp, n_bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_blk, blocks.data(), sizeof(qb::BgzfBlock) * nb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  for (int w = 0; w < 2 && e == cudaSuccess; w++) e = qb::launch_inflate_bgzf(d_comp, d_blk, nb, d_text, d_status + nb, d_status, d.main_stream);
  if (e == cudaSuccess) e = cudaEventRecord(e0, d.main_stream);
  for (int i = 0; i < iters && e == cudaSuccess; i++) e = qb::launch_inflate_bgzf(d_comp, d_blk, nb, d_text, d_status + nb, d_status, d.main_stream);
  if (e == cudaSuccess) e = cudaEventRecord(e1, d.main_stream);
  if (e == cudaSuccess) e = cudaEventSynchronize(e1);
  float ms = 0;
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
  uint32_t bad = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&bad, d_status + nb, 4, cudaMemcpyDeviceToHost);
  cudaFree(d_comp), cudaFree(d_text), cudaFree(d_blk), cudaFree(d_status);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (e != cudaSuccess) return fail(ctx, QB_ERR_CUDA, "qb_bgzf_inflate_bench: %s", cudaGetErrorString(e));
  if (bad) return fail(ctx, QB_ERR_TEXT, "a block does not inflate");
  ctx->launches += (uint64_t)iters + 2;
  *ms_per_launch = ms / (float)iters;
  if (n_text_out) *n_text_out = n_text;
  if (n_blocks_out) *n_blocks_out = nb;
  return QB_OK;
}

// Inflates whole BGZF blocks on the device and copies the text back: the decoder on its own (tests, other callers).
extern "C" int qb_bgzf_inflate(qb_ctx *ctx, const uint8_t *comp, uint64_t n_bytes, uint8_t *text_out, uint64_t text_cap_bytes,
                               uint64_t *n_text_out) {
  if (!ctx || (!comp && n_bytes) || !n_text_out) return QB_ERR_ARG;
  Device &d = ctx->dev[0];
  QB_CUDA(ctx, cudaSetDevice(d.id));
  std::vector<qb::BgzfBlock> blocks(kMaxBgzfBlocks);
  uint32_t nb = 0;
  uint64_t whole = 0, n_text = 0;
  if (!bgzf_walk(comp, n_bytes, text_cap_bytes, kMaxBgzfBlocks, blocks.data(), &nb, &whole, &n_text) || whole != n_bytes)
    return fail(ctx, QB_ERR_TEXT, "not a sequence of whole BGZF blocks that fit %llu bytes of text", (unsigned long long)text_cap_bytes);
  uint8_t *d_comp = nullptr, *d_text = nullptr;
  qb::BgzfBlock *d_blk = nullptr;
  uint32_t *d_status = nullptr;
  cudaError_t e = cudaMalloc(&d_comp, n_bytes + 64);
  if (e == cudaSuccess) e = cudaMalloc(&d_text, n_text + 64);
  if (e == cudaSuccess) e = cudaMalloc(&d_blk, sizeof(qb::BgzfBlock) * (nb + 1));
  if (e == cudaSuccess) e = cudaMalloc(&d_status, 4 * (size_t)nb + 8);
  if (e == cudaSuccess) e = cudaMemset(d_comp + n_bytes, 0, 64);
  if (e == cudaSuccess && n_bytes) e = cudaMemcpy(d_comp, comp, n_bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nb) e = cudaMemcpy(d_blk, blocks.data(), sizeof(qb::BgzfBlock) * nb, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = qb::launch_inflate_bgzf(d_comp, d_blk, nb, d_text, d_status + nb, d_status, d.main_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(d.main_stream);
  std::vector<uint32_t> status(nb + 1, 0);
  if (e == cudaSuccess) e = cudaMemcpy(status.data(), d_status, 4 * (size_t)(nb + 1), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && n_text) e = cudaMemcpy(text_out, d_text, n_text, cudaMemcpyDeviceToHost);
  cudaFree(d_comp), cudaFree(d_text), cudaFree(d_blk), cudaFree(d_status);
  if (e != cudaSuccess) return fail(ctx, QB_ERR_CUDA, "qb_bgzf_inflate: %s", cudaGetErrorString(e));
  ctx->launches += nb ? 1 : 0;
  *n_text_out = n_text;
  if (status[nb])
    for (uint32_t i = 0; i < nb; i++)
      if (status[i]) return fail(ctx, QB_ERR_TEXT, "BGZF block %u does not inflate (check %u)", i, status[i]);
  return QB_OK;
}

extern "C" int qb_text_status(qb_ctx *ctx, int mate, uint64_t *n_reads, uint64_t *tail_bytes) {
  int rc = check_mate(ctx, mate);
  if (rc) return rc;
  if ((rc = flush_all_deferred(ctx))) return rc;
  if (ctx->dev[0].text.empty()) return fail(ctx, QB_ERR_ARG, "the text path was not used");
  const Device::TextMate &m = ctx->dev[0].text[mate];
  if (n_reads) *n_reads = m.reads;
  if (tail_bytes) *tail_bytes = m.tail_at_end;
  return m.invalid ? fail(ctx, QB_ERR_TEXT, "mate %d: the text is not canonical 4-line FASTQ; count it through the host reader", mate)
                   : QB_OK;
}

int qb_sync(qb_ctx *ctx) {
  if (!ctx) return QB_ERR_ARG;
  {
    const int frc = flush_all_deferred(ctx);
    if (frc) return frc;
  }
  uint64_t seen;  // submissions up to this one are covered by the stream synchronisation below
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    seen = ctx->submit_seq;
  }
  for (Device &d : ctx->dev) {
    QB_CUDA(ctx, cudaSetDevice(d.id));
    for (Slot &s : d.slots) QB_CUDA(ctx, cudaStreamSynchronize(s.stream));
    QB_CUDA(ctx, cudaStreamSynchronize(d.main_stream));
  }
  {  // everything submitted so far is done; slots held by other threads stay theirs
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (Device &d : ctx->dev)
      for (Slot &s : d.slots)
        if (s.state == Slot::PENDING && s.seq <= seen) s.state = Slot::FREE;
  }
  ctx->cv_slot.notify_all();
  return QB_OK;
}

int qb_reset(qb_ctx *ctx, int mate) {
  int rc = check_mate(ctx, mate);
  if (rc) return rc;
  rc = qb_sync(ctx);
  if (rc) return rc;
  std::unique_lock<std::shared_mutex> lk(ctx->acc_mu);
  for (Device &d : ctx->dev) {
    QB_CUDA(ctx, cudaSetDevice(d.id));
    QB_CUDA(ctx, cudaMemset(d.acc[mate], 0, ctx->acc_u64 * 8));
    if (ctx->extras.load()) QB_CUDA(ctx, cudaMemset(d.ext[mate], 0, ((size_t)ctx->cfg.len_cap + qb::kExtrasMeanBins) * 8));
    if (!d.text.empty()) {  // a new stream starts: no carry, no error from the last one
      QB_CUDA(ctx, cudaMemset(d.text[mate].d_state, 0, sizeof(qb::TextState)));
      d.text[mate].invalid = 0;
      d.text[mate].reads = d.text[mate].tail_at_end = 0;
      d.text[mate].last_framed = nullptr;
    }
  }
  ctx->result_valid = false;
  return QB_OK;
}

int qb_nccl_unique_id(uint8_t id_out[QB_NCCL_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) == QB_NCCL_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  NcclApi *nc = nccl_api();
  if (!nc || nc->GetUniqueId(&id) != ncclSuccess) return QB_ERR_NCCL;
  memcpy(id_out, &id, sizeof id);
  return QB_OK;
}

int qb_comm_init_rank(qb_ctx *ctx, int n_ranks, int rank, const uint8_t id_in[QB_NCCL_ID_BYTES]) {
  if (!ctx || !id_in || n_ranks < 1 || rank < 0 || rank >= n_ranks) return QB_ERR_ARG;
  if (ctx->dev.size() != 1) return fail(ctx, QB_ERR_ARG, "qb_comm_init_rank needs a one-device context per rank");
  ncclUniqueId id;
  memcpy(&id, id_in, sizeof id);
  NcclApi *nc = nccl_api();
  if (!nc) return fail(ctx, QB_ERR_NCCL, "libnccl.so.2 could not be loaded");
  QB_CUDA(ctx, cudaSetDevice(ctx->dev[0].id));
  QB_NCCL(ctx, nc->CommInitRank(&ctx->rank_comm, n_ranks, id, rank));
  ctx->n_ranks = n_ranks;
  ctx->rank = rank;
  return QB_OK;
}

// Synchronises and brings the summed accumulators of EVERY mate to h_result: one grouped reduce over the devices of
// this process and / or one reduce over the ranks, one device-to-host copy.  The result is kept until new work is
// submitted, so the qb_finish() of the second mate costs a memcpy.
static int reduce_all(qb_ctx *ctx) {
  int rc = qb_sync(ctx);
  if (rc) return rc;
  NcclApi *nc = nccl_api();
  Device &root = ctx->dev[0];
  if (ctx->rank_comm && ctx->n_ranks > 1 && ctx->cur_cap < ctx->cfg.len_cap) {
    // the ranks may have grown their accumulators differently: agree on the largest row count first.  (Not needed
    // once the rows cover len_cap -- every rank is created with the same len_cap --, e.g. for len_cap <= 512.)
    QB_CUDA(ctx, cudaSetDevice(root.id));
    unsigned long long cap = ctx->cur_cap;
    QB_CUDA(ctx, cudaMemcpyAsync(root.d_scalar, &cap, 8, cudaMemcpyHostToDevice, root.main_stream));
    QB_NCCL(ctx, nc->AllReduce(root.d_scalar, root.d_scalar, 1, ncclUint64, ncclMax, ctx->rank_comm, root.main_stream));
    QB_CUDA(ctx, cudaMemcpyAsync(&cap, root.d_scalar, 8, cudaMemcpyDeviceToHost, root.main_stream));
    QB_CUDA(ctx, cudaStreamSynchronize(root.main_stream));
    if ((rc = ensure_cap(ctx, (uint32_t)cap))) return rc;
  }
  std::unique_lock<std::shared_mutex> lk(ctx->acc_mu);
  const size_t n = ctx->acc_u64 * (size_t)ctx->cfg.n_mates;
  const unsigned long long *src = root.acc_all;
  if (ctx->dev.size() > 1) {
    // one grouped ncclReduce of every mate's [rows | counters] from every device of this process to device 0
    QB_NCCL(ctx, nc->GroupStart());
    for (Device &d : ctx->dev)
      QB_NCCL(ctx, nc->Reduce(d.acc_all, d.reduce_buf, n, ncclUint64, ncclSum, 0, d.comm, d.main_stream));
    QB_NCCL(ctx, nc->GroupEnd());
    for (Device &d : ctx->dev) {
      QB_CUDA(ctx, cudaSetDevice(d.id));
      QB_CUDA(ctx, cudaStreamSynchronize(d.main_stream));
    }
    src = root.reduce_buf;
  }
  QB_CUDA(ctx, cudaSetDevice(root.id));
  if (ctx->rank_comm && ctx->n_ranks > 1) {
    // one ncclReduce across the ranks (one process per GPU) to rank 0, over NVLink/NVSwitch
    QB_NCCL(ctx, nc->Reduce(src, root.reduce_buf, n, ncclUint64, ncclSum, 0, ctx->rank_comm, root.main_stream));
    if (ctx->rank == 0) src = root.reduce_buf;
  }
  ctx->d_result = src;
  QB_CUDA(ctx, cudaMemcpyAsync(ctx->h_result, src, n * 8, cudaMemcpyDeviceToHost, root.main_stream));
  QB_CUDA(ctx, cudaStreamSynchronize(root.main_stream));
  ctx->result_valid = true;
  return QB_OK;
}

int qb_finish(qb_ctx *ctx, int mate, uint64_t *rows_out, uint64_t rows_cap, uint64_t *max_length,
              uint64_t *n_reads) {
  int rc = check_mate(ctx, mate);
  if (rc) return rc;
  if (!ctx->result_valid && (rc = reduce_all(ctx))) return rc;
  if (!ctx->dev[0].text.empty() && ctx->dev[0].text[mate].invalid)
    return fail(ctx, QB_ERR_TEXT, "mate %d: the text is not canonical 4-line FASTQ; count it through the host reader", mate);

  const uint32_t cap = ctx->cur_cap;
  unsigned long long *rows = ctx->h_result + (size_t)mate * ctx->acc_u64;
  const unsigned long long *cnt = rows + (size_t)cap * qb::kRow;
  if (cnt[qb::kCntError])
    return fail(ctx, QB_ERR_CAPACITY, "%llu tile(s)/read(s) exceeded len_cap (%u), the max_len given at submit, or the batch layout rules; result invalid",
                cnt[qb::kCntError], ctx->cfg.len_cap);
  uint64_t ml = 0;
  for (uint32_t i = 0; i < cap; i++)
    if (rows[(size_t)i * qb::kRow + qb::kColLength]) ml = i + 1;  // quack.c:194-198, derived
  if (!ctx->cfg.adapters_enabled && ml > 10) {
    // no -a: the reference counts kmer_count[10] once per read longer than 10 (quack.c:210-217)
    unsigned long long k = 0;
    for (uint64_t i = 10; i < ml; i++) k += rows[(size_t)i * qb::kRow + qb::kColLength];
    rows[(size_t)10 * qb::kRow + qb::kColKmer] = k;
  }
  if (max_length) *max_length = ml;
  if (n_reads) *n_reads = cnt[qb::kCntReads];
  if (rows_out) {
    if (ml > rows_cap) return fail(ctx, QB_ERR_CAPACITY, "rows_out holds %llu rows, need %llu",
                                   (unsigned long long)rows_cap, (unsigned long long)ml);
    memcpy(rows_out, rows, (size_t)ml * qb::kRow * 8);
  }
  return QB_OK;
}

// transform() of the reference (quack.c:230-293) on the device: binning of reads longer than 3000 bp, running sum of
// kmer_count, percentages; only the transformed rows cross the link.  Same arrays as host/render.c:qr_transform().
int qb_finish_transformed(qb_ctx *ctx, int mate, uint64_t *rows_out, uint64_t rows_cap, uint64_t *max_length,
                          uint64_t *n_reads, uint64_t *original_max_length) {
  uint64_t ml = 0, nr = 0;
  int rc = qb_finish(ctx, mate, nullptr, 0, &ml, &nr);  // (reduce, error checks, longest read)
  if (rc) return rc;
  const uint32_t n_rows = qb::transformed_rows((uint32_t)ml);
  if (original_max_length) *original_max_length = ml;
  if (max_length) *max_length = n_rows;
  if (n_reads) *n_reads = nr;
  if (!rows_out) return QB_OK;
  if (n_rows > rows_cap) return fail(ctx, QB_ERR_CAPACITY, "rows_out holds %llu rows, need %u", (unsigned long long)rows_cap, n_rows);
  if (n_rows == 0) return QB_OK;
  if (nr == 0) return fail(ctx, QB_ERR_ARG, "no reads: nothing to transform");
  Device &root = ctx->dev[0];
  QB_CUDA(ctx, cudaSetDevice(root.id));
  std::shared_lock<std::shared_mutex> lk(ctx->acc_mu);
  if (!ctx->result_valid.load() || !ctx->d_result) return fail(ctx, QB_ERR_ARG, "the result changed under qb_finish_transformed");
  unsigned long long *d_out = nullptr;
  QB_CUDA(ctx, cudaMalloc(&d_out, ((size_t)n_rows * qb::kRow + 1) * 8));
  const unsigned long long *src = ctx->d_result + (size_t)mate * ctx->acc_u64;
  cudaError_t e = qb::launch_transform(src, (uint32_t)ml, nr, !ctx->cfg.adapters_enabled, d_out, d_out + (size_t)n_rows * qb::kRow, root.main_stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(rows_out, d_out, (size_t)n_rows * qb::kRow * 8, cudaMemcpyDeviceToHost, root.main_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(root.main_stream);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(ctx, QB_ERR_CUDA, "device transform failed: %s", cudaGetErrorString(e));
  ctx->launches += 2;
  return QB_OK;
}

// ---- opt-in side outputs (no reference oracle: SURVEY.md section 0.1) ----
int qb_extras_enable(qb_ctx *ctx) {
  if (!ctx) return QB_ERR_ARG;
  if (ctx->extras.load()) return QB_OK;
  if (ctx->submit_seq != 0 || ctx->launches.load() != 0)
    return fail(ctx, QB_ERR_ARG, "qb_extras_enable must be called before the first batch is submitted");
  const size_t n = (size_t)ctx->cfg.len_cap + qb::kExtrasMeanBins;
  for (Device &d : ctx->dev) {
    QB_CUDA(ctx, cudaSetDevice(d.id));
    d.ext.assign(ctx->cfg.n_mates, nullptr);
    for (int m = 0; m < ctx->cfg.n_mates; m++) {
      QB_CUDA(ctx, cudaMalloc(&d.ext[m], n * 8));
      QB_CUDA(ctx, cudaMemset(d.ext[m], 0, n * 8));
    }
    QB_CUDA(ctx, cudaMalloc(&d.ext_reduce, n * 8));
  }
  ctx->extras.store(true);
  return QB_OK;
}

int qb_extras_finish(qb_ctx *ctx, int mate, uint64_t *n_count, uint64_t *qual_sum, uint64_t rows_cap, uint64_t *mean_hist) {
  int rc = check_mate(ctx, mate);
  if (rc) return rc;
  if (!ctx->extras.load()) return fail(ctx, QB_ERR_ARG, "qb_extras_enable was not called");
  // the heatmap rows (for the per-position quality sums) and max_length: the ordinary result
  uint64_t ml = 0, nr = 0;
  if ((rc = qb_finish(ctx, mate, nullptr, 0, &ml, &nr))) return rc;
  if (ml > rows_cap) return fail(ctx, QB_ERR_CAPACITY, "extras arrays hold %llu positions, need %llu", (unsigned long long)rows_cap, (unsigned long long)ml);
  const size_t n = (size_t)ctx->cfg.len_cap + qb::kExtrasMeanBins;
  std::vector<unsigned long long> sum(n, 0), part(n);
  for (Device &d : ctx->dev) {  // the devices of this process: summed on the host (n is small)
    QB_CUDA(ctx, cudaSetDevice(d.id));
    QB_CUDA(ctx, cudaMemcpy(part.data(), d.ext[mate], n * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) sum[i] += part[i];
  }
  if (ctx->rank_comm && ctx->n_ranks > 1) {  // one process per GPU: one more integer reduce to rank 0
    NcclApi *nc = nccl_api();
    Device &root = ctx->dev[0];
    QB_CUDA(ctx, cudaSetDevice(root.id));
    QB_CUDA(ctx, cudaMemcpyAsync(root.ext[mate], sum.data(), n * 8, cudaMemcpyHostToDevice, root.main_stream));
    QB_NCCL(ctx, nc->Reduce(root.ext[mate], root.ext_reduce, n, ncclUint64, ncclSum, 0, ctx->rank_comm, root.main_stream));
    if (ctx->rank == 0) QB_CUDA(ctx, cudaMemcpyAsync(sum.data(), root.ext_reduce, n * 8, cudaMemcpyDeviceToHost, root.main_stream));
    QB_CUDA(ctx, cudaStreamSynchronize(root.main_stream));
  }
  if (n_count) memcpy(n_count, sum.data(), (size_t)ml * 8);
  if (mean_hist) memcpy(mean_hist, sum.data() + ctx->cfg.len_cap, qb::kExtrasMeanBins * 8);
  if (qual_sum) {  // sum_s s * scores[p][s] over the raw heatmap rows (quack.c:203-204 counts)
    const unsigned long long *rows = ctx->h_result + (size_t)mate * ctx->acc_u64;
    for (uint64_t p = 0; p < ml; p++) {
      unsigned long long q = 0;
      for (uint32_t s = 1; s < 91; s++) q += (unsigned long long)s * rows[p * qb::kRow + s];
      qual_sum[p] = q;
    }
  }
  return QB_OK;
}

int qb_invalid_quality_count(qb_ctx *ctx, int mate, uint64_t *out) {
  int rc = check_mate(ctx, mate);
  if (rc) return rc;
  rc = qb_sync(ctx);
  if (rc) return rc;
  unsigned long long total = 0;
  for (Device &d : ctx->dev) {
    unsigned long long v = 0;
    QB_CUDA(ctx, cudaSetDevice(d.id));
    QB_CUDA(ctx, cudaMemcpy(&v, d.acc[mate] + (size_t)ctx->cur_cap * qb::kRow + qb::kCntInvalidQual, 8,
                            cudaMemcpyDeviceToHost));
    total += v;
  }
  *out = total;
  return QB_OK;
}

uint64_t qb_launch_count(const qb_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

int qb_kernel_counts(const qb_ctx *ctx, uint64_t *n_simple, uint64_t *n_fused) {
  if (!ctx) return QB_ERR_ARG;
  if (n_simple) *n_simple = ctx->launches_simple;
  if (n_fused) *n_fused = ctx->launches_fused;
  return QB_OK;
}

uint64_t qb_period_launch_count(const qb_ctx *ctx) { return ctx ? ctx->launches_period.load() : 0; }
uint64_t qb_flat_launch_count(const qb_ctx *ctx) { return ctx ? ctx->launches_flat.load() : 0; }
uint64_t qb_h2d_bytes(const qb_ctx *ctx) { return ctx ? ctx->h2d_bytes.load() : 0; }

int qb_profile_enable(qb_ctx *ctx, int max_launches) {
  if (!ctx || max_launches < 0) return QB_ERR_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  for (auto &r : ctx->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  ctx->prof.clear();
  ctx->prof.reserve(max_launches);  // launch_batch keeps pointers into the vector: never reallocate
  ctx->prof_cap = max_launches;
  return QB_OK;
}

int qb_profile_collect(qb_ctx *ctx, float *ms_out, uint64_t *bytes_out, int cap) {
  if (!ctx || cap < 0) return QB_ERR_ARG;
  int rc = qb_sync(ctx);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  int n = 0;
  for (auto &r : ctx->prof) {
    if (n < cap) {
      float ms = 0;
      cudaSetDevice(r.dev);
      if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) ms = -1.f;
      if (ms_out) ms_out[n] = ms;
      if (bytes_out) bytes_out[n] = r.bytes;
      n++;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  ctx->prof.clear();
  return n;
}

int qb_timer_start(qb_ctx *ctx, int device_index) {
  if (!ctx || device_index < 0 || device_index >= (int)ctx->dev.size()) return QB_ERR_ARG;
  Device &d = ctx->dev[device_index];
  QB_CUDA(ctx, cudaSetDevice(d.id));
  if (!d.t0) {
    QB_CUDA(ctx, cudaEventCreate(&d.t0));
    QB_CUDA(ctx, cudaEventCreate(&d.t1));
  }
  QB_CUDA(ctx, cudaEventRecord(d.t0, d.main_stream));
  return QB_OK;
}

int qb_timer_stop(qb_ctx *ctx, int device_index, float *ms) {
  if (!ctx || !ms || device_index < 0 || device_index >= (int)ctx->dev.size()) return QB_ERR_ARG;
  Device &d = ctx->dev[device_index];
  if (!d.t0) return fail(ctx, QB_ERR_ARG, "qb_timer_stop without qb_timer_start");
  QB_CUDA(ctx, cudaSetDevice(d.id));
  QB_CUDA(ctx, cudaEventRecord(d.t1, d.main_stream));
  QB_CUDA(ctx, cudaEventSynchronize(d.t1));
  QB_CUDA(ctx, cudaEventElapsedTime(ms, d.t0, d.t1));
  return QB_OK;
}

// ---------------------------------------------------------------------- device-resident batches

static int dbatch_alloc(qb_ctx *ctx, int di, uint32_t n_reads, uint64_t n_bytes, qb_dbatch **out) {
  if (di < 0 || di >= (int)ctx->dev.size()) return fail(ctx, QB_ERR_ARG, "device_index %d out of range", di);
  if (n_bytes > 0xFFFFFF00ull - 64) return fail(ctx, QB_ERR_CAPACITY, "a batch must stay below 4 GiB (u32 offsets)");
  Device &d = ctx->dev[di];
  QB_CUDA(ctx, cudaSetDevice(d.id));
  qb_dbatch *b = new qb_dbatch();
  memset(b, 0, sizeof *b);
  b->device_index = di;
  b->n_reads = n_reads;
  b->n_bytes = n_bytes;
  cudaError_t e;
  if ((e = cudaMalloc(&b->d_seq, pad_bytes(n_bytes))) != cudaSuccess ||
      (e = cudaMalloc(&b->d_qual, pad_bytes(n_bytes))) != cudaSuccess ||
      (e = cudaMalloc(&b->d_off, pad_reads(n_reads) * 4)) != cudaSuccess ||
      (e = cudaMalloc(&b->d_len, pad_reads(n_reads) * 4)) != cudaSuccess ||
      (e = cudaMalloc(&b->d_tiles, pad_reads(n_reads) * sizeof(uint2))) != cudaSuccess) {
    qb_dbatch_free(ctx, b);
    return fail(ctx, QB_ERR_NOMEM, "cudaMalloc for a device batch failed: %s", cudaGetErrorString(e));
  }
  QB_CUDA(ctx, cudaMemset(b->d_seq + (n_bytes & ~15ull), 0, pad_bytes(n_bytes) - (n_bytes & ~15ull)));
  QB_CUDA(ctx, cudaMemset(b->d_qual + (n_bytes & ~15ull), 0, pad_bytes(n_bytes) - (n_bytes & ~15ull)));
  QB_CUDA(ctx, cudaMemset(b->d_off, 0, pad_reads(n_reads) * 4));
  QB_CUDA(ctx, cudaMemset(b->d_len, 0, pad_reads(n_reads) * 4));
  *out = b;
  return QB_OK;
}

int qb_dbatch_upload(qb_ctx *ctx, int device_index, const uint8_t *seq, const uint8_t *qual, const uint32_t *offset,
                     const uint32_t *length, uint32_t n_reads, uint64_t n_bytes, uint32_t max_len, qb_dbatch **out) {
  if (!ctx || !out) return QB_ERR_ARG;
  qb_dbatch *b = nullptr;
  int rc = dbatch_alloc(ctx, device_index, n_reads, n_bytes, &b);
  if (rc) return rc;
  b->max_len = max_len ? max_len : ctx->cfg.len_cap;
  if (n_bytes) {
    QB_CUDA(ctx, cudaMemcpy(b->d_seq, seq, n_bytes, cudaMemcpyHostToDevice));
    QB_CUDA(ctx, cudaMemcpy(b->d_qual, qual, n_bytes, cudaMemcpyHostToDevice));
  }
  if (n_reads) {
    QB_CUDA(ctx, cudaMemcpy(b->d_off, offset, (size_t)n_reads * 4, cudaMemcpyHostToDevice));
    QB_CUDA(ctx, cudaMemcpy(b->d_len, length, (size_t)n_reads * 4, cudaMemcpyHostToDevice));
    uint32_t longest = 0;
    b->uniform_len = detect_shape(offset, length, n_reads, &b->first_offset, &b->contig_min_len, &longest);
    if (longest > b->max_len) b->contig_min_len = 0;  // (see submit_queue)
  }
  *out = b;
  return QB_OK;
}

int qb_dbatch_generate(qb_ctx *ctx, int device_index, uint64_t seed, int mate, uint64_t first_read, uint32_t n_reads,
                       uint32_t len_min, uint32_t len_max, double adapter_rate, qb_dbatch **out) {
  if (!ctx || !out || len_min == 0 || len_max < len_min) return QB_ERR_ARG;
  // lengths first (they fix the byte count), then generate in host chunks and upload
  uint64_t total = 0;
  uint32_t shortest = 0xFFFFFFFFu;
  for (uint32_t r = 0; r < n_reads; r++) {
    const uint32_t l = qb::gen_length(seed, first_read + r, len_min, len_max);
    total += l;
    shortest = l < shortest ? l : shortest;
  }
  qb_dbatch *b = nullptr;
  int rc = dbatch_alloc(ctx, device_index, n_reads, total, &b);
  if (rc) return rc;
  b->max_len = len_max;
  const uint32_t chunk = 1u << 20;
  uint8_t *hs = (uint8_t *)qb_host_alloc((size_t)chunk * len_max), *hq = (uint8_t *)qb_host_alloc((size_t)chunk * len_max);
  uint32_t *ho = (uint32_t *)qb_host_alloc((size_t)chunk * 4), *hl = (uint32_t *)qb_host_alloc((size_t)chunk * 4);
  if (!hs || !hq || !ho || !hl) {
    qb_host_free(hs), qb_host_free(hq), qb_host_free(ho), qb_host_free(hl);
    qb_dbatch_free(ctx, b);
    return fail(ctx, QB_ERR_NOMEM, "pinned staging allocation failed");
  }
  uint64_t base = 0;
  bool uniform = len_min == len_max, contiguous = true;  // the generator packs reads back to back
  for (uint32_t r0 = 0; r0 < n_reads && rc == QB_OK; r0 += chunk) {
    const uint32_t n = n_reads - r0 < chunk ? n_reads - r0 : chunk;
    uint64_t nb = 0;
    rc = qb_gen_reads(seed, mate, first_read + r0, n, len_min, len_max, adapter_rate, hs, hq, ho, hl, &nb);
    if (rc) break;
    for (uint32_t i = 0; i < n; i++) ho[i] += (uint32_t)base;
    uint32_t fo = 0, cm = 0, lg = 0;
    const uint32_t ul = detect_shape(ho, hl, n, &fo, &cm, &lg);
    if (uniform && (ul != len_min || fo != (uint32_t)base)) uniform = false;
    if (!cm || fo != (uint32_t)base) contiguous = false;  // (the generator packs reads back to back)
    if (cudaMemcpy(b->d_seq + base, hs, nb, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(b->d_qual + base, hq, nb, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(b->d_off + r0, ho, (size_t)n * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(b->d_len + r0, hl, (size_t)n * 4, cudaMemcpyHostToDevice) != cudaSuccess)
      rc = fail(ctx, QB_ERR_CUDA, "upload of generated reads failed: %s", cudaGetErrorString(cudaGetLastError()));
    base += nb;
  }
  qb_host_free(hs), qb_host_free(hq), qb_host_free(ho), qb_host_free(hl);
  if (rc) {
    qb_dbatch_free(ctx, b);
    return rc;
  }
  b->uniform_len = uniform && n_reads ? len_min : 0u;
  b->contig_min_len = contiguous && n_reads ? shortest : 0u;
  b->first_offset = 0;
  *out = b;
  return QB_OK;
}

int qb_dbatch_info(const qb_dbatch *b, uint32_t *n_reads, uint64_t *n_bytes) {
  if (!b) return QB_ERR_ARG;
  if (n_reads) *n_reads = b->n_reads;
  if (n_bytes) *n_bytes = b->n_bytes;
  return QB_OK;
}

void qb_dbatch_free(qb_ctx *ctx, qb_dbatch *b) {
  if (!b) return;
  if (ctx && b->device_index >= 0 && b->device_index < (int)ctx->dev.size()) cudaSetDevice(ctx->dev[b->device_index].id);
  cudaFree(b->d_seq);
  cudaFree(b->d_qual);
  cudaFree(b->d_off);
  cudaFree(b->d_len);
  cudaFree(b->d_tiles);
  delete b;
}

int qb_dbatch_run(qb_ctx *ctx, qb_dbatch *b, int mate) {
  int rc = check_mate(ctx, mate);
  if (rc) return rc;
  if (!b) return QB_ERR_ARG;
  Device &d = ctx->dev[b->device_index];
  QB_CUDA(ctx, cudaSetDevice(d.id));
  qb::BatchView v{b->d_seq, b->d_qual, b->d_off, b->d_len, b->n_reads, b->n_bytes, b->max_len, b->d_tiles,
                  b->uniform_len, b->first_offset, b->contig_min_len};
  return launch_batch(ctx, d, v, mate, d.main_stream);
}

int qb_dbatch_time(qb_ctx *ctx, qb_dbatch *b, int mate, int warmup, int iters, int flush_l2, float *ms_avg,
                   float *ms_min) {
  int rc = check_mate(ctx, mate);
  if (rc) return rc;
  if (!b || iters < 1) return QB_ERR_ARG;
  Device &d = ctx->dev[b->device_index];
  QB_CUDA(ctx, cudaSetDevice(d.id));
  if (flush_l2 && !d.l2_scratch) {
    d.l2_words = (256ull << 20) / 4;  // 256 MiB > 126 MB of L2
    QB_CUDA(ctx, cudaMalloc(&d.l2_scratch, d.l2_words * 4));
  }
  cudaEvent_t e0, e1;
  QB_CUDA(ctx, cudaEventCreate(&e0));
  QB_CUDA(ctx, cudaEventCreate(&e1));
  for (int i = 0; i < warmup; i++)
    if ((rc = qb_dbatch_run(ctx, b, mate))) return rc;
  QB_CUDA(ctx, cudaStreamSynchronize(d.main_stream));
  double sum = 0;
  float best = 1e30f;
  for (int i = 0; i < iters; i++) {
    if (flush_l2) QB_CUDA(ctx, qb::launch_l2_flush(d.l2_scratch, d.l2_words, d.main_stream));
    QB_CUDA(ctx, cudaEventRecord(e0, d.main_stream));
    if ((rc = qb_dbatch_run(ctx, b, mate))) return rc;
    QB_CUDA(ctx, cudaEventRecord(e1, d.main_stream));
    QB_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    QB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    sum += ms;
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms_avg) *ms_avg = (float)(sum / iters);
  if (ms_min) *ms_min = best;
  return QB_OK;
}

int qb_measure_h2d(qb_ctx *ctx, int device_index, uint64_t bytes, int iters, double *gbs) {
  if (!ctx || !gbs || device_index < 0 || device_index >= (int)ctx->dev.size() || bytes == 0) return QB_ERR_ARG;
  Device &d = ctx->dev[device_index];
  QB_CUDA(ctx, cudaSetDevice(d.id));
  void *h = nullptr, *dv = nullptr;
  QB_CUDA(ctx, cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
  memset(h, 1, bytes);
  QB_CUDA(ctx, cudaMalloc(&dv, bytes));
  cudaEvent_t e0, e1;
  QB_CUDA(ctx, cudaEventCreate(&e0));
  QB_CUDA(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int i = 0; i < iters + 1; i++) {
    QB_CUDA(ctx, cudaEventRecord(e0, d.main_stream));
    QB_CUDA(ctx, cudaMemcpyAsync(dv, h, bytes, cudaMemcpyHostToDevice, d.main_stream));
    QB_CUDA(ctx, cudaEventRecord(e1, d.main_stream));
    QB_CUDA(ctx, cudaEventSynchronize(e1));
    float ms;
    QB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (i > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dv);
  cudaFreeHost(h);
  *gbs = (double)bytes / (best * 1e-3) / 1e9;
  return QB_OK;
}

// Diagnostics: host-to-device rate of ONE pass over a list of pinned host buffers (every byte read once from host
// memory, unlike the repeated copy of qb_measure_h2d whose source may sit in the CPU's last-level cache).
int qb_measure_h2d_list(qb_ctx *ctx, int device_index, const void *const *ptrs, const uint64_t *sizes, uint32_t n, double *gbs) {
  if (!ctx || !gbs || !ptrs || !sizes || n == 0 || device_index < 0 || device_index >= (int)ctx->dev.size()) return QB_ERR_ARG;
  Device &d = ctx->dev[device_index];
  QB_CUDA(ctx, cudaSetDevice(d.id));
  uint64_t mx = 0, total = 0;
  for (uint32_t i = 0; i < n; i++) mx = sizes[i] > mx ? sizes[i] : mx, total += sizes[i];
  void *dv[2] = {nullptr, nullptr};
  QB_CUDA(ctx, cudaMalloc(&dv[0], mx));
  QB_CUDA(ctx, cudaMalloc(&dv[1], mx));
  cudaEvent_t e0, e1;
  QB_CUDA(ctx, cudaEventCreate(&e0));
  QB_CUDA(ctx, cudaEventCreate(&e1));
  QB_CUDA(ctx, cudaEventRecord(e0, d.main_stream));
  for (uint32_t i = 0; i < n; i++) QB_CUDA(ctx, cudaMemcpyAsync(dv[i & 1u], ptrs[i], sizes[i], cudaMemcpyHostToDevice, d.main_stream));
  QB_CUDA(ctx, cudaEventRecord(e1, d.main_stream));
  QB_CUDA(ctx, cudaEventSynchronize(e1));
  float ms = 0;
  QB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dv[0]);
  cudaFree(dv[1]);
  *gbs = (double)total / (ms * 1e-3) / 1e9;
  return QB_OK;
}

// tools: shared-memory pipe microbenchmarks (not part of the documented boundary)
int qb_microbench(char *report, size_t cap) {
  int sm = 0;
  if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0) != cudaSuccess) return QB_ERR_CUDA;
  return qb::run_microbench(sm, report, cap) == cudaSuccess ? QB_OK : QB_ERR_CUDA;
}

// tools: the anchor filter built for the adapter set (distinct 7-mer anchors, pass rate on random bases)
int qb_adapter_filter_info(const qb_ctx *ctx, uint32_t *n_anchors, double *density) {
  if (!ctx) return QB_ERR_ARG;
  if (n_anchors) *n_anchors = ctx->n_anchors;
  if (density) *density = ctx->anchor_density;
  return QB_OK;
}

}  // extern "C"
