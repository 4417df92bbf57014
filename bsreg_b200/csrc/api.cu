// C ABI of the bsreg Gibbs hot path (see include/bsreg_cuda.h for the mapping to the
// reference's R6 protocol).  Host-side orchestration only: device memory, streams,
// CUDA graphs, chain sharding over GPUs.  No CPU fallback: without a usable CUDA
// device every entry point fails with BSREG_ENOGPU / BSREG_ECUDA.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/bsreg_cuda.h"
#include "internal.cuh"

using namespace bsreg;

namespace {

thread_local std::string g_err;

// NVTX range (nsys / ncu --nvtx timelines): create phases, run blocks
struct Nvtx {
  explicit Nvtx(const char *name) { nvtxRangePushA(name); }
  ~Nvtx() { nvtxRangePop(); }
  Nvtx(const Nvtx &) = delete;
  Nvtx &operator=(const Nvtx &) = delete;
};

// BSREG_TIMING=1: wall-clock of the create() phases on stderr (diagnostics only); every phase is also an NVTX range
struct PhaseTimer {
  bool on = getenv("BSREG_TIMING") != nullptr, open = false;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  ~PhaseTimer() { if (open) nvtxRangePop(); }
  void begin(const char *what) {
    if (open) nvtxRangePop();
    nvtxRangePushA(what);
    open = true;
  }
  void mark(const char *what) {
    if (open) { nvtxRangePop(); open = false; }
    if (!on) return;
    static const bool host_only = getenv("BSREG_TIMING") && atoi(getenv("BSREG_TIMING")) == 2;   // 2: no device sync
    if (!host_only) cudaDeviceSynchronize();
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[bsreg timing] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(e_ == cudaErrorMemoryAllocation ? BSREG_ENOMEM : BSREG_ECUDA, "%s: %s (%s:%d)",  \
                  #expr, cudaGetErrorString(e_), __FILE__, __LINE__);                              \
  } while (0)
#define TRY(expr)            \
  do {                       \
    int r_ = (expr);         \
    if (r_ != BSREG_OK) return r_; \
  } while (0)


// ---------------------------------------------------------------------------
// Caching device allocator.  cudaMalloc / cudaFree cost milliseconds (and now and then tens of
// milliseconds) each, and a fit allocates ~40 buffers whose sizes repeat exactly from one fit to
// the next (an R session calls bm() again and again), so freed blocks are kept per device and
// reused by exact size.  Blocks are only handed back after the owning stream was synchronised.
// bsreg_release_cache() returns everything to the driver.
// ---------------------------------------------------------------------------
struct FreeBlock { int dev; size_t bytes; void *ptr; };
std::mutex g_pool_mu;
std::vector<FreeBlock> g_pool_free;
std::unordered_map<void *, std::pair<int, size_t>> g_pool_live;
size_t g_pool_cached = 0;
constexpr size_t POOL_MAX_CACHED = (size_t)16 << 30;

cudaError_t pool_malloc(void **p, size_t bytes) {
  bytes = (std::max<size_t>(bytes, 1) + 511) & ~(size_t)511;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (size_t i = 0; i < g_pool_free.size(); ++i)
      if (g_pool_free[i].dev == dev && g_pool_free[i].bytes == bytes) {
        *p = g_pool_free[i].ptr;
        g_pool_cached -= bytes;
        g_pool_free[i] = g_pool_free.back();
        g_pool_free.pop_back();
        g_pool_live[*p] = {dev, bytes};
        return cudaSuccess;
      }
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation) {            // give the cache back and retry once
    cudaGetLastError();
    std::vector<FreeBlock> drop;
    { std::lock_guard<std::mutex> lk(g_pool_mu); drop.swap(g_pool_free); g_pool_cached = 0; }
    for (const FreeBlock &b : drop) { cudaSetDevice(b.dev); cudaFree(b.ptr); }
    cudaSetDevice(dev);
    e = cudaMalloc(p, bytes);
  }
  if (e == cudaSuccess) { std::lock_guard<std::mutex> lk(g_pool_mu); g_pool_live[*p] = {dev, bytes}; }
  return e;
}

void pool_free(void *p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  auto it = g_pool_live.find(p);
  if (it == g_pool_live.end()) { cudaFree(p); return; }
  const FreeBlock b{it->second.first, it->second.second, p};
  g_pool_live.erase(it);
  if (g_pool_cached + b.bytes > POOL_MAX_CACHED) { cudaFree(p); return; }
  g_pool_cached += b.bytes;
  g_pool_free.push_back(b);
}

// device temporary of a set-up phase: returned to the pool when the scope ends -- on the error paths too -- after the
// stream that used it has drained
struct PoolTmp {
  cudaStream_t st;
  void *p = nullptr;
  explicit PoolTmp(cudaStream_t s) : st(s) {}
  PoolTmp(const PoolTmp &) = delete;
  PoolTmp &operator=(const PoolTmp &) = delete;
  ~PoolTmp() { release(); }
  cudaError_t get(size_t bytes) { return pool_malloc(&p, bytes); }
  void release() { if (p) { cudaStreamSynchronize(st); pool_free(p); p = nullptr; } }
  template <class T> T *as() const { return static_cast<T *>(p); }
};

// Small pinned host blocks (results of the trace moments): cudaMallocHost / cudaFreeHost cost hundreds of microseconds
// and synchronise the device, so blocks are recycled across fits like the device buffers above.
std::vector<void *> g_pinned_free;
constexpr size_t PINNED_BLOCK = 4096;

cudaError_t pinned_get(void **p) {
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!g_pinned_free.empty()) { *p = g_pinned_free.back(); g_pinned_free.pop_back(); return cudaSuccess; }
  }
  return cudaMallocHost(p, PINNED_BLOCK);
}

void pinned_put(void *p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  g_pinned_free.push_back(p);
}

constexpr int GRAPH_UNROLL = 24;      // Gibbs iterations per CUDA graph launch (multiple of the 3 Gram buffers)
constexpr int BLOCK_ITERS = 1000;     // iterations between host syncs / progress callbacks
constexpr size_t DRAW_BUF_BYTES = (size_t)1 << 30;

struct Shard {
  int dev = 0, sm_count = 148;
  cudaStream_t stream = nullptr, aux = nullptr;   // aux: chain-step branch of the iteration graph
  bool own_stream = false;
  int64_t seq = 0;                    // index of the next sweep: it accumulates into Gram buffer seq % 3
  std::vector<cudaEvent_t> events;
  int chain_lo = 0;                   // first chain of this shard inside the handle
  DeviceData dd{};
  ChainState cs{};
  std::vector<void *> allocs;
  RunCtl *ctl_dev = nullptr;
  void *psync = nullptr;              // counters of the persistent kernel
  double *draws_dev = nullptr;
  int save_cap = 0;
  double *p0_init = nullptr, *beta0 = nullptr;
  double *scratch = nullptr;          // partial sums for deterministic reductions
  DeltaData dl{};                     // sampled SLX connectivity parameter (dl.L > 0: Psi_SLX is a function of delta)
  double *dl_partial = nullptr;       // per-CTA partials of the per-chain sweep
  bool no_coop = false, coop_checked = false;   // the cooperative (persistent) launch was refused on this device
  cudaGraphExec_t graph = nullptr, sweep_graph = nullptr, chain_graph = nullptr;

  template <class T>
  int alloc(T **p, size_t count) {
    void *q = nullptr;
    CU(pool_malloc((void **)&q, std::max<size_t>(count, 1) * sizeof(T)));
    allocs.push_back(q);
    *p = static_cast<T *>(q);
    return BSREG_OK;
  }
};

}  // namespace

struct bsreg_handle {
  int n = 0, k = 0, m = 0, d = 0, P = 0, model = 0, prior = 0, ldet = 0, stats_mode = 0, n_chains = 0;
  bool spatial = false, slx_psi = false;
  Priors pr{};
  bsreg_options opt{};
  std::vector<Shard> shards;
  std::atomic<int64_t> launches{0};   // shards are set up by one host thread each
  int64_t nnz = 0;
  std::vector<double> node_re, node_im, node_w;
  std::vector<double> spline;         // BSREG_LDET_SPLINE: [x | y | b | c | d]
  bool have_trace_w = false;
  double trace_w = 0.0;               // tr(W), from the validation pass over the CSR
};

namespace {

int upload(Shard &s, const void *src, void *dst, size_t bytes) {
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s.stream));
  return BSREG_OK;
}

// Gram buffer rotation (internal.cuh): sweep j accumulates into buffer j % 3 and clears (j + 1) % 3
void enqueue_sweep(bsreg_handle *h, Shard &s, cudaStream_t st) {
  s.dd.gcur = (int)(s.seq % 3);
  s.dd.gzero = (int)((s.seq + 1) % 3);
  launch_sweep(s.dd, s.sm_count, st);
  ++s.seq;
  ++h->launches;
}

void enqueue_chain_step(bsreg_handle *h, Shard &s, cudaStream_t st) {
  s.dd.gcur = (int)((s.seq + 2) % 3);          // buffer of the latest sweep, (seq - 1) % 3
  launch_chain_step(s.dd, h->pr, s.cs, s.ctl_dev, s.draws_dev, st);
  ++h->launches;
}

// one Gibbs iteration = [fused sweep] + chain step, serialised on the shard's stream
int enqueue_iteration(bsreg_handle *h, Shard &s) {
  if (h->stats_mode == BSREG_STATS_STREAM) enqueue_sweep(h, s, s.stream);
  enqueue_chain_step(h, s, s.stream);
  return BSREG_OK;
}

// GRAPH_UNROLL iterations as one CUDA graph.  The sweep does not depend on the chain state, so
// inside the graph the sweeps (S) and the chain steps (C) form two chains with cross edges
//   S_i <- S_{i-1}, C_{i-2}      (buffer (i + 1) % 3, cleared by S_i, was last read by C_{i-2})
//   C_i <- C_{i-1}, S_i
// and sweep i + 1 overlaps chain step i.
int build_graph(bsreg_handle *h, Shard &s) {
  if (s.graph) return BSREG_OK;
  const bool streamed = h->stats_mode == BSREG_STATS_STREAM;
  if (s.events.empty()) {
    s.events.resize(2 * GRAPH_UNROLL);
    for (cudaEvent_t &e : s.events) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  cudaEvent_t *evS = s.events.data(), *evC = evS + GRAPH_UNROLL;
  cudaGraph_t g = nullptr;
  const int64_t before = h->launches, seq0 = s.seq;
  CU(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < GRAPH_UNROLL; ++i) {
    if (streamed) {
      if (i >= 2) CU(cudaStreamWaitEvent(s.stream, evC[i - 2], 0));
      enqueue_sweep(h, s, s.stream);
      CU(cudaEventRecord(evS[i], s.stream));
      CU(cudaStreamWaitEvent(s.aux, evS[i], 0));
      enqueue_chain_step(h, s, s.aux);
      CU(cudaEventRecord(evC[i], s.aux));
    } else {
      enqueue_chain_step(h, s, s.stream);
    }
  }
  if (streamed) CU(cudaStreamWaitEvent(s.stream, evC[GRAPH_UNROLL - 1], 0));
  h->launches = before;
  s.seq = seq0;
  CU(cudaStreamEndCapture(s.stream, &g));
  CU(cudaGraphInstantiate(&s.graph, g, 0));
  CU(cudaGraphDestroy(g));
  return BSREG_OK;
}

// The cooperative launch can be refused before anything runs (then nothing must have been touched: the Gram buffers are
// cleared only once the occupancy check has passed).
static cudaError_t launch_persistent_checked(bsreg_handle *h, Shard &s, const RunCtl &ctl, int count, bool cached, size_t gbytes) {
  if (!s.coop_checked) {
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s.dev);
    s.coop_checked = true;
    if (!coop) return cudaErrorNotSupported;
  }
  cudaError_t e = cudaMemsetAsync(s.dd.Gfix, 0, gbytes, s.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(s.psync, 0, persistent_sync_bytes(), s.stream);
  if (e != cudaSuccess) return e;
  const char *force = getenv("BSREG_FORCE_COOP_FAIL");             // test hook: behave as if the launch had been refused
  if (force && *force == '1') e = cudaErrorCooperativeLaunchTooLarge;
  else e = launch_persistent(s.dd, h->pr, s.cs, ctl, s.ctl_dev, s.draws_dev, count, cached, s.psync, s.sm_count, s.stream);
  if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
    // the buffers were cleared: rebuild the Gram matrix the one-launch path expects (one sweep into buffer 0)
    cudaGetLastError();
    s.seq = 0;
    enqueue_sweep(h, s, s.stream);
  }
  return e;
}

int enqueue_iterations_delta(bsreg_handle *h, Shard &s, int count);

int enqueue_iterations(bsreg_handle *h, Shard &s, int count, const RunCtl &ctl) {
  if (s.dl.L > 0) return enqueue_iterations_delta(h, s, count);
  const bool streamed = h->stats_mode == BSREG_STATS_STREAM;
  if (count >= 3 && !s.no_coop && persistent_supported(s.dd, h->pr, s.cs.n_chains, s.sm_count, streamed)) {
    // one cooperative launch for the whole block: sweep CTAs and chain CTAs meet at the triple-buffered Gram matrix
    const size_t gbytes = (size_t)3 * h->d * h->d * 8;
    const cudaError_t le = launch_persistent_checked(h, s, ctl, count, !streamed, gbytes);
    if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorNotSupported || le == cudaErrorLaunchOutOfResources) {
      // not every CTA can be resident here (MPS share, MIG slice, ...): this device keeps to one launch per stage
      cudaGetLastError();
      s.no_coop = true;
      return enqueue_iterations(h, s, count, ctl);
    }
    CU(le);
    // the last Gram matrix is in buffer (count - 1) % 3; the other two were consumed but not recycled: clear them
    for (int k = 0; k < 2; ++k)
      CU(cudaMemsetAsync(s.dd.Gfix + (size_t)((count + k) % 3) * h->d * h->d, 0, gbytes / 3, s.stream));
    s.seq = 3 + count % 3;
    h->launches += 1;
    return BSREG_OK;
  }
  if (!streamed && count > 1 && chain_step_loops(s.dd, h->pr)) {
    // cached statistics, general chain kernel: the whole block in ONE launch (the kernel loops; its code stays in the
    // instruction cache from the second iteration on)
    s.dd.gcur = (int)((s.seq + 2) % 3);          // buffer of the latest sweep, as enqueue_chain_step
    launch_chain_step(s.dd, h->pr, s.cs, s.ctl_dev, s.draws_dev, s.stream, count);
    ++h->launches;
    CU(cudaGetLastError());
    return BSREG_OK;
  }
  const int per_iter = streamed ? 2 : 1;
  while (count > 0) {
    if (count >= GRAPH_UNROLL && (!streamed || s.seq % 3 == 0)) {
      TRY(build_graph(h, s));
      CU(cudaGraphLaunch(s.graph, s.stream));
      h->launches += (int64_t)GRAPH_UNROLL * per_iter;
      if (streamed) s.seq += GRAPH_UNROLL;
      count -= GRAPH_UNROLL;
    } else {
      enqueue_iteration(h, s);
      --count;
    }
  }
  CU(cudaGetLastError());
  return BSREG_OK;
}

// measurement helpers: reps back-to-back launches of one kernel, as graphs of GRAPH_UNROLL where possible
int enqueue_repeated(bsreg_handle *h, Shard &s, int reps, bool sweep) {
  cudaGraphExec_t &ge = sweep ? s.sweep_graph : s.chain_graph;
  while (reps > 0) {
    if (reps >= GRAPH_UNROLL && (!sweep || s.seq % 3 == 0)) {
      if (!ge) {
        cudaGraph_t g = nullptr;
        const int64_t before = h->launches, seq0 = s.seq;
        CU(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < GRAPH_UNROLL; ++i) { if (sweep) enqueue_sweep(h, s, s.stream); else enqueue_chain_step(h, s, s.stream); }
        h->launches = before;
        s.seq = seq0;
        CU(cudaStreamEndCapture(s.stream, &g));
        CU(cudaGraphInstantiate(&ge, g, 0));
        CU(cudaGraphDestroy(g));
      }
      CU(cudaGraphLaunch(ge, s.stream));
      h->launches += GRAPH_UNROLL;
      if (sweep) s.seq += GRAPH_UNROLL;
      reps -= GRAPH_UNROLL;
    } else {
      if (sweep) enqueue_sweep(h, s, s.stream); else enqueue_chain_step(h, s, s.stream);
      --reps;
    }
  }
  CU(cudaGetLastError());
  return BSREG_OK;
}

// Chebyshev / Hutchinson trace moments, in two halves so that bsreg_create can run the device work on
// the auxiliary stream beside the upload of y and X, the relabelling and the tile packing: launch() only enqueues,
// finish() waits for the results.
struct MomentsJob {
  bool active = false;
  int order = 0, probes = 0;
  double m1 = 0.0;
  double *U = nullptr, *buf[3] = {nullptr, nullptr, nullptr}, *res = nullptr, *rowsum = nullptr, *scratch = nullptr;
  double *partial = nullptr;          // tiles mode: one row of per-CTA dot partials per order
  double *host = nullptr;             // pinned: res[0 .. order], then tr(W^2) partial at [order + 1]
  cudaEvent_t done = nullptr;
};

int trace_moments_launch(bsreg_handle *h, Shard &s, const bsreg_problem *prob_csr, int order, int probes, uint64_t seed,
                         MomentsJob &job, cudaStream_t st) {
  const int n = h->n;
  job.order = order; job.probes = probes; job.active = true;
  CU(cudaSetDevice(s.dev));
  if ((size_t)(order + 2) * 8 > PINNED_BLOCK) return fail(BSREG_EINVAL, "Chebyshev order too large");
  CU(pinned_get((void **)&job.host));
  for (int j = 0; j <= order + 1; ++j) job.host[j] = 0.0;
  CU(cudaEventCreateWithFlags(&job.done, cudaEventDisableTiming));
  CU(pool_malloc((void **)&job.res, (size_t)(order + 2) * 8));
  CU(pool_malloc((void **)&job.scratch, 4096 * 8));
  if (order >= 2) {   // tr(W^2) = sum_ij W_ij W_ji, exact from the CSR on the device
    CU(pool_malloc((void **)&job.rowsum, (size_t)n * 8));
    launch_trace_w2_rows(n, s.dd.rp, s.dd.ci, s.dd.va, job.rowsum, st);
    launch_dot(job.rowsum, nullptr, n, job.scratch, job.res + order + 1, st);
    h->launches += 3;
  }
  if (order >= 3) {
    const size_t len = (size_t)n * probes;
    CU(pool_malloc((void **)&job.U, len * 8));
    for (int b = 0; b < 3; ++b) CU(pool_malloc((void **)&job.buf[b], len * 8));
    const int parts = spmm_dot_parts(n, probes);
    CU(pool_malloc((void **)&job.partial, (size_t)(order + 1) * parts * 8));
    launch_probes(job.U, n, probes, seed, st);
    // T_1 = W U ; T_j = 2 W T_{j-1} - T_{j-2}; U stays intact for the dots u'T_j u
    launch_spmm(n, s.dd.rp, s.dd.ci, s.dd.va, job.U, probes, probes, job.buf[0], probes, 0, 1.0, nullptr, 0.0, st);
    h->launches += 2;
    const double *prev = job.U;
    double *cur = job.buf[0];
    int nxt = 1;
    for (int j = 2; j <= order; ++j) {
      double *next = job.buf[nxt];
      if (j >= 3) {          // the dot u'T_j u rides on the product
        launch_spmm_dot(n, s.dd.rp, s.dd.ci, s.dd.va, cur, probes, next, 2.0, prev, -1.0, job.U,
                        job.partial + (size_t)j * parts, nullptr, st);
        ++h->launches;
      } else {
        launch_spmm(n, s.dd.rp, s.dd.ci, s.dd.va, cur, probes, probes, next, probes, 0, 2.0, prev, -1.0, st);
        ++h->launches;
      }
      prev = cur; cur = next; nxt = (nxt + 1) % 3;
    }
    // the dots of all orders in one launch (one warp per order, fixed order) instead of one small launch per order
    launch_combine_rows(job.partial + (size_t)3 * parts, parts, order - 2, job.res + 3, st);
    ++h->launches;
  }
  CU(cudaMemcpyAsync(job.host, job.res, (size_t)(order + 2) * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaEventRecord(job.done, st));
  // tr(W): summed during the validation pass of bsreg_create; here only for the parity entry point
  if (order >= 1) {
    if (h->have_trace_w) {
      job.m1 = h->trace_w;
    } else {
      long double tr = 0.0L;
      for (int i = 0; i < n; ++i)
        for (int p = prob_csr->row_ptr[i]; p < prob_csr->row_ptr[i + 1]; ++p)
          if (prob_csr->col_idx[p] == i) tr += prob_csr->val[p];
      job.m1 = (double)tr;
    }
  }
  return BSREG_OK;
}

int trace_moments_finish(bsreg_handle *h, Shard &s, MomentsJob &job, double *moments) {
  if (!job.active) return BSREG_OK;
  const int order = job.order, n = h->n;
  CU(cudaSetDevice(s.dev));
  CU(cudaEventSynchronize(job.done));
  for (int j = 0; j <= order; ++j) moments[j] = 0.0;
  moments[0] = n;
  if (order >= 1) moments[1] = job.m1;
  if (order >= 2) moments[2] = 2.0 * job.host[order + 1] - n;
  for (int j = 3; j <= order; ++j) moments[j] = job.host[j] / job.probes;
  if (job.U) pool_free(job.U);
  for (int b = 0; b < 3; ++b) if (job.buf[b]) pool_free(job.buf[b]);
  if (job.rowsum) pool_free(job.rowsum);
  pool_free(job.res); pool_free(job.scratch);
  if (job.partial) pool_free(job.partial);
  pinned_put(job.host);
  cudaEventDestroy(job.done);
  job = MomentsJob();
  return BSREG_OK;
}

int trace_moments_dev(bsreg_handle *h, Shard &s, const bsreg_problem *prob_csr, int order, int probes, uint64_t seed,
                      double *moments) {
  MomentsJob job;
  TRY(trace_moments_launch(h, s, prob_csr, order, probes, seed, job, s.stream));
  return trace_moments_finish(h, s, job, moments);
}

void chebyshev_nodes(const std::vector<double> &mom, int n, std::vector<double> &re, std::vector<double> &im,
                     std::vector<double> &w) {
  const int q = (int)mom.size() - 1;
  re.assign(q + 1, 0.0); im.assign(q + 1, 0.0); w.assign(q + 1, 0.0);
  for (int k = 1; k <= q + 1; ++k) {
    const double theta = M_PI * (k - 0.5) / (q + 1);
    double acc = 0.5 * n;
    for (int j = 1; j <= q; ++j) acc += mom[j] * std::cos(theta * j);
    re[k - 1] = std::cos(theta);
    w[k - 1] = 2.0 / (q + 1) * acc;
  }
}

int set_nodes(bsreg_handle *h) {
  for (Shard &s : h->shards) {
    CU(cudaSetDevice(s.dev));
    const int nn = (int)h->node_re.size();
    double *re, *im, *w;
    TRY(s.alloc(&re, nn)); TRY(s.alloc(&im, nn)); TRY(s.alloc(&w, nn));
    TRY(upload(s, h->node_re.data(), re, nn * 8));
    TRY(upload(s, h->node_im.data(), im, nn * 8));
    TRY(upload(s, h->node_w.data(), w, nn * 8));
    s.dd.n_nodes = nn; s.dd.node_re = re; s.dd.node_im = im; s.dd.node_w = w;
  }
  return BSREG_OK;
}

// Does W lack locality?  (share of nonzeros further than max(1024, n/64) from the diagonal above 20 %: e.g. points in
// random order; a panel's block-diagonal kronecker W or spatially sorted observations stay as they are)
static bool wants_reorder(int n, const int *rp, const int *ci) {
  const char *env = getenv("BSREG_REORDER");                         // A/B switch: 0 = never, 1 = always (read per create)
  if (env && *env) return atoi(env) != 0;
  if (n < 4096) return false;
  const long long far = std::max(1024, n / 64);
  long long cnt = 0, tot = 0;
  const int step = std::max(1, n / 8192);                             // a sample of the rows is enough
  for (int i = 0; i < n; i += step)
    for (int q = rp[i]; q < rp[i + 1]; ++q) { ++tot; cnt += std::llabs((long long)ci[q] - i) > far; }
  return tot > 0 && cnt * 5 > tot;
}

// Breadth-first relabelling over the out-edges of W (queue order, restarted at the smallest unvisited observation):
// consecutive labels are graph neighbours of neighbours, so the rows of a 64-row tile share most of their columns.
static void bfs_order(int n, const int *rp, const int *ci, std::vector<int> &perm) {
  perm.resize(n);
  std::vector<unsigned char> seen(n, 0);
  int head = 0, tail = 0;
  for (int root = 0; root < n; ++root) {
    if (seen[root]) continue;
    seen[root] = 1; perm[tail++] = root;
    while (head < tail) {
      // the queue is the prefetch list: row pointers 8 vertices ahead, their column lists 4 ahead
      if (head + 8 < tail) __builtin_prefetch(rp + perm[head + 8]);
      if (head + 4 < tail) __builtin_prefetch(ci + rp[perm[head + 4]]);
      const int v = perm[head++];
      for (int q = rp[v]; q < rp[v + 1]; ++q) {
        const unsigned u = (unsigned)ci[q];
        if (u < (unsigned)n && !seen[u]) { seen[u] = 1; perm[tail++] = (int)u; }   // (a bad index is reported by the validation pass)
      }
    }
  }
}

// the internal ordering is the same for every shard: whichever shard thread gets there first computes it
struct SharedOrder {
  std::once_flag once;
  std::vector<int> perm;
  std::thread worker;                 // started by bsreg_create before the validation pass and the uploads
  ~SharedOrder() { if (worker.joinable()) worker.join(); }
  void start(bool wanted, int n, const int *rp, const int *ci) {
    if (!wanted) return;
    worker = std::thread([this, n, rp, ci] { if (wants_reorder(n, rp, ci)) bfs_order(n, rp, ci, perm); });
  }
  const std::vector<int> &get() {
    std::call_once(once, [this] { if (worker.joinable()) worker.join(); });
    return perm;
  }
};

int setup_shard(bsreg_handle *h, Shard &s, const bsreg_problem *p, const bsreg_options *o, int n_chains, int chain_lo,
                MomentsJob *moments_job, SharedOrder &order) {
  CU(cudaSetDevice(s.dev));
  CU(cudaDeviceGetAttribute(&s.sm_count, cudaDevAttrMultiProcessorCount, s.dev));   // (cudaGetDeviceProperties costs milliseconds)
  if (o->stream && h->shards.size() == 1) { s.stream = (cudaStream_t)o->stream; s.own_stream = false; }
  else { CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)); s.own_stream = true; }
  CU(cudaStreamCreateWithFlags(&s.aux, cudaStreamNonBlocking));
  s.chain_lo = chain_lo;
  const int n = h->n, k = h->k, m = h->m, d = h->d;
  DeviceData &dd = s.dd;
  dd.n = n; dd.m = m; dd.d = d; dd.model = h->model; dd.nnz = p->nnz;

  PhaseTimer pt;
  pt.begin("bsreg_create: W upload");
  // ---- W (CSR) ----
  int *rp = nullptr, *ci = nullptr; double *va = nullptr;
  if (p->nnz > 0 || h->model != BSREG_MODEL_LM || p->k_slx > 0) {
    // padded so that the 16-byte staged copies of the sweep may over-read past the end
    TRY(s.alloc(&rp, (size_t)n + 1 + 272)); TRY(s.alloc(&ci, (size_t)p->nnz + 8)); TRY(s.alloc(&va, (size_t)p->nnz + 4));
    CU(cudaMemsetAsync(rp, 0, ((size_t)n + 1 + 272) * 4, s.stream));
    CU(cudaMemsetAsync(ci, 0, ((size_t)p->nnz + 8) * 4, s.stream));
    CU(cudaMemsetAsync(va, 0, ((size_t)p->nnz + 4) * 8, s.stream));
    TRY(upload(s, p->row_ptr, rp, ((size_t)n + 1) * 4));
    TRY(upload(s, p->col_idx, ci, (size_t)p->nnz * 4));
    TRY(upload(s, p->val, va, (size_t)p->nnz * 8));
  }
  dd.rp = rp; dd.ci = ci; dd.va = va;
  if (moments_job != nullptr && p->nnz > 0) {
    // the trace moments need W only: their device work runs on the auxiliary stream beside the rest of the setup
    cudaEvent_t w_up;
    CU(cudaEventCreateWithFlags(&w_up, cudaEventDisableTiming));
    CU(cudaEventRecord(w_up, s.stream));
    CU(cudaStreamWaitEvent(s.aux, w_up, 0));
    CU(cudaEventDestroy(w_up));
    const int order = o->cheb_order > 0 ? o->cheb_order : 50, probes = o->cheb_probes > 0 ? o->cheb_probes : 64;
    TRY(trace_moments_launch(h, s, p, order, probes, o->seed, *moments_job, s.aux));
  }

  pt.mark("W upload");
  pt.begin("bsreg_create: y, X upload + transpose");
  // ---- y, X (column-major on the host -> row-major [n][m] in HBM) ----
  double *y, *X;
  TRY(s.alloc(&y, (size_t)n)); TRY(s.alloc(&X, (size_t)n * m));
  PoolTmp tmp(s.stream);
  CU(tmp.get((size_t)n * std::max(k, std::max(p->k_slx, 1)) * 8));
  TRY(upload(s, p->y, y, (size_t)n * 8));
  TRY(upload(s, p->X, tmp.as<double>(), (size_t)n * k * 8));
  launch_transpose(tmp.as<double>(), X, n, k, m, 0, s.stream);
  if (p->k_slx > 0 && o->slx_psi) {
    // setup_2SLX with a connectivity FUNCTION (R/12_slx.R:69-74): pattern, log distances and X_SLX stay on the device;
    // the lag columns of the design start at Psi(delta0) X_SLX (set_delta(priors$delta))
    const int64_t nz = p->nnz_slx > 0 ? p->nnz_slx : p->nnz;
    const int32_t *hrp = p->nnz_slx > 0 ? p->row_ptr_slx : p->row_ptr, *hci = p->nnz_slx > 0 ? p->col_idx_slx : p->col_idx;
    const double *hva = p->nnz_slx > 0 ? p->val_slx : p->val;
    std::vector<double> logd((size_t)nz);
    for (int64_t q = 0; q < nz; ++q) logd[q] = std::log(hva[q]);
    int *drp, *dci; double *dlog, *dxl;
    TRY(s.alloc(&drp, (size_t)n + 1)); TRY(s.alloc(&dci, (size_t)nz)); TRY(s.alloc(&dlog, (size_t)nz));
    TRY(s.alloc(&dxl, (size_t)n * p->k_slx));
    CU(cudaStreamSynchronize(s.stream));                       // tmp is reused for X_SLX
    TRY(upload(s, hrp, drp, ((size_t)n + 1) * 4)); TRY(upload(s, hci, dci, (size_t)nz * 4));
    TRY(upload(s, logd.data(), dlog, (size_t)nz * 8));
    TRY(upload(s, p->X_slx, tmp.as<double>(), (size_t)n * p->k_slx * 8));
    launch_transpose(tmp.as<double>(), dxl, n, p->k_slx, p->k_slx, 0, s.stream);
    DeltaData &dl = s.dl;
    dl.L = p->k_slx; dl.norm = o->slx_psi_norm; dl.sampled = o->slx_delta_scale > 0.0 ? 1 : 0;
    dl.rp = drp; dl.ci = dci; dl.logd = dlog; dl.Xl = dxl;
    dl.prior_a = o->slx_delta_a; dl.prior_b = o->slx_delta_b; dl.lo = o->slx_delta_min; dl.hi = o->slx_delta_max;
    launch_slx_power_lag(dl, o->slx_delta, X, m, k, n, s.stream);
    CU(cudaStreamSynchronize(s.stream));                       // `logd` is a host temporary
    h->launches += 2;
  } else if (p->k_slx > 0) {   // setup_2SLX fixed-W branch, R/12_slx.R:60-68: X <- [X, Psi_SLX %*% X_SLX]
    const int *rps = rp, *cis = ci; const double *vas = va;
    PoolTmp xs_rm(s.stream), trp(s.stream), tci(s.stream), tva(s.stream);
    CU(cudaStreamSynchronize(s.stream));
    TRY(upload(s, p->X_slx, tmp.as<double>(), (size_t)n * p->k_slx * 8));
    CU(xs_rm.get((size_t)n * p->k_slx * 8));
    launch_transpose(tmp.as<double>(), xs_rm.as<double>(), n, p->k_slx, p->k_slx, 0, s.stream);
    if (p->nnz_slx > 0) {
      CU(trp.get(((size_t)n + 1) * 4)); CU(tci.get((size_t)p->nnz_slx * 4)); CU(tva.get((size_t)p->nnz_slx * 8));
      TRY(upload(s, p->row_ptr_slx, trp.p, ((size_t)n + 1) * 4));
      TRY(upload(s, p->col_idx_slx, tci.p, (size_t)p->nnz_slx * 4));
      TRY(upload(s, p->val_slx, tva.p, (size_t)p->nnz_slx * 8));
      rps = trp.as<int>(); cis = tci.as<int>(); vas = tva.as<double>();
    }
    launch_spmm(n, rps, cis, vas, xs_rm.as<double>(), p->k_slx, p->k_slx, X, m, k, 1.0, nullptr, 0.0, s.stream);
    h->launches += 2;
  }
  // internal ordering of the observations for the tiled sweep (the Gram matrix does not depend on it): breadth-first when
  // W has no locality.  Host work, done here so that it runs beside the uploads and transposes enqueued above.
  pt.mark("y, X enqueue");
  const std::vector<int> &perm = order.get();
  pt.mark("host ordering (BFS)");
  tmp.release();                                             // drains the stream first
  dd.y = y; dd.X = X;
  h->launches += 1;
  pt.mark("y, X upload + transpose");
  pt.begin("bsreg_create: tile packing");
  // ---- tile-blocked operands of the ring sweep (sweep.cu): Zt and the per-tile CSR blobs ----
  dd.Zt = nullptr; dd.Xr = nullptr; dd.tile_ul = nullptr; dd.csr_blob = nullptr; dd.tile_off = nullptr; dd.tile_len = nullptr; dd.yp = y;
  dd.tile_max_blob = 0; dd.tile_max_nu = 0;
  dd.sw_rp = nullptr; dd.sw_ci = nullptr; dd.sw_va = nullptr; dd.sw_X = nullptr; dd.sw_y = nullptr;
  {
    constexpr int R = SWEEP_TILE_ROWS;
    const int ntiles = (n + R - 1) / R;
    const bool with_w = h->model != BSREG_MODEL_LM;
    const bool permuted = with_w && !perm.empty();
    std::vector<long long> off(ntiles + 1, 0);
    int cap_blob = 0, cap_nu = 0;
    if (with_w) {
      for (int t = 0; t < ntiles; ++t) {
        int nz = 0, max_len = 0;
        for (int r = t * R; r < std::min(t * R + R, n); ++r) {
          const int src = permuted ? perm[r] : r;
          const int len = p->row_ptr[src + 1] - p->row_ptr[src];
          nz += len;
          max_len = std::max(max_len, len);
        }
        const int cap = tile_blob_bytes(nz, nz, ell_width(nz, max_len));   // room for a tile whose columns are all distinct
        off[t + 1] = off[t] + cap;
        cap_blob = std::max(cap_blob, cap);
        cap_nu = std::max(cap_nu, nz);
      }
    }
    // SEM (d = 2M + 2 <= 46): ring sweep of semsweep.cu.  Whether a tile fits a stage depends on how many FOREIGN rows
    // it references, known only after packing: pack first, keep the result if the geometry fits.
    bool sem_ring = false;
    const char *nsr = getenv("BSREG_NO_SEM_RING");                                 // A/B switch (read per create)
    const bool no_sem_ring = nsr != nullptr && *nsr != 0 && *nsr != '0';
    if (h->model == BSREG_MODEL_SEM && m <= 22 && with_w && !no_sem_ring) {
      const int ldx = sem_ring_ldx(m);
      const size_t nrows = (size_t)ntiles * R;
      PoolTmp tperm(s.stream), tinv(s.stream), txr(s.stream), typ(s.stream), tdb(s.stream), tdoff(s.stream), tdlen(s.stream);
      int *dperm = nullptr, *dinv = nullptr;
      if (permuted) {
        std::vector<int> inv(n);
        for (int r = 0; r < n; ++r) inv[perm[r]] = r;
        CU(tperm.get((size_t)n * 4)); CU(tinv.get((size_t)n * 4));
        dperm = tperm.as<int>(); dinv = tinv.as<int>();
        TRY(upload(s, perm.data(), dperm, (size_t)n * 4));
        TRY(upload(s, inv.data(), dinv, (size_t)n * 4));
        CU(cudaStreamSynchronize(s.stream));                 // `inv` is a host temporary
      }
      CU(txr.get(nrows * ldx * 8)); CU(typ.get(nrows * 8));
      CU(tdb.get((size_t)off[ntiles] + 16)); CU(tdoff.get(((size_t)ntiles + 1) * 8)); CU(tdlen.get(((size_t)ntiles + 2) * 4));
      launch_pack_rows(X, y, dperm, txr.as<double>(), typ.as<double>(), n, m, ldx, s.stream);
      int *dmax = tdlen.as<int>() + ntiles;
      CU(cudaMemsetAsync(dmax, 0, 8, s.stream));
      TRY(upload(s, off.data(), tdoff.p, ((size_t)ntiles + 1) * 8));
      launch_pack_csr_tiles(rp, ci, va, dperm, dinv, tdoff.as<long long>(), tdlen.as<int>(), dmax, tdb.as<unsigned char>(), n,
                            true, s.stream);
      h->launches += 2;
      int mx[2] = {0, 0};
      CU(cudaMemcpyAsync(mx, dmax, 8, cudaMemcpyDeviceToHost, s.stream));
      CU(cudaStreamSynchronize(s.stream));
      // the lists of foreign rows leave the blobs for a ring of their own (fixed-size records)
      PoolTmp tul(s.stream);
      const int u_ints = sem_ul_ints(mx[1]);
      CU(tul.get((size_t)ntiles * u_ints * 4));
      CU(cudaMemsetAsync(dmax, 0, 4, s.stream));
      launch_extract_ul(tdb.as<unsigned char>(), tdoff.as<long long>(), tdlen.as<int>(), dmax, tul.as<int>(), u_ints, n, s.stream);
      ++h->launches;
      CU(cudaMemcpyAsync(mx, dmax, 4, cudaMemcpyDeviceToHost, s.stream));
      CU(cudaStreamSynchronize(s.stream));
      if (sem_ring_supported(n, m, mx[0], mx[1])) {
        sem_ring = true;
        dd.Xr = txr.as<double>(); dd.yp = typ.as<double>(); dd.tile_ul = tul.as<int>();
        dd.csr_blob = tdb.as<unsigned char>(); dd.tile_off = tdoff.as<long long>(); dd.tile_len = tdlen.as<int>();
        dd.tile_max_blob = mx[0]; dd.tile_max_nu = mx[1];
        for (PoolTmp *t : {&txr, &typ, &tdb, &tdoff, &tdlen, &tul}) { s.allocs.push_back(t->p); t->p = nullptr; }   // the shard keeps them
      }
    }
    if (sem_ring) {
    } else if (sweep_ring_supported(d, h->model, cap_blob, cap_nu)) {
      double *zt, *yp = nullptr;
      int *dperm = nullptr, *dinv = nullptr;
      PoolTmp tperm(s.stream), tinv(s.stream);
      if (permuted) {
        std::vector<int> inv(n);
        for (int r = 0; r < n; ++r) inv[perm[r]] = r;
        CU(tperm.get((size_t)n * 4)); CU(tinv.get((size_t)n * 4));
        dperm = tperm.as<int>(); dinv = tinv.as<int>();
        TRY(upload(s, perm.data(), dperm, (size_t)n * 4));
        TRY(upload(s, inv.data(), dinv, (size_t)n * 4));
        CU(cudaStreamSynchronize(s.stream));                 // `inv` is a host temporary
        TRY(s.alloc(&yp, (size_t)n));
      }
      TRY(s.alloc(&zt, (size_t)ntiles * (m + 1) * R));
      launch_pack_tiles(X, y, dperm, zt, yp, n, m, s.stream);
      h->launches += 1;
      dd.Zt = zt;
      if (permuted) dd.yp = yp;
      if (with_w) {
        unsigned char *db; long long *doff; int *dlen, *dmax;
        TRY(s.alloc(&db, (size_t)off[ntiles] + 16)); TRY(s.alloc(&doff, (size_t)ntiles + 1)); TRY(s.alloc(&dlen, (size_t)ntiles + 2));
        dmax = dlen + ntiles;
        CU(cudaMemsetAsync(dmax, 0, 8, s.stream));
        TRY(upload(s, off.data(), doff, ((size_t)ntiles + 1) * 8));
        launch_pack_csr_tiles(rp, ci, va, dperm, dinv, doff, dlen, dmax, db, n, false, s.stream);
        h->launches += 1;
        int mx[2] = {0, 0};
        CU(cudaMemcpyAsync(mx, dmax, 8, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaStreamSynchronize(s.stream));                 // `off` is a host temporary
        dd.csr_blob = db; dd.tile_off = doff; dd.tile_len = dlen; dd.tile_max_blob = mx[0]; dd.tile_max_nu = mx[1];
      } else {
        CU(cudaStreamSynchronize(s.stream));
      }
    } else if (permuted) {
      // designs the ring does not cover (SEM, 24 < d <= 88) sweep W, X, y directly: give them the internal ordering too
      std::vector<int> inv(n), rpn(n + 1, 0);
      for (int r = 0; r < n; ++r) { inv[perm[r]] = r; rpn[r + 1] = rpn[r] + (p->row_ptr[perm[r] + 1] - p->row_ptr[perm[r]]); }
      int *dperm = nullptr, *dinv = nullptr, *srp, *sci;
      double *sva, *sX, *sy;
      PoolTmp tperm(s.stream), tinv(s.stream);
      CU(tperm.get((size_t)n * 4)); CU(tinv.get((size_t)n * 4));
      dperm = tperm.as<int>(); dinv = tinv.as<int>();
      TRY(s.alloc(&srp, (size_t)n + 1)); TRY(s.alloc(&sci, (size_t)p->nnz + 8)); TRY(s.alloc(&sva, (size_t)p->nnz + 4));
      TRY(s.alloc(&sX, (size_t)n * m)); TRY(s.alloc(&sy, (size_t)n));
      TRY(upload(s, perm.data(), dperm, (size_t)n * 4));
      TRY(upload(s, inv.data(), dinv, (size_t)n * 4));
      TRY(upload(s, rpn.data(), srp, ((size_t)n + 1) * 4));
      launch_permute_problem(rp, ci, va, X, y, dperm, dinv, srp, sci, sva, sX, sy, n, m, s.stream);
      h->launches += 1;
      CU(cudaStreamSynchronize(s.stream));                   // host temporaries
      dd.sw_rp = srp; dd.sw_ci = sci; dd.sw_va = sva; dd.sw_X = sX; dd.sw_y = sy;
    }
  }

  pt.mark("tile packing");
  pt.begin("bsreg_create: lags, scales, first sweep");
  // ---- cached lags (R/15_sar.R:57, R/15_sem.R:63-64) ----
  dd.Wy = nullptr; dd.WX = nullptr;
  if (h->model != BSREG_MODEL_LM) {
    TRY(s.alloc(&dd.Wy, (size_t)n));
    launch_spmm(n, rp, ci, va, y, 1, 1, dd.Wy, 1, 0, 1.0, nullptr, 0.0, s.stream);
    ++h->launches;
  }
  if (h->model == BSREG_MODEL_SEM) {
    TRY(s.alloc(&dd.WX, (size_t)n * m));
    launch_spmm(n, rp, ci, va, X, m, m, dd.WX, m, 0, 1.0, nullptr, 0.0, s.stream);
    ++h->launches;
  }

  // ---- Gram accumulators and fixed-point scales ----
  double *norms, *gscale, *ginv;
  TRY(s.alloc(&dd.Gfix, (size_t)3 * d * d));
  const int nparts = colnorm_parts();
  TRY(s.alloc(&norms, (size_t)nparts * d)); TRY(s.alloc(&gscale, (size_t)d * d)); TRY(s.alloc(&ginv, (size_t)d * d));
  TRY(s.alloc(&s.scratch, 4096));
  CU(cudaMemsetAsync(dd.Gfix, 0, (size_t)3 * d * d * 8, s.stream));
  launch_colnorms(dd, norms, s.stream);
  ++h->launches;
  std::vector<double> hn(d, 0.0), hpart((size_t)nparts * d), hs((size_t)d * d), hi((size_t)d * d);
  CU(cudaMemcpyAsync(hpart.data(), norms, (size_t)nparts * d * 8, cudaMemcpyDeviceToHost, s.stream));
  CU(cudaStreamSynchronize(s.stream));
  for (int b = 0; b < nparts; ++b)
    for (int i = 0; i < d; ++i) hn[i] += hpart[(size_t)b * d + i];
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      const double bound = std::sqrt(hn[i] * hn[j]);
      int e = 0;
      if (bound > 0.0 && std::isfinite(bound)) e = std::ilogb(bound) + 1;
      hs[(size_t)i * d + j] = std::ldexp(1.0, e - 61);
      hi[(size_t)i * d + j] = std::ldexp(1.0, 61 - e);
    }
  TRY(upload(s, hs.data(), gscale, (size_t)d * d * 8));
  TRY(upload(s, hi.data(), ginv, (size_t)d * d * 8));
  CU(cudaStreamSynchronize(s.stream));
  dd.Gscale = gscale; dd.Ginv_scale = ginv;
  enqueue_sweep(h, s, s.stream);

  pt.mark("lags, scales, first sweep");
  pt.begin("bsreg_create: chain state");
  // ---- chain state ----
  ChainState &cs = s.cs;
  cs.n_chains = n_chains; cs.chain_offset = o->chain_offset + chain_lo; cs.seed = o->seed;
  const int A = aux_len(m);
  double *mu0;
  TRY(s.alloc(&cs.beta, (size_t)n_chains * m)); TRY(s.alloc(&cs.sigma2, (size_t)n_chains)); TRY(s.alloc(&cs.lam, (size_t)n_chains));
  TRY(s.alloc(&cs.p0, (size_t)n_chains * m)); TRY(s.alloc(&cs.aux, (size_t)n_chains * A));
  TRY(s.alloc(&cs.mh_scale, (size_t)n_chains * 2)); TRY(s.alloc(&cs.mh_cnt, (size_t)n_chains * 8));
  TRY(s.alloc(&cs.iters, (size_t)n_chains)); TRY(s.alloc(&cs.step, (size_t)n_chains));
  TRY(s.alloc(&mu0, (size_t)m)); TRY(s.alloc(&s.p0_init, (size_t)m)); TRY(s.alloc(&s.ctl_dev, 1));
  { unsigned char *ps; TRY(s.alloc(&ps, persistent_sync_bytes())); s.psync = ps; }
  CU(cudaMemsetAsync(cs.aux, 0, (size_t)n_chains * A * 8, s.stream));
  CU(cudaMemsetAsync(cs.step, 0, (size_t)n_chains * 4, s.stream));
  std::vector<double> hmu(m), hp0(m);
  for (int a = 0; a < m; ++a) {
    hmu[a] = o->ng_mu ? o->ng_mu[o->ng_mu_len == 1 ? 0 : a] : 0.0;
    hp0[a] = o->ng_precision ? o->ng_precision[o->ng_precision_len == 1 ? 0 : a] : 1e-8;
  }
  TRY(upload(s, hmu.data(), mu0, (size_t)m * 8));
  TRY(upload(s, hp0.data(), s.p0_init, (size_t)m * 8));
  if (o->ng_beta0) { TRY(s.alloc(&s.beta0, (size_t)m)); TRY(upload(s, o->ng_beta0, s.beta0, (size_t)m * 8)); }
  CU(cudaStreamSynchronize(s.stream));
  cs.mu0 = mu0;
  pt.mark("chain state");
  return BSREG_OK;
}

// Grid + spline log-determinant (R/15_sar.R:79-87): determinant grid on shard 0 (one CTA per grid point, dense LU with
// partial pivoting), FMM coefficients on the host, the five arrays to every shard.
int build_spline(bsreg_handle *h, const bsreg_options *o) {
  const int G = o->ldet_grid_len, reps = o->ldet_reps > 0 ? o->ldet_reps : 1, size = h->n / reps;
  std::vector<double> &sp = h->spline;
  sp.assign((size_t)5 * G, 0.0);
  for (int g = 0; g < G; ++g)                                   // i_seq, R/00_aux.R:105-107: from + (0:n1) * by
    sp[g] = G == 1 ? o->ldet_grid_from : o->ldet_grid_from + (o->ldet_grid_to - o->ldet_grid_from) * g / (G - 1);
  {
    Shard &s = h->shards[0];
    CU(cudaSetDevice(s.dev));
    PoolTmp A(s.stream), xs(s.stream), out(s.stream);
    CU(A.get((size_t)G * size * size * 8)); CU(xs.get((size_t)G * 8)); CU(out.get((size_t)G * 8));
    TRY(upload(s, sp.data(), xs.p, (size_t)G * 8));
    launch_grid_logdets(size, s.dd.rp, s.dd.ci, s.dd.va, G, xs.as<double>(), A.as<double>(), (double)reps, out.as<double>(), s.stream);
    h->launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(sp.data() + G, out.p, (size_t)G * 8, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
  }
  fmm_spline_coefficients(G, sp.data(), sp.data() + G, sp.data() + 2 * G, sp.data() + 3 * G, sp.data() + 4 * G);
  for (Shard &s : h->shards) {
    CU(cudaSetDevice(s.dev));
    double *d;
    TRY(s.alloc(&d, (size_t)5 * G));
    TRY(upload(s, sp.data(), d, (size_t)5 * G * 8));
    CU(cudaStreamSynchronize(s.stream));
    s.dd.spl_n = G; s.dd.spl = d;
    s.dd.n_nodes = 0; s.dd.node_re = nullptr; s.dd.node_im = nullptr; s.dd.node_w = nullptr;
  }
  return BSREG_OK;
}

// Sampled SLX connectivity parameter: per-chain state (delta, proposal scale, counters) and per-chain Gram matrices, all
// starting from the shared Gram matrix at delta0 (the design was built at Psi(delta0), setup_shard)
int init_delta(bsreg_handle *h, const bsreg_options *o) {
  for (Shard &s : h->shards) {
    if (s.dl.L == 0) continue;
    CU(cudaSetDevice(s.dev));
    const int C = s.cs.n_chains, d = h->d, L = s.dl.L;
    TRY(s.alloc(&s.dl.delta, (size_t)C)); TRY(s.alloc(&s.dl.delta_prop, (size_t)C)); TRY(s.alloc(&s.dl.scale, (size_t)C));
    TRY(s.alloc(&s.dl.cnt, (size_t)C * 4)); TRY(s.alloc(&s.dl.Gchain, (size_t)C * d * d));
    TRY(s.alloc(&s.dl_partial, (size_t)delta_sweep_ctas(s.sm_count) * C * L * (1 + h->m)));
    std::vector<double> v((size_t)C, o->slx_delta), sc((size_t)C, o->slx_delta_scale);
    TRY(upload(s, v.data(), s.dl.delta, (size_t)C * 8)); TRY(upload(s, v.data(), s.dl.delta_prop, (size_t)C * 8));
    TRY(upload(s, sc.data(), s.dl.scale, (size_t)C * 8));
    CU(cudaMemsetAsync(s.dl.cnt, 0, (size_t)C * 4 * 8, s.stream));
    s.dd.gcur = (int)((s.seq + 2) % 3);
    launch_gram_to_double(s.dd, s.dl.Gchain, s.stream);
    for (int c = 1; c < C; ++c)
      CU(cudaMemcpyAsync(s.dl.Gchain + (size_t)c * d * d, s.dl.Gchain, (size_t)d * d * 8, cudaMemcpyDeviceToDevice, s.stream));
    ++h->launches;
    CU(cudaStreamSynchronize(s.stream));                  // host temporaries
    s.dd.Gchain = s.dl.Gchain;
  }
  return BSREG_OK;
}

// One Gibbs iteration with a sampled delta: proposals -> per-chain sweep (the lags at every chain's proposal) ->
// sigma^2, beta, shrinkage (chain kernel on the chain's own Gram matrix) -> MH decision, tuning, draw store
int enqueue_iterations_delta(bsreg_handle *h, Shard &s, int count) {
  const int m = h->m, L = s.dl.L, parts = delta_sweep_ctas(s.sm_count);
  for (int i = 0; i < count; ++i) {
    if (s.dl.sampled) {
      launch_delta_propose(s.dl, s.cs, s.stream);
      launch_sweep_delta(s.dl, s.dd.X, s.dd.y, h->n, m, m - L, s.cs.n_chains, s.dl.delta_prop, s.dl_partial, s.sm_count, s.stream);
      h->launches += 2;
    }
    enqueue_chain_step(h, s, s.stream);
    launch_delta_decide(s.dd, s.dl, h->pr, s.cs, s.ctl_dev, s.draws_dev, s.dl_partial, parts, s.dl.sampled ? 0 : 2, s.stream);
    ++h->launches;
  }
  CU(cudaGetLastError());
  return BSREG_OK;
}

int init_chains(bsreg_handle *h, bool refresh_only) {
  const bsreg_options &o = h->opt;
  for (Shard &s : h->shards) {
    CU(cudaSetDevice(s.dev));
    ChainInit ci{s.beta0, o.ng_sigma0, o.sp_lambda, o.sp_lambda_scale, o.hs_lambda, o.hs_tau, o.hs_zeta, o.hs_nu,
                 o.sng_lambda, o.sng_tau, o.sng_theta, o.sng_theta_scale, s.p0_init, refresh_only};
    s.dd.gcur = (int)((s.seq + 2) % 3);
    launch_chain_init(s.dd, h->pr, s.cs, ci, s.stream);
    ++h->launches;
    CU(cudaGetLastError());
  }
  for (Shard &s : h->shards) {
    CU(cudaSetDevice(s.dev));
    CU(cudaStreamSynchronize(s.stream));
  }
  return BSREG_OK;
}

}  // namespace

extern "C" {

int bsreg_abi_version(void) { return BSREG_ABI_VERSION; }
const char *bsreg_last_error(void) { return g_err.c_str(); }

int bsreg_create(const bsreg_problem *p, const bsreg_options *o, bsreg_handle **out) {
  if (!p || !o || !out) return fail(BSREG_EINVAL, "null argument");
  *out = nullptr;
  if (p->n <= 0 || p->k <= 0 || !p->y || !p->X) return fail(BSREG_EINVAL, "need n > 0, k > 0, y and X");
  if (o->model < 0 || o->model > 2 || o->prior < 0 || o->prior > 3) return fail(BSREG_EINVAL, "bad model or prior code");
  if (o->n_chains <= 0) return fail(BSREG_EINVAL, "n_chains must be positive");
  const bool need_w = o->model != BSREG_MODEL_LM || p->k_slx > 0;
  if (need_w && (!p->row_ptr || !p->col_idx || !p->val || p->nnz <= 0))
    return fail(BSREG_EINVAL, "Please provide a connectivity matrix (CSR row_ptr/col_idx/val)");   // R/15_sar.R:52
  SharedOrder order;
  long double trace_w = 0.0L;
  if (need_w) {
    if (p->row_ptr[0] != 0 || p->row_ptr[p->n] != p->nnz) return fail(BSREG_EINVAL, "CSR row_ptr inconsistent with nnz");
    for (int i = 0; i < p->n; ++i)
      if (p->row_ptr[i + 1] < p->row_ptr[i]) return fail(BSREG_EINVAL, "CSR row_ptr must be non-decreasing");
    // the internal ordering of the observations (breadth-first relabelling when W has no locality: ~3 ms of host work at
    // N = 1e5) is computed by a helper thread beside the validation pass, the uploads and the trace moments
    const int d_gram = (o->model == BSREG_MODEL_SEM ? 2 : 1) * (p->k + p->k_slx) + 2;
    order.start(o->model != BSREG_MODEL_LM && d_gram <= 88 && p->nnz > 0, p->n, p->row_ptr, p->col_idx);
    // one pass: index range, max_i sum_j |W_ij| for the Chebyshev log-determinant below, and tr(W)
    double row_norm = 0.0;
    for (int i = 0; i < p->n; ++i) {
      double rs = 0.0;
      for (int64_t q = p->row_ptr[i]; q < p->row_ptr[i + 1]; ++q) {
        if (p->col_idx[q] < 0 || p->col_idx[q] >= p->n) return fail(BSREG_EINVAL, "CSR column index out of range");
        rs += std::fabs(p->val[q]);
        if (p->col_idx[q] == i) trace_w += p->val[q];
      }
      row_norm = std::max(row_norm, rs);
    }
    if (o->model != BSREG_MODEL_LM && o->ldet == BSREG_LDET_CHEBYSHEV && row_norm > 1.0 + 1e-8) {
      // T_j(W) is bounded only if the spectrum of W lies in [-1, 1]; min(||W||_inf, ||W||_1) bounds the spectral radius.
      // (The reference's eigenvalue method, R/15_sar.R:91-96, has no such condition: pass BSREG_LDET_NODES instead.)
      std::vector<double> cs((size_t)p->n, 0.0);
      for (int64_t q = 0; q < p->nnz; ++q) cs[p->col_idx[q]] += std::fabs(p->val[q]);
      const double col_norm = *std::max_element(cs.begin(), cs.end());
      if (col_norm > 1.0 + 1e-8)
        return fail(BSREG_EINVAL, "BSREG_LDET_CHEBYSHEV needs the spectrum of W inside [-1, 1], but ||W||_inf = %g and "
                                  "||W||_1 = %g: normalise W (e.g. row-standardise) or pass eigenvalue nodes", row_norm, col_norm);
    }
  }
  if (p->k_slx > 0 && !p->X_slx) return fail(BSREG_EINVAL, "k_slx > 0 needs X_slx");
  Nvtx range_create("bsreg_create");
  PhaseTimer pt0;
  int ndev_avail = 0;
  if (cudaGetDeviceCount(&ndev_avail) != cudaSuccess || ndev_avail == 0) {
    cudaGetLastError();
    return fail(BSREG_ENOGPU, "no CUDA device available (this library has no CPU fallback)");
  }

  pt0.mark("argument validation");
  PhaseTimer pt;
  bsreg_handle *h = new bsreg_handle();
  h->n = p->n; h->k = p->k; h->m = p->k + p->k_slx; h->model = o->model; h->prior = o->prior;
  h->ldet = o->ldet; h->stats_mode = o->stats_mode; h->n_chains = o->n_chains; h->opt = *o; h->nnz = p->nnz;
  h->spatial = o->model != BSREG_MODEL_LM;
  h->d = (o->model == BSREG_MODEL_SEM) ? 2 * h->m + 2 : h->m + 2;
  h->P = h->m + 1 + (h->spatial ? 1 : 0);
  if (h->n <= h->m) { delete h; return fail(BSREG_EINVAL, "need n > number of regressors"); }
  if (chain_smem_bytes(h->m, h->model, h->prior) > 227 * 1024) {
    delete h;
    return fail(BSREG_EINVAL, "too many regressors for the on-chip chain step (M = %d)", h->m);
  }
  if (h->spatial && o->ldet == BSREG_LDET_NONE) { delete h; return fail(BSREG_EINVAL, "spatial model needs a log-determinant method"); }
  if (h->spatial && o->ldet == BSREG_LDET_NODES && (p->n_nodes <= 0 || !p->node_re || !p->node_w)) {
    delete h; return fail(BSREG_EINVAL, "BSREG_LDET_NODES needs node_re / node_w");
  }

  h->slx_psi = o->slx_psi != 0;
  if (h->slx_psi) {
    // connectivity function Psi_SLX(delta) = dist ^ -delta (R/12_slx.R:69-74): SLX model, a handful of lagged columns
    const int64_t nz = p->nnz_slx > 0 ? p->nnz_slx : p->nnz;
    const double *dist = p->nnz_slx > 0 ? p->val_slx : p->val;
    bool ok = true;
    for (int64_t q = 0; q < nz && ok; ++q) ok = dist[q] > 0.0 && std::isfinite(dist[q]);
    const char *why = o->model != BSREG_MODEL_LM ? "sampling 'delta' is built for the SLX model (SDM / SDEM: fixed connectivity only)"
                    : p->k_slx <= 0 || p->k_slx > 8 ? "a connectivity function needs 1 to 8 columns in X_SLX"
                    : h->m + 2 > 48 ? "too many regressors for a connectivity function (M + 2 <= 48)"
                    : !ok ? "the values of the Psi_SLX pattern must be positive distances"
                    : !(o->slx_delta > o->slx_delta_min && o->slx_delta < o->slx_delta_max)
                        ? "Please provide a valid starting value for delta (spatial) via 'delta'."      // R/41_set_options.R:140
                    : nullptr;
    if (why) { delete h; return fail(BSREG_EINVAL, "%s", why); }
    h->P = h->m + 2;                                              // c(beta, sigma, delta_SLX), R/12_slx.R:117-121
  }
  if (h->spatial && o->ldet == BSREG_LDET_SPLINE) {
    const int reps = o->ldet_reps > 0 ? o->ldet_reps : 1;
    if (o->ldet_grid_len < 2 || !(o->ldet_grid_to > o->ldet_grid_from)) {
      delete h; return fail(BSREG_EINVAL, "BSREG_LDET_SPLINE needs i_lambda = c(from, to, length) with to > from and length >= 2");
    }
    if (p->n % reps != 0) { delete h; return fail(BSREG_EINVAL, "'reps' must divide the number of observations"); }
    if (p->n / reps > 4096) {
      delete h;
      return fail(BSREG_EINVAL, "BSREG_LDET_SPLINE factorises dense %d x %d matrices (limit 4096): use the eigenvalue or "
                                "Chebyshev log-determinant at this size", p->n / reps, p->n / reps);
    }
  }

  Priors &pr = h->pr;
  pr.prior = o->prior;
  pr.shape0 = o->ng_shape; pr.rate0 = o->ng_rate; pr.shape1 = o->ng_shape + 0.5 * (double)p->n;   // R/10_lm.R:152
  pr.sng_lambda_a = o->sng_lambda_a; pr.sng_lambda_b = o->sng_lambda_b;
  pr.sng_theta_scale = o->sng_theta_scale; pr.sng_theta_a = o->sng_theta_a;
  pr.lam_a = o->sp_lambda_a; pr.lam_b = o->sp_lambda_b; pr.lam_lo = o->sp_lambda_min; pr.lam_hi = o->sp_lambda_max;
  pr.lbeta_ab = std::lgamma(pr.lam_a) + std::lgamma(pr.lam_b) - std::lgamma(pr.lam_a + pr.lam_b);
  // the copied options must not keep dangling host pointers
  h->opt.ng_mu = nullptr; h->opt.ng_precision = nullptr; h->opt.ng_beta0 = nullptr; h->opt.devices = nullptr; h->opt.stream = nullptr;

  int cur = 0;
  cudaGetDevice(&cur);
  const int ndev = o->n_devices > 0 ? o->n_devices : 1;
  if (ndev > o->n_chains) { delete h; return fail(BSREG_EINVAL, "more devices than chains"); }
  h->shards.resize(ndev);
  int rc = BSREG_OK, lo = 0;
  MomentsJob mjob;
  h->have_trace_w = need_w; h->trace_w = (double)trace_w;
  const bool want_moments = h->spatial && o->ldet == BSREG_LDET_CHEBYSHEV;
  std::vector<int> cnts(ndev), los(ndev);
  for (int g = 0; g < ndev && rc == BSREG_OK; ++g) {
    Shard &s = h->shards[g];
    s.dev = o->n_devices > 0 && o->devices ? o->devices[g] : cur;
    if (s.dev < 0 || s.dev >= ndev_avail) { rc = fail(BSREG_EINVAL, "device %d not available", s.dev); break; }
    cnts[g] = o->n_chains / ndev + (g < o->n_chains % ndev ? 1 : 0);         // contiguous blocks of chain ids
    los[g] = lo;
    lo += cnts[g];
  }
  if (rc == BSREG_OK) {
    if (ndev == 1) {
      rc = setup_shard(h, h->shards[0], p, o, cnts[0], 0, want_moments ? &mjob : nullptr, order);
    } else {
      // one host thread per device: uploads, packing and the first sweep of all shards run side by side (the current
      // device and the error string are per thread; the first failure is reported)
      std::vector<int> rcs(ndev, BSREG_OK);
      std::vector<std::string> errs(ndev);
      std::vector<std::thread> workers;
      for (int g = 0; g < ndev; ++g)
        workers.emplace_back([&, g] {
          rcs[g] = setup_shard(h, h->shards[g], p, o, cnts[g], los[g], g == 0 && want_moments ? &mjob : nullptr, order);
          if (rcs[g] != BSREG_OK) errs[g] = g_err;
        });
      for (std::thread &t : workers) t.join();
      for (int g = 0; g < ndev && rc == BSREG_OK; ++g)
        if (rcs[g] != BSREG_OK) { rc = rcs[g]; g_err = errs[g]; }
    }
  }
  pt.mark("shards total");
  pt.begin("bsreg_create: log-determinant");
  // ---- log-determinant nodes ----
  if (rc == BSREG_OK && h->spatial) {
    if (o->ldet == BSREG_LDET_SPLINE) {
    } else if (o->ldet == BSREG_LDET_NODES) {
      h->node_re.assign(p->node_re, p->node_re + p->n_nodes);
      if (p->node_im) h->node_im.assign(p->node_im, p->node_im + p->n_nodes); else h->node_im.assign(p->n_nodes, 0.0);
      h->node_w.assign(p->node_w, p->node_w + p->n_nodes);
    } else {
      const int order = o->cheb_order > 0 ? o->cheb_order : 50, probes = o->cheb_probes > 0 ? o->cheb_probes : 64;
      std::vector<double> mom(order + 1);
      if (mjob.active) rc = trace_moments_finish(h, h->shards[0], mjob, mom.data());
      else rc = trace_moments_dev(h, h->shards[0], p, order, probes, o->seed, mom.data());
      if (rc == BSREG_OK) chebyshev_nodes(mom, h->n, h->node_re, h->node_im, h->node_w);
    }
    if (rc == BSREG_OK && o->ldet == BSREG_LDET_SPLINE) rc = build_spline(h, o);
    else if (rc == BSREG_OK) rc = set_nodes(h);
  }
  if (mjob.active) { std::vector<double> drop(mjob.order + 1); trace_moments_finish(h, h->shards[0], mjob, drop.data()); }   // error paths
  pt.mark("log-det nodes");
  pt.begin("bsreg_create: chain init");
  if (rc == BSREG_OK) rc = init_chains(h, false);
  if (rc == BSREG_OK && h->slx_psi) rc = init_delta(h, o);
  pt.mark("chain init");
  cudaSetDevice(cur);
  if (rc != BSREG_OK) { std::string keep = g_err; bsreg_destroy(h); g_err = keep; return rc; }
  *out = h;
  return BSREG_OK;
}

// Host-side ordering logic of bsreg_create, exposed for the CPU test-suite (no device needed): returns 1 and fills
// perm[n] (internal row r = caller's observation perm[r]) if W would be relabelled, 0 if it is local enough already.
int bsreg_debug_internal_order(int32_t n, const int32_t *row_ptr, const int32_t *col_idx, int32_t force, int32_t *perm) {
  if (n <= 0 || !row_ptr || !col_idx || !perm) return fail(BSREG_EINVAL, "bsreg_debug_internal_order: bad arguments");
  if (!force && !wants_reorder(n, row_ptr, col_idx)) return 0;
  std::vector<int> pv;
  bfs_order(n, row_ptr, col_idx, pv);
  for (int i = 0; i < n; ++i) perm[i] = pv[i];
  return 1;
}

void bsreg_destroy(bsreg_handle *h) {
  if (!h) return;
  int cur = 0;
  cudaGetDevice(&cur);
  for (Shard &s : h->shards) {
    cudaSetDevice(s.dev);
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.aux) cudaStreamSynchronize(s.aux);
    if (s.graph) cudaGraphExecDestroy(s.graph);
    if (s.sweep_graph) cudaGraphExecDestroy(s.sweep_graph);
    if (s.chain_graph) cudaGraphExecDestroy(s.chain_graph);
    for (cudaEvent_t e : s.events) cudaEventDestroy(e);
    if (s.aux) cudaStreamDestroy(s.aux);
    for (void *q : s.allocs) pool_free(q);
    if (s.draws_dev) pool_free(s.draws_dev);
    if (s.own_stream && s.stream) cudaStreamDestroy(s.stream);
  }
  cudaSetDevice(cur);
  delete h;
}

int bsreg_dims(const bsreg_handle *h, int32_t *n, int32_t *m, int32_t *p, int32_t *n_chains, int32_t *d) {
  if (!h) return fail(BSREG_EINVAL, "null handle");
  if (n) *n = h->n;
  if (m) *m = h->m;
  if (p) *p = h->P;
  if (n_chains) *n_chains = h->n_chains;
  if (d) *d = h->d;
  return BSREG_OK;
}

int bsreg_run(bsreg_handle *h, int32_t n_burn, int32_t n_save, const bsreg_mh *mh, double *draws,
              bsreg_progress_fn progress, void *user) {
  if (!h) return fail(BSREG_EINVAL, "null handle");
  if (n_burn < 0 || n_save < 0) return fail(BSREG_EINVAL, "n_burn and n_save must be non-negative");
  Nvtx range_run("bsreg_run");
  const bsreg_mh def{0.8, 0.20, 0.45, 0.8};
  if (!mh) mh = &def;
  int cur = 0;
  cudaGetDevice(&cur);
  const int64_t total = (int64_t)n_burn + n_save;
  // draw buffers
  for (Shard &s : h->shards) {
    CU(cudaSetDevice(s.dev));
    const size_t row_bytes = (size_t)h->P * s.cs.n_chains * 8;
    int cap = (int)std::min<size_t>((size_t)std::max(n_save, 1), std::max<size_t>(DRAW_BUF_BYTES / row_bytes, 1));
    if (cap > s.save_cap) {
      CU(cudaStreamSynchronize(s.stream));
      if (s.graph) { cudaGraphExecDestroy(s.graph); s.graph = nullptr; }   // the graphs bake the buffer pointer
      if (s.chain_graph) { cudaGraphExecDestroy(s.chain_graph); s.chain_graph = nullptr; }
      if (s.draws_dev) pool_free(s.draws_dev);
      s.draws_dev = nullptr;
      CU(pool_malloc((void **)&s.draws_dev, (size_t)cap * row_bytes));
      s.save_cap = cap;
    }
    CU(cudaMemsetAsync(s.cs.step, 0, (size_t)s.cs.n_chains * 4, s.stream));
  }
  RunCtl ctl{};
  ctl.n_burn = n_burn; ctl.n_save = n_save;
  ctl.tune_limit = (int)std::floor(mh->adjust_burn * (double)n_burn);   // i <= adjust_burn * n_burn, R/50_execute.R:62
  ctl.acc_lo = mh->acc_lo; ctl.acc_hi = mh->acc_hi; ctl.acc_change = mh->acc_change;
  ctl.store = 1;
  int64_t done = 0;
  int rc = BSREG_OK;
  while (done < total && rc == BSREG_OK) {
    int nb;
    const bool saving = done >= n_burn;
    Nvtx range_block(saving ? "bsreg_run: sampling block" : "bsreg_run: burn-in block");
    if (!saving) nb = (int)std::min<int64_t>(BLOCK_ITERS, n_burn - done);
    else nb = (int)std::min<int64_t>(BLOCK_ITERS, total - done);
    for (Shard &s : h->shards) {
      CU(cudaSetDevice(s.dev));
      int nbs = nb;
      if (saving) nbs = std::min(nbs, s.save_cap);
      nb = std::min(nb, nbs);
    }
    for (Shard &s : h->shards) {
      CU(cudaSetDevice(s.dev));
      ctl.save_cap = s.save_cap;
      ctl.save_base = saving ? (int)(done - n_burn) : 0;
      CU(cudaMemcpyAsync(s.ctl_dev, &ctl, sizeof ctl, cudaMemcpyHostToDevice, s.stream));
      TRY(enqueue_iterations(h, s, nb, ctl));
      if (saving && draws) {
        double *dst = draws + (size_t)s.chain_lo * h->P * n_save + ctl.save_base;
        CU(cudaMemcpy2DAsync(dst, (size_t)n_save * 8, s.draws_dev, (size_t)s.save_cap * 8, (size_t)nb * 8,
                             (size_t)h->P * s.cs.n_chains, cudaMemcpyDeviceToHost, s.stream));
      }
    }
    for (Shard &s : h->shards) {
      CU(cudaSetDevice(s.dev));
      CU(cudaStreamSynchronize(s.stream));
    }
    done += nb;
    if (progress && progress(done, total, user) != 0) rc = fail(BSREG_EABORT, "aborted by progress callback");
  }
  cudaSetDevice(cur);
  return rc;
}

int bsreg_fetch_draws(bsreg_handle *h, int32_t first, int32_t count, double *draws) {
  if (!h || !draws) return fail(BSREG_EINVAL, "null argument");
  int cur = 0;
  cudaGetDevice(&cur);
  for (Shard &s : h->shards) {
    if (first < 0 || count < 0 || first + count > s.save_cap) return fail(BSREG_EINVAL, "rows outside the device draw buffer");
    CU(cudaSetDevice(s.dev));
    double *dst = draws + (size_t)s.chain_lo * h->P * count;
    CU(cudaMemcpy2DAsync(dst, (size_t)count * 8, s.draws_dev + first, (size_t)s.save_cap * 8, (size_t)count * 8,
                         (size_t)h->P * s.cs.n_chains, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
  }
  cudaSetDevice(cur);
  return BSREG_OK;
}

int bsreg_aux_len(const bsreg_handle *h) { return h ? aux_len(h->m) : 0; }

int bsreg_get_state(bsreg_handle *h, double *params, double *p0, double *aux, double *mh_value, double *mh_scale,
                    int64_t *mh_counters, int64_t *iterations) {
  if (!h) return fail(BSREG_EINVAL, "null handle");
  int cur = 0;
  cudaGetDevice(&cur);
  const int m = h->m, P = h->P, A = aux_len(m);
  for (Shard &s : h->shards) {
    CU(cudaSetDevice(s.dev));
    const int C = s.cs.n_chains, lo = s.chain_lo;
    std::vector<double> beta((size_t)C * m), s2(C), lam(C), ax((size_t)C * A), sc((size_t)C * 2);
    CU(cudaMemcpy(beta.data(), s.cs.beta, beta.size() * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(s2.data(), s.cs.sigma2, (size_t)C * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(lam.data(), s.cs.lam, (size_t)C * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ax.data(), s.cs.aux, ax.size() * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(sc.data(), s.cs.mh_scale, sc.size() * 8, cudaMemcpyDeviceToHost));
    std::vector<double> dl_delta(C, 0.0);
    if (h->slx_psi) CU(cudaMemcpy(dl_delta.data(), s.dl.delta, (size_t)C * 8, cudaMemcpyDeviceToHost));
    for (int c = 0; c < C; ++c) {
      if (params) {
        double *dst = params + (size_t)(lo + c) * P;
        std::copy(beta.begin() + (size_t)c * m, beta.begin() + (size_t)(c + 1) * m, dst);
        dst[m] = s2[c];
        if (h->spatial) dst[m + 1] = lam[c];
        if (h->slx_psi) dst[m + 1] = dl_delta[c];
      }
      if (aux) std::copy(ax.begin() + (size_t)c * A, ax.begin() + (size_t)(c + 1) * A, aux + (size_t)(lo + c) * A);
      if (mh_value) { mh_value[2 * (lo + c)] = lam[c]; mh_value[2 * (lo + c) + 1] = ax[(size_t)c * A + 2 * m + 1]; }
      if (mh_scale) { mh_scale[2 * (lo + c)] = sc[2 * c]; mh_scale[2 * (lo + c) + 1] = sc[2 * c + 1]; }
    }
    if (p0) CU(cudaMemcpy(p0 + (size_t)lo * m, s.cs.p0, (size_t)C * m * 8, cudaMemcpyDeviceToHost));
    if (mh_counters) CU(cudaMemcpy(mh_counters + (size_t)lo * 8, s.cs.mh_cnt, (size_t)C * 8 * 8, cudaMemcpyDeviceToHost));
    if (iterations) CU(cudaMemcpy(iterations + lo, s.cs.iters, (size_t)C * 8, cudaMemcpyDeviceToHost));
  }
  cudaSetDevice(cur);
  return BSREG_OK;
}

int bsreg_set_state(bsreg_handle *h, const double *params, const double *p0, const double *aux, const double *mh_value,
                    const double *mh_scale, const int64_t *mh_counters, const int64_t *iterations) {
  if (!h) return fail(BSREG_EINVAL, "null handle");
  (void)mh_value;   // the MH values are the parameters themselves (lambda in params, theta in aux)
  int cur = 0;
  cudaGetDevice(&cur);
  const int m = h->m, P = h->P, A = aux_len(m);
  for (Shard &s : h->shards) {
    CU(cudaSetDevice(s.dev));
    CU(cudaStreamSynchronize(s.stream));
    const int C = s.cs.n_chains, lo = s.chain_lo;
    if (params) {
      std::vector<double> beta((size_t)C * m), s2(C), lam(C, 0.0);
      for (int c = 0; c < C; ++c) {
        const double *src = params + (size_t)(lo + c) * P;
        std::copy(src, src + m, beta.begin() + (size_t)c * m);
        s2[c] = src[m];
        if (h->spatial) lam[c] = src[m + 1];
      }
      CU(cudaMemcpy(s.cs.beta, beta.data(), beta.size() * 8, cudaMemcpyHostToDevice));
      CU(cudaMemcpy(s.cs.sigma2, s2.data(), (size_t)C * 8, cudaMemcpyHostToDevice));
      CU(cudaMemcpy(s.cs.lam, lam.data(), (size_t)C * 8, cudaMemcpyHostToDevice));
    }
    if (p0) CU(cudaMemcpy(s.cs.p0, p0 + (size_t)lo * m, (size_t)C * m * 8, cudaMemcpyHostToDevice));
    if (aux) CU(cudaMemcpy(s.cs.aux, aux + (size_t)lo * A, (size_t)C * A * 8, cudaMemcpyHostToDevice));
    if (mh_scale) CU(cudaMemcpy(s.cs.mh_scale, mh_scale + (size_t)lo * 2, (size_t)C * 2 * 8, cudaMemcpyHostToDevice));
    if (mh_counters) CU(cudaMemcpy(s.cs.mh_cnt, mh_counters + (size_t)lo * 8, (size_t)C * 8 * 8, cudaMemcpyHostToDevice));
    if (iterations) CU(cudaMemcpy(s.cs.iters, iterations + lo, (size_t)C * 8, cudaMemcpyHostToDevice));
  }
  cudaSetDevice(cur);
  return init_chains(h, true);    // recompute the cached log-det / rss at the restored lambda
}

int bsreg_get_delta_state(bsreg_handle *h, double *delta, double *scale, int64_t *counters) {
  if (!h) return fail(BSREG_EINVAL, "null handle");
  if (!h->slx_psi) return fail(BSREG_EINVAL, "the handle has no connectivity function (slx_psi)");
  int cur = 0;
  cudaGetDevice(&cur);
  for (Shard &s : h->shards) {
    CU(cudaSetDevice(s.dev));
    CU(cudaStreamSynchronize(s.stream));
    const int C = s.cs.n_chains, lo = s.chain_lo;
    if (delta) CU(cudaMemcpy(delta + lo, s.dl.delta, (size_t)C * 8, cudaMemcpyDeviceToHost));
    if (scale) CU(cudaMemcpy(scale + lo, s.dl.scale, (size_t)C * 8, cudaMemcpyDeviceToHost));
    if (counters) CU(cudaMemcpy(counters + (size_t)lo * 4, s.dl.cnt, (size_t)C * 4 * 8, cudaMemcpyDeviceToHost));
  }
  cudaSetDevice(cur);
  return BSREG_OK;
}

// Restores value, scale and counters; the chains' Gram matrices are rebuilt at the restored values by one per-chain sweep.
int bsreg_set_delta_state(bsreg_handle *h, const double *delta, const double *scale, const int64_t *counters) {
  if (!h) return fail(BSREG_EINVAL, "null handle");
  if (!h->slx_psi) return fail(BSREG_EINVAL, "the handle has no connectivity function (slx_psi)");
  int cur = 0;
  cudaGetDevice(&cur);
  for (Shard &s : h->shards) {
    CU(cudaSetDevice(s.dev));
    CU(cudaStreamSynchronize(s.stream));
    const int C = s.cs.n_chains, lo = s.chain_lo, L = s.dl.L;
    if (scale) CU(cudaMemcpy(s.dl.scale, scale + lo, (size_t)C * 8, cudaMemcpyHostToDevice));
    if (counters) CU(cudaMemcpy(s.dl.cnt, counters + (size_t)lo * 4, (size_t)C * 4 * 8, cudaMemcpyHostToDevice));
    if (delta) {
      CU(cudaMemcpy(s.dl.delta, delta + lo, (size_t)C * 8, cudaMemcpyHostToDevice));
      launch_sweep_delta(s.dl, s.dd.X, s.dd.y, h->n, h->m, h->m - L, C, s.dl.delta, s.dl_partial, s.sm_count, s.stream);
      launch_delta_decide(s.dd, s.dl, h->pr, s.cs, s.ctl_dev, s.draws_dev, s.dl_partial, delta_sweep_ctas(s.sm_count), 1, s.stream);
      h->launches += 2;
      CU(cudaGetLastError());
      CU(cudaStreamSynchronize(s.stream));
    }
  }
  cudaSetDevice(cur);
  return BSREG_OK;
}

// ---- parity entry points (shard 0) ----
int bsreg_eval_spmm(bsreg_handle *h, const double *B, int32_t mcols, double *out) {
  if (!h || !B || !out || mcols <= 0) return fail(BSREG_EINVAL, "bad argument");
  Shard &s = h->shards[0];
  if (!s.dd.rp) return fail(BSREG_EINVAL, "handle has no W");
  CU(cudaSetDevice(s.dev));
  const size_t len = (size_t)h->n * mcols;
  double *cm, *rm, *orm;
  CU(pool_malloc((void **)&cm, len * 8)); CU(pool_malloc((void **)&rm, len * 8)); CU(pool_malloc((void **)&orm, len * 8));
  CU(cudaMemcpyAsync(cm, B, len * 8, cudaMemcpyHostToDevice, s.stream));
  launch_transpose(cm, rm, h->n, mcols, mcols, 0, s.stream);
  launch_spmm(h->n, s.dd.rp, s.dd.ci, s.dd.va, rm, mcols, mcols, orm, mcols, 0, 1.0, nullptr, 0.0, s.stream);
  h->launches += 2;
  std::vector<double> host(len);
  CU(cudaMemcpyAsync(host.data(), orm, len * 8, cudaMemcpyDeviceToHost, s.stream));
  CU(cudaStreamSynchronize(s.stream));
  for (int i = 0; i < h->n; ++i)
    for (int c = 0; c < mcols; ++c) out[(size_t)c * h->n + i] = host[(size_t)i * mcols + c];
  pool_free(cm); pool_free(rm); pool_free(orm);
  return BSREG_OK;
}

int bsreg_eval_gram(bsreg_handle *h, double *G) {
  if (!h || !G) return fail(BSREG_EINVAL, "bad argument");
  Shard &s = h->shards[0];
  CU(cudaSetDevice(s.dev));
  enqueue_sweep(h, s, s.stream);
  double *dG;
  CU(pool_malloc((void **)&dG, (size_t)h->d * h->d * 8));
  s.dd.gcur = (int)((s.seq + 2) % 3);
  launch_gram_to_double(s.dd, dG, s.stream);
  ++h->launches;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(G, dG, (size_t)h->d * h->d * 8, cudaMemcpyDeviceToHost, s.stream));
  CU(cudaStreamSynchronize(s.stream));
  pool_free(dG);
  return BSREG_OK;
}

int bsreg_eval_residual_ss(bsreg_handle *h, int32_t n_eval, const double *beta, const double *lam, const double *rho,
                           double *ss) {
  if (!h || !beta || !lam || !rho || !ss || n_eval <= 0) return fail(BSREG_EINVAL, "bad argument");
  Shard &s = h->shards[0];
  CU(cudaSetDevice(s.dev));
  double *db, *dl, *dr, *dout, *part;
  CU(pool_malloc((void **)&db, (size_t)n_eval * h->m * 8)); CU(pool_malloc((void **)&dl, (size_t)n_eval * 8)); CU(pool_malloc((void **)&dr, (size_t)n_eval * 8));
  CU(pool_malloc((void **)&dout, (size_t)n_eval * 8)); CU(pool_malloc((void **)&part, (size_t)592 * 4 * 8));
  CU(cudaMemcpyAsync(db, beta, (size_t)n_eval * h->m * 8, cudaMemcpyHostToDevice, s.stream));
  CU(cudaMemcpyAsync(dl, lam, (size_t)n_eval * 8, cudaMemcpyHostToDevice, s.stream));
  CU(cudaMemcpyAsync(dr, rho, (size_t)n_eval * 8, cudaMemcpyHostToDevice, s.stream));
  launch_residual_ss(s.dd, n_eval, db, dl, dr, part, dout, s.stream);
  h->launches += 2 * ((n_eval + 3) / 4);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ss, dout, (size_t)n_eval * 8, cudaMemcpyDeviceToHost, s.stream));
  CU(cudaStreamSynchronize(s.stream));
  pool_free(db); pool_free(dl); pool_free(dr); pool_free(dout); pool_free(part);
  return BSREG_OK;
}

int bsreg_eval_logdet(bsreg_handle *h, int32_t n_eval, const double *v, double *ldet) {
  if (!h || !v || !ldet || n_eval <= 0) return fail(BSREG_EINVAL, "bad argument");
  Shard &s = h->shards[0];
  if (s.dd.n_nodes <= 0 && s.dd.spl_n <= 0) return fail(BSREG_EINVAL, "handle has no log-determinant nodes");
  CU(cudaSetDevice(s.dev));
  double *dv, *dout;
  CU(pool_malloc((void **)&dv, (size_t)n_eval * 8)); CU(pool_malloc((void **)&dout, (size_t)n_eval * 8));
  CU(cudaMemcpyAsync(dv, v, (size_t)n_eval * 8, cudaMemcpyHostToDevice, s.stream));
  launch_eval_logdet(s.dd, n_eval, dv, dout, s.stream);
  ++h->launches;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(ldet, dout, (size_t)n_eval * 8, cudaMemcpyDeviceToHost, s.stream));
  CU(cudaStreamSynchronize(s.stream));
  pool_free(dv); pool_free(dout);
  return BSREG_OK;
}

int bsreg_n_nodes(const bsreg_handle *h) { return h ? (int)h->node_re.size() : 0; }

int bsreg_n_spline(const bsreg_handle *h) { return h ? (int)(h->spline.size() / 5) : 0; }

int bsreg_get_spline(bsreg_handle *h, double *x, double *y, double *b, double *c, double *d) {
  if (!h) return fail(BSREG_EINVAL, "null handle");
  const size_t G = h->spline.size() / 5;
  double *dst[5] = {x, y, b, c, d};
  for (int k = 0; k < 5; ++k)
    if (dst[k]) std::copy(h->spline.begin() + k * G, h->spline.begin() + (k + 1) * G, dst[k]);
  return BSREG_OK;
}

int bsreg_get_nodes(bsreg_handle *h, double *re, double *im, double *w) {
  if (!h) return fail(BSREG_EINVAL, "null handle");
  if (re) std::copy(h->node_re.begin(), h->node_re.end(), re);
  if (im) std::copy(h->node_im.begin(), h->node_im.end(), im);
  if (w) std::copy(h->node_w.begin(), h->node_w.end(), w);
  return BSREG_OK;
}

int bsreg_eval_trace_moments(bsreg_handle *h, int32_t order, int32_t n_probes, uint64_t seed, double *moments) {
  if (!h || !moments || order < 0 || n_probes <= 0) return fail(BSREG_EINVAL, "bad argument");
  Shard &s = h->shards[0];
  if (!s.dd.rp) return fail(BSREG_EINVAL, "handle has no W");
  CU(cudaSetDevice(s.dev));
  // host copy of the CSR for the exact low-order traces
  std::vector<int32_t> rp((size_t)h->n + 1), ci((size_t)h->nnz);
  std::vector<double> va((size_t)h->nnz);
  CU(cudaMemcpy(rp.data(), s.dd.rp, rp.size() * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(ci.data(), s.dd.ci, ci.size() * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(va.data(), s.dd.va, va.size() * 8, cudaMemcpyDeviceToHost));
  bsreg_problem p{};
  p.row_ptr = rp.data(); p.col_idx = ci.data(); p.val = va.data();
  return trace_moments_dev(h, s, &p, order, n_probes, seed, moments);
}

int bsreg_eval_conditionals(bsreg_handle *h, int32_t n_eval, const double *beta, const double *sigma2, const double *lam,
                            const double *p0, const double *v, double *mu1, double *rate1, double *rss, double *logpost) {
  if (!h || n_eval <= 0 || !beta || !sigma2 || !lam || !p0 || !v) return fail(BSREG_EINVAL, "bad argument");
  Shard &s = h->shards[0];
  CU(cudaSetDevice(s.dev));
  const size_t ne = n_eval, m = h->m;
  double *buf;
  CU(pool_malloc((void **)&buf, (ne * m * 3 + ne * 6) * 8));
  double *db = buf, *dp0 = db + ne * m, *dmu1 = dp0 + ne * m, *ds2 = dmu1 + ne * m, *dl = ds2 + ne, *dv = dl + ne,
         *drate = dv + ne, *drss = drate + ne, *dlp = drss + ne;
  CU(cudaMemcpyAsync(db, beta, ne * m * 8, cudaMemcpyHostToDevice, s.stream));
  CU(cudaMemcpyAsync(dp0, p0, ne * m * 8, cudaMemcpyHostToDevice, s.stream));
  CU(cudaMemcpyAsync(ds2, sigma2, ne * 8, cudaMemcpyHostToDevice, s.stream));
  CU(cudaMemcpyAsync(dl, lam, ne * 8, cudaMemcpyHostToDevice, s.stream));
  CU(cudaMemcpyAsync(dv, v, ne * 8, cudaMemcpyHostToDevice, s.stream));
  s.dd.gcur = (int)((s.seq + 2) % 3);
  launch_eval_conditionals(s.dd, h->pr, s.cs.mu0, n_eval, db, ds2, dl, dp0, dv, dmu1, drate, drss, dlp, s.stream);
  ++h->launches;
  CU(cudaGetLastError());
  if (mu1) CU(cudaMemcpyAsync(mu1, dmu1, ne * m * 8, cudaMemcpyDeviceToHost, s.stream));
  if (rate1) CU(cudaMemcpyAsync(rate1, drate, ne * 8, cudaMemcpyDeviceToHost, s.stream));
  if (rss) CU(cudaMemcpyAsync(rss, drss, ne * 8, cudaMemcpyDeviceToHost, s.stream));
  if (logpost) CU(cudaMemcpyAsync(logpost, dlp, ne * 8, cudaMemcpyDeviceToHost, s.stream));
  CU(cudaStreamSynchronize(s.stream));
  pool_free(buf);
  return BSREG_OK;
}

int bsreg_eval_rng(int32_t kind, uint64_t seed, int32_t chain, int32_t slot, int32_t it0, int32_t n, const double *params,
                   double *out) {
  if (!out || n <= 0 || kind < 0 || kind > 5) return fail(BSREG_EINVAL, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(BSREG_ENOGPU, "no CUDA device available"); }
  double *dp, *dout;
  CU(pool_malloc((void **)&dp, 4 * 8)); CU(pool_malloc((void **)&dout, (size_t)n * 8));
  double hp[4] = {0, 0, 0, 0};
  const int np = kind == 2 ? 2 : kind == 3 ? 1 : kind == 4 ? 3 : kind == 5 ? 4 : 0;
  for (int i = 0; i < np; ++i) hp[i] = params ? params[i] : 0.0;
  CU(cudaMemcpy(dp, hp, sizeof hp, cudaMemcpyHostToDevice));
  launch_eval_rng(kind, seed, chain, slot, it0, n, dp, dout, 0);
  CU(cudaGetLastError());
  CU(cudaMemcpy(out, dout, (size_t)n * 8, cudaMemcpyDeviceToHost));
  pool_free(dp); pool_free(dout);
  return BSREG_OK;
}

// ---- measurement helpers ----
void *bsreg_stream(bsreg_handle *h) { return h ? (void *)h->shards[0].stream : nullptr; }

int bsreg_launch_sweep(bsreg_handle *h, int32_t reps) {
  if (!h) return fail(BSREG_EINVAL, "null handle");
  Shard &s = h->shards[0];
  CU(cudaSetDevice(s.dev));
  return enqueue_repeated(h, s, reps, true);
}

int bsreg_launch_chain_step(bsreg_handle *h, int32_t reps) {
  if (!h) return fail(BSREG_EINVAL, "null handle");
  Shard &s = h->shards[0];
  CU(cudaSetDevice(s.dev));
  RunCtl ctl{};
  ctl.n_burn = 1 << 30; ctl.tune_limit = 0; ctl.store = 0; ctl.acc_lo = 0.2; ctl.acc_hi = 0.45; ctl.acc_change = 0.8;
  CU(cudaMemcpyAsync(s.ctl_dev, &ctl, sizeof ctl, cudaMemcpyHostToDevice, s.stream));
  return enqueue_repeated(h, s, reps, false);
}

int bsreg_sync(bsreg_handle *h) {
  if (!h) return fail(BSREG_EINVAL, "null handle");
  for (Shard &s : h->shards) { CU(cudaSetDevice(s.dev)); CU(cudaStreamSynchronize(s.stream)); }
  return BSREG_OK;
}

// CSR once (12 nnz + 4 (N+1)) + y and X once (8 N (K+1))
int64_t bsreg_sweep_bytes(const bsreg_handle *h) {
  if (!h) return 0;
  const int64_t w = h->model == BSREG_MODEL_LM ? 0 : 12 * h->nnz + 4 * ((int64_t)h->n + 1);
  return w + 8 * (int64_t)h->n * (h->m + 1);
}

int64_t bsreg_kernel_launches(const bsreg_handle *h) { return h ? h->launches.load() : 0; }

void bsreg_release_cache(void) {
  std::vector<FreeBlock> drop;
  { std::lock_guard<std::mutex> lk(g_pool_mu); drop.swap(g_pool_free); g_pool_cached = 0; }
  int cur = 0;
  cudaGetDevice(&cur);
  for (const FreeBlock &b : drop) { cudaSetDevice(b.dev); cudaFree(b.ptr); }
  std::vector<void *> pins;
  { std::lock_guard<std::mutex> lk(g_pool_mu); pins.swap(g_pinned_free); }
  for (void *q : pins) cudaFreeHost(q);
  cudaSetDevice(cur);
}

}  // extern "C"
