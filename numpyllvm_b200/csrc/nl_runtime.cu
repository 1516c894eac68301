// nl_runtime.cu — device runtime behind the C-ABI: what the reference's scheduler
// (scheduler.cpp) and worker thread (thread.cpp) become on B200.  The
// moodycamel task queue + one pthread worker are replaced by in-order CUDA
// streams per device; the "tasks" are kernel launches, dependencies are stream
// order and events.  NCCL is bound lazily with dlopen so the library loads (and
// the host-only entry points work) on a machine without GPUs or NCCL.
#include <chrono>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and enums only; symbols are resolved with dlsym
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nothing unless a profiler is attached

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>

#include "nl_internal.h"
#include "nl_step.h"

namespace nl {

#define CUDA_TRY(expr)                                                             \
  do {                                                                             \
    cudaError_t e_ = (expr);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      set_error("%s failed: %s", #expr, cudaGetErrorString(e_));                   \
      return NL_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

// scratch arena layout (per device and stream; launches on one stream are ordered)
constexpr size_t kRedTicketOff = 0;            // u32, self-resetting
constexpr size_t kRedPartialsOff = 256;        // 8 B x kMaxGrid
constexpr size_t kMaxGrid = 8192;
constexpr size_t kTileTicketOff = kRedPartialsOff + 8 * kMaxGrid;   // u32, zeroed per compact launch
constexpr size_t kSmRankOff = kTileTicketOff + 256;                 // u32 x 256: CTAs that have started on each SM, zeroed per compact launch
constexpr size_t kTileDescOff = kSmRankOff + 1024;                  // 8 B x n_tiles, zeroed per compact launch

struct Scratch {
  char *base = nullptr;
  size_t bytes = 0;
  // compaction arenas: two of them (ticket + ranks + descriptors each) alternate between consecutive super-tile launches
  int arena = 0;                 // the one the next super-tile launch works in
  bool arena_clean[2] = {false, false};   // ticket, ranks and EVERY descriptor of the arena are zero
  int64_t arena_dirty[2] = {0, 0};        // descriptors its last user left behind (what the next launch zeroes)
};

struct Device {
  int id = -1;
  int sm_count = 0;
  cudaStream_t stream[NL_N_STREAMS] = {nullptr, nullptr, nullptr};
  Scratch scratch[NL_N_STREAMS];
  int64_t *gather = nullptr;        // world_size int64 for the count allgather
  ncclComm_t comm = nullptr;
  // NVLink mailbox of this device's rank and the table of every rank's mailbox as mapped here
  unsigned long long *box = nullptr;          // cudaMalloc (IPC-exportable), kMailboxBytes, zeroed
  unsigned long long **box_table = nullptr;   // device array [world]
  unsigned long long epoch = 0;               // fused reductions launched on this device so far
  unsigned long long *box_host_table[256] = {nullptr};   // host copy of box_table (to post an abort from the host)
  unsigned long long *status = nullptr;       // pinned host word (device-mapped): raised by a kernel that gave up on a peer
  // column blocks above kCacheMin bytes, kept by size class when nl_free() returns them (see nl_malloc)
  std::multimap<size_t, void *> cache_free;
  std::unordered_map<void *, size_t> cache_class;   // every live or cached large block -> its class size
  size_t cache_bytes = 0;
};

constexpr int kMaxWorld = 256;
constexpr size_t kMailboxBytes = 2 * (size_t)kMaxWorld * 16 + 64;   // [2 parities][world][value, epoch] + the abort word (nl_pipeline.cuh: kMailboxAbortIdx)
static unsigned long long g_peer_timeout_ns = 10ull * 1000 * 1000 * 1000;   // bound on a kernel's wait for a peer (nl_tune "peer_timeout_ms")
void set_peer_timeout_ms(int ms) { g_peer_timeout_ns = (unsigned long long)(ms > 0 ? ms : 10000) * 1000ull * 1000ull; }
static bool g_peer_ready = false;
static bool g_logical = false;            // two devices of the runtime share a GPU (NL_LOGICAL_DEVICES)
static bool g_trust_self_clean = true;    // nl_tune("compact_self_clean", 0): clear the compaction arena in front of every launch, as round 1 did
static void *g_ipc_mapped[NL_MAX_DEVICES][kMaxWorld];          // IPC mappings to close at destroy

// NVTX range around every pipeline launch (what PrintPipeline's progress lines were for, debug_printer.cpp:131-160;
// NL_NVTX=0 turns them off)
static bool nvtx_on() {
  static int on = -1;
  if (on < 0) { const char *e = getenv("NL_NVTX"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char *name) : on(nvtx_on()) { if (on) nvtxRangePushA(name); }
  ~NvtxRange() { if (on) nvtxRangePop(); }
};

static Device g_dev[NL_MAX_DEVICES];
static int g_ndev = 0;
static int g_world = 0, g_first_rank = 0;
static std::atomic<uint64_t> g_launches{0};
static LaunchInfo g_last_info = {"", 0, 0, 0};
static std::mutex g_mu;

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static int check_dev(int dev, int stream) {
  if (g_ndev == 0) { set_error("no CUDA device initialised (call nl_init; there is no CPU fallback)"); return NL_ERR_NO_DEVICE; }
  if (dev < 0 || dev >= g_ndev) { set_error("device index %d out of range (0..%d)", dev, g_ndev - 1); return NL_ERR_INVALID; }
  if (stream < 0 || stream >= NL_N_STREAMS) { set_error("stream index %d out of range", stream); return NL_ERR_INVALID; }
  return NL_OK;
}

// a kernel that gave up waiting for a peer's mailbox slot (peer_give_up) left a note: report it once
static int check_peer_status(Device &d) {
  if (!d.status || *(volatile unsigned long long *)d.status == 0) return NL_OK;
  const unsigned long long v = *(volatile unsigned long long *)d.status;
  *(volatile unsigned long long *)d.status = 0;
  set_error("fused multi-GPU finish given up on device %d: epoch %llu, rank %d did not post its slot (its launch failed, it aborted, or it "
            "took longer than the peer timeout); the result of that launch is void — destroy and re-create the communicator",
            d.id, v & 0xffffffffffffull, (int)(v >> 48) - 1);
  return NL_ERR_NCCL;
}

void release_cached_blocks(Device &d);     // below, next to nl_malloc's cache

static int ensure_scratch(Device &d, int stream, size_t need) {
  Scratch &s = d.scratch[stream];
  if (s.bytes >= need) return NL_OK;
  if (s.base) {
    CUDA_TRY(cudaStreamSynchronize(d.stream[stream]));
    CUDA_TRY(cudaFree(s.base));
    s.base = nullptr;
    s.bytes = 0;
  }
  size_t want = need + (need >> 2) + (1u << 20);
  cudaError_t e = cudaMalloc((void **)&s.base, want);
  if (e == cudaErrorMemoryAllocation) {
    // parked column blocks (nl_malloc's cache) and the pool's free blocks may hold the memory: give them back, retry
    cudaGetLastError();
    release_cached_blocks(d);
    e = cudaMalloc((void **)&s.base, want);
  }
  CUDA_TRY(e);
  CUDA_TRY(cudaMemsetAsync(s.base, 0, want, d.stream[stream]));
  s.bytes = want;
  s.arena = 0;
  s.arena_clean[0] = s.arena_clean[1] = true;
  s.arena_dirty[0] = s.arena_dirty[1] = 0;
  return NL_OK;
}

// ------------------------------------------------------------------ NCCL binding
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.handle) return NL_OK;
  // the bare soname resolves to an already-loaded libnccl (e.g. the one torch bundles) first
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) { set_error("NCCL not found: %s", dlerror()); return NL_ERR_NCCL; }
#define BIND(field, sym)                                                     \
  *(void **)(&g_nccl.field) = dlsym(h, sym);                                 \
  if (!g_nccl.field) { set_error("NCCL symbol %s missing", sym); return NL_ERR_NCCL; }
  BIND(GetUniqueId, "ncclGetUniqueId")
  BIND(CommInitRank, "ncclCommInitRank")
  BIND(CommDestroy, "ncclCommDestroy")
  BIND(AllReduce, "ncclAllReduce")
  BIND(AllGather, "ncclAllGather")
  BIND(GroupStart, "ncclGroupStart")
  BIND(GroupEnd, "ncclGroupEnd")
  BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
  g_nccl.handle = h;
  return NL_OK;
}

#define NCCL_TRY(expr)                                                             \
  do {                                                                             \
    ncclResult_t r_ = (expr);                                                      \
    if (r_ != ncclSuccess) {                                                       \
      set_error("%s failed: %s", #expr, g_nccl.GetErrorString(r_));                \
      return NL_ERR_NCCL;                                                          \
    }                                                                              \
  } while (0)

}  // namespace nl

using namespace nl;

extern "C" {

// ======================================================================= runtime
int nl_init(int n_devices, const int *device_ids) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_ndev > 0) return NL_OK;   // idempotent
  // logical devices share one CUDA context: a kernel of one device may wait in its mailbox for a kernel of another,
  // and loading that second kernel lazily at its first launch would wait for the first one (CUDA's lazy module
  // loading may synchronise the context) — load everything when the context is created
  if (!device_ids && getenv("NL_LOGICAL_DEVICES")) setenv("CUDA_MODULE_LOADING", "EAGER", 0);
  int visible = 0;
  cudaError_t e = cudaGetDeviceCount(&visible);
  if (e != cudaSuccess || visible == 0) {
    set_error("no CUDA device visible (%s); numpyllvm_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count is 0");
    cudaGetLastError();
    return NL_ERR_NO_DEVICE;
  }
  // NL_LOGICAL_DEVICES=k: k devices of the runtime laid over the visible GPUs round-robin.  Every logical device has
  // its own streams, scratch arena and mailbox, so the multi-shard paths — the in-kernel exchange between devices,
  // the count scan of filters, repartition, the sort across devices — run on a box with ONE GPU exactly as they do
  // between GPUs (peer stores land in the same HBM instead of crossing NVLink; NCCL, which refuses two ranks on
  // one GPU, is left out: the fused finish is the only one).  For the tests; tests/test_gpu_logical_devices.py.
  static int logical_ids[NL_MAX_DEVICES];
  if (!device_ids) {
    const char *env = getenv("NL_LOGICAL_DEVICES");
    const int k = env ? atoi(env) : 0;
    if (k > 0) {
      if (n_devices <= 0) n_devices = k;
      if (n_devices > NL_MAX_DEVICES) n_devices = NL_MAX_DEVICES;
      for (int d = 0; d < n_devices; d++) logical_ids[d] = d % visible;
      device_ids = logical_ids;
    }
  }
  if (n_devices <= 0) n_devices = visible;
  if (n_devices > NL_MAX_DEVICES) n_devices = NL_MAX_DEVICES;
  if (n_devices > visible && !device_ids) { set_error("%d devices requested, %d visible", n_devices, visible); return NL_ERR_INVALID; }
  g_logical = false;
  for (int d = 0; d < n_devices; d++) {
    Device &dv = g_dev[d];
    dv.id = device_ids ? device_ids[d] : d;
    CUDA_TRY(cudaSetDevice(dv.id));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dv.id));
    if (prop.major < 10) { set_error("device %d is sm_%d%d; the kernel library is built for sm_100a only", dv.id, prop.major, prop.minor); return NL_ERR_NO_DEVICE; }
    dv.sm_count = prop.multiProcessorCount;
    for (int s = 0; s < NL_N_STREAMS; s++) CUDA_TRY(cudaStreamCreateWithFlags(&dv.stream[s], cudaStreamNonBlocking));
    // column buffers come from the device's stream-ordered pool and go back to it, never to the
    // driver: thunk storage is allocated and dropped at every evaluate and cudaMalloc/cudaFree of
    // multi-GiB blocks costs tens of milliseconds
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dv.id));
    uint64_t keep = UINT64_MAX;
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    CUDA_TRY(cudaMalloc((void **)&dv.gather, sizeof(int64_t) * 256));
    for (int o = 0; o < d; o++)
      if (g_dev[o].id == dv.id) g_logical = true;
  }
  // NVLink peer access between the local devices (shard concat, peer copies)
  for (int a = 0; a < n_devices; a++)
    for (int b = 0; b < n_devices; b++) {
      if (a == b || g_dev[a].id == g_dev[b].id) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, g_dev[a].id, g_dev[b].id) == cudaSuccess && can) {
        cudaSetDevice(g_dev[a].id);
        cudaError_t pe = cudaDeviceEnablePeerAccess(g_dev[b].id, 0);
        if (pe != cudaSuccess) cudaGetLastError();   // already enabled is fine
      }
    }
  g_ndev = n_devices;
  if (g_logical) {
    // two devices of the runtime on one GPU: a cudaMalloc between the launch on one and the launch on the other would
    // wait for the first kernel, which waits in its mailbox for the second — so the scratch arenas exist up front
    // (between real GPUs an allocation on one device never waits for a kernel on another)
    for (int d = 0; d < n_devices; d++) {
      CUDA_TRY(cudaSetDevice(g_dev[d].id));
      const int rc = ensure_scratch(g_dev[d], 0, kTileDescOff + ((size_t)64 << 20));
      if (rc) { g_ndev = 0; return rc; }
    }
    CUDA_TRY(cudaDeviceSynchronize());
  }
  return NL_OK;
}

constexpr size_t kCacheMin = 1u << 20;
static std::mutex g_cache_mu;

static size_t size_class(size_t bytes) {
  size_t p = 1;
  while ((p << 1) <= bytes) p <<= 1;
  const size_t granule = (p >> 3) > (2u << 20) ? (p >> 3) : (size_t)(2u << 20);
  return (bytes + granule - 1) / granule * granule;
}

static void cache_release_all(Device &d) {     // caller holds g_cache_mu and has set the device
  for (auto &kv : d.cache_free) {
    cudaFreeAsync(kv.second, d.stream[0]);
    d.cache_class.erase(kv.second);
  }
  d.cache_free.clear();
  d.cache_bytes = 0;
}

}  // extern "C"
void nl::release_cached_blocks(Device &d) {
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    cache_release_all(d);
  }
  cudaDeviceSynchronize();
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, d.id) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
}
extern "C" {

int nl_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  nl_comm_destroy();          // communicator, IPC mappings, mailboxes and epochs: nothing stale survives into a later nl_init
  for (int d = 0; d < g_ndev; d++) {
    Device &dv = g_dev[d];
    cudaSetDevice(dv.id);
    {
      std::lock_guard<std::mutex> lk(g_cache_mu);
      cache_release_all(dv);
      dv.cache_class.clear();
    }
    cudaDeviceSynchronize();
    if (dv.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(dv.comm);
    dv.comm = nullptr;
    for (int s = 0; s < NL_N_STREAMS; s++) {
      if (dv.scratch[s].base) cudaFree(dv.scratch[s].base);
      dv.scratch[s] = Scratch();
      if (dv.stream[s]) cudaStreamDestroy(dv.stream[s]);
      dv.stream[s] = nullptr;
    }
    if (dv.gather) cudaFree(dv.gather);
    dv.gather = nullptr;
  }
  g_ndev = 0;
  g_world = 0;
  return NL_OK;
}

int nl_device_count(void) { return g_ndev; }

int nl_device_info(int dev, char *name, size_t name_len, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem) {
  int rc = check_dev(dev, 0);
  if (rc) return rc;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, g_dev[dev].id));
  if (name && name_len) snprintf(name, name_len, "%s", prop.name);
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_mem) *total_mem = prop.totalGlobalMem;
  return NL_OK;
}

int nl_device_cuda_id(int dev) {
  int rc = check_dev(dev, 0);
  if (rc) return rc;
  return g_dev[dev].id;
}

// ---- CUDA graphs: a fixed sequence of launches captured once and replayed (launch-bound small pipelines) ----
struct nl_graph_s {
  cudaGraphExec_t exec;
  int dev;
};

int nl_graph_begin(int dev) {
  int rc = check_dev(dev, 0);
  if (rc) return rc;
  Device &d = g_dev[dev];
  CUDA_TRY(cudaSetDevice(d.id));
  // what a launch may need to allocate must exist before the capture starts
  rc = ensure_scratch(d, 0, kTileDescOff + ((size_t)1 << 22));
  if (rc) return rc;
  CUDA_TRY(cudaStreamBeginCapture(d.stream[0], cudaStreamCaptureModeThreadLocal));
  return NL_OK;
}

int nl_graph_end(int dev, nl_graph *graph_out) {
  int rc = check_dev(dev, 0);
  if (rc) return rc;
  if (!graph_out) { set_error("nl_graph_end: null"); return NL_ERR_INVALID; }
  Device &d = g_dev[dev];
  CUDA_TRY(cudaSetDevice(d.id));
  cudaGraph_t g = nullptr;
  CUDA_TRY(cudaStreamEndCapture(d.stream[0], &g));
  nl_graph_s *out = new nl_graph_s;
  out->dev = dev;
  cudaError_t e = cudaGraphInstantiate(&out->exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { delete out; CUDA_TRY(e); }
  *graph_out = out;
  return NL_OK;
}

int nl_graph_launch(nl_graph graph) {
  if (!graph) { set_error("nl_graph_launch: null"); return NL_ERR_INVALID; }
  Device &d = g_dev[graph->dev];
  CUDA_TRY(cudaSetDevice(d.id));
  CUDA_TRY(cudaGraphLaunch(graph->exec, d.stream[0]));
  // a captured compaction launch works in arena 0 behind a memset of its own: whatever the replay leaves there is
  // not what the bookkeeping of the plain launches knows
  d.scratch[0].arena_clean[0] = false;
  d.scratch[0].arena_dirty[0] = -1;
  return NL_OK;
}

int nl_graph_destroy(nl_graph graph) {
  if (!graph) return NL_OK;
  cudaGraphExecDestroy(graph->exec);
  delete graph;
  return NL_OK;
}

// Large blocks are recycled by size class on top of the stream-ordered pool.  Measured on 2 x B200 with peer access
// between the pools (tools/jit_sort_bench.py, NL_DEBUG_ALLOC=1): a request that no free block of the pool fits — a
// bucket of 1 GiB + 1 MB after a 1 GiB block was freed — is REFUSED by cudaMallocAsync with 178 GB free at the driver;
// the trim-and-retry below then works but costs 14 ms, and so does every later allocation that has to map fresh
// memory.  Thunk storage, filter outputs and sort buckets come and go at every evaluate with sizes that differ by a
// few elements, so: sizes above 1 MiB are rounded up to 1/8 of their power of two (<= 12.5 % slack), nl_free() parks
// the block under its class, and the next request of the class takes it.  Both calls are compute-stream ordered
// (stream 0), so a parked block's last use precedes its next one.  nl_pool_trim() and an allocation failure empty the
// cache.
int nl_malloc(int dev, size_t bytes, void **ptr_out) {
  int rc = check_dev(dev, 0);
  if (rc) return rc;
  if (!ptr_out) { set_error("nl_malloc: null ptr_out"); return NL_ERR_INVALID; }
  static const bool trace = getenv("NL_DEBUG_ALLOC") != nullptr;    // slow or refused allocations on stderr
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  Device &d = g_dev[dev];
  size_t want = bytes ? bytes : 1;
  const bool cached = want > kCacheMin;
  if (cached) {
    want = size_class(want);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = d.cache_free.find(want);
    if (it != d.cache_free.end()) {
      *ptr_out = it->second;
      d.cache_free.erase(it);
      d.cache_bytes -= want;
      return NL_OK;
    }
  }
  // stream-ordered on the compute stream: usable there at once; other streams order themselves
  // after it with nl_stream_wait()
  const auto t0 = std::chrono::steady_clock::now();
  cudaError_t e = cudaMallocAsync(ptr_out, want, d.stream[0]);
  if (trace) {
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (ms > 0.5) {
      cudaMemPool_t pool;
      uint64_t reserved = 0, used = 0;
      if (cudaDeviceGetDefaultMemPool(&pool, d.id) == cudaSuccess) {
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
      }
      fprintf(stderr, "[nl_malloc] device %d: %zu bytes took %.2f ms (pool: %.1f MiB reserved, %.1f MiB in use)\n", dev, want, ms, reserved / 1048576.0,
              used / 1048576.0);
    }
  }
  if (e == cudaErrorMemoryAllocation) {
    // the pool (or the cache above it) may be holding blocks of the wrong sizes: give them back and retry once
    cudaGetLastError();
    if (trace) {
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      fprintf(stderr, "[nl_malloc] device %d: %zu bytes refused (driver: %.1f of %.1f MiB free, %.1f MiB parked); trimming the pool and retrying\n", dev, want,
              free_b / 1048576.0, total_b / 1048576.0, d.cache_bytes / 1048576.0);
    }
    {
      std::lock_guard<std::mutex> lk(g_cache_mu);
      cache_release_all(d);
    }
    cudaDeviceSynchronize();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, d.id) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    e = cudaMallocAsync(ptr_out, want, d.stream[0]);
  }
  if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); set_error("out of device memory allocating %zu bytes", want); return NL_ERR_NOMEM; }
  CUDA_TRY(e);
  if (cached) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    d.cache_class[*ptr_out] = want;
  }
  return NL_OK;
}

int nl_pool_trim(int dev) {
  int rc = check_dev(dev, 0);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    cache_release_all(g_dev[dev]);
  }
  CUDA_TRY(cudaDeviceSynchronize());
  cudaMemPool_t pool;
  CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, g_dev[dev].id));
  CUDA_TRY(cudaMemPoolTrimTo(pool, 0));
  return NL_OK;
}

int nl_free(int dev, void *ptr) {
  int rc = check_dev(dev, 0);
  if (rc) return rc;
  if (!ptr) return NL_OK;
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  Device &d = g_dev[dev];
  {
    // a large block waits under its size class for the next request (nl_malloc); in compute-stream order its next
    // user comes after everything already enqueued there
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = d.cache_class.find(ptr);
    if (it != d.cache_class.end()) {
      d.cache_free.emplace(it->second, ptr);
      d.cache_bytes += it->second;
      return NL_OK;
    }
  }
  // freed in compute-stream order: kernels already enqueued there still see the buffer; work on
  // the copy streams must have been ordered before this point by the caller
  CUDA_TRY(cudaFreeAsync(ptr, d.stream[0]));
  return NL_OK;
}

int nl_memset(int dev, int stream, void *ptr, int value, size_t bytes) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  CUDA_TRY(cudaMemsetAsync(ptr, value, bytes, g_dev[dev].stream[stream]));
  return NL_OK;
}

int nl_host_alloc(size_t bytes, void **ptr_out) {
  if (g_ndev == 0) { set_error("nl_host_alloc: no device initialised"); return NL_ERR_NO_DEVICE; }
  if (!ptr_out) { set_error("nl_host_alloc: null ptr_out"); return NL_ERR_INVALID; }
  cudaError_t e = cudaHostAlloc(ptr_out, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); set_error("out of pinned host memory allocating %zu bytes", bytes); return NL_ERR_NOMEM; }
  CUDA_TRY(e);
  return NL_OK;
}

int nl_host_free(void *ptr) {
  if (!ptr) return NL_OK;
  CUDA_TRY(cudaFreeHost(ptr));
  return NL_OK;
}

int nl_host_is_pinned(const void *ptr) {
  if (g_ndev == 0 || !ptr) return 0;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) { cudaGetLastError(); return 0; }
  return attr.type == cudaMemoryTypeHost ? 1 : 0;
}

static int copy_async(int dev, int stream, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (bytes == 0) return NL_OK;
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, kind, g_dev[dev].stream[stream]));
  return NL_OK;
}
int nl_memcpy_h2d(int dev, int stream, void *dst, const void *src, size_t bytes) { return copy_async(dev, stream, dst, src, bytes, cudaMemcpyHostToDevice); }
int nl_memcpy_d2h(int dev, int stream, void *dst, const void *src, size_t bytes) { return copy_async(dev, stream, dst, src, bytes, cudaMemcpyDeviceToHost); }
int nl_memcpy_d2d(int dev, int stream, void *dst, const void *src, size_t bytes) { return copy_async(dev, stream, dst, src, bytes, cudaMemcpyDeviceToDevice); }

int nl_memcpy_peer(int dst_dev, int stream, void *dst, int src_dev, const void *src, size_t bytes) {
  int rc = check_dev(dst_dev, stream);
  if (rc) return rc;
  rc = check_dev(src_dev, 0);
  if (rc) return rc;
  if (bytes == 0) return NL_OK;
  CUDA_TRY(cudaSetDevice(g_dev[dst_dev].id));
  CUDA_TRY(cudaMemcpyPeerAsync(dst, g_dev[dst_dev].id, src, g_dev[src_dev].id, bytes, g_dev[dst_dev].stream[stream]));
  return NL_OK;
}

int nl_stream_sync(int dev, int stream) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(g_dev[dev].stream[stream]));
  return check_peer_status(g_dev[dev]);
}

int nl_device_sync(int dev) {
  int rc = check_dev(dev, 0);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  CUDA_TRY(cudaDeviceSynchronize());
  return check_peer_status(g_dev[dev]);
}

int nl_stream_wait(int dev, int waiter, int signaler_dev, int signaler) {
  int rc = check_dev(dev, waiter);
  if (rc) return rc;
  rc = check_dev(signaler_dev, signaler);
  if (rc) return rc;
  cudaEvent_t ev;
  CUDA_TRY(cudaSetDevice(g_dev[signaler_dev].id));
  CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(ev, g_dev[signaler_dev].stream[signaler]));
  CUDA_TRY(cudaStreamWaitEvent(g_dev[dev].stream[waiter], ev, 0));
  CUDA_TRY(cudaEventDestroy(ev));   // released once the wait has consumed it
  return NL_OK;
}

struct nl_event_s {
  cudaEvent_t ev;
  int dev;
};

int nl_event_create(int dev, nl_event *ev_out) {
  int rc = check_dev(dev, 0);
  if (rc) return rc;
  if (!ev_out) { set_error("nl_event_create: null"); return NL_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  nl_event_s *e = new nl_event_s;
  e->dev = dev;
  cudaError_t ce = cudaEventCreate(&e->ev);
  if (ce != cudaSuccess) { delete e; CUDA_TRY(ce); }
  *ev_out = e;
  return NL_OK;
}
int nl_event_create_sync(int dev, nl_event *ev_out) {
  int rc = check_dev(dev, 0);
  if (rc) return rc;
  if (!ev_out) { set_error("nl_event_create_sync: null"); return NL_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  nl_event_s *e = new nl_event_s;
  e->dev = dev;
  cudaError_t ce = cudaEventCreateWithFlags(&e->ev, cudaEventDisableTiming);
  if (ce != cudaSuccess) { delete e; CUDA_TRY(ce); }
  *ev_out = e;
  return NL_OK;
}
int nl_event_query(nl_event ev) {
  if (!ev) { set_error("null event"); return NL_ERR_INVALID; }
  cudaError_t ce = cudaEventQuery(ev->ev);
  if (ce == cudaSuccess) return 1;
  if (ce == cudaErrorNotReady) { cudaGetLastError(); return 0; }
  CUDA_TRY(ce);
  return 0;
}
int nl_stream_wait_event(int dev, int stream, nl_event ev) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (!ev) { set_error("null event"); return NL_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  CUDA_TRY(cudaStreamWaitEvent(g_dev[dev].stream[stream], ev->ev, 0));
  return NL_OK;
}
int nl_event_destroy(nl_event ev) {
  if (!ev) return NL_OK;
  cudaEventDestroy(ev->ev);
  delete ev;
  return NL_OK;
}
int nl_event_record(nl_event ev, int stream) {
  if (!ev) { set_error("null event"); return NL_ERR_INVALID; }
  int rc = check_dev(ev->dev, stream);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(g_dev[ev->dev].id));
  CUDA_TRY(cudaEventRecord(ev->ev, g_dev[ev->dev].stream[stream]));
  return NL_OK;
}
int nl_event_sync(nl_event ev) {
  if (!ev) { set_error("null event"); return NL_ERR_INVALID; }
  CUDA_TRY(cudaEventSynchronize(ev->ev));
  return NL_OK;
}
int nl_event_elapsed_ms(nl_event start, nl_event stop, float *ms_out) {
  if (!start || !stop || !ms_out) { set_error("null event"); return NL_ERR_INVALID; }
  CUDA_TRY(cudaEventElapsedTime(ms_out, start->ev, stop->ev));
  return NL_OK;
}

// ======================================================================= hot path
// this rank cannot take part in epoch `epoch` after all: tell every peer's mailbox (their kernels poll the word)
static void post_abort(Device &d, unsigned long long epoch) {
  cudaSetDevice(d.id);
  for (int r = 0; r < g_world; r++)
    if (d.box_host_table[r])
      cudaMemcpyAsync(d.box_host_table[r] + 2 * kMaxWorld * 2, &epoch, sizeof(epoch), cudaMemcpyHostToDevice, d.stream[2]);
  cudaStreamSynchronize(d.stream[2]);
  cudaGetLastError();
}

static int launch_impl(const nl_plan *plan, void *const *outputs, const void *const *inputs, int64_t start, int64_t end,
                       int64_t *output_sizes, int thread_nr, int dev, int stream, bool fuse_finish, int64_t *peer_offsets = nullptr) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (!plan || plan->abi_version != NL_ABI_VERSION) { set_error("nl_launch_pipeline: bad plan / ABI version"); return NL_ERR_INVALID; }
  if (plan->n_steps <= 0 || plan->n_steps > NL_MAX_STEPS || plan->n_inputs > NL_MAX_INPUTS || plan->n_outputs <= 0 ||
      plan->n_outputs > NL_MAX_OUTPUTS) { set_error("nl_launch_pipeline: plan out of range"); return NL_ERR_INVALID; }
  if (!outputs || (plan->n_inputs > 0 && !inputs) || !output_sizes) { set_error("nl_launch_pipeline: null argument"); return NL_ERR_INVALID; }
  if (start < 0 || end < start || thread_nr < 0) { set_error("nl_launch_pipeline: bad range [%lld,%lld) / thread_nr", (long long)start, (long long)end); return NL_ERR_INVALID; }
  const size_t tsize = nl_dtype_size(plan->dtype);
  if (tsize == 0 || plan->dtype == NL_BOOL) { set_error("nl_launch_pipeline: bad plan dtype"); return NL_ERR_INVALID; }

  Device &d = g_dev[dev];
  CUDA_TRY(cudaSetDevice(d.id));

  KParams kp;
  memset(&kp, 0, sizeof(kp));
  const int V16 = nl_vec_elems(plan->dtype);
  bool vec16 = true;
  for (int k = 0; k < plan->n_inputs; k++) {
    kp.in[k] = inputs[k];
    kp.imm[k] = plan->in[k].imm_bits;
    kp.in_kind[k] = (uint8_t)plan->in[k].kind;
    if (plan->in[k].kind != NL_IN_SCALAR_IMM && !inputs[k]) { set_error("input %d is a null pointer", k); return NL_ERR_INVALID; }
    if (plan->in[k].kind == NL_IN_ARRAY) {
      const size_t es = nl_dtype_size(plan->in[k].dtype);
      const uintptr_t addr = (uintptr_t)inputs[k] + (uintptr_t)start * es;
      if (addr % (es * V16)) vec16 = false;
    }
  }
  for (int o = 0; o < plan->n_outputs; o++) {
    if (!outputs[o]) { set_error("output %d is a null pointer", o); return NL_ERR_INVALID; }
    kp.out[o] = outputs[o];
    kp.out_space[o] = (uint8_t)plan->out[o].index_space;
    if (plan->out[o].index_space == NL_IDX_LOOP || plan->out[o].index_space == NL_IDX_WHERE) {
      const size_t es = nl_dtype_size(plan->out[o].dtype);
      const uintptr_t addr = (uintptr_t)outputs[o] + (uintptr_t)start * es;
      if (addr % (es * V16)) vec16 = false;
    }
  }
  kp.start = start;
  kp.end = end;
  kp.out_sizes = output_sizes;
  kp.n_steps = plan->n_steps;
  kp.n_inputs = plan->n_inputs;
  kp.n_outputs = plan->n_outputs;
  kp.thread_nr = thread_nr;
  memcpy(kp.step, plan->step, sizeof(uint32_t) * (size_t)plan->n_steps);
  for (int i = 0; i < plan->n_steps; i++)
    if (step_op(plan->step[i]) == S_REDUCE) { kp.reduce_rop = (int32_t)step_a(plan->step[i]); kp.reduce_out = (int32_t)step_b(plan->step[i]); }

  if (fuse_finish) {
    // every rank launches the same sequence of fused reductions: the per-device counter is the epoch.  It is
    // committed (d.epoch) only when the launch is certain — scratch allocated, kernel selected —, and a launch that
    // still fails after that posts an abort for the epoch so that the peers' kernels stop waiting for this rank.
    kp.peer_box = d.box_table;
    kp.peer_world = g_world;
    kp.peer_rank = g_first_rank + dev;
    kp.peer_epoch = d.epoch + 1;
    kp.peer_offsets = (long long *)peer_offsets;
    kp.peer_timeout_ns = g_peer_timeout_ns;
    kp.peer_status = d.status;
  }
  if (end == start) {
    if (fuse_finish) d.epoch++;
    rc = launch_empty_range(d.stream[stream], plan, kp);
    if (rc != NL_OK && fuse_finish) post_abort(d, kp.peer_epoch);
    return rc;
  }

  const int64_t tile = pipeline_tile_elems(plan, vec16, end - start, &kp, d.sm_count);   // also selects the staged compaction pipeline
  kp.n_tiles = (end - start + tile - 1) / tile;
  const bool compact = (plan->flags & NL_PLAN_HAS_COMPACT) != 0;
  // two arenas of (ticket, ranks, descriptors) behind the reduction scratch
  const size_t arena_head = kTileDescOff - kTileTicketOff;
  const size_t arena_need = arena_head + (compact ? (size_t)kp.n_tiles * 8 * NL_DESC_STRIDE : 0);
  rc = ensure_scratch(d, stream, kTileTicketOff + 2 * arena_need + 256);
  if (rc) return rc;
  Scratch &sc = d.scratch[stream];
  char *sb = sc.base;
  const size_t arena_bytes = ((sc.bytes - kTileTicketOff) / 2) & ~(size_t)255;
  kp.red_ticket = (unsigned int *)(sb + kRedTicketOff);
  kp.red_partials = (unsigned long long *)(sb + kRedPartialsOff);
  int cur = 0;
  kp.clean_count = 0;
  kp.clean_desc = nullptr;
  kp.clean_ticket = kp.clean_rank = nullptr;
  if (compact) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CUDA_TRY(cudaStreamIsCapturing(d.stream[stream], &cap));
    const bool capturing = cap != cudaStreamCaptureStatusNone;
    const bool alternate = kp.super_k > 0 && g_trust_self_clean && !capturing;
    cur = alternate ? sc.arena : 0;
    char *ab = sb + kTileTicketOff + (size_t)cur * arena_bytes;
    if (alternate) {
      // the super-tile kernel zeroes the OTHER arena — what the previous launch left there — while it starts
      // (nl_compact_super.cuh); the host only steps in when this launch's arena is not known to be clean
      if (!sc.arena_clean[cur]) CUDA_TRY(cudaMemsetAsync(ab, 0, arena_bytes, d.stream[stream]));
      const int other = 1 - cur;
      char *ob = sb + kTileTicketOff + (size_t)other * arena_bytes;
      if (!sc.arena_clean[other]) {
        const int64_t cap_desc = (int64_t)((arena_bytes - arena_head) / (8 * NL_DESC_STRIDE));
        kp.clean_count = (sc.arena_dirty[other] < 0 || sc.arena_dirty[other] > cap_desc) ? cap_desc : sc.arena_dirty[other];
        if (kp.clean_count < 1) kp.clean_count = 1;                // (ticket and ranks are cleaned with the descriptors)
        kp.clean_ticket = (unsigned int *)ob;
        kp.clean_rank = (unsigned int *)(ob + (kSmRankOff - kTileTicketOff));
        kp.clean_desc = (unsigned long long *)(ob + arena_head);
      }
      sc.arena_clean[other] = true;
      sc.arena_clean[cur] = false;
      sc.arena_dirty[cur] = kp.n_tiles;
      sc.arena = other;
    } else {
      // the other compaction kernels, a launch that is only captured into a graph (nothing runs now, and a replay
      // must find its arena clean every time), nl_tune("compact_self_clean", 0): arena 0, cleared in front
      CUDA_TRY(cudaMemsetAsync(ab, 0, arena_head + (size_t)kp.n_tiles * 8 * NL_DESC_STRIDE, d.stream[stream]));
      if (!capturing) {
        // (the memset above covers this launch's descriptors only: what an earlier launch left beyond them stays)
        if (sc.arena_clean[0]) sc.arena_dirty[0] = kp.n_tiles;
        else if (sc.arena_dirty[0] >= 0 && sc.arena_dirty[0] < kp.n_tiles) sc.arena_dirty[0] = kp.n_tiles;
        sc.arena_clean[0] = false;
      }
    }
  }
  {
    char *ab = sb + kTileTicketOff + (size_t)cur * arena_bytes;
    kp.tile_ticket = (unsigned int *)ab;
    kp.sm_rank = (unsigned int *)(ab + (kSmRankOff - kTileTicketOff));
    kp.tile_desc = (unsigned long long *)(ab + arena_head);
  }

  LaunchInfo info;
  if (fuse_finish) d.epoch++;
  NvtxRange range(fuse_finish ? (peer_offsets ? "nl_pipeline+allgather" : "nl_pipeline+allreduce") : "nl_pipeline");
  rc = launch_pipeline_kernel(plan, kp, vec16, d.sm_count, d.stream[stream], &info);
  if (rc == NL_OK) g_last_info = info;
  else if (fuse_finish) post_abort(d, kp.peer_epoch);
  return rc;
}

int nl_launch_pipeline(const nl_plan *plan, void *const *outputs, const void *const *inputs, int64_t start,
                       int64_t end, int64_t *output_sizes, int thread_nr, int dev, int stream) {
  return launch_impl(plan, outputs, inputs, start, end, output_sizes, thread_nr, dev, stream, false);
}

int nl_launch_pipeline_allreduce(const nl_plan *plan, void *const *outputs, const void *const *inputs, int64_t start,
                                 int64_t end, int64_t *output_sizes, int dev) {
  if (!g_peer_ready) { set_error("nl_launch_pipeline_allreduce: peer mailboxes are not set up (nl_comm_init / nl_comm_mailbox_import)"); return NL_ERR_NCCL; }
  if (!plan || !(plan->flags & NL_PLAN_HAS_REDUCE)) { set_error("nl_launch_pipeline_allreduce: the plan has no reduction"); return NL_ERR_INVALID; }
  // one stream only: the mailboxes are two epochs deep and epochs must complete in launch order
  return launch_impl(plan, outputs, inputs, start, end, output_sizes, 0, dev, 0, true);
}

int nl_launch_pipeline_allgather(const nl_plan *plan, void *const *outputs, const void *const *inputs, int64_t start,
                                 int64_t end, int64_t *output_sizes, int64_t *offsets, int dev) {
  if (!g_peer_ready) { set_error("nl_launch_pipeline_allgather: peer mailboxes are not set up (nl_comm_init / nl_comm_mailbox_import)"); return NL_ERR_NCCL; }
  if (!plan || !(plan->flags & NL_PLAN_HAS_COMPACT) || (plan->flags & NL_PLAN_HAS_REDUCE)) { set_error("nl_launch_pipeline_allgather: the plan is not a filter"); return NL_ERR_INVALID; }
  if (!offsets) { set_error("nl_launch_pipeline_allgather: null offsets"); return NL_ERR_INVALID; }
  return launch_impl(plan, outputs, inputs, start, end, output_sizes, 0, dev, 0, true, offsets);
}

static int rop_of(int reduce_op) { return reduce_op == NL_OP_MAX ? R_MAX : reduce_op == NL_OP_MIN ? R_MIN : reduce_op == NL_OP_SUM ? R_SUM : -1; }

int nl_finalize_partials(int dev, int stream, int reduce_op, int dtype, const void *partials, int n, void *out) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (rop_of(reduce_op) < 0 || !partials || !out || n < 0) { set_error("nl_finalize_partials: bad argument"); return NL_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  return launch_finalize_partials(g_dev[dev].stream[stream], rop_of(reduce_op), dtype, partials, n, out);
}

int nl_assign_fill(int dev, int stream, void *dst, int dtype, int64_t lo, int64_t step, int64_t count, uint64_t value_bits) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (!dst || lo < 0 || count < 0 || step == 0) { set_error("nl_assign_fill: bad argument"); return NL_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  return launch_assign_fill(g_dev[dev].stream[stream], g_dev[dev].sm_count, dst, dtype, lo, step, count, value_bits);
}

int nl_assign_copy(int dev, int stream, void *dst, int dtype, int64_t lo, int64_t step, int64_t count, const void *src) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (!dst || !src || lo < 0 || count < 0 || step == 0) { set_error("nl_assign_copy: bad argument"); return NL_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  return launch_assign_copy(g_dev[dev].stream[stream], g_dev[dev].sm_count, dst, dtype, lo, step, count, src);
}

int nl_fill(int dev, int stream, void *dst, int dtype, int64_t first_index, int64_t count, uint64_t seed, int kind) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (!dst || count < 0) { set_error("nl_fill: bad argument"); return NL_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  return launch_fill(g_dev[dev].stream[stream], g_dev[dev].sm_count, dst, dtype, first_index, count, seed, kind);
}

int nl_gather(int dev, int stream, int dtype, void *out, const int64_t *idx, int64_t count, const void *const *shard_ptr,
              const int64_t *shard_start, int n_shards, int64_t total, int64_t *bad) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (count < 0 || n_shards < 1 || n_shards > NL_MAX_DEVICES || total < 0 || !shard_ptr || !shard_start || !bad ||
      (count > 0 && (!out || !idx))) { set_error("nl_gather: bad argument"); return NL_ERR_INVALID; }
  if (count == 0) return NL_OK;
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  return launch_gather(g_dev[dev].stream[stream], g_dev[dev].sm_count, dtype, out, idx, count, shard_ptr, shard_start, n_shards, total, bad);
}

int nl_cast(int dev, int stream, void *dst, int dst_dtype, const void *src, int src_dtype, int64_t count) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (count < 0 || (count > 0 && (!dst || !src))) { set_error("nl_cast: bad argument"); return NL_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  return launch_cast(g_dev[dev].stream[stream], g_dev[dev].sm_count, dst, dst_dtype, src, src_dtype, count);
}

int nl_checksum(int dev, int stream, const void *ptr, size_t nbytes, uint64_t *out) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (!out || (nbytes > 0 && !ptr)) { set_error("nl_checksum: bad argument"); return NL_ERR_INVALID; }
  if (((uintptr_t)ptr & 7) != 0) { set_error("nl_checksum: the buffer must be 8-byte aligned"); return NL_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  return launch_checksum(g_dev[dev].stream[stream], g_dev[dev].sm_count, ptr, nbytes, (unsigned long long *)out);
}

int nl_sort(int dev, int stream, int dtype, void *keys, void *tmp, int64_t count) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (count < 0 || (count > 0 && (!keys || !tmp))) { set_error("nl_sort: bad argument"); return NL_ERR_INVALID; }
  if (count <= 1) return NL_OK;
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  return launch_sort(g_dev[dev].stream[stream], g_dev[dev].sm_count, dtype, keys, tmp, count);
}

int nl_sort_sample(int dev, int stream, int dtype, const void *keys, int64_t count, int n_samples, uint64_t *host_keys) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (count <= 0 || !keys || n_samples <= 0 || !host_keys) { set_error("nl_sort_sample: bad argument"); return NL_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  return launch_sort_sample(g_dev[dev].stream[stream], dtype, keys, count, n_samples, (unsigned long long *)host_keys);
}

static int sort_split_any(const char *what, int mode, int dev, int stream, int dtype, const void *keys, void *out, int64_t count, const uint64_t *splitters,
                          int n_splitters, int64_t *host_counts, void *const *run_dst) {
  int rc = check_dev(dev, stream);
  if (rc) return rc;
  if (count < 0 || (count > 0 && !keys) || n_splitters < 0 || n_splitters > kSortMaxSplitters || (n_splitters > 0 && !splitters) ||
      (mode == 0 && count > 0 && !out) || (mode != 2 && !host_counts) || (mode == 2 && !run_dst)) {
    set_error("%s: bad argument (at most %d splitters)", what, kSortMaxSplitters);
    return NL_ERR_INVALID;
  }
  for (int j = 1; j < n_splitters; j++)
    if (splitters[j] < splitters[j - 1]) { set_error("%s: the splitters must ascend", what); return NL_ERR_INVALID; }
  if (count == 0) {
    if (mode != 2)
      for (int j = 0; j <= n_splitters; j++) host_counts[j] = 0;
    return NL_OK;
  }
  CUDA_TRY(cudaSetDevice(g_dev[dev].id));
  return launch_sort_split(g_dev[dev].stream[stream], g_dev[dev].sm_count, mode, dtype, keys, out, count, (const unsigned long long *)splitters, n_splitters,
                           host_counts, run_dst);
}

int nl_sort_split(int dev, int stream, int dtype, const void *keys, void *out, int64_t count, const uint64_t *splitters, int n_splitters,
                  int64_t *host_counts) {
  return sort_split_any("nl_sort_split", 0, dev, stream, dtype, keys, out, count, splitters, n_splitters, host_counts, nullptr);
}

int nl_sort_split_count(int dev, int stream, int dtype, const void *keys, int64_t count, const uint64_t *splitters, int n_splitters, int64_t *host_counts) {
  return sort_split_any("nl_sort_split_count", 1, dev, stream, dtype, keys, nullptr, count, splitters, n_splitters, host_counts, nullptr);
}

int nl_sort_split_scatter(int dev, int stream, int dtype, const void *keys, int64_t count, const uint64_t *splitters, int n_splitters, void *const *run_dst) {
  return sort_split_any("nl_sort_split_scatter", 2, dev, stream, dtype, keys, nullptr, count, splitters, n_splitters, nullptr, run_dst);
}

// ======================================================================= communicator
int nl_comm_unique_id(void *id_out) {
  if (!id_out) { set_error("nl_comm_unique_id: null"); return NL_ERR_INVALID; }
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == NL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  memcpy(id_out, &id, sizeof(id));
  return NL_OK;
}

static int setup_local_mailboxes();

int nl_comm_init(const void *id, int world_size, int first_rank) {
  if (g_ndev == 0) { set_error("nl_comm_init before nl_init"); return NL_ERR_NO_DEVICE; }
  if (!id || world_size < g_ndev || first_rank < 0 || first_rank + g_ndev > world_size || world_size > 256) { set_error("nl_comm_init: bad argument"); return NL_ERR_INVALID; }
  int rc = load_nccl();
  if (rc) return rc;
  if (g_world) { set_error("communicator already initialised"); return NL_ERR_INVALID; }
  if (g_logical) {
    // logical devices on one GPU: NCCL refuses two ranks per GPU; the mailboxes below are the communicator
    if (g_ndev != world_size) { set_error("nl_comm_init: logical devices (NL_LOGICAL_DEVICES) only within one process"); return NL_ERR_INVALID; }
  } else {
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    NCCL_TRY(g_nccl.GroupStart());
    for (int d = 0; d < g_ndev; d++) {
      CUDA_TRY(cudaSetDevice(g_dev[d].id));
      NCCL_TRY(g_nccl.CommInitRank(&g_dev[d].comm, world_size, uid, first_rank + d));
    }
    NCCL_TRY(g_nccl.GroupEnd());
  }
  g_world = world_size;
  g_first_rank = first_rank;
  g_peer_ready = false;
  if (g_ndev == world_size) return setup_local_mailboxes();   // all ranks in this process: P2P mailboxes right away
  return NL_OK;
}

int nl_comm_world_size(void) { return g_world; }

// ---- NVLink mailboxes (fused reduction finish, nl_pipeline.cuh: peer_allreduce) ----
static int ensure_mailbox(int d) {
  Device &dv = g_dev[d];
  CUDA_TRY(cudaSetDevice(dv.id));
  if (!dv.box) {
    CUDA_TRY(cudaMalloc((void **)&dv.box, kMailboxBytes));
    CUDA_TRY(cudaMemset(dv.box, 0, kMailboxBytes));
  }
  if (!dv.box_table) CUDA_TRY(cudaMalloc((void **)&dv.box_table, sizeof(unsigned long long *) * kMaxWorld));
  if (!dv.status) {
    CUDA_TRY(cudaHostAlloc((void **)&dv.status, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    *dv.status = 0;
  }
  return NL_OK;
}

// all ranks live in this process: map every local mailbox into every local device (P2P)
static int setup_local_mailboxes() {
  for (int d = 0; d < g_ndev; d++)
    for (int e = 0; e < g_ndev; e++) {
      if (d == e || g_dev[d].id == g_dev[e].id) continue;
      int can = 0;
      CUDA_TRY(cudaDeviceCanAccessPeer(&can, g_dev[d].id, g_dev[e].id));
      if (!can) return NL_OK;                      // no P2P between some pair: stay on the NCCL finish
    }
  for (int d = 0; d < g_ndev; d++) {
    int rc = ensure_mailbox(d);
    if (rc) return rc;
    for (int e = 0; e < g_ndev; e++) {
      if (d == e || g_dev[d].id == g_dev[e].id) continue;
      cudaError_t err = cudaDeviceEnablePeerAccess(g_dev[e].id, 0);
      if (err == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (err != cudaSuccess) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(err)); return NL_ERR_CUDA; }
    }
  }
  // columns come from the stream-ordered pools: peers may read and write them too (gather reads
  // remote shards with plain loads; cudaDeviceEnablePeerAccess alone does not cover pool memory)
  for (int d = 0; d < g_ndev; d++) {
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, g_dev[d].id));
    for (int e = 0; e < g_ndev; e++) {
      if (d == e || g_dev[d].id == g_dev[e].id) continue;
      cudaMemAccessDesc desc;
      memset(&desc, 0, sizeof(desc));
      desc.location.type = cudaMemLocationTypeDevice;
      desc.location.id = g_dev[e].id;
      desc.flags = cudaMemAccessFlagsProtReadWrite;
      CUDA_TRY(cudaMemPoolSetAccess(pool, &desc, 1));
    }
  }
  unsigned long long *table[kMaxWorld] = {nullptr};
  for (int d = 0; d < g_ndev; d++) table[d] = g_dev[d].box;
  for (int d = 0; d < g_ndev; d++) {
    CUDA_TRY(cudaSetDevice(g_dev[d].id));
    CUDA_TRY(cudaMemcpy(g_dev[d].box_table, table, sizeof(table[0]) * (size_t)g_ndev, cudaMemcpyHostToDevice));
    memcpy(g_dev[d].box_host_table, table, sizeof(table));
    g_dev[d].epoch = 0;
  }
  g_peer_ready = true;
  return NL_OK;
}

int nl_comm_peer_ready(void) { return g_peer_ready ? 1 : 0; }

int nl_comm_abort(int dev) {
  int rc = check_dev(dev, 0);
  if (rc) return rc;
  if (!g_peer_ready) { set_error("nl_comm_abort: no peer mailboxes"); return NL_ERR_NCCL; }
  post_abort(g_dev[dev], g_dev[dev].epoch + 1);
  return NL_OK;
}

int nl_comm_mailbox_export(int dev, void *handle_out) {
  int rc = check_dev(dev, 0);
  if (rc) return rc;
  if (!handle_out) { set_error("nl_comm_mailbox_export: null"); return NL_ERR_INVALID; }
  if (!g_world) { set_error("nl_comm_mailbox_export before nl_comm_init"); return NL_ERR_NCCL; }
  rc = ensure_mailbox(dev);
  if (rc) return rc;
  static_assert(sizeof(cudaIpcMemHandle_t) == NL_MAILBOX_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, g_dev[dev].box));
  memcpy(handle_out, &h, sizeof(h));
  return NL_OK;
}

int nl_comm_mailbox_import(const void *handles) {
  if (!handles) { set_error("nl_comm_mailbox_import: null"); return NL_ERR_INVALID; }
  if (!g_world || g_world > kMaxWorld) { set_error("nl_comm_mailbox_import before nl_comm_init"); return NL_ERR_NCCL; }
  // importing twice would restart the epochs over mailboxes that still hold the old ones (a slot left at epoch 1
  // would satisfy the first wait at once): a communicator is set up once, nl_comm_destroy() gives fresh mailboxes
  if (g_peer_ready) { set_error("nl_comm_mailbox_import: the mailboxes are already set up (nl_comm_destroy first)"); return NL_ERR_INVALID; }
  const char *hb = (const char *)handles;
  for (int d = 0; d < g_ndev; d++) {
    int rc = ensure_mailbox(d);
    if (rc) return rc;
    unsigned long long *table[kMaxWorld] = {nullptr};
    for (int r = 0; r < g_world; r++) {
      const int local = r - g_first_rank;
      if (local == d) { table[r] = g_dev[d].box; continue; }
      if (local >= 0 && local < g_ndev) {
        // another device of this process: plain peer access
        int rc2 = ensure_mailbox(local);
        if (rc2) return rc2;
        CUDA_TRY(cudaSetDevice(g_dev[d].id));
        cudaError_t err = cudaDeviceEnablePeerAccess(g_dev[local].id, 0);
        if (err == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (err != cudaSuccess) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(err)); return NL_ERR_CUDA; }
        table[r] = g_dev[local].box;
        continue;
      }
      cudaIpcMemHandle_t h;
      memcpy(&h, hb + (size_t)r * NL_MAILBOX_HANDLE_BYTES, sizeof(h));
      void *mapped = nullptr;
      CUDA_TRY(cudaSetDevice(g_dev[d].id));
      CUDA_TRY(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
      g_ipc_mapped[d][r] = mapped;
      table[r] = (unsigned long long *)mapped;
    }
    CUDA_TRY(cudaSetDevice(g_dev[d].id));
    CUDA_TRY(cudaMemcpy(g_dev[d].box_table, table, sizeof(table[0]) * (size_t)g_world, cudaMemcpyHostToDevice));
    memcpy(g_dev[d].box_host_table, table, sizeof(table));
    g_dev[d].epoch = 0;
  }
  g_peer_ready = true;
  return NL_OK;
}

int nl_comm_destroy(void) {
  for (int d = 0; d < g_ndev; d++) {
    if (g_dev[d].comm && g_nccl.CommDestroy) {
      cudaSetDevice(g_dev[d].id);
      g_nccl.CommDestroy(g_dev[d].comm);
    }
    g_dev[d].comm = nullptr;
    cudaSetDevice(g_dev[d].id);
    for (int r = 0; r < kMaxWorld; r++)
      if (g_ipc_mapped[d][r]) { cudaIpcCloseMemHandle(g_ipc_mapped[d][r]); g_ipc_mapped[d][r] = nullptr; }
    if (g_dev[d].box) { cudaFree(g_dev[d].box); g_dev[d].box = nullptr; }
    if (g_dev[d].box_table) { cudaFree(g_dev[d].box_table); g_dev[d].box_table = nullptr; }
    if (g_dev[d].status) { cudaFreeHost(g_dev[d].status); g_dev[d].status = nullptr; }
    memset(g_dev[d].box_host_table, 0, sizeof(g_dev[d].box_host_table));
    g_dev[d].epoch = 0;
  }
  g_world = 0;
  g_peer_ready = false;
  return NL_OK;
}

int nl_finish_reduce(int reduce_op, int dtype, void *const *scalar, int stream) {
  int rc = check_dev(0, stream);
  if (rc) return rc;
  if (rop_of(reduce_op) < 0 || !scalar) { set_error("nl_finish_reduce: bad argument"); return NL_ERR_INVALID; }
  if (g_world <= 1 && g_ndev == 1) return NL_OK;     // a single shard: its partial is the result
  if (!g_world) { set_error("nl_finish_reduce: %d local devices but no communicator (call nl_comm_init)", g_ndev); return NL_ERR_NCCL; }
  if (g_logical) { set_error("nl_finish_reduce: no NCCL communicator between logical devices of one GPU (use the fused finish)"); return NL_ERR_NCCL; }
  ncclDataType_t t;
  switch (dtype) {
    case NL_I32: t = ncclInt32; break;
    case NL_I64: t = ncclInt64; break;
    case NL_F32: t = ncclFloat32; break;
    case NL_F64: t = ncclFloat64; break;
    default: set_error("nl_finish_reduce: bad dtype"); return NL_ERR_INVALID;
  }
  const ncclRedOp_t op = reduce_op == NL_OP_MAX ? ncclMax : reduce_op == NL_OP_MIN ? ncclMin : ncclSum;
  NCCL_TRY(g_nccl.GroupStart());
  for (int d = 0; d < g_ndev; d++) {
    if (!scalar[d]) { g_nccl.GroupEnd(); set_error("nl_finish_reduce: null scalar for device %d", d); return NL_ERR_INVALID; }
    NCCL_TRY(g_nccl.AllReduce(scalar[d], scalar[d], 1, t, op, g_dev[d].comm, g_dev[d].stream[stream]));
  }
  NCCL_TRY(g_nccl.GroupEnd());
  return NL_OK;
}

int nl_finish_filter(const int64_t *const *count, int64_t *const *offsets, int stream) {
  int rc = check_dev(0, stream);
  if (rc) return rc;
  if (!count || !offsets) { set_error("nl_finish_filter: null"); return NL_ERR_INVALID; }
  const int world = (g_world > 0) ? g_world : 1;
  if (world == 1 && g_ndev > 1) { set_error("nl_finish_filter: %d local devices but no communicator", g_ndev); return NL_ERR_NCCL; }
  if (g_logical) { set_error("nl_finish_filter: no NCCL communicator between logical devices of one GPU (use the fused exchange)"); return NL_ERR_NCCL; }
  if (world > 1) {
    NCCL_TRY(g_nccl.GroupStart());
    for (int d = 0; d < g_ndev; d++)
      NCCL_TRY(g_nccl.AllGather(count[d], g_dev[d].gather, 1, ncclInt64, g_dev[d].comm, g_dev[d].stream[stream]));
    NCCL_TRY(g_nccl.GroupEnd());
  }
  for (int d = 0; d < g_ndev; d++) {
    CUDA_TRY(cudaSetDevice(g_dev[d].id));
    const int64_t *src = (world > 1) ? g_dev[d].gather : count[d];
    rc = launch_scan_offsets(g_dev[d].stream[stream], src, world, offsets[d]);
    if (rc) return rc;
  }
  return NL_OK;
}

// ======================================================================= introspection
int nl_plan_kernel(const nl_plan *plan, int vec16, char *name, size_t name_len, int *is_static) {
  if (!plan || plan->abi_version != NL_ABI_VERSION) { set_error("nl_plan_kernel: bad plan"); return NL_ERR_INVALID; }
  return describe_pipeline_kernel(plan, vec16 != 0, name, name_len, is_static);
}

int nl_set_static_kernels(int mode) {
  if (mode < 0 || mode > 2) { set_error("nl_set_static_kernels: mode must be 0 (interpreter only), 1 (all tiers) or 2 (families + interpreter)"); return NL_ERR_INVALID; }
  set_static_kernels_enabled(mode);
  return NL_OK;
}

int nl_tune(const char *key, int value) {
  if (!key) { set_error("nl_tune: null key"); return NL_ERR_INVALID; }
  if (!strcmp(key, "compact_kernel")) {
    // 0: by size (super-tile kernel, then staged pipeline, then plain); 1: plain kernel only;
    // 2: staged pipeline before the super-tile kernel
    if (value < 0 || value > 2) { set_error("nl_tune(compact_kernel): 0, 1 or 2"); return NL_ERR_INVALID; }
    set_compact_pipe_enabled(value == 1 ? 0 : 1);
    set_compact_prefer_pipe(value == 2 ? 1 : 0);
  } else if (!strcmp(key, "pipe_dynamic")) {
    set_compact_pipe_dynamic(value ? 1 : 0);
  } else if (!strcmp(key, "pipe_side_buffers")) {
    set_compact_pipe_side(value ? 1 : 0);
  } else if (!strcmp(key, "pipe_lag")) {
    if (value < 0 || value > 15) { set_error("nl_tune(pipe_lag): 0 (default) .. 15"); return NL_ERR_INVALID; }
    set_compact_pipe_lag(value);
  } else if (!strcmp(key, "max_ctas_per_sm")) {
    if (value < 0 || value > 32) { set_error("nl_tune(max_ctas_per_sm): 0 (occupancy) .. 32"); return NL_ERR_INVALID; }
    set_max_ctas_per_sm(value);
  } else if (!strcmp(key, "sort_algo")) {
    if (value < 0 || value > 1) { set_error("nl_tune(sort_algo): 0 one sweep per position (default), 1 count + scatter"); return NL_ERR_INVALID; }
    set_sort_algo(value);
  } else if (!strcmp(key, "compact_self_clean")) {
    g_trust_self_clean = value != 0;
  } else if (!strcmp(key, "sort_portion_log2")) {
    if (value != 0 && (value < 13 || value > 28)) { set_error("nl_tune(sort_portion_log2): 13 .. 28 (0: default 28)"); return NL_ERR_INVALID; }
    set_sort_portion_log2(value ? value : 28);
  } else if (!strcmp(key, "poll_sleep_ns")) {
    set_poll_sleep_ns(value);
  } else if (!strcmp(key, "peer_timeout_ms")) {
    set_peer_timeout_ms(value);
  } else if (!strcmp(key, "reset")) {
    set_peer_timeout_ms(0);
    set_compact_pipe_enabled(1); set_compact_prefer_pipe(0); set_compact_pipe_dynamic(0); set_compact_pipe_side(1);
    set_compact_pipe_lag(0); set_max_ctas_per_sm(0); set_poll_sleep_ns(0); set_sort_algo(0); set_sort_portion_log2(28); g_trust_self_clean = true;
  } else {
    set_error("nl_tune: unknown key '%s'", key);
    return NL_ERR_INVALID;
  }
  return NL_OK;
}

uint64_t nl_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int nl_last_launch_info(char *kernel_name, size_t name_len, int *grid, int *block, int *vec_elems) {
  if (kernel_name && name_len) snprintf(kernel_name, name_len, "%s", g_last_info.kernel);
  if (grid) *grid = g_last_info.grid;
  if (block) *block = g_last_info.block;
  if (vec_elems) *vec_elems = g_last_info.vec_elems;
  return NL_OK;
}

}  // extern "C"
