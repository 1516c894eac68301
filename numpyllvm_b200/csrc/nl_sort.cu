// nl_sort.cu — device radix sort for the `sort()` full pipeline breaker (sm_100a).
//
// The reference sorts with NumPy on the host (thunk_methods.cpp:37-62, the `base` function of a
// fullbreaker, executed by compiler.cpp:93-122); here the shard is sorted where it lives.
// Least-significant-digit radix sort, 8-bit digits, keys only, stable:
//
//   sort_histogram_kernel   one read of the keys: the 256-bin histogram of EVERY digit position.
//                           A position where all keys share one digit needs no pass at all
//                           (the reference's test data, ints in [1,100), sorts in ONE pass).
//   per remaining position: chunk_histogram_kernel (per-CTA-chunk digit counts)
//                           chunk_offsets_kernel   (where each chunk's run of each digit starts)
//                           scatter_kernel         (stable in-tile ranking with warp match,
//                                                   keys regrouped by digit in shared memory so
//                                                   that global stores form runs)
//
// Traffic per executed pass: 8 B read (counts) + 8 B read + 8 B write (scatter) per 8-byte key.
// Keys are compared through an order-preserving bijection of their bit patterns (sign flip for
// ints; the usual float flip, rotated so that negative NaNs also land at the end like NumPy's
// sort puts every NaN last); the stored values are the original bits.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "nl_abi.h"
#include "nl_internal.h"

#define CUDA_TRY(expr)                                                             \
  do {                                                                             \
    cudaError_t e_ = (expr);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      set_error("%s failed: %s", #expr, cudaGetErrorString(e_));                   \
      return NL_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

namespace nl {

#ifndef NL_SORT_ITEMS32
#define NL_SORT_ITEMS32 8
#endif
#ifndef NL_SORT_MIN_CTAS
#define NL_SORT_MIN_CTAS 4
#endif
constexpr int kSortBlock = 256;
constexpr int kSortItems = 8;                          // keys per thread per tile
constexpr int kSortTile = kSortBlock * kSortItems;     // 2048
constexpr int kSortWarps = kSortBlock / 32;

template <typename UK, int kFloat>
__device__ __forceinline__ UK order_key(UK bits) {
  constexpr int B = sizeof(UK) * 8;
  constexpr UK SIGN = (UK)1 << (B - 1);
  if constexpr (kFloat == 2) {
    return bits;                       // unsigned integers already sort by their bit patterns
  } else if constexpr (kFloat == 0) {
    return bits ^ SIGN;
  } else {
    const UK k = bits ^ ((bits >> (B - 1)) ? (UK)~(UK)0 : SIGN);
    // key of -inf: everything below it is a negative NaN; rotate those to the top
    constexpr UK NEG_INF = (sizeof(UK) == 8) ? (UK)0x000FFFFFFFFFFFFFull : (UK)0x007FFFFFu;
    return (UK)(k - NEG_INF);
  }
}

template <typename UK, int kFloat>
__global__ void __launch_bounds__(kSortBlock) sort_histogram_kernel(const UK *__restrict__ keys, long long n,
                                                                     unsigned long long *__restrict__ ghist /*[sizeof(UK)][256]*/) {
  constexpr int P = sizeof(UK);
  __shared__ unsigned int h[P][256];
  for (int i = threadIdx.x; i < P * 256; i += kSortBlock) (&h[0][0])[i] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * kSortBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kSortBlock) {
    const UK k = order_key<UK, kFloat>(keys[i]);
#pragma unroll
    for (int p = 0; p < P; p++) atomicAdd(&h[p][(k >> (8 * p)) & 0xff], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P * 256; i += kSortBlock) {
    const unsigned int c = (&h[0][0])[i];
    if (c) atomicAdd(&ghist[i], (unsigned long long)c);
  }
}

// counts of one digit position over the CTA's contiguous chunk -> hist[digit * gridDim.x + cta]
template <typename UK, int kFloat>
__global__ void __launch_bounds__(kSortBlock) chunk_histogram_kernel(const UK *__restrict__ keys, long long n, long long chunk, int shift,
                                                                      unsigned int *__restrict__ hist) {
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const long long lo = (long long)blockIdx.x * chunk;
  long long hi = lo + chunk;
  if (hi > n) hi = n;
  for (long long i = lo + threadIdx.x; i < hi; i += kSortBlock)
    atomicAdd(&h[(order_key<UK, kFloat>(keys[i]) >> shift) & 0xff], 1u);
  __syncthreads();
  hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// one CTA per digit: start of the digit in the output (from the global histogram) + running sum
// of the chunks' counts -> offs[digit * n_chunks + chunk]
__global__ void __launch_bounds__(kSortBlock) chunk_offsets_kernel(const unsigned long long *__restrict__ ghist_pass /*[256]*/,
                                                                    const unsigned int *__restrict__ hist, int n_chunks,
                                                                    long long *__restrict__ offs) {
  const int d = blockIdx.x;
  __shared__ long long base;
  __shared__ long long part[kSortBlock];
  if (threadIdx.x == 0) {
    long long b = 0;
    for (int j = 0; j < d; j++) b += (long long)ghist_pass[j];
    base = b;
  }
  // each thread owns a contiguous slice of the chunks
  const int per = (n_chunks + kSortBlock - 1) / kSortBlock;
  const int lo = threadIdx.x * per, hi = min(lo + per, n_chunks);
  long long s = 0;
  for (int c = lo; c < hi; c++) s += hist[(size_t)d * n_chunks + c];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = base;
    for (int t = 0; t < kSortBlock; t++) {
      const long long v = part[t];
      part[t] = run;
      run += v;
    }
  }
  __syncthreads();
  long long run = part[threadIdx.x];
  for (int c = lo; c < hi; c++) {
    offs[(size_t)d * n_chunks + c] = run;
    run += hist[(size_t)d * n_chunks + c];
  }
}

template <typename UK, int kFloat, int ITEMS>
__global__ void __launch_bounds__(kSortBlock, NL_SORT_MIN_CTAS) scatter_kernel(const UK *__restrict__ in, UK *__restrict__ out, long long n, long long chunk,
                                                              int shift, const long long *__restrict__ offs) {
  __shared__ unsigned int cnt[kSortWarps][256];     // per warp: keys of each digit seen so far in the tile
  __shared__ unsigned int tile_start[256];          // exclusive scan of the tile's digit counts
  __shared__ unsigned int tile_count[256];
  __shared__ long long goff[256];                   // where the next key of each digit goes
  __shared__ long long gdelta[256];                 // goff - start of the digit's run inside the tile
  __shared__ UK staged[(kSortBlock * ITEMS)];
  __shared__ unsigned int wsum[kSortWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned int lt = (1u << lane) - 1u;
  goff[tid] = offs[(size_t)tid * gridDim.x + blockIdx.x];
  const long long lo = (long long)blockIdx.x * chunk;
  long long hi = lo + chunk;
  if (hi > n) hi = n;

  for (long long base = lo; base < hi; base += (kSortBlock * ITEMS)) {
    const int valid = (int)((hi - base < (kSortBlock * ITEMS)) ? (hi - base) : (kSortBlock * ITEMS));
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) cnt[w][tid] = 0;
    __syncthreads();
    // warp-striped keys: item i of lane l in warp w is tile element w*256 + i*32 + l, so the
    // order (warp, item, lane) is the input order and the ranking below is stable
    UK key[ITEMS];
    unsigned int rank[ITEMS];
    int dig[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int e = warp * (32 * ITEMS) + i * 32 + lane;
      const bool ok = e < valid;
      key[i] = ok ? in[base + e] : (UK)0;
      dig[i] = ok ? (int)((order_key<UK, kFloat>(key[i]) >> shift) & 0xff) : -1;
    }
    // lanes holding the same digit, item by item.  Eight ballots (one per digit bit) instead of
    // match.any: measured on B200, half of the kernel's stall samples sat behind MATCH.ANY results.
    unsigned int peers[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const bool ok = dig[i] >= 0;
      unsigned int m = __ballot_sync(0xffffffffu, ok);
#pragma unroll
      for (int b = 0; b < 8; b++) {
        const bool bit = (dig[i] >> b) & 1;
        const unsigned int bal = __ballot_sync(0xffffffffu, bit);
        m &= bit ? bal : ~bal;
      }
      peers[i] = ok ? m : (1u << lane);
    }
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      unsigned int before = 0;
      if (dig[i] >= 0) before = cnt[warp][dig[i]];
      __syncwarp();
      rank[i] = before + __popc(peers[i] & lt);
      if (dig[i] >= 0 && (peers[i] & lt) == 0) cnt[warp][dig[i]] = before + __popc(peers[i]);
      __syncwarp();
    }
    __syncthreads();
    // thread d: digit d's counts of the warps -> exclusive over the warps, total of the tile
    {
      unsigned int run = 0;
#pragma unroll
      for (int w = 0; w < kSortWarps; w++) {
        const unsigned int c = cnt[w][tid];
        cnt[w][tid] = run;
        run += c;
      }
      tile_count[tid] = run;
      // exclusive scan of the 256 digit totals
      unsigned int incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) wsum[warp] = incl;
      __syncthreads();
      unsigned int wbase = 0;
#pragma unroll
      for (int w = 0; w < kSortWarps; w++)
        if (w < warp) wbase += wsum[w];
      const unsigned int ts = wbase + incl - run;
      tile_start[tid] = ts;
      // one shared-memory read per key in each of the two phases below instead of two (the ranking, not memory,
      // bounds this kernel: ncu, SM pipes 69 % / DRAM 21 %): the warps' bases absorb the digit's start in the tile,
      // and the output phase reads (global run start - tile start) as one number
#pragma unroll
      for (int w = 0; w < kSortWarps; w++) cnt[w][tid] += ts;
      gdelta[tid] = goff[tid] - (long long)ts;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
      if (dig[i] >= 0) staged[cnt[warp][dig[i]] + rank[i]] = key[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int pos = i * kSortBlock + tid;
      if (pos < valid) {
        const UK k = staged[pos];
        const int d = (int)((order_key<UK, kFloat>(k) >> shift) & 0xff);
        out[gdelta[d] + (long long)pos] = k;
      }
    }
    __syncthreads();
    goff[tid] += (long long)tile_count[tid];
  }
}

// ---- one sweep per digit position ------------------------------------------------------------------------
// The per-position counting pass is gone: a tile's digit histogram is chained to its predecessors' with a decoupled
// look-back (one 32-bit descriptor per tile and digit: 2 status bits + 30 count bits), so every executed position
// reads the keys once and writes them once — 16 B per 8-byte key instead of 24.  Tiles are handed out by a ticket,
// so every predecessor of a tile belongs to a running CTA.  The in-tile ranking is the scatter kernel's.
constexpr unsigned int kSweepAgg = 1u << 30, kSweepPrefix = 2u << 30, kSweepMask = (1u << 30) - 1u;

__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned int *p, unsigned int v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename UK, int kFloat, int ITEMS, bool kMatch>
__global__ void __launch_bounds__(kSortBlock, NL_SORT_MIN_CTAS) onesweep_kernel(const UK *__restrict__ in, UK *__restrict__ out, long long n, int shift,
                                                               const unsigned long long *__restrict__ ghist_pass /*[256]*/,
                                                               unsigned int *__restrict__ desc /*[n_tiles][256], zeroed*/,
                                                               unsigned int *__restrict__ ticket /*zeroed*/, int n_tiles) {
  constexpr int TILE = kSortBlock * ITEMS;
  __shared__ unsigned int cnt[kSortWarps][256];     // per warp: keys of each digit seen so far in the tile
  __shared__ unsigned int tile_start[256];          // exclusive scan of the tile's digit counts
  __shared__ long long goff[256];                   // where this tile's run of each digit starts in the output
  __shared__ long long digit_start[256];            // where each digit starts in the output (exclusive scan of the global histogram)
  __shared__ UK staged[TILE];
  __shared__ unsigned int wsum[kSortWarps];
  __shared__ int sh_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned int lt = (1u << lane) - 1u;
  {
    // exclusive scan of the 256 global digit counts (once per CTA)
    const long long c = (long long)ghist_pass[tid];
    long long incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    __shared__ long long wtot[kSortWarps];
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    long long wbase = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; w++)
      if (w < warp) wbase += wtot[w];
    digit_start[tid] = wbase + incl - c;
  }

  while (true) {
    __syncthreads();
    if (tid == 0) sh_tile = (int)atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) cnt[w][tid] = 0;
    __syncthreads();
    const int tile = sh_tile;
    if (tile >= n_tiles) break;
    const long long base = (long long)tile * TILE;
    const int valid = (int)((n - base < TILE) ? (n - base) : TILE);

    // warp-striped keys: item i of lane l in warp w is tile element w*32*ITEMS + i*32 + l, so the order
    // (warp, item, lane) is the input order and the ranking below is stable
    UK key[ITEMS];
    unsigned int rank[ITEMS];
    int dig[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int e = warp * (32 * ITEMS) + i * 32 + lane;
      const bool ok = e < valid;
      key[i] = ok ? in[base + e] : (UK)0;
      dig[i] = ok ? (int)((order_key<UK, kFloat>(key[i]) >> shift) & 0xff) : -1;
    }
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const bool ok = dig[i] >= 0;
      unsigned int m;
      if constexpr (kMatch) {
        m = __match_any_sync(0xffffffffu, dig[i]);
      } else {
        m = __ballot_sync(0xffffffffu, ok);
#pragma unroll
        for (int b = 0; b < 8; b++) {
          const bool bit = (dig[i] >> b) & 1;
          const unsigned int bal = __ballot_sync(0xffffffffu, bit);
          m &= bit ? bal : ~bal;
        }
      }
      unsigned int before = 0;
      if (ok) before = cnt[warp][dig[i]];
      __syncwarp();
      rank[i] = before + __popc(m & lt);
      if (ok && (m & lt) == 0) cnt[warp][dig[i]] = before + __popc(m);
      __syncwarp();
    }
    __syncthreads();
    // thread d: digit d's counts of the warps -> exclusive over the warps, total of the tile
    unsigned int run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) {
      const unsigned int c = cnt[w][tid];
      cnt[w][tid] = run;
      run += c;
    }
    // publish this tile's count of digit d, then chain it to the predecessors'
    unsigned int *mine = desc + (size_t)tile * 256 + tid;
    unsigned long long excl = 0;
    if (tile == 0) {
      st_relaxed_u32(mine, kSweepPrefix | run);
    } else {
      st_relaxed_u32(mine, kSweepAgg | run);
      for (int t = tile - 1;; t--) {
        const unsigned int *pd = desc + (size_t)t * 256 + tid;
        unsigned int d = ld_relaxed_u32(pd);
        while ((d >> 30) == 0) d = ld_relaxed_u32(pd);
        excl += d & kSweepMask;
        if ((d >> 30) == 2) break;
      }
      st_relaxed_u32(mine, kSweepPrefix | (unsigned int)((excl + run) & kSweepMask));
    }
    goff[tid] = digit_start[tid] + (long long)excl;
    // exclusive scan of the 256 digit totals of the tile
    unsigned int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned int wbase = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; w++)
      if (w < warp) wbase += wsum[w];
    tile_start[tid] = wbase + incl - run;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; i++)
      if (dig[i] >= 0) staged[tile_start[dig[i]] + cnt[warp][dig[i]] + rank[i]] = key[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int pos = i * kSortBlock + tid;
      if (pos < valid) {
        const UK k = staged[pos];
        const int d = (int)((order_key<UK, kFloat>(k) >> shift) & 0xff);
        out[goff[d] + (long long)(pos - (int)tile_start[d])] = k;
      }
    }
  }
}

// Measured on B200, 2^28 keys (tools/sort_bench.py): count + scatter 13.5 Gkeys/s (i64) / 29.6 (i32); one sweep
// 12.3 / 25.5; one sweep with match.any 10.4 / 22.0.  With 2048-key tiles the per-tile look-back (256 chains, 1 KB of
// descriptors per tile, a 128 MB memset per position) costs more than the counting pass it removes (13 % of a
// position), and the ranking — not memory — bounds the scatter either way.  The counting form stays the default.
static int g_sort_algo = 0;          // 0: count + scatter per position (default); 1: one sweep per position; 2: one sweep with match.any
void set_sort_algo(int a) { g_sort_algo = a; }

template <typename UK, int kFloat>
static int sort_impl(cudaStream_t st, int sm_count, void *keys_v, void *tmp_v, int64_t n) {
  constexpr int P = sizeof(UK);
  // keys per thread and tile in the scatter: 16 KB of keys per tile for either key width
  constexpr int ITEMS = (sizeof(UK) == 4) ? NL_SORT_ITEMS32 : kSortItems;
  constexpr int TILE = kSortBlock * ITEMS;
  UK *a = (UK *)keys_v, *b = (UK *)tmp_v;
  int n_chunks = sm_count * NL_SORT_MIN_CTAS * 2;        // two full waves of resident CTAs
  long long chunk = (n + n_chunks - 1) / n_chunks;
  chunk = (chunk + TILE - 1) / TILE * TILE;
  n_chunks = (int)((n + chunk - 1) / chunk);

  unsigned long long *ghist = nullptr;
  unsigned int *hist = nullptr;
  long long *offs = nullptr;
  CUDA_TRY(cudaMallocAsync((void **)&ghist, sizeof(unsigned long long) * P * 256, st));
  CUDA_TRY(cudaMallocAsync((void **)&hist, sizeof(unsigned int) * 256 * (size_t)n_chunks, st));
  CUDA_TRY(cudaMallocAsync((void **)&offs, sizeof(long long) * 256 * (size_t)n_chunks, st));
  // one-sweep scratch: the ticket (first 256 bytes) + one descriptor per tile and digit, re-zeroed per position
  const long long n_tiles_sweep = (n + TILE - 1) / TILE;
  const size_t sweep_bytes = 256 + (size_t)n_tiles_sweep * 256 * sizeof(unsigned int);
  unsigned int *sweep = nullptr;
  CUDA_TRY(cudaMallocAsync((void **)&sweep, sweep_bytes, st));
  CUDA_TRY(cudaMemsetAsync(ghist, 0, sizeof(unsigned long long) * P * 256, st));
  sort_histogram_kernel<UK, kFloat><<<sm_count * 8, kSortBlock, 0, st>>>(a, n, ghist);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  unsigned long long host[P * 256];
  CUDA_TRY(cudaMemcpyAsync(host, ghist, sizeof(host), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));

  UK *src = a, *dst = b;
  for (int p = 0; p < P; p++) {
    bool trivial = false;
    for (int d = 0; d < 256; d++)
      if (host[p * 256 + d] == (unsigned long long)n) trivial = true;   // every key has the same digit here
    if (trivial) continue;
    if (g_sort_algo != 0) {
      CUDA_TRY(cudaMemsetAsync(sweep, 0, sweep_bytes, st));
      const int grid = (int)((long long)sm_count * NL_SORT_MIN_CTAS < n_tiles_sweep ? (long long)sm_count * NL_SORT_MIN_CTAS : n_tiles_sweep);
      if (g_sort_algo == 2)
        onesweep_kernel<UK, kFloat, ITEMS, true><<<grid, kSortBlock, 0, st>>>(src, dst, n, 8 * p, ghist + p * 256, sweep + 64, sweep, (int)n_tiles_sweep);
      else
        onesweep_kernel<UK, kFloat, ITEMS, false><<<grid, kSortBlock, 0, st>>>(src, dst, n, 8 * p, ghist + p * 256, sweep + 64, sweep, (int)n_tiles_sweep);
      CUDA_TRY(cudaGetLastError());
      count_launch(1);
    } else {
      chunk_histogram_kernel<UK, kFloat><<<n_chunks, kSortBlock, 0, st>>>(src, n, chunk, 8 * p, hist);
      chunk_offsets_kernel<<<256, kSortBlock, 0, st>>>(ghist + p * 256, hist, n_chunks, offs);
      scatter_kernel<UK, kFloat, ITEMS><<<n_chunks, kSortBlock, 0, st>>>(src, dst, n, chunk, 8 * p, offs);
      CUDA_TRY(cudaGetLastError());
      count_launch(3);
    }
    UK *t = src; src = dst; dst = t;
  }
  if (src != a) CUDA_TRY(cudaMemcpyAsync(a, src, sizeof(UK) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaFreeAsync(sweep, st));
  CUDA_TRY(cudaFreeAsync(ghist, st));
  CUDA_TRY(cudaFreeAsync(hist, st));
  CUDA_TRY(cudaFreeAsync(offs, st));
  return NL_OK;
}

int launch_sort(void *stream, int sm_count, int dtype, void *keys, void *tmp, int64_t n) {
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    // (key kind: 0 signed integer, 1 float, 2 unsigned integer)
    case NL_I32: return sort_impl<uint32_t, 0>(st, sm_count, keys, tmp, n);
    case NL_I64: return sort_impl<unsigned long long, 0>(st, sm_count, keys, tmp, n);
    case NL_F32: return sort_impl<uint32_t, 1>(st, sm_count, keys, tmp, n);
    case NL_F64: return sort_impl<unsigned long long, 1>(st, sm_count, keys, tmp, n);
    case NL_I8: return sort_impl<uint8_t, 0>(st, sm_count, keys, tmp, n);
    case NL_I16: return sort_impl<uint16_t, 0>(st, sm_count, keys, tmp, n);
    case NL_U8: return sort_impl<uint8_t, 2>(st, sm_count, keys, tmp, n);
    case NL_U16: return sort_impl<uint16_t, 2>(st, sm_count, keys, tmp, n);
    case NL_U32: return sort_impl<uint32_t, 2>(st, sm_count, keys, tmp, n);
    case NL_U64: return sort_impl<unsigned long long, 2>(st, sm_count, keys, tmp, n);
    default: set_error("nl_sort: unsupported dtype %d", dtype); return NL_ERR_UNSUPPORTED;
  }
}

}  // namespace nl
