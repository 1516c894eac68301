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
      tile_start[tid] = wbase + incl - run;
    }
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
    __syncthreads();
    goff[tid] += (long long)tile_count[tid];
  }
}

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
    chunk_histogram_kernel<UK, kFloat><<<n_chunks, kSortBlock, 0, st>>>(src, n, chunk, 8 * p, hist);
    chunk_offsets_kernel<<<256, kSortBlock, 0, st>>>(ghist + p * 256, hist, n_chunks, offs);
    scatter_kernel<UK, kFloat, ITEMS><<<n_chunks, kSortBlock, 0, st>>>(src, dst, n, chunk, 8 * p, offs);
    CUDA_TRY(cudaGetLastError());
    count_launch(3);
    UK *t = src; src = dst; dst = t;
  }
  if (src != a) CUDA_TRY(cudaMemcpyAsync(a, src, sizeof(UK) * (size_t)n, cudaMemcpyDeviceToDevice, st));
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
