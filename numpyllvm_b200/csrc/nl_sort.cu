// nl_sort.cu — device radix sort for the `sort()` full pipeline breaker (sm_100a).
//
// The reference sorts with NumPy on the host (thunk_methods.cpp:37-62, the `base` function of a
// fullbreaker, executed by compiler.cpp:93-122); here the shard is sorted where it lives.
// Least-significant-digit radix sort, 8-bit digits, keys only, stable:
//
//   sort_histogram_kernel   one read of the keys: the 256-bin histogram of EVERY digit position.
//                           A position where all keys share one digit needs no pass at all
//                           (the reference's test data, ints in [1,100), sorts in ONE pass).
//   per remaining position  onesweep_kernel (default): the keys are read once from HBM and written once — a tile's
//                           digit counts are chained to its predecessors' with a decoupled look-back; see the
//                           comment above the kernel for the form that won the measurements (16 B per 8-byte key).
//                           nl_tune("sort_algo", 1): chunk_histogram_kernel (per-CTA-chunk digit counts) +
//                           chunk_offsets_kernel (where each chunk's run of each digit starts) + scatter_kernel
//                           (in-tile ranking, keys regrouped by digit in shared memory) — round 1's form, 24 B.
//   nl_sort_sample / nl_sort_split   the device steps of the sort across GPUs (csrc/jit/device_sort.cpp): regularly
//                           spaced samples, and one sweep of the same kernel whose "digit" is the number of splitters
//                           below the key.
//
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

// the float order key back to the float's bit pattern
template <typename UK>
__device__ __forceinline__ UK float_bits_of_order_key(UK ordered) {
  constexpr int B = sizeof(UK) * 8;
  constexpr UK SIGN = (UK)1 << (B - 1);
  constexpr UK NEG_INF = (sizeof(UK) == 8) ? (UK)0x000FFFFFFFFFFFFFull : (UK)0x007FFFFFu;
  const UK k = (UK)(ordered + NEG_INF);
  return (k & SIGN) ? (UK)(k ^ SIGN) : (UK)~k;
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

// ---- one sweep per digit position (4- and 8-byte keys) ---------------------------------------------------------
// Every executed position reads the keys once from HBM and writes them once: 16 B per 8-byte key instead of the 24 B of
// count + scatter.  A tile's 256 digit counts are chained to its predecessors' with a decoupled look-back (one 32-bit
// descriptor per tile and digit: 2 status bits + 30 count bits; tiles are handed out by a ticket, so every predecessor
// belongs to a running CTA).  What tools/sort_lab.cu measured on B200 (2^28 keys, ms per position, 64-bit keys):
//   * keys held in registers from the load to the regrouping: 1.59-1.86 whatever the ranking (eight ballots, shared
//     atomicOr, match.any 2.56) and whatever the tile (2048..8192 keys): 80 registers, 24 warps per SM, and a CTA that
//     waits 2-3 us for the look-back of a 4096-key tile with nothing else to do (the same kernel with the look-back
//     faked: 1.34);
//   * THIS form: no key stays in a register across a barrier.  The tile is read twice — once to count (digit ->
//     shared atomic on the warp's histogram), and, when the counts are public and turned into running slots of the
//     regrouped tile, again (from L2, batches of CH keys in flight) to rank: eight ballots find the lanes with the same
//     digit, one LDS + one STS per key advance the warp's slot, and the key goes straight to its place in shared
//     memory.  64-77 registers with 6144/8192-key tiles -> 1.18 ms (64-bit, 26 Gkeys/s for a full sort; CUB's
//     DeviceRadixSort 12.2) and 0.84 ms (32-bit, 71 Gkeys/s; CUB 45.5).
//   * a look-back window of 4 tiles per round trip is best (1: 3.2 ms, 2: 2.0, 8: +5 %, 16: +25 %).
// The ballots are PTX with immediate masks: ptxas then moves the eight digit bits into predicates with one R2P and
// spends VOTE + 2 LOP3 per bit (25 instructions per key; the C++ form compiled to 6 per bit).
constexpr unsigned int kSweepAgg = 1u << 30, kSweepPrefix = 2u << 30, kSweepMask = (1u << 30) - 1u;
constexpr int kSweepWindow = 4;
static long long g_sweep_portion = 1ll << 28;     // keys per launch: a tile's running digit offsets stay below 2^30 (tests shrink it)
void set_sort_portion_log2(int l) { g_sweep_portion = 1ll << l; }

__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned int *p, unsigned int v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int B>
__device__ __forceinline__ unsigned int same_bit(unsigned int d) {
  unsigned int x;
  asm("{\n\t.reg .pred p;\n\t.reg .b32 t, bal, s;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\tvote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
      "selp.b32 s, 0, 0xffffffff, p;\n\txor.b32 %0, bal, s;\n\t}"
      : "=r"(x) : "r"(d), "n"(1u << B));
  return x;
}
// lanes of the (converged) warp whose 8-bit digit equals mine
__device__ __forceinline__ unsigned int match_digit(unsigned int d) {
  return same_bit<0>(d) & same_bit<1>(d) & same_bit<2>(d) & same_bit<3>(d) & same_bit<4>(d) & same_bit<5>(d) & same_bit<6>(d) & same_bit<7>(d);
}

// What a sweep groups the keys by: one 8-bit digit of the order key (the radix passes), or the number of splitters
// below the key (the multi-GPU sample sort cuts a shard into one run per destination GPU with the same kernel).
template <typename UK>
struct RadixDigit {
  int shift;
  __device__ __forceinline__ unsigned int operator()(UK ordered) const { return (unsigned int)(ordered >> shift) & 0xff; }
};
constexpr int kMaxSplitters = 15;
template <typename UK>
struct SplitDigit {
  UK s[kMaxSplitters];           // ascending order keys; bucket j holds the keys in (s[j-1], s[j]]
  int m;
  __device__ __forceinline__ unsigned int operator()(UK ordered) const {
    unsigned int d = 0;
    // m is uniform and the splitters sit in the kernel's parameter bank: a real loop of m compares (a sort over two
    // GPUs pays one compare per key, not fifteen predicated ones)
#pragma unroll 1
    for (int j = 0; j < m; j++) d += (s[j] < ordered) ? 1u : 0u;
    return d;
  }
};

struct SweepShared {             // views into the CTA's dynamic shared memory
  void *staged;                  // the tile regrouped by digit
  unsigned int (*cnt)[256];      // per warp: digit counts, then the warp's next slot of each digit
  long long *gdelta;             // per digit: (start of the tile's run in the output) - (start of the run in the tile)
  unsigned int *wsum;
  int *next_tile;
};

// kIn: how a loaded key becomes its order key (0 signed integer, 1 float, 2 it already is one).  kOut: what is written —
// 0 the loaded bits, 1 the order key, 2 the float bits of a loaded order key.  A float sort of several positions
// converts its keys in the first sweep and back in the last one; the sweeps between them run as unsigned integers.
template <typename UK, int kIn, int kOut, int ITEMS, int CH, bool FULL, typename Digit>
__device__ __forceinline__ void sweep_tile(const UK *__restrict__ in, UK *__restrict__ out, unsigned int n, const Digit &digit, long long dstart,
                                           unsigned long long *__restrict__ bins_out, unsigned int *__restrict__ desc, unsigned int *__restrict__ ticket,
                                           int tile, int n_tiles, const SweepShared &sh) {
  constexpr int WARPS = kSortWarps, TILE = kSortBlock * ITEMS, W = kSweepWindow;
  static_assert(ITEMS % CH == 0, "the tile is read in whole batches");
  UK *staged = (UK *)sh.staged;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned int lt = (1u << lane) - 1u;
  const unsigned int base = (unsigned int)tile * TILE;
  const int valid = FULL ? TILE : (int)(n - base);
  unsigned int *cw = &sh.cnt[warp][0];
  // warp-striped keys: item i of lane l in warp w is tile element w*32*ITEMS + i*32 + l, so the order (warp, item, lane)
  // is the input order and the ranking below is stable
  const UK *src = in + base + warp * (32 * ITEMS) + lane;
#pragma unroll
  for (int i0 = 0; i0 < ITEMS; i0 += CH) {
    UK key[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (FULL) key[i] = __ldcg(src + (i0 + i) * 32);      // stays in L2 for the second read
      else key[i] = (warp * (32 * ITEMS) + (i0 + i) * 32 + lane < valid) ? src[(i0 + i) * 32] : (UK)0;
    }
#pragma unroll
    for (int i = 0; i < CH; i++) {
      const unsigned int d = digit(order_key<UK, kIn>(key[i]));
      if (FULL || (warp * (32 * ITEMS) + (i0 + i) * 32 + lane < valid)) atomicAdd(&cw[d], 1u);
    }
  }
  __syncthreads();
  // thread d: digit d's counts of the warps; the tile's count goes public before anything else happens
  unsigned int run = 0, excl, c[WARPS];
  unsigned int *mine = desc + (size_t)tile * 256 + tid;
#pragma unroll
  for (int w = 0; w < WARPS; w++) c[w] = sh.cnt[w][tid];
#pragma unroll
  for (int w = 0; w < WARPS; w++) run += c[w];
  if (tile > 0) st_relaxed_u32(mine, kSweepAgg | run);
  {
    unsigned int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) sh.wsum[warp] = incl;
    excl = incl - run;
  }
  __syncthreads();
  {
    unsigned int wbase = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++)
      if (w < warp) wbase += sh.wsum[w];
    excl += wbase;                       // where digit tid starts in the regrouped tile
    unsigned int slot = excl;
#pragma unroll
    for (int w = 0; w < WARPS; w++) {    // the counters become the warps' next free slot of the digit
      sh.cnt[w][tid] = slot;
      slot += c[w];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i0 = 0; i0 < ITEMS; i0 += CH) {
    UK key[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (FULL) key[i] = __ldcs(src + (i0 + i) * 32);
      else key[i] = (warp * (32 * ITEMS) + (i0 + i) * 32 + lane < valid) ? src[(i0 + i) * 32] : (UK)0;
    }
#pragma unroll
    for (int i = 0; i < CH; i++) {
      const UK ordered = order_key<UK, kIn>(key[i]);
      const unsigned int d = digit(ordered);
      const bool ok = FULL || (warp * (32 * ITEMS) + (i0 + i) * 32 + lane < valid);
      unsigned int m = match_digit(d);
      if (!FULL) m &= __ballot_sync(0xffffffffu, ok);
      unsigned int before = 0;
      if (ok) before = cw[d];
      __syncwarp();
      if (ok && (m & lt) == 0) cw[d] = before + __popc(m);
      if (ok) staged[before + __popc(m & lt)] = (kOut == 1) ? ordered : key[i];
      __syncwarp();
    }
  }
  if (tid == 0) *sh.next_tile = (int)atomicAdd(ticket, 1u);
  {
    // chain the tile's counts to its predecessors': W descriptors per round trip
    unsigned long long before = 0;
    if (tile > 0) {
      int t = tile - 1;
      while (true) {
        unsigned int look[W];
#pragma unroll
        for (int j = 0; j < W; j++) look[j] = (t - j >= 0) ? ld_relaxed_u32(desc + (size_t)(t - j) * 256 + tid) : kSweepPrefix;
        bool done = false;
#pragma unroll
        for (int j = 0; j < W; j++) {
          if (!done) {
            unsigned int v = look[j];
            while ((v >> 30) == 0) v = ld_relaxed_u32(desc + (size_t)(t - j) * 256 + tid);
            before += v & kSweepMask;
            if ((v >> 30) == 2) done = true;
          }
        }
        if (done) break;
        t -= W;
      }
    }
    st_relaxed_u32(mine, kSweepPrefix | (unsigned int)(before + run));
    sh.gdelta[tid] = dstart + (long long)before - (long long)excl;
    if (tile == n_tiles - 1 && bins_out) bins_out[tid] = (unsigned long long)(dstart + (long long)before + (long long)run);   // where the next launch continues
  }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < WARPS; w++) sh.cnt[w][tid] = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const int pos = i * kSortBlock + tid;
    if (FULL || pos < valid) {
      const UK k = staged[pos];
      const unsigned int d = digit((kOut == 1) ? k : order_key<UK, kIn>(k));
      out[sh.gdelta[d] + (long long)pos] = (kOut == 2) ? float_bits_of_order_key<UK>(k) : k;
    }
  }
}

template <typename UK, int kIn, int kOut, int ITEMS, int CH, typename Digit>
__device__ __noinline__ void sweep_partial_tile(const UK *__restrict__ in, UK *__restrict__ out, unsigned int n, const Digit &digit, long long dstart,
                                                unsigned long long *__restrict__ bins_out, unsigned int *__restrict__ desc, unsigned int *__restrict__ ticket,
                                                int tile, int n_tiles, const SweepShared &sh) {
  sweep_tile<UK, kIn, kOut, ITEMS, CH, false>(in, out, n, digit, dstart, bins_out, desc, ticket, tile, n_tiles, sh);
}

template <typename UK, int kIn, int kOut, int ITEMS, int CH, int MINCTAS, typename Digit>
__global__ void __launch_bounds__(kSortBlock, MINCTAS) onesweep_kernel(const UK *__restrict__ in, UK *__restrict__ out, unsigned int n, const __grid_constant__ Digit digit,
                                                                       const unsigned long long *__restrict__ bins_in /*[256]: where each digit continues in out*/,
                                                                       unsigned long long *__restrict__ bins_out /*[256] for the next launch, or null*/,
                                                                       unsigned int *__restrict__ desc /*[n_tiles][256], zeroed*/,
                                                                       unsigned int *__restrict__ ticket /*zeroed*/, int n_tiles) {
  constexpr int TILE = kSortBlock * ITEMS;
  extern __shared__ __align__(16) unsigned char sweep_smem[];
  __shared__ int sh_tile;
  SweepShared sh;
  sh.staged = sweep_smem;
  sh.cnt = (unsigned int(*)[256])((UK *)sweep_smem + TILE);
  sh.gdelta = (long long *)(sh.cnt + kSortWarps);
  sh.wsum = (unsigned int *)(sh.gdelta + 256);
  sh.next_tile = &sh_tile;
  const int tid = threadIdx.x;
  const long long dstart = (long long)bins_in[tid];
#pragma unroll
  for (int w = 0; w < kSortWarps; w++) sh.cnt[w][tid] = 0;
  if (tid == 0) sh_tile = (int)atomicAdd(ticket, 1u);
  while (true) {
    __syncthreads();
    const int tile = sh_tile;
    if (tile >= n_tiles) break;
    if ((unsigned int)(tile + 1) * TILE <= n)
      sweep_tile<UK, kIn, kOut, ITEMS, CH, true>(in, out, n, digit, dstart, bins_out, desc, ticket, tile, n_tiles, sh);
    else
      sweep_partial_tile<UK, kIn, kOut, ITEMS, CH>(in, out, n, digit, dstart, bins_out, desc, ticket, tile, n_tiles, sh);
  }
}

// one CTA per digit position: where each digit starts in the output (exclusive scan of the position's histogram)
__global__ void __launch_bounds__(kSortBlock) digit_starts_kernel(const unsigned long long *__restrict__ ghist, unsigned long long *__restrict__ bins /*[P][2][256]*/) {
  __shared__ unsigned long long part[kSortWarps];
  const int p = blockIdx.x, d = threadIdx.x, lane = d & 31, warp = d >> 5;
  const unsigned long long c = ghist[p * 256 + d];
  unsigned long long incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) part[warp] = incl;
  __syncthreads();
  unsigned long long wbase = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; w++)
    if (w < warp) wbase += part[w];
  bins[(size_t)p * 512 + d] = wbase + incl - c;
}

template <int BYTES> struct SweepShape;                  // keys per thread, keys per batch of the second read, resident CTAs per SM
template <> struct SweepShape<8> { static constexpr int items = 24, batch = 8, ctas = 3; };
template <> struct SweepShape<4> { static constexpr int items = 32, batch = 8, ctas = 4; };
template <> struct SweepShape<2> { static constexpr int items = 32, batch = 8, ctas = 4; };
template <> struct SweepShape<1> { static constexpr int items = 32, batch = 8, ctas = 4; };

static int g_sort_algo = 0;          // 0: one sweep per position (default); 1: count + scatter per position
void set_sort_algo(int a) { g_sort_algo = a; }

// one sweep of `n` keys src -> dst, grouped by `digit`, in launches of at most one portion; bins[0..255] = where each
// group starts in dst (bins[256..511] is the second half of the double buffer the launches hand over through)
template <typename UK, int kIn, int kOut, typename Digit>
static int run_sweep(cudaStream_t st, int sm_count, const UK *src, UK *dst, int64_t n, const Digit &digit, unsigned long long *bins, unsigned int **scratch_io,
                     size_t *scratch_bytes_io) {
  using Shape = SweepShape<sizeof(UK)>;
  constexpr int ITEMS = Shape::items, TILE = kSortBlock * ITEMS;
  auto kern = onesweep_kernel<UK, kIn, kOut, ITEMS, Shape::batch, Shape::ctas, Digit>;
  const size_t smem = sizeof(UK) * TILE + sizeof(unsigned int) * 256 * kSortWarps + sizeof(long long) * 256 + 64;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kSortBlock, smem));
  if (occ < 1) { set_error("nl_sort: the one-sweep kernel does not fit an SM"); return NL_ERR_CUDA; }
  const long long portion = (g_sweep_portion / TILE > 0 ? g_sweep_portion / TILE : 1) * TILE;
  const long long max_tiles = ((n < portion ? n : portion) + TILE - 1) / TILE;
  const size_t need = 256 + (size_t)max_tiles * 256 * sizeof(unsigned int);
  if (*scratch_bytes_io < need) {
    if (*scratch_io) CUDA_TRY(cudaFreeAsync(*scratch_io, st));
    *scratch_io = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)scratch_io, need, st));
    *scratch_bytes_io = need;
  }
  unsigned int *scratch = *scratch_io;
  int q = 0;
  for (long long lo = 0; lo < n; lo += portion, q++) {
    const long long len = (n - lo < portion) ? (n - lo) : portion;
    const int n_tiles = (int)((len + TILE - 1) / TILE);
    CUDA_TRY(cudaMemsetAsync(scratch, 0, 256 + (size_t)n_tiles * 256 * sizeof(unsigned int), st));
    const int grid = (int)((long long)sm_count * occ < n_tiles ? (long long)sm_count * occ : n_tiles);
    unsigned long long *b_in = bins + (q & 1) * 256, *b_out = bins + ((q + 1) & 1) * 256;
    kern<<<grid, kSortBlock, smem, st>>>(src + lo, dst, (unsigned int)len, digit, b_in, (lo + len < n) ? b_out : nullptr, scratch + 64, scratch, n_tiles);
    CUDA_TRY(cudaGetLastError());
    count_launch();
  }
  return NL_OK;
}

// all non-trivial positions of a key, one sweep each
template <typename UK, int kFloat>
static int onesweep_passes(cudaStream_t st, int sm_count, UK *&src, UK *&dst, int64_t n, const unsigned long long *ghist, const unsigned long long *host_hist) {
  constexpr int P = sizeof(UK);
  unsigned int *scratch = nullptr;
  size_t scratch_bytes = 0;
  unsigned long long *bins = nullptr;
  CUDA_TRY(cudaMallocAsync((void **)&bins, sizeof(unsigned long long) * P * 512, st));
  digit_starts_kernel<<<P, kSortBlock, 0, st>>>(ghist, bins);
  count_launch();
  int todo[P], n_todo = 0;
  for (int p = 0; p < P; p++) {
    bool trivial = false;
    for (int d = 0; d < 256; d++)
      if (host_hist[p * 256 + d] == (unsigned long long)n) trivial = true;   // every key has the same digit here
    if (!trivial) todo[n_todo++] = p;
  }
  for (int i = 0; i < n_todo; i++) {
    const int p = todo[i];
    RadixDigit<UK> digit;
    digit.shift = 8 * p;
    unsigned long long *b = bins + (size_t)p * 512;
    int rc = NL_OK;
    bool done = false;
    if constexpr (kFloat == 1) {
      // floats: order keys are made in the first sweep and turned back in the last one
      if (n_todo > 1) {
        if (i == 0) rc = run_sweep<UK, 1, 1>(st, sm_count, src, dst, n, digit, b, &scratch, &scratch_bytes);
        else if (i == n_todo - 1) rc = run_sweep<UK, 2, 2>(st, sm_count, src, dst, n, digit, b, &scratch, &scratch_bytes);
        else rc = run_sweep<UK, 2, 0>(st, sm_count, src, dst, n, digit, b, &scratch, &scratch_bytes);
        done = true;
      }
    }
    if (!done) rc = run_sweep<UK, kFloat, 0>(st, sm_count, src, dst, n, digit, b, &scratch, &scratch_bytes);
    if (rc != NL_OK) return rc;
    UK *t = src; src = dst; dst = t;
  }
  if (scratch) CUDA_TRY(cudaFreeAsync(scratch, st));
  CUDA_TRY(cudaFreeAsync(bins, st));
  return NL_OK;
}

// ---- the two device steps of the multi-GPU sample sort (csrc/jit/device_sort.cpp) -------------------------------
// order keys of k regularly spaced elements, widened to 64 bits
template <typename UK, int kFloat>
__global__ void sample_kernel(const UK *__restrict__ keys, long long n, int k, unsigned long long *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k) out[i] = (unsigned long long)order_key<UK, kFloat>(keys[(long long)(((__int128)n * i) / k)]);
}

// how many keys fall into each of the m + 1 buckets the splitters cut: per-thread counts packed 8 bits a bucket
template <typename UK, int kFloat>
__global__ void __launch_bounds__(kSortBlock) split_count_kernel(const UK *__restrict__ keys, long long n, const __grid_constant__ SplitDigit<UK> digit,
                                                                   unsigned long long *__restrict__ ghist /*[256], zeroed*/) {
  __shared__ unsigned int h[16];
  if (threadIdx.x < 16) h[threadIdx.x] = 0;
  __syncthreads();
  unsigned long long lo = 0, hi = 0;      // buckets 0..7 and 8..15, 8 bits each
  int pending = 0;
  auto flush = [&]() {
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const unsigned int cl = (unsigned int)(lo >> (8 * b)) & 0xff, ch = (unsigned int)(hi >> (8 * b)) & 0xff;
      if (cl) atomicAdd(&h[b], cl);
      if (ch) atomicAdd(&h[8 + b], ch);
    }
    lo = hi = 0;
    pending = 0;
  };
  for (long long i = (long long)blockIdx.x * kSortBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kSortBlock) {
    const unsigned int d = digit(order_key<UK, kFloat>(keys[i]));
    const unsigned long long one = 1ull << (8 * (d & 7));
    if (d < 8) lo += one; else hi += one;
    if (++pending == 255) flush();
  }
  flush();
  __syncthreads();
  if (threadIdx.x < 16 && h[threadIdx.x]) atomicAdd(&ghist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

template <typename UK, int kFloat>
static int sample_impl(cudaStream_t st, const void *keys, int64_t n, int k, unsigned long long *host_out) {
  unsigned long long *dev = nullptr;
  CUDA_TRY(cudaMallocAsync((void **)&dev, sizeof(unsigned long long) * (size_t)k, st));
  sample_kernel<UK, kFloat><<<(k + 255) / 256, 256, 0, st>>>((const UK *)keys, n, k, dev);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  CUDA_TRY(cudaMemcpyAsync(host_out, dev, sizeof(unsigned long long) * (size_t)k, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaFreeAsync(dev, st));
  return NL_OK;
}

template <typename UK>
static SplitDigit<UK> make_split_digit(const unsigned long long *splitters, int m) {
  SplitDigit<UK> digit;
  for (int j = 0; j < kMaxSplitters; j++) digit.s[j] = (j < m) ? (UK)splitters[j] : (UK)0;
  digit.m = m;
  return digit;
}

// mode 0: count + sweep into `out` (runs back to back); 1: count only; 2: sweep only, run j goes to run_dst[j]
template <typename UK, int kFloat>
static int split_impl(cudaStream_t st, int sm_count, int mode, const void *keys, void *out, int64_t n, const unsigned long long *splitters, int m,
                      int64_t *host_counts, void *const *run_dst) {
  const SplitDigit<UK> digit = make_split_digit<UK>(splitters, m);
  unsigned long long *ghist = nullptr, *bins = nullptr;
  CUDA_TRY(cudaMallocAsync((void **)&ghist, sizeof(unsigned long long) * 256, st));
  CUDA_TRY(cudaMallocAsync((void **)&bins, sizeof(unsigned long long) * 512, st));
  if (mode != 2) {
    CUDA_TRY(cudaMemsetAsync(ghist, 0, sizeof(unsigned long long) * 256, st));
    split_count_kernel<UK, kFloat><<<sm_count * 8, kSortBlock, 0, st>>>((const UK *)keys, n, digit, ghist);
    CUDA_TRY(cudaGetLastError());
    count_launch();
    static_assert(sizeof(int64_t) == sizeof(unsigned long long), "the counts are copied as they are");
    CUDA_TRY(cudaMemcpyAsync(host_counts, ghist, sizeof(unsigned long long) * (size_t)(m + 1), cudaMemcpyDeviceToHost, st));
  }
  if (mode == 0) {
    digit_starts_kernel<<<1, kSortBlock, 0, st>>>(ghist, bins);
    CUDA_TRY(cudaGetLastError());
    count_launch();
  } else if (mode == 2) {
    // where each run starts, in elements relative to run 0's destination: the sweep adds a tile's offset inside its
    // run and stores through `out + offset` — whichever GPU that address belongs to (one unified address space)
    unsigned long long rel[512];
    memset(rel, 0, sizeof(rel));
    void *base = nullptr;
    for (int j = 0; j <= m && !base; j++) base = run_dst[j];
    if (!base) { set_error("nl_sort_split_scatter: every run_dst is null"); return NL_ERR_INVALID; }
    for (int j = 0; j <= m; j++)
      if (run_dst[j]) rel[j] = (unsigned long long)(((long long)(intptr_t)run_dst[j] - (long long)(intptr_t)base) / (long long)sizeof(UK));
    CUDA_TRY(cudaMemcpyAsync(bins, rel, sizeof(rel), cudaMemcpyHostToDevice, st));   // pageable source: consumed when the call returns
    out = base;
  }
  if (mode != 1) {
    unsigned int *scratch = nullptr;
    size_t scratch_bytes = 0;
    const int rc = run_sweep<UK, kFloat, 0>(st, sm_count, (const UK *)keys, (UK *)out, n, digit, bins, &scratch, &scratch_bytes);
    if (rc != NL_OK) return rc;
    if (scratch) CUDA_TRY(cudaFreeAsync(scratch, st));
  }
  CUDA_TRY(cudaFreeAsync(ghist, st));
  CUDA_TRY(cudaFreeAsync(bins, st));
  return NL_OK;
}

template <typename UK, int kFloat>
static int sort_impl(cudaStream_t st, int sm_count, void *keys_v, void *tmp_v, int64_t n) {
  constexpr int P = sizeof(UK);
  UK *a = (UK *)keys_v, *b = (UK *)tmp_v;
  if (n <= 1) return NL_OK;
  unsigned long long *ghist = nullptr;
  CUDA_TRY(cudaMallocAsync((void **)&ghist, sizeof(unsigned long long) * P * 256, st));
  CUDA_TRY(cudaMemsetAsync(ghist, 0, sizeof(unsigned long long) * P * 256, st));
  sort_histogram_kernel<UK, kFloat><<<sm_count * 8, kSortBlock, 0, st>>>(a, n, ghist);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  unsigned long long host[P * 256];
  CUDA_TRY(cudaMemcpyAsync(host, ghist, sizeof(host), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));

  UK *src = a, *dst = b;
  if (g_sort_algo == 0) {
    const int rc = onesweep_passes<UK, kFloat>(st, sm_count, src, dst, n, ghist, host);
    if (rc != NL_OK) return rc;
    if (src != a) CUDA_TRY(cudaMemcpyAsync(a, src, sizeof(UK) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaFreeAsync(ghist, st));
    return NL_OK;
  }
  // count + scatter per position (nl_tune("sort_algo", 1))
  // keys per thread and tile in the scatter: 16 KB of keys per tile for either key width
  constexpr int ITEMS = (sizeof(UK) == 4) ? NL_SORT_ITEMS32 : kSortItems;
  constexpr int TILE = kSortBlock * ITEMS;
  int n_chunks = sm_count * NL_SORT_MIN_CTAS * 2;        // two full waves of resident CTAs
  long long chunk = (n + n_chunks - 1) / n_chunks;
  chunk = (chunk + TILE - 1) / TILE * TILE;
  n_chunks = (int)((n + chunk - 1) / chunk);
  unsigned int *hist = nullptr;
  long long *offs = nullptr;
  CUDA_TRY(cudaMallocAsync((void **)&hist, sizeof(unsigned int) * 256 * (size_t)n_chunks, st));
  CUDA_TRY(cudaMallocAsync((void **)&offs, sizeof(long long) * 256 * (size_t)n_chunks, st));
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

#define NL_SORT_DISPATCH(CALL)                                                              \
  switch (dtype) {                                                                         \
    /* (key kind: 0 signed integer, 1 float, 2 unsigned integer) */                        \
    case NL_I32: return CALL(uint32_t, 0);                                                 \
    case NL_I64: return CALL(unsigned long long, 0);                                       \
    case NL_F32: return CALL(uint32_t, 1);                                                 \
    case NL_F64: return CALL(unsigned long long, 1);                                       \
    case NL_I8: return CALL(uint8_t, 0);                                                   \
    case NL_I16: return CALL(uint16_t, 0);                                                 \
    case NL_U8: return CALL(uint8_t, 2);                                                   \
    case NL_U16: return CALL(uint16_t, 2);                                                 \
    case NL_U32: return CALL(uint32_t, 2);                                                 \
    case NL_U64: return CALL(unsigned long long, 2);                                       \
    default: set_error("nl_sort: unsupported dtype %d", dtype); return NL_ERR_UNSUPPORTED; \
  }

int launch_sort_sample(void *stream, int dtype, const void *keys, int64_t n, int k, unsigned long long *host_out) {
  cudaStream_t st = (cudaStream_t)stream;
#define NL_CALL(UK, KF) sample_impl<UK, KF>(st, keys, n, k, host_out)
  NL_SORT_DISPATCH(NL_CALL)
#undef NL_CALL
}

int launch_sort_split(void *stream, int sm_count, int mode, int dtype, const void *keys, void *out, int64_t n, const unsigned long long *splitters, int m,
                      int64_t *host_counts, void *const *run_dst) {
  cudaStream_t st = (cudaStream_t)stream;
#define NL_CALL(UK, KF) split_impl<UK, KF>(st, sm_count, mode, keys, out, n, splitters, m, host_counts, run_dst)
  NL_SORT_DISPATCH(NL_CALL)
#undef NL_CALL
}

}  // namespace nl
