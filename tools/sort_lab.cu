// sort_lab.cu — which form of the one-sweep radix pass is fastest on a B200?  (Design tool behind nl_sort.cu.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/sort_lab tools/sort_lab.cu
//   build/sort_lab [log2_keys=28] [variant filter]
//
// One kernel template — tile ticket, in-tile ranking, decoupled look-back over per-tile digit counts, regrouping in
// shared memory — swept over block size, keys per thread, resident CTAs, the way lanes with equal digits find each
// other (ballots / match.any / shared atomicOr) and the look-back's read width.  Every variant sorts 2^k random keys
// in 4 (32-bit) or 8 (64-bit) passes; the result is checked (sorted, same multiset checksums).  CUB's
// DeviceRadixSort::SortKeys is timed beside them as the library yardstick.  Test tooling only.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr unsigned kAgg = 1u << 30, kPre = 2u << 30, kMask = (1u << 30) - 1u;

__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(unsigned *p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__device__ __forceinline__ uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
template <typename UK>
__global__ void fill(UK *k, long long n, int mode) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint64_t r = splitmix(42 ^ (uint64_t)i);
    if (mode == 1) r = 1 + r % 99;
    k[i] = (UK)r;
  }
}
template <typename UK>
__global__ void check(const UK *k, long long n, unsigned long long *res /*inversions, sum, xor*/) {
  unsigned long long inv = 0, s = 0, x = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (i + 1 < n && k[i] > k[i + 1]) inv++;
    s += (unsigned long long)k[i];
    x ^= (unsigned long long)k[i] * 0x9E3779B97F4A7C15ull;
  }
  if (inv) atomicAdd(res, inv);
  atomicAdd(res + 1, s);
  atomicXor(res + 2, x);
}

// all digit positions in one read; then bins[p][d] = exclusive scan over d
template <typename UK>
__global__ void __launch_bounds__(256) hist_all(const UK *__restrict__ keys, long long n, unsigned long long *__restrict__ gh) {
  constexpr int P = sizeof(UK);
  __shared__ unsigned h[P][256];
  for (int i = threadIdx.x; i < P * 256; i += 256) (&h[0][0])[i] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const UK k = keys[i];
#pragma unroll
    for (int p = 0; p < P; p++) atomicAdd(&h[p][(k >> (8 * p)) & 0xff], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P * 256; i += 256) {
    const unsigned c = (&h[0][0])[i];
    if (c) atomicAdd(&gh[i], (unsigned long long)c);
  }
}
__global__ void scan_bins(const unsigned long long *gh, unsigned long long *bins) {   // one CTA of 256 per position
  __shared__ unsigned long long s[256];
  const int p = blockIdx.x, d = threadIdx.x;
  s[d] = gh[p * 256 + d];
  __syncthreads();
  if (d == 0) {
    unsigned long long run = 0;
    for (int j = 0; j < 256; j++) { const unsigned long long c = s[j]; s[j] = run; run += c; }
  }
  __syncthreads();
  bins[p * 256 + d] = s[d];
}

// MATCH: 0 eight ballots, 1 match.any, 2 shared atomicOr, 3 none (timing only: wrong result)
template <typename UK, int BLOCK, int ITEMS, int MATCH, int MINCTAS, int LB>
__global__ void __launch_bounds__(BLOCK, MINCTAS) sweep(const UK *__restrict__ in, UK *__restrict__ out, unsigned n, int shift,
                                                        const unsigned long long *__restrict__ bins, unsigned *__restrict__ desc,
                                                        unsigned *__restrict__ ticket, int n_tiles) {
  constexpr int WARPS = BLOCK / 32, TILE = BLOCK * ITEMS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  UK *staged = (UK *)smem_raw;
  unsigned(*cnt)[256] = (unsigned(*)[256])(staged + TILE);
  long long *gdelta = (long long *)(cnt + WARPS);
  unsigned *wsum = (unsigned *)(gdelta + 256);
  unsigned(*mm)[256] = (unsigned(*)[256])(wsum + 8);   // MATCH == 2 only
  __shared__ int sh_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  long long dstart = 0;
  if (tid < 256) dstart = (long long)bins[tid];
  if (MATCH == 2) {
    for (int i = tid; i < WARPS * 256; i += BLOCK) (&mm[0][0])[i] = 0;
  }

  while (true) {
    if (tid == 0) sh_tile = (int)atomicAdd(ticket, 1u);
    for (int i = tid; i < WARPS * 256; i += BLOCK) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int tile = sh_tile;
    if (tile >= n_tiles) break;
    const unsigned base = (unsigned)tile * TILE;
    const int valid = (int)((n - base < (unsigned)TILE) ? (n - base) : TILE);
    const bool full = valid == TILE;

    UK key[ITEMS];
    unsigned rk[(ITEMS + 1) / 2];   // two 16-bit ranks per register
    const UK *src = in + base + warp * (32 * ITEMS) + lane;
    if (full) {
#pragma unroll
      for (int i = 0; i < ITEMS; i++) key[i] = __ldcs(src + i * 32);
    } else {
#pragma unroll
      for (int i = 0; i < ITEMS; i++) key[i] = (warp * (32 * ITEMS) + i * 32 + lane < valid) ? src[i * 32] : (UK) ~(UK)0;
    }
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const unsigned d = (unsigned)(key[i] >> shift) & 0xff;
      const bool ok = full || (warp * (32 * ITEMS) + i * 32 + lane < valid);
      unsigned m;
      if (MATCH == 0) {
        m = full ? 0xffffffffu : __ballot_sync(0xffffffffu, ok);
#pragma unroll
        for (int b = 0; b < 8; b++) {
          const bool bit = (d >> b) & 1;
          const unsigned bal = __ballot_sync(0xffffffffu, bit);
          m &= bit ? bal : ~bal;
        }
      } else if (MATCH == 1) {
        m = __match_any_sync(0xffffffffu, ok ? d : 256u + lane);
      } else if (MATCH == 2) {
        if (ok) atomicOr(&mm[warp][d], 1u << lane);
        __syncwarp();
        m = ok ? mm[warp][d] : (1u << lane);
        __syncwarp();
        if (ok && (m & lt) == 0) mm[warp][d] = 0;
      } else {
        m = 1u << lane;
      }
      unsigned before = 0;
      if (ok) before = cnt[warp][d];
      __syncwarp();
      {
        const unsigned r = before + __popc(m & lt);
        if (i & 1) rk[i / 2] |= r << 16; else rk[i / 2] = r;
      }
      if (ok && (m & lt) == 0) cnt[warp][d] = before + __popc(m);
      __syncwarp();
    }
    __syncthreads();
    // digit thread d: counts of the warps -> exclusive over the warps; publish the tile's count; start the look-back
    unsigned run = 0, excl = 0, look[LB];
    unsigned *mine = desc + (size_t)tile * 256 + tid;
    if (tid < 256) {
#pragma unroll
      for (int w = 0; w < WARPS; w++) {
        const unsigned c = cnt[w][tid];
        cnt[w][tid] = run;
        run += c;
      }
      if (tile > 0) {
        st_relaxed(mine, kAgg | run);
#pragma unroll
        for (int j = 0; j < LB; j++) look[j] = (tile - 1 - j >= 0) ? ld_relaxed(mine - (size_t)(j + 1) * 256) : kPre;
      }
      unsigned incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) wsum[warp] = incl;
      excl = incl - run;   // within the warp, for now
    }
    __syncthreads();
    if (tid < 256) {
      unsigned wbase = 0;
#pragma unroll
      for (int w = 0; w < 8; w++)
        if (w < warp) wbase += wsum[w];
      const unsigned ts = wbase + excl;   // start of digit tid in the regrouped tile
#pragma unroll
      for (int w = 0; w < WARPS; w++) cnt[w][tid] += ts;
      excl = ts;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const unsigned d = (unsigned)(key[i] >> shift) & 0xff;
      const bool ok = full || (warp * (32 * ITEMS) + i * 32 + lane < valid);
      if (ok) staged[cnt[warp][d] + ((i & 1) ? (rk[i / 2] >> 16) : (rk[i / 2] & 0xffffu))] = key[i];
    }
    if (tid < 256) {
      const unsigned ts = excl;
      unsigned long long before = 0;
      if (tile > 0) {
        int t = tile - 1;
        while (true) {
          bool done = false;
#pragma unroll
          for (int j = 0; j < LB; j++) {
            if (!done) {
              unsigned v = look[j];
              while ((v >> 30) == 0) v = ld_relaxed(desc + (size_t)(t - j) * 256 + tid);
              before += v & kMask;
              if ((v >> 30) == 2) done = true;
            }
          }
          if (done) break;
          t -= LB;
#pragma unroll
          for (int j = 0; j < LB; j++) look[j] = (t - j >= 0) ? ld_relaxed(desc + (size_t)(t - j) * 256 + tid) : kPre;
        }
      }
      st_relaxed(mine, kPre | (unsigned)(before + run));
      gdelta[tid] = dstart + (long long)before - (long long)ts;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int pos = i * BLOCK + tid;
      if (full || pos < valid) {
        const UK k = staged[pos];
        const unsigned d = (unsigned)(k >> shift) & 0xff;
        out[gdelta[d] + (long long)pos] = k;
      }
    }
    __syncthreads();
  }
}


// ---- second form: full tiles without validity predicates, ballots written out, warp counters touched once ----
// lanes of the warp whose 8-bit digit equals mine: one ballot per digit bit.  Written in PTX with immediate masks so that
// ptxas moves the eight bits into predicates with ONE R2P and spends VOTE + 2 LOP3 per bit (25 instructions per key;
// the C++ form compiles to 6 per bit).
template <int B>
__device__ __forceinline__ unsigned same_bit(unsigned d) {
  unsigned x;
  asm("{\n\t.reg .pred p;\n\t.reg .b32 t, bal, s;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\tvote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
      "selp.b32 s, 0, 0xffffffff, p;\n\txor.b32 %0, bal, s;\n\t}"
      : "=r"(x) : "r"(d), "n"(1u << B));
  return x;
}
__device__ __forceinline__ unsigned match8(unsigned d) {
  return same_bit<0>(d) & same_bit<1>(d) & same_bit<2>(d) & same_bit<3>(d) & same_bit<4>(d) & same_bit<5>(d) & same_bit<6>(d) & same_bit<7>(d);
}

template <typename UK, int BLOCK, int ITEMS, int MATCH, bool FULL>
__device__ __forceinline__ void sweep2_tile(const UK *__restrict__ in, UK *__restrict__ out, unsigned n, int shift, long long dstart,
                                            unsigned *__restrict__ desc, int tile, UK *staged, unsigned (*cnt)[256], long long *gdelta,
                                            unsigned *wsum, unsigned (*mm)[256]) {
  constexpr int WARPS = BLOCK / 32, TILE = BLOCK * ITEMS, LB = 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const unsigned base = (unsigned)tile * TILE;
  const int valid = FULL ? TILE : (int)(n - base);
  unsigned *cw = &cnt[warp][0];

  UK key[ITEMS];
  unsigned rk[(ITEMS + 1) / 2];
  const UK *src = in + base + warp * (32 * ITEMS) + lane;
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    if (FULL) key[i] = __ldcs(src + i * 32);
    else key[i] = (warp * (32 * ITEMS) + i * 32 + lane < valid) ? src[i * 32] : (UK) ~(UK)0;
  }
  asm volatile("" ::: "memory");   // all loads of the tile are in flight before the ranking starts
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const unsigned d = (unsigned)(key[i] >> shift) & 0xff;
    const bool ok = FULL || (warp * (32 * ITEMS) + i * 32 + lane < valid);
    unsigned m;
    if (MATCH == 0) {
      m = match8(d);
      if (!FULL) m &= __ballot_sync(0xffffffffu, ok);
    } else {
      if (ok) atomicOr(&mm[warp][d], 1u << lane);
      __syncwarp();
      m = ok ? mm[warp][d] : (1u << lane);
      __syncwarp();
      if (ok && (m & lt) == 0) mm[warp][d] = 0;
    }
    unsigned before = 0;
    if (ok) before = cw[d];
    __syncwarp();
    const unsigned r = before + __popc(m & lt);
    if (i & 1) rk[i / 2] |= r << 16; else rk[i / 2] = r;
    if (ok && (m & lt) == 0) cw[d] = before + __popc(m);
    __syncwarp();
  }
  __syncthreads();
  unsigned run = 0, excl = 0, look[LB], c[WARPS];
  unsigned *mine = desc + (size_t)tile * 256 + tid;
  if (tid < 256) {
#pragma unroll
    for (int w = 0; w < WARPS; w++) {
      c[w] = cnt[w][tid];
      cnt[w][tid] = 0;   // ready for the next tile (every read of this tile's counts is below, from registers)
    }
#pragma unroll
    for (int w = 0; w < WARPS; w++) run += c[w];
    if (tile > 0) {
      st_relaxed(mine, kAgg | run);
#pragma unroll
      for (int j = 0; j < LB; j++) look[j] = (tile - 1 - j >= 0) ? ld_relaxed(mine - (size_t)(j + 1) * 256) : kPre;
    }
    unsigned incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    excl = incl - run;
  }
  __syncthreads();
  // the warps' bases in the regrouped tile go to a second array (wb = cnt's twin), so the counters above are already clean
  unsigned (*wb)[256] = cnt + WARPS;
  if (tid < 256) {
    unsigned wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; w++)
      if (w < warp) wbase += wsum[w];
    unsigned ts = wbase + excl;
    excl = ts;
#pragma unroll
    for (int w = 0; w < WARPS; w++) {
      wb[w][tid] = ts;
      ts += c[w];
    }
  }
  __syncthreads();
  const unsigned *wbw = &wb[warp][0];
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const unsigned d = (unsigned)(key[i] >> shift) & 0xff;
    const bool ok = FULL || (warp * (32 * ITEMS) + i * 32 + lane < valid);
    if (ok) staged[wbw[d] + ((i & 1) ? (rk[i / 2] >> 16) : (rk[i / 2] & 0xffffu))] = key[i];
  }
  if (tid < 256) {
    const unsigned ts = excl;
    unsigned long long before = 0;
    if (tile > 0) {
      int t = tile - 1;
      while (true) {
        bool done = false;
#pragma unroll
        for (int j = 0; j < LB; j++) {
          if (!done) {
            unsigned v = look[j];
            while ((v >> 30) == 0) v = ld_relaxed(desc + (size_t)(t - j) * 256 + tid);
            before += v & kMask;
            if ((v >> 30) == 2) done = true;
          }
        }
        if (done) break;
        t -= LB;
#pragma unroll
        for (int j = 0; j < LB; j++) look[j] = (t - j >= 0) ? ld_relaxed(desc + (size_t)(t - j) * 256 + tid) : kPre;
      }
    }
    st_relaxed(mine, kPre | (unsigned)(before + run));
    gdelta[tid] = dstart + (long long)before - (long long)ts;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const int pos = i * BLOCK + tid;
    if (FULL || pos < valid) {
      const UK k = staged[pos];
      const unsigned d = (unsigned)(k >> shift) & 0xff;
      out[gdelta[d] + (long long)pos] = k;
    }
  }
}

template <typename UK, int BLOCK, int ITEMS, int MATCH>
__device__ __noinline__ void sweep2_partial(const UK *__restrict__ in, UK *__restrict__ out, unsigned n, int shift, long long dstart,
                                            unsigned *__restrict__ desc, int tile, UK *staged, unsigned (*cnt)[256], long long *gdelta,
                                            unsigned *wsum, unsigned (*mm)[256]) {
  sweep2_tile<UK, BLOCK, ITEMS, MATCH, false>(in, out, n, shift, dstart, desc, tile, staged, cnt, gdelta, wsum, mm);
}

template <typename UK, int BLOCK, int ITEMS, int MATCH, int MINCTAS, int LBX>
__global__ void __launch_bounds__(BLOCK, MINCTAS) sweep2(const UK *__restrict__ in, UK *__restrict__ out, unsigned n, int shift,
                                                         const unsigned long long *__restrict__ bins, unsigned *__restrict__ desc,
                                                         unsigned *__restrict__ ticket, int n_tiles) {
  constexpr int WARPS = BLOCK / 32, TILE = BLOCK * ITEMS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  UK *staged = (UK *)smem_raw;
  unsigned(*cnt)[256] = (unsigned(*)[256])(staged + TILE);      // [2 * WARPS][256]: counters, then the warps' bases
  long long *gdelta = (long long *)(cnt + 2 * WARPS);
  unsigned *wsum = (unsigned *)(gdelta + 256);
  unsigned(*mm)[256] = (unsigned(*)[256])(wsum + 8);
  __shared__ int sh_tile;
  const int tid = threadIdx.x;
  long long dstart = 0;
  if (tid < 256) dstart = (long long)bins[tid];
  for (int i = tid; i < WARPS * 256; i += BLOCK) (&cnt[0][0])[i] = 0;
  if (MATCH == 2)
    for (int i = tid; i < WARPS * 256; i += BLOCK) (&mm[0][0])[i] = 0;
  while (true) {
    __syncthreads();
    if (tid == 0) sh_tile = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = sh_tile;
    if (tile >= n_tiles) break;
    if ((unsigned)(tile + 1) * TILE <= n)
      sweep2_tile<UK, BLOCK, ITEMS, MATCH, true>(in, out, n, shift, dstart, desc, tile, staged, cnt, gdelta, wsum, mm);
    else
      sweep2_partial<UK, BLOCK, ITEMS, MATCH>(in, out, n, shift, dstart, desc, tile, staged, cnt, gdelta, wsum, mm);
  }
}


// ---- third form: the tile's digit counts are known (and public) BEFORE the ranking starts -------------------------
// per-warp histogram with shared atomics right after the load -> aggregate published, look-back window requested ->
// the counters become running positions in the regrouped tile -> ranking writes every key straight to its slot.
// LBMODE: 0 real look-back with a window of W tiles per round, 1 none (timing only: wrong offsets)
template <typename UK, int BLOCK, int ITEMS, int W, int LBMODE, int MK, bool FULL>
__device__ __forceinline__ void sweep3_tile(const UK *__restrict__ in, UK *__restrict__ out, unsigned n, int shift, long long dstart,
                                            unsigned *__restrict__ desc, int tile, UK *staged, unsigned (*cnt)[256], long long *gdelta,
                                            unsigned *wsum, unsigned *ticket, int *sh_next) {
  constexpr int WARPS = BLOCK / 32, TILE = BLOCK * ITEMS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const unsigned base = (unsigned)tile * TILE;
  const int valid = FULL ? TILE : (int)(n - base);
  unsigned *cw = &cnt[warp][0];
  unsigned *mw = &cnt[WARPS + warp][0];   // MK == 1: the warp's match masks (all zero between two keys)

  UK key[ITEMS];
  const UK *src = in + base + warp * (32 * ITEMS) + lane;
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    if (FULL) key[i] = __ldcs(src + i * 32);
    else key[i] = (warp * (32 * ITEMS) + i * 32 + lane < valid) ? src[i * 32] : (UK) ~(UK)0;
  }
  asm volatile("" ::: "memory");
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const unsigned d = (unsigned)(key[i] >> shift) & 0xff;
    if (FULL || (warp * (32 * ITEMS) + i * 32 + lane < valid)) atomicAdd(&cw[d], 1u);
  }
  __syncthreads();
  unsigned run = 0, excl = 0, look[W], c[WARPS];
  unsigned *mine = desc + (size_t)tile * 256 + tid;
  if (tid < 256) {
#pragma unroll
    for (int w = 0; w < WARPS; w++) c[w] = cnt[w][tid];
#pragma unroll
    for (int w = 0; w < WARPS; w++) run += c[w];
    if (tile > 0 && LBMODE != 1) {
      st_relaxed(mine, kAgg | run);
      if (LBMODE == 0) {
#pragma unroll
        for (int j = 0; j < W; j++) look[j] = (tile - 1 - j >= 0) ? ld_relaxed(mine - (size_t)(j + 1) * 256) : kPre;
      }
    }
    unsigned incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    excl = incl - run;
  }
  __syncthreads();
  if (tid < 256) {
    unsigned wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; w++)
      if (w < warp) wbase += wsum[w];
    unsigned ts = wbase + excl;
    excl = ts;
#pragma unroll
    for (int w = 0; w < WARPS; w++) {
      cnt[w][tid] = ts;
      ts += c[w];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const unsigned d = (unsigned)(key[i] >> shift) & 0xff;
    const bool ok = FULL || (warp * (32 * ITEMS) + i * 32 + lane < valid);
    unsigned m;
    if (MK == 0) {
      m = match8(d);
      if (!FULL) m &= __ballot_sync(0xffffffffu, ok);
    } else {
      if (ok) atomicOr(&mw[d], 1u << lane);
      __syncwarp();
      m = ok ? mw[d] : (1u << lane);
    }
    unsigned before = 0;
    if (ok) before = cw[d];
    __syncwarp();
    if (ok && (m & lt) == 0) {
      cw[d] = before + __popc(m);
      if (MK == 1) mw[d] = 0;
    }
    if (ok) staged[before + __popc(m & lt)] = key[i];
    __syncwarp();
  }
  if (tid == 0) *sh_next = (int)atomicAdd(ticket, 1u);
  if (tid < 256) {
    const unsigned ts = excl;
    unsigned long long before = 0;
    if (LBMODE == 1) {
      before = (unsigned long long)tile * (TILE / 256);
    } else if (tile > 0) {
      int t = tile - 1;
      if (LBMODE == 2) {
#pragma unroll
        for (int j = 0; j < W; j++) look[j] = (t - j >= 0) ? ld_relaxed(desc + (size_t)(t - j) * 256 + tid) : kPre;
      }
      while (true) {
        bool done = false;
#pragma unroll
        for (int j = 0; j < W; j++) {
          if (!done) {
            unsigned v = look[j];
            while ((v >> 30) == 0) v = ld_relaxed(desc + (size_t)(t - j) * 256 + tid);
            before += v & kMask;
            if ((v >> 30) == 2) done = true;
          }
        }
        if (done) break;
        t -= W;
#pragma unroll
        for (int j = 0; j < W; j++) look[j] = (t - j >= 0) ? ld_relaxed(desc + (size_t)(t - j) * 256 + tid) : kPre;
      }
    }
    if (LBMODE != 1) st_relaxed(mine, kPre | (unsigned)(before + run));
    gdelta[tid] = dstart + (long long)before - (long long)ts;
  }
  __syncthreads();
  for (int i = tid; i < WARPS * 256; i += BLOCK) (&cnt[0][0])[i] = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const int pos = i * BLOCK + tid;
    if (FULL || pos < valid) {
      const UK k = staged[pos];
      const unsigned d = (unsigned)(k >> shift) & 0xff;
      out[gdelta[d] + (long long)pos] = k;
    }
  }
}

template <typename UK, int BLOCK, int ITEMS, int W, int LBMODE, int MK>
__device__ __noinline__ void sweep3_partial(const UK *__restrict__ in, UK *__restrict__ out, unsigned n, int shift, long long dstart,
                                            unsigned *__restrict__ desc, int tile, UK *staged, unsigned (*cnt)[256], long long *gdelta,
                                            unsigned *wsum, unsigned *ticket, int *sh_next) {
  sweep3_tile<UK, BLOCK, ITEMS, W, LBMODE, MK, false>(in, out, n, shift, dstart, desc, tile, staged, cnt, gdelta, wsum, ticket, sh_next);
}

// MATCH is reused as LBMODE here
template <typename UK, int BLOCK, int ITEMS, int LBMODE, int MINCTAS, int W, int MK>
__global__ void __launch_bounds__(BLOCK, MINCTAS) sweep3(const UK *__restrict__ in, UK *__restrict__ out, unsigned n, int shift,
                                                         const unsigned long long *__restrict__ bins, unsigned *__restrict__ desc,
                                                         unsigned *__restrict__ ticket, int n_tiles) {
  constexpr int WARPS = BLOCK / 32, TILE = BLOCK * ITEMS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  UK *staged = (UK *)smem_raw;
  unsigned(*cnt)[256] = (unsigned(*)[256])(staged + TILE);
  long long *gdelta = (long long *)(cnt + 2 * WARPS);
  unsigned *wsum = (unsigned *)(gdelta + 256);
  __shared__ int sh_tile;
  const int tid = threadIdx.x;
  long long dstart = 0;
  if (tid < 256) dstart = (long long)bins[tid];
  for (int i = tid; i < 2 * WARPS * 256; i += BLOCK) (&cnt[0][0])[i] = 0;
  if (tid == 0) sh_tile = (int)atomicAdd(ticket, 1u);
  while (true) {
    __syncthreads();
    const int tile = sh_tile;
    if (tile >= n_tiles) break;
    if ((unsigned)(tile + 1) * TILE <= n)
      sweep3_tile<UK, BLOCK, ITEMS, W, LBMODE, MK, true>(in, out, n, shift, dstart, desc, tile, staged, cnt, gdelta, wsum, ticket, &sh_tile);
    else
      sweep3_partial<UK, BLOCK, ITEMS, W, LBMODE, MK>(in, out, n, shift, dstart, desc, tile, staged, cnt, gdelta, wsum, ticket, &sh_tile);
  }
}


// ---- fifth form: no key stays in a register across a barrier — the tile is read twice (the second time from L2) ----
// count: load, digit, shared atomic.  rank: load again (batches of CH keys in flight), ballots, slot, store to the
// regrouped tile.  Few registers -> more resident CTAs or larger tiles (fewer, cheaper look-backs).
template <typename UK, int BLOCK, int ITEMS, int W, int CH, bool FULL>
__device__ __forceinline__ void sweep5_tile(const UK *__restrict__ in, UK *__restrict__ out, unsigned n, int shift, long long dstart,
                                            unsigned *__restrict__ desc, int tile, UK *staged, unsigned (*cnt)[256], long long *gdelta,
                                            unsigned *wsum, unsigned *ticket, int *sh_next) {
  constexpr int WARPS = BLOCK / 32, TILE = BLOCK * ITEMS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const unsigned base = (unsigned)tile * TILE;
  const int valid = FULL ? TILE : (int)(n - base);
  unsigned *cw = &cnt[warp][0];
  const UK *src = in + base + warp * (32 * ITEMS) + lane;
#pragma unroll
  for (int i0 = 0; i0 < ITEMS; i0 += CH) {
    UK key[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (FULL) key[i] = __ldcg(src + (i0 + i) * 32);
      else key[i] = (warp * (32 * ITEMS) + (i0 + i) * 32 + lane < valid) ? src[(i0 + i) * 32] : (UK) ~(UK)0;
    }
#pragma unroll
    for (int i = 0; i < CH; i++) {
      const unsigned d = (unsigned)(key[i] >> shift) & 0xff;
      if (FULL || (warp * (32 * ITEMS) + (i0 + i) * 32 + lane < valid)) atomicAdd(&cw[d], 1u);
    }
  }
  __syncthreads();
  unsigned run = 0, excl = 0, look[W], c[WARPS];
  unsigned *mine = desc + (size_t)tile * 256 + tid;
  if (tid < 256) {
#pragma unroll
    for (int w = 0; w < WARPS; w++) c[w] = cnt[w][tid];
#pragma unroll
    for (int w = 0; w < WARPS; w++) run += c[w];
    if (tile > 0) st_relaxed(mine, kAgg | run);
    unsigned incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    excl = incl - run;
  }
  __syncthreads();
  if (tid < 256) {
    unsigned wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; w++)
      if (w < warp) wbase += wsum[w];
    unsigned ts = wbase + excl;
    excl = ts;
#pragma unroll
    for (int w = 0; w < WARPS; w++) {
      cnt[w][tid] = ts;
      ts += c[w];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i0 = 0; i0 < ITEMS; i0 += CH) {
    UK key[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (FULL) key[i] = __ldcs(src + (i0 + i) * 32);
      else key[i] = (warp * (32 * ITEMS) + (i0 + i) * 32 + lane < valid) ? src[(i0 + i) * 32] : (UK) ~(UK)0;
    }
#pragma unroll
    for (int i = 0; i < CH; i++) {
      const unsigned d = (unsigned)(key[i] >> shift) & 0xff;
      const bool ok = FULL || (warp * (32 * ITEMS) + (i0 + i) * 32 + lane < valid);
      unsigned m = match8(d);
      if (!FULL) m &= __ballot_sync(0xffffffffu, ok);
      unsigned before = 0;
      if (ok) before = cw[d];
      __syncwarp();
      if (ok && (m & lt) == 0) cw[d] = before + __popc(m);
      if (ok) staged[before + __popc(m & lt)] = key[i];
      __syncwarp();
    }
  }
  if (tid == 0) *sh_next = (int)atomicAdd(ticket, 1u);
  if (tid < 256) {
    const unsigned ts = excl;
    unsigned long long before = 0;
    if (tile > 0) {
      int t = tile - 1;
      while (true) {
#pragma unroll
        for (int j = 0; j < W; j++) look[j] = (t - j >= 0) ? ld_relaxed(desc + (size_t)(t - j) * 256 + tid) : kPre;
        bool done = false;
#pragma unroll
        for (int j = 0; j < W; j++) {
          if (!done) {
            unsigned v = look[j];
            while ((v >> 30) == 0) v = ld_relaxed(desc + (size_t)(t - j) * 256 + tid);
            before += v & kMask;
            if ((v >> 30) == 2) done = true;
          }
        }
        if (done) break;
        t -= W;
      }
    }
    st_relaxed(mine, kPre | (unsigned)(before + run));
    gdelta[tid] = dstart + (long long)before - (long long)ts;
  }
  __syncthreads();
  for (int i = tid; i < WARPS * 256; i += BLOCK) (&cnt[0][0])[i] = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const int pos = i * BLOCK + tid;
    if (FULL || pos < valid) {
      const UK k = staged[pos];
      const unsigned d = (unsigned)(k >> shift) & 0xff;
      out[gdelta[d] + (long long)pos] = k;
    }
  }
}

template <typename UK, int BLOCK, int ITEMS, int W, int CH>
__device__ __noinline__ void sweep5_partial(const UK *__restrict__ in, UK *__restrict__ out, unsigned n, int shift, long long dstart,
                                            unsigned *__restrict__ desc, int tile, UK *staged, unsigned (*cnt)[256], long long *gdelta,
                                            unsigned *wsum, unsigned *ticket, int *sh_next) {
  sweep5_tile<UK, BLOCK, ITEMS, W, CH, false>(in, out, n, shift, dstart, desc, tile, staged, cnt, gdelta, wsum, ticket, sh_next);
}

// MATCH slot = CH (keys in flight per batch), LB slot = W
template <typename UK, int BLOCK, int ITEMS, int CH, int MINCTAS, int W>
__global__ void __launch_bounds__(BLOCK, MINCTAS) sweep5(const UK *__restrict__ in, UK *__restrict__ out, unsigned n, int shift,
                                                         const unsigned long long *__restrict__ bins, unsigned *__restrict__ desc,
                                                         unsigned *__restrict__ ticket, int n_tiles) {
  constexpr int WARPS = BLOCK / 32, TILE = BLOCK * ITEMS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  UK *staged = (UK *)smem_raw;
  unsigned(*cnt)[256] = (unsigned(*)[256])(staged + TILE);
  long long *gdelta = (long long *)(cnt + WARPS);
  unsigned *wsum = (unsigned *)(gdelta + 256);
  __shared__ int sh_tile;
  const int tid = threadIdx.x;
  long long dstart = 0;
  if (tid < 256) dstart = (long long)bins[tid];
  for (int i = tid; i < WARPS * 256; i += BLOCK) (&cnt[0][0])[i] = 0;
  if (tid == 0) sh_tile = (int)atomicAdd(ticket, 1u);
  while (true) {
    __syncthreads();
    const int tile = sh_tile;
    if (tile >= n_tiles) break;
    if ((unsigned)(tile + 1) * TILE <= n)
      sweep5_tile<UK, BLOCK, ITEMS, W, CH, true>(in, out, n, shift, dstart, desc, tile, staged, cnt, gdelta, wsum, ticket, &sh_tile);
    else
      sweep5_partial<UK, BLOCK, ITEMS, W, CH>(in, out, n, shift, dstart, desc, tile, staged, cnt, gdelta, wsum, ticket, &sh_tile);
  }
}

template <typename UK>
struct Bench {
  long long n;
  UK *src, *a, *b;
  unsigned long long *gh, *bins, *res;
  unsigned *desc;
  size_t desc_bytes;
  unsigned long long want_sum, want_xor;
  cudaEvent_t e0, e1;
  int sms;

  void init(long long n_) {
    n = n_;
    CK(cudaMalloc(&src, sizeof(UK) * n));
    CK(cudaMalloc(&a, sizeof(UK) * n));
    CK(cudaMalloc(&b, sizeof(UK) * n));
    CK(cudaMalloc(&gh, 8 * 256 * 8));
    CK(cudaMalloc(&bins, 8 * 256 * 8));
    CK(cudaMalloc(&res, 64));
    desc_bytes = 256 + (size_t)((n + 1023) / 1024) * 256 * 4;
    CK(cudaMalloc(&desc, desc_bytes));
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    sms = pr.multiProcessorCount;
    fill<UK><<<sms * 8, 256>>>(src, n, 0);
    unsigned long long h[3];
    verify(src, h);
    want_sum = h[1];
    want_xor = h[2];
  }
  void verify(const UK *k, unsigned long long *h) {
    CK(cudaMemset(res, 0, 64));
    check<UK><<<sms * 8, 256>>>(k, n, res);
    CK(cudaMemcpy(h, res, 24, cudaMemcpyDeviceToHost));
  }

  template <int VER, int BLOCK, int ITEMS, int MATCH, int MINCTAS, int LB>
  void run(const char *label) {
    constexpr int P = sizeof(UK), TILE = BLOCK * ITEMS, WARPS = BLOCK / 32;
    auto kern = VER == 5 ? sweep5<UK, BLOCK, ITEMS, MATCH, MINCTAS, LB> : VER == 4 ? sweep3<UK, BLOCK, ITEMS, MATCH, MINCTAS, LB, 1> : VER == 3 ? sweep3<UK, BLOCK, ITEMS, MATCH, MINCTAS, LB, 0> : VER == 2 ? sweep2<UK, BLOCK, ITEMS, MATCH, MINCTAS, LB> : sweep<UK, BLOCK, ITEMS, MATCH, MINCTAS, LB>;
    const size_t smem = sizeof(UK) * TILE + 4 * 256 * WARPS * ((VER >= 2 && VER <= 4) ? 2 : 1) + 8 * 256 + 32 + (MATCH == 2 ? 4 * 256 * WARPS : 0);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    const int n_tiles = (int)((n + TILE - 1) / TILE);
    const int grid = (sms * occ < n_tiles) ? sms * occ : n_tiles;
    float best = 1e30f, best_hist = 0;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaMemcpy(a, src, sizeof(UK) * n, cudaMemcpyDeviceToDevice));
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      CK(cudaMemsetAsync(gh, 0, 8 * 256 * 8));
      hist_all<UK><<<sms * 8, 256>>>(a, n, gh);
      scan_bins<<<P, 256>>>(gh, bins);
      CK(cudaEventRecord(e1));
      UK *s = a, *d = b;
      for (int p = 0; p < P; p++) {
        CK(cudaMemsetAsync(desc, 0, 256 + (size_t)n_tiles * 256 * 4));
        kern<<<grid, BLOCK, smem>>>(s, d, (unsigned)n, 8 * p, bins + p * 256, desc + 64, desc, n_tiles);
        UK *t = s; s = d; d = t;
      }
      cudaEvent_t e2;
      CK(cudaEventCreate(&e2));
      CK(cudaEventRecord(e2));
      CK(cudaDeviceSynchronize());
      CK(cudaGetLastError());
      float th, tt;
      CK(cudaEventElapsedTime(&th, e0, e1));
      CK(cudaEventElapsedTime(&tt, e0, e2));
      CK(cudaEventDestroy(e2));
      if (tt < best) { best = tt; best_hist = th; }
    }
    unsigned long long h[3];
    verify((P % 2) ? b : a, h);
    const bool okres = h[0] == 0 && h[1] == want_sum && h[2] == want_xor;
    printf("%-34s regs %3d occ %d smem %6zu  total %8.3f ms (hist %.3f, pass %.3f)  %6.2f Gkeys/s  %s\n", label, fa.numRegs, occ, smem, best,
           best_hist, (best - best_hist) / P, n / best / 1e6, (MATCH == 3 || ((VER == 3 || VER == 4) && MATCH == 1)) ? "(timing only)" : (okres ? "ok" : "WRONG"));
    fflush(stdout);
  }

  void run_cub() {
    void *tmp = nullptr;
    size_t tb = 0;
    cub::DoubleBuffer<UK> db(a, b);
    CK(cub::DeviceRadixSort::SortKeys(tmp, tb, db, (int)n));
    CK(cudaMalloc(&tmp, tb));
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaMemcpy(a, src, sizeof(UK) * n, cudaMemcpyDeviceToDevice));
      cub::DoubleBuffer<UK> d2(a, b);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      CK(cub::DeviceRadixSort::SortKeys(tmp, tb, d2, (int)n));
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float t;
      CK(cudaEventElapsedTime(&t, e0, e1));
      if (t < best) best = t;
      if (rep == 2) {
        unsigned long long h[3];
        verify(d2.Current(), h);
        printf("%-34s total %8.3f ms  %6.2f Gkeys/s  %s\n", "CUB DeviceRadixSort::SortKeys", best, n / best / 1e6,
               (h[0] == 0 && h[1] == want_sum && h[2] == want_xor) ? "ok" : "WRONG");
      }
    }
    CK(cudaFree(tmp));
  }
};

#define RUNV(VER, B, BLOCK, ITEMS, MATCH, MINCTAS, LB)                                                              \
  do {                                                                                                         \
    const char *l_ = "v" #VER " " #BLOCK "x" #ITEMS " match" #MATCH " ctas" #MINCTAS " lb" #LB;                          \
    if (!filter || strstr(l_, filter)) B.template run<VER, BLOCK, ITEMS, MATCH, MINCTAS, LB>(l_);                   \
  } while (0)

#define RUN(B, BLOCK, ITEMS, MATCH, MINCTAS, LB) RUNV(1, B, BLOCK, ITEMS, MATCH, MINCTAS, LB)

int main(int argc, char **argv) {
  const int log2n = argc > 1 ? atoi(argv[1]) : 28;
  const char *filter = argc > 2 ? argv[2] : nullptr;
  const long long n = 1ll << log2n;
  {
    printf("== 64-bit keys, n = 2^%d\n", log2n);
    Bench<unsigned long long> B;
    B.init(n);
    B.run_cub();
    RUN(B, 256, 16, 2, 3, 4);
    RUNV(5, B, 256, 24, 8, 3, 4);
    RUNV(5, B, 256, 24, 4, 3, 4);
    RUNV(5, B, 256, 24, 12, 3, 4);
    RUNV(5, B, 256, 24, 8, 3, 2);
    RUNV(5, B, 256, 24, 8, 3, 3);
    RUNV(5, B, 256, 20, 4, 3, 4);
    RUNV(5, B, 256, 20, 4, 4, 4);
    RUNV(5, B, 256, 28, 4, 3, 4);
    RUNV(5, B, 256, 32, 8, 3, 4);
    RUNV(5, B, 384, 16, 8, 2, 4);
    RUNV(5, B, 384, 16, 8, 3, 4);
    RUNV(3, B, 256, 16, 1, 3, 8);
    CK(cudaFree(B.src)); CK(cudaFree(B.a)); CK(cudaFree(B.b));
  }
  {
    printf("== 32-bit keys, n = 2^%d\n", log2n);
    Bench<unsigned int> B;
    B.init(n);
    B.run_cub();
    RUNV(5, B, 256, 32, 16, 4, 4);
    RUNV(5, B, 256, 32, 8, 4, 4);
    RUNV(5, B, 256, 32, 16, 4, 2);
    RUNV(5, B, 256, 40, 8, 3, 4);
    RUNV(5, B, 256, 48, 16, 3, 4);
    RUNV(5, B, 384, 32, 16, 2, 4);
    RUNV(5, B, 512, 16, 16, 3, 4);
  }
  return 0;
}
