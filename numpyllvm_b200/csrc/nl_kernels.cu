// nl_kernels.cu — the sm_100a kernel library behind nl_launch_pipeline().
//
// ONE kernel template covers the four paths of the engine, because in the
// reference they are one thing too: a single fused loop per pipeline
// (compiler.cpp:362-509).  pipeline_kernel<T, V, kCompact, kReduce> interprets
// the plan's accumulator-machine program (nl_step.h) over tiles of
// kBlock * kRows * V elements:
//
//   (1) elementwise chains   128-bit coalesced streaming loads (V = 16/sizeof(T)
//                            elements per access, kRows independent accesses in
//                            flight per thread per operand), operands and all
//                            intermediates stay in registers, grid-stride tiles
//   (2) filter compaction    per-row ballot/popc counts, packed warp scan,
//                            single-pass decoupled look-back across tiles
//                            (status + value packed in one 64-bit word),
//                            stable order, 64-bit offsets
//   (3) max / min / sum      per-thread accumulators -> warp shuffle tree ->
//                            block tree -> per-CTA partial -> last CTA combines
//                            all partials in a fixed order (deterministic)
//   (4) in-place assignment  masked store at the loop index (S_WHERE) and the
//                            fill / copy kernels below
//
// All are HBM-bound streaming kernels: no shared-memory staging of operands
// (no reuse), no tensor cores (nothing is a contraction).  Integer arithmetic
// is done in unsigned to get the reference's two's-complement wrap; the
// translation unit is compiled with -fmad=false so float chains round exactly
// like the reference's fmul/fadd sequence.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <limits.h>
#include <stdio.h>
#include <string.h>

#include "nl_internal.h"
#include "nl_step.h"

namespace nl {

// ------------------------------------------------------------------ type traits
template <typename T> struct TT;
template <> struct TT<int32_t> {
  using U = uint32_t; using Acc = long long; static constexpr bool is_float = false;
  __device__ static int32_t lowest() { return INT32_MIN; }
  __device__ static int32_t highest() { return INT32_MAX; }
};
template <> struct TT<long long> {
  using U = unsigned long long; using Acc = long long; static constexpr bool is_float = false;
  __device__ static long long lowest() { return LLONG_MIN; }
  __device__ static long long highest() { return LLONG_MAX; }
};
template <> struct TT<float> {
  using U = uint32_t; using Acc = double; static constexpr bool is_float = true;
  __device__ static float lowest() { return -INFINITY; }
  __device__ static float highest() { return INFINITY; }
};
template <> struct TT<double> {
  using U = unsigned long long; using Acc = double; static constexpr bool is_float = true;
  __device__ static double lowest() { return -INFINITY; }
  __device__ static double highest() { return INFINITY; }
};

template <typename T> __device__ __forceinline__ T op_mul(T a, T b) {
  if constexpr (TT<T>::is_float) return a * b; else return (T)((typename TT<T>::U)a * (typename TT<T>::U)b);
}
template <typename T> __device__ __forceinline__ T op_add(T a, T b) {
  if constexpr (TT<T>::is_float) return a + b; else return (T)((typename TT<T>::U)a + (typename TT<T>::U)b);
}
template <typename T> __device__ __forceinline__ T op_sub(T a, T b) {
  if constexpr (TT<T>::is_float) return a - b; else return (T)((typename TT<T>::U)a - (typename TT<T>::U)b);
}
// max/min with NumPy semantics: NaN is sticky
template <typename T> __device__ __forceinline__ T op_max(T cur, T l) {
  if constexpr (TT<T>::is_float) return (l > cur || l != l) ? l : cur; else return l > cur ? l : cur;
}
template <typename T> __device__ __forceinline__ T op_min(T cur, T l) {
  if constexpr (TT<T>::is_float) return (l < cur || l != l) ? l : cur; else return l < cur ? l : cur;
}
__device__ __forceinline__ long long acc_add(long long a, long long b) { return (long long)((unsigned long long)a + (unsigned long long)b); }
__device__ __forceinline__ double acc_add(double a, double b) { return a + b; }

// ------------------------------------------------------------------ vector access
template <int BYTES> struct VecOf;
template <> struct VecOf<16> { using type = uint4; };
template <> struct VecOf<8> { using type = uint2; };
template <> struct VecOf<4> { using type = unsigned int; };
template <> struct VecOf<2> { using type = unsigned short; };
template <> struct VecOf<1> { using type = unsigned char; };

// one row = V consecutive elements of one thread; streaming (evict-first) access
template <typename T, int V>
__device__ __forceinline__ void load_row(const T *__restrict__ base, int64_t idx, int nvalid, T *dst) {
  using Vec = typename VecOf<sizeof(T) * V>::type;
  if (nvalid == V) {
    union { Vec q; T t[V]; } u;
    u.q = __ldcs(reinterpret_cast<const Vec *>(base + idx));
#pragma unroll
    for (int v = 0; v < V; v++) dst[v] = u.t[v];
  } else {
#pragma unroll
    for (int v = 0; v < V; v++) dst[v] = (v < nvalid) ? base[idx + v] : T(0);
  }
}
template <typename T, int V>
__device__ __forceinline__ void store_row(T *__restrict__ base, int64_t idx, int nvalid, const T *src) {
  using Vec = typename VecOf<sizeof(T) * V>::type;
  if (nvalid == V) {
    union { Vec q; T t[V]; } u;
#pragma unroll
    for (int v = 0; v < V; v++) u.t[v] = src[v];
    __stcs(reinterpret_cast<Vec *>(base + idx), u.q);
  } else {
#pragma unroll
    for (int v = 0; v < V; v++) if (v < nvalid) base[idx + v] = src[v];
  }
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <typename T> __device__ __forceinline__ T from_bits(uint64_t bits) {
  if constexpr (sizeof(T) == 8) { union { uint64_t b; T t; } u; u.b = bits; return u.t; }
  else { union { uint32_t b; T t; } u; u.b = (uint32_t)bits; return u.t; }
}
template <typename T> __device__ __forceinline__ unsigned long long to_bits(T v) {
  if constexpr (sizeof(T) == 8) { union { unsigned long long b; T t; } u; u.t = v; return u.b; }
  else { union { uint32_t b; T t; } u; u.t = v; return u.b; }
}

constexpr unsigned long long kDescValueMask = (1ull << 62) - 1;
constexpr unsigned long long kDescAggregate = 1ull << 62;
constexpr unsigned long long kDescPrefix = 2ull << 62;

// ------------------------------------------------------------------ the pipeline kernel
template <typename T, int V, bool kCompact, bool kReduce>
__global__ void __launch_bounds__(kBlock, kMinBlocks) pipeline_kernel(const __grid_constant__ KParams p) {
  constexpr int U = kRows;
  constexpr int E = U * V;
  constexpr int NW = kBlock / 32;
  constexpr int64_t TILE = (int64_t)kBlock * E;
  constexpr uint32_t ROWMASK = (1u << V) - 1u;
  static_assert(E <= 32, "one predicate bit per element");
  static_assert(U * NW == 32, "the tile scan assumes one warp covers all (row, warp) totals");
  static_assert(U <= 4 && V * 32 < 256, "packed 8-bit row counters");
  using Tr = TT<T>;
  using Acc = typename Tr::Acc;

  __shared__ uint32_t s_wtot[NW];        // per-warp packed row totals
  __shared__ uint32_t s_off[U * NW];     // exclusive offset of (row, warp) inside the tile
  __shared__ long long s_prefix;         // exclusive prefix of this tile (elements kept before it)
  __shared__ long long s_tile;
  __shared__ unsigned long long s_red[NW];
  __shared__ int s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // output cardinalities that do not depend on the data (compiler.cpp:491-508)
  if (blockIdx.x == 0 && tid == 0) {
    for (int o = 0; o < p.n_outputs; o++) {
      if (p.out_space[o] == NL_IDX_LOOP || p.out_space[o] == NL_IDX_WHERE) p.out_sizes[o] = p.end;
      else if (p.out_space[o] == NL_IDX_SHARD) p.out_sizes[o] = 1;
    }
  }

  Acc rsum = 0;
  T rext = (p.reduce_rop == R_MIN) ? Tr::highest() : Tr::lowest();

  int64_t tile = blockIdx.x;
  while (true) {
    if constexpr (kCompact) {
      // tiles are handed out in order so every predecessor of a tile is already owned by a
      // running CTA: the look-back below can never wait on work that has not started
      if (tid == 0) s_tile = (long long)atomicAdd(p.tile_ticket, 1u);
      __syncthreads();
      tile = s_tile;
    }
    if (tile >= p.n_tiles) break;

    const int64_t tile_base = p.start + tile * TILE;
    int64_t row_idx[U];
    int nvalid[U];
    uint32_t keep = 0;
#pragma unroll
    for (int u = 0; u < U; u++) {
      row_idx[u] = tile_base + (int64_t)(u * kBlock + tid) * V;
      const int64_t rem = p.end - row_idx[u];
      nvalid[u] = rem >= V ? V : (rem > 0 ? (int)rem : 0);
      keep |= ((1u << nvalid[u]) - 1u) << (u * V);
    }

    T acc[E], t0[E], t1[E], t2[E], x[E];
#pragma unroll
    for (int e = 0; e < E; e++) { acc[e] = T(0); t0[e] = T(0); t1[e] = T(0); t2[e] = T(0); }
    uint32_t pacc = 0, p0 = 0, p1 = 0, p2 = 0;
    int64_t cidx[U];
#pragma unroll
    for (int u = 0; u < U; u++) cidx[u] = 0;

    auto fetch = [&](int k, T(&dst)[E]) {
      const int kind = p.in_kind[k];
      if (kind == NL_IN_ARRAY) {
        const T *base = reinterpret_cast<const T *>(p.in[k]);
#pragma unroll
        for (int u = 0; u < U; u++) load_row<T, V>(base, row_idx[u], nvalid[u], &dst[u * V]);
      } else {
        const T s = (kind == NL_IN_SCALAR_IMM) ? from_bits<T>(p.imm[k]) : *reinterpret_cast<const T *>(p.in[k]);
#pragma unroll
        for (int e = 0; e < E; e++) dst[e] = s;
      }
    };
    auto operand = [&](uint32_t b) {
      if (b & SRC_TEMP) {
        switch (b & 0x7f) {
          case 0:
#pragma unroll
            for (int e = 0; e < E; e++) x[e] = t0[e];
            break;
          case 1:
#pragma unroll
            for (int e = 0; e < E; e++) x[e] = t1[e];
            break;
          default:
#pragma unroll
            for (int e = 0; e < E; e++) x[e] = t2[e];
            break;
        }
      } else {
        fetch((int)b, x);
      }
    };

#pragma unroll 1
    for (int pc = 0; pc < p.n_steps; pc++) {
      const uint32_t s = p.step[pc];
      const uint32_t a = step_a(s), b = step_b(s);
      switch (step_op(s)) {
        case S_LOAD:
          fetch((int)a, acc);
          break;
        case S_LOADT:
          switch (a) {
            case 0:
#pragma unroll
              for (int e = 0; e < E; e++) acc[e] = t0[e];
              break;
            case 1:
#pragma unroll
              for (int e = 0; e < E; e++) acc[e] = t1[e];
              break;
            default:
#pragma unroll
              for (int e = 0; e < E; e++) acc[e] = t2[e];
              break;
          }
          break;
        case S_SAVE:
          switch (a) {
            case 0:
#pragma unroll
              for (int e = 0; e < E; e++) t0[e] = acc[e];
              break;
            case 1:
#pragma unroll
              for (int e = 0; e < E; e++) t1[e] = acc[e];
              break;
            default:
#pragma unroll
              for (int e = 0; e < E; e++) t2[e] = acc[e];
              break;
          }
          break;
        case S_BIN:
          operand(b);
          switch (a) {
            case V_MUL:
#pragma unroll
              for (int e = 0; e < E; e++) acc[e] = op_mul<T>(acc[e], x[e]);
              break;
            case V_ADD:
#pragma unroll
              for (int e = 0; e < E; e++) acc[e] = op_add<T>(acc[e], x[e]);
              break;
            case V_SUB:
#pragma unroll
              for (int e = 0; e < E; e++) acc[e] = op_sub<T>(acc[e], x[e]);
              break;
            default:
#pragma unroll
              for (int e = 0; e < E; e++) acc[e] = op_sub<T>(x[e], acc[e]);
              break;
          }
          break;
        case S_CMP: {
          operand(b);
          uint32_t m = 0;
          switch (a) {
            case C_GE:
#pragma unroll
              for (int e = 0; e < E; e++) m |= (uint32_t)(acc[e] >= x[e]) << e;
              break;
            case C_GT:
#pragma unroll
              for (int e = 0; e < E; e++) m |= (uint32_t)(acc[e] > x[e]) << e;
              break;
            case C_LE:
#pragma unroll
              for (int e = 0; e < E; e++) m |= (uint32_t)(acc[e] <= x[e]) << e;
              break;
            case C_LT:
#pragma unroll
              for (int e = 0; e < E; e++) m |= (uint32_t)(acc[e] < x[e]) << e;
              break;
            case C_EQ:
#pragma unroll
              for (int e = 0; e < E; e++) m |= (uint32_t)(acc[e] == x[e]) << e;
              break;
            default:
#pragma unroll
              for (int e = 0; e < E; e++) m |= (uint32_t)(acc[e] != x[e]) << e;
              break;
          }
          pacc = m;
          break;
        }
        case S_PLOAD: {
          const int kind = p.in_kind[a];
          uint32_t m = 0;
          if (kind == NL_IN_ARRAY) {
            const unsigned char *base = reinterpret_cast<const unsigned char *>(p.in[a]);
#pragma unroll
            for (int u = 0; u < U; u++) {
              unsigned char bytes[V];
              load_row<unsigned char, V>(base, row_idx[u], nvalid[u], bytes);
#pragma unroll
              for (int v = 0; v < V; v++) m |= (uint32_t)(bytes[v] != 0) << (u * V + v);
            }
          } else {
            const bool on = (kind == NL_IN_SCALAR_IMM) ? ((p.imm[a] & 0xff) != 0)
                                                       : (*reinterpret_cast<const unsigned char *>(p.in[a]) != 0);
            m = on ? 0xffffffffu : 0u;
          }
          pacc = m;
          break;
        }
        case S_PLOADT:
          pacc = (a == 0) ? p0 : (a == 1) ? p1 : p2;
          break;
        case S_PSAVE:
          if (a == 0) p0 = pacc; else if (a == 1) p1 = pacc; else p2 = pacc;
          break;
        case S_PBIN: {
          const uint32_t q = (b == 0) ? p0 : (b == 1) ? p1 : p2;
          pacc = (a == P_AND) ? (pacc & q) : (a == P_OR) ? (pacc | q) : (pacc ^ q);
          break;
        }
        case S_PNOT:
          pacc = ~pacc;
          break;
        case S_FILTER:
          // `if (!mask) goto for.inc` (thunk_as_mapping.cpp:11): the element leaves the pipeline
          keep &= pacc;
          if constexpr (kCompact) {
            // rank of every kept element in (row, thread, v) order == its input order
            uint32_t packed = 0;
#pragma unroll
            for (int u = 0; u < U; u++) packed |= (uint32_t)__popc((keep >> (u * V)) & ROWMASK) << (8 * u);
            uint32_t incl = packed;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
              if (lane >= d) incl += n;
            }
            const uint32_t lane_excl = incl - packed;
            if (lane == 31) s_wtot[warp] = incl;
            __syncthreads();
            if (warp == 0) {
              // lane l <-> (row u = l / NW, warp w = l % NW): row-major == input order
              const int u = lane / NW, w = lane % NW;
              const uint32_t val = (s_wtot[w] >> (8 * u)) & 0xffu;
              uint32_t sc = val;
#pragma unroll
              for (int d = 1; d < 32; d <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, sc, d);
                if (lane >= d) sc += n;
              }
              s_off[lane] = sc - val;
              const unsigned long long agg = __shfl_sync(0xffffffffu, sc, 31);
              // decoupled look-back over the predecessors' descriptors
              unsigned long long excl = 0;
              if (tile == 0) {
                if (lane == 0) st_relaxed_u64(&p.tile_desc[0], kDescPrefix | agg);
              } else {
                if (lane == 0) st_relaxed_u64(&p.tile_desc[tile], kDescAggregate | agg);
                int64_t j = tile - 1;
                while (true) {
                  const int64_t q = j - lane;
                  unsigned long long d = kDescPrefix;          // before tile 0: an empty prefix
                  if (q >= 0) {
                    d = ld_relaxed_u64(&p.tile_desc[q]);
                    while ((d >> 62) == 0) d = ld_relaxed_u64(&p.tile_desc[q]);
                  }
                  const uint32_t has_prefix = __ballot_sync(0xffffffffu, (d >> 62) == 2);
                  const int first = has_prefix ? (__ffs(has_prefix) - 1) : 31;
                  unsigned long long v = (lane <= first) ? (d & kDescValueMask) : 0ull;
#pragma unroll
                  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                  excl += v;
                  if (has_prefix) break;
                  j -= 32;
                }
                if (lane == 0) st_relaxed_u64(&p.tile_desc[tile], kDescPrefix | ((excl + agg) & kDescValueMask));
              }
              if (lane == 0) {
                s_prefix = (long long)excl;
                if (tile == p.n_tiles - 1) {
                  for (int o = 0; o < p.n_outputs; o++)
                    if (p.out_space[o] == NL_IDX_COMPACT) p.out_sizes[o] = p.start + (long long)(excl + agg);
                }
              }
            }
            __syncthreads();
            const int64_t tile_excl = s_prefix;
#pragma unroll
            for (int u = 0; u < U; u++)
              cidx[u] = p.start + tile_excl + (int64_t)s_off[u * NW + warp] + (int64_t)((lane_excl >> (8 * u)) & 0xffu);
          }
          break;
        case S_STORE: {
          T *base = reinterpret_cast<T *>(p.out[a]);
          if (b == NL_IDX_LOOP) {
#pragma unroll
            for (int u = 0; u < U; u++) store_row<T, V>(base, row_idx[u], nvalid[u], &acc[u * V]);
          } else {
#pragma unroll
            for (int u = 0; u < U; u++) {
              int64_t ci = cidx[u];
#pragma unroll
              for (int v = 0; v < V; v++)
                if ((keep >> (u * V + v)) & 1u) base[ci++] = acc[u * V + v];
            }
          }
          break;
        }
        case S_PSTORE: {
          unsigned char *base = reinterpret_cast<unsigned char *>(p.out[a]);
          if (b == NL_IDX_LOOP) {
#pragma unroll
            for (int u = 0; u < U; u++) {
              unsigned char bytes[V];
#pragma unroll
              for (int v = 0; v < V; v++) bytes[v] = (unsigned char)((pacc >> (u * V + v)) & 1u);   // canonical 0x01 (Q4)
              store_row<unsigned char, V>(base, row_idx[u], nvalid[u], bytes);
            }
          } else {
#pragma unroll
            for (int u = 0; u < U; u++) {
              int64_t ci = cidx[u];
#pragma unroll
              for (int v = 0; v < V; v++)
                if ((keep >> (u * V + v)) & 1u) base[ci++] = (unsigned char)((pacc >> (u * V + v)) & 1u);
            }
          }
          break;
        }
        case S_WHERE: {
          T *base = reinterpret_cast<T *>(p.out[a]);
          const uint32_t m = pacc & keep;
#pragma unroll
          for (int u = 0; u < U; u++) {
            const uint32_t rowbits = (m >> (u * V)) & ROWMASK;
            if (rowbits == ROWMASK) {
              store_row<T, V>(base, row_idx[u], V, &acc[u * V]);
            } else {
#pragma unroll
              for (int v = 0; v < V; v++)
                if ((rowbits >> v) & 1u) base[row_idx[u] + v] = acc[u * V + v];
            }
          }
          break;
        }
        case S_REDUCE:
          if constexpr (kReduce) {
            switch (a) {
              case R_SUM:
#pragma unroll
                for (int e = 0; e < E; e++) if ((keep >> e) & 1u) rsum = acc_add(rsum, (Acc)acc[e]);
                break;
              case R_MAX:
#pragma unroll
                for (int e = 0; e < E; e++) if ((keep >> e) & 1u) rext = op_max<T>(rext, acc[e]);
                break;
              default:
#pragma unroll
                for (int e = 0; e < E; e++) if ((keep >> e) & 1u) rext = op_min<T>(rext, acc[e]);
                break;
            }
          }
          break;
        default:
          break;
      }
    }
    if constexpr (!kCompact) tile += gridDim.x;
  }

  if constexpr (kReduce) {
    // block tree in a fixed order, then the last CTA to finish folds every CTA's partial
    const int rop = p.reduce_rop;
    auto block_reduce = [&](Acc sv, T ev) -> unsigned long long {
      if (rop == R_SUM) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sv = acc_add(sv, __shfl_xor_sync(0xffffffffu, sv, o));
      } else if (rop == R_MAX) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ev = op_max<T>(ev, __shfl_xor_sync(0xffffffffu, ev, o));
      } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ev = op_min<T>(ev, __shfl_xor_sync(0xffffffffu, ev, o));
      }
      __syncthreads();
      if (lane == 0) s_red[warp] = (rop == R_SUM) ? to_bits<Acc>(sv) : to_bits<T>(ev);
      __syncthreads();
      unsigned long long out = 0;
      if (warp == 0) {
        if (rop == R_SUM) {
          Acc v = (lane < NW) ? from_bits<Acc>(s_red[lane]) : Acc(0);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v = acc_add(v, __shfl_xor_sync(0xffffffffu, v, o));
          out = to_bits<Acc>(v);
        } else {
          T v = (lane < NW) ? from_bits<T>(s_red[lane]) : (rop == R_MAX ? Tr::lowest() : Tr::highest());
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
            v = (rop == R_MAX) ? op_max<T>(v, __shfl_xor_sync(0xffffffffu, v, o)) : op_min<T>(v, __shfl_xor_sync(0xffffffffu, v, o));
          out = to_bits<T>(v);
        }
      }
      return out;   // valid in warp 0
    };

    const unsigned long long mine = block_reduce(rsum, rext);
    if (tid == 0) {
      p.red_partials[blockIdx.x] = mine;
      __threadfence();
      const unsigned int t = atomicAdd(p.red_ticket, 1u);
      s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      Acc sv = 0;
      T ev = (rop == R_MIN) ? Tr::highest() : Tr::lowest();
      for (unsigned int i = tid; i < gridDim.x; i += kBlock) {
        const unsigned long long bits = ld_relaxed_u64(&p.red_partials[i]);
        if (rop == R_SUM) sv = acc_add(sv, from_bits<Acc>(bits));
        else if (rop == R_MAX) ev = op_max<T>(ev, from_bits<T>(bits));
        else ev = op_min<T>(ev, from_bits<T>(bits));
      }
      const unsigned long long total = block_reduce(sv, ev);
      if (tid == 0) {
        T *out = reinterpret_cast<T *>(p.out[p.reduce_out]);
        // the store casts the accumulator to the column type (compiler.cpp:256-260)
        out[p.thread_nr] = (rop == R_SUM) ? (T)from_bits<Acc>(total) : from_bits<T>(total);
        *p.red_ticket = 0u;
      }
    }
  }
}

// A launch over an empty range still has to define its outputs: cardinalities and the
// reduction identity (llvm_initialize_max / _sum with zero iterations).
template <typename T>
__global__ void empty_range_kernel(const __grid_constant__ KParams p) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int o = 0; o < p.n_outputs; o++) {
    if (p.out_space[o] == NL_IDX_SHARD) {
      p.out_sizes[o] = 1;
      T *out = reinterpret_cast<T *>(p.out[o]);
      out[p.thread_nr] = (p.reduce_rop == R_SUM) ? T(0) : (p.reduce_rop == R_MAX ? TT<T>::lowest() : TT<T>::highest());
    } else if (p.out_space[o] == NL_IDX_COMPACT) {
      p.out_sizes[o] = p.start;
    } else {
      p.out_sizes[o] = p.end;
    }
  }
}

// ------------------------------------------------------------------ small kernels
template <typename T>
__global__ void finalize_partials_kernel(int rop, const T *__restrict__ partials, int n, T *__restrict__ out) {
  // llvm_finalize_sum_data (thunk_methods.cpp:133-151): fold the partials in order j = 0..n-1
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  T d = (rop == R_SUM) ? T(0) : (rop == R_MAX ? TT<T>::lowest() : TT<T>::highest());
  for (int j = 0; j < n; j++) {
    if (rop == R_SUM) d = op_add<T>(d, partials[j]);
    else if (rop == R_MAX) d = op_max<T>(d, partials[j]);
    else d = op_min<T>(d, partials[j]);
  }
  out[0] = d;
}

template <typename T>
__global__ void __launch_bounds__(256) assign_fill_kernel(T *__restrict__ dst, int64_t lo, int64_t step, int64_t count, T value) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) dst[lo + k * step] = value;
}
// contiguous fill: 128-bit stores on the aligned body
template <typename T>
__global__ void __launch_bounds__(256) assign_fill_vec_kernel(T *__restrict__ dst, int64_t count, T value) {
  constexpr int V = 16 / sizeof(T);
  const uintptr_t addr = reinterpret_cast<uintptr_t>(dst);
  int64_t head = (int64_t)(((16 - (addr & 15)) & 15) / sizeof(T));
  if (head > count) head = count;
  const int64_t nvec = (count - head) / V;
  const int64_t tail0 = head + nvec * V;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  union { uint4 q; T t[V]; } u;
#pragma unroll
  for (int v = 0; v < V; v++) u.t[v] = value;
  uint4 *body = reinterpret_cast<uint4 *>(dst + head);
  for (int64_t k = gtid; k < nvec; k += stride) __stcs(body + k, u.q);
  if (gtid < head) dst[gtid] = value;
  if (gtid < count - tail0) dst[tail0 + gtid] = value;
}
template <typename T>
__global__ void __launch_bounds__(256) assign_copy_kernel(T *__restrict__ dst, int64_t lo, int64_t step, int64_t count, const T *__restrict__ src) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) dst[lo + k * step] = __ldcs(src + k);
}

__global__ void scan_offsets_kernel(const int64_t *__restrict__ counts, int n, int64_t *__restrict__ offsets) {
  // the memmove gather of compiler.cpp:36-48, reduced to where each shard's run starts
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int64_t run = 0;
  for (int j = 0; j < n; j++) { offsets[j] = run; run += counts[j]; }
  offsets[n] = run;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
// synthetic inputs, bit-identical to the oracle's nlo_fill
template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(T *__restrict__ dst, int64_t first_index, int64_t count, uint64_t seed, int kind) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint64_t z = splitmix64(seed ^ (uint64_t)(first_index + k));
    T v;
    if constexpr (sizeof(T) == 1) {
      v = (T)(z & 1);
    } else if constexpr (!TT<T>::is_float) {
      v = (kind == 0) ? (T)(1 + (((z >> 32) * 99ULL) >> 32)) : (T)(long long)z;
    } else if constexpr (sizeof(T) == 4) {
      const float f = __uint_as_float(0x3f800000u | (uint32_t)(z >> 41));
      v = (kind == 3) ? f - 1.0f : f;
    } else {
      const double d = __longlong_as_double((long long)(0x3ff0000000000000ULL | (z >> 12)));
      v = (kind == 3) ? d - 1.0 : d;
    }
    dst[k] = v;
  }
}
template <> struct TT<unsigned char> { static constexpr bool is_float = false; };

// ------------------------------------------------------------------ host-side dispatch
typedef void (*pipeline_fn)(const KParams);

template <typename T>
static pipeline_fn pick_T(bool vec16, bool compact, bool reduce) {
  constexpr int V16 = 16 / sizeof(T);
  if (vec16) {
    if (compact) return reduce ? pipeline_kernel<T, V16, true, true> : pipeline_kernel<T, V16, true, false>;
    return reduce ? pipeline_kernel<T, V16, false, true> : pipeline_kernel<T, V16, false, false>;
  }
  if (compact) return reduce ? pipeline_kernel<T, 1, true, true> : pipeline_kernel<T, 1, true, false>;
  return reduce ? pipeline_kernel<T, 1, false, true> : pipeline_kernel<T, 1, false, false>;
}

static pipeline_fn pick(int dtype, bool vec16, bool compact, bool reduce) {
  switch (dtype) {
    case NL_I32: return pick_T<int32_t>(vec16, compact, reduce);
    case NL_I64: return pick_T<long long>(vec16, compact, reduce);
    case NL_F32: return pick_T<float>(vec16, compact, reduce);
    case NL_F64: return pick_T<double>(vec16, compact, reduce);
  }
  return nullptr;
}

static const char *dt_name(int d) {
  static const char *n[] = {"bool", "i32", "i64", "f32", "f64"};
  return (d >= 0 && d <= 4) ? n[d] : "?";
}

#define CUDA_TRY(expr)                                                             \
  do {                                                                             \
    cudaError_t e_ = (expr);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      set_error("%s failed: %s", #expr, cudaGetErrorString(e_));                   \
      return NL_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

int pipeline_kernel_occupancy(int dtype, bool vec16, bool compact, bool reduce, int *blocks_per_sm, int *regs) {
  pipeline_fn fn = pick(dtype, vec16, compact, reduce);
  if (!fn) { set_error("no pipeline kernel for dtype %d", dtype); return NL_ERR_UNSUPPORTED; }
  int bps = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, kBlock, 0));
  cudaFuncAttributes attr;
  CUDA_TRY(cudaFuncGetAttributes(&attr, fn));
  if (blocks_per_sm) *blocks_per_sm = bps;
  if (regs) *regs = attr.numRegs;
  return NL_OK;
}

int launch_pipeline_kernel(const nl_plan *plan, const KParams &kp, bool vec16, int sm_count, void *stream, LaunchInfo *info) {
  const bool compact = (plan->flags & NL_PLAN_HAS_COMPACT) != 0;
  const bool reduce = (plan->flags & NL_PLAN_HAS_REDUCE) != 0;
  pipeline_fn fn = pick(plan->dtype, vec16, compact, reduce);
  if (!fn) { set_error("no pipeline kernel for dtype %d", plan->dtype); return NL_ERR_UNSUPPORTED; }
  int bps = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, kBlock, 0));
  if (bps < 1) bps = 1;
  int64_t grid = (int64_t)sm_count * bps;
  if (grid > kp.n_tiles) grid = kp.n_tiles;
  if (grid < 1) grid = 1;
  fn<<<(unsigned)grid, kBlock, 0, (cudaStream_t)stream>>>(kp);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  if (info) {
    const int v = vec16 ? (int)(16 / nl_dtype_size(plan->dtype)) : 1;
    snprintf(info->kernel, sizeof(info->kernel), "pipeline_kernel<%s,V=%d,compact=%d,reduce=%d>", dt_name(plan->dtype), v, (int)compact, (int)reduce);
    info->grid = (int)grid;
    info->block = kBlock;
    info->vec_elems = v;
  }
  return NL_OK;
}

int launch_empty_range(void *stream, const nl_plan *plan, const KParams &kp) {
  switch (plan->dtype) {
    case NL_I32: empty_range_kernel<int32_t><<<1, 32, 0, (cudaStream_t)stream>>>(kp); break;
    case NL_I64: empty_range_kernel<long long><<<1, 32, 0, (cudaStream_t)stream>>>(kp); break;
    case NL_F32: empty_range_kernel<float><<<1, 32, 0, (cudaStream_t)stream>>>(kp); break;
    case NL_F64: empty_range_kernel<double><<<1, 32, 0, (cudaStream_t)stream>>>(kp); break;
    default: set_error("bad dtype"); return NL_ERR_INVALID;
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

int launch_finalize_partials(void *stream, int rop, int dtype, const void *partials, int n, void *out) {
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case NL_I32: finalize_partials_kernel<int32_t><<<1, 32, 0, st>>>(rop, (const int32_t *)partials, n, (int32_t *)out); break;
    case NL_I64: finalize_partials_kernel<long long><<<1, 32, 0, st>>>(rop, (const long long *)partials, n, (long long *)out); break;
    case NL_F32: finalize_partials_kernel<float><<<1, 32, 0, st>>>(rop, (const float *)partials, n, (float *)out); break;
    case NL_F64: finalize_partials_kernel<double><<<1, 32, 0, st>>>(rop, (const double *)partials, n, (double *)out); break;
    default: set_error("bad dtype"); return NL_ERR_INVALID;
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

static unsigned grid_for(int64_t work_items, int sm_count) {
  int64_t blocks = (work_items + 255) / 256;
  const int64_t cap = (int64_t)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

template <typename T>
static void fill_dispatch(cudaStream_t st, int sm_count, void *dst, int64_t lo, int64_t step, int64_t count, uint64_t bits) {
  T value;
  memcpy(&value, &bits, sizeof(T));
  if (step == 1 && count >= 1024) {
    assign_fill_vec_kernel<T><<<grid_for(count / (16 / sizeof(T)), sm_count), 256, 0, st>>>((T *)dst + lo, count, value);
  } else {
    assign_fill_kernel<T><<<grid_for(count, sm_count), 256, 0, st>>>((T *)dst, lo, step, count, value);
  }
}

int launch_assign_fill(void *stream, int sm_count, void *dst, int dtype, int64_t lo, int64_t step, int64_t count, uint64_t bits) {
  cudaStream_t st = (cudaStream_t)stream;
  if (count <= 0) return NL_OK;
  switch (dtype) {
    case NL_BOOL: fill_dispatch<unsigned char>(st, sm_count, dst, lo, step, count, bits); break;
    case NL_I32: case NL_F32: fill_dispatch<uint32_t>(st, sm_count, dst, lo, step, count, bits); break;
    case NL_I64: case NL_F64: fill_dispatch<unsigned long long>(st, sm_count, dst, lo, step, count, bits); break;
    default: set_error("bad dtype"); return NL_ERR_INVALID;
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

int launch_assign_copy(void *stream, int sm_count, void *dst, int dtype, int64_t lo, int64_t step, int64_t count, const void *src) {
  cudaStream_t st = (cudaStream_t)stream;
  if (count <= 0) return NL_OK;
  const unsigned g = grid_for(count, sm_count);
  switch (dtype) {
    case NL_BOOL: assign_copy_kernel<unsigned char><<<g, 256, 0, st>>>((unsigned char *)dst, lo, step, count, (const unsigned char *)src); break;
    case NL_I32: case NL_F32: assign_copy_kernel<uint32_t><<<g, 256, 0, st>>>((uint32_t *)dst, lo, step, count, (const uint32_t *)src); break;
    case NL_I64: case NL_F64: assign_copy_kernel<unsigned long long><<<g, 256, 0, st>>>((unsigned long long *)dst, lo, step, count, (const unsigned long long *)src); break;
    default: set_error("bad dtype"); return NL_ERR_INVALID;
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

int launch_scan_offsets(void *stream, const int64_t *counts, int n, int64_t *offsets) {
  scan_offsets_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counts, n, offsets);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

int launch_fill(void *stream, int sm_count, void *dst, int dtype, int64_t first_index, int64_t count, uint64_t seed, int kind) {
  cudaStream_t st = (cudaStream_t)stream;
  if (count <= 0) return NL_OK;
  const unsigned g = grid_for(count, sm_count);
  switch (dtype) {
    case NL_BOOL: fill_kernel<unsigned char><<<g, 256, 0, st>>>((unsigned char *)dst, first_index, count, seed, kind); break;
    case NL_I32: fill_kernel<int32_t><<<g, 256, 0, st>>>((int32_t *)dst, first_index, count, seed, kind); break;
    case NL_I64: fill_kernel<long long><<<g, 256, 0, st>>>((long long *)dst, first_index, count, seed, kind); break;
    case NL_F32: fill_kernel<float><<<g, 256, 0, st>>>((float *)dst, first_index, count, seed, kind); break;
    case NL_F64: fill_kernel<double><<<g, 256, 0, st>>>((double *)dst, first_index, count, seed, kind); break;
    default: set_error("bad dtype"); return NL_ERR_INVALID;
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

}  // namespace nl
