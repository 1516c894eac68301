// nl_pipeline.cuh — the fused pipeline kernel template (sm_100a).
//
// In the reference every pipeline is ONE fused loop (compiler.cpp:362-509); here it
// is ONE kernel template, pipeline_kernel<T, V, kCompact, kReduce, Prog>, that runs
// the plan's accumulator-machine program (nl_step.h) over tiles of
// kBlock * kRows * V elements.  The program driver `Prog` comes in two kinds that
// share the same step executor (Tile::exec):
//
//   DynProg               reads the step words from the kernel parameters and
//                         dispatches with warp-uniform branches — the general
//                         engine for arbitrary thunk expressions;
//   StaticProg<steps...>  the step words are template arguments, so dispatch,
//                         unused temporaries and register moves fold away at build
//                         time — the "kernel templates instantiated at build time"
//                         for the pipeline shapes in nl_static.cuh.
//
// The four paths:
//   (1) elementwise chains   128-bit coalesced streaming loads (V = 16/sizeof(T)
//                            elements per access, kRows independent accesses in
//                            flight per thread per operand), every intermediate in
//                            registers, grid-stride tiles, 128-bit streaming stores
//   (2) filter compaction    per-row popc counts, packed warp scan, single-pass
//                            decoupled look-back across tiles (status + value in one
//                            64-bit word), stable order, 64-bit offsets
//   (3) max / min / sum      per-thread tree -> running accumulator -> warp shuffle
//                            tree -> block tree -> per-CTA partial -> the last CTA
//                            folds all partials in a fixed order (deterministic)
//   (4) in-place assignment  masked store at the loop index (S_WHERE)
//
// All are HBM-bound streaming kernels: no shared-memory staging of operands (no
// reuse) and no tensor cores (nothing is a contraction).  Integer arithmetic is done
// in unsigned for two's-complement wrap; translation units are compiled with
// -fmad=false so float chains round exactly like the reference's fmul/fadd sequence.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>

#include <utility>

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
// the narrow and the unsigned integer types (getLLVMType's i8 / i16, compiler.cpp:156-167, and the unsigned
// types it leaves as a todo, :168-172): every operation wraps to the column type like NumPy's, sums run in a
// 64-bit accumulator of the same signedness and are cast back on store
template <> struct TT<signed char> {
  using U = unsigned char; using Acc = long long; static constexpr bool is_float = false;
  __device__ static signed char lowest() { return SCHAR_MIN; }
  __device__ static signed char highest() { return SCHAR_MAX; }
};
template <> struct TT<short> {
  using U = unsigned short; using Acc = long long; static constexpr bool is_float = false;
  __device__ static short lowest() { return SHRT_MIN; }
  __device__ static short highest() { return SHRT_MAX; }
};
template <> struct TT<unsigned char> {
  using U = unsigned char; using Acc = unsigned long long; static constexpr bool is_float = false;
  __device__ static unsigned char lowest() { return 0; }
  __device__ static unsigned char highest() { return UCHAR_MAX; }
};
template <> struct TT<unsigned short> {
  using U = unsigned short; using Acc = unsigned long long; static constexpr bool is_float = false;
  __device__ static unsigned short lowest() { return 0; }
  __device__ static unsigned short highest() { return USHRT_MAX; }
};
template <> struct TT<unsigned int> {
  using U = unsigned int; using Acc = unsigned long long; static constexpr bool is_float = false;
  __device__ static unsigned int lowest() { return 0; }
  __device__ static unsigned int highest() { return UINT_MAX; }
};
template <> struct TT<unsigned long long> {
  using U = unsigned long long; using Acc = unsigned long long; static constexpr bool is_float = false;
  __device__ static unsigned long long lowest() { return 0; }
  __device__ static unsigned long long highest() { return ULLONG_MAX; }
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
// true division: floats only (the plan compiler rejects integer columns); IEEE, correctly rounded
template <typename T> __device__ __forceinline__ T op_div(T a, T b) {
  if constexpr (TT<T>::is_float) return a / b; else return T(0);
}
// NumPy floor_divide.  Ints: rounds towards -inf, x // 0 == 0, INT_MIN // -1 wraps.  Floats: npy_divmod
// (fmod-based, so the quotient is exact where a / b would round to the wrong side of an integer).
// (not inlined: fmod and 64-bit integer division are long sequences, and the interpreter would
// otherwise carry 64 unrolled copies of them)
template <typename T> __device__ __noinline__ T op_floordiv(T a, T b) {
  if constexpr (TT<T>::is_float) {
    if (b == T(0)) return a / b;
    T mod = fmod(a, b);
    T div = (a - mod) / b;
    if (mod != T(0) && ((b < T(0)) != (mod < T(0)))) div -= T(1);
    if (div != T(0)) {
      T fl = floor(div);
      if (div - fl > T(0.5)) fl += T(1);
      return fl;
    }
    return copysign(T(0), a / b);
  } else {
    using U = typename TT<T>::U;
    if (b == T(0)) return T(0);
    if constexpr (T(-1) < T(0)) {
      if (b == T(-1)) return (T)((U)0 - (U)a);
      T q = (T)(a / b);
      if ((a % b != T(0)) && ((a < T(0)) != (b < T(0)))) q -= T(1);
      return q;
    } else {
      return (T)(a / b);             // unsigned: truncation is the floor
    }
  }
}
// max/min with NumPy semantics: NaN is sticky
template <typename T> __device__ __forceinline__ T op_max(T cur, T l) {
  if constexpr (TT<T>::is_float) return (l > cur || l != l) ? l : cur; else return l > cur ? l : cur;
}
template <typename T> __device__ __forceinline__ T op_min(T cur, T l) {
  if constexpr (TT<T>::is_float) return (l < cur || l != l) ? l : cur; else return l < cur ? l : cur;
}
__device__ __forceinline__ long long acc_add(long long a, long long b) { return (long long)((unsigned long long)a + (unsigned long long)b); }
__device__ __forceinline__ unsigned long long acc_add(unsigned long long a, unsigned long long b) { return a + b; }
__device__ __forceinline__ double acc_add(double a, double b) { return a + b; }

// elements of one wide row of a thread (host twin: nl_vec_elems)
template <typename T> constexpr int wide_v() { return sizeof(T) >= 4 ? (int)(16 / sizeof(T)) : 4; }

// ------------------------------------------------------------------ vector access
template <int BYTES> struct VecOf;
template <> struct VecOf<16> { using type = uint4; };
template <> struct VecOf<8> { using type = uint2; };
template <> struct VecOf<4> { using type = unsigned int; };
template <> struct VecOf<2> { using type = unsigned short; };
template <> struct VecOf<1> { using type = unsigned char; };

// the same row read with normal L2 residency (ld.global.cg): used when the row will be read again
// shortly (count pass of the super-tile compaction kernel)
template <typename T, int V>
__device__ __forceinline__ void load_vec_l2(const T *__restrict__ ptr, T *dst) {
  using Vec = typename VecOf<sizeof(T) * V>::type;
  union { Vec q; T t[V]; uint4 r; } u;
#ifdef NL_L2_EVICT_LAST
  if constexpr (sizeof(Vec) == 16) {
    // L2 eviction priority "last" for rows the store pass reads again
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(u.r.x), "=r"(u.r.y), "=r"(u.r.z), "=r"(u.r.w)
                 : "l"(ptr), "l"(pol));
  } else
#endif
  {
    u.q = __ldcg(reinterpret_cast<const Vec *>(ptr));
  }
#pragma unroll
  for (int v = 0; v < V; v++) dst[v] = u.t[v];
}

// one row = V consecutive elements of one thread; streaming (evict-first) 128-bit access
template <typename T, int V>
__device__ __forceinline__ void load_vec(const T *__restrict__ ptr, T *dst) {
  using Vec = typename VecOf<sizeof(T) * V>::type;
  union { Vec q; T t[V]; } u;
  u.q = __ldcs(reinterpret_cast<const Vec *>(ptr));
#pragma unroll
  for (int v = 0; v < V; v++) dst[v] = u.t[v];
}
template <typename T, int V>
__device__ __forceinline__ void store_vec(T *__restrict__ ptr, const T *src) {
  using Vec = typename VecOf<sizeof(T) * V>::type;
  union { Vec q; T t[V]; } u;
#pragma unroll
  for (int v = 0; v < V; v++) u.t[v] = src[v];
  __stcs(reinterpret_cast<Vec *>(ptr), u.q);
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
  else if constexpr (sizeof(T) == 4) { union { uint32_t b; T t; } u; u.b = (uint32_t)bits; return u.t; }
  else if constexpr (sizeof(T) == 2) { union { uint16_t b; T t; } u; u.b = (uint16_t)bits; return u.t; }
  else { union { uint8_t b; T t; } u; u.b = (uint8_t)bits; return u.t; }
}
template <typename T> __device__ __forceinline__ unsigned long long to_bits(T v) {
  if constexpr (sizeof(T) == 8) { union { unsigned long long b; T t; } u; u.t = v; return u.b; }
  else if constexpr (sizeof(T) == 4) { union { uint32_t b; T t; } u; u.t = v; return u.b; }
  else if constexpr (sizeof(T) == 2) { union { uint16_t b; T t; } u; u.t = v; return u.b; }
  else { union { uint8_t b; T t; } u; u.t = v; return u.b; }
}

constexpr unsigned long long kDescValueMask = (1ull << 62) - 1;
constexpr unsigned long long kDescAggregate = 1ull << 62;
constexpr unsigned long long kDescPrefix = 2ull << 62;
// one tile descriptor per 128-byte line: neighbouring tiles are published and polled by different
// SMs at the same time; sharing a line between them serialises in the L2 slice
constexpr int kDescStride = NL_DESC_STRIDE;

// ------------------------------------------------------------------ NVLink mailbox all-reduce
// Fused multi-GPU finish of a reduction: the CTA that folded this GPU's partials writes the
// result straight into a mailbox slot on EVERY peer GPU (plain stores over NVLink P2P / IPC
// mappings, released with a system-scope flag), then waits until its own mailbox holds the
// slot of every rank for this epoch and folds them in rank order — the same order on every
// GPU, so all ranks end up with bit-identical results and no NCCL launch follows the kernel.
//
// Mailbox of one rank: [2 parities][world][value u64, epoch u64].  Two parities suffice: a rank
// can start epoch e + 2 only after it finished e + 1, which needed every peer's e + 1 slot,
// which a peer writes only after it finished reading epoch e.
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// A wait on a peer is BOUNDED: a rank whose launch failed (or never came) must not hang the other
// seven kernels.  Every mailbox ends in an abort word; the waiter gives up when it reads an abort for
// this epoch (or an earlier one) or when peer_timeout_ns have passed, posts the abort to every peer
// itself — so they stop, too — and raises the device's status word in pinned host memory, which the
// host turns into NL_ERR_NCCL at its next synchronisation (nl_runtime.cu: check_peer_status).
constexpr int kMailboxAbortIdx = 2 * 256 * 2;       // u64 index of the abort word: behind [2 parities][256 ranks][value, epoch]
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
static __device__ __noinline__ void peer_give_up(const KParams &p, int waiting_for) {
  for (int dst = 0; dst < p.peer_world; dst++) st_release_sys_u64(p.peer_box[dst] + kMailboxAbortIdx, p.peer_epoch);
  if (p.peer_status) {
    st_relaxed_sys_u64(p.peer_status, (p.peer_epoch & 0xffffffffffffull) | ((unsigned long long)(waiting_for + 1) << 48));
    __threadfence_system();
  }
}
// lane 0 of the exchanging warp: wait until `flag` carries this launch's epoch; false: aborted / timed out
__device__ __forceinline__ bool peer_wait(const KParams &p, const unsigned long long *flag, int src) {
  unsigned int spins = 0;
  unsigned long long t0 = 0;
  while (ld_acquire_sys_u64(flag) != p.peer_epoch) {
    if ((++spins & 255u) == 0) {
      const unsigned long long ab = ld_relaxed_sys_u64(p.peer_box[p.peer_rank] + kMailboxAbortIdx);
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      if ((ab != 0 && ab <= p.peer_epoch) || now - t0 > p.peer_timeout_ns) {
        peer_give_up(p, src);
        return false;
      }
    }
  }
  return true;
}

// Called by ONE warp (all 32 lanes); `mine` = this GPU's result as accumulator bits (Acc for
// sums, T for max/min).  Returns the global result bits in every lane.
template <typename T>
__device__ __forceinline__ unsigned long long peer_allreduce(const KParams &p, unsigned long long mine, int rop, int lane) {
  using Acc = typename TT<T>::Acc;
  using Tr = TT<T>;
  const int W = p.peer_world, r = p.peer_rank;
  const unsigned long long epoch = p.peer_epoch;
  const size_t par_off = (size_t)(epoch & 1ull) * (size_t)W * 2;
  for (int dst = lane; dst < W; dst += 32) {
    unsigned long long *slot = p.peer_box[dst] + par_off + (size_t)r * 2;
    st_relaxed_sys_u64(slot, mine);
    st_release_sys_u64(slot + 1, epoch);
  }
  const unsigned long long *own = p.peer_box[r] + par_off;
  Acc sv = 0;
  T ev = (rop == R_MIN) ? Tr::highest() : Tr::lowest();
  for (int src = 0; src < W; src++) {                     // rank order: deterministic, identical everywhere
    unsigned long long bits = 0;
    int ok = 1;
    if (lane == 0) {
      ok = peer_wait(p, own + (size_t)src * 2 + 1, src) ? 1 : 0;
      bits = ld_relaxed_sys_u64(own + (size_t)src * 2);
    }
    if (!__shfl_sync(0xffffffffu, ok, 0)) break;           // aborted: the host reports NL_ERR_NCCL, the value is void
    bits = __shfl_sync(0xffffffffu, bits, 0);
    if (rop == R_SUM) sv = acc_add(sv, from_bits<Acc>(bits));
    else if (rop == R_MAX) ev = op_max<T>(ev, from_bits<T>(bits));
    else ev = op_min<T>(ev, from_bits<T>(bits));
  }
  return (rop == R_SUM) ? to_bits<Acc>(sv) : to_bits<T>(ev);
}

// The filter's finish over the same mailboxes: every rank posts how many elements its shard
// kept; offsets[r] = where rank r's run starts in the concatenated result, offsets[world] =
// total (what the memmove gather of compiler.cpp:36-48 computes).  Called by one warp.
__device__ __forceinline__ void peer_scan_counts(const KParams &p, unsigned long long count, int lane) {
  const int W = p.peer_world, r = p.peer_rank;
  const unsigned long long epoch = p.peer_epoch;
  const size_t par_off = (size_t)(epoch & 1ull) * (size_t)W * 2;
  for (int dst = lane; dst < W; dst += 32) {
    unsigned long long *slot = p.peer_box[dst] + par_off + (size_t)r * 2;
    st_relaxed_sys_u64(slot, count);
    st_release_sys_u64(slot + 1, epoch);
  }
  if (lane == 0) {
    const unsigned long long *own = p.peer_box[r] + par_off;
    long long run = 0;
    for (int src = 0; src < W; src++) {
      if (!peer_wait(p, own + (size_t)src * 2 + 1, src)) break;      // aborted: the host reports NL_ERR_NCCL
      p.peer_offsets[src] = run;
      run += (long long)ld_relaxed_sys_u64(own + (size_t)src * 2);
    }
    p.peer_offsets[W] = run;
  }
}

// shared memory of one CTA (tiny: scan totals, the look-back result, the reduce tree)
struct TileShared {
  uint32_t wtot[kBlock / 32];          // elements each warp keeps in this tile
  uint32_t off[kBlock / 32];           // exclusive offset of each warp's run inside the tile
  long long prefix;                    // elements kept before this tile
  long long tile;                      // dynamically claimed tile id
  unsigned long long red[kBlock / 32];
  int last;
};

// ------------------------------------------------------------------ per-tile machine state
//
// Element order inside a tile is warp-blocked: warp w owns the contiguous run
// [w*32*E, (w+1)*32*E) of the tile and row u of lane l is the V consecutive elements at
// w*32*E + u*32*V + l*V — every row access of a warp is one contiguous 32*V-element segment
// (512 B for 128-bit rows), and the input order is (warp, row, lane, v).
//
// BLK is the number of threads that share a tile (the CTA, or the consumer warps of the staged
// compaction pipeline); kStaged: full tiles of column inputs are read from the shared-memory
// slot a TMA producer filled (nl_compact_pipe.cuh) instead of from global memory.
template <typename T, int V, int U, bool kCompact, bool kReduce, int BLK = kBlock, bool kStaged = false, bool kTemps = true>
struct Tile {
  static constexpr int E = U * V;
  static constexpr int NW = BLK / 32;
  static constexpr int ROWSTRIDE = 32 * V;                 // elements between consecutive rows of a lane
  static constexpr int64_t TILE = (int64_t)BLK * E;
  static constexpr uint32_t ROWMASK = (1u << V) - 1u;
  static constexpr uint32_t FULLMASK = (E == 32) ? 0xffffffffu : ((1u << E) - 1u);
  static constexpr int NWORDS = (U + 3) / 4;               // packed 8-bit per-row counters
  static constexpr int LOOKBACK = 1;                       // descriptor windows (of 32) read per round trip
  static_assert(E <= 32, "one predicate bit per element");
  static_assert(V * 32 < 256, "packed 8-bit row counters");
  using Tr = TT<T>;
  using Acc = typename Tr::Acc;

  // kTemps = false: the program never saves a value temporary (plan->n_temps == 0); the arrays
  // shrink to nothing and the interpreter fits three CTAs per SM
  T acc[E], t0[kTemps ? E : 1], t1[kTemps ? E : 1];
  uint32_t pacc, p0, p1, p2;
  uint32_t keep, valid;
  int64_t cbase;             // compacted index of this warp's run (64-bit)
  uint32_t crow[U];          // + offset of each of this lane's rows inside the run
  int64_t thr_base, tile;
  bool full;
  bool keep_in_l2 = false;      // column rows are loaded with normal L2 residency (they will be read again)
  const unsigned char *stage;   // kStaged: shared-memory slot of the current tile (column 0)
  uint32_t stage_off;           // kStaged: byte offset of this lane's row 0 inside a staged column
  // running reduction state (lives across tiles)
  Acc rsum;
  T rext;

  __device__ __forceinline__ void begin_tile(const KParams &p, int64_t tile_id, int tid) {
    tile = tile_id;
    const int64_t tile_base = p.start + tile_id * TILE;
    // block-uniform: every row of every thread is in range -> straight-line vector accesses,
    // all kRows loads of an operand are issued back to back before anything consumes them
    full = (tile_base + TILE <= p.end);
    thr_base = tile_base + (int64_t)(tid >> 5) * (32 * E) + (int64_t)(tid & 31) * V;   // row u at thr_base + u*ROWSTRIDE
    stage_off = (uint32_t)(((tid >> 5) * (32 * E) + (tid & 31) * V) * sizeof(T));
    keep = FULLMASK;
    if (!full) {
      keep = 0;
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int64_t rem = p.end - (thr_base + (int64_t)u * ROWSTRIDE);
        const int nv = rem >= V ? V : (rem > 0 ? (int)rem : 0);
        keep |= ((1u << nv) - 1u) << (u * V);
      }
    }
    valid = keep;
    pacc = p0 = p1 = p2 = 0;
#pragma unroll
    for (int e = 0; e < E; e++) acc[e] = T(0);
    if constexpr (kTemps) {
#pragma unroll
      for (int e = 0; e < E; e++) { t0[e] = T(0); t1[e] = T(0); }
    }
    cbase = 0;
#pragma unroll
    for (int u = 0; u < U; u++) crow[u] = 0;
  }

  // x <- temporary | column rows | broadcast scalar
  __device__ __forceinline__ void fetch(const KParams &p, uint32_t src, T (&x)[E]) const {
    if (src & SRC_TEMP) {
      if constexpr (kTemps) {
        if ((src & 0x7f) == 0) {
#pragma unroll
          for (int e = 0; e < E; e++) x[e] = t0[e];
        } else {
#pragma unroll
          for (int e = 0; e < E; e++) x[e] = t1[e];
        }
      }
    } else {
      const int kind = p.in_kind[src];
      if (kind == NL_IN_ARRAY) {
        if constexpr (kStaged) {
          if (full && p.stage_col[src] != 0xff) {
            // the tile's column sits in shared memory: 128-bit LDS per row, conflict-free
            const T *sbase = reinterpret_cast<const T *>(stage + (size_t)p.stage_col[src] * (size_t)(TILE * sizeof(T)) + stage_off);
#pragma unroll
            for (int u = 0; u < U; u++) {
              using Vec = typename VecOf<sizeof(T) * V>::type;
              union { Vec q; T t[V]; } w;
              w.q = *reinterpret_cast<const Vec *>(sbase + u * ROWSTRIDE);
#pragma unroll
              for (int v = 0; v < V; v++) x[u * V + v] = w.t[v];
            }
            return;
          }
        }
        const T *base = reinterpret_cast<const T *>(p.in[src]) + thr_base;
        if (full) {
          if (keep_in_l2) {
#pragma unroll
            for (int u = 0; u < U; u++) load_vec_l2<T, V>(base + u * ROWSTRIDE, &x[u * V]);
          } else {
#pragma unroll
            for (int u = 0; u < U; u++) load_vec<T, V>(base + u * ROWSTRIDE, &x[u * V]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < E; e++) x[e] = ((valid >> e) & 1u) ? base[(e / V) * ROWSTRIDE + (e % V)] : T(0);
        }
      } else {
        const T sc = (kind == NL_IN_SCALAR_IMM) ? from_bits<T>(p.imm[src]) : *reinterpret_cast<const T *>(p.in[src]);
#pragma unroll
        for (int e = 0; e < E; e++) x[e] = sc;
      }
    }
  }

  // crow[u] <- rank of this thread's first kept element of row u inside the warp's run (rows in
  // order, lanes in order inside a row); returns the warp's total.  Per-row counts travel as
  // packed 8-bit fields, four rows per word, so one shuffle scan serves four rows.
  __device__ __forceinline__ uint32_t rank_rows(int lane) {
    (void)lane;
    uint32_t packed[NWORDS], incl[NWORDS];
#pragma unroll
    for (int w = 0; w < NWORDS; w++) {
      // per-row popcounts of four rows at once, then spread into 8-bit fields
      const int rows = (U - 4 * w >= 4) ? 4 : (U - 4 * w);
      const uint32_t x = (keep >> (4 * w * V)) & ((rows * V >= 32) ? 0xffffffffu : ((1u << (rows * V)) - 1u));
      if constexpr (V == 1) {
        packed[w] = (x | (x << 7) | (x << 14) | (x << 21)) & 0x01010101u;
      } else if constexpr (V == 2) {
        const uint32_t c = x - ((x >> 1) & 0x55u);
        packed[w] = (c | (c << 6) | (c << 12) | (c << 18)) & 0x03030303u;
      } else if constexpr (V == 4) {
        uint32_t c = x - ((x >> 1) & 0x5555u);
        c = (c & 0x3333u) + ((c >> 2) & 0x3333u);
        packed[w] = (c & 0x0000000fu) | ((c << 4) & 0x00000f00u) | ((c << 8) & 0x000f0000u) | ((c << 12) & 0x0f000000u);
      } else {
        packed[w] = 0;
#pragma unroll
        for (int r = 0; r < 4; r++)
          if (4 * w + r < U) packed[w] |= (uint32_t)__popc((keep >> ((4 * w + r) * V)) & ROWMASK) << (8 * r);
      }
    }
#pragma unroll
    for (int w = 0; w < NWORDS; w++) {
      incl[w] = packed[w];
      // inclusive scan over the lanes; the shuffle's own predicate (source lane in range) guards
      // the add, so no lane compare is needed
#pragma unroll
      for (int d = 1; d < 32; d <<= 1)
        asm volatile("{\n\t.reg .u32 r;\n\t.reg .pred p;\n\tshfl.sync.up.b32 r|p, %0, %1, 0x0, 0xffffffff;\n\t@p add.u32 %0, %0, r;\n\t}"
                     : "+r"(incl[w])
                     : "r"(d));
    }
    // row bases inside the warp's run: running sum of the row totals (lane 31's inclusive counts)
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < NWORDS; w++) {
      const uint32_t totw = __shfl_sync(0xffffffffu, incl[w], 31);
      const uint32_t exw = incl[w] - packed[w];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (4 * w + r < U) {
          crow[4 * w + r] = run + ((exw >> (8 * r)) & 0xffu);
          run += (totw >> (8 * r)) & 0xffu;
        }
      }
    }
    return run;
  }

  // `if (!mask) goto for.inc` (thunk_as_mapping.cpp:11) + the order-preserving compaction index
  __device__ __forceinline__ void filter_step(const KParams &p, TileShared &sh, int lane, int warp) {
    keep &= pacc;
    if constexpr (kCompact) {
      // rank of every kept element in (warp, row, lane, v) order == its input order
      const uint32_t run = rank_rows(lane);
      if (lane == 0) sh.wtot[warp] = run;
      __syncthreads();
      if (warp == 0) {
        const uint32_t val = (lane < NW) ? sh.wtot[lane] : 0u;
        uint32_t sc = val;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
          const uint32_t nb = __shfl_up_sync(0xffffffffu, sc, d);
          if (lane >= d) sc += nb;
        }
        if (lane < NW) sh.off[lane] = sc - val;
        const unsigned long long agg = __shfl_sync(0xffffffffu, sc, NW - 1);
        // decoupled look-back: LOOKBACK windows of 32 predecessor descriptors are fetched per
        // round trip (the window, not the tile, limits how fast the prefix frontier can advance)
        unsigned long long excl = 0;
        if (tile == 0) {
          if (lane == 0) st_relaxed_u64(&p.tile_desc[0], kDescPrefix | agg);
        } else {
          if (lane == 0) st_relaxed_u64(&p.tile_desc[tile * kDescStride], kDescAggregate | agg);
          int64_t j = tile - 1;
          bool done = false;
          while (!done) {
            unsigned long long d[LOOKBACK];
#pragma unroll
            for (int k = 0; k < LOOKBACK; k++) {
              const int64_t q = j - lane - 32 * k;
              d[k] = (q >= 0) ? ld_relaxed_u64(&p.tile_desc[q * kDescStride]) : kDescPrefix;   // before tile 0: an empty prefix
            }
#pragma unroll
            for (int k = 0; k < LOOKBACK; k++) {
              if (done) break;
              const int64_t q = j - lane - 32 * k;
              while ((d[k] >> 62) == 0) {
                d[k] = ld_relaxed_u64(&p.tile_desc[q * kDescStride]);
              }
              const uint32_t has_prefix = __ballot_sync(0xffffffffu, (d[k] >> 62) == 2);
              const int first = has_prefix ? (__ffs(has_prefix) - 1) : 31;
              unsigned long long v = (lane <= first) ? (d[k] & kDescValueMask) : 0ull;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
              excl += v;
              if (has_prefix) done = true;
            }
            j -= 32 * LOOKBACK;
          }
          if (lane == 0) st_relaxed_u64(&p.tile_desc[tile * kDescStride], kDescPrefix | ((excl + agg) & kDescValueMask));
        }
        if (lane == 0) {
          sh.prefix = (long long)excl;
          if (tile == p.n_tiles - 1) {
            for (int o = 0; o < p.n_outputs; o++)
              if (p.out_space[o] == NL_IDX_COMPACT) p.out_sizes[o] = p.start + (long long)(excl + agg);
          }
        }
        // multi-GPU: the shard's count goes to the peers from inside the kernel
        if (tile == p.n_tiles - 1 && p.peer_offsets != nullptr) peer_scan_counts(p, excl + agg, lane);
      }
      __syncthreads();
      cbase = p.start + sh.prefix + (int64_t)sh.off[warp];
    }
  }

  __device__ __forceinline__ bool is_scalar(const KParams &p, uint32_t src) const {
    return !(src & SRC_TEMP) && p.in_kind[src] != NL_IN_ARRAY;
  }
  __device__ __forceinline__ T scalar_of(const KParams &p, uint32_t src) const {
    return (p.in_kind[src] == NL_IN_SCALAR_IMM) ? from_bits<T>(p.imm[src]) : *reinterpret_cast<const T *>(p.in[src]);
  }
  // kInlineOnly (shape families): the operator is a run-time field, floor division never reaches here
  // (family_step_ok) — its out-of-line sequence would cost every family kernel registers for nothing
  template <bool kInlineOnly = false, class X>
  __device__ __forceinline__ void bin_apply(uint32_t vop, X x) {
    if constexpr (kInlineOnly) {
      switch (vop) {
        case V_MUL:
#pragma unroll
          for (int e = 0; e < E; e++) acc[e] = op_mul<T>(acc[e], x(e));
          break;
        case V_ADD:
#pragma unroll
          for (int e = 0; e < E; e++) acc[e] = op_add<T>(acc[e], x(e));
          break;
        case V_SUB:
#pragma unroll
          for (int e = 0; e < E; e++) acc[e] = op_sub<T>(acc[e], x(e));
          break;
        case V_RSUB:
#pragma unroll
          for (int e = 0; e < E; e++) acc[e] = op_sub<T>(x(e), acc[e]);
          break;
        case V_DIV:
#pragma unroll
          for (int e = 0; e < E; e++) acc[e] = op_div<T>(acc[e], x(e));
          break;
        default:
#pragma unroll
          for (int e = 0; e < E; e++) acc[e] = op_div<T>(x(e), acc[e]);
          break;
      }
      return;
    }
    switch (vop) {
      case V_MUL:
#pragma unroll
        for (int e = 0; e < E; e++) acc[e] = op_mul<T>(acc[e], x(e));
        break;
      case V_ADD:
#pragma unroll
        for (int e = 0; e < E; e++) acc[e] = op_add<T>(acc[e], x(e));
        break;
      case V_SUB:
#pragma unroll
        for (int e = 0; e < E; e++) acc[e] = op_sub<T>(acc[e], x(e));
        break;
      case V_RSUB:
#pragma unroll
        for (int e = 0; e < E; e++) acc[e] = op_sub<T>(x(e), acc[e]);
        break;
      case V_DIV:
#pragma unroll
        for (int e = 0; e < E; e++) acc[e] = op_div<T>(acc[e], x(e));
        break;
      case V_RDIV:
#pragma unroll
        for (int e = 0; e < E; e++) acc[e] = op_div<T>(x(e), acc[e]);
        break;
      case V_FDIV:
#pragma unroll
        for (int e = 0; e < E; e++) acc[e] = op_floordiv<T>(acc[e], x(e));
        break;
      default:
#pragma unroll
        for (int e = 0; e < E; e++) acc[e] = op_floordiv<T>(x(e), acc[e]);
        break;
    }
  }
  template <class X>
  __device__ __forceinline__ uint32_t cmp_apply(uint32_t cop, X x) const {
    uint32_t m = 0;
    switch (cop) {
      case C_GE:
#pragma unroll
        for (int e = 0; e < E; e++) m |= (uint32_t)(acc[e] >= x(e)) << e;
        break;
      case C_GT:
#pragma unroll
        for (int e = 0; e < E; e++) m |= (uint32_t)(acc[e] > x(e)) << e;
        break;
      case C_LE:
#pragma unroll
        for (int e = 0; e < E; e++) m |= (uint32_t)(acc[e] <= x(e)) << e;
        break;
      case C_LT:
#pragma unroll
        for (int e = 0; e < E; e++) m |= (uint32_t)(acc[e] < x(e)) << e;
        break;
      case C_EQ:
#pragma unroll
        for (int e = 0; e < E; e++) m |= (uint32_t)(acc[e] == x(e)) << e;
        break;
      default:
#pragma unroll
        for (int e = 0; e < E; e++) m |= (uint32_t)(acc[e] != x(e)) << e;
        break;
    }
    return m;
  }

  // One step of the accumulator machine.  With compile-time constant (op, a, b) — the
  // StaticProg driver — everything but the selected case folds away.
  // kValuesOnly (phase B of the staged compaction pipeline): the keep mask is already known, so
  // steps that only produce predicates are skipped.
  // kOperand (shape families): the BIN / CMP operand is known to be a scalar (FK_SCALAR) or a column
  // (FK_COLUMN) at build time — a scalar-only step then carries no second register set; FK_EITHER decides
  // at run time like the interpreter
  template <bool kValuesOnly = false, bool kInlineOnly = false, uint32_t kOperand = FK_EITHER>
  __device__ __forceinline__ void exec(const KParams &p, TileShared &sh, uint32_t op, uint32_t a, uint32_t b,
                                       int lane, int warp) {
    switch (op) {
      case S_LOAD:
        if (is_scalar(p, a)) {
          const T sc = scalar_of(p, a);
#pragma unroll
          for (int e = 0; e < E; e++) acc[e] = sc;
        } else {
          fetch(p, a, acc);
        }
        break;
      case S_LOADT:
        fetch(p, SRC_TEMP | a, acc);
        break;
      case S_SAVE:
        if constexpr (kTemps) {
          if (a == 0) {
#pragma unroll
            for (int e = 0; e < E; e++) t0[e] = acc[e];
          } else {
#pragma unroll
            for (int e = 0; e < E; e++) t1[e] = acc[e];
          }
        }
        break;
      case S_BIN:
        // a scalar operand is used straight from its (uniform) register: broadcasting it into E
        // registers first cost 16 moves per step
        if (kOperand == FK_SCALAR || (kOperand == FK_EITHER && is_scalar(p, b))) {
          const T sc = scalar_of(p, b);
          bin_apply<kInlineOnly>(a, [&](int) { return sc; });
        } else {
          T x[E];
          fetch(p, b, x);
          bin_apply<kInlineOnly>(a, [&](int e) { return x[e]; });
        }
        break;
      case S_CMP:
        if constexpr (!kValuesOnly) {
          if (kOperand == FK_SCALAR || (kOperand == FK_EITHER && is_scalar(p, b))) {
            const T sc = scalar_of(p, b);
            pacc = cmp_apply(a, [&](int) { return sc; });
          } else {
            T x[E];
            fetch(p, b, x);
            pacc = cmp_apply(a, [&](int e) { return x[e]; });
          }
        }
        break;
      case S_PLOAD: {
        if constexpr (kValuesOnly) break;
        const int kind = p.in_kind[a];
        uint32_t m = 0;
        if (kind == NL_IN_ARRAY) {
          const unsigned char *base = reinterpret_cast<const unsigned char *>(p.in[a]) + thr_base;
          if (full) {
            unsigned char bytes[E];
#pragma unroll
            for (int u = 0; u < U; u++) load_vec<unsigned char, V>(base + u * ROWSTRIDE, &bytes[u * V]);
#pragma unroll
            for (int e = 0; e < E; e++) m |= (uint32_t)(bytes[e] != 0) << e;
          } else {
#pragma unroll
            for (int e = 0; e < E; e++)
              if ((valid >> e) & 1u) m |= (uint32_t)(base[(e / V) * ROWSTRIDE + (e % V)] != 0) << e;
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
        filter_step(p, sh, lane, warp);
        break;
      case S_STORE: {
        if (b == NL_IDX_LOOP) {
          T *base = reinterpret_cast<T *>(p.out[a]) + thr_base;
          if (full) {
#pragma unroll
            for (int u = 0; u < U; u++) store_vec<T, V>(base + u * ROWSTRIDE, &acc[u * V]);
          } else {
#pragma unroll
            for (int e = 0; e < E; e++)
              if ((valid >> e) & 1u) base[(e / V) * ROWSTRIDE + (e % V)] = acc[e];
          }
        } else {
          T *base = reinterpret_cast<T *>(p.out[a]) + cbase;
#pragma unroll
          for (int u = 0; u < U; u++) {
            uint32_t ci = crow[u];
#pragma unroll
            for (int v = 0; v < V; v++)
              if ((keep >> (u * V + v)) & 1u) __stcs(&base[ci++], acc[u * V + v]);   // streaming: must not evict rows waiting in L2 for the store pass
          }
        }
        break;
      }
      case S_PSTORE: {
        if (b == NL_IDX_LOOP) {
          unsigned char *base = reinterpret_cast<unsigned char *>(p.out[a]) + thr_base;
          if (full) {
#pragma unroll
            for (int u = 0; u < U; u++) {
              unsigned char bytes[V];
#pragma unroll
              for (int v = 0; v < V; v++) bytes[v] = (unsigned char)((pacc >> (u * V + v)) & 1u);   // canonical 0x01 (Q4)
              store_vec<unsigned char, V>(base + u * ROWSTRIDE, bytes);
            }
          } else {
#pragma unroll
            for (int e = 0; e < E; e++)
              if ((valid >> e) & 1u) base[(e / V) * ROWSTRIDE + (e % V)] = (unsigned char)((pacc >> e) & 1u);
          }
        } else {
          unsigned char *base = reinterpret_cast<unsigned char *>(p.out[a]) + cbase;
#pragma unroll
          for (int u = 0; u < U; u++) {
            uint32_t ci = crow[u];
#pragma unroll
            for (int v = 0; v < V; v++)
              if ((keep >> (u * V + v)) & 1u) base[ci++] = (unsigned char)((pacc >> (u * V + v)) & 1u);
          }
        }
        break;
      }
      case S_WHERE: {
        T *base = reinterpret_cast<T *>(p.out[a]) + thr_base;
        const uint32_t m = pacc & keep;
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint32_t rowbits = (m >> (u * V)) & ROWMASK;
          if (rowbits == ROWMASK) {
            store_vec<T, V>(base + u * ROWSTRIDE, &acc[u * V]);
          } else {
#pragma unroll
            for (int v = 0; v < V; v++)
              if ((rowbits >> v) & 1u) base[u * ROWSTRIDE + v] = acc[u * V + v];
          }
        }
        break;
      }
      case S_REDUCE:
        if constexpr (kReduce) {
          // fold this thread's E elements in a fixed tree first (short dependency chain),
          // then one update of the running accumulator
          const bool all = (keep == FULLMASK);
          if (a == R_SUM) {
            Acc w[E];
            if (all) {
#pragma unroll
              for (int e = 0; e < E; e++) w[e] = (Acc)acc[e];
            } else {
#pragma unroll
              for (int e = 0; e < E; e++) w[e] = ((keep >> e) & 1u) ? (Acc)acc[e] : Acc(0);
            }
#pragma unroll
            for (int st = E / 2; st > 0; st >>= 1)
#pragma unroll
              for (int i = 0; i < st; i++) w[i] = acc_add(w[i], w[i + st]);
            rsum = acc_add(rsum, w[0]);
          } else if (a == R_MAX) {
            T w[E];
            if (all) {
#pragma unroll
              for (int e = 0; e < E; e++) w[e] = acc[e];
            } else {
#pragma unroll
              for (int e = 0; e < E; e++) w[e] = ((keep >> e) & 1u) ? acc[e] : Tr::lowest();
            }
#pragma unroll
            for (int st = E / 2; st > 0; st >>= 1)
#pragma unroll
              for (int i = 0; i < st; i++) w[i] = op_max<T>(w[i], w[i + st]);
            rext = op_max<T>(rext, w[0]);
          } else {
            T w[E];
            if (all) {
#pragma unroll
              for (int e = 0; e < E; e++) w[e] = acc[e];
            } else {
#pragma unroll
              for (int e = 0; e < E; e++) w[e] = ((keep >> e) & 1u) ? acc[e] : Tr::highest();
            }
#pragma unroll
            for (int st = E / 2; st > 0; st >>= 1)
#pragma unroll
              for (int i = 0; i < st; i++) w[i] = op_min<T>(w[i], w[i + st]);
            rext = op_min<T>(rext, w[0]);
          }
        }
        break;
      default:
        break;
    }
  }
};

// ------------------------------------------------------------------ program drivers
template <bool kUsesTemps>
struct DynProgT {
  static constexpr bool is_static = false;
  static constexpr int min_blocks_stream = 1;
  static constexpr bool uses_temps = kUsesTemps;
  template <class TileT>
  static __device__ __forceinline__ void run(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
#pragma unroll 1
    for (int pc = 0; pc < p.n_steps; pc++) {
      const uint32_t s = p.step[pc];
      t.exec(p, sh, step_op(s), step_a(s), step_b(s), lane, warp);
    }
  }
};
using DynProg = DynProgT<true>;        // the general interpreter: two value temporaries in registers
using DynProgLight = DynProgT<false>;  // plans without value temporaries: fewer registers, 3 CTAs per SM

template <uint32_t... S>
struct StaticProg {
  static constexpr bool is_static = true;
  static constexpr bool uses_temps = true;
  static constexpr int min_blocks_stream = 1;        // resident CTAs per SM the register allocation must allow (non-compaction)
  static constexpr int n_steps = sizeof...(S);
  template <class TileT>
  static __device__ __forceinline__ void run(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
    (t.exec(p, sh, step_op(S), step_a(S), step_b(S), lane, warp), ...);
  }
  static bool matches(const nl_plan *plan) {
    constexpr uint32_t mine[] = {S...};
    if (plan->n_steps != n_steps) return false;
    for (int i = 0; i < n_steps; i++)
      if (plan->step[i] != mine[i]) return false;
    return true;
  }
};

// A shape FAMILY (nl_step.h: family_key): the step kinds, temporaries and stores are template
// arguments like StaticProg's, while operators and input indices are read from the launch's own step
// words — one build-time kernel for every expression of the same shape.
template <uint32_t... K>
struct FamilyProg {
  static constexpr bool is_static = true;
  // an "either" operand keeps a column-sized register set in reserve at every step: cap the allocation so that
  // two CTAs stay resident (the exact shapes need 84 registers for the same rows)
  static constexpr int min_blocks_stream = 2;
  static constexpr int n_steps = sizeof...(K);
  static constexpr bool uses_temps = ((step_op(K) == S_SAVE) || ...);
  NL_HD static constexpr uint32_t key_at(int i) {
    constexpr uint32_t k[] = {K...};
    return k[i];
  }
  template <int I, bool kValuesOnly, class TileT>
  static __device__ __forceinline__ void step(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
    constexpr uint32_t k = key_at(I);
    constexpr uint32_t op = step_op(k);
    const uint32_t s = p.step[I];                    // this launch's operators and inputs
    uint32_t a = step_a(k), b = step_b(k);
    if constexpr (op == S_LOAD || op == S_PLOAD || op == S_REDUCE || op == S_PBIN) a = step_a(s);
    if constexpr ((op == S_BIN || op == S_CMP) && !(step_b(k) & SRC_TEMP)) {
      a = step_a(s);
      b = step_b(s);
      t.template exec<kValuesOnly, true, step_b(k)>(p, sh, op, a, b, lane, warp);      // scalar / column / either: from the key
    } else {
      if constexpr (op == S_BIN || op == S_CMP) a = step_a(s);
      t.template exec<kValuesOnly, true>(p, sh, op, a, b, lane, warp);
    }
  }
  template <class TileT, size_t... I>
  static __device__ __forceinline__ void run_seq(TileT &t, const KParams &p, TileShared &sh, int lane, int warp, std::index_sequence<I...>) {
    (step<(int)I, false>(t, p, sh, lane, warp), ...);
  }
  template <class TileT>
  static __device__ __forceinline__ void run(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
    run_seq(t, p, sh, lane, warp, std::make_index_sequence<n_steps>{});
  }
  static bool matches(const nl_plan *plan) {
    constexpr uint32_t mine[] = {K...};
    if (plan->n_steps != n_steps) return false;
    for (int i = 0; i < n_steps; i++) {
      const uint32_t s = plan->step[i], op = step_op(s);
      const bool column = (op == S_BIN || op == S_CMP) && !(step_b(s) & SRC_TEMP) && plan->in[step_b(s)].kind == NL_IN_ARRAY;
      if (!family_key_fits(family_key(s, column), mine[i]) || !family_step_ok(s)) return false;
    }
    return true;
  }
};

// ------------------------------------------------------------------ the kernel
template <typename T, int V, int U, bool kCompact, bool kReduce, class Prog>
__global__ void __launch_bounds__(kBlock, Prog::is_static ? (kCompact ? kMinBlocksCompact : Prog::min_blocks_stream) : (Prog::uses_temps ? kMinBlocks : kMinBlocksLight))
pipeline_kernel(const __grid_constant__ KParams p) {
  using TileT = Tile<T, V, U, kCompact, kReduce, kBlock, false, Prog::uses_temps>;
  using Tr = TT<T>;
  using Acc = typename Tr::Acc;
  constexpr int NW = kBlock / 32;
  __shared__ TileShared sh;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // output cardinalities that do not depend on the data (compiler.cpp:491-508)
  if (blockIdx.x == 0 && tid == 0) {
    for (int o = 0; o < p.n_outputs; o++) {
      if (p.out_space[o] == NL_IDX_LOOP || p.out_space[o] == NL_IDX_WHERE) p.out_sizes[o] = p.end;
      else if (p.out_space[o] == NL_IDX_SHARD) p.out_sizes[o] = 1;
    }
  }

  TileT t;
  t.rsum = 0;
  t.rext = (p.reduce_rop == R_MIN) ? Tr::highest() : Tr::lowest();

  // Plain modes stride the tiles over the grid.  Compaction hands them out with a ticket counter:
  // tiles are claimed in order, so every predecessor of a tile is owned by a running CTA and the
  // look-back never waits on work that has not started.  A ticket is taken only when the CTA is
  // ready to start the tile: measured on B200, claiming one tile ahead (to prefetch it) is ~7x
  // slower, because the parked tiles chain the CTAs' look-backs into one serial dependency.
  int64_t tile = blockIdx.x;
  while (true) {
    if constexpr (kCompact) {
      if (tid == 0) sh.tile = (long long)atomicAdd(p.tile_ticket, 1u);
      __syncthreads();
      tile = sh.tile;
    }
    if (tile >= p.n_tiles) break;
    if constexpr (!Prog::is_static && !kCompact && Prog::uses_temps) {
      // (general interpreter only: measured +15 % there, -8 % on the light variant, whose third
      // resident CTA already covers the latency)
      // The interpreter issues the loads of one operand at a time (a step cannot start before
      // the previous one is decoded), so DRAM latency is exposed once per step.  The rows this
      // CTA will read in its NEXT tile are requested into L2 now: one prefetch per 128-byte
      // line, while the current tile computes.
      const int64_t nt = tile + gridDim.x;
      if (nt < p.n_tiles - 1 && (lane & 7) == 0) {
        const int64_t first = p.start + nt * TileT::TILE + (int64_t)warp * (32 * TileT::E) + (int64_t)lane * V;
        for (int k = 0; k < p.n_inputs; k++) {
          if (p.in_kind[k] != NL_IN_ARRAY) continue;
          const T *col = reinterpret_cast<const T *>(p.in[k]) + first;
#pragma unroll
          for (int u = 0; u < U; u++) asm volatile("prefetch.global.L2 [%0];" ::"l"(col + u * TileT::ROWSTRIDE));
        }
      }
    }
    t.begin_tile(p, tile, tid);
    Prog::run(t, p, sh, lane, warp);
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
      if (lane == 0) sh.red[warp] = (rop == R_SUM) ? to_bits<Acc>(sv) : to_bits<T>(ev);
      __syncthreads();
      unsigned long long out = 0;
      if (warp == 0) {
        if (rop == R_SUM) {
          Acc v = (lane < NW) ? from_bits<Acc>(sh.red[lane]) : Acc(0);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v = acc_add(v, __shfl_xor_sync(0xffffffffu, v, o));
          out = to_bits<Acc>(v);
        } else {
          T v = (lane < NW) ? from_bits<T>(sh.red[lane]) : (rop == R_MAX ? Tr::lowest() : Tr::highest());
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
            v = (rop == R_MAX) ? op_max<T>(v, __shfl_xor_sync(0xffffffffu, v, o)) : op_min<T>(v, __shfl_xor_sync(0xffffffffu, v, o));
          out = to_bits<T>(v);
        }
      }
      return out;   // valid in warp 0
    };

    const unsigned long long mine = block_reduce(t.rsum, t.rext);
    if (tid == 0) {
      p.red_partials[blockIdx.x] = mine;
      __threadfence();
      const unsigned int tk = atomicAdd(p.red_ticket, 1u);
      sh.last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (sh.last) {
      __threadfence();
      Acc sv = 0;
      T ev = (rop == R_MIN) ? Tr::highest() : Tr::lowest();
      for (unsigned int i = tid; i < gridDim.x; i += kBlock) {
        const unsigned long long bits = ld_relaxed_u64(&p.red_partials[i]);
        if (rop == R_SUM) sv = acc_add(sv, from_bits<Acc>(bits));
        else if (rop == R_MAX) ev = op_max<T>(ev, from_bits<T>(bits));
        else ev = op_min<T>(ev, from_bits<T>(bits));
      }
      unsigned long long total = block_reduce(sv, ev);
      // multi-GPU: exchange with the peers over NVLink inside this kernel (warp 0 holds `total`)
      if (p.peer_box != nullptr && warp == 0) total = peer_allreduce<T>(p, total, rop, lane);
      if (tid == 0) {
        T *out = reinterpret_cast<T *>(p.out[p.reduce_out]);
        // the store casts the accumulator to the column type (compiler.cpp:256-260)
        out[p.thread_nr] = (rop == R_SUM) ? (T)from_bits<Acc>(total) : from_bits<T>(total);
        *p.red_ticket = 0u;
      }
    }
  }
}

// ------------------------------------------------------------------ static program registry
typedef void (*pipeline_fn)(const KParams);

struct StaticEntry {
  const char *name;
  bool (*matches)(const nl_plan *plan);
  int dtype;
  bool vec16;
  int rows;            // U of the instantiation: rows per thread per tile
  pipeline_fn fn;
  pipeline_fn pipe;    // staged compaction pipeline (nl_compact_pipe.cuh) or nullptr
  int pipe_rows;       // ... and the rows per thread it was instantiated for
  pipeline_fn super;   // super-tile compaction kernel (nl_compact_super.cuh) or nullptr
  int super_rows;
};

// per-dtype tables, defined in nl_static_<dtype>.cu
const StaticEntry *static_table_i32(int *n);
const StaticEntry *static_table_i64(int *n);
const StaticEntry *static_table_f32(int *n);
const StaticEntry *static_table_f64(int *n);
// per-dtype interpreter-driven instantiations, defined in nl_dynamic_<dtype>.cu (the staged pipeline only for
// the 4- and 8-byte types: its TMA ring is laid out in 16-byte rows)
#define NL_DECLARE_DYNAMIC(dt, T, sfx)                                                     \
  pipeline_fn dynamic_kernel_##sfx(bool vec16, bool compact, bool reduce, bool uses_temps); \
  pipeline_fn dynamic_pipe_##sfx(int rows);                                                 \
  pipeline_fn dynamic_super_##sfx(bool uses_temps);
NL_FOR_VALUE_DTYPES(NL_DECLARE_DYNAMIC)
#undef NL_DECLARE_DYNAMIC
const StaticEntry *static_table_i32_b(int *n);
const StaticEntry *static_table_i64_b(int *n);
const StaticEntry *static_table_f32_b(int *n);
const StaticEntry *static_table_f64_b(int *n);
// shape families (nl_family.cuh), two halves per dtype, defined in nl_family_<dtype>[_b].cu
const StaticEntry *family_table_i32(int *n);
const StaticEntry *family_table_i64(int *n);
const StaticEntry *family_table_f32(int *n);
const StaticEntry *family_table_f64(int *n);
const StaticEntry *family_table_i32_b(int *n);
const StaticEntry *family_table_i64_b(int *n);
const StaticEntry *family_table_f32_b(int *n);
const StaticEntry *family_table_f64_b(int *n);

}  // namespace nl
