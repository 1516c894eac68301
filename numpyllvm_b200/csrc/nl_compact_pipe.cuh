// nl_compact_pipe.cuh — warp-specialised compaction pipeline (sm_100a).
//
// Why it exists (measured on B200): in the one-CTA-per-tile-at-a-time
// compaction kernel every tile pays ticket -> DRAM loads -> look-back -> stores serially
// (~4-5 us per 32 KB tile); with the registers holding the tile that caps the kernel at
// 0.5-0.7 of the HBM roofline.  Claiming tiles early to prefetch them is far worse, because a
// parked tile delays the publication of its aggregate and every later tile's look-back needs it.
//
// This kernel removes the serialisation without delaying aggregates.  One persistent CTA per SM:
//
//   producer warp   claims tile tickets in order and streams each tile's columns into a ring
//                   of shared-memory slots with TMA bulk copies (cp.async.bulk + mbarrier
//                   complete_tx) — DRAM traffic never stops for a look-back;
//   consumer warps  phase A of slot k: run the pipeline up to the filter from shared memory,
//                   count the survivors of their own rows; the LAST warp to finish publishes
//                   the tile aggregate to the global descriptor straight away.
//                   phase B of slot k - lag: once the prefix is known, recompute from shared
//                   memory, rank and store the compacted values.  Warps never meet at a barrier;
//   scan warps      for slots in turn: exclusive offsets of the consumer warps, the decoupled
//                   look-back over the global descriptors, publish the inclusive prefix, hand the
//                   tile's base offset to phase B.
//
// Tiles waiting for their prefix are parked in shared memory (the ring), not in registers, and the
// aggregate of a tile is public one phase-A after its data lands.  Progress: tickets are claimed
// in order; phase A of a tile depends on nothing but its own data; the globally oldest tile
// waiting for a prefix therefore always finds every earlier aggregate published.
#pragma once
#include <cstdio>
#include <type_traits>
#include <utility>

#include "nl_pipeline.cuh"

namespace nl {

#ifndef NL_PIPE_WARPS
#define NL_PIPE_WARPS 8
#endif
#ifndef NL_PIPE_ROWS
#define NL_PIPE_ROWS 8
#endif
constexpr int kPipeConsumerWarps = NL_PIPE_WARPS;
#ifndef NL_PIPE_LOOKBACK
#define NL_PIPE_LOOKBACK 1
#endif
constexpr int kPipeLookback = NL_PIPE_LOOKBACK;  // windows of 32 predecessor descriptors fetched per look-back round trip
constexpr int kPipeRows = NL_PIPE_ROWS;          // 16-byte vectors per consumer thread and tile
constexpr int kPipeScanWarps = 4;
constexpr int kPipeConsumers = kPipeConsumerWarps * 32;                       // 512
constexpr int kPipeThreads = kPipeConsumers + 32 + kPipeScanWarps * 32;      // + producer + scan
constexpr int kPipeMaxSlots = 8;
#ifndef NL_PIPE_WAIT_NS
#define NL_PIPE_WAIT_NS 2000
#endif
constexpr uint32_t kPipeWaitHintNs = NL_PIPE_WAIT_NS;   // suspend-time hint of the mbarrier waits
constexpr int kPipeArriveShift = 20;          // SlotMeta::agg: survivors below this bit, warp arrivals above
constexpr int kPipeMeta = 16;                // bookkeeping ring (power of two), see SlotMeta
constexpr int kPipeRowsNarrow = 4;               // ... of the instantiation for two or three staged columns (half the slot size)
constexpr int pipe_side_bytes(int rows) { return 96 * rows; }   // side buffer of one consumer warp per lag entry: 3/16 of its rows

// ---- mbarrier / TMA bulk copy (PTX) ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or
// `ns` elapse, instead of returning to a spin loop that eats issue slots of the working warps
__device__ __forceinline__ bool mbar_try_wait_for(uint64_t *bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait_for(bar, parity, kPipeWaitHintNs)) {
  }
}
// global -> shared bulk copy; completion is signalled on `bar` as transferred bytes
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// per-tile bookkeeping in shared memory.  It lives in its own ring of kPipeMeta entries, indexed
// by the iteration, so that a data slot can go back to the producer before its tile is stored.
// An entry is rewritten kPipeMeta iterations later; a consumer warp that released data slot
// k - S has finished every phase up to B(k - S - 1 - lag), so S + lag + 1 <= kPipeMeta keeps the
// entry of iteration k - kPipeMeta out of use when the producer reaches iteration k.
struct SlotMeta {
  long long tile;                       // tile id (>= n_tiles: end of work)
  long long prefix;                     // elements kept before this tile (written by the scan warp)
  unsigned int agg;                     // survivors of the tile so far | warps done with phase A << kPipeArriveShift
  unsigned int pad;
  unsigned int wtot[kPipeConsumerWarps];   // per consumer warp: elements kept
  unsigned int off[kPipeConsumerWarps];    // exclusive scan of wtot (scan warp)
};

struct PipeShared {
  uint64_t full[kPipeMaxSlots];     // producer -> consumers: the slot's bytes have landed
  uint64_t empty[kPipeMaxSlots];    // consumers -> producer: slot reusable (16 warp arrivals)
  uint64_t counted[kPipeMeta];      // consumers -> scan warp: phase A done by all warps
  uint64_t ranked[kPipeMeta];       // scan warp -> consumers: prefix + warp offsets are valid
  SlotMeta meta[kPipeMeta];
  int end_iter;                     // iteration that carried the end marker (-1: not seen yet)
};

// One step after the filter applied to a survivor that sits in the side buffer (lane-major, not
// in its tile position): only scalar operands and compacted stores can run there — the host
// checks that (pipe_eligible) before it enables the side buffer.
template <typename T>
__device__ __forceinline__ void sparse_step(const KParams &p, uint32_t op, uint32_t a, uint32_t b, T &v, long long pos) {
  if (op == S_BIN) {
    const T sc = (p.in_kind[b] == NL_IN_SCALAR_IMM) ? from_bits<T>(p.imm[b]) : *reinterpret_cast<const T *>(p.in[b]);
    switch (a) {
      case V_MUL: v = op_mul<T>(v, sc); break;
      case V_ADD: v = op_add<T>(v, sc); break;
      case V_SUB: v = op_sub<T>(v, sc); break;
      case V_RSUB: v = op_sub<T>(sc, v); break;
      case V_DIV: v = op_div<T>(v, sc); break;
      case V_RDIV: v = op_div<T>(sc, v); break;
      case V_FDIV: v = op_floordiv<T>(v, sc); break;
      default: v = op_floordiv<T>(sc, v); break;
    }
  } else if (op == S_STORE) {
    reinterpret_cast<T *>(p.out[a])[pos] = v;
  }
}

// ---- program drivers split at the filter step -----------------------------------------------------
struct DynSplit {
  template <bool kValuesOnly, class TileT>
  static __device__ __forceinline__ void before(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
#pragma unroll 1
    for (int pc = 0; pc < p.filter_step; pc++) {
      const uint32_t s = p.step[pc];
      t.template exec<kValuesOnly>(p, sh, step_op(s), step_a(s), step_b(s), lane, warp);
    }
  }
  template <class TileT>
  static __device__ __forceinline__ void after(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
#pragma unroll 1
    for (int pc = p.filter_step + 1; pc < p.n_steps; pc++) {
      const uint32_t s = p.step[pc];
      t.exec(p, sh, step_op(s), step_a(s), step_b(s), lane, warp);
    }
  }
  template <typename T>
  static __device__ __forceinline__ void after_sparse(const KParams &p, T v, long long pos) {
#pragma unroll 1
    for (int pc = p.filter_step + 1; pc < p.n_steps; pc++) {
      const uint32_t s = p.step[pc];
      sparse_step<T>(p, step_op(s), step_a(s), step_b(s), v, pos);
    }
  }
  // store pass of the super-tile kernel: do the steps before the filter have to run again? (host: store_pass_needs_prefix)
  static __device__ __forceinline__ bool needs_prefix(const KParams &p) { return p.store_prefix != 0; }
  static constexpr bool stash_ok = false;        // (the interpreter keeps the two-pass form)
  static constexpr int stash_output = 0;
};

template <uint32_t... S>
struct StaticSplit {
  static constexpr int N = sizeof...(S);
  NL_HD static constexpr uint32_t step_at(int i) {
    constexpr uint32_t st[] = {S...};
    return st[i];
  }
  NL_HD static constexpr int filter_index() {
    constexpr uint32_t st[] = {S...};
    for (int i = 0; i < N; i++)
      if (step_op(st[i]) == S_FILTER) return i;
    return N;
  }
  template <bool kValuesOnly, class TileT, int I0, size_t... I>
  static __device__ __forceinline__ void run_seq(TileT &t, const KParams &p, TileShared &sh, int lane, int warp,
                                                 std::index_sequence<I...>) {
    (t.template exec<kValuesOnly>(p, sh, std::integral_constant<uint32_t, step_op(step_at(I0 + (int)I))>::value,
                                  std::integral_constant<uint32_t, step_a(step_at(I0 + (int)I))>::value,
                                  std::integral_constant<uint32_t, step_b(step_at(I0 + (int)I))>::value, lane, warp), ...);
  }
  template <typename T, int I0, size_t... I>
  static __device__ __forceinline__ void run_sparse(const KParams &p, T &v, long long pos, std::index_sequence<I...>) {
    (sparse_step<T>(p, std::integral_constant<uint32_t, step_op(step_at(I0 + (int)I))>::value,
                    std::integral_constant<uint32_t, step_a(step_at(I0 + (int)I))>::value,
                    std::integral_constant<uint32_t, step_b(step_at(I0 + (int)I))>::value, v, pos), ...);
  }
  template <bool kValuesOnly, class TileT>
  static __device__ __forceinline__ void before(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
    run_seq<kValuesOnly, TileT, 0>(t, p, sh, lane, warp, std::make_index_sequence<filter_index()>{});
  }
  template <class TileT>
  static __device__ __forceinline__ void after(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
    run_seq<false, TileT, filter_index() + 1>(t, p, sh, lane, warp, std::make_index_sequence<N - filter_index() - 1>{});
  }
  template <typename T>
  static __device__ __forceinline__ void after_sparse(const KParams &p, T v, long long pos) {
    run_sparse<T, filter_index() + 1>(p, v, pos, std::make_index_sequence<N - filter_index() - 1>{});
  }
  NL_HD static constexpr bool needs_prefix_c() {
    constexpr uint32_t st[] = {S...};
    return store_pass_needs_prefix(st, N, filter_index());
  }
  static __device__ __forceinline__ bool needs_prefix(const KParams &) { return needs_prefix_c(); }
  // Nothing but the compacted store follows the filter (`a[pred(a)]`): the survivors of a sparse warp-tile can wait
  // in shared memory between the passes instead of being re-read and re-ranked (compact_super_kernel: stash)
  NL_HD static constexpr bool stash_ok_c() {
    constexpr uint32_t st[] = {S...};
    return filter_index() + 2 == N && step_op(st[N - 1]) == S_STORE && step_b(st[N - 1]) == NL_IDX_COMPACT;
  }
  static constexpr bool stash_ok = stash_ok_c();
  NL_HD static constexpr int stash_output_c() {
    constexpr uint32_t st[] = {S...};
    return (int)step_a(st[N - 1]);
  }
  static constexpr int stash_output = stash_output_c();
};

// The same split for a shape family (FamilyProg): the step kinds on either side of the filter are
// build-time, operators and inputs come from the launch.
template <uint32_t... K>
struct FamilySplit {
  using Prog = FamilyProg<K...>;
  static constexpr int N = sizeof...(K);
  NL_HD static constexpr int filter_index() {
    constexpr uint32_t st[] = {K...};
    for (int i = 0; i < N; i++)
      if (step_op(st[i]) == S_FILTER) return i;
    return N;
  }
  template <bool kValuesOnly, class TileT, int I0, size_t... I>
  static __device__ __forceinline__ void run_seq(TileT &t, const KParams &p, TileShared &sh, int lane, int warp, std::index_sequence<I...>) {
    (Prog::template step<I0 + (int)I, kValuesOnly>(t, p, sh, lane, warp), ...);
  }
  template <bool kValuesOnly, class TileT>
  static __device__ __forceinline__ void before(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
    run_seq<kValuesOnly, TileT, 0>(t, p, sh, lane, warp, std::make_index_sequence<filter_index()>{});
  }
  template <class TileT>
  static __device__ __forceinline__ void after(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
    run_seq<false, TileT, filter_index() + 1>(t, p, sh, lane, warp, std::make_index_sequence<N - filter_index() - 1>{});
  }
  NL_HD static constexpr bool needs_prefix_c() {
    constexpr uint32_t st[] = {K...};
    return store_pass_needs_prefix(st, N, filter_index());
  }
  static __device__ __forceinline__ bool needs_prefix(const KParams &) { return needs_prefix_c(); }
  // Nothing but the compacted store follows the filter (`a[pred(a)]`): the survivors of a sparse warp-tile can wait
  // in shared memory between the passes instead of being re-read and re-ranked (compact_super_kernel: stash)
  // (a family knows which operands are scalars: operators on the survivors may follow the filter — `(a*2+1)[a > t]`)
  NL_HD static constexpr bool stash_ok_c() {
    constexpr uint32_t st[] = {K...};
    if (filter_index() + 2 > N || step_op(st[N - 1]) != S_STORE || step_b(st[N - 1]) != NL_IDX_COMPACT) return false;
    for (int i = filter_index() + 1; i < N - 1; i++)
      if (step_op(st[i]) != S_BIN || (step_b(st[i]) & SRC_TEMP) || step_b(st[i]) != FK_SCALAR) return false;
    return true;
  }
  template <typename T>
  static __device__ __forceinline__ void after_sparse(const KParams &p, T v, long long pos) {
#pragma unroll
    for (int i = filter_index() + 1; i < N; i++) {
      const uint32_t s = p.step[i];
      sparse_step<T>(p, step_op(s), step_a(s), step_b(s), v, pos);
    }
  }
  static constexpr bool stash_ok = stash_ok_c();
  NL_HD static constexpr int stash_output_c() {
    constexpr uint32_t st[] = {K...};
    return (int)step_a(st[N - 1]);
  }
  static constexpr int stash_output = stash_output_c();
};

// Optional per-role time breakdown (cycles), printed by CTA 0 when built with -DNL_PIPE_STATS.
#ifdef NL_PIPE_STATS
#define PSTAT_DECL(...) long long __VA_ARGS__
#define PSTAT_T0(v) const long long v = clock64()
#define PSTAT_ADD(acc, v) acc += clock64() - (v)
#define PSTAT_INC(acc, n) acc += (n)
#else
#define PSTAT_DECL(...)
#define PSTAT_T0(v)
#define PSTAT_ADD(acc, v)
#define PSTAT_INC(acc, n)
#endif

// ---- the kernel ---------------------------------------------------------------------------------
template <typename T, int V, class Split, int U = kPipeRows>
__global__ void __launch_bounds__(kPipeThreads, 1) compact_pipe_kernel(const __grid_constant__ KParams p) {
  constexpr int kPipeSideBytes = 96 * U;           // == pipe_side_bytes(U)
  using TileT = Tile<T, V, U, true, false, kPipeConsumers, true>;
  constexpr uint32_t kColBytes = (uint32_t)(TileT::TILE * sizeof(T));
  constexpr int kSideElems = kPipeSideBytes / (int)sizeof(T);      // side-buffer capacity of one warp, elements

  extern __shared__ __align__(128) unsigned char dyn_smem[];
  PipeShared &ps = *reinterpret_cast<PipeShared *>(dyn_smem);
  // [PipeShared | keep bits of every consumer thread, per lag entry | side buffers, per lag entry | ring of tile slots]
  const int S = p.n_slots, D = p.lag;
  const int LD = D + 1;              // lag entries: phase A of tile k fills one while phase B of tile k - D reads another
  using KeepT = typename std::conditional<(TileT::E <= 16), uint16_t, uint32_t>::type;   // one bit per element of a thread
  KeepT *keepbits = reinterpret_cast<KeepT *>(dyn_smem + ((sizeof(PipeShared) + 127) / 128) * 128);
  T *side = reinterpret_cast<T *>(keepbits + (size_t)LD * kPipeConsumers);
  unsigned char *ring = reinterpret_cast<unsigned char *>(side) + (p.side_cap ? (size_t)LD * kPipeConsumerWarps * kPipeSideBytes : 0);
  __shared__ TileShared sh_unused;   // exec() wants one; the staged phases never touch it

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t slot_bytes = kColBytes * (uint32_t)p.n_stage_cols;

  if (tid == 0) {
    ps.end_iter = -1;
    for (int s = 0; s < S; s++) {
      mbar_init(&ps.full[s], 1);
      mbar_init(&ps.empty[s], kPipeConsumerWarps);
    }
    for (int s = 0; s < kPipeMeta; s++) {
      mbar_init(&ps.counted[s], kPipeConsumerWarps);
      mbar_init(&ps.ranked[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (blockIdx.x == 0) {
      for (int o = 0; o < p.n_outputs; o++)
        if (p.out_space[o] == NL_IDX_LOOP) p.out_sizes[o] = p.end;
    }
  }
  __syncthreads();

  if (warp == kPipeConsumerWarps) {
    // ===================== producer: tickets + TMA bulk loads =====================
    if (lane == 0) {
      int s = 0;                        // slot / parity are carried, not divided out
      uint32_t par = 0;
      PSTAT_DECL(st_ticket = 0, st_empty = 0, st_n = 0);
      PSTAT_T0(st_begin);
      for (int k = 0;; k++, s = (s + 1 == S) ? 0 : s + 1, par ^= (s == 0) ? 1u : 0u) {
        PSTAT_T0(c0);
        const long long tile = (long long)atomicAdd(p.tile_ticket, 1u);
        PSTAT_ADD(st_ticket, c0 + (tile & 0));
        PSTAT_T0(c1);
        mbar_wait(&ps.empty[s], par ^ 1u);          // slot free (passes at once in the first round)
        PSTAT_ADD(st_empty, c1);
        PSTAT_INC(st_n, 1);
        SlotMeta &m = ps.meta[k & (kPipeMeta - 1)];
        m.tile = tile;
        m.agg = 0;
        if (tile >= p.n_tiles) {
          mbar_arrive(&ps.full[s]);                 // end marker travels through the pipeline
#ifdef NL_PIPE_STATS
          if (blockIdx.x == 0)
            printf("producer: tiles %lld total %lld cyc | per tile: ticket %lld, wait empty %lld\n", st_n, clock64() - st_begin,
                   st_ticket / st_n, st_empty / st_n);
#endif
          break;
        }
        const long long tile_base = p.start + tile * TileT::TILE;
        const bool full_tile = tile_base + TileT::TILE <= p.end;
        if (!full_tile) {
          mbar_arrive(&ps.full[s]);                 // the ragged last tile is read from global memory
          continue;
        }
        mbar_arrive_expect_tx(&ps.full[s], slot_bytes);
        unsigned char *dst = ring + (size_t)s * slot_bytes;
        for (int c = 0; c < p.n_inputs; c++) {
          if (p.stage_col[c] == 0xff) continue;
          const unsigned char *src = reinterpret_cast<const unsigned char *>(p.in[c]) + (size_t)tile_base * sizeof(T);
          bulk_g2s(dst + (size_t)p.stage_col[c] * kColBytes, src, kColBytes, &ps.full[s]);
        }
      }
    }
  } else if (warp > kPipeConsumerWarps) {
    // ===================== scan warps: warp offsets + decoupled look-back =====================
    // Scan warp sw takes iterations sw, sw + kPipeScanWarps, ...  It reaches iteration `it` only
    // after it - kPipeScanWarps was counted, and the consumers count in order, so the barrier it
    // waits on is never more than one phase behind (an mbarrier parity wait further ahead than
    // that would pass immediately).
    const int sw = warp - kPipeConsumerWarps - 1;
    PSTAT_DECL(st_counted = 0, st_look = 0, st_windows = 0, st_spins = 0, st_n = 0);
    for (int it = sw;; it += kPipeScanWarps) {
      PSTAT_T0(c0);
      const int me = it & (kPipeMeta - 1);
      const uint32_t par = (uint32_t)((it / kPipeMeta) & 1);
      // iterations after the end marker are never counted: leave instead of waiting for them
      bool stop = false;
      while (!mbar_try_wait_for(&ps.counted[me], par, kPipeWaitHintNs)) {
        const int e = *reinterpret_cast<volatile int *>(&ps.end_iter);
        if (e >= 0 && it > e) { stop = true; break; }
      }
      if (stop) break;
      PSTAT_ADD(st_counted, c0);
      PSTAT_T0(c1);
      SlotMeta &m = ps.meta[me];
      const long long tile = m.tile;
      if (tile >= p.n_tiles) break;
      // exclusive offsets of the consumer warps inside the tile (input order = warp order)
      const uint32_t val = (lane < kPipeConsumerWarps) ? m.wtot[lane] : 0u;
      uint32_t sc = val;
#pragma unroll
      for (int d = 1; d < kPipeConsumerWarps; d <<= 1) {
        const uint32_t nb = __shfl_up_sync(0xffffffffu, sc, d);
        if (lane >= d) sc += nb;
      }
      if (lane < kPipeConsumerWarps) m.off[lane] = sc - val;
      const unsigned long long agg = __shfl_sync(0xffffffffu, sc, kPipeConsumerWarps - 1);
      // look back over the global descriptors (this tile's aggregate was published by the last
      // consumer warp of phase A)
      unsigned long long excl = 0;
      if (tile > 0) {
        long long j = tile - 1;
        bool done = false;
        while (!done) {
          unsigned long long d[kPipeLookback];
#pragma unroll
          for (int k = 0; k < kPipeLookback; k++) {
            const long long q = j - lane - 32 * k;
            d[k] = (q >= 0) ? ld_relaxed_u64(&p.tile_desc[q * kDescStride]) : kDescPrefix;
          }
#pragma unroll
          for (int k = 0; k < kPipeLookback; k++) {
            if (done) break;
            const long long q = j - lane - 32 * k;
            while ((d[k] >> 62) == 0) { d[k] = ld_relaxed_u64(&p.tile_desc[q * kDescStride]); PSTAT_INC(st_spins, 1); }
            PSTAT_INC(st_windows, 1);
            const uint32_t has_prefix = __ballot_sync(0xffffffffu, (d[k] >> 62) == 2);
            const int first = has_prefix ? (__ffs(has_prefix) - 1) : 31;
            unsigned long long v = (lane <= first) ? (d[k] & kDescValueMask) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            excl += v;
            if (has_prefix) done = true;
          }
          j -= 32 * kPipeLookback;
        }
      }
      if (lane == 0) {
        st_relaxed_u64(&p.tile_desc[tile * kDescStride], kDescPrefix | ((excl + agg) & kDescValueMask));
        m.prefix = (long long)excl;
        if (tile == p.n_tiles - 1) {
          for (int o = 0; o < p.n_outputs; o++)
            if (p.out_space[o] == NL_IDX_COMPACT) p.out_sizes[o] = p.start + (long long)(excl + agg);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ps.ranked[me]);    // release: prefix and off[] are visible to phase B
      // multi-GPU: the shard's count goes to the peers from inside the kernel
      if (tile == p.n_tiles - 1 && p.peer_offsets != nullptr) peer_scan_counts(p, excl + agg, lane);
      PSTAT_ADD(st_look, c1);
      PSTAT_INC(st_n, 1);
    }
#ifdef NL_PIPE_STATS
    if (blockIdx.x == 0 && lane == 0 && st_n)
      printf("scan warp %d: tiles %lld | per tile: wait counted %lld, offsets+look-back %lld cyc, windows %.2f, lane-0 respins %.2f\n", sw,
             st_n, st_counted / st_n, st_look / st_n, (double)st_windows / st_n, (double)st_spins / st_n);
#endif
  } else {
    // ===================== consumer warps =====================
    // Per iteration k: phase A of tile k (count), then phase B of tile k - D (store).  A warp whose
    // rows kept at most side_cap elements copies them — already ranked — into its side buffer
    // and hands the data slot back in phase A, so at low selectivity the whole ring is prefetch;
    // otherwise the slot stays parked until phase B recomputes the values from it.
    TileT t;
    t.rsum = 0;
    t.rext = T(0);
    const uint32_t side_cap = (uint32_t)p.side_cap;
    int a_end = -1;                                   // iteration at which the end marker was seen
    int sa = 0, sb = 0;                               // data slots of phase A / phase B: carried (no divisions)
    int la = 0, lb = 0;                               // lag entries of phase A / phase B
    uint32_t para = 0;
    PSTAT_DECL(st_full = 0, st_a = 0, st_ranked = 0, st_b = 0, st_n = 0);
    PSTAT_T0(st_begin);
    for (int k = 0;; k++) {
      // ---- phase A of iteration k: count (first: the tile's aggregate must go public early) ----
      if (a_end < 0) {
        const int s = sa;
        const uint32_t par = para;
        sa = (sa + 1 == S) ? 0 : sa + 1;
        para ^= (sa == 0) ? 1u : 0u;
        KeepT *my_keepbits = keepbits + la * kPipeConsumers + tid;
        T *my_side = side + (size_t)(la * kPipeConsumerWarps + warp) * kSideElems;
        la = (la + 1 == LD) ? 0 : la + 1;
        const int me = k & (kPipeMeta - 1);
        PSTAT_T0(c0);
        mbar_wait(&ps.full[s], par);
        PSTAT_ADD(st_full, c0);
        PSTAT_T0(c1);
        SlotMeta &m = ps.meta[me];
        const long long tile = m.tile;
        if (tile >= p.n_tiles) {
          a_end = k;
          if (warp == 0 && lane == 0) *reinterpret_cast<volatile int *>(&ps.end_iter) = k;
          if (lane == 0) mbar_arrive(&ps.counted[me]);
        } else {
          t.begin_tile(p, tile, tid);
          t.stage = ring + (size_t)s * slot_bytes;
          Split::template before<false>(t, p, sh_unused, lane, warp);
          t.keep &= t.pacc;
          const uint32_t cnt = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(t.keep));
          const bool release_now = cnt <= side_cap;     // warp-uniform; phase B takes the same branch
          if (lane == 0) {
            m.wtot[warp] = cnt;
            // one shared atomic carries both the warp's count (low bits) and its arrival (high
            // bits); the warp that sees 15 earlier arrivals publishes the tile aggregate now,
            // not after the look-back
            const unsigned int old = atomicAdd(&m.agg, cnt | (1u << kPipeArriveShift));
            if ((old >> kPipeArriveShift) == kPipeConsumerWarps - 1) {
              const unsigned long long agg = (unsigned long long)((old & ((1u << kPipeArriveShift) - 1u)) + cnt);
              if (tile > 0) st_relaxed_u64(&p.tile_desc[tile * kDescStride], kDescAggregate | agg);
              else st_relaxed_u64(&p.tile_desc[0], kDescPrefix | agg);
            }
            mbar_arrive(&ps.counted[me]);
          }
          if (release_now) {
            if (cnt != 0) {
              t.rank_rows(lane);
              if (t.keep) {
#pragma unroll
                for (int u = 0; u < U; u++) {
                  uint32_t ci = t.crow[u];
#pragma unroll
                  for (int v = 0; v < V; v++)
                    if ((t.keep >> (u * V + v)) & 1u) my_side[ci++] = t.acc[u * V + v];
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ps.empty[s]);
          } else {
            *my_keepbits = (KeepT)t.keep;     // parked with the tile for phase B
          }
        }
        PSTAT_ADD(st_a, c1);
        PSTAT_INC(st_n, 1);
      }
      // ---- phase B of iteration k - D: store ----
      const int kb = k - D;
      if (kb >= 0) {
        if (a_end >= 0 && kb >= a_end) break;
        PSTAT_T0(c2);
        const int s = sb;
        sb = (sb + 1 == S) ? 0 : sb + 1;
        const KeepT *my_keepbits = keepbits + lb * kPipeConsumers + tid;
        const T *my_side = side + (size_t)(lb * kPipeConsumerWarps + warp) * kSideElems;
        lb = (lb + 1 == LD) ? 0 : lb + 1;
        const int me = kb & (kPipeMeta - 1);
        SlotMeta &m = ps.meta[me];
        mbar_wait(&ps.ranked[me], (uint32_t)((kb / kPipeMeta) & 1));
        PSTAT_ADD(st_ranked, c2);
        PSTAT_T0(c3);
        const uint32_t cnt = m.wtot[warp];              // warp-uniform
        if (cnt != 0) {
          const long long cbase = p.start + m.prefix + (long long)m.off[warp];
          if (cnt <= side_cap) {
            // survivors wait in the side buffer in output order: coalesced stores
#pragma unroll
            for (int r = 0; r < (kSideElems + 31) / 32; r++) {
              const uint32_t i = (uint32_t)lane + 32u * r;
              if (i < cnt) Split::template after_sparse<T>(p, my_side[i], cbase + i);
            }
          } else {
            const uint32_t my_keep = *my_keepbits;
            t.begin_tile(p, m.tile, tid);
            t.stage = ring + (size_t)s * slot_bytes;
            // values are recomputed from shared memory only by the threads that kept something
            // (the steps before the filter are per-thread: no warp collectives inside)
            if (my_keep) Split::template before<true>(t, p, sh_unused, lane, warp);
            t.keep = my_keep;
            t.rank_rows(lane);
            t.cbase = cbase;
            if (my_keep) Split::after(t, p, sh_unused, lane, warp);
            __syncwarp();
            if (lane == 0) mbar_arrive(&ps.empty[s]);
          }
        }
        PSTAT_ADD(st_b, c3);
      }
    }
#ifdef NL_PIPE_STATS
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kPipeConsumerWarps - 1) && st_n)
      printf("consumer warp %d: tiles %lld total %lld cyc | per tile: wait full %lld, phase A %lld, wait ranked %lld, phase B %lld\n", warp,
             st_n, clock64() - st_begin, st_full / st_n, st_a / st_n, st_ranked / st_n, st_b / st_n);
#endif
  }
}

typedef void (*pipe_fn)(const KParams);

}  // namespace nl
