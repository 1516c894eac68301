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
#include <type_traits>
#include <utility>

#include "nl_pipeline.cuh"

namespace nl {

constexpr int kPipeConsumerWarps = 16;
constexpr int kPipeScanWarps = 4;
constexpr int kPipeConsumers = kPipeConsumerWarps * 32;                       // 512
constexpr int kPipeThreads = kPipeConsumers + 32 + kPipeScanWarps * 32;      // + producer + scan
constexpr int kPipeMaxSlots = 16;

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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// for the roles that are not on the critical path (producer, scan warps): back off between polls
// so the spin does not take issue slots from the consumer warps
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(100);
}
// global -> shared bulk copy; completion is signalled on `bar` as transferred bytes
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// per-slot bookkeeping in shared memory
struct SlotMeta {
  long long tile;                       // tile id (>= n_tiles: end of work)
  long long prefix;                     // elements kept before this tile (written by the scan warp)
  unsigned int agg;                     // elements kept in this tile (atomic sum of the warps)
  unsigned int arrived;                 // consumer warps done with phase A
  unsigned int wtot[kPipeConsumerWarps];   // per consumer warp: elements kept
  unsigned int off[kPipeConsumerWarps];    // exclusive scan of wtot (scan warp)
};

struct PipeShared {
  uint64_t full[kPipeMaxSlots];     // producer -> consumers: the slot's bytes have landed
  uint64_t counted[kPipeMaxSlots];  // consumers -> scan warp: phase A done by all warps
  uint64_t ranked[kPipeMaxSlots];   // scan warp -> consumers: prefix + warp offsets are valid
  uint64_t empty[kPipeMaxSlots];    // consumers -> producer: phase B done, slot reusable
  SlotMeta meta[kPipeMaxSlots];
  int end_iter;                     // iteration that carried the end marker (-1: not seen yet)
};

// ---- program drivers split at the filter step -----------------------------------------------------
struct DynSplit {
  template <class TileT>
  static __device__ __forceinline__ void before(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
#pragma unroll 1
    for (int pc = 0; pc < p.filter_step; pc++) {
      const uint32_t s = p.step[pc];
      t.exec(p, sh, step_op(s), step_a(s), step_b(s), lane, warp);
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
  template <class TileT, int I0, size_t... I>
  static __device__ __forceinline__ void run_seq(TileT &t, const KParams &p, TileShared &sh, int lane, int warp,
                                                 std::index_sequence<I...>) {
    (t.exec(p, sh, std::integral_constant<uint32_t, step_op(step_at(I0 + (int)I))>::value,
            std::integral_constant<uint32_t, step_a(step_at(I0 + (int)I))>::value,
            std::integral_constant<uint32_t, step_b(step_at(I0 + (int)I))>::value, lane, warp), ...);
  }
  template <class TileT>
  static __device__ __forceinline__ void before(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
    run_seq<TileT, 0>(t, p, sh, lane, warp, std::make_index_sequence<filter_index()>{});
  }
  template <class TileT>
  static __device__ __forceinline__ void after(TileT &t, const KParams &p, TileShared &sh, int lane, int warp) {
    run_seq<TileT, filter_index() + 1>(t, p, sh, lane, warp, std::make_index_sequence<N - filter_index() - 1>{});
  }
};

// ---- the kernel ---------------------------------------------------------------------------------
template <typename T, int V, class Split>
__global__ void __launch_bounds__(kPipeThreads, 1) compact_pipe_kernel(const __grid_constant__ KParams p) {
  constexpr int U = kRows;
  using TileT = Tile<T, V, U, true, false, kPipeConsumers, true>;
  constexpr uint32_t kColBytes = (uint32_t)(TileT::TILE * sizeof(T));

  extern __shared__ __align__(128) unsigned char dyn_smem[];
  PipeShared &ps = *reinterpret_cast<PipeShared *>(dyn_smem);
  // [PipeShared | keep bits of every consumer thread, per slot | ring of tile slots]
  uint32_t *keepbits = reinterpret_cast<uint32_t *>(dyn_smem + ((sizeof(PipeShared) + 127) / 128) * 128);
  unsigned char *ring = reinterpret_cast<unsigned char *>(keepbits) + (size_t)p.n_slots * kPipeConsumers * sizeof(uint32_t);
  __shared__ TileShared sh_unused;   // exec() wants one; the staged phases never touch it

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = p.n_slots, D = p.lag;
  const uint32_t slot_bytes = kColBytes * (uint32_t)p.n_stage_cols;

  if (tid == 0) {
    ps.end_iter = -1;
    for (int s = 0; s < S; s++) {
      mbar_init(&ps.full[s], 1);
      mbar_init(&ps.counted[s], kPipeConsumerWarps);
      mbar_init(&ps.ranked[s], 1);
      mbar_init(&ps.empty[s], kPipeConsumerWarps);
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
      const int B = p.batch;
      long long batch_base = 0;
      int s = 0, in_batch = 0;          // slot / parity / position in the batch are carried, not divided out
      uint32_t par = 0;
      for (;; s = (s + 1 == S) ? 0 : s + 1, par ^= (s == 0) ? 1u : 0u, in_batch = (in_batch + 1 == B) ? 0 : in_batch + 1) {
        // one atomic claims a batch of B consecutive tiles: the ticket round trip is paid once
        // per batch and consecutive tiles let the scan warp chain their prefixes locally
        if (in_batch == 0) batch_base = (long long)atomicAdd(p.tile_ticket, (unsigned int)B);
        mbar_wait_relaxed(&ps.empty[s], par ^ 1u);  // slot free (passes at once in the first round)
        const long long tile = batch_base + in_batch;
        SlotMeta &m = ps.meta[s];
        m.tile = tile;
        m.agg = 0;
        m.arrived = 0;
        if (tile >= p.n_tiles) {
          mbar_arrive(&ps.full[s]);                 // end marker travels through the pipeline
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
        for (int k = 0; k < p.n_inputs; k++) {
          if (p.stage_col[k] == 0xff) continue;
          const unsigned char *src = reinterpret_cast<const unsigned char *>(p.in[k]) + (size_t)tile_base * sizeof(T);
          bulk_g2s(dst + (size_t)p.stage_col[k] * kColBytes, src, kColBytes, &ps.full[s]);
        }
      }
    }
  } else if (warp > kPipeConsumerWarps) {
    // ===================== scan warps: warp offsets + decoupled look-back =====================
    // A scan warp strides over the iterations; its stride must not exceed the ring depth, or its
    // first wait on a slot would target a phase the barrier has not even started (an mbarrier
    // parity wait more than one phase ahead passes immediately).  Surplus scan warps stay idle.
    const int sw = warp - kPipeConsumerWarps - 1;
    const int B = p.batch;
    int n_scan = S / B;                       // stride in iterations = n_scan * B <= S
    if (n_scan > kPipeScanWarps) n_scan = kPipeScanWarps;
    if (n_scan < 1) n_scan = 1;
    unsigned long long chain_excl = 0, chain_agg = 0;   // prefix state carried along a batch
    for (int it = sw * B; sw < n_scan; it = ((it + 1) % B == 0) ? (it + 1 + (n_scan - 1) * B) : (it + 1)) {
      const int s = it % S;
      const uint32_t par = (uint32_t)((it / S) & 1);
      // iterations after the end marker are never counted: leave instead of waiting for them
      bool stop = false;
      while (!mbar_try_wait(&ps.counted[s], par)) {
        const int e = *reinterpret_cast<volatile int *>(&ps.end_iter);
        if (e >= 0 && it > e) { stop = true; break; }
        __nanosleep(100);
      }
      if (stop) break;
      SlotMeta &m = ps.meta[s];
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
      // the first tile of a batch looks back over the global descriptors (its aggregate was
      // published by the last consumer warp of phase A); the following tiles of the batch are
      // consecutive, so their prefix is the previous one plus the previous aggregate
      unsigned long long excl = 0;
      if (it % B != 0) {
        excl = chain_excl + chain_agg;
      } else if (tile > 0) {
        long long j = tile - 1;
        bool done = false;
        while (!done) {
          unsigned long long d[TileT::LOOKBACK];
#pragma unroll
          for (int k = 0; k < TileT::LOOKBACK; k++) {
            const long long q = j - lane - 32 * k;
            d[k] = (q >= 0) ? ld_relaxed_u64(&p.tile_desc[q * kDescStride]) : kDescPrefix;
          }
#pragma unroll
          for (int k = 0; k < TileT::LOOKBACK; k++) {
            if (done) break;
            const long long q = j - lane - 32 * k;
            while ((d[k] >> 62) == 0) d[k] = ld_relaxed_u64(&p.tile_desc[q * kDescStride]);
            const uint32_t has_prefix = __ballot_sync(0xffffffffu, (d[k] >> 62) == 2);
            const int first = has_prefix ? (__ffs(has_prefix) - 1) : 31;
            unsigned long long v = (lane <= first) ? (d[k] & kDescValueMask) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            excl += v;
            if (has_prefix) done = true;
          }
          j -= 32 * TileT::LOOKBACK;
        }
      }
      chain_excl = excl;
      chain_agg = agg;
      if (lane == 0) {
        st_relaxed_u64(&p.tile_desc[tile * kDescStride], kDescPrefix | ((excl + agg) & kDescValueMask));
        m.prefix = (long long)excl;
        if (tile == p.n_tiles - 1) {
          for (int o = 0; o < p.n_outputs; o++)
            if (p.out_space[o] == NL_IDX_COMPACT) p.out_sizes[o] = p.start + (long long)(excl + agg);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ps.ranked[s]);     // release: prefix and off[] are visible to phase B
    }
  } else {
    // ===================== consumer warps =====================
    TileT t;
    t.rsum = 0;
    t.rext = T(0);
    int a_end = -1;                                   // iteration at which the end marker was seen
    int sa = 0, sb = 0;                               // slots of phase A / phase B, carried (no divisions)
    uint32_t para = 0, parb = 0;
    for (int k = 0;; k++) {
      // ---- phase A of iteration k: count ----
      if (a_end < 0) {
        const int s = sa;
        const uint32_t par = para;
        sa = (sa + 1 == S) ? 0 : sa + 1;
        para ^= (sa == 0) ? 1u : 0u;
        mbar_wait(&ps.full[s], par);
        SlotMeta &m = ps.meta[s];
        const long long tile = m.tile;
        if (tile >= p.n_tiles) {
          a_end = k;
          if (warp == 0 && lane == 0) *reinterpret_cast<volatile int *>(&ps.end_iter) = k;
          if (lane == 0) mbar_arrive(&ps.counted[s]);
        } else {
          t.begin_tile(p, tile, tid);
          t.stage = ring + (size_t)s * slot_bytes;
          Split::before(t, p, sh_unused, lane, warp);
          t.keep &= t.pacc;
          // survivors of this warp's rows (one REDUX); the keep bits are parked with the tile so
          // that phase B needs the data only where something survived
          keepbits[s * kPipeConsumers + tid] = t.keep;
          const uint32_t cnt = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(t.keep));
          if (lane == 0) {
            m.wtot[warp] = cnt;
            atomicAdd(&m.agg, cnt);
            __threadfence_block();
            const unsigned int n = atomicAdd(&m.arrived, 1u);
            if (n == kPipeConsumerWarps - 1) {
              // last warp of the tile: its aggregate goes public now, not after the look-back
              const unsigned long long agg = atomicAdd(&m.agg, 0u);
              if (tile > 0) st_relaxed_u64(&p.tile_desc[tile * kDescStride], kDescAggregate | agg);
              else st_relaxed_u64(&p.tile_desc[0], kDescPrefix | agg);
            }
            mbar_arrive(&ps.counted[s]);
          }
        }
      }
      // ---- phase B of iteration k - D: rank and store ----
      const int kb = k - D;
      if (kb >= 0) {
        if (a_end >= 0 && kb >= a_end) break;
        const int s = sb;
        const uint32_t par = parb;
        sb = (sb + 1 == S) ? 0 : sb + 1;
        parb ^= (sb == 0) ? 1u : 0u;
        mbar_wait(&ps.ranked[s], par);
        SlotMeta &m = ps.meta[s];
        const long long tile = m.tile;
        const uint32_t my_keep = keepbits[s * kPipeConsumers + tid];
        const bool warp_has = m.wtot[warp] != 0;        // warp-uniform: nothing to rank or store otherwise
        if (warp_has) {
        t.begin_tile(p, tile, tid);
        t.stage = ring + (size_t)s * slot_bytes;
        // values are recomputed from shared memory only by the threads that kept something
        // (the steps before the filter are per-thread: no warp collectives inside)
        if (my_keep) Split::before(t, p, sh_unused, lane, warp);
        t.keep = my_keep;
        // rank of every kept element in (row, lane, v) order inside the warp's run
        uint32_t packed[TileT::NWORDS], incl[TileT::NWORDS];
#pragma unroll
        for (int w = 0; w < TileT::NWORDS; w++) packed[w] = 0;
#pragma unroll
        for (int u = 0; u < U; u++)
          packed[u / 4] |= (uint32_t)__popc((t.keep >> (u * V)) & TileT::ROWMASK) << (8 * (u % 4));
#pragma unroll
        for (int w = 0; w < TileT::NWORDS; w++) {
          incl[w] = packed[w];
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t nb = __shfl_up_sync(0xffffffffu, incl[w], d);
            if (lane >= d) incl[w] += nb;
          }
        }
        uint32_t run = 0;
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint32_t tot = (__shfl_sync(0xffffffffu, incl[u / 4], 31) >> (8 * (u % 4))) & 0xffu;
          const uint32_t ex = ((incl[u / 4] - packed[u / 4]) >> (8 * (u % 4))) & 0xffu;
          t.crow[u] = run + ex;
          run += tot;
        }
        t.cbase = p.start + m.prefix + (long long)m.off[warp];
        if (my_keep) Split::after(t, p, sh_unused, lane, warp);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ps.empty[s]);
      }
    }
  }
}

typedef void (*pipe_fn)(const KParams);

}  // namespace nl
