// nl_compact_super.cuh — compaction over super-tiles (sm_100a).
//
// Measured on B200: a compaction kernel's rate is
//       resident CTAs per SM  x  bytes a CTA moves per tile  /  latency of one tile,
// and the latency (ticket, loads, look-back, stores: 4-6 us) hardly depends on the bytes.  The
// plain kernel keeps a tile's values in registers across the look-back, so its tile is bounded by
// the register file (32 KB) and it stops at 0.7-0.8 of the roofline; `b[a > t]`, which moves two
// columns per tile with the same registers, runs at 0.97-1.0.
//
// This kernel gives every filter that ratio.  A CTA claims a SUPER-tile of K consecutive tiles:
//   count pass   for each of the K tiles: run the program up to the filter, keep ONLY the keep
//                mask (one bit per element) and the warp's survivor count; rows are loaded with
//                normal L2 residency;
//   one look-back for the whole super-tile (K x fewer descriptors, K x fewer latencies);
//   store pass   for each tile: re-run the value part of the program — its rows come back from L2,
//                not from HBM (K x 32 KB per CTA, a few tens of MB chip-wide, far below the 126 MB
//                L2) — rank, run the steps after the filter, store compacted.
// HBM traffic stays one read of the inputs and one write of the survivors; no shared-memory ring, no
// warp specialisation, and the same code serves build-time shapes and the step interpreter.
#pragma once
#include "nl_compact_pipe.cuh"   // StaticSplit / DynSplit: the program split at the filter step

namespace nl {

#ifndef NL_SUPER_K
#define NL_SUPER_K 4
#endif
constexpr int kSuperK = NL_SUPER_K;        // tiles per super-tile

template <typename T, int V, int U, class Split, bool kTemps, int kMinCtas>
__global__ void __launch_bounds__(kBlock, kMinCtas) compact_super_kernel(const __grid_constant__ KParams p) {
  using TileT = Tile<T, V, U, true, false, kBlock, false, kTemps>;
  constexpr int NW = kBlock / 32;
  constexpr int K = kSuperK;
  static_assert(K * NW <= 32, "one warp scans all (tile, warp) counts of a super-tile");
  __shared__ TileShared sh;
  __shared__ uint32_t wtot[K][NW];     // survivors of warp w in tile k
  __shared__ uint32_t woff[K][NW];     // exclusive scan over (k, w): offset of the warp's run inside the super-tile
  // The stash: when only scalar operators and the compacted store follow the filter and a warp-tile kept at most a
  // quarter of its elements, the count pass ranks the survivors and parks them here, in output order; the store
  // pass is then one coalesced copy — no second read of the rows from L2, no second ranking, no scattered
  // predicated stores.  (ncu, a[a > t] at 10 %: the store pass was 55 % of the instructions and all of the second
  // L2 read; this is the kernel that runs into the 1000 W power cap in long runs, so instructions saved are also
  // clocks kept: 10 % selectivity 0.82 -> 0.96 of the roofline inside the bench run, 1 % 0.89 -> 0.99.  Dense
  // tiles take the two-pass route as before; a second, stash-free form of the count loop for CTAs that meet
  // dense data was measured and dropped: no gain at 50-90 %, 3-9 % loss at 1-10 %.)
  constexpr bool kStash = Split::stash_ok;
  constexpr int kStashCap = kStash ? (32 * TileT::E) / 4 : 1;
  __shared__ T stash[kStash ? K : 1][kStash ? NW : 1][kStashCap];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long n_sub = (p.end - p.start + TileT::TILE - 1) / TileT::TILE;

  // which CTA of its SM is this?  (the kernel is persistent: all CTAs are resident from the start)
  if (tid == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    sh.last = (int)atomicAdd(&p.sm_rank[smid & 255u], 1u);
  }
  __syncthreads();
  const bool may_retire = sh.last >= p.super_keep_ctas;
  __syncthreads();

  if (blockIdx.x == 0 && tid == 0) {
    for (int o = 0; o < p.n_outputs; o++)
      if (p.out_space[o] == NL_IDX_LOOP) p.out_sizes[o] = p.end;
  }

  // Two arenas (ticket, per-SM ranks, descriptors) alternate between consecutive launches on a stream: this launch
  // works in one and, while it starts, zeroes what the previous launch left in the other — all CTAs share the work, no
  // memset launch stands in front of a filter (3-4 us of a 0.2 ms kernel when 8 GPUs share a 2^30-element filter), and
  // nothing happens at the kernel's end.  (A first form had the LAST CTA to finish zero the launch's own arena: 65,536
  // descriptors stored by one CTA after everybody else had left, and the extra code at the loop's exit, cost the 2^30
  // filters 3-4 %.)  The host clears an arena itself after a compaction kernel that does not take part, while a
  // launch is captured into a graph, and after a (re)allocation: nl_runtime.cu.
  if (p.clean_count > 0) {
    for (long long i = (long long)blockIdx.x * kBlock + tid; i < p.clean_count; i += (long long)gridDim.x * kBlock)
      p.clean_desc[i * kDescStride] = 0ull;
    if (blockIdx.x == 0) {
      p.clean_rank[tid & 255] = 0u;
      if (tid == 0) *p.clean_ticket = 0u;
    }
  }

  TileT t;
  t.rsum = 0;
  t.rext = T(0);
  while (true) {
    if (tid == 0) sh.tile = (long long)atomicAdd(p.tile_ticket, 1u);
    __syncthreads();
    const long long super = sh.tile;
    if (super >= p.n_tiles) break;

    // ---- count pass ----
    // (Requesting the rows of tiles 1..K-1 into L2 with prefetch.global.L2 before tile 0 is read — the tiles are
    // counted one after the other through the same registers — was measured and dropped: 3-9 % SLOWER at every
    // selectivity and size, round 2.)
    uint32_t keepm[K];
    t.keep_in_l2 = true;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const long long sub = super * K + k;
      keepm[k] = 0;
      uint32_t cnt = 0;
      if (sub < n_sub) {
        t.begin_tile(p, sub, tid);
        Split::template before<false>(t, p, sh, lane, warp);
        keepm[k] = t.keep & t.pacc;
        cnt = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(keepm[k]));
        if constexpr (kStash) {
          if (cnt != 0 && cnt <= (uint32_t)kStashCap) {          // warp-uniform
            t.keep = keepm[k];
            t.rank_rows(lane);
            T *slot = stash[k][warp];
#pragma unroll
            for (int u = 0; u < U; u++) {
              uint32_t ci = t.crow[u];
#pragma unroll
              for (int v = 0; v < V; v++)
                if ((keepm[k] >> (u * V + v)) & 1u) slot[ci++] = t.acc[u * V + v];
            }
          }
        }
      }
      if (lane == 0) wtot[k][warp] = cnt;
    }
    __syncthreads();

    // ---- offsets of the (tile, warp) runs + one decoupled look-back for the super-tile ----
    if (warp == 0) {
      const uint32_t val = (lane < K * NW) ? (&wtot[0][0])[lane] : 0u;
      uint32_t sc = val;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t nb = __shfl_up_sync(0xffffffffu, sc, d);
        if (lane >= d) sc += nb;
      }
      if (lane < K * NW) (&woff[0][0])[lane] = sc - val;
      const unsigned long long agg = __shfl_sync(0xffffffffu, sc, 31);
      unsigned long long excl = 0;
      if (super == 0) {
        if (lane == 0) st_relaxed_u64(&p.tile_desc[0], kDescPrefix | agg);
      } else {
        if (lane == 0) st_relaxed_u64(&p.tile_desc[super * kDescStride], kDescAggregate | agg);
        long long j = super - 1;
        bool done = false;
        while (!done) {
          const long long q = j - lane;
          unsigned long long d = (q >= 0) ? ld_relaxed_u64(&p.tile_desc[q * kDescStride]) : kDescPrefix;
          while ((d >> 62) == 0) {
            // the predecessor has not published yet: polling flat out costs issue slots and L2 bandwidth — and power,
            // which is what throttles this kernel in long runs
            if (p.poll_sleep_ns) __nanosleep(p.poll_sleep_ns);
            d = ld_relaxed_u64(&p.tile_desc[q * kDescStride]);
          }
          const uint32_t has_prefix = __ballot_sync(0xffffffffu, (d >> 62) == 2);
          const int first = has_prefix ? (__ffs(has_prefix) - 1) : 31;
          unsigned long long v = (lane <= first) ? (d & kDescValueMask) : 0ull;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          excl += v;
          if (has_prefix) done = true;
          j -= 32;
        }
        if (lane == 0) st_relaxed_u64(&p.tile_desc[super * kDescStride], kDescPrefix | ((excl + agg) & kDescValueMask));
      }
      if (lane == 0) {
        sh.prefix = (long long)excl;
        // dense super-tile (more than 70 % kept)?  see the end of the loop
        sh.last = (agg * 10ull > (unsigned long long)(K * TileT::TILE) * 7ull) ? 1 : 0;
        if (super == p.n_tiles - 1) {
          for (int o = 0; o < p.n_outputs; o++)
            if (p.out_space[o] == NL_IDX_COMPACT) p.out_sizes[o] = p.start + (long long)(excl + agg);
        }
      }
      // multi-GPU: the shard's count goes to the peers from inside the kernel
      if (super == p.n_tiles - 1 && p.peer_offsets != nullptr) peer_scan_counts(p, excl + agg, lane);
    }
    __syncthreads();

    // ---- store pass: values come back from L2 ----
    // (Letting warps 1..7 fetch, recompute and rank their first two-pass tile while warp 0 walks the look-back chain —
    // 14 % of the stall samples sit at that barrier — was measured and dropped: 2-3 % slower on dense input, and the
    // reordered store pass alone cost the family kernels up to 10 %, round 2.)
    t.keep_in_l2 = false;
    const long long base_out = p.start + sh.prefix;
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (wtot[k][warp] == 0) continue;             // warp-uniform
      if constexpr (kStash) {
        const uint32_t cnt = wtot[k][warp];
        if (cnt <= (uint32_t)kStashCap) {
          // (whatever follows the filter — scalar operators, the store — runs on the bare survivors)
          const T *slot = stash[k][warp];
          const long long at = base_out + (long long)woff[k][warp];
          for (uint32_t i = lane; i < cnt; i += 32) Split::template after_sparse<T>(p, slot[i], at + (long long)i);
          continue;
        }
      }
      const long long sub = super * K + k;
      t.begin_tile(p, sub, tid);
      // (only when a value from before the filter is still needed: `(a*2+1)[a > t]` reloads its column after it)
      if (keepm[k] && Split::needs_prefix(p)) Split::template before<true>(t, p, sh, lane, warp);
      t.keep = keepm[k];
      t.rank_rows(lane);
      t.cbase = base_out + (long long)woff[k][warp];
      if (keepm[k]) Split::after(t, p, sh, lane, warp);
    }
    const bool dense = sh.last != 0;
    __syncthreads();                                // sh.tile / wtot are rewritten by the next round
    // The rows of K tiles per resident CTA wait in L2 between the two passes.  When most elements
    // survive, the survivors written meanwhile push that window (4 CTAs/SM x 128 KB = 76 MB) out of
    // L2 and the store pass re-reads from HBM (measured +36 % DRAM reads at 90 %).  The CTA that
    // started fourth on its SM then retires: three per SM keep the window at 57 MB (0.85 -> 0.94 of
    // the roofline at 90 %), while sparse inputs keep all four (0.90 vs 0.80 at 1 %).
    if (dense && may_retire) break;
  }
}

}  // namespace nl
