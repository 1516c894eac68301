// device_sort.cpp — see device_sort.hpp.
//
// One GPU: copy the shard, nl_sort() it (LSD radix sort, one sweep per digit position, nl_sort.cu).
// Several GPUs: sample sort over NVLink, every key radix-sorted ONCE —
//   1. one small kernel per GPU reads regularly spaced samples of its (unsorted) shard; the host sorts them and picks
//      G - 1 splitters (in the kernels' key order, so NaNs stay last);
//   2. every GPU counts how many of its keys fall between the splitters (nl_sort_split_count), all GPUs side by side;
//   3. every GPU cuts its shard with ONE sweep of the sort's own kernel — "number of splitters below the key" as the
//      digit — that stores run j straight into its place in GPU j's bucket (nl_sort_split_scatter): the exchange
//      happens inside the kernel, over NVLink peer memory, while it regroups;
//   4. every GPU sorts what it received — bucket j holds exactly the elements between splitters j - 1 and j, so the
//      buckets in order are the sorted array;
//   5. the buckets are re-cut into the engine's regular partition (peer copies again), so the result can meet any
//      other column of the same length in a fused pipeline.
// (Round 1 sorted every shard first, found the cuts with ~1,500 single-element reads + synchronisations, and sorted
// every bucket again.)
#include "device_sort.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Buf {
  int dev = -1;
  void *ptr = nullptr;
  int64_t count = 0;
};

void free_bufs(std::vector<Buf> &v) {
  for (Buf &b : v)
    if (b.ptr) { nl_free(b.dev, b.ptr); b.ptr = nullptr; }
}

int sync_devices(int G) {
  int rc = NL_OK;
  Py_BEGIN_ALLOW_THREADS
  for (int g = 0; g < G && rc == NL_OK; g++) rc = nl_stream_sync(g, 0);
  Py_END_ALLOW_THREADS
  return rc;
}

// copy the global element range [lo, hi) of the concatenation of `src` into dst (on dst_dev)
int gather_range(const std::vector<Buf> &src, const std::vector<int64_t> &src_start, size_t es, int64_t lo, int64_t hi,
                 int dst_dev, void *dst) {
  for (size_t j = 0; j < src.size(); j++) {
    const int64_t a = std::max(lo, src_start[j]), b = std::min(hi, src_start[j] + src[j].count);
    if (a >= b) continue;
    const char *from = (const char *)src[j].ptr + (size_t)(a - src_start[j]) * es;
    char *to = (char *)dst + (size_t)(a - lo) * es;
    const int rc = (src[j].dev == dst_dev) ? nl_memcpy_d2d(dst_dev, 0, to, from, (size_t)(b - a) * es)
                                           : nl_memcpy_peer(dst_dev, 0, to, src[j].dev, from, (size_t)(b - a) * es);
    if (rc != NL_OK) return rc;
  }
  return NL_OK;
}

// regular storage of `total` elements filled from the concatenation of `runs`
Storage *regular_from_runs(int dtype, const std::vector<Buf> &runs, int64_t total) {
  const size_t es = nl_dtype_size(dtype);
  Storage *out = storage_new(dtype, total);
  if (storage_alloc_shards(out) < 0) { storage_decref(out); return nullptr; }
  std::vector<int64_t> start(runs.size(), 0);
  int64_t run = 0;
  for (size_t j = 0; j < runs.size(); j++) { start[j] = run; run += runs[j].count; }
  for (Shard &sh : out->shards) {
    if (sh.capacity == 0) continue;
    const int rc = gather_range(runs, start, es, sh.start, sh.start + sh.capacity, sh.dev, sh.ptr);
    if (rc != NL_OK) { nljit_raise(rc, "jit: repartition copy failed"); storage_decref(out); return nullptr; }
  }
  const int rc = sync_devices(nljit_device_count());
  if (rc != NL_OK) { nljit_raise(rc, "jit: repartition failed"); storage_decref(out); return nullptr; }
  return out;
}

}  // namespace

int storage_copy_range(Storage *src, int64_t lo, int64_t hi, int dst_dev, void *dst) {
  std::vector<Buf> runs;
  std::vector<int64_t> start;
  int64_t run = 0;
  for (const Shard &sh : src->shards) {
    Buf b;
    b.dev = sh.dev; b.ptr = sh.ptr; b.count = sh.count;
    runs.push_back(b);
    start.push_back(run);
    run += sh.count;
  }
  return gather_range(runs, start, nl_dtype_size(src->dtype), lo, hi, dst_dev, dst);
}

Storage *storage_repartition(Storage *in) {
  if (nljit_ensure_runtime() < 0) return nullptr;
  if (storage_settle(in) < 0) return nullptr;
  int rc = sync_devices(nljit_device_count());
  if (rc != NL_OK) { nljit_raise(rc, "jit: device synchronisation failed"); return nullptr; }
  std::vector<Buf> runs;
  for (const Shard &sh : in->shards) {
    Buf b;
    b.dev = sh.dev; b.ptr = sh.ptr; b.count = sh.count;
    runs.push_back(b);
  }
  return regular_from_runs(in->dtype, runs, in->cardinality());
}

namespace {

// a sorted copy of one shard (tmp freed before returning); the caller synchronises
int sorted_copy(const Shard &sh, int dtype, size_t es, Buf *out) {
  out->dev = sh.dev;
  out->count = sh.count;
  if (sh.count == 0) return NL_OK;
  void *tmp = nullptr;
  int rc = nl_malloc(sh.dev, (size_t)sh.count * es, &out->ptr);
  if (rc == NL_OK) rc = nl_malloc(sh.dev, (size_t)sh.count * es, &tmp);
  if (rc == NL_OK) rc = nl_memcpy_d2d(sh.dev, 0, out->ptr, sh.ptr, (size_t)sh.count * es);
  if (rc == NL_OK) {
    Py_BEGIN_ALLOW_THREADS
    rc = nl_sort(sh.dev, 0, dtype, out->ptr, tmp, sh.count);
    Py_END_ALLOW_THREADS
  }
  if (tmp) nl_free(sh.dev, tmp);
  return rc;
}

}  // namespace

Storage *device_sort(Storage *in) {
  if (nljit_ensure_runtime() < 0) return nullptr;
  const int G = nljit_device_count();
  const int dtype = in->dtype;
  const size_t es = nl_dtype_size(dtype);
  const int64_t total = in->cardinality();
  int rc = NL_OK;
  auto fail = [&](int code, const char *what) -> Storage * { nljit_raise(code, what); return nullptr; };
  if (storage_settle(in) < 0) return nullptr;
  // evaluation no longer ends with idle devices: peers read these shards over NVLink below
  if (G > 1 && (rc = sync_devices(G)) != NL_OK) return fail(rc, "jit: device synchronisation failed");

  const size_t S = in->shards.size();
  int nonempty = 0;
  for (const Shard &sh : in->shards) nonempty += sh.count > 0;
  if (G == 1) {
    // one device: the copy the sort works on IS the result column (one shard of exactly `total` elements)
    Storage *out = storage_new(dtype, total);
    if (storage_alloc_shards(out) < 0) { storage_decref(out); return nullptr; }
    const Shard &src = in->shards[0];
    Shard &dst = out->shards[0];
    if (total > 0) {
      void *tmp = nullptr;
      rc = nl_memcpy_d2d(dst.dev, 0, dst.ptr, src.ptr, (size_t)total * es);
      if (rc == NL_OK && total > 1) rc = nl_malloc(dst.dev, (size_t)total * es, &tmp);
      if (rc == NL_OK && total > 1) {
        Py_BEGIN_ALLOW_THREADS
        rc = nl_sort(dst.dev, 0, dtype, dst.ptr, tmp, total);
        Py_END_ALLOW_THREADS
      }
      if (tmp) nl_free(dst.dev, tmp);
      if (rc == NL_OK) rc = sync_devices(G);
      if (rc != NL_OK) { storage_decref(out); return fail(rc, "jit: device sort failed"); }
    }
    return out;
  }
  if (nonempty <= 1) {
    // one run on one of several devices: sort a copy, then lay it out as a regular column
    std::vector<Buf> sorted(S);
    for (size_t g = 0; g < S && rc == NL_OK; g++) rc = sorted_copy(in->shards[g], dtype, es, &sorted[g]);
    if (rc == NL_OK) rc = sync_devices(G);
    Storage *out = (rc == NL_OK) ? regular_from_runs(dtype, sorted, total) : fail(rc, "jit: device sort failed");
    free_bufs(sorted);
    return out;
  }
  if (G - 1 > 15) return fail(NL_ERR_UNSUPPORTED, "jit: sort() across more than 16 devices");

  // NL_DEBUG_SORT=1: wall-clock time of every step on stderr
  const bool trace = getenv("NL_DEBUG_SORT") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!trace) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[jit sort] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  // page-locked scratch for what the devices report: samples and run lengths
  const int kSamples = 256;
  uint64_t *host_samples = nullptr;
  int64_t *host_counts = nullptr;
  rc = nl_host_alloc(sizeof(uint64_t) * S * kSamples + sizeof(int64_t) * S * (size_t)(G + 1), (void **)&host_samples);
  if (rc != NL_OK) return fail(rc, "jit: page-locked allocation failed");
  host_counts = (int64_t *)(host_samples + S * kSamples);
  lap("page-locked scratch");
  std::vector<Buf> bucket((size_t)G);
  auto cleanup = [&]() { free_bufs(bucket); nl_host_free(host_samples); };

  // 1. splitters from regularly spaced samples of every shard
  for (size_t g = 0; g < S && rc == NL_OK; g++)
    if (in->shards[g].count > 0)
      rc = nl_sort_sample(in->shards[g].dev, 0, dtype, in->shards[g].ptr, in->shards[g].count, kSamples, host_samples + g * kSamples);
  if (rc == NL_OK) rc = sync_devices(G);
  if (rc != NL_OK) { cleanup(); return fail(rc, "jit: device sort failed (sampling)"); }
  std::vector<uint64_t> samples;
  for (size_t g = 0; g < S; g++)
    if (in->shards[g].count > 0) samples.insert(samples.end(), host_samples + g * kSamples, host_samples + (g + 1) * kSamples);
  std::sort(samples.begin(), samples.end());
  std::vector<uint64_t> splitter((size_t)G - 1);
  for (int j = 1; j < G; j++) splitter[(size_t)j - 1] = samples[(samples.size() * (size_t)j) / (size_t)G];
  lap("samples and splitters");

  // 2. how many keys of every shard fall between the splitters (one counting pass per GPU, all side by side)
  for (size_t g = 0; g < S && rc == NL_OK; g++) {
    const Shard &sh = in->shards[g];
    for (int j = 0; j <= G; j++) host_counts[g * (size_t)(G + 1) + (size_t)j] = 0;
    if (sh.count == 0) continue;
    rc = nl_sort_split_count(sh.dev, 0, dtype, sh.ptr, sh.count, splitter.data(), G - 1, host_counts + g * (size_t)(G + 1));
  }
  if (rc == NL_OK) rc = sync_devices(G);
  if (rc != NL_OK) { cleanup(); return fail(rc, "jit: device sort failed (counting)"); }
  lap("counts at the splitters");

  // 3. bucket j on device j = run j of every shard, in shard order.  The sweep that cuts a shard stores every run
  //    straight into its place in the destination GPU's bucket — the exchange happens inside the kernel, over NVLink,
  //    while it regroups (first version of the round: grouped locally, then peer copies — one more pass over every key)
  std::vector<std::vector<int64_t>> at(S, std::vector<int64_t>((size_t)G, 0));    // where run j of shard g starts in bucket j
  for (int j = 0; j < G; j++) {
    bucket[(size_t)j].dev = j;
    for (size_t g = 0; g < S; g++) {
      at[g][(size_t)j] = bucket[(size_t)j].count;
      bucket[(size_t)j].count += host_counts[g * (size_t)(G + 1) + (size_t)j];
    }
    if (bucket[(size_t)j].count == 0) continue;
    rc = nl_malloc(j, (size_t)bucket[(size_t)j].count * es, &bucket[(size_t)j].ptr);
    if (rc != NL_OK) { cleanup(); return fail(rc, "jit: device allocation failed"); }
  }
  rc = sync_devices(G);                 // the buckets exist before a peer writes into them
  for (size_t g = 0; g < S && rc == NL_OK; g++) {
    const Shard &sh = in->shards[g];
    if (sh.count == 0) continue;
    void *dst[16];
    for (int j = 0; j < G; j++)
      dst[j] = bucket[(size_t)j].ptr ? (void *)((char *)bucket[(size_t)j].ptr + (size_t)at[g][(size_t)j] * es) : nullptr;
    rc = nl_sort_split_scatter(sh.dev, 0, dtype, sh.ptr, sh.count, splitter.data(), G - 1, dst);
  }
  if (rc == NL_OK) rc = sync_devices(G);
  if (rc != NL_OK) { cleanup(); return fail(rc, "jit: device sort failed (exchange)"); }
  lap("cut + exchange (one sweep)");

  // 4. sort the buckets — one host thread per device: nl_sort() synchronises its stream once (it reads the histogram
  //    that tells it which digit positions to skip), and G of those waits in a row were a third of the sort at 8 GPUs
  {
    std::vector<void *> tmp((size_t)G, nullptr);
    std::vector<int> rcs((size_t)G, NL_OK);
    for (int j = 0; j < G && rc == NL_OK; j++)
      if (bucket[(size_t)j].count > 1) rc = nl_malloc(j, (size_t)bucket[(size_t)j].count * es, &tmp[(size_t)j]);
    lap("  scratch for the buckets");
    if (rc == NL_OK) {
      Py_BEGIN_ALLOW_THREADS
      std::vector<std::thread> workers;
      for (int j = 0; j < G; j++) {
        if (bucket[(size_t)j].count <= 1) continue;
        workers.emplace_back([&, j]() {
          Buf &b = bucket[(size_t)j];
          int r = nl_sort(j, 0, dtype, b.ptr, tmp[(size_t)j], b.count);
          if (r == NL_OK) r = nl_stream_sync(j, 0);
          rcs[(size_t)j] = r;
        });
      }
      for (std::thread &w : workers) w.join();
      Py_END_ALLOW_THREADS
      for (int j = 0; j < G; j++)
        if (rcs[(size_t)j] != NL_OK) rc = rcs[(size_t)j];
    }
    for (int j = 0; j < G; j++)
      if (tmp[(size_t)j]) nl_free(j, tmp[(size_t)j]);
    if (rc != NL_OK) { cleanup(); return fail(rc, "jit: device sort failed"); }
  }
  lap("bucket sorts");

  // 5. regular partition
  Storage *out = regular_from_runs(dtype, bucket, total);
  cleanup();
  lap("regular partition");
  return out;
}
