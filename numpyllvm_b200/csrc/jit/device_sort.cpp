// device_sort.cpp — see device_sort.hpp.
//
// One GPU: copy the shard, nl_sort() it (LSD radix sort, nl_sort.cu).
// Several GPUs: sample sort over NVLink —
//   1. every GPU radix-sorts a copy of its shard;
//   2. the host reads a few regularly spaced samples of every sorted shard and picks G - 1
//      splitters (in the kernels' key order, so NaNs stay last);
//   3. every sorted shard is cut at the splitters (binary search over single-element reads);
//   4. piece j of every shard goes to GPU j by peer copy, in shard order;
//   5. every GPU sorts what it received — bucket j now holds exactly the elements between
//      splitters j - 1 and j, so the buckets in order are the sorted array;
//   6. the buckets are re-cut into the engine's regular partition (peer copies again), so the
//      result can meet any other column of the same length in a fused pipeline.
#include "device_sort.hpp"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

// the kernels' order-preserving bijection (nl_sort.cu: order_key), for the host-side splitters
uint64_t host_order_key(uint64_t bits, int dtype) {
  switch (dtype) {
    case NL_I8: return (uint64_t)((uint8_t)bits ^ 0x80u);
    case NL_I16: return (uint64_t)((uint16_t)bits ^ 0x8000u);
    case NL_U8: return (uint64_t)(uint8_t)bits;
    case NL_U16: return (uint64_t)(uint16_t)bits;
    case NL_U32: return (uint64_t)(uint32_t)bits;
    case NL_U64: return bits;
    case NL_I32: return (uint64_t)((uint32_t)bits ^ 0x80000000u);
    case NL_I64: return bits ^ 0x8000000000000000ull;
    case NL_F32: {
      const uint32_t b = (uint32_t)bits;
      const uint32_t k = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
      return (uint64_t)(uint32_t)(k - 0x007fffffu);
    }
    default: {
      const uint64_t k = bits ^ ((bits >> 63) ? ~0ull : 0x8000000000000000ull);
      return k - 0x000fffffffffffffull;
    }
  }
}

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

// element `i` of a device buffer as its order key
int read_key(const Buf &b, size_t es, int dtype, int64_t i, uint64_t *key) {
  uint64_t bits = 0;
  int rc = nl_memcpy_d2h(b.dev, 0, &bits, (const char *)b.ptr + (size_t)i * es, es);
  if (rc == NL_OK) rc = nl_stream_sync(b.dev, 0);
  if (rc != NL_OK) return rc;
  *key = host_order_key(bits, dtype);
  return NL_OK;
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

  // 1. a sorted copy of every shard
  std::vector<Buf> sorted((size_t)in->shards.size());
  for (size_t g = 0; g < in->shards.size(); g++) {
    const Shard &sh = in->shards[g];
    sorted[g].dev = sh.dev;
    sorted[g].count = sh.count;
    if (sh.count == 0) continue;
    void *tmp = nullptr;
    rc = nl_malloc(sh.dev, (size_t)sh.count * es, &sorted[g].ptr);
    if (rc == NL_OK) rc = nl_malloc(sh.dev, (size_t)sh.count * es, &tmp);
    if (rc == NL_OK) rc = nl_memcpy_d2d(sh.dev, 0, sorted[g].ptr, sh.ptr, (size_t)sh.count * es);
    if (rc == NL_OK) {
      Py_BEGIN_ALLOW_THREADS
      rc = nl_sort(sh.dev, 0, dtype, sorted[g].ptr, tmp, sh.count);
      Py_END_ALLOW_THREADS
    }
    if (tmp) nl_free(sh.dev, tmp);
    if (rc != NL_OK) { free_bufs(sorted); return fail(rc, "jit: device sort failed"); }
  }
  int nonempty = 0;
  for (const Buf &b : sorted) nonempty += b.count > 0;
  if (G == 1 || nonempty <= 1) {
    // one sorted run: it only has to be laid out as a regular column
    rc = sync_devices(G);
    Storage *out = (rc == NL_OK) ? regular_from_runs(dtype, sorted, total) : fail(rc, "jit: device sort failed");
    free_bufs(sorted);
    return out;
  }

  // 2. splitters from regularly spaced samples
  const int kSamples = 64;
  std::vector<uint64_t> samples;
  for (const Buf &b : sorted) {
    for (int i = 0; i < kSamples && b.count > 0; i++) {
      uint64_t key;
      rc = read_key(b, es, dtype, (int64_t)(((__int128)b.count * i) / kSamples), &key);
      if (rc != NL_OK) { free_bufs(sorted); return fail(rc, "jit: device sort failed (sampling)"); }
      samples.push_back(key);
    }
  }
  std::sort(samples.begin(), samples.end());
  std::vector<uint64_t> splitter((size_t)G - 1);
  for (int j = 1; j < G; j++) splitter[(size_t)j - 1] = samples[(samples.size() * (size_t)j) / (size_t)G];

  // 3. cut[g][j] = first element of sorted shard g whose key is > splitter j - 1 (cut[g][0] = 0)
  std::vector<std::vector<int64_t>> cut(sorted.size(), std::vector<int64_t>((size_t)G + 1, 0));
  for (size_t g = 0; g < sorted.size(); g++) {
    cut[g][(size_t)G] = sorted[g].count;
    for (int j = 1; j < G; j++) {
      int64_t lo = cut[g][(size_t)j - 1], hi = sorted[g].count;
      while (lo < hi) {
        const int64_t mid = lo + (hi - lo) / 2;
        uint64_t key;
        rc = read_key(sorted[g], es, dtype, mid, &key);
        if (rc != NL_OK) { free_bufs(sorted); return fail(rc, "jit: device sort failed (partition)"); }
        if (key <= splitter[(size_t)j - 1]) lo = mid + 1; else hi = mid;
      }
      cut[g][(size_t)j] = lo;
    }
  }

  // 4. bucket j on device j: piece j of every shard, in shard order
  std::vector<Buf> bucket((size_t)G);
  for (int j = 0; j < G; j++) {
    bucket[(size_t)j].dev = j;
    for (size_t g = 0; g < sorted.size(); g++) bucket[(size_t)j].count += cut[g][(size_t)j + 1] - cut[g][(size_t)j];
    if (bucket[(size_t)j].count == 0) continue;
    rc = nl_malloc(j, (size_t)bucket[(size_t)j].count * es, &bucket[(size_t)j].ptr);
    if (rc != NL_OK) { free_bufs(sorted); free_bufs(bucket); return fail(rc, "jit: device allocation failed"); }
  }
  rc = sync_devices(G);
  for (int j = 0; j < G && rc == NL_OK; j++) {
    int64_t at = 0;
    for (size_t g = 0; g < sorted.size() && rc == NL_OK; g++) {
      const int64_t n = cut[g][(size_t)j + 1] - cut[g][(size_t)j];
      if (n == 0) continue;
      const char *from = (const char *)sorted[g].ptr + (size_t)cut[g][(size_t)j] * es;
      char *to = (char *)bucket[(size_t)j].ptr + (size_t)at * es;
      rc = (sorted[g].dev == j) ? nl_memcpy_d2d(j, 0, to, from, (size_t)n * es) : nl_memcpy_peer(j, 0, to, sorted[g].dev, from, (size_t)n * es);
      at += n;
    }
  }
  if (rc == NL_OK) rc = sync_devices(G);
  free_bufs(sorted);
  if (rc != NL_OK) { free_bufs(bucket); return fail(rc, "jit: device sort failed (exchange)"); }

  // 5. sort the buckets
  for (int j = 0; j < G; j++) {
    Buf &b = bucket[(size_t)j];
    if (b.count <= 1) continue;
    void *tmp = nullptr;
    rc = nl_malloc(j, (size_t)b.count * es, &tmp);
    if (rc == NL_OK) {
      Py_BEGIN_ALLOW_THREADS
      rc = nl_sort(j, 0, dtype, b.ptr, tmp, b.count);
      Py_END_ALLOW_THREADS
    }
    if (tmp) nl_free(j, tmp);
    if (rc != NL_OK) { free_bufs(bucket); return fail(rc, "jit: device sort failed"); }
  }
  rc = sync_devices(G);
  if (rc != NL_OK) { free_bufs(bucket); return fail(rc, "jit: device sort failed"); }

  // 6. regular partition
  Storage *out = regular_from_runs(dtype, bucket, total);
  free_bufs(bucket);
  return out;
}
