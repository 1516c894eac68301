// tma_ring_bench.cu — how fast can ONE persistent CTA per SM stream global memory into a shared-
// memory ring with cp.async.bulk?  (Design question behind nl_compact_pipe.cuh: is the producer
// side of the staged compaction pipeline a ceiling below the HBM roofline?)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_ring_bench tools/tma_ring_bench.cu
//   build/tma_ring_bench [log2_bytes=33]
//
// Sweeps slot size, ring depth, bytes per bulk copy and the number of issuing lanes.  Consumers
// only touch one word per 128 bytes of each slot (so the data really has to be in shared memory)
// and hand the slot back.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kConsumerWarps = 8;
constexpr int kMaxSlots = 16;

struct Shared {
  uint64_t full[kMaxSlots];
  uint64_t empty[kMaxSlots];
};

// tiles are handed out statically (tile = it * gridDim + blockIdx): no ticket, no look-back
__global__ void __launch_bounds__((kConsumerWarps + 1) * 32, 1)
ring_kernel(const unsigned char *src, long long n_tiles, int slot_bytes, int slots, int chunk_bytes, int issue_lanes,
            unsigned long long *sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  Shared &sh = *reinterpret_cast<Shared *>(smem);
  unsigned char *ring = smem + 1024;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < slots; s++) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int chunks = slot_bytes / chunk_bytes;
  if (warp == kConsumerWarps) {
    int s = 0; uint32_t par = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (lane == 0) {
        mbar_wait(&sh.empty[s], par ^ 1u);
        mbar_arrive_expect_tx(&sh.full[s], (uint32_t)slot_bytes);
      }
      __syncwarp();
      const unsigned char *g = src + (size_t)tile * slot_bytes;
      unsigned char *d = ring + (size_t)s * slot_bytes;
      for (int c = lane; c < chunks; c += issue_lanes)
        if (lane < issue_lanes) bulk_g2s(d + (size_t)c * chunk_bytes, g + (size_t)c * chunk_bytes, (uint32_t)chunk_bytes, &sh.full[s]);
      s = (s + 1 == slots) ? 0 : s + 1;
      par ^= (s == 0) ? 1u : 0u;
    }
  } else {
    int s = 0; uint32_t par = 0;
    unsigned long long acc = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(&sh.full[s], par);
      const unsigned char *d = ring + (size_t)s * slot_bytes;
      for (int off = (warp * 32 + lane) * 128; off < slot_bytes; off += kConsumerWarps * 32 * 128)
        acc += *reinterpret_cast<const unsigned long long *>(d + off);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.empty[s]);
      s = (s + 1 == slots) ? 0 : s + 1;
      par ^= (s == 0) ? 1u : 0u;
    }
    if (acc == 0x1234567887654321ull) *sink = acc;
  }
}

// reference point: plain 16-byte streaming loads, same bytes
__global__ void __launch_bounds__(256) ldg_kernel(const uint4 *src, long long n16, unsigned long long *sink) {
  unsigned long long acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x * 4) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const long long j = i + (long long)k * gridDim.x * blockDim.x;
      v[k] = (j < n16) ? __ldcs(src + j) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) acc += v[k].x + v[k].w;
  }
  if (acc == 0x1234567887654321ull) *sink = acc;
}

int main(int argc, char **argv) {
  const int log2b = argc > 1 ? atoi(argv[1]) : 33;
  const size_t bytes = (size_t)1 << log2b;
  unsigned char *src; unsigned long long *sink;
  CK(cudaMalloc(&src, bytes));
  CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(src, 1, bytes));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  auto time = [&](auto launch) {
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 5; r++) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / 5;
  };
  {
    const float ms = time([&] { ldg_kernel<<<sms * 8, 256>>>((const uint4 *)src, (long long)(bytes / 16), sink); });
    printf("ldg 16B x4 grid=%d: %.3f ms  %.0f GB/s\n", sms * 8, ms, bytes / ms / 1e6);
  }
  const int slot_kb[] = {16, 32, 64};
  for (int sk : slot_kb) {
    const int slot_bytes = sk * 1024;
    const int max_slots = (226 * 1024 - 1024) / slot_bytes;
    for (int slots : {2, 3, 4, 6, 8, 12}) {
      if (slots > max_slots || slots > kMaxSlots) continue;
      for (int chunk_kb : {2, 8, 32, 64}) {
        const int chunk = chunk_kb * 1024;
        if (chunk > slot_bytes) continue;
        for (int lanes : {1, 8}) {
          if (lanes > slot_bytes / chunk) continue;
          const long long n_tiles = (long long)(bytes / slot_bytes);
          const size_t smem = 1024 + (size_t)slots * slot_bytes;
          const float ms = time([&] {
            ring_kernel<<<sms, (kConsumerWarps + 1) * 32, smem>>>(src, n_tiles, slot_bytes, slots, chunk, lanes, sink);
          });
          CK(cudaGetLastError());
          printf("ring slot=%2dKB slots=%2d chunk=%2dKB lanes=%d: %.3f ms  %.0f GB/s\n", sk, slots, chunk_kb, lanes, ms, bytes / ms / 1e6);
        }
      }
    }
  }
  return 0;
}
