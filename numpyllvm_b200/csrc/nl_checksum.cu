// nl_checksum.cu — nl_checksum(): three 64-bit checksums of a device buffer, the device half of the
// full-size parity check (last row: "64-bit checksums computed chunk-wise on host and on
// device").  The buffer is read as little-endian 64-bit words w_0 .. w_{n-1} (a trailing partial word
// is zero-padded):
//     out[0] = sum  w_i                (mod 2^64)   which values
//     out[1] = xor  w_i                             (a second, independent witness)
//     out[2] = sum  w_i * (2 i + 1)    (mod 2^64)   and in which ORDER (a stable compaction)
// All three are sums in a commutative ring, so the host can form them chunk by chunk from the
// oracle's output without holding the column:  out[2] of a chunk that starts at word b is
// local[2] + 2 b local[0].  One streaming pass: 128-bit loads, warp shuffles, one atomic per warp.
#include <cuda_runtime.h>

#include "nl_internal.h"

namespace nl {

__global__ void __launch_bounds__(256) checksum_kernel(const unsigned long long *__restrict__ w, unsigned long long n_words,
                                                       unsigned long long tail_word, int has_tail, unsigned long long *__restrict__ out) {
  unsigned long long s = 0, x = 0, p = 0;
  const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    const ulonglong2 *w2 = reinterpret_cast<const ulonglong2 *>(w);
    const unsigned long long n2 = n_words >> 1;
    for (unsigned long long k = gtid; k < n2; k += stride) {
      const ulonglong2 v = __ldcs(w2 + k);
      s += v.x + v.y;
      x ^= v.x ^ v.y;
      p += v.x * (4 * k + 1) + v.y * (4 * k + 3);
    }
    if (gtid == 0 && (n_words & 1)) {
      const unsigned long long v = w[n_words - 1];
      s += v; x ^= v; p += v * (2 * (n_words - 1) + 1);
    }
  } else {
    for (unsigned long long k = gtid; k < n_words; k += stride) {
      const unsigned long long v = __ldcs(w + k);
      s += v; x ^= v; p += v * (2 * k + 1);
    }
  }
  if (gtid == 0 && has_tail) { s += tail_word; x ^= tail_word; p += tail_word * (2 * n_words + 1); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    x ^= __shfl_xor_sync(0xffffffffu, x, o);
    p += __shfl_xor_sync(0xffffffffu, p, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 0, s);
    atomicXor(out + 1, x);
    atomicAdd(out + 2, p);
  }
}

int launch_checksum(void *stream, int sm_count, const void *ptr, size_t nbytes, unsigned long long *out) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, 3 * sizeof(unsigned long long), st);
  if (e != cudaSuccess) { set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e)); return NL_ERR_CUDA; }
  const unsigned long long n_words = nbytes / 8;
  const int tail = (int)(nbytes % 8);
  unsigned long long tail_word = 0;
  if (tail) {
    // the last 1..7 bytes, zero-padded (fetched through the stream so the value is the data's at this point)
    e = cudaMemcpyAsync(&tail_word, (const char *)ptr + n_words * 8, (size_t)tail, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("nl_checksum: tail read failed: %s", cudaGetErrorString(e)); return NL_ERR_CUDA; }
  }
  unsigned long long want = (n_words / 2 + 255) / 256;
  const unsigned long long cap = (unsigned long long)sm_count * 8;
  unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
  checksum_kernel<<<grid, 256, 0, st>>>((const unsigned long long *)ptr, n_words, tail_word, tail != 0, out);
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("checksum kernel launch failed: %s", cudaGetErrorString(e)); return NL_ERR_CUDA; }
  count_launch();
  return NL_OK;
}

}  // namespace nl
