// nl_cast.cu — nl_cast(): convert a column to another element type, dst[i] = (D)src[i].
//
// The reference gives a binary operation the dtype of its LEFT operand and leaves "todo: correct type"
// (thunk_as_number.cpp:78), and its materialise store only knows a signed integer cast (compiler.cpp:256-260,
// "FIXME float casts").  Here two columns of different types meet through NumPy's promotion rule: the host
// casts the narrower operand to the common type with this kernel (one streaming pass: read S, write D) and the
// fused pipeline then runs in that type.  C conversion semantics, which are NumPy's astype() for every
// in-range value (integer narrowing wraps; float -> integer truncates towards zero).
#include <cuda_runtime.h>

#include "nl_internal.h"

namespace nl {

template <typename S, typename D>
__global__ void __launch_bounds__(256) cast_kernel(const S *__restrict__ src, D *__restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; base < n; base += stride) {
    // four consecutive elements per thread (a warp covers a contiguous run of 128 elements on either side)
    S v[4];
    const int m = (n - base >= 4) ? 4 : (int)(n - base);
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = (j < m) ? __ldcs(src + base + j) : S(0);
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (j < m) __stcs(dst + base + j, (D)v[j]);
  }
}

template <typename S>
static int cast_from(cudaStream_t st, unsigned grid, const void *src, void *dst, int dst_dtype, long long n) {
  switch (dst_dtype) {
#define NL_CASE(dt, D, sfx) case dt: cast_kernel<S, D><<<grid, 256, 0, st>>>((const S *)src, (D *)dst, n); return NL_OK;
    NL_FOR_VALUE_DTYPES(NL_CASE)
#undef NL_CASE
    case NL_BOOL: return NL_ERR_UNSUPPORTED;
  }
  return NL_ERR_INVALID;
}

int launch_cast(void *stream, int sm_count, void *dst, int dst_dtype, const void *src, int src_dtype, int64_t n) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return NL_OK;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  int rc = NL_ERR_INVALID;
  switch (src_dtype) {
#define NL_CASE(dt, S, sfx) case dt: rc = cast_from<S>(st, (unsigned)blocks, src, dst, dst_dtype, n); break;
    NL_FOR_VALUE_DTYPES(NL_CASE)
#undef NL_CASE
    case NL_BOOL: rc = cast_from<unsigned char>(st, (unsigned)blocks, src, dst, dst_dtype, n); break;   // 0x00 / 0x01
  }
  if (rc == NL_ERR_UNSUPPORTED) { set_error("nl_cast: casting to bool is not supported (compare with 0 instead)"); return rc; }
  if (rc != NL_OK) { set_error("nl_cast: bad dtype"); return rc; }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("cast kernel launch failed: %s", cudaGetErrorString(e)); return NL_ERR_CUDA; }
  count_launch();
  return NL_OK;
}

}  // namespace nl
