// Writers of the operand-image format (train_tc.cuh has the format's description): shared by every kernel that leaves a
// matrix as the operand of a later tensor-core GEMM.  An image of X (R rows x K columns) is X split into bf16 hi and lo
// planes, each [K/64 chunks][R_pad rows][128 bytes], the 16-byte units of a row XOR-swizzled by (row & 7).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bcnf {

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

__device__ __forceinline__ float warp_sum_tc(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// write 8 consecutive columns (n0 .. n0+7, n0 % 8 == 0) of `row` into an image
__device__ __forceinline__ void img_store8(unsigned char* img, long long plane, int rpad, int row, int n0, const float (&v)[8]) {
  const long long off = ((long long)(n0 >> 6) * rpad + row) * 128 + ((((n0 & 63) >> 3) ^ (row & 7)) << 4);
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hi[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    const float h0 = __uint_as_float(hi[i] << 16), h1 = __uint_as_float(hi[i] & 0xffff0000u);
    lo[i] = pack_bf16x2(v[2 * i] - h0, v[2 * i + 1] - h1);
  }
  *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(img + plane + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
// single element (kernels whose threads own one column each)
__device__ __forceinline__ void img_store1(unsigned char* img, long long plane, int rpad, int row, int n, float v) {
  const long long off = ((long long)(n >> 6) * rpad + row) * 128 + ((((n & 63) >> 3) ^ (row & 7)) << 4) + ((n & 7) << 1);
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  *reinterpret_cast<__nv_bfloat16*>(img + off) = h;
  *reinterpret_cast<__nv_bfloat16*>(img + plane + off) = __float2bfloat16_rn(v - __bfloat162float(h));
}

}  // namespace bcnf
