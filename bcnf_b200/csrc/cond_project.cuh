// Condition projection  P = h . W1h^T + b1  for all conditioner networks of the stack at once.
//
// The reference concatenates the features to every row and multiplies them inside each
// block's first Linear (torch.cat([y, h]) -> nn.Linear, cnf.py:101-104), i.e. once per
// (sample, instance) row and per block.  Here the h-columns of all first Linears are stacked
// into one (C, PW) matrix and applied once per conditioning instance.
#pragma once
#include "common.cuh"

namespace bcnf {

constexpr int kProjBM = 64, kProjBN = 64, kProjBK = 16;

// P (M, N) = h (M, K) @ Wp (K, N) + bp (N);  N multiple of 4.
__global__ void __launch_bounds__(256)
cond_project_kernel(const float* __restrict__ h, const float* __restrict__ Wp, const float* __restrict__ bp,
                    float* __restrict__ P, long long M, int N, int K) {
  __shared__ float As[kProjBK][kProjBM + 4];   // h tile, transposed: As[k][m]
  __shared__ __align__(16) float Bs[kProjBK][kProjBN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;      // 16 x 16 threads, 4 x 4 outputs each
  const long long m0 = (long long)blockIdx.y * kProjBM;
  const int n0 = blockIdx.x * kProjBN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += kProjBK) {
    // h tile: 64 rows x 16 k, k contiguous in memory
    for (int e = tid; e < kProjBM * kProjBK; e += 256) {
      const int m = e >> 4, k = e & 15;
      const long long gm = m0 + m;
      As[k][m] = (gm < M && k0 + k < K) ? __ldg(h + gm * K + k0 + k) : 0.f;
    }
    for (int e = tid; e < kProjBK * kProjBN / 4; e += 256) {
      const int k = e >> 4, q = e & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < K && n0 + q * 4 < N) v = __ldg(reinterpret_cast<const float4*>(Wp + (size_t)(k0 + k) * N + n0 + q * 4));
      *reinterpret_cast<float4*>(&Bs[k][q * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kProjBK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n < N) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + n));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long m = m0 + ty * 4 + i;
      if (m < M)
        *reinterpret_cast<float4*>(P + m * N + n) =
            make_float4(acc[i][0] + b4.x, acc[i][1] + b4.y, acc[i][2] + b4.z, acc[i][3] + b4.w);
    }
  }
}

}  // namespace bcnf
