// Training primitives of the conditioner MLP (fp32): one strided, tiled SGEMM with fused epilogues.
//
// The training step of the reference (Trainer._train_batch, src/bcnf/train/trainer.py:244-277) runs the
// conditioner as nn.Linear -> nn.GELU -> nn.Dropout per hidden layer (cnf.py:78-83) and back-propagates
// through it with autograd: per Linear one addmm forward, one mm for the data gradient, one mm for the
// weight gradient, a column sum for the bias gradient, plus separate gelu / dropout kernels.  Here the
// three GEMMs are one kernel with the surrounding element-wise work fused into its epilogue:
//
//   forward        h_l   = drop(gelu(h_{l-1} W_l^T + b_l)),  pre-activation saved       (EPI_BIAS_GELU_DROP)
//   data gradient  da_{l-1} = (da_l W_l) * gelu'(a_{l-1}) * mask_{l-1} / (1-p)            (EPI_DGELU_DROP)
//   weight gradient dW_l += da_l^T h_{l-1}                                                (EPI_NONE, beta = 1)
//
// Dropout masks are a counter-based hash of (seed, layer id, row, column): regenerated, never stored.
// Parameters stay in the reference's own layout ((out, in) row-major), so gradients land directly in
// tensors shaped like the parameters.
#pragma once
#include "common.cuh"

namespace bcnf {

enum TrainEpi : int { TEPI_NONE = 0, TEPI_BIAS = 1, TEPI_BIAS_GELU_DROP = 2, TEPI_DGELU_DROP = 3 };

struct GemmArgs {
  const float* A;  // A(i, r) = A[i*as0 + r*as1]
  const float* B;  // B(r, j) = B[r*bs0 + j*bs1]
  float* C;        // C(i, j) = C[i*cs0 + j]
  int M, N, K;
  long long as0, as1, bs0, bs1, cs0;
  float beta;          // C = epi(acc) + beta * C   (beta in {0, 1})
  int epi;
  const float* bias;   // [N]
  float* save;         // TEPI_BIAS_GELU_DROP: pre-activation out, same indexing as C
  const float* saved;  // TEPI_DGELU_DROP: pre-activation in, same indexing as C
  unsigned long long seed;
  unsigned int layer_uid;
  float p_drop;
  const unsigned long long* seed_ptr;   // optional device-resident seed word, XORed into `seed` (CUDA-graph replays)
  float* colsum;           // optional [N]: colsum[j] += sum_i C(i, j) of the values written (bias gradient of the layer below)
  float* ws;               // optional split-K workspace (fp32, zero on entry, left zero) and its size in floats
  unsigned int* counters;  // optional split-K tile counters (zero on entry, left zero) and their number
  long long ws_floats;
  int n_counters;
  int split_k;             // 0: library decides (needs ws / counters); 1: never split; n > 1: at most n splits
  // operand images (train_tc.cuh): when a_img and b_img are given the TMA-fed kernel runs and A / B are not read
  const unsigned char* a_img; long long a_plane; int a_rpad; int img_mn;   // img_mn: weight-gradient mode, see train_tc.cuh
  const unsigned char* b_img; long long b_plane; int b_rpad; int pad1;
  unsigned char* c_img; long long c_plane; int c_rpad; int pad2;     // optional: image of the values written to C
};

// uniform [0,1) from (seed, layer, element): two rounds of a 64-bit mix (splitmix64 finaliser)
__host__ __device__ __forceinline__ float dropout_uniform(unsigned long long seed, unsigned int layer_uid,
                                                          unsigned long long elem) {
  unsigned long long x = seed ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(layer_uid + 1u)) ^ (elem * 0xD1B54A32D192ED03ull);
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (float)(x >> 40) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float dgelu_erf(float x) {
  // d/dx [0.5 x (1 + erf(x/sqrt2))] = Phi(x) + x phi(x)
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return fmaf(x, pdf, cdf);
}

// same derivative with the branch-free erf of gelu_erf_fast (|erf error| <= 1.1e-7) and an ex2-based Gaussian
__device__ __forceinline__ float dgelu_erf_fast(float x) {
  const float t = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
  float p = -7.638800192e-07f;
  p = fmaf(p, t, 1.656447655e-05f);
  p = fmaf(p, t, -1.540620985e-04f);
  p = fmaf(p, t, 7.796742463e-04f);
  p = fmaf(p, t, -2.041655680e-03f);
  p = fmaf(p, t, -2.589820766e-04f);
  p = fmaf(p, t, 2.797563118e-02f);
  p = fmaf(p, t, -1.483925716e-01f);
  p = fmaf(p, t, -9.184330629e-01f);
  p = fmaf(p, t, -1.627907331e+00f);
  float e, gs;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p * t));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(gs) : "f"(-0.72134752044448170368f * x * x));   // exp(-x^2/2)
  const float cdf = fmaf(0.5f, copysignf(1.0f - e, x), 0.5f);
  return fmaf(x * 0.39894228040143267794f, gs, cdf);
}

template <int BM, int BN>
__global__ void __launch_bounds__(256)
train_gemm_kernel(const GemmArgs g) {
  constexpr int BK = 32;
  constexpr int TM = BM / 16, TN = BN / 16;     // per-thread micro tile (16 x 16 threads)
  constexpr int LA = BM * BK / 256, LB = BN * BK / 256;   // global loads per thread and tile
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const unsigned long long seed = g.seed_ptr ? (g.seed ^ *g.seed_ptr) : g.seed;
  float acc[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;

  const bool a_r_fast = g.as1 == 1;   // r contiguous in A
  const bool b_j_fast = g.bs1 == 1;   // j contiguous in B
  float ra[LA], rb[LB];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int t = 0; t < LA; ++t) {
      const int e = tid + t * 256;
      int i, r;
      if (a_r_fast) { i = e / BK; r = e % BK; } else { r = e / BM; i = e % BM; }
      const int gi = i0 + i, gr = k0 + r;
      ra[t] = (gi < g.M && gr < g.K) ? __ldg(g.A + gi * g.as0 + gr * g.as1) : 0.f;
    }
#pragma unroll
    for (int t = 0; t < LB; ++t) {
      const int e = tid + t * 256;
      int j, r;
      if (b_j_fast) { r = e / BN; j = e % BN; } else { j = e / BK; r = e % BK; }
      const int gj = j0 + j, gr = k0 + r;
      rb[t] = (gj < g.N && gr < g.K) ? __ldg(g.B + gr * g.bs0 + gj * g.bs1) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int t = 0; t < LA; ++t) {
      const int e = tid + t * 256;
      int i, r;
      if (a_r_fast) { i = e / BK; r = e % BK; } else { r = e / BM; i = e % BM; }
      As[buf][r][i] = ra[t];
    }
#pragma unroll
    for (int t = 0; t < LB; ++t) {
      const int e = tid + t * 256;
      int j, r;
      if (b_j_fast) { r = e / BN; j = e % BN; } else { j = e / BK; r = e % BK; }
      Bs[buf][r][j] = rb[t];
    }
  };

  // double-buffered K loop: the global loads of tile k+1 are in flight while tile k is multiplied
  load_tile(0);
  store_tile(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    const bool more = k0 + BK < g.K;
    if (more) load_tile(k0 + BK);
#pragma unroll
    for (int r = 0; r < BK; ++r) {
      float av[TM], bv[TN];
#pragma unroll
      for (int a = 0; a < TM; ++a) av[a] = As[buf][r][ty * TM + a];
#pragma unroll
      for (int b = 0; b < TN; ++b) bv[b] = Bs[buf][r][tx * TN + b];
#pragma unroll
      for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    if (more) {
      store_tile(buf ^ 1);     // the other buffer was last read one iteration ago, before the previous barrier
      __syncthreads();
      buf ^= 1;
    }
  }

  const float keep_scale = g.p_drop > 0.f ? 1.0f / (1.0f - g.p_drop) : 1.0f;
#pragma unroll
  for (int a = 0; a < TM; ++a) {
    const int i = i0 + ty * TM + a;
    if (i >= g.M) continue;
#pragma unroll
    for (int b = 0; b < TN; ++b) {
      const int j = j0 + tx * TN + b;
      if (j >= g.N) continue;
      const long long idx = i * g.cs0 + j;
      float v = acc[a][b];
      if (g.epi == TEPI_BIAS) {
        v += __ldg(g.bias + j);
      } else if (g.epi == TEPI_BIAS_GELU_DROP) {
        v += __ldg(g.bias + j);
        g.save[idx] = v;
        v = gelu_erf(v);
        if (g.p_drop > 0.f)
          v = dropout_uniform(seed, g.layer_uid, (unsigned long long)i * (unsigned)g.N + (unsigned)j) >= g.p_drop
                  ? v * keep_scale : 0.f;
      } else if (g.epi == TEPI_DGELU_DROP) {
        v *= dgelu_erf(__ldg(g.saved + idx));
        if (g.p_drop > 0.f)
          v = dropout_uniform(seed, g.layer_uid, (unsigned long long)i * (unsigned)g.N + (unsigned)j) >= g.p_drop
                  ? v * keep_scale : 0.f;
      }
      if (g.beta != 0.f) v += g.beta * g.C[idx];
      g.C[idx] = v;
    }
  }
}

// out[j] = sum_i X[i*ldx + j] (+ beta * out[j]): bias gradients and ActNorm reductions.
// Block = 32 columns x 8 row groups; each thread sums every 8th row (coalesced across the 32 columns), then the 8
// partial sums of a column meet in shared memory.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int M, int N, long long ldx,
                                                     float* __restrict__ out, float beta) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (j < N) {
    int i = ty;
    for (; i + 24 < M; i += 32) {
      s0 += X[i * ldx + j]; s1 += X[(i + 8) * ldx + j]; s2 += X[(i + 16) * ldx + j]; s3 += X[(i + 24) * ldx + j];
    }
    for (; i < M; i += 8) s0 += X[i * ldx + j];
  }
  part[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && j < N) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += part[r][tx];
    out[j] = beta != 0.f ? fmaf(beta, out[j], s) : s;
  }
}

__global__ void dropout_mask_kernel(float* __restrict__ out, int M, int N, unsigned long long seed, unsigned int layer_uid, float p,
                                    const unsigned long long* __restrict__ seed_ptr) {
  if (seed_ptr) seed ^= *seed_ptr;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)M * N) return;
  const float scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  out[e] = (p > 0.f && dropout_uniform(seed, layer_uid, (unsigned long long)e) < p) ? 0.f : scale;
}

// ---- Trainer step fusion (reference trainer.py:267-277) -----------------------------------
// Adam over ONE flat parameter / gradient / moment blob (torch.optim.Adam semantics, no amsgrad): a single pass over
// 4 x n floats instead of one multi-tensor launch per chunk of parameter tensors.  hyper = [lr, beta1, beta2, eps,
// weight_decay] and step (the 1-based step count, already advanced) live on the device: a captured step replays with
// the current values.  n is a multiple of 4 (views are 64-element aligned; padding holds zeros and stays zero).
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n4,
                                                        const float* __restrict__ hyper, const float* __restrict__ step) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float t = *step;
  const float bc1 = 1.0f - powf(b1, t), bc2_sqrt = sqrtf(1.0f - powf(b2, t));
  const float step_size = lr / bc1;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    float* pe = reinterpret_cast<float*>(&pp);
    float* me = reinterpret_cast<float*>(&mm);
    float* ve = reinterpret_cast<float*>(&vv);
    const float* ge = reinterpret_cast<const float*>(&gg);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = wd != 0.f ? fmaf(wd, pe[k], ge[k]) : ge[k];
      me[k] = fmaf(1.0f - b1, gk - me[k], me[k]);                       // exp_avg.lerp_(grad, 1 - beta1)
      ve[k] = fmaf(1.0f - b2, gk * gk, b2 * ve[k]);                     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(ve[k]) / bc2_sqrt + eps;
      pe[k] -= step_size * (me[k] / denom);
    }
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
}

// NLL of a batch and the gradients it sends back, in one launch (utils.py:49-53): loss = mean_b(0.5 sum_j z^2 - logdet),
// dz = z / B, dlogdet = -1 / B.  One block, rows strided over its threads, fixed reduction order.
__global__ void __launch_bounds__(1024) nll_kernel(const float* __restrict__ z, const float* __restrict__ ld, int B, int D,
                                                   float* __restrict__ loss, float* __restrict__ dz, float* __restrict__ dld) {
  __shared__ float part[32];
  const float inv = 1.0f / (float)B;
  float acc = 0.f;
  for (int r = threadIdx.x; r < B; r += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < D; ++j) {
      const float x = z[(long long)r * D + j];
      s = fmaf(x, x, s);
      dz[(long long)r * D + j] = x * inv;
    }
    acc += 0.5f * s - ld[r];
    dld[r] = -inv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *loss = s * inv;
  }
}

}  // namespace bcnf
