// Transformer condition encoder (reference src/bcnf/models/feature_network.py:183-307), inference: everything that is
// not a Linear.  The Linears (q/k/v merged, fc_out, the two FFN layers, the output layer) run on the CTA-pair GEMM on
// operand images (gemm_img2.cuh); the three kernels here produce those images:
//
//   trf_embed_kernel   tokens (B, T, F) -> x = tokens . Wf^T + bf (+ positional table)        : fp32 x + image of x
//   trf_attn_kernel    q | k | v (rows, 3E) fp32 -> softmax(q k^T / sqrt(hd)) v per instance, head : image of the context
//   trf_add_ln_kernel  x <- LayerNorm(x + y) * gamma + beta  (post-norm block, :255-259)      : fp32 x + image of x
//
// All arithmetic is fp32; an image is the bf16 hi / lo split of the fp32 value (train_tc.cuh: img_store8), so the GEMM
// that reads it sees the value to 2^-17.  HBM-bound kernels: every element is read once and written once (+ its image).
#pragma once
#include "common.cuh"
#include "train_tc.cuh"

namespace bcnf {

struct TrfEmbedArgs {
  const float* tokens;   // (rows, F), rows = B * T
  const float* Wf;       // (E, F)  nn.Linear weight
  const float* bf;       // (E)
  const float* pos;      // (T, E) positional table or null (feature_network.py:291-301)
  float* x;              // (rows, E)
  unsigned char* x_img; long long plane; int rpad;
  long long rows; int T, F, E;
};

// one thread per 8 consecutive columns of one row
__global__ void __launch_bounds__(256) trf_embed_kernel(const TrfEmbedArgs a) {
  const int groups = a.E >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.rows * groups) return;
  const long long row = idx / groups;
  const int n0 = (int)(idx - row * groups) << 3;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __ldg(a.bf + n0 + i);
  const float* tok = a.tokens + row * a.F;
  for (int f = 0; f < a.F; ++f) {
    const float t = __ldg(tok + f);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(t, __ldg(a.Wf + (long long)(n0 + i) * a.F + f), v[i]);
  }
  if (a.pos) {
    const float* p = a.pos + (long long)(row % a.T) * a.E + n0;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += __ldg(p + i);
  }
  float4* dst = reinterpret_cast<float4*>(a.x + row * a.E + n0);
  dst[0] = make_float4(v[0], v[1], v[2], v[3]);
  dst[1] = make_float4(v[4], v[5], v[6], v[7]);
  img_store8(a.x_img, a.plane, a.rpad, (int)row, n0, v);
}

struct TrfAttnArgs {
  const float* qkv;      // (rows, 3E): q | k | v, head h in columns [h*hd, (h+1)*hd) of each third
  unsigned char* ctx_img; long long plane; int rpad;   // image of the concatenated heads (rows, E)
  long long n_inst; int T, E, heads;
  float scale;           // 1 / sqrt(hd)
};

constexpr int kTrfAttnThreads = 256;
constexpr int kTrfMaxT = 64;

// One CTA per instance, one THREAD per (head, query) pair: a warp takes a head, its lanes the queries (lane, lane + 32).
// k and v of the instance's T tokens are staged in shared memory and read as warp-wide broadcasts (all lanes of a warp
// read the same k_j / v_j of their head); q_i, the T scores and the context row live in registers.  Per pair
// 2 * T * HD FMAs and T exps; the first version (one warp per pair, lanes over keys) spent 3.6x the instructions on
// half-empty warps: 2.56 ms per call at 16 384 instances x 30 tokens x 8 heads, this one is bound by its 1 GB of traffic.
// TB = compile-time bound of T (scores stay in registers).
template <int HD, int TB>
__global__ void __launch_bounds__(kTrfAttnThreads, 2) trf_attn_kernel(const TrfAttnArgs a) {
  extern __shared__ __align__(16) float smem_attn[];
  const int T = a.T, E = a.E;
  float* ks = smem_attn;                // [T][E]
  float* vs = ks + T * E;               // [T][E]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kTrfAttnThreads / 32;
  for (long long inst = blockIdx.x; inst < a.n_inst; inst += gridDim.x) {
    const float* src = a.qkv + inst * T * 3 * E;
    const int e4 = E >> 2;
    for (int i = tid; i < T * 2 * e4; i += kTrfAttnThreads) {
      const int t = i / (2 * e4), c4 = i - t * 2 * e4;          // columns E .. 3E of token t: k | v
      const float4 w = __ldg(reinterpret_cast<const float4*>(src + (long long)t * 3 * E + E) + c4);
      float* dst = c4 < e4 ? ks + t * E + (c4 << 2) : vs + t * E + ((c4 - e4) << 2);
      *reinterpret_cast<float4*>(dst) = w;
    }
    __syncthreads();
    for (int h = warp; h < a.heads; h += nwarps) {
      for (int i = lane; i < T; i += 32) {
        float q[HD];
        const float4* qp = reinterpret_cast<const float4*>(src + (long long)i * 3 * E + h * HD);
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) {
          const float4 w = __ldg(qp + c);
          q[4 * c] = w.x * a.scale; q[4 * c + 1] = w.y * a.scale; q[4 * c + 2] = w.z * a.scale; q[4 * c + 3] = w.w * a.scale;
        }
        float sc[TB];
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < TB; ++j) {
          if (j < T) {
            const float4* kp = reinterpret_cast<const float4*>(ks + j * E + h * HD);
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < HD / 4; ++c) {
              const float4 w = kp[c];
              acc = fmaf(q[4 * c], w.x, acc); acc = fmaf(q[4 * c + 1], w.y, acc);
              acc = fmaf(q[4 * c + 2], w.z, acc); acc = fmaf(q[4 * c + 3], w.w, acc);
            }
            sc[j] = acc;
            m = fmaxf(m, acc);
          }
        }
        float l = 0.f;
        float ctx[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) ctx[c] = 0.f;
#pragma unroll
        for (int j = 0; j < TB; ++j) {
          if (j < T) {
            const float p = expf(sc[j] - m);
            l += p;
            const float4* vp = reinterpret_cast<const float4*>(vs + j * E + h * HD);
#pragma unroll
            for (int c = 0; c < HD / 4; ++c) {
              const float4 w = vp[c];
              ctx[4 * c] = fmaf(p, w.x, ctx[4 * c]); ctx[4 * c + 1] = fmaf(p, w.y, ctx[4 * c + 1]);
              ctx[4 * c + 2] = fmaf(p, w.z, ctx[4 * c + 2]); ctx[4 * c + 3] = fmaf(p, w.w, ctx[4 * c + 3]);
            }
          }
        }
        const float inv = 1.0f / l;
        const int row = (int)(inst * T + i);
#pragma unroll
        for (int g = 0; g < HD / 8; ++g) {
          float w[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) w[u] = ctx[8 * g + u] * inv;
          img_store8(a.ctx_img, a.plane, a.rpad, row, h * HD + 8 * g, w);
        }
      }
    }
    __syncthreads();
  }
}

struct TrfAddLnArgs {
  float* x;              // (rows, E) in / out
  const float* y;        // (rows, E): the sublayer output (bias included)
  const float* gamma; const float* beta;
  unsigned char* x_img; long long plane; int rpad;
  long long rows; int E; float eps;
};

constexpr int kTrfLnMaxGroups = 4;      // 8-column groups per lane: E <= 32 * 4 * 8 = 1024

// one warp per row: nn.LayerNorm over the last axis (biased variance, eps inside the square root)
__global__ void __launch_bounds__(256) trf_add_ln_kernel(const TrfAddLnArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const int groups = a.E >> 3;
  float v[kTrfLnMaxGroups][8];
  float sum = 0.f;
#pragma unroll
  for (int g = 0; g < kTrfLnMaxGroups; ++g) {
    const int grp = lane + 32 * g;
    if (grp < groups) {
      const float4* xp = reinterpret_cast<const float4*>(a.x + row * a.E + grp * 8);
      const float4* yp = reinterpret_cast<const float4*>(a.y + row * a.E + grp * 8);
      const float4 x0 = xp[0], x1 = xp[1], y0 = __ldg(yp), y1 = __ldg(yp + 1);
      v[g][0] = x0.x + y0.x; v[g][1] = x0.y + y0.y; v[g][2] = x0.z + y0.z; v[g][3] = x0.w + y0.w;
      v[g][4] = x1.x + y1.x; v[g][5] = x1.y + y1.y; v[g][6] = x1.z + y1.z; v[g][7] = x1.w + y1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += v[g][i];
    }
  }
  const float mean = warp_sum_tc(sum) / (float)a.E;
  float sq = 0.f;
#pragma unroll
  for (int g = 0; g < kTrfLnMaxGroups; ++g)
    if (lane + 32 * g < groups) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[g][i] - mean; sq = fmaf(d, d, sq); }
    }
  const float rstd = 1.0f / sqrtf(warp_sum_tc(sq) / (float)a.E + a.eps);
#pragma unroll
  for (int g = 0; g < kTrfLnMaxGroups; ++g) {
    const int grp = lane + 32 * g;
    if (grp < groups) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o[i] = fmaf((v[g][i] - mean) * rstd, __ldg(a.gamma + grp * 8 + i), __ldg(a.beta + grp * 8 + i));
      float4* dst = reinterpret_cast<float4*>(a.x + row * a.E + grp * 8);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
      img_store8(a.x_img, a.plane, a.rpad, (int)row, grp * 8, o);
    }
  }
}

}  // namespace bcnf
