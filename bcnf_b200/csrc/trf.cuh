// Transformer condition encoder (reference src/bcnf/models/feature_network.py:183-307): everything that is not a Linear.
// The Linears (q/k/v merged, fc_out, the two FFN layers, the output layer) run on the CTA-pair GEMM on operand images
// (gemm_img2.cuh); the kernels here produce those images.  Inference (feature_tc.py) and the Trainer's step (trf_train.py):
//
//   trf_embed_kernel    tokens (B, T, F) -> x = (tokens . Wf^T + bf) * mask (+ positional table)   : fp32 x + image of x
//   trf_attn_kernel     q | k | v (rows, 3E) fp32 -> softmax(q k^T / sqrt(hd)) v per instance, head : image of the context
//                                                                                                     (+ fp32 copy)
//   trf_add_ln_kernel   x_out = LayerNorm(x + mask * y) * gamma + beta  (post-norm block, :255-259) : fp32 x_out + image
//                                                                                   (+ LayerNorm input and statistics)
//   trf_gelu_kernel     a = gelu(u)                                              (training)        : fp32 a + image of a
//   trf_attn_bwd_kernel d ctx -> d q | d k | d v, probabilities recomputed       (training)        : fp32
//   trf_ln_param_grad_kernel  d gamma, d beta of nn.LayerNorm                    (training)        : fp32, atomics
//
// `mask` = dropout multipliers (training; null in inference).  All arithmetic is fp32; an image is the bf16 hi / lo split
// of the fp32 value (img_store.cuh: img_store8), so the GEMM that reads it sees the value to 2^-17.  In inference these are
// HBM-bound kernels: every element is read once and written once (+ its image).
#pragma once
#include "common.cuh"
#include "img_store.cuh"

namespace bcnf {

struct TrfEmbedArgs {
  const float* tokens;   // (rows, F), rows = B * T
  const float* Wf;       // (E, F)  nn.Linear weight
  const float* bf;       // (E)
  const float* pos;      // (T, E) positional table or null (feature_network.py:291-301)
  const float* mask;     // (rows, E) dropout multipliers (0 or 1 / (1 - p)) applied to the Linear's output (:288), or null
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
  if (a.mask) {
    const float* mk = a.mask + row * a.E + n0;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= __ldg(mk + i);
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
  float* ctx;            // optional fp32 copy of the same (training: the weight gradient of fc_out reads it)
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
          if (a.ctx) {
            float4* cp = reinterpret_cast<float4*>(a.ctx + (long long)row * E + h * HD + 8 * g);
            cp[0] = make_float4(w[0], w[1], w[2], w[3]);
            cp[1] = make_float4(w[4], w[5], w[6], w[7]);
          }
        }
      }
    }
    __syncthreads();
  }
}

struct TrfAddLnArgs {
  const float* x;        // (rows, E): the residual input
  const float* y;        // (rows, E): the sublayer output (bias included)
  const float* mask;     // (rows, E) dropout multipliers applied to y (self.dropout, :255-259), or null
  const float* gamma; const float* beta;
  float* s;              // optional (rows, E): x + mask * y, the LayerNorm input (saved for the backward pass)
  float* mean; float* rstd;   // optional (rows): LayerNorm statistics (saved for the backward pass)
  float* x_out;          // (rows, E); may alias x
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
      const long long o = row * a.E + grp * 8;
      const float4* xp = reinterpret_cast<const float4*>(a.x + o);
      const float4* yp = reinterpret_cast<const float4*>(a.y + o);
      const float4 x0 = xp[0], x1 = xp[1];
      float4 y0 = __ldg(yp), y1 = __ldg(yp + 1);
      if (a.mask) {
        const float4 m0 = __ldg(reinterpret_cast<const float4*>(a.mask + o)), m1 = __ldg(reinterpret_cast<const float4*>(a.mask + o) + 1);
        y0.x *= m0.x; y0.y *= m0.y; y0.z *= m0.z; y0.w *= m0.w;
        y1.x *= m1.x; y1.y *= m1.y; y1.z *= m1.z; y1.w *= m1.w;
      }
      v[g][0] = x0.x + y0.x; v[g][1] = x0.y + y0.y; v[g][2] = x0.z + y0.z; v[g][3] = x0.w + y0.w;
      v[g][4] = x1.x + y1.x; v[g][5] = x1.y + y1.y; v[g][6] = x1.z + y1.z; v[g][7] = x1.w + y1.w;
      if (a.s) {
        float4* sp = reinterpret_cast<float4*>(a.s + o);
        sp[0] = make_float4(v[g][0], v[g][1], v[g][2], v[g][3]);
        sp[1] = make_float4(v[g][4], v[g][5], v[g][6], v[g][7]);
      }
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
  if (lane == 0 && a.mean) { a.mean[row] = mean; a.rstd[row] = rstd; }
#pragma unroll
  for (int g = 0; g < kTrfLnMaxGroups; ++g) {
    const int grp = lane + 32 * g;
    if (grp < groups) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o[i] = fmaf((v[g][i] - mean) * rstd, __ldg(a.gamma + grp * 8 + i), __ldg(a.beta + grp * 8 + i));
      float4* dst = reinterpret_cast<float4*>(a.x_out + row * a.E + grp * 8);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
      img_store8(a.x_img, a.plane, a.rpad, (int)row, grp * 8, o);
    }
  }
}

// ---- training only ---------------------------------------------------------------------------------------------------
// a = gelu(u) (nn.GELU, exact erf) as fp32 (the weight gradient of the second FFN Linear reads it) and as the operand image
// of that Linear.  u, the pre-activation, stays in memory for the backward pass.  One thread per 8 columns.
struct TrfGeluArgs {
  const float* u; float* a;        // (rows, N); a optional
  unsigned char* a_img; long long plane; int rpad;
  long long rows; int N;
};
__global__ void __launch_bounds__(256) trf_gelu_kernel(const TrfGeluArgs g) {
  const int groups = g.N >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.rows * groups) return;
  const long long row = idx / groups;
  const int n0 = (int)(idx - row * groups) << 3;
  const float4* up = reinterpret_cast<const float4*>(g.u + row * g.N + n0);
  const float4 u0 = __ldg(up), u1 = __ldg(up + 1);
  float v[8] = {gelu_erf(u0.x), gelu_erf(u0.y), gelu_erf(u0.z), gelu_erf(u0.w),
                gelu_erf(u1.x), gelu_erf(u1.y), gelu_erf(u1.z), gelu_erf(u1.w)};
  if (g.a) {
    float4* ap = reinterpret_cast<float4*>(g.a + row * g.N + n0);
    ap[0] = make_float4(v[0], v[1], v[2], v[3]);
    ap[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  img_store8(g.a_img, g.plane, g.rpad, (int)row, n0, v);
}

// d gamma[c] += sum_rows g * (s - mean) * rstd, d beta[c] += sum_rows g  of nn.LayerNorm: column reductions over all token
// rows that only the optimizer reads (ATen's GammaBetaBackward kernel takes 64 us for 7680 x 128).  A block reduces a slab
// of rows per column in registers and adds its partial sums atomically; the caller zeroes the outputs.
struct TrfLnParamGradArgs {
  const float* g; const float* s; const float* mean; const float* rstd;
  float* dgamma; float* dbeta;
  long long rows; int E; int slab;
};
__global__ void __launch_bounds__(256) trf_ln_param_grad_kernel(const TrfLnParamGradArgs a) {
  const long long r0 = (long long)blockIdx.x * a.slab;
  const long long r1 = r0 + a.slab < a.rows ? r0 + a.slab : a.rows;
  for (int c = threadIdx.x; c < a.E; c += blockDim.x) {
    float dg = 0.f, db = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const float gv = __ldg(a.g + r * a.E + c);
      const float xh = (__ldg(a.s + r * a.E + c) - __ldg(a.mean + r)) * __ldg(a.rstd + r);
      dg = fmaf(gv, xh, dg);
      db += gv;
    }
    atomicAdd(a.dgamma + c, dg);
    atomicAdd(a.dbeta + c, db);
  }
}

// Backward of the attention core (feature_network.py:207-226): d ctx -> d q | d k | d v, one CTA per instance, a warp per
// head.  Phase 1, lane = query i: recompute p_i. = softmax(q_i k^T / sqrt(hd)), dP_ij = dctx_i . v_j,
// dS_ij = p_ij (dP_ij - sum_j p_ij dP_ij), dq_i = sum_j dS_ij k_j / sqrt(hd); p and dS of the head go to the warp's own
// shared-memory tiles.  Phase 2, lane = key j: dk_j = sum_i dS_ij q_i / sqrt(hd), dv_j = sum_i p_ij dctx_i.  T <= 32.
struct TrfAttnBwdArgs {
  const float* qkv;      // (rows, 3E) saved by the forward pass
  const float* dctx;     // (rows, E)
  float* dqkv;           // (rows, 3E)
  long long n_inst; int T, E, heads;
  int hpg;               // heads per CTA (blockIdx.y selects the group; blockDim.x = 32 * hpg): at batch 256 two half-size CTAs
                         // per instance fill the SMs in one wave where one full-size CTA per instance needs two
  float scale;
};

template <int HD>
__global__ void __launch_bounds__(kTrfAttnThreads, 1) trf_attn_bwd_kernel(const TrfAttnBwdArgs a) {
  extern __shared__ __align__(16) float smem_attn[];
  const int T = a.T, E = a.E, ldp = T + 1;
  const int W = a.hpg * HD;             // columns of this CTA's heads
  const int c0 = blockIdx.y * W;        // first of them
  const int nthr = blockDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* qs = smem_attn;                // [T][W]
  float* ks = qs + T * W;
  float* vs = ks + T * W;
  float* ds = vs + T * W;               // d ctx
  float* pw = ds + T * W + warp * 2 * T * ldp;   // this warp's p [T][T+1]
  float* sw = pw + T * ldp;                      // this warp's dS [T][T+1]
  for (long long inst = blockIdx.x; inst < a.n_inst; inst += gridDim.x) {
    const float* src = a.qkv + inst * T * 3 * E;
    const float* dsrc = a.dctx + inst * T * E;
    float* dout = a.dqkv + inst * T * 3 * E + c0;
    const int w4 = W >> 2;
    for (int i = tid; i < T * 4 * w4; i += nthr) {
      const int t = i / (4 * w4), r = i - t * 4 * w4, part = r / w4, c4 = r - part * w4;
      const float* from = part < 3 ? src + (long long)t * 3 * E + part * E + c0 : dsrc + (long long)t * E + c0;
      const float4 w = __ldg(reinterpret_cast<const float4*>(from) + c4);
      float* dst = (part == 0 ? qs : part == 1 ? ks : part == 2 ? vs : ds) + t * W + (c4 << 2);
      *reinterpret_cast<float4*>(dst) = w;
    }
    __syncthreads();
    {
      const int h = warp;               // head inside the group: columns [h * HD, (h + 1) * HD) of the staged tiles
      const int E_ = W;                 // row pitch of the staged tiles
      const int i = lane;
      if (i < T) {
        float q[HD], dc[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) { q[c] = qs[i * E_ + h * HD + c] * a.scale; dc[c] = ds[i * E_ + h * HD + c]; }
        float sc[32], dp[32];
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < T) {
            const float4* kp = reinterpret_cast<const float4*>(ks + j * E_ + h * HD);
            const float4* vp = reinterpret_cast<const float4*>(vs + j * E_ + h * HD);
            float acc = 0.f, acd = 0.f;
#pragma unroll
            for (int c = 0; c < HD / 4; ++c) {
              const float4 w = kp[c], z = vp[c];
              acc = fmaf(q[4 * c], w.x, acc); acc = fmaf(q[4 * c + 1], w.y, acc);
              acc = fmaf(q[4 * c + 2], w.z, acc); acc = fmaf(q[4 * c + 3], w.w, acc);
              acd = fmaf(dc[4 * c], z.x, acd); acd = fmaf(dc[4 * c + 1], z.y, acd);
              acd = fmaf(dc[4 * c + 2], z.z, acd); acd = fmaf(dc[4 * c + 3], z.w, acd);
            }
            sc[j] = acc; dp[j] = acd;
            m = fmaxf(m, acc);
          }
        }
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < T) { sc[j] = expf(sc[j] - m); l += sc[j]; }
        const float inv = 1.0f / l;
        float delta = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < T) { sc[j] *= inv; delta = fmaf(sc[j], dp[j], delta); }
        float dq[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) dq[c] = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < T) {
            const float dsij = sc[j] * (dp[j] - delta);
            pw[i * ldp + j] = sc[j];
            sw[i * ldp + j] = dsij;
            const float4* kp = reinterpret_cast<const float4*>(ks + j * E_ + h * HD);
#pragma unroll
            for (int c = 0; c < HD / 4; ++c) {
              const float4 w = kp[c];
              dq[4 * c] = fmaf(dsij, w.x, dq[4 * c]); dq[4 * c + 1] = fmaf(dsij, w.y, dq[4 * c + 1]);
              dq[4 * c + 2] = fmaf(dsij, w.z, dq[4 * c + 2]); dq[4 * c + 3] = fmaf(dsij, w.w, dq[4 * c + 3]);
            }
          }
        }
        float4* qo = reinterpret_cast<float4*>(dout + (long long)i * 3 * E + h * HD);
#pragma unroll
        for (int c = 0; c < HD / 4; ++c)
          qo[c] = make_float4(dq[4 * c] * a.scale, dq[4 * c + 1] * a.scale, dq[4 * c + 2] * a.scale, dq[4 * c + 3] * a.scale);
      }
      __syncwarp();
      const int j = lane;
      if (j < T) {
        float dk[HD], dv[HD];
#pragma unroll
        for (int c = 0; c < HD; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
        for (int i2 = 0; i2 < T; ++i2) {
          const float p = pw[i2 * ldp + j], dsij = sw[i2 * ldp + j];
          const float4* qp = reinterpret_cast<const float4*>(qs + i2 * E_ + h * HD);
          const float4* dp4 = reinterpret_cast<const float4*>(ds + i2 * E_ + h * HD);
#pragma unroll
          for (int c = 0; c < HD / 4; ++c) {
            const float4 w = qp[c], z = dp4[c];
            dk[4 * c] = fmaf(dsij, w.x, dk[4 * c]); dk[4 * c + 1] = fmaf(dsij, w.y, dk[4 * c + 1]);
            dk[4 * c + 2] = fmaf(dsij, w.z, dk[4 * c + 2]); dk[4 * c + 3] = fmaf(dsij, w.w, dk[4 * c + 3]);
            dv[4 * c] = fmaf(p, z.x, dv[4 * c]); dv[4 * c + 1] = fmaf(p, z.y, dv[4 * c + 1]);
            dv[4 * c + 2] = fmaf(p, z.z, dv[4 * c + 2]); dv[4 * c + 3] = fmaf(p, z.w, dv[4 * c + 3]);
          }
        }
        float4* ko = reinterpret_cast<float4*>(dout + (long long)j * 3 * E + E + h * HD);
        float4* vo = reinterpret_cast<float4*>(dout + (long long)j * 3 * E + 2 * E + h * HD);
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) {
          ko[c] = make_float4(dk[4 * c] * a.scale, dk[4 * c + 1] * a.scale, dk[4 * c + 2] * a.scale, dk[4 * c + 3] * a.scale);
          vo[c] = make_float4(dv[4 * c], dv[4 * c + 1], dv[4 * c + 2], dv[4 * c + 3]);
        }
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

}  // namespace bcnf
