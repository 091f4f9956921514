// Training step of the coupling stack: everything around the hidden-layer GEMMs, one kernel per coupling
// block and direction instead of ~25 element-wise / sliver-GEMM launches.
//
// Reference path (Trainer._train_batch -> CondRealNVP_v2.forward in train mode -> autograd):
//   ConditionalNestedNeuralNetwork.forward cnf.py:98-107, ConditionalAffineCouplingLayer.forward cnf.py:165-196,
//   ActNorm.forward cnf.py:348-351, OrthonormalTransformation.forward cnf.py:333-336.
//
//   train_pre_kernel       first Linear of a conditioner on the network's own half: pre1 = y_src W1a^T + P
//                          (P = h W1h^T + b1, one GEMM per network off the dependency chain), a1 = drop(gelu(pre1))
//   train_post_kernel      last Linear (N = 2 dout <= 64) + chunk + tanh + exp + affine update + log-det row sum,
//                          then the ActNorm / orthonormal mixing that follow the coupling in model.layers
//   train_post_bwd_kernel  their backward: gradients through the glue ops (ActNorm parameter gradients reduced
//                          over the batch), the affine update, tanh, and the data gradient of the last Linear with
//                          gelu' * dropout mask of the last hidden layer applied
//   train_pre_bwd_kernel   data gradient of the first Linear w.r.t. the network's own half of y
// One warp owns one row; D <= 64.
#pragma once
#include "train_ops.cuh"
#include "train_tc.cuh"

namespace bcnf {

constexpr int kGlueWarps = 8;
constexpr int kGlueMaxOps = 4;
enum GlueOpType : int { GLUE_ORTHO = 0, GLUE_ACTNORM = 1 };

struct GlueOp {
  int type;
  const float* p0;   // ortho: Q (D, D) row-major;  actnorm: scale (D)
  const float* p1;   // actnorm: bias (D)
  float* save;       // actnorm: its input x (B, D): written by the forward, read by the backward
  float* g0;         // backward, actnorm: d scale (D), accumulated atomically (caller zeroes)
  float* g1;         // backward, actnorm: d bias (D)
};

struct TrainPreArgs {
  const float* y; long long y_pitch;
  int B, D, src0, din;
  const float* W1; long long w1_pitch;
  const float* P; long long p_pitch;
  int H;
  float* pre; float* act; long long pitch;
  unsigned long long seed; unsigned int layer_uid; float p_drop; const unsigned long long* seed_ptr;
  unsigned char* act_img; long long img_plane; int img_rpad;     // optional image of act (train_tc.cuh)
};

struct TrainPostArgs {
  const float* a; long long a_pitch;      // null: glue ops only
  const float* Wout; const float* bout;
  int B, D, H, dst0, dout;
  const float* y_in; float* y_out;
  float* ld;
  float* ls_save; float* ydst_save;
  int n_ops; GlueOp ops[kGlueMaxOps];
};

struct TrainPostBwdArgs {
  const float* dz_in; float* dz_out;
  const float* dld;
  int B, D, H, dst0, dout;
  const float* ls_save; const float* ydst_save;
  const float* Wout;                      // null: glue ops only
  const float* pre; long long pitch;
  float* d_o;
  float* d_pre;
  unsigned long long seed; unsigned int layer_uid; float p_drop; const unsigned long long* seed_ptr;
  int n_ops; GlueOp ops[kGlueMaxOps];
  unsigned char* dpre_img; long long img_plane; int img_rpad;    // optional image of d_pre
};

struct TrainPreBwdArgs {
  const float* d_pre; long long pitch;
  const float* W1; long long w1_pitch;
  int B, D, H, src0, din;
  float* dz;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- first Linear (own half) + P + GELU + dropout: thread = one hidden unit j, block = kPreRows rows --------------
constexpr int kPreRows = 8;
__global__ void __launch_bounds__(128) train_pre_kernel(const TrainPreArgs a) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int i0 = blockIdx.y * kPreRows;
  __shared__ float ys[kPreRows][32];
  for (int e = threadIdx.x; e < kPreRows * 32; e += 128) {
    const int r = e >> 5, k = e & 31, i = i0 + r;
    ys[r][k] = (i < a.B && k < a.din) ? a.y[i * a.y_pitch + a.src0 + k] : 0.f;
  }
  __syncthreads();
  if (j >= a.H) return;
  float w[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) w[k] = k < a.din ? __ldg(a.W1 + j * a.w1_pitch + k) : 0.f;
  const unsigned long long seed = a.seed_ptr ? (a.seed ^ *a.seed_ptr) : a.seed;
  const float keep_scale = a.p_drop > 0.f ? 1.0f / (1.0f - a.p_drop) : 1.0f;
  for (int r = 0; r < kPreRows; ++r) {
    const int i = i0 + r;
    if (i >= a.B) break;
    float s = __ldg(a.P + i * a.p_pitch + j);
#pragma unroll
    for (int k = 0; k < 32; ++k) s = fmaf(ys[r][k], w[k], s);
    a.pre[i * a.pitch + j] = s;
    float v = gelu_erf(s);
    if (a.p_drop > 0.f)
      v = dropout_uniform(seed, a.layer_uid, (unsigned long long)i * (unsigned)a.H + (unsigned)j) >= a.p_drop ? v * keep_scale : 0.f;
    a.act[i * a.pitch + j] = v;
    if (a.act_img) img_store1(a.act_img, a.img_plane, a.img_rpad, i, j, v);
  }
}

// ---- glue ops on a row held in shared memory (one warp) -------------------------------------------------------
__device__ __forceinline__ void glue_forward(const GlueOp& op, float* yr, float* tmp, int D, long long i, float& ld, int lane) {
  if (op.type == GLUE_ORTHO) {
    for (int j = lane; j < D; j += 32) {
      float s = 0.f;
      for (int k = 0; k < D; ++k) s = fmaf(yr[k], __ldg(op.p0 + k * D + j), s);   // y @ Q, cnf.py:335
      tmp[j] = s;
    }
    __syncwarp();
    for (int j = lane; j < D; j += 32) yr[j] = tmp[j];
    __syncwarp();
  } else {
    float lsum = 0.f;
    for (int j = lane; j < D; j += 32) {
      const float s = __ldg(op.p0 + j), x = yr[j];
      if (op.save) op.save[i * D + j] = x;
      yr[j] = fmaf(s, x, __ldg(op.p1 + j));                                        // cnf.py:349
      lsum += logf(fabsf(s));
    }
    ld += warp_sum(lsum);                                                          // cnf.py:350
    __syncwarp();
  }
}

constexpr int kGlueKT = 128;     // K tile of the last Linear staged in shared memory

// cooperative load of Wout[m, k0 .. k0+KT) for m < no into w_s[m][kk] (zero past H)
__device__ __forceinline__ void load_wout_tile(float (*w_s)[kGlueKT], const float* __restrict__ Wout, int no, int H, int k0) {
  // thread = (column kk, row parity): all loads of a thread are independent and issued back to back (8 per batch)
  const int kk = threadIdx.x & (kGlueKT - 1), m0 = threadIdx.x / kGlueKT;
  constexpr int MS = 32 * kGlueWarps / kGlueKT;          // rows covered per pass (2)
  const bool k_ok = k0 + kk < H;
  const float* src = Wout + k0 + kk;
#pragma unroll
  for (int mb = 0; mb < 64; mb += 8 * MS) {
    if (mb >= no) break;
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int m = mb + u * MS + m0;
      v[u] = (k_ok && m < no) ? __ldg(src + (long long)m * H) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int m = mb + u * MS + m0;
      if (m < no) w_s[m][kk] = v[u];
    }
  }
}

__global__ void __launch_bounds__(32 * kGlueWarps) train_post_kernel(const TrainPostArgs a) {
  __shared__ float y_s[kGlueWarps][64], t_s[kGlueWarps][64];
  __shared__ float w_s[64][kGlueKT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * kGlueWarps + warp;
  const bool valid = i < a.B;
  float* yr = y_s[warp];
  float* tmp = t_s[warp];
  if (valid) for (int j = lane; j < a.D; j += 32) yr[j] = a.y_in[i * a.D + j];
  __syncwarp();
  float ld = 0.f;
  if (a.a) {
    // o = a Wout^T + bout: Wout staged tile by tile in shared memory, lanes stride over k, butterfly reduction
    const int no = 2 * a.dout;
    float acc[64];
#pragma unroll
    for (int m = 0; m < 64; ++m) acc[m] = 0.f;
    const float* ar = a.a + (valid ? i : 0) * a.a_pitch;
    for (int k0 = 0; k0 < a.H; k0 += kGlueKT) {
      __syncthreads();
      load_wout_tile(w_s, a.Wout, no, a.H, k0);
      float x[kGlueKT / 32];
#pragma unroll
      for (int t = 0; t < kGlueKT / 32; ++t) { const int k = k0 + lane + 32 * t; x[t] = (valid && k < a.H) ? ar[k] : 0.f; }
      __syncthreads();
#pragma unroll
      for (int t = 0; t < kGlueKT / 32; ++t)
#pragma unroll
        for (int m = 0; m < 64; ++m)
          if (m < no) acc[m] = fmaf(x[t], w_s[m][lane + 32 * t], acc[m]);
    }
#pragma unroll
    for (int m = 0; m < 64; ++m)
      if (m < no) acc[m] = warp_sum(acc[m]);
    // lane m < dout owns element m of the transformed half: t = o[m], log s = tanh(o[dout + m])  (cnf.py:104-107)
    float t = 0.f, s_raw = 0.f;
#pragma unroll
    for (int m = 0; m < 64; ++m) {
      if (m < no) {
        if (m == lane) t = acc[m];
        if (m == lane + a.dout) s_raw = acc[m];
      }
    }
    float ls = 0.f;
    if (valid && lane < a.dout) {
      t += __ldg(a.bout + lane);
      ls = tanhf(s_raw + __ldg(a.bout + a.dout + lane));
      const float yd = yr[a.dst0 + lane];
      a.ls_save[i * a.dout + lane] = ls;
      a.ydst_save[i * a.dout + lane] = yd;
      yr[a.dst0 + lane] = fmaf(expf(ls), yd, t);                                   // cnf.py:179 / :184
    }
    ld += warp_sum(ls);                                                            // cnf.py:190 / :193
    __syncwarp();
  }
  if (!valid) return;
  for (int o = 0; o < a.n_ops; ++o) glue_forward(a.ops[o], yr, tmp, a.D, i, ld, lane);
  for (int j = lane; j < a.D; j += 32) a.y_out[i * a.D + j] = yr[j];
  if (lane == 0) a.ld[i] += ld;
}

__global__ void __launch_bounds__(32 * kGlueWarps) train_post_bwd_kernel(const TrainPostBwdArgs a) {
  __shared__ float d_s[kGlueWarps][64], t_s[kGlueWarps][64];
  __shared__ float red_s[2][kGlueMaxOps][64];       // block-level partial sums of the ActNorm parameter gradients
  __shared__ float w_s[64][kGlueKT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * kGlueWarps + warp;
  const bool valid = i < a.B;
  for (int e = threadIdx.x; e < 2 * kGlueMaxOps * 64; e += blockDim.x) (&red_s[0][0][0])[e] = 0.f;
  __syncthreads();
  float* dr = d_s[warp];
  float* tmp = t_s[warp];
  const float dld = valid ? a.dld[i] : 0.f;
  if (valid) {
    for (int j = lane; j < a.D; j += 32) dr[j] = a.dz_in[i * a.D + j];
    __syncwarp();
    for (int o = a.n_ops - 1; o >= 0; --o) {
      const GlueOp& op = a.ops[o];
      if (op.type == GLUE_ORTHO) {
        for (int k = lane; k < a.D; k += 32) {
          float s = 0.f;
          for (int j = 0; j < a.D; ++j) s = fmaf(dr[j], __ldg(op.p0 + k * a.D + j), s);   // dz @ Q^T
          tmp[k] = s;
        }
        __syncwarp();
        for (int k = lane; k < a.D; k += 32) dr[k] = tmp[k];
        __syncwarp();
      } else {
        for (int j = lane; j < a.D; j += 32) {
          const float s = __ldg(op.p0 + j), d = dr[j];
          // d/ds [s x + b] and d/ds sum_j log|s_j| (the log-det term reaches every row's loss)
          atomicAdd(&red_s[0][o][j], fmaf(d, op.save[i * a.D + j], dld / s));
          atomicAdd(&red_s[1][o][j], d);
          dr[j] = d * s;
        }
        __syncwarp();
      }
    }
  }
  if (a.Wout) {
    const int no = 2 * a.dout;
    // affine update and tanh backward; lane m < dout owns element m
    float dt = 0.f, dso = 0.f;
    if (valid && lane < a.dout) {
      const float ls = a.ls_save[i * a.dout + lane], yd = a.ydst_save[i * a.dout + lane];
      const float e = expf(ls), dn = dr[a.dst0 + lane];
      dt = dn;
      dso = fmaf(dn * yd, e, dld) * (1.0f - ls * ls);
      dr[a.dst0 + lane] = dn * e;
      a.d_o[i * no + lane] = dt;
      a.d_o[i * no + a.dout + lane] = dso;
    }
    // d a = d_o Wout (Wout staged tile by tile), then gelu'(pre) * dropout mask of the last hidden layer
    float dov[64];
#pragma unroll
    for (int m = 0; m < 64; ++m) {
      dov[m] = 0.f;
      if (m < no) dov[m] = m < a.dout ? __shfl_sync(0xffffffffu, dt, m & 31) : __shfl_sync(0xffffffffu, dso, (m - a.dout) & 31);
    }
    const unsigned long long seed = a.seed_ptr ? (a.seed ^ *a.seed_ptr) : a.seed;
    const float keep_scale = a.p_drop > 0.f ? 1.0f / (1.0f - a.p_drop) : 1.0f;
    for (int k0 = 0; k0 < a.H; k0 += kGlueKT) {
      __syncthreads();
      load_wout_tile(w_s, a.Wout, no, a.H, k0);
      __syncthreads();
      if (!valid) continue;
#pragma unroll
      for (int t = 0; t < kGlueKT / 32; ++t) {
        const int k = k0 + lane + 32 * t;
        if (k >= a.H) break;
        float s = 0.f;
#pragma unroll
        for (int m = 0; m < 64; ++m)
          if (m < no) s = fmaf(dov[m], w_s[m][lane + 32 * t], s);
        s *= dgelu_erf(a.pre[i * a.pitch + k]);
        if (a.p_drop > 0.f)
          s = dropout_uniform(seed, a.layer_uid, (unsigned long long)i * (unsigned)a.H + (unsigned)k) >= a.p_drop ? s * keep_scale : 0.f;
        a.d_pre[i * a.pitch + k] = s;
        if (a.dpre_img) img_store1(a.dpre_img, a.img_plane, a.img_rpad, (int)i, k, s);
      }
    }
  }
  if (valid) {
    __syncwarp();
    for (int j = lane; j < a.D; j += 32) a.dz_out[i * a.D + j] = dr[j];
  }
  __syncthreads();
  for (int o = 0; o < a.n_ops; ++o) {
    if (a.ops[o].type != GLUE_ACTNORM) continue;
    for (int j = threadIdx.x; j < a.D; j += blockDim.x) {
      atomicAdd(a.ops[o].g0 + j, red_s[0][o][j]);
      atomicAdd(a.ops[o].g1 + j, red_s[1][o][j]);
    }
  }
}

constexpr int kGlueJT = 128;     // rows of W1[:, :din] staged per tile in train_pre_bwd_kernel
__global__ void __launch_bounds__(32 * kGlueWarps) train_pre_bwd_kernel(const TrainPreBwdArgs a) {
  __shared__ float w_s[kGlueJT][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * kGlueWarps + warp;
  const bool valid = i < a.B;
  float acc[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) acc[k] = 0.f;
  for (int j0 = 0; j0 < a.H; j0 += kGlueJT) {
    __syncthreads();
    {
      // thread = (row jj of the tile, column parity): its loads are independent, issued back to back
      const int jj = threadIdx.x & (kGlueJT - 1), kq = threadIdx.x / kGlueJT;    // kq in {0, 1}
      const bool j_ok = j0 + jj < a.H;
      const float* src = a.W1 + (long long)(j0 + jj) * a.w1_pitch;
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) { const int k = 2 * u + kq; v[u] = (j_ok && k < a.din) ? __ldg(src + k) : 0.f; }
#pragma unroll
      for (int u = 0; u < 16; ++u) w_s[jj][2 * u + kq] = v[u];
    }
    __syncthreads();
    if (!valid) continue;
#pragma unroll
    for (int t = 0; t < kGlueJT / 32; ++t) {
      const int jj = lane + 32 * t, j = j0 + jj;
      const float d = j < a.H ? a.d_pre[i * a.pitch + j] : 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < a.din) acc[k] = fmaf(d, w_s[jj][k], acc[k]);
    }
  }
  if (!valid) return;
  float mine = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (k < a.din) {
      const float s = warp_sum(acc[k]);
      if (k == lane) mine = s;
    }
  }
  if (lane < a.din) a.dz[i * a.D + a.src0 + lane] += mine;
}

}  // namespace bcnf
