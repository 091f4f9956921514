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
  float pv[kPreRows];
#pragma unroll
  for (int r = 0; r < kPreRows; ++r) pv[r] = i0 + r < a.B ? __ldg(a.P + (i0 + r) * a.p_pitch + j) : 0.f;
#pragma unroll
  for (int r = 0; r < kPreRows; ++r) {
    const int i = i0 + r;
    if (i >= a.B) break;
    float s = pv[r];
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
#pragma unroll 8
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

// ---- asynchronous staging of small parameter blocks into shared memory: 4-byte cp.async (no alignment
// requirement beyond the element), every copy of a thread in flight at once, zero fill where !valid ----
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Wout[m, k0 .. k0+kt) for m < no  ->  w_s[m*kt + kk]   (whole matrix when kt >= H: the usual case)
__device__ __forceinline__ void stage_wout(float* w_s, const float* __restrict__ Wout, int no, int H, int k0, int kt) {
  for (int e = threadIdx.x; e < no * kt; e += 32 * kGlueWarps) {
    const int m = e / kt, kk = e - m * kt;
    const bool ok = k0 + kk < H;
    cp_async4(w_s + e, Wout + (long long)m * H + (ok ? k0 + kk : 0), ok);
  }
}

extern __shared__ float glue_dyn_smem[];

// stage one row of a (B, pitch) matrix, columns [k0, k0 + kt), into shared memory (zero past H)
__device__ __forceinline__ void stage_row(float* x_s, const float* __restrict__ row, int H, int k0, int kt, int lane) {
  for (int kk = lane; kk < kt; kk += 32) {
    const bool ok = k0 + kk < H;
    cp_async4(x_s + kk, row + (ok ? k0 + kk : 0), ok);
  }
}

// Dynamic shared memory of the post kernels: w_s[no][kt + 1] (odd pitch: lanes = output rows read conflict-free),
// then one kt-float row buffer per warp.
__global__ void __launch_bounds__(32 * kGlueWarps) train_post_kernel(const TrainPostArgs a, const int kt) {
  __shared__ float y_s[kGlueWarps][64], t_s[kGlueWarps][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * kGlueWarps + warp;
  const bool valid = i < a.B;
  const int no = 2 * a.dout, wp = kt + 1;
  float* w_s = glue_dyn_smem;
  float* x_s = glue_dyn_smem + no * wp + warp * kt;
  float* yr = y_s[warp];
  float* tmp = t_s[warp];
  if (valid) for (int j = lane; j < a.D; j += 32) yr[j] = a.y_in[i * a.D + j];
  __syncwarp();
  float ld = 0.f;
  if (a.a) {
    // o = a Wout^T + bout.  Lane m owns outputs m and m + 32: it walks k with x[k] broadcast from shared memory.
    const float* ar = a.a + (valid ? i : 0) * a.a_pitch;
    const int m0 = lane < no ? lane : 0, m1 = lane + 32 < no ? lane + 32 : 0;
    float o0a = 0.f, o0b = 0.f, o1a = 0.f, o1b = 0.f;
    for (int k0 = 0; k0 < a.H; k0 += kt) {
      if (k0 > 0) __syncthreads();
      for (int e = threadIdx.x; e < no * kt; e += 32 * kGlueWarps) {
        const int m = e / kt, kk = e - m * kt;
        const bool ok = k0 + kk < a.H;
        cp_async4(w_s + m * wp + kk, a.Wout + (long long)m * a.H + (ok ? k0 + kk : 0), ok);
      }
      stage_row(x_s, ar, valid ? a.H : 0, k0, kt, lane);
      cp_async_wait_all();
      __syncthreads();
      const float* w0 = w_s + m0 * wp;
      const float* w1 = w_s + m1 * wp;
      if (no <= 32) {
#pragma unroll 4
        for (int kk = 0; kk < kt; kk += 2) {
          o0a = fmaf(x_s[kk], w0[kk], o0a);
          o0b = fmaf(x_s[kk + 1], w0[kk + 1], o0b);
        }
      } else {
#pragma unroll 4
        for (int kk = 0; kk < kt; kk += 2) {
          const float xa = x_s[kk], xb = x_s[kk + 1];
          o0a = fmaf(xa, w0[kk], o0a); o0b = fmaf(xb, w0[kk + 1], o0b);
          o1a = fmaf(xa, w1[kk], o1a); o1b = fmaf(xb, w1[kk + 1], o1b);
        }
      }
    }
    const float o0 = o0a + o0b, o1 = o1a + o1b;      // outputs lane and lane + 32
    // lane m < dout owns element m of the transformed half: t = o[m], log s = tanh(o[dout + m])  (cnf.py:104-107)
    const int si = a.dout + lane;                      // index of this lane's s output
    const float s_lo = __shfl_sync(0xffffffffu, o0, si & 31), s_hi = __shfl_sync(0xffffffffu, o1, si & 31);
    float ls = 0.f;
    if (valid && lane < a.dout) {
      const float t = o0 + __ldg(a.bout + lane);
      ls = tanhf((si < 32 ? s_lo : s_hi) + __ldg(a.bout + si));
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

// Dynamic shared memory: w_s[no][kt] (lanes = consecutive k: conflict-free)
__global__ void __launch_bounds__(32 * kGlueWarps) train_post_bwd_kernel(const TrainPostBwdArgs a, const int kt) {
  __shared__ float d_s[kGlueWarps][64], t_s[kGlueWarps][64], do_s[kGlueWarps][64];
  __shared__ float red_s[2][kGlueMaxOps][64];       // block-level partial sums of the ActNorm parameter gradients
  float* w_s = glue_dyn_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * kGlueWarps + warp;
  const bool valid = i < a.B;
  const int no = 2 * a.dout;
  if (a.Wout) stage_wout(w_s, a.Wout, no, a.H, 0, kt);
  for (int e = threadIdx.x; e < 2 * kGlueMaxOps * 64; e += blockDim.x) (&red_s[0][0][0])[e] = 0.f;
  __syncthreads();
  float* dr = d_s[warp];
  float* tmp = t_s[warp];
  float* dov = do_s[warp];
  const float dld = valid ? a.dld[i] : 0.f;
  if (valid) {
    for (int j = lane; j < a.D; j += 32) dr[j] = a.dz_in[i * a.D + j];
    __syncwarp();
    for (int o = a.n_ops - 1; o >= 0; --o) {
      const GlueOp& op = a.ops[o];
      if (op.type == GLUE_ORTHO) {
        for (int k = lane; k < a.D; k += 32) {
          float s = 0.f;
#pragma unroll 8
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
    // affine update and tanh backward; lane m < dout owns element m; d_o kept in shared memory for the broadcast below
    if (valid && lane < a.dout) {
      const float ls = a.ls_save[i * a.dout + lane], yd = a.ydst_save[i * a.dout + lane];
      const float e = expf(ls), dn = dr[a.dst0 + lane];
      const float dso = fmaf(dn * yd, e, dld) * (1.0f - ls * ls);
      dr[a.dst0 + lane] = dn * e;
      a.d_o[i * no + lane] = dn;
      a.d_o[i * no + a.dout + lane] = dso;
      dov[lane] = dn;
      dov[a.dout + lane] = dso;
    }
    __syncwarp();
    // d a = d_o Wout (Wout staged in shared memory), then gelu'(pre) * dropout mask of the last hidden layer
    const unsigned long long seed = a.seed_ptr ? (a.seed ^ *a.seed_ptr) : a.seed;
    const float keep_scale = a.p_drop > 0.f ? 1.0f / (1.0f - a.p_drop) : 1.0f;
    for (int k0 = 0; k0 < a.H; k0 += kt) {
      if (k0 > 0) { __syncthreads(); stage_wout(w_s, a.Wout, no, a.H, k0, kt); }
      cp_async_wait_all();
      __syncthreads();
      if (!valid) continue;
      for (int kk = lane; kk < kt && k0 + kk < a.H; kk += 64) {
        // two columns per pass: independent chains
        const int k = k0 + kk, kb = kk + 32;
        const bool second = kb < kt && k0 + kb < a.H;
        const float pre0 = a.pre[i * a.pitch + k], pre1 = second ? a.pre[i * a.pitch + k + 32] : 0.f;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 2
        for (int m = 0; m < no; ++m) {
          const float dv = dov[m];
          s0 = fmaf(dv, w_s[m * kt + kk], s0);
          s1 = fmaf(dv, w_s[m * kt + (second ? kb : kk)], s1);
        }
        s0 *= dgelu_erf(pre0);
        s1 *= dgelu_erf(pre1);
        if (a.p_drop > 0.f) {
          s0 = dropout_uniform(seed, a.layer_uid, (unsigned long long)i * (unsigned)a.H + (unsigned)k) >= a.p_drop ? s0 * keep_scale : 0.f;
          s1 = dropout_uniform(seed, a.layer_uid, (unsigned long long)i * (unsigned)a.H + (unsigned)(k + 32)) >= a.p_drop ? s1 * keep_scale : 0.f;
        }
        a.d_pre[i * a.pitch + k] = s0;
        if (a.dpre_img) img_store1(a.dpre_img, a.img_plane, a.img_rpad, (int)i, k, s0);
        if (second) {
          a.d_pre[i * a.pitch + k + 32] = s1;
          if (a.dpre_img) img_store1(a.dpre_img, a.img_plane, a.img_rpad, (int)i, k + 32, s1);
        }
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

// dz[:, src] += d_pre . W1[:, :din].  Lane k < din owns column k: it walks j with d_pre[j] broadcast from shared memory.
// Dynamic shared memory: w_s[jt][wp] (the first din columns of W1; lanes = consecutive k), then one jt-float row per warp.
__global__ void __launch_bounds__(32 * kGlueWarps) train_pre_bwd_kernel(const TrainPreBwdArgs a, const int jt, const int wp) {
  float* w_s = glue_dyn_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* x_s = glue_dyn_smem + jt * wp + warp * jt;
  const long long i = (long long)blockIdx.x * kGlueWarps + warp;
  const bool valid = i < a.B;
  const float* dr = a.d_pre + (valid ? i : 0) * a.pitch;
  const int kl = lane < a.din ? lane : 0;
  float acc0 = 0.f, acc1 = 0.f;
  for (int j0 = 0; j0 < a.H; j0 += jt) {
    if (j0 > 0) __syncthreads();
    for (int e = threadIdx.x; e < jt * a.din; e += 32 * kGlueWarps) {
      const int jj = e / a.din, k = e - jj * a.din;
      const bool ok = j0 + jj < a.H;
      cp_async4(w_s + jj * wp + k, a.W1 + (long long)(ok ? j0 + jj : 0) * a.w1_pitch + k, ok);
    }
    stage_row(x_s, dr, valid ? a.H : 0, j0, jt, lane);
    cp_async_wait_all();
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < jt; jj += 2) {
      acc0 = fmaf(x_s[jj], w_s[jj * wp + kl], acc0);
      acc1 = fmaf(x_s[jj + 1], w_s[(jj + 1) * wp + kl], acc1);
    }
  }
  if (valid && lane < a.din) a.dz[i * a.D + a.src0 + lane] += acc0 + acc1;
}

}  // namespace bcnf
