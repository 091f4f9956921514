// Layer-by-layer execution of the coupling stack on the CTA-pair GEMM (gemm_img2.cuh).
//
// Same scope and arithmetic as flow_tc.cuh (reference cnf.py:479-488, :500-506 and callees; 3-pass bf16 split with
// fp32 accumulation, fp32 everywhere outside the GEMM operands), different schedule.  The fused kernel keeps a 64-row
// tile per CTA on chip through all layers; its A tiles are too short for the tensor pipe to run at full rate
// (64 rows per CTA exposes the shared-memory read of A in every MMA).  Here a batch of rows
// goes through the stack one layer at a time: the hidden-layer GEMMs run as 256 x 256 tiles at the full MMA rate
// and hand the activations on as bf16 hi/lo operand images (L2-resident: two ping-pong images of one batch), and
// ONE glue kernel per conditioner network does everything between the last Linear of one network and the first
// hidden GEMM of the next:
//   finish the previous coupling (t, tanh/exp, affine update or its inverse, log-det row sum, cnf.py:175-213) from
//   the 2*dout outputs of its last Linear, apply the ActNorm / mixing layers in between (cnf.py:333-354), and
//   evaluate the next network's first Linear on its own half of y plus the hoisted condition projection P,
//   GELU, split, image.
#pragma once
#include "common.cuh"
#include "train_tc.cuh"

namespace bcnf {

constexpr int kLgRows = 32;                 // rows per CTA
constexpr int kLgThreads = 256;
constexpr int kLgMaxOps = 6;
constexpr int kLgDin = 12;                  // own-half inputs of a conditioner held in registers (D <= 24)

struct LayeredGlueArgs {
  float* Y;                 // (rows, DP) state of the batch
  float* LD;                // (rows)
  const float* in;          // first call of a batch: state is loaded from here (rows, D) and LD starts at 0
  float* out;               // last call: the state is written here (rows, D) ...
  float* logdet_out;        // ... and the log-det here (may be null)
  int n_rows;               // rows of this batch
  long long row_base;       // index of the batch's first row in the whole call (row -> instance)
  int D, DP;
  // coupling to finish: o = last Linear output incl. bias, t in [0, dout), s in [dop, dop + dout)
  const float* o; int o_ld;
  int prev_dst0, prev_dout, prev_dop, prev_inverse;
  // ActNorm / mixing ops between the two couplings, in execution order
  int n_ops; int op_type[kLgMaxOps]; const float* op_par[kLgMaxOps];
  // first Linear of the next conditioner network
  int has_next, src0, din, H1, H1p;
  const float* W1a;         // [dinp][H1p] input-major
  const float* P; long long PW; int proj_off;
  const int* row2inst; long long inst_period;
  unsigned char* img; long long img_plane; int img_rpad;
  int passes;
};

// Dynamic shared memory: the block's slice of the output image, [plane][K chunk][32 rows][128 B] in image layout; each
// 4 KB (plane, chunk) piece is contiguous in the image too and leaves with one bulk store (a lane-per-unit 16-byte
// store pattern touches a different line per lane: ~2 cycles per store, measured).
extern __shared__ __align__(128) unsigned char lg_stage[];

__global__ void __launch_bounds__(kLgThreads) flow_layered_glue_kernel(const LayeredGlueArgs a) {
  __shared__ float y_s[kLgRows][65];
  __shared__ float o_s[kLgRows][33];
  __shared__ float ld_s[kLgRows];
  __shared__ long long inst_s[kLgRows];      // row -> conditioning instance (one 64-bit modulo per row, not per thread)
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * kLgRows;
  const int D = a.D;

  // ---- phase A, block-parallel: load the state, finish the previous coupling, glue ops, write the state back ----
  for (int e = tid; e < kLgRows * 64; e += kLgThreads) {
    const int rl = e >> 6, j = e & 63, r = r0 + rl;
    float v = 0.f;
    if (r < a.n_rows && j < D) v = a.in ? a.in[(long long)r * D + j] : a.Y[(long long)r * a.DP + j];
    y_s[rl][j] = v;
  }
  if (tid < kLgRows) {
    ld_s[tid] = (!a.in && r0 + tid < a.n_rows) ? a.LD[r0 + tid] : 0.f;
    long long inst = 0;
    if (a.has_next && r0 + tid < a.n_rows) {
      const long long rg = a.row_base + r0 + tid;
      inst = a.row2inst ? (long long)a.row2inst[rg] : (a.inst_period > 0 ? rg % a.inst_period : rg);
    }
    inst_s[tid] = inst;
  }
  if (a.o)
    for (int e = tid; e < kLgRows * 32; e += kLgThreads) {
      const int rl = e >> 5, c = e & 31, r = r0 + rl;
      o_s[rl][c] = r < a.n_rows ? a.o[(long long)r * a.o_ld + c] : 0.f;
    }
  __syncthreads();
  if (a.o) {
    for (int e = tid; e < kLgRows * a.prev_dout; e += kLgThreads) {
      const int rl = e / a.prev_dout, m = e - rl * a.prev_dout;
      const float t = o_s[rl][m];
      const float ls = tanhf(o_s[rl][a.prev_dop + m]);                                             // cnf.py:107
      const float yd = y_s[rl][a.prev_dst0 + m];
      y_s[rl][a.prev_dst0 + m] = a.prev_inverse ? (yd - t) * expf(-ls) : fmaf(expf(ls), yd, t);    // cnf.py:204 / :179
      o_s[rl][a.prev_dop + m] = ls;
    }
    __syncthreads();
    if (tid < kLgRows) {
      float s = 0.f;
      for (int m = 0; m < a.prev_dout; ++m) s += o_s[tid][a.prev_dop + m];                         // cnf.py:190, fixed order
      ld_s[tid] += s;
    }
    __syncthreads();
  }
  for (int k = 0; k < a.n_ops; ++k) {
    const float* w = a.op_par[k];
    if (a.op_type[k] == DOP_MIX) {
      float acc[3];                                       // kLgRows * D <= 32 * 24 = 768 = 3 passes of 256 threads
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int e = tid + u * kLgThreads;
        acc[u] = 0.f;
        if (e < kLgRows * D) {
          const int rl = e / D, j = e - rl * D;
          float s = 0.f;
#pragma unroll 4
          for (int i = 0; i < D; ++i) s = fmaf(y_s[rl][i], __ldg(w + i * a.DP + j), s);            // y @ M
          acc[u] = s;
        }
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int e = tid + u * kLgThreads;
        if (e < kLgRows * D) { const int rl = e / D, j = e - rl * D; y_s[rl][j] = acc[u]; }
      }
    } else {
      for (int e = tid; e < kLgRows * D; e += kLgThreads) {
        const int rl = e / D, j = e - rl * D;
        const float sc = __ldg(w + j), b = __ldg(w + a.DP + j);
        y_s[rl][j] = a.op_type[k] == DOP_ACTNORM_FWD ? fmaf(sc, y_s[rl][j], b) : __fdiv_rn(y_s[rl][j] - b, sc);
      }
      if (tid < kLgRows) ld_s[tid] += __ldg(w + 2 * a.DP);
    }
    __syncthreads();
  }
  for (int e = tid; e < kLgRows * D; e += kLgThreads) {
    const int rl = e / D, j = e - rl * D, r = r0 + rl;
    if (r < a.n_rows) {
      if (a.out) a.out[(long long)r * D + j] = y_s[rl][j];
      else a.Y[(long long)r * a.DP + j] = y_s[rl][j];
    }
  }
  if (tid < kLgRows && r0 + tid < a.n_rows) {
    if (a.out) { if (a.logdet_out) a.logdet_out[r0 + tid] = ld_s[tid]; }
    else a.LD[r0 + tid] = ld_s[tid];
  }
  if (!a.has_next) return;

  // ---- phase B: unit = 8 consecutive hidden units (weights in registers) x every third row ----
  const int groups = a.H1p >> 3;
  const int n_chunks = (a.H1p + 63) >> 6;             // the image is padded to whole 64-column chunks (zeros)
  for (int unit = tid; unit < n_chunks * 8 * 3; unit += kLgThreads) {
    const int g = unit / 3, par = unit - g * 3;
    const int j0 = g * 8;
    const bool live = g < groups;
    float w[kLgDin][8];
#pragma unroll
    for (int k = 0; k < kLgDin; ++k) {
      if (k < a.din && live) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.W1a + (long long)k * a.H1p + j0));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.W1a + (long long)k * a.H1p + j0 + 4));
        w[k][0] = w0.x; w[k][1] = w0.y; w[k][2] = w0.z; w[k][3] = w0.w; w[k][4] = w1.x; w[k][5] = w1.y; w[k][6] = w1.z; w[k][7] = w1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) w[k][e] = 0.f;
      }
    }
    auto load_p = [&](int rl, float4& p0, float4& p1) {
      const int r = r0 + rl;
      p0 = make_float4(0.f, 0.f, 0.f, 0.f); p1 = p0;
      if (rl < kLgRows && r < a.n_rows && live) {
        const float* p = a.P + inst_s[rl] * a.PW + a.proj_off + j0;
        p0 = __ldg(reinterpret_cast<const float4*>(p)); p1 = __ldg(reinterpret_cast<const float4*>(p + 4));
      }
    };
    float4 pa0, pa1, pb0, pb1;
    load_p(par, pa0, pa1);
    for (int rl = par; rl < kLgRows; rl += 3) {
      load_p(rl + 3, pb0, pb1);                        // next row's projection slice in flight during this row's math
      const int r = r0 + rl;
      float v[8] = {pa0.x, pa0.y, pa0.z, pa0.w, pa1.x, pa1.y, pa1.z, pa1.w};
      if (r < a.n_rows && live) {
        const float* yr = y_s[rl] + a.src0;
#pragma unroll
        for (int k = 0; k < kLgDin; ++k) {
          const float x = yr[k];                       // (beyond din: zero weights; y_s rows are zero-padded to 64)
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = fmaf(x, w[k][e], v[e]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = j0 + e < a.H1 ? gelu_erf_fast(v[e]) : 0.f;
      }
      const int off = ((j0 >> 6) * kLgRows + rl) * 128 + ((((j0 & 63) >> 3) ^ (rl & 7)) << 4);    // r0 % 8 == 0: same swizzle
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        hi[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
        const float h0 = __uint_as_float(hi[e] << 16), h1 = __uint_as_float(hi[e] & 0xffff0000u);
        lo[e] = pack_bf16x2(v[2 * e] - h0, v[2 * e + 1] - h1);
      }
      *reinterpret_cast<uint4*>(lg_stage + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (a.passes == 3) *reinterpret_cast<uint4*>(lg_stage + n_chunks * kLgRows * 128 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      pa0 = pb0; pa1 = pb1;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  {
    const int planes = a.passes == 3 ? 2 : 1;
    if (tid < n_chunks * planes) {
      const int plane = tid / n_chunks, chunk = tid - plane * n_chunks;
      unsigned char* dst = a.img + (long long)plane * a.img_plane + ((long long)chunk * a.img_rpad + r0) * 128;
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(lg_stage + (plane * n_chunks + chunk) * kLgRows * 128);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(kLgRows * 128) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
}

}  // namespace bcnf
