// Row-tile fused coupling stack, fp32 FMA, any width (generic path and fp32-exact mode).
//
// Same scope as flow_rowthread.cuh (reference cnf.py:479-488, :500-506 and callees), for
// conditioners too wide to keep a row's hidden vector in one thread's registers.  One CTA owns
// R rows from the first layer to the last: y, the log-det accumulator and both activation
// buffers stay in shared memory; HBM sees y/P once in and z/logdet once out.  Weights are read
// straight from L2 (k-major, so a warp reads 512 contiguous bytes per k) and reused across the
// R rows of the tile from registers.
#pragma once
#include "common.cuh"

namespace bcnf {

constexpr int kTiledThreads = 256;

enum GemmInit { INIT_BIAS = 0, INIT_PROJ = 1 };
enum GemmEpi { EPI_NONE = 0, EPI_GELU = 1 };

// out[r][j] = epi(init[r][j] + sum_k in[r][k] * W[k*N + j])   r < R, j < N, k < K
// N, K multiples of 4; in rows zero-padded up to K; W global, k-major.
template <int R, int INIT, int EPI>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ in_s, int in_pitch, int K,
                                          const float* __restrict__ W, int N,
                                          const float* __restrict__ bias,     // INIT_BIAS: global [N]
                                          const float* const* prow_s,         // INIT_PROJ: smem [R] row pointers
                                          float* __restrict__ out_s, int out_pitch) {
  constexpr int RT = R / 8;            // rows per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = warp * RT;
  const int nq = N >> 2;
  for (int q0 = 0; q0 < nq; q0 += 64) {
    const int qa = q0 + lane, qb = q0 + 32 + lane;
    const bool va = qa < nq, vb = qb < nq;
    const int ja = qa << 2, jb = qb << 2;
    float acc[RT][8];
#pragma unroll
    for (int rr = 0; rr < RT; ++rr) {
      float4 ia = make_float4(0.f, 0.f, 0.f, 0.f), ib = ia;
      if (INIT == INIT_BIAS) {
        if (va) ia = __ldg(reinterpret_cast<const float4*>(bias + ja));
        if (vb) ib = __ldg(reinterpret_cast<const float4*>(bias + jb));
      } else {
        const float* p = prow_s[r0 + rr];
        if (va) ia = __ldg(reinterpret_cast<const float4*>(p + ja));
        if (vb) ib = __ldg(reinterpret_cast<const float4*>(p + jb));
      }
      acc[rr][0] = ia.x; acc[rr][1] = ia.y; acc[rr][2] = ia.z; acc[rr][3] = ia.w;
      acc[rr][4] = ib.x; acc[rr][5] = ib.y; acc[rr][6] = ib.z; acc[rr][7] = ib.w;
    }
    const float* wa = W + ja;
    const float* wb = W + jb;
#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += 4) {
      float4 wva[4], wvb[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        wva[kk] = va ? __ldg(reinterpret_cast<const float4*>(wa + (size_t)(k0 + kk) * N)) : make_float4(0, 0, 0, 0);
        wvb[kk] = vb ? __ldg(reinterpret_cast<const float4*>(wb + (size_t)(k0 + kk) * N)) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int rr = 0; rr < RT; ++rr) {
        const float4 x4 = *reinterpret_cast<const float4*>(in_s + (r0 + rr) * in_pitch + k0);
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          acc[rr][0] = fmaf(xs[kk], wva[kk].x, acc[rr][0]);
          acc[rr][1] = fmaf(xs[kk], wva[kk].y, acc[rr][1]);
          acc[rr][2] = fmaf(xs[kk], wva[kk].z, acc[rr][2]);
          acc[rr][3] = fmaf(xs[kk], wva[kk].w, acc[rr][3]);
          acc[rr][4] = fmaf(xs[kk], wvb[kk].x, acc[rr][4]);
          acc[rr][5] = fmaf(xs[kk], wvb[kk].y, acc[rr][5]);
          acc[rr][6] = fmaf(xs[kk], wvb[kk].z, acc[rr][6]);
          acc[rr][7] = fmaf(xs[kk], wvb[kk].w, acc[rr][7]);
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < RT; ++rr) {
      if (EPI == EPI_GELU) {
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[rr][c] = gelu_erf(acc[rr][c]);
      }
      float* o = out_s + (r0 + rr) * out_pitch;
      if (va) *reinterpret_cast<float4*>(o + ja) = make_float4(acc[rr][0], acc[rr][1], acc[rr][2], acc[rr][3]);
      if (vb) *reinterpret_cast<float4*>(o + jb) = make_float4(acc[rr][4], acc[rr][5], acc[rr][6], acc[rr][7]);
    }
  }
}

// Shared-memory carve-up (floats): y[R][YP] | y2[R][YP] | xin[R][XP] | ts[R][TP] | act0[R][AP] | act1[R][AP]
// | ld[R] | prow pointers[R]
struct TiledSmem {
  int YP, XP, TP, AP;
  __host__ __device__ size_t bytes(int R) const {
    return sizeof(float) * ((size_t)R * (2 * YP + XP + TP + 2 * AP) + R) + sizeof(void*) * R;
  }
};

template <int R>
__global__ void __launch_bounds__(kTiledThreads)
flow_tiled_kernel(const FlowArgs a, const StackDims sd, const TiledSmem lay) {
  extern __shared__ __align__(128) unsigned char smem_tl[];
  float* y_s = reinterpret_cast<float*>(smem_tl);
  float* y2_s = y_s + R * lay.YP;
  float* xin_s = y2_s + R * lay.YP;
  float* ts_s = xin_s + R * lay.XP;
  float* act0 = ts_s + R * lay.TP;
  float* act1 = act0 + R * lay.AP;
  float* ld_s = act1 + R * lay.AP;
  const float** prow_s = reinterpret_cast<const float**>(ld_s + R);

  const int tid = threadIdx.x;
  const int D = sd.D;
  const long long n_tiles = (a.n_rows + R - 1) / R;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * R;
    for (int e = tid; e < R * lay.YP; e += kTiledThreads) {
      const int r = e / lay.YP, j = e - r * lay.YP;
      const long long row = row0 + r;
      y_s[e] = (row < a.n_rows && j < D) ? flow_input(a, row, j, D) : 0.f;
    }
    if (tid < R) {
      const long long row = row0 + tid;
      ld_s[tid] = 0.f;
      prow_s[tid] = a.P + (row < a.n_rows ? row_instance(a, row) : 0) * (long long)sd.PW;
    }
    __syncthreads();

    for (int oi = 0; oi < a.n_ops; ++oi) {
      const DevOp op = a.ops[oi];
      const float* w = a.blob + op.off;
      if (op.type == DOP_HALF) {
        const HalfLayout& hl = sd.half[op.src];
        const int in0 = op.src == 0 ? 0 : sd.Da;
        const int out0 = op.src == 0 ? sd.Da : 0;
        // stage the conditioner's own-half input, zero padded to a multiple of 4
        for (int e = tid; e < R * lay.XP; e += kTiledThreads) {
          const int r = e / lay.XP, j = e - r * lay.XP;
          xin_s[e] = j < hl.din ? y_s[r * lay.YP + in0 + j] : 0.f;
        }
        if (tid < R) prow_s[tid] += op.proj_off;
        __syncthreads();
        float* cur = act0;
        float* nxt = act1;
        tile_gemm<R, INIT_PROJ, EPI_GELU>(xin_s, lay.XP, hl.dinp, w + hl.off_w[0], hl.hp[0], nullptr, prow_s,
                                          cur, lay.AP);
        __syncthreads();
        if (tid < R) prow_s[tid] -= op.proj_off;
        for (int l = 1; l < hl.L; ++l) {
          tile_gemm<R, INIT_BIAS, EPI_GELU>(cur, lay.AP, hl.hp[l - 1], w + hl.off_w[l], hl.hp[l], w + hl.off_b[l],
                                            nullptr, nxt, lay.AP);
          __syncthreads();
          float* t = cur; cur = nxt; nxt = t;
        }
        tile_gemm<R, INIT_BIAS, EPI_NONE>(cur, lay.AP, hl.hp[hl.L - 1], w + hl.off_wout, 2 * hl.dop,
                                          w + hl.off_bout, nullptr, ts_s, lay.TP);
        __syncthreads();
        if (tid < R) {
          // affine update + log-det row sum (cnf.py:179, :190 / :204)
          const float* ts = ts_s + tid * lay.TP;
          float* yr = y_s + tid * lay.YP + out0;
          float ls_sum = 0.f;
          for (int j = 0; j < hl.dout; ++j) {
            const float ls = tanhf(ts[hl.dop + j]);
            ls_sum += ls;
            if (!op.inverse) yr[j] = fmaf(expf(ls), yr[j], ts[j]);
            else             yr[j] = (yr[j] - ts[j]) * expf(-ls);
          }
          ld_s[tid] += ls_sum;
        }
        __syncthreads();
      } else if (op.type == DOP_MIX) {
        for (int e = tid; e < R * D; e += kTiledThreads) {
          const int r = e / D, j = e - r * D;
          float s = 0.f;
          for (int i = 0; i < D; ++i) s = fmaf(y_s[r * lay.YP + i], __ldg(w + i * sd.DP + j), s);
          y2_s[r * lay.YP + j] = s;
        }
        __syncthreads();
        for (int e = tid; e < R * D; e += kTiledThreads) {
          const int r = e / D, j = e - r * D;
          y_s[r * lay.YP + j] = y2_s[r * lay.YP + j];
        }
        __syncthreads();
      } else {
        for (int e = tid; e < R * D; e += kTiledThreads) {
          const int r = e / D, j = e - r * D;
          const float s = __ldg(w + j), b = __ldg(w + sd.DP + j);
          float v = y_s[r * lay.YP + j];
          v = op.type == DOP_ACTNORM_FWD ? fmaf(s, v, b) : __fdiv_rn(v - b, s);
          y_s[r * lay.YP + j] = v;
        }
        if (tid < R) ld_s[tid] += __ldg(w + 2 * sd.DP);
        __syncthreads();
      }
    }

    for (int e = tid; e < R * D; e += kTiledThreads) {
      const int r = e / D, j = e - r * D;
      const long long row = row0 + r;
      if (row < a.n_rows) flow_output(a, row, j, D, y_s[r * lay.YP + j]);
    }
    if (a.logdet && tid < R && row0 + tid < a.n_rows) a.logdet[row0 + tid] = ld_s[tid];
    __syncthreads();
  }
}

}  // namespace bcnf
