// Row-per-thread fused coupling stack (narrow conditioners: padded width 16 or 32).
//
// Replaces the per-layer Python loop of CondRealNVP_v2.forward / .inverse
// (reference cnf.py:479-488, :500-506) and everything it calls for one row:
// ActNorm (cnf.py:348-354), the conditioner MLP (cnf.py:98-107), the affine update and
// log-det row sum (cnf.py:175-196, :198-213) and the orthonormal mixing (cnf.py:333-339).
//
// One thread owns one row: y (D floats), the hidden vector (HP floats) and the log-det
// accumulator live in registers from the first layer to the last; HBM is read once (y, the
// projection slices P) and written once (z, logdet).  Parameters are streamed chunk by chunk
// into shared memory with TMA bulk copies (cp.async.bulk + mbarrier, double buffered); every
// lane of a warp reads the same weight address, so shared-memory reads are broadcasts.
#pragma once
#include "common.cuh"

namespace bcnf {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Shared-memory reads by 32-bit shared-window address.  The parameter buffers are reached through
// runtime-selected pointers, which the compiler would otherwise treat as generic (LD.E instead of
// LDS: measured 44 % of all stall samples in the first version of this kernel).
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// acc[j] += x * w[j], j < N, w = shared-window byte address (broadcast reads)
template <int N>
__device__ __forceinline__ void axpy_row(float (&acc)[N], float x, uint32_t w) {
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    float4 v = lds4(w + 4 * j);
    acc[j + 0] = fmaf(x, v.x, acc[j + 0]);
    acc[j + 1] = fmaf(x, v.y, acc[j + 1]);
    acc[j + 2] = fmaf(x, v.z, acc[j + 2]);
    acc[j + 3] = fmaf(x, v.w, acc[j + 3]);
  }
}

// Packed variant: acc holds N/2 fp32 pairs; one FFMA2 updates two output columns with the same input x.
__device__ __forceinline__ void lds2x2(uint32_t addr, f32x2& a, f32x2& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
template <int N>
__device__ __forceinline__ void axpy_row2(f32x2 (&acc)[N / 2], float x, uint32_t w) {
  const f32x2 xx = pack2(x, x);
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    f32x2 w01, w23;
    lds2x2(w + 4 * j, w01, w23);
    acc[j / 2] = fma2(xx, w01, acc[j / 2]);
    acc[j / 2 + 1] = fma2(xx, w23, acc[j / 2 + 1]);
  }
}

// R rows per thread: every weight load (a warp-wide broadcast from shared memory) feeds R packed FMAs per pair of
// columns instead of one -- the first version of this kernel was issue-bound on those loads (ncu r01: FMA pipe 53 %,
// issue 70 %, LSU 35 %; one LDS.128 per two FFMA2).
template <int N, int R>
__device__ __forceinline__ void axpy_rows2(f32x2 (&acc)[R][N / 2], const float (&x)[R], uint32_t w) {
  f32x2 xx[R];
#pragma unroll
  for (int r = 0; r < R; ++r) xx[r] = pack2(x[r], x[r]);
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    f32x2 w01, w23;
    lds2x2(w + 4 * j, w01, w23);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      acc[r][j / 2] = fma2(xx[r], w01, acc[r][j / 2]);
      acc[r][j / 2 + 1] = fma2(xx[r], w23, acc[r][j / 2 + 1]);
    }
  }
}

// One conditioner network + affine update of the other half (cnf.py:98-107, :178-190, :203-204), R rows at once.
template <int D, int HP, int SRC, int R>
__device__ __forceinline__ void half_coupling_rows(float (&y)[R][D], float (&ld)[R], uint32_t w, const HalfLayout& hl,
                                                   const float* const (&prow)[R], int proj_off, int inverse) {
  constexpr int DA = (D + 1) / 2, DB = D / 2;
  constexpr int DIN = SRC == 0 ? DA : DB;
  constexpr int DOUT = SRC == 0 ? DB : DA;
  constexpr int IN0 = SRC == 0 ? 0 : DA;
  constexpr int OUT0 = SRC == 0 ? DA : 0;
  constexpr int DOP = (DOUT + 3) / 4 * 4;

  f32x2 acc[R][HP / 2];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < HP; j += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(prow[r] + proj_off + j));
      acc[r][j / 2] = pack2(v.x, v.y);
      acc[r][j / 2 + 1] = pack2(v.z, v.w);
    }
  {
    const uint32_t w1 = w + 4u * hl.off_w[0];
#pragma unroll
    for (int i = 0; i < DIN; ++i) {
      float x[R];
#pragma unroll
      for (int r = 0; r < R; ++r) x[r] = y[r][IN0 + i];
      axpy_rows2<HP, R>(acc, x, w1 + 4u * (i * HP));
    }
  }
  const int L = hl.L;
  for (int l = 1;; ++l) {
    float hcur[R][HP];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < HP; j += 2) unpack2(gelu_erf_fast2(acc[r][j / 2]), hcur[r][j], hcur[r][j + 1]);
    if (l >= L) {
      f32x2 ts2[R][DOP];
      const uint32_t wo = w + 4u * hl.off_wout;
      const uint32_t bo = w + 4u * hl.off_bout;
#pragma unroll
      for (int j = 0; j < 2 * DOP; j += 4) {
        f32x2 b01, b23;
        lds2x2(bo + 4 * j, b01, b23);
#pragma unroll
        for (int r = 0; r < R; ++r) { ts2[r][j / 2] = b01; ts2[r][j / 2 + 1] = b23; }
      }
#pragma unroll
      for (int k = 0; k < HP; ++k) {
        float x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = hcur[r][k];
        axpy_rows2<2 * DOP, R>(ts2, x, wo + 4u * (k * (2 * DOP)));
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float ts[2 * DOP];
#pragma unroll
        for (int j = 0; j < 2 * DOP; j += 2) unpack2(ts2[r][j / 2], ts[j], ts[j + 1]);
        float ls_sum = 0.f;
#pragma unroll
        for (int j = 0; j < DOUT; ++j) {
          float ls = tanhf(ts[DOP + j]);                                       // cnf.py:107
          ls_sum += ls;
          if (!inverse) y[r][OUT0 + j] = fmaf(expf(ls), y[r][OUT0 + j], ts[j]);   // cnf.py:179
          else          y[r][OUT0 + j] = (y[r][OUT0 + j] - ts[j]) * expf(-ls);    // cnf.py:204
        }
        ld[r] += ls_sum;                                                       // cnf.py:190, :488
      }
      return;
    }
    const uint32_t wl = w + 4u * hl.off_w[l];
    const uint32_t bl = w + 4u * hl.off_b[l];
#pragma unroll
    for (int j = 0; j < HP; j += 4) {
      f32x2 b01, b23;
      lds2x2(bl + 4 * j, b01, b23);
#pragma unroll
      for (int r = 0; r < R; ++r) { acc[r][j / 2] = b01; acc[r][j / 2 + 1] = b23; }
    }
#pragma unroll
    for (int k = 0; k < HP; ++k) {
      float x[R];
#pragma unroll
      for (int r = 0; r < R; ++r) x[r] = hcur[r][k];
      axpy_rows2<HP, R>(acc, x, wl + 4u * (k * HP));
    }
  }
}

// (single-row form kept for reference / A-B timing: BCNF_ROWTHREAD_R=1)
template <int D, int HP, int SRC>
__device__ __forceinline__ void half_coupling(float (&y)[D], float& ld, uint32_t w,
                                              const HalfLayout& hl, const float* __restrict__ prow,
                                              int inverse) {
  constexpr int DA = (D + 1) / 2, DB = D / 2;
  constexpr int DIN = SRC == 0 ? DA : DB;
  constexpr int DOUT = SRC == 0 ? DB : DA;
  constexpr int IN0 = SRC == 0 ? 0 : DA;
  constexpr int OUT0 = SRC == 0 ? DA : 0;
  constexpr int DOP = (DOUT + 3) / 4 * 4;

  f32x2 acc[HP / 2];
  // first Linear: the condition part (W1h.h + b1) was hoisted into P (bcnf_cond_project)
#pragma unroll
  for (int j = 0; j < HP; j += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(prow + j));
    acc[j / 2] = pack2(v.x, v.y);
    acc[j / 2 + 1] = pack2(v.z, v.w);
  }
  {
    const uint32_t w1 = w + 4u * hl.off_w[0];
#pragma unroll
    for (int i = 0; i < DIN; ++i) axpy_row2<HP>(acc, y[IN0 + i], w1 + 4u * (i * HP));
  }
  const int L = hl.L;
  for (int l = 1;; ++l) {
    float hcur[HP];
#pragma unroll
    for (int j = 0; j < HP; j += 2)      // nn.GELU() (exact-erf form), two columns per instruction; Dropout = identity in eval
      unpack2(gelu_erf_fast2(acc[j / 2]), hcur[j], hcur[j + 1]);
    if (l >= L) {
      // last Linear -> (t, s); t = first DOUT outputs, s = last DOUT (chunk(2, dim=1), cnf.py:104)
      f32x2 ts2[DOP];
      const uint32_t wo = w + 4u * hl.off_wout;
      const uint32_t bo = w + 4u * hl.off_bout;
#pragma unroll
      for (int j = 0; j < 2 * DOP; j += 4) lds2x2(bo + 4 * j, ts2[j / 2], ts2[j / 2 + 1]);
#pragma unroll
      for (int k = 0; k < HP; ++k) axpy_row2<2 * DOP>(ts2, hcur[k], wo + 4u * (k * (2 * DOP)));
      float ts[2 * DOP];
#pragma unroll
      for (int j = 0; j < 2 * DOP; j += 2) unpack2(ts2[j / 2], ts[j], ts[j + 1]);
      float ls_sum = 0.f;
#pragma unroll
      for (int j = 0; j < DOUT; ++j) {
        float ls = tanhf(ts[DOP + j]);                              // cnf.py:107
        ls_sum += ls;
        if (!inverse) y[OUT0 + j] = fmaf(expf(ls), y[OUT0 + j], ts[j]);   // cnf.py:179
        else          y[OUT0 + j] = (y[OUT0 + j] - ts[j]) * expf(-ls);    // cnf.py:204
      }
      ld += ls_sum;                                                 // cnf.py:190, :488
      return;
    }
    const uint32_t wl = w + 4u * hl.off_w[l];
    const uint32_t bl = w + 4u * hl.off_b[l];
#pragma unroll
    for (int j = 0; j < HP; j += 4) lds2x2(bl + 4 * j, acc[j / 2], acc[j / 2 + 1]);
#pragma unroll
    for (int k = 0; k < HP; ++k) axpy_row2<HP>(acc, hcur[k], wl + 4u * (k * HP));
  }
}

constexpr int kRowThreadBlock = 128;

template <int D, int HP, int R>
__global__ void __launch_bounds__(kRowThreadBlock)
flow_rowthread_kernel(const FlowArgs a, const StackDims sd, const int chunk_cap_bytes) {
  constexpr int DP = (D + 3) / 4 * 4;
  constexpr int kTileRows = R * kRowThreadBlock;         // a thread owns rows tid, tid + 128, ... of its tile
  extern __shared__ __align__(128) unsigned char smem_rt[];
  float* buf[2] = {reinterpret_cast<float*>(smem_rt), reinterpret_cast<float*>(smem_rt + chunk_cap_bytes)};
  const uint32_t buf_addr0 = smem_u32(smem_rt);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_rt + 2 * (size_t)chunk_cap_bytes);

  const int tid = threadIdx.x;
  const long long n_tiles = (a.n_rows + kTileRows - 1) / kTileRows;
  if ((long long)blockIdx.x >= n_tiles) return;
  const long long my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const long long total = my_tiles * a.n_chunks;   // chunk instances this CTA consumes

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int g = 0; g < 2 && g < total; ++g) {
      const Chunk c = a.chunks[g % a.n_chunks];
      mbar_expect_tx(&bars[g], (uint32_t)c.bytes);
      tma_bulk_g2s(buf[g], a.blob + c.off, (uint32_t)c.bytes, &bars[g]);
    }
  }

  long long g = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    long long row[R];
    bool valid[R];
    float y[R][D];
    float ld[R];
    const float* prow[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      row[r] = tile * kTileRows + r * kRowThreadBlock + tid;
      valid[r] = row[r] < a.n_rows;
      ld[r] = 0.f;
      prow[r] = a.P;
      if (valid[r]) {
#pragma unroll
        for (int j = 0; j < D; ++j) y[r][j] = flow_input(a, row[r], j, D);
        prow[r] = a.P + row_instance(a, row[r]) * (long long)sd.PW;
      } else {
#pragma unroll
        for (int j = 0; j < D; ++j) y[r][j] = 0.f;
      }
    }

    for (int ci = 0; ci < a.n_chunks; ++ci, ++g) {
      const int b = (int)(g & 1);
      mbar_wait(&bars[b], (uint32_t)((g >> 1) & 1));
      const Chunk c = a.chunks[ci];
      const uint32_t base = buf_addr0 + (uint32_t)b * (uint32_t)chunk_cap_bytes;
      for (int oi = 0; oi < c.n_ops; ++oi) {
        const DevOp op = a.ops[c.first_op + oi];
        const uint32_t w = base + 4u * (uint32_t)(op.off - c.off);
        if (op.type == DOP_HALF) {
          if (op.src == 0) half_coupling_rows<D, HP, 0, R>(y, ld, w, sd.half[0], prow, op.proj_off, op.inverse);
          else             half_coupling_rows<D, HP, 1, R>(y, ld, w, sd.half[1], prow, op.proj_off, op.inverse);
        } else if (op.type == DOP_MIX) {
          // y <- y @ M, M = Q (forward, cnf.py:335) or Q^T (inverse, cnf.py:339)
          f32x2 o2[R][DP / 2];
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < DP / 2; ++j) o2[r][j] = 0ull;
#pragma unroll
          for (int i = 0; i < D; ++i) {
            float x[R];
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = y[r][i];
            axpy_rows2<DP, R>(o2, x, w + 4u * (i * DP));
          }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float o[DP];
#pragma unroll
            for (int j = 0; j < DP; j += 2) unpack2(o2[r][j / 2], o[j], o[j + 1]);
#pragma unroll
            for (int j = 0; j < D; ++j) y[r][j] = o[j];
          }
        } else if (op.type == DOP_ACTNORM_FWD) {
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const float sc = lds1(w + 4 * j), bi = lds1(w + 4 * (DP + j));
#pragma unroll
            for (int r = 0; r < R; ++r) y[r][j] = fmaf(sc, y[r][j], bi);                               // cnf.py:349
          }
          const float c0 = lds1(w + 4 * (2 * DP));
#pragma unroll
          for (int r = 0; r < R; ++r) ld[r] += c0;                                                     // cnf.py:350
        } else {  // DOP_ACTNORM_INV
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const float sc = lds1(w + 4 * j), bi = lds1(w + 4 * (DP + j));
#pragma unroll
            for (int r = 0; r < R; ++r) y[r][j] = __fdiv_rn(y[r][j] - bi, sc);                         // cnf.py:354
          }
          const float c0 = lds1(w + 4 * (2 * DP));
#pragma unroll
          for (int r = 0; r < R; ++r) ld[r] += c0;
        }
      }
      __syncthreads();   // every thread is done reading buf[b]
      if (tid == 0 && g + 2 < total) {
        const Chunk n = a.chunks[(int)((g + 2) % a.n_chunks)];
        fence_proxy_async();
        mbar_expect_tx(&bars[b], (uint32_t)n.bytes);
        tma_bulk_g2s(buf[b], a.blob + n.off, (uint32_t)n.bytes, &bars[b]);
      }
    }

#pragma unroll
    for (int r = 0; r < R; ++r)
      if (valid[r]) {
#pragma unroll
        for (int j = 0; j < D; ++j) flow_output(a, row[r], j, D, y[r][j]);
        if (a.logdet) a.logdet[row[r]] = ld[r];
      }
  }
}

}  // namespace bcnf
