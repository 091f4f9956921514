// Shared host/device definitions of the packed coupling-stack program.
//
// A stack (CondRealNVP_v2.layers, reference cnf.py:392-423) is compiled at set_params time
// into a linear "program" of device ops, one program per direction, and one fp32 parameter
// blob per direction laid out in program order so that any run of consecutive ops is one
// contiguous byte range (the row-per-thread kernel streams such ranges with TMA bulk copies).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef BCNF_MAX_SIZE
#define BCNF_MAX_SIZE 64
#endif
#define BCNF_TC_MAX_LAYERS 10

namespace bcnf {

enum DevOpType : int {
  DOP_ACTNORM_FWD = 0,  // y = s*y + b ; ld += c          blob: s[DP] b[DP] c[4]
  DOP_ACTNORM_INV = 1,  // y = (y-b)/s ; ld += c(=-?)     same blob
  DOP_MIX = 2,          // y = y @ M                       blob: M[D][DP]   (M = Q or Q^T)
  DOP_HALF = 3          // one conditioner network + affine update of the other half
};

struct DevOp {
  int type;
  int src;        // DOP_HALF: 0 = conditioner reads first half (nn_a), 1 = second half (nn_b)
  int proj_off;   // DOP_HALF: column offset of this network's slice in P
  int inverse;    // DOP_HALF: 0 = z = exp(ls)*y + t ; 1 = y = (z - t)*exp(-ls)
  long long off;  // float offset of the op's parameters in the blob of its direction
};

struct Chunk {      // a run of consecutive ops whose parameters are streamed together
  long long off;    // float offset in blob (16-byte aligned)
  int bytes;        // multiple of 16
  int first_op;
  int n_ops;
  int pad;
};

// Parameter layout of one half-coupling inside the blob (all counts in floats).
//   W1a  [DINP][HP0]                       rows >= din are zero
//   for l = 1..L-1:  W_l [HP(l-1)][HP(l)] , b_l [HP(l)]
//   Wout [HP(L-1)][2*DOP] , bout [2*DOP]   t in columns [0,dout), s in [DOP, DOP+dout)
// Matrices are stored input-major (k-major): element (k, j) multiplies input k into output j.
struct HalfLayout {
  int din, dinp;     // conditioner's own-half input width, padded to 4
  int dout, dop;     // width of the half being transformed, padded to 4
  int L;
  int hp[8];         // padded hidden widths (multiples of 16)
  int h[8];          // true hidden widths
  int off_w[8];      // off_w[0] = W1a; off_w[l] = W_l
  int off_b[8];      // off_b[l] = b_l (l >= 1)
  int off_wout, off_bout;
  int total;         // floats, multiple of 4
};

struct StackDims {
  int D, DP;         // flow dimension, padded to 4
  int Da, Db;        // ceil(D/2), floor(D/2)   (torch.chunk(2), cnf.py:175)
  int C;
  int PW;            // projection width = sum over half-couplings of HP0
  HalfLayout half[2];  // [0]: nn_a (reads a, updates b)   [1]: nn_b (reads b, updates a)
};

static inline __host__ __device__ int round_up(int x, int m) { return (x + m - 1) / m * m; }

__device__ __forceinline__ float gelu_erf(float x) {
  // nn.GELU() default (exact erf), as instantiated by LayerFactory (factories.py:65-66)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// Branch-free GELU for the tensor-core epilogue: erfc(t) = 2^(-t*q(t)) with q a degree-9 minimax
// fit of -log2(erfc(t))/t on [0, 4] (|erf error| <= 1.1e-7 evaluated in fp32, i.e. the rounding
// level of erf itself; fitted by the script quoted in DESIGN.md).  17 instructions vs ~30 for erff.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float t = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
  float p = -7.638800192e-07f;           // coefficients of q(t), already multiplied by -log2(e)
  p = fmaf(p, t, 1.656447655e-05f);
  p = fmaf(p, t, -1.540620985e-04f);
  p = fmaf(p, t, 7.796742463e-04f);
  p = fmaf(p, t, -2.041655680e-03f);
  p = fmaf(p, t, -2.589820766e-04f);
  p = fmaf(p, t, 2.797563118e-02f);
  p = fmaf(p, t, -1.483925716e-01f);
  p = fmaf(p, t, -9.184330629e-01f);
  p = fmaf(p, t, -1.627907331e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p * t));
  const float erf_v = copysignf(1.0f - e, x);
  const float hx = 0.5f * x;
  return fmaf(hx, erf_v, hx);
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 process two fp32 lanes per instruction) ----------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// gelu_erf_fast on two values at once: the Horner chain, the products and the final fma run packed
// (about 10.5 instructions per value instead of 17).  Same polynomial, same results bit for bit.
__device__ __forceinline__ f32x2 gelu_erf_fast2(f32x2 x) {
  float x0, x1;
  unpack2(x, x0, x1);
  const float t0 = fminf(fabsf(x0) * 0.70710678118654752440f, 4.0f);
  const float t1 = fminf(fabsf(x1) * 0.70710678118654752440f, 4.0f);
  const f32x2 t = pack2(t0, t1);
  f32x2 p = pack2(-7.638800192e-07f, -7.638800192e-07f);
  p = fma2(p, t, pack2(1.656447655e-05f, 1.656447655e-05f));
  p = fma2(p, t, pack2(-1.540620985e-04f, -1.540620985e-04f));
  p = fma2(p, t, pack2(7.796742463e-04f, 7.796742463e-04f));
  p = fma2(p, t, pack2(-2.041655680e-03f, -2.041655680e-03f));
  p = fma2(p, t, pack2(-2.589820766e-04f, -2.589820766e-04f));
  p = fma2(p, t, pack2(2.797563118e-02f, 2.797563118e-02f));
  p = fma2(p, t, pack2(-1.483925716e-01f, -1.483925716e-01f));
  p = fma2(p, t, pack2(-9.184330629e-01f, -9.184330629e-01f));
  p = fma2(p, t, pack2(-1.627907331e+00f, -1.627907331e+00f));
  float a0, a1;
  unpack2(mul2(p, t), a0, a1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const f32x2 erf_v = pack2(copysignf(1.0f - e0, x0), copysignf(1.0f - e1, x1));
  const f32x2 hx = mul2(x, pack2(0.5f, 0.5f));
  return fma2(hx, erf_v, hx);
}

// The same on four pairs in lock step: every Horner step is written for all four pairs before the next one, so the
// four dependent chains (10 packed FMAs each) are interleaved in the instruction stream instead of running one after the
// other (an epilogue warp has three others on its scheduler: a chain issued alone waits for its own FMA latency most of
// the time).  Same operations on every value as gelu_erf_fast2: identical results.
__device__ __forceinline__ void gelu_erf_fast2x4(f32x2 (&v)[4]) {
  float x0[4], x1[4];
  f32x2 t[4], p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unpack2(v[i], x0[i], x1[i]);
    t[i] = pack2(fminf(fabsf(x0[i]) * 0.70710678118654752440f, 4.0f), fminf(fabsf(x1[i]) * 0.70710678118654752440f, 4.0f));
    p[i] = pack2(-7.638800192e-07f, -7.638800192e-07f);
  }
#define BCNF_GELU_STEP(c)                                      \
  _Pragma("unroll") for (int i = 0; i < 4; ++i) p[i] = fma2(p[i], t[i], pack2(c, c));
  BCNF_GELU_STEP(1.656447655e-05f)
  BCNF_GELU_STEP(-1.540620985e-04f)
  BCNF_GELU_STEP(7.796742463e-04f)
  BCNF_GELU_STEP(-2.041655680e-03f)
  BCNF_GELU_STEP(-2.589820766e-04f)
  BCNF_GELU_STEP(2.797563118e-02f)
  BCNF_GELU_STEP(-1.483925716e-01f)
  BCNF_GELU_STEP(-9.184330629e-01f)
  BCNF_GELU_STEP(-1.627907331e+00f)
#undef BCNF_GELU_STEP
  float e0[4], e1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a0, a1;
    unpack2(mul2(p[i], t[i]), a0, a1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0[i]) : "f"(a0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1[i]) : "f"(a1));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const f32x2 erf_v = pack2(copysignf(1.0f - e0[i], x0[i]), copysignf(1.0f - e1[i], x1[i]));
    const f32x2 hx = mul2(v[i], pack2(0.5f, 0.5f));
    v[i] = fma2(hx, erf_v, hx);
  }
}

// Launch arguments common to the flow kernels.
struct FlowArgs {
  const float* in;        // (n_rows, D)
  float* out;             // (n_rows, D)
  float* logdet;          // (n_rows) or null
  const float* P;         // (n_inst, PW)
  const int* row2inst;    // or null
  long long inst_period;  // used when row2inst == null; 0 = identity
  long long n_rows;
  const float* blob;
  const DevOp* ops;
  int n_ops;
  const Chunk* chunks;
  int n_chunks;
  long long* trace;       // debug: clock64 stamps of the tensor-core pipeline (null in normal runs)
  long long blob_floats;  // size of `blob` (floats)
  // in == null: the input is drawn inside the kernel, z[row, j] = sigma * N(0, 1) from Philox4x32-10 keyed by `seed` at
  // counter row * D + j (cnf.py:566,578,584 draw it with torch.randn on the CPU and copy it over)
  unsigned long long seed;
  float sigma;
  // rank_out != null: the result is not written; instead rank_out[inst, j] += (x[row, j] < rank_y[inst, j]), the
  // reduction of compute_y_hat_ranks (src/bcnf/eval/calibration.py:44-48), inst = the row's conditioning instance
  const float* rank_y;    // (n_inst, D)
  int* rank_out;          // (n_inst, D), zeroed by the caller
};

// ---- counter-based normal deviates: Philox4x32-10 (Salmon et al. 2011) + Box-Muller ----------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// element e of the (n_rows, D) standard-normal field of `seed`: word e & 3 of the Philox block at counter e >> 2
// (two Box-Muller pairs per block).  A pure function of (seed, e): the same value in every kernel, tile and launch.
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned long long e) {
  uint32_t w[4];
  const unsigned long long ctr = e >> 2;
  philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0x6263u, 0x6e66u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
  const int lane = (int)(e & 3ull);
  const uint32_t a = w[lane & 2], b = w[(lane & 2) + 1];
  const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);        // (0, 1]
  const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);                  // [0, 1)
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return r * ((lane & 1) ? sn : cs);
}
// input element (row, j): read, or drawn in place when the caller passed no input
__device__ __forceinline__ float flow_input(const FlowArgs& a, long long row, int j, int D) {
  const long long e = row * D + j;
  return a.in ? __ldg(a.in + e) : a.sigma * philox_normal(a.seed, (unsigned long long)e);
}

__device__ __forceinline__ long long row_instance(const FlowArgs& a, long long r) {
  if (a.row2inst) return (long long)a.row2inst[r];
  if (a.inst_period > 0) return r % a.inst_period;
  return r;
}

// result element (row, j): written, or folded into the rank counters of its conditioning instance
__device__ __forceinline__ void flow_output(const FlowArgs& a, long long row, int j, int D, float v) {
  if (a.rank_out) {
    const long long i = row_instance(a, row) * D + j;
    if (v < __ldg(a.rank_y + i)) atomicAdd(a.rank_out + i, 1);
  } else {
    a.out[row * D + j] = v;
  }
}

}  // namespace bcnf
