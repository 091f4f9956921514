// Fused coupling stack, second generation: 128 rows per CTA, activations streamed through L2.
//
// Same scope and arithmetic as flow_tc.cuh (reference cnf.py:479-488, :500-506 and callees: conditioner MLP
// cnf.py:98-107, affine coupling cnf.py:165-213, ActNorm cnf.py:348-354, orthonormal mixing cnf.py:333-339;
// 3-pass bf16 split with fp32 accumulation, fp32 everywhere outside the GEMM operands).  What changes is where the
// activations live.  flow_tc.cuh keeps a layer's activations in shared memory, which caps a CTA at 64 rows (TMEM
// cannot hold 128 rows x 528 fp32 columns next to them) and therefore caps the tensor pipe at 58 % : an SS-mode
// tcgen05.mma re-reads its A rows for every N chunk, and with 64 rows per CTA that read costs more shared-memory
// cycles than the MMA lasts.  Here a CTA PAIR owns a 256-row tile (128 rows per CTA, the
// M = 256 cta_group::2 MMA of gemm_img2.cuh) and a layer's
// activations are an OPERAND IMAGE (bf16 hi / lo planes, [K/64][128 rows][128 B], SWIZZLE_128B tile layout) in a
// per-CTA scratch that never leaves L2:
//
//   * producer warp: per N chunk (<= 256 columns) and 64-wide K stage, bulk copies of the CTA's 128 rows of the
//     activation image and of its half of the weight image tile into a 3-stage ring (64 KB per stage);
//   * one issuing thread (leader CTA): tcgen05.mma.cta_group::2 M = 256, N = chunk, K = 16 into one of two
//     256-column TMEM accumulators;
//   * sixteen epilogue warps per CTA drain the other accumulator: + bias (folded into the GEMM where the layer
//     width leaves a padding column; else added here; + the hoisted condition projection P for the first Linear)
//     -> exact-erf GELU, four value pairs in lock step (packed fp32x2) -> bf16 hi / lo split.  A thread owns a row
//     and 16 columns of every 64-column image chunk: its two 16-byte units are ONE aligned 32-byte sector of the swizzled image row, so it
//     stores them straight to the scratch (st.global.v8.b32, a full L2 sector per lane).  No staging, no barrier
//     inside a layer; after an N chunk every warp fences and counts itself, and the producer starts the next
//     layer's K stages as soon as the N chunk they read has been counted sixteen times, so the MMAs of layer l+1
//     overlap the epilogue of the last N chunk of layer l (which is the short one: 528 = 256 + 256 + 16);
//   * the last Linear (N = 2 x dout padded to 16) leaves t and s in TMEM; the epilogue warps apply tanh / exp, the
//     affine update, the log-det row sum, ActNorm and the orthonormal mixing in fp32, keep y in shared memory
//     between networks, and stage the next network's own-half input as image chunk 0.
//
// A 528-wide layer is three N chunks (256, 256, 16).  With 128 A rows per CTA an SS-mode MMA runs at the math floor
// for every N >= 128 (tools/ts_probe.cu); the third chunk re-streams A for 16 columns and is bound by the L2 round
// trip of its stages.  Per CTA and hidden layer the L2 traffic is 3 x 270 KB of activations + 608 KB of weights in,
// 270 KB out; the scratch of all 148 CTAs is 85 MB of the 126 MB L2, and a layer's input image is dropped from L2
// (discard.global.L2) once the layer's MMAs have completed.  Measurements and what was tried: DESIGN.md section 6.1.
#pragma once
#include "flow_tc.cuh"
#include "gemm_img2.cuh"

namespace bcnf {

constexpr int kS2Stages = 3;
constexpr int kS2EpiWarps = 16;
constexpr int kS2EpiThreads = 32 * kS2EpiWarps;
constexpr int kS2FirstEpi = 2;                  // warp 0 producer, 1 issuer (leader) / relay (peer); 18 warps = 576 threads,
                                                // so a thread may use 112 registers (96 with two more, idle, warps)
constexpr int kS2Threads = 32 * kS2FirstEpi + kS2EpiThreads;
constexpr int kS2Rows = 128;                    // rows per CTA (256 per pair)
constexpr int kS2Tile = kS2Rows * 128;          // bytes of one (128 rows x 64 k) bf16 tile / image chunk
constexpr int kS2SlotCols = 256;                // TMEM: two accumulator slots
constexpr int kS2MaxChunks = 4;
constexpr int kS2YPitch = 25;                   // floats per row of the y state (D <= 24; odd: no bank conflicts)
constexpr int kS2TsPitch = 25;                  // floats per row of the (t | s) scratch (2 * dop <= 24 columns; odd)
constexpr int kS2GparOff = kS2Rows * (kS2YPitch + kS2TsPitch) * 4;   // glue scratch: [ts][y][staged ActNorm / mixing parameters]
constexpr int kS2GparBytes = 2 * kS2Rows * 128 - kS2GparOff;         // 7168
constexpr int kS2GlueMax = 6;                   // ActNorm / mixing layers between two conditioner networks
constexpr int kS2MiscBytes = 2048;

struct S2Layer {
  int n_chunks;                 // N chunks (<= 256 columns each, multiples of 16; all but the last multiples of 64)
  int chunk_n[kS2MaxChunks];
  int n_kst;                    // 64-wide K stages
  int last_ksteps;              // K = 16 steps that carry data in the last stage
  int n_img;                    // 64-column image chunks of this layer's OUTPUT (hidden layers; 0 for the last Linear)
  int w_rpad;                   // rows of the weight image
  int bias_k;                   // >= 0: the bias is folded into the GEMM as weight column bias_k (the input image carries
                                // a constant 1 there: a free padding column of the previous layer); -1: added in the epilogue
  int one_col;                  // >= 0: output column that the epilogue overwrites with 1 (= bias_k of the next layer)
  int pad;
  long long w_off;              // byte offset of the weight image inside the network's block
  long long w_plane;            // bytes between its hi and lo plane
};

struct S2Half {
  int L;                        // hidden layers; layer[0] first Linear, [1..L-1] hidden, [L] last Linear
  int n_last;                   // N of the last Linear: 2 * dop padded to 16 (t in columns [0, dout), s in [dop, dop + dout))
  S2Layer layer[kTcMaxLayers];
  long long net_bytes;          // bytes of one network's weight images
};

struct S2Dims {
  S2Half half[2];
  int n_halfops, two_way;
  int a_kchunks;                // image chunks per activation buffer
  long long a_plane;            // bytes of one plane of a CTA's activation buffer
  int stage_bytes, b_off, b_lo_off;   // stage: [A hi][A lo][B hi][B lo]
  int stg_off, misc_off, smem_bytes;
  long long cta_bytes;          // scratch bytes per CTA: 2 buffers x planes x a_plane
  int debug;                    // timing experiments only (results are wrong): 1 skip the GELU arithmetic, 2 skip the
                                // activation stores, 4 skip the release of the output groups, 8 skip the TMEM loads,
                                // 16 keep dead activation lines in L2 (no discard)
};

// ---- waits with a watchdog: a protocol bug traps (the launch fails) instead of hanging the GPU --------------------
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __noinline__ void s2_timeout(unsigned int* dbg, uint32_t code, uint32_t aux) {
  if (dbg) { dbg[0] = code; dbg[1] = aux; dbg[2] = blockIdx.x; dbg[3] = threadIdx.x; __threadfence_system(); }
  __trap();
}
// SPIN = true: test_wait polling (single hot waiter); false: try_wait (suspending; many waiters)
template <bool SPIN>
__device__ __forceinline__ void s2_wait(uint64_t* bar, uint32_t parity, unsigned int* dbg, uint32_t code) {
  if (SPIN ? mbar_test(bar, parity) : mbar_try(bar, parity)) return;
  const unsigned long long t0 = g2_now();
  uint32_t n = 0;
  for (;;) {
    if (SPIN ? mbar_test(bar, parity) : mbar_try(bar, parity)) return;
    if (((++n) & 1023u) == 0 && g2_now() - t0 > 4000000000ull) s2_timeout(dbg, code, parity);
  }
}
// debug trace (FlowArgs::trace, block 0 only): role r writes stamp k of its event idx at trace[r * 8192 + idx * 4 + k]
__device__ __forceinline__ void s2_stamp(long long* trace, int role, uint32_t idx, int k) {
  if (trace && blockIdx.x == 0 && idx < 2048u) trace[role * 8192 + idx * 4 + k] = clock64();
}
// The input image of a layer is dead once the layer's MMAs have completed: drop its (dirty) lines from L2 instead of
// letting them be written back to HBM before the next-but-one layer overwrites them (ncu: 16.9 GB of DRAM writes per
// 75 776 rows, 77 % of all activation bytes, before this).  One 128-byte line = one row of one image chunk.
__device__ __forceinline__ void discard_line(const unsigned char* p) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void s2_epi_sync() { asm volatile("bar.sync 3, %0;" ::"n"(kS2EpiThreads) : "memory"); }
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// wait for the loads issued so far; the destination registers are operands, so no use of them can move above the wait
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}
// read of the projection slice P[instance, network]: 256-bit loads that bypass L1.  (An L2 evict-first hint was tried
// and dropped: with M samples per instance the same slice is read by every tile that holds rows of the instance, and
// the CTA pairs walk the networks roughly in step, so the reads after the first are L2 hits only if the line stays.)
__device__ __forceinline__ void ldg_stream8(const float* p, float4& lo, float4& hi) {
  uint32_t r[8];
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
  lo = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
  hi = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
}

// one aligned 32-byte sector = two 16-byte units of an image row (swapped when the row's swizzle phase is odd)
__device__ __forceinline__ void st_global_sector(unsigned char* p, const uint32_t (&w)[8], bool swap) {
  if (swap)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]),
                 "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  else
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

// (acc + add) -> GELU -> bf16 hi / lo for 8 columns: four packed words per plane, written at hi[o..o+3] / lo[o..o+3]
template <int NPASS, bool ADD, int O>
__device__ __forceinline__ void s2_gelu_pack8(const uint32_t* r, const float4& b0, const float4& b1, uint32_t (&hi)[8], uint32_t (&lo)[8]) {
  f32x2 v[4];
  if (ADD) {
    v[0] = add2(pack2(__uint_as_float(r[0]), __uint_as_float(r[1])), pack2(b0.x, b0.y));
    v[1] = add2(pack2(__uint_as_float(r[2]), __uint_as_float(r[3])), pack2(b0.z, b0.w));
    v[2] = add2(pack2(__uint_as_float(r[4]), __uint_as_float(r[5])), pack2(b1.x, b1.y));
    v[3] = add2(pack2(__uint_as_float(r[6]), __uint_as_float(r[7])), pack2(b1.z, b1.w));
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
  }
  gelu_erf_fast2x4(v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float x0, x1;
    unpack2(v[i], x0, x1);
    hi[O + i] = pack_bf16x2(x0, x1);
    lo[O + i] = 0u;
    if (NPASS == 3) {
      const f32x2 h = pack2(__uint_as_float(hi[O + i] << 16), __uint_as_float(hi[O + i] & 0xffff0000u));
      float l0, l1;
      unpack2(fma2(h, pack2(-1.0f, -1.0f), v[i]), l0, l1);      // v - hi in one packed FMA (no sign-flip LOP3s)
      lo[O + i] = pack_bf16x2(l0, l1);
    }
  }
}

// Column oc (0..15, else nothing) of the 16 this thread has just stored := 1.0 (hi 0x3f80, lo 0): the constant input of
// the next layer's folded bias.  A second, 2-byte store by the same thread to the same location (ordered after the sector
// store); patching the packed words in registers instead made the compiler index them through local memory.
template <int NPASS>
__device__ __forceinline__ void s2_store_one(int oc, unsigned char* img_chunk, long long plane, int row, int part) {
  if (oc >= 0 && oc < 16) {
    const int unit = (part * 2 + (oc >> 3)) ^ (row & 7);
    unsigned char* p = img_chunk + row * 128 + unit * 16 + (oc & 7) * 2;
    *reinterpret_cast<volatile unsigned short*>(p) = (unsigned short)0x3f80;
    if (NPASS == 3) *reinterpret_cast<volatile unsigned short*>(p + plane) = (unsigned short)0;
  }
}

// Folded biases: weight image column k of row n := bias[n] (bf16 hi / lo), for a batch of (network, layer) images.
struct S2BiasCol {
  const float* bias;      // [n]
  unsigned char* img;     // hi plane of the weight image ([K/64][rpad][128 B], rows = output unit)
  long long plane;
  int rpad, n, k, pad;
};
__global__ void s2_bias_col_kernel(const S2BiasCol* __restrict__ descs) {
  const S2BiasCol d = descs[blockIdx.x];
  for (int n = threadIdx.x; n < d.n; n += blockDim.x) {
    const float v = d.bias[n];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const int kk = d.k & 63;
    const long long off = ((long long)(d.k >> 6) * d.rpad + n) * 128 + ((((kk >> 3) ^ (n & 7)) << 4)) + (kk & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(d.img + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(d.img + d.plane + off) = lo;
  }
}

template <int NPASS>
__global__ void __launch_bounds__(kS2Threads, 1)
flow_tc2_kernel(const FlowArgs a, const StackDims sd, const S2Dims d2, const unsigned char* __restrict__ w_img,
                const long long* __restrict__ w_off, unsigned char* __restrict__ act, unsigned int* __restrict__ dbg) {
  constexpr int PL = NPASS == 3 ? 2 : 1;
  extern __shared__ __align__(1024) unsigned char smem_s2[];
  unsigned char* glue = smem_s2 + d2.stg_off;                    // 32 KB: (t | s) rows and the y state of the tile
  unsigned char* misc = smem_s2 + d2.misc_off;
  uint64_t* full = reinterpret_cast<uint64_t*>(misc);            // [3] own TMA
  uint64_t* peer_full = full + 4;                                // [3] leader: the peer's stage has landed
  uint64_t* empty = peer_full + 4;                               // [3] multicast commit
  uint64_t* acc_full = empty + 4;                                // [2] multicast commit
  uint64_t* tmem_empty = acc_full + 2;                           // [2] leader: every epilogue warp of both CTAs drained the slot
  uint64_t* gpar_bar = tmem_empty + 2;                           // [1] glue parameters have landed
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(misc + 192);
  unsigned int* done_warps = reinterpret_cast<unsigned int*>(misc + 196);   // epilogue warps x output groups made visible
  int* gl_tab = reinterpret_cast<int*>(misc + 200);              // [1 + 2 * kS2GlueMax]: count, then (type, float offset) per op
  const float** prow_s = reinterpret_cast<const float**>(misc + 256);       // [128] projection row of each tile row

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const long long n_tiles = (a.n_rows + 2 * kS2Rows - 1) / (2 * kS2Rows);
  const long long n_pairs = gridDim.x >> 1, pair_id = blockIdx.x >> 1;
  const long long my_tiles = pair_id < n_tiles ? (n_tiles - pair_id + n_pairs - 1) / n_pairs : 0;
  unsigned char* act_cta = act + (long long)blockIdx.x * d2.cta_bytes;
  const long long buf_bytes = (long long)PL * d2.a_plane;
  const int L = d2.half[0].L;

  if (tid == 0) {
    for (int s = 0; s < kS2Stages; ++s) { mbar_init(&full[s], 1); mbar_init(&peer_full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&tmem_empty[b], 2 * kS2EpiWarps); }
    mbar_init(gpar_bar, 1);
    *done_warps = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);

  if (warp == 0) {
    // ===================== producer: own 128 rows of the activation image, own half of the weight tile ==============
    if (lane == 0) {
      uint32_t it = 0, job = 0, grp_base = 0;       // grp_base: output groups that precede the input of the current job
      for (long long ti = 0; ti < my_tiles; ++ti) {
        for (int oi = 0; oi < a.n_ops; ++oi) {
          const DevOp op = a.ops[oi];
          if (op.type != DOP_HALF) continue;
          const S2Half& hl = d2.half[op.src];
          const unsigned char* wnet = w_img + w_off[oi];
          for (int l = 0; l <= hl.L; ++l, ++job) {
            const S2Layer& ly = hl.layer[l];
            const unsigned char* abuf = act_cta + (long long)(job & 1u) * buf_bytes;
            const uint32_t in_groups = l == 0 ? 1u : (uint32_t)hl.layer[l - 1].n_chunks;
            int n0 = 0;
            for (int c = 0; c < ly.n_chunks; ++c) {
              const int cn = ly.chunk_n[c];
              const uint32_t b_bytes = (uint32_t)(cn >> 1) * 128u;
              const unsigned char* wsrc = wnet + ly.w_off + (long long)(n0 + (int)cta * (cn >> 1)) * 128;
              for (int kc = 0; kc < ly.n_kst; ++kc, ++it) {
                if (c == 0) {
                  // image chunk kc of the input was written by this CTA's own epilogue as part of output group
                  // that holds its last column: wait until all sixteen epilogue warps have made that group visible
                  uint32_t g_in = 0;        // output group of the previous layer that holds the last column of K stage kc
                  if (l > 0) {
                    const S2Layer& lp = hl.layer[l - 1];
                    int end = lp.chunk_n[0];
                    while (g_in + 1u < in_groups && end <= 64 * kc + 63) end += lp.chunk_n[++g_in];
                  }
                  const uint32_t need = (grp_base + g_in + 1u) * (uint32_t)kS2EpiWarps;
                  if (*reinterpret_cast<volatile unsigned int*>(done_warps) < need) {
                    const unsigned long long t0 = g2_now();
                    uint32_t n = 0;
                    while (*reinterpret_cast<volatile unsigned int*>(done_warps) < need)
                      if (((++n) & 1023u) == 0 && g2_now() - t0 > 4000000000ull) s2_timeout(dbg, 0x100u, need);
                  }
                  // (this thread only reads the data through the async proxy: the proxy fence orders those reads after
                  // the count; an acquire fence at GPU scope would also invalidate the SM's L1 at every K stage)
                  asm volatile("fence.proxy.async;" ::: "memory");
                  if (kc == 0) s2_stamp(a.trace, 2, job, 0);
                  if (kc == ly.n_kst - 1) s2_stamp(a.trace, 2, job, 1);
                }
                const int s = (int)(it % kS2Stages);
                const uint32_t use = it / kS2Stages;
                if (use > 0) s2_wait<false>(&empty[s], (use - 1) & 1u, dbg, 0x101u);
                unsigned char* st = smem_s2 + (size_t)s * d2.stage_bytes;
                mbar_expect_tx(&full[s], (uint32_t)PL * ((uint32_t)kS2Tile + b_bytes));
                tma_bulk_g2s(st, abuf + (long long)kc * kS2Tile, kS2Tile, &full[s]);
                tma_bulk_g2s(st + d2.b_off, wsrc + (long long)kc * ly.w_rpad * 128, b_bytes, &full[s]);
                if (NPASS == 3) {
                  tma_bulk_g2s(st + kS2Tile, abuf + d2.a_plane + (long long)kc * kS2Tile, kS2Tile, &full[s]);
                  tma_bulk_g2s(st + d2.b_lo_off, wsrc + ly.w_plane + (long long)kc * ly.w_rpad * 128, b_bytes, &full[s]);
                }
              }
              n0 += cn;
            }
            grp_base += in_groups;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (!leader) {
      // ===================== relay: tell the leader when this CTA's stage has landed =====================
      if (lane == 0) {
        long long per_tile = 0;
        for (int hi = 0; hi < d2.n_halfops; ++hi) {
          const S2Half& hl = d2.half[d2.two_way ? (hi & 1) : 0];
          for (int l = 0; l <= hl.L; ++l) per_tile += (long long)hl.layer[l].n_chunks * hl.layer[l].n_kst;
        }
        const long long total = per_tile * my_tiles;
        const uint32_t peer_bar0 = mapa_u32(smem_u32(&peer_full[0]), 0);
        for (long long it = 0; it < total; ++it) {
          const int s = (int)(it % kS2Stages);
          s2_wait<true>(&full[s], (uint32_t)(it / kS2Stages) & 1u, dbg, 0x200u);
          mbar_arrive_remote(peer_bar0 + 8u * (uint32_t)s);
        }
      }
    } else {
      // ===================== MMA issuer =====================
      // The whole warp walks the loops (every value below is warp-uniform, so the descriptors stay in uniform
      // registers); one elected lane issues the MMAs and the commits.
      const uint32_t st_addr = smem_u32(smem_s2);
      uint32_t it = 0, nchunk = 0;
      for (long long ti = 0; ti < my_tiles; ++ti)
        for (int hi = 0; hi < d2.n_halfops; ++hi) {
          const S2Half& hl = d2.half[d2.two_way ? (hi & 1) : 0];
          for (int l = 0; l <= hl.L; ++l) {
            const S2Layer& ly = hl.layer[l];
            for (int c = 0; c < ly.n_chunks; ++c, ++nchunk) {
              const uint32_t slot = nchunk & 1u;
              if (lane == 0) s2_stamp(a.trace, 0, nchunk, 0);
              if (nchunk >= 2) s2_wait<true>(&tmem_empty[slot], ((nchunk >> 1) - 1u) & 1u, dbg, 0x300u);
              tc_fence_after();
              if (lane == 0) s2_stamp(a.trace, 0, nchunk, 1);
              const uint32_t acc = tmem_base + slot * (uint32_t)kS2SlotCols;
              const uint32_t idesc = make_idesc_m256(ly.chunk_n[c]);
              for (int kc = 0; kc < ly.n_kst; ++kc, ++it) {
                const int s = (int)(it % kS2Stages);
                const uint32_t par = (it / kS2Stages) & 1u;
                s2_wait<true>(&full[s], par, dbg, 0x301u);
                s2_wait<true>(&peer_full[s], par, dbg, 0x302u);
                tc_fence_after();
                if (lane == 0 && kc == 0) s2_stamp(a.trace, 0, nchunk, 2);
                const int ksteps = kc == ly.n_kst - 1 ? ly.last_ksteps : 4;
                const uint32_t base = st_addr + (uint32_t)s * (uint32_t)d2.stage_bytes;
                const uint64_t ah = make_smem_desc(base), bh = make_smem_desc(base + (uint32_t)d2.b_off);
                const uint64_t al = make_smem_desc(base + kS2Tile), bl = make_smem_desc(base + (uint32_t)d2.b_lo_off);
                if (elect_one_sync()) {
                  for (int k = 0; k < ksteps; ++k) {
                    umma_2sm(acc, ah + 2 * k, bh + 2 * k, idesc, (kc | k) == 0 ? 0u : 1u);
                    if (NPASS == 3) {
                      umma_2sm(acc, al + 2 * k, bh + 2 * k, idesc, 1u);
                      umma_2sm(acc, ah + 2 * k, bl + 2 * k, idesc, 1u);
                    }
                  }
                  umma_commit_2sm(&empty[s], 3);
                  if (kc == ly.n_kst - 1) umma_commit_2sm(&acc_full[slot], 3);
                }
                __syncwarp();
                if (lane == 0 && kc == ly.n_kst - 1) s2_stamp(a.trace, 0, nchunk, 3);
              }
            }
          }
        }
    }
  } else if (warp >= kS2FirstEpi) {
    // ===================== epilogue warps ================================================================
    const int et = tid - 32 * kS2FirstEpi;         // 0..511
    const int q = warp & 3;                        // TMEM lane quarter of this warp
    const int part = (warp - kS2FirstEpi) >> 2;    // 4 warps share a quarter: column group / glue column class
    const int row = q * 32 + lane;                 // row of this CTA's 128 held by this thread's TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t tmem_empty_leader = mapa_u32(smem_u32(tmem_empty), 0);
    const int D = sd.D;
    float* ts_s = reinterpret_cast<float*>(glue);                            // [128][25]
    float* y_s = reinterpret_cast<float*>(glue + kS2Rows * kS2TsPitch * 4);  // [128][25]: the state y of the tile's rows
    const float* gpar_s = reinterpret_cast<const float*>(glue + kS2GparOff); // parameters of the next ActNorm / mixing run
    uint32_t gl_runs = 0;                          // parameter runs fetched so far (phase of gpar_bar)
    uint32_t nchunk = 0, job = 0;
    float ld_acc = 0.f;                            // log-det of this row (row owners: part == 0)
    // byte offset, inside a 128-byte image row, of the 32-byte sector holding this thread's two 16-byte units
    // (logical units 2*part, 2*part + 1, XOR-swizzled by row & 7: always one aligned pair, swapped when row is odd)
    const int sec_off = row * 128 + ((((part * 2) ^ (row & 7)) >> 1) << 5);
    uint32_t swap_u = (uint32_t)row & 1u;
    asm volatile("" : "+r"(swap_u));             // kept in a register (the compiler re-read %tid for it in every chunk)
    const bool sec_swap = swap_u != 0u;

    // every thread has finished its stores of one output group: make them visible, count the warp
    // (a release at GPU scope by one lane, cumulative over the warp through __syncwarp: the stores are acknowledged by
    // L2 before the count moves.  No acquire, so no L1 invalidation -- a full __threadfence() per lane cost 11 % of the
    // kernel's warp time in membar stalls and kept evicting the bias / mixing matrices from L1.  A CTA-scope release,
    // with or without a writer-side fence.proxy.async.global, measured the same within run-to-run noise (111.6 / 110.9 /
    // 111.8 ms per 500 000 rows): the GPU-scope form, whose L2 acknowledgement does not lean on request ordering, stays.)
    auto publish_group = [&]() {
      __syncwarp();
      if (d2.debug & 4) { if (lane == 0) atomicAdd(done_warps, 1u); return; }
      if (lane == 0) asm volatile("red.release.gpu.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(done_warps)) : "memory");
    };

    long long ti = 0;
    long long row_g = 0;                           // global row of this thread in the current tile
    long long row_base = 0;                        // first global row of this CTA's 128
    bool valid = false;
    int oi = 0;
    // The CTA's 128 rows of y are one contiguous range of the (n_rows, D) array: every thread moves the elements
    // e = et + 512 u of it (coalesced, all loads in flight at once) instead of a row owner walking its row.
    constexpr int kYPer = (kS2Rows * (kS2YPitch - 1) + kS2EpiThreads - 1) / kS2EpiThreads;   // 7

    // The ActNorm / mixing layers ops[first..) up to the next conditioner network: their parameters are one contiguous
    // range of the blob (common.cuh).  One thread reads the op table and starts a bulk copy of the range into shared
    // memory -- called a whole network ahead of the use, so that the dependent global reads (op table -> parameters:
    // 17-22 k cycles under a loaded L2, measured) are off the critical path.  Visible to the others after the next
    // epilogue barrier.
    auto fetch_glue = [&](int first) {
      if (et == 0) {
        int n = 0, k = first;
        long long off0 = 0;
        while (k < a.n_ops && n < kS2GlueMax) {
          const DevOp o = a.ops[k];
          if (o.type == DOP_HALF) break;
          if (n == 0) off0 = o.off;
          gl_tab[1 + 2 * n] = o.type;
          gl_tab[2 + 2 * n] = (int)(o.off - off0);
          ++n; ++k;
        }
        const long long off_end = k < a.n_ops ? a.ops[k].off : a.blob_floats;
        gl_tab[0] = n;
        if (n > 0) {
          const uint32_t bytes = (uint32_t)(off_end - off0) * 4u;
          mbar_expect_tx(gpar_bar, bytes);
          tma_bulk_g2s(glue + kS2GparOff, a.blob + off0, bytes, gpar_bar);
        }
      }
    };
    // load a fresh tile: y from the input, log-det 0, projection row pointers
    auto fresh_tile = [&]() {
      const long long tile = pair_id + ti * n_pairs;
      row_g = tile * (2 * kS2Rows) + (long long)cta * kS2Rows + row;
      valid = row_g < a.n_rows;
      row_base = row_g - row;
      const long long lim = (a.n_rows - row_base) * D;       // elements of this CTA's range that exist
#pragma unroll
      for (int u = 0; u < kYPer; ++u) {
        const int e = et + kS2EpiThreads * u;
        if (e < kS2Rows * D) {
          const int r = e / D;
          y_s[r * kS2YPitch + (e - r * D)] = e < lim ? flow_input(a, row_base + r, e - r * D, D) : 0.f;
        }
      }
      if (part == 0) {
        ld_acc = 0.f;
        prow_s[row] = a.P + (valid ? row_instance(a, row_g) : 0) * (long long)sd.PW;
      }
      fetch_glue(0);
      s2_epi_sync();
      oi = 0;
    };
    // apply the fetched run (cnf.py:333-354); y in y_s, all 512 threads
    auto glue_ops = [&]() {
      const int n = gl_tab[0];
      if (n > 0) { s2_wait<false>(gpar_bar, gl_runs & 1u, dbg, 0x503u); ++gl_runs; }
      else s2_epi_sync();          // (everyone has read the table before the next fetch rewrites it)
      for (int g = 0; g < n; ++g, ++oi) {
        const int type = gl_tab[1 + 2 * g];
        const float* w = gpar_s + gl_tab[2 + 2 * g];
        if (type == DOP_MIX) {
          float o[6];
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int j = part + 4 * u;
            float s = 0.f;
            if (j < D) {
#pragma unroll 4
              for (int i = 0; i < D; ++i) s = fmaf(y_s[row * kS2YPitch + i], w[i * sd.DP + j], s);   // y @ M
            }
            o[u] = s;
          }
          s2_epi_sync();
#pragma unroll
          for (int u = 0; u < 6; ++u) { const int j = part + 4 * u; if (j < D) y_s[row * kS2YPitch + j] = o[u]; }
        } else {
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int j = part + 4 * u;
            if (j < D) {
              const float s = w[j], b = w[sd.DP + j], y = y_s[row * kS2YPitch + j];
              y_s[row * kS2YPitch + j] = type == DOP_ACTNORM_FWD ? fmaf(s, y, b) : __fdiv_rn(y - b, s);
            }
          }
          if (part == 0) ld_acc += w[2 * sd.DP];
        }
        s2_epi_sync();
      }
    };
    // y_s -> output (coalesced), log-det
    auto store_result = [&]() {
      const long long lim = (a.n_rows - row_base) * D;
#pragma unroll
      for (int u = 0; u < kYPer; ++u) {
        const int e = et + kS2EpiThreads * u;
        if (e < kS2Rows * D && e < lim) { const int r = e / D; flow_output(a, row_base + r, e - r * D, D, y_s[r * kS2YPitch + (e - r * D)]); }
      }
      if (part == 0 && valid && a.logdet) a.logdet[row_g] = ld_acc;
    };
    // own-half input of the conditioner network ops[oi], zero padded to one K = 16 step, as the first 32-byte sector
    // of image chunk 0 of the buffer its first Linear reads (the MMA of that layer reads nothing else of the chunk)
    auto write_x_in = [&]() {
      if (part == 0) {
        const DevOp op = a.ops[oi];
        const HalfLayout& hl = sd.half[op.src];
        const int in0 = op.src == 0 ? 0 : sd.Da;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v0 = 2 * i < hl.din ? y_s[row * kS2YPitch + in0 + 2 * i] : 0.f;
          const float v1 = 2 * i + 1 < hl.din ? y_s[row * kS2YPitch + in0 + 2 * i + 1] : 0.f;
          hi[i] = pack_bf16x2(v0, v1);
          const float h0 = __uint_as_float(hi[i] << 16), h1 = __uint_as_float(hi[i] & 0xffff0000u);
          lo[i] = pack_bf16x2(v0 - h0, v1 - h1);
        }
        unsigned char* dst = act_cta + (long long)(job & 1u) * buf_bytes + sec_off;     // part == 0: units 0, 1
        st_global_sector(dst, hi, sec_swap);
        if (NPASS == 3) st_global_sector(dst + d2.a_plane, lo, sec_swap);
      }
      publish_group();
    };

    if (my_tiles > 0) {
      fresh_tile();
      glue_ops();
      for (;;) {
        // ---- here ops[oi] is a conditioner network; y_s holds the current state ----
        fetch_glue(oi + 1);             // (read after the barriers of this network's last Linear)
        write_x_in();
        const DevOp op = a.ops[oi];
        const float* w = a.blob + op.off;
        const HalfLayout& hl = sd.half[op.src];
        const S2Half& tl = d2.half[op.src];
        const int out0 = op.src == 0 ? sd.Da : 0;

        // ---- hidden layers: TMEM -> (+P | +bias) -> GELU -> bf16 hi/lo, one 32-byte sector per plane and 16 columns,
        //      stored straight into the L2-resident image of the next layer's input.  No barrier inside a layer: the
        //      warps only meet at the accumulator hand-offs.
        for (int l = 0; l < L; ++l, ++job) {
          const S2Layer& ly = tl.layer[l];
          unsigned char* obuf = act_cta + (long long)((job + 1u) & 1u) * buf_bytes + sec_off;
          const float* add = l == 0 ? prow_s[row] + op.proj_off : w + hl.off_b[l];
          int n0 = 0;
          for (int c = 0; c < ly.n_chunks; ++c, ++nchunk) {
            const int cn = ly.chunk_n[c];
            const uint32_t slot = nchunk & 1u;
            if (et == 0) s2_stamp(a.trace, 1, nchunk, 0);
            const int n_ic = (cn + 63) >> 6;
            const uint32_t tcol = lane_addr + slot * (uint32_t)kS2SlotCols + (uint32_t)(part * 16);
            // column of this thread's 16 that the epilogue replaces by the constant 1 of the next layer's folded bias
            // (warp-uniform test: one 16-column group of the layer at most)
            if (l > 0 && ly.bias_k >= 0) {
              // ---- bias already in the accumulator: TMEM -> GELU -> split -> store, the next TMEM load in flight ----
              s2_wait<false>(&acc_full[slot], (nchunk >> 1) & 1u, dbg, 0x501u);
              tc_fence_after();
              if (et == 0) s2_stamp(a.trace, 1, nchunk, 1);
              // (tcgen05.ld takes a dozen cycles: the load of the next 16 columns is issued into the same registers once
              // the arithmetic has consumed them, and completes behind the two sector stores; a second register set
              // held across the whole GELU made the compiler serialise half of the Horner chains)
              uint32_t r[16];
              if (part * 16 < cn && !(d2.debug & 8)) tmem_ld16_issue(tcol, r);
              unsigned char* dst = obuf + (long long)(n0 >> 6) * kS2Tile;
              asm volatile("" : "+l"(dst));          // a live pointer, not re-derived from the kernel parameters per chunk
              for (int ic = 0; ic < n_ic; ++ic, dst += kS2Tile) {
                const int colc = ic * 64 + part * 16;
                const bool more = ic + 1 < n_ic && colc + 64 < cn;
                if (colc < cn) {
                  if (!(d2.debug & 8)) tmem_ld16_wait(r);
                  uint32_t hw[8], lw[8];
                  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (d2.debug & 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { hw[i] = r[i]; lw[i] = r[8 + i]; }
                  } else {
                    s2_gelu_pack8<NPASS, false, 0>(r, z4, z4, hw, lw);
                    s2_gelu_pack8<NPASS, false, 4>(r + 8, z4, z4, hw, lw);
                  }
                  if (more && !(d2.debug & 8)) tmem_ld16_issue(tcol + (uint32_t)(ic + 1) * 64u, r);
                  if (!(d2.debug & 2)) {
                  st_global_sector(dst, hw, sec_swap);
                  if (NPASS == 3) st_global_sector(dst + d2.a_plane, lw, sec_swap);
                  }
                  s2_store_one<NPASS>(ly.one_col - (n0 + colc), dst - sec_off, d2.a_plane, row, part);
                }
              }
            } else {
              // ---- add vector (the row's projection slice for the first Linear: an L2 / HBM read; else the bias) fetched
              //      one image chunk ahead of its use, the first one before the wait for the accumulator ----
              auto load_add = [&](int ic, float4 (&b)[4]) {
                const int colc = ic * 64 + part * 16;
                if (ic < n_ic && colc < cn) {
                  const int n = n0 + colc;
                  if (l == 0) {
                    ldg_stream8(add + n, b[0], b[1]);
                    ldg_stream8(add + n + 8, b[2], b[3]);
                  } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) b[u] = __ldg(reinterpret_cast<const float4*>(add + n + 4 * u));
                  }
                }
              };
              float4 b[4];
              load_add(0, b);
              s2_wait<false>(&acc_full[slot], (nchunk >> 1) & 1u, dbg, 0x501u);
              tc_fence_after();
              if (et == 0) s2_stamp(a.trace, 1, nchunk, 1);
              unsigned char* dst = obuf + (long long)(n0 >> 6) * kS2Tile;
              asm volatile("" : "+l"(dst));
              for (int ic = 0; ic < n_ic; ++ic, dst += kS2Tile) {
                const int colc = ic * 64 + part * 16;
                float4 bn[4];
                load_add(ic + 1, bn);
                if (colc < cn) {
                  uint32_t r[16];
                  tmem_ld16_issue(tcol + (uint32_t)ic * 64u, r);
                  tmem_ld16_wait(r);
                  uint32_t hw[8], lw[8];
                  s2_gelu_pack8<NPASS, true, 0>(r, b[0], b[1], hw, lw);
                  s2_gelu_pack8<NPASS, true, 4>(r + 8, b[2], b[3], hw, lw);
                  st_global_sector(dst, hw, sec_swap);
                  if (NPASS == 3) st_global_sector(dst + d2.a_plane, lw, sec_swap);
                  s2_store_one<NPASS>(ly.one_col - (n0 + colc), dst - sec_off, d2.a_plane, row, part);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) b[u] = bn[u];
              }
            }
            if (c == ly.n_chunks - 1 && !(d2.debug & 16)) {
              // every MMA of this layer has completed (its last accumulator is what we have just drained): the input
              // image is dead -- the four threads of a row drop the row's line of every fourth chunk each (ordered
              // before the next writes to these lines by the release below and the accumulator hand-off of the layer
              // after next; one thread per row doing all of them put 2-3 k cycles into the layer's critical path)
              const unsigned char* ibuf = act_cta + (long long)(job & 1u) * buf_bytes + row * 128;
              const int in_img = l == 0 ? 1 : tl.layer[l - 1].n_img;
              for (int kc = part; kc < in_img; kc += 4) {
                discard_line(ibuf + (long long)kc * kS2Tile);
                if (NPASS == 3) discard_line(ibuf + d2.a_plane + (long long)kc * kS2Tile);
              }
            }
            // slot drained, group written: release the accumulator, make the group visible to the producer
            tc_fence_before();
            publish_group();
            if (lane == 0) mbar_arrive_remote(tmem_empty_leader + 8u * slot);
            if (et == 0) s2_stamp(a.trace, 1, nchunk, 2);
            n0 += cn;
          }
        }

        // ---- last Linear: (t | s) from TMEM, affine update, log-det, glue, next network's input ----
        {
          const uint32_t slot = nchunk & 1u;
          if (et == 0) s2_stamp(a.trace, 1, nchunk, 0);
          s2_wait<false>(&acc_full[slot], (nchunk >> 1) & 1u, dbg, 0x502u);
          tc_fence_after();
          if (et == 0) s2_stamp(a.trace, 1, nchunk, 1);
          if (!(d2.debug & 16)) {          // the last Linear's input image is dead too
            const unsigned char* ibuf = act_cta + (long long)(job & 1u) * buf_bytes + row * 128;
            for (int kc = part; kc < tl.layer[L - 1].n_img; kc += 4) {
              discard_line(ibuf + (long long)kc * kS2Tile);
              if (NPASS == 3) discard_line(ibuf + d2.a_plane + (long long)kc * kS2Tile);
            }
          }
          if (part == 0) {
            uint32_t r0[16], r1[16];
            tmem_ld16_issue(lane_addr + slot * (uint32_t)kS2SlotCols, r0);
            tmem_ld16_issue(lane_addr + slot * (uint32_t)kS2SlotCols + 16u, r1);
            tmem_ld_wait();
            const float* bo = w + hl.off_bout;
            const bool addb = tl.layer[L].bias_k < 0;      // else the bias came with the accumulator
#pragma unroll
            for (int c = 0; c < 16; ++c) {       // (a row of ts_s holds 2 * dop <= 24 columns: nothing beyond them is written)
              if (c < 2 * hl.dop) ts_s[row * kS2TsPitch + c] = __uint_as_float(r0[c]) + (addb ? __ldg(bo + c) : 0.f);
              if (16 + c < 2 * hl.dop) ts_s[row * kS2TsPitch + 16 + c] = __uint_as_float(r1[c]) + (addb ? __ldg(bo + 16 + c) : 0.f);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(tmem_empty_leader + 8u * slot);
          if (et == 0) s2_stamp(a.trace, 3, 2u * (job / (uint32_t)(L + 1)), 0);
          s2_epi_sync();
          if (et == 0) s2_stamp(a.trace, 3, 2u * (job / (uint32_t)(L + 1)), 1);
          ++nchunk; ++job;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int m = part + 4 * u;
            if (m < hl.dout) {
              const float t = ts_s[row * kS2TsPitch + m];
              const float ls = tanhf(ts_s[row * kS2TsPitch + hl.dop + m]);                          // cnf.py:107
              const float yd = y_s[row * kS2YPitch + out0 + m];
              y_s[row * kS2YPitch + out0 + m] = op.inverse ? (yd - t) * expf(-ls) : fmaf(expf(ls), yd, t);   // cnf.py:204 / :179
              ts_s[row * kS2TsPitch + hl.dop + m] = ls;
            }
          }
          if (et == 0) s2_stamp(a.trace, 3, 2u * ((job - 1u) / (uint32_t)(L + 1)), 2);
          s2_epi_sync();
          if (et == 0) s2_stamp(a.trace, 3, 2u * ((job - 1u) / (uint32_t)(L + 1)), 3);
          if (part == 0) {
            float s = 0.f;
            for (int m = 0; m < hl.dout; ++m) s += ts_s[row * kS2TsPitch + hl.dop + m];             // cnf.py:190, fixed order
            ld_acc += s;
          }
        }
        ++oi;
        if (et == 0) s2_stamp(a.trace, 3, 2u * ((job - 1u) / (uint32_t)(L + 1)) + 1u, 0);
        glue_ops();
        if (et == 0) s2_stamp(a.trace, 3, 2u * ((job - 1u) / (uint32_t)(L + 1)) + 1u, 1);
        if (et == 0) s2_stamp(a.trace, 1, nchunk - 1u, 2);
        if (oi == a.n_ops) {
          store_result();
          ++ti;
          if (ti == my_tiles) break;
          s2_epi_sync();                  // every thread is done with y_s before the next tile overwrites it
          fresh_tile();
          glue_ops();
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

}  // namespace bcnf
