// Row-tile fused coupling stack on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Same scope as the other flow kernels (reference cnf.py:479-488, :500-506 and callees) for
// conditioners wide enough to be real GEMMs.  A CTA PAIR (cluster of 2, cta_group::2) owns 128
// rows -- 64 per CTA -- from the first layer to the last:
//
//   * activations of the current layer live in shared memory as bf16 K-major SWIZZLE_128B tiles
//     (one 64-row x 64-column tile per 128 bytes x 64 rows), as a hi part and, in the 3-pass
//     mode, a lo part (x = hi + lo to ~16 mantissa bits);
//   * weights are pre-split into hi/lo bf16 and pre-swizzled at pack time into exactly the
//     shared-memory tile images the MMA wants, so the producer warp streams them with plain
//     TMA bulk copies (cp.async.bulk + mbarrier) through a ring of stages; each CTA of the pair
//     loads its half of the N rows of every tile, so a weight byte crosses L2->SM once per 128 rows;
//   * one warp per N chunk of the leader CTA issues tcgen05.mma.cta_group::2 (M = 128 over the pair,
//     N <= 256 per instruction, K = 16; a single thread needs ~106 cycles per issue, more than such an MMA
//     lasts), accumulating fp32 in TMEM: 64 rows x N per CTA = N/2 columns; the accumulators of successive
//     layers rotate through a ring of TMEM slots;
//   * sixteen epilogue warps per CTA read TMEM (tcgen05.ld), add the bias (or the hoisted condition
//     projection P for the first layer), apply exact-erf GELU in fp32 (packed f32x2 arithmetic), split to
//     bf16 hi/lo and write the next layer's activation tiles in place, handing them to the issuers one
//     N chunk at a time so that the next layer's MMAs overlap the rest of the epilogue;
//   * the last Linear (N = 2 x 16) leaves t and s in TMEM; the row-owner threads apply
//     tanh/exp, the affine update, the log-det row sum, ActNorm and the orthonormal mixing in
//     fp32 on y kept in shared memory.
// Cross-CTA hand-offs are mbarrier arrivals with CTA-scope semantics (see mbar_arrive_remote): what they
// announce is only read by the arriving CTA's own half of the 2-CTA MMA.
//
// Arithmetic: NPASS = 3 computes a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (fp32-class accuracy),
// NPASS = 1 only a_hi*w_hi (plain bf16 inputs).  Accumulation, bias, GELU, tanh, exp, the
// affine update and the log-det are fp32 in both modes.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "img_store.cuh"   // pack_bf16x2
#include "flow_rowthread.cuh"   // mbarrier / bulk-copy helpers

namespace bcnf {

constexpr int kTcEpiWarps = 16;
constexpr int kTcEpiThreads = kTcEpiWarps * 32;
constexpr int kTcIssuers = 4;                    // one MMA-issuing warp per N chunk (<= 4 chunks of <= 256 columns)
constexpr int kTcFirstEpiWarp = 1 + kTcIssuers;
// warp 0 producer; warps 1..4 MMA issuers (leader CTA; warp 1 of the peer relays stage arrivals); then epilogue
constexpr int kTcThreads = 32 * kTcFirstEpiWarp + kTcEpiThreads;
constexpr int kTcRows = 64;             // rows per CTA (128 per CTA pair)
constexpr int kTcATile = kTcRows * 128; // bytes of one 64-row x 64-col bf16 activation tile
constexpr int kTcMaxLayers = BCNF_TC_MAX_LAYERS;

struct TcLayer {
  int np;            // padded N (multiple of 16)
  int n_chunks;      // N is processed in chunks of <= 256 columns
  int chunk_n[4];
  int kc;            // number of K chunks of the WEIGHT tiles (TcDims::kw columns each)
  int last_ksteps;   // K=16 steps that carry data in the last chunk (1..kw/16)
};

struct TcHalfLayout {
  int L;                       // hidden layers; layer[0] first Linear, [1..L-1] hidden, [L] last Linear
  int doh;                     // padded width of t (and of s) in the last Linear: N = 2 * doh
  TcLayer layer[kTcMaxLayers];
  long long stream_bytes;      // bytes of the weight tile stream of one conditioner network
};

struct TcDims {
  TcHalfLayout half[2];
  int kw;              // K columns per weight tile: 64 (SWIZZLE_128B)
  int a_chunks;        // 64-column activation tiles per part
  int stage_bytes;     // bytes of one weight stage (hi [+ lo] of the largest half tile)
  int n_stages;
  int off_alo, off_stage, off_y, off_ts, off_misc;   // shared-memory carve-up (bytes)
  int yp, tsp;         // pitches (floats) of y_s and ts_s
  int smem_bytes;
  int regions;         // TMEM is cut into `regions` slots of `region_cols` columns; the accumulator of N chunk c of the
  int region_cols;     // i-th layer executed lives in slot (sum of chunks of earlier layers + c) mod regions, so that
                       // the next layer's MMAs can start while this layer's accumulators are still being drained
  int n_halfops;       // conditioner networks per pass over the stack (same in both directions)
  int two_way;         // their nn_a / nn_b pattern: a,b,a,b,... (two_way) or a,a,a,... -- kept in the kernel
                       // parameters so that the MMA issuers never depend on values loaded from global memory
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  // default semantics (.release.cta), as CUTLASS' ClusterBarrier::arrive(cta_id): the data the arrival
  // announces lives in the arriving CTA's own shared memory and is only ever read there (by its half of
  // the 2-CTA MMA), after this thread's fence.proxy.async + CTA barrier; a cluster-scope release costs a
  // few thousand cycles per hand-off (measured) and orders nothing more that matters here
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITC_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONEC_%=;\n"
      "bra WAITC_%=;\n"
      "DONEC_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// spin on test_wait (no hardware suspend): lowest wake-up latency, for the single hot hand-off per layer
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory"); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B: 8-row x 128-byte atoms, 1024 bytes apart (cute::UMMA::SmemDescriptor)
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
  d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  // kind::f16, A = B = bf16 (1), D = f32 (1), both K-major, M = 128 over the CTA pair
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  // arrives (once the MMAs issued so far by this thread retire) on the barrier at this offset in every CTA of the mask
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// issue a TMEM load of 8 consecutive columns of this thread's lane (no wait)
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  tmem_ld8_issue(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}


// store 8 consecutive activations (columns n0..n0+7 of `row`) as bf16 hi (and lo) into the tiles
template <int NPASS>
__device__ __forceinline__ void store_act8(unsigned char* a_hi, unsigned char* a_lo, int row, int n0, const float (&v)[8]) {
  const int tile = n0 >> 6, c16 = (n0 & 63) >> 3;
  const int off = tile * kTcATile + row * 128 + ((c16 ^ (row & 7)) << 4);
  float hi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) hi[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
  *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]),
                                                     pack_bf16x2(hi[4], hi[5]), pack_bf16x2(hi[6], hi[7]));
  if (NPASS == 3)
    *reinterpret_cast<uint4*>(a_lo + off) =
        make_uint4(pack_bf16x2(v[0] - hi[0], v[1] - hi[1]), pack_bf16x2(v[2] - hi[2], v[3] - hi[3]),
                   pack_bf16x2(v[4] - hi[4], v[5] - hi[5]), pack_bf16x2(v[6] - hi[6], v[7] - hi[7]));
}

// (+ add) -> GELU -> bf16 hi/lo split of 8 accumulator columns, packed two values per instruction, stored as
// one 16-byte chunk per part into the swizzled activation tiles
template <int NPASS>
__device__ __forceinline__ void gelu_store8(unsigned char* a_hi, unsigned char* a_lo, int row, int n0,
                                            const uint32_t (&r)[8], const float4& b0, const float4& b1) {
  const int tile = n0 >> 6, c16 = (n0 & 63) >> 3;
  const int off = tile * kTcATile + row * 128 + ((c16 ^ (row & 7)) << 4);
  f32x2 v[4];
  v[0] = gelu_erf_fast2(add2(pack2(__uint_as_float(r[0]), __uint_as_float(r[1])), pack2(b0.x, b0.y)));
  v[1] = gelu_erf_fast2(add2(pack2(__uint_as_float(r[2]), __uint_as_float(r[3])), pack2(b0.z, b0.w)));
  v[2] = gelu_erf_fast2(add2(pack2(__uint_as_float(r[4]), __uint_as_float(r[5])), pack2(b1.x, b1.y)));
  v[3] = gelu_erf_fast2(add2(pack2(__uint_as_float(r[6]), __uint_as_float(r[7])), pack2(b1.z, b1.w)));
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float x0, x1;
    unpack2(v[i], x0, x1);
    hi[i] = pack_bf16x2(x0, x1);                                        // cvt.rn.bf16x2.f32
    if (NPASS == 3) {
      const f32x2 h = pack2(__uint_as_float(hi[i] << 16), __uint_as_float(hi[i] & 0xffff0000u));
      float l0, l1;
      unpack2(add2(v[i], h ^ 0x8000000080000000ull), l0, l1);           // v - hi, both lanes
      lo[i] = pack_bf16x2(l0, l1);
    }
  }
  *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (NPASS == 3) *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

template <int NPASS>
__global__ void __launch_bounds__(kTcThreads, 1)
flow_tc_kernel(const FlowArgs a, const StackDims sd, const TcDims td, const unsigned char* __restrict__ tc_blob,
               const long long* __restrict__ tc_off) {
  extern __shared__ __align__(1024) unsigned char smem_tc[];
  unsigned char* a_hi = smem_tc;
  unsigned char* a_lo = smem_tc + td.off_alo;
  unsigned char* stage0 = smem_tc + td.off_stage;
  float* y_s = reinterpret_cast<float*>(smem_tc + td.off_y);
  float* ts_s = reinterpret_cast<float*>(smem_tc + td.off_ts);
  unsigned char* misc = smem_tc + td.off_misc;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(misc);              // [8]
  uint64_t* w_peer = w_full + 8;                                     // [8] leader only
  uint64_t* w_empty = w_peer + 8;                                    // [8]
  uint64_t* acc_full = w_empty + 8;                                  // [4] one per N chunk / issuer
  uint64_t* a_ready = acc_full + 4;                                  // [1] leader only: input of a first Linear written
  uint64_t* a_chunk = a_ready + 1;                                   // [4] leader only: N chunk c of a hidden layer's output written
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(a_chunk + 4);
  float* ld_s = reinterpret_cast<float*>(tmem_ptr_s + 2);            // [64]
  const float** prow_s = reinterpret_cast<const float**>(ld_s + kTcRows);   // [64]

  // warp index and TMEM base as lane-0 broadcasts: provably warp-uniform, so ptxas keeps the MMA operands in
  // uniform registers instead of wrapping every tcgen05.mma in a uniformisation loop (tools/mma_probe.cu:
  // 139 -> 106 cycles per issue)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t cta = cluster_ctarank();         // rank in the CTA pair
  const uint32_t prank = cta & 1u, lead_rank = 0u;
  const bool leader = prank == 0;                 // rank 0 of the pair issues the MMAs
  const int n_stages = td.n_stages;
  const long long n_tiles = (a.n_rows + 2 * kTcRows - 1) / (2 * kTcRows);
  const long long n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  // iterations whose tile index is past the end run on padding rows and store nothing
  const long long n_iter = (n_tiles + n_clusters - 1) / n_clusters;
  const uint16_t mask_all = 3, mask_pair = 3;

  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_peer[s], 1); mbar_init(&w_empty[s], 1); }
    for (int j = 0; j < kTcIssuers; ++j) mbar_init(&acc_full[j], 1);
    mbar_init(a_ready, 2);
    for (int j = 0; j < kTcIssuers; ++j) mbar_init(&a_chunk[j], 2);
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
    // ===================== weight producer (each CTA streams its half of every tile) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long iter = 0; iter < n_iter; ++iter) {
        for (int oi = 0; oi < a.n_ops; ++oi) {
          const DevOp op = a.ops[oi];
          if (op.type != DOP_HALF) continue;
          const TcHalfLayout& hl = td.half[op.src];
          const unsigned char* src = tc_blob + tc_off[oi];
          for (int l = 0; l <= hl.L; ++l) {
            const TcLayer& ly = hl.layer[l];
            // K-major tile order: (kc, nc); issuer nc consumes every n_chunks-th tile
            for (int kc = 0; kc < ly.kc; ++kc) {
              for (int nc = 0; nc < ly.n_chunks; ++nc, ++it) {
                const uint32_t rows_b = (uint32_t)(ly.chunk_n[nc] >> 1) * (uint32_t)(2 * td.kw);
                const int s = it % n_stages;
                const uint32_t use = it / n_stages;
                if (use > 0) mbar_wait_cluster(&w_empty[s], (use - 1) & 1);
                const uint32_t bytes = rows_b * (NPASS == 3 ? 2u : 1u);
                mbar_expect_tx(&w_full[s], bytes);     // every CTA arms its own barrier for every tile
                unsigned char* dst = stage0 + (size_t)s * td.stage_bytes;
                const unsigned char* half = src + (size_t)prank * 2u * rows_b;
                tma_bulk_g2s(dst, half, bytes, &w_full[s]);
                src += 4u * rows_b;
              }
            }
          }
        }
      }
    }
  } else if (warp < kTcFirstEpiWarp) {
    if (!leader) {
      // ===================== relay: tell the leader when this CTA's half of a stage has landed ===========
      // one lane per stage, so the waits of different stages overlap instead of serialising
      if (warp == 1 && lane < n_stages) {
        long long per_tile = 0;
        for (int oi = 0; oi < a.n_ops; ++oi) {
          const DevOp op = a.ops[oi];
          if (op.type != DOP_HALF) continue;
          const TcHalfLayout& hl = td.half[op.src];
          for (int l = 0; l <= hl.L; ++l) per_tile += (long long)hl.layer[l].n_chunks * hl.layer[l].kc;
        }
        const long long total = per_tile * n_iter;
        const uint32_t peer_bar = mapa_u32(smem_u32(&w_peer[lane]), lead_rank);
        uint32_t use = 0;
        for (long long it = lane; it < total; it += n_stages, ++use) {
          mbar_wait(&w_full[lane], use & 1);
          mbar_arrive_remote(peer_bar);
        }
      }
    } else if (lane == 0) {
      // ===================== MMA issuers (leader CTA): warp 1+j issues every MMA of N chunk j ==============
      // Several issuing threads keep the tensor pipe fed while each one waits on barriers and builds
      // descriptors; chunks write disjoint TMEM columns, so their relative order does not matter.
      const int j = warp - 1;
      const uint32_t a_hi_addr = smem_u32(a_hi), a_lo_addr = smem_u32(a_lo), st_addr = smem_u32(stage0);
      // Hand-offs from the epilogue: `a_ready` once per conditioner network (input of its first Linear), and
      // `a_chunk[c]` once per hidden layer and N chunk c of its output.  Each barrier completes at most once
      // between two waits of every issuer (the next completion needs this layer's MMAs), so parities never alias.
      uint32_t xin_use = 0, chunk_use[kTcIssuers] = {0, 0, 0, 0}, lay_cnt = 0;
      int s = 0;            // ring position of the next tile of the stream (all chunks)
      uint32_t par = 0;
      uint32_t reg0 = 0;    // TMEM slot of chunk 0 of the current layer
      for (long long iter = 0; iter < n_iter; ++iter)
        for (int hi = 0; hi < td.n_halfops; ++hi) {
          const TcHalfLayout& hl = td.half[td.two_way ? (hi & 1) : 0];
          int n_prev = 1;                // hand-offs that make up the input of the current layer (1 for the first Linear)
          for (int l = 0; l <= hl.L; ++l) {
            const TcLayer& ly = hl.layer[l];
            const int nch = ly.n_chunks;
            ++lay_cnt;
            int got = 0;                       // hand-offs of this layer's input consumed so far
            auto wait_inputs = [&](int target) {
              while (got < target) {
                if (l == 0) { mbar_wait(a_ready, xin_use & 1); ++xin_use; }
                else { mbar_wait(&a_chunk[got], chunk_use[got] & 1); ++chunk_use[got]; }
                ++got;
              }
            };
            if (j >= nch) {
              // not my layer: nothing to wait for -- only keep the hand-off, ring and slot counters in step (an idle
              // issuer must not sit on a parity wait while the barrier can complete twice)
              if (l == 0) ++xin_use;
              else for (int c = 0; c < n_prev; ++c) ++chunk_use[c];
              int adv = nch * ly.kc;
              while (adv > 0) { const int d = adv < n_stages - s ? adv : n_stages - s; s += d; adv -= d; if (s == n_stages) { s = 0; par ^= 1; } }
            } else {
              // my accumulator slot; it may still hold chunk c_shared of the previous layer (not yet drained)
              const uint32_t col = ((reg0 + (uint32_t)j) % (uint32_t)td.regions) * (uint32_t)td.region_cols;
              const int c_shared = l == 0 ? -1 : n_prev + j - td.regions;
              const TcLayer& lp = hl.layer[l > 0 ? l - 1 : 0];
              const bool tr = a.trace && j == 0 && blockIdx.x == 0 && iter == 0 && lay_cnt <= 32;
              const int cn = ly.chunk_n[j];
              const uint32_t idesc = make_idesc(cn);
              const uint32_t rows_b = (uint32_t)(cn >> 1) * (uint32_t)(2 * td.kw);
              const int steps_per_tile = td.kw >> 4;
              for (int kc = 0; kc < ly.kc; ++kc) {
                // phases needed: every output chunk of the previous layer that overlaps A columns of this K tile
                int need = 0;
                if (l > 0) {
                  const int col_end = min((kc + 1) * td.kw, lp.np);
                  int pre = 0;
                  for (int c = 0; c < n_prev; ++c) { pre += lp.chunk_n[c]; if (pre >= col_end) { need = c; break; } need = c; }
                  if (kc == 0 && c_shared > need) need = c_shared;
                }
                wait_inputs(need + 1);
                tc_fence_after();
                if (tr && kc == 0) a.trace[(lay_cnt - 1) * 8 + 0] = clock64();
                // my tile of this K step sits j positions further in the ring
                int sj = s + j; uint32_t pj = par;
                while (sj >= n_stages) { sj -= n_stages; pj ^= 1; }
                mbar_wait(&w_full[sj], pj);
                mbar_wait_cluster(&w_peer[sj], pj);
                tc_fence_after();
                if (tr && kc == 0) a.trace[(lay_cnt - 1) * 8 + 1] = clock64();
                const int ksteps = kc == ly.kc - 1 ? ly.last_ksteps : steps_per_tile;
                const int g0 = kc * steps_per_tile;                 // first K=16 step of this weight tile
                const uint64_t ah = make_smem_desc(a_hi_addr + (g0 >> 2) * kTcATile) + 2 * (g0 & 3);
                const uint64_t al = make_smem_desc(a_lo_addr + (g0 >> 2) * kTcATile) + 2 * (g0 & 3);
                const uint64_t wh = make_smem_desc(st_addr + sj * td.stage_bytes);
                const uint64_t wl = make_smem_desc(st_addr + sj * td.stage_bytes + rows_b);
                for (int k = 0; k < ksteps; ++k) {
                  const uint32_t first = (kc | k) == 0 ? 0u : 1u;
                  umma_2sm(tmem_base + col, ah + 2 * k, wh + 2 * k, idesc, first);
                  if (NPASS == 3) {
                    umma_2sm(tmem_base + col, al + 2 * k, wh + 2 * k, idesc, 1u);
                    umma_2sm(tmem_base + col, ah + 2 * k, wl + 2 * k, idesc, 1u);
                  }
                }
                umma_commit_2sm(&w_empty[sj], mask_all);   // frees the stage in both CTAs
                s += nch;
                while (s >= n_stages) { s -= n_stages; par ^= 1; }
              }
              wait_inputs(n_prev);                         // (already true after the last K tile)
              umma_commit_2sm(&acc_full[j], mask_pair);    // chunk j of the layer output complete in TMEM of both CTAs
              if (tr) a.trace[(lay_cnt - 1) * 8 + 2] = clock64();
            }
            reg0 = (reg0 + (uint32_t)nch) % (uint32_t)td.regions;
            n_prev = nch;                // the epilogue of this layer hands over one N chunk at a time
          }
        }
    }
  } else {
    // ===================== epilogue warps ================================================================
    const int et = tid - 32 * kTcFirstEpiWarp;     // 0..kTcEpiThreads-1
    const int q = warp & 3;                        // TMEM lane quarter this warp may touch
    const int part = (warp - kTcFirstEpiWarp) >> 2;   // 4 warps share a quarter: each takes a quarter of the column groups
    const int row = ((q & 1) << 5) + lane;         // row of this CTA held by this thread's TMEM lane
    const int nhalf = q >> 1;                      // 2x2 layout: lanes 64..127 hold the second N half of a chunk
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int D = sd.D;
    uint32_t reg0 = 0;                             // TMEM slot of chunk 0 of the current layer (same walk as the issuers)
    uint32_t acc_cnt = 0;                          // layers seen (trace index)
    uint32_t acc_use[kTcIssuers] = {0, 0, 0, 0};   // phase counters of the per-chunk accumulator barriers
    const uint32_t a_ready_leader = mapa_u32(smem_u32(a_ready), lead_rank);
    const uint32_t a_chunk_leader = mapa_u32(smem_u32(a_chunk), lead_rank);

    for (long long iter = 0; iter < n_iter; ++iter) {
      const long long tile = iter * n_clusters + cluster_id;   // >= n_tiles: padding iteration
      const long long row0 = tile * (2 * kTcRows) + (long long)prank * kTcRows;
      if (et < kTcRows) {
        const long long r = row0 + et;
        const bool valid = r < a.n_rows;
        for (int j = 0; j < td.yp; ++j) y_s[et * td.yp + j] = (valid && j < D) ? flow_input(a, r, j, D) : 0.f;
        ld_s[et] = 0.f;
        prow_s[et] = a.P + (valid ? row_instance(a, r) : 0) * (long long)sd.PW;
      }
      epi_bar_sync();

      for (int oi = 0; oi < a.n_ops; ++oi) {
        const DevOp op = a.ops[oi];
        const float* w = a.blob + op.off;
        if (op.type != DOP_HALF) {
          if (et < kTcRows) {
            float* yr = y_s + et * td.yp;
            if (op.type == DOP_MIX) {
              float o[BCNF_MAX_SIZE];
              for (int j = 0; j < D; ++j) {
                float s = 0.f;
                for (int i = 0; i < D; ++i) s = fmaf(yr[i], __ldg(w + i * sd.DP + j), s);
                o[j] = s;
              }
              for (int j = 0; j < D; ++j) yr[j] = o[j];
            } else {
              for (int j = 0; j < D; ++j) {
                const float s = __ldg(w + j), b = __ldg(w + sd.DP + j);
                yr[j] = op.type == DOP_ACTNORM_FWD ? fmaf(s, yr[j], b) : __fdiv_rn(yr[j] - b, s);
              }
              ld_s[et] += __ldg(w + 2 * sd.DP);
            }
          }
          continue;
        }
        const HalfLayout& hl = sd.half[op.src];
        const TcHalfLayout& tl = td.half[op.src];
        const int in0 = op.src == 0 ? 0 : sd.Da;
        const int out0 = op.src == 0 ? sd.Da : 0;
        // ---- input of the first Linear: own half of y, zero padded to the K=16 steps in use ----
        if (et < kTcRows) {
          const float* yr = y_s + et * td.yp + in0;
          const int kcols = ((hl.din + 15) >> 4) << 4;
          for (int n0 = 0; n0 < kcols; n0 += 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (n0 + i) < hl.din ? yr[n0 + i] : 0.f;
            store_act8<NPASS>(a_hi, a_lo, et, n0, v);
          }
        }
        fence_proxy_async();
        epi_bar_sync();
        if (et == 0) mbar_arrive_remote(a_ready_leader);

        // ---- hidden layers: TMEM -> (+P | +bias) -> GELU -> bf16 hi/lo tiles of the next layer ----
        for (int l = 0; l < tl.L; ++l) {
          const TcLayer& ly = tl.layer[l];
#pragma unroll
          for (int j = 0; j < kTcIssuers; ++j)
            if (j < ly.n_chunks) { mbar_wait_cluster(&acc_full[j], acc_use[j] & 1); ++acc_use[j]; }
          ++acc_cnt;
          tc_fence_after();
          const bool tr = a.trace && blockIdx.x == 0 && iter == 0 && acc_cnt <= 32 && et == 0;
          const long long t_acc = tr ? clock64() : 0;
          const float* add = l == 0 ? prow_s[row] + op.proj_off : w + hl.off_b[l];
          int coff = 0;
          long long t_own = 0, t_bar = 0;
          for (int nc = 0; nc < ly.n_chunks; ++nc) {
            const int cn = ly.chunk_n[nc];
            const uint32_t col = ((reg0 + (uint32_t)nc) % (uint32_t)td.regions) * (uint32_t)td.region_cols;
            const int groups = cn >> 4;                 // 8-column groups in this thread's half chunk
            const int g0 = (groups * part) >> 2, g1 = (groups * (part + 1)) >> 2;
            const int nbase = coff + nhalf * (cn >> 1);
            for (int g = g0; g < g1; g += 2) {
              const bool two = g + 1 < g1;
              const int n0 = nbase + g * 8;
              // bias / projection values first (global, L1/L2), then two TMEM loads under one wait
              float4 b[4];
              b[0] = __ldg(reinterpret_cast<const float4*>(add + n0));
              b[1] = __ldg(reinterpret_cast<const float4*>(add + n0 + 4));
              if (two) {
                b[2] = __ldg(reinterpret_cast<const float4*>(add + n0 + 8));
                b[3] = __ldg(reinterpret_cast<const float4*>(add + n0 + 12));
              }
              uint32_t r0[8], r1[8];
              tmem_ld8_issue(lane_addr + col + g * 8, r0);
              if (two) tmem_ld8_issue(lane_addr + col + g * 8 + 8, r1);
              tmem_ld_wait();
              gelu_store8<NPASS>(a_hi, a_lo, row, n0, r0, b[0], b[1]);
              if (two) gelu_store8<NPASS>(a_hi, a_lo, row, n0 + 8, r1, b[2], b[3]);
            }
            coff += cn;
            // this N chunk of the next layer's input is complete (and its accumulator slot drained): one a_ready phase
            if (tr && nc == ly.n_chunks - 1) t_own = clock64();
            tc_fence_before();
            fence_proxy_async();
            epi_bar_sync();
            if (tr && nc == ly.n_chunks - 1) t_bar = clock64();
            if (et == 0) mbar_arrive_remote(a_chunk_leader + 8u * (uint32_t)nc);
          }
          reg0 = (reg0 + (uint32_t)ly.n_chunks) % (uint32_t)td.regions;
          if (tr) {   // stamps are written after the arrivals so that the trace's global stores do not delay them
            a.trace[(acc_cnt - 1) * 8 + 3] = t_acc;
            a.trace[(acc_cnt - 1) * 8 + 4] = t_own;
            a.trace[(acc_cnt - 1) * 8 + 5] = t_bar;
          }
        }

        // ---- last Linear: (t | s) from TMEM, then the affine update on the row-owner threads ----
        mbar_wait_cluster(&acc_full[0], acc_use[0] & 1);   // the last Linear is a single chunk
        ++acc_use[0];
        ++acc_cnt;
        tc_fence_after();
        {
          const int doh = tl.doh;
          const int groups = doh >> 3;                  // 8-column groups per half (t or s)
          for (int g = part; g < groups; g += 4) {
            float v[8];
            tmem_ld8(lane_addr + (reg0 % (uint32_t)td.regions) * (uint32_t)td.region_cols + g * 8, v);
            const int j0 = g * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int j = j0 + i;
              const float b = j < hl.dout ? __ldg(w + hl.off_bout + nhalf * hl.dop + j) : 0.f;
              ts_s[row * td.tsp + nhalf * doh + j] = v[i] + b;
            }
          }
          reg0 = (reg0 + 1u) % (uint32_t)td.regions;
          tc_fence_before();
          epi_bar_sync();
          if (et < kTcRows) {
            const float* ts = ts_s + et * td.tsp;
            float* yr = y_s + et * td.yp + out0;
            float ls_sum = 0.f;
            for (int j = 0; j < hl.dout; ++j) {
              const float ls = tanhf(ts[doh + j]);                         // cnf.py:107
              ls_sum += ls;
              if (!op.inverse) yr[j] = fmaf(expf(ls), yr[j], ts[j]);       // cnf.py:179
              else             yr[j] = (yr[j] - ts[j]) * expf(-ls);        // cnf.py:204
            }
            ld_s[et] += ls_sum;                                            // cnf.py:190
          }
        }
      }

      if (et < kTcRows) {
        const long long r = row0 + et;
        if (r < a.n_rows) {
          for (int j = 0; j < D; ++j) flow_output(a, r, j, D, y_s[et * td.yp + j]);
          if (a.logdet) a.logdet[r] = ld_s[et];
        }
      }
      epi_bar_sync();
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

// ---- pack: fp32 (out, in) weights -> bf16 hi/lo pre-swizzled tile images -----------------------------
struct TcPackDesc {
  const float* w;       // Linear weight (out, in) row-major
  unsigned char* dst;   // tile base: [cta 0: hi, lo][cta 1: hi, lo]
  int pitch;            // in_features of the Linear
  int col0;             // first input column (skips nothing for the conditioner's own half)
  int k_valid;          // real K
  int n_valid;          // real N (hidden) / dout (last Linear)
  int out_mode;         // 1: last Linear, rows [0,dout) -> n in [0,doh), rows [dout,2dout) -> n in [doh, 2doh)
  int doh;
  int chunk_off, chunk_n, kc;
  int kw;               // K columns per tile: 64 or 32
};

__global__ void tc_pack_kernel(const TcPackDesc* __restrict__ descs) {
  const TcPackDesc d = descs[blockIdx.x];
  const int half_rows = d.chunk_n >> 1;
  const int row_bytes = 2 * d.kw;                 // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
  const int cpr = d.kw >> 3;                      // 16-byte chunks per row
  const uint32_t rows_b = (uint32_t)half_rows * (uint32_t)row_bytes;
  for (int e = threadIdx.x; e < d.chunk_n * cpr; e += blockDim.x) {
    const int nl_all = e / cpr, c16 = e - nl_all * cpr;
    const int r = nl_all / half_rows, nl = nl_all - r * half_rows;
    const int n = d.chunk_off + nl_all;
    int src_row = -1;
    if (!d.out_mode) { if (n < d.n_valid) src_row = n; }
    else if (n < d.doh) { if (n < d.n_valid) src_row = n; }
    else if (n - d.doh < d.n_valid) src_row = d.n_valid + (n - d.doh);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float v[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int k = d.kc * d.kw + c16 * 8 + p * 2 + t;
        v[t] = (src_row >= 0 && k < d.k_valid) ? d.w[(size_t)src_row * d.pitch + d.col0 + k] : 0.f;
      }
      const float h0 = __bfloat162float(__float2bfloat16_rn(v[0])), h1 = __bfloat162float(__float2bfloat16_rn(v[1]));
      hi[p] = pack_bf16x2(h0, h1);
      lo[p] = pack_bf16x2(v[0] - h0, v[1] - h1);
    }
    // hardware swizzle: 16-byte chunk index XOR row-in-atom (128B mode: row & 7; 64B mode: (row >> 1) & 3)
    const int sw = d.kw == 64 ? (nl & 7) : ((nl >> 1) & 3);
    unsigned char* base = d.dst + (size_t)r * 2u * rows_b + (size_t)nl * row_bytes + ((c16 ^ sw) << 4);
    *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + rows_b) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

}  // namespace bcnf
