// Fused coupling stack, second generation: 128 rows per CTA, activations streamed through L2.
//
// Same scope and arithmetic as flow_tc.cuh (reference cnf.py:479-488, :500-506 and callees: conditioner MLP
// cnf.py:98-107, affine coupling cnf.py:165-213, ActNorm cnf.py:348-354, orthonormal mixing cnf.py:333-339;
// 3-pass bf16 split with fp32 accumulation, fp32 everywhere outside the GEMM operands).  What changes is where the
// activations live.  flow_tc.cuh keeps a layer's activations in shared memory, which caps a CTA at 64 rows (TMEM
// cannot hold 128 rows x 528 fp32 columns next to them) and therefore caps the tensor pipe at 58 % : an SS-mode
// tcgen05.mma re-reads its A rows for every N chunk, and with 64 rows per CTA that read costs more shared-memory
// cycles than the MMA lasts.  Here a CTA PAIR owns a 256-row tile (128 rows per CTA, the
// M = 256 cta_group::2 MMA of gemm_img2.cuh: operand reads fit the 64 B/clk port for N >= 192) and a layer's
// activations are an OPERAND IMAGE (bf16 hi / lo planes, [K/64][128 rows][128 B], SWIZZLE_128B tile layout) in a
// per-CTA scratch that never leaves L2:
//
//   * producer warp: per N chunk (<= 256 columns) and 64-wide K stage, bulk copies of the CTA's 128 rows of the
//     activation image and of its half of the weight image tile into a 3-stage ring (64 KB per stage);
//   * one issuing thread (leader CTA): tcgen05.mma.cta_group::2 M = 256, N = chunk, K = 16 into one of two
//     256-column TMEM accumulators;
//   * sixteen epilogue warps per CTA drain the other accumulator: + bias (or + the hoisted condition projection P
//     for the first Linear) -> exact-erf GELU -> bf16 hi / lo split -> one 64-column image chunk (16 KB per plane)
//     staged in shared memory in image layout;
//   * a store thread writes each staged chunk to the scratch with a bulk store and publishes its completion; the
//     producer starts the next layer's K stages as soon as the chunks they read are complete, so the MMAs of layer
//     l+1 overlap the epilogue of the last N chunk of layer l;
//   * the last Linear (N = 2 x dout padded to 16) leaves t and s in TMEM; the epilogue warps apply tanh / exp, the
//     affine update, the log-det row sum, ActNorm and the orthonormal mixing in fp32, keep y in the output buffer
//     between networks, and stage the next network's own-half input as image chunk 0.
//
// A 528-wide layer is three N chunks (192, 192, 144): 322 shared-memory cycles per K = 16 step against 264 tensor
// cycles, an 82 % ceiling instead of 58 %.  Per CTA and hidden layer the L2 traffic is 3 x 270 KB of activations +
// 557 KB of weights in, 270 KB out; the live scratch of all 148 CTAs is ~ 60 MB of the 126 MB L2.
#pragma once
#include "flow_tc.cuh"
#include "gemm_img2.cuh"

namespace bcnf {

constexpr int kS2Stages = 3;
constexpr int kS2EpiWarps = 16;
constexpr int kS2EpiThreads = 32 * kS2EpiWarps;
constexpr int kS2FirstEpi = 4;                  // warp 0 producer, 1 issuer (leader) / relay (peer), 2 store, 3 idle
constexpr int kS2Threads = 32 * kS2FirstEpi + kS2EpiThreads;
constexpr int kS2Rows = 128;                    // rows per CTA (256 per pair)
constexpr int kS2Tile = kS2Rows * 128;          // bytes of one (128 rows x 64 k) bf16 tile / image chunk
constexpr int kS2SlotCols = 256;                // TMEM: two accumulator slots
constexpr int kS2MaxChunks = 4;
constexpr int kS2YPitch = 29;                   // floats per row of the glue scratch (D <= 28; odd: no bank conflicts)
constexpr int kS2MiscBytes = 2048;

struct S2Layer {
  int n_chunks;                 // N chunks (<= 256 columns each, multiples of 16; all but the last multiples of 64)
  int chunk_n[kS2MaxChunks];
  int n_kst;                    // 64-wide K stages
  int last_ksteps;              // K = 16 steps that carry data in the last stage
  int n_img;                    // 64-column image chunks of this layer's OUTPUT (hidden layers; 0 for the last Linear)
  int w_rpad;                   // rows of the weight image
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
// streaming read of the projection slice: every P value is read once per (row, network); keep it out of the way of the
// activation scratch in L2
__device__ __forceinline__ void ldg_stream8(const float* p, float4& lo, float4& hi) {
  uint32_t r[8];
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
  lo = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
  hi = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
}

// (acc + add) -> GELU -> bf16 hi / lo for 8 columns: one 16-byte unit per plane
template <int NPASS>
__device__ __forceinline__ void s2_gelu_pack8(const uint32_t* r, const float4& b0, const float4& b1, uint4& hi4, uint4& lo4) {
  f32x2 v[4];
  v[0] = gelu_erf_fast2(add2(pack2(__uint_as_float(r[0]), __uint_as_float(r[1])), pack2(b0.x, b0.y)));
  v[1] = gelu_erf_fast2(add2(pack2(__uint_as_float(r[2]), __uint_as_float(r[3])), pack2(b0.z, b0.w)));
  v[2] = gelu_erf_fast2(add2(pack2(__uint_as_float(r[4]), __uint_as_float(r[5])), pack2(b1.x, b1.y)));
  v[3] = gelu_erf_fast2(add2(pack2(__uint_as_float(r[6]), __uint_as_float(r[7])), pack2(b1.z, b1.w)));
  uint32_t hi[4], lo[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float x0, x1;
    unpack2(v[i], x0, x1);
    hi[i] = pack_bf16x2(x0, x1);
    if (NPASS == 3) {
      const f32x2 h = pack2(__uint_as_float(hi[i] << 16), __uint_as_float(hi[i] & 0xffff0000u));
      float l0, l1;
      unpack2(add2(v[i], h ^ 0x8000000080000000ull), l0, l1);
      lo[i] = pack_bf16x2(l0, l1);
    }
  }
  hi4 = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  lo4 = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

template <int NPASS>
__global__ void __launch_bounds__(kS2Threads, 1)
flow_tc2_kernel(const FlowArgs a, const StackDims sd, const S2Dims d2, const unsigned char* __restrict__ w_img,
                const long long* __restrict__ w_off, unsigned char* __restrict__ act, unsigned int* __restrict__ dbg) {
  constexpr int PL = NPASS == 3 ? 2 : 1;
  extern __shared__ __align__(1024) unsigned char smem_s2[];
  unsigned char* stg = smem_s2 + d2.stg_off;                     // 2 x 16 KB: staged image chunk (hi, lo) / glue scratch
  unsigned char* misc = smem_s2 + d2.misc_off;
  uint64_t* full = reinterpret_cast<uint64_t*>(misc);            // [3] own TMA
  uint64_t* peer_full = full + 4;                                // [3] leader: the peer's stage has landed
  uint64_t* empty = peer_full + 4;                               // [3] multicast commit
  uint64_t* acc_full = empty + 4;                                // [2] multicast commit
  uint64_t* tmem_empty = acc_full + 2;                           // [2] leader: both epilogues drained the slot
  uint64_t* stg_full = tmem_empty + 2;                           // [1] epilogue -> store thread
  uint64_t* stg_empty = stg_full + 1;                            // [1] store thread -> epilogue
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(misc + 192);
  volatile uint32_t* done_cnt = reinterpret_cast<volatile uint32_t*>(misc + 196);   // image chunks whose stores are complete
  unsigned long long* stg_dst = reinterpret_cast<unsigned long long*>(misc + 208);  // destination of the staged chunk (hi plane)
  const float** prow_s = reinterpret_cast<const float**>(misc + 256);               // [128] projection row of each tile row

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
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&tmem_empty[b], 2); }
    mbar_init(stg_full, 1);
    mbar_init(stg_empty, 1);
    *done_cnt = 0u;
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
      uint32_t it = 0, job = 0, grp_base = 0;
      for (long long ti = 0; ti < my_tiles; ++ti) {
        for (int oi = 0; oi < a.n_ops; ++oi) {
          const DevOp op = a.ops[oi];
          if (op.type != DOP_HALF) continue;
          const S2Half& hl = d2.half[op.src];
          const unsigned char* wnet = w_img + w_off[oi];
          for (int l = 0; l <= hl.L; ++l, ++job) {
            const S2Layer& ly = hl.layer[l];
            const unsigned char* abuf = act_cta + (long long)(job & 1u) * buf_bytes;
            const uint32_t in_img = l == 0 ? 1u : (uint32_t)hl.layer[l - 1].n_img;
            int n0 = 0;
            for (int c = 0; c < ly.n_chunks; ++c) {
              const int cn = ly.chunk_n[c];
              const uint32_t b_bytes = (uint32_t)(cn >> 1) * 128u;
              const unsigned char* wsrc = wnet + ly.w_off + (long long)(n0 + (int)cta * (cn >> 1)) * 128;
              for (int kc = 0; kc < ly.n_kst; ++kc, ++it) {
                if (c == 0) {
                  // the image chunk this stage reads must have been stored (by this CTA's own epilogue)
                  const uint32_t need = grp_base + (uint32_t)kc + 1u;
                  if (*done_cnt < need) {
                    const unsigned long long t0 = g2_now();
                    uint32_t n = 0;
                    while (*done_cnt < need)
                      if (((++n) & 1023u) == 0 && g2_now() - t0 > 4000000000ull) s2_timeout(dbg, 0x100u, need);
                  }
                  asm volatile("fence.proxy.async;" ::: "memory");
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
            grp_base += in_img;
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
              if (nchunk >= 2) s2_wait<true>(&tmem_empty[slot], ((nchunk >> 1) - 1u) & 1u, dbg, 0x300u);
              tc_fence_after();
              const uint32_t acc = tmem_base + slot * (uint32_t)kS2SlotCols;
              const uint32_t idesc = make_idesc_m256(ly.chunk_n[c]);
              for (int kc = 0; kc < ly.n_kst; ++kc, ++it) {
                const int s = (int)(it % kS2Stages);
                const uint32_t par = (it / kS2Stages) & 1u;
                s2_wait<true>(&full[s], par, dbg, 0x301u);
                s2_wait<true>(&peer_full[s], par, dbg, 0x302u);
                tc_fence_after();
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
              }
            }
          }
        }
    }
  } else if (warp == 2) {
    // ===================== store thread: staged image chunk -> scratch, completion published to the producer ========
    if (lane == 0) {
      uint32_t per_tile = 0;
      for (int hi = 0; hi < d2.n_halfops; ++hi) {
        const S2Half& hl = d2.half[d2.two_way ? (hi & 1) : 0];
        per_tile += 1u;
        for (int l = 0; l < hl.L; ++l) per_tile += (uint32_t)hl.layer[l].n_img;
      }
      const uint32_t total = per_tile * (uint32_t)my_tiles;
      uint32_t issued = 0, published = 0;
      while (issued < total) {
        const uint32_t par = issued & 1u;
        if (!mbar_test(stg_full, par)) {
          if (published < issued) {      // idle: finish what is in flight so that the producer can go on
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence_block();
            *done_cnt = issued;
            published = issued;
          }
          s2_wait<true>(stg_full, par, dbg, 0x400u);
        }
        unsigned char* dst = reinterpret_cast<unsigned char*>(stg_dst[0]);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(stg)), "n"(kS2Tile) : "memory");
        if (NPASS == 3)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + d2.a_plane), "r"(smem_u32(stg + kS2Tile)), "n"(kS2Tile) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        ++issued;
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the staging buffer may be refilled
        mbar_arrive_local(stg_empty);
        asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");          // everything but the newest store is complete
        if (published < issued - 1u) {
          asm volatile("fence.proxy.async;" ::: "memory");
          __threadfence_block();
          *done_cnt = issued - 1u;
          published = issued - 1u;
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      *done_cnt = issued;
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
    float* ts_s = reinterpret_cast<float*>(stg);                    // [128][32]  (glue scratch, hi half of the staging)
    float* y_s = reinterpret_cast<float*>(stg + kS2Tile);           // [128][29]  (lo half)
    uint32_t nchunk = 0, stg_use = 0, job = 0;
    float ld_acc = 0.f;                            // log-det of this row (row owners: part == 0)

    auto staging_free = [&]() { if (stg_use > 0) s2_wait<false>(stg_empty, (stg_use - 1u) & 1u, dbg, 0x500u); };
    auto hand_over = [&](unsigned char* dst, bool release_slot, uint32_t slot) {
      fence_proxy_async();
      if (release_slot) tc_fence_before();
      s2_epi_sync();
      if (et == 0) {
        stg_dst[0] = reinterpret_cast<unsigned long long>(dst);
        mbar_arrive_local(stg_full);
        if (release_slot) mbar_arrive_remote(tmem_empty_leader + 8u * slot);
      }
      ++stg_use;
    };

    long long ti = 0;
    long long row_g = 0;                           // global row of this thread in the current tile
    bool valid = false;
    int oi = 0;

    // load a fresh tile: y from the input, log-det 0, projection row pointers
    auto fresh_tile = [&]() {
      const long long tile = pair_id + ti * n_pairs;
      row_g = tile * (2 * kS2Rows) + (long long)cta * kS2Rows + row;
      valid = row_g < a.n_rows;
      staging_free();
      if (part == 0) {
        for (int j = 0; j < D; ++j) y_s[row * kS2YPitch + j] = valid ? __ldg(a.in + row_g * D + j) : 0.f;
        ld_acc = 0.f;
        prow_s[row] = a.P + (valid ? row_instance(a, row_g) : 0) * (long long)sd.PW;
      }
      s2_epi_sync();
      oi = 0;
    };
    // ActNorm / mixing layers up to the next conditioner network (cnf.py:333-354); y in y_s, all 512 threads
    auto glue_ops = [&]() {
      while (oi < a.n_ops) {
        const DevOp op = a.ops[oi];
        if (op.type == DOP_HALF) break;
        const float* w = a.blob + op.off;
        if (op.type == DOP_MIX) {
          float o[7];
#pragma unroll
          for (int u = 0; u < 7; ++u) {
            const int j = part + 4 * u;
            float s = 0.f;
            if (j < D)
              for (int i = 0; i < D; ++i) s = fmaf(y_s[row * kS2YPitch + i], __ldg(w + i * sd.DP + j), s);   // y @ M
            o[u] = s;
          }
          s2_epi_sync();
#pragma unroll
          for (int u = 0; u < 7; ++u) { const int j = part + 4 * u; if (j < D) y_s[row * kS2YPitch + j] = o[u]; }
        } else {
#pragma unroll
          for (int u = 0; u < 7; ++u) {
            const int j = part + 4 * u;
            if (j < D) {
              const float s = __ldg(w + j), b = __ldg(w + sd.DP + j), y = y_s[row * kS2YPitch + j];
              y_s[row * kS2YPitch + j] = op.type == DOP_ACTNORM_FWD ? fmaf(s, y, b) : __fdiv_rn(y - b, s);
            }
          }
          if (part == 0) ld_acc += __ldg(w + 2 * sd.DP);
        }
        s2_epi_sync();
        ++oi;
      }
    };
    // y_s -> output buffer (state between networks, and the final result)
    auto store_y = [&](bool final_) {
      if (part == 0 && valid) {
        for (int j = 0; j < D; ++j) a.out[row_g * D + j] = y_s[row * kS2YPitch + j];
        if (final_ && a.logdet) a.logdet[row_g] = ld_acc;
      }
    };
    // own-half input of the conditioner network ops[oi] as image chunk 0 of the buffer its first Linear reads
    auto stage_x_in = [&]() {
      const DevOp op = a.ops[oi];
      const HalfLayout& hl = sd.half[op.src];
      const int in0 = op.src == 0 ? 0 : sd.Da;
      float v[16];
      if (part == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = i < hl.din ? y_s[row * kS2YPitch + in0 + i] : 0.f;
      }
      s2_epi_sync();                                // y_s / ts_s are dead from here: the staging is rewritten as an image
      if (part == 0) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          hi[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          const float h0 = __uint_as_float(hi[i] << 16), h1 = __uint_as_float(hi[i] & 0xffff0000u);
          lo[i] = pack_bf16x2(v[2 * i] - h0, v[2 * i + 1] - h1);
        }
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int off = row * 128 + ((u ^ (row & 7)) << 4);
          const uint4 h4 = u == 0 ? make_uint4(hi[0], hi[1], hi[2], hi[3]) : (u == 1 ? make_uint4(hi[4], hi[5], hi[6], hi[7]) : z4);
          const uint4 l4 = u == 0 ? make_uint4(lo[0], lo[1], lo[2], lo[3]) : (u == 1 ? make_uint4(lo[4], lo[5], lo[6], lo[7]) : z4);
          *reinterpret_cast<uint4*>(stg + off) = h4;
          if (NPASS == 3) *reinterpret_cast<uint4*>(stg + kS2Tile + off) = l4;
        }
      }
      hand_over(act_cta + (long long)(job & 1u) * buf_bytes, false, 0u);     // read by job `job`
    };

    if (my_tiles > 0) {
      fresh_tile();
      glue_ops();
      for (;;) {
        // ---- here ops[oi] is a conditioner network; y_s holds the current state ----
        store_y(false);
        stage_x_in();
        const DevOp op = a.ops[oi];
        const float* w = a.blob + op.off;
        const HalfLayout& hl = sd.half[op.src];
        const S2Half& tl = d2.half[op.src];
        const int out0 = op.src == 0 ? sd.Da : 0;

        // ---- hidden layers: TMEM -> (+P | +bias) -> GELU -> bf16 hi/lo image chunks of the next layer's input ----
        for (int l = 0; l < L; ++l, ++job) {
          const S2Layer& ly = tl.layer[l];
          unsigned char* obuf = act_cta + (long long)((job + 1u) & 1u) * buf_bytes;
          const float* add = l == 0 ? prow_s[row] + op.proj_off : w + hl.off_b[l];
          int n0 = 0;
          for (int c = 0; c < ly.n_chunks; ++c, ++nchunk) {
            const int cn = ly.chunk_n[c];
            const uint32_t slot = nchunk & 1u;
            s2_wait<false>(&acc_full[slot], (nchunk >> 1) & 1u, dbg, 0x501u);
            tc_fence_after();
            const int n_ic = (cn + 63) >> 6;
            for (int ic = 0; ic < n_ic; ++ic) {
              const int colc = ic * 64 + part * 16;
              uint4 h4[2], l4[2];
              if (colc < cn) {
                const int n = n0 + colc;
                float4 b[4];
                if (l == 0) {
                  ldg_stream8(add + n, b[0], b[1]);
                  ldg_stream8(add + n + 8, b[2], b[3]);
                } else {
#pragma unroll
                  for (int u = 0; u < 4; ++u) b[u] = __ldg(reinterpret_cast<const float4*>(add + n + 4 * u));
                }
                uint32_t r[16];
                tmem_ld16_issue(lane_addr + slot * (uint32_t)kS2SlotCols + (uint32_t)colc, r);
                tmem_ld_wait();
                s2_gelu_pack8<NPASS>(r, b[0], b[1], h4[0], l4[0]);
                s2_gelu_pack8<NPASS>(r + 8, b[2], b[3], h4[1], l4[1]);
              } else {
                h4[0] = h4[1] = l4[0] = l4[1] = make_uint4(0u, 0u, 0u, 0u);
              }
              staging_free();
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int off = row * 128 + (((part * 2 + u) ^ (row & 7)) << 4);
                *reinterpret_cast<uint4*>(stg + off) = h4[u];
                if (NPASS == 3) *reinterpret_cast<uint4*>(stg + kS2Tile + off) = l4[u];
              }
              hand_over(obuf + (long long)((n0 >> 6) + ic) * kS2Tile, ic == n_ic - 1, slot);
            }
            n0 += cn;
          }
        }

        // ---- last Linear: (t | s) from TMEM, affine update, log-det, glue, next network's input ----
        {
          const uint32_t slot = nchunk & 1u;
          s2_wait<false>(&acc_full[slot], (nchunk >> 1) & 1u, dbg, 0x502u);
          tc_fence_after();
          staging_free();
          if (part == 0) {
            uint32_t r0[16], r1[16];
            tmem_ld16_issue(lane_addr + slot * (uint32_t)kS2SlotCols, r0);
            tmem_ld16_issue(lane_addr + slot * (uint32_t)kS2SlotCols + 16u, r1);
            tmem_ld_wait();
            const float* bo = w + hl.off_bout;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              ts_s[row * 32 + c] = __uint_as_float(r0[c]) + (c < 2 * hl.dop ? __ldg(bo + c) : 0.f);
              ts_s[row * 32 + 16 + c] = __uint_as_float(r1[c]) + (16 + c < 2 * hl.dop ? __ldg(bo + 16 + c) : 0.f);
            }
            for (int j = 0; j < D; ++j) y_s[row * kS2YPitch + j] = valid ? a.out[row_g * D + j] : 0.f;
          }
          tc_fence_before();
          s2_epi_sync();
          if (et == 0) mbar_arrive_remote(tmem_empty_leader + 8u * slot);
          ++nchunk; ++job;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int m = part + 4 * u;
            if (m < hl.dout) {
              const float t = ts_s[row * 32 + m];
              const float ls = tanhf(ts_s[row * 32 + hl.dop + m]);                                  // cnf.py:107
              const float yd = y_s[row * kS2YPitch + out0 + m];
              y_s[row * kS2YPitch + out0 + m] = op.inverse ? (yd - t) * expf(-ls) : fmaf(expf(ls), yd, t);   // cnf.py:204 / :179
              ts_s[row * 32 + hl.dop + m] = ls;
            }
          }
          s2_epi_sync();
          if (part == 0) {
            float s = 0.f;
            for (int m = 0; m < hl.dout; ++m) s += ts_s[row * 32 + hl.dop + m];                     // cnf.py:190, fixed order
            ld_acc += s;
          }
        }
        ++oi;
        glue_ops();
        if (oi == a.n_ops) {
          store_y(true);
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
