// Training GEMMs of the conditioner MLP on the tensor cores (tcgen05 + TMEM), fp32 in / fp32 out.
//
// Same contract as train_gemm_kernel (train_ops.cuh): C = epi(A.B) (+ C) with arbitrary element strides for A
// and B, so that the forward (x W^T), the data gradient (d W) and the weight gradient (d^T x) of a Linear
// (reference: nn.Linear inside ConditionalNestedNeuralNetwork, cnf.py:78-83, back-propagated by autograd in
// Trainer._train_batch, trainer.py:268) are one kernel reading the activations and the parameters where they
// lie, in the reference's own (out, in) layout.  Arithmetic is the 3-pass bf16 split of the inference kernel
// (flow_tc.cuh): a = a_hi + a_lo, b = b_hi + b_lo in bf16, a_hi.b_hi + a_lo.b_hi + a_hi.b_lo accumulated in
// fp32 in TMEM -- fp32-class accuracy (~2^-17 per product) at a third of the bf16 tensor rate.
//
// One CTA computes a 128 x BN tile of C (tcgen05.mma.cta_group::1, M = 128, N = BN, K = 16):
//   * 16 loader warps in two groups of 8 alternate over the 64-wide K chunks of the CTA's K range: each thread
//     issues all of its global loads for the chunk (coalesced along whichever index is contiguous in memory),
//     splits the values to bf16 hi / lo and writes K-major SWIZZLE_128B operand tiles into a ring of stages --
//     the transposition of an "MN-major" operand (d^T, W read along its input index) happens in this store, so
//     the MMA only ever sees K-major tiles;
//   * lane 0 of warp 0 issues the MMAs and commits each stage back to the loaders;
//   * the loader warps then turn into the epilogue: tcgen05.ld (lane = row of the tile), + bias / GELU /
//     dropout / gelu' * mask, store.  With split-K (gridDim.z > 1) the raw accumulators are reduced into an fp32
//     workspace with red.global.add and the last CTA of a tile (per-tile arrival counter) applies the
//     epilogue; workspace and counters are left zeroed for the next launch.
//   * optionally the column sums of the A operand's rows... (see colsum below) are folded in: when g.colsum is
//     set, the epilogue also accumulates sum_i C(i, j) of the FINAL values into colsum[j] (bias gradients of
//     the layer below come from the data-gradient GEMM that produces d).
#pragma once
#include "flow_tc.cuh"
#include "train_ops.cuh"

namespace bcnf {

constexpr int kTgGroupWarps = 8;
constexpr int kTgGroupThreads = 32 * kTgGroupWarps;
constexpr int kTgGroups = 2;
constexpr int kTgThreads = 32 + kTgGroups * kTgGroupThreads;   // warp 0: TMEM allocation + MMA issue
constexpr int kTgBM = 128;
constexpr int kTgATile = kTgBM * 128;                           // bytes: 128 rows x 64 k, bf16

template <int BN>
struct TgCfg {
  static constexpr int b_tile = BN * 128;
  static constexpr int stage = 2 * kTgATile + 2 * b_tile;       // A hi, A lo, B hi, B lo
  static constexpr int stages = BN <= 64 ? 4 : 3;
  static constexpr int bar_off = stages * stage;
  static constexpr int smem = bar_off + 128;
  static constexpr int tmem_cols = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ void umma_1sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_1sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tg_group_sync(int group) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(kTgGroupThreads) : "memory");
}

// Global -> registers: the (ROWS x 64) operand tile X(row, k) = src[row*s_row + k*s_k], zero outside
// row < rows_valid, k < k_valid.  Unit = 8 consecutive k of one row; UPT units per thread.
//   k contiguous (s_k == 1): 8 lanes cover the 64 k of a row with vector loads (width by alignment);
//   otherwise: lanes run along the rows, each of the 8 loads of a unit is a coalesced warp access.
template <int ROWS>
__device__ __forceinline__ void tg_load_regs(float (&v)[ROWS / 32][8], const float* __restrict__ src, long long s_row,
                                             long long s_k, int rows_valid, int k_valid, int t, int vec) {
  constexpr int UPT = ROWS / 32;
#pragma unroll
  for (int p = 0; p < UPT; ++p) {
    const int u = t + p * kTgGroupThreads;
    int row, grp;
    if (s_k == 1) { grp = u & 7; row = u >> 3; } else { row = u % ROWS; grp = u / ROWS; }
    const int k0 = grp * 8;
    const float* q = src + row * s_row + k0 * s_k;
    const bool row_ok = row < rows_valid;
    if (s_k == 1 && row_ok && k0 + 8 <= k_valid && vec > 1) {
      if (vec == 4) {
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(q));
        const float4 x1 = __ldg(reinterpret_cast<const float4*>(q + 4));
        v[p][0] = x0.x; v[p][1] = x0.y; v[p][2] = x0.z; v[p][3] = x0.w;
        v[p][4] = x1.x; v[p][5] = x1.y; v[p][6] = x1.z; v[p][7] = x1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x = __ldg(reinterpret_cast<const float2*>(q + 2 * i));
          v[p][2 * i] = x.x; v[p][2 * i + 1] = x.y;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[p][i] = (row_ok && k0 + i < k_valid) ? __ldg(q + i * s_k) : 0.f;
    }
  }
}

template <int ROWS>
__device__ __forceinline__ void tg_store_regs(const float (&v)[ROWS / 32][8], unsigned char* hi, unsigned char* lo,
                                              bool k_fast, int t) {
  constexpr int UPT = ROWS / 32;
#pragma unroll
  for (int p = 0; p < UPT; ++p) {
    const int u = t + p * kTgGroupThreads;
    int row, grp;
    if (k_fast) { grp = u & 7; row = u >> 3; } else { row = u % ROWS; grp = u / ROWS; }
    store_act8<3>(hi, lo, row, grp * 8, v[p]);     // tile 0: byte offset row*128 + ((grp ^ (row & 7)) << 4)
  }
}

// widest vector load usable for rows of `base` with pitch `s_row` floats (k contiguous)
static inline __host__ __device__ int tg_vec_width(const float* base, long long s_row) {
  const unsigned long long a = (unsigned long long)base;
  if ((a & 15) == 0 && (s_row & 3) == 0) return 4;
  if ((a & 7) == 0 && (s_row & 1) == 0) return 2;
  return 1;
}

struct TgExtra {
  float* ws;               // split-K: fp32 workspace, >= tiles_m*128 x tiles_n*BN floats per launch, zero on entry and on exit
  unsigned int* counters;  // split-K: one per C tile, zero on entry and on exit
  float* colsum;           // optional: colsum[j] += sum_i C(i, j) of the final values (atomic; caller zeroes)
  int chunks_per_split;    // K chunks (64 wide) per blockIdx.z
  long long* trace;        // debug: clock64 stamps of CTA (0,0,0) (null in normal runs)
};

template <int BN>
__global__ void __launch_bounds__(kTgThreads, 1)
train_tc_gemm_kernel(const GemmArgs g, const TgExtra x) {
  using Cfg = TgCfg<BN>;
  extern __shared__ __align__(1024) unsigned char smem_tg[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_tg + Cfg::bar_off);   // [stages] one arrival per fill (group leader)
  uint64_t* empty = full + 4;                                             // [stages] tcgen05.commit
  uint64_t* acc_full = empty + 4;                                         // [1]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(acc_full + 1);
  uint32_t* flag_s = tmem_ptr_s + 1;

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int i0 = blockIdx.y * kTgBM, j0 = blockIdx.x * BN;
  const int n_kc = (g.K + 63) >> 6;
  const int kc0 = blockIdx.z * x.chunks_per_split;
  const int kc1 = min(n_kc, kc0 + x.chunks_per_split);
  const int n_it = kc1 - kc0;      // >= 1 (host guarantees)
  const bool tr = x.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (tr && tid == 0) x.trace[0] = clock64();

  if (tid == 0) {
    for (int s = 0; s < Cfg::stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "n"(Cfg::tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);
  if (tr && tid == 0) x.trace[1] = clock64();

  if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BN);
      const uint32_t st_addr = smem_u32(smem_tg);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % Cfg::stages;
        mbar_wait(&full[s], (uint32_t)(it / Cfg::stages) & 1u);
        tc_fence_after();
        if (tr && it < 8) x.trace[8 + it] = clock64();
        const int krem = g.K - (kc0 + it) * 64;
        const int ksteps = krem >= 64 ? 4 : (krem + 15) >> 4;
        const uint32_t base = st_addr + (uint32_t)s * Cfg::stage;
        const uint64_t ah = make_smem_desc(base), al = make_smem_desc(base + kTgATile);
        const uint64_t bh = make_smem_desc(base + 2 * kTgATile), bl = make_smem_desc(base + 2 * kTgATile + Cfg::b_tile);
        for (int k = 0; k < ksteps; ++k) {
          umma_1sm(tmem_base, ah + 2 * k, bh + 2 * k, idesc, (it | k) == 0 ? 0u : 1u);
          umma_1sm(tmem_base, al + 2 * k, bh + 2 * k, idesc, 1u);
          umma_1sm(tmem_base, ah + 2 * k, bl + 2 * k, idesc, 1u);
        }
        umma_commit_1sm(&empty[s]);
      }
      umma_commit_1sm(acc_full);
    }
  } else {
    // ===================== loaders (two groups alternate over the K chunks), then epilogue =====================
    const int group = (warp - 1) / kTgGroupWarps;
    const int t = tid - 32 - group * kTgGroupThreads;
    const bool a_k_fast = g.as1 == 1, b_k_fast = g.bs0 == 1;
    // A tile: rows = i (stride as0), k = r (stride as1).  B tile: rows = j (stride bs1), k = r (stride bs0).
    const float* a_base = g.A + (long long)i0 * g.as0;
    const float* b_base = g.B + (long long)j0 * g.bs1;
    const int a_vec = a_k_fast ? tg_vec_width(a_base, g.as0) : 1;
    const int b_vec = b_k_fast ? tg_vec_width(b_base, g.bs1) : 1;
    const int a_rows = g.M - i0, b_rows = g.N - j0;
    for (int it = group; it < n_it; it += kTgGroups) {
      const int kc = kc0 + it;
      const int s = it % Cfg::stages;
      const int use = it / Cfg::stages;
      float va[kTgBM / 32][8], vb[BN / 32][8];
      tg_load_regs<kTgBM>(va, a_base + (long long)kc * 64 * g.as1, g.as0, g.as1, a_rows, g.K - kc * 64, t, a_vec);
      tg_load_regs<BN>(vb, b_base + (long long)kc * 64 * g.bs0, g.bs1, g.bs0, b_rows, g.K - kc * 64, t, b_vec);
      if (tr && t == 0 && it < 8) x.trace[16 + it] = clock64();
      if (use > 0) mbar_wait(&empty[s], (uint32_t)(use - 1) & 1u);
      unsigned char* st = smem_tg + (size_t)s * Cfg::stage;
      tg_store_regs<kTgBM>(va, st, st + kTgATile, a_k_fast, t);
      tg_store_regs<BN>(vb, st + 2 * kTgATile, st + 2 * kTgATile + Cfg::b_tile, b_k_fast, t);
      if (tr && t == 0 && it < 8) x.trace[24 + it] = clock64();
      fence_proxy_async();
      tg_group_sync(group);
      if (t == 0) mbar_arrive_local(&full[s]);
      if (tr && t == 0 && it < 8) x.trace[32 + it] = clock64();
    }

    // ---- epilogue: lane quarter q of TMEM = rows q*32..q*32+31 of the tile; 4 warps share a quarter ----
    const int q = warp & 3;
    const int part = (warp - 1) >> 2;                 // 0..3: which quarter of the BN columns
    constexpr int CPP = BN / 4;                       // columns per part (8, 16 or 32)
    const int i = i0 + q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * CPP);
    const unsigned long long seed = g.seed_ptr ? (g.seed ^ *g.seed_ptr) : g.seed;
    const float keep_scale = g.p_drop > 0.f ? 1.0f / (1.0f - g.p_drop) : 1.0f;
    if (tr && tid == 32) x.trace[2] = clock64();
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (tr && tid == 32) x.trace[3] = clock64();

    bool finish = true;
    const bool split = gridDim.z > 1;
    float* wsp = nullptr;
    if (split) {
      // reduce the raw partial sums into the workspace tile; the last CTA to arrive finishes the tile
      const long long tile_id = (long long)blockIdx.y * gridDim.x + blockIdx.x;
      wsp = x.ws + tile_id * (long long)(kTgBM * BN) + (long long)(q * 32 + lane) * BN + part * CPP;
#pragma unroll
      for (int c = 0; c < CPP; c += 8) {
        float v[8];
        tmem_ld8(taddr + c, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(wsp + c + e, v[e]);
      }
      __threadfence();
      // all loader/epilogue threads of this CTA have published their adds before the counter moves
      asm volatile("bar.sync 3, %0;" ::"n"(kTgGroups * kTgGroupThreads) : "memory");
      if (tid == 32) {
        const unsigned int prev = atomicAdd(x.counters + tile_id, 1u);
        const bool last = prev == gridDim.z - 1;
        if (last) x.counters[tile_id] = 0u;
        *flag_s = last ? 1u : 0u;
      }
      asm volatile("bar.sync 3, %0;" ::"n"(kTgGroups * kTgGroupThreads) : "memory");
      finish = *flag_s != 0u;
      if (finish) __threadfence();
    }

    if (finish) {
#pragma unroll
      for (int c = 0; c < CPP; c += 8) {
        float v[8];
        if (split) {
#pragma unroll
          for (int e = 0; e < 8; ++e) { v[e] = __ldcg(wsp + c + e); __stcg(wsp + c + e, 0.f); }
        } else {
          tmem_ld8(taddr + c, v);
        }
        float cs[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = j0 + part * CPP + c + e;
          float val = v[e];
          const bool ok = i < g.M && j < g.N;
          if (ok) {
            const long long idx = (long long)i * g.cs0 + j;
            if (g.epi == TEPI_BIAS) {
              val += __ldg(g.bias + j);
            } else if (g.epi == TEPI_BIAS_GELU_DROP) {
              val += __ldg(g.bias + j);
              g.save[idx] = val;
              val = gelu_erf(val);
              if (g.p_drop > 0.f)
                val = dropout_uniform(seed, g.layer_uid, (unsigned long long)i * (unsigned)g.N + (unsigned)j) >= g.p_drop
                          ? val * keep_scale : 0.f;
            } else if (g.epi == TEPI_DGELU_DROP) {
              val *= dgelu_erf(__ldg(g.saved + idx));
              if (g.p_drop > 0.f)
                val = dropout_uniform(seed, g.layer_uid, (unsigned long long)i * (unsigned)g.N + (unsigned)j) >= g.p_drop
                          ? val * keep_scale : 0.f;
            }
            if (g.beta != 0.f) val += g.beta * g.C[idx];
            g.C[idx] = val;
          } else {
            val = 0.f;
          }
          cs[e] = val;
        }
        if (x.colsum) {
          // column sums over the 32 rows of this warp, then one atomic per column and warp
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float sum = cs[e];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            cs[e] = sum;
          }
          if (lane < 8) {
            const int j = j0 + part * CPP + c + lane;
            float mine = cs[0];
#pragma unroll
            for (int e = 1; e < 8; ++e) mine = lane == e ? cs[e] : mine;
            if (j < g.N) atomicAdd(x.colsum + j, mine);
          }
        }
      }
    }
  }

  if (tr && tid == 32) x.trace[4] = clock64();
  tc_fence_before();
  __syncthreads();
  if (tr && tid == 0) x.trace[5] = clock64();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::tmem_cols));
  }
}


// =====================================================================================================
// Second kernel: operands as pre-split bf16 "images", fed by TMA.
//
// The trace of the kernel above (tools/tc_gemm_trace.py) shows where a batch-256 GEMM spends its time: issuing the
// register-staged loads (1500 cycles per pair of K chunks), waiting for them and converting (2700), and one
// thread issuing 12 MMAs per chunk at >100 cycles each.  An IMAGE of a matrix X (R rows x K columns) is X split
// into bf16 hi and lo planes, each stored as [K/64 chunks][R_pad rows][128 bytes] with the 16-byte units of a row
// XOR-swizzled by (row & 7): exactly the K-major SWIZZLE_128B operand tile of tcgen05.mma, so a (rows x 64)
// operand tile is ONE contiguous block per plane and the producer warp moves it with a bulk copy -- no registers,
// no conversion, four stages deep.  Whoever produces a matrix writes its image: the epilogue of this kernel
// (activations, gradients), img_pack_kernel (parameters after the optimizer step, network inputs).
// Three warps issue the MMAs, one per pass of the 3-term split, each into its own accumulator; the epilogue adds them.
// =====================================================================================================


struct ImgArgs {
  const unsigned char* a_img; long long a_plane; int a_rpad;   // A(i, r): rows i (M), chunks over r (K)
  const unsigned char* b_img; long long b_plane; int b_rpad;   // B(r, j): rows j (N), chunks over r (K)
  unsigned char* c_img; long long c_plane; int c_rpad;         // optional: image of the values written to C
  float* colsum;
};

constexpr int kT2EpiWarps = 16;
constexpr int kT2Threads = 32 * (4 + kT2EpiWarps);   // warp 0 producer, warps 1-3 issuers (one per pass), 16 epilogue warps

template <int BN>
struct T2Cfg {
  static constexpr int a_tile = kTgBM * 128, b_tile = BN * 128;
  static constexpr int stage = 2 * a_tile + 2 * b_tile;
  static constexpr int stages = BN <= 64 ? 4 : 3;
  static constexpr int bar_off = stages * stage;
  static constexpr int smem = bar_off + 128;
  static constexpr int tmem_cols = 3 * BN <= 128 ? 128 : (3 * BN <= 256 ? 256 : 512);
};

// (img_store8 / img_store1, the writers of the image format: img_store.cuh)

template <int BN>
__global__ void __launch_bounds__(kT2Threads, 1)
train_tc2_gemm_kernel(const GemmArgs g, const ImgArgs im) {
  using Cfg = T2Cfg<BN>;
  extern __shared__ __align__(1024) unsigned char smem_t2[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_t2 + Cfg::bar_off);   // [stages] producer: expect_tx + 4 bulk copies
  uint64_t* empty = full + 4;                                             // [stages] 3 commits (one per issuer)
  uint64_t* acc_full = empty + 4;                                         // [1] 3 commits
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int i0 = blockIdx.y * kTgBM, j0 = blockIdx.x * BN;
  const int n_it = (g.K + 63) >> 6;

  if (tid == 0) {
    for (int s = 0; s < Cfg::stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 3); }
    mbar_init(acc_full, 3);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "n"(Cfg::tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);

  if (warp == 0) {
    // ===================== producer: 4 bulk copies per K chunk =====================
    if (lane == 0) {
      const unsigned char* a_src = im.a_img + (long long)i0 * 128;
      const unsigned char* b_src = im.b_img + (long long)j0 * 128;
      for (int it = 0; it < n_it; ++it) {
        const int s = it % Cfg::stages, use = it / Cfg::stages;
        if (use > 0) mbar_wait(&empty[s], (uint32_t)(use - 1) & 1u);
        unsigned char* st = smem_t2 + (size_t)s * Cfg::stage;
        mbar_expect_tx(&full[s], (uint32_t)Cfg::stage);
        const long long ao = (long long)it * im.a_rpad * 128, bo = (long long)it * im.b_rpad * 128;
        tma_bulk_g2s(st, a_src + ao, Cfg::a_tile, &full[s]);
        tma_bulk_g2s(st + Cfg::a_tile, a_src + im.a_plane + ao, Cfg::a_tile, &full[s]);
        tma_bulk_g2s(st + 2 * Cfg::a_tile, b_src + bo, Cfg::b_tile, &full[s]);
        tma_bulk_g2s(st + 2 * Cfg::a_tile + Cfg::b_tile, b_src + im.b_plane + bo, Cfg::b_tile, &full[s]);
      }
    }
  } else if (warp < 4) {
    // ===================== MMA issuers: warp 1+p issues pass p into accumulator p =====================
    if (lane == 0) {
      const int p = warp - 1;                       // 0: a_hi.b_hi   1: a_lo.b_hi   2: a_hi.b_lo
      const uint32_t idesc = make_idesc(BN);
      const uint32_t st_addr = smem_u32(smem_t2);
      const uint32_t a_off = p == 1 ? Cfg::a_tile : 0, b_off = 2 * Cfg::a_tile + (p == 2 ? Cfg::b_tile : 0);
      const uint32_t acc = tmem_base + (uint32_t)(p * BN);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % Cfg::stages;
        mbar_wait(&full[s], (uint32_t)(it / Cfg::stages) & 1u);
        tc_fence_after();
        const int krem = g.K - it * 64;
        const int ksteps = krem >= 64 ? 4 : (krem + 15) >> 4;
        const uint32_t base = st_addr + (uint32_t)s * Cfg::stage;
        const uint64_t ad = make_smem_desc(base + a_off), bd = make_smem_desc(base + b_off);
        for (int k = 0; k < ksteps; ++k) umma_1sm(acc, ad + 2 * k, bd + 2 * k, idesc, (it | k) == 0 ? 0u : 1u);
        umma_commit_1sm(&empty[s]);
      }
      umma_commit_1sm(acc_full);
    }
  } else {
    // ===================== epilogue: lane quarter q = rows q*32.., four warps per quarter split the columns ==========
    const int q = warp & 3, part = (warp - 4) >> 2;
    constexpr int CPP = BN / 4;
    const int row = q * 32 + lane, i = i0 + row;
    const int jbase = j0 + part * CPP;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * CPP);
    const unsigned long long seed = g.seed_ptr ? (g.seed ^ *g.seed_ptr) : g.seed;
    const float keep_scale = g.p_drop > 0.f ? 1.0f / (1.0f - g.p_drop) : 1.0f;
    // operands of the epilogue are fetched while the main loop runs: bias / saved pre-activations / old C
    float pf[CPP];
#pragma unroll
    for (int c = 0; c < CPP; ++c) {
      const int j = jbase + c;
      const bool ok = i < g.M && j < g.N;
      pf[c] = 0.f;
      if (g.epi == TEPI_BIAS || g.epi == TEPI_BIAS_GELU_DROP) { if (j < g.N) pf[c] = __ldg(g.bias + j); }
      else if (g.epi == TEPI_DGELU_DROP) { if (ok) pf[c] = dgelu_erf_fast(__ldg(g.saved + (long long)i * g.cs0 + j)); }
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < CPP; c += 8) {
      uint32_t r0[8], r1[8], r2[8];
      tmem_ld8_issue(taddr + c, r0);
      tmem_ld8_issue(taddr + BN + c, r1);
      tmem_ld8_issue(taddr + 2 * BN + c, r2);
      tmem_ld_wait();
      float out[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int j = jbase + c + e;
        float val = (__uint_as_float(r0[e]) + __uint_as_float(r1[e])) + __uint_as_float(r2[e]);
        if (i < g.M && j < g.N) {
          const long long idx = (long long)i * g.cs0 + j;
          if (g.epi == TEPI_BIAS) {
            val += pf[c + e];
          } else if (g.epi == TEPI_BIAS_GELU_DROP) {
            val += pf[c + e];
            g.save[idx] = val;
            val = gelu_erf_fast(val);
            if (g.p_drop > 0.f)
              val = dropout_uniform(seed, g.layer_uid, (unsigned long long)i * (unsigned)g.N + (unsigned)j) >= g.p_drop
                        ? val * keep_scale : 0.f;
          } else if (g.epi == TEPI_DGELU_DROP) {
            val *= pf[c + e];
            if (g.p_drop > 0.f)
              val = dropout_uniform(seed, g.layer_uid, (unsigned long long)i * (unsigned)g.N + (unsigned)j) >= g.p_drop
                        ? val * keep_scale : 0.f;
          }
          if (g.beta != 0.f) val += g.beta * g.C[idx];
          g.C[idx] = val;
        } else {
          val = 0.f;
        }
        out[e] = val;
      }
      const int n0 = jbase + c;
      if (im.c_img && i < im.c_rpad && n0 < (im.c_plane / ((long long)im.c_rpad * 128)) * 64)
        img_store8(im.c_img, im.c_plane, im.c_rpad, i, n0, out);
      if (im.colsum) {
#pragma unroll
        for (int e = 0; e < 8; ++e) out[e] = warp_sum_tc(out[e]);
        if (lane < 8) {
          float mine = out[0];
#pragma unroll
          for (int e = 1; e < 8; ++e) mine = lane == e ? out[e] : mine;
          if (n0 + lane < g.N) atomicAdd(im.colsum + n0 + lane, mine);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::tmem_cols));
  }
}

// =====================================================================================================
// Third kernel: weight gradients dW = d^T . x from the SAME images, read as MN-major operands.
//
// dW(i, j) = sum_r d(r, i) x(r, j): the contraction runs over the batch rows r.  An image chunk [rows r][64 columns] is
// exactly the canonical MN-major SWIZZLE_128B operand tile of tcgen05.mma (cute::UMMA make_umma_desc<Major::MN>:
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -- 64 consecutive M (or N) elements per 128-byte row, 8-row swizzle
// atoms along K): no second copy of the activations, no transposition.  A tile = 128 columns of d (two image chunks,
// LBO = one chunk apart) x 64 rows; B tile = BN columns of x (BN/64 chunks) x 64 rows; one K = 16 step advances the
// descriptors by 16 rows (2 KB).  Otherwise the structure of the kernel above: producer warp with bulk copies, three
// issuing warps (one per pass, own accumulator), epilogue warps add the accumulators and store fp32 (+ beta C).
// =====================================================================================================
template <int BN>
struct T3Cfg {
  static constexpr int a_tile = 2 * 64 * 128, b_tile = (BN / 64) * 64 * 128;   // per plane: chunks x 64 rows x 128 B
  static constexpr int stage = 2 * a_tile + 2 * b_tile;
  static constexpr int stages = BN <= 64 ? 4 : 3;
  static constexpr int bar_off = stages * stage;
  static constexpr int smem = bar_off + 128;
  static constexpr int tmem_cols = 3 * BN <= 256 ? 256 : 512;
};

__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  // MN-major, SWIZZLE_128B: 64-element (128-byte) MN blocks lbo_bytes apart, 8-row K groups 1024 bytes apart
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_mn(int n) {
  // as make_idesc, with both operands MN-major (bits 15, 16)
  return make_idesc(n) | (1u << 15) | (1u << 16);
}

template <int BN>
__global__ void __launch_bounds__(kT2Threads, 1)
train_tc3_dw_kernel(const GemmArgs g, const ImgArgs im) {
  using Cfg = T3Cfg<BN>;
  extern __shared__ __align__(1024) unsigned char smem_t3[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_t3 + Cfg::bar_off);
  uint64_t* empty = full + 4;
  uint64_t* acc_full = empty + 4;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int i0 = blockIdx.y * kTgBM, j0 = blockIdx.x * BN;       // i: columns of d (image a), j: columns of x (image b)
  const int n_it = (g.K + 63) >> 6;                              // K = batch rows
  const int a_chunks_img = (int)(im.a_plane / ((long long)im.a_rpad * 128));
  const int b_chunks_img = (int)(im.b_plane / ((long long)im.b_rpad * 128));
  // chunks of this tile that exist in the images (columns past the image only feed discarded rows / columns of dW)
  const int a_n = min(2, a_chunks_img - (i0 >> 6)), b_n = min(BN / 64, b_chunks_img - (j0 >> 6));

  if (tid == 0) {
    for (int s = 0; s < Cfg::stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 3); }
    mbar_init(acc_full, 3);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)), "n"(Cfg::tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_it; ++it) {
        const int s = it % Cfg::stages, use = it / Cfg::stages;
        if (use > 0) mbar_wait(&empty[s], (uint32_t)(use - 1) & 1u);
        unsigned char* st = smem_t3 + (size_t)s * Cfg::stage;
        mbar_expect_tx(&full[s], (uint32_t)(2 * (a_n + b_n) * 64 * 128));
        const long long r_off = (long long)it * 64 * 128;        // 64 batch rows further down every chunk
        for (int c = 0; c < a_n; ++c) {
          const unsigned char* src = im.a_img + ((long long)((i0 >> 6) + c) * im.a_rpad) * 128 + r_off;
          tma_bulk_g2s(st + c * 8192, src, 8192, &full[s]);
          tma_bulk_g2s(st + Cfg::a_tile + c * 8192, src + im.a_plane, 8192, &full[s]);
        }
        for (int c = 0; c < b_n; ++c) {
          const unsigned char* src = im.b_img + ((long long)((j0 >> 6) + c) * im.b_rpad) * 128 + r_off;
          tma_bulk_g2s(st + 2 * Cfg::a_tile + c * 8192, src, 8192, &full[s]);
          tma_bulk_g2s(st + 2 * Cfg::a_tile + Cfg::b_tile + c * 8192, src + im.b_plane, 8192, &full[s]);
        }
      }
    }
  } else if (warp < 4) {
    if (lane == 0) {
      const int p = warp - 1;                       // 0: a_hi.b_hi   1: a_lo.b_hi   2: a_hi.b_lo
      const uint32_t idesc = make_idesc_mn(BN);
      const uint32_t st_addr = smem_u32(smem_t3);
      const uint32_t a_off = p == 1 ? Cfg::a_tile : 0, b_off = 2 * Cfg::a_tile + (p == 2 ? Cfg::b_tile : 0);
      const uint32_t acc = tmem_base + (uint32_t)(p * BN);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % Cfg::stages;
        mbar_wait(&full[s], (uint32_t)(it / Cfg::stages) & 1u);
        tc_fence_after();
        const int krem = g.K - it * 64;
        const int ksteps = krem >= 64 ? 4 : (krem + 15) >> 4;
        const uint32_t base = st_addr + (uint32_t)s * Cfg::stage;
        const uint64_t ad = make_smem_desc_mn(base + a_off, 8192), bd = make_smem_desc_mn(base + b_off, 8192);
        for (int k = 0; k < ksteps; ++k) umma_1sm(acc, ad + 128 * k, bd + 128 * k, idesc, (it | k) == 0 ? 0u : 1u);   // +2048 B per step
        umma_commit_1sm(&empty[s]);
      }
      umma_commit_1sm(acc_full);
    }
  } else {
    const int q = warp & 3, part = (warp - 4) >> 2;
    constexpr int CPP = BN / 4;
    const int i = i0 + q * 32 + lane;
    const int jbase = j0 + part * CPP;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * CPP);
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < CPP; c += 8) {
      uint32_t r0[8], r1[8], r2[8];
      tmem_ld8_issue(taddr + c, r0);
      tmem_ld8_issue(taddr + BN + c, r1);
      tmem_ld8_issue(taddr + 2 * BN + c, r2);
      tmem_ld_wait();
      if (i < g.M) {
        float* crow = g.C + (long long)i * g.cs0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = jbase + c + e;
          if (j < g.N) {
            float val = (__uint_as_float(r0[e]) + __uint_as_float(r1[e])) + __uint_as_float(r2[e]);
            if (g.beta != 0.f) val += g.beta * crow[j];
            crow[j] = val;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::tmem_cols));
  }
}

// ---- fp32 matrix -> image (parameters in either orientation, network inputs) -----------------------------------
struct ImgPackDesc {
  const float* src;        // X(row, k) = src[row*s_row + k*s_k]
  long long s_row, s_k;
  int rows, k;             // valid extent; the image is zero outside
  unsigned char* dst;
  long long plane;         // bytes between the hi and the lo plane ( = chunks * rpad * 128 )
  int rpad;                // rows of the image (multiple of 32)
  int chunks;              // 64-wide K chunks of the image
};
constexpr int kImgPackMax = 64;
struct ImgPackBatch { ImgPackDesc d[kImgPackMax]; };

__global__ void __launch_bounds__(kTgGroupThreads) img_pack_kernel(const ImgPackBatch batch) {
  const ImgPackDesc& d = batch.d[blockIdx.y];
  const int r0 = blockIdx.x * 32;
  if (r0 >= d.rpad) return;
  const int t = threadIdx.x;
  const bool k_fast = d.s_k == 1;
  const float* base = d.src + (long long)r0 * d.s_row;
  const int vec = k_fast ? tg_vec_width(base, d.s_row) : 1;
  int row, grp;
  if (k_fast) { grp = t & 7; row = t >> 3; } else { row = t & 31; grp = t >> 5; }
  // four chunks of loads in flight per thread before the first store
  for (int kc0 = 0; kc0 < d.chunks; kc0 += 4) {
    float v[4][1][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (kc0 + u < d.chunks)
        tg_load_regs<32>(v[u], base + (long long)(kc0 + u) * 64 * d.s_k, d.s_row, d.s_k, d.rows - r0, d.k - (kc0 + u) * 64, t, vec);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (kc0 + u < d.chunks) img_store8(d.dst, d.plane, d.rpad, r0 + row, (kc0 + u) * 64 + grp * 8, v[u][0]);
  }
}

}  // namespace bcnf
