// Condition projection P = h . W1h^T + b1 on the tensor cores (same machinery as flow_tc.cuh).
//
// For NLL / log-prob evaluation every row has its own conditioning instance, so the hoisted part
// of the first Linear (torch.cat([y, h]) -> nn.Linear, cnf.py:101-104) is 39 % of all MACs of a
// *_large config and an fp32 FMA GEMM would dominate the step.  A CTA pair
// computes, per work item, the projection of 128 instances (64 per CTA) for ONE conditioner
// network (N = H1 padded, <= 4 chunks of <= 256 columns, accumulators in TMEM):
//   * the epilogue/converter warps read h (fp32) from global, split it into bf16 hi/lo and write
//     K-major SWIZZLE_128B tiles into a ring of A stages;
//   * the producer warp streams the pre-swizzled bf16 hi/lo tiles of W1h with TMA bulk copies;
//   * one issuing warp per N chunk runs tcgen05.mma.cta_group::2 over the K chunks (3 passes for
//     bf16x3), then the converter warps turn into the epilogue: TMEM -> + b1 -> P (fp32, global).
// Work items are ordered network-major so that concurrently running pairs share the same weight
// tiles in L2.
#pragma once
#include "flow_tc.cuh"

namespace bcnf {

struct ProjTcDims {
  TcLayer layer[2];        // N chunking / K chunks of the projection for nn_a and nn_b networks
  long long stream_bytes[2];
  int a_stages, b_stages;
  int stage_bytes;         // one B tile stage (hi [+ lo])
  int off_b, off_misc, smem_bytes;
  int C;                   // number of condition features (K)
  int PW;
  int two_way;             // networks alternate nn_a, nn_b (else all nn_a): lets the issuers stay on kernel parameters
};

struct ProjNet {           // one conditioner network of the stack
  long long stream_off;    // byte offset of its W1h tile stream
  int proj_off;            // column offset in P
  int src;                 // 0: nn_a, 1: nn_b
};

template <int NPASS>
__global__ void __launch_bounds__(kTcThreads, 1)
proj_tc_kernel(const float* __restrict__ h, float* __restrict__ P, const float* __restrict__ bproj,
               const unsigned char* __restrict__ blob, const ProjNet* __restrict__ nets, const int n_nets,
               const long long n_inst, const ProjTcDims pd) {
  extern __shared__ __align__(1024) unsigned char smem_pj[];
  constexpr int kAStage = (NPASS == 3 ? 2 : 1) * kTcATile;     // hi [+ lo] of a 64-row x 64-k tile
  unsigned char* a_st = smem_pj;
  unsigned char* b_st = smem_pj + pd.off_b;
  unsigned char* misc = smem_pj + pd.off_misc;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(misc);   // [8]
  uint64_t* w_peer = w_full + 8;                          // [8] leader
  uint64_t* w_empty = w_peer + 8;                         // [8]
  uint64_t* a_full = w_empty + 8;                         // [4] leader, count 2
  uint64_t* a_empty = a_full + 4;                         // [4] count = n_chunks of the item... fixed 4 arrivals (see below)
  uint64_t* acc_full = a_empty + 4;                       // [4]
  uint64_t* tmem_free = acc_full + 4;                     // [1] leader, count 2
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tmem_free + 1);

  // warp index and TMEM base as lane-0 broadcasts: provably warp-uniform, so ptxas keeps the MMA operands in
  // uniform registers instead of wrapping every tcgen05.mma in a uniformisation loop (tools/mma_probe.cu:
  // 139 -> 106 cycles per issue)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int nbs = pd.b_stages, nas = pd.a_stages;
  const long long n_mt = (n_inst + 2 * kTcRows - 1) / (2 * kTcRows);
  const long long n_items = n_mt * n_nets;
  const long long n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const int kc_total = (pd.C + 63) / 64;
  const int last_ksteps = (pd.C - 64 * (kc_total - 1) + 15) / 16;

  if (tid == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_peer[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&a_full[s], 2); mbar_init(&a_empty[s], kTcIssuers); mbar_init(&acc_full[s], 1); }
    mbar_init(tmem_free, 2);
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
    // ===================== W1h tile producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long item = cluster_id; item < n_items; item += n_clusters) {
        const ProjNet net = nets[item / n_mt];
        const TcLayer& ly = pd.layer[net.src];
        const unsigned char* src = blob + net.stream_off;
        for (int kc = 0; kc < kc_total; ++kc)
          for (int nc = 0; nc < ly.n_chunks; ++nc, ++it) {
            const uint32_t rows_b = (uint32_t)(ly.chunk_n[nc] >> 1) * 128u;
            const int s = it % nbs;
            const uint32_t use = it / nbs;
            if (use > 0) mbar_wait_cluster(&w_empty[s], (use - 1) & 1);
            const uint32_t bytes = rows_b * (NPASS == 3 ? 2u : 1u);
            mbar_expect_tx(&w_full[s], bytes);
            tma_bulk_g2s(b_st + (size_t)s * pd.stage_bytes, src + (size_t)cta * 2u * rows_b, bytes, &w_full[s]);
            src += 4u * rows_b;
          }
      }
    }
  } else if (warp < kTcFirstEpiWarp) {
    if (!leader) {
      if (warp == 1 && lane < nbs) {      // relay: one lane per B stage
        long long total = 0;
        for (long long item = cluster_id; item < n_items; item += n_clusters)
          total += (long long)kc_total * pd.layer[nets[item / n_mt].src].n_chunks;
        const uint32_t peer_bar = mapa_u32(smem_u32(&w_peer[lane]), 0);
        uint32_t use = 0;
        for (long long it = lane; it < total; it += nbs, ++use) {
          mbar_wait(&w_full[lane], use & 1);
          mbar_arrive_remote(peer_bar);
        }
      }
    } else if (lane == 0) {
      // ===================== MMA issuers: warp 1+j owns N chunk j =====================
      const int j = warp - 1;
      const uint32_t a_addr = smem_u32(a_st), b_addr = smem_u32(b_st);
      int s = 0; uint32_t par = 0;           // B ring position of the next K step's first tile
      uint32_t a_it = 0, item_cnt = 0;
      for (long long item = cluster_id; item < n_items; item += n_clusters, ++item_cnt) {
        const TcLayer& ly = pd.layer[pd.two_way ? (int)((item / n_mt) & 1) : 0];
        const int nch = ly.n_chunks;
        if (item_cnt > 0) mbar_wait_cluster(tmem_free, (item_cnt - 1) & 1);   // previous item's accumulators drained
        tc_fence_after();
        uint32_t col = 0;
        for (int c = 0; c < j && c < nch; ++c) col += (uint32_t)(ly.chunk_n[c] >> 1);
        const int cn = j < nch ? ly.chunk_n[j] : 16;
        const uint32_t idesc = make_idesc(cn);
        const uint32_t rows_b = (uint32_t)(cn >> 1) * 128u;
        for (int kc = 0; kc < kc_total; ++kc, ++a_it) {
          const int sa = a_it % nas;
          mbar_wait_cluster(&a_full[sa], (a_it / nas) & 1);
          tc_fence_after();
          if (j < nch) {
            int sj = s + j; uint32_t pj = par;
            while (sj >= nbs) { sj -= nbs; pj ^= 1; }
            mbar_wait(&w_full[sj], pj);
            mbar_wait_cluster(&w_peer[sj], pj);
            tc_fence_after();
            const int ksteps = kc == kc_total - 1 ? last_ksteps : 4;
            const uint64_t ah = make_smem_desc(a_addr + sa * kAStage);
            const uint64_t al = make_smem_desc(a_addr + sa * kAStage + kTcATile);
            const uint64_t wh = make_smem_desc(b_addr + sj * pd.stage_bytes);
            const uint64_t wl = make_smem_desc(b_addr + sj * pd.stage_bytes + rows_b);
            for (int k = 0; k < ksteps; ++k) {
              const uint32_t first = (kc | k) == 0 ? 0u : 1u;
              umma_2sm(tmem_base + col, ah + 2 * k, wh + 2 * k, idesc, first);
              if (NPASS == 3) {
                umma_2sm(tmem_base + col, al + 2 * k, wh + 2 * k, idesc, 1u);
                umma_2sm(tmem_base + col, ah + 2 * k, wl + 2 * k, idesc, 1u);
              }
            }
            umma_commit_2sm(&w_empty[sj], 3);
          }
          // every issuer (also the ones without a chunk) reports the A stage: a_empty counts kTcIssuers arrivals
          umma_commit_2sm(&a_empty[sa], 3);
          s += nch;
          while (s >= nbs) { s -= nbs; par ^= 1; }
        }
        if (j < nch) umma_commit_2sm(&acc_full[j], 3);
      }
    }
  } else {
    // ===================== converter + epilogue warps =====================
    const int et = tid - 32 * kTcFirstEpiWarp;
    const int q = warp & 3;
    const int part = (warp - kTcFirstEpiWarp) >> 2;
    const int row = ((q & 1) << 5) + lane;
    const int nhalf = q >> 1;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_full_leader0 = mapa_u32(smem_u32(&a_full[0]), 0);
    const uint32_t tmem_free_leader = mapa_u32(smem_u32(tmem_free), 0);
    uint32_t a_it = 0;
    uint32_t acc_use[kTcIssuers] = {0, 0, 0, 0};
    const int crow = et >> 3, cgrp = et & 7;       // converter mapping: 64 rows x 8 groups of 8 k

    for (long long item = cluster_id; item < n_items; item += n_clusters) {
      const long long net_i = item / n_mt, mt = item - net_i * n_mt;
      const ProjNet net = nets[net_i];
      const TcLayer& ly = pd.layer[net.src];
      const long long inst0 = mt * (2 * kTcRows) + (long long)cta * kTcRows;
      // ---- feed the A ring: h[inst0 .. +64, kc*64 .. +64] as bf16 hi/lo tiles ----
      for (int kc = 0; kc < kc_total; ++kc, ++a_it) {
        const int sa = a_it % nas;
        const uint32_t use = a_it / nas;
        if (use > 0) mbar_wait_cluster(&a_empty[sa], (use - 1) & 1);
        const long long inst = inst0 + crow;
        const int k0 = kc * 64 + cgrp * 8;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        if (inst < n_inst) {
          const float* src = h + inst * (long long)pd.C + k0;
          if (k0 + 8 <= pd.C && (pd.C & 3) == 0) {
            const float4 x0 = __ldg(reinterpret_cast<const float4*>(src));
            const float4 x1 = __ldg(reinterpret_cast<const float4*>(src + 4));
            v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) if (k0 + i < pd.C) v[i] = __ldg(src + i);
          }
        }
        unsigned char* st = a_st + (size_t)sa * kAStage;
        store_act8<NPASS>(st, st + kTcATile, crow, cgrp * 8, v);
        fence_proxy_async();
        epi_bar_sync();
        if (et == 0) mbar_arrive_remote(a_full_leader0 + 8u * (uint32_t)sa);
      }
      // ---- epilogue: TMEM -> + b1 -> P ----
#pragma unroll
      for (int j = 0; j < kTcIssuers; ++j)
        if (j < ly.n_chunks) { mbar_wait_cluster(&acc_full[j], acc_use[j] & 1); ++acc_use[j]; }
      tc_fence_after();
      const long long inst = inst0 + row;
      uint32_t col = 0;
      int coff = 0;
      for (int nc = 0; nc < ly.n_chunks; ++nc) {
        const int cn = ly.chunk_n[nc];
        const int groups = cn >> 4;
        const int g0 = (groups * part) >> 2, g1 = (groups * (part + 1)) >> 2;
        const int nbase = coff + nhalf * (cn >> 1);
        for (int g = g0; g < g1; ++g) {
          const int n0 = net.proj_off + nbase + g * 8;
          float v[8];
          tmem_ld8(lane_addr + col + g * 8, v);
          if (inst < n_inst) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bproj + n0));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bproj + n0 + 4));
            float* dst = P + inst * (long long)pd.PW + n0;
            *reinterpret_cast<float4*>(dst) = make_float4(v[0] + b0.x, v[1] + b0.y, v[2] + b0.z, v[3] + b0.w);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4] + b1.x, v[5] + b1.y, v[6] + b1.z, v[7] + b1.w);
          }
        }
        col += (uint32_t)(cn >> 1);
        coff += cn;
      }
      tc_fence_before();
      epi_bar_sync();
      if (et == 0) mbar_arrive_remote(tmem_free_leader);
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
