// C = A . B^T (+ bias) on operand images, CTA pairs: the throughput GEMM of the package.
//
// Used for the condition projection P = h . W1h^T + b1 of every conditioner network at once
// (torch.cat([y, h]) -> nn.Linear hoisted out of the stack, cnf.py:101-104; 39 % of all MACs of a *_large
// log-prob evaluation): M = instances, N = sum of padded first-layer widths, K = C.
//
// Operands are images (train_tc.cuh: bf16 hi / lo planes, [K/64][rows][128 B], SWIZZLE_128B tile layout), so a
// producer warp feeds the MMAs with plain bulk copies.  A CTA pair (cluster of 2, tcgen05.mma.cta_group::2)
// computes 256 x 256 tiles of C: each CTA stages its own 128 rows of A and its half (128 rows) of the B tile
// (64 KB per stage for the 3-pass split, 3 stages), the leader's issuer thread runs M = 256, N = 256, K = 16 MMAs
// -- 128 tensor-pipe cycles each, longer than one thread's issue interval, and 64 B/clk of shared-memory operand
// reads per CTA, the ratio of the large CUTLASS tiles -- into one of two 256-column TMEM accumulators, and eight
// epilogue warps per CTA drain the other one (lane = row, + bias, fp32 stores) while the next tile's MMAs run.
// Persistent: pairs walk the tile list round-robin, n fastest, so concurrent pairs share the same rows of A in L2.
#pragma once
#include "flow_tc.cuh"
#include "train_tc.cuh"

namespace bcnf {

constexpr int kG2Stages = 3;
constexpr int kG2EpiWarps = 8;
constexpr int kG2Threads = 32 * (4 + kG2EpiWarps);   // warp 0 producer, warp 1 issuer (leader) / relay (peer), 2-3 idle
constexpr int kG2Tile = 128 * 128;                   // bytes: 128 rows x 64 k, bf16

template <int NPASS>
struct G2Cfg {
  static constexpr int planes = NPASS == 3 ? 2 : 1;
  static constexpr int stage = 2 * planes * kG2Tile;  // A (own 128 rows) + B (own half) per plane
  static constexpr int stg_off = kG2Stages * stage;   // epilogue staging: one 128-row x 64-column chunk, hi and lo plane
  static constexpr int bar_off = stg_off + 2 * kG2Tile;
  static constexpr int smem = bar_off + 256;
};

struct G2Args {
  const unsigned char* a_img; long long a_plane; int a_rpad;   // rows = M index, chunks over K
  const unsigned char* b_img; long long b_plane; int b_rpad;   // rows = N index, chunks over K
  float* C; long long ldc;
  const float* bias;        // [N] or null
  int M, N, K;
  // epilogue mode 1 (C == null): out = gelu(acc + bias) written as an operand image (rows = M index, k = N index):
  // the A operand of the next layer's GEMM; columns >= N inside the image are written as zeros
  unsigned char* c_img; long long c_plane; int c_rpad;
  // A operand gathered chunk by chunk (a_tab_n > 0): K chunk kc is the [a_rpad rows][128 B] block a_tab_hi[kc] / a_tab_lo[kc]
  // -- lets one GEMM read [x_t | h_(t-1)] from the buffers where earlier launches left them
  const unsigned char* a_tab_hi[16]; const unsigned char* a_tab_lo[16]; int a_tab_n;
  // epilogue mode 2, LSTM cell (nn.LSTM gate order i, f, g, o; columns interleaved n = 4*unit + gate, so an 8-column group
  // holds two whole units): c = sigmoid(f) c + sigmoid(i) tanh(g); h = sigmoid(o) tanh(c).  cell / hsum are fp32
  // [N/8 groups][state_rows][2] (a lane = a row: coalesced); a 256-column tile = 64 units = ONE 64-column chunk of the h
  // image, staged in shared memory and written with a bulk store to h_hi[tile] / h_lo[tile].
  int lstm;
  float* cell; float* hsum; long long state_rows;
  unsigned char* h_hi[4]; unsigned char* h_lo[4];
  int debug;                // bit 0: skip the stores of the epilogue (timing experiments)
  unsigned long long* trace;   // debug: globaltimer stamps [pair][tile (<16)][4] (null in normal runs)
};

__device__ __forceinline__ unsigned long long g2_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// logistic and tanh through ex2 / rcp (MUFU): absolute error ~1e-7, a fifth of the instructions of tanhf / full division
__device__ __forceinline__ float g2_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float g2_tanh(float x) { return fmaf(2.0f, g2_sigmoid(2.0f * x), -1.0f); }
__device__ __forceinline__ void g2_epi_sync() { asm volatile("bar.sync 2, %0;" ::"n"(32 * kG2EpiWarps) : "memory"); }
__device__ __forceinline__ uint32_t make_idesc_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// Tile walk of a CTA pair.  Wide N (the projection): round-robin over the tile list, n fastest, so the pairs running
// at the same time share rows of A in L2.  Narrow N (a hidden layer: 3 column tiles): a pair owns whole ROW tiles and
// walks their column tiles back to back, so its 256 rows of A are fetched from L2 by one SM pair three times in a row
// instead of by three pairs at different times (the activation image of a batch is as large as a third of L2).
struct G2Walk {
  long long n_tiles, n_pairs, pair_id;
  int tiles_m, tiles_n, by_rows;
  __device__ __forceinline__ long long count() const {
    if (!by_rows) return pair_id < n_tiles ? (n_tiles - pair_id + n_pairs - 1) / n_pairs : 0;
    const long long rows = pair_id < tiles_m ? (tiles_m - pair_id + n_pairs - 1) / n_pairs : 0;
    return rows * tiles_n;
  }
  __device__ __forceinline__ void at(long long i, int& m, int& n) const {
    if (!by_rows) { const long long t = pair_id + i * n_pairs; m = (int)(t / tiles_n); n = (int)(t % tiles_n); }
    else { m = (int)(pair_id + (i / tiles_n) * n_pairs); n = (int)(i % tiles_n); }
  }
};

template <int NPASS>
__global__ void __launch_bounds__(kG2Threads, 1)
gemm_img2_kernel(const G2Args g) {
  using Cfg = G2Cfg<NPASS>;
  extern __shared__ __align__(1024) unsigned char smem_g2[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_g2 + Cfg::bar_off);   // [S] own TMA
  uint64_t* peer_full = full + 4;                                         // [S] leader: the peer's stage has landed
  uint64_t* empty = peer_full + 4;                                        // [S] multicast commit
  uint64_t* acc_full = empty + 4;                                         // [2] multicast commit
  uint64_t* tmem_empty = acc_full + 2;                                    // [2] leader: both epilogues drained the buffer
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int n_kc = (g.K + 63) >> 6;
  const int tiles_n = (g.N + 255) >> 8, tiles_m = (g.M + 255) >> 8;
  const long long n_tiles = (long long)tiles_m * tiles_n;
  const long long n_pairs = gridDim.x >> 1, pair_id = blockIdx.x >> 1;
  G2Walk walk;
  walk.n_tiles = n_tiles; walk.n_pairs = n_pairs; walk.pair_id = pair_id; walk.tiles_m = tiles_m; walk.tiles_n = tiles_n;
  walk.by_rows = tiles_n <= 4 && tiles_m >= n_pairs / 2;
  const long long my_tiles = walk.count();

  if (tid == 0) {
    for (int s = 0; s < kG2Stages; ++s) { mbar_init(&full[s], 1); mbar_init(&peer_full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&tmem_empty[b], 2); }
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
    // ===================== producer: own 128 rows of A, own half of the B tile =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long ti = 0; ti < my_tiles; ++ti) {
        int tm, tn;
        walk.at(ti, tm, tn);
        const int m0 = tm * 256 + (int)cta * 128, n0 = tn * 256 + (int)cta * 128;
        const unsigned char* a_src = g.a_img + (long long)m0 * 128;
        const unsigned char* b_src = g.b_img + (long long)n0 * 128;
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int s = it % kG2Stages;
          const uint32_t use = it / kG2Stages;
          if (use > 0) mbar_wait_cluster(&empty[s], (use - 1) & 1);
          unsigned char* st = smem_g2 + (size_t)s * Cfg::stage;
          mbar_expect_tx(&full[s], (uint32_t)Cfg::stage);
          const long long ao = (long long)kc * g.a_rpad * 128, bo = (long long)kc * g.b_rpad * 128;
          const unsigned char* ahi = g.a_tab_n > 0 ? g.a_tab_hi[kc] + (long long)m0 * 128 : a_src + ao;
          tma_bulk_g2s(st, ahi, kG2Tile, &full[s]);
          tma_bulk_g2s(st + Cfg::planes * kG2Tile, b_src + bo, kG2Tile, &full[s]);
          if (NPASS == 3) {
            const unsigned char* alo = g.a_tab_n > 0 ? g.a_tab_lo[kc] + (long long)m0 * 128 : a_src + g.a_plane + ao;
            tma_bulk_g2s(st + kG2Tile, alo, kG2Tile, &full[s]);
            tma_bulk_g2s(st + 3 * kG2Tile, b_src + g.b_plane + bo, kG2Tile, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (!leader) {
      // ===================== relay: tell the leader when this CTA's stage has landed =====================
      // One thread, stages in the order they fill, spinning on test_wait: divergent lanes blocked in try_wait
      // (a suspending wait) stall each other for the length of the hardware time-out (measured: 10-40x slower runs).
      if (lane == 0) {
        const long long total = my_tiles * n_kc;
        const uint32_t peer_bar0 = mapa_u32(smem_u32(&peer_full[0]), 0);
        for (long long it = 0; it < total; ++it) {
          const int s = (int)(it % kG2Stages);
          mbar_spin(&full[s], (uint32_t)(it / kG2Stages) & 1u);
          mbar_arrive_remote(peer_bar0 + 8u * (uint32_t)s);
        }
      }
    } else if (lane == 0) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = make_idesc_m256(256);
      const uint32_t st_addr = smem_u32(smem_g2);
      uint32_t it = 0, tile_it = 0;
      for (long long ti = 0; ti < my_tiles; ++ti, ++tile_it) {
        const uint32_t buf = tile_it & 1u;
        if (tile_it >= 2) mbar_wait_cluster(&tmem_empty[buf], ((tile_it >> 1) - 1) & 1);
        tc_fence_after();
        if (g.trace && tile_it < 16) g.trace[(pair_id * 16 + tile_it) * 4 + 0] = g2_now();
        const uint32_t acc = tmem_base + buf * 256u;
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int s = it % kG2Stages;
          const uint32_t par = (it / kG2Stages) & 1u;
          mbar_wait(&full[s], par);
          mbar_wait_cluster(&peer_full[s], par);
          tc_fence_after();
          const int krem = g.K - kc * 64;
          const int ksteps = krem >= 64 ? 4 : (krem + 15) >> 4;
          const uint32_t base = st_addr + (uint32_t)s * Cfg::stage;
          const uint64_t ah = make_smem_desc(base), bh = make_smem_desc(base + Cfg::planes * kG2Tile);
          const uint64_t al = make_smem_desc(base + kG2Tile), bl = make_smem_desc(base + 3 * kG2Tile);
          for (int k = 0; k < ksteps; ++k) {
            umma_2sm(acc, ah + 2 * k, bh + 2 * k, idesc, (kc | k) == 0 ? 0u : 1u);
            if (NPASS == 3) {
              umma_2sm(acc, al + 2 * k, bh + 2 * k, idesc, 1u);
              umma_2sm(acc, ah + 2 * k, bl + 2 * k, idesc, 1u);
            }
          }
          umma_commit_2sm(&empty[s], 3);
        }
        umma_commit_2sm(&acc_full[buf], 3);
        if (g.trace && tile_it < 16) g.trace[(pair_id * 16 + tile_it) * 4 + 1] = g2_now();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: lane = row of this CTA's 128, two warps per lane quarter x 128 columns ==========
    const int q = warp & 3, part = (warp - 4) >> 2;
    const int et = tid - 128;
    const uint32_t tmem_empty_leader = mapa_u32(smem_u32(tmem_empty), 0);
    uint32_t tile_it = 0;
    for (long long ti = 0; ti < my_tiles; ++ti, ++tile_it) {
      const uint32_t buf = tile_it & 1u;
      int tm, tn;
      walk.at(ti, tm, tn);
      const long long i = (long long)tm * 256 + (long long)cta * 128 + q * 32 + lane;
      const int n_base = tn * 256 + part * 128;
      mbar_wait_cluster(&acc_full[buf], (tile_it >> 1) & 1);
      tc_fence_after();
      if (g.trace && leader && et == 0 && tile_it < 16) g.trace[(pair_id * 16 + tile_it) * 4 + 2] = g2_now();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u + (uint32_t)(part * 128);
      if (g.lstm) {
        // ---- LSTM cell: gates -> (c, h); h as the next step's operand image chunk ----
        unsigned char* stg = smem_g2 + Cfg::stg_off;
        const int row = q * 32 + lane;
        const long long row0 = (long long)tm * 256 + (long long)cta * 128;
        const long long grow = row0 + row;
        const bool row_ok = grow < g.M;
        if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous tile's store has read the staging
        g2_epi_sync();
#pragma unroll 1
        for (int c = 0; c < 128; c += 16) {
          uint32_t r0[8], r1[8];
          tmem_ld8_issue(taddr + c, r0);
          tmem_ld8_issue(taddr + c + 8, r1);
          tmem_ld_wait();
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            const uint32_t* r = h8 == 0 ? r0 : r1;
            const int gcol = tn * 256 + part * 128 + c + h8 * 8;         // first gate column of this pair of units
            float hv[2] = {0.f, 0.f};
            if (row_ok && gcol < g.N) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + gcol));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(g.bias + gcol + 4));
              const long long so = ((long long)(gcol >> 3) * g.state_rows + grow) * 2;
              float2 cp = *reinterpret_cast<const float2*>(g.cell + so);
              const float pre[8] = {__uint_as_float(r[0]) + b0.x, __uint_as_float(r[1]) + b0.y, __uint_as_float(r[2]) + b0.z,
                                    __uint_as_float(r[3]) + b0.w, __uint_as_float(r[4]) + b1.x, __uint_as_float(r[5]) + b1.y,
                                    __uint_as_float(r[6]) + b1.z, __uint_as_float(r[7]) + b1.w};
              float cn[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const float ig = g2_sigmoid(pre[4 * u]), fg = g2_sigmoid(pre[4 * u + 1]);
                const float gg = g2_tanh(pre[4 * u + 2]), og = g2_sigmoid(pre[4 * u + 3]);
                cn[u] = fmaf(fg, u == 0 ? cp.x : cp.y, ig * gg);
                hv[u] = og * g2_tanh(cn[u]);
              }
              *reinterpret_cast<float2*>(g.cell + so) = make_float2(cn[0], cn[1]);
              if (g.hsum) {
                float2 hs = *reinterpret_cast<const float2*>(g.hsum + so);
                hs.x += hv[0]; hs.y += hv[1];
                *reinterpret_cast<float2*>(g.hsum + so) = hs;
              }
            }
            const uint32_t hi = pack_bf16x2(hv[0], hv[1]);
            const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
            const int ucol = part * 32 + ((c + h8 * 8) >> 2);              // unit column inside the 64-unit chunk (even)
            const int off = row * 128 + (((ucol >> 3) ^ (row & 7)) << 4) + ((ucol & 7) << 1);
            *reinterpret_cast<uint32_t*>(stg + off) = hi;
            if (NPASS == 3) *reinterpret_cast<uint32_t*>(stg + kG2Tile + off) = pack_bf16x2(hv[0] - h0, hv[1] - h1);
          }
        }
        fence_proxy_async();
        g2_epi_sync();
        if (et == 0) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g.h_hi[tn] + row0 * 128), "r"(smem_u32(stg)), "n"(kG2Tile) : "memory");
          if (NPASS == 3)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g.h_lo[tn] + row0 * 128), "r"(smem_u32(stg + kG2Tile)), "n"(kG2Tile) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        (void)i;
      } else if (g.c_img) {
        // ---- gelu(acc + bias) -> bf16 hi/lo image of the next layer's input ----
        // A lane owns a row, so direct stores would touch 32 different lines per warp instruction (measured: 2 cycles
        // per 16-byte store, 11-14 us per tile).  The CTA's 128 rows x one 64-column chunk are ONE contiguous 16 KB
        // block of the image (per plane): stage it in shared memory in image layout and hand it to a bulk store.
        unsigned char* stg = smem_g2 + Cfg::stg_off;
        const int c_cols = (int)(g.c_plane / ((long long)g.c_rpad * 128)) * 64;
        const int row = q * 32 + lane;
        const long long row0 = (long long)tm * 256 + (long long)cta * 128;
        const uint32_t tchunk = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u + (uint32_t)(part * 32);
        const int n_tile0 = tn * 256;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int nc0 = n_tile0 + cc * 64;                 // first column of this chunk (multiple of 64)
          if (nc0 >= c_cols) break;
          uint32_t r[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) tmem_ld8_issue(tchunk + cc * 64 + u * 8, r[u]);
          tmem_ld_wait();
          uint4 hi4[4], lo4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int n = nc0 + part * 32 + u * 8;
            f32x2 v[4];
            if (n + 8 <= g.N) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(g.bias + n + 4));
              v[0] = gelu_erf_fast2(add2(pack2(__uint_as_float(r[u][0]), __uint_as_float(r[u][1])), pack2(b0.x, b0.y)));
              v[1] = gelu_erf_fast2(add2(pack2(__uint_as_float(r[u][2]), __uint_as_float(r[u][3])), pack2(b0.z, b0.w)));
              v[2] = gelu_erf_fast2(add2(pack2(__uint_as_float(r[u][4]), __uint_as_float(r[u][5])), pack2(b1.x, b1.y)));
              v[3] = gelu_erf_fast2(add2(pack2(__uint_as_float(r[u][6]), __uint_as_float(r[u][7])), pack2(b1.z, b1.w)));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int na = n + 2 * e, nb = na + 1;
                const float xa = na < g.N ? gelu_erf_fast(__uint_as_float(r[u][2 * e]) + __ldg(g.bias + na)) : 0.f;
                const float xb = nb < g.N ? gelu_erf_fast(__uint_as_float(r[u][2 * e + 1]) + __ldg(g.bias + nb)) : 0.f;
                v[e] = pack2(xa, xb);
              }
            }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x0, x1;
              unpack2(v[e], x0, x1);
              hi[e] = pack_bf16x2(x0, x1);
              const f32x2 hf = pack2(__uint_as_float(hi[e] << 16), __uint_as_float(hi[e] & 0xffff0000u));
              float l0, l1;
              unpack2(add2(v[e], hf ^ 0x8000000080000000ull), l0, l1);
              lo[e] = pack_bf16x2(l0, l1);
            }
            hi4[u] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            lo4[u] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          // the previous chunk's bulk stores must have read the staging buffers before they are overwritten
          if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          g2_epi_sync();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int off = row * 128 + (((part * 4 + u) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(stg + off) = hi4[u];
            if (NPASS == 3) *reinterpret_cast<uint4*>(stg + kG2Tile + off) = lo4[u];
          }
          fence_proxy_async();
          g2_epi_sync();
          if (et == 0) {
            unsigned char* dst = g.c_img + ((long long)(nc0 >> 6) * g.c_rpad + row0) * 128;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(stg)), "n"(kG2Tile) : "memory");
            if (NPASS == 3)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + g.c_plane), "r"(smem_u32(stg + kG2Tile)), "n"(kG2Tile) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        (void)i;
      } else if ((g.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 && !(g.debug & 2)) {
        // ---- fp32 tile, stores coalesced through a per-warp transpose ----
        // A lane owns a row of the accumulator, so direct stores touch 32 lines per warp instruction, half a sector
        // each (the fallback below: ~10 us per tile, which bounds every GEMM with K < ~600).  Instead each warp parks
        // 32 rows x 32 columns in its own 4 KB of the staging area (16-byte units XOR-swizzled by row & 7: conflict-free
        // both ways) and writes them back with 8 lanes per row: every store instruction covers four whole 128-byte lines.
        unsigned char* wst = smem_g2 + Cfg::stg_off + (warp - 4) * 4096;
        const long long row_w0 = (long long)tm * 256 + (long long)cta * 128 + q * 32;     // first row of this warp
        const int rr0 = lane >> 3, uu = lane & 7;
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          if (n_base + c >= g.N) break;
          uint32_t r[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) tmem_ld8_issue(taddr + c + u * 8, r[u]);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int n = n_base + c + u * 8;
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (g.bias) {
              if (n + 8 <= g.N) {
                b0 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
                b1 = __ldg(reinterpret_cast<const float4*>(g.bias + n + 4));
              } else {
                float bb[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) bb[e] = n + e < g.N ? __ldg(g.bias + n + e) : 0.f;
                b0 = make_float4(bb[0], bb[1], bb[2], bb[3]);
                b1 = make_float4(bb[4], bb[5], bb[6], bb[7]);
              }
            }
            *reinterpret_cast<float4*>(wst + lane * 128 + (((2 * u) ^ (lane & 7)) << 4)) =
                make_float4(__uint_as_float(r[u][0]) + b0.x, __uint_as_float(r[u][1]) + b0.y,
                            __uint_as_float(r[u][2]) + b0.z, __uint_as_float(r[u][3]) + b0.w);
            *reinterpret_cast<float4*>(wst + lane * 128 + (((2 * u + 1) ^ (lane & 7)) << 4)) =
                make_float4(__uint_as_float(r[u][4]) + b1.x, __uint_as_float(r[u][5]) + b1.y,
                            __uint_as_float(r[u][6]) + b1.z, __uint_as_float(r[u][7]) + b1.w);
          }
          __syncwarp();
          const int n = n_base + c + uu * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rr0;
            const float4 v = *reinterpret_cast<const float4*>(wst + rr * 128 + ((uu ^ (rr & 7)) << 4));
            const long long gi = row_w0 + rr;
            if (gi < g.M) {
              float* dst = g.C + gi * g.ldc + n;
              if (n + 4 <= g.N) {
                *reinterpret_cast<float4*>(dst) = v;
              } else {
                if (n < g.N) dst[0] = v.x;
                if (n + 1 < g.N) dst[1] = v.y;
                if (n + 2 < g.N) dst[2] = v.z;
              }
            }
          }
          __syncwarp();
        }
        (void)i;
      } else {
      float* crow = g.C + i * g.ldc;
#pragma unroll 1
      for (int c = 0; c < 128; c += 16) {
        uint32_t r0[8], r1[8];
        tmem_ld8_issue(taddr + c, r0);
        tmem_ld8_issue(taddr + c + 8, r1);
        tmem_ld_wait();
        if (i < g.M && !(g.debug & 1)) {
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            const uint32_t* r = h8 == 0 ? r0 : r1;
            const int n = n_base + c + h8 * 8;
            if (n + 8 <= g.N && (g.ldc & 3) == 0) {
              float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
              if (g.bias) { b0 = __ldg(reinterpret_cast<const float4*>(g.bias + n)); b1 = __ldg(reinterpret_cast<const float4*>(g.bias + n + 4)); }
              *reinterpret_cast<float4*>(crow + n) = make_float4(__uint_as_float(r[0]) + b0.x, __uint_as_float(r[1]) + b0.y,
                                                                 __uint_as_float(r[2]) + b0.z, __uint_as_float(r[3]) + b0.w);
              *reinterpret_cast<float4*>(crow + n + 4) = make_float4(__uint_as_float(r[4]) + b1.x, __uint_as_float(r[5]) + b1.y,
                                                                     __uint_as_float(r[6]) + b1.z, __uint_as_float(r[7]) + b1.w);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (n + e < g.N) crow[n + e] = __uint_as_float(r[e]) + (g.bias ? __ldg(g.bias + n + e) : 0.f);
            }
          }
        }
      }
      }
      tc_fence_before();
      g2_epi_sync();
      if (et == 0) mbar_arrive_remote(tmem_empty_leader + 8u * buf);
      if (g.trace && leader && et == 0 && tile_it < 16) g.trace[(pair_id * 16 + tile_it) * 4 + 3] = g2_now();
    }
  }

  if (tid == 128) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // epilogue thread 0: its bulk stores are complete
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

}  // namespace bcnf
