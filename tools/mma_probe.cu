// Microbenchmark: cycles per tcgen05.mma for the operand shapes the fused kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_probe tools/mma_probe.cu && tools/mma_probe
// Operands are zero-filled shared-memory tiles (K-major SWIZZLE_128B); one or several threads of the leader
// CTA issue R MMAs back to back, commit, and wait; cycles = (clock64 at completion - at first issue) / R.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define BCNF_PROBE
#include "../bcnf_b200/csrc/flow_tc.cuh"
using namespace bcnf;

// mode 0: cta_group::2, M=128 (64 rows per CTA);  mode 1: cta_group::1, M=128 (128 rows in one CTA)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

// uniform = 1: every lane of the issuing warp runs the loop (operands stay warp-uniform), one elected lane issues
__global__ void __launch_bounds__(256, 1) probe(int mode, int N, int R, int issuers, int uniform, long long* out, uint32_t tptr_const) {
  const int same_b = 0;
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tptr;
  // canonical warp index: a lane-0 broadcast makes the value provably warp-uniform for ptxas
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    if (mode == 0) { asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "n"(512));
                     asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;"); }
    else           { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "n"(512));
                     asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  // a 512-column allocation can only start at column 0 of lane 0: use the constant so that the address is
  // provably warp-uniform (a value loaded from shared memory makes ptxas wrap every MMA in a uniformisation loop)
  if (threadIdx.x == 0 && cta == 0) out[7] = tptr;
  const uint32_t tm = tptr_const;
  const uint32_t a_addr = smem_u32(sm), b_addr = smem_u32(sm) + 64 * 1024;   // A: 64 KB region, B: 128 KB region
  if (cta == 0 && (uniform || lane == 0) && warp >= 1 && warp <= issuers) {
    const int j = warp - 1;
    const uint32_t idesc = make_idesc(N);
    const long long t0 = clock64();
    for (int r = 0; r < R; ++r) {
      const int ka = r & 3, ta = (r >> 2) & 3;                    // rotate over 4 K slices x 4 tiles
      const uint64_t ad = make_smem_desc(a_addr + ta * 16384) + 2 * ka;
      const int tb = same_b ? 0 : ((r >> 2) & 3);
      const uint64_t bd = make_smem_desc(b_addr + tb * 32768) + 2 * ka;
      if (mode == 0) umma_2sm(tm + j * 128, ad, bd, idesc, r > 0);
      else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                        ::"r"(tm + j * 128), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(r > 0)) : "memory");
    }
    const long long t1 = clock64();
    if (lane == 0) {
      if (mode == 0) umma_commit_2sm(&bar[j], 1);
      else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[j])) : "memory");
    }
    if (uniform) __syncwarp();
    mbar_wait(&bar[j], 0);
    const long long t2 = clock64();
    if (lane == 0) { out[j * 2 + 0] = t1 - t0; out[j * 2 + 1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (warp == 0) {
    if (mode == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512));
  }
}

int main(int argc, char** argv) {
  // usage: mma_probe mode N issuers uniform
  if (argc < 5) { printf("usage: mma_probe mode N issuers uniform\n"); return 2; }
  const int mode = atoi(argv[1]), N = atoi(argv[2]), issuers = atoi(argv[3]), uniform = atoi(argv[4]);
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int R = 2000;
  cudaLaunchConfig_t cfg{}; cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = mode == 0 ? 1 : 0; cfg.gridDim = dim3(mode == 0 ? 2 : 1); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 200 * 1024;
  cudaMemset(d, 0, 64);
  cudaError_t e = cudaLaunchKernelEx(&cfg, probe, mode, N, R, issuers, uniform, d, (uint32_t)0);
  if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1; }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  const double ideal = mode == 0 ? 128.0 * N / 512.0 : 128.0 * N / 256.0;
  printf("mode %d N %3d issuers %d uniform %d | issue %7.1f cyc/MMA | done %7.1f cyc/MMA per issuer => %.1f overall | ideal %.0f\n",
         mode, N, issuers, uniform, (double)h[0] / R, (double)h[1] / R, (double)h[1] / R / issuers, ideal);
  printf("  tmem base returned by tcgen05.alloc: 0x%llx\n", h[7]);
  return 0;
}
