// C ABI of the Transformer condition encoder's kernels (include/bcnf_b200.h: bcnf_trf_*): a translation unit of its own,
// compiled in parallel with bcnf_abi.cu (the attention kernels are instantiated per head width and token bound).
#include "abi_common.cuh"

#include <algorithm>
#include <cmath>

#include "trf.cuh"

using namespace bcnf;

// ---- Transformer condition encoder: the kernels between its Linears (trf.cuh) ------------------------------------------
static int trf_img_check(const char* who, const void* img, int64_t plane, int32_t rpad, int64_t rows, int32_t E) {
  if (!img) return fail(BCNF_E_ARG, "%s: null image", who);
  if (E < 8 || E % 8 || E > 1024) return fail(BCNF_E_ARG, "%s: E=%d must be a multiple of 8 in [8, 1024]", who, E);
  if (rows < 0 || rows > 0x7fffff00LL) return fail(BCNF_E_ARG, "%s: rows=%lld", who, (long long)rows);
  if (rpad % 32 || rpad < rows || plane < (long long)((E + 63) / 64) * rpad * 128)
    return fail(BCNF_E_ARG, "%s: image too small (rpad=%d plane=%lld for %lld rows x %d)", who, rpad, (long long)plane, (long long)rows, E);
  return BCNF_OK;
}

extern "C" int bcnf_trf_embed(const float* tokens, const float* Wf, const float* bf, const float* pos, const float* mask,
                              int64_t rows, int32_t T, int32_t F, int32_t E, float* x, void* x_img, int64_t plane, int32_t rpad,
                              int32_t device, void* stream) {
  NVTX_RANGE("bcnf_trf_embed");
  if (!tokens || !Wf || !bf || !x || T < 1 || F < 1) return fail(BCNF_E_ARG, "bcnf_trf_embed: bad argument");
  if (int rc = trf_img_check("bcnf_trf_embed", x_img, plane, rpad, rows, E)) return rc;
  if (rows == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  TrfEmbedArgs a;
  a.tokens = tokens; a.Wf = Wf; a.bf = bf; a.pos = pos; a.mask = mask; a.x = x;
  a.x_img = (unsigned char*)x_img; a.plane = plane; a.rpad = rpad; a.rows = rows; a.T = T; a.F = F; a.E = E;
  const long long threads = rows * (E / 8);
  trf_embed_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

static int trf_head_check(const char* who, int32_t T, int32_t E, int32_t heads, int t_max) {
  if (T < 1 || T > t_max || heads < 1 || E % heads)
    return fail(BCNF_E_ARG, "%s: bad argument (T=%d must be in [1, %d], E=%d divisible by heads=%d)", who, T, t_max, E, heads);
  const int hd = E / heads;
  if (hd != 8 && hd != 16 && hd != 32 && hd != 64) return fail(BCNF_E_UNSUPPORTED, "%s: head width %d (supported: 8, 16, 32, 64)", who, hd);
  return BCNF_OK;
}

extern "C" int bcnf_trf_attention(const float* qkv, int64_t n_inst, int32_t T, int32_t E, int32_t heads, float* ctx, void* ctx_img,
                                  int64_t plane, int32_t rpad, int32_t device, void* stream) {
  NVTX_RANGE("bcnf_trf_attention");
  if (!qkv || n_inst < 0) return fail(BCNF_E_ARG, "bcnf_trf_attention: bad argument");
  if (int rc = trf_head_check("bcnf_trf_attention", T, E, heads, kTrfMaxT)) return rc;
  if (int rc = trf_img_check("bcnf_trf_attention", ctx_img, plane, rpad, n_inst * T, E)) return rc;
  if (n_inst == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  const int hd = E / heads;
  const size_t smem = sizeof(float) * (size_t)T * 2 * E;
  if (smem > 200 * 1024) return fail(BCNF_E_UNSUPPORTED, "bcnf_trf_attention: T=%d x E=%d does not fit shared memory", T, E);
  void (*kern)(const TrfAttnArgs) = nullptr;
  const bool small = T <= 32;
  // (33..64 tokens keep 64 scores per thread in registers: instantiated for head widths 8 and 16 only)
  if (!small && hd > 16) return fail(BCNF_E_UNSUPPORTED, "bcnf_trf_attention: T=%d > 32 needs a head width <= 16 (got %d)", T, hd);
  int slot = 0;
  switch (hd) {
    case 8: kern = small ? trf_attn_kernel<8, 32> : trf_attn_kernel<8, 64>; slot = 0; break;
    case 16: kern = small ? trf_attn_kernel<16, 32> : trf_attn_kernel<16, 64>; slot = 1; break;
    case 32: kern = trf_attn_kernel<32, 32>; slot = 2; break;
    default: kern = trf_attn_kernel<64, 32>; slot = 3; break;
  }
  slot = 2 * slot + (small ? 0 : 1);
  static size_t configured[64][8] = {};
  if (smem > 48 * 1024 && (device >= 64 || configured[device][slot] < smem)) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (device < 64) configured[device][slot] = smem;
  }
  int n_sm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
  TrfAttnArgs a;
  a.qkv = qkv; a.ctx_img = (unsigned char*)ctx_img; a.plane = plane; a.rpad = rpad; a.ctx = ctx;
  a.n_inst = n_inst; a.T = T; a.E = E; a.heads = heads; a.scale = 1.0f / sqrtf((float)hd);
  const long long grid = std::min<long long>(n_inst, (long long)n_sm * 32);
  kern<<<(unsigned)grid, kTrfAttnThreads, smem, (cudaStream_t)stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

extern "C" int bcnf_trf_attention_bwd(const float* qkv, const float* dctx, int64_t n_inst, int32_t T, int32_t E, int32_t heads,
                                      float* dqkv, int32_t device, void* stream) {
  NVTX_RANGE("bcnf_trf_attention_bwd");
  if (!qkv || !dctx || !dqkv || n_inst < 0 || E % 4) return fail(BCNF_E_ARG, "bcnf_trf_attention_bwd: bad argument");
  if (int rc = trf_head_check("bcnf_trf_attention_bwd", T, E, heads, 32)) return rc;
  if (n_inst == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  const int hd = E / heads;
  // heads per CTA: half an instance's heads when that leaves whole warps' worth (two half-size CTAs per instance fill the
  // SMs in one wave at batch 256), never more than the 8 warps of a block
  int hpg = heads;
  if (heads > 4 && heads % 2 == 0) hpg = heads / 2;
  while (hpg > kTrfAttnThreads / 32) {
    if (hpg % 2) return fail(BCNF_E_UNSUPPORTED, "bcnf_trf_attention_bwd: %d heads cannot be grouped into blocks of <= 8", heads);
    hpg /= 2;
  }
  const int groups = heads / hpg;
  const size_t smem = sizeof(float) * ((size_t)4 * T * hpg * hd + (size_t)hpg * 2 * T * (T + 1));
  if (smem > 200 * 1024) return fail(BCNF_E_UNSUPPORTED, "bcnf_trf_attention_bwd: T=%d x E=%d does not fit shared memory", T, E);
  void (*kern)(const TrfAttnBwdArgs) = hd == 8 ? trf_attn_bwd_kernel<8> : hd == 16 ? trf_attn_bwd_kernel<16>
                                       : hd == 32 ? trf_attn_bwd_kernel<32> : trf_attn_bwd_kernel<64>;
  const int slot = hd == 8 ? 0 : hd == 16 ? 1 : hd == 32 ? 2 : 3;
  static size_t configured[64][4] = {};
  if (smem > 48 * 1024 && (device >= 64 || configured[device][slot] < smem)) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (device < 64) configured[device][slot] = smem;
  }
  int n_sm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
  TrfAttnBwdArgs a;
  a.qkv = qkv; a.dctx = dctx; a.dqkv = dqkv; a.n_inst = n_inst; a.T = T; a.E = E; a.heads = heads; a.hpg = hpg;
  a.scale = 1.0f / sqrtf((float)hd);
  const long long grid = std::min<long long>(n_inst, (long long)n_sm * 8);
  kern<<<dim3((unsigned)grid, (unsigned)groups), 32 * hpg, smem, (cudaStream_t)stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

extern "C" int bcnf_trf_ln_param_grad(const float* g, const float* s, const float* mean, const float* rstd, int64_t rows,
                                      int32_t E, float* dgamma, float* dbeta, int32_t device, void* stream) {
  NVTX_RANGE("bcnf_trf_ln_param_grad");
  if (!g || !s || !mean || !rstd || !dgamma || !dbeta || rows < 0 || E < 1) return fail(BCNF_E_ARG, "bcnf_trf_ln_param_grad: bad argument");
  if (rows == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  TrfLnParamGradArgs a;
  a.g = g; a.s = s; a.mean = mean; a.rstd = rstd; a.dgamma = dgamma; a.dbeta = dbeta; a.rows = rows; a.E = E; a.slab = 32;
  const long long blocks = (rows + a.slab - 1) / a.slab;
  if (blocks > 0x7fffffffLL) return fail(BCNF_E_ARG, "bcnf_trf_ln_param_grad: rows=%lld", (long long)rows);
  trf_ln_param_grad_kernel<<<(unsigned)blocks, E <= 128 ? 128 : 256, 0, (cudaStream_t)stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

extern "C" int bcnf_trf_add_layernorm(const float* x, const float* y, const float* mask, const float* gamma, const float* beta,
                                      float eps, int64_t rows, int32_t E, float* s, float* mean, float* rstd, float* x_out,
                                      void* x_img, int64_t plane, int32_t rpad, int32_t device, void* stream) {
  NVTX_RANGE("bcnf_trf_add_layernorm");
  if (!x || !y || !gamma || !beta || !x_out || (mean == nullptr) != (rstd == nullptr))
    return fail(BCNF_E_ARG, "bcnf_trf_add_layernorm: null argument");
  if (int rc = trf_img_check("bcnf_trf_add_layernorm", x_img, plane, rpad, rows, E)) return rc;
  if (rows == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  TrfAddLnArgs a;
  a.x = x; a.y = y; a.mask = mask; a.gamma = gamma; a.beta = beta; a.s = s; a.mean = mean; a.rstd = rstd; a.x_out = x_out;
  a.x_img = (unsigned char*)x_img; a.plane = plane; a.rpad = rpad; a.rows = rows; a.E = E; a.eps = eps;
  trf_add_ln_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

extern "C" int bcnf_trf_gelu(const float* u, int64_t rows, int32_t N, float* a_out, void* a_img, int64_t plane, int32_t rpad,
                             int32_t device, void* stream) {
  NVTX_RANGE("bcnf_trf_gelu");
  if (!u) return fail(BCNF_E_ARG, "bcnf_trf_gelu: null argument");
  if (int rc = trf_img_check("bcnf_trf_gelu", a_img, plane, rpad, rows, N)) return rc;
  if (rows == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  TrfGeluArgs g;
  g.u = u; g.a = a_out; g.a_img = (unsigned char*)a_img; g.plane = plane; g.rpad = rpad; g.rows = rows; g.N = N;
  const long long threads = rows * (N / 8);
  trf_gelu_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

