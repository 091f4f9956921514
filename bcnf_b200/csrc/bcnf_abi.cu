// C ABI of bcnf_b200 (include/bcnf_b200.h): handle management, parameter packing, dispatch.
#include "../../include/bcnf_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "abi_common.cuh"
#include "common.cuh"
#include "cond_project.cuh"
#include "flow_rowthread.cuh"
#include "flow_tc.cuh"
#include "flow_tiled.cuh"
#include "train_ops.cuh"
#include "train_tc.cuh"
#include "train_glue.cuh"
#include "gemm_img2.cuh"
#include "flow_tc2.cuh"
#include "resim.cuh"

using namespace bcnf;

// ----------------------------------------------------------------------------------------------
// error reporting (abi_common.cuh: bcnf_fail, NVTX_RANGE, DEVICE_GUARD, CUDA_TRY)
// ----------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int bcnf_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ----------------------------------------------------------------------------------------------
// parameter packing: one table-driven kernel per set_params call
// ----------------------------------------------------------------------------------------------
struct PackDesc {
  const float* src;
  float* dst;
  int src_pitch, rows, cols, dst_pitch;
  int mode;  // 0 copy: dst[r*dp + c] = src[r*sp + c]; 1 transpose: dst[c*dp + r] = src[r*sp + c];
             // 2 actnorm log-det: dst[0] = sum_r log|src[r]|
  int pad;
};

__global__ void pack_kernel(const PackDesc* __restrict__ descs) {
  const PackDesc d = descs[blockIdx.x];
  if (d.mode == 2) {
    if (blockIdx.y == 0 && threadIdx.x == 0) {
      float s = 0.f;   // torch.sum(torch.log(torch.abs(scale))), cnf.py:350
      for (int r = 0; r < d.rows; ++r) s += logf(fabsf(d.src[r]));
      d.dst[0] = s;
    }
    return;
  }
  const long long n = (long long)d.rows * d.cols;
  for (long long e = (long long)blockIdx.y * blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.y * blockDim.x) {
    const int r = (int)(e / d.cols), c = (int)(e - (long long)r * d.cols);
    const float v = d.src[(size_t)r * d.src_pitch + c];
    if (d.mode == 0) d.dst[(size_t)r * d.dst_pitch + c] = v;
    else             d.dst[(size_t)c * d.dst_pitch + r] = v;
  }
}

// ----------------------------------------------------------------------------------------------
// handle
// ----------------------------------------------------------------------------------------------
struct Program {
  std::vector<DevOp> ops;
  std::vector<Chunk> chunks;
  DevOp* d_ops = nullptr;
  Chunk* d_chunks = nullptr;
  float* d_blob = nullptr;
  long long blob_floats = 0;
  int max_chunk_bytes = 0;
};

struct bcnf_flow {
  bcnf_flow_desc_t desc;
  std::vector<int> op_types;
  StackDims sd;
  int n_half = 0;
  Program prog[2];  // 0 forward, 1 inverse
  float* d_wproj = nullptr;  // (C, PW)
  float* d_bproj = nullptr;  // (PW)
  PackDesc* d_pack = nullptr;
  PackDesc* h_pack = nullptr;  // pinned
  int pack_cap = 0;
  int kernel = BCNF_KERNEL_TILED;
  int rows_per_cta = 0;
  int num_sms = 148;
  int max_smem_optin = 0;
  bool params_set = false;
  long long packed_bytes = 0;
  long long macs_row = 0, macs_inst = 0;
  // tiled launch configuration
  TiledSmem tiled_lay;
  int tiled_R = 32;
  // tcgen05 path
  TcDims td;
  int npass = 0;                       // 0 = not a tensor-core handle
  unsigned char* d_tc_blob = nullptr;  // bf16 hi/lo weight tile images (direction independent)
  long long tc_blob_bytes = 0;
  long long* d_tc_off[2] = {nullptr, nullptr};   // per direction, per device op: byte offset of its tile stream
  std::vector<long long> tc_off_by_layer;        // per (layer index, net 0/1): offset in d_tc_blob
  TcPackDesc* d_tc_pack = nullptr;
  TcPackDesc* h_tc_pack = nullptr;
  int tc_pack_cap = 0;
  // CTA-pair GEMM on operand images (gemm_img2.cuh): image of Wproj (rows = projection column, k = condition
  // feature), rebuilt by set_params, and a scratch image of the h rows of one slice of instances
  unsigned char* d_wproj_img = nullptr;
  long long wproj_plane = 0;
  int wproj_rpad = 0;
  unsigned char* d_h_img = nullptr;
  long long h_img_bytes = 0;
  // second-generation fused kernel (flow_tc2.cuh): weight images of every Linear of every conditioner network, the
  // per-direction table of their offsets, and one activation scratch per stream that has run the kernel
  // switches read from the environment ONCE, at create (tests / A-B timing / debugging; never on the launch path)
  int env_rowthread_r = 0;                    // BCNF_ROWTHREAD_R: rows per thread of the row-per-thread kernel (1 or 2)
  bool env_proj_fma = false;                  // BCNF_PROJ_FMA=1: fp32 FMA projection kernel on a tensor-core handle
  int env_tc2_debug = 0;                      // BCNF_TC2_DEBUG: timing experiments of flow_tc2 (wrong results)
  std::string env_tc2_trace, env_tc_trace;    // BCNF_TC2_TRACE / BCNF_TC_TRACE: file that receives a pipeline trace
  bool tc1_ok = false;                        // the first-generation kernel's plan (f.td) is valid
  bool s2_ok = false;
  bool s2_use = false;                        // dispatch forward / inverse to it (decided once, at create)
  S2Dims s2;
  unsigned char* d_s2_img = nullptr;
  long long s2_img_bytes = 0;
  long long* d_s2_off[2] = {nullptr, nullptr};
  struct S2Scratch { cudaStream_t stream; unsigned char* act; };
  std::vector<S2Scratch> s2_scratch;
  unsigned int* d_s2_dbg = nullptr;            // device alias of h_s2_dbg (mapped pinned memory: readable after a trap)
  unsigned int* h_s2_dbg = nullptr;
  S2BiasCol* d_s2_bias = nullptr;
  int s2_ctas = 0;
};

static const int kRowThreadChunkCap = 20 * 1024;  // bytes per streamed parameter chunk

static void build_half_layout(HalfLayout& hl, int din, int dout, int L, const int* hidden) {
  memset(&hl, 0, sizeof(hl));
  hl.din = din; hl.dinp = round_up(din, 4);
  hl.dout = dout; hl.dop = round_up(dout, 4);
  hl.L = L;
  int off = 0;
  for (int l = 0; l < L; ++l) { hl.h[l] = hidden[l]; hl.hp[l] = round_up(hidden[l], 16); }
  hl.off_w[0] = off; off += hl.dinp * hl.hp[0];
  for (int l = 1; l < L; ++l) {
    hl.off_w[l] = off; off += hl.hp[l - 1] * hl.hp[l];
    hl.off_b[l] = off; off += hl.hp[l];
  }
  hl.off_wout = off; off += hl.hp[L - 1] * 2 * hl.dop;
  hl.off_bout = off; off += 2 * hl.dop;
  hl.total = round_up(off, 4);
}

static int op_param_floats(const bcnf_flow& f, const DevOp& op) {
  switch (op.type) {
    case DOP_ACTNORM_FWD:
    case DOP_ACTNORM_INV: return 2 * f.sd.DP + 4;
    case DOP_MIX: return f.sd.D * f.sd.DP;
    default: return f.sd.half[op.src].total;
  }
}

// Compile the layer list into the device program of one direction.
static void build_program(bcnf_flow& f, int dir) {
  Program& p = f.prog[dir];
  p.ops.clear();
  const int n = (int)f.op_types.size();
  // projection slice of each conditioner network, in forward layer order
  std::vector<int> proj_a(n, -1), proj_b(n, -1);
  int pw = 0;
  for (int i = 0; i < n; ++i)
    if (f.op_types[i] == BCNF_OP_COUPLING) {
      proj_a[i] = pw; pw += f.sd.half[0].hp[0];
      if (f.desc.two_way) { proj_b[i] = pw; pw += f.sd.half[1].hp[0]; }
    }
  f.sd.PW = pw;
  for (int s = 0; s < n; ++s) {
    const int i = dir == 0 ? s : n - 1 - s;   // inverse walks reversed(self.layers), cnf.py:500
    DevOp op{};
    switch (f.op_types[i]) {
      case BCNF_OP_ACTNORM:
        op.type = dir == 0 ? DOP_ACTNORM_FWD : DOP_ACTNORM_INV;
        p.ops.push_back(op);
        break;
      case BCNF_OP_ORTHO:
        op.type = DOP_MIX;
        p.ops.push_back(op);
        break;
      default:
        // forward: nn_a then nn_b (cnf.py:178-184); the reference's inverse ALSO runs nn_a
        // first, on z_a, and then nn_b on the recovered y_b (cnf.py:203-208) -- reproduced.
        op.type = DOP_HALF; op.inverse = dir; op.src = 0; op.proj_off = proj_a[i];
        p.ops.push_back(op);
        if (f.desc.two_way) { op.src = 1; op.proj_off = proj_b[i]; p.ops.push_back(op); }
        break;
    }
  }
  long long off = 0;
  for (auto& op : p.ops) { op.off = off; off += op_param_floats(f, op); }
  p.blob_floats = off;
  // chunks for the streaming kernel: greedy runs of ops up to the byte cap
  p.chunks.clear();
  p.max_chunk_bytes = 0;
  Chunk cur{0, 0, 0, 0, 0};
  for (int i = 0; i < (int)p.ops.size(); ++i) {
    const int bytes = op_param_floats(f, p.ops[i]) * 4;
    if (cur.n_ops > 0 && cur.bytes + bytes > kRowThreadChunkCap) {
      p.chunks.push_back(cur);
      cur = Chunk{p.ops[i].off, 0, i, 0, 0};
    }
    cur.bytes += bytes;
    cur.n_ops += 1;
    if (cur.bytes > p.max_chunk_bytes) p.max_chunk_bytes = cur.bytes;
  }
  if (cur.n_ops > 0) p.chunks.push_back(cur);
}

static void free_program(Program& p) {
  if (p.d_ops) cudaFree(p.d_ops);
  if (p.d_chunks) cudaFree(p.d_chunks);
  if (p.d_blob) cudaFree(p.d_blob);
  p.d_ops = nullptr; p.d_chunks = nullptr; p.d_blob = nullptr;
}

extern "C" int bcnf_abi_version(void) { return BCNF_ABI_VERSION; }
extern "C" const char* bcnf_last_error(void) { return g_err; }

extern "C" int bcnf_flow_destroy(bcnf_flow_t* f) {
  if (!f) return BCNF_OK;
  DeviceGuard guard(f->desc.device);
  free_program(f->prog[0]);
  free_program(f->prog[1]);
  if (f->d_wproj) cudaFree(f->d_wproj);
  if (f->d_bproj) cudaFree(f->d_bproj);
  if (f->d_pack) cudaFree(f->d_pack);
  if (f->h_pack) cudaFreeHost(f->h_pack);
  if (f->d_tc_blob) cudaFree(f->d_tc_blob);
  for (int d = 0; d < 2; ++d) if (f->d_tc_off[d]) cudaFree(f->d_tc_off[d]);
  if (f->d_tc_pack) cudaFree(f->d_tc_pack);
  if (f->h_tc_pack) cudaFreeHost(f->h_tc_pack);
  if (f->d_wproj_img) cudaFree(f->d_wproj_img);
  if (f->d_h_img) cudaFree(f->d_h_img);
  if (f->d_s2_img) cudaFree(f->d_s2_img);
  for (int d = 0; d < 2; ++d) if (f->d_s2_off[d]) cudaFree(f->d_s2_off[d]);
  for (auto& sc : f->s2_scratch) if (sc.act) cudaFree(sc.act);
  if (f->h_s2_dbg) cudaFreeHost(f->h_s2_dbg);
  if (f->d_s2_bias) cudaFree(f->d_s2_bias);
  delete f;
  return BCNF_OK;
}

// ---- tcgen05 path: layer / chunk structure of one conditioner network -------------------------------
static void tc_set_chunks(TcLayer& ly) {
  const int n = (ly.np + 255) / 256;
  const int base = round_up((ly.np + n - 1) / n, 16);
  ly.n_chunks = n;
  for (int i = 0; i < 4; ++i) ly.chunk_n[i] = 0;
  for (int i = 0, left = ly.np; i < n; ++i) { ly.chunk_n[i] = std::min(base, left); left -= ly.chunk_n[i]; }
}

static void tc_build_half(TcHalfLayout& tl, const HalfLayout& hl, int kw) {
  memset(&tl, 0, sizeof(tl));
  tl.L = hl.L;
  tl.doh = round_up(hl.dout, 8);
  for (int l = 0; l <= hl.L; ++l) {
    TcLayer& ly = tl.layer[l];
    const int k_real = l == 0 ? hl.din : hl.h[l - 1];
    ly.np = l == hl.L ? 2 * tl.doh : hl.hp[l];
    ly.kc = (k_real + kw - 1) / kw;
    ly.last_ksteps = (k_real - kw * (ly.kc - 1) + 15) / 16;
    tc_set_chunks(ly);
  }
  long long bytes = 0;
  for (int l = 0; l <= hl.L; ++l)
    for (int nc = 0; nc < tl.layer[l].n_chunks; ++nc)
      bytes += (long long)tl.layer[l].kc * 4 * (tl.layer[l].chunk_n[nc] / 2) * (2 * kw);
  tl.stream_bytes = bytes;
}

// Returns 0 and fills f.td if the stack fits the tensor-core kernel, else a reason string.
static const char* tc_plan(bcnf_flow& f, int npass) {
  const StackDims& sd = f.sd;
  TcDims& td = f.td;
  memset(&td, 0, sizeof(td));
  int a_chunks = 1, max_half_rows = 8, max_doh = 8;
  td.kw = 64;                 // weight tiles are 64 K columns wide (SWIZZLE_128B)
  for (int s = 0; s < 2; ++s) {
    const HalfLayout& hl = sd.half[s];
    if (hl.din > 64) return "own-half width > 64";
    for (int l = 0; l < hl.L; ++l) {
      if (hl.h[l] < 48) return "hidden width < 48: the row-per-thread / tiled FMA kernels are the better fit";
      if (hl.hp[l] > 1024) return "hidden width > 1024";
    }
    tc_build_half(td.half[s], hl, td.kw);
    for (int l = 0; l <= hl.L; ++l) {
      const TcLayer& ly = td.half[s].layer[l];
      if (l >= 1) a_chunks = std::max(a_chunks, (hl.h[l - 1] + 63) / 64);
      for (int nc = 0; nc < ly.n_chunks; ++nc) max_half_rows = std::max(max_half_rows, ly.chunk_n[nc] / 2);
    }
    max_doh = std::max(max_doh, td.half[s].doh);
  }
  td.a_chunks = a_chunks;
  td.stage_bytes = round_up(max_half_rows * 2 * td.kw * (npass == 3 ? 2 : 1), 1024);
  td.off_alo = npass == 3 ? a_chunks * kTcATile : 0;
  td.off_stage = (npass == 3 ? 2 : 1) * a_chunks * kTcATile;
  td.yp = sd.DP;
  td.tsp = 2 * max_doh;
  const int tail = round_up(kTcRows * td.yp * 4, 16) + round_up(kTcRows * td.tsp * 4, 16) + 2048;
  const int room = f.max_smem_optin - td.off_stage - tail;
  td.n_stages = std::min(8, room / td.stage_bytes);
  if (td.n_stages < 2) return "activation tiles + 2 weight stages exceed shared memory";
  td.off_y = td.off_stage + td.n_stages * td.stage_bytes;
  td.off_ts = td.off_y + round_up(kTcRows * td.yp * 4, 16);
  td.off_misc = td.off_ts + round_up(kTcRows * td.tsp * 4, 16);
  td.smem_bytes = td.off_misc + 2048;
  {
    int max_half_cols = 8;
    for (int s2 = 0; s2 < 2; ++s2)
      for (int l = 0; l <= td.half[s2].L; ++l)
        for (int nc = 0; nc < td.half[s2].layer[l].n_chunks; ++nc)
          max_half_cols = std::max(max_half_cols, td.half[s2].layer[l].chunk_n[nc] / 2);
    td.region_cols = max_half_cols;
    td.regions = std::min(8, 512 / max_half_cols);
    if (td.regions < 4) return "accumulator slots do not fit in TMEM";     // every layer has <= 4 N chunks
  }
  td.n_halfops = f.n_half;
  td.two_way = f.desc.two_way ? 1 : 0;
  return nullptr;
}


// ---- second-generation fused kernel (flow_tc2.cuh): layer / chunk structure, shared-memory carve-up ---------------
static void s2_set_chunks(S2Layer& ly, int np) {
  // N chunks: a pattern of widths (multiples of 64, so that every chunk starts on an image chunk), the last one taking
  // what is left.  Every tcgen05.mma re-reads its 128 A rows whatever its N, so the split does not change the operand
  // traffic of a layer; it decides what overlaps what (flow_tc2.cuh: the accumulator of chunk c + 2 needs the epilogue
  // of chunk c, the next layer's last K stages need the epilogue of the last chunk).  BCNF_TC2_SPLIT="256,192" (read
  // when the handle is created) overrides the default for experiments.
  int pat[kS2MaxChunks] = {256, 256, 256, 256};
  if (const char* e = getenv("BCNF_TC2_SPLIT")) {
    int i = 0;
    for (const char* q = e; *q && i < kS2MaxChunks; ++i) {
      const int v = atoi(q);
      if (v >= 64 && v <= 256 && v % 64 == 0) pat[i] = v;
      while (*q && *q != ',') ++q;
      if (*q == ',') ++q;
    }
    for (; i > 0 && i < kS2MaxChunks; ++i) pat[i] = pat[i - 1];
  }
  for (int i = 0; i < kS2MaxChunks; ++i) ly.chunk_n[i] = 0;
  int n = 0, left = np;
  while (left > 0 && n < kS2MaxChunks) { ly.chunk_n[n] = std::min(pat[n], left); left -= ly.chunk_n[n]; ++n; }
  ly.n_chunks = left > 0 ? kS2MaxChunks + 1 : n;       // too many chunks: s2_plan refuses
}

// Returns 0 and fills f.s2 if the stack fits the kernel, else a reason string.
static const char* s2_plan(bcnf_flow& f, int npass) {
  const StackDims& sd = f.sd;
  S2Dims& d = f.s2;
  memset(&d, 0, sizeof(d));
  if (sd.D > kS2YPitch - 1) return "size > 24";
  {
    // ActNorm / mixing parameters between two conditioner networks are staged in shared memory: longest run
    long long run = 0, best = 0;
    for (const auto& op : f.prog[0].ops) {
      if (op.type == DOP_HALF) { run = 0; continue; }
      run += op.type == DOP_MIX ? (long long)sd.D * sd.DP : 2 * sd.DP + 4;
      best = std::max(best, run);
    }
    if (best * 4 > kS2GparBytes) return "ActNorm / mixing parameters between two couplings exceed the staging area";
  }
  const int PL = npass == 3 ? 2 : 1;
  int a_kchunks = 1;
  for (int s = 0; s < 2; ++s) {
    const HalfLayout& hl = sd.half[s];
    S2Half& tl = d.half[s];
    if (hl.din > 16) return "own-half width > 16";
    if (2 * hl.dop > kS2TsPitch - 1) return "last Linear wider than 24 columns";
    if (hl.L + 1 > kTcMaxLayers) return "too many hidden layers";
    tl.L = hl.L;
    tl.n_last = round_up(2 * hl.dop, 16);
    long long off = 0;
    for (int l = 0; l <= hl.L; ++l) {
      S2Layer& ly = tl.layer[l];
      if (l < hl.L && (hl.h[l] < 48 || hl.hp[l] > 1024)) return "hidden width outside [48, 1024]";
      // bias folded into the GEMM where the previous layer's width leaves a free padding column (526 -> 528): the
      // input image carries a constant 1 in column h and the weight image the bias there (no bias reads in the epilogue)
      const bool fold = l >= 1 && hl.h[l - 1] % 16 != 0 && !getenv("BCNF_TC2_NOFOLD");
      ly.bias_k = fold ? hl.h[l - 1] : -1;
      ly.one_col = -1;
      if (fold) tl.layer[l - 1].one_col = hl.h[l - 1];
      const int k_real = l == 0 ? hl.din : hl.h[l - 1] + (fold ? 1 : 0);
      const int np = l == hl.L ? tl.n_last : hl.hp[l];
      s2_set_chunks(ly, np);
      if (ly.n_chunks > kS2MaxChunks) return "too many N chunks";
      ly.n_kst = (k_real + 63) / 64;
      ly.last_ksteps = (k_real - 64 * (ly.n_kst - 1) + 15) / 16;
      ly.n_img = l < hl.L ? (np + 63) / 64 : 0;
      ly.w_rpad = round_up(np, 32);
      ly.w_plane = (long long)ly.n_kst * ly.w_rpad * 128;
      ly.w_off = off;
      off += 2 * ly.w_plane;
      a_kchunks = std::max(a_kchunks, ly.n_img);
    }
    tl.net_bytes = off;
  }
  d.n_halfops = f.n_half;
  d.two_way = f.desc.two_way ? 1 : 0;
  d.a_kchunks = a_kchunks;
  d.a_plane = (long long)a_kchunks * kS2Tile;
  d.b_off = PL * kS2Tile;
  d.b_lo_off = d.b_off + kS2Tile;
  d.stage_bytes = 2 * PL * kS2Tile;
  d.stg_off = kS2Stages * d.stage_bytes;
  d.misc_off = d.stg_off + 2 * kS2Tile;
  d.smem_bytes = d.misc_off + kS2MiscBytes;
  d.cta_bytes = 2LL * PL * d.a_plane;
  if (d.smem_bytes > f.max_smem_optin) return "stages + staging exceed shared memory";
  return nullptr;
}

template <typename K>
static int opt_in_smem(K kernel, size_t bytes) {
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}
// cudaFuncSetAttribute is per device: remember what each device has been opted in to (one table per kernel instance)
template <typename K>
static int opt_in_smem_once(K kernel, size_t bytes, size_t (&configured)[64]) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || configured[dev] < bytes) {
    if (int rc = opt_in_smem(kernel, bytes)) return rc;
    if (dev >= 0 && dev < 64) configured[dev] = bytes;
  }
  return 0;
}

extern "C" int bcnf_flow_create(const bcnf_flow_desc_t* desc, const int32_t* op_types, bcnf_flow_t** out) {
  NVTX_RANGE("bcnf_flow_create");
  if (!desc || !op_types || !out) return fail(BCNF_E_ARG, "bcnf_flow_create: null argument");
  *out = nullptr;
  if (desc->size < 2 || desc->size > BCNF_MAX_SIZE)
    return fail(BCNF_E_UNSUPPORTED, "size=%d outside [2, %d]", desc->size, BCNF_MAX_SIZE);
  if (desc->n_conditions < 1)
    return fail(BCNF_E_UNSUPPORTED, "n_conditions=%d: the reference's unconditional path is broken "
                                    "(cnf.py:480-485) and is not provided", desc->n_conditions);
  if (desc->n_hidden < 1 || desc->n_hidden > BCNF_MAX_HIDDEN_LAYERS)
    return fail(BCNF_E_UNSUPPORTED, "len(nested_sizes)=%d outside [1, %d]", desc->n_hidden, BCNF_MAX_HIDDEN_LAYERS);
  for (int l = 0; l < desc->n_hidden; ++l)
    if (desc->hidden[l] < 1 || desc->hidden[l] > BCNF_MAX_HIDDEN)
      return fail(BCNF_E_UNSUPPORTED, "nested_sizes[%d]=%d outside [1, %d]", l, desc->hidden[l], BCNF_MAX_HIDDEN);
  if (desc->n_ops < 1) return fail(BCNF_E_ARG, "empty layer list");
  if (desc->precision < BCNF_PREC_FP32 || desc->precision > BCNF_PREC_BF16)
    return fail(BCNF_E_ARG, "unknown precision %d", desc->precision);
  for (int i = 0; i < desc->n_ops; ++i)
    if (op_types[i] < 0 || op_types[i] > 2)
      return fail(BCNF_E_ARG, "layer %d has unknown type %d", i, op_types[i]);   // cnf.py:485

  DEVICE_GUARD(desc->device);
  bcnf_flow* f = new (std::nothrow) bcnf_flow();
  if (!f) return fail(BCNF_E_NOMEM, "out of host memory");
  f->desc = *desc;
  f->op_types.assign(op_types, op_types + desc->n_ops);
  StackDims& sd = f->sd;
  memset(&sd, 0, sizeof(sd));
  sd.D = desc->size; sd.DP = round_up(sd.D, 4);
  sd.Da = (sd.D + 1) / 2; sd.Db = sd.D / 2;
  sd.C = desc->n_conditions;
  build_half_layout(sd.half[0], sd.Da, sd.Db, desc->n_hidden, desc->hidden);   // nn_a, cnf.py:136
  build_half_layout(sd.half[1], sd.Db, sd.Da, desc->n_hidden, desc->hidden);   // nn_b, cnf.py:150
  build_program(*f, 0);
  build_program(*f, 1);
  f->n_half = 0;
  for (auto& op : f->prog[0].ops) f->n_half += op.type == DOP_HALF;

  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, desc->device));
  f->num_sms = prop.multiProcessorCount;
  f->max_smem_optin = (int)prop.sharedMemPerBlockOptin;

  // algorithmic work
  {
    long long per_half[2];
    for (int s = 0; s < 2; ++s) {
      const HalfLayout& hl = sd.half[s];
      long long m = (long long)hl.din * hl.h[0];
      for (int l = 1; l < hl.L; ++l) m += (long long)hl.h[l - 1] * hl.h[l];
      m += (long long)hl.h[hl.L - 1] * 2 * hl.dout;
      per_half[s] = m;
    }
    f->macs_row = 0; f->macs_inst = 0;
    for (auto& op : f->prog[0].ops) {
      if (op.type == DOP_HALF) { f->macs_row += per_half[op.src]; f->macs_inst += (long long)sd.C * sd.half[op.src].h[0]; }
      else if (op.type == DOP_MIX) f->macs_row += (long long)sd.D * sd.D;
      else f->macs_row += sd.D;
    }
  }

  // kernel selection
  bool uniform = true;
  for (int l = 1; l < desc->n_hidden; ++l) uniform &= desc->hidden[l] == desc->hidden[0];
  const int hp0 = sd.half[0].hp[0];
  const char* force = getenv("BCNF_FORCE_KERNEL");   // "tiled": exercise the generic path on narrow stacks
  const bool allow_rowthread = !(force && strcmp(force, "tiled") == 0);
  const bool rowthread_ok = allow_rowthread && uniform && (hp0 == 16 || hp0 == 32) && (sd.D == 19 || sd.D == 21) &&
                            std::max(f->prog[0].max_chunk_bytes, f->prog[1].max_chunk_bytes) <= kRowThreadChunkCap;
  if (desc->precision != BCNF_PREC_FP32 && allow_rowthread) {   // BCNF_FORCE_KERNEL=tiled pins the fp32 tiled kernel
    const int npass = desc->precision == BCNF_PREC_BF16X3 ? 3 : 1;
    if (prop.major != 10) { delete f; return fail(BCNF_E_UNSUPPORTED, "tcgen05 path needs an sm_100 device (got sm_%d%d)", prop.major, prop.minor); }
    // two generations of the fused kernel: the second (128 rows per CTA, activations through L2, flow_tc2.cuh) where the
    // stack fits it, else the first (64 rows per CTA, activations in shared memory, flow_tc.cuh);
    // BCNF_FLOW_TC=1 pins the first generation (A/B timing, tests of both).  Read once, here.
    const char* why1 = tc_plan(*f, npass);
    const char* why2 = s2_plan(*f, npass);
    if (why1 && why2) { delete f; return fail(BCNF_E_UNSUPPORTED, "tensor-core path: %s", why2); }
    f->tc1_ok = why1 == nullptr;
    f->s2_ok = why2 == nullptr;
    f->npass = npass;
    f->kernel = BCNF_KERNEL_TCGEN05;
    f->rows_per_cta = kTcRows;
    const char* gen = getenv("BCNF_FLOW_TC");
    f->s2_use = f->s2_ok && !(gen && atoi(gen) == 1 && f->tc1_ok);
    f->env_proj_fma = getenv("BCNF_PROJ_FMA") != nullptr;
    if (const char* e = getenv("BCNF_TC2_DEBUG")) f->env_tc2_debug = atoi(e);
    if (const char* e = getenv("BCNF_TC2_TRACE")) f->env_tc2_trace = e;
    if (const char* e = getenv("BCNF_TC_TRACE")) f->env_tc_trace = e;
    f->s2_ctas = 2 * (f->num_sms / 2);
    if (f->s2_use) f->rows_per_cta = kS2Rows;
  } else if (rowthread_ok) {
    f->kernel = BCNF_KERNEL_ROWTHREAD;
    if (const char* e = getenv("BCNF_ROWTHREAD_R")) f->env_rowthread_r = atoi(e) == 1 ? 1 : (atoi(e) == 2 ? 2 : 0);
    const int R = f->env_rowthread_r > 0 ? f->env_rowthread_r : (hp0 == 16 ? 2 : 1);
    f->rows_per_cta = R * kRowThreadBlock;
  } else {
    f->kernel = BCNF_KERNEL_TILED;
    int hpmax = 0;
    for (int l = 0; l < desc->n_hidden; ++l) hpmax = std::max(hpmax, sd.half[0].hp[l]);
    TiledSmem lay;
    lay.YP = round_up(sd.D, 4);
    lay.XP = std::max(sd.half[0].dinp, sd.half[1].dinp);
    lay.TP = 2 * std::max(sd.half[0].dop, sd.half[1].dop);
    lay.AP = hpmax + 4;
    f->tiled_lay = lay;
    f->tiled_R = lay.bytes(32) <= (size_t)f->max_smem_optin ? 32 : 16;
    if (lay.bytes(f->tiled_R) > (size_t)f->max_smem_optin) {
      delete f;
      return fail(BCNF_E_UNSUPPORTED, "row tile needs %zu bytes of shared memory", lay.bytes(16));
    }
    f->rows_per_cta = f->tiled_R;
  }

  // device storage
  long long bytes = 0;
  for (int d = 0; d < 2; ++d) {
    Program& p = f->prog[d];
    if (cudaMalloc(&p.d_blob, p.blob_floats * 4) != cudaSuccess ||
        cudaMalloc(&p.d_ops, p.ops.size() * sizeof(DevOp)) != cudaSuccess ||
        cudaMalloc(&p.d_chunks, p.chunks.size() * sizeof(Chunk)) != cudaSuccess) {
      cudaGetLastError();
      bcnf_flow_destroy(f);
      return fail(BCNF_E_NOMEM, "device allocation of %lld bytes failed", p.blob_floats * 4);
    }
    cudaMemcpy(p.d_ops, p.ops.data(), p.ops.size() * sizeof(DevOp), cudaMemcpyHostToDevice);
    cudaMemcpy(p.d_chunks, p.chunks.data(), p.chunks.size() * sizeof(Chunk), cudaMemcpyHostToDevice);
    bytes += p.blob_floats * 4;
  }
  const size_t wproj_bytes = (size_t)sd.C * sd.PW * 4;
  if (cudaMalloc(&f->d_wproj, wproj_bytes) != cudaSuccess || cudaMalloc(&f->d_bproj, (size_t)sd.PW * 4) != cudaSuccess) {
    cudaGetLastError();
    bcnf_flow_destroy(f);
    return fail(BCNF_E_NOMEM, "device allocation of %zu bytes failed", wproj_bytes);
  }
  bytes += wproj_bytes + (size_t)sd.PW * 4;
  // pack table: per direction, per op at most 2*(L+1)+2 descriptors; projection 2 per half
  f->pack_cap = 2 * (int)f->prog[0].ops.size() * (2 * desc->n_hidden + 6) + 4 * f->n_half + 16;
  if (cudaMalloc(&f->d_pack, f->pack_cap * sizeof(PackDesc)) != cudaSuccess ||
      cudaMallocHost(&f->h_pack, f->pack_cap * sizeof(PackDesc)) != cudaSuccess) {
    cudaGetLastError();
    bcnf_flow_destroy(f);
    return fail(BCNF_E_NOMEM, "pack table allocation failed");
  }
  if (f->npass && f->tc1_ok) {
    // one tile stream per conditioner network, in forward layer order; both directions index into it
    const int n = (int)f->op_types.size();
    f->tc_off_by_layer.assign(2 * n, -1);
    long long off = 0;
    int n_tiles = 0;
    for (int i = 0; i < n; ++i)
      if (f->op_types[i] == BCNF_OP_COUPLING)
        for (int s = 0; s < (desc->two_way ? 2 : 1); ++s) {
          f->tc_off_by_layer[2 * i + s] = off;
          off += f->td.half[s].stream_bytes;
          for (int l = 0; l <= f->td.half[s].L; ++l) n_tiles += f->td.half[s].layer[l].n_chunks * f->td.half[s].layer[l].kc;
        }
    f->tc_blob_bytes = off;
    f->tc_pack_cap = n_tiles;
    bool ok = cudaMalloc(&f->d_tc_blob, off) == cudaSuccess &&
              cudaMalloc(&f->d_tc_pack, (size_t)n_tiles * sizeof(TcPackDesc)) == cudaSuccess &&
              cudaMallocHost(&f->h_tc_pack, (size_t)n_tiles * sizeof(TcPackDesc)) == cudaSuccess;
    for (int d = 0; d < 2 && ok; ++d) {
      Program& p = f->prog[d];
      std::vector<long long> offs(p.ops.size(), 0);
      int oi = 0;
      for (int sidx = 0; sidx < n; ++sidx) {
        const int i = d == 0 ? sidx : n - 1 - sidx;
        if (f->op_types[i] == BCNF_OP_COUPLING) {
          offs[oi++] = f->tc_off_by_layer[2 * i];
          if (desc->two_way) offs[oi++] = f->tc_off_by_layer[2 * i + 1];
        } else {
          oi++;
        }
      }
      ok = cudaMalloc(&f->d_tc_off[d], offs.size() * sizeof(long long)) == cudaSuccess;
      if (ok) cudaMemcpy(f->d_tc_off[d], offs.data(), offs.size() * sizeof(long long), cudaMemcpyHostToDevice);
    }
    if (!ok) {
      cudaGetLastError();
      bcnf_flow_destroy(f);
      return fail(BCNF_E_NOMEM, "device allocation of %lld bytes (bf16 weight tiles) failed", off);
    }
    bytes += off;
  }
  f->packed_bytes = bytes;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { bcnf_flow_destroy(f); return fail((int)e, "create: %s", cudaGetErrorString(e)); }
  *out = f;
  return BCNF_OK;
}

// Watchdog record of the second-generation fused kernel: {wait code, aux, block, thread} of the first wait that timed
// out (all zero in a healthy run).  Host memory: still readable after the trap has killed the context.
extern "C" int bcnf_flow_debug_words(const bcnf_flow_t* f, uint32_t* out4) {
  if (!f || !out4) return fail(BCNF_E_ARG, "bcnf_flow_debug_words: null argument");
  for (int i = 0; i < 4; ++i) out4[i] = f->h_s2_dbg ? f->h_s2_dbg[i] : 0u;
  return BCNF_OK;
}

extern "C" int bcnf_flow_info(const bcnf_flow_t* f, bcnf_flow_info_t* info) {
  if (!f || !info) return fail(BCNF_E_ARG, "bcnf_flow_info: null argument");
  info->kernel = f->kernel;
  info->proj_width = f->sd.PW;
  info->n_half_couplings = f->n_half;
  info->rows_per_cta = f->rows_per_cta;
  info->packed_bytes = f->packed_bytes;
  info->macs_per_row = f->macs_row;
  info->macs_per_instance = f->macs_inst;
  return BCNF_OK;
}

// Emit the pack descriptors of one conditioner network into `v`.
static int emit_half(const bcnf_flow& f, const HalfLayout& hl, float* dst, const float* const* w,
                     const float* const* b, int proj_off, bool with_proj, std::vector<PackDesc>& v) {
  const int L = hl.L, C = f.sd.C;
  for (int j = 0; j <= L; ++j)
    if (!w || !b || !w[j] || !b[j]) return fail(BCNF_E_ARG, "coupling layer is missing Linear %d parameters", j);
  const int in_total = hl.din + C;   // sizes[0] += n_conditions, cnf.py:72
  // first Linear, own-half columns -> W1a (k-major)
  v.push_back({w[0], dst + hl.off_w[0], in_total, hl.h[0], hl.din, hl.hp[0], 1, 0});
  if (with_proj) {
    // first Linear, feature columns -> projection matrix; bias -> projection bias
    v.push_back({w[0] + hl.din, f.d_wproj + proj_off, in_total, hl.h[0], C, f.sd.PW, 1, 0});
    v.push_back({b[0], f.d_bproj + proj_off, hl.h[0], 1, hl.h[0], hl.hp[0], 0, 0});
  }
  for (int l = 1; l < L; ++l) {
    v.push_back({w[l], dst + hl.off_w[l], hl.h[l - 1], hl.h[l], hl.h[l - 1], hl.hp[l], 1, 0});
    v.push_back({b[l], dst + hl.off_b[l], hl.h[l], 1, hl.h[l], hl.hp[l], 0, 0});
  }
  // last Linear: rows [0, dout) are t, rows [dout, 2 dout) are s (chunk(2, dim=1), cnf.py:104)
  const int hin = hl.h[L - 1];
  v.push_back({w[L], dst + hl.off_wout, hin, hl.dout, hin, 2 * hl.dop, 1, 0});
  v.push_back({w[L] + (size_t)hl.dout * hin, dst + hl.off_wout + hl.dop, hin, hl.dout, hin, 2 * hl.dop, 1, 0});
  v.push_back({b[L], dst + hl.off_bout, hl.dout, 1, hl.dout, 2 * hl.dop, 0, 0});
  v.push_back({b[L] + hl.dout, dst + hl.off_bout + hl.dop, hl.dout, 1, hl.dout, 2 * hl.dop, 0, 0});
  return 0;
}

// Emit the bf16 tile-image descriptors of one conditioner network.
static void emit_tc_half(const bcnf_flow& f, int s, const float* const* w, long long off, std::vector<TcPackDesc>& v) {
  const HalfLayout& hl = f.sd.half[s];
  const TcHalfLayout& tl = f.td.half[s];
  unsigned char* dst = f.d_tc_blob + off;
  for (int l = 0; l <= hl.L; ++l) {
    const TcLayer& ly = tl.layer[l];
    TcPackDesc d{};
    d.w = w[l];
    d.pitch = l == 0 ? hl.din + f.sd.C : hl.h[l - 1];
    d.col0 = 0;
    d.k_valid = l == 0 ? hl.din : hl.h[l - 1];
    d.out_mode = l == hl.L;
    d.n_valid = l == hl.L ? hl.dout : hl.h[l];
    d.doh = tl.doh;
    for (int kc = 0; kc < ly.kc; ++kc) {          // K-major stream order, as the kernel consumes it
      int coff = 0;
      for (int nc = 0; nc < ly.n_chunks; ++nc) {
        d.dst = dst; d.chunk_off = coff; d.chunk_n = ly.chunk_n[nc]; d.kc = kc; d.kw = f.td.kw;
        v.push_back(d);
        dst += (size_t)4 * (ly.chunk_n[nc] / 2) * (2 * f.td.kw);
        coff += ly.chunk_n[nc];
      }
    }
  }
}

// Weight images of every Linear of every conditioner network for flow_tc2.cuh, made from the forward program's fp32
// blob (input-major matrices: W1a [DINP][HP0], W_l [HP(l-1)][HP(l)], last Linear [HP(L-1)][2*DOP] with t | s halves),
// and the per-direction table: device op index -> byte offset of its network's images.
static int build_s2_images(bcnf_flow* f, cudaStream_t stream) {
  if (!f->s2_ok) return BCNF_OK;
  const StackDims& sd = f->sd;
  const Program& p = f->prog[0];
  std::vector<long long> net_off;          // per conditioner network, forward program order
  long long bytes = 0;
  for (const auto& op : p.ops)
    if (op.type == DOP_HALF) { net_off.push_back(bytes); bytes += f->s2.half[op.src].net_bytes; }
  if (bytes > f->s2_img_bytes) {
    if (f->d_s2_img) CUDA_TRY(cudaFree(f->d_s2_img));
    f->d_s2_img = nullptr; f->s2_img_bytes = 0;
    CUDA_TRY(cudaMalloc(&f->d_s2_img, (size_t)bytes));
    f->s2_img_bytes = bytes;
  }
  std::vector<ImgPackDesc> descs;
  {
    size_t h = 0;
    for (const auto& op : p.ops) {
      if (op.type != DOP_HALF) continue;
      const HalfLayout& hl = sd.half[op.src];
      const S2Half& tl = f->s2.half[op.src];
      for (int l = 0; l <= hl.L; ++l) {
        const S2Layer& ly = tl.layer[l];
        ImgPackDesc d;
        const int n_out = l < hl.L ? hl.hp[l] : 2 * hl.dop;
        d.src = p.d_blob + op.off + (l == 0 ? hl.off_w[0] : (l < hl.L ? hl.off_w[l] : hl.off_wout));
        d.s_row = 1; d.s_k = n_out; d.rows = n_out; d.k = l == 0 ? hl.dinp : hl.hp[l - 1];
        d.dst = f->d_s2_img + net_off[h] + ly.w_off; d.plane = ly.w_plane; d.rpad = ly.w_rpad; d.chunks = ly.n_kst;
        descs.push_back(d);
      }
      ++h;
    }
  }
  for (size_t b0 = 0; b0 < descs.size(); b0 += kImgPackMax) {
    ImgPackBatch batch;
    const int nb = (int)std::min<size_t>(kImgPackMax, descs.size() - b0);
    int max_blocks = 0;
    for (int i = 0; i < nb; ++i) { batch.d[i] = descs[b0 + i]; max_blocks = std::max(max_blocks, descs[b0 + i].rpad / 32); }
    img_pack_kernel<<<dim3(max_blocks, nb), kTgGroupThreads, 0, stream>>>(batch);
    CUDA_TRY(cudaGetLastError());
  }
  // folded biases: one weight-image column per (network, layer)
  {
    std::vector<S2BiasCol> cols;
    size_t h = 0;
    for (const auto& op : p.ops) {
      if (op.type != DOP_HALF) continue;
      const HalfLayout& hl = sd.half[op.src];
      const S2Half& tl = f->s2.half[op.src];
      for (int l = 1; l <= hl.L; ++l) {
        const S2Layer& ly = tl.layer[l];
        if (ly.bias_k < 0) continue;
        S2BiasCol c;
        c.bias = p.d_blob + op.off + (l < hl.L ? hl.off_b[l] : hl.off_bout);
        c.img = f->d_s2_img + net_off[h] + ly.w_off; c.plane = ly.w_plane; c.rpad = ly.w_rpad;
        c.n = l < hl.L ? hl.hp[l] : 2 * hl.dop; c.k = ly.bias_k; c.pad = 0;
        cols.push_back(c);
      }
      ++h;
    }
    if (!cols.empty()) {
      if (!f->d_s2_bias) CUDA_TRY(cudaMalloc(&f->d_s2_bias, cols.size() * sizeof(S2BiasCol)));
      CUDA_TRY(cudaMemcpyAsync(f->d_s2_bias, cols.data(), cols.size() * sizeof(S2BiasCol), cudaMemcpyHostToDevice, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));       // (cols is a host temporary)
      s2_bias_col_kernel<<<(unsigned)cols.size(), 256, 0, stream>>>(f->d_s2_bias);
      CUDA_TRY(cudaGetLastError());
    }
  }
  // offset tables (static per handle: built once)
  if (!f->d_s2_off[0]) {
    const int n = (int)f->op_types.size();
    // forward network index of each (layer, sub-network)
    std::vector<long long> by_layer(2 * n, -1);
    {
      size_t h = 0;
      for (int i = 0; i < n; ++i)
        if (f->op_types[i] == BCNF_OP_COUPLING)
          for (int s2 = 0; s2 < (f->desc.two_way ? 2 : 1); ++s2) by_layer[2 * i + s2] = net_off[h++];
    }
    for (int dir = 0; dir < 2; ++dir) {
      std::vector<long long> offs(f->prog[dir].ops.size(), 0);
      int oi = 0;
      for (int sidx = 0; sidx < n; ++sidx) {
        const int i = dir == 0 ? sidx : n - 1 - sidx;
        if (f->op_types[i] == BCNF_OP_COUPLING) {
          offs[oi++] = by_layer[2 * i];
          if (f->desc.two_way) offs[oi++] = by_layer[2 * i + 1];
        } else {
          oi++;
        }
      }
      CUDA_TRY(cudaMalloc(&f->d_s2_off[dir], offs.size() * sizeof(long long)));
      CUDA_TRY(cudaMemcpyAsync(f->d_s2_off[dir], offs.data(), offs.size() * sizeof(long long), cudaMemcpyHostToDevice, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));      // (offs is a host temporary; create-time only)
    }
  }
  if (!f->h_s2_dbg) {
    CUDA_TRY(cudaHostAlloc(&f->h_s2_dbg, 64, cudaHostAllocMapped));
    memset(f->h_s2_dbg, 0, 64);
    CUDA_TRY(cudaHostGetDevicePointer(&f->d_s2_dbg, f->h_s2_dbg, 0));
  }
  return BCNF_OK;
}

// Activation scratch of the stream `stream` (one per stream that runs the kernel: two launches of one handle on
// different streams may overlap).  The first stream's scratch is allocated by set_params; another stream's on its
// first call, which is therefore not capturable.
static int s2_scratch_for(bcnf_flow* f, cudaStream_t stream, unsigned char** out) {
  for (auto& sc : f->s2_scratch)
    if (sc.stream == stream) { *out = sc.act; return BCNF_OK; }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  CUDA_TRY(cudaStreamIsCapturing(stream, &cap));
  if (cap != cudaStreamCaptureStatusNone)
    return fail(BCNF_E_STATE, "the activation scratch of this stream is not allocated yet; run the flow once on it "
                              "outside CUDA-graph capture");
  unsigned char* act = nullptr;
  CUDA_TRY(cudaMalloc(&act, (size_t)f->s2_ctas * (size_t)f->s2.cta_bytes));
  f->s2_scratch.push_back({stream, act});
  *out = act;
  return BCNF_OK;
}

extern "C" int bcnf_flow_set_params(bcnf_flow_t* f, const bcnf_op_params_t* ops, void* stream_) {
  NVTX_RANGE("bcnf_flow_set_params");
  if (!f || !ops) return fail(BCNF_E_ARG, "bcnf_flow_set_params: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  DEVICE_GUARD(f->desc.device);
  const int n = (int)f->op_types.size();
  std::vector<PackDesc> v;
  v.reserve(f->pack_cap);
  const StackDims& sd = f->sd;
  for (int dir = 0; dir < 2; ++dir) {
    Program& p = f->prog[dir];
    CUDA_TRY(cudaMemsetAsync(p.d_blob, 0, p.blob_floats * 4, stream));
    int oi = 0;
    for (int s = 0; s < n; ++s) {
      const int i = dir == 0 ? s : n - 1 - s;
      const bcnf_op_params_t& src = ops[i];
      if (src.type != f->op_types[i])
        return fail(BCNF_E_ARG, "layer %d: type %d does not match the handle's %d", i, src.type, f->op_types[i]);
      float* dst = p.d_blob + p.ops[oi].off;
      if (src.type == BCNF_OP_ACTNORM) {
        if (!src.scale || !src.bias) return fail(BCNF_E_ARG, "layer %d: ActNorm needs scale and bias", i);
        v.push_back({src.scale, dst, sd.D, 1, sd.D, sd.DP, 0, 0});
        v.push_back({src.bias, dst + sd.DP, sd.D, 1, sd.D, sd.DP, 0, 0});
        v.push_back({src.scale, dst + 2 * sd.DP, 1, sd.D, 1, 1, 2, 0});
        oi += 1;
      } else if (src.type == BCNF_OP_ORTHO) {
        if (!src.q) return fail(BCNF_E_ARG, "layer %d: orthonormal_matrix missing", i);
        v.push_back({src.q, dst, sd.D, sd.D, sd.D, sd.DP, dir == 0 ? 0 : 1, 0});   // Q or Q^T
        oi += 1;
      } else {
        int rc = emit_half(*f, sd.half[0], dst, src.w_a, src.b_a, p.ops[oi].proj_off, dir == 0, v);
        if (rc) return rc;
        oi += 1;
        if (f->desc.two_way) {
          float* dst_b = p.d_blob + p.ops[oi].off;
          rc = emit_half(*f, sd.half[1], dst_b, src.w_b, src.b_b, p.ops[oi].proj_off, dir == 0, v);
          if (rc) return rc;
          oi += 1;
        }
      }
    }
  }
  if ((int)v.size() > f->pack_cap) return fail(BCNF_E_STATE, "internal: pack table overflow");
  CUDA_TRY(cudaMemsetAsync(f->d_wproj, 0, (size_t)sd.C * sd.PW * 4, stream));
  CUDA_TRY(cudaMemsetAsync(f->d_bproj, 0, (size_t)sd.PW * 4, stream));
  // the pinned staging table is reused: wait for the previous upload on this handle
  CUDA_TRY(cudaStreamSynchronize(stream));
  memcpy(f->h_pack, v.data(), v.size() * sizeof(PackDesc));
  CUDA_TRY(cudaMemcpyAsync(f->d_pack, f->h_pack, v.size() * sizeof(PackDesc), cudaMemcpyHostToDevice, stream));
  pack_kernel<<<dim3((unsigned)v.size(), 8), 256, 0, stream>>>(f->d_pack);
  CUDA_TRY(cudaGetLastError());
  if (f->npass) {
    std::vector<TcPackDesc> tv;
    tv.reserve(f->tc_pack_cap);
    for (int i = 0; i < n && f->tc1_ok; ++i)
      if (f->op_types[i] == BCNF_OP_COUPLING) {
        emit_tc_half(*f, 0, ops[i].w_a, f->tc_off_by_layer[2 * i], tv);
        if (f->desc.two_way) emit_tc_half(*f, 1, ops[i].w_b, f->tc_off_by_layer[2 * i + 1], tv);
      }
    if ((int)tv.size() > f->tc_pack_cap) return fail(BCNF_E_STATE, "internal: tile pack table overflow");
    if (!tv.empty()) {
      CUDA_TRY(cudaStreamSynchronize(stream));
      memcpy(f->h_tc_pack, tv.data(), tv.size() * sizeof(TcPackDesc));
      CUDA_TRY(cudaMemcpyAsync(f->d_tc_pack, f->h_tc_pack, tv.size() * sizeof(TcPackDesc), cudaMemcpyHostToDevice, stream));
      tc_pack_kernel<<<(unsigned)tv.size(), 256, 0, stream>>>(f->d_tc_pack);
      CUDA_TRY(cudaGetLastError());
    }
    // image of Wproj for the CTA-pair projection GEMM: rows = projection column j, k = condition feature
    const int chunks = (sd.C + 63) / 64;
    f->wproj_rpad = (sd.PW + 255) / 256 * 256;
    f->wproj_plane = (long long)chunks * f->wproj_rpad * 128;
    if (!f->d_wproj_img) CUDA_TRY(cudaMalloc(&f->d_wproj_img, (size_t)(2 * f->wproj_plane)));
    if (!f->d_h_img) {     // scratch image of one slice of instances' features (bcnf_cond_project never allocates)
      f->h_img_bytes = 2LL * chunks * 32768 * 128;
      CUDA_TRY(cudaMalloc(&f->d_h_img, (size_t)f->h_img_bytes));
    }
    ImgPackBatch batch;
    ImgPackDesc& d = batch.d[0];
    d.src = f->d_wproj; d.s_row = 1; d.s_k = sd.PW; d.rows = sd.PW; d.k = sd.C;
    d.dst = f->d_wproj_img; d.plane = f->wproj_plane; d.rpad = f->wproj_rpad; d.chunks = chunks;
    img_pack_kernel<<<dim3(f->wproj_rpad / 32, 1), kTgGroupThreads, 0, stream>>>(batch);
    CUDA_TRY(cudaGetLastError());
    if (int rc = build_s2_images(f, stream)) return rc;
    if (f->s2_ok && f->s2_scratch.empty()) {
      unsigned char* act = nullptr;
      if (int rc = s2_scratch_for(f, stream, &act)) return rc;
    }
  }
  f->params_set = true;
  return BCNF_OK;
}

template <int NPASS>
static int launch_gemm_img2(const G2Args& g, int num_sms, cudaStream_t stream) {
  using Cfg = G2Cfg<NPASS>;
  auto kern = gemm_img2_kernel<NPASS>;
  static bool attr_set[64] = {};
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_set[dev]) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem));
    if (dev < 64) attr_set[dev] = true;
  }
  const long long tiles = (long long)((g.M + 255) / 256) * ((g.N + 255) / 256);
  const int n_kc = (g.K + 63) / 64;
  if (g.a_rpad % 128 || g.a_rpad < (g.M + 255) / 256 * 256 || g.a_plane < (long long)n_kc * g.a_rpad * 128)
    return fail(BCNF_E_ARG, "gemm_img2: A image too small (rpad=%d for M=%d: needs a multiple of 256 rows)", g.a_rpad, g.M);
  if (g.b_rpad % 128 || g.b_rpad < (g.N + 255) / 256 * 256 || g.b_plane < (long long)n_kc * g.b_rpad * 128)
    return fail(BCNF_E_ARG, "gemm_img2: B image too small (rpad=%d for N=%d: needs a multiple of 256 rows)", g.b_rpad, g.N);
  const int pairs = (int)std::min<long long>(tiles, num_sms / 2);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(kG2Threads);
  cfg.dynamicSmemBytes = Cfg::smem; cfg.stream = stream;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, g));
  return BCNF_OK;
}

// gelu(A . B^T + bias) written as an operand image (the A operand of the next Linear): feature MLPs on the tensor cores
extern "C" int bcnf_gemm_img_gelu(const void* a_img, int64_t a_plane, int32_t a_rpad, const void* b_img, int64_t b_plane,
                                  int32_t b_rpad, const float* bias, void* c_img, int64_t c_plane, int32_t c_rpad, int32_t M,
                                  int32_t N, int32_t K, int32_t passes, int32_t device, void* stream) {
  if (!a_img || !b_img || !c_img || !bias || M < 0 || N < 0 || K < 1) return fail(BCNF_E_ARG, "bcnf_gemm_img_gelu: bad argument");
  if (passes != 1 && passes != 3) return fail(BCNF_E_ARG, "bcnf_gemm_img_gelu: passes must be 1 (bf16) or 3 (bf16x3)");
  if (c_rpad % 256 || c_rpad < (M + 255) / 256 * 256 || c_plane % ((long long)c_rpad * 128) || c_plane / ((long long)c_rpad * 128) * 64 < N)
    return fail(BCNF_E_ARG, "bcnf_gemm_img_gelu: output image too small (rpad=%d plane=%lld for M=%d N=%d)", c_rpad, (long long)c_plane, M, N);
  if (((uintptr_t)bias & 15) != 0) return fail(BCNF_E_ARG, "bcnf_gemm_img_gelu: bias must be 16-byte aligned");
  if (M == 0 || N == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  int n_sm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
  G2Args g;
  memset(&g, 0, sizeof(g));
  g.a_img = (const unsigned char*)a_img; g.a_plane = a_plane; g.a_rpad = a_rpad;
  g.b_img = (const unsigned char*)b_img; g.b_plane = b_plane; g.b_rpad = b_rpad;
  g.bias = bias; g.M = M; g.N = N; g.K = K;
  g.c_img = (unsigned char*)c_img; g.c_plane = c_plane; g.c_rpad = c_rpad;
  return passes == 3 ? launch_gemm_img2<3>(g, n_sm, (cudaStream_t)stream) : launch_gemm_img2<1>(g, n_sm, (cudaStream_t)stream);
}

// One time step of one LSTM layer / direction: gates = [x_t | h_(t-1)] . Wcat^T + b on the CTA-pair GEMM with the cell
// update in the epilogue (gemm_img2.cuh, mode 2)
static_assert(sizeof(bcnf_lstm_step_t) == 16 * 8 * 2 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 4 * 8 * 2 + 16, "bcnf_lstm_step_t layout");
extern "C" int bcnf_lstm_step(const bcnf_lstm_step_t* a, int32_t device, void* stream) {
  if (!a || !a->b_img || !a->bias || !a->cell) return fail(BCNF_E_ARG, "bcnf_lstm_step: null argument");
  if (a->n_chunks < 1 || a->n_chunks > 16) return fail(BCNF_E_ARG, "bcnf_lstm_step: %d operand chunks (1..16)", a->n_chunks);
  if (a->passes != 1 && a->passes != 3) return fail(BCNF_E_ARG, "bcnf_lstm_step: passes must be 1 or 3");
  if (a->N < 8 || a->N % 8 || a->N > 1024) return fail(BCNF_E_ARG, "bcnf_lstm_step: N = 4 * hidden must be a multiple of 8, <= 1024");
  if (a->M < 0 || a->a_rpad % 256 || a->a_rpad < (a->M + 255) / 256 * 256 || a->state_rows < a->M)
    return fail(BCNF_E_ARG, "bcnf_lstm_step: bad row counts (M=%d a_rpad=%d state_rows=%lld)", a->M, a->a_rpad, (long long)a->state_rows);
  if (((uintptr_t)a->bias & 15) != 0) return fail(BCNF_E_ARG, "bcnf_lstm_step: bias must be 16-byte aligned");
  const int tiles_n = (a->N + 255) / 256;
  for (int k = 0; k < a->n_chunks; ++k)
    if (!a->a_hi[k] || (a->passes == 3 && !a->a_lo[k])) return fail(BCNF_E_ARG, "bcnf_lstm_step: operand chunk %d missing", k);
  for (int t = 0; t < tiles_n; ++t)
    if (!a->h_hi[t] || (a->passes == 3 && !a->h_lo[t])) return fail(BCNF_E_ARG, "bcnf_lstm_step: output chunk %d missing", t);
  if (a->M == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  int n_sm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
  G2Args g;
  memset(&g, 0, sizeof(g));
  g.a_tab_n = a->n_chunks;
  for (int k = 0; k < a->n_chunks; ++k) { g.a_tab_hi[k] = (const unsigned char*)a->a_hi[k]; g.a_tab_lo[k] = (const unsigned char*)a->a_lo[k]; }
  g.a_img = g.a_tab_hi[0]; g.a_rpad = a->a_rpad; g.a_plane = (long long)a->n_chunks * a->a_rpad * 128;   // (size checks only)
  g.b_img = (const unsigned char*)a->b_img; g.b_plane = a->b_plane; g.b_rpad = a->b_rpad;
  g.bias = a->bias; g.M = a->M; g.N = a->N; g.K = a->n_chunks * 64;
  g.lstm = 1; g.cell = a->cell; g.hsum = a->hsum; g.state_rows = a->state_rows;
  for (int t = 0; t < 4; ++t) { g.h_hi[t] = (unsigned char*)a->h_hi[t]; g.h_lo[t] = (unsigned char*)a->h_lo[t]; }
  return a->passes == 3 ? launch_gemm_img2<3>(g, n_sm, (cudaStream_t)stream) : launch_gemm_img2<1>(g, n_sm, (cudaStream_t)stream);
}

static void* g_g2_trace = nullptr;
// debug: device buffer (74 x 16 x 4 uint64) that receives globaltimer stamps of the next bcnf_gemm_img launches
extern "C" int bcnf_gemm_img_set_trace(void* device_buffer) { g_g2_trace = device_buffer; return BCNF_OK; }

// C = A . B^T (+ bias) on operand images with the CTA-pair kernel (tests, tools)
extern "C" int bcnf_gemm_img(const void* a_img, int64_t a_plane, int32_t a_rpad, const void* b_img, int64_t b_plane,
                             int32_t b_rpad, float* C, int64_t ldc, const float* bias, int32_t M, int32_t N, int32_t K,
                             int32_t passes, int32_t device, void* stream) {
  if (!a_img || !b_img || !C || M < 0 || N < 0 || K < 1) return fail(BCNF_E_ARG, "bcnf_gemm_img: bad argument");
  if (passes != 1 && passes != 3) return fail(BCNF_E_ARG, "bcnf_gemm_img: passes must be 1 (bf16) or 3 (bf16x3)");
  if (M == 0 || N == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  int n_sm = 0;     // (cudaGetDeviceProperties takes milliseconds per call)
  CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
  G2Args g;
  memset(&g, 0, sizeof(g));
  g.a_img = (const unsigned char*)a_img; g.a_plane = a_plane; g.a_rpad = a_rpad;
  g.b_img = (const unsigned char*)b_img; g.b_plane = b_plane; g.b_rpad = b_rpad;
  g.C = C; g.ldc = ldc; g.bias = bias; g.M = M; g.N = N; g.K = K;
  static const int g2_debug = getenv("BCNF_G2_DEBUG") ? atoi(getenv("BCNF_G2_DEBUG")) : 0;   // read once per process
  g.debug = g2_debug;
  g.trace = (unsigned long long*)g_g2_trace;
  return passes == 3 ? launch_gemm_img2<3>(g, n_sm, (cudaStream_t)stream) : launch_gemm_img2<1>(g, n_sm, (cudaStream_t)stream);
}

static const long long kProjSlice = 32768;   // instances per launch of the projection GEMM (scratch image: one slice)

extern "C" int bcnf_cond_project(bcnf_flow_t* f, const float* h, int64_t n_inst, float* P, void* stream_) {
  NVTX_RANGE("bcnf_cond_project");
  if (!f || !h || !P) return fail(BCNF_E_ARG, "bcnf_cond_project: null argument");
  if (n_inst < 0) return fail(BCNF_E_ARG, "n_inst=%lld", (long long)n_inst);
  if (!f->params_set) return fail(BCNF_E_STATE, "bcnf_flow_set_params has not been called");
  if (n_inst == 0) return BCNF_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  DEVICE_GUARD(f->desc.device);
  if (f->npass && !f->env_proj_fma) {
    // CTA-pair GEMM on operand images, a slice of instances at a time: h -> image (scratch), P = h_img . Wproj_img^T + b
    const int chunks = (f->sd.C + 63) / 64;
    const long long slice = kProjSlice;
    const long long cap_rows = std::min<long long>((n_inst + 255) / 256 * 256, slice);
    const long long need = 2 * (long long)chunks * cap_rows * 128;
    // the scratch image holds one slice of instances and was sized for a full slice by bcnf_flow_set_params: no
    // allocation and no synchronisation here, whatever n_inst is (graph capture at any instance count)
    if (need > f->h_img_bytes) return fail(BCNF_E_STATE, "internal: projection scratch image smaller than one slice");
    for (long long m0 = 0; m0 < n_inst; m0 += slice) {
      const long long m = std::min<long long>(slice, n_inst - m0);
      const int rpad = (int)((m + 255) / 256 * 256);
      const long long plane = (long long)chunks * rpad * 128;
      ImgPackBatch batch;
      ImgPackDesc& d = batch.d[0];
      d.src = h + m0 * f->sd.C; d.s_row = f->sd.C; d.s_k = 1; d.rows = (int)m; d.k = f->sd.C;
      d.dst = f->d_h_img; d.plane = plane; d.rpad = rpad; d.chunks = chunks;
      img_pack_kernel<<<dim3(rpad / 32, 1), kTgGroupThreads, 0, stream>>>(batch);
      CUDA_TRY(cudaGetLastError());
      G2Args g;
      memset(&g, 0, sizeof(g));
      g.a_img = f->d_h_img; g.a_plane = plane; g.a_rpad = rpad;
      g.b_img = f->d_wproj_img; g.b_plane = f->wproj_plane; g.b_rpad = f->wproj_rpad;
      g.C = P + m0 * f->sd.PW; g.ldc = f->sd.PW; g.bias = f->d_bproj; g.M = (int)m; g.N = f->sd.PW; g.K = f->sd.C; g.debug = 0; g.trace = nullptr;
      if (int rc = f->npass == 3 ? launch_gemm_img2<3>(g, f->num_sms, stream) : launch_gemm_img2<1>(g, f->num_sms, stream)) return rc;
    }
    return BCNF_OK;
  }
  const int N = f->sd.PW, K = f->sd.C;
  const long long max_rows = 65535LL * kProjBM;
  for (long long m0 = 0; m0 < n_inst; m0 += max_rows) {
    const long long m = std::min<long long>(max_rows, n_inst - m0);
    dim3 grid((N + kProjBN - 1) / kProjBN, (unsigned)((m + kProjBM - 1) / kProjBM));
    cond_project_kernel<<<grid, 256, 0, stream>>>(h + m0 * K, f->d_wproj, f->d_bproj, P + m0 * N, m, N, K);
  }
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

template <int D, int HP, int R>
static int launch_rowthread(bcnf_flow* f, const FlowArgs& a, int dir, cudaStream_t stream) {
  const Program& p = f->prog[dir];
  const int cap = round_up(std::max(f->prog[0].max_chunk_bytes, f->prog[1].max_chunk_bytes), 128);
  const size_t smem = 2 * (size_t)cap + 16;
  auto kern = flow_rowthread_kernel<D, HP, R>;
  static size_t configured[64] = {};
  if (int rc = opt_in_smem_once(kern, smem, configured)) return rc;
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kRowThreadBlock, smem));
  if (occ < 1) return fail(BCNF_E_UNSUPPORTED, "row-per-thread kernel does not fit on an SM");
  const long long tiles = (a.n_rows + R * kRowThreadBlock - 1) / (R * kRowThreadBlock);
  const int grid = (int)std::min<long long>(tiles, (long long)f->num_sms * occ);
  (void)p;
  kern<<<grid, kRowThreadBlock, smem, stream>>>(a, f->sd, cap);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

template <int R>
static int launch_tiled(bcnf_flow* f, const FlowArgs& a, cudaStream_t stream) {
  const size_t smem = f->tiled_lay.bytes(R);
  auto kern = flow_tiled_kernel<R>;
  static size_t configured[64] = {};
  if (int rc = opt_in_smem_once(kern, smem, configured)) return rc;
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kTiledThreads, smem));
  if (occ < 1) return fail(BCNF_E_UNSUPPORTED, "tiled kernel does not fit on an SM (%zu bytes)", smem);
  const long long tiles = (a.n_rows + R - 1) / R;
  const int grid = (int)std::min<long long>(tiles, (long long)f->num_sms * occ);
  kern<<<grid, kTiledThreads, smem, stream>>>(a, f->sd, f->tiled_lay);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

template <int NPASS>
static int launch_tc(bcnf_flow* f, const FlowArgs& a, int dir, cudaStream_t stream) {
  const size_t smem = (size_t)f->td.smem_bytes;
  auto kern = flow_tc_kernel<NPASS>;
  static size_t configured[64] = {};
  if (int rc = opt_in_smem_once(kern, smem, configured)) return rc;
  const long long tiles = (a.n_rows + 2 * kTcRows - 1) / (2 * kTcRows);
  const int csize = 2;                  // one CTA pair per 128-row tile
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cfg.gridDim = dim3((unsigned)csize);
  int max_clusters = 0;
  CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
  if (max_clusters < 1) return fail(BCNF_E_UNSUPPORTED, "no cluster of %d CTAs fits on this device", csize);
  const long long want = (tiles + csize / 2 - 1) / (csize / 2);
  const int clusters = (int)std::min<long long>(want, max_clusters);
  cfg.gridDim = dim3((unsigned)(csize * clusters));
  const StackDims sd = f->sd;
  const TcDims td = f->td;
  const unsigned char* blob = f->d_tc_blob;
  const long long* offs = f->d_tc_off[dir];
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a, sd, td, blob, offs));
  return BCNF_OK;
}


template <int NPASS>
static int launch_tc2(bcnf_flow* f, const FlowArgs& a, int dir, cudaStream_t stream) {
  const size_t smem = (size_t)f->s2.smem_bytes;
  auto kern = flow_tc2_kernel<NPASS>;
  static size_t configured[64] = {};
  if (int rc = opt_in_smem_once(kern, smem, configured)) return rc;
  unsigned char* act = nullptr;
  if (int rc = s2_scratch_for(f, stream, &act)) return rc;
  const long long tiles = (a.n_rows + 2 * kS2Rows - 1) / (2 * kS2Rows);
  const int pairs = (int)std::min<long long>(tiles, f->s2_ctas / 2);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(kS2Threads);
  cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  const StackDims sd = f->sd;
  S2Dims d2 = f->s2;
  d2.debug = f->env_tc2_debug;          // timing experiments (wrong results); 0 outside them
  const unsigned char* img = f->d_s2_img;
  const long long* offs = f->d_s2_off[dir];
  unsigned int* dbg = f->d_s2_dbg;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a, sd, d2, img, offs, act, dbg));
  return BCNF_OK;
}

// what bcnf_flow_sample / bcnf_flow_sample_ranks add to a plain forward / inverse call
struct FlowExtra {
  bool draw = false;                 // in == null: z is drawn inside the kernel
  unsigned long long seed = 0;
  float sigma = 1.f;
  const float* rank_y = nullptr;     // rank reduction instead of the output
  int* rank_out = nullptr;
};

static int run_flow(bcnf_flow_t* f, int dir, const float* in, const float* P, const int32_t* row2inst,
                    int64_t inst_period, int64_t n_rows, float* out, float* logdet, void* stream_,
                    const FlowExtra& ex = FlowExtra()) {
  if (!f) return fail(BCNF_E_ARG, "bcnf_flow_%s: null handle", dir ? "inverse" : "forward");
  if (n_rows < 0 || inst_period < 0) return fail(BCNF_E_ARG, "negative size");
  if (!f->params_set) return fail(BCNF_E_STATE, "bcnf_flow_set_params has not been called");
  if (n_rows == 0) return BCNF_OK;   // empty batch: nothing to read or write
  if ((!in && !ex.draw) || !P || (!out && !ex.rank_out))
    return fail(BCNF_E_ARG, "bcnf_flow_%s: null argument", dir ? "inverse" : "forward");
  cudaStream_t stream = (cudaStream_t)stream_;
  DEVICE_GUARD(f->desc.device);
  const Program& p = f->prog[dir];
  FlowArgs a;
  a.in = in; a.out = out; a.logdet = logdet; a.P = P; a.row2inst = row2inst;
  a.inst_period = inst_period; a.n_rows = n_rows;
  a.blob = p.d_blob; a.ops = p.d_ops; a.n_ops = (int)p.ops.size();
  a.chunks = p.d_chunks; a.n_chunks = (int)p.chunks.size();
  a.trace = nullptr;
  a.blob_floats = p.blob_floats;
  a.seed = ex.seed; a.sigma = ex.sigma; a.rank_y = ex.rank_y; a.rank_out = ex.rank_out;
  if (const char* tp = f->kernel == BCNF_KERNEL_TCGEN05 && f->s2_use && !f->env_tc2_trace.empty() ? f->env_tc2_trace.c_str() : nullptr) {
    // debug aid: dump the clock64 stamps of block 0's issuer / epilogue / producer to the named file (synchronises!)
    const size_t n = 4 * 8192;
    long long* d_tr = nullptr;
    CUDA_TRY(cudaMalloc(&d_tr, n * sizeof(long long)));
    CUDA_TRY(cudaMemsetAsync(d_tr, 0, n * sizeof(long long), stream));
    a.trace = d_tr;
    int rc = f->npass == 3 ? launch_tc2<3>(f, a, dir, stream) : launch_tc2<1>(f, a, dir, stream);
    std::vector<long long> h(n);
    CUDA_TRY(cudaStreamSynchronize(stream));
    CUDA_TRY(cudaMemcpy(h.data(), d_tr, n * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(d_tr);
    if (FILE* fp = fopen(tp, "w")) {
      for (int role = 0; role < 4; ++role)
        for (int i = 0; i < 2048; ++i) {
          const long long* e = &h[role * 8192 + i * 4];
          if (e[0] | e[1] | e[2] | e[3]) fprintf(fp, "%d %d %lld %lld %lld %lld\n", role, i, e[0], e[1], e[2], e[3]);
        }
      fclose(fp);
    }
    return rc;
  }
  if (const char* tp = f->env_tc_trace.empty() ? nullptr : f->env_tc_trace.c_str()) {
    // debug aid: dump clock64 stamps of the first tile's pipeline to the named file (synchronises!)
    if (f->kernel == BCNF_KERNEL_TCGEN05 && !f->s2_use) {
      long long* d_tr = nullptr;
      CUDA_TRY(cudaMalloc(&d_tr, 64 * 8 * sizeof(long long)));
      CUDA_TRY(cudaMemsetAsync(d_tr, 0, 64 * 8 * sizeof(long long), stream));
      a.trace = d_tr;
      int rc = f->npass == 3 ? launch_tc<3>(f, a, dir, stream) : launch_tc<1>(f, a, dir, stream);
      std::vector<long long> h(64 * 8);
      CUDA_TRY(cudaStreamSynchronize(stream));
      CUDA_TRY(cudaMemcpy(h.data(), d_tr, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      cudaFree(d_tr);
      if (FILE* fp = fopen(tp, "w")) {
        for (int i = 0; i < 64; ++i) {
          for (int j = 0; j < 8; ++j) fprintf(fp, "%lld ", h[i * 8 + j]);
          fprintf(fp, "\n");
        }
        fclose(fp);
      }
      return rc;
    }
  }
  if (f->kernel == BCNF_KERNEL_TCGEN05) {
    if (f->s2_use) return f->npass == 3 ? launch_tc2<3>(f, a, dir, stream) : launch_tc2<1>(f, a, dir, stream);
    if (!f->tc1_ok) return fail(BCNF_E_STATE, "internal: no fused tensor-core kernel planned for this stack");
    return f->npass == 3 ? launch_tc<3>(f, a, dir, stream) : launch_tc<1>(f, a, dir, stream);
  }
  if (f->kernel == BCNF_KERNEL_ROWTHREAD) {
    const int hp = f->sd.half[0].hp[0];
    // rows per thread: 2 where the registers allow it (width 16), 1 for width 32; BCNF_ROWTHREAD_R (read at create) overrides
    const int R = f->env_rowthread_r > 0 ? f->env_rowthread_r : (hp == 16 ? 2 : 1);
    if (f->sd.D == 19 && hp == 16) return R == 2 ? launch_rowthread<19, 16, 2>(f, a, dir, stream) : launch_rowthread<19, 16, 1>(f, a, dir, stream);
    if (f->sd.D == 19 && hp == 32) return R == 2 ? launch_rowthread<19, 32, 2>(f, a, dir, stream) : launch_rowthread<19, 32, 1>(f, a, dir, stream);
    if (f->sd.D == 21 && hp == 16) return R == 2 ? launch_rowthread<21, 16, 2>(f, a, dir, stream) : launch_rowthread<21, 16, 1>(f, a, dir, stream);
    if (f->sd.D == 21 && hp == 32) return R == 2 ? launch_rowthread<21, 32, 2>(f, a, dir, stream) : launch_rowthread<21, 32, 1>(f, a, dir, stream);
    return fail(BCNF_E_STATE, "internal: no row-per-thread instance for D=%d HP=%d", f->sd.D, hp);
  }
  if (f->tiled_R == 32) return launch_tiled<32>(f, a, stream);
  return launch_tiled<16>(f, a, stream);
}

// Posterior sampling with the latent drawn inside the kernel (bcnf_b200.h): inverse pass on z = sigma * N(0, 1).
extern "C" int bcnf_flow_sample(bcnf_flow_t* f, uint64_t seed, float sigma, const float* P, const int32_t* row2inst,
                                int64_t inst_period, int64_t n_rows, float* x, float* logdet, void* stream) {
  NVTX_RANGE("bcnf_flow_sample");
  if (!x) return fail(BCNF_E_ARG, "bcnf_flow_sample: null output");
  FlowExtra ex;
  ex.draw = true; ex.seed = seed; ex.sigma = sigma;
  return run_flow(f, 1, nullptr, P, row2inst, inst_period, n_rows, x, logdet, stream, ex);
}

// Calibration ranks fused behind the sampler: ranks[i, j] += #{rows r of instance i : x[r, j] < y[i, j]}; the samples
// themselves are never written.  z drawn inside the kernel as in bcnf_flow_sample, or read from `z` if it is not null.
extern "C" int bcnf_flow_sample_ranks(bcnf_flow_t* f, const float* z, uint64_t seed, float sigma, const float* P,
                                      const int32_t* row2inst, int64_t inst_period, int64_t n_rows, const float* y,
                                      int32_t* ranks, void* stream) {
  NVTX_RANGE("bcnf_flow_sample_ranks");
  if (!y || !ranks) return fail(BCNF_E_ARG, "bcnf_flow_sample_ranks: null argument");
  FlowExtra ex;
  ex.draw = z == nullptr; ex.seed = seed; ex.sigma = sigma; ex.rank_y = y; ex.rank_out = ranks;
  return run_flow(f, 1, z, P, row2inst, inst_period, n_rows, nullptr, nullptr, stream, ex);
}

// Re-simulation of n parameter sets (bcnf_b200.h; reference src/bcnf/simulation/physics.py:53-165).
extern "C" int bcnf_resimulate(const double* params, int64_t n, int32_t n_steps, double dt, int32_t substeps,
                               int32_t break_on_impact, double* x_out, int32_t device, void* stream) {
  NVTX_RANGE("bcnf_resimulate");
  if (n < 0 || n_steps < 1 || substeps < 1 || !(dt > 0.0)) return fail(BCNF_E_ARG, "bcnf_resimulate: bad size");
  if (n == 0) return BCNF_OK;
  if (!params || !x_out) return fail(BCNF_E_ARG, "bcnf_resimulate: null argument");
  DEVICE_GUARD(device);
  const int threads = 128;
  const long long blocks = (n + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return fail(BCNF_E_UNSUPPORTED, "bcnf_resimulate: too many trajectories for one launch");
  resim_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(params, n, n_steps, dt, substeps, break_on_impact, x_out);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

extern "C" int bcnf_flow_forward(bcnf_flow_t* f, const float* y, const float* P, const int32_t* row2inst,
                                 int64_t inst_period, int64_t n_rows, float* z, float* logdet, void* stream) {
  NVTX_RANGE("bcnf_flow_forward");
  return run_flow(f, 0, y, P, row2inst, inst_period, n_rows, z, logdet, stream);
}

extern "C" int bcnf_flow_inverse(bcnf_flow_t* f, const float* z, const float* P, const int32_t* row2inst,
                                 int64_t inst_period, int64_t n_rows, float* x, float* logdet, void* stream) {
  NVTX_RANGE("bcnf_flow_inverse");
  return run_flow(f, 1, z, P, row2inst, inst_period, n_rows, x, logdet, stream);
}

// ----------------------------------------------------------------------------------------------
// training primitives
// ----------------------------------------------------------------------------------------------
static_assert(sizeof(bcnf_gemm_args_t) == sizeof(GemmArgs), "bcnf_gemm_args_t and GemmArgs must match");

static int g_train_gemm_mode = 0;
static long long* g_train_trace = nullptr;   // debug: device buffer of 64 clock64 stamps (bcnf_train_gemm_trace)

// Debug aid (tools/tc_gemm_check.py): run one tensor-core GEMM with the pipeline stamps of CTA (0,0,0) recorded.
extern "C" int bcnf_train_gemm_trace(const bcnf_gemm_args_t* args, int32_t device, void* stream, int64_t* out64) {
  DEVICE_GUARD(device);
  long long* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 64 * sizeof(long long)));
  CUDA_TRY(cudaMemset(d, 0, 64 * sizeof(long long)));
  g_train_trace = d;
  const int rc = bcnf_train_gemm(args, device, stream);
  g_train_trace = nullptr;
  if (rc == BCNF_OK) {
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    CUDA_TRY(cudaMemcpy(out64, d, 64 * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  cudaFree(d);
  return rc;
}

extern "C" int bcnf_train_set_gemm_mode(int32_t mode) {
  const int old = g_train_gemm_mode;
  g_train_gemm_mode = mode;
  return old;
}

template <int BN>
static int launch_train_tc_bn(const GemmArgs& g, cudaStream_t stream) {
  using Cfg = TgCfg<BN>;
  static bool attr_set[64] = {};
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    CUDA_TRY(cudaFuncSetAttribute(train_tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem));
    attr_set[dev] = true;
  }
  const int tiles_m = (g.M + kTgBM - 1) / kTgBM, tiles_n = (g.N + BN - 1) / BN;
  const long long tiles = (long long)tiles_m * tiles_n;
  const int n_kc = (g.K + 63) / 64;
  // split K over gridDim.z when the tiles alone leave most SMs idle: at batch 256 a CTA's time is the L2 -> SM
  // transfer of its operand panels (about 64 B/clk per SM), so spreading K over more SMs is what shortens it
  int split = 1;
  if (g.split_k != 1 && g.ws && g.counters && tiles <= g.n_counters && tiles * kTgBM * BN <= g.ws_floats && tiles < 120) {
    split = (int)(148 / tiles);
    if (g.split_k > 1 && split > g.split_k) split = g.split_k;
    if (split > n_kc) split = n_kc;
    if (split < 1) split = 1;
  }
  TgExtra x;
  x.ws = g.ws;
  x.counters = g.counters;
  x.colsum = g.colsum;
  x.trace = g_train_trace;
  x.chunks_per_split = (n_kc + split - 1) / split;
  split = (n_kc + x.chunks_per_split - 1) / x.chunks_per_split;
  dim3 grid(tiles_n, tiles_m, split);
  train_tc_gemm_kernel<BN><<<grid, kTgThreads, Cfg::smem, stream>>>(g, x);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

template <int BN>
static int launch_train_tc2_bn(const GemmArgs& g, cudaStream_t stream) {
  using Cfg = T2Cfg<BN>;
  static bool attr_set[64] = {};
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    CUDA_TRY(cudaFuncSetAttribute(train_tc2_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem));
    attr_set[dev] = true;
  }
  const int tiles_m = (g.M + kTgBM - 1) / kTgBM, tiles_n = (g.N + BN - 1) / BN;
  const int n_kc = (g.K + 63) / 64;
  if (g.a_rpad % 128 || g.a_rpad < tiles_m * kTgBM || g.a_plane < (long long)n_kc * g.a_rpad * 128)
    return fail(BCNF_E_ARG, "bcnf_train_gemm: A image too small (rpad=%d plane=%lld for M=%d K=%d)", g.a_rpad, g.a_plane, g.M, g.K);
  if (g.b_rpad < tiles_n * BN || g.b_plane < (long long)n_kc * g.b_rpad * 128)
    return fail(BCNF_E_ARG, "bcnf_train_gemm: B image too small (rpad=%d plane=%lld for N=%d K=%d)", g.b_rpad, g.b_plane, g.N, g.K);
  if (g.c_img && (g.c_rpad <= 0 || g.c_plane % ((long long)g.c_rpad * 128)))
    return fail(BCNF_E_ARG, "bcnf_train_gemm: bad C image (rpad=%d plane=%lld)", g.c_rpad, g.c_plane);
  ImgArgs im;
  im.a_img = g.a_img; im.a_plane = g.a_plane; im.a_rpad = g.a_rpad;
  im.b_img = g.b_img; im.b_plane = g.b_plane; im.b_rpad = g.b_rpad;
  im.c_img = g.c_img; im.c_plane = g.c_plane; im.c_rpad = g.c_rpad;
  im.colsum = g.colsum;
  dim3 grid(tiles_n, tiles_m);
  train_tc2_gemm_kernel<BN><<<grid, kT2Threads, Cfg::smem, stream>>>(g, im);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

template <int BN>
static int launch_train_tc3_bn(const GemmArgs& g, cudaStream_t stream) {
  using Cfg = T3Cfg<BN>;
  static bool attr_set[64] = {};
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    CUDA_TRY(cudaFuncSetAttribute(train_tc3_dw_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem));
    attr_set[dev] = true;
  }
  const int n_kc = (g.K + 63) / 64;
  if (g.a_rpad <= 0 || g.b_rpad <= 0 || g.a_rpad < n_kc * 64 || g.b_rpad < n_kc * 64 || g.a_plane % ((long long)g.a_rpad * 128) ||
      g.b_plane % ((long long)g.b_rpad * 128) || g.a_plane / ((long long)g.a_rpad * 128) * 64 < g.M ||
      g.b_plane / ((long long)g.b_rpad * 128) * 64 < g.N)
    return fail(BCNF_E_ARG, "bcnf_train_gemm: images too small for the weight-gradient mode (M=%d N=%d K=%d)", g.M, g.N, g.K);
  ImgArgs im;
  memset(&im, 0, sizeof(im));
  im.a_img = g.a_img; im.a_plane = g.a_plane; im.a_rpad = g.a_rpad;
  im.b_img = g.b_img; im.b_plane = g.b_plane; im.b_rpad = g.b_rpad;
  dim3 grid((g.N + BN - 1) / BN, (g.M + kTgBM - 1) / kTgBM);
  train_tc3_dw_kernel<BN><<<grid, kT2Threads, Cfg::smem, stream>>>(g, im);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

static int launch_train_tc2(const GemmArgs& g, int force_bn, cudaStream_t stream) {
  if (g.img_mn) {
    if (g.epi != TEPI_NONE || g.colsum || g.c_img) return fail(BCNF_E_ARG, "bcnf_train_gemm: the weight-gradient mode has no epilogue options");
    return force_bn == 128 ? launch_train_tc3_bn<128>(g, stream) : launch_train_tc3_bn<64>(g, stream);
  }
  const long long tm = (g.M + kTgBM - 1) / kTgBM;
  int bn = force_bn;
  // at a few hundred rows the CTA's time is the L2 -> SM transfer of its operand panels: narrow tiles, more SMs
  if (bn == 0) bn = tm * ((g.N + 127) / 128) >= 148 ? 128 : (tm * ((g.N + 63) / 64) >= 148 ? 64 : 32);
  if (bn == 128) return launch_train_tc2_bn<128>(g, stream);
  if (bn == 64) return launch_train_tc2_bn<64>(g, stream);
  if (bn == 32) return launch_train_tc2_bn<32>(g, stream);
  return fail(BCNF_E_ARG, "bcnf_train_gemm: tile width %d (32, 64 or 128)", bn);
}

static_assert(sizeof(bcnf_img_pack_desc_t) == sizeof(ImgPackDesc), "bcnf_img_pack_desc_t and ImgPackDesc must match");

extern "C" int bcnf_img_pack(const bcnf_img_pack_desc_t* descs, int32_t n, int32_t device, void* stream) {
  if (n < 0 || (n > 0 && !descs)) return fail(BCNF_E_ARG, "bcnf_img_pack: bad argument");
  DEVICE_GUARD(device);
  for (int b0 = 0; b0 < n; b0 += kImgPackMax) {
    ImgPackBatch batch;
    const int nb = n - b0 < kImgPackMax ? n - b0 : kImgPackMax;
    int max_blocks = 0;
    for (int i = 0; i < nb; ++i) {
      const bcnf_img_pack_desc_t& d = descs[b0 + i];
      if (!d.src || !d.dst || d.rpad <= 0 || d.rpad % 32 || d.chunks <= 0 || d.plane < (long long)d.chunks * d.rpad * 128 ||
          d.rows < 0 || d.rows > d.rpad || d.k < 0 || d.k > d.chunks * 64)
        return fail(BCNF_E_ARG, "bcnf_img_pack: bad descriptor %d", b0 + i);
      memcpy(&batch.d[i], &d, sizeof(ImgPackDesc));
      if (d.rpad / 32 > max_blocks) max_blocks = d.rpad / 32;
    }
    img_pack_kernel<<<dim3(max_blocks, nb), kTgGroupThreads, 0, (cudaStream_t)stream>>>(batch);
    CUDA_TRY(cudaGetLastError());
  }
  return BCNF_OK;
}

static int launch_train_tc(const GemmArgs& g, int force_bn, cudaStream_t stream) {
  const long long tm = (g.M + kTgBM - 1) / kTgBM;
  int bn = force_bn;
  if (bn == 0) {
    if (tm * ((g.N + 127) / 128) >= 148) bn = 128;
    else bn = 64;
  }
  if (bn == 128) return launch_train_tc_bn<128>(g, stream);
  if (bn == 64) return launch_train_tc_bn<64>(g, stream);
  if (bn == 32) return launch_train_tc_bn<32>(g, stream);
  return fail(BCNF_E_ARG, "bcnf_train_gemm: tile width %d (32, 64 or 128)", bn);
}

extern "C" int bcnf_train_gemm(const bcnf_gemm_args_t* args, int32_t device, void* stream_) {
  const bool images = args && args->a_img && args->b_img;
  if (!args || !args->C || (!images && (!args->A || !args->B))) return fail(BCNF_E_ARG, "bcnf_train_gemm: null argument");
  if (args->M < 0 || args->N < 0 || args->K < 0) return fail(BCNF_E_ARG, "bcnf_train_gemm: negative size");
  if (args->epilogue < 0 || args->epilogue > 3) return fail(BCNF_E_ARG, "bcnf_train_gemm: unknown epilogue %d", args->epilogue);
  if ((args->epilogue == BCNF_EPI_BIAS || args->epilogue == BCNF_EPI_BIAS_GELU_DROP) && !args->bias)
    return fail(BCNF_E_ARG, "bcnf_train_gemm: bias missing");
  if (args->epilogue == BCNF_EPI_BIAS_GELU_DROP && !args->save) return fail(BCNF_E_ARG, "bcnf_train_gemm: save missing");
  if (args->epilogue == BCNF_EPI_DGELU_DROP && !args->saved) return fail(BCNF_E_ARG, "bcnf_train_gemm: saved missing");
  if (args->p_drop < 0.f || args->p_drop >= 1.f) return fail(BCNF_E_ARG, "bcnf_train_gemm: p_drop=%f", args->p_drop);
  if (args->M == 0 || args->N == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  GemmArgs g;
  memcpy(&g, args, sizeof(g));
  cudaStream_t stream = (cudaStream_t)stream_;
  const int mode = g_train_gemm_mode & 0xF, force_bn = (g_train_gemm_mode >> 4) & 0xFF;
  if (images) {
    if (g.K < 1) return fail(BCNF_E_ARG, "bcnf_train_gemm: K = 0 with operand images");
    return launch_train_tc2(g, force_bn, stream);
  }
  // tensor cores when the problem fills a 128-row tile reasonably; slivers (K = 10 own-half inputs, the 18-wide last
  // Linear and its gradients) stay on the FMA kernel
  const bool tc_legal = g.K >= 16;
  const bool tc_auto = g.M >= 64 && g.N >= 32 && g.K >= 64;
  if (tc_legal && (mode == 2 || (mode == 0 && tc_auto))) return launch_train_tc(g, force_bn, stream);
  // small problems: 32x32 tiles put more CTAs in flight (a 256-row batch is 4 tiles of 64)
  const long long ctas64 = (long long)((g.M + 63) / 64) * ((g.N + 63) / 64);
  if (ctas64 >= 296) {
    dim3 grid((g.N + 63) / 64, (g.M + 63) / 64);
    train_gemm_kernel<64, 64><<<grid, 256, 0, stream>>>(g);
  } else {
    dim3 grid((g.N + 31) / 32, (g.M + 31) / 32);
    train_gemm_kernel<32, 32><<<grid, 256, 0, stream>>>(g);
  }
  CUDA_TRY(cudaGetLastError());
  if (g.colsum) {   // same contract as the tensor-core epilogue: column sums of the values just written
    colsum_kernel<<<(g.N + 31) / 32, 256, 0, stream>>>(g.C, g.M, g.N, g.cs0, g.colsum, 1.0f);
    CUDA_TRY(cudaGetLastError());
  }
  return BCNF_OK;
}

extern "C" int bcnf_train_colsum(const float* X, int32_t M, int32_t N, int64_t ldx, float* out, float beta,
                                 int32_t device, void* stream_) {
  if (!X || !out) return fail(BCNF_E_ARG, "bcnf_train_colsum: null argument");
  if (M < 0 || N < 0) return fail(BCNF_E_ARG, "bcnf_train_colsum: negative size");
  if (N == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  colsum_kernel<<<(N + 31) / 32, 256, 0, (cudaStream_t)stream_>>>(X, M, N, ldx, out, beta);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

// Adam over a flat blob (bcnf_b200.h; reference step: torch.optim.Adam in trainer.py:271).
extern "C" int bcnf_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, const float* step,
                              int32_t device, void* stream_) {
  if (!p || !g || !m || !v || !hyper || !step) return fail(BCNF_E_ARG, "bcnf_adam_flat: null argument");
  if (n < 0 || n % 4) return fail(BCNF_E_ARG, "bcnf_adam_flat: n must be a non-negative multiple of 4");
  if ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) != 0) return fail(BCNF_E_ARG, "bcnf_adam_flat: 16-byte alignment");
  if (n == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  int n_sm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
  const long long n4 = n / 4;
  const long long blocks = std::min<long long>((n4 + 255) / 256, (long long)n_sm * 8);
  adam_flat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(p, g, m, v, n4, hyper, step);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

// NLL loss of a batch + its gradient seeds (bcnf_b200.h; reference inn_nll_loss, utils.py:49-53).
extern "C" int bcnf_train_nll(const float* z, const float* logdet, int32_t B, int32_t D, float* loss, float* dz, float* dlogdet,
                              int32_t device, void* stream_) {
  if (!z || !logdet || !loss || !dz || !dlogdet) return fail(BCNF_E_ARG, "bcnf_train_nll: null argument");
  if (B < 1 || D < 1) return fail(BCNF_E_ARG, "bcnf_train_nll: empty batch");
  DEVICE_GUARD(device);
  const int threads = B >= 1024 ? 1024 : (B + 31) / 32 * 32;
  nll_kernel<<<1, threads, 0, (cudaStream_t)stream_>>>(z, logdet, B, D, loss, dz, dlogdet);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

extern "C" int bcnf_train_dropout_mask(float* out, int32_t M, int32_t N, uint64_t seed, uint32_t layer_uid, float p_drop,
                                       const uint64_t* seed_ptr, int32_t device, void* stream_) {
  if (!out) return fail(BCNF_E_ARG, "bcnf_train_dropout_mask: null argument");
  if (M < 0 || N < 0 || p_drop < 0.f || p_drop >= 1.f) return fail(BCNF_E_ARG, "bcnf_train_dropout_mask: bad argument");
  const long long n = (long long)M * N;
  if (n == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  dropout_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(out, M, N, seed, layer_uid, p_drop,
                                                                                            (const unsigned long long*)seed_ptr);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

// ----------------------------------------------------------------------------------------------
// fused training kernels around the hidden-layer GEMMs
// ----------------------------------------------------------------------------------------------
static_assert(sizeof(bcnf_train_pre_args_t) == sizeof(TrainPreArgs), "bcnf_train_pre_args_t / TrainPreArgs");
static_assert(sizeof(bcnf_train_post_args_t) == sizeof(TrainPostArgs), "bcnf_train_post_args_t / TrainPostArgs");
static_assert(sizeof(bcnf_train_post_bwd_args_t) == sizeof(TrainPostBwdArgs), "bcnf_train_post_bwd_args_t / TrainPostBwdArgs");
static_assert(sizeof(bcnf_train_pre_bwd_args_t) == sizeof(TrainPreBwdArgs), "bcnf_train_pre_bwd_args_t / TrainPreBwdArgs");

constexpr size_t kGlueSmemBudget = 96 * 1024;   // dynamic shared memory of a glue kernel (parameter tile)

static int glue_allow_smem(const void* kernel, size_t bytes) {
  if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGlueSmemBudget));
  return BCNF_OK;
}

// K tile of the last Linear staged in shared memory: the whole matrix when it fits the budget
// (row_bufs: per-warp staging rows of kt floats next to the tile; pad: extra floats per tile row)
static int glue_wout_tile(int no, int H, const void* kernel, int row_bufs, int pad, int* kt, size_t* smem) {
  int t = (H + 31) / 32 * 32;
  auto bytes = [&](int tt) { return ((size_t)no * (tt + pad) + (size_t)row_bufs * tt) * 4; };
  while (bytes(t) > kGlueSmemBudget && t > 32) t = (t / 2 + 31) / 32 * 32;
  *kt = t;
  *smem = bytes(t);
  return glue_allow_smem(kernel, *smem);
}

static int check_glue_dims(const char* what, int B, int D, int H, int half0, int half_n, int n_ops) {
  if (B < 0 || D < 1 || D > 64 || H < 0) return fail(BCNF_E_ARG, "%s: bad size (B=%d D=%d H=%d; D <= 64)", what, B, D, H);
  if (half_n < 0 || half_n > 32 || half0 < 0 || half0 + half_n > D) return fail(BCNF_E_ARG, "%s: bad half [%d, %d) of %d", what, half0, half0 + half_n, D);
  if (n_ops < 0 || n_ops > kGlueMaxOps) return fail(BCNF_E_ARG, "%s: %d glue ops (max %d)", what, n_ops, kGlueMaxOps);
  return BCNF_OK;
}

extern "C" int bcnf_train_pre(const bcnf_train_pre_args_t* args, int32_t device, void* stream) {
  if (!args || !args->y || !args->W1 || !args->P || !args->pre || !args->act) return fail(BCNF_E_ARG, "bcnf_train_pre: null argument");
  if (int rc = check_glue_dims("bcnf_train_pre", args->B, args->D, args->H, args->src0, args->din, 0)) return rc;
  if (args->p_drop < 0.f || args->p_drop >= 1.f) return fail(BCNF_E_ARG, "bcnf_train_pre: p_drop=%f", args->p_drop);
  if (args->B == 0 || args->H == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  TrainPreArgs a;
  memcpy(&a, args, sizeof(a));
  dim3 grid((a.H + 127) / 128, (a.B + kPreRows - 1) / kPreRows);
  train_pre_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

extern "C" int bcnf_train_post(const bcnf_train_post_args_t* args, int32_t device, void* stream) {
  if (!args || !args->y_in || !args->y_out || !args->ld) return fail(BCNF_E_ARG, "bcnf_train_post: null argument");
  if (args->a && (!args->Wout || !args->bout || !args->ls_save || !args->ydst_save)) return fail(BCNF_E_ARG, "bcnf_train_post: null argument");
  if (int rc = check_glue_dims("bcnf_train_post", args->B, args->D, args->H, args->dst0, args->dout, args->n_ops)) return rc;
  if (args->B == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  TrainPostArgs a;
  memcpy(&a, args, sizeof(a));
  int kt = 0;
  size_t smem = 0;
  if (a.a) { if (int rc = glue_wout_tile(2 * a.dout, a.H, (const void*)train_post_kernel, kGlueWarps, 1, &kt, &smem)) return rc; }
  train_post_kernel<<<(a.B + kGlueWarps - 1) / kGlueWarps, 32 * kGlueWarps, smem, (cudaStream_t)stream>>>(a, kt);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

extern "C" int bcnf_train_post_bwd(const bcnf_train_post_bwd_args_t* args, int32_t device, void* stream) {
  if (!args || !args->dz_in || !args->dz_out || !args->dld) return fail(BCNF_E_ARG, "bcnf_train_post_bwd: null argument");
  if (args->Wout && (!args->ls_save || !args->ydst_save || !args->pre || !args->d_o || !args->d_pre))
    return fail(BCNF_E_ARG, "bcnf_train_post_bwd: null argument");
  if (int rc = check_glue_dims("bcnf_train_post_bwd", args->B, args->D, args->H, args->dst0, args->dout, args->n_ops)) return rc;
  for (int o = 0; o < args->n_ops; ++o)
    if (args->ops[o].type == GLUE_ACTNORM && (!args->ops[o].save || !args->ops[o].g0 || !args->ops[o].g1))
      return fail(BCNF_E_ARG, "bcnf_train_post_bwd: ActNorm op %d lacks save / gradient buffers", o);
  if (args->B == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  TrainPostBwdArgs a;
  memcpy(&a, args, sizeof(a));
  int kt = 0;
  size_t smem = 0;
  if (a.Wout) { if (int rc = glue_wout_tile(2 * a.dout, a.H, (const void*)train_post_bwd_kernel, 0, 0, &kt, &smem)) return rc; }
  train_post_bwd_kernel<<<(a.B + kGlueWarps - 1) / kGlueWarps, 32 * kGlueWarps, smem, (cudaStream_t)stream>>>(a, kt);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}

extern "C" int bcnf_train_pre_bwd(const bcnf_train_pre_bwd_args_t* args, int32_t device, void* stream) {
  if (!args || !args->d_pre || !args->W1 || !args->dz) return fail(BCNF_E_ARG, "bcnf_train_pre_bwd: null argument");
  if (int rc = check_glue_dims("bcnf_train_pre_bwd", args->B, args->D, args->H, args->src0, args->din, 0)) return rc;
  if (args->B == 0) return BCNF_OK;
  DEVICE_GUARD(device);
  TrainPreBwdArgs a;
  memcpy(&a, args, sizeof(a));
  const int wp = a.din > 0 ? a.din : 1;    // lanes = consecutive columns: any pitch is conflict-free
  int jt = (a.H + 31) / 32 * 32;
  while (((size_t)jt * wp + (size_t)kGlueWarps * jt) * 4 > kGlueSmemBudget && jt > 32) jt = (jt / 2 + 31) / 32 * 32;
  const size_t smem = ((size_t)jt * wp + (size_t)kGlueWarps * jt) * 4;
  if (int rc = glue_allow_smem((const void*)train_pre_bwd_kernel, smem)) return rc;
  train_pre_bwd_kernel<<<(a.B + kGlueWarps - 1) / kGlueWarps, 32 * kGlueWarps, smem, (cudaStream_t)stream>>>(a, jt, wp);
  CUDA_TRY(cudaGetLastError());
  return BCNF_OK;
}
