/*
 * bcnf_b200 -- C ABI of the B200-native CondRealNVP_v2 coupling-stack path.
 *
 * The reference (psaegert/bcnf) has no plugin / FFI interface: its boundary is the Python
 * class CondRealNVP_v2 (src/bcnf/models/cnf.py:357-588) and its state_dict layout.  This
 * header is the ABI that sits directly beneath that class; bcnf_b200/cnf.py binds it with
 * ctypes.  Every entry point names the reference code it replaces.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers to contiguous row-major fp32, 16-byte aligned bases;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are
 *     asynchronous on it, make no hidden synchronisation and start no host threads;
 *   - return value 0 = success, < 0 = argument/validation error (BCNF_E_*), > 0 = cudaError_t;
 *     nothing throws across the ABI; bcnf_last_error() returns the message of the last
 *     failing call on the calling thread;
 *   - PyTorch (or any caller) owns inputs, outputs and the model parameters; the library owns
 *     only the packed copy of the parameters held by a bcnf_flow_t.
 */
#ifndef BCNF_B200_H
#define BCNF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCNF_ABI_VERSION 2
#define BCNF_MAX_HIDDEN_LAYERS 8
#define BCNF_MAX_SIZE 64         /* largest supported flow dimension D */
#define BCNF_MAX_HIDDEN 1024     /* largest supported conditioner width */

enum {
  BCNF_OK = 0,
  BCNF_E_ARG = -1,          /* null pointer, negative size, bad enum */
  BCNF_E_UNSUPPORTED = -2,  /* shape outside the supported envelope */
  BCNF_E_STATE = -3,        /* parameters not set yet */
  BCNF_E_NOMEM = -4
};

/* Layer kinds of CondRealNVP_v2.layers (cnf.py:392-423). */
typedef enum {
  BCNF_OP_ACTNORM = 0,   /* ActNorm, cnf.py:342-354 */
  BCNF_OP_COUPLING = 1,  /* ConditionalAffineCouplingLayer, cnf.py:110-213 */
  BCNF_OP_ORTHO = 2      /* OrthonormalTransformation, cnf.py:312-339 */
} bcnf_op_type_t;

/* Arithmetic of the conditioner GEMMs. */
typedef enum {
  BCNF_PREC_FP32 = 0,    /* fp32 FMA everywhere (all widths) */
  BCNF_PREC_BF16X3 = 1,  /* tcgen05, 3-term bf16 split, fp32 accumulate: fp32-class accuracy */
  BCNF_PREC_BF16 = 2     /* tcgen05, single bf16 pass, fp32 accumulate: stated bf16 tolerance */
} bcnf_precision_t;

/* Which kernel family a handle dispatches to (bcnf_flow_info). */
typedef enum {
  BCNF_KERNEL_ROWTHREAD = 0, /* one row per thread, hidden vector in registers (width <= 32) */
  BCNF_KERNEL_TILED = 1,     /* row tile per CTA, fp32 FMA, activations in shared memory */
  BCNF_KERNEL_TCGEN05 = 2    /* row tile per CTA, tcgen05 MMA with TMEM accumulators */
} bcnf_kernel_t;

/* Constructor arguments of CondRealNVP_v2 that shape the stack (cnf.py:358-375). */
typedef struct {
  int32_t size;                            /* D   (`size`) */
  int32_t n_conditions;                    /* C   (`n_conditions`, > 0) */
  int32_t n_hidden;                        /* L   (len(nested_sizes)), 1..BCNF_MAX_HIDDEN_LAYERS */
  int32_t hidden[BCNF_MAX_HIDDEN_LAYERS];  /* H_1..H_L (`nested_sizes`) */
  int32_t two_way;                         /* `two_way` */
  int32_t n_ops;                           /* len(layers) of the (sub-)stack to run */
  int32_t precision;                       /* bcnf_precision_t */
  int32_t device;                          /* CUDA device ordinal */
} bcnf_flow_desc_t;

/* Parameters of one layer, in the reference's own storage layout (torch.nn.Linear.weight is
 * (out, in) row-major).  Unused members are NULL. */
typedef struct {
  int32_t type;              /* bcnf_op_type_t */
  int32_t reserved;
  const float* scale;        /* ACTNORM  (D)      layers.{i}.scale              cnf.py:345 */
  const float* bias;         /* ACTNORM  (D)      layers.{i}.bias               cnf.py:346 */
  const float* q;            /* ORTHO    (D, D)   layers.{i}.orthonormal_matrix cnf.py:323 */
  const float* const* w_a;   /* COUPLING L+1 ptrs layers.{i}.nn_a.nn.{j}.weight cnf.py:78-85 */
  const float* const* b_a;   /* COUPLING L+1 ptrs layers.{i}.nn_a.nn.{j}.bias */
  const float* const* w_b;   /* two_way only      layers.{i}.nn_b.nn.{j}.weight cnf.py:150-160 */
  const float* const* b_b;
} bcnf_op_params_t;

typedef struct bcnf_flow bcnf_flow_t;

typedef struct {
  int32_t kernel;            /* bcnf_kernel_t used by forward / inverse */
  int32_t proj_width;        /* floats per instance of the condition projection P */
  int32_t n_half_couplings;  /* conditioner networks in the stack (2 per two-way coupling) */
  int32_t rows_per_cta;      /* row tile of the flow kernel */
  int64_t packed_bytes;      /* device bytes owned by the handle */
  int64_t macs_per_row;      /* algorithmic multiply-accumulates per row, projection hoisted */
  int64_t macs_per_instance; /* multiply-accumulates of the projection, per instance */
} bcnf_flow_info_t;

int bcnf_abi_version(void);
const char* bcnf_last_error(void);

/* Build a handle for a layer sequence `op_types[0..n_ops)` (the order of
 * CondRealNVP_v2.layers, cnf.py:392-423, or any contiguous slice of it -- a single
 * ConditionalAffineCouplingLayer is a 1-op stack).  Allocates the packed-parameter storage. */
int bcnf_flow_create(const bcnf_flow_desc_t* desc, const int32_t* op_types, bcnf_flow_t** out);
int bcnf_flow_destroy(bcnf_flow_t* flow);
int bcnf_flow_info(const bcnf_flow_t* flow, bcnf_flow_info_t* info);
/* Diagnostics of the fused tensor-core kernel: every wait inside it carries a watchdog that records
 * {wait code, aux, block, thread} and traps (the launch fails with a CUDA error) instead of hanging the
 * device if a hand-off never arrives.  All zero in a healthy run.  No reference counterpart. */
int bcnf_flow_debug_words(const bcnf_flow_t* flow, uint32_t* out4);

/* (Re)pack the parameters from the caller's tensors (device pointers).  Call after
 * load_state_dict / every optimiser step.  Replaces nothing in the reference: it is the
 * price of not reading nn.Module parameters one ATen op at a time. */
int bcnf_flow_set_params(bcnf_flow_t* flow, const bcnf_op_params_t* ops, void* stream);

/* P[i, :] = W1h . h[i, :] + b1 for every conditioner network of the stack: the part of the
 * first Linear of ConditionalNestedNeuralNetwork that multiplies the condition features
 * (torch.cat([y, h]) then nn.Linear, cnf.py:101-104), hoisted so that it is computed once
 * per conditioning instance instead of once per (sample, instance) row.
 *   h: (n_inst, C)   P: (n_inst, proj_width) */
int bcnf_cond_project(bcnf_flow_t* flow, const float* h, int64_t n_inst, float* P, void* stream);

/* Row r uses instance row2inst[r] if row2inst != NULL, else r % inst_period if
 * inst_period > 0 (the tiling of `c.repeat(m, 1)` in _sample(outer=True), cnf.py:579),
 * else r. */

/* Layer loop of CondRealNVP_v2.forward, cnf.py:476-488:  z (n_rows, D) and, if logdet != NULL,
 * the accumulated log|det J| (n_rows) that the reference leaves in model.log_det_J. */
int bcnf_flow_forward(bcnf_flow_t* flow, const float* y, const float* P, const int32_t* row2inst,
                      int64_t inst_period, int64_t n_rows, float* z, float* logdet, void* stream);

/* Layer loop of CondRealNVP_v2.inverse, cnf.py:500-506: x (n_rows, D).  If logdet != NULL it
 * receives the log|det| of the forward map at x (i.e. minus the inverse's own). */
int bcnf_flow_inverse(bcnf_flow_t* flow, const float* z, const float* P, const int32_t* row2inst,
                      int64_t inst_period, int64_t n_rows, float* x, float* logdet, void* stream);

/* Posterior sampling with the latent drawn INSIDE the kernel: x = inverse(sigma * z), z ~ N(0, 1) from Philox4x32-10
 * keyed by `seed` at counter row * D + j -- a pure function of (seed, row, j), independent of tiling and launch.
 * Replaces `sigma * torch.randn(...)` on the CPU generator plus the host-to-device copy of z in
 * CondRealNVP_v2._sample (cnf.py:566, :578, :584) and the z read of bcnf_flow_inverse (no z array exists at all).
 * The reference's own draw order is still available: pass its z to bcnf_flow_inverse. */
int bcnf_flow_sample(bcnf_flow_t* flow, uint64_t seed, float sigma, const float* P, const int32_t* row2inst,
                     int64_t inst_period, int64_t n_rows, float* x, float* logdet, void* stream);

/* Calibration ranks fused behind the sampler (compute_y_hat_ranks, src/bcnf/eval/calibration.py:33-48):
 *   ranks[i, j] += #{ rows r whose instance is i : x[r, j] < y[i, j] }
 * with x the inverse pass as above; the (M, N, D) sample tensor the reference materialises, copies to the host and
 * reduces there is never written.  z == NULL: drawn in the kernel (seed, sigma); else read (n_rows, D).
 *   y: (n_inst, D)   ranks: int32 (n_inst, D), zeroed (or carrying earlier chunks' counts) by the caller. */
int bcnf_flow_sample_ranks(bcnf_flow_t* flow, const float* z, uint64_t seed, float sigma, const float* P,
                           const int32_t* row2inst, int64_t inst_period, int64_t n_rows, const float* y,
                           int32_t* ranks, void* stream);

/* Re-simulation of sampled parameter sets: physics_ODE_simulation
 * (src/bcnf/simulation/physics.py:53-165) for n parameter rows at once, one thread per trajectory, instead of one
 * scipy.integrate.odeint call per trajectory in a process pool (resimulation.py:21-59).
 *   params: (n, 19) fp64, columns in the keyword order of physics_ODE_simulation:
 *           x0_x x0_y x0_z v0_x v0_y v0_z g_x g_y g_z w_x w_y w_z b m rho r a_x a_y a_z
 *   x_out:  (n, n_steps, 3) fp64, n_steps = len(np.arange(0, T, dt)); x_out[:, 0] = x0
 *   substeps: RK4 steps per output interval (16 leaves < 1e-9 of scale vs the reference's LSODA at 1.49e-8). */
int bcnf_resimulate(const double* params, int64_t n, int32_t n_steps, double dt, int32_t substeps,
                    int32_t break_on_impact, double* x_out, int32_t device, void* stream);

/* ---- training primitives (Trainer._train_batch, src/bcnf/train/trainer.py:244-277) ------------------
 * The conditioner's Linear -> GELU -> Dropout chain (cnf.py:78-83) forward and backward as one strided
 * SGEMM with fused epilogues; parameters and gradients stay in the reference's (out, in) layout. */
typedef enum {
  BCNF_EPI_NONE = 0,            /* C = A.B (+ beta C)                                        */
  BCNF_EPI_BIAS = 1,            /* C = A.B + bias[j]                         (last Linear)   */
  BCNF_EPI_BIAS_GELU_DROP = 2,  /* save = A.B + bias[j]; C = dropout(gelu(save))  (nn.Linear, nn.GELU, nn.Dropout) */
  BCNF_EPI_DGELU_DROP = 3       /* C = (A.B) * gelu'(saved) * dropout mask        (their autograd backward)        */
} bcnf_epilogue_t;

typedef struct {
  const float* A;        /* A(i, r) = A[i*as0 + r*as1] */
  const float* B;        /* B(r, j) = B[r*bs0 + j*bs1] */
  float* C;              /* C(i, j) = C[i*cs0 + j]     */
  int32_t M, N, K;
  int64_t as0, as1, bs0, bs1, cs0;
  float beta;            /* 0 or 1 */
  int32_t epilogue;      /* bcnf_epilogue_t */
  const float* bias;     /* [N] */
  float* save;           /* BCNF_EPI_BIAS_GELU_DROP: pre-activations out (indexed like C) */
  const float* saved;    /* BCNF_EPI_DGELU_DROP: pre-activations in (indexed like C) */
  uint64_t seed;         /* dropout: mask(i, j) = hash(seed, layer_uid, i*N + j) >= p_drop */
  uint32_t layer_uid;
  float p_drop;
  const uint64_t* seed_ptr;  /* optional DEVICE word XORed into seed: lets a captured CUDA graph draw new masks per replay */
  float* colsum;         /* optional [N]: colsum[j] += sum_i C(i, j) of the values written (atomic; the caller zeroes it):
                          * the bias gradient of the Linear whose output gradient this GEMM produces */
  float* ws;             /* optional split-K workspace: ws_floats fp32, all zero on entry, left all zero */
  uint32_t* counters;    /* optional split-K tile counters: n_counters words, all zero on entry, left all zero */
  int64_t ws_floats;
  int32_t n_counters;
  int32_t split_k;       /* 0: the library decides (only splits when ws and counters are given); 1: never; n > 1: at most n */
  /* Operand images.  An image of a matrix X (rows x k) is X split into bf16 hi / lo planes, each stored as
   * [ceil(k/64) chunks][rpad rows][128 bytes], 16-byte units of a row XOR-swizzled by (row & 7), zero outside the
   * valid extent: the shared-memory operand tile of tcgen05.mma, fetched by TMA bulk copies.  lo plane = base + plane.
   * With a_img (rows = i, k = r; rpad multiple of 128, >= M rounded up) and b_img (rows = j, k = r; rpad >= N rounded
   * up to 128) the TMA-fed kernel runs and A / B may be NULL.  c_img: the values written to C are also written as an
   * image (rows = i, k = j) -- the A operand of the next GEMM of the chain.  Images come from bcnf_img_pack, the
   * c_img of a previous GEMM, or the act_img / dpre_img outputs of bcnf_train_pre / bcnf_train_post_bwd. */
  const void* a_img; int64_t a_plane; int32_t a_rpad;
  int32_t img_mn;        /* 1: weight-gradient mode C(i, j) = sum_r a(r, i) b(r, j): both images are read MN-major (rows = r, the
                          * contraction index; M / N = their column counts, K = their row count) */
  const void* b_img; int64_t b_plane; int32_t b_rpad; int32_t pad1;
  void* c_img; int64_t c_plane; int32_t c_rpad; int32_t pad2;
} bcnf_gemm_args_t;

int bcnf_train_gemm(const bcnf_gemm_args_t* args, int32_t device, void* stream);

/* fp32 matrix -> operand image: X(row, k) = src[row*s_row + k*s_k] for row < rows, k < k (zero elsewhere). */
typedef struct {
  const float* src;
  int64_t s_row, s_k;
  int32_t rows, k;
  void* dst;
  int64_t plane;       /* chunks * rpad * 128 */
  int32_t rpad;        /* multiple of 32 */
  int32_t chunks;
} bcnf_img_pack_desc_t;
int bcnf_img_pack(const bcnf_img_pack_desc_t* descs, int32_t n, int32_t device, void* stream);
/* C (M, N; row pitch ldc) = A . B^T (+ bias[N]) on operand images with the CTA-pair kernel (256 x 256 tiles,
 * tcgen05.mma.cta_group::2): the GEMM behind bcnf_cond_project on tensor-core handles.  a_img: rows = M index,
 * b_img: rows = N index, both with rpad a multiple of 256 covering M resp. N; passes = 3 (bf16x3) or 1 (bf16). */
int bcnf_gemm_img(const void* a_img, int64_t a_plane, int32_t a_rpad, const void* b_img, int64_t b_plane, int32_t b_rpad,
                  float* C, int64_t ldc, const float* bias, int32_t M, int32_t N, int32_t K, int32_t passes,
                  int32_t device, void* stream);
/* Same GEMM with the epilogue gelu(acc + bias) -> operand image c_img (rows = M index, k = N index; columns >= N of the
 * image are written as zeros): Linear + nn.GELU whose output is the A operand of the next Linear -- the hidden layers of
 * the layer-by-layer stack schedule and of FullyConnectedFeatureNetwork (reference feature_network.py:114-145). */
int bcnf_gemm_img_gelu(const void* a_img, int64_t a_plane, int32_t a_rpad, const void* b_img, int64_t b_plane, int32_t b_rpad,
                       const float* bias, void* c_img, int64_t c_plane, int32_t c_rpad, int32_t M, int32_t N, int32_t K,
                       int32_t passes, int32_t device, void* stream);
/* One time step of one nn.LSTM layer and direction (LSTMFeatureNetwork, reference feature_network.py:148-178):
 * gates = [x_t | h_(t-1)] . Wcat^T + bias on the CTA-pair GEMM, cell update in the epilogue.  The A operand is gathered
 * from n_chunks 64-column image chunks (a_hi / a_lo: [a_rpad rows][128 B] each) -- the chunks of x_t and of the
 * previous step's h.  Wcat image rows and bias are gate-interleaved: n = 4 * unit + gate (gate order i, f, g, o).
 * cell (in/out) and hsum (optional, += h) are fp32 [N/8][state_rows][2]; h_t leaves as one image chunk per 64 units. */
typedef struct {
  const void* a_hi[16]; const void* a_lo[16];
  const void* b_img; int64_t b_plane;
  const float* bias;
  float* cell; float* hsum; int64_t state_rows;
  int32_t n_chunks, a_rpad, b_rpad, M;
  void* h_hi[4]; void* h_lo[4];
  int32_t N, passes, pad0, pad1;
} bcnf_lstm_step_t;
int bcnf_lstm_step(const bcnf_lstm_step_t* args, int32_t device, void* stream);
/* Transformer condition encoder (reference src/bcnf/models/feature_network.py:183-307): the pieces between its Linears,
 * each leaving the operand image the next bcnf_gemm_img / bcnf_gemm_img_gelu reads (drivers: bcnf_b200/feature_tc.py:
 * transformer_token0 for inference, bcnf_b200/trf_train.py for the Trainer's step).  rows = instances * T tokens;
 * E = trf_size (multiple of 8, <= 1024); images as above with rpad a multiple of 256 covering rows.  Pointers marked
 * "opt" may be NULL; they are the training-mode extras (dropout multipliers in, tensors saved for the backward out).
 *   bcnf_trf_embed:         x = (tokens . Wf^T + bf) * mask[opt] (+ pos[t], the (T, E) positional table, opt)  (:287-301)
 *                           -> x fp32 + image
 *   bcnf_trf_attention:     per instance and head softmax(q k^T / sqrt(E / heads)) v from qkv (rows, 3E) = q | k | v
 *                           (:207-226, no mask) -> image of the concatenated heads (+ fp32 ctx, opt); E / heads in
 *                           {8, 16, 32, 64} for T <= 32, in {8, 16} for T <= 64
 *   bcnf_trf_attention_bwd: its backward, d ctx (rows, E) -> d qkv (rows, 3E), probabilities recomputed from qkv; T <= 32
 *   bcnf_trf_add_layernorm: x_out = LayerNorm(x + mask[opt] * y) * gamma + beta (post-norm block, :255-259; eps as
 *                           nn.LayerNorm) -> x_out fp32 (may alias x) + image; s = x + mask * y, mean, rstd (rows) opt
 *   bcnf_trf_gelu:          a = gelu(u) (nn.GELU, erf) for u (rows, N), N a multiple of 8 -> image (+ fp32 a, opt)
 *   bcnf_trf_ln_param_grad: dgamma[c] += sum_rows g * (s - mean) * rstd, dbeta[c] += sum_rows g  (nn.LayerNorm's parameter
 *                           gradients from the s / mean / rstd bcnf_trf_add_layernorm saved; the caller zeroes the outputs) */
int bcnf_trf_embed(const float* tokens, const float* Wf, const float* bf, const float* pos, const float* mask, int64_t rows,
                   int32_t T, int32_t F, int32_t E, float* x, void* x_img, int64_t plane, int32_t rpad, int32_t device,
                   void* stream);
int bcnf_trf_attention(const float* qkv, int64_t n_inst, int32_t T, int32_t E, int32_t heads, float* ctx, void* ctx_img,
                       int64_t plane, int32_t rpad, int32_t device, void* stream);
int bcnf_trf_attention_bwd(const float* qkv, const float* dctx, int64_t n_inst, int32_t T, int32_t E, int32_t heads,
                           float* dqkv, int32_t device, void* stream);
int bcnf_trf_add_layernorm(const float* x, const float* y, const float* mask, const float* gamma, const float* beta, float eps,
                           int64_t rows, int32_t E, float* s, float* mean, float* rstd, float* x_out, void* x_img,
                           int64_t plane, int32_t rpad, int32_t device, void* stream);
int bcnf_trf_gelu(const float* u, int64_t rows, int32_t N, float* a, void* a_img, int64_t plane, int32_t rpad, int32_t device,
                  void* stream);
int bcnf_trf_ln_param_grad(const float* g, const float* s, const float* mean, const float* rstd, int64_t rows, int32_t E,
                           float* dgamma, float* dbeta, int32_t device, void* stream);
/* Debug aid (tools/gemm_img_check.py --trace): device buffer of 74 x 16 x 4 uint64 for the per-tile globaltimer stamps
 * of the following bcnf_gemm_img launches; NULL switches it off. */
int bcnf_gemm_img_set_trace(void* device_buffer);
/* Kernel selection of bcnf_train_gemm, for tests and A/B timing.  bits 0-3: 0 = automatic (tensor cores when the
 * problem fills a 128-row tile, fp32 FMA for slivers), 1 = fp32 FMA kernel only, 2 = tcgen05 kernel wherever it
 * is legal; bits 4-11: force the tensor-core tile width BN (32, 64 or 128; 0 = automatic).  Returns the old mode. */
int bcnf_train_set_gemm_mode(int32_t mode);
/* Debug aid: bcnf_train_gemm once, synchronously, returning 64 clock64 stamps of CTA (0,0,0) of the tensor-core kernel
 * (layout: tools/tc_gemm_check.py). */
int bcnf_train_gemm_trace(const bcnf_gemm_args_t* args, int32_t device, void* stream, int64_t* out64);
/* out[j] = sum_i X[i*ldx + j] + beta*out[j]   (bias gradients: the sum(0) of autograd's Linear backward) */
int bcnf_train_colsum(const float* X, int32_t M, int32_t N, int64_t ldx, float* out, float beta, int32_t device, void* stream);
/* Trainer step fusion (reference src/bcnf/train/trainer.py:267-277).
 * bcnf_adam_flat: one torch.optim.Adam update (no amsgrad) of n parameters held in one flat blob, with their gradients
 * and both moments in blobs of the same layout: p, g, m, v device pointers (16-byte aligned, n a multiple of 4);
 * hyper = device float[5] {lr, beta1, beta2, eps, weight_decay}; step = device float holding the 1-based step count of
 * THIS update (the caller advances it).  Replaces optimizer.step() (trainer.py:271).
 * bcnf_train_nll: loss[0] = mean_b(0.5 sum_j z[b, j]^2 - logdet[b]) (inn_nll_loss, src/bcnf/utils.py:49-53) together
 * with the gradients it sends back, dz = z / B and dlogdet = -1 / B, in one launch (trainer.py:267). */
int bcnf_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, const float* step,
                   int32_t device, void* stream);
int bcnf_train_nll(const float* z, const float* logdet, int32_t B, int32_t D, float* loss, float* dz, float* dlogdet,
                   int32_t device, void* stream);
/* the multiplicative dropout mask (0 or 1/(1-p)) the fused epilogues apply, materialised for tests */
int bcnf_train_dropout_mask(float* out, int32_t M, int32_t N, uint64_t seed, uint32_t layer_uid, float p_drop,
                            const uint64_t* seed_ptr, int32_t device, void* stream);

/* ---- fused training kernels around the hidden-layer GEMMs (one launch per coupling block and direction) ----------
 * Reference: ConditionalNestedNeuralNetwork.forward cnf.py:98-107, ConditionalAffineCouplingLayer.forward
 * cnf.py:165-196, ActNorm.forward cnf.py:348-351, OrthonormalTransformation.forward cnf.py:333-336, and their autograd
 * backward (trainer.py:268).  All pointers are device pointers, fp32; D <= 64. */
typedef struct {
  int32_t type;        /* 0: y = y @ Q (p0 = Q (D, D));  1: ActNorm y = p0 * y + p1, log-det += sum log|p0| */
  const float* p0;
  const float* p1;
  float* save;         /* ActNorm: its input (B, D), written forward, read backward */
  float* g0;           /* backward, ActNorm: gradient of scale (D), accumulated (the caller zeroes it) */
  float* g1;           /* backward, ActNorm: gradient of bias (D) */
} bcnf_glue_op_t;

/* pre = y[:, src0:src0+din] . W1[:, :din]^T + P ;  act = dropout(gelu(pre))     (P = h . W1[:, din:]^T + b1) */
typedef struct {
  const float* y; int64_t y_pitch;
  int32_t B, D, src0, din;
  const float* W1; int64_t w1_pitch;
  const float* P; int64_t p_pitch;
  int32_t H;
  float* pre; float* act; int64_t pitch;
  uint64_t seed; uint32_t layer_uid; float p_drop; const uint64_t* seed_ptr;
  void* act_img; int64_t img_plane; int32_t img_rpad;   /* optional: also write act as an operand image (see bcnf_img_*) */
} bcnf_train_pre_args_t;

/* o = a . Wout^T + bout; t, log s = o[:dout], tanh(o[dout:]); y[dst] = exp(log s) * y[dst] + t; ld += sum log s;
 * then the n_ops glue ops that follow the coupling in model.layers.  a == NULL: glue ops only. */
typedef struct {
  const float* a; int64_t a_pitch;
  const float* Wout; const float* bout;
  int32_t B, D, H, dst0, dout;
  const float* y_in; float* y_out;
  float* ld;
  float* ls_save; float* ydst_save;     /* (B, dout) each, for the backward */
  int32_t n_ops; bcnf_glue_op_t ops[4];
} bcnf_train_post_args_t;

/* backward of bcnf_train_post: dz_in, dld -> dz_out (the conditioner path into y[src] is added by
 * bcnf_train_pre_bwd), d_o (B, 2 dout) and d_pre = (d_o . Wout) * gelu'(pre) * dropout mask of the last hidden layer. */
typedef struct {
  const float* dz_in; float* dz_out;
  const float* dld;
  int32_t B, D, H, dst0, dout;
  const float* ls_save; const float* ydst_save;
  const float* Wout;                    /* NULL: glue ops only */
  const float* pre; int64_t pitch;
  float* d_o;
  float* d_pre;
  uint64_t seed; uint32_t layer_uid; float p_drop; const uint64_t* seed_ptr;
  int32_t n_ops; bcnf_glue_op_t ops[4];
  void* dpre_img; int64_t img_plane; int32_t img_rpad;  /* optional: also write d_pre as an operand image */
} bcnf_train_post_bwd_args_t;

/* dz[:, src0:src0+din] += d_pre . W1[:, :din] */
typedef struct {
  const float* d_pre; int64_t pitch;
  const float* W1; int64_t w1_pitch;
  int32_t B, D, H, src0, din;
  float* dz;
} bcnf_train_pre_bwd_args_t;

int bcnf_train_pre(const bcnf_train_pre_args_t* args, int32_t device, void* stream);
int bcnf_train_post(const bcnf_train_post_args_t* args, int32_t device, void* stream);
int bcnf_train_post_bwd(const bcnf_train_post_bwd_args_t* args, int32_t device, void* stream);
int bcnf_train_pre_bwd(const bcnf_train_pre_bwd_args_t* args, int32_t device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BCNF_B200_H */
