/*
 * libgoat_sm100 -- C ABI of the B200 (sm_100a) kernels behind the GOAT cross-modal hot path.
 *
 * The reference (CrystalSixone/VLN-GOAT) has no FFI / plugin layer for this path: every op is a
 * stock torch call inside Python nn.Modules.  The drop-in boundary is
 * therefore the Python module API (vln_goat_b200.modules mirrors the reference class names and
 * state_dict keys) and THIS header is the operator interface those modules bind through ctypes.
 * Each entry point cites the reference code it replaces.  P/ = pretrain_src/, M/ = map_nav_src/.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise.
 *  - the caller owns every buffer (inputs, outputs, workspaces); the library allocates nothing
 *    and keeps no pointer after return.  Workspace sizes come from goat_*_workspace_bytes().
 *  - all work is enqueued on the stream passed in; no implicit synchronisation, no default
 *    stream, safe under CUDA-graph capture.
 *  - every function returns 0 on success, a goat_status_t otherwise; goat_last_error() returns
 *    a thread-local human-readable message.  Nothing throws or exits.
 *  - 16-byte alignment is required for tensor base pointers (checked, never copied).
 *  - dtypes: GOAT_F32 runs fp32 SIMT kernels (parity mode, 1e-5); GOAT_F16 / GOAT_BF16 run the
 *    tcgen05 tensor-core kernels with fp32 accumulation (1e-3 mode).
 */
#ifndef GOAT_SM100_H_
#define GOAT_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* goat_stream_t; /* cudaStream_t */

typedef enum { GOAT_OK = 0, GOAT_ERR_INVALID = 1, GOAT_ERR_CUDA = 2, GOAT_ERR_UNSUPPORTED = 3 } goat_status_t;
typedef enum { GOAT_F32 = 0, GOAT_F16 = 1, GOAT_BF16 = 2 } goat_dtype_t;

/* epilogue activation of goat_gemm */
typedef enum {
  GOAT_ACT_NONE = 0,
  GOAT_ACT_GELU = 1,  /* erf-GELU, P/model/Bert_backbone.py:41-47; writes the pre-activation to aux_out */
  GOAT_ACT_RELU = 2,  /* ClsPrediction, P/model/pretrain_goat.py:27-38 */
  GOAT_ACT_DGELU = 3, /* multiply by gelu'(aux_in)  (backward of GELU) */
  GOAT_ACT_DRELU = 4, /* multiply by [aux_in > 0]   (backward of ReLU; aux_in = ReLU output) */
  GOAT_ACT_TANH = 5   /* BertPooler, P/model/Bert_backbone.py:783-795 */
} goat_act_t;

int goat_version(void);
const char* goat_last_error(void);
/* 1 if the tcgen05 path is usable on the current device (compute capability 10.x) */
int goat_device_supported(void);

/* ------------------------------------------------------------------------------------------
 * goat_gemm: out[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )
 *
 * Replaces every nn.Linear on the path (Q/K/V/out projections P/model/Bert_backbone.py:170-172,
 * :302; FFN :348,:362; heads) in forward (A = activations, B = weight [N,K]) and, with the
 * major-ness flags, both backward products (dX = dY W: b_mn_major=1; dW = dY^T X: both =1).
 *   a_mn_major = 0: A[m,k] = A[m*lda + k]   (K contiguous)     1: A[m,k] = A[k*lda + m]
 *   b_mn_major = 0: B[n,k] = B[n*ldb + k]   (K contiguous)     1: B[n,k] = B[k*ldb + n]
 * epilogue, in this order: v = alpha*acc + bias[n]; activation (see goat_act_t); dropout(p, seed)
 * with 1/(1-p) scaling; + res[m,n]; store as out_dtype into out (and a second copy in `dtype`
 * into out2 when given).
 * dtype F16/BF16 -> tcgen05.mma (TMA-fed, TMEM accumulators) when K, lda, ldb are multiples of 8
 * and K >= 16; otherwise, and always for F32, the SIMT kernel.  force_simt=1 selects the SIMT
 * kernel (used by tests as an on-device cross-check).
 * M > 128 runs the CTA-pair kernel (gemm_umma2.cu): a cluster of two CTAs shares one 256 x 256 (or 256 x 128) tile
 * through tcgen05.mma.cta_group::2, each CTA staging its 128 rows of A and half of B; M <= 128 runs the single-CTA
 * 128 x 128 kernel.  Both are persistent with two TMEM accumulators (the epilogue of one tile overlaps the main loop of
 * the next) and are launched with programmatic dependent launch (GOAT_PDL=0 disables it).
 * ------------------------------------------------------------------------------------------ */
typedef struct goat_gemm_args {
  int M, N, K;
  int dtype;      /* operand dtype of A, B, aux_in, aux_out, out2 */
  int a_mn_major, b_mn_major;
  int lda, ldb;
  const void* A;
  const void* B;
  const float* bias; /* [N] or NULL */
  const float* res;  /* [M, ldres] fp32 or NULL */
  int ldres;
  const void* aux_in; /* [M, ldaux] or NULL */
  void* aux_out;      /* [M, ldaux] or NULL */
  int ldaux;
  void* out;
  int ldc;
  int out_dtype; /* GOAT_F32 or == dtype */
  void* out2;    /* optional, dtype `dtype` */
  int ldc2;
  int act;
  float alpha;
  float drop_p;
  uint64_t drop_seed;
  const uint64_t* drop_seed_ptr; /* optional DEVICE word added to drop_seed at run time (CUDA-graph safe reseeding) */
  int force_simt;
  int accumulate; /* 1: out (fp32) += alpha * A B^T, reduced with fp32 atomics; the tcgen05 kernel then also splits K
                     across CTAs when M x N alone cannot fill the GPU (weight gradients: K = tokens).  No bias / res /
                     act / dropout / out2 in this mode; the caller zero-initialises out. */
  const void* B_lo; /* optional second weight operand with B's layout: acc = A B^T + A B_lo^T in one launch (the K loop runs
                       twice over A).  B = fp16(W), B_lo = fp16(W - B): the weight then carries ~22 mantissa bits, which
                       removes the systematic weight-rounding error of the 16-bit forward pass (profiles/
                       r02_fp16_error_budget.txt).  tcgen05 path only, B K-major, no split-K. */
} goat_gemm_args;
int goat_gemm(const goat_gemm_args* args, goat_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Attention core: O = softmax(scale * Q K^T + kmask[b,k] + bias[b,q,k]) V   per (batch, head)
 *
 * Replaces P/model/Bert_backbone.py:247-290 (BertSelfAttention / RobertaSelfAttention) and the
 * core of nn.MultiheadAttention in the pano encoder (P/model/transformer.py:174-177).
 * Q/K/V/O are token-major: element (b, t, h, d) at base[b*sb + t*ld + h*D + d], so a fused
 * [tokens, 3*hidden] projection output is consumed in place and O lands already head-merged
 * (no permute/contiguous, :288-290).  kmask is the additive key mask (0 / -10000 of
 * P/model/ops.py:25-34, or -inf for the pano encoder's key_padding_mask); bias is the additive
 * [B,Nq,Nk] term (graph_sprels, :690-691).  lse[b,h,q] (log-sum-exp of the scaled scores) is
 * saved for backward.  drop_p/seed: attention-probability dropout (:280).
 * D must be 64.  F16/BF16 with 16-byte aligned bases and ld/sb multiples of 8 run on tcgen05 (TMA-staged Q/K/V tiles, S
 * and O accumulators in TMEM): the persistent pipelined kernels for Nq, Nk <= 128, 128-row query tiles x 128-key chunks
 * beyond (long instructions); everything else, and F32, runs the fp32-math SIMT kernels.
 * ------------------------------------------------------------------------------------------ */
typedef struct goat_attn_args {
  int B, heads, Nq, Nk, D;
  int dtype;
  const void* Q;
  const void* K;
  const void* V;
  int ldq, ldk, ldv;
  long long sbq, sbk, sbv;
  const float* kmask; /* [B, Nk] or NULL */
  const float* bias;  /* [B, Nq, Nk] or NULL */
  float scale;
  void* O;
  int ldo;
  long long sbo;
  float* lse; /* [B, heads, Nq] */
  float drop_p;
  uint64_t drop_seed;
  const uint64_t* drop_seed_ptr; /* optional DEVICE word added to drop_seed at run time */
  /* backward only */
  const void* dO; /* same layout as O */
  void* dQ;       /* same layout as Q (ldq, sbq) */
  void* dK;       /* same layout as K */
  void* dV;       /* same layout as V */
  float* dbias;   /* [B, Nq, Nk] fp32, ACCUMULATED into (caller zeroes), or NULL */
  int force_simt; /* 1: run the fp32-math SIMT kernels even when the tcgen05 path is eligible (cross-check) */
} goat_attn_args;
int goat_attn_core_fwd(const goat_attn_args* args, goat_stream_t stream);
int goat_attn_core_bwd(const goat_attn_args* args, goat_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (biased variance), fp32 statistics.
 * Replaces nn.LayerNorm / BertLayerNorm (P/model/Bert_backbone.py:303,309,363,369,805; eps is
 * per call site: 1e-12 or 1e-5, SURVEY.md 8a note 5).  The residual add that precedes it in the
 * reference (:309) is fused into the producing goat_gemm epilogue (res).
 * fwd: x [M,H] (x_dtype) -> y32 (fp32, optional) and/or y16 (y16_dtype, optional); mean/rstd [M].
 * bwd: dy [M,H] fp32 (+ optional dy2 fp32 added to it) -> dx32 = LN'(dy) (+ dres if given),
 *      dx16 (optional copy, dtype dx16_dtype = F16/BF16/F32, of the LN' term only, masked by
 *      dropout(drop_p, seed) for the GEMM that produced x), dgamma/dbeta [H] (written, not
 *      accumulated), dcolsum [H] optional = column sum of the dx16 values (the bias gradient of the
 *      producing Linear).
 *      workspace: goat_layernorm_bwd_workspace_bytes(M, H).
 * ------------------------------------------------------------------------------------------ */
int goat_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, float eps, float* y32,
                       void* y16, int y16_dtype, float* mean, float* rstd, int M, int H, goat_stream_t stream);
size_t goat_layernorm_bwd_workspace_bytes(int M, int H);
int goat_layernorm_bwd(const float* dy, const void* x, int x_dtype, const float* gamma, const float* mean,
                       const float* rstd, const float* dres, float* dx32, void* dx16, int dx16_dtype, float drop_p,
                       uint64_t drop_seed, const uint64_t* drop_seed_ptr, float* dgamma, float* dbeta, float* dcolsum,
                       void* workspace, int M, int H, goat_stream_t stream);
/* Same backward, but dgamma / dbeta / dcolsum are ACCUMULATED into with vector fp32 atomics by the one kernel (no
 * workspace, no finalize launch): pass zero-initialised vectors or flat-gradient views that already hold other
 * contributions.  H must be 768 and every tensor 16-byte aligned, else GOAT_ERR_UNSUPPORTED. */
int goat_layernorm_bwd_acc(const float* dy, const void* x, int x_dtype, const float* gamma, const float* mean,
                           const float* rstd, const float* dres, float* dx32, void* dx16, int dx16_dtype, float drop_p,
                           uint64_t drop_seed, const uint64_t* drop_seed_ptr, float* dgamma, float* dbeta,
                           float* dcolsum, int M, int H, goat_stream_t stream);

/* column sum: out[n] = sum_m x[m*ld + n]  (bias gradients).  workspace: goat_colsum_workspace_bytes(M,N) */
size_t goat_colsum_workspace_bytes(int M, int N);
int goat_colsum(const void* x, int dtype, int M, int N, int ld, float* out, void* workspace, goat_stream_t stream);
/* out[n] += sum_m x[m,n] in ONE kernel with fp32 atomics (out zero-initialised or holding an earlier partial sum, e.g. a
 * flat-gradient view the optimizer cleared).  N and ld multiples of 16/sizeof(dtype) elements, x 16-byte aligned.
 * Same reference sites as goat_colsum (bias gradients of every nn.Linear on the path). */
int goat_colsum_acc(const void* x, int dtype, int M, int N, int ld, float* out, goat_stream_t stream);

/* dtype conversion of n contiguous elements (fp32 master weights -> 16-bit operands, activations in/out) */
int goat_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, goat_stream_t stream);
/* same, with the dropout mask of element index i = 0..n-1 applied (regenerates the mask a goat_gemm epilogue
 * used on a contiguous [M,N] output: index m*N+n).  Backward of the pre-LN pano layers (P/model/transformer.py:170-182). */
int goat_dropout_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, float drop_p,
                      uint64_t drop_seed, const uint64_t* drop_seed_ptr, goat_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused optimizer step over ONE flat fp32 parameter buffer (the caller lays the model's parameters out
 * back to back; weight-decayed tensors first).  Replaces torch.nn.utils.clip_grad_norm_ +
 * the per-tensor AdamW loop (P/train_r2r_goat.py:349-366, P/optim/adamw.py:53-110; decay groups
 * P/optim/misc.py:12-22).
 *   goat_sumsq: partial[i] = sum of squares of a slice of g; *nparts_out (HOST int) = number of slices.
 *               partial needs goat_sumsq_workspace_bytes() bytes.
 *   goat_adamw_step: norm = sqrt(sum partial) * hp[8]; coef = hp[8] * min(1, hp[7] / (norm + 1e-6));
 *               g' = coef*g; m = b1 m + (1-b1) g'; v = b2 v + (1-b2) g'^2;
 *               p -= lr * sqrt(bc2)/bc1 * m / (sqrt(v) + eps); p -= lr*wd*p for the first n_decay elements;
 *               shadow (optional, F16/BF16) = p rounded -- the operand copy the tcgen05 GEMMs read.
 *               hp is a DEVICE array of 9 floats {lr, beta1, beta2, eps, weight_decay, 1-beta1^t, 1-beta2^t,
 *               max_grad_norm (<=0: off), grad pre-scale}; norm_out (optional, device) receives the norm.
 *               zero_grad=1 also clears g (optimizer.zero_grad(), P/train_r2r_goat.py:366) so that the next
 *               step's split-K weight-gradient GEMMs can accumulate into it with atomics.
 *   scaler (optional, DEVICE array of 5 floats {scale, clean steps, overflowed, skipped, steps taken}): the dynamic loss
 *               scale of the fp16 path -- torch.cuda.amp.GradScaler in the reference (P/train_r2r_goat.py:279,325,
 *               351-363).  Gradients are additionally divided by scaler[0]; a non-finite norm SKIPS the update (only
 *               g is cleared); bias corrections come from scaler[4] + 1.  goat_scaler_update (one thread) then
 *               applies GradScaler.update(): scale *= backoff after an overflow, *= growth after `interval` clean steps.
 *   shadow_lo (optional, same dtype as shadow): shadow_lo = round(p - shadow), the second term of the split weight
 *               operand (goat_gemm_args.B_lo).
 * ------------------------------------------------------------------------------------------ */
size_t goat_sumsq_workspace_bytes(void);
int goat_sumsq(const float* g, long long n, float* partial, int* nparts_out, goat_stream_t stream);
int goat_adamw_step(float* p, float* g, float* m, float* v, void* shadow, int shadow_dtype, long long n,
                    long long n_decay, const float* hp, const float* partial, int nparts, float* norm_out,
                    int zero_grad, const float* scaler, void* shadow_lo, goat_stream_t stream);
int goat_scaler_update(float* scaler, const float* partial, int nparts, float growth, float backoff, int interval,
                       goat_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The data-parallel gradient exchange FUSED with the optimizer step over NVLink peer memory (one process per GPU).
 * Replaces DDP's bucketed all-reduce + a full AdamW on every rank (P/utils/misc.py:52-58, P/optim/adamw.py:64-110,
 * P/train_r2r_goat.py:349-366).  Rank r owns elements [lo, lo+n) of the flat buffers.
 *   goat_peer_export: CUDA IPC handle (GOAT_PEER_HANDLE_BYTES bytes) of the device allocation `ptr` lies in, and
 *               ptr's byte offset inside it.  The allocation must be a plain cudaMalloc one (GOAT_ERR_CUDA otherwise).
 *   goat_peer_open / goat_peer_close: map / unmap another process's allocation -> its base address in this process.
 *               An allocation can be mapped once per process: callers cache by handle.
 *   goat_peer_reduce_sumsq: g_shard_out[i] = sum over ranks q (in rank order) of g_peers[q][lo + i], i < n -- peer loads
 *               through NVLink -- and partial[] = per-CTA sums of squares of the result (goat_sumsq's format).
 *               g_peers is a HOST array of `world` device pointers (own buffer included, at index rank).
 *               The caller must have synchronised the ranks (every rank's backward finished) before the launch.
 *   goat_adamw_step_peers: goat_adamw_step's arithmetic on the shard (g_shard indexed from 0; m, v and p indexed by flat
 *               element), and the new values are STORED TO EVERY RANK, 4 bytes per element: shadow + shadow_lo (16-bit
 *               operand copies, optional) for elements < n_fp32_from, the fp32 value for elements >= n_fp32_from (0: all
 *               of them; their 16-bit copies are re-derived by each rank with goat_split_cast after the barrier); the own
 *               rank's fp32 master is always written.  A non-finite norm with a scaler changes nothing.  The caller
 *               synchronises the ranks again before anybody reads the buffers, and clears its own gradient buffer.
 *   goat_peer_barrier / goat_peer_sum_scalar: rank synchronisation through flags in peer memory (one 32-thread CTA, no
 *               NCCL call).  Every rank owns a zero-initialised signal block of goat_peer_signal_bytes() bytes;
 *               signal_peers is the HOST array of all ranks' blocks.  `epoch` must be the same on every rank for a given
 *               call and increase by one per call (either function).  After the barrier every rank's work enqueued before
 *               ITS call (peer stores included) is visible to work enqueued after this one.  goat_peer_sum_scalar also
 *               exchanges one float: out[0] = sum over ranks, in rank order, of sum(partial[0..nparts)) -- the same
 *               bits on every rank (the global squared gradient norm).  out may alias partial.
 *               The kernels spin until every rank arrives: all ranks must make the same sequence of calls.
 * ------------------------------------------------------------------------------------------ */
#define GOAT_MAX_PEERS 8
#define GOAT_PEER_HANDLE_BYTES 64
/* hi = round(x) to F16/BF16, lo (optional) = round(x - hi): the split operand copies of an fp32 range (local). */
int goat_split_cast(const float* x, void* hi, void* lo, int dtype, long long n, goat_stream_t stream);
size_t goat_peer_signal_bytes(void);
int goat_peer_barrier(void* const* signal_peers, int world, int rank, unsigned int epoch, goat_stream_t stream);
int goat_peer_sum_scalar(void* const* signal_peers, int world, int rank, unsigned int epoch, const float* partial,
                         int nparts, float* out, goat_stream_t stream);
int goat_peer_export(const void* ptr, void* handle_out, unsigned long long* offset_out);
int goat_peer_open(const void* handle, void** base_out);
int goat_peer_close(void* base);
int goat_peer_reduce_sumsq(void* const* g_peers, int world, long long lo, long long n, float* g_shard_out,
                           float* partial, int* nparts_out, goat_stream_t stream);
int goat_adamw_step_peers(void* const* p_peers, void* const* shadow_peers, void* const* shadow_lo_peers,
                          int shadow_dtype, int world, int rank, const float* g_shard, float* m, float* v,
                          long long lo, long long n, long long n_decay, long long n_fp32_from, const float* hp,
                          const float* partial, int nparts, float* norm_out, const float* scaler, goat_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Attention pooling over the token axis of x [B,N,H] (fp32), softmax WITHOUT a mask (the reference pools over
 * padding too, SURVEY.md 8a note 6).
 *   mode 0  adaptive panorama fusion  P/model/vilmodel_goat.py:354-362, M/models/vilmodel_GOAT.py:727-733:
 *           s_n = tanh(x_n . w + bias[0]);  a = softmax_n(s);  out = sum_n a_n x_n
 *   mode 1  CFP pooling               P/model/pretrain_goat.py:502-515, M/models/vilmodel_GOAT.py:905-918:
 *           s_n = tanh(x_n) . w;            a = softmax_n(s);  out = tanh(sum_n a_n x_n)
 * fwd saves a [B,N] (and s [B,N] in mode 0).  bwd: dx [B,N,H] written; dw [H] and db [1] ACCUMULATED (caller zeroes).
 * n_valid (optional DEVICE int): pool over the first min(N, *n_valid) tokens only -- the reference pools over the
 * batch's own padded length, so buffers padded further (static shapes for CUDA graphs) must not add tokens.
 * ------------------------------------------------------------------------------------------ */
int goat_attn_pool_fwd(const float* x, const float* w, const float* bias, int mode, int B, int N, int H, float* out,
                       float* a, float* s, const int* n_valid, goat_stream_t stream);
int goat_attn_pool_bwd(const float* dout, const float* x, const float* w, const float* a, const float* s, const float* out,
                       int mode, int B, int N, int H, float* dx, float* dw, float* db, const int* n_valid,
                       goat_stream_t stream);

/* p(z)-weighted dictionary sum of the back-door adjustment: out[b,:] = sum_n p[b,n] x[b,n,:]
 * (M/models/vilmodel_GOAT.py:664-665 image z-dict, :107-111 instruction z-dicts).  p is data (no gradient). */
int goat_wsum_fwd(const float* x, const float* p, int B, int N, int H, float* out, goat_stream_t stream);
int goat_wsum_bwd(const float* dout, const float* p, int B, int N, int H, float* dx, goat_stream_t stream);

/* "door" gate of BACL-text / FACL (M/models/vilmodel_GOAT.py:145-148, :548-552):
 *   g = sigmoid(aug . wa + ba[0] + ori . wo + bo[0]);  out = g * aug + (1 - g) * ori      rows = tokens [M,H]
 * bwd: daug / dori written; dwa, dwo [H], dba, dbo [1] ACCUMULATED (caller zeroes). */
int goat_door_gate_fwd(const float* aug, const float* ori, const float* wa, const float* ba, const float* wo,
                       const float* bo, int M, int H, float* out, float* gate, goat_stream_t stream);
int goat_door_gate_bwd(const float* dout, const float* aug, const float* ori, const float* wa, const float* wo,
                       const float* gate, int M, int H, float* daug, float* dori, float* dwa, float* dwo, float* dba,
                       float* dbo, goat_stream_t stream);

/* Row-wise softmax cross-entropy, F.cross_entropy(reduction='none') of the SAP / MLM / CFP losses
 * (P/model/pretrain_goat.py:213-215, :348-350, :522-532; M/r2r/agent_base.py:133).  Element (i,j) of the logits is
 * logits[i*stride_row + j*stride_col], so the transposed InfoNCE term needs no copy.  -inf logits are legal (masked
 * actions).  labels int64; label == ignore_index gives loss 0 and no gradient.  lse [M] is saved for backward.
 * bwd: dlogits(i,j) = dloss_i (softmax_ij - [j == label_i]), written or (accumulate=1) added at dlogits strides. */
int goat_xent_fwd(const float* logits, long long stride_row, long long stride_col, const long long* labels, int M, int N,
                  long long ignore_index, float* loss, float* lse, goat_stream_t stream);
int goat_xent_bwd(const float* dloss, const float* logits, long long stride_row, long long stride_col,
                  const long long* labels, const float* lse, int M, int N, long long ignore_index, float* dlogits,
                  long long dstride_row, long long dstride_col, int accumulate, goat_stream_t stream);

/* The same cross-entropy over a class axis that is only ever materialised one column chunk at a time -- the MLM head's
 * tied 768 -> 50265 vocabulary projection (P/model/Bert_backbone.py:813-829, loss P/model/pretrain_goat.py:209-224), whose
 * [n, 50265] fp32 logits the reference writes to memory and reads back three times.
 *   fwd (per chunk [c0, c0+Nc) of the logits, `first` = 1 on the first chunk): running row maximum / sum of exponentials
 *       (run_max, run_sum [M]) and the label's logit (picked [M]) are updated; afterwards lse = run_max + log(run_sum) and
 *       loss = lse - picked (0 for label == ignore_index).
 *   bwd (per recomputed chunk): out[i,j] = dloss_i (exp(logit_ij - lse_i) - [c0 + j == label_i]) converted to out_dtype
 *       with leading dimension ldo -- directly the operand of the dgrad / wgrad GEMMs.  Labels outside [0, n_classes)
 *       other than ignore_index give no gradient. */
int goat_xent_chunk_fwd(const float* logits, long long ld, const long long* labels, int M, int Nc, long long c0, int first,
                        float* run_max, float* run_sum, float* picked, goat_stream_t stream);
int goat_xent_chunk_bwd(const float* dloss, const float* logits, long long ld, const long long* labels, const float* lse,
                        int M, int Nc, long long c0, long long ignore_index, long long n_classes, void* out, int out_dtype,
                        long long ldo, goat_stream_t stream);

/* Gather-and-reduce over index lists: out[r,:] = scale * sum_{k<K, idx[r,k]>=0} src[idx[r,k],:], scale = 1 (sum) or
 * 1/#valid (mean).  Replaces the host Python loops of global-map aggregation (P/model/vilmodel_goat.py:430-468: mean
 * of the candidate-view embeddings that observed an unvisited node) and of the SAP / navigation logit fusion
 * (P/model/pretrain_goat.py:328-345, M/models/vilmodel_GOAT.py:797-813) -- the index lists are built once on the
 * host from the viewpoint-id strings.  idx int32 [R,K], -1 = empty slot.  bwd ACCUMULATES into dsrc (caller zeroes). */
int goat_segment_reduce_fwd(const float* src, const int* idx, int R, int K, int H, int mean, float* out,
                            goat_stream_t stream);
int goat_segment_reduce_bwd(const float* dout, const int* idx, int R, int K, int H, int mean, float* dsrc,
                            goat_stream_t stream);

/* RoBERTa input embeddings before LayerNorm (P/model/Bert_backbone.py:87-116): out[m,:] = word[ids[m]] + pos[m % L] +
 * type[0]  (position ids are arange(L) from 0, token types all 0).  bwd ACCUMULATES into the three tables (any may be
 * NULL); rows equal to padding_idx (< 0: none) of the word and position tables receive no gradient, as with
 * nn.Embedding(padding_idx=...) (the fine-tune config inherits roberta's pad_token_id = 1). */
int goat_embed_fwd(const long long* ids, const float* word, const float* pos, const float* type, int M, int L, int H,
                   float* out, goat_stream_t stream);
int goat_embed_bwd(const float* dout, const long long* ids, int M, int L, int H, long long padding_idx, float* dword,
                   float* dpos, float* dtype, goat_stream_t stream);

/* GPU-resident feature bank: the reference reads the pre-extracted 36-view CLIP/ViT features of every
 * trajectory step from HDF5 on the host and copies [S,36,768] fp32 to the device per batch (P/data/dataset.py:811-818,
 * P/data/loader.py:78-87).  Here the whole table lives in HBM once, in 16 bit, and a batch carries one int per step:
 * out[r, :] = src[idx[r], :] over rows of row_bytes (a multiple of 16) bytes; idx < 0 gives a zero row (padded steps). */
int goat_gather_rows(const void* src, const int* idx, int R, long long row_bytes, void* out, goat_stream_t stream);

/* Backward of the activation behind a head / pooler nn.Linear (ClsPrediction ReLU, BertPooler tanh, the GELU of the
 * MLM / CFP transforms: P/model/pretrain_goat.py:27-38, P/model/Bert_backbone.py:783-811):
 *   out[i] = dy[i] * act'(ref[i]) converted to out_dtype, one pass.  ref = the forward OUTPUT y (fp32) for RELU / TANH
 *   (y > 0, 1 - y^2) and the stored pre-activation (ref_dtype) for GELU.  act: GOAT_ACT_RELU / _TANH / _GELU / _NONE. */
int goat_act_grad(const float* dy, const void* ref, int ref_dtype, int act, void* out, int out_dtype, long long n,
                  goat_stream_t stream);
/* out[i] = act(x[i]) in fp32 (exact erf GELU): the head transforms keep their fp32 pre-activation -- applying GELU in a
 * 16-bit GEMM epilogue would round the pre-activation to 16 bits first, which costs these fp32-out heads ~1e-3. */
int goat_act_fwd(const float* x, int act, float* out, long long n, goat_stream_t stream);

/* Spatial-relation bias of the global-map self-attention: sprel_linear = nn.Linear(1, 1) applied to every pairwise
 * distance (P/model/vilmodel_goat.py:499-501, M/models/vilmodel_GOAT.py:473-476): out[i] = d[i] * w[0] + b[0].
 * bwd: dw[0] += sum dout[i] d[i], db[0] += sum dout[i]  (ACCUMULATED; the distances are data). */
int goat_sprel_fwd(const float* d, const float* w, const float* b, float* out, long long n, goat_stream_t stream);
int goat_sprel_bwd(const float* dout, const float* d, float* dw, float* db, long long n, goat_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GOAT_SM100_H_ */
