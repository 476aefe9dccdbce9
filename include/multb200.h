/* multb200.h -- C ABI of libmultb200.so: hand-written sm_100a kernels for the dynamic
 * MulT transformer stacks (weight-sliced multi-head attention, dynamic Linear /
 * LayerNorm, sinusoidal position embedding) of duyubo/Multimodal-Transformer-Robustness.
 *
 * The reference has no FFI: its boundary for this path is the Python class API of the
 * package `modules`.  These entry points are what that API's
 * forward/backward bind to; each one cites the reference code it replaces (paths
 * relative to the reference root).  Rules of the boundary:
 *   - plain pointers + sizes only (no torch types); every pointer is a DEVICE pointer
 *     to fp32 data unless stated otherwise; `stream` is a cudaStream_t passed as void*.
 *   - every op is GROUPED: it takes an array of `n` problem descriptors (host memory,
 *     n <= MTB_MAX_GROUP) and runs all of them in ONE kernel launch -- this is how all
 *     active fusion branches of a stage run concurrently (north-star item (d)).
 *   - the library never owns tensor memory and keeps no mutable global state apart from
 *     a per-thread last-error string.
 *   - return value 0 = success; negative = error (message via mtb_last_error()).
 *   - activations are row-major [tokens, features] with an explicit leading dimension;
 *     a seq-first [L, B, E] tensor is the matrix [L*B, E] (token t = l*B + b).
 */
#ifndef MULTB200_H
#define MULTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTB_ABI_VERSION 15
#define MTB_MAX_GROUP 24

/* Dropout RNG: Philox4x32-7.  Element `i` of a dropout site is kept iff
 * philox(key = seed, counter = offset + i/4)[i % 4] >= (uint32)(p * 2^32).
 * If `dev` is non-null it points to two device uint64 {seed_add, offset_add} that are
 * added to seed/offset at run time (lets a captured CUDA graph draw fresh masks on every
 * replay).  Replaces F.dropout at modules/dynamic_transformer.py:68,77-78,173,182,185 and
 * modules/dynamic_multihead_attention.py:110. */
typedef struct {
  uint64_t seed;
  uint64_t offset;
  const uint64_t* dev;
} mtb_rng;

/* ---- library ---------------------------------------------------------------------- */
int mtb_abi_version(void);
const char* mtb_last_error(void);
int mtb_sm_count(void);
/* which GEMM engine mtb_linear_* uses: 0 = fp32 CUDA-core (parity mode, 1e-5),
 * 1 = tcgen05 TF32 tensor-core (TMA + TMEM), 2 = the same engine for callers that keep activations in bf16: problems
 * flagged in_bf16 / out_bf16 run tcgen05.mma.kind::f16 on bf16 operands, unflagged ones run as in mode 1.
 * Returns the previous mode. */
int mtb_set_gemm_mode(int mode);
int mtb_get_gemm_mode(void);
/* which attention core mtb_attn_* uses: 0 = fp32 CUDA-core flash kernels, 1 = tcgen05 / TMEM flash
 * kernels (TF32 QK^T, PV, dP, dQ, dK, dV on the tensor core), -1 (default) = follow the GEMM engine. */
int mtb_set_attn_mode(int mode);
int mtb_get_attn_mode(void);
/* force-load every kernel of the library into the current CUDA context (CUDA loads modules lazily;
 * without this the first use of each kernel variant stalls a training step by milliseconds) */
int mtb_preload(void);
/* number of kernels this library has launched so far in this process (bench bookkeeping) */
uint64_t mtb_launch_count(void);

/* materialise the keep-mask (1 byte per element, 1 = keep) a kernel would use for a
 * dropout site; test support for injecting the same masks into the CPU oracle. */
int mtb_dropout_mask(mtb_rng rng, float p, int64_t n, uint8_t* keep, void* stream);
/* rng_dev[1] += delta  (advance the per-replay offset of a captured graph) */
int mtb_rng_advance(uint64_t* rng_dev, uint64_t delta, void* stream);

/* ---- (a1,a2) embed: y = scale * x + PE(pos(x[...,0])), then dropout ------------------
 * modules/dynamic_transformer.py:64-68,72-78 + modules/position_embedding.py:8-27,45-83.
 * x is a strided [L, B, E] view (element strides sl, sb, se); y is contiguous [L*B, E].
 * pos = l+1 where x[l,b,0] != 0 else 0 (zero positional vector); PE(p, c) =
 * sin(p*w) for even c, cos(p*w) for odd c, w = exp(-(c/2) * ln(1e4) / (E/2 - 1)).
 * The sinusoid is evaluated in-register (no table, no host copy). */
typedef struct {
  const float* x; int64_t sl, sb, se;
  float* y;
  int L, B, E;
  float scale; float p; mtb_rng rng;
} mtb_embed_desc;
int mtb_embed_fwd(const mtb_embed_desc* d, int n, void* stream);
/* dx[L*B, E] (contiguous) = scale * keep/(1-p) * dy ; d->x = dy, d->y = dx */
int mtb_embed_bwd(const mtb_embed_desc* d, int n, void* stream);

/* ---- strided sum / copy: dst[T,E] = (accumulate ? dst : 0) + sum_i src_i[T,E], n_src in 1..3 ------
 * The fusion DAG's glue (src/dynamic_models2.py:242 torch.cat, :257 h[-1], and the gradient fan-in
 * of a branch output consumed by several later branches) without materialising concatenations. */
typedef struct {
  const float* src[3]; int64_t ld_src[3]; int n_src;
  float* dst; int64_t ld_dst;
  int T, E; int accumulate;
  int src_bf16[3]; int dst_bf16;   /* bf16 data path: that operand is bfloat16 (sums are formed in fp32; doubles as the cast op) */
} mtb_addn_desc;
int mtb_addn(const mtb_addn_desc* d, int n, void* stream);

/* ---- (a3,a10) dropout + residual + LayerNorm ------------------------------------------
 * modules/dynamic_transformer.py:163,169-170,173-178,185-187,87 + modules/dynamic_layers.py:61-67.
 *   x_new = res + dropout_p(a)      (a == NULL: x_new = res, nothing written to x_new)
 *   y     = LayerNorm(x_new) * gamma[idx] + beta[idx]    (gamma == NULL: no LayerNorm)
 * idx (int32[E], may be NULL) gathers the affine parameters (active_mask).
 * mean/rstd ([T] each, may be NULL in inference) are saved for backward. */
typedef struct {
  const float* res; int64_t ld_res;
  const float* a;   int64_t ld_a;
  float* x_new;     int64_t ld_x;
  float* y;         int64_t ld_y;
  const float* gamma; const float* beta; const int32_t* idx;
  float* mean; float* rstd;
  int T, E;
  float eps; float p; mtb_rng rng;
  int a_bf16, y_bf16;   /* bf16 data path: `a` (a GEMM output) / `y` (a GEMM input) are bfloat16; the residual stream
                           (res, x_new) and the statistics always stay fp32, like autocast's LayerNorm */
} mtb_resln_desc;
int mtb_resln_fwd(const mtb_resln_desc* d, int n, void* stream);

/* backward of the above.
 *   g    = d_xnew (may be NULL) + LayerNorm_backward(dy; x_new, mean, rstd, gamma[idx])
 *   d_res = g ;  d_a = g * keep/(1-p)   (d_a may be NULL)
 *   dgamma/dbeta (may be NULL: masked LayerNorm gets no gradient) are
 *   ACCUMULATED (+=) at [idx]; so is the optional dbias (column sums of d_a). */
typedef struct {
  const float* dy;     int64_t ld_dy;
  const float* d_xnew; int64_t ld_dx;
  const float* x_new;  int64_t ld_x;
  const float* mean; const float* rstd;
  const float* gamma; const int32_t* idx;
  float* d_res; int64_t ld_dres;
  float* d_a;   int64_t ld_da;
  float* dgamma; float* dbeta;
  int T, E;
  float p; mtb_rng rng;
  float* dbias;    /* optional: dbias[idx[c]] += sum_t d_a[t, c] -- the bias gradient of the linear layer that
                      produced `a` (out-projection / fc2), fused here so no separate column-sum pass is needed */
  int dy_bf16, da_bf16;   /* bf16 data path: dy (a dgrad GEMM output) / d_a (a GEMM input) are bfloat16 */
} mtb_resln_bwd_desc;
int mtb_resln_bwd(const mtb_resln_bwd_desc* d, int n, void* stream);

/* ---- (a6,a7,a9) sliced / gathered linear ----------------------------------------------
 * modules/dynamic_multihead_attention.py:259-282 (_in_proj/_out_proj) and
 * modules/dynamic_layers.py:15-25 (DynamicLinear):
 *   Y[M,N] = act( X[M,K] . W'^T + b' ),  W'[n,k] = W[row(n)*ldw + col(k)],  b'[n] = b[row(n)]
 * row(n) = row_idx ? row_idx[n] : n ; col(k) = col_idx ? col_idx[k] : k  (int32 device
 * arrays).  Head/dim prefix slicing of the [3,H,hd,E] in-projection and the column
 * slicing of the out-projection are expressed as index arrays; plain prefix slices just
 * use a smaller N/K with the full ldw.
 * act: 0 = none, 1 = ReLU followed by dropout(p) (modules/dynamic_transformer.py:181-182). */
/* Optional block structure of an index array: idx == concat_s [seg[s]*len, (seg[s]+1)*len), s < n.
 * n == 0 means "unknown / not block structured".  The reference's active_mask gathers are unions
 * of d-wide blocks (src/dynamic_models2.py:243-251); when the host states that structure the
 * tensor-core engine addresses the blocks through TMA instead of falling back to the fp32 engine. */
#define MTB_MAX_SEGS 16
typedef struct { int32_t len; int32_t n; int32_t seg[MTB_MAX_SEGS]; } mtb_segs;

typedef struct {
  const float* X; int64_t ldx;
  const float* W; int64_t ldw;
  const float* bias;
  const int32_t* row_idx; const int32_t* col_idx;
  float* Y; int64_t ldy;
  int M, N, K;
  int act; float p; mtb_rng rng;
  mtb_segs row_segs, col_segs;
  /* bf16 data path (tensor-core engine only, gemm mode 2): in_bf16 -- X and W point to bfloat16 data (W: the bf16 shadow
   * of the fp32 master weight, same shape and leading dimension in ELEMENTS); out_bf16 -- Y is bfloat16.  Accumulation,
   * bias, ReLU and dropout stay fp32; the result is rounded to nearest even once.  bias is always fp32. */
  int in_bf16, out_bf16;
} mtb_linear_desc;
int mtb_linear_fwd(const mtb_linear_desc* d, int n, void* stream);

/* backward.  dY' = dY                         (act == 0)
 *            dY' = dY * [Yact > 0] / (1-p)    (act == 1; Yact is the forward output)
 *   dX[M,K]  (+)= dY' . W'            (dX may be NULL; accumulate_dx: += instead of =)
 *   dW[row(n)*ldw + col(k)] += dY'^T . X   and   db[row(n)] += colsum(dY')
 *   (dW/db may be NULL; they are always ACCUMULATED into full-size, caller-zeroed grads so
 *    rows/cols outside the active slice keep an explicit zero gradient). */
typedef struct {
  const float* dY; int64_t ldy;
  const float* Yact; int64_t ldyact;
  const float* X; int64_t ldx;
  const float* W; int64_t ldw;
  const int32_t* row_idx; const int32_t* col_idx;
  float* dX; int64_t lddx; int accumulate_dx;
  float* dW; float* db;
  int M, N, K;
  int act; float p;
  float* scratch;   /* [M*N] floats, required by the tensor-core engine when act == 1 (holds dY') */
  mtb_segs row_segs, col_segs;
  /* bf16 data path: in_bf16 -- dY, Yact, X, W and scratch are bfloat16; dx_bf16 -- dX is bfloat16.  dW / db are ALWAYS
   * fp32 (accumulated into the fp32 gradient arena, like autocast's weight gradients after the cast back). */
  int in_bf16, dx_bf16;
} mtb_linear_bwd_desc;
int mtb_linear_bwd(const mtb_linear_bwd_desc* d, int n, void* stream);

/* ---- (a4,a5) fused attention core -----------------------------------------------------
 * modules/dynamic_multihead_attention.py:91-116 + modules/transformer.py:145-157:
 *   S = scale * q k^T ; S[i,j] = -inf where j - i >= 1 + |Lk - Lq| ; P = softmax_fp32(S) ;
 *   P~ = dropout_p(P) ; o = P~ v.      Scores are never written to HBM.
 * q/k/v/o are token-major: row (l*B + b) with leading dimension ld*, head h occupies
 * columns [h*hd, (h+1)*hd) -- i.e. the GEMM output layout, so no transposes exist.
 * lse[(b*H + h)*Lq + i] = log-sum-exp of row i (saved for backward).
 * dropout element index = ((b*H + h)*Lq + i)*Lk + j. */
typedef struct {
  const float* q; int64_t ldq;
  const float* k; int64_t ldk;
  const float* v; int64_t ldv;
  float* o; int64_t ldo;
  float* lse;
  int Lq, Lk, B, H, hd;
  float scale; float p; mtb_rng rng;
  /* optional [B*H*Lq, ceil(Lk/32)] words: when p > 0 the forward kernel stores the dropout keep bits of every score it
   * computed (bit j%32 of word j/32 of row (b*H+h)*Lq + i); the backward kernels read them instead of re-drawing the
   * Philox stream twice.  NULL: not stored / re-drawn.  Tensor-core engine only (the fp32 engine ignores it). */
  uint32_t* keep_bits;
  int bf16;   /* bf16 data path (tensor-core engine only), bit mask: 1 = q, k, v are bfloat16; 2 = o is bfloat16.
                 Scores, softmax and lse stay fp32.  (The plan executor keeps q / k / v in fp32 -- their 25-element head
                 rows fall below cp.async's 4-byte granularity at 2 bytes per element -- and writes o in bf16.) */
} mtb_attn_desc;
int mtb_attn_fwd(const mtb_attn_desc* d, int n, void* stream);

typedef struct {
  const float* q; int64_t ldq;
  const float* k; int64_t ldk;
  const float* v; int64_t ldv;
  const float* o; int64_t ldo;
  const float* d_o; int64_t lddo;
  const float* lse;
  float* delta;            /* scratch [B*H*Lq] */
  float* dq; int64_t lddq;
  float* dk; int64_t lddk;
  float* dv; int64_t lddv;
  int Lq, Lk, B, H, hd;
  float scale; float p; mtb_rng rng;
  const uint32_t* keep_bits;   /* optional: the words the forward kernel stored (see mtb_attn_desc) */
  int bf16;   /* bf16 data path, bit mask: 1 = q, k, v; 2 = o; 4 = d_o; 8 = dq, dk, dv are bfloat16.  lse / delta stay fp32 */
} mtb_attn_bwd_desc;
int mtb_attn_bwd(const mtb_attn_bwd_desc* d, int n, void* stream);

/* ---- plan executor: a whole list of grouped launches in one call -----------------------
 * North-star item (d) / SURVEY.md 8b `mtb_stage_launch`: the host plan executor (mtb200/engine.py) compiles one
 * sampled configuration into flat lists of grouped launches; this entry point issues such a list natively, so the
 * per-launch cost is one driver call instead of one interpreter round trip.  `kind` selects the entry point the
 * descriptor array belongs to; ops flagged `side` (deferred weight gradients, which feed nothing downstream) are
 * issued on `side_stream`, ordered after everything issued before them on `stream`; `stream` re-joins `side_stream`
 * before the call returns.  side_stream == NULL runs everything on `stream`.  Stops at the first failing op.
 * The fork / join events are created once per process on the current device: one device per process (the model this
 * library is built for: one rank per GPU), calls from a single host thread. */
enum { MTB_OP_EMBED_FWD = 0, MTB_OP_EMBED_BWD = 1, MTB_OP_ADDN = 2, MTB_OP_RESLN_FWD = 3, MTB_OP_RESLN_BWD = 4,
       MTB_OP_LINEAR_FWD = 5, MTB_OP_LINEAR_BWD = 6, MTB_OP_ATTN_FWD = 7, MTB_OP_ATTN_BWD = 8 };
typedef struct {
  int32_t kind;          /* MTB_OP_* */
  int32_t n;             /* descriptors in `descs` (<= MTB_MAX_GROUP) */
  const void* descs;     /* host array of the matching descriptor type */
  int32_t side;          /* 1: may run on the side stream */
  int32_t reserved;
} mtb_op;
int mtb_run_ops(const mtb_op* ops, int n_ops, void* stream, void* side_stream);
/* The same op list captured into a CUDA graph (side ops become a parallel branch when use_side != 0) and replayed with
 * one launch.  Valid as long as every address in the descriptors stays valid -- the plan executor's stage batches only
 * reference persistent regions, parameters and the device-side dropout counter, so a replay draws fresh dropout masks.
 * The descriptor arrays themselves are copied into the graph (kernel parameters) and may be freed after capture.
 * mtb_graph_capture returns 0 and a handle, or non-zero when this driver cannot capture the list (run it eagerly). */
int mtb_graph_capture(const mtb_op* ops, int n_ops, int use_side, void** handle);
int mtb_graph_launch(void* handle, void* stream);
int mtb_graph_destroy(void* handle);

/* ---- fused gradient clip + Adam over the flat arenas --------------------------------
 * Replaces `torch.nn.utils.clip_grad_norm_(model.parameters(), clip); optimizer.step()` of
 * src/train.py:181-182 (optimizer = torch.optim.Adam, src/train.py:51) for parameters whose
 * gradients live in one flat fp32 arena.  Static tables (built once per model) cut every
 * parameter into chunks: chunk c covers `chunk_n[c]` elements of parameter `chunk_pid[c]`
 * starting at `chunk_param[c]` (parameter storage) and at arena offset `chunk_off[c]` (same
 * offset in grad / exp_avg / exp_avg_sq).  `active[pid] != 0` marks the parameters that
 * received a gradient this step (torch skips `.grad is None`); only those are clipped and
 * updated and only their `steps[pid]` counters advance (per-parameter bias correction, as
 * torch keeps it).  scalars[0] = total gradient norm (what clip_grad_norm_ returns),
 * scalars[1] = clip coefficient applied.  max_norm <= 0 disables clipping.  Three launches. */
typedef struct {
  float* const* chunk_param;     /* [n_chunks] device array of device pointers */
  const int64_t* chunk_off;      /* [n_chunks] */
  const int32_t* chunk_n;        /* [n_chunks] */
  const int32_t* chunk_pid;      /* [n_chunks] */
  const uint8_t* active;         /* [n_params] */
  int32_t* steps;                /* [n_params] */
  float* grad; float* exp_avg; float* exp_avg_sq;   /* flat arenas */
  float* partial;                /* scratch [n_chunks] */
  float* scalars;                /* out [2] */
  int n_chunks, n_params;
  float lr, beta1, beta2, eps, weight_decay, max_norm;
  uint16_t* shadow;              /* optional bf16 shadow arena (same offsets as grad): shadow[off + i] = bf16(param[i]) is
                                    rewritten for every updated element, so the bf16 data path needs no separate cast pass */
} mtb_adam_desc;
int mtb_adam_step(const mtb_adam_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MULTB200_H */
