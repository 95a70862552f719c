/*
 * libshgvqa — C ABI of the B200-native (sm_100a) hot path of SHG-VQA.
 *
 * The reference (aurooj/SHG-VQA) has no FFI layer: its hot path is reached through
 * Python nn.Module symbols.  Each entry point below names the
 * reference computation it replaces (paths relative to AGQA/src/).  Conventions:
 *   - plain pointers and sizes; all pointers are DEVICE pointers unless named host_*;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), does not
 *     allocate, and is CUDA-graph capturable;
 *   - returns 0 on success or a negative SHG_E* code; shg_last_error() gives the text;
 *   - matrices are row-major; activations bf16 (uint16_t storage), master weights and
 *     gradients fp32, statistics fp32, LSAP arithmetic fp64 (as scipy's).
 */
#ifndef SHGVQA_H
#define SHGVQA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHG_OK 0
#define SHG_EINVAL (-1)      /* bad argument                                  */
#define SHG_ESHAPE (-2)      /* unsupported shape / alignment                 */
#define SHG_ECUDA (-3)       /* CUDA runtime / driver error                   */
#define SHG_EDEVICE (-4)     /* not an sm_100 device                          */

const char* shg_last_error(void);
int shg_version(void);
/* Refuses anything that is not compute capability 10.x (there is no fallback path). */
int shg_device_check(int device);
/* ------------------------------------------------------------------ Hungarian matching
 * Replaces the CPU loop `[linear_sum_assignment(c[i]) for i, c in enumerate(C.split(sizes, -1))]`
 * (lxrt/matcher.py:76-80; scipy's rectangular_lsap, tie-break included).  One CTA per frame.
 *
 * cost      fp32 [n_frames, q_f, ld]: frame f owns rows [f*q_f,(f+1)*q_f) and the column range
 *           [col_off[f], col_off[f+1]) — the layout of `C.view(bs*clip_len, q, -1)` (matcher.py:76).
 * rows/cols int32 [n_frames, q_f]: matched (query, target) pairs, query ascending, -1 padded.
 * n_match   int32 [n_frames]; status int32 [n_frames]: 0 ok, -1 infeasible, -2 NaN/-inf entry
 *           (scipy raises ValueError for both), -3 more than 32 targets in the frame, -4 (fused matcher only) a label
 *           outside [0, n_cls) — an IndexError in the reference (lxrt/matcher.py:73).
 * q_f <= 32 and every frame's target count <= 32.
 */
int shg_lsap_batched(const float* cost, int n_frames, int q_f, int ld, const int32_t* col_off,
                     int32_t* rows, int32_t* cols, int32_t* n_match, int32_t* status, void* stream);

/* Fused matcher front half + solver: softmax over classes, gather of the target columns, cost = -p,
 * LSAP (lxrt/matcher.py:59-80) and, optionally, the scatter of matched labels into `target_classes`
 * (tasks/agqaHGQA.py:216-221; background class 0 elsewhere).  logits fp32 or bf16 [n_frames*q_f, n_cls]
 * (ld_logits elements per row); labels int32, CSR by frame through label_off[n_frames+1].
 * row_max / row_lse fp32 and row_argmax int32 [n_frames*q_f]: softmax statistics reused by the set loss
 * (row_argmax may be NULL).  target_classes int32 [n_frames*q_f] or NULL. */
int shg_match_from_logits(const void* logits, int logits_bf16, int n_frames, int q_f, int n_cls, int ld_logits,
                          const int32_t* labels, const int32_t* label_off, int32_t* rows, int32_t* cols,
                          int32_t* n_match, int32_t* status, float* row_max, float* row_lse, int32_t* row_argmax,
                          int32_t* target_classes, void* stream);

/* Weighted cross-entropy over the matched targets and its gradient (tasks/agqaHGQA.py:223:
 * F.cross_entropy(weight=empty_weight) = sum w_t nll / sum w_t with w[0] = eos_coef, else 1) plus the
 * counts behind class_error (tasks/agqaHGQA.py:228).  fwd ADDS into out[0]=sum w*nll, out[1]=sum w,
 * out[2]=#matched, out[3]=#matched whose argmax is the label (caller zeroes `out`); bwd writes
 * dlogits = grad_scale * w_t (softmax - onehot) / out[1] as bf16 or fp32.
 * Error reporting without a synchronisation: `status` (int32 [n_frames], as written by shg_match_from_logits; may be
 * NULL) and target classes outside [0, n_cls) are COUNTED into *err_out (fp32 device scalar, may be NULL): the
 * reference raises for those inputs (scipy ValueError on NaN/-inf costs, IndexError on a label >= n_cls); the host
 * turns a non-zero count into an exception when it reads the loss sums.  Such rows add nothing to the loss or the
 * gradient. */
int shg_set_loss_fwd(const void* logits, int logits_bf16, int n_rows, int n_cls, int ld_logits,
                     const int32_t* target_classes, const float* row_lse, const int32_t* row_argmax, float eos_coef,
                     float* out, const int32_t* status, int n_frames, float* err_out, void* stream);
int shg_set_loss_bwd(const void* logits, int logits_bf16, int n_rows, int n_cls, int ld_logits,
                     const int32_t* target_classes, const float* row_lse, float eos_coef, const float* out,
                     float grad_scale, void* dlogits, int dlogits_bf16, int ld_dlogits, void* stream);

/* ------------------------------------------------------------------ dense projections (tcgen05 + TMA)
 * D[M,N] = epilogue( A . B^T ) with bf16 operands staged by TMA (128B swizzle) and fp32 accumulation in TMEM.
 * Replaces every nn.Linear on the path (lxrt/modeling_capsbert.py:373-375,427,466,481;
 * lxrt/transformer.py:189-193 incl. packed in_proj; tasks/agqa_model.py:105-140) and their backward.
 *
 *   a_mn_major = 0: A is [M,K] row-major (lda elements per row).  1: A is stored [K,M] row-major (A^T view).
 *   b_mn_major = 0: B is [N,K] row-major (nn.Linear weight).      1: B is stored [K,N] row-major.
 *   forward        : A = x [M,K],  B = W [N,K]                      (0,0)
 *   dgrad          : A = dy [M,N'], B = W stored [N',K'] as [K,N]    (0,1)  -> dx = dy . W
 *   wgrad          : A = dy stored [M',N] as [K,M], B = x stored [M',K'] as [K,N]  (1,1) -> dW = dy^T . x
 *
 * Epilogue, in this order: + bias[N] (fp32, may be NULL) ; optional store of the pre-activation (bf16, `preact`,
 * same ld as d) ; activation ; x act'(mul_aux) (backward of GELU / ReLU) ; + residual (bf16 [M,N], ldr) ; store as bf16 or fp32 — or, with
 * SHG_GEMM_ACCUMULATE, atomically add fp32 into d (split-K wgrad: `splits` CTAs share one output tile).
 */
#define SHG_ACT_NONE 0
#define SHG_ACT_GELU_ERF 1   /* lxrt/modeling_capsbert.py:127-133 */
#define SHG_ACT_RELU 2       /* lxrt/transformer.py:229 */
#define SHG_ACT_TANH 3       /* BertPooler2, lxrt/modeling_capsbert.py:1531 */
#define SHG_MUL_GELU_GRAD 1
#define SHG_MUL_RELU_MASK 2
#define SHG_GEMM_OUT_F32 1
#define SHG_GEMM_ACCUMULATE 2
typedef struct {
    const void* a; int64_t lda; int a_mn_major;
    const void* b; int64_t ldb; int b_mn_major;
    void* d; int64_t ldd;
    int m, n, k;
    const float* bias;
    const void* residual; int64_t ldr;
    void* preact;
    int act;
    const void* mul_aux;   /* bf16 [M,N], ld = ldd: backward through an activation, applied after act/bias */
    int mul_mode;          /* SHG_MUL_GELU_GRAD: d *= gelu'(mul_aux) ; SHG_MUL_RELU_MASK: d = mul_aux > 0 ? d*mul_scale : 0 */
    float mul_scale;       /* 0 = 1; 1/(1-p) when mul_aux is a ReLU output that also went through dropout */
    float dropout_p;       /* training-mode nn.Dropout on the projection output (after act, before residual) */
    uint32_t dropout_salt; /* call-site id; the mask is a pure function of (*dropout_seed, salt, row*N + col) */
    const uint32_t* dropout_seed;   /* device scalar; NULL = dropout off */
    int flags;
    int splits;        /* 0 = choose (split-K only with SHG_GEMM_ACCUMULATE); > 1 requires SHG_GEMM_ACCUMULATE */
    int block_n;       /* 0 = choose; else 64|128|192|256, optionally | 1024 to force CTA-pair (256-row) tiles
                          (128/256 only) or | 2048 to forbid them */
    /* grouped product (the two DETR decoders in lock-step, lxrt/transformer.py:95-124 run twice in
     * tasks/agqa_model.py:220-249): rows [0, split_m) of A use b / bias, rows [split_m, m) use b2 / bias2 — two
     * nn.Linear layers of the same shape in one launch.  b2 NULL = ordinary product.  split_m % 128 == 0. */
    const void* b2;
    const float* bias2;
    int split_m;
    /* grouped split-K accumulation (the same two layers' weight gradients in one launch): with SHG_GEMM_ACCUMULATE and
     * splits = 0, K rows [0, k_split) accumulate into d and rows [k_split, k) into d2 (same shape / ld as d).
     * d2 NULL = ordinary accumulation.  k_split % 64 == 0. */
    void* d2;
    int k_split;
    /* optional scratch for the tail-split schedule of CTA-pair tiles (the tiles of a last, at most quarter-full wave
     * are cut along K over the idle SMs; partial sums meet here): caller-owned device memory, 16-byte aligned, ALL ZERO
     * on entry and left all zero on return, used by one launch at a time (one buffer per stream).  NULL / too small =
     * plain schedule.  18 MB covers every shape (37 tiles x 2 x 128 x 256 fp32 + 1 KB of counters); the bs32 step
     * needs 2.1 MB. */
    void* workspace;
    int64_t workspace_bytes;
    /* 0 = one CTA per SM (148).  > 0 caps the persistent grid at this many CTAs (rounded down to an even count for
     * CTA-pair tiles; the tail-split schedule is not used then).  The data-parallel step sets 148 - R while a
     * collective that owns R SMs is in flight: a 148-CTA persistent grid would leave R of its CTAs waiting for an SM
     * until the others have finished their share (two waves for one wave of work). */
    int max_ctas;
} shg_gemm_args;
int shg_gemm(const shg_gemm_args* args, void* stream);
/* Debug aid: CTA 0 of subsequent shg_gemm launches writes 7 %globaltimer phase stamps to buf (8 x u64 device
 * memory); NULL = off. */
int shg_debug_gemm_stamps(void* buf);

/* ------------------------------------------------------------------ attention (tcgen05, head dim 64)
 * O = softmax(scale * Q K^T + mask) V per (batch, head).  Replaces BertAttention's matmul/softmax/matmul
 * (lxrt/modeling_capsbert.py:394-417) and nn.MultiheadAttention inside the DETR decoders
 * (lxrt/transformer.py:220-227).  q/k/v/o are bf16 row-major [batch*len, ld] with head h in columns
 * [64h, 64h+64) counted from the given pointer (so packed QKV / hoisted K|V buffers are read in place).
 * mask_mode 0: none; 1: additive fp32 key mask [batch, lk] (BERT's (1-mask)*-10000, :1834);
 * 2: block-causal by frame — key j visible to query i iff j / k_per_frame <= i / q_per_frame
 * (lxrt/entry.py:114-121).  lse (fp32 [batch, heads, lq], may be NULL) saves log-sum-exp for the backward.
 * lk <= 416 (all keys of a head are resident on chip; no online softmax). */
typedef struct {
    const void* q; int64_t ldq;
    const void* k; int64_t ldk;
    const void* v; int64_t ldv;
    void* o; int64_t ldo;
    float* lse;
    int batch, heads, lq, lk;
    float scale;
    int mask_mode;
    const float* key_mask;
    int q_per_frame, k_per_frame;
    float* o_f32; int64_t ldof;   /* optional fp32 copy of o [batch*lq, ldof] for the backward's delta (may be NULL) */
    /* training-mode dropout on the attention probabilities (lxrt/modeling_capsbert.py:381,402 p = 0.1;
     * nn.MultiheadAttention(dropout=0.15), lxrt/transformer.py:192-193): o = (dropout(softmax(s)) v); the mask is a
     * pure function of (*dropout_seed, dropout_salt, batch, head, query, key).  dropout_seed NULL = off. */
    float dropout_p; uint32_t dropout_salt; const uint32_t* dropout_seed;
    uint32_t* dropout_bits;   /* optional out: the keep mask as bit words [batch, heads, lq, ceil(lk/32)] (bit k%32 of word
                                 k/32 = key k kept), so that the backward reads 1 bit instead of re-hashing per element */
} shg_attn_args;
int shg_attn_fwd(const shg_attn_args* args, void* stream);
/* Key length from which shg_attn_fwd uses the chunked kernel (online softmax over 128-key chunks, two CTAs per SM)
 * instead of the whole-row kernel; default 257 (the 393-token visual memory), also SHG_ATTN_FWD2_MIN_LK.  Both
 * kernels give the same results within the tolerance stated in tests/test_gpu_kernels.py. */
int shg_set_attn_fwd_chunked_min_lk(int min_lk);

/* Backward of the above (autograd of lxrt/modeling_capsbert.py:394-417 / nn.MultiheadAttention).  Recomputes
 * the probabilities from q, k and the saved lse; d_o is the gradient of o; writes dq/dk/dv (bf16, same head
 * layout, any leading dimension — e.g. column slices of a packed dQKV buffer) and delta (fp32 [batch,heads,lq]
 * scratch).  Two launches: a per-query-tile pass for dq and a per-key-tile pass for dk, dv. */
typedef struct {
    const void* q; int64_t ldq;
    const void* k; int64_t ldk;
    const void* v; int64_t ldv;
    const void* o; int64_t ldo;
    const void* d_o; int64_t lddo;
    const float* lse;
    float* delta;
    void* dq; int64_t lddq;
    void* dk; int64_t lddk;
    void* dv; int64_t lddv;
    int batch, heads, lq, lk;
    float scale;
    int mask_mode;
    const float* key_mask;
    int q_per_frame, k_per_frame;
    const float* o_f32; int64_t ldof;   /* fp32 o from the forward: delta = rowsum(d_o * o_f32) in one sweep; NULL:
                                           delta from an extra sweep over the keys (same accuracy, ~40 % more work) */
    float dropout_p; uint32_t dropout_salt; const uint32_t* dropout_seed;   /* same triple as the forward call */
    const uint32_t* dropout_bits;   /* the forward's keep-bit words, or NULL: the mask is regenerated from the triple */
} shg_attn_bwd_args;
int shg_attn_bwd(const shg_attn_bwd_args* args, void* stream);
/* Problems with a single query tile (lq <= 128, o_f32 given) take ONE merged pass that also produces dk and dv
 * (default on; 0 restores the two-pass schedule, also SHG_ATTN_MERGE=0 in the environment). */
int shg_set_attn_bwd_merge(int on);
/* Problems with several query tiles (lq > 128, o_f32 given, lk <= 512) run the two passes on the second-generation
 * kernels: 64-wide chunks, two CTAs per SM, the bf16 P / dS operands kept in tensor memory (default on; 0 restores
 * the first-generation 128-wide kernels, also SHG_ATTN_BWD2=0 in the environment). */
int shg_set_attn_bwd_chunked(int on);

/* ------------------------------------------------------------------ LayerNorm family (HBM-bound)
 * y = LN(x (+ residual)) * gamma + beta, rows of width n (768 / 1536), statistics fp32.  Replaces
 * BertLayerNorm after BertAttOutput/BertOutput (lxrt/modeling_capsbert.py:431-435,485-489; eps 1e-12),
 * the DETR norms (lxrt/transformer.py:199-201; eps 1e-5) and the head norm (tasks/agqa_model.py:108).
 * Optional second output y2 = y + addend (the decoder's `tgt + query_pos`, lxrt/transformer.py:220,224);
 * addend rows repeat with period addend_rows (0 = same shape as x).  mean/rstd (fp32 [rows]) may be NULL. */
int shg_layernorm_fwd(const void* x, const void* residual, const float* gamma, const float* beta, float eps,
                      void* y, void* y2, const void* addend, int addend_rows, float* mean, float* rstd,
                      int rows, int n, void* stream);

/* Backward of the above.  g = dy (+ dy2, the gradient arriving over the residual branch); x (+ residual) is the
 * forward input; dx bf16; dgamma/dbeta fp32 [n] are ACCUMULATED; dxsum (fp32 [n], may be NULL) accumulates the
 * column sums of dx = the bias gradient of the Linear that produced x. */
/* Training-mode dropout hooks (seed: device scalar, NULL = off): in_p/in_salt — the forward output went through
 * nn.Dropout (embeddings), so dy is masked first; out_p/out_salt + dx_drop — the forward input was
 * dropout(dense_out) + residual, so dx_drop (bf16, = dx * mask/(1-p)) is the gradient the projection sees and
 * dxsum accumulates it, while dx stays the residual-branch gradient. */
int shg_layernorm_bwd(const void* dy, const void* dy2, const void* x, const void* residual, const float* gamma,
                      const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta, float* dxsum,
                      int rows, int n, void* dx_drop, float out_p, uint32_t out_salt, float in_p, uint32_t in_salt,
                      const uint32_t* seed, void* stream);

/* Grouped forms: rows [0, split_row) use (gamma, beta, dgamma, dbeta, dxsum), rows [split_row, rows) the second set —
 * the same LayerNorm position of the relation and action decoders (lxrt/transformer.py:199-201) in one launch over
 * their stacked token rows.  split_row >= rows degenerates to the plain call. */
int shg_layernorm_fwd_grouped(const void* x, const void* residual, const float* gamma, const float* beta, float eps,
                              void* y, void* y2, const void* addend, int addend_rows, float* mean, float* rstd, int rows,
                              int n, int split_row, const float* gamma2, const float* beta2, void* stream);
int shg_layernorm_bwd_grouped(const void* dy, const void* dy2, const void* x, const void* residual, const float* gamma,
                              const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta, float* dxsum,
                              int rows, int n, void* dx_drop, float out_p, uint32_t out_salt, float in_p,
                              uint32_t in_salt, const uint32_t* seed, int split_row, const float* gamma2,
                              float* dgamma2, float* dbeta2, float* dxsum2, void* stream);

/* out[n] += column sums of the bf16 matrix x [rows, n] (ld elements per row, ld % 8 == 0): bias gradients. */
int shg_colsum_accum(const void* x, int64_t ld, int rows, int n, float* out, void* stream);

/* Embedding gathers fused with LayerNorm: BertEmbeddings (lxrt/modeling_capsbert.py:341-355; word[ids] +
 * pos[t] + type[seg]) and HGEmbeddings (:313-325; word_ids == NULL: row t of the table, pos == NULL).
 * Tables fp32 [*, n]; ids int64 [rows]; t = row % seq_len.  y bf16; pre (bf16 pre-LN sum), mean, rstd are
 * the saved tensors for the backward (may be NULL). */
int shg_embed_ln_fwd(const float* word, const int64_t* word_ids, const float* pos, const float* type,
                     const int64_t* type_ids, const float* gamma, const float* beta, float eps, void* y, void* pre,
                     float* mean, float* rstd, int rows, int seq_len, int n, float drop_p, uint32_t drop_salt,
                     const uint32_t* seed, void* stream);
/* Scatter-add of the gradient wrt the pre-LN sum (bf16 [rows,n]) into the fp32 table gradients;
 * *_pad = padding_idx whose row receives no gradient (nn.Embedding(padding_idx=0)), or -1. */
int shg_embed_scatter_bwd(const void* de, const int64_t* word_ids, const int64_t* type_ids, float* dword,
                          int word_pad, float* dpos, int pos_pad, float* dtype, int type_pad, int rows, int seq_len,
                          int n, void* stream);

/* y[r,:] = x[r,:] + add[r % add_rows,:] (bf16): type tokens of CrossEncoder (lxrt/modeling_capsbert.py:1168-1181),
 * learned visual position rows (:1019-1024). */
int shg_add_rows(const void* x, const void* add, int add_rows, void* y, int rows, int n, void* stream);
/* dx = dy * act'(aux) over n bf16 elements: act = SHG_ACT_GELU_ERF (aux = saved pre-activation),
 * SHG_ACT_TANH / SHG_ACT_RELU (aux = saved output).  Backward of the MLP heads and the pooler. */
int shg_act_bwd(const void* dy, const void* aux, void* dx, int64_t n, int act, void* stream);
/* Answer loss: out[0] += BCEWithLogitsLoss()(x, target) * cols (tasks/agqaHGQA.py:344-345), x fp32 [rows, cols]
 * (ldx), target fp32 [rows, cols]; dx (bf16, lddx; may be NULL) = grad_scale * d loss / d x. */
int shg_bce_logits(const float* x, int64_t ldx, const float* target, int rows, int cols, float grad_scale, float* out,
                   void* dx, int64_t lddx, void* stream);
/* Strided block copy (bf16): dst[i0*d0 + i1*d1 + e] (+)= src[i0*s0 + i1*s1 + e] (+ add[e]), e < chunk.  Assembles
 * the hypergraph token sequence [CLS | per frame: 3 action + 8 relation slots] with its type tokens
 * (tasks/agqa_model.py:255-258, lxrt/modeling_capsbert.py:1168-1184), picks [CLS] rows for BertPooler2, and routes
 * gradients back (accumulate != 0 adds into dst). */
int shg_copy3d(const void* src, void* dst, const void* add, int n0, int n1, int chunk, int64_t s0, int64_t s1,
               int64_t d0, int64_t d1, int accumulate, void* stream);
/* fp32 -> bf16 shadow copy of parameters / inputs. */
int shg_f32_to_bf16(const float* x, void* y, int64_t n, void* stream);

/* ------------------------------------------------------------------ optimiser (HBM-bound, flat buffers)
 * Replaces clip_grad_norm_(params, 5.) + BertAdam.step() (tasks/agqaHGQA.py:391-392, lxrt/optimization.py:101-180:
 * no bias correction, decoupled weight decay on every tensor, lr from warmup_linear).  The buffers are flat fp32
 * arrays; the "live" parameters (those that receive gradients — the reference skips p.grad is None) are given as
 * segments: seg_start[i] = offset of segment i, seg_prefix[i] = number of live elements before it (both multiples
 * of 4, device int64), total = live element count.
 * shg_grad_sumsq: out[0] = sum g^2, bit-reproducible (fixed-order two-stage sum: data-parallel replicas must derive
 * the same clip coefficient from the same all-reduced gradients); workspace = 2048 floats, zero-initialised once by
 * the caller, left zero-countered by every call.  shg_bert_adam_step: g' = g * grad_scale * min(1, max_norm/(|g*grad_scale|+1e-6))
 * (max_norm <= 0: no clipping), then the Adam update in place; refreshes the bf16 shadow (first n_shadow elements
 * of the flat space) and optionally zeroes the gradient for the next step.  lr_dev (device float, may be NULL)
 * overrides lr so that a captured CUDA graph follows the warm-up schedule. */
int shg_grad_sumsq(const float* grad, const int64_t* seg_start, const int64_t* seg_prefix, int n_seg,
                   int64_t total, float* out, float* workspace, void* stream);
int shg_bert_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow, int64_t n_shadow,
                       const int64_t* seg_start, const int64_t* seg_prefix, int n_seg, int64_t total,
                       const float* sumsq, float max_norm, float lr, const float* lr_dev, float b1, float b2, float eps,
                       float weight_decay, float grad_scale, int zero_grad, void* stream);

/* ------------------------------------------------------------------ Conv3d visual front-end (tcgen05 implicit GEMM)
 * Replaces VisualFeatEncoder.conv and the token assembly after it (lxrt/modeling_capsbert.py:989-996, 1037-1073):
 *   nn.Sequential(ZeroPad2d(1), Conv3d(2048, 768, (5,3,3)), GeLU(), ZeroPad2d(1), Conv3d(768, 768, (5,3,3)), GeLU())
 *   x = cat(cls_token, img.permute(0,2,3,4,1).view(B, t*H*W, c)); x = position_encoding(x); dropout(x)
 * and their backward.  Activations are channels-last bf16 [batch, t, 7, 7, c] (the spatial size is fixed at 7x7, the
 * slow_r50 output); the weight is bf16 packed [c_out, 45, c_in] with tap = (dt*3 + dh)*3 + dw — i.e.
 * weight.permute(0,2,3,4,1) of the reference's [c_out, c_in, 5, 3, 3].  No im2col buffer: 5-D TMA boxes with
 * out-of-bounds zero fill are the padded, shifted operand tiles.  t_out = t_in - 4; t_in a multiple of 4, >= 8;
 * c_in, c_out multiples of 256. */
typedef struct {
    const void* x;          /* bf16 [batch, t_in, 7, 7, c_in] */
    const void* w;          /* bf16 packed [c_out, 45, c_in] */
    const float* bias;      /* fp32 [c_out] or NULL */
    void* y;                /* bf16: GELU(conv) for output position pos = (t*7 + h)*7 + w of clip b is written to row
                               b*y_rows_per_clip + y_row_off + pos (ldy elements per row) — (393, 1) drops the second
                               conv's output straight into the [B*393, 768] token matrix behind the cls row */
    int64_t ldy;
    int y_rows_per_clip;    /* 0 = t_out*49 */
    int y_row_off;
    void* preact;           /* bf16 [batch*t_out*49, c_out]: conv + bias before the GELU, saved for the backward; or NULL */
    const float* pos_emb;   /* fp32 [>= y_row_off + t_out*49, c_out]: row (y_row_off + pos) is added after the GELU
                               (LearnedPositionalEncoding, lxrt/modeling_capsbert.py:1066); or NULL */
    int batch, t_in, c_in, c_out;
    float dropout_p;        /* nn.Dropout on the token (:1068); mask = f(*dropout_seed, salt, out_row*c_out + channel) */
    uint32_t dropout_salt;
    const uint32_t* dropout_seed;
} shg_conv3d_args;
int shg_conv3d_fwd(const shg_conv3d_args* args, void* stream);
/* dx [batch, t_in, 7, 7, c_in] bf16 = conv_transpose(dy [batch, t_in-4, 7, 7, c_out], w) (x gelu'(preact) when preact,
 * bf16 [batch*t_in*49, c_in] = the saved pre-activation of the layer that produced x, is given). */
int shg_conv3d_dgrad(const void* dy, const void* w, const void* preact, void* dx, int batch, int t_in, int c_in,
                     int c_out, void* stream);
/* dw fp32 packed [c_out, 45, c_in] += sum over positions of dy (x) shifted x. */
int shg_conv3d_wgrad(const void* dy, const void* x, float* dw, int batch, int t_in, int c_in, int c_out, void* stream);
/* fp32 [batch, channels, positions] (the backbone's NCDHW output, positions = T*H*W) -> bf16 channels-last
 * [batch, positions, channels], and the way back for the gradient handed to the backbone. */
int shg_ncdhw_to_ndhwc(const float* x, void* y, int batch, int channels, int positions, void* stream);
int shg_ndhwc_to_ncdhw(const void* x, float* y, int batch, int channels, int positions, void* stream);
/* tokens[b*rows_per_clip, :] = dropout(cls + pe[0]) (bf16, ld elements per row) — the cls row of every clip. */
int shg_frontend_cls(const float* cls, const float* pe, void* tokens, int64_t ld, int batch, int rows_per_clip, int n,
                     float drop_p, uint32_t drop_salt, const uint32_t* seed, void* stream);
/* Backward of the token assembly.  d_tok bf16 [batch*rows_per_clip, n] (ld): g = d_tok * dropmask/(1-p);
 * d_pe fp32 [rows_per_clip, n] += sum_b g; d_cls fp32 [n] += sum_b g[b, 0]; d_pre bf16 [batch*(rows_per_clip-1), n] =
 * g[b, 1:] * gelu'(preact) — the gradient wrt the second conv's pre-activation. */
int shg_frontend_bwd(const void* d_tok, int64_t ld, const void* preact, void* d_pre, float* d_pe, float* d_cls, int batch,
                     int rows_per_clip, int n, float drop_p, uint32_t drop_salt, const uint32_t* seed, void* stream);

#ifdef __cplusplus
}
#endif
#endif
