/*
 * b200t5 -- C ABI of the B200-native (sm_100a) replacement for the hot path of catie-aq/flashT5.
 *
 * One shared library (libb200t5.so), plain pointers and sizes, no torch types.  Every entry point
 * is what the reference's Python launcher for that kernel would bind through ctypes (see
 * INTEGRATION.md).  Citations are relative to the reference checkout.
 *
 * Conventions
 *   - all pointers are DEVICE pointers on `device`; the library never allocates, frees or
 *     synchronises; every kernel is enqueued on `stream` (a cudaStream_t passed as void*);
 *     workspaces are caller-allocated (query the size first); entry points are CUDA-graph safe.
 *   - strides are in ELEMENTS, order (batch, head, seq, dim); the last (dim) stride must be 1.
 *   - return value: 0 on success, a negative B200T5_ERR_* code otherwise; the message is kept in
 *     thread-local storage and returned by b200t5_last_error().
 *   - there is NO CPU fallback: with no sm_100 device every compute entry point fails with
 *     B200T5_ERR_CUDA / B200T5_ERR_UNSUPPORTED.
 */
#ifndef B200T5_H_
#define B200T5_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200T5_ABI_VERSION 3   /* 3: attn flags (was reserved0), b200t5_attn_fwd_workspace_bytes */

#if defined(__GNUC__)
#define B200T5_API __attribute__((visibility("default")))
#else
#define B200T5_API
#endif

enum {
    B200T5_OK = 0,
    B200T5_ERR_INVALID = -1,     /* bad shape / stride / pointer alignment / null pointer          */
    B200T5_ERR_UNSUPPORTED = -2, /* dtype or head dim outside {fp16,bf16} x {16,32,64,128}, non sm_100 device */
    B200T5_ERR_CUDA = -3,        /* a CUDA runtime / driver call or kernel launch failed             */
    B200T5_ERR_WORKSPACE = -4    /* workspace missing or too small                                   */
};

enum { B200T5_F16 = 0, B200T5_BF16 = 1, B200T5_F32 = 2 };

/* ------------------------------------------------------------------------------------------------
 * Attention with additive bias.
 * Replaces torch.ops.flasht5.flash_attn_v2_fwd / _bwd and their Triton kernels
 *   src/model/ops/flash_attention_v2_bias.py:27-80 (fwd launcher), :327-483 (_fwd_kernel),
 *   :91-217 (bwd launcher), :516-556 (_bwd_preprocess), :559-745 (_bwd_kv_kernel),
 *   :748-905 (_bwd_q_kernel), :214-215 (ds.sum(0)).
 *
 * q:(B,H,M,D)  k,v:(B,H,N,D)  bias:(bias_B,bias_H,M,N) with bias_B in {1,B}, bias_H in {1,H},
 * or bias == NULL (cross attention).  o has q's shape; lse:(B,H,M) fp32 contiguous.
 * Semantics: S = sm_scale * Q K^T + bias, bottom-right aligned causal
 * mask (visible iff n <= m + N - M), lse = ln sum exp S (natural log), rows with no visible key
 * give o = 0 and lse = -inf.
 * ---------------------------------------------------------------------------------------------- */
/* flags of b200t5_attn_params: with B200T5_ATTN_DETERMINISTIC the backward gives every key block its own slice of the dQ
 * surface and every batch element its own slice of the dS surface, so that no two partial results ever meet in an
 * L2 reduce-add and the final sums run in a fixed order: dQ and dBias are then bitwise reproducible run to run (dK, dV, O, L
 * always are), at the price of a larger workspace (the workspace query honours the flag). */
#define B200T5_ATTN_DETERMINISTIC 1
/* B200T5_ATTN_DBIAS_F32 (backward, head dims 16 / 32 / 64): `dbias` points at an fp32 tensor of bias's shape (strides in fp32
 * elements); the sum over the broadcast dimensions is delivered unrounded, so that a data-parallel caller can all-reduce it
 * across ranks and round ONCE afterwards (flasht5_b200/data_parallel.py). */
#define B200T5_ATTN_DBIAS_F32 2
/* B200T5_ATTN_DBIAS_ACCUMULATE (with B200T5_ATTN_DBIAS_F32): dbias += the batch-summed dS instead of dbias = ...  The layers of a
 * T5 stack share one position-bias tensor (reference modeling_flash_t5.py:452-455), so autograd adds L - 1 dense (H, M, N)
 * gradients per stack; with this flag every layer's backward adds into ONE persistent fp32 buffer and the caller hands
 * autograd a single gradient.  Causal-masked entries are left untouched. */
#define B200T5_ATTN_DBIAS_ACCUMULATE 4

typedef struct b200t5_attn_params {
    /* problem */
    int32_t B, H, M, N, D;
    int32_t dtype;        /* B200T5_F16 | B200T5_BF16: q,k,v,bias,o,do,dq,dk,dv,dbias all share it */
    int32_t causal;       /* 0 | 1 */
    int32_t bias_B;       /* 1 or B  (ignored when bias == NULL) */
    int32_t bias_H;       /* 1 or H */
    float sm_scale;
    int32_t device;       /* CUDA device ordinal the pointers live on */
    int32_t flags;        /* 0, or B200T5_ATTN_DETERMINISTIC (backward only; was reserved0 = 0 in ABI version 1 callers) */
    void* stream;         /* cudaStream_t */

    /* forward operands */
    const void* q;  int64_t q_strides[4];
    const void* k;  int64_t k_strides[4];
    const void* v;  int64_t v_strides[4];
    const void* bias; int64_t bias_strides[4];   /* strides of the size-1 dims are ignored */
    void* o;        int64_t o_strides[4];        /* fwd: output; bwd: input */
    float* lse;                                   /* fwd: output; bwd: input */

    /* backward operands (ignored by b200t5_attn_fwd) */
    const void* dout; int64_t do_strides[4];
    void* dq;       int64_t dq_strides[4];
    void* dk;       int64_t dk_strides[4];
    void* dv;       int64_t dv_strides[4];
    void* dbias;    int64_t dbias_strides[4];    /* bias's shape; NULL iff bias == NULL */
    void* workspace;                              /* >= b200t5_attn_bwd_workspace_bytes() bytes, 256-B aligned */
    size_t workspace_bytes;
} b200t5_attn_params;

B200T5_API int b200t5_attn_fwd(const b200t5_attn_params* p);
/* Optional forward workspace: non-zero only for a bias whose rows a TMA tensor map cannot address (N % 8 != 0, a row stride
 * that is not a multiple of 8 elements, an unaligned base).  Given >= this many bytes in p->workspace (256-B aligned) the
 * forward copies the bias once into rows padded to 8 elements and streams it with TMA; without it the bias is read
 * element by element (correct, ~2x slower forward).  The reference has no such restriction to mirror (Triton pointers). */
B200T5_API size_t b200t5_attn_fwd_workspace_bytes(const b200t5_attn_params* p);
B200T5_API size_t b200t5_attn_bwd_workspace_bytes(const b200t5_attn_params* p);
B200T5_API int b200t5_attn_bwd(const b200t5_attn_params* p);

/* ------------------------------------------------------------------------------------------------
 * RMSNorm.  Replaces torch.ops.flasht5.rmsnorm_triton_fwd / _bwd
 *   src/model/ops/rms_norm.py:134-174 (+ kernel :25-66), :186-236 (+ kernel :68-131).
 * x, y, dy, dx: (rows, n) with unit last stride; w, dw: (n); rstd: (rows) fp32.
 * x_dtype / w_dtype: B200T5_F16 | B200T5_BF16 | B200T5_F32 (y, dy, dx use x_dtype; dw uses w_dtype).
 * bwd needs a workspace of b200t5_rmsnorm_bwd_workspace_bytes(n) bytes (fp32 partial dW rows).
 * ---------------------------------------------------------------------------------------------- */
B200T5_API int b200t5_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int64_t rows, int64_t n,
                       int64_t x_row_stride, int64_t y_row_stride, float eps, int x_dtype, int w_dtype,
                       int device, void* stream);
B200T5_API size_t b200t5_rmsnorm_bwd_workspace_bytes(int64_t n);
B200T5_API int b200t5_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, void* dw,
                       void* workspace, size_t workspace_bytes, int64_t rows, int64_t n, int64_t dy_row_stride,
                       int64_t x_row_stride, int64_t dx_row_stride, int x_dtype, int w_dtype, int device,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cross-entropy + z-loss.  Replaces torch.ops.flasht5.cross_entropy_triton_fwd / _bwd
 *   src/model/ops/cross_entropy_loss.py:164-217 (+ kernel :40-111), :228-274 (+ kernel :119-162).
 * logits: (rows, vocab) unit last stride, dtype in {F16,BF16,F32}; labels: (rows) int64;
 * losses, z_losses, lse: (rows) fp32.  lse_is_input != 0: `lse` holds a precomputed log-sum-exp that is
 * used instead of being recomputed (reference PRECOMPUTED_LSE, :66-76; requires smoothing == 0 and
 * logit_scale == 1).  dlogits may alias logits (in-place backward).
 * dlosses: (rows) fp32 with element stride dloss_stride.
 * ---------------------------------------------------------------------------------------------- */
B200T5_API int b200t5_ce_fwd(const void* logits, const int64_t* labels, float* losses, float* z_losses, float* lse,
                  int lse_is_input, int64_t rows, int64_t vocab, int64_t row_stride, float smoothing,
                  float logit_scale, float lse_square_scale, int64_t ignore_index, int dtype, int device,
                  void* stream);
B200T5_API int b200t5_ce_bwd(const void* logits, const int64_t* labels, const float* lse, const float* dlosses,
                  int64_t dloss_stride, void* dlogits, int64_t rows, int64_t vocab, int64_t row_stride,
                  int64_t dlogits_row_stride, float smoothing, float logit_scale, float lse_square_scale,
                  int64_t ignore_index, int dtype, int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * T5 relative-position bias producer (the step that builds the attention operator's `bias` input).
 * Replaces RelativePositionalEncoding.compute_bias, src/utils/positional_encoding.py:73-102 (bucket -> embedding
 * gather -> (1,H,M,N)) and the scatter-add backward of that gather.
 * table: (num_buckets, H) row-major, dtype table_dtype.  lut: int32[lut_len], lut[rel + lut_zero] = bucket of the
 * relative position rel = mem_pos[n] - ctx_pos[m] (out-of-range indices are clamped); the caller builds it with the
 * reference formula (:25-71).  ctx_pos: int32[M] or NULL (= 0..M-1); mem_pos: int32[N] or NULL (= 0..N-1).
 * bias / dbias: (1, H, M, N) contiguous, dtype bias_dtype / dbias_dtype.  dtable: (num_buckets, H) fp32, overwritten.
 * num_buckets <= 256.  All dtypes in {F16, BF16, F32}.
 * ---------------------------------------------------------------------------------------------- */
B200T5_API int b200t5_t5_bias_fwd(const void* table, const int32_t* lut, int32_t lut_zero, int32_t lut_len,
                                  const int32_t* ctx_pos, const int32_t* mem_pos, void* bias, int32_t H, int32_t M,
                                  int32_t N, int32_t num_buckets, int table_dtype, int bias_dtype, int device,
                                  void* stream);
B200T5_API int b200t5_t5_bias_bwd(const void* dbias, const int32_t* lut, int32_t lut_zero, int32_t lut_len,
                                  const int32_t* ctx_pos, const int32_t* mem_pos, float* dtable, int32_t H, int32_t M,
                                  int32_t N, int32_t num_buckets, int dbias_dtype, int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention with the T5 relative-position bias computed inside the kernels (no (1,H,M,N) bias tensor, no
 * (1,H,M,N) gradient handed back): the operator behind the reference's `attention_type == "fa2_rpe"` call
 *   src/model/modeling_flash_t5.py:275-279  flash_attn_func(q, k, v, softmax_scale=..., causal=...,
 *       rpe_weights=relative_attention_bias.weight.t(), rpe_max_distance=...)
 * (the flash-attention fork that implements it is not part of the reference checkout; the semantics are those of
 * the dense path with bias = RelativePositionalEncoding.compute_bias(M, N), src/utils/positional_encoding.py:73-102,
 * cast to the q dtype, and dtable = the gradient that autograd would scatter back into the embedding table).
 *
 * bias[h, m, n] = table[lut[(n - m) + lut_zero], h].  The caller passes the bucket lookup table (built with the
 * reference formula :25-71) and the two relative positions beyond which it is constant:
 *   every rel <= const_lo has lut == lut[const_lo + lut_zero]; every rel >= const_hi has lut == lut[const_hi + lut_zero]
 *   (const_lo < const_hi; -(M-1) <= const_lo and const_hi <= N-1 always qualify).
 * `band` is a caller-allocated fp32 buffer of H * b200t5_rpe_band_len(const_lo, const_hi) elements:
 * b200t5_rpe_band() fills it (bias per head over the relative positions const_lo-255 .. const_hi+255, rounded to
 * the attention dtype), b200t5_attn_rpe_fwd / _bwd read it.  One band serves every layer that shares the table.
 * Limits: band_len <= 8192 (B200T5_ERR_UNSUPPORTED otherwise: use the dense path), num_buckets <= 256.
 * b200t5_attn_rpe_fwd / _bwd take the same b200t5_attn_params as the dense entry points with bias == NULL and
 * dbias == NULL; _bwd additionally writes dtable (num_buckets, H) fp32 row-major (overwritten) and needs a
 * workspace of b200t5_attn_rpe_bwd_workspace_bytes() bytes.
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200t5_rpe_params {
    const void* table;             /* element (bucket, head) at table[bucket * table_stride_b + head * table_stride_h] */
    int64_t table_stride_b, table_stride_h;
    int32_t table_dtype;           /* B200T5_F16 | B200T5_BF16 | B200T5_F32 */
    int32_t num_buckets;
    const int32_t* lut;            /* int32[lut_len]; must cover rel in [-(M-1), N-1] */
    int32_t lut_zero, lut_len;
    int32_t const_lo, const_hi;
    float* band;                   /* (H, band_len) fp32 */
    float* dtable;                 /* backward only */
} b200t5_rpe_params;

B200T5_API int b200t5_rpe_band_len(int32_t const_lo, int32_t const_hi);
B200T5_API int b200t5_rpe_band(const b200t5_rpe_params* r, int32_t H, int io_dtype, int device, void* stream);
B200T5_API int b200t5_attn_rpe_fwd(const b200t5_attn_params* p, const b200t5_rpe_params* r);
B200T5_API size_t b200t5_attn_rpe_bwd_workspace_bytes(const b200t5_attn_params* p, const b200t5_rpe_params* r);
B200T5_API int b200t5_attn_rpe_bwd(const b200t5_attn_params* p, const b200t5_rpe_params* r);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-tensor AdamWScale step.
 * Replaces AdamWScale._adamwscaled / _foreach_adamwscaled, src/utils/adamw_scaled.py:154-211, :213-281: Adam moments,
 * step size x max(1e-3, rms(parameter)), optional Kahan compensation for 16-bit parameters, decoupled weight decay --
 * three launches for all tensors of a (parameter dtype, state dtype, kahan) group, no host synchronisation.
 * `tensors`: DEVICE array of n_tensors descriptors; `chunk_tensor`: DEVICE int32[n_chunks], the tensor every
 * b200t5_adamw_chunk_elems()-element chunk belongs to (chunks of a tensor are consecutive, first_chunk is its first).
 * The caller computes, with the reference's own expressions (:173-180),
 *   ss_base  = lr [* sqrt(1 - beta2^t) / (1 - beta1^t)]      as fp32
 *   ss_floor = that step size x 1e-3                          (used when rms(p) <= 1e-3)
 *   round_step_to_p: 1 when the step size x rms product is a tensor of the parameter dtype in the reference (no bias
 *                    correction: python float x 16-bit rms tensor), 0 when it is fp32 (bias correction on).
 * p_dtype / state_dtype: B200T5_F16 | B200T5_BF16 | B200T5_F32; state_dtype == p_dtype, or 16-bit states under fp32
 * parameters (use_state_dtype); kahan requires a 16-bit p_dtype.  Gradients have the parameter dtype.
 * workspace: b200t5_adamw_workspace_bytes(n_tensors, n_chunks) bytes, 256-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200t5_adamw_tensor {
    void* p;
    const void* g;
    void* m;
    void* v;
    void* comp;             /* Kahan compensation (parameter dtype), NULL unless kahan */
    int64_t numel;
    int32_t first_chunk;
    float sqrt_numel;       /* (float)(numel ** 0.5) */
    float ss_base;
    float ss_floor;
    float neg_lr_wd;        /* -(lr * weight_decay) as fp32; 0 = no decay */
    int32_t reserved;
} b200t5_adamw_tensor;

B200T5_API int b200t5_adamw_chunk_elems(void);
B200T5_API size_t b200t5_adamw_workspace_bytes(int32_t n_tensors, int32_t n_chunks);
B200T5_API int b200t5_adamw_scale_step(const b200t5_adamw_tensor* tensors, int32_t n_tensors, const int32_t* chunk_tensor,
                                       int32_t n_chunks, void* workspace, size_t workspace_bytes, int p_dtype,
                                       int state_dtype, int kahan, float beta1, float beta2, float eps,
                                       int round_step_to_p, int device, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Library state
 * ---------------------------------------------------------------------------------------------- */
B200T5_API int b200t5_abi_version(void);
B200T5_API const char* b200t5_last_error(void);          /* thread-local, never NULL */
B200T5_API uint64_t b200t5_launch_count(void);           /* kernels launched by this library since load (all threads) */
/* 1 if `device` is an sm_100 part this library can run on, 0 otherwise, negative on CUDA error */
B200T5_API int b200t5_device_supported(int device);

/* ------------------------------------------------------------------------------------------------
 * Per-kernel device timing (measurement hook for bench.py's roofline leg; off by default).
 * While enabled, the two main attention kernels are bracketed with CUDA events recorded on the launch
 * stream.  b200t5_profile_collect() must be called after the caller has synchronised that stream; it
 * writes up to `cap` (kernel id, milliseconds) pairs in launch order, returns how many, and clears
 * the list.  Kernel ids: 1 = attention forward, 2 = attention backward (the fused dQ/dK/dV/dS kernel).
 * Not thread-safe against concurrent launches; not for use under CUDA-graph capture.
 * ---------------------------------------------------------------------------------------------- */
enum { B200T5_KERNEL_ATTN_FWD = 1, B200T5_KERNEL_ATTN_BWD = 2 };
B200T5_API int b200t5_profile_enable(int enable);
B200T5_API int b200t5_profile_collect(int* kernel_ids, float* ms, int cap);

#ifdef __cplusplus
}
#endif
#endif /* B200T5_H_ */
