// Internal kernel-parameter blocks and launcher prototypes shared by the .cu files.
// (The public C ABI is include/b200t5.h; nothing here is exported.)
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200t5 {

// Bias mode 3: bias[h, m, n] = band[h][(n - m) - band_lo], a Toeplitz bias given as one fp32 row per head over
// the relative positions band_lo .. band_lo + band_len - 1 (values already rounded to the io dtype).  Every
// rel <= const_lo shares the value at const_lo and every rel >= const_hi the value at const_hi, so a 128x128 tile
// whose relative positions all lie on one side needs no per-element lookup.  band_lo = const_lo - 255 and
// band_len = const_hi - const_lo + 511: any tile that is not constant stays inside the band.
struct RpeBand {
    const float* band;      // (H, band_len) fp32
    int band_lo, band_len;
    int const_lo, const_hi;
    // backward, optional: tiles that lie entirely beyond a constant end do not store dS; the CTA adds the sum of its
    // dS over such tiles to dconst[h * 2 + side] (side 0: rel <= const_lo, 1: rel >= const_hi), fp32, pre-zeroed
    float* dconst;          // NULL: every tile stores dS
};
constexpr int kRpeBandPad = 255;        // 2 * 128 - 1 relative positions per tile
constexpr int kRpeMaxBandLen = 4096;    // 16 KB of shared memory (what the backward kernel can spare beside its K double buffer and a 5-slot ring)

struct AttnFwdKernelParams {
    CUtensorMap map_q;      // (D, M, H, B)   box (min(D,64), 128, 1, 1)
    CUtensorMap map_k;      // (D, N, H, B)
    CUtensorMap map_v;      // (D, N, H, B)
    CUtensorMap map_bias;   // (N, M, Hb, Bb) box (64, 128, 1, 1)            [bias mode 1]
    CUtensorMap map_o;      // (D, M, H, B) box (64, 128, 1, 1): the output tile (TMA store; D = 64)
    const void* bias;       // raw pointer                                   [bias mode 2]
    int64_t bias_sb, bias_sh, bias_sm, bias_sn;   // element strides          [bias mode 2]
    void* o;
    int64_t o_sb, o_sh, o_sm;                     // element strides, last dim contiguous
    float* lse;             // (B, H, M) contiguous
    int B, H, M, N;
    int num_m_blocks;
    int bias_b_bcast, bias_h_bcast;
    float sm_scale;
    // in-kernel T5 relative-position bias                                     [bias mode 3]
    RpeBand rpe;
};

struct AttnBwdKernelParams {
    CUtensorMap map_q;      // (D, M, H, B)
    CUtensorMap map_k;      // (D, N, H, B)
    CUtensorMap map_v;      // (D, N, H, B)
    CUtensorMap map_do;     // (D, M, H, B)
    CUtensorMap map_bias;   // (N, M, Hb, Bb)                                 [bias mode 1]
                            // (the v3 kernel, D <= 64, reads a repacked copy through `bias` instead)
    CUtensorMap map_ds;     // (N, M, H, B) 16-bit dS workspace, row pitch = N rounded up to 8   [bias modes 1, 2]
                            // v3 kernel (D <= 64): the TRANSPOSED surface (M, N, H, G), row pitch = M rounded up to 8, box (32, 128)
    const void* k;          // raw K / V pointers + element strides: the v3 kernel copies its key block into TMEM itself
    int64_t k_sb, k_sh, k_sn;
    const void* v;
    int64_t v_sb, v_sh, v_sn;
    CUtensorMap map_dq;     // 16-bit dQ group surface (D, M, H, dq_groups * B), box (min(D,64), 128, 1, 1), reduce-add
    CUtensorMap map_dk_st, map_dv_st;   // v3 kernel, D = 64: dK / dV (D, N, H, B) with box (D / 2, 128): one warpgroup's half of a tile (TMA store)
    const void* bias;       // [bias mode 2]
    int64_t bias_sb, bias_sh, bias_sm, bias_sn;
    void* dk;
    int64_t dk_sb, dk_sh, dk_sn;
    void* dv;
    int64_t dv_sb, dv_sh, dv_sn;
    const float* lse;       // (B, H, M)        [D = 128 kernel]
    const float* delta;     // (B, H, M)        [D = 128 kernel]
    const float* nl;        // (B, H, m_pad) -L * log2e (-inf: P = 0)   [v3 kernel: statistics as the kernel consumes them]
    const float* ndelta;    // (B, H, m_pad) -delta
    int m_pad;              // M rounded up to 128
    int B, H, M, N;
    int num_m_blocks, num_n_blocks;
    int bias_b_bcast, bias_h_bcast;
    float sm_scale;
    // dS surface: (N, M, H, ds_groups) 16-bit.  Batch b adds its tile into group b % ds_groups with a TMA
    // reduce-add (ds_use_reduce = 1, surface pre-zeroed) or owns its slice and stores (ds_use_reduce = 0).
    int ds_groups, ds_use_reduce;
    // dQ surface: key block nb adds its partial dQ tile (rounded to the io dtype) into group nb % dq_groups;
    // groups hold <= 4 key blocks each and are summed in fp32 by the convert kernel.
    int dq_groups;
    // in-kernel T5 relative-position bias                                     [bias mode 3]
    RpeBand rpe;
};

cudaError_t launch_bias_align_copy(const void* bias, const int64_t* strides, void* dst, int Bb, int Hb, int M, int N, int pitch,
                                   cudaStream_t stream);
cudaError_t launch_attn_fwd(const AttnFwdKernelParams& kp, int D, bool bf16, int bias_mode, bool causal,
                            cudaStream_t stream);
cudaError_t launch_attn_bwd(const AttnBwdKernelParams& kp, int D, bool bf16, int bias_mode, bool causal,
                            cudaStream_t stream);        // the D = 128 kernel (attn_bwd.cu)
// transposed-formulation kernel (attn_bwd_v3.cu), D <= 64, bias modes 0, 1 (through the transposed copy), 3
cudaError_t launch_attn_bwd_v3(const AttnBwdKernelParams& kp, int D, bool bf16, int bias_mode, bool causal,
                               cudaStream_t stream);
// Repacked copy of a dense bias for the v3 kernel: out[bh][key block of 128][query block of 32][c = 0..3][n = 0..127][8]
// = bias[bh][32 * qb + 8 * c + e][128 * kb + n] (16-bit elements, arbitrary input strides, zero beyond M / N)
cudaError_t launch_bias_repack(const void* bias, const int64_t* strides, void* bias_p, int Bb, int Hb, int M, int N,
                               cudaStream_t stream);
// v3 kernel helpers fused pairwise (independent work in one launch, block-index split):
//   pre : delta + zero-fills (+ the bias repack when bias != NULL)        post : dQ conversion (+ transposing dBias reduce when dbias != NULL)
//   (the row statistics are written in the form the v3 kernel consumes: nl = -L * log2e and ndelta = -delta, rows padded to m_pad)
cudaError_t launch_attn_bwd_pre_fused(const void* o, const int64_t* o_strides, const void* dout, const int64_t* do_strides,
                                      const float* lse, float* nl_out, float* ndelta_out, int m_pad, void* dq_ws, int dq_groups,
                                      int B, int H, int M, int N, int D, bool bf16, void* zero_ptr, size_t zero_bytes,
                                      const void* bias, const int64_t* bias_strides, void* bias_p, int Bb, int Hb,
                                      cudaStream_t stream);
cudaError_t launch_attn_bwd_post_fused(const void* dq_ws, int dq_groups, void* dq, const int64_t* dq_strides, int B, int H, int M,
                                       int N, int D, float sm_scale, bool bf16, const void* ds_t, int m_pitch, void* dbias,
                                       const int64_t* dbias_strides, int G, int reduce_b, int reduce_h, bool causal, bool out_f32,
                                       cudaStream_t stream, bool accumulate = false);
// dbias[bb,hb,m,n] = sum over broadcast batch-group / head of ds_t[g,h,n,m] (the transposed surface of the v3 kernel);
// fp32 accumulation, one rounding; causal-masked entries are written as 0 without being read
// out_f32: dbias is an fp32 tensor (unrounded sums) instead of the io dtype; accumulate (fp32 only): dbias += the sums
cudaError_t launch_dbias_reduce_t(const void* ds_t, int m_pitch, void* dbias, const int64_t* dbias_strides, int G, int H,
                                  int M, int N, int reduce_b, int reduce_h, bool causal, bool bf16, bool out_f32,
                                  cudaStream_t stream, bool accumulate = false);

// delta[b,h,m] = sum_d O*dO  (fp32); also zero-fills the 16-bit dQ group surface (dq_groups, B, H, M, D).
// `zero_ptr` / `zero_bytes` (multiple of 16, may be 0): an extra surface to zero-fill in the same launch.
cudaError_t launch_attn_bwd_preprocess(const void* o, const int64_t* o_strides, const void* dout,
                                       const int64_t* do_strides, float* delta, void* dq_ws, int dq_groups, int B,
                                       int H, int M, int D, bool bf16, void* zero_ptr, size_t zero_bytes,
                                       cudaStream_t stream);
// dq convert + dBias reduce (one fused launch when possible); dbias may be NULL
cudaError_t launch_attn_bwd_finalize(const void* dq_ws, int dq_groups, void* dq, const int64_t* dq_strides, int B, int H,
                                     int M, int N, int D, float sm_scale, bool bf16, const void* ds_ws, int ws_pitch,
                                     void* dbias, const int64_t* dbias_strides, int G, int reduce_b, int reduce_h,
                                     bool causal, cudaStream_t stream);
// dq[b,h,m,:] = 16bit(sm_scale * sum_g dq_ws[g,b,h,m,:])
cudaError_t launch_attn_bwd_dq_convert(const void* dq_ws, int dq_groups, void* dq, const int64_t* dq_strides, int B,
                                       int H, int M, int D, float sm_scale, bool bf16, cudaStream_t stream);
// dbias[bb,hb,m,n] = sum over broadcast batch/head of ds_ws[b,h,m,n]; causal-masked entries are 0 (never read).
cudaError_t launch_dbias_reduce(const void* ds_ws, int ws_pitch, void* dbias, const int64_t* dbias_strides, int B, int H,
                                int M, int N, int reduce_b, int reduce_h, bool causal, bool bf16, cudaStream_t stream);

cudaError_t launch_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int rows, int n,
                               int64_t x_row_stride, int64_t y_row_stride, float eps, int x_dtype, int w_dtype,
                               cudaStream_t stream);
constexpr int kRmsnormMaxPartials = 296;   // 2 CTAs per SM x 148 SMs: rows of the fp32 partial-dW workspace
cudaError_t launch_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, void* dw,
                               float* dw_partial, int rows, int n, int64_t dy_row_stride, int64_t x_row_stride,
                               int64_t dx_row_stride, int x_dtype, int w_dtype, cudaStream_t stream);
cudaError_t launch_ce_fwd(const void* logits, const int64_t* labels, float* losses, float* z_losses, float* lse,
                          bool lse_is_input, int rows, int vocab, int64_t row_stride, float smoothing, float logit_scale,
                          float lse_square_scale, int64_t ignore_index, int dtype, cudaStream_t stream);
cudaError_t launch_ce_bwd(const void* logits, const int64_t* labels, const float* lse, const float* dlosses,
                          int64_t dloss_stride, void* dlogits, int rows, int vocab, int64_t row_stride,
                          int64_t dlogits_row_stride, float smoothing, float logit_scale, float lse_square_scale,
                          int64_t ignore_index, int dtype, cudaStream_t stream);

// T5 relative-position bias producer (t5_bias.cu).  lut[rel + lut_zero] = bucket, rel = mem_pos[n] - ctx_pos[m]
// (positions default to 0..M-1 / 0..N-1 when the pointers are NULL).  table: (num_buckets, H); bias: (1, H, M, N)
// contiguous; dtable: (num_buckets, H) fp32, zeroed by the launcher.
cudaError_t launch_t5_bias_fwd(const void* table, const int32_t* lut, int lut_zero, int lut_len, const int32_t* ctx_pos,
                               const int32_t* mem_pos, void* bias, int H, int M, int N, int num_buckets, int table_dtype,
                               int bias_dtype, cudaStream_t stream);
cudaError_t launch_t5_bias_bwd(const void* dbias, const int32_t* lut, int lut_zero, int lut_len, const int32_t* ctx_pos,
                               const int32_t* mem_pos, float* dtable, int H, int M, int N, int num_buckets, int dbias_dtype,
                               cudaStream_t stream);

// Band of bias values for bias mode 3 (RpeBand): band[h][j] = io_round(table[lut[clamp(band_lo + j + lut_zero)], h]).
cudaError_t launch_rpe_band(const void* table, int64_t stride_b, int64_t stride_h, int table_dtype, const int32_t* lut,
                            int lut_zero, int lut_len, float* band, int H, int band_lo, int band_len, int io_dtype,
                            cudaStream_t stream);

// (developer path) table gradient straight from the non-constant tiles of the dS group surface; zeroes dtable first
// `transposed`: ds_ws is the (G, H, N, M) surface of the v3 kernel (rows = keys) instead of (G, H, M, N)
cudaError_t launch_rpe_dtable_band(const void* ds_ws, int pitch, int G, int H, int M, int N, const int32_t* lut, int lut_zero,
                                   int lut_len, int const_lo, int const_hi, float* dtable, int num_buckets, bool causal,
                                   bool bf16, bool transposed, cudaStream_t stream);
// dtable[lut[const_lo + lut_zero], h] += dconst[h][0];  dtable[lut[const_hi + lut_zero], h] += dconst[h][1]
cudaError_t launch_rpe_dtable_add_const(float* dtable, const float* dconst, const int32_t* lut, int lut_zero, int lut_len,
                                        int const_lo, int const_hi, int H, cudaStream_t stream);

// Fused multi-tensor AdamWScale step (adamw.cu).  One descriptor per parameter tensor (device array); mirrors
// b200t5_adamw_tensor of the public header (checked with a static_assert in api.cu).
struct AdamwTensor {
    void* p;
    const void* g;
    void* m;
    void* v;
    void* comp;             // Kahan compensation, NULL unless enabled
    int64_t numel;
    int32_t first_chunk;    // index of the tensor's first chunk in the launch
    float sqrt_numel;       // (float)(numel ** 0.5)
    float ss_base;          // step size before the rms factor
    float ss_floor;         // step size when rms(p) <= 1e-3
    float neg_lr_wd;        // -(lr * weight_decay); 0 = no decay
    int32_t reserved;
};
int adamw_chunk_elems();
cudaError_t launch_adamw_step(const AdamwTensor* tensors, int n_tensors, const int32_t* chunk_tensor, int n_chunks,
                              float* chunk_sumsq, float* neg_step, int p_dtype, int state_dtype, bool kahan, float beta1,
                              float beta2, float eps, bool round_step_to_p, cudaStream_t stream);

// Launch counter (every kernel launched by this library bumps it; read through the C ABI).
void count_launch(int n = 1);

}  // namespace b200t5
