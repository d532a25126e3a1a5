// C ABI of libb200t5.so (declared in include/b200t5.h): argument validation, TMA tensor-map
// construction from runtime strides, workspace carving and kernel launches.  No allocation, no
// synchronisation, no CPU fallback.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/b200t5.h"
#include "common.cuh"
#include "kernels.h"

namespace b200t5 {

// ------------------------------------------------------------------------------------------
// error / bookkeeping
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int fail_cuda(cudaError_t e, const char* what) {
    return fail(B200T5_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------------
// optional per-kernel event timing (b200t5_profile_*)
// ------------------------------------------------------------------------------------------
struct ProfRecord {
    int id;
    cudaEvent_t start, stop;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRecord> g_prof;

struct ProfScope {
    bool active = false;
    ProfRecord rec{};
    cudaStream_t stream;
    ProfScope(int id, cudaStream_t s) : stream(s) {
        if (!g_prof_on) return;
        if (cudaEventCreate(&rec.start) != cudaSuccess) return;
        if (cudaEventCreate(&rec.stop) != cudaSuccess) {
            cudaEventDestroy(rec.start);
            return;
        }
        rec.id = id;
        active = true;
        cudaEventRecord(rec.start, stream);
    }
    ~ProfScope() {
        if (!active) return;
        cudaEventRecord(rec.stop, stream);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back(rec);
    }
};

// RAII: make `device` current for the duration of a call
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            switched = err == cudaSuccess;
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

static int device_cc(int device, int* major) {
    static std::mutex mu;
    static int cache[64];
    static bool have[64];
    if (device < 0 || device >= 64) return fail(B200T5_ERR_INVALID, "device ordinal %d out of range", device);
    std::lock_guard<std::mutex> lk(mu);
    if (!have[device]) {
        int mj = 0;
        cudaError_t e = cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, device);
        if (e != cudaSuccess) return fail_cuda(e, "cudaDeviceGetAttribute");
        cache[device] = mj;
        have[device] = true;
    }
    *major = cache[device];
    return 0;
}

static int require_sm100(int device) {
    int mj = 0;
    int rc = device_cc(device, &mj);
    if (rc) return rc;
    if (mj != 10)
        return fail(B200T5_ERR_UNSUPPORTED, "device %d has compute capability %d.x; this library is sm_100a only", device,
                    mj);
    return 0;
}

// ------------------------------------------------------------------------------------------
// TMA tensor maps
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

static CUtensorMapSwizzle swizzle_for_row_bytes(int row_bytes) {
    return row_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// 4-D map over a (d3, d2, d1, d0) tensor; d0 innermost and contiguous.  `strides` are ELEMENT strides of
// d1, d2, d3.  Size-1 dims get a synthetic (legal) stride.  box = (box0, box1, 1, 1).
static int make_map_4d(CUtensorMap* map, const void* base, int elem_bytes, CUtensorMapDataType dt, uint64_t d0,
                       uint64_t d1, uint64_t d2, uint64_t d3, int64_t s1, int64_t s2, int64_t s3, uint32_t box0,
                       uint32_t box1, const char* what, bool swizzle_only_128 = false) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(B200T5_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {d0, d1, d2, d3};
    int64_t es[3] = {s1, s2, s3};
    cuuint64_t strides[3];
    uint64_t natural = d0 * static_cast<uint64_t>(elem_bytes);
    for (int i = 0; i < 3; ++i) {
        uint64_t sb = static_cast<uint64_t>(es[i]) * elem_bytes;
        if (dims[i + 1] == 1) sb = (natural + 15) / 16 * 16;                  // never dereferenced beyond index 0
        else if (es[i] <= 0) return fail(B200T5_ERR_INVALID, "%s: stride %d is %lld for a dimension of size %llu (broadcast views must be materialised or given as a size-1 dimension)", what, i + 1, (long long)es[i], (unsigned long long)dims[i + 1]);
        if (sb % 16 != 0) return fail(B200T5_ERR_INVALID, "%s: stride %d (%lld elements) is not 16-byte aligned", what, i + 1, (long long)es[i]);
        strides[i] = sb;
        natural = sb * dims[i + 1];
    }
    if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return fail(B200T5_ERR_INVALID, "%s: base pointer is not 16-byte aligned", what);
    cuuint32_t box[4] = {box0, box1, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const int row_bytes = static_cast<int>(box0) * elem_bytes;
    const CUtensorMapSwizzle swz =
        (swizzle_only_128 && row_bytes < 128) ? CU_TENSOR_MAP_SWIZZLE_NONE : swizzle_for_row_bytes(row_bytes);
    CUresult r = enc(map, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200T5_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed with CUresult %d", what, (int)r);
    return 0;
}

static bool strides_tma_ok(const void* ptr, const int64_t* s, int b_dim, int h_dim) {
    if (reinterpret_cast<uintptr_t>(ptr) % 16 != 0) return false;
    if (s[3] != 1) return false;
    // a stride of 0 (an expanded view) on a dimension larger than 1 cannot be expressed in a tensor map
    if (s[2] % 8 != 0 || s[2] <= 0) return false;
    if (h_dim > 1 && (s[1] % 8 != 0 || s[1] <= 0)) return false;
    if (b_dim > 1 && (s[0] % 8 != 0 || s[0] <= 0)) return false;
    return true;
}

// ------------------------------------------------------------------------------------------
// attention
// ------------------------------------------------------------------------------------------
static int validate_common(const b200t5_attn_params* p) {
    if (!p) return fail(B200T5_ERR_INVALID, "params is NULL");
    if (p->B < 1 || p->H < 1 || p->M < 1 || p->N < 1) return fail(B200T5_ERR_INVALID, "B, H, M, N must be >= 1 (got %d, %d, %d, %d)", p->B, p->H, p->M, p->N);
    if (!(p->D == 16 || p->D == 32 || p->D == 64 || p->D == 128))
        return fail(B200T5_ERR_UNSUPPORTED, "head dim %d not in {16, 32, 64, 128}", p->D);   // reference :233-234
    if (!(p->dtype == B200T5_F16 || p->dtype == B200T5_BF16)) return fail(B200T5_ERR_UNSUPPORTED, "dtype %d is not fp16/bf16", p->dtype);
    if (!p->q || !p->k || !p->v || !p->o || !p->lse) return fail(B200T5_ERR_INVALID, "q, k, v, o, lse must be non-NULL");
    if (p->bias) {
        if (!(p->bias_B == 1 || p->bias_B == p->B) || !(p->bias_H == 1 || p->bias_H == p->H))
            return fail(B200T5_ERR_INVALID, "bias batch/head dims (%d, %d) must be 1 or (B, H) = (%d, %d)", p->bias_B, p->bias_H, p->B, p->H);
    }
    if ((int64_t)p->B * p->H * ((p->M + 127) / 128) > 0x7FFFFFFFLL || (int64_t)p->B * p->H * ((p->N + 127) / 128) > 0x7FFFFFFFLL)
        return fail(B200T5_ERR_INVALID, "grid too large");
    if (!strides_tma_ok(p->q, p->q_strides, p->B, p->H) || !strides_tma_ok(p->k, p->k_strides, p->B, p->H) ||
        !strides_tma_ok(p->v, p->v_strides, p->B, p->H) || !strides_tma_ok(p->o, p->o_strides, p->B, p->H))
        return fail(B200T5_ERR_INVALID, "q, k, v, o need unit last stride, 16-byte aligned base and other strides that are multiples of 8 elements");
    return 0;
}

static int bias_mode_of(const b200t5_attn_params* p) {
    if (!p->bias) return 0;
    return strides_tma_ok(p->bias, p->bias_strides, p->bias_B, p->bias_H) ? 1 : 2;
}

static int round_up8(int x) { return (x + 7) / 8 * 8; }
// forward workspace: only for a bias whose rows a tensor map cannot address (the aligned copy)
static size_t fwd_workspace_bytes(const b200t5_attn_params* p) {
    if (bias_mode_of(p) != 2) return 0;
    return ((size_t)p->bias_B * p->bias_H * p->M * (size_t)round_up8(p->N) * 2 + 255) / 256 * 256;
}
static int check_dtype3(int dt, const char* what) {
    if (dt == B200T5_F16 || dt == B200T5_BF16 || dt == B200T5_F32) return 0;
    return fail(B200T5_ERR_UNSUPPORTED, "%s dtype %d not in {fp16, bf16, fp32}", what, dt);
}

// Backward of the relative-position operator (compile-time; the A/B builds of round 2 chose level 2):
//   0: every tile stores dS, dense (1, H, M, N) scratch, producer backward folds it into the table gradient
//   1: half tiles entirely beyond a constant end of the bucket table keep their dS in the kernel (two fp32 sums per CTA)
//   2: as 1, and the table gradient is folded straight from the non-constant tiles of the dS surface (no dense scratch)
#ifndef B200T5_RPE_SKIP_LEVEL
#define B200T5_RPE_SKIP_LEVEL 2
#endif
static int rpe_skip_const_level() { return B200T5_RPE_SKIP_LEVEL; }
static bool rpe_skip_const_enabled() { return rpe_skip_const_level() != 0; }

// ---- in-kernel relative-position bias (bias mode 3) ----
static int rpe_band_len(int const_lo, int const_hi) { return const_hi - const_lo + 2 * kRpeBandPad + 1; }

static int validate_rpe(const b200t5_rpe_params* r, bool need_table, bool need_bwd) {
    if (!r) return fail(B200T5_ERR_INVALID, "rpe params is NULL");
    if (r->const_lo >= r->const_hi) return fail(B200T5_ERR_INVALID, "rpe: const_lo (%d) must be < const_hi (%d)", r->const_lo, r->const_hi);
    if ((int64_t)r->const_hi - r->const_lo > kRpeMaxBandLen) return fail(B200T5_ERR_UNSUPPORTED, "rpe: %lld distinct relative positions between const_lo and const_hi; at most %d fit the kernels (use the dense-bias path)", (long long)r->const_hi - r->const_lo, kRpeMaxBandLen - 2 * kRpeBandPad - 1);
    if (rpe_band_len(r->const_lo, r->const_hi) > kRpeMaxBandLen) return fail(B200T5_ERR_UNSUPPORTED, "rpe: band of %d relative positions exceeds %d (use the dense-bias path)", rpe_band_len(r->const_lo, r->const_hi), kRpeMaxBandLen);
    if (!r->band) return fail(B200T5_ERR_INVALID, "rpe: band buffer is NULL");
    if (reinterpret_cast<uintptr_t>(r->band) % 4 != 0) return fail(B200T5_ERR_INVALID, "rpe: band buffer is not 4-byte aligned");
    if (need_table || need_bwd) {
        if (!r->lut || r->lut_len < 1) return fail(B200T5_ERR_INVALID, "rpe: lut is NULL or empty");
        if (r->num_buckets < 1 || r->num_buckets > 256) return fail(B200T5_ERR_UNSUPPORTED, "rpe: num_buckets %d not in [1, 256]", r->num_buckets);
    }
    if (need_table && !r->table) return fail(B200T5_ERR_INVALID, "rpe: table is NULL");
    if (need_bwd && !r->dtable) return fail(B200T5_ERR_INVALID, "rpe: dtable is NULL");
    return 0;
}

static void fill_rpe_band(RpeBand* rb, const b200t5_rpe_params* r) {
    rb->band = r->band;
    rb->const_lo = r->const_lo;
    rb->const_hi = r->const_hi;
    rb->band_lo = r->const_lo - kRpeBandPad;
    rb->band_len = rpe_band_len(r->const_lo, r->const_hi);
}

}  // namespace b200t5

using namespace b200t5;

static int attn_fwd_impl(const b200t5_attn_params* p, const b200t5_rpe_params* rpe) {
    int rc = validate_common(p);
    if (rc) return rc;
    if (rpe) {
        if (p->bias) return fail(B200T5_ERR_INVALID, "the relative-position entry points take bias == NULL");
        if ((rc = validate_rpe(rpe, false, false))) return rc;
    }
    if ((rc = require_sm100(p->device))) return rc;
    DeviceGuard guard(p->device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");

    const CUtensorMapDataType dt = p->dtype == B200T5_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const uint32_t boxd = p->D >= 64 ? 64 : p->D;
    AttnFwdKernelParams kp;
    memset(&kp, 0, sizeof(kp));
    if ((rc = make_map_4d(&kp.map_q, p->q, 2, dt, p->D, p->M, p->H, p->B, p->q_strides[2], p->q_strides[1], p->q_strides[0], boxd, 128, "q"))) return rc;
    if ((rc = make_map_4d(&kp.map_k, p->k, 2, dt, p->D, p->N, p->H, p->B, p->k_strides[2], p->k_strides[1], p->k_strides[0], boxd, 128, "k"))) return rc;
    if ((rc = make_map_4d(&kp.map_v, p->v, 2, dt, p->D, p->N, p->H, p->B, p->v_strides[2], p->v_strides[1], p->v_strides[0], boxd, 128, "v"))) return rc;
    if (p->D == 64) {
        if (!strides_tma_ok(p->o, p->o_strides, p->B, p->H)) return fail(B200T5_ERR_INVALID, "o needs unit last stride, 16-byte aligned base and other strides that are multiples of 8 elements");
        if ((rc = make_map_4d(&kp.map_o, p->o, 2, dt, p->D, p->M, p->H, p->B, p->o_strides[2], p->o_strides[1], p->o_strides[0], boxd, 128, "o"))) return rc;
    }
    int mode = rpe ? 3 : bias_mode_of(p);
    if (mode == 2 && p->workspace && p->workspace_bytes >= fwd_workspace_bytes(p) && reinterpret_cast<uintptr_t>(p->workspace) % 256 == 0) {
        // rows a tensor map cannot address: aligned copy in the caller's workspace, then the TMA path
        const int pitch = round_up8(p->N);
        cudaError_t ce = launch_bias_align_copy(p->bias, p->bias_strides, p->workspace, p->bias_B, p->bias_H, p->M, p->N, pitch,
                                                static_cast<cudaStream_t>(p->stream));
        if (ce != cudaSuccess) return fail_cuda(ce, "bias_align_copy launch");
        if ((rc = make_map_4d(&kp.map_bias, p->workspace, 2, dt, p->N, p->M, p->bias_H, p->bias_B, pitch, (int64_t)p->M * pitch, (int64_t)p->bias_H * p->M * pitch, 64, 128, "bias (aligned copy)"))) return rc;
        mode = 1;
    } else if (mode == 1) {
        if ((rc = make_map_4d(&kp.map_bias, p->bias, 2, dt, p->N, p->M, p->bias_H, p->bias_B, p->bias_strides[2], p->bias_strides[1], p->bias_strides[0], 64, 128, "bias"))) return rc;
    } else if (mode == 2) {
        kp.bias = p->bias;
        kp.bias_sb = p->bias_strides[0];
        kp.bias_sh = p->bias_strides[1];
        kp.bias_sm = p->bias_strides[2];
        kp.bias_sn = p->bias_strides[3];
    } else if (mode == 3) {
        fill_rpe_band(&kp.rpe, rpe);
    }
    kp.o = p->o;
    kp.o_sb = p->o_strides[0];
    kp.o_sh = p->o_strides[1];
    kp.o_sm = p->o_strides[2];
    kp.lse = p->lse;
    kp.B = p->B; kp.H = p->H; kp.M = p->M; kp.N = p->N;
    kp.num_m_blocks = (p->M + 127) / 128;
    kp.bias_b_bcast = p->bias ? (p->bias_B == 1) : 1;
    kp.bias_h_bcast = p->bias ? (p->bias_H == 1) : 1;
    kp.sm_scale = p->sm_scale;
    cudaError_t e;
    {
        ProfScope prof(B200T5_KERNEL_ATTN_FWD, static_cast<cudaStream_t>(p->stream));
        const bool bf16 = p->dtype == B200T5_BF16, causal = p->causal != 0;
        cudaStream_t stream = static_cast<cudaStream_t>(p->stream);
        e = launch_attn_fwd(kp, p->D, bf16, mode, causal, stream);
    }
    if (e != cudaSuccess) return fail_cuda(e, "attn_fwd launch");
    return 0;
}

extern "C" int b200t5_attn_fwd(const b200t5_attn_params* p) { return attn_fwd_impl(p, nullptr); }
extern "C" size_t b200t5_attn_fwd_workspace_bytes(const b200t5_attn_params* p) {
    if (!p || p->B < 1 || p->H < 1 || p->M < 1 || p->N < 1 || p->D < 1 || !p->bias) return 0;
    return fwd_workspace_bytes(p);
}
extern "C" int b200t5_attn_rpe_fwd(const b200t5_attn_params* p, const b200t5_rpe_params* r) {
    if (!r) return fail(B200T5_ERR_INVALID, "rpe params is NULL");
    return attn_fwd_impl(p, r);
}

namespace {
struct BwdWorkspace {
    size_t delta_off, dq_off, ds_off, ds_bytes, bias_t_off, dbias_off, dconst_off, total;
    int ds_pitch, ds_groups, ds_use_reduce, dq_groups;
    bool transposed;        // D <= 64: the v3 kernel (dS surface and bias copy are (.., N, M)); D = 128: (.., M, N)
};
// has_rpe: the bias is the in-kernel relative-position bias, i.e. a (1, H, M, N) bias whose dense gradient is only
// an intermediate (kept in the workspace and folded into the (num_buckets, H) table gradient).
BwdWorkspace bwd_workspace_layout(const b200t5_attn_params* p, bool has_rpe = false, bool rpe_skip = false) {
    BwdWorkspace w;
    auto align = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t rows = (size_t)p->B * p->H * p->M;
    w.transposed = p->D <= 64;
    w.ds_pitch = w.transposed ? round_up8(p->M) : round_up8(p->N);
    w.delta_off = 0;
    // D = 128: delta (B, H, M).  v3 kernel: -L * log2e and -delta, each (B, H, M rounded up to 128)
    const size_t m_pad = (size_t)(p->M + 127) / 128 * 128;
    w.dq_off = w.transposed ? 2 * align((size_t)p->B * p->H * m_pad * sizeof(float)) : align(rows * sizeof(float));
    // dQ group surface: <= 4 key blocks accumulate (16-bit, at L2) into one group; <= 8 groups
    const int nnb = (p->N + 127) / 128;
    const bool deterministic = (p->flags & B200T5_ATTN_DETERMINISTIC) != 0;
    int gq = (nnb + 3) / 4;
    if (gq > 8) gq = 8;
    if (deterministic) gq = nnb;           // one key block per group: a single addend per slot, summed in fixed order later
    w.dq_groups = gq;
    w.ds_off = w.dq_off + align((size_t)gq * rows * p->D * 2);
    // dS surface: per-batch bias -> one slice per batch (plain stores); batch-broadcast bias -> the batch is
    // folded into ds_groups slices of <= 8 (<= B/16 for huge B) batches each by TMA reduce-add in the bias dtype,
    // and the slices are summed in fp32 afterwards.  Keeps the surface L2-sized (67 MB at the headline shape
    // instead of 537 MB) while bounding the 16-bit accumulation depth.
    w.ds_groups = 0;
    w.ds_use_reduce = 0;
    size_t ds_bytes = 0;
    if (p->bias || has_rpe) {
        // (with the constant-tile skip of the relative-position path some tiles are never written: the surface must be
        //  the zero-filled, reduce-added kind even for a single batch)
        if (((has_rpe || p->bias_B == 1) && p->B > 1) || rpe_skip) {
            int g = (p->B + 7) / 8;
            if (g > 16) g = 16;
            if (deterministic) g = p->B;   // one batch element per group (still zero-filled + reduce-added: unwritten tiles)
            w.ds_groups = g;
            w.ds_use_reduce = 1;
        } else {
            w.ds_groups = p->B;
        }
        ds_bytes = (size_t)w.ds_groups * p->H * (size_t)(w.transposed ? p->N : p->M) * (size_t)w.ds_pitch * 2;
    }
    w.ds_bytes = ds_bytes;
    w.bias_t_off = w.ds_off + align(ds_bytes);
    // repacked dense bias of the v3 kernel: [bias_B][bias_H][key blocks of 128][query blocks of 32][4][128][8] 16-bit
    size_t bias_t_bytes = (w.transposed && p->bias) ? (size_t)p->bias_B * p->bias_H * (size_t)((p->N + 127) / 128) * (size_t)(4 * ((p->M + 127) / 128)) * 4 * 128 * 16 : 0;   // whole 128-query tiles: the kernel reads every sub-tile of its last tile
    // D = 128 with bias rows a tensor map cannot address: the same region holds the aligned copy (rows padded to 8 elements)
    if (!w.transposed && p->bias && bias_mode_of(p) == 2) bias_t_bytes = fwd_workspace_bytes(p);
    w.dbias_off = w.bias_t_off + align(bias_t_bytes);
    const bool dense_scratch = has_rpe && !(w.transposed && rpe_skip && rpe_skip_const_level() >= 2);
    w.dconst_off = w.dbias_off + (dense_scratch ? align((size_t)p->H * p->M * (size_t)p->N * 2) : 0);
    w.total = w.dconst_off + (has_rpe ? align((size_t)p->H * 2 * sizeof(float)) : 0);
    return w;
}
}  // namespace

extern "C" size_t b200t5_attn_bwd_workspace_bytes(const b200t5_attn_params* p) {
    if (!p || p->B < 1 || p->H < 1 || p->M < 1 || p->N < 1 || p->D < 1) return 0;
    return bwd_workspace_layout(p).total;
}

extern "C" size_t b200t5_attn_rpe_bwd_workspace_bytes(const b200t5_attn_params* p, const b200t5_rpe_params* r) {
    if (!p || !r || p->B < 1 || p->H < 1 || p->M < 1 || p->N < 1 || p->D < 1) return 0;
    return bwd_workspace_layout(p, true, rpe_skip_const_enabled() && p->D <= 64).total;
}

static int attn_bwd_impl(const b200t5_attn_params* p, const b200t5_rpe_params* rpe) {
    int rc = validate_common(p);
    if (rc) return rc;
    if (!p->dout || !p->dq || !p->dk || !p->dv) return fail(B200T5_ERR_INVALID, "dout, dq, dk, dv must be non-NULL");
    if ((p->bias != nullptr) != (p->dbias != nullptr)) return fail(B200T5_ERR_INVALID, "dbias must be given exactly when bias is");
    if (rpe) {
        if (p->bias) return fail(B200T5_ERR_INVALID, "the relative-position entry points take bias == NULL and dbias == NULL");
        if ((rc = validate_rpe(rpe, false, true))) return rc;
        if (rpe->lut_zero < p->M - 1 || rpe->lut_len - 1 - rpe->lut_zero < p->N - 1)
            return fail(B200T5_ERR_INVALID, "rpe: lut (zero %d, len %d) does not cover relative positions %d..%d", rpe->lut_zero, rpe->lut_len, -(p->M - 1), p->N - 1);
    }
    if (!strides_tma_ok(p->dout, p->do_strides, p->B, p->H) || !strides_tma_ok(p->dq, p->dq_strides, p->B, p->H) ||
        !strides_tma_ok(p->dk, p->dk_strides, p->B, p->H) || !strides_tma_ok(p->dv, p->dv_strides, p->B, p->H))
        return fail(B200T5_ERR_INVALID, "dout, dq, dk, dv need unit last stride, 16-byte aligned base and other strides that are multiples of 8 elements");
    if ((p->flags & B200T5_ATTN_DBIAS_ACCUMULATE) && !(p->flags & B200T5_ATTN_DBIAS_F32)) return fail(B200T5_ERR_INVALID, "B200T5_ATTN_DBIAS_ACCUMULATE needs B200T5_ATTN_DBIAS_F32 (the accumulator is an fp32 tensor)");
    if ((p->flags & B200T5_ATTN_DBIAS_F32) && (p->D > 64 || rpe)) return fail(B200T5_ERR_UNSUPPORTED, "B200T5_ATTN_DBIAS_F32 is implemented for head dims 16 / 32 / 64 of the dense-bias operator");
    const bool rpe_skip = rpe != nullptr && rpe_skip_const_enabled() && p->D <= 64;   // the D = 128 kernel stores every tile
    const BwdWorkspace w = bwd_workspace_layout(p, rpe != nullptr, rpe_skip);
    if (!p->workspace || p->workspace_bytes < w.total) return fail(B200T5_ERR_WORKSPACE, "workspace of %zu bytes needed, %zu given", w.total, p->workspace ? p->workspace_bytes : (size_t)0);
    if (reinterpret_cast<uintptr_t>(p->workspace) % 256 != 0) return fail(B200T5_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    if ((rc = require_sm100(p->device))) return rc;
    DeviceGuard guard(p->device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");

    cudaStream_t stream = static_cast<cudaStream_t>(p->stream);
    const bool bf16 = p->dtype == B200T5_BF16;
    const bool causal = p->causal != 0;
    uint8_t* ws = static_cast<uint8_t*>(p->workspace);
    float* delta = reinterpret_cast<float*>(ws + w.delta_off);
    void* dq_ws = ws + w.dq_off;
    void* ds_ws = ws + w.ds_off;
    const int G = w.ds_groups > 0 ? w.ds_groups : 1;

    // delta, zero-fill of the dQ group surface and (when dS is reduce-added) of the dS group surface -- and, for the v3 kernel
    // with a dense bias, the repacked bias copy -- in one launch
    const int m_pad = (p->M + 127) / 128 * 128;
    float* nl = delta;                                                              // v3: records of [64 x nl | 64 x -delta] per 64 padded rows
    float* ndelta = reinterpret_cast<float*>(ws + w.dq_off / 2);
    cudaError_t e;
    if (w.transposed)
        e = launch_attn_bwd_pre_fused(p->o, p->o_strides, p->dout, p->do_strides, p->lse, nl, ndelta, m_pad, dq_ws, w.dq_groups, p->B, p->H, p->M,
                                      p->N, p->D, bf16, w.ds_use_reduce ? ds_ws : nullptr, w.ds_use_reduce ? (w.ds_bytes + 15) / 16 * 16 : 0,
                                      rpe ? nullptr : p->bias, p->bias_strides, ws + w.bias_t_off, p->bias_B, p->bias_H, stream);
    else
        e = launch_attn_bwd_preprocess(p->o, p->o_strides, p->dout, p->do_strides, delta, dq_ws, w.dq_groups, p->B, p->H, p->M, p->D, bf16,
                                       w.ds_use_reduce ? ds_ws : nullptr, w.ds_use_reduce ? (w.ds_bytes + 15) / 16 * 16 : 0, stream);
    if (e != cudaSuccess) return fail_cuda(e, "attn_bwd_preprocess launch");

    const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const uint32_t boxd = p->D >= 64 ? 64 : p->D;
    AttnBwdKernelParams kp;
    memset(&kp, 0, sizeof(kp));
    const uint32_t qrows = w.transposed ? 64 : 128;         // the v3 kernel loads Q / dO per 64-query half tile
    if ((rc = make_map_4d(&kp.map_q, p->q, 2, dt, p->D, p->M, p->H, p->B, p->q_strides[2], p->q_strides[1], p->q_strides[0], boxd, qrows, "q"))) return rc;
    if ((rc = make_map_4d(&kp.map_k, p->k, 2, dt, p->D, p->N, p->H, p->B, p->k_strides[2], p->k_strides[1], p->k_strides[0], boxd, 128, "k"))) return rc;
    // (v3: V travels through the Q / dO ring as two 64-key halves)
    if ((rc = make_map_4d(&kp.map_v, p->v, 2, dt, p->D, p->N, p->H, p->B, p->v_strides[2], p->v_strides[1], p->v_strides[0], boxd, w.transposed ? 64 : 128, "v"))) return rc;
    if ((rc = make_map_4d(&kp.map_do, p->dout, 2, dt, p->D, p->M, p->H, p->B, p->do_strides[2], p->do_strides[1], p->do_strides[0], boxd, qrows, "dout"))) return rc;
    if ((rc = make_map_4d(&kp.map_dq, dq_ws, 2, dt, p->D, p->M, p->H, (uint64_t)w.dq_groups * p->B, p->D, (int64_t)p->M * p->D, (int64_t)p->H * p->M * p->D, boxd, 128, "dq group surface", true))) return rc;
    kp.dq_groups = w.dq_groups;
    if (w.transposed && p->D == 64) {
        // dK / dV leave the v3 kernel through TMA stores of [128 keys][32 columns] boxes staged in shared memory (thread = key row
        // stores straight to global memory cost one L1 request per 16 bytes: ~2 500 cycles per work item)
        if ((rc = make_map_4d(&kp.map_dk_st, p->dk, 2, dt, p->D, p->N, p->H, p->B, p->dk_strides[2], p->dk_strides[1], p->dk_strides[0], 32, 128, "dk"))) return rc;
        if ((rc = make_map_4d(&kp.map_dv_st, p->dv, 2, dt, p->D, p->N, p->H, p->B, p->dv_strides[2], p->dv_strides[1], p->dv_strides[0], 32, 128, "dv"))) return rc;
    }
    // D <= 64: every dense bias goes through the transposed copy (which also absorbs unaligned rows); D = 128: TMA or pointers
    int mode = rpe ? 3 : (p->bias ? (w.transposed ? 1 : bias_mode_of(p)) : 0);
    float* dconst = nullptr;
    if (mode == 3) {
        fill_rpe_band(&kp.rpe, rpe);
        if (rpe_skip) {
            dconst = reinterpret_cast<float*>(ws + w.dconst_off);
            e = cudaMemsetAsync(dconst, 0, (size_t)p->H * 2 * sizeof(float), stream);
            if (e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync(dconst)");
            kp.rpe.dconst = dconst;
        }
    }
    if (w.transposed) {
        if (mode == 1) kp.bias = ws + w.bias_t_off;          // the repacked copy written by the fused pre-kernel above
        if (mode != 0) {
            if ((rc = make_map_4d(&kp.map_ds, ds_ws, 2, dt, p->M, p->N, p->H, G, w.ds_pitch, (int64_t)p->N * w.ds_pitch, (int64_t)p->H * p->N * w.ds_pitch, 32, 128, "dS workspace (transposed)"))) return rc;
        }
        kp.k = p->k; kp.k_sb = p->k_strides[0]; kp.k_sh = p->k_strides[1]; kp.k_sn = p->k_strides[2];
        kp.v = p->v; kp.v_sb = p->v_strides[0]; kp.v_sh = p->v_strides[1]; kp.v_sn = p->v_strides[2];
    } else {
        if (mode == 2) {
            // unaligned bias rows: aligned copy in the workspace, then the TMA path (as in the forward)
            const int pitch = round_up8(p->N);
            e = launch_bias_align_copy(p->bias, p->bias_strides, ws + w.bias_t_off, p->bias_B, p->bias_H, p->M, p->N, pitch, stream);
            if (e != cudaSuccess) return fail_cuda(e, "bias_align_copy launch");
            if ((rc = make_map_4d(&kp.map_bias, ws + w.bias_t_off, 2, dt, p->N, p->M, p->bias_H, p->bias_B, pitch, (int64_t)p->M * pitch, (int64_t)p->bias_H * p->M * pitch, 64, 128, "bias (aligned copy)"))) return rc;
            mode = 1;
        } else if (mode == 1) {
            if ((rc = make_map_4d(&kp.map_bias, p->bias, 2, dt, p->N, p->M, p->bias_H, p->bias_B, p->bias_strides[2], p->bias_strides[1], p->bias_strides[0], 64, 128, "bias"))) return rc;
        }
        if (mode != 0) {
            // dS tiles always go to the (B, H, M, n_pad) 16-bit workspace through TMA stores
            if ((rc = make_map_4d(&kp.map_ds, ds_ws, 2, dt, p->N, p->M, p->H, G, w.ds_pitch, (int64_t)p->M * w.ds_pitch, (int64_t)p->H * p->M * w.ds_pitch, 64, 128, "dS workspace"))) return rc;
        }
    }
    kp.ds_groups = G;
    kp.ds_use_reduce = w.ds_use_reduce;
    kp.dk = p->dk; kp.dk_sb = p->dk_strides[0]; kp.dk_sh = p->dk_strides[1]; kp.dk_sn = p->dk_strides[2];
    kp.dv = p->dv; kp.dv_sb = p->dv_strides[0]; kp.dv_sh = p->dv_strides[1]; kp.dv_sn = p->dv_strides[2];
    kp.lse = p->lse;
    kp.delta = delta;
    kp.nl = nl;
    kp.ndelta = ndelta;
    kp.m_pad = m_pad;
    kp.B = p->B; kp.H = p->H; kp.M = p->M; kp.N = p->N;
    kp.num_m_blocks = (p->M + 127) / 128;
    kp.num_n_blocks = (p->N + 127) / 128;
    kp.bias_b_bcast = p->bias ? (p->bias_B == 1) : 1;
    kp.bias_h_bcast = p->bias ? (p->bias_H == 1) : (rpe ? 0 : 1);
    kp.sm_scale = p->sm_scale;
    {
        ProfScope prof(B200T5_KERNEL_ATTN_BWD, stream);
        e = w.transposed ? launch_attn_bwd_v3(kp, p->D, bf16, mode, causal, stream) : launch_attn_bwd(kp, p->D, bf16, mode, causal, stream);
    }
    if (e != cudaSuccess) return fail_cuda(e, "attn_bwd launch");

    if (rpe && dconst && w.transposed && rpe_skip_const_level() >= 2) {
        // dQ conversion alone, then the table gradient straight from the non-constant tiles of the surface
        e = launch_attn_bwd_dq_convert(dq_ws, w.dq_groups, p->dq, p->dq_strides, p->B, p->H, p->M, p->D, p->sm_scale, bf16, stream);
        if (e != cudaSuccess) return fail_cuda(e, "attn_bwd_dq_convert launch");
        e = launch_rpe_dtable_band(ds_ws, w.ds_pitch, G, p->H, p->M, p->N, rpe->lut, rpe->lut_zero, rpe->lut_len,
                                   rpe->const_lo, rpe->const_hi, rpe->dtable, rpe->num_buckets, causal, bf16, true, stream);
        if (e != cudaSuccess) return fail_cuda(e, "rpe_dtable_band launch");
        e = launch_rpe_dtable_add_const(rpe->dtable, dconst, rpe->lut, rpe->lut_zero, rpe->lut_len, rpe->const_lo, rpe->const_hi, p->H, stream);
        if (e != cudaSuccess) return fail_cuda(e, "rpe_dtable_add_const launch");
        return 0;
    }
    // the dense gradient: dBias itself, or (relative-position operator) a (1, H, M, N) scratch in the workspace that the
    // producer's segmented sum folds into the (num_buckets, H) table gradient
    void* dbias_out = rpe ? static_cast<void*>(ws + w.dbias_off) : (mode != 0 ? p->dbias : nullptr);
    const int64_t scratch_strides[4] = {(int64_t)p->H * p->M * p->N, (int64_t)p->M * p->N, p->N, 1};
    const int64_t* dbias_strides = rpe ? scratch_strides : p->dbias_strides;
    const int reduce_b = rpe ? 1 : (p->bias_B == 1), reduce_h = rpe ? 0 : (p->bias_H == 1);
    if (w.transposed) {
        e = launch_attn_bwd_post_fused(dq_ws, w.dq_groups, p->dq, p->dq_strides, p->B, p->H, p->M, p->N, p->D, p->sm_scale, bf16, ds_ws,
                                       w.ds_pitch, dbias_out, dbias_strides, G, reduce_b, reduce_h, causal,
                                       !rpe && (p->flags & B200T5_ATTN_DBIAS_F32) != 0, stream,
                                       !rpe && (p->flags & B200T5_ATTN_DBIAS_ACCUMULATE) != 0);
        if (e != cudaSuccess) return fail_cuda(e, "attn_bwd finalize launch");
    } else {
        e = launch_attn_bwd_finalize(dq_ws, w.dq_groups, p->dq, p->dq_strides, p->B, p->H, p->M, p->N, p->D, p->sm_scale, bf16,
                                     ds_ws, w.ds_pitch, dbias_out, dbias_strides, G, reduce_b, reduce_h, causal, stream);
        if (e != cudaSuccess) return fail_cuda(e, "attn_bwd_finalize launch");
    }
    if (rpe) {
        e = launch_t5_bias_bwd(dbias_out, rpe->lut, rpe->lut_zero, rpe->lut_len, nullptr, nullptr, rpe->dtable, p->H, p->M, p->N,
                               rpe->num_buckets, p->dtype, stream);
        if (e != cudaSuccess) return fail_cuda(e, "t5_bias_bwd launch");
        if (dconst) {
            e = launch_rpe_dtable_add_const(rpe->dtable, dconst, rpe->lut, rpe->lut_zero, rpe->lut_len, rpe->const_lo, rpe->const_hi, p->H, stream);
            if (e != cudaSuccess) return fail_cuda(e, "rpe_dtable_add_const launch");
        }
    }
    return 0;
}

extern "C" int b200t5_attn_bwd(const b200t5_attn_params* p) { return attn_bwd_impl(p, nullptr); }
extern "C" int b200t5_attn_rpe_bwd(const b200t5_attn_params* p, const b200t5_rpe_params* r) {
    if (!r) return fail(B200T5_ERR_INVALID, "rpe params is NULL");
    return attn_bwd_impl(p, r);
}

extern "C" int b200t5_rpe_band_len(int32_t const_lo, int32_t const_hi) {
    if (const_lo >= const_hi) return 0;
    const int64_t n = (int64_t)const_hi - const_lo + 2 * kRpeBandPad + 1;
    return n > 0x7FFFFFFF ? 0 : (int)n;
}

extern "C" int b200t5_rpe_band(const b200t5_rpe_params* r, int32_t H, int io_dtype, int device, void* stream) {
    int rc = validate_rpe(r, true, false);
    if (rc) return rc;
    if (H < 1) return fail(B200T5_ERR_INVALID, "H must be >= 1");
    if (!(io_dtype == B200T5_F16 || io_dtype == B200T5_BF16)) return fail(B200T5_ERR_UNSUPPORTED, "io dtype %d is not fp16/bf16", io_dtype);
    if ((rc = check_dtype3(r->table_dtype, "table"))) return rc;
    if ((rc = require_sm100(device))) return rc;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    cudaError_t e = launch_rpe_band(r->table, r->table_stride_b, r->table_stride_h, r->table_dtype, r->lut, r->lut_zero, r->lut_len,
                                    r->band, H, r->const_lo - kRpeBandPad, rpe_band_len(r->const_lo, r->const_hi), io_dtype,
                                    static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "rpe_band launch");
    return 0;
}

// ------------------------------------------------------------------------------------------
// RMSNorm / cross-entropy
// ------------------------------------------------------------------------------------------

extern "C" int b200t5_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int64_t rows, int64_t n,
                                  int64_t x_row_stride, int64_t y_row_stride, float eps, int x_dtype, int w_dtype,
                                  int device, void* stream) {
    if (!x || !w || !y || !rstd) return fail(B200T5_ERR_INVALID, "x, w, y, rstd must be non-NULL");
    if (rows < 0 || n < 1 || n > 65536 || rows > 0x7FFFFFFF) return fail(B200T5_ERR_INVALID, "bad rows/n (%lld, %lld)", (long long)rows, (long long)n);
    int rc;
    if ((rc = check_dtype3(x_dtype, "x")) || (rc = check_dtype3(w_dtype, "w"))) return rc;
    if ((rc = require_sm100(device))) return rc;
    if (rows == 0) return 0;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    cudaError_t e = launch_rmsnorm_fwd(x, w, y, rstd, (int)rows, (int)n, x_row_stride, y_row_stride, eps, x_dtype, w_dtype, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "rmsnorm_fwd launch");
    return 0;
}

extern "C" size_t b200t5_rmsnorm_bwd_workspace_bytes(int64_t n) {
    if (n < 1) return 0;
    return (size_t)kRmsnormMaxPartials * (size_t)n * sizeof(float);
}

extern "C" int b200t5_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, void* dw,
                                  void* workspace, size_t workspace_bytes, int64_t rows, int64_t n,
                                  int64_t dy_row_stride, int64_t x_row_stride, int64_t dx_row_stride, int x_dtype,
                                  int w_dtype, int device, void* stream) {
    if (!dy || !x || !w || !rstd || !dx || !dw) return fail(B200T5_ERR_INVALID, "dy, x, w, rstd, dx, dw must be non-NULL");
    if (rows < 0 || n < 1 || n > 65536 || rows > 0x7FFFFFFF) return fail(B200T5_ERR_INVALID, "bad rows/n (%lld, %lld)", (long long)rows, (long long)n);
    int rc;
    if ((rc = check_dtype3(x_dtype, "x")) || (rc = check_dtype3(w_dtype, "w"))) return rc;
    const size_t need = b200t5_rmsnorm_bwd_workspace_bytes(n);
    if (!workspace || workspace_bytes < need) return fail(B200T5_ERR_WORKSPACE, "workspace of %zu bytes needed", need);
    if ((rc = require_sm100(device))) return rc;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    cudaError_t e = launch_rmsnorm_bwd(dy, x, w, rstd, dx, dw, static_cast<float*>(workspace), (int)rows, (int)n, dy_row_stride, x_row_stride, dx_row_stride, x_dtype, w_dtype, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "rmsnorm_bwd launch");
    return 0;
}

extern "C" int b200t5_ce_fwd(const void* logits, const int64_t* labels, float* losses, float* z_losses, float* lse,
                             int lse_is_input, int64_t rows, int64_t vocab, int64_t row_stride, float smoothing,
                             float logit_scale, float lse_square_scale, int64_t ignore_index, int dtype, int device,
                             void* stream) {
    if (!logits || !labels || !losses || !z_losses || !lse) return fail(B200T5_ERR_INVALID, "logits, labels, losses, z_losses, lse must be non-NULL");
    if (rows < 0 || vocab < 1 || rows > 0x7FFFFFFF || vocab > 0x7FFFFFFF) return fail(B200T5_ERR_INVALID, "bad rows/vocab");
    if (lse_is_input && (smoothing != 0.f || logit_scale != 1.f)) return fail(B200T5_ERR_INVALID, "a precomputed lse needs smoothing == 0 and logit_scale == 1");
    int rc;
    if ((rc = check_dtype3(dtype, "logits"))) return rc;
    if ((rc = require_sm100(device))) return rc;
    if (rows == 0) return 0;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    cudaError_t e = launch_ce_fwd(logits, labels, losses, z_losses, lse, lse_is_input != 0, (int)rows, (int)vocab, row_stride, smoothing, logit_scale, lse_square_scale, ignore_index, dtype, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "ce_fwd launch");
    return 0;
}

extern "C" int b200t5_ce_bwd(const void* logits, const int64_t* labels, const float* lse, const float* dlosses,
                             int64_t dloss_stride, void* dlogits, int64_t rows, int64_t vocab, int64_t row_stride,
                             int64_t dlogits_row_stride, float smoothing, float logit_scale, float lse_square_scale,
                             int64_t ignore_index, int dtype, int device, void* stream) {
    if (!logits || !labels || !lse || !dlosses || !dlogits) return fail(B200T5_ERR_INVALID, "logits, labels, lse, dlosses, dlogits must be non-NULL");
    if (rows < 0 || vocab < 1 || rows > 0x7FFFFFFF || vocab > 0x7FFFFFFF) return fail(B200T5_ERR_INVALID, "bad rows/vocab");
    int rc;
    if ((rc = check_dtype3(dtype, "logits"))) return rc;
    if ((rc = require_sm100(device))) return rc;
    if (rows == 0) return 0;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    cudaError_t e = launch_ce_bwd(logits, labels, lse, dlosses, dloss_stride, dlogits, (int)rows, (int)vocab, row_stride, dlogits_row_stride, smoothing, logit_scale, lse_square_scale, ignore_index, dtype, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "ce_bwd launch");
    return 0;
}

// ------------------------------------------------------------------------------------------
// T5 relative-position bias producer
// ------------------------------------------------------------------------------------------
static int check_t5_args(const void* a, const void* b, const int32_t* lut, int lut_len, int H, int M, int N, int nb) {
    if (!a || !b || !lut) return fail(B200T5_ERR_INVALID, "table/dbias, bias/dtable and lut must be non-NULL");
    if (H < 1 || M < 1 || N < 1 || lut_len < 1) return fail(B200T5_ERR_INVALID, "H, M, N, lut_len must be >= 1");
    if (nb < 1 || nb > 256) return fail(B200T5_ERR_UNSUPPORTED, "num_buckets %d not in [1, 256]", nb);
    return 0;
}

extern "C" int b200t5_t5_bias_fwd(const void* table, const int32_t* lut, int32_t lut_zero, int32_t lut_len,
                                  const int32_t* ctx_pos, const int32_t* mem_pos, void* bias, int32_t H, int32_t M,
                                  int32_t N, int32_t num_buckets, int table_dtype, int bias_dtype, int device,
                                  void* stream) {
    int rc;
    if ((rc = check_t5_args(table, bias, lut, lut_len, H, M, N, num_buckets))) return rc;
    if ((rc = check_dtype3(table_dtype, "table")) || (rc = check_dtype3(bias_dtype, "bias"))) return rc;
    if ((rc = require_sm100(device))) return rc;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    cudaError_t e = launch_t5_bias_fwd(table, lut, lut_zero, lut_len, ctx_pos, mem_pos, bias, H, M, N, num_buckets, table_dtype, bias_dtype, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "t5_bias_fwd launch");
    return 0;
}

extern "C" int b200t5_t5_bias_bwd(const void* dbias, const int32_t* lut, int32_t lut_zero, int32_t lut_len,
                                  const int32_t* ctx_pos, const int32_t* mem_pos, float* dtable, int32_t H, int32_t M,
                                  int32_t N, int32_t num_buckets, int dbias_dtype, int device, void* stream) {
    int rc;
    if ((rc = check_t5_args(dbias, dtable, lut, lut_len, H, M, N, num_buckets))) return rc;
    if ((rc = check_dtype3(dbias_dtype, "dbias"))) return rc;
    if ((rc = require_sm100(device))) return rc;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    cudaError_t e = launch_t5_bias_bwd(dbias, lut, lut_zero, lut_len, ctx_pos, mem_pos, dtable, H, M, N, num_buckets, dbias_dtype, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "t5_bias_bwd launch");
    return 0;
}

// ------------------------------------------------------------------------------------------
// fused multi-tensor AdamWScale step
// ------------------------------------------------------------------------------------------
static_assert(sizeof(b200t5_adamw_tensor) == sizeof(AdamwTensor) && offsetof(b200t5_adamw_tensor, neg_lr_wd) == offsetof(AdamwTensor, neg_lr_wd) &&
              offsetof(b200t5_adamw_tensor, numel) == offsetof(AdamwTensor, numel), "AdamwTensor must mirror b200t5_adamw_tensor");

extern "C" int b200t5_adamw_chunk_elems(void) { return adamw_chunk_elems(); }

extern "C" size_t b200t5_adamw_workspace_bytes(int32_t n_tensors, int32_t n_chunks) {
    if (n_tensors < 0 || n_chunks < 0) return 0;
    auto align = [](size_t x) { return (x + 255) / 256 * 256; };
    return align((size_t)n_chunks * sizeof(float)) + align((size_t)n_tensors * sizeof(float));
}

extern "C" int b200t5_adamw_scale_step(const b200t5_adamw_tensor* tensors, int32_t n_tensors, const int32_t* chunk_tensor,
                                       int32_t n_chunks, void* workspace, size_t workspace_bytes, int p_dtype,
                                       int state_dtype, int kahan, float beta1, float beta2, float eps,
                                       int round_step_to_p, int device, void* stream) {
    if (n_tensors < 0 || n_chunks < 0) return fail(B200T5_ERR_INVALID, "negative tensor / chunk count");
    if (n_tensors == 0 || n_chunks == 0) return 0;
    if (!tensors || !chunk_tensor) return fail(B200T5_ERR_INVALID, "tensors and chunk_tensor must be non-NULL");
    int rc;
    if ((rc = check_dtype3(p_dtype, "parameter")) || (rc = check_dtype3(state_dtype, "state"))) return rc;
    if (kahan && p_dtype == B200T5_F32) return fail(B200T5_ERR_INVALID, "Kahan compensation applies to 16-bit parameters only (reference :107-111)");
    if (p_dtype != B200T5_F32 && state_dtype != p_dtype) return fail(B200T5_ERR_UNSUPPORTED, "16-bit parameters keep their states in the parameter dtype");
    if (!(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f) || !(eps >= 0.f)) return fail(B200T5_ERR_INVALID, "betas must be in [0, 1) and eps >= 0");
    const size_t need = b200t5_adamw_workspace_bytes(n_tensors, n_chunks);
    if (!workspace || workspace_bytes < need) return fail(B200T5_ERR_WORKSPACE, "workspace of %zu bytes needed", need);
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return fail(B200T5_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    if ((rc = require_sm100(device))) return rc;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail_cuda(guard.err, "cudaSetDevice");
    float* chunk_sumsq = static_cast<float*>(workspace);
    float* neg_step = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + ((size_t)n_chunks * sizeof(float) + 255) / 256 * 256);
    cudaError_t e = launch_adamw_step(reinterpret_cast<const AdamwTensor*>(tensors), n_tensors, chunk_tensor, n_chunks, chunk_sumsq,
                                      neg_step, p_dtype, state_dtype, kahan != 0, beta1, beta2, eps, round_step_to_p != 0,
                                      static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda(e, "adamw_step launch");
    return 0;
}

// ------------------------------------------------------------------------------------------
// library state
// ------------------------------------------------------------------------------------------
extern "C" int b200t5_abi_version(void) { return B200T5_ABI_VERSION; }
extern "C" const char* b200t5_last_error(void) { return g_err; }
extern "C" uint64_t b200t5_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int b200t5_profile_enable(int enable) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = enable != 0;
    if (!g_prof_on) {
        for (auto& r : g_prof) {
            cudaEventDestroy(r.start);
            cudaEventDestroy(r.stop);
        }
        g_prof.clear();
    }
    return 0;
}
extern "C" int b200t5_profile_collect(int* kernel_ids, float* ms, int cap) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    int n = 0;
    for (auto& r : g_prof) {
        float t = -1.f;
        cudaError_t e = cudaEventElapsedTime(&t, r.start, r.stop);
        if (e != cudaSuccess) {
            cudaGetLastError();
            t = -1.f;                       // caller did not synchronise
        }
        if (n < cap && kernel_ids && ms) {
            kernel_ids[n] = r.id;
            ms[n] = t;
            ++n;
        }
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    g_prof.clear();
    return n;
}
extern "C" int b200t5_device_supported(int device) {
    int mj = 0;
    int rc = device_cc(device, &mj);
    if (rc) return rc;
    return mj == 10 ? 1 : 0;
}
