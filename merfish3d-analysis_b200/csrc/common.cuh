// common.cuh -- context, error handling and launch accounting for libm3d_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/m3d_b200.h"

#define M3D_NUM_SMS_DEFAULT 148

// ---------------------------------------------------------------- errors
extern thread_local std::string g_m3d_error;
int m3d_fail(int code, const char* fmt, ...);

#define M3D_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return m3d_fail(M3D_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,     \
                            cudaGetErrorString(e__));                                      \
    } while (0)

#define M3D_CHECK_LAUNCH() M3D_CUDA(cudaGetLastError())

// ---------------------------------------------------------------- kernel families (launch accounting)
enum M3dKernel {
    KF_WEIGHT = 0,
    KF_LOWPASS_Z,
    KF_LOWPASS_YX,
    KF_LOWPASS_GENERIC,
    KF_DECODE_GATE,
    KF_DECODE_SEARCH,
    KF_DECODE_DENSE,
    KF_CCL_COLLECT,
    KF_CCL_INIT,
    KF_CCL_MERGE,
    KF_CCL_COMPRESS,
    KF_CCL_SELECT,
    KF_CCL_SORT,
    KF_CCL_ASSIGN,
    KF_CCL_SCAN,
    KF_CCL_SCATTER,
    KF_CCL_INTERFACE,
    KF_FEATURES,
    KF_SELECT_HIST,
    KF_REPLACE_ABOVE,
    KF_WARP_AFFINE,
    KF_TABLE_HIST,
    KF_TABLE_GRID,
    KF_TABLE_OVERLAP,
    KF_TABLE_WITHIN,
    KF_CENTROID,
    KF_EIGVALS,
    KF_TABLE_CELLS,
    KF_RESET_FG,
    KF_ZARR_UNSHUFFLE,
    KF_ZARR_FILL,
    KF_ZARR_LZ4,
    KF_ZARR_ZSTD,
    KF_COUNT
};

// ---------------------------------------------------------------- scratch buffer
struct Scratch {
    void* ptr = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        // grow with headroom so tile-to-tile size jitter does not reallocate
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e != cudaSuccess) {
            ptr = nullptr;
            return m3d_fail(M3D_ERR_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return 0;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

// ---------------------------------------------------------------- per-call decode parameters (by value)
// Exact-arithmetic parameters.  Without normalisation vectors bkg = 0 and nrm = 1 (identity).
struct DecodeParams {
    float bkg[M3D_MAX_BITS];
    float nrm[M3D_MAX_BITS];
    float rcp[M3D_MAX_BITS];  // RN(1 / nrm) when the reciprocal form of the division is proven exact, else 0 (voxel_math.cuh)
    int rcp_all;              // every live bit has rcp != 0
    float pix_thr, mag_lo, mag_hi;
    int n_bits;
    int K;
    int mode;          // 0: generic rows (direct scan); 1: every row's non-zeros share one value
                       // (proxy scan); 2: mode 1 + one on-bit count and value for all rows
                       // (top-w selection + hash lookup, warp-cooperative fallback)
    int max_on;        // padded on-bit list length (modes 1, 2); the on-bit count w in mode 2
    float cval;        // mode 2: the common non-zero entry (1/sqrt(w))
    int hash_bits;     // mode 2: log2 of the mask -> codeword hash table size
    const float* codebook;       // K x M3D_MAX_BITS fp32 (zero padded)
    const uint8_t* onbits;       // K x max_on bit indices (pad = zero slot)
    const float* cw_a;           // K: ||c_k||^2
    const float* cw_g;           // K: 2*c_k
    const float* cw_c;           // K: c_k (value of the non-zero entries)
    const uint32_t* cw_mask;     // K: on-bit mask
    const uint8_t* excluded;     // K flags
    const uint32_t* hash_keys;   // 2^hash_bits on-bit masks (0 = empty)
    const int16_t* hash_vals;    // 2^hash_bits codeword indices
};

// Parameters of the conservative streaming gate (decode_gate_kernel).
struct GateParams {
    float bkg[M3D_MAX_BITS];
    float rcp[M3D_MAX_BITS];  // ~1/nrm; 0 for padding bits
    float lo2, hi2;           // squared-magnitude window, widened by the proven error margin
    int n_bits;
    int all_candidates;       // 1: vectors outside the margin proof -> every voxel goes to the exact kernel
};

struct m3d_ctx {
    int device = 0;
    int num_sms = M3D_NUM_SMS_DEFAULT;
    int n_bits = 0;
    int nb_pad = 0;  // 16 / 24 / 32
    int K = 0;
    int max_on = 0;
    int mode = 0;
    float cval = 0.f;
    int hash_bits = 0;
    // device-side codebook artefacts
    float* d_codebook = nullptr;
    uint8_t* d_onbits = nullptr;
    float* d_cw_a = nullptr;
    float* d_cw_g = nullptr;
    float* d_cw_c = nullptr;
    uint32_t* d_cw_mask = nullptr;
    uint8_t* d_excluded = nullptr;
    uint32_t* d_hash_keys = nullptr;
    int16_t* d_hash_vals = nullptr;
    // normalisation state (identity when use_norm == 0)
    float bkg[M3D_MAX_BITS];
    float nrm[M3D_MAX_BITS];
    int use_norm = 0;
    float pix_thr = 0.f, mag_lo = 0.f, mag_hi = 0.f;
    // scratch
    Scratch s_cand;      // decode candidates
    Scratch s_rec_x;     // float16 scaled values of foreground voxels (search -> features hand-off)
    Scratch s_rec_md;    // float16 magnitude | distance of foreground voxels
    Scratch s_counters;  // small device counters
    Scratch s_lp_tmp;    // low-pass intermediate volume
    Scratch s_fg;        // CCL foreground list
    Scratch s_parent;    // dense union-find parents
    Scratch s_aux;       // dense per-root aux (area / id)
    Scratch s_roots;     // surviving roots (unsorted, sorted)
    Scratch s_area;      // area by id, offsets, cursors
    Scratch s_vox;       // voxel lists grouped by component
    Scratch s_sort;      // cub temp storage
    unsigned long long* h_pinned = nullptr;  // pinned host mailbox for counts
    // label state (valid between m3d_label and m3d_features)
    int64_t lab_dims[3] = {0, 0, 0};
    int64_t lab_n_fg = 0;
    int64_t lab_n_features = -1;
    int lab_max_px = 0;
    int lab_rec_valid = 0;
    size_t sparse_cap_override = 0;  // m3d_set_sparse_capacity (tests); 0 = automatic  // records of the last m3d_decode_label cover every foreground voxel
    // persistent decoded image (m3d_decode_label_persistent): the previous call's foreground list is still in
    // s_fg, so the next call resets those voxels instead of re-filling the whole image with -1
    const void* prev_decoded = nullptr;
    size_t prev_n_vox = 0;
    int prev_fg_valid = 0;
    int gate_skip_background = 0;  // set around the gate launch only
    int lowpass_f32 = 0;           // m3d_set_lowpass_mode: 1 = float32 FMA accumulation (opt-in, lowpass.cu)
    // accounting
    int64_t launches[KF_COUNT];
    // optional per-kernel-family device timing (m3d_set_timing): event pairs recorded on the
    // launch stream around every launch, resolved lazily by m3d_kernel_time_ms
    int timing = 0;
    struct TimedSpan { int k; cudaEvent_t a, b; };
    std::vector<TimedSpan> spans;
    std::mutex span_mu;  // launches may come from the prefetch thread as well
    double time_ms[KF_COUNT];
    DecodeParams params() const;
    GateParams gate_params() const;
};

// Scoped launch accounting: construct right before a launch (or a library call that launches),
// destruct right after.  Counts the launch and, when timing is on, brackets it with events.
struct KernelScope {
    m3d_ctx* ctx;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    int k;
    KernelScope(m3d_ctx* c, M3dKernel kf, cudaStream_t s, int64_t n = 1) : ctx(c), st(s), k((int)kf) {
        __atomic_fetch_add(&ctx->launches[k], n, __ATOMIC_RELAXED);  // the prefetch thread launches too
        if (ctx->timing) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, st);
        }
    }
    ~KernelScope() {
        if (a) {
            cudaEventRecord(b, st);
            std::lock_guard<std::mutex> lk(ctx->span_mu);
            ctx->spans.push_back({k, a, b});
        }
    }
};

#define M3D_LAUNCH(ctx, kf, st, ...)            \
    do {                                        \
        KernelScope ks__((ctx), (kf), (st));    \
        __VA_ARGS__;                            \
    } while (0)

// capacity (entries) of the sparse hand-off buffers: ample for real data (a few % foreground),
// bounded so that degenerate all-foreground inputs cannot exhaust HBM -- overflow entries take the
// recompute paths
static inline size_t m3d_sparse_capacity(const m3d_ctx* ctx, size_t n_vox) {
    if (ctx->sparse_cap_override) return ctx->sparse_cap_override;
    size_t c = n_vox / 16;
    if (c < ((size_t)1 << 20)) c = (size_t)1 << 20;
    return c < n_vox ? c : n_vox;
}

static inline int ceil_div_i64(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
