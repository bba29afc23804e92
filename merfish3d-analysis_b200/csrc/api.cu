// api.cu -- context lifetime, normalisation state, order-statistic helpers (C ABI).
#include <math.h>
#include <stdarg.h>

#include "common.cuh"

thread_local std::string g_m3d_error;

int m3d_fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_m3d_error = buf;
    return code;
}

static const char* const kKernelNames[KF_COUNT] = {
    "weight_kernel",       "lowpass_z_kernel",    "lowpass_yx_kernel",   "lowpass_axis_generic_kernel",
    "decode_gate_kernel",  "decode_search_kernel", "decode_dense_kernel", "ccl_collect_kernel",
    "ccl_init_kernel",     "ccl_merge_kernel",    "ccl_compress_kernel", "ccl_select_kernel",
    "cub_radix_sort",      "ccl_assign_kernel",   "cub_exclusive_sum",   "ccl_scatter_kernel",
    "ccl_interface_kernel",   "features_kernel",     "select_hist_kernel",  "replace_above_kernel", "warp_affine_kernel",
    "table_hist3d_kernel", "table_grid_kernels", "table_overlap_kernel", "table_within_kernels",
    "centroid_stats_kernel", "inertia_eigvals_kernel", "assign_cells_kernel", "reset_foreground_kernel",
    "zarr_unshuffle_place_kernel", "zarr_fill_chunk_kernel", "blosc_lz4_decode_kernel", "blosc_zstd_decode_kernel"};

DecodeParams m3d_ctx::params() const {
    DecodeParams P;
    memset(&P, 0, sizeof(P));
    for (int b = 0; b < M3D_MAX_BITS; ++b) {
        const bool live = use_norm && b < n_bits;
        P.bkg[b] = live ? bkg[b] : 0.f;
        P.nrm[b] = live ? nrm[b] : 1.f;
        // reciprocal form of (s - bkg) / nrm (voxel_math.cuh, div_by_rcp): only for ordinary vectors -- |nrm| in
        // [2^-40, 2^40] with a significand that is not all ones, bkg = 0 or |bkg| in [2^-20, 2^30] (so that an integer
        // sample minus bkg is 0 or at least 2^-24 in magnitude); anything else keeps the IEEE division (rcp = 0)
        const float nm = P.nrm[b], bg = P.bkg[b];
        uint32_t nbits;
        memcpy(&nbits, &nm, 4);
        const float an = fabsf(nm), ab = fabsf(bg);
        const bool ok = an >= 9.0949470177e-13f && an <= 1.0995116278e+12f && (nbits & 0x7FFFFFu) != 0x7FFFFFu &&
                        (bg == 0.f || (ab >= 9.5367431640625e-07f && ab <= 1073741824.0f));
        P.rcp[b] = ok ? 1.0f / nm : 0.f;  // float division on the host: correctly rounded
    }
    P.rcp_all = 1;
    for (int b = 0; b < n_bits; ++b)
        if (P.rcp[b] == 0.f) P.rcp_all = 0;
    P.pix_thr = pix_thr;
    P.mag_lo = mag_lo;
    P.mag_hi = mag_hi;
    P.n_bits = n_bits;
    P.K = K;
    P.mode = mode;
    P.max_on = max_on;
    P.cval = cval;
    P.hash_bits = hash_bits;
    P.codebook = d_codebook;
    P.onbits = d_onbits;
    P.cw_a = d_cw_a;
    P.cw_g = d_cw_g;
    P.cw_c = d_cw_c;
    P.cw_mask = d_cw_mask;
    P.excluded = d_excluded;
    P.hash_keys = d_hash_keys;
    P.hash_vals = d_hash_vals;
    return P;
}

// Conservative gate window.  The streaming kernel evaluates q = clamp((s - bkg) * rcp, 0, 1)
// and acc = sum fma(q, q, acc): (s - bkg) is the reference's own first rounding, the
// reciprocal multiply is within 1.5 ulp of the IEEE quotient, so acc is within ~3e-6
// (relative) of the reference's squared norm; the exact kernel compares sqrt(acc_exact)
// against the thresholds, one more half-ulp.  A 1e-5 relative widening is >3x that bound.
// Outside the proof's preconditions (tiny/huge/non-finite vectors, tiny thresholds) every
// voxel is sent to the exact kernel instead.
GateParams m3d_ctx::gate_params() const {
    GateParams G;
    memset(&G, 0, sizeof(G));
    G.n_bits = n_bits;
    int all = 0;
    for (int b = 0; b < n_bits; ++b) {
        const float bg = use_norm ? bkg[b] : 0.f;
        const float nm = use_norm ? nrm[b] : 1.f;
        const float an = fabsf(nm);
        if (!(an >= 1.0e-30f && an <= 1.0e30f) || !(fabsf(bg) <= 3.0e38f)) all = 1;
        G.bkg[b] = bg;
        G.rcp[b] = (float)(1.0 / (double)nm);
    }
    const double lo = (double)mag_lo, hi = (double)mag_hi;
    if (!(lo >= 1.0e-3) || !(hi >= lo) || !(lo <= 1.0e15)) all = 1;  // also catches NaN thresholds
    G.lo2 = (float)(lo * lo * (1.0 - 1.0e-5));
    const double h2 = hi * hi * (1.0 + 1.0e-5);
    G.hi2 = h2 > 3.0e38 ? __builtin_inff() : (float)h2;
    G.all_candidates = all;
    return G;
}

extern "C" int m3d_abi_version(void) { return M3D_ABI_VERSION; }
extern "C" const char* m3d_last_error(void) { return g_m3d_error.c_str(); }

extern "C" int m3d_create(int device, int n_bits, int n_codewords, const float* codebook_unit_host,
                          const int32_t* excluded_host, int n_excluded, m3d_ctx** out) {
    if (!out || !codebook_unit_host) return m3d_fail(M3D_ERR_ARG, "m3d_create: null argument");
    if (n_bits < 1 || n_bits > M3D_MAX_BITS) return m3d_fail(M3D_ERR_ARG, "m3d_create: n_bits must be in [1, %d]", M3D_MAX_BITS);
    if (n_codewords < 1 || n_codewords > 32767) return m3d_fail(M3D_ERR_ARG, "m3d_create: n_codewords must be in [1, 32767] (int16 decoded image)");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev < 1)
        return m3d_fail(M3D_ERR_CUDA, "m3d_create: no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return m3d_fail(M3D_ERR_ARG, "m3d_create: device %d of %d", device, n_dev);
    M3D_CUDA(cudaSetDevice(device));
    m3d_ctx* ctx = new m3d_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    M3D_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    ctx->n_bits = n_bits;
    ctx->nb_pad = n_bits <= 8 ? 8 : (n_bits <= 16 ? 16 : (n_bits <= 24 ? 24 : 32));
    ctx->K = n_codewords;
    memset(ctx->launches, 0, sizeof(ctx->launches));
    for (int i = 0; i < KF_COUNT; ++i) ctx->time_ms[i] = 0.0;
    for (int b = 0; b < M3D_MAX_BITS; ++b) {
        ctx->bkg[b] = 0.f;
        ctx->nrm[b] = 1.f;
    }
    const int K = n_codewords;
    // padded codebook + structure analysis
    std::vector<float> cb((size_t)K * M3D_MAX_BITS, 0.f);
    std::vector<float> cw_a(K), cw_g(K), cw_c(K);
    std::vector<uint32_t> mask(K, 0u);
    int binary = 1, uniform = 1, max_on = 1, first_on = -1;
    float first_c = 0.f;
    for (int k = 0; k < K; ++k) {
        float cval = 0.f;
        int n_on = 0;
        double nn = 0.0;
        for (int b = 0; b < n_bits; ++b) {
            float v = codebook_unit_host[(size_t)k * n_bits + b];
            cb[(size_t)k * M3D_MAX_BITS + b] = v;
            nn += (double)v * (double)v;
            if (v != 0.f) {
                if (n_on == 0) cval = v;
                else if (v != cval) binary = 0;
                mask[k] |= (1u << b);
                ++n_on;
            }
            if (!(v == v)) binary = 0;
        }
        if (n_on > max_on) max_on = n_on;
        if (k == 0) {
            first_on = n_on;
            first_c = cval;
        } else if (n_on != first_on || cval != first_c) {
            uniform = 0;
        }
        cw_a[k] = (float)nn;
        cw_g[k] = 2.f * cval;
        cw_c[k] = cval;
    }
    if (max_on > 16) binary = 0;  // proxy path keeps short on-bit lists only
    // mode 2 needs positive entries (S ordering) and at least one off bit per row
    if (!binary || first_on < 1 || first_on >= n_bits || !(first_c > 0.f)) uniform = 0;
    ctx->mode = binary ? (uniform ? 2 : 1) : 0;
    ctx->max_on = max_on;
    ctx->cval = first_c;
    std::vector<uint8_t> on((size_t)K * max_on, (uint8_t)ctx->nb_pad);  // pad -> zero slot
    for (int k = 0; k < K; ++k) {
        int j = 0;
        for (int b = 0; b < n_bits && j < max_on; ++b)
            if (mask[k] & (1u << b)) on[(size_t)k * max_on + j++] = (uint8_t)b;
    }
    // mask -> first codeword index, open addressing (mode 2)
    int hash_bits = 4;
    while ((1 << hash_bits) < 4 * K) ++hash_bits;
    ctx->hash_bits = hash_bits;
    std::vector<uint32_t> hkeys((size_t)1 << hash_bits, 0u);
    std::vector<int16_t> hvals((size_t)1 << hash_bits, (int16_t)-1);
    if (ctx->mode == 2) {
        const uint32_t hm = (1u << hash_bits) - 1u;
        for (int k = 0; k < K; ++k) {
            uint32_t h = (mask[k] * 2654435761u) >> (32 - hash_bits);
            while (hkeys[h] != 0u && hkeys[h] != mask[k]) h = (h + 1u) & hm;
            if (hkeys[h] == 0u) {
                hkeys[h] = mask[k];
                hvals[h] = (int16_t)k;  // duplicates keep the first (lowest) index, like argmin
            }
        }
    }
    std::vector<uint8_t> excl(K, 0);
    for (int i = 0; i < n_excluded; ++i) {
        if (!excluded_host) break;
        int idx = excluded_host[i];
        if (idx < 0 || idx >= K) {
            delete ctx;
            return m3d_fail(M3D_ERR_ARG, "m3d_create: excluded index %d out of range", idx);
        }
        excl[idx] = 1;
    }
#define UP(dst, vec, T)                                                                          \
    M3D_CUDA(cudaMalloc((void**)&ctx->dst, vec.size() * sizeof(T)));                            \
    M3D_CUDA(cudaMemcpy(ctx->dst, vec.data(), vec.size() * sizeof(T), cudaMemcpyHostToDevice));
    UP(d_codebook, cb, float)
    UP(d_onbits, on, uint8_t)
    UP(d_cw_a, cw_a, float)
    UP(d_cw_g, cw_g, float)
    UP(d_cw_c, cw_c, float)
    UP(d_cw_mask, mask, uint32_t)
    UP(d_excluded, excl, uint8_t)
    UP(d_hash_keys, hkeys, uint32_t)
    UP(d_hash_vals, hvals, int16_t)
#undef UP
    M3D_CUDA(cudaMallocHost((void**)&ctx->h_pinned, 64));
    *out = ctx;
    return M3D_OK;
}

void m3d_release_upload_ring(m3d_ctx* ctx);  // upload.cu
void m3d_release_zarr_ring(m3d_ctx* ctx);    // zarrio.cu

extern "C" int m3d_destroy(m3d_ctx* ctx) {
    if (!ctx) return M3D_OK;
    cudaSetDevice(ctx->device);
    m3d_release_upload_ring(ctx);
    m3d_release_zarr_ring(ctx);
    for (auto& sp : ctx->spans) {
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    ctx->spans.clear();
    cudaFree(ctx->d_codebook);
    cudaFree(ctx->d_onbits);
    cudaFree(ctx->d_cw_a);
    cudaFree(ctx->d_cw_g);
    cudaFree(ctx->d_cw_c);
    cudaFree(ctx->d_cw_mask);
    cudaFree(ctx->d_excluded);
    cudaFree(ctx->d_hash_keys);
    cudaFree(ctx->d_hash_vals);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    ctx->s_cand.release();
    ctx->s_rec_x.release();
    ctx->s_rec_md.release();
    ctx->s_counters.release();
    ctx->s_lp_tmp.release();
    ctx->s_fg.release();
    ctx->s_parent.release();
    ctx->s_aux.release();
    ctx->s_roots.release();
    ctx->s_area.release();
    ctx->s_vox.release();
    ctx->s_sort.release();
    delete ctx;
    return M3D_OK;
}

extern "C" int m3d_set_normalization(m3d_ctx* ctx, const float* background_host, const float* normalization_host) {
    if (!ctx) return m3d_fail(M3D_ERR_ARG, "m3d_set_normalization: null ctx");
    ctx->lab_rec_valid = 0;
    if (!background_host || !normalization_host) {
        ctx->use_norm = 0;
        return M3D_OK;
    }
    ctx->use_norm = 1;
    for (int b = 0; b < ctx->n_bits; ++b) {
        ctx->bkg[b] = background_host[b];
        ctx->nrm[b] = normalization_host[b];
    }
    return M3D_OK;
}

extern "C" int m3d_set_thresholds(m3d_ctx* ctx, float pixel_threshold, float magnitude_lo, float magnitude_hi) {
    if (!ctx) return m3d_fail(M3D_ERR_ARG, "m3d_set_thresholds: null ctx");
    ctx->lab_rec_valid = 0;
    ctx->pix_thr = pixel_threshold;
    ctx->mag_lo = magnitude_lo;
    ctx->mag_hi = magnitude_hi;
    return M3D_OK;
}

extern "C" int m3d_set_sparse_capacity(m3d_ctx* ctx, int64_t entries) {
    if (!ctx || entries < 0) return m3d_fail(M3D_ERR_ARG, "m3d_set_sparse_capacity: bad argument");
    ctx->sparse_cap_override = (size_t)entries;
    ctx->lab_rec_valid = 0;
    return M3D_OK;
}

extern "C" int64_t m3d_launch_count(m3d_ctx* ctx) {
    if (!ctx) return 0;
    int64_t n = 0;
    for (int i = 0; i < KF_COUNT; ++i) n += ctx->launches[i];
    return n;
}

extern "C" const char* m3d_kernel_name(int i) { return (i >= 0 && i < KF_COUNT) ? kKernelNames[i] : nullptr; }
static void resolve_spans(m3d_ctx* ctx) {
    std::lock_guard<std::mutex> lk(ctx->span_mu);
    for (auto& sp : ctx->spans) {
        cudaEventSynchronize(sp.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) ctx->time_ms[sp.k] += (double)ms;
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    ctx->spans.clear();
}
extern "C" int m3d_set_timing(m3d_ctx* ctx, int enable) {
    if (!ctx) return m3d_fail(M3D_ERR_ARG, "m3d_set_timing: null ctx");
    cudaSetDevice(ctx->device);
    resolve_spans(ctx);
    ctx->timing = enable ? 1 : 0;
    return M3D_OK;
}
extern "C" int m3d_reset_counters(m3d_ctx* ctx) {
    if (!ctx) return m3d_fail(M3D_ERR_ARG, "m3d_reset_counters: null ctx");
    cudaSetDevice(ctx->device);
    resolve_spans(ctx);
    memset(ctx->launches, 0, sizeof(ctx->launches));
    for (int i = 0; i < KF_COUNT; ++i) ctx->time_ms[i] = 0.0;
    return M3D_OK;
}
extern "C" double m3d_kernel_time_ms(m3d_ctx* ctx, int i) {
    if (!ctx || i < 0 || i >= KF_COUNT) return 0.0;
    cudaSetDevice(ctx->device);
    resolve_spans(ctx);
    return ctx->time_ms[i];
}
extern "C" int64_t m3d_kernel_launches(m3d_ctx* ctx, int i) {
    return (ctx && i >= 0 && i < KF_COUNT) ? ctx->launches[i] : 0;
}

// ------------------------------------------------------------------ order statistics (PD:1113-1177)
namespace {

__device__ __forceinline__ uint32_t ordered_key(float v) {
    uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// One digit histogram of the radix select: 4 bytes per element and pass, so the pass should run at HBM speed -- and the
// optimiser's percentile seed makes ~26 passes over every filtered bit volume.  The first version (scalar loop, runtime
// predicate / clip / prefix switches, one shared-memory atomic per element) ran at a third of that: it was bound by its
// ~20 instructions per element, not by atomics (warp-aggregating them with match.any changed nothing: 231 -> 241 ms for
// 2 tiles).  Now: the switches are template parameters, loads are 128-bit with two in flight per thread, and
//   * passes WITH a prefix (digits 2 and 3) only test `(key & mask) == value` per element -- the few matching elements
//     take the atomic in a rare branch;
//   * passes WITHOUT a prefix (first digit, counts) count every element: image data lands in a handful of bins there, so
//     each thread run-length-accumulates its consecutive elements and flushes a run with one atomic.
template <int PRED, bool CLIP, bool PREFIX>
__global__ void __launch_bounds__(256, 8)
select_hist_kernel(const float* __restrict__ data, size_t n, float sub, float cutoff, uint32_t prefix_mask,
                   uint32_t prefix_value, int shift, unsigned long long* __restrict__ hist) {
    __shared__ unsigned int s_hist[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) s_hist[i] = 0u;
    __syncthreads();
    uint32_t run_bin = 0xFFFFFFFFu, run_len = 0u;  // !PREFIX: the current run of equal bins
    auto add = [&](float raw) {
        float v = __fsub_rn(raw, sub);
        if (CLIP && v < 0.f) v = 0.f;
        bool ok = true;
        if (PRED == 1) ok = v < cutoff;
        else if (PRED == 2) ok = v > cutoff;
        const uint32_t key = ordered_key(v);
        if (PREFIX) {
            if (ok && (key & prefix_mask) == prefix_value) atomicAdd(&s_hist[(key >> shift) & 2047u], 1u);
        } else {
            const uint32_t bin = ok ? ((key >> shift) & 2047u) : 0xFFFFFFFFu;
            if (bin == run_bin) {
                ++run_len;
            } else {
                if (run_bin != 0xFFFFFFFFu) atomicAdd(&s_hist[run_bin], run_len);
                run_bin = bin;
                run_len = 1u;
            }
        }
    };
    const size_t stride = (size_t)gridDim.x * 256;
    const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t n4 = ((reinterpret_cast<uintptr_t>(data) & 15u) == 0) ? n / 4 : 0;
    const float4* data4 = reinterpret_cast<const float4*>(data);
    size_t i = tid;
    for (; i + stride < n4; i += 2 * stride) {
        const float4 q0 = __ldcs(data4 + i), q1 = __ldcs(data4 + i + stride);
        add(q0.x);
        add(q0.y);
        add(q0.z);
        add(q0.w);
        add(q1.x);
        add(q1.y);
        add(q1.z);
        add(q1.w);
    }
    if (i < n4) {
        const float4 q0 = __ldcs(data4 + i);
        add(q0.x);
        add(q0.y);
        add(q0.z);
        add(q0.w);
    }
    for (size_t j = 4 * n4 + tid; j < n; j += stride) add(__ldcs(data + j));  // unaligned volumes, the last n % 4 elements
    if (!PREFIX && run_bin != 0xFFFFFFFFu) atomicAdd(&s_hist[run_bin], run_len);
    __syncthreads();
    for (int k = threadIdx.x; k < 2048; k += 256)
        if (s_hist[k]) atomicAdd(hist + k, (unsigned long long)s_hist[k]);
}

// The same digit histogram for up to SELECT_BATCH small multisets at once (the optimiser's 2 x bits per-iteration medians:
// ~25 k values each): blockIdx.y = query, the blocks of a query stride over its values.  One launch per level instead of
// one per multiset -- the kernels were microseconds each, the launches and the ctypes calls were the cost.
constexpr int SELECT_BATCH = 128;
struct SelectBatch {
    const float* data[SELECT_BATCH];
    long long n[SELECT_BATCH];
    uint32_t prefix_mask[SELECT_BATCH];
    uint32_t prefix_value[SELECT_BATCH];
};

__global__ void __launch_bounds__(256)
select_hist_batch_kernel(const __grid_constant__ SelectBatch B, int shift, unsigned long long* __restrict__ hist) {
    __shared__ unsigned int s_hist[2048];
    const int q = blockIdx.y;
    const size_t n = (size_t)B.n[q];
    if ((size_t)blockIdx.x * 256 >= n) return;  // uniform per block
    for (int i = threadIdx.x; i < 2048; i += 256) s_hist[i] = 0u;
    __syncthreads();
    const float* __restrict__ data = B.data[q];
    const uint32_t pm = B.prefix_mask[q], pv = B.prefix_value[q];
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float v = __ldg(data + i);
        const uint32_t key = ordered_key(v);
        if (v == v && (key & pm) == pv) atomicAdd(&s_hist[(key >> shift) & 2047u], 1u);  // NaN = no entry
    }
    __syncthreads();
    unsigned long long* out = hist + (size_t)q * 2048;
    for (int i = threadIdx.x; i < 2048; i += 256)
        if (s_hist[i]) atomicAdd(out + i, (unsigned long long)s_hist[i]);
}

__global__ void __launch_bounds__(256)
replace_above_kernel(float* __restrict__ data, size_t n, float thr, float value) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        if (data[i] > thr) data[i] = value;
}

}  // namespace

extern "C" int m3d_select_hist(m3d_ctx* ctx, const float* data_dev, int64_t n, float sub, int clip0, int pred,
                               float cutoff, uint32_t prefix_mask, uint32_t prefix_value, int shift,
                               unsigned long long* hist_dev, void* stream) {
    if (!ctx || !data_dev || !hist_dev || n < 0) return m3d_fail(M3D_ERR_ARG, "m3d_select_hist: bad argument");
    if (shift < 0 || shift > 31) return m3d_fail(M3D_ERR_ARG, "m3d_select_hist: shift %d", shift);
    if (n == 0) return M3D_OK;
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    size_t want = ((size_t)n + 255) / 256;
    size_t cap = (size_t)ctx->num_sms * 8;
    int blocks = (int)(want < cap ? want : cap);
    const bool prefix = prefix_mask != 0u;
    const bool clip = clip0 != 0;
#define M3D_SELECT_CASE(P_, C_, X_)                                                                                  \
    if (pred == P_ && clip == C_ && prefix == X_) {                                                                  \
        M3D_LAUNCH(ctx, KF_SELECT_HIST, st,                                                                           \
                   (select_hist_kernel<P_, C_, X_><<<blocks, 256, 0, st>>>(data_dev, (size_t)n, sub, cutoff, prefix_mask, \
                                                                          prefix_value, shift, hist_dev)));          \
        launched = true;                                                                                             \
    }
    bool launched = false;
    M3D_SELECT_CASE(0, false, false)
    M3D_SELECT_CASE(0, false, true)
    M3D_SELECT_CASE(0, true, false)
    M3D_SELECT_CASE(0, true, true)
    M3D_SELECT_CASE(1, false, false)
    M3D_SELECT_CASE(1, false, true)
    M3D_SELECT_CASE(1, true, false)
    M3D_SELECT_CASE(1, true, true)
    M3D_SELECT_CASE(2, false, false)
    M3D_SELECT_CASE(2, false, true)
    M3D_SELECT_CASE(2, true, false)
    M3D_SELECT_CASE(2, true, true)
#undef M3D_SELECT_CASE
    if (!launched) return m3d_fail(M3D_ERR_ARG, "m3d_select_hist: pred %d", pred);
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

extern "C" int m3d_select_hist_batch(m3d_ctx* ctx, int n_queries, const float* const* data_dev, const int64_t* n,
                                     const uint32_t* prefix_mask, const uint32_t* prefix_value, int shift,
                                     unsigned long long* hist_dev, void* stream) {
    if (!ctx || n_queries < 0 || !hist_dev || (n_queries > 0 && (!data_dev || !n || !prefix_mask || !prefix_value)))
        return m3d_fail(M3D_ERR_ARG, "m3d_select_hist_batch: bad argument");
    if (shift < 0 || shift > 31) return m3d_fail(M3D_ERR_ARG, "m3d_select_hist_batch: shift %d", shift);
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    for (int q0 = 0; q0 < n_queries; q0 += SELECT_BATCH) {
        const int nq = n_queries - q0 < SELECT_BATCH ? n_queries - q0 : SELECT_BATCH;
        SelectBatch B;
        memset(&B, 0, sizeof(B));
        long long n_max = 0;
        for (int j = 0; j < nq; ++j) {
            if (n[q0 + j] < 0 || (n[q0 + j] > 0 && !data_dev[q0 + j]))
                return m3d_fail(M3D_ERR_ARG, "m3d_select_hist_batch: query %d", q0 + j);
            B.data[j] = data_dev[q0 + j];
            B.n[j] = n[q0 + j];
            B.prefix_mask[j] = prefix_mask[q0 + j];
            B.prefix_value[j] = prefix_value[q0 + j];
            if (B.n[j] > n_max) n_max = B.n[j];
        }
        if (n_max == 0) continue;
        long long bx = (n_max + 256 * 16 - 1) / (256 * 16);  // ~16 values per thread
        const long long cap = (long long)ctx->num_sms * 8;
        if (bx > cap) bx = cap;
        M3D_LAUNCH(ctx, KF_SELECT_HIST, st,
                   select_hist_batch_kernel<<<dim3((unsigned)bx, (unsigned)nq), 256, 0, st>>>(B, shift, hist_dev + (size_t)q0 * 2048));
        M3D_CHECK_LAUNCH();
    }
    return M3D_OK;
}

extern "C" int m3d_replace_above(m3d_ctx* ctx, float* data_dev, int64_t n, float threshold, float value, void* stream) {
    if (!ctx || !data_dev || n < 0) return m3d_fail(M3D_ERR_ARG, "m3d_replace_above: bad argument");
    if (n == 0) return M3D_OK;
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    size_t want = ((size_t)n + 255) / 256;
    size_t cap = (size_t)ctx->num_sms * 8;
    int blocks = (int)(want < cap ? want : cap);
    M3D_LAUNCH(ctx, KF_REPLACE_ABOVE, st,
               replace_above_kernel<<<blocks, 256, 0, st>>>(data_dev, (size_t)n, threshold, value));
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}
