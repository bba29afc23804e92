// centroid.cu -- per-on-bit intensity-weighted centroid statistics of decoded components
// (PD:2701-2906 `_add_on_bit_weighted_centroids` /
// `_plane_wise_weighted_centroid_statistics`).
//
// Reference: for every bit, label image dilated along z (max over a z_support window), weights
// max(intensity, 0) in float32, per-label float64 bincount sums of w, w*z, w*y, w*x and a float32
// running maximum over the UNdilated labels -- 16 full-volume passes, one per bit.
// Here: one pass.  A thread owns one (y, x) column and walks z with a register ring of labels, so the
// label image is read exactly once, coalesced along x; only the few voxels whose dilated label is
// foreground touch the stack, and only for the on-bits of that label's codeword (the reference never
// reads the other (label, bit) pairs).  float64 atomics: summation order differs from bincount's raster
// order, i.e. results agree to float64 round-off (~1e-16 relative), not bit-for-bit.
#include "common.cuh"

namespace {

constexpr int CT_THREADS = 128;
constexpr int MAX_HALF = 15;  // z_support up to 31

template <typename T>
__device__ __forceinline__ float load_f32(const T* p);
template <>
__device__ __forceinline__ float load_f32<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_f32<uint16_t>(const uint16_t* p) { return (float)__ldg(p); }

// per-voxel accumulation, shared by the windowed walk below
template <typename T>
__device__ __forceinline__ void accumulate_voxel(int32_t cl, int32_t own, int z, int y, int x, size_t v_idx, size_t n_vox,
                                                 const T* __restrict__ stack, const int16_t* __restrict__ label_code,
                                                 const uint32_t* __restrict__ cw_mask, int K, int n_bits,
                                                 long long minlength, double* __restrict__ sums,
                                                 float* __restrict__ peak) {
    if (cl > 0 && cl < minlength) {
        const int code = label_code[cl];
        if (code >= 0 && code < K) {
            uint32_t m = cw_mask[code];
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                if (b >= n_bits) break;
                const float w = fmaxf(load_f32<T>(stack + (size_t)b * n_vox + v_idx), 0.f);
                double* s = sums + ((size_t)cl * n_bits + b) * 4;
                atomicAdd(s + 0, (double)w);
                atomicAdd(s + 1, (double)w * (double)z);
                atomicAdd(s + 2, (double)__fmul_rn(w, (float)y));
                atomicAdd(s + 3, (double)__fmul_rn(w, (float)x));
            }
        }
    }
    if (own > 0 && own < minlength) {
        const int code = label_code[own];
        if (code >= 0 && code < K) {
            uint32_t m = cw_mask[code];
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                if (b >= n_bits) break;
                const float w = fmaxf(load_f32<T>(stack + (size_t)b * n_vox + v_idx), 0.f);
                // w >= 0: the int ordering of the bit patterns is the float ordering
                atomicMax(reinterpret_cast<int*>(peak + (size_t)own * n_bits + b), __float_as_int(w));
            }
        }
    }
}

// HALF >= 0: compile-time window, labels of planes z-HALF .. z+HALF live in registers and shift by one per
// step (no local memory).  HALF < 0: run-time half width with an indexed ring (any z_support up to 31).
template <typename T, int HALF>
__global__ void __launch_bounds__(CT_THREADS)
centroid_stats_kernel(const int32_t* __restrict__ labels, const T* __restrict__ stack, int Z, int Y, int X, int half_rt,
                      const int16_t* __restrict__ label_code, const uint32_t* __restrict__ cw_mask, int K, int n_bits,
                      long long minlength, double* __restrict__ sums, float* __restrict__ peak) {
    const long long col = (long long)blockIdx.x * CT_THREADS + threadIdx.x;
    const long long plane = (long long)Y * X;
    if (col >= plane) return;
    const int y = (int)(col / X), x = (int)(col % X);
    const size_t n_vox = (size_t)Z * plane;
    if constexpr (HALF >= 0) {
        constexpr int W = 2 * HALF + 1;
        int32_t win[W];  // win[i] = labels[z - HALF + i] (0 outside the volume)
#pragma unroll
        for (int i = 0; i < W; ++i) {
            const int zz = i - HALF - 1;  // state before the first shift
            win[i] = (zz >= 0 && zz < Z) ? __ldg(labels + (size_t)zz * plane + col) : 0;
        }
        constexpr int U = 4;  // planes fetched per batch: keeps U independent loads in flight per thread
        for (int zb = 0; zb < Z; zb += U) {
            int32_t nxt[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int zin = zb + j + HALF;
                nxt[j] = zin < Z ? __ldg(labels + (size_t)zin * plane + col) : 0;
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int z = zb + j;
#pragma unroll
                for (int i = 0; i < W - 1; ++i) win[i] = win[i + 1];
                win[W - 1] = nxt[j];
                if (z >= Z) continue;
                int32_t cl = 0;
#pragma unroll
                for (int i = 0; i < W; ++i) cl = win[i] > cl ? win[i] : cl;
                const int32_t own = win[HALF];
                if (cl <= 0 && own <= 0) continue;
                accumulate_voxel<T>(cl, own, z, y, x, (size_t)z * plane + col, n_vox, stack, label_code, cw_mask, K,
                                    n_bits, minlength, sums, peak);
            }
        }
    } else {
        const int half = half_rt;
        int32_t ring[2 * MAX_HALF + 1];
        const int W = 2 * half + 1;
#pragma unroll
        for (int i = 0; i < 2 * MAX_HALF + 1; ++i) ring[i] = 0;
        // ring[zz % W] holds labels[zz]; preload planes 0 .. half-1
        for (int zz = 0; zz < half && zz < Z; ++zz) ring[zz % W] = labels[(size_t)zz * plane + col];
        for (int z = 0; z < Z; ++z) {
            const int zin = z + half;
            if (zin < Z) ring[zin % W] = labels[(size_t)zin * plane + col];
            const int z0 = z - half < 0 ? 0 : z - half, z1 = z + half >= Z ? Z - 1 : z + half;
            int32_t cl = 0;
            for (int zz = z0; zz <= z1; ++zz) {
                const int32_t v = ring[zz % W];
                cl = v > cl ? v : cl;
            }
            const int32_t own = ring[z % W];
            if (cl <= 0 && own <= 0) continue;
            accumulate_voxel<T>(cl, own, z, y, x, (size_t)z * plane + col, n_vox, stack, label_code, cw_mask, K, n_bits,
                                minlength, sums, peak);
        }
    }
}

template <typename T>
void launch_centroid(m3d_ctx* ctx, int blocks, cudaStream_t st, int half, const int32_t* labels, const T* stack, int Z,
                     int Y, int X, const int16_t* label_code, long long minlength, double* sums, float* peak) {
#define M3D_CT(H)                                                                                                   \
    centroid_stats_kernel<T, H><<<blocks, CT_THREADS, 0, st>>>(labels, stack, Z, Y, X, half, label_code, ctx->d_cw_mask, \
                                                               ctx->K, ctx->n_bits, minlength, sums, peak)
    KernelScope ks(ctx, KF_CENTROID, st);
    switch (half) {
        case 0: M3D_CT(0); break;
        case 1: M3D_CT(1); break;
        case 2: M3D_CT(2); break;
        case 3: M3D_CT(3); break;  // the reference default, centroid_z_support = 7
        default: M3D_CT(-1); break;
    }
#undef M3D_CT
}

}  // namespace

extern "C" int m3d_centroid_statistics(m3d_ctx* ctx, const int32_t* labels_dev, const void* stack_dev, int dtype,
                                       const int64_t dims[3], int z_support, const int16_t* label_code_dev,
                                       int64_t minlength, double* sums_dev, float* peak_dev, void* stream) {
    if (!ctx || !labels_dev || !stack_dev || !dims || !label_code_dev || !sums_dev || !peak_dev)
        return m3d_fail(M3D_ERR_ARG, "m3d_centroid_statistics: null argument");
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || minlength < 1)
        return m3d_fail(M3D_ERR_ARG, "m3d_centroid_statistics: bad dims / minlength");
    if (dtype != M3D_DTYPE_U16 && dtype != M3D_DTYPE_F32)
        return m3d_fail(M3D_ERR_ARG, "m3d_centroid_statistics: dtype %d", dtype);
    int half = z_support / 2;
    if (half < 0) half = 0;
    if (half > MAX_HALF) return m3d_fail(M3D_ERR_ARG, "m3d_centroid_statistics: z_support %d exceeds %d", z_support, 2 * MAX_HALF + 1);
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t cells = (size_t)minlength * ctx->n_bits;
    M3D_CUDA(cudaMemsetAsync(sums_dev, 0, cells * 4 * sizeof(double), st));
    M3D_CUDA(cudaMemsetAsync(peak_dev, 0, cells * sizeof(float), st));
    const long long plane = (long long)dims[1] * dims[2];
    const int blocks = (int)((plane + CT_THREADS - 1) / CT_THREADS);
    if (dtype == M3D_DTYPE_U16)
        launch_centroid<uint16_t>(ctx, blocks, st, half, labels_dev, reinterpret_cast<const uint16_t*>(stack_dev),
                                  (int)dims[0], (int)dims[1], (int)dims[2], label_code_dev, (long long)minlength,
                                  sums_dev, peak_dev);
    else
        launch_centroid<float>(ctx, blocks, st, half, labels_dev, reinterpret_cast<const float*>(stack_dev),
                               (int)dims[0], (int)dims[1], (int)dims[2], label_code_dev, (long long)minlength, sums_dev,
                               peak_dev);
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

// ------------------------------------------------------------------ inertia-tensor eigenvalues
// scikit-image `inertia_tensor_eigvals` (PD:3038-3047 requests it): eigenvalues of the 3x3 inertia
// tensor built from the normalised second central moments, clipped at 0, descending.  The reference
// gets them from LAPACK on the host, one Python call per region; here one thread per table row runs
// cyclic Jacobi rotations in float64 (converged to round-off in <= 6 sweeps for 3x3).
namespace {

__global__ void __launch_bounds__(128)
inertia_eigvals_kernel(const double* __restrict__ table, long long n_rows, long long stride, double* __restrict__ out) {
    const long long r = (long long)blockIdx.x * 128 + threadIdx.x;
    if (r >= n_rows) return;
    const double* t = table + r * stride;
    const double n = t[1];
    const double zz = t[6], yy = t[7], xx = t[8], zy = t[9], zx = t[10], yx = t[11];
    double a00 = (yy + xx) / n, a11 = (zz + xx) / n, a22 = (zz + yy) / n;
    double a01 = -zy / n, a02 = -zx / n, a12 = -yx / n;
    for (int sweep = 0; sweep < 8; ++sweep) {
        const double off = fabs(a01) + fabs(a02) + fabs(a12);
        if (off == 0.0) break;
#define M3D_JACOBI(app, aqq, apq, arp, arq)                                              \
    if (apq != 0.0) {                                                                    \
        const double theta = (aqq - app) / (2.0 * apq);                                  \
        const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0)); \
        const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;                          \
        app -= tt * apq;                                                                 \
        aqq += tt * apq;                                                                 \
        apq = 0.0;                                                                       \
        const double rp = arp, rq = arq;                                                 \
        arp = c * rp - s * rq;                                                           \
        arq = s * rp + c * rq;                                                           \
    }
        M3D_JACOBI(a00, a11, a01, a02, a12)
        M3D_JACOBI(a00, a22, a02, a01, a12)
        M3D_JACOBI(a11, a22, a12, a01, a02)
#undef M3D_JACOBI
    }
    double e0 = fmax(a00, 0.0), e1 = fmax(a11, 0.0), e2 = fmax(a22, 0.0), tmp;
    if (e0 < e1) { tmp = e0; e0 = e1; e1 = tmp; }
    if (e1 < e2) { tmp = e1; e1 = e2; e2 = tmp; }
    if (e0 < e1) { tmp = e0; e0 = e1; e1 = tmp; }
    out[r * 3 + 0] = e0;
    out[r * 3 + 1] = e1;
    out[r * 3 + 2] = e2;
}

}  // namespace

extern "C" int m3d_inertia_eigvals(m3d_ctx* ctx, const double* table_dev, int64_t n_rows, int64_t n_cols,
                                   double* eigvals_dev, void* stream) {
    if (!ctx || n_rows < 0 || n_cols < M3D_TABLE_FIXED_COLS)
        return m3d_fail(M3D_ERR_ARG, "m3d_inertia_eigvals: bad argument");
    if (n_rows == 0) return M3D_OK;
    if (!table_dev || !eigvals_dev) return m3d_fail(M3D_ERR_ARG, "m3d_inertia_eigvals: null argument");
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int blocks = (int)((n_rows + 127) / 128);
    M3D_LAUNCH(ctx, KF_EIGVALS, st,
               inertia_eigvals_kernel<<<blocks, 128, 0, st>>>(table_dev, (long long)n_rows, (long long)n_cols, eigvals_dev));
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}
