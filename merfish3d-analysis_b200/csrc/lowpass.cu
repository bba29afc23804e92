// lowpass.cu -- separable Gaussian low-pass (replaces PD:1948-2024).
//
// Semantics = scipy.ndimage.gaussian_filter(float32 volume, sigma, mode='reflect',
// truncate=4): axes filtered in order z, y, x; every 1-D pass accumulates in float64 with
// SciPy's symmetric formula  o = x[c]*w[c]; for j=-r..-1: o += (x[c+j] + x[c-j]) * w[c+j]
// and stores float32.  The predictor weighting float32(readout)*float32(predictor)
// (PD:1879-1881) is fused into the first pass's loads.
//
// Kernels
//   lowpass_z_kernel<T,R>   : one thread per (y,x) column marching along z with a register
//                             ring of 2R+1 float64 samples (loop unrolled by the ring length so
//                             every ring index is static).  Reads each input sample once
//                             (+2R reflected planes), writes float32.  Bound by the FP64 pipe.
//   lowpass_yx_kernel<T,RY,RX>: 32x64 output tile per CTA; the (32+2RY)x(64+2RX) input tile is
//                             staged in shared memory as float64, y pass -> float32-rounded
//                             intermediate in shared memory -> x pass -> float32 store.
//   lowpass_axis_generic_kernel: any radius <= 64 / any shape, one thread per output; used
//                             when no templated radius matches.
//
// Opt-in second arithmetic (m3d_set_lowpass_mode(ctx, 1), M3D_LOWPASS_ACCUM=float32): weights cast to float32 and every
// 1-D pass accumulated in float32 with one FMA per tap, taps in ascending order -- what cupyx.scipy.ndimage's correlate
// kernel does for float32 images as far as its source is remembered (weights dtype = promote(input, float32), `sum +=
// value * w` under NVRTC's default FMA contraction; the reference calls it at PD:1972-1979).  CuPy is not installed
// here and has no vendored source, so this mode is NOT pinned to the reference; the default stays SciPy's float64
// arithmetic, which the reference-generated goldens pin.  Same kernels, `F32` template flag.
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace {

constexpr int LP_MAX_RADIUS = 64;

struct Weights {
    int r;
    double w[LP_MAX_RADIUS + 1];  // w[i] = weight at offset -(r - i) ... i.e. w[0]=edge, w[r]=centre
};
struct Weights32 {  // the float32 mode's full kernel: w[i] = float32(weight at offset i - r), i = 0 .. 2 r
    float w[2 * LP_MAX_RADIUS + 1];
};
static Weights32 make_weights32(const Weights& W) {
    Weights32 F;
    memset(&F, 0, sizeof(F));
    for (int i = 0; i <= 2 * W.r && W.r >= 0; ++i) F.w[i] = (float)W.w[i <= W.r ? i : 2 * W.r - i];
    return F;
}

// SciPy _gaussian_kernel1d (order 0), float64; symmetric so only the first r+1 entries are kept.
static Weights make_weights(double sigma) {
    Weights W;
    int lw = (int)(4.0 * sigma + 0.5);
    if (lw > LP_MAX_RADIUS) lw = -1;
    W.r = lw;
    if (lw < 0) return W;
    std::vector<double> phi(2 * lw + 1);
    double sum = 0.0;
    const double s2 = sigma * sigma;
    for (int i = -lw; i <= lw; ++i) {
        phi[i + lw] = exp(-0.5 / s2 * (double)(i * i));
    }
    // NumPy's phi.sum() is a pairwise sum; for <= 128 elements with n >= 8 it uses eight
    // interleaved accumulators.  Reproduce it so the normalised weights match bit for bit.
    const int n = 2 * lw + 1;
    if (n < 8) {
        sum = 0.0;
        for (int i = 0; i < n; ++i) sum += phi[i];
    } else {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = phi[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += phi[i + j];
        sum = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) sum += phi[i];
    }
    for (int i = 0; i <= lw; ++i) W.w[i] = phi[i] / sum;  // phi[::-1] is symmetric
    return W;
}

__device__ __forceinline__ int reflect_index(int j, int n) {
    if (j >= 0 && j < n) return j;
    const int per = 2 * n;
    int m = j % per;
    if (m < 0) m += per;
    return (m >= n) ? (per - 1 - m) : m;
}

// reflect for indices that overshoot the end by less than one period: no integer division on the
// marching kernels' per-plane path (the general form stays behind a real branch)
__device__ __forceinline__ int reflect_high(int j, int n) {
    if (j < n) return j;
    if (j < 2 * n) return 2 * n - 1 - j;
    return reflect_index(j, n);
}

template <typename T>
__device__ __forceinline__ float in_as_f32(const T* p, size_t i);
template <>
__device__ __forceinline__ float in_as_f32<uint16_t>(const uint16_t* p, size_t i) { return (float)__ldg(p + i); }
template <>
__device__ __forceinline__ float in_as_f32<float>(const float* p, size_t i) { return __ldg(p + i); }

// A sample that is loaded now and used later stays in a full 32-bit register until then.  (Held as `uint16_t` the
// compiler packs it into half a register with a PRMT right behind the LDG -- an instruction that waits for the load,
// which is exactly what a prefetch is meant to avoid: the z kernel spent 10 of 17 cycles per issue in long_scoreboard.)
template <typename T>
__device__ __forceinline__ uint32_t ld_raw32(const T* p);
template <>
__device__ __forceinline__ uint32_t ld_raw32<uint16_t>(const uint16_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <>
__device__ __forceinline__ uint32_t ld_raw32<float>(const float* p) { return __float_as_uint(__ldg(p)); }
template <typename T>
__device__ __forceinline__ float raw32_as_f32(uint32_t v);
template <>
__device__ __forceinline__ float raw32_as_f32<uint16_t>(uint32_t v) { return (float)v; }  // < 65536: exact, = (float)(uint16_t)
template <>
__device__ __forceinline__ float raw32_as_f32<float>(uint32_t v) { return __uint_as_float(v); }

template <typename T, bool PRED>
__device__ __forceinline__ double load_weighted(const T* in, const float* pred, size_t i) {
    float v = in_as_f32<T>(in, i);
    if (PRED) v = __fmul_rn(v, __ldg(pred + i));
    return (double)v;
}

// ------------------------------------------------------------------ z pass, register ring
// One thread per (y,x) column marching along z.  The 2R+1 samples of the window live in a register ring (loop
// unrolled by the ring length so every index is static); the sample that enters the ring is requested PF planes
// ahead and stays raw until it is needed.
//
// What decided the speed of this kernel (profiles/r2_lowpass.txt): the unrolled steps must be BRANCH-FREE.  With a
// bound check and a reflection branch per step, ptxas gave every prefetch load the same scoreboard, so the first use of
// the OLDEST sample waited for the YOUNGEST load as well -- a queue five deep in the source and one deep in the
// hardware (70 % of the stall samples on that conversion, float64 pipe 43 % busy).  Whole chunks of 2R+1 outputs now run
// without a branch: the plane address marches (forwards, then backwards through the reflected tail: a uniform select),
// the loads rotate over several scoreboards, and a step is 37 float64 operations plus ~12 others.  The checked form is
// kept for volumes thinner than the window; a partial last chunk only adds the bound check.  (Staging the samples through shared memory
// with cp.async commit groups -- counted waits, eight planes in flight -- was measured too: 1.40 ms against 1.22 ms.)

// f(integral_constant<int, 0>) ... f(integral_constant<int, N - 1>): a loop whose index is a compile-time constant
template <int N, int I = 0, typename F>
__device__ __forceinline__ void unroll_steps(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        unroll_steps<N, I + 1>(f);
    }
}

template <typename T, int R, bool PRED, bool F32>
__global__ void __launch_bounds__(128, 5)
lowpass_z_kernel(const T* __restrict__ in, const float* __restrict__ pred, float* __restrict__ out,
                 int Z, size_t plane, Weights W, Weights32 W32) {
    constexpr int RING = 2 * R + 1;
    constexpr int PF = (RING % 5 == 0) ? 5 : ((RING % 3 == 0) ? 3 : 1);  // divides the unroll length
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= plane) return;
    using A = typename std::conditional<F32, float, double>::type;
    A ring[RING];
    uint32_t pre[PF];  // prefetched samples stay raw (ld_raw32): converting them on arrival would make the
    float pre_w[PF];   // warp wait on the very load the prefetch is meant to hide
    // slot of ext[j] is (j + R) mod RING
#pragma unroll
    for (int i = 0; i < 2 * R; ++i)
        ring[i] = (A)load_weighted<T, PRED>(in, pred, (size_t)reflect_index(i - R, Z) * plane + c);
#pragma unroll
    for (int i = 0; i < PF; ++i) {
        const size_t a = (size_t)reflect_index(R + i, Z) * plane + c;
        pre[i] = ld_raw32<T>(in + a);
        pre_w[i] = PRED ? __ldg(pred + a) : 1.f;
    }
    // marching address of (reflected) plane o + R + PF: valid while that index stays below 2 Z
    size_t a_fwd = (size_t)reflect_index(R + PF, Z) * plane + c;
    float* po = out + c;
    // one output step.  CHECK: o may lie past the end (partial last chunk).  MARCH: o + R + PF + 1 < 2 Z holds for every
    // o < Z, so the address of the reflected plane marches and needs no division.
    auto step = [&](auto check_tag, auto march_tag, auto u_tag, int o) {
        constexpr bool CHECK = decltype(check_tag)::value, MARCH = decltype(march_tag)::value;
        constexpr int u = decltype(u_tag)::value;
        if (CHECK && o >= Z) return;
        {
            float v = raw32_as_f32<T>(pre[u % PF]);
            if (PRED) v = __fmul_rn(v, pre_w[u % PF]);
            ring[(u + 2 * R) % RING] = (A)v;
        }
        {   // the sample for output o + PF
            const int zi = o + R + PF;
            size_t a;
            if (MARCH) {
                a = a_fwd;  // next: one plane on, the last plane once more (reflection), then back down
                a_fwd = zi + 1 < Z ? a_fwd + plane : (zi + 1 == Z ? a_fwd : a_fwd - plane);
            } else {
                a = (size_t)reflect_high(zi, Z) * plane + c;
            }
            pre[u % PF] = ld_raw32<T>(in + a);
            if (PRED) pre_w[u % PF] = __ldg(pred + a);
        }
        if constexpr (F32) {
            float acc = 0.f;
#pragma unroll
            for (int jj = -R; jj <= R; ++jj) acc = __fmaf_rn(ring[(u + R + jj + RING) % RING], W32.w[jj + R], acc);
            *po = acc;
        } else {
            double acc = __dmul_rn(ring[(u + R) % RING], W.w[R]);
#pragma unroll
            for (int jj = -R; jj < 0; ++jj) {
                double pr = __dadd_rn(ring[(u + R + jj + RING) % RING], ring[(u + R - jj) % RING]);
                acc = __dadd_rn(acc, __dmul_rn(pr, W.w[jj + R]));
            }
            *po = (float)acc;
        }
        po += plane;
    };
    auto chunk = [&](auto check_tag, auto march_tag, int o0) {
        unroll_steps<RING>([&](auto u_tag) { step(check_tag, march_tag, u_tag, o0 + decltype(u_tag)::value); });
    };
    int o0 = 0;
    if (Z >= R + PF + 1) {
        for (; o0 + RING <= Z; o0 += RING) chunk(std::false_type{}, std::true_type{}, o0);
        if (o0 < Z) chunk(std::true_type{}, std::true_type{}, o0);
    } else {
        for (; o0 < Z; o0 += RING) chunk(std::true_type{}, std::false_type{}, o0);
    }
}

// ------------------------------------------------------------------ fused y,x pass, shared tile
// Tile of YX_TY x (64 - 2*RX) outputs per CTA.  The float32 input tile (with halo) is staged in
// shared memory once; both passes are REGISTER BLOCKED: a work item loads 8 + 2R consecutive
// samples along the filter axis (conflict-free), converts them to float64 once and produces 8
// outputs with SciPy's symmetric formula -- ~3 shared-memory words per output instead of 20, so the
// kernel is bound by the float64 pipe, not by shared memory.
//   y pass: item = (column, group of 8 rows); lanes = consecutive columns.
//   x pass: item = (row, group of 8 columns); lanes = consecutive rows (row stride 73 words is
//           coprime with the 32 banks); results go back through shared memory for coalesced stores.
constexpr int YX_TY = 32;
constexpr int YX_W = 64;             // staged columns = outputs + 2*RX
constexpr int YX_STRIDE = YX_W + 9;  // 73 words: odd and coprime with 32
constexpr int YX_THREADS = 256;
constexpr int YX_BLK = 8;            // outputs per work item

template <typename T, int RY, int RX, bool PRED, bool F32>
__global__ void __launch_bounds__(YX_THREADS, F32 ? 4 : 5)
lowpass_yx_kernel(const T* __restrict__ in, const float* __restrict__ pred, float* __restrict__ out,
                  int Y, int X, Weights WY, Weights WX, Weights32 WY32, Weights32 WX32) {
    using A = typename std::conditional<F32, float, double>::type;
    constexpr int IN_H = YX_TY + 2 * RY;
    constexpr int TX = YX_W - 2 * RX;
    constexpr int XG = (TX + YX_BLK - 1) / YX_BLK;  // column groups (the last may be partial)
    static_assert(YX_TY % YX_BLK == 0 && TX > 0, "tile must split into 8-output items");
    __shared__ float s_in[IN_H][YX_STRIDE];
    __shared__ float s_mid[YX_TY][YX_STRIDE];
    float (*s_out)[YX_STRIDE] = s_in;  // output stage (the input tile is dead by then)
    const int x0 = blockIdx.x * TX;
    const int y0 = blockIdx.y * YX_TY;
    const size_t zoff = (size_t)blockIdx.z * (size_t)Y * X;
    {
        // thread = (column lx, row phase): the column reflection is computed once, all of the
        // thread's loads are issued before the first shared-memory store
        constexpr int ROWS_PER_PASS = YX_THREADS / YX_W;            // 4
        constexpr int N_LD = (IN_H + ROWS_PER_PASS - 1) / ROWS_PER_PASS;
        const int lx = threadIdx.x % YX_W, lr = threadIdx.x / YX_W;
        const int gx = reflect_index(x0 + lx - RX, X);
        uint32_t raw[N_LD];
        float wv[N_LD];
        if (y0 - RY >= 0 && y0 - RY + N_LD * ROWS_PER_PASS <= Y) {
            // tile rows (and the rows the last, partly unused pass touches) all exist: one address, then a row stride
            size_t a = zoff + (size_t)(y0 - RY + lr) * X + gx;
#pragma unroll
            for (int k = 0; k < N_LD; ++k) {
                raw[k] = ld_raw32<T>(in + a);
                wv[k] = PRED ? __ldg(pred + a) : 1.f;
                a += (size_t)ROWS_PER_PASS * X;
            }
        } else {
#pragma unroll
            for (int k = 0; k < N_LD; ++k) {
                const int ly = lr + k * ROWS_PER_PASS;
                const int gy = reflect_index(y0 + (ly < IN_H ? ly : IN_H - 1) - RY, Y);
                const size_t a = zoff + (size_t)gy * X + gx;
                raw[k] = ld_raw32<T>(in + a);
                wv[k] = PRED ? __ldg(pred + a) : 1.f;
            }
        }
#pragma unroll
        for (int k = 0; k < N_LD; ++k) {
            const int ly = lr + k * ROWS_PER_PASS;
            float v = raw32_as_f32<T>(raw[k]);
            if (PRED) v = __fmul_rn(v, wv[k]);
            if (ly < IN_H) s_in[ly][lx] = v;
        }
    }
    __syncthreads();
    // ---- y pass: YX_W columns x (YX_TY / 8) row groups = 256 items
    for (int it = threadIdx.x; it < YX_W * (YX_TY / YX_BLK); it += YX_THREADS) {
        const int col = it % YX_W, r0 = (it / YX_W) * YX_BLK;
        A v[YX_BLK + 2 * RY];
#pragma unroll
        for (int k = 0; k < YX_BLK + 2 * RY; ++k) v[k] = (A)s_in[r0 + k][col];
#pragma unroll
        for (int o = 0; o < YX_BLK; ++o) {
            if constexpr (F32) {
                float acc = 0.f;
#pragma unroll
                for (int jj = -RY; jj <= RY; ++jj) acc = __fmaf_rn(v[o + RY + jj], WY32.w[jj + RY], acc);
                s_mid[r0 + o][col] = acc;
            } else {
                double acc = __dmul_rn(v[o + RY], WY.w[RY]);
#pragma unroll
                for (int jj = -RY; jj < 0; ++jj)
                    acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(v[o + RY + jj], v[o + RY - jj]), WY.w[jj + RY]));
                s_mid[r0 + o][col] = (float)acc;
            }
        }
    }
    __syncthreads();
    // ---- x pass: YX_TY rows x (TX / 8) column groups; lanes = rows
    for (int it = threadIdx.x; it < YX_TY * XG; it += YX_THREADS) {
        const int row = it % YX_TY, c0 = (it / YX_TY) * YX_BLK;
        A v[YX_BLK + 2 * RX];
#pragma unroll
        for (int k = 0; k < YX_BLK + 2 * RX; ++k) v[k] = (c0 + k < YX_W) ? (A)s_mid[row][c0 + k] : (A)0;
#pragma unroll
        for (int o = 0; o < YX_BLK; ++o) {
            if (c0 + o >= TX) break;
            if constexpr (F32) {
                float acc = 0.f;
#pragma unroll
                for (int jj = -RX; jj <= RX; ++jj) acc = __fmaf_rn(v[o + RX + jj], WX32.w[jj + RX], acc);
                s_out[row][c0 + o] = acc;  // the input tile is dead: reuse it as the output stage
            } else {
                double acc = __dmul_rn(v[o + RX], WX.w[RX]);
#pragma unroll
                for (int jj = -RX; jj < 0; ++jj)
                    acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(v[o + RX + jj], v[o + RX - jj]), WX.w[jj + RX]));
                s_out[row][c0 + o] = (float)acc;  // the input tile is dead: reuse it as the output stage
            }
        }
    }
    __syncthreads();
    {
        // a warp stores whole rows (lane = column, then column + 32): one row pointer that advances by the warp count
        constexpr int WARPS = YX_THREADS / 32;
        const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
        const bool c0_ok = lane < TX && x0 + lane < X, c1_ok = lane + 32 < TX && x0 + lane + 32 < X;
        float* po = out + zoff + (size_t)(y0 + wrp) * X + x0 + lane;
#pragma unroll
        for (int ly = wrp; ly < YX_TY; ly += WARPS) {
            if (y0 + ly < Y) {
                if (c0_ok) po[0] = s_out[ly][lane];
                if (c1_ok) po[32] = s_out[ly][lane + 32];
            }
            po += (size_t)WARPS * X;
        }
    }
}

// ------------------------------------------------------------------ generic single-axis pass
template <typename T, bool PRED>
__global__ void __launch_bounds__(256)
lowpass_axis_generic_kernel(const T* __restrict__ in, const float* __restrict__ pred,
                            float* __restrict__ out, size_t total, int len, size_t stride, Weights W, int f32,
                            Weights32 W32) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int pos = (int)((i / stride) % (size_t)len);
    const size_t base = i - (size_t)pos * stride;
    const int r = W.r;
    if (f32) {
        float acc32 = 0.f;
        for (int jj = -r; jj <= r; ++jj) {
            const float a = (float)load_weighted<T, PRED>(in, pred, base + (size_t)reflect_index(pos + jj, len) * stride);
            acc32 = __fmaf_rn(a, W32.w[jj + r], acc32);
        }
        out[i] = acc32;
        return;
    }
    double acc = __dmul_rn(load_weighted<T, PRED>(in, pred, i), W.w[r]);
    for (int jj = -r; jj < 0; ++jj) {
        double a = load_weighted<T, PRED>(in, pred, base + (size_t)reflect_index(pos + jj, len) * stride);
        double b = load_weighted<T, PRED>(in, pred, base + (size_t)reflect_index(pos - jj, len) * stride);
        acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(a, b), W.w[jj + r]));
    }
    out[i] = (float)acc;
}

template <typename T, bool PRED>
int run_generic_axis(m3d_ctx* ctx, const T* in, const float* pred, float* out, size_t total, int len,
                     size_t stride, const Weights& W, cudaStream_t st) {
    int blocks = (int)((total + 255) / 256);
    const Weights32 W32 = make_weights32(W);
    M3D_LAUNCH(ctx, KF_LOWPASS_GENERIC, st,
               lowpass_axis_generic_kernel<T, PRED><<<blocks, 256, 0, st>>>(in, pred, out, total, len, stride, W,
                                                                            ctx->lowpass_f32, W32));
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

template <typename T, bool PRED>
int run_z(m3d_ctx* ctx, const T* in, const float* pred, float* out, int Z, int Y, int X, const Weights& W,
          cudaStream_t st) {
    const size_t plane = (size_t)Y * X;
    const int blocks = (int)((plane + 127) / 128);
    if (W.r != 12 && W.r != 8 && W.r != 6 && W.r != 4 && W.r != 2)
        return run_generic_axis<T, PRED>(ctx, in, pred, out, (size_t)Z * plane, Z, plane, W, st);
    KernelScope ks(ctx, KF_LOWPASS_Z, st);
    const Weights32 W32 = make_weights32(W);
#define M3D_Z_CASE(R_)                                                                                          \
    case R_:                                                                                                    \
        if (ctx->lowpass_f32) lowpass_z_kernel<T, R_, PRED, true><<<blocks, 128, 0, st>>>(in, pred, out, Z, plane, W, W32); \
        else lowpass_z_kernel<T, R_, PRED, false><<<blocks, 128, 0, st>>>(in, pred, out, Z, plane, W, W32);      \
        break;
    switch (W.r) {
        M3D_Z_CASE(12)
        M3D_Z_CASE(8)
        M3D_Z_CASE(6)
        M3D_Z_CASE(4)
        M3D_Z_CASE(2)
    }
#undef M3D_Z_CASE
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

// y then x on `n_planes` planes.  Returns 1 when no templated instance matches.
template <typename T, bool PRED>
int run_yx_fast(m3d_ctx* ctx, const T* in, const float* pred, float* out, int n_planes, int Y, int X,
                const Weights& WY, const Weights& WX, cudaStream_t st) {
    const int tx = YX_W - 2 * WX.r;  // outputs per tile row (see lowpass_yx_kernel)
    dim3 grid((X + tx - 1) / tx, (Y + YX_TY - 1) / YX_TY, n_planes);
    if (grid.y > 65535 || grid.z > 65535) return 1;
    if (!((WY.r == 4 && WX.r == 4) || (WY.r == 2 && WX.r == 2) || (WY.r == 6 && WX.r == 6))) return 1;
    KernelScope ks(ctx, KF_LOWPASS_YX, st);
    const Weights32 WY32 = make_weights32(WY), WX32 = make_weights32(WX);
#define M3D_YX_CASE(R_)                                                                                                   \
    if (WY.r == R_ && WX.r == R_) {                                                                                       \
        if (ctx->lowpass_f32)                                                                                             \
            lowpass_yx_kernel<T, R_, R_, PRED, true><<<grid, YX_THREADS, 0, st>>>(in, pred, out, Y, X, WY, WX, WY32, WX32); \
        else                                                                                                              \
            lowpass_yx_kernel<T, R_, R_, PRED, false><<<grid, YX_THREADS, 0, st>>>(in, pred, out, Y, X, WY, WX, WY32, WX32); \
    }
    M3D_YX_CASE(4)
    M3D_YX_CASE(2)
    M3D_YX_CASE(6)
#undef M3D_YX_CASE
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

template <typename T, bool PRED>
int lowpass_volume(m3d_ctx* ctx, const T* in, const float* pred, int Z, int Y, int X, const Weights& WZ,
                   const Weights& WY, const Weights& WX, int mode2d, float* out, cudaStream_t st) {
    const size_t vol = (size_t)Z * Y * X;
    if (mode2d) {
        int rc = run_yx_fast<T, PRED>(ctx, in, pred, out, Z, Y, X, WY, WX, st);
        if (rc <= 0) return rc;
        if (ctx->s_lp_tmp.ensure(vol * sizeof(float))) return M3D_ERR_CUDA;
        float* tmp = reinterpret_cast<float*>(ctx->s_lp_tmp.ptr);
        rc = run_generic_axis<T, PRED>(ctx, in, pred, tmp, vol, Y, (size_t)X, WY, st);
        if (rc) return rc;
        return run_generic_axis<float, false>(ctx, tmp, nullptr, out, vol, X, 1, WX, st);
    }
    if (ctx->s_lp_tmp.ensure(2 * vol * sizeof(float))) return M3D_ERR_CUDA;
    float* tmp = reinterpret_cast<float*>(ctx->s_lp_tmp.ptr);
    float* tmp2 = tmp + vol;
    int rc = run_z<T, PRED>(ctx, in, pred, tmp, Z, Y, X, WZ, st);
    if (rc) return rc;
    rc = run_yx_fast<float, false>(ctx, tmp, nullptr, out, Z, Y, X, WY, WX, st);
    if (rc <= 0) return rc;
    rc = run_generic_axis<float, false>(ctx, tmp, nullptr, tmp2, vol, Y, (size_t)X, WY, st);
    if (rc) return rc;
    return run_generic_axis<float, false>(ctx, tmp2, nullptr, out, vol, X, 1, WX, st);
}

template <typename T>
int lowpass_all(m3d_ctx* ctx, const T* in, const float* pred, int n_vols, int Z, int Y, int X,
                const double sigma[3], int mode2d, float* out, cudaStream_t st) {
    Weights WZ = make_weights(sigma[0]);
    Weights WY = make_weights(sigma[1]);
    Weights WX = make_weights(sigma[2]);
    if ((!mode2d && WZ.r < 0) || WY.r < 0 || WX.r < 0)
        return m3d_fail(M3D_ERR_ARG, "m3d_lowpass: sigma too large (radius > %d)", LP_MAX_RADIUS);
    const size_t vol = (size_t)Z * Y * X;
    for (int v = 0; v < n_vols; ++v) {
        int rc = pred ? lowpass_volume<T, true>(ctx, in + v * vol, pred + v * vol, Z, Y, X, WZ, WY, WX, mode2d,
                                                out + v * vol, st)
                      : lowpass_volume<T, false>(ctx, in + v * vol, nullptr, Z, Y, X, WZ, WY, WX, mode2d,
                                                 out + v * vol, st);
        if (rc) return rc;
    }
    return M3D_OK;
}

// readout (uint16) x predictor (float32) -> float32 (PD:1879-1881).  Streaming: 8 elements per thread (one 128-bit
// load of samples, two of weights, two 128-bit stores) when the three pointers allow it, scalar otherwise / for the tail.
__global__ void __launch_bounds__(256)
weight_kernel(const uint16_t* __restrict__ r, const float* __restrict__ p, float* __restrict__ o, size_t n, size_t n8) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n8) {
        const uint4 q = __ldcs(reinterpret_cast<const uint4*>(r) + t);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        float v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[2 * j] = (float)(w[j] & 0xFFFFu);
            v[2 * j + 1] = (float)(w[j] >> 16);
        }
        if (p) {
            const float4 a = __ldcs(reinterpret_cast<const float4*>(p) + 2 * t);
            const float4 b = __ldcs(reinterpret_cast<const float4*>(p) + 2 * t + 1);
            v[0] = __fmul_rn(v[0], a.x);
            v[1] = __fmul_rn(v[1], a.y);
            v[2] = __fmul_rn(v[2], a.z);
            v[3] = __fmul_rn(v[3], a.w);
            v[4] = __fmul_rn(v[4], b.x);
            v[5] = __fmul_rn(v[5], b.y);
            v[6] = __fmul_rn(v[6], b.z);
            v[7] = __fmul_rn(v[7], b.w);
        }
        float4* dst = reinterpret_cast<float4*>(o) + 2 * t;
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    // elements past the vector body (all of them when a pointer is not 16-byte aligned: n8 = 0)
    const size_t i = 8 * n8 + t;
    if (i < n) {
        const float s = (float)__ldg(r + i);
        o[i] = p ? __fmul_rn(s, __ldg(p + i)) : s;
    }
}

// ------------------------------------------------------------------ decode-time affine warp
// scipy.ndimage.affine_transform(order=1, mode='constant', cval=0) semantics (the reference's
// warp_array_to_reference_gpu, utils/multiview_registration.py:797-902): input coordinate
// c_a = ((offset_a + z*m_a0) + y*m_a1) + x*m_a2 in float64; a sample whose coordinate falls outside
// [0, dim-1] on any axis is cval; otherwise the 8 taps are accumulated in float64 in z,y,x nesting
// order as ((v*wz)*wy)*wx, taps outside the image contributing 0.  The predictor multiply of
// PD:1879-1881 is fused into the tap loads.
struct AffineParams {
    double m[9];
    double off[3];
};

// scipy.ndimage order-1 sampling (NI_GeometricTransform, mode='constant', cval=0) of one volume at a float64
// coordinate: outside [0, dim-1] on any axis -> 0; weights (1-f, f) per axis; taps summed in z,y,x nesting order
// with value * wz * wy * wx, all in float64, like SciPy.  The predictor multiply (float32) is fused into the taps.
template <typename T, bool PRED>
__device__ __forceinline__ double sample_order1(const T* __restrict__ in, const float* __restrict__ pred, int Z, int Y, int X,
                                                const double (&c)[3]) {
    const int dims[3] = {Z, Y, X};
    const size_t plane = (size_t)Y * X;
    int st[3];
    double w[3][2];
    bool outside = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        outside |= !(c[a] >= 0.0 && c[a] <= (double)(dims[a] - 1));  // NaN coordinates are outside too
        const double fl = floor(c[a]);
        st[a] = (int)fl;
        const double f = __dadd_rn(c[a], -fl);
        w[a][0] = __dadd_rn(1.0, -f);
        w[a][1] = f;
    }
    double t = 0.0;
    if (!outside) {
#pragma unroll
        for (int dz = 0; dz < 2; ++dz)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int iz = st[0] + dz, iy = st[1] + dy, ix = st[2] + dx;
                    double v = 0.0;
                    if (iz < Z && iy < Y && ix < X) {  // lower bounds hold: coordinates are >= 0
                        const size_t j = (size_t)iz * plane + (size_t)iy * X + ix;
                        float fv = in_as_f32<T>(in, j);
                        if (PRED) fv = __fmul_rn(fv, __ldg(pred + j));
                        v = (double)fv;
                    }
                    const double term = __dmul_rn(__dmul_rn(__dmul_rn(v, w[0][dz]), w[1][dy]), w[2][dx]);
                    t = __dadd_rn(t, term);
                }
    }
    return t;
}

// Affine + SOFIMA flow warp (utils/multiview_registration.py:905-1131): the coarse flow field (3 channels X,Y,Z
// on a strided grid) is interpolated at every output voxel, added to the voxel index, pushed through the
// physical affine and the moving image is sampled once.  All coordinate arithmetic is float32 in the reference's
// order of operations; both interpolations are SciPy order-1 in float64.
struct FlowParams {
    float t[12];          // rows 0..2 of the 4x4 physical transform (z, y, x)
    float spacing[3];     // z, y, x
    float origin[3];
    float stride[3];      // flow grid stride, z, y, x
    float box_start[3];   // reference coordinate of the first flow sample, z, y, x
    int fz, fy, fx;       // flow grid dims
};

template <typename T, bool PRED>
__global__ void __launch_bounds__(256)
warp_flow_kernel(const T* __restrict__ in, const float* __restrict__ pred, int Z, int Y, int X,
                 const float* __restrict__ flow, FlowParams F, int OY, int OX, int oz0, size_t n_out,
                 float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_out) return;
    const size_t oplane = (size_t)OY * OX;
    const int oz = (int)(i / oplane);
    const size_t rem = i - (size_t)oz * oplane;
    const int y = (int)(rem / OX), x = (int)(rem - (size_t)y * OX);
    const float g[3] = {(float)(oz0 + oz), (float)y, (float)x};
    double fc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) fc[a] = (double)__fdiv_rn(__fsub_rn(g[a], F.box_start[a]), F.stride[a]);
    const size_t fvol = (size_t)F.fz * F.fy * F.fx;
    // channels are X, Y, Z displacements; deformed index = identity + flow (float32)
    const float dx = __fadd_rn(g[2], (float)sample_order1<float, false>(flow, nullptr, F.fz, F.fy, F.fx, fc));
    const float dy = __fadd_rn(g[1], (float)sample_order1<float, false>(flow + fvol, nullptr, F.fz, F.fy, F.fx, fc));
    const float dz = __fadd_rn(g[0], (float)sample_order1<float, false>(flow + 2 * fvol, nullptr, F.fz, F.fy, F.fx, fc));
    const float p[3] = {__fadd_rn(__fmul_rn(dz, F.spacing[0]), F.origin[0]), __fadd_rn(__fmul_rn(dy, F.spacing[1]), F.origin[1]),
                        __fadd_rn(__fmul_rn(dx, F.spacing[2]), F.origin[2])};
    double c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float m = __fmul_rn(F.t[4 * a], p[0]);
        m = __fadd_rn(m, __fmul_rn(F.t[4 * a + 1], p[1]));
        m = __fadd_rn(m, __fmul_rn(F.t[4 * a + 2], p[2]));
        m = __fadd_rn(m, F.t[4 * a + 3]);
        c[a] = (double)__fdiv_rn(__fsub_rn(m, F.origin[a]), F.spacing[a]);
    }
    out[i] = (float)sample_order1<T, PRED>(in, pred, Z, Y, X, c);
}

template <typename T, bool PRED>
__global__ void __launch_bounds__(256)
warp_affine_kernel(const T* __restrict__ in, const float* __restrict__ pred, int Z, int Y, int X, int oz0,
                   size_t n_out, AffineParams A, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_out) return;
    const size_t plane = (size_t)Y * X;
    const int oz = (int)(i / plane);
    const size_t rem = i - (size_t)oz * plane;
    const int y = (int)(rem / X), x = (int)(rem - (size_t)y * X);
    const double zc = (double)(oz0 + oz), yc = (double)y, xc = (double)x;
    const int dims[3] = {Z, Y, X};
    int st[3];
    double w[3][2];
    bool outside = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double c = __dadd_rn(A.off[a], __dmul_rn(zc, A.m[3 * a]));
        c = __dadd_rn(c, __dmul_rn(yc, A.m[3 * a + 1]));
        c = __dadd_rn(c, __dmul_rn(xc, A.m[3 * a + 2]));
        outside |= !(c >= 0.0 && c <= (double)(dims[a] - 1));  // NaN coordinates are outside too
        const double fl = floor(c);
        st[a] = (int)fl;
        const double f = __dadd_rn(c, -fl);
        w[a][0] = __dadd_rn(1.0, -f);
        w[a][1] = f;
    }
    double t = 0.0;
    if (!outside) {
#pragma unroll
        for (int dz = 0; dz < 2; ++dz)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int iz = st[0] + dz, iy = st[1] + dy, ix = st[2] + dx;
                    double v = 0.0;
                    if (iz < Z && iy < Y && ix < X) {  // lower bounds hold: coordinates are >= 0
                        const size_t j = (size_t)iz * plane + (size_t)iy * X + ix;
                        float fv = in_as_f32<T>(in, j);
                        if (PRED) fv = __fmul_rn(fv, __ldg(pred + j));
                        v = (double)fv;
                    }
                    const double term = __dmul_rn(__dmul_rn(__dmul_rn(v, w[0][dz]), w[1][dy]), w[2][dx]);
                    t = __dadd_rn(t, term);
                }
    }
    out[i] = (float)t;
}

}  // namespace

extern "C" int m3d_warp_affine(m3d_ctx* ctx, const void* in_dev, int in_dtype, const float* predictor_dev,
                               const int64_t dims[3], const double matrix_host[9], const double offset_host[3],
                               int64_t out_z0, int64_t out_nz, float* out_dev, void* stream) {
    if (!ctx || !in_dev || !out_dev || !dims || !matrix_host || !offset_host)
        return m3d_fail(M3D_ERR_ARG, "m3d_warp_affine: null argument");
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || dims[0] > 0x7fffffff || dims[1] > 0x7fffffff ||
        dims[2] > 0x7fffffff || out_nz <= 0 || out_z0 < -0x7fffffffll || out_z0 > 0x7fffffffll)
        return m3d_fail(M3D_ERR_ARG, "m3d_warp_affine: bad dims");
    if (in_dtype != M3D_DTYPE_U16 && in_dtype != M3D_DTYPE_F32)
        return m3d_fail(M3D_ERR_ARG, "m3d_warp_affine: dtype %d", in_dtype);
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    AffineParams A;
    for (int i = 0; i < 9; ++i) A.m[i] = matrix_host[i];
    for (int i = 0; i < 3; ++i) A.off[i] = offset_host[i];
    const int Z = (int)dims[0], Y = (int)dims[1], X = (int)dims[2];
    const size_t n_out = (size_t)out_nz * Y * X;
    const size_t blocks = (n_out + 255) / 256;
    if (blocks > 0x7fffffffull) return m3d_fail(M3D_ERR_ARG, "m3d_warp_affine: grid too large");
    KernelScope ks(ctx, KF_WARP_AFFINE, st);
    if (in_dtype == M3D_DTYPE_U16) {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(in_dev);
        if (predictor_dev)
            warp_affine_kernel<uint16_t, true><<<(unsigned)blocks, 256, 0, st>>>(p, predictor_dev, Z, Y, X, (int)out_z0, n_out, A, out_dev);
        else
            warp_affine_kernel<uint16_t, false><<<(unsigned)blocks, 256, 0, st>>>(p, nullptr, Z, Y, X, (int)out_z0, n_out, A, out_dev);
    } else {
        const float* p = reinterpret_cast<const float*>(in_dev);
        if (predictor_dev)
            warp_affine_kernel<float, true><<<(unsigned)blocks, 256, 0, st>>>(p, predictor_dev, Z, Y, X, (int)out_z0, n_out, A, out_dev);
        else
            warp_affine_kernel<float, false><<<(unsigned)blocks, 256, 0, st>>>(p, nullptr, Z, Y, X, (int)out_z0, n_out, A, out_dev);
    }
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

extern "C" int m3d_warp_flow(m3d_ctx* ctx, const void* in_dev, int in_dtype, const float* predictor_dev,
                             const int64_t dims[3], const float transform_host[16], const float spacing_host[3],
                             const float origin_host[3], const float* flow_dev, const int64_t flow_dims[3],
                             const float stride_zyx_host[3], const float box_start_zyx_host[3],
                             const int64_t out_dims[3], int64_t out_z0, int64_t out_nz, float* out_dev, void* stream) {
    if (!ctx || !in_dev || !out_dev || !dims || !transform_host || !spacing_host || !origin_host || !flow_dev ||
        !flow_dims || !stride_zyx_host || !box_start_zyx_host || !out_dims)
        return m3d_fail(M3D_ERR_ARG, "m3d_warp_flow: null argument");
    for (int a = 0; a < 3; ++a)
        if (dims[a] <= 0 || dims[a] > 0x7fffffff || flow_dims[a] <= 0 || flow_dims[a] > 0x7fffffff || out_dims[a] <= 0 ||
            out_dims[a] > 0x7fffffff)
            return m3d_fail(M3D_ERR_ARG, "m3d_warp_flow: bad dims");
    if (out_nz <= 0 || out_z0 < 0 || out_z0 + out_nz > out_dims[0]) return m3d_fail(M3D_ERR_ARG, "m3d_warp_flow: bad z range");
    if (in_dtype != M3D_DTYPE_U16 && in_dtype != M3D_DTYPE_F32) return m3d_fail(M3D_ERR_ARG, "m3d_warp_flow: dtype %d", in_dtype);
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    FlowParams F;
    for (int i = 0; i < 12; ++i) F.t[i] = transform_host[i];
    for (int a = 0; a < 3; ++a) {
        F.spacing[a] = spacing_host[a];
        F.origin[a] = origin_host[a];
        F.stride[a] = stride_zyx_host[a];
        F.box_start[a] = box_start_zyx_host[a];
    }
    F.fz = (int)flow_dims[0];
    F.fy = (int)flow_dims[1];
    F.fx = (int)flow_dims[2];
    const int Z = (int)dims[0], Y = (int)dims[1], X = (int)dims[2], OY = (int)out_dims[1], OX = (int)out_dims[2];
    const size_t n_out = (size_t)out_nz * OY * OX;
    const size_t blocks = (n_out + 255) / 256;
    if (blocks > 0x7fffffffull) return m3d_fail(M3D_ERR_ARG, "m3d_warp_flow: grid too large");
    KernelScope ks(ctx, KF_WARP_AFFINE, st);
#define M3D_WF(TT, PP, ptr) \
    warp_flow_kernel<TT, PP><<<(unsigned)blocks, 256, 0, st>>>(ptr, predictor_dev, Z, Y, X, flow_dev, F, OY, OX, (int)out_z0, n_out, out_dev)
    if (in_dtype == M3D_DTYPE_U16) {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(in_dev);
        if (predictor_dev) M3D_WF(uint16_t, true, p);
        else M3D_WF(uint16_t, false, p);
    } else {
        const float* p = reinterpret_cast<const float*>(in_dev);
        if (predictor_dev) M3D_WF(float, true, p);
        else M3D_WF(float, false, p);
    }
#undef M3D_WF
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

extern "C" int m3d_lowpass(m3d_ctx* ctx, const void* in_dev, int in_dtype, const float* predictor_dev,
                           int n_vols, const int64_t dims[3], const double sigma[3], int mode2d,
                           float* out_dev, void* stream) {
    if (!ctx || !in_dev || !out_dev || !dims || !sigma) return m3d_fail(M3D_ERR_ARG, "m3d_lowpass: null argument");
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || n_vols <= 0)
        return m3d_fail(M3D_ERR_ARG, "m3d_lowpass: bad dims");
    if (dims[0] > 0x7fffffff || dims[1] > 0x7fffffff || dims[2] > 0x7fffffff)
        return m3d_fail(M3D_ERR_ARG, "m3d_lowpass: dims exceed int32");
    for (int a = mode2d ? 1 : 0; a < 3; ++a)
        if (!(sigma[a] > 0.0)) return m3d_fail(M3D_ERR_ARG, "m3d_lowpass: sigma must be > 0 (inactive filters are the caller's no-op)");
    if ((const void*)out_dev == in_dev) return m3d_fail(M3D_ERR_ARG, "m3d_lowpass: in-place not supported");
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int Z = (int)dims[0], Y = (int)dims[1], X = (int)dims[2];
    if (in_dtype == M3D_DTYPE_U16)
        return lowpass_all<uint16_t>(ctx, reinterpret_cast<const uint16_t*>(in_dev), predictor_dev, n_vols, Z, Y, X,
                                     sigma, mode2d, out_dev, st);
    if (in_dtype == M3D_DTYPE_F32)
        return lowpass_all<float>(ctx, reinterpret_cast<const float*>(in_dev), predictor_dev, n_vols, Z, Y, X, sigma,
                                  mode2d, out_dev, st);
    return m3d_fail(M3D_ERR_ARG, "m3d_lowpass: dtype %d", in_dtype);
}

extern "C" int m3d_set_lowpass_mode(m3d_ctx* ctx, int mode) {
    if (!ctx) return m3d_fail(M3D_ERR_ARG, "m3d_set_lowpass_mode: null ctx");
    if (mode != 0 && mode != 1) return m3d_fail(M3D_ERR_ARG, "m3d_set_lowpass_mode: mode must be 0 (float64) or 1 (float32)");
    ctx->lowpass_f32 = mode;
    return M3D_OK;
}

extern "C" int m3d_weight(m3d_ctx* ctx, const uint16_t* readout_dev, const float* predictor_dev, int64_t n,
                          float* out_dev, void* stream) {
    if (!ctx || !readout_dev || !out_dev || n <= 0) return m3d_fail(M3D_ERR_ARG, "m3d_weight: bad argument");
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool aligned = ((reinterpret_cast<uintptr_t>(readout_dev) | reinterpret_cast<uintptr_t>(out_dev) |
                           reinterpret_cast<uintptr_t>(predictor_dev)) & 15u) == 0;
    const size_t n8 = aligned ? (size_t)n / 8 : 0;
    const size_t tail = (size_t)n - 8 * n8;
    const size_t threads = n8 > tail ? n8 : tail;  // every thread takes one vector and / or one tail element
    const size_t nblk = (threads + 255) / 256;
    if (nblk > 0x7fffffffull) return m3d_fail(M3D_ERR_ARG, "m3d_weight: %lld elements need too large a grid", (long long)n);
    const int blocks = (int)nblk;
    M3D_LAUNCH(ctx, KF_WEIGHT, st,
               weight_kernel<<<blocks, 256, 0, st>>>(readout_dev, predictor_dev, out_dev, (size_t)n, n8));
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}
