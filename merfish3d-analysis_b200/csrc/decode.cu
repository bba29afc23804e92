// decode.cu -- fused per-voxel barcode decode (replaces PD:2354-2643).
//
// Fast path (production, result images not requested), two kernels:
//   decode_gate_kernel   : streaming pass over the (bits,z,y,x) stack, 4 voxels per thread,
//                          128-bit / 64-bit vector loads.  scale -> clip -> L2 norm ->
//                          magnitude gates.  Writes decoded = -1 everywhere and appends the
//                          voxels that pass the gates to a candidate list.  HBM-bound:
//                          n_bits*sizeof(in) + 2 bytes per voxel.
//   decode_search_kernel : one thread per candidate; recomputes the trace, L2-normalises,
//                          finds the nearest codeword (codebook resident in shared memory,
//                          proxy score + exact float32 re-evaluation), applies the pixel
//                          gate and exclusions, overwrites decoded[v].
// Dense path (reference-complete images, return_results=True):
//   decode_dense_kernel  : one thread per voxel, search everywhere, writes decoded int16 and
//                          magnitude / distance / scaled float16 after round(.,5).
#include "voxel_math.cuh"

namespace {

constexpr int GATE_THREADS = 256;
constexpr int SEARCH_THREADS = 128;

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
        float4 v = __ldcs(reinterpret_cast<const float4*>(p));
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
};
template <>
struct Vec4<uint16_t> {
    static __device__ __forceinline__ void load(const uint16_t* p, float (&o)[4]) {
        uint2 v = __ldcs(reinterpret_cast<const uint2*>(p));
        // exact uint16 -> float32 through the 2^23 magic number (PRMT + FADD, no I2F)
        o[0] = __uint_as_float(__byte_perm(v.x, 0x4B000000u, 0x7410)) - 8388608.0f;
        o[1] = __uint_as_float(__byte_perm(v.x, 0x4B000000u, 0x7432)) - 8388608.0f;
        o[2] = __uint_as_float(__byte_perm(v.y, 0x4B000000u, 0x7410)) - 8388608.0f;
        o[3] = __uint_as_float(__byte_perm(v.y, 0x4B000000u, 0x7432)) - 8388608.0f;
    }
};

template <typename T>
struct IsFloatIn { static constexpr bool value = false; };
template <>
struct IsFloatIn<float> { static constexpr bool value = true; };

// ------------------------------------------------------------------ gate (streaming) kernel
template <typename T, int NB, bool SAFE>
__global__ void __launch_bounds__(GATE_THREADS)
decode_gate_kernel(const T* __restrict__ stack, size_t n_vox, DecodeParams P,
                   int16_t* __restrict__ decoded, uint32_t* __restrict__ cand,
                   unsigned int* __restrict__ cand_count) {
    constexpr bool NANCLIP = IsFloatIn<T>::value;
    const size_t v0 = ((size_t)blockIdx.x * GATE_THREADS + threadIdx.x) * 4;
    const bool active = v0 < n_vox;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
        float raw[NB][4];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b < P.n_bits) Vec4<T>::load(stack + (size_t)b * n_vox + v0, raw[b]);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b < P.n_bits) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float x = scale_clip<SAFE, NANCLIP>(raw[b][j], P.bkg[b], P.nrm[b], P.rcp[b], P.use_norm);
                    float s = __fmul_rn(x, x);
                    acc[j] = (b == 0) ? s : __fadd_rn(acc[j], s);
                }
            }
        }
        // decoded = -1 for the whole vector; the search kernel overwrites decoded voxels
        uint2 neg;
        neg.x = 0xFFFFFFFFu;
        neg.y = 0xFFFFFFFFu;
        __stcs(reinterpret_cast<uint2*>(decoded + v0), neg);
    }
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float mag = __fsqrt_rn(acc[j]);
        // zero norm is reported as -1 (PD:2462) and can only pass when mag_lo <= -1
        if (acc[j] == 0.f) mag = -1.f;
        bool pass = active && (mag >= P.mag_lo) && (mag <= P.mag_hi);
        unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            unsigned base = 0;
            if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(cand_count, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (pass) cand[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)(v0 + j);
        }
    }
}

// scalar variant for volumes whose voxel count is not a multiple of 4
template <typename T, int NB, bool SAFE>
__global__ void __launch_bounds__(GATE_THREADS)
decode_gate_scalar_kernel(const T* __restrict__ stack, size_t n_vox, DecodeParams P,
                          int16_t* __restrict__ decoded, uint32_t* __restrict__ cand,
                          unsigned int* __restrict__ cand_count) {
    constexpr bool NANCLIP = IsFloatIn<T>::value;
    const size_t v = (size_t)blockIdx.x * GATE_THREADS + threadIdx.x;
    const bool active = v < n_vox;
    float acc = 0.f;
    if (active) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b < P.n_bits) {
                float x = scale_clip<SAFE, NANCLIP>(load_elem(stack, (size_t)b * n_vox + v), P.bkg[b],
                                                    P.nrm[b], P.rcp[b], P.use_norm);
                float s = __fmul_rn(x, x);
                acc = (b == 0) ? s : __fadd_rn(acc, s);
            }
        }
        decoded[v] = -1;
    }
    float mag = __fsqrt_rn(acc);
    if (acc == 0.f) mag = -1.f;
    bool pass = active && (mag >= P.mag_lo) && (mag <= P.mag_hi);
    unsigned m = __ballot_sync(0xffffffffu, pass);
    const unsigned lane = threadIdx.x & 31u;
    if (m) {
        unsigned base = 0;
        if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(cand_count, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (pass) cand[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)v;
    }
}

// ------------------------------------------------------------------ shared-memory codebook staging
struct SearchSmem {
    float* xs;       // [(NB+1)][BLOCK]
    uint8_t* on;     // [K][max_on]
    float* a;        // [K]
    float* g;        // [K]
};

template <int NB, int BLOCK>
__device__ __forceinline__ SearchSmem stage_codebook(unsigned char* smem, const DecodeParams& P) {
    SearchSmem s;
    s.xs = reinterpret_cast<float*>(smem);
    s.a = s.xs + (NB + 1) * BLOCK;
    s.g = s.a + P.K;
    s.on = reinterpret_cast<uint8_t*>(s.g + P.K);
    if (P.binary) {
        for (int i = threadIdx.x; i < P.K; i += BLOCK) {
            s.a[i] = P.cw_a[i];
            s.g[i] = P.cw_g[i];
        }
        for (int i = threadIdx.x; i < P.K * P.max_on; i += BLOCK) s.on[i] = P.onbits[i];
    }
    s.xs[NB * BLOCK + threadIdx.x] = 0.f;  // zero slot addressed by padded on-bit entries
    __syncthreads();
    return s;
}

static size_t search_smem_bytes(int nb, int block, int K, int max_on) {
    return (size_t)(nb + 1) * block * 4 + (size_t)K * 8 + (size_t)K * max_on + 16;
}

// full per-voxel decode from the raw trace; shared by search and dense kernels
template <typename T, int NB, bool SAFE, int BLOCK>
__device__ __forceinline__ void decode_one(const T* __restrict__ stack, size_t n_vox, size_t v,
                                           const DecodeParams& P, const SearchSmem& S, float (&x)[NB],
                                           float& mag, float& d, int& k) {
    constexpr bool NANCLIP = IsFloatIn<T>::value;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        x[b] = 0.f;
        if (b < P.n_bits)
            x[b] = scale_clip<SAFE, NANCLIP>(load_elem(stack, (size_t)b * n_vox + v), P.bkg[b], P.nrm[b],
                                             P.rcp[b], P.use_norm);
    }
    float n = l2_norm<NB>(x);
    float xh[NB];
    mag = unit_vector<NB, SAFE>(x, n, xh);
    float* col = S.xs + threadIdx.x;
#pragma unroll
    for (int b = 0; b < NB; ++b) col[b * BLOCK] = xh[b];
    nearest_codeword<NB>(xh, P, col, BLOCK, S.on, S.a, S.g, d, k);
}

template <typename T, int NB, bool SAFE>
__global__ void __launch_bounds__(SEARCH_THREADS)
decode_search_kernel(const T* __restrict__ stack, size_t n_vox, DecodeParams P,
                     int16_t* __restrict__ decoded, const uint32_t* __restrict__ cand,
                     const unsigned int* __restrict__ cand_count) {
    extern __shared__ __align__(16) unsigned char smem[];
    SearchSmem S = stage_codebook<NB, SEARCH_THREADS>(smem, P);
    const unsigned n = *cand_count;
    for (unsigned i = blockIdx.x * SEARCH_THREADS + threadIdx.x; i < n; i += gridDim.x * SEARCH_THREADS) {
        size_t v = cand[i];
        float x[NB];
        float mag, d;
        int k;
        decode_one<T, NB, SAFE, SEARCH_THREADS>(stack, n_vox, v, P, S, x, mag, d, k);
        int16_t dec = apply_gates(d, k, mag, P);
        if (dec >= 0) decoded[v] = dec;
    }
}

template <typename T, int NB, bool SAFE>
__global__ void __launch_bounds__(SEARCH_THREADS)
decode_dense_kernel(const T* __restrict__ stack, size_t n_vox, DecodeParams P,
                    int16_t* __restrict__ decoded, __half* __restrict__ magnitude,
                    __half* __restrict__ distance, __half* __restrict__ scaled) {
    extern __shared__ __align__(16) unsigned char smem[];
    SearchSmem S = stage_codebook<NB, SEARCH_THREADS>(smem, P);
    for (size_t v = (size_t)blockIdx.x * SEARCH_THREADS + threadIdx.x; v < n_vox;
         v += (size_t)gridDim.x * SEARCH_THREADS) {
        float x[NB];
        float mag, d;
        int k;
        decode_one<T, NB, SAFE, SEARCH_THREADS>(stack, n_vox, v, P, S, x, mag, d, k);
        decoded[v] = apply_gates(d, k, mag, P);
        if (magnitude) magnitude[v] = round5_f16(mag);
        if (distance) distance[v] = round5_f16(d);
        if (scaled) {
#pragma unroll
            for (int b = 0; b < NB; ++b)
                if (b < P.n_bits) scaled[(size_t)b * n_vox + v] = round5_f16(x[b]);
        }
    }
}

template <typename T, int NB, bool SAFE>
int launch_decode(m3d_ctx* ctx, const T* stack, size_t n_vox, int16_t* decoded, __half* mag, __half* dist,
                  __half* scaled, cudaStream_t st) {
    DecodeParams P = ctx->params();
    const bool dense = (mag != nullptr) || (dist != nullptr) || (scaled != nullptr);
    if (dense) {
        size_t smem = search_smem_bytes(NB, SEARCH_THREADS, P.K, P.max_on);
        auto kern = decode_dense_kernel<T, NB, SAFE>;
        M3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        size_t want = (n_vox + SEARCH_THREADS - 1) / SEARCH_THREADS;
        size_t cap = (size_t)ctx->num_sms * 16;
        int blocks = (int)(want < cap ? want : cap);
        if (blocks < 1) blocks = 1;
        M3D_LAUNCH(ctx, KF_DECODE_DENSE, st,
                   kern<<<blocks, SEARCH_THREADS, smem, st>>>(stack, n_vox, P, decoded, mag, dist, scaled));
        M3D_CHECK_LAUNCH();
        return M3D_OK;
    }
    // fast path
    if (ctx->s_cand.ensure(n_vox * sizeof(uint32_t))) return M3D_ERR_CUDA;
    if (ctx->s_counters.ensure(256)) return M3D_ERR_CUDA;
    unsigned int* cand_count = reinterpret_cast<unsigned int*>(ctx->s_counters.ptr);
    uint32_t* cand = reinterpret_cast<uint32_t*>(ctx->s_cand.ptr);
    M3D_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(unsigned int), st));
    const bool vec = (n_vox % 4 == 0) && ((reinterpret_cast<uintptr_t>(stack) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(decoded) & 7u) == 0);
    if (vec) {
        size_t units = n_vox / 4;
        int blocks = (int)((units + GATE_THREADS - 1) / GATE_THREADS);
        M3D_LAUNCH(ctx, KF_DECODE_GATE, st,
                   decode_gate_kernel<T, NB, SAFE><<<blocks, GATE_THREADS, 0, st>>>(stack, n_vox, P, decoded, cand, cand_count));
    } else {
        int blocks = (int)((n_vox + GATE_THREADS - 1) / GATE_THREADS);
        M3D_LAUNCH(ctx, KF_DECODE_GATE, st,
                   decode_gate_scalar_kernel<T, NB, SAFE><<<blocks, GATE_THREADS, 0, st>>>(stack, n_vox, P, decoded, cand, cand_count));
    }
    M3D_CHECK_LAUNCH();
    {
        size_t smem = search_smem_bytes(NB, SEARCH_THREADS, P.K, P.max_on);
        auto kern = decode_search_kernel<T, NB, SAFE>;
        M3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int blocks = ctx->num_sms * 8;
        M3D_LAUNCH(ctx, KF_DECODE_SEARCH, st,
                   kern<<<blocks, SEARCH_THREADS, smem, st>>>(stack, n_vox, P, decoded, cand, cand_count));
        M3D_CHECK_LAUNCH();
    }
    return M3D_OK;
}

template <typename T, bool SAFE>
int dispatch_nb(m3d_ctx* ctx, const T* stack, size_t n_vox, int16_t* decoded, __half* mag, __half* dist,
                __half* scaled, cudaStream_t st) {
    switch (ctx->nb_pad) {
        case 16: return launch_decode<T, 16, SAFE>(ctx, stack, n_vox, decoded, mag, dist, scaled, st);
        case 24: return launch_decode<T, 24, SAFE>(ctx, stack, n_vox, decoded, mag, dist, scaled, st);
        case 32: return launch_decode<T, 32, SAFE>(ctx, stack, n_vox, decoded, mag, dist, scaled, st);
    }
    return m3d_fail(M3D_ERR_ARG, "unsupported padded bit count %d", ctx->nb_pad);
}

}  // namespace

extern "C" int m3d_decode(m3d_ctx* ctx, const void* stack_dev, int dtype, const int64_t dims[3],
                          int16_t* decoded_dev, uint16_t* magnitude_f16_dev, uint16_t* distance_f16_dev,
                          uint16_t* scaled_f16_dev, void* stream) {
    if (!ctx || !stack_dev || !decoded_dev || !dims) return m3d_fail(M3D_ERR_ARG, "m3d_decode: null argument");
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return m3d_fail(M3D_ERR_ARG, "m3d_decode: bad dims");
    const size_t n_vox = (size_t)dims[0] * dims[1] * dims[2];
    if (n_vox >= 0xFFFFFFF0ull) return m3d_fail(M3D_ERR_ARG, "m3d_decode: volume exceeds 2^32 voxels per call");
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    __half* mag = reinterpret_cast<__half*>(magnitude_f16_dev);
    __half* dist = reinterpret_cast<__half*>(distance_f16_dev);
    __half* scaled = reinterpret_cast<__half*>(scaled_f16_dev);
    const bool safe = ctx->use_norm && ctx->safe_div;
    if (dtype == M3D_DTYPE_U16) {
        const uint16_t* s = reinterpret_cast<const uint16_t*>(stack_dev);
        return safe ? dispatch_nb<uint16_t, true>(ctx, s, n_vox, decoded_dev, mag, dist, scaled, st)
                    : dispatch_nb<uint16_t, false>(ctx, s, n_vox, decoded_dev, mag, dist, scaled, st);
    } else if (dtype == M3D_DTYPE_F32) {
        const float* s = reinterpret_cast<const float*>(stack_dev);
        return safe ? dispatch_nb<float, true>(ctx, s, n_vox, decoded_dev, mag, dist, scaled, st)
                    : dispatch_nb<float, false>(ctx, s, n_vox, decoded_dev, mag, dist, scaled, st);
    }
    return m3d_fail(M3D_ERR_ARG, "m3d_decode: dtype %d", dtype);
}
