// decode.cu -- fused per-voxel barcode decode (replaces PD:2354-2643).
//
// Production path (result images not requested), two kernels:
//   decode_gate_kernel   : ONE streaming pass over the (bits,z,y,x) stack.  Every thread issues all
//                          of its 128-bit loads (one per bit plane) before touching any of them,
//                          evaluates a conservative squared-magnitude window, writes decoded = -1
//                          with a 128-bit store and appends possible foreground voxels to a
//                          candidate list (warp-aggregated).  It never decides a result: the
//                          window is widened by a proven margin so the list is a superset.
//                          HBM-bound: n_bits*sizeof(in) + 2 bytes per voxel.
//   decode_search_kernel : one lane per candidate.  Exact reference arithmetic (IEEE division,
//                          sequential sums), exact magnitude gates, nearest codeword with the
//                          codebook resident in shared memory:
//                            mode 2 (all rows share on-bit count w and value, e.g. MHD4): top-(w+1)
//                              selection in registers + hash lookup of the top-w mask; voxels that
//                              this cannot settle are finished warp-cooperatively (32 lanes split
//                              the codebook, shuffle arg-max, exact re-evaluation of the near-ties);
//                            mode 1 / 0: proxy scan with exact re-evaluation / direct scan.
//                          Winners overwrite decoded[v]; optionally the foreground list and the
//                          union-find slots for the labelling stage are emitted in the same pass.
// Dense path (reference-complete images, return_results=True):
//   decode_dense_kernel  : one thread per voxel, search everywhere, writes decoded int16 and
//                          magnitude / distance / scaled float16 after round(.,5).
#include <cuda_bf16.h>

#include "voxel_math.cuh"

namespace {

constexpr int GATE_THREADS = 256;
constexpr int SEARCH_THREADS = 128;
constexpr int XS_STRIDE = SEARCH_THREADS + 1;  // odd stride: conflict-free rows AND columns
constexpr int COOP_SWITCH = 12;                // more unresolved lanes than this -> lane-parallel scan

// ------------------------------------------------------------------ vector access
template <typename T, int VPT>
struct Vec;
template <>
struct Vec<uint16_t, 8> {
    using type = uint4;
    static __device__ __forceinline__ float get(const uint4& v, int j) {
        const uint32_t w = (j < 2) ? v.x : (j < 4) ? v.y : (j < 6) ? v.z : v.w;
        // exact uint16 -> float32 through the 2^23 magic number (PRMT + FADD, no I2F)
        return __uint_as_float(__byte_perm(w, 0x4B000000u, (j & 1) ? 0x7432 : 0x7410)) - 8388608.0f;
    }
};
template <>
struct Vec<uint16_t, 4> {
    using type = uint2;
    static __device__ __forceinline__ float get(const uint2& v, int j) {
        const uint32_t w = (j < 2) ? v.x : v.y;
        return __uint_as_float(__byte_perm(w, 0x4B000000u, (j & 1) ? 0x7432 : 0x7410)) - 8388608.0f;
    }
};
template <>
struct Vec<uint16_t, 1> {
    using type = unsigned short;
    static __device__ __forceinline__ float get(const unsigned short& v, int) { return (float)v; }
};
template <>
struct Vec<float, 4> {
    using type = float4;
    static __device__ __forceinline__ float get(const float4& v, int j) {
        return (j == 0) ? v.x : (j == 1) ? v.y : (j == 2) ? v.z : v.w;
    }
};
template <>
struct Vec<float, 2> {
    using type = float2;
    static __device__ __forceinline__ float get(const float2& v, int j) { return j == 0 ? v.x : v.y; }
};
template <>
struct Vec<float, 1> {
    using type = float;
    static __device__ __forceinline__ float get(const float& v, int) { return v; }
};

template <int VPT>
__device__ __forceinline__ void store_background(int16_t* p);
template <>
__device__ __forceinline__ void store_background<8>(int16_t* p) {
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
}
template <>
__device__ __forceinline__ void store_background<4>(int16_t* p) {
    __stcs(reinterpret_cast<uint2*>(p), make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu));
}
template <>
__device__ __forceinline__ void store_background<2>(int16_t* p) {
    __stcs(reinterpret_cast<unsigned int*>(p), 0xFFFFFFFFu);
}
template <>
__device__ __forceinline__ void store_background<1>(int16_t* p) {
    *p = (int16_t)-1;
}

// ------------------------------------------------------------------ gate (streaming) kernel
template <typename T, int NB, int VPT>
__global__ void __launch_bounds__(GATE_THREADS)
decode_gate_kernel(const T* __restrict__ stack, size_t n_vox, size_t n_units, GateParams G,
                   int16_t* __restrict__ decoded, uint32_t* __restrict__ cand,
                   unsigned int* __restrict__ cand_count, int write_background) {
    using V = typename Vec<T, VPT>::type;
    const size_t unit = (size_t)blockIdx.x * GATE_THREADS + threadIdx.x;
    const bool active = unit < n_units;
    const size_t v0 = (active ? unit : n_units - 1) * VPT;  // clamped: loads stay unpredicated
    V raw[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int pb = (b < G.n_bits) ? b : (G.n_bits - 1);  // padding planes re-read the last one (L1 hit)
        raw[b] = __ldcs(reinterpret_cast<const V*>(stack + (size_t)pb * n_vox + v0));
    }
    float acc[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) acc[j] = 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const float bg = G.bkg[b], rc = G.rcp[b];  // rcp == 0 on padding planes
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            float q = __fmul_rn(__fsub_rn(Vec<T, VPT>::get(raw[b], j), bg), rc);
            q = fminf(fmaxf(q, 0.f), 1.f);
            acc[j] = __fmaf_rn(q, q, acc[j]);
        }
    }
    if (active && write_background) store_background<VPT>(decoded + v0);  // skipped when the image is already clean
    unsigned bits = 0;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const bool pass = G.all_candidates || (acc[j] >= G.lo2 && acc[j] <= G.hi2);
        bits |= (pass ? 1u : 0u) << j;
    }
    if (!active) bits = 0;
    const unsigned cnt = __popc(bits);
    if (__any_sync(0xffffffffu, cnt != 0)) {
        const unsigned lane = threadIdx.x & 31u;
        unsigned incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += n;
        }
        unsigned base = 0;
        if (lane == 31) base = atomicAdd(cand_count, incl);
        base = __shfl_sync(0xffffffffu, base, 31);
        unsigned pos = base + incl - cnt;
#pragma unroll
        for (int j = 0; j < VPT; ++j)
            if ((bits >> j) & 1u) cand[pos++] = (uint32_t)(v0 + j);
    }
}

// ------------------------------------------------------------------ shared-memory codebook staging
struct SearchSmem {
    float* xs;           // [(NB+1)][XS_STRIDE]  xh of every thread's voxel; row NB is the zero slot
    float* a;            // [K]   (mode 1/2 proxy scan)
    float* g;            // [K]
    uint32_t* hkeys;     // [1 << hash_bits] (mode 2)
    uint32_t* masks;     // [round_up(K, 8)] on-bit masks, 0 beyond K (mode 2, tensor-core on-bit sums)
    uint4* afrag;        // [ceil(K/16)][KS][32] A fragments (codeword rows x bit columns, bf16 0/1) of the on-bit matrix
    uint32_t* cand_bits; // [warps][ceil(K/16)][4][4] ballot words of the tensor-core marking pass (mode 2)
    uint16_t* cand_cum;  // [warps][ceil(K/16)][32] per voxel: marked codewords in the tiles before tile i
    int16_t* hvals;      // [1 << hash_bits]
    uint8_t* on;         // [K][max_on]
};

static size_t search_smem_bytes(int nb, const DecodeParams& P) {
    size_t n = (size_t)(nb + 1) * XS_STRIDE * 4;
    if (P.mode >= 1) n += (size_t)P.K * 8 + (size_t)P.K * P.max_on;
    if (P.mode == 2)
        n += ((size_t)1 << P.hash_bits) * 6 + (size_t)((P.K + 7) & ~7) * 4 +
             (size_t)(SEARCH_THREADS / 32) * ((P.K + 15) / 16) * (64 + 64) +
             (size_t)((P.K + 15) / 16) * ((nb + 15) / 16) * 32 * 16 + 16;
    return n + 16;
}

template <int NB>
__device__ __forceinline__ SearchSmem stage_codebook(unsigned char* smem, const DecodeParams& P) {
    SearchSmem s;
    s.xs = reinterpret_cast<float*>(smem);
    float* p = s.xs + (NB + 1) * XS_STRIDE;
    s.a = s.g = nullptr;
    s.hkeys = nullptr;
    s.masks = nullptr;
    s.cand_bits = nullptr;
    s.cand_cum = nullptr;
    s.afrag = nullptr;
    s.hvals = nullptr;
    s.on = nullptr;
    if (P.mode >= 1) {
        s.a = p;
        s.g = p + P.K;
        p += 2 * P.K;
        for (int i = threadIdx.x; i < P.K; i += SEARCH_THREADS) {
            s.a[i] = P.cw_a[i];
            s.g[i] = P.cw_g[i];
        }
    }
    if (P.mode == 2) {
        const int hs = 1 << P.hash_bits;
        const int kpad = (P.K + 7) & ~7;
        constexpr int KS = (NB + 15) / 16;
        const int n_mt = (P.K + 15) >> 4;
        while (reinterpret_cast<uintptr_t>(p) & 15u) ++p;  // 16-byte alignment for the fragment table and the ballot words
        s.afrag = reinterpret_cast<uint4*>(p);
        // A fragment of mma.m16n8k16 (row-major 16 x 16, bf16): lane (g = lane / 4, t = lane % 4) holds
        // a0 = (row g, cols 2t, 2t+1), a1 = (row g+8, same cols), a2 = (row g, cols 2t+8, 2t+9), a3 = (row g+8, ...);
        // columns = bits 16 ks + ...; 0x3F80 = bf16(1.0); rows >= K are zero.  Row g of tile i is codeword 16 i + 2 g and
        // row g + 8 is codeword 16 i + 2 g + 1, so that the ballot words of the marking pass interleave into ascending k.
        for (int e = threadIdx.x; e < n_mt * KS * 32; e += SEARCH_THREADS) {
            const int lane = e & 31, ks = (e >> 5) % KS, i = (e >> 5) / KS;
            const int g = lane >> 2, t = lane & 3;
            const int k0 = 16 * i + 2 * g, k1 = k0 + 1;
            const uint32_t m0 = (k0 < P.K) ? P.cw_mask[k0] : 0u, m1 = (k1 < P.K) ? P.cw_mask[k1] : 0u;
            const int c = 16 * ks + 2 * t;
            auto pack = [](uint32_t m, int col) -> uint32_t {
                return (((m >> col) & 1u) ? 0x3F80u : 0u) | (((m >> (col + 1)) & 1u) ? 0x3F800000u : 0u);
            };
            s.afrag[e] = make_uint4(pack(m0, c), pack(m1, c), pack(m0, c + 8), pack(m1, c + 8));
        }
        s.cand_bits = reinterpret_cast<uint32_t*>(s.afrag + n_mt * KS * 32);
        s.cand_cum = reinterpret_cast<uint16_t*>(s.cand_bits + (SEARCH_THREADS / 32) * n_mt * 16);
        s.hkeys = reinterpret_cast<uint32_t*>(s.cand_cum + (SEARCH_THREADS / 32) * n_mt * 32);
        s.masks = s.hkeys + hs;
        s.hvals = reinterpret_cast<int16_t*>(s.masks + kpad);
        for (int i = threadIdx.x; i < hs; i += SEARCH_THREADS) {
            s.hkeys[i] = P.hash_keys[i];
            s.hvals[i] = P.hash_vals[i];
        }
        for (int i = threadIdx.x; i < kpad; i += SEARCH_THREADS) s.masks[i] = (i < P.K) ? P.cw_mask[i] : 0u;
        s.on = reinterpret_cast<uint8_t*>(s.hvals + hs);
    } else if (P.mode == 1) {
        s.on = reinterpret_cast<uint8_t*>(p);
    }
    if (P.mode >= 1)
        for (int i = threadIdx.x; i < P.K * P.max_on; i += SEARCH_THREADS) s.on[i] = P.onbits[i];
    s.xs[NB * XS_STRIDE + threadIdx.x] = 0.f;  // zero slot addressed by padded on-bit entries
    __syncthreads();
    return s;
}

// ------------------------------------------------------------------ exact per-voxel trace
// raw values of one voxel: every bit-plane load is issued before the (branchy) IEEE divisions
template <typename T, int NB>
__device__ __forceinline__ void load_stack_trace(const T* __restrict__ stack, size_t n_vox, size_t v,
                                                 const DecodeParams& P, float (&s)[NB]) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int pb = (b < P.n_bits) ? b : (P.n_bits - 1);
        s[b] = load_elem(stack, (size_t)pb * n_vox + v);
    }
}

template <int NB, bool INT_IN>
__device__ __forceinline__ void finish_trace(const float (&s)[NB], const DecodeParams& P, float (&x)[NB],
                                             float (&xh)[NB], float& mag) {
    if (INT_IN && P.rcp_all) {
        // every bit has a guarded reciprocal and integer samples need no operand check: straight-line hot path
#pragma unroll
        for (int b = 0; b < NB; ++b)
            x[b] = (b < P.n_bits) ? clip01_nan(div_by_rcp(__fsub_rn(s[b], P.bkg[b]), P.nrm[b], P.rcp[b])) : 0.f;
    } else {
#pragma unroll
        for (int b = 0; b < NB; ++b)
            x[b] = (b < P.n_bits) ? scale_clip<INT_IN>(s[b], P.bkg[b], P.nrm[b], P.rcp[b]) : 0.f;
    }
    const float n = l2_norm<NB>(x);
    mag = unit_vector<NB>(x, n, xh, INT_IN && P.rcp_all);
}

template <typename T, int NB>
__device__ __forceinline__ void exact_trace(const T* __restrict__ stack, size_t n_vox, size_t v,
                                            const DecodeParams& P, float (&x)[NB], float (&xh)[NB], float& mag) {
    float s[NB];
    load_stack_trace<T, NB>(stack, n_vox, v, P, s);
    finish_trace<NB, sizeof(T) == 2>(s, P, x, xh, mag);  // 2-byte samples = uint16: integers in [0, 65535]
}

// ------------------------------------------------------------------ lane-parallel scans (modes 0, 1, 2)
// Returns first-argmin index and direct-form distance.  col = this thread's xs column.
template <int NB>
__device__ __forceinline__ void scan_codebook(const float (&xh)[NB], const DecodeParams& P, const SearchSmem& S,
                                              const float* __restrict__ col, float& d_out, int& k_out) {
    float best_d = __int_as_float(0x7f800000);
    int best_k = 0;
    if (P.mode >= 1) {
        float best_p = __int_as_float(0x7f800000);
        const int mo = P.max_on;
        for (int k = 0; k < P.K; ++k) {
            const uint8_t* on = S.on + k * mo;
            float sum = 0.f;
            for (int j = 0; j < mo; ++j) sum += col[on[j] * XS_STRIDE];
            const float p = S.a[k] - S.g[k] * sum;
            if (p <= best_p + M3D_PROXY_MARGIN) {
                const float d = direct_distance_binary<NB>(xh, __ldg(P.cw_mask + k), __ldg(P.cw_c + k));
                if (d < best_d) {
                    best_d = d;
                    best_k = k;
                }
            }
            best_p = fminf(best_p, p);
        }
        // NaN traces: every comparison is false; NumPy's argmin returns the first NaN (index 0)
        if (!(best_d == best_d) || best_d == __int_as_float(0x7f800000)) {
            best_d = direct_distance<NB>(xh, P.codebook);
            best_k = 0;
        }
    } else {
        for (int k = 0; k < P.K; ++k) {
            const float d = direct_distance<NB>(xh, P.codebook + (size_t)k * M3D_MAX_BITS);
            if (d < best_d || k == 0) {
                best_d = d;
                best_k = k;
            }
        }
    }
    d_out = best_d;
    k_out = best_k;
}

// ------------------------------------------------------------------ mode 2 fast path
// Top-(w+1) selection on order-preserving integer keys (xh >= 0): the low 5 mantissa bits carry
// the bit index, so every key is unique and "remove the maximum" is an equality test.  If the
// top-w on-bit set is a codeword and beats every other w-subset by more than the float32
// uncertainty (T_w - T_{w+1} > M3D_SUM_MARGIN), it is THE argmin; its distance is then evaluated
// with the exact direct form.  Returns false when the voxel needs a real search.
template <int NB>
__device__ __forceinline__ bool topw_lookup(const float (&xh)[NB], const DecodeParams& P, const SearchSmem& S,
                                            float& d_out, int& k_out) {
    uint32_t key[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) key[b] = (__float_as_uint(xh[b]) & ~31u) | (uint32_t)b;
    uint32_t mask = 0, kmax = 0;
    float last = 0.f;
    const int w = P.max_on;
    for (int i = 0; i <= w; ++i) {
        kmax = key[0];
#pragma unroll
        for (int b = 1; b < NB; ++b) kmax = max(kmax, key[b]);
#pragma unroll
        for (int b = 0; b < NB; ++b) key[b] = (key[b] == kmax) ? 0u : key[b];
        if (i < w) {
            mask |= 1u << (kmax & 31u);
            last = __uint_as_float(kmax & ~31u);
        }
    }
    const float next = __uint_as_float(kmax & ~31u);
    if (!(last - next > M3D_SUM_MARGIN)) return false;
    const uint32_t hm = (1u << P.hash_bits) - 1u;
    uint32_t h = (mask * 2654435761u) >> (32 - P.hash_bits);
    while (true) {
        const uint32_t kk = S.hkeys[h];
        if (kk == mask) break;
        if (kk == 0u) return false;
        h = (h + 1u) & hm;
    }
    k_out = S.hvals[h];
    d_out = direct_distance_binary<NB>(xh, mask, P.cval);
    return true;
}

// Warp-cooperative search for the voxel held by lane `src` (mode 2): the 32 lanes split the
// codebook, shuffle-reduce the maximum on-bit sum, re-evaluate every codeword within the margin
// with the exact direct form and shuffle-reduce (distance, index) lexicographically.
__device__ __forceinline__ void coop_search(int src, const DecodeParams& P, const SearchSmem& S, int warp_col0,
                                            float& d_out, int& k_out) {
    const unsigned lane = threadIdx.x & 31u;
    const float* col = S.xs + warp_col0 + src;
    const int mo = P.max_on;
    float smax = -1.f;
    for (int k = lane; k < P.K; k += 32) {
        const uint8_t* on = S.on + k * mo;
        float sum = 0.f;
        for (int j = 0; j < mo; ++j) sum += col[on[j] * XS_STRIDE];
        smax = fmaxf(smax, sum);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    float best_d = __int_as_float(0x7f800000);
    int best_k = 0x7fffffff;
    for (int k = lane; k < P.K; k += 32) {
        const uint8_t* on = S.on + k * mo;
        float sum = 0.f;
        for (int j = 0; j < mo; ++j) sum += col[on[j] * XS_STRIDE];
        if (sum >= smax - M3D_SUM_MARGIN) {
            const float d = direct_distance_binary_smem(col, XS_STRIDE, P.n_bits, __ldg(P.cw_mask + k), P.cval);
            if (d < best_d) {  // k ascends within a lane: ties keep the lower index
                best_d = d;
                best_k = k;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, best_d, o);
        const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
        if (od < best_d || (od == best_d && ok < best_k)) {
            best_d = od;
            best_k = ok;
        }
    }
    d_out = best_d;
    k_out = best_k;
}

// ------------------------------------------------------------------ mode 2, dense regime: tensor-core on-bit sums
// When most lanes of a warp need a real search (every voxel passes the magnitude gate: the optimiser's first
// iteration with percentile-seeded vectors, reference-complete result images, all-foreground data), the
// voxels x bits x codewords contraction  S = Xh . M^T  (M = 0/1 on-bit matrix) runs on the tensor cores:
// warp-level mma.sync m16n8k16, bf16 operands, float32 accumulators, everything in registers.  Xh is split
// into bf16 hi + lo parts (M is exact in bf16), so |S~ - S| <= w * 2^-16 + accumulation noise =: eps.
// The MMA result never decides: pass 1 finds each voxel's largest S~, pass 2 recomputes the tiles and MARKS
// every codeword with S~ >= max - (M3D_SUM_MARGIN + 2 eps) -- a superset of the codewords coop_search would
// re-evaluate -- as warp-ballot words in shared memory; the warp then evaluates the marked (voxel, codeword) pairs
// with the exact float32 direct form, one pair per lane and round (evaluate_marked_pairs_warp), keeping every
// voxel's lexicographic (d, k) minimum = NumPy's first arg-min.  Clipped traces make exact ties among a dozen
// codewords common (noise-level normalisation saturates many bits at 1), so the candidate set really is a set,
// not a single winner -- and a very uneven one, which is why the pairs are shared out over the lanes.
// (tcgen05 / TMEM is not warranted here: the contraction is only 16-32 deep and the kernel is bound by the
// per-voxel IEEE arithmetic and the exact re-evaluations, not by MMA issue; the warp-level form keeps operands
// and accumulators in registers with no shared-memory descriptors or TMEM round trip.)
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    // cvt.rn.bf16x2.f32: both values in one full-rate instruction (the scalar conversion is an XU-pipe F2F each)
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);  // .x (low half) = x0
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xFFFF0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(__fsub_rn(x0, h0), __fsub_rn(x1, h1));
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void mma_bf16_m16n8k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// For the 32 voxels whose unit traces sit in this warp's xs columns: write into S.cand_bits (ballot words, layout at
// pass 2 below) the set of codewords that can be the exact arg-min (empty for NaN traces).  All lanes must call.
//
// Roles in D = A . B (m16n8k16): A = 16 codewords x 16 bits of the on-bit matrix (exact in bf16; the fragments are
// precomputed once per block, S.afrag: one 128-bit shared-memory load per tile), B = 16 bits x 8 voxels (the traces,
// bf16 hi + lo, built once per 32 voxels and kept in registers), D = on-bit sums of 16 codewords x 8 voxels.  Lane
// (g, t) then holds codewords 16 i + 2 g, 16 i + 2 g + 1 (fragment rows g, g + 8: see stage_codebook) for voxels
// 8 j + 2t, 8 j + 2t + 1.
template <int NB>
__device__ __forceinline__ void mma_mark_candidates_warp(const DecodeParams& P, const SearchSmem& S, int warp_col0) {
    constexpr int KS = (NB + 15) / 16;
    const unsigned lane = threadIdx.x & 31u;
    const int g = (int)(lane >> 2), t = (int)(lane & 3u);
    const int n_mt = (P.K + 15) >> 4;
    const float window = M3D_SUM_MARGIN + 2.f * ((float)P.max_on * 1.5259e-5f + 4.0e-6f);
    // B fragments (rows = bits 2t, 2t+1 | 2t+8, 2t+9 of the k step, column = voxel 8 j + g)
    uint32_t bhi[4][KS][2], blo[4][KS][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float* col = S.xs + warp_col0 + 8 * j + g;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int b = ks * 16 + h * 8 + 2 * t;
                const int r0 = (b < NB ? b : NB) * XS_STRIDE, r1 = (b + 1 < NB ? b + 1 : NB) * XS_STRIDE;  // NB = zero row
                split_bf16x2(col[r0], col[r1], bhi[j][ks][h], blo[j][ks][h]);
            }
        }
    }
    auto sums = [&](const uint4 (&a)[KS], int j, float (&d)[4]) {
        d[0] = d[1] = d[2] = d[3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const uint32_t af[4] = {a[ks].x, a[ks].y, a[ks].z, a[ks].w};
            mma_bf16_m16n8k16(d, af, bhi[j][ks][0], bhi[j][ks][1]);
            mma_bf16_m16n8k16(d, af, blo[j][ks][0], blo[j][ks][1]);
        }
    };
    // pass 1: largest on-bit sum of every voxel (sums are >= 0, padding rows give 0: no bound checks needed)
    float vmax[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) vmax[j][0] = vmax[j][1] = 0.f;
    const uint4* af = S.afrag + lane;
    for (int i = 0; i < n_mt; ++i) {
        uint4 a[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) a[ks] = af[(i * KS + ks) * 32];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float d[4];
            sums(a, j, d);
            vmax[j][0] = fmaxf(vmax[j][0], fmaxf(d[0], d[2]));
            vmax[j][1] = fmaxf(vmax[j][1], fmaxf(d[1], d[3]));
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float v = vmax[j][c];  // the other codeword rows of this voxel live in the lanes with the same t
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
            vmax[j][c] = v - window;  // from here on: the cut
        }
    // pass 2: the candidate sets.  Every comparison `sum >= cut` goes straight into a warp ballot: bit 4 g + t of the word
    // for accumulator slot q belongs to lane (g, t), i.e. to codeword row g (q = 0, 1) or g + 8 (q = 2, 3) of tile i and
    // voxel 8 j + 2 t + (q & 1).  Lane 0 stores the four words of (i, j) with one 128-bit store, ordered {q0, q2, q1, q3}
    // so that the two words a voxel needs are adjacent.  No shuffles, no atomics, no per-lane bit assembly; the sets need
    // no clearing beforehand.  NaN sums compare false: nothing is marked.  Padding rows (k >= K) are dropped by the reader.
    uint4* wb = reinterpret_cast<uint4*>(S.cand_bits) + (size_t)(warp_col0 >> 5) * n_mt * 4;
    for (int i = 0; i < n_mt; ++i) {
        uint4 a[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) a[ks] = af[(i * KS + ks) * 32];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float d[4];
            sums(a, j, d);
            const unsigned q0 = __ballot_sync(0xffffffffu, d[0] >= vmax[j][0]);
            const unsigned q1 = __ballot_sync(0xffffffffu, d[1] >= vmax[j][1]);
            const unsigned q2 = __ballot_sync(0xffffffffu, d[2] >= vmax[j][0]);
            const unsigned q3 = __ballot_sync(0xffffffffu, d[3] >= vmax[j][1]);
            if (lane == 0) wb[i * 4 + j] = make_uint4(q0, q2, q1, q3);
        }
    }
    __syncwarp();
}

// sqrt(sum_b (xh_b - c_b)^2) for one (voxel, codeword) pair, xh read from the voxel's shared-memory column: the direct
// form's own operations in its own order (x - 0 = x exactly, so the select moves into the subtrahend)
template <int NB>
__device__ __forceinline__ float pair_distance(const float* __restrict__ col, int n_bits, uint32_t mask, float c_val) {
    float acc = 0.f;
#pragma unroll
    for (int b0 = 0; b0 < NB; b0 += 8) {
        if (b0 < n_bits) {  // uniform; rows n_bits .. NB - 1 hold zeros
#pragma unroll
            for (int b = b0; b < b0 + 8; ++b) {
                const float t = __fsub_rn(col[b * XS_STRIDE], ((mask >> b) & 1u) ? c_val : 0.f);
                const float term = __fmul_rn(t, t);
                acc = (b == 0) ? term : __fadd_rn(acc, term);
            }
        }
    }
    return __fsqrt_rn(acc);
}

// Exact evaluation of the marked candidates of a warp's 32 voxels, shared by the 32 lanes.  The candidate sets are very
// uneven (mean ~3 codewords per voxel, ~25 for the unluckiest lane of a warp: clipped traces tie exactly), so instead of
// every lane walking its own set, the (voxel, codeword) pairs are numbered in voxel order / ascending k and lane p of a
// round evaluates pair 32 r + p from the voxel's shared-memory column.  Nothing is queued: a pair's voxel comes from a
// shuffle binary search over the inclusive counts, its tile from a binary search over the voxel's per-tile prefix
// counts (written by the owner lane while counting), its codeword from a rank select inside the tile's word -- all
// branch-free.  Lanes holding pairs of the same voxel are adjacent: a segmented shuffle scan keeps the lexicographic
// (distance, k) minimum (= NumPy's first arg-min) and every owner lane pulls the value of its segment's last lane
// (a later round only brings larger k: strict improvement only).
// `active`: this lane's voxel takes part.  Returns the owner's result in d / k (k = -1: nothing was marked) and the
// size of its candidate set.
template <int NB>
__device__ __forceinline__ void evaluate_marked_pairs_warp(bool active, const DecodeParams& P, const SearchSmem& S,
                                                           int warp_col0, float& d_out, int& k_out, unsigned& cnt_out) {
    const unsigned lane = threadIdx.x & 31u;
    const int n_mt = (P.K + 15) >> 4;
    const uint2* words = reinterpret_cast<const uint2*>(S.cand_bits) + (size_t)(warp_col0 >> 5) * n_mt * 8;
    uint16_t* cum = S.cand_cum + (size_t)(warp_col0 >> 5) * n_mt * 32;
    const int rem = P.K - 16 * (n_mt - 1);  // codewords in the last tile: 1 .. 16
    const uint32_t last_valid = (rem >= 16) ? 0x33333333u
                                            : ((0x33333333u & ((1u << (4 * (rem >> 1))) - 1u)) |
                                               ((rem & 1) ? (1u << (4 * (rem >> 1))) : 0u));
    // candidates of voxel (lane) vx in tile i: bit 4 g <-> codeword 16 i + 2 g, bit 4 g + 1 <-> 16 i + 2 g + 1 (ascending k)
    auto tile_word = [&](unsigned vx, int i) -> uint32_t {
        const uint2 q = words[i * 8 + (vx >> 3) * 2 + (vx & 1u)];
        const unsigned tsh = (vx >> 1) & 3u;
        const uint32_t w = ((q.x >> tsh) & 0x11111111u) | (((q.y >> tsh) & 0x11111111u) << 1);
        return (i == n_mt - 1) ? (w & last_valid) : w;
    };
    unsigned cnt = 0;
    for (int i = 0; i < n_mt; ++i) {
        cum[i * 32 + lane] = (uint16_t)cnt;
        cnt += __popc(tile_word(lane, i));
    }
    if (!active) cnt = 0;
    unsigned incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += n;
    }
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned excl = incl - cnt;
    __syncwarp();
    int top = 1;  // largest power of two below n_mt (binary search stride)
    while (2 * top < n_mt) top *= 2;
    const float* cols = S.xs + warp_col0;
    float best_d = __int_as_float(0x7f800000);
    int best_k = -1;
    for (unsigned q0 = 0; q0 < total; q0 += 32) {
        const unsigned q = q0 + lane;
        const bool have = q < total;
        // voxel of pair q = number of lanes whose inclusive count is <= q (31 at most while q < total)
        unsigned v = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const unsigned t = __shfl_sync(0xffffffffu, incl, (int)(v + step - 1));
            if (t <= q) v += step;
        }
        v &= 31u;
        unsigned r = q - __shfl_sync(0xffffffffu, excl, (int)v);  // rank of the pair inside the voxel's set
        float d = __int_as_float(0x7f800000);
        int k = 0;
        if (have) {
            // tile: the last i with cum[i][v] <= r
            int ti = 0;
            for (int step = top; step > 0; step >>= 1) {
                const int j = ti + step;
                if (j < n_mt && (unsigned)cum[j * 32 + v] <= r) ti = j;
            }
            r -= cum[ti * 32 + v];
            uint32_t w = tile_word(v, ti);
            // rank select: position of the r-th set bit (each nibble holds at most two, in bits 0 and 1)
            int pos = 0;
            unsigned c = __popc(w & 0xFFFFu);
            if (r >= c) {
                pos = 16;
                r -= c;
            }
            c = __popc((w >> pos) & 0xFFu);
            if (r >= c) {
                pos += 8;
                r -= c;
            }
            c = __popc((w >> pos) & 0xFu);
            if (r >= c) {
                pos += 4;
                r -= c;
            }
            pos += (r >= ((w >> pos) & 1u)) ? 1 : 0;
            k = 16 * ti + ((pos >> 1) | (pos & 1));
            d = pair_distance<NB>(cols + v, P.n_bits, S.masks[k], P.cval);
        }
        const unsigned vtag = have ? v : 0xFFFFu;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {  // lower lanes of a segment hold lower k: they win ties
            const float od = __shfl_up_sync(0xffffffffu, d, o);
            const int ok = __shfl_up_sync(0xffffffffu, k, o);
            const unsigned ov = __shfl_up_sync(0xffffffffu, vtag, o);
            if (lane >= (unsigned)o && ov == vtag && od <= d) {
                d = od;
                k = ok;
            }
        }
        // owner side: the last lane of my voxel's segment in this round, if it has one
        const bool mine = cnt != 0u && incl > q0 && excl < q0 + 32u;
        const unsigned tail = min(incl - 1u - q0, 31u);
        const float od = __shfl_sync(0xffffffffu, d, (int)(tail & 31u));
        const int ok = __shfl_sync(0xffffffffu, k, (int)(tail & 31u));
        if (mine && od < best_d) {
            best_d = od;
            best_k = ok;
        }
    }
    d_out = best_d;
    k_out = best_k;
    cnt_out = cnt;
}

// nearest codeword for the voxels of one warp; `want` = this lane holds a voxel to search.
// All 32 lanes must call (warp-synchronous).  `dense_mode` is the warp's memory between calls: while the previous
// batch went through the tensor-core path and few of its voxels had a single candidate, the top-w attempt (which
// would fail for most lanes) is skipped and every voxel goes straight to marking + pair evaluation.
template <int NB>
__device__ __forceinline__ void nearest_codeword_warp(bool want, const float (&xh)[NB], const DecodeParams& P,
                                                      const SearchSmem& S, float& d, int& k, bool& dense_mode) {
    float* col = S.xs + threadIdx.x;
#pragma unroll
    for (int b = 0; b < NB; ++b) col[b * XS_STRIDE] = xh[b];
    __syncwarp();
    if (P.mode == 2) {
        bool done = !want;
        const int n_want = __popc(__ballot_sync(0xffffffffu, want));
        if (!(dense_mode && n_want > COOP_SWITCH)) {
            if (want) done = topw_lookup<NB>(xh, P, S, d, k);
        }
        unsigned pending = __ballot_sync(0xffffffffu, !done);
        const int warp_col0 = (int)(threadIdx.x & ~31u);
        dense_mode = false;
        if (__popc(pending) > COOP_SWITCH) {
            // dense regime: on-bit sums of all 32 voxels on the tensor cores, exact distance for the settled ones
            mma_mark_candidates_warp<NB>(P, S, warp_col0);
            float ed;
            int ek;
            unsigned n_cand;
            evaluate_marked_pairs_warp<NB>(!done, P, S, warp_col0, ed, ek, n_cand);
            // voxels with a single candidate are the ones the top-w attempt settles: keep skipping it while they are few
            const int n_single = __popc(__ballot_sync(0xffffffffu, !done && n_cand == 1u));
            dense_mode = (__popc(pending) - n_single) > COOP_SWITCH;
            if (!done && ek >= 0) {
                k = ek;
                d = ed;
                done = true;
            }
            pending = __ballot_sync(0xffffffffu, !done);
            if (__popc(pending) > COOP_SWITCH) {
                if (!done) scan_codebook<NB>(xh, P, S, col, d, k);
                pending = 0u;
            }
        }
        {
            while (pending) {
                const int src = __ffs(pending) - 1;
                pending &= pending - 1u;
                float cd;
                int ck;
                coop_search(src, P, S, warp_col0, cd, ck);
                if ((int)(threadIdx.x & 31u) == src) {
                    d = cd;
                    k = ck;
                }
            }
        }
    } else if (want) {
        scan_codebook<NB>(xh, P, S, col, d, k);
    }
    __syncwarp();
}

// ------------------------------------------------------------------ candidate search kernel
struct FgSink {  // optional hand-off to the labelling / regionprops stages (m3d_decode_label)
    uint32_t* fg;
    unsigned int* fg_count;
    uint32_t* parent;
    uint32_t* aux;
    __half* rec_x;      // [rec_cap][NB] float16 scaled image values (round(.,5)) of foreground voxels
    uint32_t* rec_md;   // [rec_cap]     float16 magnitude | float16 distance << 16
    unsigned rec_cap;
};

template <typename T, int NB>
#ifndef M3D_SEARCH_MINB
#define M3D_SEARCH_MINB 5  // 96 registers, 20 warps per SM
#endif
__global__ void __launch_bounds__(SEARCH_THREADS, M3D_SEARCH_MINB)
decode_search_kernel(const T* __restrict__ stack, size_t n_vox, DecodeParams P, int16_t* __restrict__ decoded,
                     const uint32_t* __restrict__ cand, const unsigned int* __restrict__ cand_count, FgSink sink) {
    extern __shared__ __align__(16) unsigned char smem[];
    SearchSmem S = stage_codebook<NB>(smem, P);
    const unsigned n = *cand_count;
    const unsigned n_round = (n + 31u) & ~31u;
    bool dense_mode = false;
    for (unsigned i = blockIdx.x * SEARCH_THREADS + threadIdx.x; i < n_round; i += gridDim.x * SEARCH_THREADS) {
        const bool valid = i < n;
        const size_t v = valid ? cand[i] : 0;
        float x[NB], xh[NB];
        float mag = -1.f, d = 0.f;
        int k = 0;
        bool want = false;
        if (valid) {
            exact_trace<T, NB>(stack, n_vox, v, P, x, xh, mag);
            want = (mag >= P.mag_lo) && (mag <= P.mag_hi);  // exact magnitude gates (PD:2613-2614)
        } else {
#pragma unroll
            for (int b = 0; b < NB; ++b) xh[b] = 0.f;
        }
        nearest_codeword_warp<NB>(want, xh, P, S, d, k, dense_mode);
        bool fg = false;
        if (want) {
            const int16_t dec = apply_gates(d, k, mag, P);
            if (dec >= 0) {
                decoded[v] = dec;
                fg = true;
            }
        }
        if (sink.fg) {
            if (fg) {
                sink.parent[v] = (uint32_t)v;
                sink.aux[v] = 0u;
            }
            const unsigned m = __ballot_sync(0xffffffffu, fg);
            if (m) {
                const unsigned lane = threadIdx.x & 31u;
                unsigned base = 0;
                const int leader = __ffs(m) - 1;
                if ((int)lane == leader) base = atomicAdd(sink.fg_count, (unsigned)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (fg) {
                    const unsigned idx = base + __popc(m & ((1u << lane) - 1u));
                    sink.fg[idx] = (uint32_t)v;
                    if (idx < sink.rec_cap) {  // per-voxel values the regionprops stage would recompute
                        sink.rec_md[idx] = (uint32_t)__half_as_ushort(round5_f16(mag)) |
                                           ((uint32_t)__half_as_ushort(round5_f16(d)) << 16);
                        uint32_t hw[NB / 2];
#pragma unroll
                        for (int b = 0; b < NB; b += 2)
                            hw[b / 2] = (uint32_t)__half_as_ushort(round5_f16(x[b])) |
                                        ((uint32_t)__half_as_ushort(round5_f16(x[b + 1])) << 16);
                        uint4* dst = reinterpret_cast<uint4*>(sink.rec_x + (size_t)idx * NB);
#pragma unroll
                        for (int q = 0; q < NB / 8; ++q)
                            dst[q] = make_uint4(hw[4 * q], hw[4 * q + 1], hw[4 * q + 2], hw[4 * q + 3]);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------ dense kernel (all result images)
template <typename T, int NB>
__global__ void __launch_bounds__(SEARCH_THREADS)
decode_dense_kernel(const T* __restrict__ stack, size_t n_vox, DecodeParams P, int16_t* __restrict__ decoded,
                    __half* __restrict__ magnitude, __half* __restrict__ distance, __half* __restrict__ scaled) {
    extern __shared__ __align__(16) unsigned char smem[];
    SearchSmem S = stage_codebook<NB>(smem, P);
    const size_t n_round = (n_vox + 31u) & ~(size_t)31u;
    bool dense_mode = false;
    for (size_t v = (size_t)blockIdx.x * SEARCH_THREADS + threadIdx.x; v < n_round;
         v += (size_t)gridDim.x * SEARCH_THREADS) {
        const bool valid = v < n_vox;
        float x[NB], xh[NB];
        float mag = -1.f, d = 0.f;
        int k = 0;
        if (valid) {
            exact_trace<T, NB>(stack, n_vox, v, P, x, xh, mag);
        } else {
#pragma unroll
            for (int b = 0; b < NB; ++b) x[b] = xh[b] = 0.f;
        }
        // the reference stores the distance of EVERY voxel (PD:2630-2632): search everywhere.
        // NaN traces skip the ordered searches: NumPy's argmin of an all-NaN column is index 0.
        bool nan_trace = false;
#pragma unroll
        for (int b = 0; b < NB; ++b) nan_trace |= !(xh[b] == xh[b]);
        nearest_codeword_warp<NB>(valid && !nan_trace, xh, P, S, d, k, dense_mode);
        if (valid && nan_trace) {
            d = direct_distance<NB>(xh, P.codebook);
            k = 0;
        }
        if (valid) {
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
}

// ------------------------------------------------------------------ launch helpers
template <typename T, int NB, int VPT>
int launch_gate(m3d_ctx* ctx, const T* stack, size_t n_vox, int16_t* decoded, uint32_t* cand,
                unsigned int* cand_count, cudaStream_t st) {
    const GateParams G = ctx->gate_params();
    const size_t n_units = n_vox / VPT;
    const size_t blocks = (n_units + GATE_THREADS - 1) / GATE_THREADS;
    if (blocks > 0x7fffffffull) return m3d_fail(M3D_ERR_ARG, "m3d_decode: grid too large");
    M3D_LAUNCH(ctx, KF_DECODE_GATE, st,
               decode_gate_kernel<T, NB, VPT><<<(unsigned)blocks, GATE_THREADS, 0, st>>>(
                   stack, n_vox, n_units, G, decoded, cand, cand_count, ctx->gate_skip_background ? 0 : 1));
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

template <typename T>
struct GateWidth;
template <>
struct GateWidth<uint16_t> {  // 128-bit loads; halve when the register file would not hold all planes
    static constexpr int wide(int nb) { return nb <= 24 ? 8 : 4; }
    static constexpr int narrow(int nb) { return nb <= 24 ? 4 : 1; }
};
template <>
struct GateWidth<float> {
    static constexpr int wide(int nb) { return nb <= 24 ? 4 : 2; }
    static constexpr int narrow(int nb) { return nb <= 24 ? 2 : 1; }
};

template <typename T, int NB>
int run_gate(m3d_ctx* ctx, const T* stack, size_t n_vox, int16_t* decoded, uint32_t* cand,
             unsigned int* cand_count, cudaStream_t st) {
    constexpr int W = GateWidth<T>::wide(NB);
    constexpr int N = GateWidth<T>::narrow(NB);
    const uintptr_t sa = reinterpret_cast<uintptr_t>(stack), da = reinterpret_cast<uintptr_t>(decoded);
    auto ok = [&](int vpt) {
        return (n_vox % vpt == 0) && (sa % (vpt * sizeof(T)) == 0) && (da % (vpt * sizeof(int16_t)) == 0);
    };
    if (ok(W)) return launch_gate<T, NB, W>(ctx, stack, n_vox, decoded, cand, cand_count, st);
    if (N > 1 && ok(N)) return launch_gate<T, NB, N>(ctx, stack, n_vox, decoded, cand, cand_count, st);
    return launch_gate<T, NB, 1>(ctx, stack, n_vox, decoded, cand, cand_count, st);
}

template <typename T, int NB>
int launch_decode(m3d_ctx* ctx, const T* stack, size_t n_vox, int16_t* decoded, __half* mag, __half* dist,
                  __half* scaled, const FgSink& sink, cudaStream_t st) {
    const DecodeParams P = ctx->params();
    const size_t smem = search_smem_bytes(NB, P);
    if (smem > 200 * 1024) return m3d_fail(M3D_ERR_CAPACITY, "m3d_decode: codebook needs %zu B of shared memory", smem);
    const bool dense = (mag != nullptr) || (dist != nullptr) || (scaled != nullptr);
    if (dense) {
        auto kern = decode_dense_kernel<T, NB>;
        M3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const size_t want = (n_vox + SEARCH_THREADS - 1) / SEARCH_THREADS;
        const size_t cap = (size_t)ctx->num_sms * 16;
        const int blocks = (int)(want < cap ? want : cap);
        M3D_LAUNCH(ctx, KF_DECODE_DENSE, st,
                   kern<<<blocks, SEARCH_THREADS, smem, st>>>(stack, n_vox, P, decoded, mag, dist, scaled));
        M3D_CHECK_LAUNCH();
        return M3D_OK;
    }
    if (ctx->s_cand.ensure(n_vox * sizeof(uint32_t))) return M3D_ERR_CUDA;
    if (ctx->s_counters.ensure(256)) return M3D_ERR_CUDA;
    unsigned int* cand_count = reinterpret_cast<unsigned int*>(ctx->s_counters.ptr);
    uint32_t* cand = reinterpret_cast<uint32_t*>(ctx->s_cand.ptr);
    M3D_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(unsigned int), st));
    int rc = run_gate<T, NB>(ctx, stack, n_vox, decoded, cand, cand_count, st);
    if (rc) return rc;
    auto kern = decode_search_kernel<T, NB>;
    M3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = ctx->num_sms * 10;  // two rounds of the 5 resident blocks per SM
    M3D_LAUNCH(ctx, KF_DECODE_SEARCH, st,
               kern<<<blocks, SEARCH_THREADS, smem, st>>>(stack, n_vox, P, decoded, cand, cand_count, sink));
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

template <typename T>
int dispatch_nb(m3d_ctx* ctx, const T* stack, size_t n_vox, int16_t* decoded, __half* mag, __half* dist,
                __half* scaled, const FgSink& sink, cudaStream_t st) {
    switch (ctx->nb_pad) {
        case 8: return launch_decode<T, 8>(ctx, stack, n_vox, decoded, mag, dist, scaled, sink, st);
        case 16: return launch_decode<T, 16>(ctx, stack, n_vox, decoded, mag, dist, scaled, sink, st);
        case 24: return launch_decode<T, 24>(ctx, stack, n_vox, decoded, mag, dist, scaled, sink, st);
        case 32: return launch_decode<T, 32>(ctx, stack, n_vox, decoded, mag, dist, scaled, sink, st);
    }
    return m3d_fail(M3D_ERR_ARG, "unsupported padded bit count %d", ctx->nb_pad);
}

}  // namespace

// shared with extract.cu (m3d_decode_label)
int m3d_check_decode_args(m3d_ctx* ctx, const void* stack_dev, int dtype, const int64_t dims[3],
                          int16_t* decoded_dev, size_t* n_vox) {
    if (!ctx || !stack_dev || !decoded_dev || !dims) return m3d_fail(M3D_ERR_ARG, "m3d_decode: null argument");
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return m3d_fail(M3D_ERR_ARG, "m3d_decode: bad dims");
    if (dtype != M3D_DTYPE_U16 && dtype != M3D_DTYPE_F32) return m3d_fail(M3D_ERR_ARG, "m3d_decode: dtype %d", dtype);
    *n_vox = (size_t)dims[0] * dims[1] * dims[2];
    if (*n_vox >= 0xFFFFFFF0ull) return m3d_fail(M3D_ERR_ARG, "m3d_decode: volume exceeds 2^32 voxels per call");
    return M3D_OK;
}

int m3d_decode_internal(m3d_ctx* ctx, const void* stack_dev, int dtype, size_t n_vox, int16_t* decoded_dev,
                        uint32_t* fg, unsigned int* fg_count, uint32_t* parent, uint32_t* aux, void* rec_x,
                        uint32_t* rec_md, unsigned rec_cap, cudaStream_t st) {
    const FgSink sink{fg, fg_count, parent, aux, reinterpret_cast<__half*>(rec_x), rec_md, rec_cap};
    if (dtype == M3D_DTYPE_U16)
        return dispatch_nb<uint16_t>(ctx, reinterpret_cast<const uint16_t*>(stack_dev), n_vox, decoded_dev, nullptr,
                                     nullptr, nullptr, sink, st);
    return dispatch_nb<float>(ctx, reinterpret_cast<const float*>(stack_dev), n_vox, decoded_dev, nullptr, nullptr,
                              nullptr, sink, st);
}

extern "C" int m3d_decode(m3d_ctx* ctx, const void* stack_dev, int dtype, const int64_t dims[3],
                          int16_t* decoded_dev, uint16_t* magnitude_f16_dev, uint16_t* distance_f16_dev,
                          uint16_t* scaled_f16_dev, void* stream) {
    size_t n_vox = 0;
    int rc = m3d_check_decode_args(ctx, stack_dev, dtype, dims, decoded_dev, &n_vox);
    if (rc) return rc;
    ctx->prev_fg_valid = 0;  // this call may rewrite a persistent decoded image without leaving a foreground list
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    __half* mag = reinterpret_cast<__half*>(magnitude_f16_dev);
    __half* dist = reinterpret_cast<__half*>(distance_f16_dev);
    __half* scaled = reinterpret_cast<__half*>(scaled_f16_dev);
    const FgSink none{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0u};
    if (dtype == M3D_DTYPE_U16)
        return dispatch_nb<uint16_t>(ctx, reinterpret_cast<const uint16_t*>(stack_dev), n_vox, decoded_dev, mag, dist,
                                     scaled, none, st);
    return dispatch_nb<float>(ctx, reinterpret_cast<const float*>(stack_dev), n_vox, decoded_dev, mag, dist, scaled,
                              none, st);
}
