// extract.cu -- connected components, size filters and regionprops (replaces PD:2908-3062).
//
// Labelling (m3d_label).  The decoded image is sparse (a few % foreground), so everything
// after one streaming pass over it works on the compacted foreground list:
//   ccl_collect_kernel : stream decoded (int16, 128-bit loads), append foreground voxels to a
//                        list and initialise their union-find parent / aux slots.  Parent and
//                        aux are dense uint32 arrays indexed by voxel but only foreground
//                        entries are ever touched.
//   ccl_merge_kernel   : per foreground voxel, union with each equal-valued backward
//                        neighbour (13 of 26 in 3-D, 4 of 8 per plane in 2-D); lock-free
//                        atomicMin union-find, so every root is the component's smallest
//                        linear index = the raster-first voxel = the canonical identity.
//   ccl_compress_kernel: flatten to roots, count areas.
//   ccl_select_kernel  : roots whose area passes both size filters are appended to a list,
//                        which is radix-sorted (CUB) -> canonical ids in raster order.
//   ccl_assign / scan / scatter : area by id, exclusive offsets, voxels grouped by component.
// Regionprops (m3d_features): one warp per component sorts its voxel list (bitonic, shared
// memory) so that every float32 accumulation runs in raster order exactly like NumPy's
// reductions in scikit-image; per-voxel values are recomputed from the input stack with the
// same device functions the decode kernels use (voxel_math.cuh).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "voxel_math.cuh"

namespace {

constexpr uint32_t DROPPED = 0xFFFFFFFFu;        // failed the minimum-size filter
constexpr uint32_t DROPPED_LARGE = 0xFFFFFFFEu;  // failed the maximum-size filter
constexpr int FEAT_MAX_PX = 2048;  // upper bound on maximum_pixels supported by the warp-sort path

// counters layout inside ctx->s_counters (uint32 each)
enum { CNT_CAND = 0, CNT_FG = 1, CNT_ROOTS = 2 };

__device__ __forceinline__ void warp_append(bool pass, uint32_t value, uint32_t* list, unsigned int* counter) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m) {
        unsigned base = 0;
        const int leader = __ffs(m) - 1;
        if ((int)lane == leader) base = atomicAdd(counter, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (pass) list[base + __popc(m & ((1u << lane) - 1u))] = value;
    }
}

// ------------------------------------------------------------------ collect + init
__global__ void __launch_bounds__(256)
ccl_collect_kernel(const int16_t* __restrict__ decoded, size_t n_vox, int vec, uint32_t* __restrict__ fg,
                   unsigned int* __restrict__ fg_count, uint32_t* __restrict__ parent,
                   uint32_t* __restrict__ aux) {
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    int16_t vals[8];
    const size_t v0 = t * 8;
    if (vec && v0 + 8 <= n_vox) {
        uint4 q = __ldcs(reinterpret_cast<const uint4*>(decoded + v0));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            vals[2 * j] = (int16_t)(w[j] & 0xFFFFu);
            vals[2 * j + 1] = (int16_t)(w[j] >> 16);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) vals[j] = (v0 + j < n_vox) ? decoded[v0 + j] : (int16_t)-1;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool is_fg = vals[j] != -1;
        if (is_fg) {
            parent[v0 + j] = (uint32_t)(v0 + j);
            aux[v0 + j] = 0u;
        }
        warp_append(is_fg, (uint32_t)(v0 + j), fg, fg_count);
    }
}

// ------------------------------------------------------------------ union-find
// find for the merge phase: a start node that sat more than one hop from its root is pointed at the root afterwards
// (atomicMin keeps parents moving towards smaller indices whatever else happens to the node meanwhile).  Sparse
// production foreground never takes the extra atomic (components of a few dozen voxels, chains of one or two hops);
// smooth data decoded with noise-level vectors -- the optimiser's first iteration on low-passed tiles -- forms
// components of millions of voxels whose chains otherwise grow with the component.
__device__ __forceinline__ uint32_t uf_find_shorten(uint32_t* L, uint32_t a0) {
    uint32_t a = a0, p = __ldcg(L + a);
    int hops = 0;
    while (p != a) {
        a = p;
        p = __ldcg(L + a);
        ++hops;
    }
    if (hops > 1) atomicMin(L + a0, a);
    return a;
}

// Playne-Hawick style lock-free union; roots always move towards smaller indices.
__device__ __forceinline__ void uf_union(uint32_t* L, uint32_t a, uint32_t b) {
    a = uf_find_shorten(L, a);
    b = uf_find_shorten(L, b);
    while (a != b) {
        if (a < b) {
            uint32_t t = a;
            a = b;
            b = t;
        }  // a > b
        uint32_t old = atomicMin(L + a, b);
        a = (old == a) ? b : old;
    }
}

// Backward half of the 26-neighbourhood, ordered by how many of the others each one touches.
// If neighbour j has the voxel's value, every other same-valued neighbour adjacent to j is already
// joined to j by its own backward edge, so one union covers them all (ADJ masks below); a blob
// interior voxel needs a single union through (-1,0,0) instead of 13.
__host__ __device__ constexpr int nb_off(int j, int axis) {
    constexpr int t[13][3] = {{-1, 0, 0},  {0, -1, 0},  {0, 0, -1}, {-1, -1, 0}, {-1, 1, 0},
                              {-1, 0, -1}, {-1, 0, 1},  {0, -1, -1}, {0, -1, 1}, {-1, -1, -1},
                              {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1}};
    return t[j][axis];
}
__host__ __device__ constexpr uint32_t nb_adj(int j) {
    uint32_t m = 0;
    for (int k = 0; k < 13; ++k) {
        bool near = true;
        for (int a = 0; a < 3; ++a) {
            const int d = nb_off(j, a) - nb_off(k, a);
            if (d > 1 || d < -1) near = false;
        }
        if (near) m |= 1u << k;
    }
    return m;
}
__constant__ int8_t c_nb_off[13][3] = {{-1, 0, 0},  {0, -1, 0},  {0, 0, -1}, {-1, -1, 0}, {-1, 1, 0},
                                       {-1, 0, -1}, {-1, 0, 1},  {0, -1, -1}, {0, -1, 1}, {-1, -1, -1},
                                       {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1}};
__constant__ uint32_t c_nb_adj[13] = {nb_adj(0), nb_adj(1), nb_adj(2),  nb_adj(3),  nb_adj(4),  nb_adj(5), nb_adj(6),
                                      nb_adj(7), nb_adj(8), nb_adj(9), nb_adj(10), nb_adj(11), nb_adj(12)};
static_assert(nb_adj(0) == 0x1FFFu, "(-1,0,0) touches every backward neighbour");

__global__ void __launch_bounds__(256)
ccl_merge_kernel(const int16_t* __restrict__ decoded, const uint32_t* __restrict__ fg,
                 const unsigned int* __restrict__ fg_count, uint32_t* __restrict__ parent, int Z, int Y, int X,
                 int mode2d) {
    const unsigned n = *fg_count;
    const uint32_t plane = (uint32_t)Y * (uint32_t)X;
    for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const uint32_t v = fg[i];
        const int16_t val = decoded[v];
        const int z = (int)(v / plane);
        const uint32_t rem = v - (uint32_t)z * plane;
        const int y = (int)(rem / (uint32_t)X);
        const int x = (int)(rem - (uint32_t)y * (uint32_t)X);
        // all 13 neighbour loads are issued before any is used
        int16_t nv[13];
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            const int dz = nb_off(j, 0), dy = nb_off(j, 1), dx = nb_off(j, 2);
            const bool inb = (z + dz >= 0) && (y + dy >= 0) && (y + dy < Y) && (x + dx >= 0) && (x + dx < X) &&
                             !(mode2d && dz != 0);
            const long long off = (long long)dz * plane + (long long)dy * X + dx;
            nv[j] = inb ? decoded[(long long)v + off] : (int16_t)-2;
        }
        uint32_t same = 0;
#pragma unroll
        for (int j = 0; j < 13; ++j) same |= (nv[j] == val ? 1u : 0u) << j;
        while (same) {
            const int j = __ffs(same) - 1;
            same &= ~c_nb_adj[j];
            const long long off = (long long)c_nb_off[j][0] * plane + (long long)c_nb_off[j][1] * X + c_nb_off[j][2];
            uf_union(parent, v, (uint32_t)((long long)v + off));
        }
    }
}

// Roots and areas.  The union phase is over, so a voxel may be pointed straight at its root (plain store: every value a
// concurrent find can read is an ancestor).  Areas are counted per warp first: lanes that share a root (neighbours in
// the foreground list usually do, and a component of millions of voxels otherwise sends millions of atomics to ONE
// address) elect their lowest lane, which adds the group's size.
__global__ void __launch_bounds__(256)
ccl_compress_kernel(const uint32_t* __restrict__ fg, const unsigned int* __restrict__ fg_count,
                    uint32_t* __restrict__ parent, uint32_t* __restrict__ aux, uint32_t* __restrict__ root_of) {
    const unsigned n = *fg_count;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned n_round = ((n + 31u) / 32u) * 32u;
    for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < n_round; i += gridDim.x * 256) {
        uint32_t r = 0xFFFFFFFFu;
        if (i < n) {
            const uint32_t v = fg[i];
            uint32_t a = v, p = __ldcg(parent + a);
            int hops = 0;
            while (p != a) {
                a = p;
                p = __ldcg(parent + a);
                ++hops;
            }
            if (hops > 1) parent[v] = a;
            r = a;
            root_of[i] = r;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, r);
        if (i < n && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(aux + r, (unsigned)__popc(peers));
    }
}

__global__ void __launch_bounds__(256)
ccl_select_kernel(const uint32_t* __restrict__ fg, const unsigned int* __restrict__ fg_count,
                  const uint32_t* __restrict__ root_of, uint32_t* __restrict__ aux, uint32_t min_keep,
                  uint32_t max_keep, uint32_t* __restrict__ roots, unsigned int* __restrict__ root_count) {
    const unsigned n = *fg_count;
    const unsigned stride = gridDim.x * 256;
    const unsigned n_round = ((n + 31u) / 32u) * 32u;
    for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < n_round; i += stride) {
        bool keep = false;
        uint32_t v = 0;
        if (i < n) {
            v = fg[i];
            if (root_of[i] == v) {
                const uint32_t a = aux[v];
                keep = (a >= min_keep) && (a <= max_keep);
                if (!keep) aux[v] = (a > max_keep) ? DROPPED_LARGE : DROPPED;
            }
        }
        warp_append(keep, v, roots, root_count);
    }
}

__global__ void __launch_bounds__(256)
ccl_assign_kernel(const uint32_t* __restrict__ roots_sorted, unsigned n_roots, uint32_t* __restrict__ aux,
                  uint32_t* __restrict__ area_by_id) {
    const unsigned id = blockIdx.x * 256 + threadIdx.x;
    if (id < n_roots) {
        const uint32_t r = roots_sorted[id];
        area_by_id[id] = aux[r];
        aux[r] = id;
    }
}

__global__ void __launch_bounds__(256)
ccl_scatter_kernel(const uint32_t* __restrict__ fg, const unsigned int* __restrict__ fg_count,
                   const uint32_t* __restrict__ root_of, const uint32_t* __restrict__ aux,
                   const uint32_t* __restrict__ offs, uint32_t* __restrict__ cursor, uint32_t* __restrict__ vox,
                   int32_t* __restrict__ labels) {
    const unsigned n = *fg_count;
    for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const uint32_t id = aux[root_of[i]];
        if (id == DROPPED_LARGE) {
            // z-slab sharding: a neighbouring slab must learn that this voxel belongs to an
            // oversized component (its merged component is oversized too) -> label -1
            if (labels) labels[fg[i]] = -1;
        } else if (id != DROPPED) {
            const uint32_t v = fg[i];
            const uint32_t slot = atomicAdd(cursor + id, 1u);
            vox[offs[id] + slot] = i;  // position in the foreground list (also indexes the records)
            if (labels) labels[v] = (int32_t)(id + 1u);
        }
    }
}

// ------------------------------------------------------------------ regionprops
// NumPy's float32 pairwise sum (numpy/_core/src/umath/loops_utils.h.src), n <= 2048.
__device__ float np_pairwise_sum(const float* a, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    } else if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return __fadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
    }
}

constexpr int FEAT_WARPS = 4;
constexpr int FEAT_TILE_STRIDE = M3D_MAX_BITS + 1;

static size_t features_smem_bytes(int cap_px) {
    // per warp: sorted (voxel, fg index) keys + magnitude list + one 32-voxel value tile
    return (size_t)FEAT_WARPS * ((size_t)cap_px * 12 + 32 * FEAT_TILE_STRIDE * 4);
}

struct FeatRecords {
    const __half* rec_x;     // [n_fg][NB]
    const uint32_t* rec_md;  // [n_fg]
};

// One warp per component.  Phase A (lane = voxel): every lane produces one voxel's float16-rounded
// scaled values, magnitude and distance to the component's codeword -- read from the records the
// search kernel wrote (FROM_REC) or recomputed with the decode kernels' device functions -- and
// stages the per-bit values in a shared tile.  Phase B (lane = bit): each lane adds its bit's 32
// staged values SEQUENTIALLY in raster order, the order NumPy's axis-0 reduction uses inside
// scikit-image's intensity_mean.  Integer coordinate sums and the distance minimum are order
// independent and warp-reduced.
template <typename T, int NB, bool FROM_REC>
__global__ void __launch_bounds__(FEAT_WARPS * 32)
features_kernel(const T* __restrict__ stack, size_t n_vox, int Y, int X, DecodeParams P,
                const int16_t* __restrict__ decoded, const uint32_t* __restrict__ fg,
                const uint32_t* __restrict__ vox, const uint32_t* __restrict__ offs,
                const uint32_t* __restrict__ area_by_id, unsigned n_feat, int optimize_mode, int cap_px,
                FeatRecords R, double* __restrict__ table, int n_cols) {
    extern __shared__ __align__(16) unsigned char fsm[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    unsigned char* base = fsm + (size_t)warp * ((size_t)cap_px * 12 + 32 * FEAT_TILE_STRIDE * 4);
    unsigned long long* sk = reinterpret_cast<unsigned long long*>(base);
    float* sm = reinterpret_cast<float*>(base + (size_t)cap_px * 8);
    float* tile = reinterpret_cast<float*>(base + (size_t)cap_px * 12);
    const uint32_t plane = (uint32_t)Y * (uint32_t)X;
    for (unsigned id = blockIdx.x * FEAT_WARPS + warp; id < n_feat; id += gridDim.x * FEAT_WARPS) {
        const int n = (int)area_by_id[id];
        const uint32_t* seg = vox + offs[id];
        int Pn = 32;
        while (Pn < n) Pn <<= 1;
        for (int i = lane; i < Pn; i += 32) {
            unsigned long long key = ~0ull;
            if (i < n) {
                const uint32_t fi = seg[i];
                key = ((unsigned long long)fg[fi] << 32) | fi;
            }
            sk[i] = key;
        }
        __syncwarp();
        // bitonic sort, ascending voxel index (raster order)
        for (int k = 2; k <= Pn; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = lane; i < Pn; i += 32) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = sk[i], b = sk[ixj];
                        const bool up = ((i & k) == 0);
                        if ((a > b) == up) {
                            sk[i] = b;
                            sk[ixj] = a;
                        }
                    }
                }
                __syncwarp();
            }
        }
        const uint32_t v_first = (uint32_t)(sk[0] >> 32);
        const int z0 = (int)(v_first / plane);
        const uint32_t rem0 = v_first - (uint32_t)z0 * plane;
        const int y0 = (int)(rem0 / (uint32_t)X);
        const int x0 = (int)(rem0 - (uint32_t)y0 * (uint32_t)X);
        const int dec_id = decoded[v_first];
        const float* crow = P.codebook + (size_t)(dec_id < 0 ? 0 : dec_id) * M3D_MAX_BITS;
        float bit_acc = 0.f;
        float dist_min = __int_as_float(0x7f800000);
        long long sz = 0, sy = 0, sx = 0, szz = 0, syy = 0, sxx = 0, szy = 0, szx = 0, syx = 0;
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int j = c0 + lane;
            if (j < n) {
                const unsigned long long key = sk[j];
                const uint32_t v = (uint32_t)(key >> 32);
                const uint32_t fi = (uint32_t)key;
                float mag16, d16;
                if (FROM_REC) {
                    const uint32_t md = __ldg(R.rec_md + fi);
                    mag16 = __half2float(__ushort_as_half((unsigned short)(md & 0xFFFFu)));
                    d16 = __half2float(__ushort_as_half((unsigned short)(md >> 16)));
                    if (optimize_mode) {
                        float raw[NB];
#pragma unroll
                        for (int b = 0; b < NB; ++b) {
                            const int pb = (b < P.n_bits) ? b : (P.n_bits - 1);
                            raw[b] = load_elem(stack, (size_t)pb * n_vox + v);
                        }
#pragma unroll
                        for (int b = 0; b < NB; ++b) tile[lane * FEAT_TILE_STRIDE + b] = (b < P.n_bits) ? raw[b] : 0.f;
                    } else {
                        const uint4* src = reinterpret_cast<const uint4*>(R.rec_x + (size_t)fi * NB);
#pragma unroll
                        for (int q = 0; q < NB / 8; ++q) {
                            const uint4 t = __ldg(src + q);
                            const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                tile[lane * FEAT_TILE_STRIDE + q * 8 + e] = __half2float(
                                    __ushort_as_half((unsigned short)((w[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu)));
                        }
                    }
                } else {
                    float x[NB], xh[NB], raw[NB];
                    // issue every bit-plane load before the (branchy) IEEE divisions consume them
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const int pb = (b < P.n_bits) ? b : (P.n_bits - 1);
                        raw[b] = load_elem(stack, (size_t)pb * n_vox + v);
                    }
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        x[b] = (b < P.n_bits) ? scale_clip<sizeof(T) == 2>(raw[b], P.bkg[b], P.nrm[b], P.rcp[b]) : 0.f;
                        const float val = optimize_mode ? raw[b] : __half2float(round5_f16(x[b]));
                        tile[lane * FEAT_TILE_STRIDE + b] = (b < P.n_bits) ? val : 0.f;
                    }
                    const float nrm2 = l2_norm<NB>(x);
                    const float mag = unit_vector<NB>(x, nrm2, xh, sizeof(T) == 2 && P.rcp_all);
                    mag16 = __half2float(round5_f16(mag));
                    d16 = __half2float(round5_f16(direct_distance<NB>(xh, crow)));
                }
                dist_min = fminf(dist_min, d16);
                sm[j] = mag16;
                const int z = (int)(v / plane);
                const uint32_t rem = v - (uint32_t)z * plane;
                const int y = (int)(rem / (uint32_t)X);
                const int xx = (int)(rem - (uint32_t)y * (uint32_t)X);
                sz += z; sy += y; sx += xx;
                const long long dz = z - z0, dy = y - y0, dx = xx - x0;
                szz += dz * dz; syy += dy * dy; sxx += dx * dx;
                szy += dz * dy; szx += dz * dx; syx += dy * dx;
            }
            __syncwarp();
            if (lane < P.n_bits) {
                const int cnt = min(32, n - c0);
                for (int jj = 0; jj < cnt; ++jj) {
                    const float val = tile[jj * FEAT_TILE_STRIDE + lane];
                    bit_acc = (c0 == 0 && jj == 0) ? val : __fadd_rn(bit_acc, val);
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dist_min = fminf(dist_min, __shfl_xor_sync(0xffffffffu, dist_min, o));
            sz += __shfl_xor_sync(0xffffffffu, sz, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            szz += __shfl_xor_sync(0xffffffffu, szz, o);
            syy += __shfl_xor_sync(0xffffffffu, syy, o);
            sxx += __shfl_xor_sync(0xffffffffu, sxx, o);
            szy += __shfl_xor_sync(0xffffffffu, szy, o);
            szx += __shfl_xor_sync(0xffffffffu, szx, o);
            syx += __shfl_xor_sync(0xffffffffu, syx, o);
        }
        double* row = table + (size_t)id * n_cols;
        if (lane < P.n_bits) {
            float mean = __fdiv_rn(bit_acc, (float)n);
            if (!optimize_mode) mean = __half2float(__float2half_rn(mean));
            row[M3D_TABLE_FIXED_COLS + lane] = (double)mean;
        }
        if (lane == 0) {
            const double dn = (double)n;
            const float msum = np_pairwise_sum(sm, n);
            const float mmean = __half2float(__float2half_rn(__fdiv_rn(msum, (float)n)));
            row[0] = (double)v_first;
            row[1] = dn;
            row[2] = (double)dec_id;
            row[3] = (double)sz / dn;
            row[4] = (double)sy / dn;
            row[5] = (double)sx / dn;
            // central second moments from first-voxel-relative integer sums (exact in int64)
            const double mz = (double)(sz - (long long)z0 * n), my = (double)(sy - (long long)y0 * n),
                         mx = (double)(sx - (long long)x0 * n);
            row[6] = (double)szz - mz * mz / dn;
            row[7] = (double)syy - my * my / dn;
            row[8] = (double)sxx - mx * mx / dn;
            row[9] = (double)szy - mz * my / dn;
            row[10] = (double)szx - mz * mx / dn;
            row[11] = (double)syx - my * mx / dn;
            row[12] = (double)dist_min;
            row[13] = (double)mmean;
        }
        __syncwarp();
    }
}

template <typename T, int NB, bool FROM_REC>
int launch_features(m3d_ctx* ctx, const T* stack, size_t n_vox, int Y, int X, const int16_t* decoded,
                    int optimize_mode, double* table, cudaStream_t st) {
    const unsigned n_feat = (unsigned)ctx->lab_n_features;
    const DecodeParams P = ctx->params();
    const uint32_t* area_by_id = reinterpret_cast<const uint32_t*>(ctx->s_area.ptr);
    const uint32_t* offs = area_by_id + n_feat;
    const uint32_t* vox = reinterpret_cast<const uint32_t*>(ctx->s_vox.ptr);
    const uint32_t* fg = reinterpret_cast<const uint32_t*>(ctx->s_fg.ptr);
    const FeatRecords R{reinterpret_cast<const __half*>(ctx->s_rec_x.ptr),
                        reinterpret_cast<const uint32_t*>(ctx->s_rec_md.ptr)};
    const int n_cols = M3D_TABLE_FIXED_COLS + ctx->n_bits;
    int cap_px = 32;
    while (cap_px < ctx->lab_max_px) cap_px <<= 1;
    const size_t smem = features_smem_bytes(cap_px);
    auto kern = features_kernel<T, NB, FROM_REC>;
    M3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (int)((n_feat + FEAT_WARPS - 1) / FEAT_WARPS);
    const int cap = ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    M3D_LAUNCH(ctx, KF_FEATURES, st,
               kern<<<blocks, FEAT_WARPS * 32, smem, st>>>(stack, n_vox, Y, X, P, decoded, fg, vox, offs, area_by_id,
                                                           n_feat, optimize_mode, cap_px, R, table, n_cols));
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

template <typename T, int NB>
int launch_features_any(m3d_ctx* ctx, const T* stack, size_t n_vox, int Y, int X, const int16_t* decoded,
                        int optimize_mode, double* table, cudaStream_t st) {
    // records are only valid for the stack / vectors / thresholds of the m3d_decode_label call
    // that wrote them; m3d_label and parameter changes clear the flag
    if (ctx->lab_rec_valid)
        return launch_features<T, NB, true>(ctx, stack, n_vox, Y, X, decoded, optimize_mode, table, st);
    return launch_features<T, NB, false>(ctx, stack, n_vox, Y, X, decoded, optimize_mode, table, st);
}

template <typename T>
int dispatch_features(m3d_ctx* ctx, const T* stack, size_t n_vox, int Y, int X, const int16_t* decoded,
                      int optimize_mode, double* table, cudaStream_t st) {
    switch (ctx->nb_pad) {
        case 8: return launch_features_any<T, 8>(ctx, stack, n_vox, Y, X, decoded, optimize_mode, table, st);
        case 16: return launch_features_any<T, 16>(ctx, stack, n_vox, Y, X, decoded, optimize_mode, table, st);
        case 24: return launch_features_any<T, 24>(ctx, stack, n_vox, Y, X, decoded, optimize_mode, table, st);
        case 32: return launch_features_any<T, 32>(ctx, stack, n_vox, Y, X, decoded, optimize_mode, table, st);
    }
    return m3d_fail(M3D_ERR_ARG, "unsupported padded bit count %d", ctx->nb_pad);
}

// ------------------------------------------------------------------ z-slab interface
// One thread per voxel of the upper slab's FIRST plane: every equal-valued 26-neighbour in the lower
// slab's LAST plane (dz = -1, dy, dx in {-1,0,1}) yields a (label_hi, label_lo) equivalence.  Labels
// are the per-slab canonical ids + 1 (0 = background / dropped small, -1 = oversized component).
// Consecutive duplicates are suppressed per thread; the host de-duplicates the rest.
__global__ void __launch_bounds__(256)
ccl_interface_kernel(const int16_t* __restrict__ dec_lo, const int32_t* __restrict__ lab_lo,
                     const int16_t* __restrict__ dec_hi, const int32_t* __restrict__ lab_hi, int Y, int X,
                     int32_t* __restrict__ pairs, unsigned capacity, unsigned int* __restrict__ n_pairs) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)Y * X) return;
    const int16_t val = dec_hi[i];
    const int32_t lh = lab_hi[i];
    if (val < 0 || lh == 0) return;
    const int y = (int)(i / X), x = (int)(i - (size_t)y * X);
    int32_t last = 0;
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= Y) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= X) continue;
            const size_t j = (size_t)yy * X + xx;
            const int32_t ll = lab_lo[j];
            if (dec_lo[j] == val && ll != 0 && ll != last) {
                last = ll;
                const unsigned slot = atomicAdd(n_pairs, 1u);
                if (slot < capacity) {
                    pairs[2 * (size_t)slot] = lh;
                    pairs[2 * (size_t)slot + 1] = ll;
                }
            }
        }
    }
}

}  // namespace

extern "C" int m3d_interface_pairs(m3d_ctx* ctx, const int16_t* decoded_lo_dev, const int32_t* labels_lo_dev,
                                   const int16_t* decoded_hi_dev, const int32_t* labels_hi_dev, int64_t Y, int64_t X,
                                   int32_t* pairs_dev, int64_t capacity, int64_t* n_pairs_out, void* stream) {
    if (!ctx || !decoded_lo_dev || !labels_lo_dev || !decoded_hi_dev || !labels_hi_dev || !pairs_dev || !n_pairs_out)
        return m3d_fail(M3D_ERR_ARG, "m3d_interface_pairs: null argument");
    if (Y <= 0 || X <= 0 || capacity <= 0 || Y * X >= 0x7fffffffll)
        return m3d_fail(M3D_ERR_ARG, "m3d_interface_pairs: bad dims");
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (ctx->s_counters.ensure(256)) return M3D_ERR_CUDA;
    unsigned int* counter = reinterpret_cast<unsigned int*>(ctx->s_counters.ptr) + 8;
    M3D_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    const int blocks = (int)(((size_t)Y * X + 255) / 256);
    M3D_LAUNCH(ctx, KF_CCL_INTERFACE, st,
               ccl_interface_kernel<<<blocks, 256, 0, st>>>(decoded_lo_dev, labels_lo_dev, decoded_hi_dev, labels_hi_dev,
                                                            (int)Y, (int)X, pairs_dev, (unsigned)capacity, counter));
    M3D_CHECK_LAUNCH();
    unsigned int* h = reinterpret_cast<unsigned int*>(ctx->h_pinned) + 8;
    M3D_CUDA(cudaMemcpyAsync(h, counter, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    M3D_CUDA(cudaStreamSynchronize(st));
    *n_pairs_out = (int64_t)*h;  // may exceed capacity: the caller retries with a larger buffer
    return M3D_OK;
}

// ====================================================================== host entry points
// defined in decode.cu
int m3d_check_decode_args(m3d_ctx* ctx, const void* stack_dev, int dtype, const int64_t dims[3],
                          int16_t* decoded_dev, size_t* n_vox);
int m3d_decode_internal(m3d_ctx* ctx, const void* stack_dev, int dtype, size_t n_vox, int16_t* decoded_dev,
                        uint32_t* fg, unsigned int* fg_count, uint32_t* parent, uint32_t* aux, void* rec_x,
                        uint32_t* rec_md, unsigned rec_cap, cudaStream_t st);

namespace {

// persistent decoded image: put the previous tile's foreground voxels back to background (-1)
__global__ void __launch_bounds__(256)
reset_foreground_kernel(const uint32_t* __restrict__ fg, unsigned n, int16_t* __restrict__ decoded) {
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) decoded[fg[i]] = (int16_t)-1;
}

struct LabelScratch {
    unsigned int* counters;
    uint32_t *fg, *root_of, *parent, *aux;
};

int label_prepare(m3d_ctx* ctx, const int64_t dims[3], int maximum_pixels, int32_t* labels_dev, size_t* n_vox_out,
                  LabelScratch* L, cudaStream_t st) {
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return m3d_fail(M3D_ERR_ARG, "m3d_label: bad dims");
    const size_t n_vox = (size_t)dims[0] * dims[1] * dims[2];
    if (n_vox >= 0xFFFFFFF0ull) return m3d_fail(M3D_ERR_ARG, "m3d_label: volume exceeds 2^32 voxels per call");
    if (maximum_pixels < 1 || maximum_pixels > FEAT_MAX_PX)
        return m3d_fail(M3D_ERR_ARG, "m3d_label: maximum_pixels must be in [1, %d]", FEAT_MAX_PX);
    ctx->lab_n_features = -1;
    ctx->lab_rec_valid = 0;
    if (ctx->s_counters.ensure(256)) return M3D_ERR_CUDA;
    if (ctx->s_fg.ensure(2 * n_vox * sizeof(uint32_t))) return M3D_ERR_CUDA;  // fg list + root_of
    if (ctx->s_parent.ensure(n_vox * sizeof(uint32_t))) return M3D_ERR_CUDA;
    if (ctx->s_aux.ensure(n_vox * sizeof(uint32_t))) return M3D_ERR_CUDA;
    L->counters = reinterpret_cast<unsigned int*>(ctx->s_counters.ptr);
    L->fg = reinterpret_cast<uint32_t*>(ctx->s_fg.ptr);
    L->root_of = L->fg + n_vox;
    L->parent = reinterpret_cast<uint32_t*>(ctx->s_parent.ptr);
    L->aux = reinterpret_cast<uint32_t*>(ctx->s_aux.ptr);
    M3D_CUDA(cudaMemsetAsync(L->counters + CNT_FG, 0, 2 * sizeof(unsigned int), st));
    if (labels_dev) M3D_CUDA(cudaMemsetAsync(labels_dev, 0, n_vox * sizeof(int32_t), st));
    *n_vox_out = n_vox;
    return M3D_OK;
}

// union-find over the foreground list -> size filters -> canonical ids -> voxels grouped by id
int label_finish(m3d_ctx* ctx, const int16_t* decoded_dev, const int64_t dims[3], size_t n_vox, int mode2d,
                 double minimum_pixels, int maximum_pixels, int32_t* labels_dev, int64_t* n_features_out,
                 const LabelScratch& L, cudaStream_t st) {
    const int Z = (int)dims[0], Y = (int)dims[1], X = (int)dims[2];
    unsigned int* counters = L.counters;
    const int sparse_blocks = ctx->num_sms * 8;
    M3D_LAUNCH(ctx, KF_CCL_MERGE, st,
               ccl_merge_kernel<<<sparse_blocks, 256, 0, st>>>(decoded_dev, L.fg, counters + CNT_FG, L.parent, Z, Y, X,
                                                               mode2d ? 1 : 0));
    M3D_CHECK_LAUNCH();
    M3D_LAUNCH(ctx, KF_CCL_COMPRESS, st,
               ccl_compress_kernel<<<sparse_blocks, 256, 0, st>>>(L.fg, counters + CNT_FG, L.parent, L.aux, L.root_of));
    M3D_CHECK_LAUNCH();

    // PD:2976-2989: drop area > maximum_pixels; drop area <= max(int(minimum_pixels)-1, 0)
    long long max_size = (long long)minimum_pixels - 1;  // (long long) truncates toward zero like Python's int()
    if (max_size < 0) max_size = 0;
    const uint32_t min_keep = (uint32_t)(max_size + 1);
    if (ctx->s_roots.ensure((n_vox / min_keep + 64) * 2 * sizeof(uint32_t))) return M3D_ERR_CUDA;
    uint32_t* roots = reinterpret_cast<uint32_t*>(ctx->s_roots.ptr);
    const size_t roots_cap = ctx->s_roots.cap / (2 * sizeof(uint32_t));
    uint32_t* roots_sorted = roots + roots_cap;
    M3D_LAUNCH(ctx, KF_CCL_SELECT, st,
               ccl_select_kernel<<<sparse_blocks, 256, 0, st>>>(L.fg, counters + CNT_FG, L.root_of, L.aux, min_keep,
                                                                (uint32_t)maximum_pixels, roots, counters + CNT_ROOTS));
    M3D_CHECK_LAUNCH();

    unsigned int* h_counts = reinterpret_cast<unsigned int*>(ctx->h_pinned);
    M3D_CUDA(cudaMemcpyAsync(h_counts, counters + CNT_FG, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    M3D_CUDA(cudaStreamSynchronize(st));
    const unsigned n_fg = h_counts[0], n_roots = h_counts[1];
    ctx->lab_n_fg = n_fg;
    ctx->lab_dims[0] = Z; ctx->lab_dims[1] = Y; ctx->lab_dims[2] = X;
    ctx->lab_max_px = maximum_pixels;
    if (n_roots == 0) {
        ctx->lab_n_features = 0;
        *n_features_out = 0;
        return M3D_OK;
    }
    // canonical order = ascending first-voxel index
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, roots, roots_sorted, (int)n_roots, 0, 32, st);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_roots, st);
    if (scan_bytes > tmp_bytes) tmp_bytes = scan_bytes;
    if (ctx->s_sort.ensure(tmp_bytes)) return M3D_ERR_CUDA;
    {
        KernelScope ks(ctx, KF_CCL_SORT, st);
        M3D_CUDA(cub::DeviceRadixSort::SortKeys(ctx->s_sort.ptr, tmp_bytes, roots, roots_sorted, (int)n_roots, 0, 32, st));
    }
    // area_by_id | offs | cursor
    if (ctx->s_area.ensure((size_t)n_roots * 3 * sizeof(uint32_t))) return M3D_ERR_CUDA;
    uint32_t* area_by_id = reinterpret_cast<uint32_t*>(ctx->s_area.ptr);
    uint32_t* offs = area_by_id + n_roots;
    uint32_t* cursor = offs + n_roots;
    M3D_LAUNCH(ctx, KF_CCL_ASSIGN, st,
               ccl_assign_kernel<<<(n_roots + 255) / 256, 256, 0, st>>>(roots_sorted, n_roots, L.aux, area_by_id));
    M3D_CHECK_LAUNCH();
    {
        KernelScope ks(ctx, KF_CCL_SCAN, st);
        M3D_CUDA(cub::DeviceScan::ExclusiveSum(ctx->s_sort.ptr, tmp_bytes, area_by_id, offs, (int)n_roots, st));
    }
    M3D_CUDA(cudaMemsetAsync(cursor, 0, (size_t)n_roots * sizeof(uint32_t), st));
    if (ctx->s_vox.ensure((size_t)n_fg * sizeof(uint32_t))) return M3D_ERR_CUDA;
    uint32_t* vox = reinterpret_cast<uint32_t*>(ctx->s_vox.ptr);
    M3D_LAUNCH(ctx, KF_CCL_SCATTER, st,
               ccl_scatter_kernel<<<sparse_blocks, 256, 0, st>>>(L.fg, counters + CNT_FG, L.root_of, L.aux, offs, cursor,
                                                                 vox, labels_dev));
    M3D_CHECK_LAUNCH();
    ctx->lab_n_features = n_roots;
    *n_features_out = n_roots;
    return M3D_OK;
}

}  // namespace

extern "C" int m3d_label(m3d_ctx* ctx, const int16_t* decoded_dev, const int64_t dims[3], int mode2d,
                         double minimum_pixels, int maximum_pixels, int32_t* labels_dev,
                         int64_t* n_features_out, void* stream) {
    if (!ctx || !decoded_dev || !dims || !n_features_out) return m3d_fail(M3D_ERR_ARG, "m3d_label: null argument");
    ctx->prev_fg_valid = 0;  // the foreground list is rebuilt from another image
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    size_t n_vox = 0;
    LabelScratch L;
    int rc = label_prepare(ctx, dims, maximum_pixels, labels_dev, &n_vox, &L, st);
    if (rc) return rc;
    {
        const int vec = ((reinterpret_cast<uintptr_t>(decoded_dev) & 15u) == 0) ? 1 : 0;
        const size_t threads = (n_vox + 7) / 8;
        const int blocks = (int)((threads + 255) / 256);
        M3D_LAUNCH(ctx, KF_CCL_COLLECT, st,
                   ccl_collect_kernel<<<blocks, 256, 0, st>>>(decoded_dev, n_vox, vec, L.fg, L.counters + CNT_FG, L.parent,
                                                              L.aux));
        M3D_CHECK_LAUNCH();
    }
    return label_finish(ctx, decoded_dev, dims, n_vox, mode2d, minimum_pixels, maximum_pixels, labels_dev,
                        n_features_out, L, st);
}

static int decode_label_impl(m3d_ctx* ctx, const void* stack_dev, int dtype, const int64_t dims[3],
                             int16_t* decoded_dev, int mode2d, double minimum_pixels, int maximum_pixels,
                             int32_t* labels_dev, int64_t* n_features_out, int persistent, void* stream) {
    if (!n_features_out) return m3d_fail(M3D_ERR_ARG, "m3d_decode_label: null argument");
    size_t n_vox = 0;
    int rc = m3d_check_decode_args(ctx, stack_dev, dtype, dims, decoded_dev, &n_vox);
    if (rc) return rc;
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // the caller vouches that decoded_dev still holds exactly what the previous persistent call left there
    const bool reuse = persistent && ctx->prev_fg_valid && ctx->prev_decoded == (const void*)decoded_dev &&
                       ctx->prev_n_vox == n_vox && ctx->s_fg.ptr != nullptr;
    const unsigned prev_fg = reuse ? (unsigned)ctx->lab_n_fg : 0u;
    ctx->prev_fg_valid = 0;
    LabelScratch L;
    rc = label_prepare(ctx, dims, maximum_pixels, labels_dev, &n_vox, &L, st);
    if (rc) return rc;
    if (reuse && prev_fg) {
        const int blocks = (int)std::min<size_t>((prev_fg + 255u) / 256u, (size_t)ctx->num_sms * 8);
        M3D_LAUNCH(ctx, KF_RESET_FG, st, reset_foreground_kernel<<<blocks, 256, 0, st>>>(L.fg, prev_fg, decoded_dev));
        M3D_CHECK_LAUNCH();
    }
    // the search kernel emits the foreground list and initialises the union-find slots: no
    // second pass over the decoded image
    const size_t rec_cap = m3d_sparse_capacity(ctx, n_vox);
    if (ctx->s_rec_x.ensure(rec_cap * (size_t)ctx->nb_pad * 2)) return M3D_ERR_CUDA;
    if (ctx->s_rec_md.ensure(rec_cap * sizeof(uint32_t))) return M3D_ERR_CUDA;
    ctx->gate_skip_background = reuse ? 1 : 0;
    rc = m3d_decode_internal(ctx, stack_dev, dtype, n_vox, decoded_dev, L.fg, L.counters + CNT_FG, L.parent, L.aux,
                             ctx->s_rec_x.ptr, reinterpret_cast<uint32_t*>(ctx->s_rec_md.ptr), (unsigned)rec_cap, st);
    ctx->gate_skip_background = 0;
    if (rc) return rc;
    rc = label_finish(ctx, decoded_dev, dims, n_vox, mode2d, minimum_pixels, maximum_pixels, labels_dev,
                      n_features_out, L, st);
    if (rc) return rc;
    // the regionprops stage may read the search kernel's records instead of recomputing traces
    ctx->lab_rec_valid = ((size_t)ctx->lab_n_fg <= rec_cap) ? 1 : 0;
    if (persistent) {
        ctx->prev_decoded = decoded_dev;
        ctx->prev_n_vox = n_vox;
        ctx->prev_fg_valid = 1;
    }
    return M3D_OK;
}

extern "C" int m3d_decode_label(m3d_ctx* ctx, const void* stack_dev, int dtype, const int64_t dims[3],
                                int16_t* decoded_dev, int mode2d, double minimum_pixels, int maximum_pixels,
                                int32_t* labels_dev, int64_t* n_features_out, void* stream) {
    return decode_label_impl(ctx, stack_dev, dtype, dims, decoded_dev, mode2d, minimum_pixels, maximum_pixels, labels_dev,
                             n_features_out, 0, stream);
}

extern "C" int m3d_decode_label_persistent(m3d_ctx* ctx, const void* stack_dev, int dtype, const int64_t dims[3],
                                           int16_t* decoded_dev, int mode2d, double minimum_pixels, int maximum_pixels,
                                           int32_t* labels_dev, int64_t* n_features_out, void* stream) {
    return decode_label_impl(ctx, stack_dev, dtype, dims, decoded_dev, mode2d, minimum_pixels, maximum_pixels, labels_dev,
                             n_features_out, 1, stream);
}

extern "C" int m3d_features(m3d_ctx* ctx, const void* stack_dev, int dtype, const int64_t dims[3],
                            const int16_t* decoded_dev, int optimize_mode, double* table_dev, int64_t n_rows,
                            void* stream) {
    if (!ctx || !stack_dev || !dims || !decoded_dev) return m3d_fail(M3D_ERR_ARG, "m3d_features: null argument");
    if (ctx->lab_n_features < 0) return m3d_fail(M3D_ERR_STATE, "m3d_features: call m3d_label first");
    if (dims[0] != ctx->lab_dims[0] || dims[1] != ctx->lab_dims[1] || dims[2] != ctx->lab_dims[2])
        return m3d_fail(M3D_ERR_STATE, "m3d_features: dims differ from the last m3d_label call");
    if (n_rows < ctx->lab_n_features)
        return m3d_fail(M3D_ERR_CAPACITY, "m3d_features: table holds %lld rows, need %lld", (long long)n_rows,
                        (long long)ctx->lab_n_features);
    if (ctx->lab_n_features == 0) return M3D_OK;
    if (!table_dev) return m3d_fail(M3D_ERR_ARG, "m3d_features: null table");
    if (dtype != M3D_DTYPE_U16 && dtype != M3D_DTYPE_F32) return m3d_fail(M3D_ERR_ARG, "m3d_features: dtype %d", dtype);
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t n_vox = (size_t)dims[0] * dims[1] * dims[2];
    if (dtype == M3D_DTYPE_U16)
        return dispatch_features<uint16_t>(ctx, reinterpret_cast<const uint16_t*>(stack_dev), n_vox, (int)dims[1],
                                           (int)dims[2], decoded_dev, optimize_mode, table_dev, st);
    return dispatch_features<float>(ctx, reinterpret_cast<const float*>(stack_dev), n_vox, (int)dims[1], (int)dims[2],
                                    decoded_dev, optimize_mode, table_dev, st);
}
