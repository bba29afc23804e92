// table.cu -- post-decode transcript-table stage on the device.
//
// The rows are decoded transcripts (1e5 .. 1e7 per experiment), not voxels; all three kernels are
// latency / atomics bound and tiny next to the decode, but they replace the reference's per-pair
// Python loops (PD:4137-4177, PD:4179-4363) and its host histogramming (PD:3656-3742):
//
//   table_hist3d_kernel   : float32 searchsorted(side='right') - 1 on three feature axes, flat bin
//                           per row, all / blank histograms (shared-memory privatised atomics).
//   grid keys + CUB sort  : uniform grid hash (cell = search radius) over float64 coordinates.
//   overlap_kernel        : PD:4137 -- a row is dropped iff some row of ANOTHER tile within the 3-D
//                           radius has the smaller (distance_min, row index).  The reference's loop
//                           over cKDTree pairs reduces to exactly this per-row predicate, so no
//                           pair list and no atomics are needed.
//   within_union_kernel   : PD:4179 -- same tile, same gene, XY distance <= r_xy, 0 < |dz| <= r_z:
//                           lock-free union-find (root = smallest row index), then per-cluster
//                           lexicographic arg-min of (distance_min, row index) in two atomicMin
//                           passes; every other member is dropped.
// Squared distances use the operation order of SciPy's cKDTree (sum of squares, one rounding per
// multiply and add, compared with r*r) so inclusive-radius ties fall the same way.
#include <algorithm>

#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int TB = 256;
constexpr int MAX_EDGES = 64;

struct Edges3 {
    float e[3][MAX_EDGES];
    int n[3];
};

__device__ __forceinline__ int bin_right(const float* e, int n, float v) {
    // numpy.searchsorted(e, v, side='right') - 1; NaN sorts after everything
    if (v != v) return n - 1;
    int c = 0;
    for (int i = 0; i < n; ++i) c += (e[i] <= v) ? 1 : 0;
    return c - 1;
}

__global__ void __launch_bounds__(TB)
table_hist3d_kernel(const float* __restrict__ v0, const float* __restrict__ v1, const float* __restrict__ v2,
                    const uint8_t* __restrict__ blank, size_t n, Edges3 E, int n_bins, int32_t* __restrict__ flat_bin,
                    int32_t* __restrict__ all_hist, int32_t* __restrict__ blank_hist) {
    extern __shared__ int s_hist[];  // [2][n_bins] when it fits, else unused
    const bool priv = n_bins * 2 * sizeof(int) <= 40 * 1024;
    if (priv) {
        for (int i = threadIdx.x; i < 2 * n_bins; i += TB) s_hist[i] = 0;
        __syncthreads();
    }
    const int d1 = E.n[1] - 1, d2 = E.n[2] - 1;
    for (size_t i = (size_t)blockIdx.x * TB + threadIdx.x; i < n; i += (size_t)gridDim.x * TB) {
        const float a = v0[i], b = v1[i], c = v2[i];
        const bool finite = (fabsf(a) <= 3.4028235e38f) && (fabsf(b) <= 3.4028235e38f) && (fabsf(c) <= 3.4028235e38f);
        const int b0 = bin_right(E.e[0], E.n[0], a);
        const int b1 = bin_right(E.e[1], E.n[1], b);
        const int b2 = bin_right(E.e[2], E.n[2], c);
        const bool in = finite && b0 >= 0 && b0 < E.n[0] - 1 && b1 >= 0 && b1 < d1 && b2 >= 0 && b2 < d2;
        const int flat = in ? (b0 * d1 + b1) * d2 + b2 : -1;
        flat_bin[i] = flat;
        if (in) {
            if (priv) {
                atomicAdd(&s_hist[flat], 1);
                if (blank[i]) atomicAdd(&s_hist[n_bins + flat], 1);
            } else {
                atomicAdd(&all_hist[flat], 1);
                if (blank[i]) atomicAdd(&blank_hist[flat], 1);
            }
        }
    }
    if (priv) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_bins; i += TB) {
            if (s_hist[i]) atomicAdd(&all_hist[i], s_hist[i]);
            if (s_hist[n_bins + i]) atomicAdd(&blank_hist[i], s_hist[n_bins + i]);
        }
    }
}

// ------------------------------------------------------------------ bounds + grid keys
__device__ __forceinline__ unsigned long long ordered_u64(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double unordered_f64(unsigned long long k) {
    unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    double d;
    memcpy(&d, &u, sizeof(d));
    return d;
}

// mm[0..2] = min z,y,x   mm[3..5] = max z,y,x (ordered keys)   mm[6] = count of non-finite coordinates
__global__ void __launch_bounds__(TB)
table_bounds_kernel(const double* __restrict__ zyx, size_t n, unsigned long long* __restrict__ mm) {
    __shared__ unsigned long long s[7];
    if (threadIdx.x < 3) s[threadIdx.x] = ~0ull;
    else if (threadIdx.x < 7) s[threadIdx.x] = 0ull;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * TB + threadIdx.x; i < n; i += (size_t)gridDim.x * TB) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double v = zyx[i * 3 + a];
            if (!(fabs(v) <= 1.0e300)) {
                atomicAdd(&s[6], 1ull);
            } else {
                const unsigned long long k = ordered_u64(v);
                atomicMin(&s[a], k);
                atomicMax(&s[3 + a], k);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(&mm[threadIdx.x], s[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(&mm[threadIdx.x], s[threadIdx.x]);
    else if (threadIdx.x == 6 && s[6]) atomicAdd(&mm[6], s[6]);
}

struct GridSpec {
    double origin[3];
    double inv_cell[3];
    int use_z;      // 1: 3-D cells (overlap); 0: (tile, y, x) cells (within tile)
};

__device__ __forceinline__ long long cell_of(double v, double origin, double inv) {
    return (long long)floor((v - origin) * inv);
}

// overlap: key = cz:21 | cy:21 | cx:21.   within: key = tile:20 | cy:22 | cx:22.
__global__ void __launch_bounds__(TB)
table_keys_kernel(const double* __restrict__ zyx, const int32_t* __restrict__ tile, size_t n, GridSpec G,
                  unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
    const size_t i = (size_t)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    const unsigned long long cy = (unsigned long long)cell_of(zyx[i * 3 + 1], G.origin[1], G.inv_cell[1]);
    const unsigned long long cx = (unsigned long long)cell_of(zyx[i * 3 + 2], G.origin[2], G.inv_cell[2]);
    unsigned long long k;
    if (G.use_z) {
        const unsigned long long cz = (unsigned long long)cell_of(zyx[i * 3 + 0], G.origin[0], G.inv_cell[0]);
        k = (cz << 42) | (cy << 21) | cx;
    } else {
        k = ((unsigned long long)(uint32_t)tile[i] << 44) | (cy << 22) | cx;
    }
    keys[i] = k;
    idx[i] = (uint32_t)i;
}

__device__ __forceinline__ size_t lower_bound_u64(const unsigned long long* __restrict__ a, size_t n,
                                                  unsigned long long key) {
    size_t lo = 0, hi = n;
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------ PD:4137 tile-overlap duplicates
__global__ void __launch_bounds__(TB)
overlap_kernel(const double* __restrict__ zyx, const int32_t* __restrict__ tile, const double* __restrict__ dmin,
               size_t n, GridSpec G, double r2, const unsigned long long* __restrict__ keys,
               const uint32_t* __restrict__ idx, uint8_t* __restrict__ drop) {
    const size_t s = (size_t)blockIdx.x * TB + threadIdx.x;
    if (s >= n) return;
    const uint32_t i = idx[s];
    const double z = zyx[(size_t)i * 3], y = zyx[(size_t)i * 3 + 1], x = zyx[(size_t)i * 3 + 2];
    const long long cz = cell_of(z, G.origin[0], G.inv_cell[0]);
    const long long cy = cell_of(y, G.origin[1], G.inv_cell[1]);
    const long long cx = cell_of(x, G.origin[2], G.inv_cell[2]);
    const int ti = tile[i];
    const double di = dmin[i];
    bool lose = false;
    for (long long dz = -1; dz <= 1 && !lose; ++dz) {
        for (long long dy = -1; dy <= 1 && !lose; ++dy) {
            const long long nz = cz + dz, ny = cy + dy;
            if (nz < 0 || ny < 0) continue;
            const long long x0 = cx > 0 ? cx - 1 : 0;
            const unsigned long long k0 = ((unsigned long long)nz << 42) | ((unsigned long long)ny << 21) | (unsigned long long)x0;
            const unsigned long long k1 = ((unsigned long long)nz << 42) | ((unsigned long long)ny << 21) | (unsigned long long)(cx + 1);
            for (size_t p = lower_bound_u64(keys, n, k0); p < n && keys[p] <= k1; ++p) {
                const uint32_t j = idx[p];
                if (j == i || tile[j] == ti) continue;
                const double a = z - zyx[(size_t)j * 3], b = y - zyx[(size_t)j * 3 + 1], c = x - zyx[(size_t)j * 3 + 2];
                double d2 = __dmul_rn(a, a);
                d2 = __dadd_rn(d2, __dmul_rn(b, b));
                d2 = __dadd_rn(d2, __dmul_rn(c, c));
                if (d2 <= r2) {
                    const double dj = dmin[j];
                    if (dj < di || (dj == di && j < i)) {
                        lose = true;
                        break;
                    }
                }
            }
        }
    }
    drop[i] = lose ? 1 : 0;
}

// ------------------------------------------------------------------ PD:4179 within-tile clusters
__device__ __forceinline__ uint32_t uf_find(uint32_t* __restrict__ parent, uint32_t a) {
    // clusters are tiny (the same molecule seen in 2-3 planes): plain pointer chasing, no compression.
    // Parents only ever decrease, so a stale read just costs one more step.
    volatile uint32_t* vp = parent;
    uint32_t p = vp[a];
    while (p != a) {
        a = p;
        p = vp[a];
    }
    return a;
}

__device__ __forceinline__ void uf_union(uint32_t* __restrict__ parent, uint32_t a, uint32_t b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) {
            const uint32_t t = a;
            a = b;
            b = t;
        }  // a > b: hang a under b
        const uint32_t old = atomicCAS(&parent[a], a, b);
        if (old == a) return;
    }
}

__global__ void __launch_bounds__(TB)
iota_kernel(uint32_t* __restrict__ p, size_t n) {
    const size_t i = (size_t)blockIdx.x * TB + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(TB)
within_union_kernel(const double* __restrict__ zyx, const int32_t* __restrict__ tile, const int32_t* __restrict__ gene,
                    size_t n, GridSpec G, double rxy2, double rz, const unsigned long long* __restrict__ keys,
                    const uint32_t* __restrict__ idx, uint32_t* __restrict__ parent) {
    const size_t s = (size_t)blockIdx.x * TB + threadIdx.x;
    if (s >= n) return;
    const uint32_t i = idx[s];
    const double z = zyx[(size_t)i * 3], y = zyx[(size_t)i * 3 + 1], x = zyx[(size_t)i * 3 + 2];
    const long long cy = cell_of(y, G.origin[1], G.inv_cell[1]);
    const long long cx = cell_of(x, G.origin[2], G.inv_cell[2]);
    const unsigned long long tk = (unsigned long long)(uint32_t)tile[i] << 44;
    const int gi = gene[i];
    for (long long dy = -1; dy <= 1; ++dy) {
        const long long ny = cy + dy;
        if (ny < 0) continue;
        const long long x0 = cx > 0 ? cx - 1 : 0;
        const unsigned long long k0 = tk | ((unsigned long long)ny << 22) | (unsigned long long)x0;
        const unsigned long long k1 = tk | ((unsigned long long)ny << 22) | (unsigned long long)(cx + 1);
        for (size_t p = lower_bound_u64(keys, n, k0); p < n && keys[p] <= k1; ++p) {
            const uint32_t j = idx[p];
            if (j <= i || gene[j] != gi) continue;  // each unordered pair once
            const double b = y - zyx[(size_t)j * 3 + 1], c = x - zyx[(size_t)j * 3 + 2];
            const double d2 = __dadd_rn(__dmul_rn(b, b), __dmul_rn(c, c));
            if (!(d2 <= rxy2)) continue;
            const double az = fabs(z - zyx[(size_t)j * 3]);
            if (az > 0.0 && az <= rz) uf_union(parent, i, j);
        }
    }
}

__global__ void __launch_bounds__(TB)
within_best_d_kernel(const double* __restrict__ dmin, size_t n, uint32_t* __restrict__ parent,
                     unsigned long long* __restrict__ best_d, uint32_t* __restrict__ best_i) {
    const size_t i = (size_t)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = uf_find(parent, (uint32_t)i);
    parent[i] = r;
    atomicMin(&best_d[r], ordered_u64(dmin[i]));
}

__global__ void __launch_bounds__(TB)
within_best_i_kernel(const double* __restrict__ dmin, size_t n, const uint32_t* __restrict__ parent,
                     const unsigned long long* __restrict__ best_d, uint32_t* __restrict__ best_i) {
    const size_t i = (size_t)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = parent[i];
    if (ordered_u64(dmin[i]) == best_d[r]) atomicMin(&best_i[r], (uint32_t)i);
}

__global__ void __launch_bounds__(TB)
within_drop_kernel(size_t n, const uint32_t* __restrict__ parent, const uint32_t* __restrict__ best_i,
                   uint8_t* __restrict__ drop) {
    const size_t i = (size_t)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    drop[i] = best_i[parent[i]] != (uint32_t)i ? 1 : 0;
}

// ------------------------------------------------------------------ PD:4076-4135 cell assignment
// cell_id = 1 + index of the first polygon (lowest index) that contains the point, 0 when none does.
// Candidates come from a uniform grid over the polygon bounding boxes (the reference uses an R-tree on the same
// boxes); containment is the even-odd crossing rule in float64 = shapely's `contains` for every point that is
// not exactly on a polygon boundary.
__global__ void __launch_bounds__(TB)
assign_cells_kernel(const double* __restrict__ yx, size_t n, const double* __restrict__ verts,
                    const long long* __restrict__ offs, const double* __restrict__ bbox,
                    const int32_t* __restrict__ cell_start, const int32_t* __restrict__ cell_polys, double oy, double ox,
                    double inv_cell, int gy, int gx, int32_t* __restrict__ cell_id) {
    const size_t i = (size_t)blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    const double py = yx[2 * i], px = yx[2 * i + 1];
    int out = 0;
    const double fy = floor((py - oy) * inv_cell), fx = floor((px - ox) * inv_cell);
    if (fy >= 0.0 && fx >= 0.0 && fy < (double)gy && fx < (double)gx) {
        const int c = (int)fy * gx + (int)fx;
        for (int e = cell_start[c]; e < cell_start[c + 1] && out == 0; ++e) {
            const int p = cell_polys[e];
            const double* bb = bbox + 4 * (size_t)p;
            if (py < bb[0] || px < bb[1] || py > bb[2] || px > bb[3]) continue;
            const long long a = offs[p], b = offs[p + 1];
            bool inside = false;
            long long j = b - 1;
            for (long long k = a; k < b; j = k++) {
                const double yi = verts[2 * k], xi = verts[2 * k + 1];
                const double yj = verts[2 * j], xj = verts[2 * j + 1];
                if ((yi > py) != (yj > py)) {
                    const double xc = __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(xj, xi), __dsub_rn(py, yi)), __dsub_rn(yj, yi)), xi);
                    if (px < xc) inside = !inside;
                }
            }
            if (inside) out = p + 1;
        }
    }
    cell_id[i] = out;
}

// shared set-up: bounds -> grid spec -> sorted (key, row) lists in ctx scratch
int build_grid(m3d_ctx* ctx, const double* zyx, const int32_t* tile, size_t n, double cell_zyx[3], int use_z,
               GridSpec* G_out, unsigned long long** keys_out, uint32_t** idx_out, cudaStream_t st, const char* who) {
    if (ctx->s_counters.ensure(256)) return M3D_ERR_CUDA;
    unsigned long long* mm = reinterpret_cast<unsigned long long*>(ctx->s_counters.ptr) + 8;
    unsigned long long init[7] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull, 0ull};
    M3D_CUDA(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const int blocks = (int)std::min<size_t>((n + TB - 1) / TB, (size_t)ctx->num_sms * 8);
    M3D_LAUNCH(ctx, KF_TABLE_GRID, st, table_bounds_kernel<<<blocks, TB, 0, st>>>(zyx, n, mm));
    M3D_CHECK_LAUNCH();
    unsigned long long h[7];
    M3D_CUDA(cudaMemcpyAsync(h, mm, sizeof(h), cudaMemcpyDeviceToHost, st));
    M3D_CUDA(cudaStreamSynchronize(st));
    if (h[6]) return m3d_fail(M3D_ERR_ARG, "%s: %llu non-finite coordinates", who, h[6]);
    GridSpec G;
    memset(&G, 0, sizeof(G));
    G.use_z = use_z;
    const int bits = use_z ? 21 : 22;
    for (int a = 0; a < 3; ++a) {
        const double lo = unordered_f64(h[a]), hi = unordered_f64(h[3 + a]);
        G.origin[a] = lo;
        G.inv_cell[a] = 1.0 / cell_zyx[a];
        if (a == 0 && !use_z) continue;
        if ((hi - lo) * G.inv_cell[a] >= (double)((1ll << bits) - 2))
            return m3d_fail(M3D_ERR_ARG, "%s: coordinate extent %.3g on axis %d exceeds %d-bit cells of %.3g", who,
                            hi - lo, a, bits, cell_zyx[a]);
    }
    const size_t key_bytes = ((n * sizeof(unsigned long long) + 255) / 256) * 256;
    const size_t idx_bytes = ((n * sizeof(uint32_t) + 255) / 256) * 256;
    if (ctx->s_vox.ensure(2 * key_bytes + 2 * idx_bytes)) return M3D_ERR_CUDA;
    char* base = reinterpret_cast<char*>(ctx->s_vox.ptr);
    unsigned long long* keys_in = reinterpret_cast<unsigned long long*>(base);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(base + key_bytes);
    uint32_t* idx_in = reinterpret_cast<uint32_t*>(base + 2 * key_bytes);
    uint32_t* idx = reinterpret_cast<uint32_t*>(base + 2 * key_bytes + idx_bytes);
    const int nb = (int)((n + TB - 1) / TB);
    M3D_LAUNCH(ctx, KF_TABLE_GRID, st, table_keys_kernel<<<nb, TB, 0, st>>>(zyx, tile, n, G, keys_in, idx_in));
    M3D_CHECK_LAUNCH();
    size_t tmp = 0;
    M3D_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys_in, keys, idx_in, idx, (int)n, 0, 64, st));
    if (ctx->s_sort.ensure(tmp)) return M3D_ERR_CUDA;
    {
        KernelScope ks(ctx, KF_CCL_SORT, st);
        M3D_CUDA(cub::DeviceRadixSort::SortPairs(ctx->s_sort.ptr, tmp, keys_in, keys, idx_in, idx, (int)n, 0, 64, st));
    }
    *G_out = G;
    *keys_out = keys;
    *idx_out = idx;
    return M3D_OK;
}

}  // namespace

extern "C" int m3d_table_hist3d(m3d_ctx* ctx, const float* v0_dev, const float* v1_dev, const float* v2_dev,
                                const uint8_t* blank_dev, int64_t n, const float* edges0_host, int n0,
                                const float* edges1_host, int n1, const float* edges2_host, int n2,
                                int32_t* flat_bin_dev, int32_t* all_hist_dev, int32_t* blank_hist_dev, void* stream) {
    if (!ctx || !v0_dev || !v1_dev || !v2_dev || !blank_dev || !flat_bin_dev || !all_hist_dev || !blank_hist_dev ||
        !edges0_host || !edges1_host || !edges2_host || n < 0)
        return m3d_fail(M3D_ERR_ARG, "m3d_table_hist3d: null argument");
    const int ns[3] = {n0, n1, n2};
    const float* es[3] = {edges0_host, edges1_host, edges2_host};
    Edges3 E;
    memset(&E, 0, sizeof(E));
    for (int a = 0; a < 3; ++a) {
        if (ns[a] < 2 || ns[a] > MAX_EDGES)
            return m3d_fail(M3D_ERR_ARG, "m3d_table_hist3d: axis %d needs 2..%d edges, got %d", a, MAX_EDGES, ns[a]);
        E.n[a] = ns[a];
        for (int i = 0; i < ns[a]; ++i) E.e[a][i] = es[a][i];
    }
    if (n == 0) return M3D_OK;
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int n_bins = (n0 - 1) * (n1 - 1) * (n2 - 1);
    const size_t smem = (size_t)n_bins * 2 * sizeof(int) <= 40 * 1024 ? (size_t)n_bins * 2 * sizeof(int) : 0;
    const int blocks = (int)std::min<size_t>(((size_t)n + TB - 1) / TB, (size_t)ctx->num_sms * 4);
    M3D_LAUNCH(ctx, KF_TABLE_HIST, st,
               table_hist3d_kernel<<<blocks, TB, smem, st>>>(v0_dev, v1_dev, v2_dev, blank_dev, (size_t)n, E, n_bins,
                                                             flat_bin_dev, all_hist_dev, blank_hist_dev));
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

extern "C" int m3d_overlap_duplicates(m3d_ctx* ctx, const double* zyx_dev, const int32_t* tile_dev,
                                      const double* distance_min_dev, int64_t n, double radius, uint8_t* drop_dev,
                                      void* stream) {
    if (!ctx || !zyx_dev || !tile_dev || !distance_min_dev || !drop_dev || n < 0)
        return m3d_fail(M3D_ERR_ARG, "m3d_overlap_duplicates: null argument");
    if (!(radius > 0.0) || n >= 0x7fffffffll) return m3d_fail(M3D_ERR_ARG, "m3d_overlap_duplicates: bad radius / n");
    if (n == 0) return M3D_OK;
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    GridSpec G;
    unsigned long long* keys = nullptr;
    uint32_t* idx = nullptr;
    // cells a hair wider than the radius: neighbours within r are then always in adjacent cells, whatever
    // the rounding of (v - origin) * inv_cell
    const double c = radius * (1.0 + 1.0e-6);
    double cell[3] = {c, c, c};
    int rc = build_grid(ctx, zyx_dev, tile_dev, (size_t)n, cell, 1, &G, &keys, &idx, st, "m3d_overlap_duplicates");
    if (rc) return rc;
    const int nb = (int)(((size_t)n + TB - 1) / TB);
    M3D_LAUNCH(ctx, KF_TABLE_OVERLAP, st,
               overlap_kernel<<<nb, TB, 0, st>>>(zyx_dev, tile_dev, distance_min_dev, (size_t)n, G, radius * radius,
                                                  keys, idx, drop_dev));
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

extern "C" int m3d_within_tile_duplicates(m3d_ctx* ctx, const double* zyx_dev, const int32_t* tile_dev,
                                          const int32_t* gene_dev, const double* distance_min_dev, int64_t n,
                                          double radius_xy, double radius_z, uint8_t* drop_dev, void* stream) {
    if (!ctx || !zyx_dev || !tile_dev || !gene_dev || !distance_min_dev || !drop_dev || n < 0)
        return m3d_fail(M3D_ERR_ARG, "m3d_within_tile_duplicates: null argument");
    if (!(radius_xy > 0.0) || !(radius_z >= 0.0) || n >= 0x7fffffffll)
        return m3d_fail(M3D_ERR_ARG, "m3d_within_tile_duplicates: bad radius / n");
    if (n == 0) return M3D_OK;
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    GridSpec G;
    unsigned long long* keys = nullptr;
    uint32_t* idx = nullptr;
    const double c = radius_xy * (1.0 + 1.0e-6);
    double cell[3] = {1.0, c, c};
    int rc = build_grid(ctx, zyx_dev, tile_dev, (size_t)n, cell, 0, &G, &keys, &idx, st, "m3d_within_tile_duplicates");
    if (rc) return rc;
    const size_t nn = (size_t)n;
    const size_t a4 = ((nn * 4 + 255) / 256) * 256, a8 = ((nn * 8 + 255) / 256) * 256;
    if (ctx->s_roots.ensure(2 * a4 + a8)) return M3D_ERR_CUDA;
    char* base = reinterpret_cast<char*>(ctx->s_roots.ptr);
    uint32_t* parent = reinterpret_cast<uint32_t*>(base);
    uint32_t* best_i = reinterpret_cast<uint32_t*>(base + a4);
    unsigned long long* best_d = reinterpret_cast<unsigned long long*>(base + 2 * a4);
    const int nb = (int)((nn + TB - 1) / TB);
    M3D_LAUNCH(ctx, KF_TABLE_WITHIN, st, iota_kernel<<<nb, TB, 0, st>>>(parent, nn));
    M3D_CUDA(cudaMemsetAsync(best_i, 0xFF, nn * 4, st));
    M3D_CUDA(cudaMemsetAsync(best_d, 0xFF, nn * 8, st));
    M3D_LAUNCH(ctx, KF_TABLE_WITHIN, st,
               within_union_kernel<<<nb, TB, 0, st>>>(zyx_dev, tile_dev, gene_dev, nn, G, radius_xy * radius_xy,
                                                       radius_z, keys, idx, parent));
    M3D_LAUNCH(ctx, KF_TABLE_WITHIN, st,
               within_best_d_kernel<<<nb, TB, 0, st>>>(distance_min_dev, nn, parent, best_d, best_i));
    M3D_LAUNCH(ctx, KF_TABLE_WITHIN, st,
               within_best_i_kernel<<<nb, TB, 0, st>>>(distance_min_dev, nn, parent, best_d, best_i));
    M3D_LAUNCH(ctx, KF_TABLE_WITHIN, st, within_drop_kernel<<<nb, TB, 0, st>>>(nn, parent, best_i, drop_dev));
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}

extern "C" int m3d_assign_cells(m3d_ctx* ctx, const double* yx_dev, int64_t n, const double* verts_yx_dev,
                                const int64_t* poly_offsets_dev, const double* bbox_dev, const int32_t* cell_start_dev,
                                const int32_t* cell_polys_dev, double origin_y, double origin_x, double cell_size,
                                int grid_y, int grid_x, int32_t* cell_id_dev, void* stream) {
    if (!ctx || n < 0 || grid_y < 1 || grid_x < 1 || !(cell_size > 0.0))
        return m3d_fail(M3D_ERR_ARG, "m3d_assign_cells: bad argument");
    if (n == 0) return M3D_OK;
    if (!yx_dev || !verts_yx_dev || !poly_offsets_dev || !bbox_dev || !cell_start_dev || !cell_polys_dev || !cell_id_dev)
        return m3d_fail(M3D_ERR_ARG, "m3d_assign_cells: null argument");
    M3D_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nb = (int)(((size_t)n + TB - 1) / TB);
    M3D_LAUNCH(ctx, KF_TABLE_CELLS, st,
               assign_cells_kernel<<<nb, TB, 0, st>>>(yx_dev, (size_t)n, verts_yx_dev,
                                                       reinterpret_cast<const long long*>(poly_offsets_dev), bbox_dev,
                                                       cell_start_dev, cell_polys_dev, origin_y, origin_x, 1.0 / cell_size,
                                                       grid_y, grid_x, cell_id_dev));
    M3D_CHECK_LAUNCH();
    return M3D_OK;
}
