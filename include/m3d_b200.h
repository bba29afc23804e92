/*
 * m3d_b200.h -- C ABI of the B200-native PixelDecoder hot path (libm3d_b200.so).
 *
 * The reference (QI2lab/merfish3d-analysis 0.13.0) has no FFI layer: its hot path is the
 * Python class `PixelDecoder` calling CuPy / cuVS / cuCIM / scikit-image.  Each entry
 * point below replaces the private method(s) cited beside it ("PD" =
 * src/merfish3danalysis/PixelDecoder.py).  The host mirror of the class
 * (merfish3d-analysis_b200/PixelDecoder.py) binds these with ctypes; INTEGRATION.md shows
 * the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer named *_dev is DEVICE memory owned by the caller (torch tensors in the
 *     Python host); the library never frees or retains it past the call's stream work.
 *     Pointers named *_host are plain host memory read before the call returns.
 *   - `stream` is a cudaStream_t passed as void*; calls enqueue on it and only synchronise
 *     where a host-visible count is returned (documented per call).
 *   - dims = {z, y, x}; image layout is C-order (bits, z, y, x) exactly like the
 *     reference's `_image_data` (PD:1894).
 *   - return value 0 = ok, negative = error; m3d_last_error() gives the text
 *     (thread-local).  There is NO CPU fallback: every call fails if no CUDA device.
 */
#ifndef M3D_B200_H
#define M3D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3D_ABI_VERSION 1

#define M3D_DTYPE_U16 0
#define M3D_DTYPE_F32 1

#define M3D_OK 0
#define M3D_ERR_ARG -1
#define M3D_ERR_CUDA -2
#define M3D_ERR_CAPACITY -3
#define M3D_ERR_STATE -4

#define M3D_MAX_BITS 32
/* feature-table columns before the per-bit means (see m3d_features) */
#define M3D_TABLE_FIXED_COLS 14

typedef struct m3d_ctx m3d_ctx;

int m3d_abi_version(void);
const char* m3d_last_error(void);

/* PixelDecoder.__init__ / _normalize_codebook (PD:449-557, PD:879-906): one context per
 * process/GPU.  codebook_unit_host = (K, n_bits) float32 unit codewords (the reference's
 * `_decoding_matrix` cast to float32, PD:2497-2498); excluded_host = codeword row indices
 * whose winning assignments are suppressed (PD:863-877). */
int m3d_create(int device, int n_bits, int n_codewords, const float* codebook_unit_host,
               const int32_t* excluded_host, int n_excluded, m3d_ctx** out);
int m3d_destroy(m3d_ctx* ctx);

/* _scale_pixel_traces vectors (PD:2577-2592).  NULL, NULL = decode unnormalised. */
int m3d_set_normalization(m3d_ctx* ctx, const float* background_host,
                          const float* normalization_host);
/* pixel_assignment_threshold (PD:778-785) and magnitude_threshold (PD:2613-2614); compared
 * in float32 like the reference (NEP-50 weak Python scalars). */
int m3d_set_thresholds(m3d_ctx* ctx, float pixel_threshold, float magnitude_lo,
                       float magnitude_hi);

/* Host -> device staging for `_load_bit_data` (PD:1861-1894 hands the GPU NumPy arrays, i.e. pageable memory).
 * Page-locked sources are copied with one cudaMemcpyAsync; pageable ones go through a ring of pinned slots
 * owned by the context, filled by a few host threads while earlier slots drain over PCIe (~49 GB/s instead
 * of the ~11 GB/s of a pageable cudaMemcpy).  The call returns when every byte has been STAGED or enqueued:
 * the sources may be released, the copies complete in order on `stream`.  m3d_upload_batch runs all pieces
 * (e.g. the bit volumes of one tile) through one pipeline so the ring never idles between them. */
int m3d_upload(m3d_ctx* ctx, const void* src_host, void* dst_dev, int64_t n_bytes, void* stream);
int m3d_upload_batch(m3d_ctx* ctx, int n_pieces, const void* const* src_host, void* const* dst_dev,
                     const int64_t* n_bytes, void* stream);
/* Same, and `on_piece(piece, user)` is called on the calling thread as soon as piece `piece` is completely
 * enqueued on `stream` (pieces complete in order).  The loader uses it to record an event and start the
 * per-bit low-pass of that volume on another stream while the next bit volumes are still crossing PCIe. */
typedef void (*m3d_piece_callback)(int piece, void* user);
int m3d_upload_batch_cb(m3d_ctx* ctx, int n_pieces, const void* const* src_host, void* const* dst_dev,
                        const int64_t* n_bytes, void* stream, m3d_piece_callback on_piece, void* user);

/* ---- image store -> HBM.  Replaces `_load_from_zarr_array` (qi2labDataStore.py:2235-2267:
 * tensorstore decodes level "0" of an OME-NGFF v0.5 image into a NumPy array) + the cp.asarray that follows it in
 * `_load_bit_data` (PD:1861-1881), for the arrays `_create_array_tensorstore_qi2lab` writes (DS:1425-1529: Zarr
 * v3, regular chunks, `bytes` + `blosc` {zstd | lz4, bitshuffle} or `zstd` or no compression, optionally inside
 * `sharding_indexed` shards).  The host side (zarr_store.py) reads `zarr.json` and lists the chunks; the library
 * reads each chunk's bytes into a page-locked slot -- Blosc-LZ4 frames as they are (the device decodes their
 * streams, one warp each), everything else entropy-decoded by a pool of host threads (system libzstd) -- and the
 * device undoes the Blosc bit / byte shuffle and places the chunk into the destination volume (cropped at the
 * array edge and to the requested window).  Unwritten chunks = fill value.  The call returns when everything is
 * enqueued; when the device decoded any frame it first waits for `stream` to collect the kernels' error flag. */
#define M3D_ZARR_RAW 0    /* `bytes` only */
#define M3D_ZARR_BLOSC 1  /* `bytes` + `blosc` */
#define M3D_ZARR_ZSTD 2   /* `bytes` + `zstd` */
#define M3D_ZARR_ABSENT 3 /* never written (a shard's index says so): fill value, the file is not touched */
typedef struct m3d_zarr_chunk {
    const char* path;       /* file that holds the encoded chunk: a chunk file, or a shard file */
    int64_t offset;         /* byte offset of the encoded chunk in that file */
    int64_t length;         /* encoded length; < 0: up to the end of the file.  A file that does not exist = unwritten */
    int32_t codec;          /* M3D_ZARR_* */
    int32_t elem_size;      /* bytes per element: 1, 2, 4 or 8 (little endian) */
    int64_t chunk_shape[3]; /* (z, y, x) of the stored chunk (edge chunks are stored whole) */
    int64_t origin[3];      /* where chunk element (0,0,0) lands in the destination; may be negative / past the end */
    void* dst;              /* destination volume, C order, dst_shape elements of elem_size bytes */
    int64_t dst_shape[3];
    uint64_t fill_bits;     /* the array's fill value, as the bit pattern of one element */
    int32_t piece;          /* callback group (e.g. the bit volume the chunk belongs to); ascending */
    int32_t reserved;
} m3d_zarr_chunk;
/* Device destinations.  Returns when every chunk is decoded and ENQUEUED on `stream` (copies and kernels complete
 * in order there); `on_piece(piece, user)` is called on the calling thread as soon as the last chunk of a piece
 * is enqueued, exactly like m3d_upload_batch_cb. */
int m3d_zarr_read_chunks(m3d_ctx* ctx, int n_chunks, const m3d_zarr_chunk* chunks, void* stream,
                         m3d_piece_callback on_piece, void* user);
/* Host destinations (`dst` = host memory), same decode, un-shuffle and placement on the calling threads: what
 * `tensorstore.read().result()` returns to callers that want a NumPy array (shape probes, the normalisation
 * sampler).  Needs no GPU. */
int m3d_zarr_read_chunks_host(int n_chunks, const m3d_zarr_chunk* chunks);
/* One Blosc-1 frame on the host.  info: out = {nbytes, blocksize, cbytes, typesize, flags, codec id}.
 * encode: cname 4 = zstd, 1 = lz4; shuffle 0 none / 1 byte / 2 bit; blocksize 0 = c-blosc's 256 KiB (zstd,
 * clevel 5); blocks are never split (header flag 0x10), which every Blosc-1 reader accepts. */
int m3d_blosc_info(const void* frame, int64_t n_bytes, int64_t out[6]);
int m3d_blosc_decode_host(const void* frame, int64_t n_bytes, void* dst, int64_t dst_capacity);
int64_t m3d_blosc_encode_bound(int64_t n_bytes, int64_t blocksize);
int m3d_blosc_encode_host(const void* src, int64_t n_bytes, int typesize, int cname, int clevel, int shuffle,
                          int64_t blocksize, void* dst, int64_t dst_capacity, int64_t* out_bytes);
/* plain zstd frame (the Zarr v3 `zstd` codec): compress != 0 encodes at `level`, else decodes */
int m3d_zstd_host(int compress, const void* src, int64_t n_bytes, void* dst, int64_t dst_capacity, int level,
                  int64_t* out_bytes);
/* One zstd frame decoded by the library's OWN decoder (csrc/zstd_decode.cuh: RFC 8878, allocation-free, written to
 * run on the device next) instead of libzstd.  Test hook: the product path does not call it yet. */
int m3d_zstd_decode_builtin(const void* src, int64_t n_bytes, void* dst, int64_t dst_capacity, int64_t* out_bytes);
/* The same through the team-of-lanes arrangement (csrc/zstd_lanes.cuh: one lane per Huffman stream, shared copies)
 * with a team of one: the host pin of the code the second device version runs. */
int m3d_zstd_decode_builtin_lanes(const void* src, int64_t n_bytes, void* dst, int64_t dst_capacity, int64_t* out_bytes);

/* _load_bit_data weighting (PD:1879-1881): out = float32(readout) * float32(predictor). */
int m3d_weight(m3d_ctx* ctx, const uint16_t* readout_dev, const float* predictor_dev,
               int64_t n, float* out_dev, void* stream);

/* Decode-time warp of one bit volume into the round-1 frame (PD:1882-1889 ->
 * utils/decode_warping.py:184-245 -> utils/multiview_registration.py:797-902):
 * scipy.ndimage.affine_transform(float32(in) * predictor, matrix, offset, order=1,
 * mode='constant', cval=0) evaluated in float64 like SciPy.  matrix_host (3x3 row-major) and
 * offset_host are the PIXEL-space matrix_px / offset_px of the reference.  in_dev holds the full
 * (z,y,x) volume; only output planes [out_z0, out_z0 + out_nz) are written to out_dev (float32,
 * (out_nz,y,x)), which is how z_range cropping and z-slab sharding request their planes.
 * predictor_dev is nullable.  Bits whose round carries a SOFIMA flow field go through m3d_warp_flow. */
int m3d_warp_affine(m3d_ctx* ctx, const void* in_dev, int in_dtype, const float* predictor_dev,
                    const int64_t dims[3], const double matrix_host[9], const double offset_host[3],
                    int64_t out_z0, int64_t out_nz, float* out_dev, void* stream);

/* Affine + SOFIMA flow-field warp of one bit volume (PD:1882-1889 -> utils/decode_warping.py:248-305 ->
 * utils/multiview_registration.py:905-1131).  flow_dev = (3, fz, fy, fx) float32, channels X, Y, Z displacements in
 * reference pixels on a grid of stride `stride_zyx` whose first sample sits at `box_start_zyx`.  Per output voxel:
 * flow interpolated (SciPy order-1, constant 0 outside), added to the voxel index, mapped through
 * transform (4x4 row-major float32, physical z,y,x) with spacing / origin, and the moving volume sampled once
 * (SciPy order-1, constant 0) -- float32 coordinate arithmetic in the reference's order of operations.
 * Output planes [out_z0, out_z0 + out_nz) of the (out_dims) reference grid are written to out_dev. */
int m3d_warp_flow(m3d_ctx* ctx, const void* in_dev, int in_dtype, const float* predictor_dev,
                  const int64_t dims[3], const float transform_host[16], const float spacing_host[3],
                  const float origin_host[3], const float* flow_dev, const int64_t flow_dims[3],
                  const float stride_zyx_host[3], const float box_start_zyx_host[3],
                  const int64_t out_dims[3], int64_t out_z0, int64_t out_nz, float* out_dev, void* stream);

/* _lp_filter / _lowpass_image (PD:1948-2024): per-volume Gaussian, reflect boundary,
 * radius int(4*sigma+0.5), axis order z,y,x, fp64 accumulation, fp32 result per pass
 * (scipy.ndimage.gaussian_filter semantics).  mode2d!=0 filters y,x only (per plane).
 * in_dev = n_vols volumes of `in_dtype`; predictor_dev (nullable, float32, same shape) is
 * multiplied in first (PD:1879-1881).  out_dev may not alias in_dev. */
int m3d_lowpass(m3d_ctx* ctx, const void* in_dev, int in_dtype, const float* predictor_dev,
                int n_vols, const int64_t dims[3], const double sigma[3], int mode2d,
                float* out_dev, void* stream);

/* Arithmetic of m3d_lowpass.  0 (default): float64 accumulation with SciPy's symmetric formula -- what the NumPy/SciPy
 * stand-in behind the reference-generated goldens computes.  1: float32 weights and float32 FMA accumulation, taps in
 * ascending order -- what cupyx.scipy.ndimage.gaussian_filter (the call at PD:1972-1979) is believed to do for float32
 * images; NOT pinned (CuPy is not available to the build), opt-in, HBM-bound instead of float64-pipe-bound. */
int m3d_set_lowpass_mode(m3d_ctx* ctx, int mode);

/* _decode_pixels (PD:2523-2643) fused: scale -> clip -> L2 norm -> nearest codeword ->
 * pixel gate -> magnitude gates -> exclusion.  decoded_dev int16 (z,y,x) is always
 * written.  magnitude/distance (float16 (z,y,x)) and scaled (float16 (bits,z,y,x)) are
 * the reference's result images after round(.,5); pass all three NULL for the
 * production fast path (search only where the magnitude gate passes; features recompute
 * what they need). */
int m3d_decode(m3d_ctx* ctx, const void* stack_dev, int dtype, const int64_t dims[3],
               int16_t* decoded_dev, uint16_t* magnitude_f16_dev, uint16_t* distance_f16_dev,
               uint16_t* scaled_f16_dev, void* stream);

/* _extract_barcodes labelling + size filters (PD:2946-2989): equal-value connected
 * components (26-conn 3D / 8-conn per plane when mode2d), drop area > maximum_pixels and
 * area <= max(int(minimum_pixels)-1, 0).  Surviving components get canonical ids 0..n-1 in
 * raster order of their first voxel.  labels_dev (nullable, int32 (z,y,x)) receives id+1
 * (0 = background or dropped for being too small, -1 = dropped for being too large).  Synchronises the stream to return *n_features_out. */
int m3d_label(m3d_ctx* ctx, const int16_t* decoded_dev, const int64_t dims[3], int mode2d,
              double minimum_pixels, int maximum_pixels, int32_t* labels_dev,
              int64_t* n_features_out, void* stream);

/* Fused production path: m3d_decode (no result images) + m3d_label in one call.  The candidate
 * search kernel hands the decoded foreground voxels straight to the labelling stage, so the
 * decoded image is never re-read.  Same results as the two separate calls. */
int m3d_decode_label(m3d_ctx* ctx, const void* stack_dev, int dtype, const int64_t dims[3],
                     int16_t* decoded_dev, int mode2d, double minimum_pixels, int maximum_pixels,
                     int32_t* labels_dev, int64_t* n_features_out, void* stream);

/* Same, for a decoded image the caller keeps ACROSS calls (one persistent int16 buffer per GPU): when decoded_dev
 * and dims equal the previous m3d_decode_label_persistent call's and nothing else wrote to the buffer or ran
 * m3d_label / m3d_decode on this context in between, the previous tile's foreground voxels are reset to -1 from
 * the foreground list still held by the context and the streaming pass skips its dense -1 fill (0.8 GB of the
 * 14.2 GB it moves per config-2 tile).  The result is identical to m3d_decode_label. */
int m3d_decode_label_persistent(m3d_ctx* ctx, const void* stack_dev, int dtype, const int64_t dims[3],
                                int16_t* decoded_dev, int mode2d, double minimum_pixels, int maximum_pixels,
                                int32_t* labels_dev, int64_t* n_features_out, void* stream);

/* Z-slab sharding of one volume (no reference counterpart; contract: same result as the unsharded
 * volume).  Given the LAST decoded / label plane of the lower slab and the FIRST plane of
 * the upper slab (labels = ids + 1 from m3d_label / m3d_decode_label with a labels image; 0 =
 * background or dropped-small, -1 = dropped because oversized), emits (label_hi, label_lo) pairs of
 * components joined across the interface under the 26-neighbour equal-value rule.  pairs_dev =
 * capacity x 2 int32; *n_pairs_out may exceed capacity (then only `capacity` pairs were stored).
 * Synchronises the stream. */
int m3d_interface_pairs(m3d_ctx* ctx, const int16_t* decoded_lo_dev, const int32_t* labels_lo_dev,
                        const int16_t* decoded_hi_dev, const int32_t* labels_hi_dev, int64_t Y, int64_t X,
                        int32_t* pairs_dev, int64_t capacity, int64_t* n_pairs_out, void* stream);

/* _extract_barcodes regionprops (PD:2991-3062) for the components of the last m3d_label /
 * m3d_decode_label call, one row per component in canonical order.  table_dev = (n_rows, 14 + n_bits)
 * float64, columns:
 *   0 first_voxel (linear index)  1 area  2 decoded_id
 *   3..5 centroid z,y,x           6..11 central 2nd moments zz,yy,xx,zy,zx,yx (sums)
 *   12 distance_min               13 magnitude_mean      14.. per-bit intensity_mean
 * optimize_mode=0: bit means over the float16 scaled image (result float16-rounded);
 * optimize_mode=1: over the raw float32 input stack (PD:2935-2941). */
int m3d_features(m3d_ctx* ctx, const void* stack_dev, int dtype, const int64_t dims[3],
                 const int16_t* decoded_dev, int optimize_mode, double* table_dev,
                 int64_t n_rows, void* stream);

/* _global_normalization_vectors order statistics (PD:1113-1177).  One radix-select pass:
 * histogram of bits [shift, shift+11) of the order-preserving uint32 key of
 * v = clip0 ? max(x - sub, 0) : x - sub, over elements whose key matches
 * (key & prefix_mask) == prefix_value and that satisfy the predicate
 * (pred 0: all, 1: v < cutoff, 2: v > cutoff).  hist_dev = 2048 uint64 counters,
 * accumulated (caller zeroes).  */
int m3d_select_hist(m3d_ctx* ctx, const float* data_dev, int64_t n, float sub, int clip0,
                    int pred, float cutoff, uint32_t prefix_mask, uint32_t prefix_value,
                    int shift, unsigned long long* hist_dev, void* stream);
/* Several digit histograms in ONE launch: the optimiser's per-iteration statistic (PD:1290-1356: per-bit medians of
 * on-bit / off-bit mean intensities) walks 2 x bits small multisets together, one histogram per multiset and level.
 * Query q adds the histogram of data[q][0 .. n[q]) -- keys with (key & prefix_mask[q]) == prefix_value[q], digit
 * (key >> shift) & 2047 -- to hist_dev + 2048 * q.  The four arrays are HOST arrays of n_queries entries; entries with
 * n == 0 are skipped (data may be NULL).  No predicate, no subtraction; NaN values are "no entry" and are not counted
 * (pandas' median(skipna=True) of the reference's sparse frames), so a multiset may be handed over as a dense column
 * with NaN in the rows that do not belong to it. */
int m3d_select_hist_batch(m3d_ctx* ctx, int n_queries, const float* const* data_dev, const int64_t* n,
                          const uint32_t* prefix_mask, const uint32_t* prefix_value, int shift,
                          unsigned long long* hist_dev, void* stream);
/* hot-pixel replacement (PD:1072-1074): data[data > threshold] = value, in place. */
int m3d_replace_above(m3d_ctx* ctx, float* data_dev, int64_t n, float threshold, float value,
                      void* stream);

/* ---- post-decode transcript-table stage; rows = decoded transcripts ----
 *
 * _filter_all_barcodes_blank_fraction binning (PD:3656-3742): per row
 * b_a = searchsorted(edges_a, v_a, side='right') - 1 on float32 values and float32 edges for the three
 * feature axes (magnitude_mean, area, distance_min); rows with a non-finite feature or a bin outside
 * [0, n_a - 2] are out of range.  Writes flat_bin_dev[i] = ravel_multi_index(b0, b1, b2) or -1 and
 * accumulates the all-rows and blank-rows histograms ((n0-1)(n1-1)(n2-1) int32 each, caller zeroes).
 * edges_*_host: ascending, 2..64 entries. */
int m3d_table_hist3d(m3d_ctx* ctx, const float* v0_dev, const float* v1_dev, const float* v2_dev,
                     const uint8_t* blank_dev, int64_t n, const float* edges0_host, int n0,
                     const float* edges1_host, int n1, const float* edges2_host, int n2,
                     int32_t* flat_bin_dev, int32_t* all_hist_dev, int32_t* blank_hist_dev, void* stream);

/* _remove_duplicates_in_tile_overlap (PD:4137-4177).  zyx_dev = (n,3) float64 global coordinates.
 * drop_dev[i] = 1 iff some row j of a different tile lies within `radius` (3-D Euclidean, inclusive,
 * float64 sum-of-squares vs radius^2 like SciPy's cKDTree) and (distance_min[j], j) < (distance_min[i], i)
 * -- the outcome of the reference's loop over query_pairs.  distance_min must not be NaN.
 * Synchronises the stream once (coordinate bounds for the grid hash). */
int m3d_overlap_duplicates(m3d_ctx* ctx, const double* zyx_dev, const int32_t* tile_dev,
                           const double* distance_min_dev, int64_t n, double radius, uint8_t* drop_dev,
                           void* stream);

/* _remove_duplicates_within_tile (PD:4179-4363, 2-D decode mode).  Rows are neighbours iff same tile,
 * same gene code, XY distance <= radius_xy and 0 < |dz| <= radius_z; per connected cluster every row but
 * the one with the smallest (distance_min, row index) gets drop_dev[i] = 1.  gene_dev = any int32
 * coding of gene_id equality.  Synchronises the stream once. */
int m3d_within_tile_duplicates(m3d_ctx* ctx, const double* zyx_dev, const int32_t* tile_dev,
                               const int32_t* gene_dev, const double* distance_min_dev, int64_t n,
                               double radius_xy, double radius_z, uint8_t* drop_dev, void* stream);

/* _add_on_bit_weighted_centroids / _plane_wise_weighted_centroid_statistics (PD:2701-2906):
 * per component and per ON bit of its codeword, over the label image dilated along z by a
 * `z_support`-plane maximum window: sums of w, w*z, w*y, w*x with w = max(float32(intensity), 0)
 * (float64 accumulators; w*y and w*x are float32 products like the reference) and, over the undilated
 * labels, the float32 peak of w.  labels_dev = ids + 1 from m3d_label / m3d_decode_label ((z,y,x) int32;
 * <= 0 = background); label_code_dev[l] = codeword row of label l (int16, minlength entries, < 0 = skip).
 * sums_dev = (minlength, n_bits, 4) float64 {w, wz, wy, wx}; peak_dev = (minlength, n_bits) float32;
 * both are zeroed by the call; entries of bits that are off in the label's codeword stay zero (the
 * reference never reads them).  One pass for all bits.  float64 atomics: equal to the reference's raster
 * bincount to round-off, not bit-for-bit. */
int m3d_centroid_statistics(m3d_ctx* ctx, const int32_t* labels_dev, const void* stack_dev, int dtype,
                            const int64_t dims[3], int z_support, const int16_t* label_code_dev,
                            int64_t minlength, double* sums_dev, float* peak_dev, void* stream);

/* scikit-image `inertia_tensor_eigvals` of every row of an m3d_features table (PD:3038-3047): eigenvalues of
 * the inertia tensor from columns 1 (area) and 6..11 (central second moments), clipped at 0, descending;
 * eigvals_dev = (n_rows, 3) float64.  table_dev rows have n_cols (>= 14) float64 entries. */
int m3d_inertia_eigvals(m3d_ctx* ctx, const double* table_dev, int64_t n_rows, int64_t n_cols,
                        double* eigvals_dev, void* stream);

/* _assign_cells (PD:4076-4135): cell_id_dev[i] = 1 + index of the lowest-numbered polygon containing point i
 * (yx_dev = (n,2) float64 global_y, global_x), 0 when none does.  Polygons: verts_yx_dev (total,2) float64 with
 * poly_offsets_dev (P+1) int64, bbox_dev (P,4) = ymin,xmin,ymax,xmax.  Candidate lookup through a uniform grid
 * built by the caller: cell (gy,gx) of size cell_size at (origin_y, origin_x) lists the polygons whose box
 * touches it in cell_polys_dev[cell_start_dev[c] .. cell_start_dev[c+1]) (ascending polygon index).  Containment
 * = even-odd crossing rule in float64 (shapely `contains` away from polygon boundaries). */
int m3d_assign_cells(m3d_ctx* ctx, const double* yx_dev, int64_t n, const double* verts_yx_dev,
                     const int64_t* poly_offsets_dev, const double* bbox_dev, const int32_t* cell_start_dev,
                     const int32_t* cell_polys_dev, double origin_y, double origin_x, double cell_size,
                     int grid_y, int grid_x, int32_t* cell_id_dev, void* stream);

/* Capacity (entries) of the search -> regionprops record buffers; 0 = automatic
 * (max(2^20, n_vox/16)).  When the foreground exceeds it the regionprops kernel recomputes the
 * traces instead; results are identical.  Exposed so tests can force the overflow path. */
int m3d_set_sparse_capacity(m3d_ctx* ctx, int64_t entries);

/* number of kernels this context has launched since creation (bench.py gpu_launches). */
int64_t m3d_launch_count(m3d_ctx* ctx);
/* name of the i-th kernel family; NULL when i is out of range. */
const char* m3d_kernel_name(int i);
/* launches of the i-th kernel family by this context (0 when i is out of range). */
int64_t m3d_kernel_launches(m3d_ctx* ctx, int i);
/* Per-family device timing for bench.py's roofline: when enabled, every launch is bracketed
 * by CUDA events on its own stream; m3d_kernel_time_ms synchronises on the recorded events
 * and returns the family's accumulated milliseconds.  m3d_reset_counters zeroes launches and
 * times. */
int m3d_set_timing(m3d_ctx* ctx, int enable);
int m3d_reset_counters(m3d_ctx* ctx);
double m3d_kernel_time_ms(m3d_ctx* ctx, int i);

#ifdef __cplusplus
}
#endif
#endif /* M3D_B200_H */
