"""CPU oracle: NumPy/SciPy restatement of merfish3d-analysis' PixelDecoder hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Every function cites the reference lines it restates.  ``PD`` abbreviates
``/root/reference/src/merfish3danalysis/PixelDecoder.py`` (package 0.13.0).

Pin status
----------
PINNED against the reference's own code.  ``tests/golden/reference_shims.py`` executes
``/root/reference/src/merfish3danalysis/PixelDecoder.py`` UNMODIFIED with its GPU wheels
(CuPy / cupyx / cuVS / cuCIM / scikit-image, none installed here) answered by NumPy / SciPy,
and ``tests/golden/make_reference_golden.py`` stores what it computes for seeded tiles
(raw 3-D, low-pass + predictor weights, 2-D mode, exclusions + z crop) and for a 3-tile x
3-iteration ``optimize_normalization_by_decoding`` run.  ``tests/test_cpu_reference_golden.py``
requires this oracle to reproduce those fixtures: decoded / magnitude / distance / scaled images
bit-exact, every table column identical (eigenvalues to 1e-9), normalisation vectors identical;
one more case runs the reference live when ``/root/reference`` is present.  Also restated from
upstream's known-answer tests: exclusion semantics, ``_warp_pixel``, codebook thresholds
(``tests/test_optimization_codeword_exclusions.py:114-120``,
``tests/test_pixeldecoder_coordinates.py:6-41``, ``PD:778-791``).

What the pin cannot reach (third-party GPU kernels, not vendored, no GPU here): cuVS computes
the Euclidean distance in expanded form (this oracle and the stand-in use the direct form,
float32, sequential over bits -- O(1e-7) apart, inside the north-star 1e-6 tie allowance) and
``cupyx.scipy.ndimage.gaussian_filter`` may accumulate float32 where SciPy accumulates
float64.  ``correlate1d_restated`` is checked bit-for-bit against SciPy.
"""

from __future__ import annotations

import hashlib
from itertools import product
from typing import Sequence

import numpy as np
import pandas as pd
from scipy import ndimage as ndi
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

F32 = np.float32

DEFAULT_DECODE_LOWPASS_SIGMA = (3.0, 1.0, 1.0)  # PD:128
DEFAULT_DECODE_MAGNITUDE_THRESHOLD = (1.5, 10.0)  # PD:129
DEFAULT_2D_MINIMUM_PIXELS = 7.0  # PD:130
DEFAULT_3D_MINIMUM_PIXELS = 16.0  # PD:131
MAXIMUM_PIXELS = 500  # PD:2909


# --------------------------------------------------------------------------- codebook
def load_codebook(codebook: pd.DataFrame, n_bits: int) -> dict:
    """PD:756-800 -- drop 1-on-bit rows, derive both thresholds from the code geometry."""
    df = codebook.copy()
    df = df.fillna(0)
    bit_columns = df.columns[1 : n_bits + 1]
    on_counts = df.loc[:, bit_columns].to_numpy(dtype=np.int8).sum(axis=1)
    df = df.loc[on_counts != 1].reset_index(drop=True)
    matrix = df.loc[:, bit_columns].to_numpy(dtype=int)
    on = int(np.median(on_counts[on_counts != 1]))
    pixel_thr = float(np.sqrt(2.0 - 2.0 * ((on - 2.0) / np.sqrt(on * (on - 2.0)))))
    transcript_thr = float(np.sqrt(2.0 - 2.0 * (on / np.sqrt(on * (on + 2.0)))))
    gene_ids = df.iloc[:, 0].tolist()
    blank_count = int(
        df.iloc[:, 0].astype("string").str.lower().str.startswith("blank", na=False).sum()
    )
    return dict(
        matrix=matrix,
        gene_ids=gene_ids,
        on_bit_count=on,
        pixel_assignment_threshold=pixel_thr,
        transcript_distance_threshold=transcript_thr,
        blank_count=blank_count,
    )


def normalize_codebook(matrix: np.ndarray) -> np.ndarray:
    """PD:879-906 -- rows / ||row||_2 (zero norm -> 1); float64 like the reference."""
    m = np.asarray(matrix)
    mag = np.linalg.norm(m, axis=1, keepdims=True)
    mag[mag == 0] = 1
    return m / mag


def codebook_fingerprint(matrix: np.ndarray, gene_ids: Sequence[str]) -> str:
    """PD:843-853."""
    digest = hashlib.sha256()
    m = np.ascontiguousarray(matrix, dtype=np.int8)
    digest.update(np.asarray(m.shape, dtype=np.int64).tobytes())
    for gene_id in gene_ids:
        encoded = str(gene_id).encode("utf-8")
        digest.update(len(encoded).to_bytes(8, byteorder="little"))
        digest.update(encoded)
    digest.update(m.tobytes())
    return digest.hexdigest()


def suppress_excluded(decoded: np.ndarray, nearest: np.ndarray, excluded: Sequence[int]) -> None:
    """PD:863-877 -- excluded *winners* become background; no fall-through."""
    if not len(excluded):
        return
    decoded[np.isin(nearest, np.asarray(excluded, dtype=nearest.dtype))] = -1


def warp_pixel(p, spacing, origin, affine, camera_to_stage=None) -> np.ndarray:
    """PD:2645-2683."""
    phys = p * spacing + origin
    if camera_to_stage is not None:
        phys = (np.asarray(camera_to_stage) @ np.array([*list(phys), 1]))[:-1]
    return (np.array(affine) @ np.array([*list(phys), 1]))[:-1]


# --------------------------------------------------------------------------- input + low-pass
def weight_readout(readout: np.ndarray, predictor: np.ndarray | None) -> np.ndarray:
    """PD:1879-1881 -- float32(readout) * float32(predictor)."""
    r = np.asarray(readout, dtype=F32)
    if predictor is None:
        return r
    return r * np.asarray(predictor, dtype=F32)


def warp_px_arguments(transform_zyx_um, spacing_zyx_um):
    """utils/multiview_registration.py:857-870 -- physical 4x4 transform -> affine_transform
    ``matrix`` / ``offset`` in pixel units, float32 like the reference (origin = 0)."""
    spacing = np.asarray(spacing_zyx_um, dtype=F32)
    origin = np.zeros(3, dtype=F32)
    transform = np.asarray(transform_zyx_um, dtype=F32)
    linear_um = transform[:3, :3]
    translation_um = transform[:3, 3]
    matrix_px = (linear_um * spacing[np.newaxis, :]) / spacing[:, np.newaxis]
    offset_px = (linear_um @ origin + translation_um - origin) / spacing
    return matrix_px, offset_px


def warp_to_reference(image: np.ndarray, transform_zyx_um, spacing_zyx_um) -> np.ndarray:
    """utils/decode_warping.py:118-182 + utils/multiview_registration.py:797-902 -- order-1 affine
    resampling into the round-1 frame; identity transforms return the input unchanged.
    ``cupyx.scipy.ndimage.affine_transform`` is restated by SciPy's (same family, float64 inside)."""
    image = np.asarray(image, dtype=F32)
    if transform_zyx_um is None or np.allclose(transform_zyx_um, np.eye(4, dtype=F32)):
        return image
    matrix_px, offset_px = warp_px_arguments(transform_zyx_um, spacing_zyx_um)
    return ndi.affine_transform(image, matrix_px, offset=offset_px, output_shape=image.shape, order=1,
                                mode="constant", cval=0.0).astype(F32, copy=False)


def warp_to_reference_with_flow(image: np.ndarray, transform_zyx_um, spacing_zyx_um, flow_xyz: np.ndarray,
                                stride_zyx, box_start_xyz, reference_shape=None) -> np.ndarray:
    """utils/decode_warping.py:248-305 + utils/multiview_registration.py:985-1131 -- affine + SOFIMA flow warp:
    the flow (channels X, Y, Z on a strided grid) is interpolated at every output voxel, added to the voxel
    index, pushed through the physical affine and the moving image is sampled once; float32 arrays throughout,
    ``cupyx.scipy.ndimage.map_coordinates`` restated by SciPy's (order 1, constant 0)."""
    image = np.asarray(image, dtype=F32)
    ref_shape = tuple(int(v) for v in (image.shape if reference_shape is None else reference_shape))
    spacing = np.asarray(spacing_zyx_um, dtype=F32)
    origin = np.zeros(3, dtype=F32)
    transform = np.asarray(transform_zyx_um, dtype=F32)
    flow = np.asarray(flow_xyz, dtype=F32)
    stride = np.asarray(stride_zyx, dtype=F32)
    box_start_zyx = np.asarray(box_start_xyz, dtype=F32)[[2, 1, 0]]
    gz, gy, gx = np.meshgrid(np.arange(ref_shape[0], dtype=F32), np.arange(ref_shape[1], dtype=F32),
                             np.arange(ref_shape[2], dtype=F32), indexing="ij")
    flow_coords = np.stack([(gz - box_start_zyx[0]) / stride[0], (gy - box_start_zyx[1]) / stride[1],
                            (gx - box_start_zyx[2]) / stride[2]], axis=0)
    moved = [ident + ndi.map_coordinates(flow[c], flow_coords, order=1, mode="constant", cval=0.0)
             for c, ident in enumerate((gx, gy, gz))]
    pz = moved[2] * spacing[0] + origin[0]
    py = moved[1] * spacing[1] + origin[1]
    px = moved[0] * spacing[2] + origin[2]
    src = []
    for a in range(3):
        m = transform[a, 0] * pz + transform[a, 1] * py + transform[a, 2] * px + transform[a, 3]
        src.append((m - origin[a]) / spacing[a])
    out = ndi.map_coordinates(image, np.stack(src, axis=0), order=1, mode="constant", cval=0.0)
    return out.astype(F32, copy=False)


def lowpass_active(sigma) -> bool:
    """PD:1969, PD:4543-4546."""
    return sigma is not None and not np.any(np.asarray(sigma, dtype=float) == 0)


def lowpass_image(image: np.ndarray, sigma, is_3d: bool) -> np.ndarray:
    """PD:1948-1980 -- one bit volume (z, y, x) float32."""
    if not lowpass_active(sigma):
        return image
    if is_3d:
        return ndi.gaussian_filter(image, sigma=sigma)
    out = np.empty_like(image)
    for z in range(image.shape[0]):
        out[z] = ndi.gaussian_filter(image[z], sigma=(sigma[1], sigma[2]))
    return out


def lowpass_stack(stack: np.ndarray, sigma, is_3d: bool) -> np.ndarray:
    """PD:1982-2024 -- per-bit low-pass of a (bits, z, y, x) float32 stack."""
    if not lowpass_active(sigma):
        return stack
    out = np.empty_like(stack)
    for b in range(stack.shape[0]):
        out[b] = lowpass_image(np.asarray(stack[b], dtype=F32), sigma, is_3d)
    return out


def lowpass_image_float32_accumulation(image: np.ndarray, sigma, is_3d: bool) -> np.ndarray:
    """The OPT-IN float32 arithmetic of ``m3d_lowpass`` (``m3d_set_lowpass_mode(ctx, 1)``), restated: per axis (z, y, x;
    y, x only in 2-D mode) the float64 SciPy weights are cast to float32 and every output is ``acc = fma(x[i + j], w[j],
    acc)`` over the taps j = -r .. r in ascending order, float32 throughout, reflect boundary.  This is what
    cupyx.scipy.ndimage's correlate kernel is believed to compute for float32 images (the reference's call: PD:1972-1979);
    CuPy is not available here, so this restatement pins the KERNEL to its documented arithmetic, not to the reference.
    The fused multiply-add is emulated in extended precision (exact product, one rounding)."""
    out = np.asarray(image, dtype=F32)
    axes = (0, 1, 2) if is_3d else (1, 2)
    sig = tuple(float(v) for v in sigma)
    if not is_3d:
        sig = sig[-2:] if len(sig) == 3 else sig
    for axis, s in zip(axes, sig if is_3d else sig):
        w64, r = gaussian_kernel1d(s)
        w32 = w64.astype(F32)
        n = out.shape[axis]
        idx = np.arange(-r, n + r)
        per = 2 * n
        m = np.mod(idx, per)
        src = np.where(m >= n, per - 1 - m, m)  # scipy 'reflect': d c b a | a b c d | d c b a
        ext = np.take(out, src, axis=axis)
        acc = np.zeros(out.shape, dtype=F32)
        for j in range(2 * r + 1):
            x = np.take(ext, np.arange(j, j + n), axis=axis)
            prod = x.astype(np.longdouble) * np.longdouble(w32[j])  # exact: 24 x 24 bits
            acc = (prod + acc.astype(np.longdouble)).astype(F32)   # one rounding (64-bit significand holds the sum)
        out = acc
    return out


def gaussian_kernel1d(sigma: float, truncate: float = 4.0) -> tuple[np.ndarray, int]:
    """SciPy ``_gaussian_kernel1d`` (order 0): float64 weights, radius int(truncate*sigma+0.5)."""
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    x = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / (sd * sd) * x**2)
    phi = phi / phi.sum()
    return phi[::-1].copy(), lw


def correlate1d_restated(a: np.ndarray, sigma: float, axis: int) -> np.ndarray:
    """Restatement of SciPy ``NI_Correlate1D`` (symmetric branch, mode='reflect').

    float64 line buffers; ``o = x[c]*w[c]; for j=-r..-1: o += (x[c+j]+x[c-j])*w[c+j]``;
    result stored as float32.  Checked bit-for-bit against scipy in tests; the CUDA
    low-pass follows the same fp64 formula.
    """
    w, r = gaussian_kernel1d(sigma)
    a = np.moveaxis(np.asarray(a, dtype=F32), axis, 0)
    n = a.shape[0]
    ad = a.astype(np.float64)
    per = 2 * n
    idx = np.mod(np.arange(-r, n + r), per)
    idx = np.where(idx >= n, per - 1 - idx, idx)
    ext = ad[idx]
    out = ext[r : r + n] * w[r]
    for jj in range(-r, 0):
        out = out + (ext[r + jj : r + jj + n] + ext[r - jj : r - jj + n]) * w[r + jj]
    return np.moveaxis(out.astype(F32), 0, axis)


# --------------------------------------------------------------------------- per-voxel decode
def scale_traces(traces: np.ndarray, bkg: np.ndarray, nrm: np.ndarray) -> np.ndarray:
    """PD:2399-2401 -- (x - bkg[:,None]) / nrm[:,None], float32, true division."""
    b = np.asarray(bkg, dtype=F32)[: traces.shape[0], None]
    n = np.asarray(nrm, dtype=F32)[: traces.shape[0], None]
    return (traces - b) / n


def normalize_traces(traces: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """PD:2459-2462 -- L2 norm over bits; 0 -> inf for the divide, reported as -1."""
    norms = np.linalg.norm(traces, axis=0)
    norms = np.where(norms == 0, np.inf, norms).astype(F32)
    with np.errstate(invalid="ignore"):
        normalized = traces / norms
    norms = np.where(norms == np.inf, -1, norms).astype(F32)
    return normalized, norms


def nearest_codeword(xh: np.ndarray, codebook_unit: np.ndarray, chunk: int = 1 << 15):
    """PD:2500-2513 -- Euclidean distance to every codeword, first argmin, min.

    Documented choice: direct form sqrt(sum_b (xh_b - c_b)^2) in float32,
    accumulated sequentially over bits with separate multiply and add.
    """
    C = np.asarray(codebook_unit, dtype=F32)
    K, B = C.shape
    N = xh.shape[1]
    dmin = np.empty(N, dtype=F32)
    imin = np.empty(N, dtype=np.int64)
    for s in range(0, N, chunk):
        x = xh[:, s : s + chunk]
        t = x[0][None, :] - C[:, 0][:, None]
        acc = t * t
        for b in range(1, B):
            t = x[b][None, :] - C[:, b][:, None]
            acc = acc + t * t
        d = np.sqrt(acc)
        i = np.argmin(d, axis=0)
        imin[s : s + chunk] = i
        dmin[s : s + chunk] = d[i, np.arange(d.shape[1])]
    return dmin, imin


def decode_pixels(
    stack: np.ndarray,
    codebook_unit: np.ndarray,
    bkg: np.ndarray | None,
    nrm: np.ndarray | None,
    pixel_threshold: float,
    magnitude_threshold: Sequence[float] = DEFAULT_DECODE_MAGNITUDE_THRESHOLD,
    excluded: Sequence[int] = (),
) -> dict:
    """PD:2523-2643 -- plane-by-plane decode of a (bits, z, y, x) float32 stack."""
    B, Z, Y, X = stack.shape
    decoded = np.zeros((Z, Y, X), dtype=np.int16)
    magnitude = np.zeros((Z, Y, X), dtype=np.float16)
    distance = np.zeros((Z, Y, X), dtype=np.float16)
    scaled = np.zeros((B, Z, Y, X), dtype=np.float16)
    for z in range(Z):
        traces = np.asarray(stack[:, z], dtype=F32).reshape(B, -1)
        if bkg is not None and nrm is not None:
            with np.errstate(divide="ignore", invalid="ignore"):
                traces = scale_traces(traces, bkg, nrm)
        traces = np.clip(traces, 0.0, 1.0)  # PD:2429
        xh, mag = normalize_traces(traces)
        d, idx = nearest_codeword(xh, codebook_unit)
        dec = np.full(d.shape[0], -1, dtype=np.int16)  # PD:2610-2614
        m = d <= pixel_threshold
        dec[m] = idx[m]
        dec[mag < magnitude_threshold[0]] = -1
        dec[mag > magnitude_threshold[1]] = -1
        suppress_excluded(dec, idx, tuple(excluded))  # PD:2615-2619
        decoded[z] = dec.reshape(Y, X)
        magnitude[z] = np.round(mag, 5).reshape(Y, X)  # PD:2624-2626 (fp16 on store)
        scaled[:, z] = np.round(traces, 5).reshape(B, Y, X)
        distance[z] = np.round(d, 5).reshape(Y, X)
    return dict(decoded=decoded, magnitude=magnitude, distance=distance, scaled=scaled)


# --------------------------------------------------------------------------- CCL
def _neighbour_offsets(is_3d: bool):
    if is_3d:
        offs = [o for o in product((-1, 0, 1), repeat=3) if o > (0, 0, 0)]
    else:
        offs = [(0, dy, dx) for dy, dx in product((-1, 0, 1), repeat=2) if (dy, dx) > (0, 0)]
    return offs


def label_decoded(decoded: np.ndarray, is_3d: bool) -> np.ndarray:
    """PD:2946-2972 -- equal-value components, background -1.

    3D: 26-connectivity over the volume.  2D: 8-connectivity per plane with a running
    offset.  Both number components 1.. in raster order of their first voxel
    (scikit-image numbering; for 2D the running offset yields the same z-major order).
    """
    Z, Y, X = decoded.shape
    flat = decoded.ravel()
    fg = np.flatnonzero(flat != -1)
    labels = np.zeros(flat.shape[0], dtype=np.int32)
    if fg.size == 0:
        return labels.reshape(Z, Y, X)
    lookup = np.full(flat.shape[0], -1, dtype=np.int64)
    lookup[fg] = np.arange(fg.size)
    z, y, x = np.unravel_index(fg, (Z, Y, X))
    rows, cols = [], []
    for dz, dy, dx in _neighbour_offsets(is_3d):
        nz, ny, nx = z + dz, y + dy, x + dx
        ok = (nz >= 0) & (nz < Z) & (ny >= 0) & (ny < Y) & (nx >= 0) & (nx < X)
        src = np.flatnonzero(ok)
        nlin = (nz[src] * Y + ny[src]) * X + nx[src]
        same = flat[nlin] == flat[fg[src]]
        rows.append(src[same])
        cols.append(lookup[nlin[same]])
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    g = coo_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(fg.size, fg.size))
    _, comp = connected_components(g, directed=False)
    # renumber by first occurrence in raster order (fg is ascending)
    _, first = np.unique(comp, return_index=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(order.size, dtype=np.int32)
    rank[order] = np.arange(1, order.size + 1, dtype=np.int32)
    labels[fg] = rank[comp]
    return labels.reshape(Z, Y, X)


def filter_label_sizes(labels: np.ndarray, minimum_pixels, maximum_pixels: int = MAXIMUM_PIXELS):
    """PD:2976-2989 -- drop count > max; drop count <= max(int(min)-1, 0); no renumbering."""
    flat = labels.ravel().copy()
    counts = np.bincount(flat)
    large = np.flatnonzero(counts > maximum_pixels)
    large = large[large != 0]
    if large.size:
        flat[np.isin(flat, large)] = 0
    max_size = max(int(minimum_pixels) - 1, 0)
    counts = np.bincount(flat)
    small = np.flatnonzero(counts <= max_size)
    small = small[small != 0]
    if small.size:
        flat[np.isin(flat, small)] = 0
    return flat.reshape(labels.shape)


# --------------------------------------------------------------------------- regionprops
def _inertia_eigvals(coords: np.ndarray) -> np.ndarray:
    """Appendix B of SURVEY.md (scikit-image ``inertia_tensor_eigvals``, restated)."""
    n = coords.shape[0]
    c = coords.mean(axis=0)
    d = coords.astype(np.float64) - c
    mu2 = (d * d).sum(axis=0)
    T = np.zeros((3, 3))
    for a in range(3):
        T[a, a] = (mu2.sum() - mu2[a]) / n
    for a, b in ((0, 1), (0, 2), (1, 2)):
        v = -(d[:, a] * d[:, b]).sum() / n
        T[a, b] = v
        T[b, a] = v
    ev = np.linalg.eigvalsh(T)
    ev = np.clip(ev, 0, None)
    return np.sort(ev)[::-1]


def region_table(
    labels: np.ndarray,
    distance: np.ndarray,
    magnitude: np.ndarray,
    intensity: np.ndarray,
) -> pd.DataFrame:
    """PD:2991-3062 -- per-label features in ascending-label order.

    ``intensity`` is the (bits, z, y, x) image whose (z, y, x, bits) view the reference
    hands to scikit-image: float16 scaled image in normal decode, float32 raw
    low-passed image in optimiser mode.  Means are taken with ``np.mean`` on the same
    array shapes scikit-image uses, so NumPy's reduction order and float16 rounding are
    reproduced by construction.
    """
    B = intensity.shape[0]
    shape = labels.shape
    flat = labels.ravel()
    fg = np.flatnonzero(flat)
    order = np.argsort(flat[fg], kind="stable")
    vox = fg[order]
    labs = flat[vox]
    uniq, start = np.unique(labs, return_index=True)
    end = np.append(start[1:], labs.size)
    inten_flat = intensity.reshape(B, -1)
    dist_flat = distance.ravel()
    mag_flat = magnitude.ravel()
    rows = []
    for lab, s, e in zip(uniq, start, end):
        v = vox[s:e]
        coords = np.stack(np.unravel_index(v, shape), axis=1)
        cen = coords.mean(axis=0)
        means = np.mean(np.ascontiguousarray(inten_flat[:, v].T), axis=0)
        ev = _inertia_eigvals(coords)
        row = dict(label=int(lab), area=float(e - s), z=cen[0], y=cen[1], x=cen[2])
        for b in range(B):
            row[f"intensity_mean-{b}"] = means[b]
        for k in range(3):
            row[f"inertia_tensor_eigvals-{k}"] = ev[k]
        row["distance_min"] = np.min(dist_flat[v].astype(F32))
        row["magnitude_mean"] = np.mean(mag_flat[v], axis=0)
        row["first_voxel"] = int(v[0])
        rows.append(row)
    cols = (
        ["label", "area", "z", "y", "x"]
        + [f"intensity_mean-{b}" for b in range(B)]
        + [f"inertia_tensor_eigvals-{k}" for k in range(3)]
        + ["distance_min", "magnitude_mean", "first_voxel"]
    )
    if not rows:
        return pd.DataFrame({c: [] for c in cols})
    return pd.DataFrame(rows, columns=cols)


CENTROID_SUFFIXES = ("center_z", "center_y", "center_x", "intensity_sum", "intensity_peak", "voxel_count")


def plane_wise_centroid_statistics(labels: np.ndarray, intensity: np.ndarray, z_support: int, minlength: int):
    """PD:2833-2906 restated in its full-volume form (the one upstream's own test compares with,
    tests/test_optimization_codeword_exclusions.py:153-205): labels dilated along z by a
    ``z_support`` maximum window, weights max(intensity, 0) in float32, float64 bincount sums."""
    half = max(int(z_support) // 2, 0)
    dil = ndi.maximum_filter1d(labels, size=2 * half + 1, axis=0, mode="reflect") if half else labels
    w = np.maximum(np.asarray(intensity, dtype=F32), F32(0))
    zc = np.arange(labels.shape[0], dtype=np.float64)[:, None, None]
    yc = np.arange(labels.shape[1], dtype=F32)[None, :, None]
    xc = np.arange(labels.shape[2], dtype=F32)[None, None, :]
    flat = dil.ravel()
    weight = np.bincount(flat, weights=w.ravel(), minlength=minlength)
    zsum = np.bincount(flat, weights=(w.astype(np.float64) * zc).ravel(), minlength=minlength)
    ysum = np.bincount(flat, weights=(w * yc).ravel(), minlength=minlength)
    xsum = np.bincount(flat, weights=(w * xc).ravel(), minlength=minlength)
    peak = np.zeros(minlength, dtype=F32)
    np.maximum.at(peak, labels.ravel(), w.ravel())
    return weight, zsum, ysum, xsum, peak


def on_bit_centroid_columns(df: pd.DataFrame, labels: np.ndarray, intensity: np.ndarray, on_sel: np.ndarray,
                            z_support: int = 7, epsilon: float = 1e-6, z_offset: float | None = None) -> pd.DataFrame:
    """PD:2701-2831.  ``df`` holds ``label`` and the decoded-space centroid ``z, y, x``;
    ``intensity`` is (bits, z, y, x)."""
    n_bits = intensity.shape[0]
    extra = {f"bit{b:02d}_{s}": np.full(len(df), np.nan) for b in range(1, n_bits + 1) for s in CENTROID_SUFFIXES}
    if len(df) == 0 or labels.max() < 1:
        return pd.DataFrame(extra, index=df.index)
    zs = min(int(z_support), labels.shape[0])
    if zs % 2 == 0:
        zs -= 1
    minlength = int(labels.max()) + 1
    lab = np.clip(df["label"].to_numpy(dtype=np.int64), 0, minlength - 1)
    fallback = df[["z", "y", "x"]].to_numpy(dtype=np.float64)
    area_by_label = np.bincount(labels.ravel(), minlength=minlength).astype(F32)
    for b in range(1, n_bits + 1):
        rows = np.flatnonzero(np.any(on_sel == b, axis=1))
        if rows.size == 0:
            continue
        weight, zsum, ysum, xsum, peak = plane_wise_centroid_statistics(labels, intensity[b - 1], zs, minlength)
        wsum = weight[lab[rows]].astype(np.float64)
        den = np.maximum(wsum, F32(epsilon))
        centers = np.column_stack((zsum[lab[rows]] / den, ysum[lab[rows]] / den, xsum[lab[rows]] / den))
        invalid = (~np.all(np.isfinite(centers), axis=1)) | (wsum <= 0)
        centers[invalid] = fallback[rows][invalid]
        if z_offset is not None:
            centers[:, 0] = float(z_offset) + centers[:, 0]
        area = area_by_label[lab[rows]].astype(np.float64)
        pk = peak[lab[rows]].astype(np.float64)
        missing = (~np.isfinite(wsum)) | (wsum <= 0)
        area[missing] = 0.0
        pk[~np.isfinite(pk)] = 0.0
        wsum[missing] = 0.0
        for k, sfx in enumerate(("center_z", "center_y", "center_x")):
            extra[f"bit{b:02d}_{sfx}"][rows] = centers[:, k]
        extra[f"bit{b:02d}_intensity_sum"][rows] = wsum
        extra[f"bit{b:02d}_intensity_peak"][rows] = pk
        extra[f"bit{b:02d}_voxel_count"][rows] = area
    return pd.DataFrame(extra, index=df.index)


def annotate_table(
    df: pd.DataFrame,
    decoded: np.ndarray,
    codebook_matrix: np.ndarray,
    gene_ids: Sequence[str],
    n_bits: int,
    tile_idx: int,
    transcript_threshold: float,
    spacing,
    origin,
    affine,
    camera_to_stage,
    z_offset: float | None = None,
    centroid_inputs: tuple | None = None,
) -> pd.DataFrame:
    """PD:3066-3177 -- codeword/gene annotation, coordinates, signal stats, transcript gate.

    ``centroid_inputs`` = (labels, intensity (bits,z,y,x), z_support, epsilon) switches on the
    optimiser's per-on-bit centroid columns (PD:3117-3125)."""
    df = df.copy()
    dec_flat = decoded.ravel()
    ids = dec_flat[df["first_voxel"].to_numpy(dtype=np.int64)].astype(np.int32)
    df["decoded_id"] = ids
    df = df[df["decoded_id"] >= 0].reset_index(drop=True)
    df["barcode_id"] = df["decoded_id"].astype(np.int32) + 1
    df["gene_id"] = [gene_ids[i] for i in df["decoded_id"].to_numpy(dtype=np.int32)]
    df["tile_idx"] = tile_idx
    cb_bool = np.asarray(codebook_matrix).astype(bool, copy=False)
    on0 = np.argsort(~cb_bool, axis=1)[:, :4].astype(np.int32)  # PD:3101, verbatim
    on_sel = (on0 + 1)[df["decoded_id"].to_numpy(dtype=np.int32)]
    for k in range(4):
        df[f"on_bit_{k + 1}"] = on_sel[:, k] if len(df) else np.zeros(0, dtype=np.int32)
    if centroid_inputs is not None:
        lab_img, inten, zs, eps = centroid_inputs
        df = pd.concat([df, on_bit_centroid_columns(df, lab_img, inten, on_sel, zs, eps, z_offset)], axis=1)
    if z_offset is not None:
        df["z"] = float(z_offset) + df["z"]  # PD:3126-3127
    df["tile_z"] = np.round(df["z"], 0).astype(int)
    df["tile_y"] = np.round(df["y"], 0).astype(int)
    df["tile_x"] = np.round(df["x"], 0).astype(int)
    pts = df[["z", "y", "x"]].to_numpy(dtype=np.float64, copy=True)
    for i in range(pts.shape[0]):  # PD:3134-3141
        pts[i, :] = warp_pixel(pts[i, :].copy(), spacing, origin, affine, camera_to_stage)
    df["global_z"] = np.round(pts[:, 0], 2)
    df["global_y"] = np.round(pts[:, 1], 2)
    df["global_x"] = np.round(pts[:, 2], 2)
    for i in range(1, n_bits + 1):
        df = df.rename(columns={f"intensity_mean-{i - 1}": f"bit{i:02d}_mean_intensity"})
    bit_cols = [f"bit{i:02d}_mean_intensity" for i in range(1, n_bits + 1)]
    bit_means = df[bit_cols].to_numpy(dtype=np.float64)
    total = bit_means.sum(axis=1)
    on_idx = df[[f"on_bit_{k}" for k in range(1, 5)]].to_numpy(dtype=np.int32) - 1
    sig = np.take_along_axis(bit_means, on_idx, axis=1).sum(axis=1) if len(df) else total
    df["signal_mean"] = sig / 4.0
    df["bkd_mean"] = (total - sig) / float(n_bits - 4)
    df["s-b_mean"] = df["signal_mean"] - df["bkd_mean"]
    df = df.drop(columns=["label", "decoded_id", "first_voxel"])
    df = df[df["distance_min"] <= transcript_threshold].reset_index(drop=True)
    return df


def extract_barcodes(
    decode_out: dict,
    intensity: np.ndarray,
    codebook_matrix: np.ndarray,
    gene_ids: Sequence[str],
    is_3d: bool,
    minimum_pixels,
    transcript_threshold: float,
    tile_idx: int = 0,
    spacing=(1.0, 1.0, 1.0),
    origin=(0.0, 0.0, 0.0),
    affine=None,
    camera_to_stage=None,
    z_offset=None,
    maximum_pixels: int = MAXIMUM_PIXELS,
    return_labels: bool = False,
    collect_centroids: tuple | None = None,
):
    """PD:2908-3201.  ``collect_centroids`` = (z_support, epsilon) adds the optimiser's per-on-bit
    centroid columns (only meaningful with the raw intensity image, PD:3117-3125)."""
    labels = label_decoded(decode_out["decoded"], is_3d)
    labels = filter_label_sizes(labels, minimum_pixels, maximum_pixels)
    tab = region_table(labels, decode_out["distance"], decode_out["magnitude"], intensity)
    n_bits = intensity.shape[0]
    df = annotate_table(
        tab,
        decode_out["decoded"],
        codebook_matrix,
        gene_ids,
        n_bits,
        tile_idx,
        transcript_threshold,
        np.asarray(spacing, dtype=F32),
        np.asarray(origin, dtype=F32),
        np.eye(4, dtype=F32) if affine is None else np.asarray(affine, dtype=F32),
        np.eye(4, dtype=F32) if camera_to_stage is None else np.asarray(camera_to_stage, dtype=F32),
        z_offset,
        None if collect_centroids is None else (labels, intensity, collect_centroids[0], collect_centroids[1]),
    )
    if return_labels:
        return df, labels
    return df


def decode_tile(
    readout: np.ndarray,
    predictor: np.ndarray | None,
    cb: dict,
    bkg,
    nrm,
    is_3d: bool = True,
    lowpass_sigma=DEFAULT_DECODE_LOWPASS_SIGMA,
    magnitude_threshold=DEFAULT_DECODE_MAGNITUDE_THRESHOLD,
    minimum_pixels=None,
    excluded=(),
    optimize_mode: bool = False,
    tile_idx: int = 0,
    bit_transforms_zyx_um=None,
    collect_centroids: tuple | None = None,
    bit_flows=None,
    **coords,
):
    """PD:4471-4579 -- one tile end to end; returns (table, images dict).

    ``bit_transforms_zyx_um``: optional per-bit physical 4x4 decode-time transforms (PD:1882-1889);
    ``bit_flows``: optional per-bit ``(flow, stride_zyx, box_start_xyz)`` SOFIMA fields (None = affine only)."""
    if minimum_pixels is None:
        minimum_pixels = DEFAULT_3D_MINIMUM_PIXELS if is_3d else DEFAULT_2D_MINIMUM_PIXELS
    stack = weight_readout(readout, predictor)
    if bit_transforms_zyx_um is not None:
        spacing_um = coords.get("spacing", (1.0, 1.0, 1.0))
        vols = []
        for b in range(stack.shape[0]):
            fl = None if bit_flows is None else bit_flows[b]
            if fl is None:
                vols.append(warp_to_reference(stack[b], bit_transforms_zyx_um[b], spacing_um))
            else:  # (flow (3,fz,fy,fx), stride_zyx, box_start_xyz)
                vols.append(warp_to_reference_with_flow(stack[b], bit_transforms_zyx_um[b], spacing_um, *fl))
        stack = np.stack(vols)
    if is_3d and stack.shape[1] < 2:
        raise ValueError("decode_mode='3d' requires at least two z planes after applying z_range.")
    stack = lowpass_stack(stack, lowpass_sigma, is_3d)
    unit = normalize_codebook(cb["matrix"][:, : stack.shape[0]])
    out = decode_pixels(
        stack, unit, bkg, nrm, cb["pixel_assignment_threshold"], magnitude_threshold, excluded
    )
    intensity = stack if optimize_mode else out["scaled"]  # PD:2935-2941
    df = extract_barcodes(
        out,
        intensity,
        cb["matrix"],
        cb["gene_ids"],
        is_3d,
        minimum_pixels,
        cb["transcript_distance_threshold"],
        tile_idx=tile_idx,
        collect_centroids=collect_centroids if optimize_mode else None,
        **coords,
    )
    out["image"] = stack
    return df, out


# --------------------------------------------------------------------------- normalisation
def global_normalization_vectors(
    tile_stacks: Sequence[np.ndarray],
    n_bits_total: int,
    is_3d: bool = True,
    lowpass_sigma=DEFAULT_DECODE_LOWPASS_SIGMA,
    z_slice: slice = slice(None),
    low_percentile_cut: float = 10.0,
    high_percentile_cut: float = 90.0,
    hot_pixel_threshold: float = 50000,
):
    """PD:981-1199 -- percentile-seeded background / normalisation vectors.

    ``tile_stacks[t]`` is the weighted (bits, z, y, x) float32 stack of tile t BEFORE
    hot-pixel replacement, z-crop and low-pass.
    """
    nrm = np.ones(n_bits_total, dtype=F32)
    bkg = np.zeros(n_bits_total, dtype=F32)
    for b in range(n_bits_total):
        imgs = []
        for st in tile_stacks:
            img = np.array(st[b], dtype=F32, copy=True)
            med = np.median(img[img.shape[0] // 2]).astype(F32)  # PD:1072-1074
            img[img > hot_pixel_threshold] = med
            img = img[z_slice]
            img = lowpass_image(img, lowpass_sigma, is_3d)
            imgs.append(np.asarray(img, dtype=F32))
        low = []
        for img in imgs:
            px = img.reshape(-1)
            if px.size == 0:
                continue
            cut = np.percentile(px, low_percentile_cut)
            low.append(px[px < cut].astype(F32))
        low = np.concatenate(low) if low else np.empty((0,), dtype=F32)
        bkg[b] = np.median(low) if low.size > 0 else 0
        high = []
        for img in imgs:
            cur = img - bkg[b]
            cur[cur < 0] = 0
            px = cur.reshape(-1)
            if px.size == 0:
                continue
            cut = np.percentile(px, high_percentile_cut)
            high.append(px[px > cut].astype(F32))
        high = np.concatenate(high) if high else np.empty((0,), dtype=F32)
        nrm[b] = np.median(high) if high.size > 0 else 1
    return nrm, bkg


def iterative_normalization_vectors(df: pd.DataFrame, n_bits: int):
    """PD:1263-1368 -- per-bit medians of on-bit / off-bit feature means, rounded to 0.1.

    Returns (normalization, background) float32, or None when the reference would keep
    the previous vectors (no non-blank transcripts, PD:1302-1311).
    """
    keep = ~df["gene_id"].astype("string").str.lower().str.startswith("blank", na=False)
    d = df[keep]
    bit_cols = [c for c in d.columns if c.startswith("bit") and c.endswith("_mean_intensity")]
    if d.empty or not bit_cols:
        return None
    inten, back = [], []
    for _i, row in d.iterrows():  # PD:1315-1333, verbatim structure
        sel = [f"bit{int(row[f'on_bit_{k}']):02d}_mean_intensity" for k in range(1, 5)]
        inten.append({c: (row[c] if c in sel else None) for c in bit_cols})
        back.append({c: (row[c] if c not in sel else None) for c in bit_cols})
    di = pd.DataFrame(inten)
    db = pd.DataFrame(back)
    di = di.reindex(sorted(di.columns), axis=1)
    db = db.reindex(sorted(db.columns), axis=1)
    nv = np.round(di.median(skipna=True).to_numpy(dtype=F32, copy=True), 1)
    bv = np.round(db.median(skipna=True).to_numpy(dtype=F32, copy=True), 1)
    nv = np.nan_to_num(nv, 1.0)
    nv = np.where(nv == 0.0, 1.0, nv)
    bv = np.nan_to_num(bv, 0.0)
    return nv.astype(F32), bv.astype(F32)


def optimize_normalization(
    tiles: Sequence[tuple[np.ndarray, np.ndarray | None]],
    cb: dict,
    n_iterations: int,
    is_3d: bool = True,
    lowpass_sigma=DEFAULT_DECODE_LOWPASS_SIGMA,
    magnitude_threshold=DEFAULT_DECODE_MAGNITUDE_THRESHOLD,
    minimum_pixels=None,
    excluded=(),
):
    """PD:4581-4757 -- global seed, then n iterations of decode -> medians.

    ``tiles`` = [(readout, predictor), ...] already restricted to the chosen subset.
    Returns dict(global=(nrm,bkg), iterative=(nrm,bkg), history=[...]).
    """
    n_bits = tiles[0][0].shape[0]
    stacks = [weight_readout(r, p) for r, p in tiles]
    g_nrm, g_bkg = global_normalization_vectors(stacks, n_bits, is_3d, lowpass_sigma)
    it_nrm = it_bkg = None
    history = []
    for it in range(n_iterations):
        if it == 0:
            nrm, bkg = g_nrm, g_bkg
        else:
            nrm, bkg = it_nrm, it_bkg
        dfs = []
        for t, (r, p) in enumerate(tiles):
            df, _ = decode_tile(
                r, p, cb, bkg, nrm, is_3d, lowpass_sigma, magnitude_threshold,
                minimum_pixels, excluded, optimize_mode=True, tile_idx=t,
            )
            dfs.append(df)
        pooled = pd.concat(dfs, ignore_index=True) if dfs else pd.DataFrame()
        res = iterative_normalization_vectors(pooled, n_bits) if len(pooled) else None
        if res is None:
            if it_nrm is None:  # PD:1272-1287
                it_nrm = np.round(g_nrm[:n_bits], 1).astype(F32)
                it_bkg = np.round(g_bkg[:n_bits], 1).astype(F32)
        else:
            it_nrm, it_bkg = res
        history.append((it_nrm.copy(), it_bkg.copy(), len(pooled)))
    return dict(global_=(g_nrm, g_bkg), iterative=(it_nrm, it_bkg), history=history)
