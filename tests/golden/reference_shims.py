"""Run the UNMODIFIED reference ``PixelDecoder`` on the CPU by standing in for its GPU wheels.

TEST INFRASTRUCTURE ONLY (used by ``make_reference_golden.py`` and the live-reference CPU test).

``/root/reference/src/merfish3danalysis/PixelDecoder.py`` imports CuPy, cupyx, cuVS, cuCIM,
scikit-image, rtree, roifile and shapely at module level (PD:109-123); none is installed in the
build image and there is no GPU.  The reference's own code for the hot path -- thresholds,
per-plane loop, gates, rounding, int16/float16 stores, size filters, table annotation, the
normalisation estimators, the optimiser loop -- is pure Python over those libraries' array API,
so it runs verbatim once the third-party calls are answered by their NumPy / SciPy
counterparts (BASELINE.json: "the reference decode math on NumPy/SciPy"):

    cupy                                 -> numpy (same array API; Device/Stream/pools are no-ops)
    cupyx.scipy.ndimage.gaussian_filter  -> scipy.ndimage.gaussian_filter
    cuvs.distance.pairwise_distance      -> float32 direct form sqrt(sum_k (x_k - c_k)^2)
    cucim.skimage.measure.label          -> scikit-image semantics (equal-value components,
                                            numbered in raster order of the first voxel) on
                                            scipy.ndimage.label
    cucim ... remove_small_objects       -> scikit-image >= 0.26 ``max_size`` semantics (<= removed)
    cucim / skimage regionprops_table    -> scikit-image's published formulas (area, centroid,
                                            intensity_mean/min, inertia_tensor_eigvals)
    rtree / roifile / shapely / skimage.draw -> import-only stubs (cell assignment is off-path)

What this pins: everything the reference itself wrote.  What it cannot pin: the exact rounding of
the replaced third-party kernels (cuVS expanded-form distance, cupyx float32 accumulation) --
stated in DESIGN.md section 3.
"""

from __future__ import annotations

import contextlib
import sys
import types
from pathlib import Path

import numpy as np
import scipy.ndimage as ndi

REFERENCE_SRC = Path("/root/reference/src")


def reference_available() -> bool:
    return (REFERENCE_SRC / "merfish3danalysis" / "PixelDecoder.py").exists()


# ----------------------------------------------------------------------------- cupy -> numpy
class _Pool:
    def free_all_blocks(self):
        return None

    def used_bytes(self):
        return 0

    def total_bytes(self):
        return 0


class _NullStream:
    def synchronize(self):
        return None


class _Stream:
    null = _NullStream()


class _Device(contextlib.AbstractContextManager):
    def __init__(self, *_a, **_k):
        pass

    def __exit__(self, *exc):
        return False

    def use(self):
        return None

    def synchronize(self):
        return None


class _DeviceScalar(np.ndarray):
    """0-d result of a reduction: CuPy returns a device array with ``.get()`` (PD:2756)."""

    def get(self):
        return np.asarray(self)[()]


def _reduction(fn):
    def wrapped(*a, **k):
        r = fn(*a, **k)
        return np.asarray(r).view(_DeviceScalar) if np.ndim(r) == 0 else r

    return wrapped


def _make_cupy() -> types.ModuleType:
    cp = types.ModuleType("cupy")
    cp.max = _reduction(np.max)

    def __getattr__(name):  # everything else is the NumPy function of the same name
        return getattr(np, name)

    cp.__getattr__ = __getattr__
    cp.ndarray = np.ndarray
    cp.asnumpy = lambda a, *args, **kw: np.asarray(a)
    cp.get_array_module = lambda *a: np
    cp.get_default_memory_pool = lambda: _Pool()
    cp.get_default_pinned_memory_pool = lambda: _Pool()
    cuda = types.ModuleType("cupy.cuda")
    cuda.Device = _Device
    cuda.Stream = _Stream
    runtime = types.ModuleType("cupy.cuda.runtime")
    runtime.getDeviceCount = lambda: 1
    cuda.runtime = runtime
    cp.cuda = cuda
    return cp


# ----------------------------------------------------------------------------- cuVS
def pairwise_distance(X, Y, out=None, metric="euclidean"):
    assert metric == "euclidean"
    X = np.asarray(X, dtype=np.float32)
    Y = np.asarray(Y, dtype=np.float32)
    res = np.empty((X.shape[0], Y.shape[0]), dtype=np.float32) if out is None else out
    for k in range(Y.shape[0]):
        acc = np.zeros(X.shape[0], dtype=np.float32)
        for b in range(X.shape[1]):
            diff = X[:, b] - Y[k, b]
            acc += diff * diff
        res[:, k] = np.sqrt(acc)
    return res


# ----------------------------------------------------------------------------- scikit-image family
def label(image, background=None, return_num=False, connectivity=None):
    """skimage.measure.label: neighbours are connected iff they hold the same value."""
    image = np.asarray(image)
    nd = image.ndim
    connectivity = nd if connectivity is None else int(connectivity)
    structure = ndi.generate_binary_structure(nd, connectivity)
    background = 0 if background is None else background
    first = []  # (first raster index, value, component id within value)
    per_value = {}
    for v in np.unique(image):
        if v == background:
            continue
        lab, n = ndi.label(image == v, structure=structure)
        per_value[v] = lab
        if n:
            flat = lab.ravel()
            pos = np.flatnonzero(flat)
            firsts = np.full(n + 1, flat.size, dtype=np.int64)
            np.minimum.at(firsts, flat[pos], pos)
            first.extend((int(firsts[c]), v, c) for c in range(1, n + 1))
    first.sort()
    out = np.zeros(image.shape, dtype=np.int64)
    remap = {}
    for new, (_p, v, c) in enumerate(first, start=1):
        remap.setdefault(v, {})[c] = new
    for v, lab in per_value.items():
        table = np.zeros(int(lab.max()) + 1, dtype=np.int64)
        for c, new in remap.get(v, {}).items():
            table[c] = new
        out += table[lab]
    if return_num:
        return out, len(first)
    return out


def remove_small_objects(ar, min_size=64, connectivity=1, *, max_size=None, out=None):
    """Labelled-array branch of skimage.morphology.remove_small_objects.

    scikit-image >= 0.26: ``max_size`` removes objects with area <= max_size (the reference
    passes ``max_size=max(int(minimum_pixels) - 1, 0)``, PD:2987-2989)."""
    ar = np.asarray(ar)
    out = ar.copy() if out is None else out
    counts = np.bincount(ar.ravel())
    if max_size is None:
        too_small = counts < min_size
    else:
        too_small = counts <= max_size
    mask = too_small[ar]
    out[mask] = 0
    return out


def _regionprops_rows(label_image, intensity_image, properties):
    label_image = np.asarray(label_image)
    nd = label_image.ndim
    objects = ndi.find_objects(label_image)
    cols: dict[str, list] = {}

    def put(name, value):
        cols.setdefault(name, []).append(value)

    for i, sl in enumerate(objects):
        if sl is None:
            continue
        lab = i + 1
        region = label_image[sl] == lab
        coords = np.argwhere(region) + np.array([s.start for s in sl])  # int64, raster order
        for prop in properties:
            if prop == "label":
                put("label", lab)
            elif prop == "area":
                put("area", float(np.sum(region)))
            elif prop == "centroid":
                c = coords.mean(axis=0)
                for d in range(nd):
                    put(f"centroid-{d}", c[d])
            elif prop in ("intensity_mean", "intensity_min"):
                img = np.asarray(intensity_image)[sl]
                vals = img[region]  # (n,) or (n, channels), raster order
                red = np.mean(vals, axis=0) if prop == "intensity_mean" else np.min(vals, axis=0)
                if vals.ndim == 1:
                    put(prop, red)
                else:
                    for ch in range(vals.shape[1]):
                        put(f"{prop}-{ch}", red[ch])
            elif prop == "inertia_tensor_eigvals":
                local = np.argwhere(region).astype(np.float64)
                mu0 = float(local.shape[0])
                cen = local.mean(axis=0)
                d = local - cen
                mu2 = d.T @ d  # second central moments mu_{ij}
                diag = np.diag(mu2)
                tensor = -mu2 / mu0
                for a in range(nd):
                    tensor[a, a] = (diag.sum() - diag[a]) / mu0
                ev = np.linalg.eigvalsh(tensor)
                ev = np.clip(ev, 0, None, out=ev)
                ev = sorted(ev, reverse=True)
                for d_ in range(nd):
                    put(f"inertia_tensor_eigvals-{d_}", ev[d_])
            else:
                raise NotImplementedError(prop)
    if not cols:
        names = []
        for prop in properties:
            if prop == "centroid":
                names += [f"centroid-{d}" for d in range(nd)]
            elif prop == "inertia_tensor_eigvals":
                names += [f"inertia_tensor_eigvals-{d}" for d in range(nd)]
            elif prop in ("intensity_mean", "intensity_min") and intensity_image is not None \
                    and np.asarray(intensity_image).ndim > nd:
                names += [f"{prop}-{c}" for c in range(np.asarray(intensity_image).shape[-1])]
            else:
                names.append(prop)
        return {n: np.array([]) for n in names}
    return {k: np.asarray(v) for k, v in cols.items()}


def regionprops_table(label_image, intensity_image=None, properties=("label", "bbox"), **_kw):
    return _regionprops_rows(label_image, intensity_image, list(properties))


# ----------------------------------------------------------------------------- installation
def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("off-path dependency stub")


def install() -> None:
    """Insert the stand-in modules and put the reference's ``src`` on ``sys.path``."""
    if "cupy" in sys.modules and getattr(sys.modules["cupy"], "_m3d_shim", False):
        return
    cp = _make_cupy()
    cp._m3d_shim = True
    sys.modules["cupy"] = cp
    sys.modules["cupy.cuda"] = cp.cuda
    sys.modules["cupy.cuda.runtime"] = cp.cuda.runtime
    _stub("cupyx")
    _stub("cupyx.scipy")
    _stub("cupyx.scipy.ndimage", gaussian_filter=ndi.gaussian_filter, affine_transform=ndi.affine_transform,
          map_coordinates=ndi.map_coordinates)
    sys.modules["cupyx.scipy"].ndimage = sys.modules["cupyx.scipy.ndimage"]
    sys.modules["cupyx"].scipy = sys.modules["cupyx.scipy"]
    _stub("cuvs")
    _stub("cuvs.distance", pairwise_distance=pairwise_distance)
    _stub("cucim")
    _stub("cucim.skimage")
    _stub("cucim.skimage.measure", label=label, regionprops_table=regionprops_table)
    _stub("cucim.skimage.morphology", remove_small_objects=remove_small_objects)
    _stub("skimage")
    _stub("skimage.measure", label=label, regionprops_table=regionprops_table)
    _stub("skimage.draw", polygon=_Unavailable)
    _stub("skimage.morphology", remove_small_objects=remove_small_objects)
    _stub("rtree", index=types.SimpleNamespace(Index=_Unavailable))
    _stub("roifile", roiread=_Unavailable)
    _stub("shapely")
    _stub("shapely.geometry", Point=_Unavailable, Polygon=_Unavailable)
    # the real datastore needs zarr/tensorstore; the decoder only uses the type as an annotation
    if "merfish3danalysis.qi2labDataStore" not in sys.modules:
        try:
            import merfish3danalysis.qi2labDataStore  # noqa: F401
        except Exception:
            for k in [k for k in sys.modules if k.startswith("merfish3danalysis")]:
                del sys.modules[k]
            if str(REFERENCE_SRC) not in sys.path:
                sys.path.insert(0, str(REFERENCE_SRC))
            import merfish3danalysis  # noqa: F401  (lazy package: imports nothing heavy)

            _stub("merfish3danalysis.qi2labDataStore", qi2labDataStore=object)


def load_reference_pixeldecoder():
    """The reference's own ``PixelDecoder`` class, executed from /root/reference (unmodified)."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    if str(REFERENCE_SRC) not in sys.path:
        sys.path.insert(0, str(REFERENCE_SRC))
    install()
    import importlib

    mod = importlib.import_module("merfish3danalysis.PixelDecoder")
    return mod.PixelDecoder


@contextlib.contextmanager
def pandas2_semantics():
    """The reference writes into ``DataFrame.to_numpy()`` results (PD:3133-3141), which pandas < 3
    returns writable; pandas 3 (copy-on-write, installed here) hands back read-only views.  Inside
    this context ``to_numpy`` returns a writable copy in that case -- values are unchanged."""
    import pandas as pd

    orig = pd.DataFrame.to_numpy

    def to_numpy(self, *a, **k):
        out = orig(self, *a, **k)
        if not out.flags.writeable:
            out = out.copy()
        return out

    pd.DataFrame.to_numpy = to_numpy
    try:
        yield
    finally:
        pd.DataFrame.to_numpy = orig
