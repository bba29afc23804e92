"""Z-slab sharding of ONE volume across GPUs (or across sequential passes on one GPU).

No reference counterpart (the reference never splits a tile, PD:2561); the correctness contract
is "identical to the unsharded volume" (SURVEY.md section 8e): decoded image, component ids in
raster order of their first voxel, areas after BOTH size filters applied to the merged
components, and every feature column.

Per slab (one rank, or one pass):
  1. load planes [z0, z1) (+ a low-pass halo of radius(sigma_z) planes, discarded after the
     filter), decode + label with the minimum-size filter disabled (a component cut by an
     interface may be completed by the neighbour), features for the slab's components;
  2. the slab's LAST decoded / label plane goes to the next slab (NCCL send/recv), which emits the
     cross-interface equivalences with ``m3d_interface_pairs`` (26-neighbour, equal value);
  3. every rank resolves the (small) equivalence list with the same host union-find, applies the
     size filters to the merged areas, and drops components joined to an oversized one;
  4. components that do cross an interface are re-assembled from per-voxel records (float16 scaled
     values / raw values, magnitude, distance -- produced by the dense decode kernel on the
     gathered voxels, i.e. the same device arithmetic) and reduced with NumPy in raster order, which
     is the reference's own reduction (scikit-image ``intensity_mean`` = ``np.mean(axis=0)``).
     Only those boundary components (O(interface area)) take this path; all others keep the
     rows the regionprops kernel wrote.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from ._capi import M3D_TABLE_FIXED_COLS


def split_z(n_planes: int, n_slabs: int) -> list[tuple[int, int]]:
    """Contiguous, balanced z ranges (empty slabs dropped)."""
    edges = np.linspace(0, n_planes, n_slabs + 1).round().astype(int)
    return [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]


def lowpass_z_radius(sigma) -> int:
    """Planes of halo the 3-D low-pass needs on each side (scipy: int(4 * sigma + 0.5))."""
    return int(4.0 * float(sigma[0]) + 0.5)


@dataclass
class SlabResult:
    z0: int
    z1: int
    shape_yx: tuple
    table: np.ndarray  # (n, 14 + bits) float64, first_voxel / centroid in SLAB coordinates
    pairs: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), dtype=np.int32))  # (id_here+1, id_prev+1 | -1)
    poisoned_here: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))  # ids joined to an oversized neighbour


def decode_slab(ctx, stack, mode2d: bool, maximum_pixels: int, optimize_mode: bool):
    """Phase 1 on one device-resident slab.  Returns (decoded, labels, table)."""
    import torch

    shape = tuple(stack.shape[1:])
    decoded = torch.empty(shape, dtype=torch.int16, device=stack.device)
    labels = torch.zeros(shape, dtype=torch.int32, device=stack.device)
    n = ctx.decode_label(stack, decoded, mode2d, 1.0, int(maximum_pixels), labels=labels)
    table = ctx.features(stack, decoded, optimize_mode, n).cpu().numpy()
    return decoded, labels, table


def interface(ctx, prev_planes, decoded, labels):
    """Phase 2: equivalences between this slab's first plane and the previous slab's last plane.

    Returns (pairs (id_here, id_prev) 0-based, poisoned ids here, poisoned ids in the previous slab):
    'poisoned' = joined to a component the other slab dropped for exceeding maximum_pixels."""
    dec_lo, lab_lo = prev_planes
    raw = ctx.interface_pairs(dec_lo, lab_lo, decoded[0].contiguous(), labels[0].contiguous())
    hi, lo = raw[:, 0].astype(np.int64), raw[:, 1].astype(np.int64)
    ok = (hi > 0) & (lo > 0)
    pairs = np.stack([hi[ok] - 1, lo[ok] - 1], axis=1) if ok.any() else np.zeros((0, 2), dtype=np.int64)
    poisoned_here = np.unique(hi[(hi > 0) & (lo < 0)] - 1)
    poisoned_prev = np.unique(lo[(lo > 0) & (hi < 0)] - 1)
    return pairs, poisoned_here, poisoned_prev


def resolve(areas, pairs, poisoned, minimum_pixels: float, maximum_pixels: int):
    """Phase 3 (pure, identical on every rank).

    areas[r]    : (n_r,) component areas of slab r (by slab-local id)
    pairs[r]    : (m, 2) (id in slab r, id in slab r-1) equivalences (empty for r = 0)
    poisoned[r] : ids in slab r joined to an oversized component of a neighbour
    Returns (keep_local[r] bool (n_r,) for components that do NOT cross an interface,
             groups: list of lists of (r, id) for crossing components that survive)."""
    n_slabs = len(areas)
    offs = np.concatenate([[0], np.cumsum([len(a) for a in areas])]).astype(np.int64)
    total = int(offs[-1])
    parent = np.arange(total, dtype=np.int64)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    crossing = np.zeros(total, dtype=bool)
    for r in range(1, n_slabs):
        for a, b in np.asarray(pairs[r], dtype=np.int64).reshape(-1, 2):
            ka, kb = int(offs[r] + a), int(offs[r - 1] + b)
            crossing[ka] = crossing[kb] = True
            ra, rb = find(ka), find(kb)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
    bad = np.zeros(total, dtype=bool)
    for r in range(n_slabs):
        p = np.asarray(poisoned[r], dtype=np.int64)
        bad[offs[r] + p] = True
        crossing[offs[r] + p] = True
    max_size = max(int(minimum_pixels) - 1, 0)  # PD:2987: remove area <= max(int(min)-1, 0)
    all_area = np.concatenate(areas) if total else np.zeros(0)
    keep_local = []
    for r in range(n_slabs):
        a = np.asarray(areas[r])
        k = (a > max_size) & (a <= maximum_pixels) & ~crossing[offs[r] : offs[r + 1]]
        keep_local.append(k)
    groups = {}
    for k in np.flatnonzero(crossing):
        groups.setdefault(find(int(k)), []).append(int(k))
    out = []
    for members in groups.values():
        area = float(all_area[members].sum())
        if bad[members].any() or area <= max_size or area > maximum_pixels:
            continue
        slab_of = [int(np.searchsorted(offs, m, side="right") - 1) for m in members]
        out.append(sorted((r, int(m - offs[r])) for r, m in zip(slab_of, members)))
    return keep_local, out


def slab_records(ctx, stack, decoded, labels, ids, optimize_mode: bool):
    """Phase 4a: per-voxel records of the given slab-local component ids.

    Values come from the dense decode kernel run on the gathered voxels: float16 scaled values,
    magnitude and distance exactly as the reference's result images hold them (PD:2621-2632).
    Returns dict id -> dict(lin (n,), vals (n, bits) float16|float32, mag (n,) f16, dist (n,) f16, dec)."""
    import torch

    if len(ids) == 0:
        return {}
    want = torch.as_tensor(np.asarray(ids, dtype=np.int32) + 1, device=labels.device)
    flat = labels.reshape(-1)
    lin = torch.nonzero(torch.isin(flat, want)).reshape(-1)  # ascending = raster order
    lab = flat[lin]
    n_bits = stack.shape[0]
    src = stack.view(torch.int16) if stack.dtype == torch.uint16 else stack  # torch cannot index uint16
    mini = src.reshape(n_bits, -1)[:, lin].contiguous().reshape(n_bits, 1, 1, -1)
    if stack.dtype == torch.uint16:
        mini = mini.view(torch.uint16)
    dec = torch.empty((1, 1, lin.numel()), dtype=torch.int16, device=stack.device)
    mag = torch.empty((1, 1, lin.numel()), dtype=torch.float16, device=stack.device)
    dist = torch.empty_like(mag)
    scaled = torch.empty(tuple(mini.shape), dtype=torch.float16, device=stack.device)
    ctx.decode(mini, dec, mag, dist, scaled)
    vals = mini.to(torch.float32) if optimize_mode else scaled
    vals = vals.reshape(n_bits, -1).T.contiguous().cpu().numpy()
    lin_h, lab_h = lin.cpu().numpy(), lab.cpu().numpy()
    mag_h, dist_h = mag.reshape(-1).cpu().numpy(), dist.reshape(-1).cpu().numpy()
    dec_h = decoded.reshape(-1)[lin].cpu().numpy()
    out = {}
    order = np.argsort(lab_h, kind="stable")
    bounds = np.flatnonzero(np.diff(lab_h[order])) + 1
    for chunk in np.split(order, bounds):
        cid = int(lab_h[chunk[0]]) - 1
        out[cid] = dict(lin=lin_h[chunk], vals=vals[chunk], mag=mag_h[chunk], dist=dist_h[chunk], dec=int(dec_h[chunk[0]]))
    return out


def merged_row(parts, shape_yx, n_bits: int, optimize_mode: bool) -> np.ndarray:
    """Phase 4b: one feature-table row from the records of a cross-slab component.

    ``parts`` = [(z0 of the slab, record dict), ...] in ascending slab order, so concatenation is
    the raster order of the whole volume.  Column layout = m3d_features (include/m3d_b200.h)."""
    Y, X = shape_yx
    plane = Y * X
    lin = np.concatenate([rec["lin"].astype(np.int64) + z0 * plane for z0, rec in parts])
    vals = np.concatenate([rec["vals"] for _z, rec in parts], axis=0)
    mag = np.concatenate([rec["mag"] for _z, rec in parts])
    dist = np.concatenate([rec["dist"] for _z, rec in parts])
    n = lin.size
    z, rem = np.divmod(lin, plane)
    y, x = np.divmod(rem, X)
    row = np.zeros(M3D_TABLE_FIXED_COLS + n_bits, dtype=np.float64)
    row[0] = float(lin[0])
    row[1] = float(n)
    row[2] = float(parts[0][1]["dec"])
    sz, sy, sx = int(z.sum()), int(y.sum()), int(x.sum())
    row[3], row[4], row[5] = sz / n, sy / n, sx / n
    # central second moments exactly like the device kernel: integer sums relative to the first voxel
    dz, dy, dx = z - z[0], y - y[0], x - x[0]
    mz, my, mx = float(int(dz.sum())), float(int(dy.sum())), float(int(dx.sum()))
    dn = float(n)
    row[6] = float(int((dz * dz).sum())) - mz * mz / dn
    row[7] = float(int((dy * dy).sum())) - my * my / dn
    row[8] = float(int((dx * dx).sum())) - mx * mx / dn
    row[9] = float(int((dz * dy).sum())) - mz * my / dn
    row[10] = float(int((dz * dx).sum())) - mz * mx / dn
    row[11] = float(int((dy * dx).sum())) - my * mx / dn
    row[12] = float(np.min(dist.astype(np.float32)))  # PD:2991-2995 min of float32(distance image)
    row[13] = float(np.mean(mag, axis=0))  # float16 in, float16 out (scikit-image intensity_mean)
    row[M3D_TABLE_FIXED_COLS :] = np.mean(np.ascontiguousarray(vals), axis=0).astype(np.float64)
    return row


def assemble(slabs: list[SlabResult], keep_local, merged_rows, n_bits: int) -> np.ndarray:
    """Final table: surviving slab-local rows moved to volume coordinates + merged rows, in
    canonical order (ascending first voxel)."""
    parts = []
    for s, keep in zip(slabs, keep_local):
        if s.table.shape[0] == 0:
            continue
        t = s.table[keep].copy()
        plane = s.shape_yx[0] * s.shape_yx[1]
        t[:, 0] += float(s.z0) * plane
        # centroid z = (sum z_local + n * z0) / n from the exact integer sum, so that the value is
        # the one the unsharded kernel computes (adding z0 to the rounded mean could differ by an ulp)
        n = t[:, 1]
        t[:, 3] = (np.rint(t[:, 3] * n) + n * float(s.z0)) / n
        parts.append(t)
    if merged_rows:
        parts.append(np.stack(merged_rows))
    if not parts:
        return np.zeros((0, M3D_TABLE_FIXED_COLS + n_bits), dtype=np.float64)
    tab = np.concatenate(parts, axis=0)
    return tab[np.argsort(tab[:, 0], kind="stable")]


# ---------------------------------------------------------------------------------------------- exchange
# What the ranks exchange travels as flat float64 vectors through tensor collectives (counts first, then one
# padded buffer per rank -- the same scheme as PixelDecoder._gather_tables), never as pickled Python objects:
# every number below is an integer < 2^53, a float16 / float32 value or a float64 table entry, so the float64
# container is exact.

def pack_summary(summary: dict) -> np.ndarray:
    """``{slab: dict(z0, z1, shape_yx, areas, pairs, poisoned_here, poisoned_prev)}`` -> float64 vector."""
    parts = [np.asarray([len(summary)], dtype=np.float64)]
    for r in sorted(summary):
        s_ = summary[r]
        areas = np.asarray(s_["areas"], dtype=np.float64).reshape(-1)
        pairs = np.asarray(s_["pairs"], dtype=np.float64).reshape(-1)
        ph = np.asarray(s_["poisoned_here"], dtype=np.float64).reshape(-1)
        pp = np.asarray(s_["poisoned_prev"], dtype=np.float64).reshape(-1)
        head = [r, s_["z0"], s_["z1"], s_["shape_yx"][0], s_["shape_yx"][1], areas.size, pairs.size // 2, ph.size, pp.size]
        parts += [np.asarray(head, dtype=np.float64), areas, pairs, ph, pp]
    return np.concatenate(parts)


def unpack_summary(vec: np.ndarray) -> dict:
    vec = np.asarray(vec, dtype=np.float64)
    out, at = {}, 1
    for _ in range(int(vec[0])):
        r, z0, z1, sy, sx, na, npair, nph, npp = (int(v) for v in vec[at : at + 9])
        at += 9
        areas = vec[at : at + na].copy()
        at += na
        pairs = vec[at : at + 2 * npair].astype(np.int64).reshape(-1, 2)
        at += 2 * npair
        ph = vec[at : at + nph].astype(np.int64)
        at += nph
        pp = vec[at : at + npp].astype(np.int64)
        at += npp
        out[r] = dict(z0=z0, z1=z1, shape_yx=(sy, sx), areas=areas, pairs=pairs, poisoned_here=ph, poisoned_prev=pp)
    return out


def pack_records(records: dict, local_tabs: dict, n_bits: int) -> np.ndarray:
    """Per-voxel records of crossing components ``{(slab, id): dict(lin, vals, mag, dist, dec)}`` and the kept
    slab-local feature rows ``{slab: (n, 14 + bits) float64}`` -> float64 vector."""
    parts = [np.asarray([len(records), len(local_tabs), n_bits], dtype=np.float64)]
    for (r, cid) in sorted(records):
        rec = records[(r, cid)]
        n = int(rec["lin"].size)
        kind = 1 if np.asarray(rec["vals"]).dtype == np.float32 else 0  # 0: float16 scaled image, 1: float32 raw
        parts += [np.asarray([r, cid, n, rec["dec"], kind], dtype=np.float64),
                  np.asarray(rec["lin"], dtype=np.float64).reshape(-1),
                  np.asarray(rec["mag"], dtype=np.float64).reshape(-1),
                  np.asarray(rec["dist"], dtype=np.float64).reshape(-1),
                  np.asarray(rec["vals"], dtype=np.float64).reshape(-1)]
    for r in sorted(local_tabs):
        t = np.ascontiguousarray(local_tabs[r], dtype=np.float64)
        parts += [np.asarray([r, t.shape[0]], dtype=np.float64), t.reshape(-1)]
    return np.concatenate(parts)


def unpack_records(vec: np.ndarray):
    vec = np.asarray(vec, dtype=np.float64)
    n_rec, n_tab, n_bits = (int(v) for v in vec[:3])
    at = 3
    records, tabs = {}, {}
    for _ in range(n_rec):
        r, cid, n, dec, kind = (int(v) for v in vec[at : at + 5])
        at += 5
        lin = vec[at : at + n].astype(np.int64)
        at += n
        mag = vec[at : at + n].astype(np.float16)
        at += n
        dist = vec[at : at + n].astype(np.float16)
        at += n
        vals = vec[at : at + n * n_bits].astype(np.float32 if kind else np.float16).reshape(n, n_bits)
        at += n * n_bits
        records[(r, cid)] = dict(lin=lin, vals=vals, mag=mag, dist=dist, dec=dec)
    width = M3D_TABLE_FIXED_COLS + n_bits
    for _ in range(n_tab):
        r, n = int(vec[at]), int(vec[at + 1])
        at += 2
        tabs[r] = vec[at : at + n * width].reshape(n, width).copy()
        at += n * width
    return records, tabs


def _collective_device(dist):
    import torch

    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_gather_vectors(dist, vec: np.ndarray) -> list[np.ndarray]:
    """Variable-length float64 vectors of every rank, on every rank: sizes first, then one padded all_gather."""
    import torch

    dev = _collective_device(dist)
    world = dist.get_world_size()
    n = torch.tensor([int(vec.size)], dtype=torch.int64, device=dev)
    sizes = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(t.item()) for t in sizes]
    n_max = max(max(sizes), 1)
    send = torch.zeros(n_max, dtype=torch.float64, device=dev)
    if vec.size:
        send[: vec.size] = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64)).to(dev)
    parts = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(parts, send)
    return [p[:c].cpu().numpy() for p, c in zip(parts, sizes)]


def gather_vectors(dist, vec: np.ndarray, dst: int = 0):
    """Variable-length float64 vectors of every rank on rank ``dst`` (None elsewhere)."""
    import torch

    dev = _collective_device(dist)
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([int(vec.size)], dtype=torch.int64, device=dev)
    sizes = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(t.item()) for t in sizes]
    n_max = max(max(sizes), 1)
    send = torch.zeros(n_max, dtype=torch.float64, device=dev)
    if vec.size:
        send[: vec.size] = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64)).to(dev)
    parts = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, parts, dst=dst)
    if rank != dst:
        return None
    return [p[:c].cpu().numpy() for p, c in zip(parts, sizes)]
