"""Synthetic floor plans and robot waypoint paths.

The reference ships one building scan (floorplans/32_data/floor_img.png, 2167x1608) with
hand-drawn waypoint sets (tight_paths/, some_overlap/, minimal_overlap/) made with its
matplotlib point selector.  This module generates equivalent *assets* procedurally — a grid
of rooms joined by doors, and per-robot room tours expressed as normalised waypoints in
[-1, 1]^2 — so the density experiments run with no data files; ``write_dataset`` lays them out
exactly like the reference's ``data_dir`` (``floor_img.png`` + ``<subdir>/<k>.npy``), and a real
``32_data`` directory can be used instead.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np
from PIL import Image


def make_floorplan(nx: int = 720, ny: int = 540, rooms: Tuple[int, int] = (4, 3), wall: int = 8,
                   door: int = 60, seed: int = 0):
    """uint8 image [ny, nx] (255 = wall, 0 = free) plus the room/door geometry."""
    rng = np.random.default_rng(seed)
    img = np.zeros((ny, nx), dtype=np.uint8)
    cols, rows = rooms
    xs = np.linspace(0, nx, cols + 1).astype(int)
    ys = np.linspace(0, ny, rows + 1).astype(int)
    doors = {}
    for c in range(1, cols):  # vertical interior walls with one door per room boundary
        x = xs[c]
        img[:, x - wall // 2: x + wall // 2] = 255
        for r in range(rows):
            cy = int(rng.uniform(ys[r] + door, ys[r + 1] - door))
            img[cy - door // 2: cy + door // 2, x - wall // 2: x + wall // 2] = 0
            doors[((c - 1, r), (c, r))] = (x, cy)
    for r in range(1, rows):
        y = ys[r]
        img[y - wall // 2: y + wall // 2, :] = 255
        for c in range(cols):
            cx = int(rng.uniform(xs[c] + door, xs[c + 1] - door))
            img[y - wall // 2: y + wall // 2, cx - door // 2: cx + door // 2] = 0
            doors[((c, r - 1), (c, r))] = (cx, y)
    for c in range(1, cols):  # re-open wall crossings blocked by horizontal walls? keep walls solid at crossings
        for r in range(1, rows):
            img[ys[r] - wall // 2: ys[r] + wall // 2, xs[c] - wall // 2: xs[c] + wall // 2] = 255
    # a few pillars so rooms are not empty boxes
    for c in range(cols):
        for r in range(rows):
            if rng.random() < 0.6:
                px = int(rng.uniform(xs[c] + 0.3 * (xs[c + 1] - xs[c]), xs[c] + 0.7 * (xs[c + 1] - xs[c])))
                py = int(rng.uniform(ys[r] + 0.3 * (ys[r + 1] - ys[r]), ys[r] + 0.7 * (ys[r + 1] - ys[r])))
                img[py - 10: py + 10, px - 10: px + 10] = 255
    centers = {(c, r): (0.5 * (xs[c] + xs[c + 1]), 0.5 * (ys[r] + ys[r + 1])) for c in range(cols) for r in range(rows)}
    # nudge room centres off pillars
    for key, (cx, cy) in centers.items():
        if img[int(cy), int(cx)] or img[int(cy) - 14: int(cy) + 14, int(cx) - 14: int(cx) + 14].any():
            centers[key] = (cx + 0.22 * (xs[1] - xs[0]), cy + 0.22 * (ys[1] - ys[0]))
    return img, {"xs": xs, "ys": ys, "doors": doors, "centers": centers, "rooms": rooms}


def _door_between(geo, a, b):
    d = geo["doors"].get((a, b)) or geo["doors"].get((b, a))
    return d


def _tour(geo, start, length, rng) -> List[Tuple[int, int]]:
    cols, rows = geo["rooms"]
    cur, out, prev = start, [start], None
    for _ in range(length - 1):
        nbrs = [(cur[0] + dx, cur[1] + dy) for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1))]
        nbrs = [n for n in nbrs if 0 <= n[0] < cols and 0 <= n[1] < rows]
        choices = [n for n in nbrs if n != prev] or nbrs
        nxt = choices[int(rng.integers(len(choices)))]
        out.append(nxt)
        prev, cur = cur, nxt
    return out


def make_waypoints(img: np.ndarray, geo, n_paths: int, rooms_per_path: int = 8, seed: int = 0) -> List[np.ndarray]:
    """``n_paths`` arrays ``[K, 2]`` of normalised waypoints (x, y in [-1, 1], image-centred)."""
    rng = np.random.default_rng(seed)
    ny, nx = img.shape
    cols, rows = geo["rooms"]
    from scipy.ndimage import binary_dilation
    from scipy.interpolate import interp1d

    blocked = binary_dilation(img > 127, iterations=12)   # keep a 12 px clearance from every wall

    def is_free(norm):
        i = np.arange(len(norm))
        ii = np.linspace(0, i.max(), 40 * i.max())
        x = interp1d(i, norm[:, 0], kind="cubic")(ii) * nx / 2 + nx / 2
        y = interp1d(i, norm[:, 1], kind="cubic")(ii) * ny / 2 + ny / 2
        xi = np.clip(np.round(x).astype(int), 0, nx - 1)
        yi = np.clip(np.round(y).astype(int), 0, ny - 1)
        return not blocked[yi, xi].any()

    paths = []
    attempts = 0
    while len(paths) < n_paths:
        attempts += 1
        if attempts > 200 * n_paths:
            raise RuntimeError("could not place collision-free robot paths; use fewer rooms_per_path")
        start = (int(rng.integers(cols)), int(rng.integers(rows)))
        tour = _tour(geo, start, rooms_per_path, rng)
        pts = [geo["centers"][tour[0]]]
        for a, b in zip(tour[:-1], tour[1:]):
            dx, dy = _door_between(geo, a, b)
            ca, cb = geo["centers"][a], geo["centers"][b]
            # approach, cross and leave the door along its normal so the cubic path stays clear of the wall
            if a[0] != b[0]:
                s = np.sign(cb[0] - ca[0])
                pts += [(dx - s * 40, dy), (dx, dy), (dx + s * 40, dy)]
            else:
                s = np.sign(cb[1] - ca[1])
                pts += [(dx, dy - s * 40), (dx, dy), (dx, dy + s * 40)]
            pts.append(cb)
        p = np.asarray(pts, dtype=np.float64)
        norm = np.stack([(p[:, 0] - nx / 2) / (nx / 2), (p[:, 1] - ny / 2) / (ny / 2)], axis=1)
        if is_free(norm):
            paths.append(norm)
    return paths


def write_dataset(data_dir: str, n_paths: int = 8, subdir: str = "tight_paths", seed: int = 0, **kw) -> str:
    """Create ``data_dir/floor_img.png`` and ``data_dir/<subdir>/<k>.npy`` (reference layout)."""
    os.makedirs(os.path.join(data_dir, subdir), exist_ok=True)
    img, geo = make_floorplan(seed=seed, **kw)
    Image.fromarray(img, mode="L").save(os.path.join(data_dir, "floor_img.png"))
    for k, wp in enumerate(make_waypoints(img, geo, n_paths, seed=seed + 1)):
        np.save(os.path.join(data_dir, subdir, f"{k + 1}.npy"), wp)
    return data_dir
