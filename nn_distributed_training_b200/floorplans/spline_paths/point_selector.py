"""Waypoint path authoring tool (reference: floorplans/spline_paths/point_selector.py:1-274 — an interactive
matplotlib polygon editor over the floor plan: drag vertices, ``i`` inserts a vertex on the nearest edge, ``d`` deletes
the vertex under the cursor, ``t`` toggles the vertex markers, the cubic-spline trajectory is redrawn live, and the
closed vertex loop is saved as the next ``<k>.npy`` of a waypoint directory, :262-270).

Two layers here:

* ``WaypointPath`` — the editing model, no GUI dependency (unit-tested on a CPU box): a CLOSED loop of vertices in
  normalised ``[-1, 1]^2`` floor-plan coordinates (first == last, as the reference stores them), hit testing with a pixel
  tolerance, move (the duplicated end point follows), insert on the nearest segment, delete, spline preview through
  ``floorplans.lidar.interpolate_waypoints``, free-space check against the floor image, numbered save.
* ``PathEditor`` — the matplotlib front end (optional dependency) wiring mouse / key events to the model, with the same
  key bindings as the reference tool.

    python -m nn_distributed_training_b200.floorplans.spline_paths.point_selector <floor_img.png> <out.npy | out_dir/>
        [--points "x0,y0;x1,y1;..."]   pixel coordinates, non-interactive (builds, checks and saves the path)
        [--load path.npy]               start the editor from an existing waypoint file
        [--spline-res 30]
"""
from __future__ import annotations

import glob
import os
import sys
from typing import Callable, Optional, Tuple

import numpy as np
from PIL import Image

from ..lidar import interpolate_waypoints


def normalise(points_px: np.ndarray, nx: int, ny: int) -> np.ndarray:
    return np.stack([(points_px[:, 0] - nx / 2) / (nx / 2), (points_px[:, 1] - ny / 2) / (ny / 2)], axis=1)


def path_is_free(img: np.ndarray, waypoints: np.ndarray, spline_res: int = 30) -> bool:
    ny, nx = img.shape
    traj = interpolate_waypoints(waypoints[:, 0], waypoints[:, 1], spline_res)
    x = np.clip(np.round(traj[:, 0] * nx / 2 + nx / 2).astype(int), 0, nx - 1)
    y = np.clip(np.round(traj[:, 1] * ny / 2 + ny / 2).astype(int), 0, ny - 1)
    return bool((img[y, x] < 0.5).all())


def point_segment_distance(p, a, b) -> float:
    """Euclidean distance from point ``p`` to the segment ``a-b``."""
    p, a, b = (np.asarray(v, dtype=np.float64) for v in (p, a, b))
    ab = b - a
    den = float(ab @ ab)
    t = 0.0 if den == 0.0 else float(np.clip((p - a) @ ab / den, 0.0, 1.0))
    return float(np.linalg.norm(p - (a + t * ab)))


class WaypointPath:
    """Closed waypoint loop ``xy[0] == xy[-1]`` in normalised coordinates.  ``to_pixels`` maps a data-space point to the
    space in which the pick tolerance ``epsilon`` is measured (display pixels in the GUI; identity in tests)."""

    def __init__(self, xy: np.ndarray, epsilon: float = 5.0, to_pixels: Optional[Callable[[np.ndarray], np.ndarray]] = None):
        xy = np.asarray(xy, dtype=np.float64).reshape(-1, 2)
        if len(xy) < 3:
            raise ValueError("a path needs at least 3 distinct vertices")
        if not np.allclose(xy[0], xy[-1]):
            xy = np.vstack([xy, xy[:1]])
        self.xy = xy.copy()
        self.epsilon = float(epsilon)
        self.to_pixels = to_pixels or (lambda a: np.asarray(a, dtype=np.float64))

    @classmethod
    def circle(cls, radius: float = 0.2, n: int = 3, **kw) -> "WaypointPath":
        """The reference tool's start shape: a small loop around the origin."""
        th = np.linspace(0.0, 2 * np.pi, n + 1)
        return cls(np.column_stack([radius * np.cos(th), radius * np.sin(th)]), **kw)

    # ---- queries ---------------------------------------------------------------------------------------------------
    @property
    def n_vertices(self) -> int:
        return len(self.xy) - 1                      # distinct vertices (the last row repeats the first)

    def hit_test(self, point) -> Optional[int]:
        """Index of the vertex within ``epsilon`` (pixel space) of ``point``, nearest first; ``None`` if none is."""
        d = np.linalg.norm(self.to_pixels(self.xy) - self.to_pixels(np.asarray(point, dtype=np.float64)), axis=1)
        i = int(np.argmin(d))
        return i if d[i] <= self.epsilon else None

    def spline(self, res: int = 30) -> np.ndarray:
        """Cubic trajectory through the loop, ``[res * (len - 1), 2]`` (what the lidar datasets drive along)."""
        return interpolate_waypoints(self.xy[:, 0], self.xy[:, 1], res)

    # ---- edits -----------------------------------------------------------------------------------------------------
    def move(self, ind: int, point) -> None:
        """Move vertex ``ind``; the duplicated closing vertex follows its twin."""
        last = len(self.xy) - 1
        self.xy[ind] = point
        if ind == 0:
            self.xy[last] = point
        elif ind == last:
            self.xy[0] = point

    def delete(self, ind: int) -> bool:
        """Delete vertex ``ind`` (refused when only three distinct vertices remain); keeps the loop closed."""
        if self.n_vertices <= 3:
            return False
        last = len(self.xy) - 1
        if ind in (0, last):
            self.xy = self.xy[1:last]
            self.xy = np.vstack([self.xy, self.xy[:1]])
        else:
            self.xy = np.delete(self.xy, ind, axis=0)
        return True

    def insert(self, point) -> Optional[int]:
        """Insert ``point`` as a new vertex on the first segment within ``epsilon`` of it; returns its index."""
        px = self.to_pixels(self.xy)
        p = self.to_pixels(np.asarray(point, dtype=np.float64))
        for i in range(len(px) - 1):
            if point_segment_distance(p, px[i], px[i + 1]) <= self.epsilon:
                self.xy = np.insert(self.xy, i + 1, np.asarray(point, dtype=np.float64), axis=0)
                return i + 1
        return None

    # ---- persistence -----------------------------------------------------------------------------------------------
    def save(self, target: str) -> str:
        """``target`` ending in ``.npy``: that file; otherwise a directory, saved as the next ``<k>.npy`` in it
        (the reference numbers waypoint files 1.npy, 2.npy, ... per directory)."""
        if target.endswith(".npy"):
            path = target
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        else:
            os.makedirs(target, exist_ok=True)
            path = os.path.join(target, f"{len(glob.glob(os.path.join(target, '*.npy'))) + 1}.npy")
        np.save(path, self.xy)
        return path


class PathEditor:
    """matplotlib front end: drag a vertex with the left button; ``i`` insert at the cursor (on an edge), ``d`` delete the
    vertex under the cursor, ``t`` toggle markers, ``enter`` / closing the window ends the session."""

    def __init__(self, img: np.ndarray, path: WaypointPath, spline_res: int = 30):
        import matplotlib.pyplot as plt   # optional dependency, imported late

        self.plt, self.img, self.path, self.res = plt, img, path, spline_res
        ny, nx = img.shape
        self.fig, self.ax = plt.subplots()
        self.ax.imshow(img, cmap="gray_r", extent=(-1, 1, 1, -1))
        self.ax.set_title("drag: move vertex   i: insert   d: delete   t: toggle markers   enter: done")
        path.to_pixels = lambda a: self.ax.transData.transform(np.asarray(a, dtype=np.float64).reshape(-1, 2)).reshape(np.shape(a))
        (self.edges,) = self.ax.plot(*path.xy.T, "-", color="tab:blue", lw=1)
        (self.verts,) = self.ax.plot(*path.xy.T, "o", color="tab:red", ms=6)
        (self.curve,) = self.ax.plot(*path.spline(spline_res).T, "-", color="tab:orange", lw=2)
        self._drag: Optional[int] = None
        c = self.fig.canvas
        c.mpl_connect("button_press_event", self._press)
        c.mpl_connect("button_release_event", self._release)
        c.mpl_connect("motion_notify_event", self._motion)
        c.mpl_connect("key_press_event", self._key)

    def _redraw(self):
        self.edges.set_data(*self.path.xy.T)
        self.verts.set_data(*self.path.xy.T)
        self.curve.set_data(*self.path.spline(self.res).T)
        free = path_is_free(self.img, self.path.xy, self.res)
        self.curve.set_color("tab:orange" if free else "tab:red")      # red: the trajectory crosses a wall
        self.fig.canvas.draw_idle()

    def _press(self, ev):
        if ev.inaxes is self.ax and ev.button == 1 and self.verts.get_visible():
            self._drag = self.path.hit_test((ev.xdata, ev.ydata))

    def _release(self, ev):
        if ev.button == 1:
            self._drag = None

    def _motion(self, ev):
        if self._drag is not None and ev.inaxes is self.ax and ev.button == 1:
            self.path.move(self._drag, (ev.xdata, ev.ydata))
            self._redraw()

    def _key(self, ev):
        if ev.key == "enter":
            self.plt.close(self.fig)
            return
        if ev.inaxes is not self.ax:
            return
        if ev.key == "t":
            self.verts.set_visible(not self.verts.get_visible())
            self._drag = None
        elif ev.key == "d":
            ind = self.path.hit_test((ev.xdata, ev.ydata))
            if ind is not None:
                self.path.delete(ind)
        elif ev.key == "i":
            self.path.insert((ev.xdata, ev.ydata))
        self._redraw()

    def run(self) -> WaypointPath:
        self.plt.show()
        return self.path


def _arg(argv, flag, default=None):
    return argv[argv.index(flag) + 1] if flag in argv else default


def main(argv=None) -> Tuple[str, np.ndarray]:
    argv = sys.argv if argv is None else argv
    img = np.asarray(Image.open(argv[1]).convert("L")).astype(float) / 255.0
    target = argv[2]
    res = int(_arg(argv, "--spline-res", 30))
    if "--points" in argv:
        pts = np.asarray([[float(v) for v in p.split(",")] for p in _arg(argv, "--points").split(";")])
        if len(pts) < 4:
            raise SystemExit("need at least 4 waypoints for a cubic path")
        path = WaypointPath(normalise(pts, img.shape[1], img.shape[0]))
        if "--open" in argv:                 # keep an open polyline exactly as given (round-1 behaviour of this tool)
            path.xy = path.xy[:-1]
    else:
        start = WaypointPath(np.load(_arg(argv, "--load"))) if "--load" in argv else WaypointPath.circle()
        path = PathEditor(img, start, res).run()
        if input("Save the current trajectory? (y/n): ").strip().lower() != "y":
            print("no save!")
            return "", path.xy
    if not path_is_free(img, path.xy, res):
        print("WARNING: the interpolated trajectory crosses a wall; a lidar scan from it will fail")
    out = path.save(target)
    print("saved", out, path.xy.shape)
    return out, path.xy


if __name__ == "__main__":
    main()
