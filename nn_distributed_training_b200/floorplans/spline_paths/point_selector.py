"""Waypoint path authoring tool (reference: floorplans/spline_paths/point_selector.py — an
interactive matplotlib polygon editor that saves waypoint ``.npy`` files, :262-270).

Interactive mode needs matplotlib (click to add points on the floor plan, ``enter`` to save);
without it, ``--points "x0,y0;x1,y1;..."`` (pixel coordinates) builds the same normalised
``[-1,1]^2`` waypoint file and verifies that the interpolated trajectory stays in free space.

    python -m nn_distributed_training_b200.floorplans.spline_paths.point_selector <floor_img.png> <out.npy> \
        [--points "120,80;300,90;310,400"] [--spline-res 30]
"""
from __future__ import annotations

import sys

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


def interactive(img: np.ndarray):
    import matplotlib.pyplot as plt   # noqa: optional dependency

    pts = []
    fig, ax = plt.subplots()
    ax.imshow(img, cmap="gray_r")
    line, = ax.plot([], [], "o-r")

    def on_click(ev):
        if ev.inaxes is ax and ev.xdata is not None:
            pts.append((ev.xdata, ev.ydata))
            line.set_data(*zip(*pts))
            fig.canvas.draw_idle()

    def on_key(ev):
        if ev.key == "enter":
            plt.close(fig)
        elif ev.key == "backspace" and pts:
            pts.pop()
            line.set_data(*zip(*pts)) if pts else line.set_data([], [])
            fig.canvas.draw_idle()

    fig.canvas.mpl_connect("button_press_event", on_click)
    fig.canvas.mpl_connect("key_press_event", on_key)
    plt.show()
    return np.asarray(pts, dtype=np.float64)


def main(argv=None):
    argv = sys.argv if argv is None else argv
    img = np.asarray(Image.open(argv[1]).convert("L")).astype(float) / 255.0
    res = int(argv[argv.index("--spline-res") + 1]) if "--spline-res" in argv else 30
    if "--points" in argv:
        pts = np.asarray([[float(v) for v in p.split(",")] for p in argv[argv.index("--points") + 1].split(";")])
    else:
        pts = interactive(img)
    if len(pts) < 4:
        raise SystemExit("need at least 4 waypoints for a cubic path")
    wp = normalise(pts, img.shape[1], img.shape[0])
    if not path_is_free(img, wp, res):
        print("WARNING: the interpolated trajectory crosses a wall; a lidar scan from it will fail")
    np.save(argv[2], wp)
    print("saved", argv[2], wp.shape)


if __name__ == "__main__":
    main()
