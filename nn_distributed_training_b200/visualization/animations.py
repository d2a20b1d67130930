"""Animations and scaling figures from saved results (reference: visualization/animations/mnist_anim.ipynb,
visualization/animations/density_anim.ipynb, visualization/scaling_plots.ipynb).

The reference notebooks render with matplotlib + ffmpeg.  Neither is a dependency here: frames are drawn with
PIL and written as animated GIFs, the numbers behind the scaling figure are returned as plain tables, and the
optional matplotlib figure is produced only when matplotlib is importable.

    python -m nn_distributed_training_b200.visualization.animations mnist   <run_dir> out.gif [--problem dinno --node 0]
    python -m nn_distributed_training_b200.visualization.animations accuracy <run_dir> out.gif
    python -m nn_distributed_training_b200.visualization.animations density <run_dir> out.gif [--node 0]
    python -m nn_distributed_training_b200.visualization.animations scaling <scaling_dir> [--plot out.png]
"""
from __future__ import annotations

import glob
import os
import sys
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from PIL import Image, ImageDraw

from ..experiments.common import read_gpickle
from ..utils import graph_generation
from .results import _stack, density_image, load_results

GOOD, BAD = (0, 255, 127), (255, 69, 0)          # springgreen / orangered, as in the notebook


# ------------------------------------------------------------------------------------- MNIST grid ----
def _node_vec(entry, node) -> torch.Tensor:
    """``validation_as_vector[i]`` is ``{node: bool [V, 1]}`` (the reference's layout) or a ``[N, V]`` tensor."""
    v = entry[node] if isinstance(entry, dict) else torch.as_tensor(entry)[node]
    return torch.as_tensor(v).reshape(-1).bool()


def _all_nodes(entry) -> torch.Tensor:
    keys = sorted(entry) if isinstance(entry, dict) else range(len(entry))
    return torch.stack([_node_vec(entry, k) for k in keys])


def pick_grid_indices(vvecs: Sequence, num_total: int = 100, num_wrong: int = 2, seed: int = 100) -> torch.Tensor:
    """Validation indices for the grid: the first ``num_total - num_wrong`` samples plus ``num_wrong`` samples
    that *every* node still gets wrong at the last evaluation (mnist_anim.ipynb, cell 5), shuffled."""
    last = _all_nodes(vvecs[-1])                                 # [N, V]
    all_wrong = torch.nonzero((~last).all(0)).reshape(-1)
    g = torch.Generator().manual_seed(seed)
    all_wrong = all_wrong[torch.randperm(all_wrong.numel(), generator=g)][:num_wrong]
    taken = set(all_wrong.tolist())       # (the notebook can show such a sample twice; here the grid has no repeats)
    head = torch.as_tensor([i for i in range(last.shape[1]) if i not in taken][: num_total - all_wrong.numel()], dtype=torch.long)
    inds = torch.cat([head, all_wrong])
    return inds[torch.randperm(inds.numel(), generator=g)]


def _digit_tile(img: np.ndarray, ok: bool, cell: int, border: int) -> Image.Image:
    a = np.asarray(img, dtype=np.float32).reshape(28, 28)
    a = (255 * (a - a.min()) / max(float(a.max() - a.min()), 1e-9)).astype(np.uint8)
    tile = Image.new("RGB", (cell, cell), GOOD if ok else BAD)
    inner = Image.fromarray(a, mode="L").resize((cell - 2 * border, cell - 2 * border), Image.NEAREST).convert("RGB")
    tile.paste(inner, (border, border))
    return tile


def mnist_grid_frames(metrics: dict, val_images, node: int = 0, grid: Tuple[int, int] = (10, 10), cell: int = 40,
                      border: int = 4, inds: Optional[torch.Tensor] = None) -> List[Image.Image]:
    """One frame per evaluation: a grid of validation digits framed green / red by whether ``node`` classifies
    them correctly at that evaluation (``validation_as_vector`` metric)."""
    vv = metrics["validation_as_vector"]
    n = grid[0] * grid[1]
    if inds is None:
        inds = pick_grid_indices(vv, num_total=n)
    inds = torch.as_tensor(inds).reshape(-1)[:n]
    frames = []
    for v in vv:
        correct = _node_vec(v, node)
        frame = Image.new("RGB", (grid[1] * cell, grid[0] * cell), (0, 0, 0))
        for q, idx in enumerate(inds.tolist()):
            frame.paste(_digit_tile(np.asarray(val_images[idx]), bool(correct[idx]), cell, border),
                        ((q % grid[1]) * cell, (q // grid[1]) * cell))
        frames.append(frame)
    return frames


# ----------------------------------------------------------------------------------- line figures ----
def _axes(size, xlim, ylim, title=""):
    W, H = size
    im = Image.new("RGB", size, (255, 255, 255))
    d = ImageDraw.Draw(im)
    box = (50, 25, W - 15, H - 35)
    d.rectangle(box, outline=(0, 0, 0))
    for f in (0.0, 0.25, 0.5, 0.75, 1.0):
        y = box[3] - f * (box[3] - box[1]); x = box[0] + f * (box[2] - box[0])
        d.line([(box[0], y), (box[2], y)], fill=(225, 225, 225))
        d.text((4, y - 5), f"{ylim[0] + f * (ylim[1] - ylim[0]):.2f}", fill=(0, 0, 0))
        d.text((x - 10, box[3] + 6), f"{xlim[0] + f * (xlim[1] - xlim[0]):.0f}", fill=(0, 0, 0))
    d.text((box[0], 6), title, fill=(0, 0, 0))

    def to_px(x, y):
        return (box[0] + (x - xlim[0]) / max(xlim[1] - xlim[0], 1e-9) * (box[2] - box[0]),
                box[3] - (y - ylim[0]) / max(ylim[1] - ylim[0], 1e-9) * (box[3] - box[1]))
    return im, d, to_px


def accuracy_frames(metrics: dict, evaluate_frequency: int = 20, size=(640, 320), centralized: Optional[float] = None):
    """Growing mean top-1 curve with the min/max band over nodes (mnist_anim.ipynb, cell 8)."""
    acc = _stack(metrics["top1_accuracy"])
    t = np.arange(acc.shape[0]) * evaluate_frequency
    mean, lo, hi = acc.mean(1), acc.min(1), acc.max(1)
    frames = []
    for i in range(1, len(t) + 1):
        im, d, px = _axes(size, (0, max(float(t[-1]), 1.0)), (0.0, 1.0), "validation accuracy (mean, min-max over nodes)")
        if centralized is not None:
            d.line([px(0, centralized), px(t[-1], centralized)], fill=(75, 0, 130), width=2)
        if i > 1:
            band = [px(t[j], hi[j]) for j in range(i)] + [px(t[j], lo[j]) for j in reversed(range(i))]
            d.polygon(band, fill=(255, 214, 170))
            d.line([px(t[j], mean[j]) for j in range(i)], fill=(255, 140, 0), width=2)
        frames.append(im)
    return frames


# ---------------------------------------------------------------------------------------- density ----
def density_frames(metrics: dict, node: int = 0, scale: int = 4) -> List[Image.Image]:
    """The learned occupancy map of ``node`` at every evaluation (``mesh_grid_density`` on the every-8th-pixel mesh;
    density_anim.ipynb: MeshAnimation)."""
    frames = []
    for i in range(len(metrics["mesh_grid_density"])):
        a = np.clip(density_image(metrics, node=node, index=i), 0.0, 1.0)
        im = Image.fromarray((255 * (1.0 - a)).astype(np.uint8), mode="L")
        frames.append(im.resize((im.width * scale, im.height * scale), Image.NEAREST).convert("RGB"))
    return frames


def save_gif(frames: List[Image.Image], out: str, fps: int = 15) -> str:
    if not frames:
        raise ValueError("no frames to write")
    frames[0].save(out, save_all=True, append_images=frames[1:], duration=max(int(1000 / fps), 20), loop=0)
    return out


# ---------------------------------------------------------------------------------------- scaling ----
def scaling_table(scaling_dir: str, evaluate_frequency: int = 100, thresholds=(0.5, 0.9, 0.97)) -> List[dict]:
    """Per trial of a ``dist_mnist_scaling`` run (``<k>.gpickle`` + ``<k>_results.pt``): number of nodes, algebraic
    connectivity and the first evaluated round whose mean top-1 reaches each threshold — the data of
    visualization/scaling_plots.ipynb (``None`` where the threshold is never reached; the notebook's argmax
    silently reports round 0 there)."""
    rows = []
    for g in sorted(glob.glob(os.path.join(scaling_dir, "*.gpickle")), key=lambda p: int(os.path.basename(p).split(".")[0])):
        k = os.path.basename(g).split(".")[0]
        res = os.path.join(scaling_dir, f"{k}_results.pt")
        if not os.path.exists(res):
            continue
        graph = read_gpickle(g)
        m = torch.load(res, map_location="cpu", weights_only=False)
        mean = _stack(m["top1_accuracy"]).mean(1)
        row = {"trial": int(k), "N": graph.number_of_nodes(), "fiedler": float(graph_generation.fiedler_value(graph_generation.adjacency(graph)))}
        for th in thresholds:
            hit = np.nonzero(mean >= th)[0]
            row[f"rounds_to_{int(round(th * 100))}"] = int(hit[0] * evaluate_frequency) if hit.size else None
        row["final_mean_top1"] = float(mean[-1])
        rows.append(row)
    return rows


def plot_scaling(rows: List[dict], out: str, x: str = "N") -> Optional[str]:
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        print("matplotlib is not installed: table only")
        return None
    keys = [k for k in rows[0] if k.startswith("rounds_to_")]
    fig, ax = plt.subplots(figsize=(5, 4), tight_layout=True)
    for k, mk, c in zip(keys, "so^", ("darkgrey", "dimgrey", "black")):
        ax.plot([r[x] for r in rows], [r[k] if r[k] is not None else np.nan for r in rows], marker=mk, c=c, label=k[10:] + "%")
    ax.set_xlabel("Number of Robots" if x == "N" else "Fiedler Value"); ax.set_ylabel("Iterations to Reach"); ax.grid(); ax.legend()
    fig.savefig(out)
    return out


# -------------------------------------------------------------------------------------------- CLI ----
def _arg(argv, flag, default):
    return type(default)(argv[argv.index(flag) + 1]) if flag in argv else default


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    kind, path = argv[0], argv[1]
    if kind == "scaling":
        rows = scaling_table(path, evaluate_frequency=_arg(argv, "--every", 100))
        for r in rows:
            print(r)
        if "--plot" in argv:
            plot_scaling(rows, argv[argv.index("--plot") + 1], x=_arg(argv, "--x", "N"))
        return rows
    out = argv[2]
    res = load_results(path)
    name = _arg(argv, "--problem", sorted(res)[0])
    m, node = res[name], _arg(argv, "--node", 0)
    if kind == "mnist":
        from ..data.mnist import load_mnist
        val, _ = load_mnist(_arg(argv, "--data", "../data"), train=False)
        frames = mnist_grid_frames(m, val.x.reshape(len(val), -1), node=node)
    elif kind == "accuracy":
        frames = accuracy_frames(m, evaluate_frequency=_arg(argv, "--every", 20))
    elif kind == "density":
        frames = density_frames(m, node=node)
    else:
        raise SystemExit(f"unknown animation '{kind}'")
    print(save_gif(frames, out, fps=_arg(argv, "--fps", 15)), len(frames), "frames")
    return out


if __name__ == "__main__":
    main()
