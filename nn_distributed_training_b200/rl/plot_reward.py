"""Reward / agreement curves from the saved ``.npy/.npz`` artefacts (reference: RL/plot_reward.py,
RL/dist_rl/plot_reward.py, RL/plot_agreements.py).

File-name conventions understood (the reference's and this package's):

* centralized PPO:      ``avg_ep_rews_<ID>.npy`` + ``timesteps_<ID>.npy``
* distributed trainers: ``avg_ep_rews_<alg>_<ID>.npy`` + ``timesteps_<alg>_<ID>.npy`` with ``alg`` in
  ``dinno | cadmm | dsgd | dsgt`` (``cadmm`` is the reference's older name for DiNNO; both are read as one series)
* agreements:           ``agreements_<alg>_<ID>.npz`` with ``agree_0 .. agree_{N-1}``

Like the reference's figure, runs of the same method are combined into a mean curve with a min/max band.  matplotlib is
used when it is installed; otherwise the same figure is drawn with PIL (PNG) so the tool works on a bare training box.

    python -m nn_distributed_training_b200.rl.plot_reward <trained_dir> [out.png] [--centralized <dir>]
"""
from __future__ import annotations

import glob
import os
import re
import sys
from typing import Dict, List, Optional, Tuple

import numpy as np

ALIASES = {"dinno": ("dinno", "cadmm"), "dsgd": ("dsgd",), "dsgt": ("dsgt",)}
COLORS = {"centralized": (75, 0, 130), "dinno": (255, 140, 0), "dsgt": (50, 205, 50), "dsgd": (128, 0, 128)}
LABELS = {"centralized": "Centralized", "dinno": "DiNNO", "dsgt": "DSGT", "dsgd": "DSGD"}


def load_runs(directory: str, alg: Optional[str]) -> List[Tuple[np.ndarray, np.ndarray]]:
    """``[(timesteps, avg_ep_rews)]`` of every run of ``alg`` (``None`` / ``"centralized"``: the centralized PPO files)."""
    runs = []
    if alg in (None, "centralized"):
        for f in sorted(glob.glob(os.path.join(directory, "avg_ep_rews_*.npy"))):
            m = re.fullmatch(r"avg_ep_rews_(\d+)\.npy", os.path.basename(f))
            t = os.path.join(directory, f"timesteps_{m.group(1)}.npy") if m else None
            if t and os.path.exists(t):
                runs.append((np.load(t), np.load(f)))
        return runs
    for name in ALIASES.get(alg, (alg,)):
        for f in sorted(glob.glob(os.path.join(directory, f"avg_ep_rews_{name}_*.npy"))):
            ID = f.rsplit("_", 1)[1].split(".")[0]
            t = os.path.join(directory, f"timesteps_{name}_{ID}.npy")
            if os.path.exists(t):
                runs.append((np.load(t), np.load(f)))
    return runs


def combine(runs) -> Optional[Dict[str, np.ndarray]]:
    """Mean / min / max over runs on the shortest common length (the reference stacks equally long runs)."""
    if not runs:
        return None
    n = min(len(r[1]) for r in runs)
    arr = np.vstack([np.asarray(r[1][:n], dtype=np.float64) for r in runs])
    return {"t": np.asarray(runs[0][0][:n], dtype=np.float64), "mean": arr.mean(0), "min": arr.min(0), "max": arr.max(0),
            "runs": len(runs)}


def collect(directory="./trained", centralized_dir: Optional[str] = None, algs=("dinno", "dsgt", "dsgd")):
    series = {}
    c = combine(load_runs(centralized_dir or directory, None))
    if c is not None:
        series["centralized"] = c
    for alg in algs:
        c = combine(load_runs(directory, alg))
        if c is not None:
            series[alg] = c
    return series


def summarize(directory="./trained", algs=("dinno", "dsgd", "dsgt"), centralized_dir: Optional[str] = None):
    out = {}
    for name, c in collect(directory, centralized_dir, algs).items():
        out[name] = {"runs": int(c["runs"]), "final_mean": float(c["mean"][-1]), "max": float(c["max"].max())}
        print(f"{name}: {c['runs']} run(s), final avg episode reward {c['mean'][-1]:.1f}, best {c['max'].max():.1f}")
    return out


def _plot_pil(series, out, xlabel, ylabel, band=True):
    from PIL import ImageDraw  # noqa: F401  (PIL is a hard dependency of the package: lidar floor plans)
    from ..visualization.animations import _axes
    xs = np.concatenate([c["t"] for c in series.values()])
    lo = min(float(c["min"].min()) for c in series.values())
    hi = max(float(c["max"].max()) for c in series.values())
    pad = 0.05 * max(hi - lo, 1e-9)
    im, d, px = _axes((800, 480), (float(xs.min()), float(xs.max())), (lo - pad, hi + pad), f"{ylabel} vs {xlabel}")
    for k, (name, c) in enumerate(series.items()):
        col = COLORS.get(name, (0, 0, 0))
        if band and c["runs"] > 1:
            light = tuple(int(255 - 0.35 * (255 - v)) for v in col)
            d.polygon([px(x, y) for x, y in zip(c["t"], c["max"])] + [px(x, y) for x, y in zip(c["t"][::-1], c["min"][::-1])], fill=light)
        if len(c["t"]) > 1:
            d.line([px(x, y) for x, y in zip(c["t"], c["mean"])], fill=col, width=2)
        d.text((60, 30 + 14 * k), LABELS.get(name, name), fill=col)
    im.save(out)
    return out


def plot(directory="./trained", algs=("dinno", "dsgt", "dsgd"), out="RL_reward.png", centralized_dir: Optional[str] = None):
    series = collect(directory, centralized_dir, algs)
    if not series:
        print("no reward curves found under", directory)
        return None
    summarize(directory, algs, centralized_dir)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        if out.endswith(".svg"):
            out = out[:-4] + ".png"
        return _plot_pil(series, out, "Timestep", "Average Episode Reward")
    fig, ax = plt.subplots(figsize=(10, 8), tight_layout=True)
    for name, c in series.items():
        col = tuple(v / 255 for v in COLORS.get(name, (0, 0, 0)))
        ax.plot(c["t"], c["mean"], c=col, label=LABELS.get(name, name))
        if c["runs"] > 1:
            ax.fill_between(c["t"], c["max"], c["min"], color=col, alpha=0.3)
    ax.legend(); ax.set_xlabel("Timestep"); ax.set_ylabel("Average Episode Reward"); ax.grid(zorder=0)
    fig.savefig(out)
    return out


def plot_agreements(path, out="RL_agreement.png", timesteps: Optional[str] = None):
    """``agreements_<alg>_<ID>.npz``: distance of every node's normalised parameters to their mean, per iteration
    (reference: RL/plot_agreements.py; x axis = ``timesteps_<alg>_<ID>.npy`` when given / found next to the file)."""
    z = np.load(path)
    keys = sorted(z.files)
    if timesteps is None:
        cand = path.replace("agreements_", "timesteps_").replace(".npz", ".npy")
        timesteps = cand if os.path.exists(cand) else None
    n = min(len(z[k]) for k in keys)
    t = np.load(timesteps)[:n] if timesteps else np.arange(n)
    print({k: float(z[k][n - 1]) for k in keys})
    series = {k: {"t": np.asarray(t, dtype=np.float64), "mean": np.asarray(z[k][:n], dtype=np.float64),
                  "min": np.asarray(z[k][:n], dtype=np.float64), "max": np.asarray(z[k][:n], dtype=np.float64), "runs": 1}
              for k in keys}
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        if out.endswith(".svg"):
            out = out[:-4] + ".png"
        return _plot_pil(series, out, "Timestep", "Distance to Mean Parameter Value", band=False)
    fig, ax = plt.subplots(figsize=(10, 8), tight_layout=True)
    for k in keys:
        ax.plot(t, z[k][:n], label=k)
    ax.set_xlabel("Timestep"); ax.set_ylabel("Distance to Mean Parameter Value"); ax.grid(zorder=0); ax.legend()
    fig.savefig(out)
    return out


if __name__ == "__main__":
    args = sys.argv[1:]
    cdir = args[args.index("--centralized") + 1] if "--centralized" in args else None
    pos = [a for i, a in enumerate(args) if not a.startswith("--") and (i == 0 or args[i - 1] != "--centralized")]
    plot(pos[0] if pos else "./trained", out=pos[1] if len(pos) > 1 else "RL_reward.png", centralized_dir=cdir)
