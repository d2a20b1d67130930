"""Reward / agreement curves from the saved ``.npy/.npz`` artefacts (reference:
RL/plot_reward.py, RL/dist_rl/plot_reward.py, RL/plot_agreements.py).  Uses matplotlib when it
is installed and otherwise prints a text summary, so it works on a bare training box."""
from __future__ import annotations

import glob
import os
import sys

import numpy as np


def load_runs(directory, alg):
    runs = []
    for f in sorted(glob.glob(os.path.join(directory, f"avg_ep_rews_{alg}_*.npy"))):
        ID = f.rsplit("_", 1)[1].split(".")[0]
        t = os.path.join(directory, f"timesteps_{alg}_{ID}.npy")
        if os.path.exists(t):
            runs.append((np.load(t), np.load(f)))
    return runs


def summarize(directory="./trained", algs=("dinno", "dsgd", "dsgt")):
    out = {}
    for alg in algs:
        runs = load_runs(directory, alg)
        if runs:
            finals = [r[1][-1] for r in runs]
            out[alg] = {"runs": len(runs), "final_mean": float(np.mean(finals)), "max": float(max(r[1].max() for r in runs))}
            print(f"{alg}: {len(runs)} run(s), final avg episode reward {np.mean(finals):.1f}, best {out[alg]['max']:.1f}")
    return out


def plot(directory="./trained", algs=("dinno", "dsgd", "dsgt"), out="RL_reward.svg"):
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        print("matplotlib not installed: text summary only")
        return summarize(directory, algs)
    fig, ax = plt.subplots(figsize=(6, 4))
    for alg in algs:
        for t, r in load_runs(directory, alg):
            ax.plot(t, r, label=alg, alpha=0.7)
    ax.set_xlabel("environment steps"); ax.set_ylabel("average episode reward"); ax.legend()
    fig.savefig(out)
    return out


def plot_agreements(path, out="RL_agreement.svg"):
    z = np.load(path)
    keys = sorted(z.files)
    print({k: float(z[k][-1]) for k in keys})
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return {k: z[k] for k in keys}
    fig, ax = plt.subplots(figsize=(6, 4))
    for k in keys:
        ax.semilogy(z[k], label=k)
    ax.set_xlabel("iteration"); ax.set_ylabel("distance to mean of normalised parameters"); ax.legend()
    fig.savefig(out)
    return out


if __name__ == "__main__":
    plot(sys.argv[1] if len(sys.argv) > 1 else "./trained")
