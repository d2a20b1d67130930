"""Summarise a dist_online_dense run directory (validation-loss curve per algorithm) -> markdown on stdout.
usage: python scripts/summarize_density_run.py <results_dir>"""
import glob
import os
import sys

import torch

d = sys.argv[1]
runs = sorted(glob.glob(os.path.join(d, "*_results.pt")))
print("| problem | evaluations | val loss start (mean over robots) | @25% | @50% | @75% | final mean | final min - max | train-loss MA final |")
print("|---|---|---|---|---|---|---|---|---|")
for p in runs:
    if os.path.basename(p).startswith("solo"):
        continue
    r = torch.load(p, weights_only=False)
    vl = torch.stack([torch.as_tensor(v, dtype=torch.float64) for v in r["validation_loss"]])
    n = vl.shape[0]
    tl = r.get("train_loss_moving_average", [None])[-1]
    tlm = "-" if tl is None else f"{float(torch.as_tensor(tl).mean()):.4f}"
    print(f"| {os.path.basename(p)[:-11]} | {n} | {vl[0].mean():.3f} | {vl[n // 4].mean():.3f} | {vl[n // 2].mean():.3f} | {vl[3 * n // 4].mean():.3f} | "
          f"{vl[-1].mean():.3f} | {vl[-1].min():.3f} - {vl[-1].max():.3f} | {tlm} |")
solo = os.path.join(d, "solo_results.pt")
if os.path.exists(solo):
    s = torch.load(solo, weights_only=False)
    vals = [float(torch.as_tensor(v["validation_loss"])) for v in s.values()]
    print(f"\nindividual training (no communication), 1 epoch: validation loss per robot {', '.join(f'{v:.2f}' for v in vals)}")
