"""CubiCasa5k pre-processing: floor-plan image -> wall mask -> signed-distance tensor, plus a
train/test split of the processed plans (reference: floorplans/cubi_preproc.py:1-96, which
needs scikit-image; scipy.ndimage provides the same distance transform here).

    python -m nn_distributed_training_b200.floorplans.cubi_preproc <cubicasa_root> <out_dir> [--max N]
"""
from __future__ import annotations

import glob
import os
import sys

import numpy as np
import torch
from PIL import Image
from scipy import ndimage


def wall_mask(img: np.ndarray, thresh: float = 0.5) -> np.ndarray:
    """Dark pixels of a rendered plan are walls."""
    g = img.astype(np.float64)
    if g.ndim == 3:
        g = g[..., :3].mean(-1)
    g = g / (255.0 if g.max() > 1.0 else 1.0)
    return g < thresh


def signed_distance(mask: np.ndarray) -> np.ndarray:
    """Positive outside walls, negative inside (pixels)."""
    out = ndimage.distance_transform_edt(~mask)
    inn = ndimage.distance_transform_edt(mask)
    return out - inn


def process_plan(path: str, size=(512, 512)) -> torch.Tensor:
    img = Image.open(path).convert("L").resize(size)
    return torch.from_numpy(signed_distance(wall_mask(np.asarray(img)))).float()


def main(argv=None):
    argv = sys.argv if argv is None else argv
    root, out = argv[1], argv[2]
    limit = int(argv[argv.index("--max") + 1]) if "--max" in argv else None
    files = sorted(glob.glob(os.path.join(root, "**", "F1_scaled.png"), recursive=True))[:limit]
    os.makedirs(out, exist_ok=True)
    sdfs = [process_plan(f) for f in files]
    if not sdfs:
        print("no plans found under", root)
        return
    data = torch.stack(sdfs)
    perm = torch.randperm(len(data))
    n_test = max(1, len(data) // 10)
    torch.save(data[perm[n_test:]], os.path.join(out, "sdf_train.pt"))
    torch.save(data[perm[:n_test]], os.path.join(out, "sdf_test.pt"))
    print(f"processed {len(data)} plans -> {out}")


if __name__ == "__main__":
    main()
