"""CubiCasa5k pre-processing: black/white floor-plan renderings -> signed-distance tensors + a train/test split
(reference: floorplans/cubi_preproc.py:1-96).

Same artefacts as the reference's script, written to ``target_dir``:

* ``<image stem>.pt``   float32 SDF of the image resized so that its shorter side is ``small_sidelen``: distance to the
  nearest boundary pixel divided by the image height, **negative inside** the bright (== 1) region, positive outside
  (``SDFTransform``, reference :12-35);
* ``pzcounts.pt``       ``{name.pt: {"npixels": int, "nzeros": int}}`` (number of exactly-zero SDF pixels, :72-76);
* ``split_sets.pt``     ``{"train": [names], "test": [names]}``, shuffled 90 / 10 (:81-88).

The reference finds the boundary with scikit-image's Canny detector; on a thresholded (binary) image that is the set
of pixels where the value changes, which is computed here directly with scipy.ndimage (scikit-image and torchvision
are not dependencies of this package).

    python -m nn_distributed_training_b200.floorplans.cubi_preproc <source_dir> <target_dir> [--sidelen 512] [--max N]
                                                                   [--seed S] [--no-overwrite]
"""
from __future__ import annotations

import os
import random
import sys
from typing import Dict

import numpy as np
import torch
from PIL import Image
from scipy import ndimage

IMAGE_EXT = (".png", ".jpg", ".jpeg", ".bmp", ".tif", ".tiff")


def binarize(img: np.ndarray) -> np.ndarray:
    """Grey image in [0, 1] (or 0..255) -> {0, 1} with the reference's 0.5 threshold."""
    g = np.asarray(img, dtype=np.float64)
    if g.ndim == 3:
        g = g[..., :3].mean(-1)
    if g.max() > 1.0:
        g = g / 255.0
    return (g >= 0.5).astype(np.float64)


def boundary(img_bin: np.ndarray) -> np.ndarray:
    """Pixels of the bright region that touch the dark region (4-neighbourhood): the edge map of a binary image."""
    fg = img_bin == 1.0
    return fg & ~ndimage.binary_erosion(fg, border_value=1)


def sdf_from_binary(img_bin: np.ndarray) -> torch.Tensor:
    """Distance to the boundary, negative where the image is 1, divided by the image height."""
    edge = boundary(img_bin)
    if not edge.any():                       # uniform image: no boundary, distance is undefined -> zeros
        return torch.zeros(img_bin.shape, dtype=torch.float32)
    sdf = ndimage.distance_transform_edt(~edge)
    sdf[img_bin == 1.0] *= -1.0
    sdf /= float(img_bin.shape[0])
    return torch.as_tensor(sdf, dtype=torch.float32)


def load_resized(path: str, small_sidelen: int) -> np.ndarray:
    """``transforms.Resize(small_sidelen)``: the shorter side becomes ``small_sidelen``, aspect ratio kept."""
    img = Image.open(path).convert("L")
    w, h = img.size
    if w <= h:
        size = (small_sidelen, max(1, int(round(h * small_sidelen / w))))
    else:
        size = (max(1, int(round(w * small_sidelen / h))), small_sidelen)
    return np.asarray(img.resize(size, Image.BILINEAR), dtype=np.float64) / 255.0


def process_plan(path: str, small_sidelen: int = 512) -> torch.Tensor:
    return sdf_from_binary(binarize(load_resized(path, small_sidelen)))


def cubi_preprocess(source_dir: str, target_dir: str, small_sidelen: int = 512, overwrite: bool = True,
                    limit: int | None = None, seed: int | None = None) -> Dict[str, list]:
    os.makedirs(target_dir, exist_ok=True)
    fnames = sorted(f for f in os.listdir(source_dir) if f.lower().endswith(IMAGE_EXT))[:limit]
    counts_path = os.path.join(target_dir, "pzcounts.pt")
    pzcounts = torch.load(counts_path, weights_only=False) if os.path.isfile(counts_path) else {}
    for i, name in enumerate(fnames):
        name_str = os.path.splitext(name)[0] + ".pt"
        save_pth = os.path.join(target_dir, name_str)
        if os.path.isfile(save_pth) and not overwrite:
            continue
        sdf = process_plan(os.path.join(source_dir, name), small_sidelen)
        torch.save(sdf, save_pth)
        pzcounts[name_str] = {"npixels": int(sdf.numel()), "nzeros": int((sdf == 0.0).sum())}
        if (i + 1) % 100 == 0:
            print("Progress: ", i, " / ", len(fnames))
    ks = list(pzcounts.keys())
    random.Random(seed).shuffle(ks)
    delim = int(len(ks) * 0.9)
    split_sets = {"train": ks[:delim], "test": ks[delim:]}
    torch.save(split_sets, os.path.join(target_dir, "split_sets.pt"))
    torch.save(pzcounts, counts_path)
    return split_sets


def main(argv=None):
    argv = list(sys.argv if argv is None else argv)
    if len(argv) < 3:
        print(__doc__)
        return None

    def opt(flag, default, cast=int):
        return cast(argv[argv.index(flag) + 1]) if flag in argv else default
    split = cubi_preprocess(argv[1], argv[2], small_sidelen=opt("--sidelen", 512), overwrite="--no-overwrite" not in argv,
                            limit=opt("--max", None), seed=opt("--seed", None))
    print(f"processed {len(split['train']) + len(split['test'])} plans -> {argv[2]}")
    return split


if __name__ == "__main__":
    main()
