"""Device-resident pool for the reference's cached dataset.

``cache_dataset.py:104-135`` of the reference writes one ``torch.save`` dict per training image,
``{"height", "width", "focal_length", "ray_bundle": (2, N, 3) | (2, H, W, 3), "target": (N, 3|4) | (H, W, 3|4)}``, and one per
validation image, ``{"height", "width", "focal_length", "ray_origins", "ray_directions", "target"}``.  Its training loop then
does, EVERY iteration, ``np.random.choice(train_paths)`` -> ``torch.load`` -> ``.to(device)`` -> ``np.random.choice`` of the
ray indices (``train_nerf.py:175-193``): a disk read, a host-to-device copy of a whole image's rays and a host-side
permutation per step.  This pool reads the same files ONCE, keeps every image's rays and targets on the device and draws
the batch there.  Only the on-disk format is shared with the reference; nothing here touches the render path."""
from __future__ import annotations

import glob
import os
from typing import List, Optional, Tuple

import torch

__all__ = ["CachedRayPool", "CachedValidationSet"]


def _load(path):
    try:
        return torch.load(path, map_location="cpu", weights_only=False)   # torch >= 2.6 defaults to weights_only=True
    except TypeError:  # older torch without the keyword
        return torch.load(path, map_location="cpu")


class CachedRayPool:
    """All ``<cachedir>/train/*.data`` files as one device-resident pool of (origin, direction, target) rows.

    ``sample(n)`` mirrors the reference's two-stage draw: one cached image uniformly at random, then ``n`` distinct rays
    of it (``replace=False``; an image with fewer than ``n`` rays raises, like ``np.random.choice`` does there).
    ``sample_global(n)`` draws from the union of all images instead."""

    def __init__(self, paths: List[str], device="cuda"):
        if not paths:
            raise FileNotFoundError("CachedRayPool: no .data files given")
        ros, rds, tgts, self.meta, self.offsets = [], [], [], [], [0]
        for p in paths:
            d = _load(p)
            rb = d["ray_bundle"]
            ro, rd = rb[0].reshape(-1, 3).float(), rb[1].reshape(-1, 3).float()
            tgt = d["target"][..., :3].reshape(-1, 3).float()
            if not (ro.shape == rd.shape == tgt.shape):
                raise RuntimeError(f"CachedRayPool: {p}: ray_bundle {tuple(rb.shape)} and target {tuple(d['target'].shape)} disagree")
            ros.append(ro); rds.append(rd); tgts.append(tgt)
            self.meta.append((int(d["height"]), int(d["width"]), float(d["focal_length"])))
            self.offsets.append(self.offsets[-1] + ro.shape[0])
        self.device = torch.device(device)
        self.ray_origins = torch.cat(ros).to(self.device)
        self.ray_directions = torch.cat(rds).to(self.device)
        self.targets = torch.cat(tgts).to(self.device)
        self.paths = list(paths)

    @classmethod
    def from_dir(cls, cachedir: str, split: str = "train", device="cuda") -> "CachedRayPool":
        return cls(sorted(glob.glob(os.path.join(cachedir, split, "*.data"))), device=device)

    def __len__(self) -> int:
        return self.offsets[-1]

    @property
    def num_images(self) -> int:
        return len(self.meta)

    def sample(self, n: int, generator: Optional[torch.Generator] = None):
        """-> (height, width, focal_length, ray_origins (n, 3), ray_directions (n, 3), target (n, 3)) of ONE cached image."""
        i = int(torch.randint(self.num_images, (1,), generator=generator, device=generator.device if generator is not None else "cpu").item())
        lo, hi = self.offsets[i], self.offsets[i + 1]
        if n > hi - lo:
            raise ValueError(f"CachedRayPool: cannot take {n} distinct rays from an image that cached {hi - lo}")
        gdev = generator.device if generator is not None else self.device
        sel = lo + torch.randperm(hi - lo, generator=generator, device=gdev)[:n].to(self.device)
        h, w, f = self.meta[i]
        return h, w, f, self.ray_origins[sel], self.ray_directions[sel], self.targets[sel]

    def sample_global(self, n: int, generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """n rows drawn (with replacement) from the union of all cached images; all images must share (H, W, focal)."""
        gdev = generator.device if generator is not None else self.device
        sel = torch.randint(len(self), (n,), generator=generator, device=gdev).to(self.device)
        return self.ray_origins[sel], self.ray_directions[sel], self.targets[sel]


class CachedValidationSet:
    """``<cachedir>/val/*.data``: whole images ``(H, W, 3)`` rays + target, kept on the device (``train_nerf.py:301-317``)."""

    def __init__(self, paths: List[str], device="cuda"):
        if not paths:
            raise FileNotFoundError("CachedValidationSet: no .data files given")
        self.items = []
        dev = torch.device(device)
        for p in paths:
            d = _load(p)
            self.items.append((int(d["height"]), int(d["width"]), float(d["focal_length"]), d["ray_origins"].float().to(dev),
                               d["ray_directions"].float().to(dev), d["target"].float().to(dev)))

    @classmethod
    def from_dir(cls, cachedir: str, device="cuda") -> "CachedValidationSet":
        return cls(sorted(glob.glob(os.path.join(cachedir, "val", "*.data"))), device=device)

    def __len__(self) -> int:
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]

    def choice(self, generator: Optional[torch.Generator] = None):
        gdev = generator.device if generator is not None else "cpu"
        return self.items[int(torch.randint(len(self.items), (1,), generator=generator, device=gdev).item())]
