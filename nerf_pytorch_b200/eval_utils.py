"""Full-image rendering on the fused path (the reference's render loops: eval_nerf.py:156-190, the validation block of
train_nerf.py:287-371).

``render_image`` generates the rays of EVERY pixel on the device from the pose (csrc/raygen.cu: no (H, W, 3) ray
tensors built by a chain of torch ops, view directions / NDC / near / far packed in the same kernel), renders them in
chunks under ``torch.no_grad()`` through ``predict_and_render_radiance`` and returns the reference's 6-tuple in image
shape.  ``cast_to_image`` / ``cast_to_disparity_image`` reproduce eval_nerf.py:23-36 on the device (uint8 tensors)."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops, train_utils


def render_image(height, width, focal_length, tform_cam2world, model_coarse, model_fine, options, mode="validation",
                 encode_position_fn=None, encode_direction_fn=None, *, rays_per_chunk: Optional[int] = None,
                 device=None, impl: Optional[int] = None):
    """One pose -> (rgb_coarse (H,W,3), disp_coarse (H,W), acc_coarse (H,W), rgb_fine, disp_fine, acc_fine).

    Like the reference, the per-chunk sampler options come from ``options.nerf.train`` unless
    ``train_utils.COMPAT_MODE_QUIRK`` is switched off (the reference forgets to forward ``mode``, train_utils.py:171-181);
    the chunk size comes from ``options.nerf.<mode>.chunksize`` (``rays_per_chunk`` overrides it: the fused kernels
    need no small chunks, 180 GB of HBM hold a whole 800 x 800 image's intermediates)."""
    if device is None:
        device = next(model_coarse.parameters()).device
    use_viewdirs = bool(options.nerf.use_viewdirs)
    with torch.no_grad():
        rays = ops.gen_rays(tform_cam2world, height, width, focal_length, None, device,
                            ndc=options.dataset.no_ndc is False, near=options.dataset.near, far=options.dataset.far,
                            use_viewdirs=use_viewdirs)
        n = rays.shape[0]
        chunk = int(rays_per_chunk or getattr(options.nerf, mode).chunksize)
        inner_mode = "train" if train_utils.COMPAT_MODE_QUIRK else mode
        parts = [train_utils.predict_and_render_radiance(
            rays[i:i + chunk], model_coarse, model_fine, options, mode=inner_mode,
            encode_position_fn=encode_position_fn,
            encode_direction_fn=encode_direction_fn if use_viewdirs else None, impl=impl)
            for i in range(0, n, chunk)]
    cols = [torch.cat(c, 0) if c[0] is not None else None for c in zip(*parts)]
    shapes = [(height, width, 3), (height, width), (height, width)] * 2
    return tuple(c.view(s) if c is not None else None for c, s in zip(cols, shapes))


def cast_to_image(rgb: torch.Tensor) -> torch.Tensor:
    """eval_nerf.py:23-30: (H, W, 3) floats in [0, 1] -> uint8 (H, W, 3) the way ToPILImage does it (x * 255, truncated)."""
    return rgb[..., :3].mul(255).clamp(0, 255).to(torch.uint8)


def cast_to_disparity_image(disp: torch.Tensor) -> torch.Tensor:
    """eval_nerf.py:33-36: min-max normalised disparity as uint8.  Rays that hit nothing have NaN disparity in the
    reference (0 / 0, volume_rendering_utils.py:48); they are excluded from the min / max and drawn as 0 here."""
    finite = torch.isfinite(disp)
    d = torch.where(finite, disp, torch.zeros_like(disp))
    if finite.any():
        lo, hi = disp[finite].min(), disp[finite].max()
        d = (d - lo) / (hi - lo).clamp_min(1e-30)
    return (d.clamp(0, 1) * 255).masked_fill(~finite, 0).to(torch.uint8)
