"""Host-side mirror of the reference's nerf/nerf_helpers.py for the names the hot path exports.

Same names, argument meaning and return conventions as the reference (file:line cited per
function); ray generation and the tiny scalar helpers stay plain torch (host plumbing), the
per-ray math (encoding, resampling, exclusive cumprod inside compositing) runs in the CUDA library."""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import ops


def img2mse(img_src, img_tgt):
    """nerf/nerf_helpers.py:9-10."""
    return torch.nn.functional.mse_loss(img_src, img_tgt)


def mse2psnr(mse):
    """nerf/nerf_helpers.py:13-17."""
    if mse == 0:
        mse = 1e-5
    return -10.0 * math.log10(mse)


def get_minibatches(inputs: torch.Tensor, chunksize: Optional[int] = 1024 * 8):
    """nerf/nerf_helpers.py:20-25."""
    return [inputs[i : i + chunksize] for i in range(0, inputs.shape[0], chunksize)]


def meshgrid_xy(tensor1: torch.Tensor, tensor2: torch.Tensor):
    """nerf/nerf_helpers.py:28-40."""
    ii, jj = torch.meshgrid(tensor1, tensor2, indexing="ij")
    return ii.transpose(-1, -2), jj.transpose(-1, -2)


def get_ray_bundle(height: int, width: int, focal_length: float, tform_cam2world: torch.Tensor):
    """nerf/nerf_helpers.py:67-110: (H, W, 3) origins and un-normalised directions."""
    ii, jj = meshgrid_xy(
        torch.arange(width, dtype=tform_cam2world.dtype, device=tform_cam2world.device),
        torch.arange(height, dtype=tform_cam2world.dtype, device=tform_cam2world.device),
    )
    directions = torch.stack(
        [(ii - width * 0.5) / focal_length, -(jj - height * 0.5) / focal_length, -torch.ones_like(ii)], dim=-1
    )
    ray_directions = torch.sum(directions[..., None, :] * tform_cam2world[:3, :3], dim=-1)
    ray_origins = tform_cam2world[:3, -1].expand(ray_directions.shape)
    return ray_origins, ray_directions


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """nerf/nerf_helpers.py:170-197."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1.0 / (W / (2.0 * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1.0 / (H / (2.0 * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1.0 + 2.0 * near / rays_o[..., 2]
    d0 = -1.0 / (W / (2.0 * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1.0 / (H / (2.0 * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2.0 * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


class Embedder:
    """What ``get_embedding_function`` returns: callable like the reference's lambda
    (nerf/nerf_helpers.py:160-167) but carrying its parameters, so the fused path can read them
    instead of treating the encoder as opaque."""

    def __init__(self, num_encoding_functions=6, include_input=True, log_sampling=True):
        self.num_encoding_functions = int(num_encoding_functions)
        self.include_input = bool(include_input)
        self.log_sampling = bool(log_sampling)

    def __call__(self, x):
        return positional_encoding(x, self.num_encoding_functions, self.include_input, self.log_sampling)

    def __repr__(self):
        return (f"Embedder(L={self.num_encoding_functions}, include_input={self.include_input}, "
                f"log_sampling={self.log_sampling})")


def get_embedding_function(num_encoding_functions=6, include_input=True, log_sampling=True):
    """nerf/nerf_helpers.py:160-167."""
    return Embedder(num_encoding_functions, include_input, log_sampling)


def positional_encoding(tensor, num_encoding_functions=6, include_input=True, log_sampling=True) -> torch.Tensor:
    """nerf/nerf_helpers.py:113-157 on the CUDA library (inputs (..., 3), fp32, CUDA)."""
    if tensor.shape[-1] != 3:
        raise NotImplementedError("nerfb200 positional_encoding: last dimension must be 3")
    arch = ops.ArchSpec(n_freq_xyz=num_encoding_functions, include_input_xyz=include_input,
                        log_sampling_xyz=log_sampling)
    flat = tensor.reshape(-1, 3).contiguous().float()
    out = ops.encode(arch, 0, flat)
    return out.reshape(*tensor.shape[:-1], out.shape[-1])


def sample_pdf(z_vals, weights, num_samples, det=False, u: Optional[torch.Tensor] = None):
    """Hierarchical resampling as the render driver uses it (train_utils.py:96-105 of the
    reference): given the COARSE depths ``z_vals`` (N, Nc) and compositing weights (N, Nc) it forms
    the mid-point bins and interior weights, runs sample_pdf_2 (nerf_helpers.py:260-302) and returns
    the sorted union of coarse depths and new samples, (N, Nc + num_samples).

    Differs from the reference's ``sample_pdf_2(bins, weights[...,1:-1], ...)`` call shape on
    purpose: slicing, mid-points, inverse CDF and the sort are one CUDA kernel here."""
    n_rays = z_vals.shape[0]
    if det:
        u = torch.linspace(0.0, 1.0, steps=num_samples, dtype=torch.float32, device=z_vals.device)
    elif u is None:
        u = torch.rand(n_rays, num_samples, dtype=torch.float32, device=z_vals.device)
    return ops.sample_pdf_merge(z_vals.contiguous(), weights.contiguous(), u.contiguous(), num_samples)
