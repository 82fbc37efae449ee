"""CPU restatement of tiny_nerf.py's private pipeline (BASELINE.json configs[0])  --  TEST
INFRASTRUCTURE, never imported by the product.  Pinned bit-for-bit against the unmodified reference by
oracle/make_golden_tiny.py (vectors in tests/golden/tiny_nerf.npz).

The tiny pipeline shares only `get_ray_bundle`, `positional_encoding`, `cumprod_exclusive` and `get_minibatches` with the
main path (tiny_nerf.py:9); its sampler, compositing and model are its own.  The CUDA product does not implement this
configuration: the reference marks it CPU-only plumbing, and it is covered here at the oracle level."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import nerf_oracle as O


def compute_query_points_from_rays(ray_origins, ray_directions, near_thresh, far_thresh, num_samples, randomize=True,
                                   rand: Optional[torch.Tensor] = None):
    """tiny_nerf.py:12-65: a uniform depth grid linspace(near, far, S), jittered by U[0,1) * (far - near) / S per (ray, sample)
    (so the jitter can cross into the next cell, unlike the stratified sampler of train_utils.py:45-65); points = o + d * depth.
    `rand` injects the uniform tensor the reference draws with torch.rand(noise_shape)."""
    depth_values = torch.linspace(near_thresh, far_thresh, num_samples).to(ray_origins)
    if randomize is True:
        noise_shape = list(ray_origins.shape[:-1]) + [num_samples]
        u = torch.rand(noise_shape) if rand is None else rand
        depth_values = depth_values + u.to(ray_origins) * (far_thresh - near_thresh) / num_samples
    query_points = ray_origins[..., None, :] + ray_directions[..., None, :] * depth_values[..., :, None]
    return query_points, depth_values


def render_volume_density(radiance_field, ray_origins, depth_values):
    """tiny_nerf.py:68-107: sigma = relu(raw[3]) (no noise), rgb = sigmoid(raw[:3]), dists NOT scaled by the direction norm
    (volume_rendering_utils.py:19 does scale), last dist 1e10, alpha = 1 - exp(-sigma dists),
    weights = alpha * cumprod_exclusive(1 - alpha + 1e-10); rgb / depth / acc maps, no white background, no disparity."""
    sigma_a = torch.nn.functional.relu(radiance_field[..., 3])
    rgb = torch.sigmoid(radiance_field[..., :3])
    one_e_10 = torch.tensor([1e10], dtype=ray_origins.dtype, device=ray_origins.device)
    dists = torch.cat((depth_values[..., 1:] - depth_values[..., :-1], one_e_10.expand(depth_values[..., :1].shape)), dim=-1)
    alpha = 1.0 - torch.exp(-sigma_a * dists)
    weights = alpha * O.cumprod_exclusive(1.0 - alpha + 1e-10)
    rgb_map = (weights[..., None] * rgb).sum(dim=-2)
    depth_map = (weights * depth_values).sum(dim=-1)
    acc_map = weights.sum(-1)
    return rgb_map, depth_map, acc_map


def very_tiny_nerf_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """tiny_nerf.py:162-181 `VeryTinyNerfModel`: relu(layer1) -> relu(layer2) -> layer3 (ReLU AFTER layer1, unlike
    FlexibleNeRFModel, models.py:238); parameters layer{1,2,3}.{weight,bias}."""
    lin = torch.nn.functional.linear
    x = torch.relu(lin(x, sd["layer1.weight"], sd["layer1.bias"]))
    x = torch.relu(lin(x, sd["layer2.weight"], sd["layer2.bias"]))
    return lin(x, sd["layer3.weight"], sd["layer3.bias"])


def init_very_tiny_nerf(filter_size=128, num_encoding_functions=6, generator: Optional[torch.Generator] = None):
    """nn.Linear default init of tiny_nerf.py:166-175's three layers (3 + 6 L -> filter -> filter -> 4)."""
    sd = {}
    for name, fin, fout in (("layer1", 3 + 3 * 2 * num_encoding_functions, filter_size), ("layer2", filter_size, filter_size),
                            ("layer3", filter_size, 4)):
        bound = 1.0 / fin ** 0.5
        sd[name + ".weight"] = (torch.rand(fout, fin, generator=generator) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(fout, generator=generator) * 2 - 1) * bound
    return sd


def run_one_iter_of_tinynerf(height, width, focal_length, tform_cam2world, near_thresh, far_thresh, depth_samples_per_ray,
                             num_encoding_functions, chunksize, sd, rand: Optional[torch.Tensor] = None):
    """tiny_nerf.py:109-155: rays of every pixel -> jittered grid -> positional_encoding(points, L) -> model in chunks of
    `chunksize` POINTS (get_minibatches) -> compositing; returns the predicted (H, W, 3) image."""
    ray_origins, ray_directions = O.get_ray_bundle(height, width, focal_length, tform_cam2world)
    query_points, depth_values = compute_query_points_from_rays(ray_origins, ray_directions, near_thresh, far_thresh,
                                                                depth_samples_per_ray, rand=rand)
    flat = query_points.reshape((-1, 3))
    enc = O.positional_encoding(flat, num_encoding_functions)
    preds = [very_tiny_nerf_forward(sd, enc[i:i + chunksize]) for i in range(0, enc.shape[0], chunksize)]
    radiance_field = torch.cat(preds, dim=0).reshape(list(query_points.shape[:-1]) + [4])
    rgb_predicted, _, _ = render_volume_density(radiance_field, ray_origins, depth_values)
    return rgb_predicted
