"""Tiny driver for ncu: a few launches of the fused MLP forward (fine pass size) + one whole train step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nerf_pytorch_b200 import ops
impl = int(os.environ.get("IMPL", "0")); archname = os.environ.get("ARCH", "A1")
kw = {"A0": dict(num_layers=4, hidden=128, skip_every=4), "A1": dict(num_layers=8, hidden=128, skip_every=3),
      "A2": dict(num_layers=8, hidden=256, skip_every=4)}[archname]
arch = ops.ArchSpec(n_freq_xyz=10, n_freq_dir=4, **kw)
N, S = int(os.environ.get("NRAYS", "4096")), 192
torch.manual_seed(0)
flat = torch.randn(arch.flat_param_count(), device="cuda") * 0.05
blob = ops.pack_weights(arch, flat)
d = torch.randn(N, 3, device="cuda"); d[:, 2] = -1
o = torch.tensor([[0.0, -2.0, 3.4]], device="cuda").expand(N, 3)
rays = torch.cat([o, d, torch.full((N, 1), 2.0, device="cuda"), torch.full((N, 1), 6.0, device="cuda"), d / d.norm(dim=-1, keepdim=True)], -1).contiguous()
z = torch.sort(torch.rand(N, S, device="cuda") * 4 + 2, -1).values.contiguous()
if os.environ.get("STASH", "1") == "0":
    for _ in range(3):
        ops.mlp_fwd(arch, blob, rays, z, impl=impl)
for _ in range(3):
    raw, stash = ops.mlp_fwd(arch, blob, rays, z, impl=impl, want_stash=True)
G = torch.randn_like(raw)
if os.environ.get("BWD", "1") == "1":
    for _ in range(2):
        ops.mlp_bwd(arch, blob, rays, z, G, stash, impl=impl)
torch.cuda.synchronize()
