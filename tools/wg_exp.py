"""Fine-pass timing of the backward MLP kernels (dgrad + wgrad, 4096 x 192 points, A1).  Diagnostic only."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nerf_pytorch_b200 import ops, _lib
lib = _lib.load()
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N, S = 4096, 192
arch = ops.ArchSpec(n_freq_xyz=10, n_freq_dir=4, num_layers=8, hidden=128, skip_every=3)
blob = ops.pack_weights(arch, torch.randn(arch.flat_param_count(), device="cuda") * 0.05)
d = torch.randn(N, 3, device="cuda"); d[:, 2] = -1
o = torch.tensor([[0.0, -2.0, 3.4]], device="cuda").expand(N, 3)
rays = torch.cat([o, d, torch.full((N, 1), 2.0, device="cuda"), torch.full((N, 1), 6.0, device="cuda"), d / d.norm(dim=-1, keepdim=True)], -1).contiguous()
z = torch.sort(torch.rand(N, S, device="cuda") * 4 + 2, -1).values.contiguous()
raw, stash = ops.mlp_fwd(arch, blob, rays, z, impl=1, want_stash=True)
G = torch.randn_like(raw)
t = timeit(lambda: ops.mlp_bwd(arch, blob, rays, z, G, stash, impl=1), n=10, warm=3)
print(f"A1 bwd (dgrad + wgrad), 4096x192 points: {t:.3f} ms")
