"""Dev tool: cycle breakdown of the fused tcgen05 backward (library built with make EXTRA=-DNERFB200_PROF)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerf_pytorch_b200 import _lib, ops  # noqa: E402

arch = ops.ArchSpec(num_layers=8, hidden=128, skip_every=3, n_freq_xyz=10) if os.environ.get("ARCH", "A1") == "A1" else ops.ArchSpec(n_freq_xyz=10)
torch.manual_seed(0)
flat = (torch.rand(arch.flat_param_count(), device="cuda") - 0.5) * 0.2
blob = ops.pack_weights(arch, flat)
n, s = 4096, 192
d = torch.randn(n, 3, device="cuda")
rays = torch.cat([torch.randn(n, 3, device="cuda") * 0.1, d, torch.full((n, 1), 2.0, device="cuda"),
                  torch.full((n, 1), 6.0, device="cuda"), d / d.norm(dim=-1, keepdim=True)], -1).contiguous()
z = torch.sort(torch.rand(n, s, device="cuda") * 4 + 2, -1).values.contiguous()
raw, stash = ops.mlp_fwd(arch, blob, rays, z, impl=1, want_stash=True)
G = torch.randn_like(raw)
lib = _lib.load()
buf = (C.c_ulonglong * 32)()
for _ in range(2):
    ops.mlp_bwd(arch, blob, rays, z, G, stash, impl=1)
lib.nerfb200_prof_read(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.mlp_bwd(arch, blob, rays, z, G, stash, impl=1)
e1.record()
torch.cuda.synchronize()
lib.nerfb200_prof_read(buf, 0)
v = list(buf)
ev = max(v[7], 1)
names = {0: "epi: wait bar_acc (chain done)", 1: "epi: part A (acc -> hi/lo registers)", 4: "epi: hi/lo -> tensor memory + arrive bar_a",
         2: "epi: wait job_done (G tile free)", 5: "epi: part B (registers -> G tile)", 3: "drain: wait job_done", 6: "drain: chunk loops",
         8: "mma: wait bar_a", 9: "mma: wait w_full", 10: "mma: wait bar_g", 11: "mma: wait xl_full", 12: "mma: wait xh_full",
         13: "mma: wait acc_free", 14: "mma: chain section (incl. waits)", 15: "mma: jobs section (incl. waits)",
         16: "xprod: wait xl_free", 17: "xprod: wait xh_free", 18: "wprod: wait w_empty"}
print(f"kernel+unpack {e0.elapsed_time(e1):.3f} ms; events of CTA 0: {ev}")
for k in sorted(names):
    print(f"  {names[k]:36s} {v[k] / ev:10.0f} cycles / event")
