"""Dev tool: cycle breakdown of the tcgen05 forward chain kernel (library built with make EXTRA=-DNERFB200_PROF)."""
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
lib = _lib.load()
buf = (C.c_ulonglong * 32)()
names = {0: "epi: wait bar_acc", 1: "epi: tcgen05.ld + wait (both chunks)", 2: "epi: chunk arithmetic + tcgen05.st issue (both chunks)",
         3: "epi: wait::st + fence + arrive", 5: "epi: whole layer event", 8: "mma: wait bar_a", 9: "mma: wait bar_full (all stages)",
         10: "mma: whole MMA (incl. waits)"}
for stash in (False, True):
    for _ in range(2):
        ops.mlp_fwd(arch, blob, rays, z, impl=1, want_stash=stash)
    lib.nerfb200_prof_read_fwd(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.mlp_fwd(arch, blob, rays, z, impl=1, want_stash=stash)
    e1.record()
    torch.cuda.synchronize()
    lib.nerfb200_prof_read_fwd(buf, 0)
    v = list(buf)
    ev, tiles, mmas = max(v[7], 1), max(v[6], 1), max(v[12], 1)
    print(f"forward (stash={stash}) {e0.elapsed_time(e1):.3f} ms; slot-0 tiles of CTA 0: {tiles}, layer events {ev}, MMAs (both slots) {mmas}")
    print(f"  epi: prologue (event 0)                  {v[4] / tiles:10.0f} cycles / tile")
    for k in sorted(names):
        den = mmas if k >= 8 else ev
        print(f"  {names[k]:40s} {v[k] / den:10.0f} cycles / {'MMA' if k >= 8 else 'event'}")
    print(f"  producer: wait bar_empty                 {v[11] / mmas:10.0f} cycles / MMA")
