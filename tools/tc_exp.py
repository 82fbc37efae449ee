"""Timing experiments on the tcgen05 chain kernels: per-CTA cycle counters (nerfb200_debug_tc_profile) and the
debug switches of nerfb200_debug_tc_flags (weight copies off / MMAs off).  Diagnostic only, not a bench."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nerf_pytorch_b200 import ops, _lib
lib = _lib.load()
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N, S = 4096, 192
arch = ops.ArchSpec(n_freq_xyz=10, n_freq_dir=4, num_layers=4, hidden=128, skip_every=4)
blob = ops.pack_weights(arch, torch.randn(arch.flat_param_count(), device="cuda") * 0.05)
d = torch.randn(N, 3, device="cuda"); d[:, 2] = -1
o = torch.tensor([[0.0, -2.0, 3.4]], device="cuda").expand(N, 3)
rays = torch.cat([o, d, torch.full((N, 1), 2.0, device="cuda"), torch.full((N, 1), 6.0, device="cuda"), d / d.norm(dim=-1, keepdim=True)], -1).contiguous()
z = torch.sort(torch.rand(N, S, device="cuda") * 4 + 2, -1).values.contiguous()
prof = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
lib.nerfb200_debug_tc_profile(ctypes.c_void_p(prof.data_ptr()))
for flags, name, st in ((0, "normal", False), (2, "no MMAs", False), (0, "normal+stash", True), (2, "no MMAs+stash", True)):
    lib.nerfb200_debug_tc_flags(flags)
    t = timeit(lambda: ops.mlp_fwd(arch, blob, rays, z, impl=1, want_stash=st))
    pr = prof.view(148, 8).double().mean(0) / 41.5
    print(f"A0 fwd S=192 [{name}]: {t:.3f} ms  per tile: prologue {pr[0]:.0f} wait-mma {pr[1]:.0f} epilogue {pr[2]:.0f} total {pr[3]:.0f} | per layer: tmem-ld {pr[4]/6:.0f} chunks {pr[5]/6:.0f} st-wait {pr[6]/6:.0f} head/bar {pr[7]/6:.0f}")
lib.nerfb200_debug_tc_flags(0)

# dgrad breakdown (A1, fine pass)
arch1 = ops.ArchSpec(n_freq_xyz=10, n_freq_dir=4, num_layers=8, hidden=128, skip_every=3)
blob1 = ops.pack_weights(arch1, torch.randn(arch1.flat_param_count(), device="cuda") * 0.05)
raw, stash = ops.mlp_fwd(arch1, blob1, rays, z, impl=1, want_stash=True)
pr = prof.view(148, 8).double().mean(0) / 41.5
print(f"A1 fwd+stash per tile: prologue {pr[0]:.0f} wait-mma {pr[1]:.0f} epilogue {pr[2]:.0f} total {pr[3]:.0f} | per layer(10): tmem-ld {pr[4]/10:.0f} chunks {pr[5]/10:.0f} st-wait {pr[6]/10:.0f} head/bar {pr[7]/10:.0f}")
G = torch.randn_like(raw)
for flags in (0, 2):
    lib.nerfb200_debug_tc_flags(flags)
    t = timeit(lambda: ops.mlp_bwd(arch1, blob1, rays, z, G, stash, impl=1), n=5)
    pr = prof.view(148, 8).double().mean(0) / 41.5
    print(f"A1 bwd flags={flags}: {t:.3f} ms; dgrad per tile: prologue {pr[0]:.0f} wait-mma {pr[1]:.0f} epilogue {pr[2]:.0f} total {pr[3]:.0f} | per layer(10): tmem-ld {pr[4]/10:.0f} chunks {pr[5]/10:.0f} st-wait {pr[6]/10:.0f} head/bar {pr[7]/10:.0f}")
lib.nerfb200_debug_tc_flags(0)
