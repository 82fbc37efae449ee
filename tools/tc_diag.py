"""tcgen05 vs fp32 CUDA-core kernels on the golden cases: max errors and fine-pass timings.  Diagnostic only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import Case
from nerf_pytorch_b200 import ops
from test_stage_parity_gpu import _arch

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

c = Case("lego_a0_train")
rays, _, aux = c.aux()
arch = _arch(c)
blob = ops.pack_weights(arch, ops.flatten_state_dict(arch, c.sd_c, "cuda"))
z = aux["z_coarse"].cuda().contiguous()
raw0, st0 = ops.mlp_fwd(arch, blob, rays.cuda(), z, impl=0, want_stash=True)
print("simt done", flush=True)
raw1, st1 = ops.mlp_fwd(arch, blob, rays.cuda(), z, impl=1, want_stash=True)
torch.cuda.synchronize()
print("tc done", flush=True)
want = aux["raw_coarse"]
print("scale", want.abs().max().item(), "err simt", (raw0.cpu()-want).abs().max().item(), "err tc", (raw1.cpu()-want).abs().max().item())
P = z.numel()
s0, s1 = st0.view(-1), st1.view(-1)
off = 0
for gi, n in enumerate([128, 128, 128, 128, 128, 64]):
    a, b = s0[off:off+P*n].view(P, n), s1[off:off+P*n].view(P, n)
    d = (a-b).abs()
    print(f" stash layer {gi}: scale {a.abs().max().item():.3e} maxdiff {d.max().item():.3e} bad rows {(d.max(1).values > 1e-3*a.abs().max()).sum().item()} bad cols {(d.max(0).values > 1e-3*a.abs().max()).sum().item()}")
    off += P*n
print("raw sample simt", raw0[0,0].tolist(), "tc", raw1[0,0].tolist())
# timing at bench size
N = 4096
for archname, kw in (("A0", dict(num_layers=4, hidden=128, skip_every=4)), ("A1", dict(num_layers=8, hidden=128, skip_every=3))):
    arch = ops.ArchSpec(n_freq_xyz=10, n_freq_dir=4, **kw)
    flat = torch.randn(arch.flat_param_count(), device="cuda") * 0.05
    blob = ops.pack_weights(arch, flat)
    d = torch.randn(N, 3, device="cuda"); d[:, 2] = -1
    o = torch.tensor([[0.0, -2.0, 3.4]], device="cuda").expand(N, 3)
    rays = torch.cat([o, d, torch.full((N, 1), 2.0, device="cuda"), torch.full((N, 1), 6.0, device="cuda"), d / d.norm(dim=-1, keepdim=True)], -1).contiguous()
    macs = {"A0": 83840, "A1": 165504}[archname]
    for S in (64, 192):
        z = torch.sort(torch.rand(N, S, device="cuda") * 4 + 2, -1).values.contiguous()
        for impl in (0, 1):
            t = timeit(lambda: ops.mlp_fwd(arch, blob, rays, z, impl=impl))
            ts = timeit(lambda: ops.mlp_fwd(arch, blob, rays, z, impl=impl, want_stash=True))
            print(f"{archname} S={S} impl={impl}: {t:.3f} ms ({2*macs*N*S/t/1e9:.1f} TFLOP/s useful), with stash {ts:.3f} ms", flush=True)
        r0 = ops.mlp_fwd(arch, blob, rays, z, impl=0); r1 = ops.mlp_fwd(arch, blob, rays, z, impl=1)
        print("   max diff", (r0-r1).abs().max().item(), "scale", r0.abs().max().item())

# cycle breakdown of the tcgen05 kernel (epilogue thread 0 of every CTA)
import ctypes
from nerf_pytorch_b200 import _lib
prof = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
_lib.load().nerfb200_debug_tc_profile(ctypes.c_void_p(prof.data_ptr()))
arch = ops.ArchSpec(n_freq_xyz=10, n_freq_dir=4, num_layers=4, hidden=128, skip_every=4)
blob = ops.pack_weights(arch, torch.randn(arch.flat_param_count(), device="cuda") * 0.05)
z = torch.sort(torch.rand(N, 192, device="cuda") * 4 + 2, -1).values.contiguous()
ops.mlp_fwd(arch, blob, rays, z, impl=1); torch.cuda.synchronize()
pr = prof.view(148, 8).double()[:, :4]
print("A0 S=192 per-CTA cycles: prologue %.0f  wait-mma %.0f  epilogue %.0f  total %.0f (tiles/CTA %.1f)" % (*pr.mean(0).tolist(), 4096*192/128/148))
_lib.load().nerfb200_debug_tc_profile(None)

# backward timing: SIMT vs (SIMT dgrad + TC wgrad)
for archname, kw in (("A0", dict(num_layers=4, hidden=128, skip_every=4)), ("A1", dict(num_layers=8, hidden=128, skip_every=3))):
    arch = ops.ArchSpec(n_freq_xyz=10, n_freq_dir=4, **kw)
    blob = ops.pack_weights(arch, torch.randn(arch.flat_param_count(), device="cuda") * 0.05)
    z = torch.sort(torch.rand(N, 192, device="cuda") * 4 + 2, -1).values.contiguous()
    raw, stash = ops.mlp_fwd(arch, blob, rays, z, impl=1, want_stash=True)
    G = torch.randn_like(raw)
    for impl in (0, 1):
        t = timeit(lambda: ops.mlp_bwd(arch, blob, rays, z, G, stash, impl=impl), n=5, warm=2)
        print(f"{archname} S=192 mlp_bwd impl={impl}: {t:.3f} ms", flush=True)
    g0, _ = ops.mlp_bwd(arch, blob, rays, z, G, stash, impl=0); g1, _ = ops.mlp_bwd(arch, blob, rays, z, G, stash, impl=1)
    print("   grad max diff", (g0-g1).abs().max().item(), "scale", g0.abs().max().item())
