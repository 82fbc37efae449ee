import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from helpers import Case
from nerf_pytorch_b200 import ops
from oracle import nerf_oracle as O
from test_stage_parity_gpu import _arch

name = "a0_noview_coarse_only"
c = Case(name)
rays, _, aux = c.aux()
arch = _arch(c)
sd = c.sd_c
z = aux["z_coarse"]
gen = torch.Generator().manual_seed(11)
G = torch.randn(z.shape[0], z.shape[1], 4, generator=gen)
P = z.numel()
pts = (rays[:, None, :3] + rays[:, None, 3:6] * z[..., None]).reshape(-1, 3)
enc = O.positional_encoding(pts.double(), *c.enc_xyz)
sd64 = {k: v.double() for k, v in sd.items()}
pre = []
h = F.linear(enc, sd64["layer1.weight"], sd64["layer1.bias"]); h.requires_grad_(True); h.retain_grad(); pre.append(h)
x = h
for i in range(3):
    y = F.linear(x, sd64[f"layers_xyz.{i}.weight"], sd64[f"layers_xyz.{i}.bias"]); y.retain_grad(); pre.append(y)
    x = F.relu(y)
out = F.linear(x, sd64["fc_out.weight"], sd64["fc_out.bias"])
(out * G.reshape(-1, 4).double()).sum().backward()
flat = ops.flatten_state_dict(arch, sd, "cuda")
blob = ops.pack_weights(arch, flat)
raw, stash = ops.mlp_fwd(arch, blob, rays.cuda(), z.cuda().contiguous(), want_stash=True)
flat_grad, gst = ops.mlp_bwd(arch, blob, rays.cuda(), z.cuda().contiguous(), G.cuda().contiguous(), stash)
stash, gst = stash.cpu().view(4, P, 128), gst.cpu().view(4, P, 128)
print("raw err", (raw.cpu().double().reshape(-1,4) - out.detach()).abs().max().item())
for l in range(4):
    act = pre[l].detach() if l == 0 else F.relu(pre[l].detach())
    print(f"layer {l}: stash err {(stash[l].double()-act).abs().max().item():.2e}", end="  ")
    want = pre[l].grad
    d = (gst[l].double() - want).abs()
    bad = d > 1e-4 * want.abs().max()
    print(f"dY err max {d.max().item():.2e} (scale {want.abs().max().item():.2e}) bad frac {bad.double().mean().item():.4f}")
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("   bad rows:", rows[:20].tolist(), "n", len(rows), " bad cols:", cols[:20].tolist(), "n", len(cols))
        r, cc = bad.nonzero()[0].tolist()
        print("   example", r, cc, "got", gst[l][r, cc].item(), "want", want[r, cc].item(), "stash", stash[l][r, cc].item(), "pre64", pre[l][r, cc].item())
