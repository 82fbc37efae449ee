import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import Case
from nerf_pytorch_b200 import ops
from oracle import nerf_oracle as O
from test_stage_parity_gpu import _arch

for name in ["lego_a0_train", "a1_skip_lindisp", "a0_noview_coarse_only"]:
    c = Case(name)
    rays, _, aux = c.aux()
    arch = _arch(c)
    sd = c.sd_c
    z = aux["z_coarse"]
    gen = torch.Generator().manual_seed(11)
    G = torch.randn(z.shape[0], z.shape[1], 4, generator=gen)
    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    pts = rays[:, None, :3].double() + rays[:, None, 3:6].double() * z[..., None].double()
    raw64 = O.run_network(sd64, pts, rays.double(), 1 << 20, c.enc_xyz, c.enc_dir if c.use_viewdirs else None)
    (raw64 * G.double()).sum().backward()
    flat = ops.flatten_state_dict(arch, sd, "cuda")
    blob = ops.pack_weights(arch, flat)
    raw, stash = ops.mlp_fwd(arch, blob, rays.cuda(), z.cuda().contiguous(), want_stash=True)
    flat_grad, gst = ops.mlp_bwd(arch, blob, rays.cuda(), z.cuda().contiguous(), G.cuda().contiguous(), stash)
    flat_grad = flat_grad.cpu()
    print("==", name, "raw err", (raw.cpu().double() - raw64.detach()).abs().max().item())
    for lname, w_off, b_off, fin, fout in arch.flat_layout():
        gw = flat_grad[w_off:w_off + fin * fout].view(fout, fin)
        gb = flat_grad[b_off:b_off + fout]
        want = sd64[lname + ".weight"].grad
        d = (gw.double() - want).abs()
        i = d.argmax().item()
        colerr = d.max(0).values
        print(f"  {lname:14s} W rel {d.max().item()/(want.abs().max().item()+1e-30):.2e} at (n={i//fin},k={i%fin}) "
              f"bad cols>1e-3*max: {(colerr > 1e-3*want.abs().max()).nonzero().flatten().tolist()[:12]} "
              f"b rel {(gb.double()-sd64[lname+'.bias'].grad).abs().max().item()/(sd64[lname+'.bias'].grad.abs().max().item()+1e-30):.2e}")
