"""Diagnostic timings of every stage at the BASELINE size (4096 rays, 64+128 samples).  Not a bench."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import Case
import nerf_pytorch_b200 as nb
from nerf_pytorch_b200 import ops

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main():
    print(torch.cuda.get_device_name(0))
    c = Case("lego_a0_train")
    N, NC, NF = 4096, 64, 128
    for archname, kw in (("A0", dict(num_layers=4, hidden=128, skip_every=4)), ("A1", dict(num_layers=8, hidden=128, skip_every=3)),
                         ("A2", dict(num_layers=8, hidden=256, skip_every=4))):
        arch = ops.ArchSpec(n_freq_xyz=10, n_freq_dir=4, **kw)
        flat = torch.randn(arch.flat_param_count(), device="cuda") * 0.05
        if archname == "A0":
            flat = ops.flatten_state_dict(arch, c.sd_f, "cuda")
        blob = ops.pack_weights(arch, flat)
        g = torch.Generator(device="cuda").manual_seed(0)
        d = torch.randn(N, 3, device="cuda", generator=g); d[:, 2] = -1
        o = torch.tensor([[0.0, -2.0, 3.4]], device="cuda").expand(N, 3)
        vd = d / d.norm(dim=-1, keepdim=True)
        rays = torch.cat([o, d, torch.full((N, 1), 2.0, device="cuda"), torch.full((N, 1), 6.0, device="cuda"), vd], -1).contiguous()
        for S in (NC, NC + NF):
            z = torch.sort(torch.rand(N, S, device="cuda", generator=g) * 4 + 2, -1).values.contiguous()
            t = timeit(lambda: ops.mlp_fwd(arch, blob, rays, z))
            macs = {"A0": 83840, "A1": 165504, "A2": 593408}[archname]
            print(f"{archname} mlp_fwd simt S={S}: {t:.3f} ms  -> {2*macs*N*S/t/1e9:.2f} TFLOP/s")
            raw, stash = ops.mlp_fwd(arch, blob, rays, z, want_stash=True)
            t = timeit(lambda: ops.mlp_fwd(arch, blob, rays, z, want_stash=True))
            print(f"{archname} mlp_fwd+stash S={S}: {t:.3f} ms")
            G = torch.randn(N, S, 4, device="cuda")
            t = timeit(lambda: ops.mlp_bwd(arch, blob, rays, z, G, stash))
            print(f"{archname} mlp_bwd simt S={S}: {t:.3f} ms -> {4*macs*N*S/t/1e9:.2f} TFLOP/s")
            noise = torch.randn(N, S, device="cuda")
            t = timeit(lambda: ops.composite_fwd(raw, z, rays, noise, 0.2, False))
            print(f"composite_fwd S={S}: {t*1e3:.1f} us")
            gout = torch.randn(N, 8, device="cuda")
            t = timeit(lambda: ops.composite_bwd(raw, z, rays, noise, gout, 0.2, False))
            print(f"composite_bwd S={S}: {t*1e3:.1f} us")
        zc = torch.sort(torch.rand(N, NC, device="cuda") * 4 + 2, -1).values.contiguous()
        w = torch.rand(N, NC, device="cuda"); u = torch.rand(N, NF, device="cuda")
        t = timeit(lambda: ops.sample_pdf_merge(zc, w, u, NF))
        print(f"sample_pdf_merge: {t*1e3:.1f} us")
        del stash
    # whole path through the API
    def mk(sd):
        m = nb.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4); m.load_state_dict(sd); return m.cuda()
    mc, mf = mk(c.sd_c), mk(c.sd_f)
    epf, edf = nb.get_embedding_function(10), nb.get_embedding_function(4)
    from oracle import nerf_oracle as O
    opt = O.make_options(num_coarse=64, num_fine=128)
    ro, rd = rays[:, :3].contiguous(), rays[:, 3:6].contiguous()
    tgt = torch.rand(N, 3, device="cuda")
    def fwd():
        with torch.no_grad():
            return nb.run_one_iter_of_nerf(400, 400, 555.5, mc, mf, ro, rd, opt, encode_position_fn=epf, encode_direction_fn=edf)
    def step():
        out = nb.run_one_iter_of_nerf(400, 400, 555.5, mc, mf, ro, rd, opt, encode_position_fn=epf, encode_direction_fn=edf)
        loss = torch.nn.functional.mse_loss(out[0], tgt) + torch.nn.functional.mse_loss(out[3], tgt)
        mc.zero_grad(); mf.zero_grad(); loss.backward()
    t = timeit(fwd); print(f"API forward (no_grad) 4096 rays: {t:.3f} ms -> {N/t*1e3:.0f} rays/s")
    t = timeit(step); print(f"API fwd+bwd 4096 rays: {t:.3f} ms -> {N/t*1e3:.0f} rays/s")
    t0 = time.time(); 
    for _ in range(10): step()
    torch.cuda.synchronize(); print(f"wall fwd+bwd: {(time.time()-t0)*100:.3f} ms/step")

if __name__ == "__main__":
    main()
