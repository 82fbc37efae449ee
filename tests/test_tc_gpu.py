"""GPU: the tcgen05 kernels against the fp32 CUDA-core kernels and the oracle.

The three-term splits (fp16x2 with a scaled residual in the chain kernels, 3xTF32 in wgrad) keep ~21 bits of every
product (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulate), so the bar is the same as for the fp32 path:
relative 1e-4 of the output scale, per north_star."""
import pytest
import torch

from helpers import CASES, Case, frac_close
from test_stage_parity_gpu import _arch

pytestmark = pytest.mark.gpu

TC_CASES = ["lego_a0_train", "fern_a0_ndc", "a1_skip_lindisp", "a0_noview_coarse_only"]


@pytest.mark.parametrize("name", TC_CASES)
def test_tc_forward_matches_simt_and_oracle(name):
    from nerf_pytorch_b200 import ops

    c = Case(name)
    rays, _, aux = c.aux()
    arch = _arch(c)
    for tag, sd in (("coarse", c.sd_c), ("fine", c.sd_f)):
        if sd is None:
            continue
        blob = ops.pack_weights(arch, ops.flatten_state_dict(arch, sd, "cuda"))
        z = aux["z_" + tag].cuda().contiguous()
        raw0, st0 = ops.mlp_fwd(arch, blob, rays.cuda(), z, impl=ops.IMPL_SIMT, want_stash=True)
        raw1, st1 = ops.mlp_fwd(arch, blob, rays.cuda(), z, impl=ops.IMPL_TC, want_stash=True)
        torch.cuda.synchronize()
        want = aux["raw_" + tag]
        scale = want.abs().max().item()
        e_tc = (raw1.cpu().double() - want.double()).abs().max().item()
        e_simt = (raw0.cpu().double() - want.double()).abs().max().item()
        assert e_tc <= 1e-4 * scale + 1e-5, (tag, e_tc, e_simt, scale)
        assert frac_close(raw1.cpu(), want, rtol=1e-4, atol=1e-5 * max(1.0, scale)) > 0.999, tag
        # hidden activations + encodings (the stash the backward consumes) agree with the fp32 kernel's,
        # and the ReLU bit masks (last section of the stash) agree except where an activation is ~0
        P = z.numel()
        mask_words = sum(o // 32 for _, _, _, _, o in arch.flat_layout() if o >= 64)
        n_float = st0.numel() - P * mask_words
        s_scale = st0[:n_float].abs().max().item()
        assert (st1[:n_float] - st0[:n_float]).abs().max().item() <= 1e-4 * s_scale + 1e-6, tag
        m0, m1 = st0[n_float:].view(torch.int32), st1[n_float:].view(torch.int32)
        flips = (m0 ^ m1).ne(0).float().mean().item()
        assert flips < 1e-3, (tag, flips)


def test_tc_ragged_tail_and_sizes():
    """Tiles that straddle rays / a ragged last tile / 1 ray: same results as the fp32 kernel."""
    from nerf_pytorch_b200 import ops

    c = Case("lego_a0_train")
    arch = _arch(c)
    blob = ops.pack_weights(arch, ops.flatten_state_dict(arch, c.sd_f, "cuda"))
    g = torch.Generator().manual_seed(5)
    for n, s in ((1, 64), (3, 192), (7, 50), (33, 17), (129, 100)):
        d = torch.randn(n, 3, generator=g)
        rays = torch.cat([torch.randn(n, 3, generator=g) * 0.1 + torch.tensor([0.0, -2.0, 3.0]), d,
                          torch.full((n, 1), 2.0), torch.full((n, 1), 6.0), d / d.norm(dim=-1, keepdim=True)], -1).cuda()
        z = torch.sort(torch.rand(n, s, generator=g) * 4 + 2, -1).values.cuda().contiguous()
        r0 = ops.mlp_fwd(arch, blob, rays.contiguous(), z, impl=ops.IMPL_SIMT)
        r1 = ops.mlp_fwd(arch, blob, rays.contiguous(), z, impl=ops.IMPL_TC)
        scale = r0.abs().max().item()
        assert (r1 - r0).abs().max().item() <= 1e-4 * scale + 1e-5, (n, s, (r1 - r0).abs().max().item(), scale)


def test_tc_backward_ragged_sizes():
    """Point counts that are not multiples of the 32-point wgrad stage / the 128-point tile (partial bulk copies,
    partial last tile): dgrad + wgrad on tcgen05 vs the fp32 CUDA-core kernels."""
    from nerf_pytorch_b200 import ops

    for case in ("lego_a0_train", "a1_skip_lindisp"):
        c = Case(case)
        arch = _arch(c)
        blob = ops.pack_weights(arch, ops.flatten_state_dict(arch, c.sd_f, "cuda"))
        g = torch.Generator().manual_seed(7)
        for n, s in ((1, 16), (7, 50), (33, 17), (5, 37), (129, 100)):
            d = torch.randn(n, 3, generator=g)
            rays = torch.cat([torch.randn(n, 3, generator=g) * 0.1 + torch.tensor([0.0, -2.0, 3.0]), d,
                              torch.full((n, 1), 2.0), torch.full((n, 1), 6.0), d / d.norm(dim=-1, keepdim=True)],
                             -1).cuda().contiguous()
            z = torch.sort(torch.rand(n, s, generator=g) * 4 + 2, -1).values.cuda().contiguous()
            G = torch.randn(n, s, 4, generator=g).cuda()
            _, stash = ops.mlp_fwd(arch, blob, rays, z, impl=ops.IMPL_TC, want_stash=True)
            g0, _ = ops.mlp_bwd(arch, blob, rays, z, G, stash, impl=ops.IMPL_SIMT)
            g1, _ = ops.mlp_bwd(arch, blob, rays, z, G, stash, impl=ops.IMPL_TC)
            for lname, w_off, b_off, fin, fout in arch.flat_layout():
                for off, cnt, what in ((w_off, fin * fout, "weight"), (b_off, fout, "bias")):
                    a, b = g0[off:off + cnt], g1[off:off + cnt]
                    scale = a.abs().max().item() + 1e-30
                    assert (a - b).abs().max().item() <= 1e-4 * scale, (case, n, s, lname, what,
                                                                        (a - b).abs().max().item(), scale)


def test_tc_unsupported_hidden_256_is_refused():
    from nerf_pytorch_b200 import ops

    arch = ops.ArchSpec(num_layers=8, hidden=256, skip_every=4, n_freq_xyz=10)
    blob = ops.pack_weights(arch, torch.zeros(arch.flat_param_count(), device="cuda"))
    rays = torch.zeros(4, 11, device="cuda")
    z = torch.ones(4, 64, device="cuda")
    with pytest.raises(NotImplementedError):
        ops.mlp_fwd(arch, blob, rays, z, impl=ops.IMPL_TC)


@pytest.mark.parametrize("name", ["lego_a0_train", "a1_skip_lindisp"])
def test_tc_end_to_end_against_reference_golden(name):
    import nerf_pytorch_b200 as nb
    from test_render_parity_gpu import build_models

    c = Case(name)
    mc, mf, epf, edf = build_models(c)
    rnd = {k: v.cuda() for k, v in c.randoms.items()}
    with torch.no_grad():
        out = nb.run_one_iter_of_nerf(c.H, c.W, c.focal, mc, mf, c.ro.cuda(), c.rd.cuda(), c.options, mode=c.mode,
                                      encode_position_fn=epf, encode_direction_fn=edf, randoms=rnd, impl=1)
    for k in (0, 2, 3, 5):
        assert frac_close(out[k].cpu(), c.outputs[k], rtol=1e-4, atol=2e-5) > 0.97, k


@pytest.mark.parametrize("name", TC_CASES)
def test_tc_wgrad_matches_fp32_kernel(name):
    """Weight gradients from the tcgen05 wgrad kernel (3xTF32) vs the fp32 CUDA-core kernel, same dY / stash."""
    from nerf_pytorch_b200 import ops

    c = Case(name)
    rays, _, aux = c.aux()
    arch = _arch(c)
    blob = ops.pack_weights(arch, ops.flatten_state_dict(arch, c.sd_c, "cuda"))
    z = aux["z_coarse"].cuda().contiguous()
    gen = torch.Generator().manual_seed(11)
    G = torch.randn(z.shape[0], z.shape[1], 4, generator=gen).cuda()
    raw, stash = ops.mlp_fwd(arch, blob, rays.cuda(), z, want_stash=True)
    g0, gs0 = ops.mlp_bwd(arch, blob, rays.cuda(), z, G, stash, impl=ops.IMPL_SIMT)
    g1, gs1 = ops.mlp_bwd(arch, blob, rays.cuda(), z, G, stash, impl=ops.IMPL_TC)
    torch.cuda.synchronize()
    # dgrad chain (tcgen05) vs fp32 kernel: per-layer pre-activation gradients
    n_act = sum(o for _, _, _, _, o in arch.flat_layout() if o >= 64) * z.numel()
    gscale = gs0[:n_act].abs().max().item()
    assert (gs1[:n_act] - gs0[:n_act]).abs().max().item() <= 1e-4 * gscale, ((gs1[:n_act] - gs0[:n_act]).abs().max().item(), gscale)
    for lname, w_off, b_off, fin, fout in arch.flat_layout():
        for off, n, what in ((w_off, fin * fout, "weight"), (b_off, fout, "bias")):
            a, b = g0[off:off + n], g1[off:off + n]
            scale = a.abs().max().item() + 1e-30
            assert (a - b).abs().max().item() <= 1e-4 * scale, (lname, what, (a - b).abs().max().item(), scale)
