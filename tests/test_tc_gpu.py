"""GPU: the tcgen05 kernels against the oracle (fp64 where the comparison is a gradient) and the fp32 CUDA-core kernels.

The three-term fp16x2 splits (x = hi + lo 2^-11, x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulate) keep ~22 bits of
every product, so the bar is the fp32 path's: the MEASURED error is ~5e-6 of the tensor's scale on the shipped
checkpoints; the asserts below are written against that (with headroom for the sum over 128..320 terms), far inside the
1e-4 relative of north_star, and the worst element is bounded, not just a quantile."""
import pytest
import torch

from helpers import CASES, Case, frac_close
from test_stage_parity_gpu import _arch

pytestmark = pytest.mark.gpu

TC_CASES = ["lego_a0_train", "fern_a0_ndc", "a1_skip_lindisp", "a0_noview_coarse_only"]
TC_TRAIN_CASES = ["lego_a0_train", "fern_a0_ndc", "a1_skip_lindisp"]  # the fused backward needs the view-dependent heads


def gemm_layers(arch):
    """(name, out_features, cum_n) of the gemm layers in the library's order (heads skipped)."""
    out, cum = [], 0
    for name, _, _, _, fout in arch.flat_layout():
        if fout >= 64:
            out.append((name, fout, cum))
            cum += fout
    return out, cum


def decode_tc_stash_layer(stash, n_points, cum_n, n):
    """Operand tiles of the tcgen05 forward (csrc/tc_common.cuh "Operand tiles") -> activations [n_points][n] (fp64)."""
    tiles = (n_points + 127) // 128
    p_pad = tiles * 128
    raw = stash.view(torch.int16)[2 * p_pad * cum_n: 2 * p_pad * (cum_n + n)]
    t = raw.view(tiles, 2, 16, n // 8, 8, 8).view(torch.float16).double()     # [tile][hi|lo][pb][fb][p%8][f%8]
    val = (t[:, 0] + t[:, 1] / 2048.0) * 16.0
    return val.permute(0, 1, 3, 2, 4).reshape(p_pad, n)[:n_points]


def decode_simt_stash_layer(stash, n_points, cum_n, n):
    """fp32 rows with 16-byte chunks XOR-swizzled by (point & 7) inside each 128-byte segment (csrc/common.cuh swz_col)."""
    x = stash[n_points * cum_n: n_points * (cum_n + n)].view(n_points, n // 32, 8, 4)
    pt = torch.arange(n_points, device=stash.device) & 7
    q = torch.arange(8, device=stash.device)
    idx = (q[None, :] ^ pt[:, None])                                   # stored position of logical chunk q
    return torch.gather(x, 2, idx[:, None, :, None].expand(n_points, n // 32, 8, 4)).reshape(n_points, n).double()


def oracle_mlp_grads(c, sd, rays, z, G):
    """fp64 autograd of the oracle MLP: d(sum(raw * G))/d(params)."""
    from oracle import nerf_oracle as O

    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    pts = (rays[:, None, :3] + rays[:, None, 3:6] * z[..., None]).double()  # fp32 points, then promoted (train_utils.py:67)
    raw64 = O.run_network(sd64, pts, rays.double(), 1 << 20, c.enc_xyz, c.enc_dir if c.use_viewdirs else None)
    (raw64 * G.double()).sum().backward()
    return {k: v.grad for k, v in sd64.items()}


def oracle_mlp_acts(c, sd, rays, z):
    """fp64 activations of every gemm layer of the oracle MLP, in the library's gemm order: [P][n] each."""
    from oracle import nerf_oracle as O

    sd64 = {k: v.double() for k, v in sd.items()}
    pts = (rays[:, None, :3] + rays[:, None, 3:6] * z[..., None]).double().reshape(-1, 3)
    emb = O.positional_encoding(pts, *c.enc_xyz)
    dim_xyz = emb.shape[-1]
    if c.use_viewdirs:
        vd = rays[:, None, -3:].double().expand(z.shape[0], z.shape[1], 3).reshape(-1, 3)
        emb = torch.cat((emb, O.positional_encoding(vd, *c.enc_dir)), -1)
    _, acts = O.flexible_nerf_forward(sd64, emb, dim_xyz, return_acts=True)
    out, i = [acts["h0"]], 1
    while f"h{i}" in acts:
        out.append(acts[f"h{i}"])
        i += 1
    if c.use_viewdirs:
        out += [acts["feat"], acts["d"]]
    return out


def tc_mask_view(arch, stash, n_points, layer_index, layers, cum_total):
    """The ReLU bit-mask words of gemm layer `layer_index` inside a tcgen05 forward stash: int32 [P][n/32]."""
    p_pad = (n_points + 127) // 128 * 128
    enc_tile_w = (arch.dim_xyz + 15) & ~15
    dir_pad = (arch.dim_dir + 7) & ~7
    mask_base = cum_total + enc_tile_w + dir_pad
    mask_cum = sum(n // 32 for _, n, _ in layers[:layer_index])
    n = layers[layer_index][1]
    start = p_pad * (mask_base + mask_cum)
    return stash.view(torch.int32)[start:start + n_points * (n // 32)].view(n_points, n // 32)


def pack_mask(act):
    """bool [P][n] -> int32 words [P][n/32], bit j of word w = column 32 w + j."""
    P, n = act.shape
    b = act.view(P, n // 32, 32).to(torch.int64)
    w = (b << torch.arange(32, device=act.device, dtype=torch.int64)).sum(-1)
    return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)


@pytest.mark.parametrize("name", TC_CASES)
def test_tc_forward_matches_simt_and_oracle(name):
    from nerf_pytorch_b200 import ops

    c = Case(name)
    rays, _, aux = c.aux()
    arch = _arch(c)
    layers, _ = gemm_layers(arch)
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
        # worst element: 3e-5 of the output scale (measured ~5e-6; the fp32 reference itself is ~2e-6 from fp64)
        assert e_tc <= 3e-5 * scale + 1e-6, (tag, e_tc, e_simt, scale)
        assert frac_close(raw1.cpu(), want, rtol=1e-4, atol=1e-5 * max(1.0, scale)) > 0.9999, tag
        # the stash the backward consumes: every layer's activation tile decodes to the fp32 kernel's activations
        P = z.numel()
        for lname, n, cum in layers:
            a0 = decode_simt_stash_layer(st0, P, cum, n)
            a1 = decode_tc_stash_layer(st1, P, cum, n)
            s = a0.abs().max().item()
            assert (a1 - a0).abs().max().item() <= 3e-5 * s + 1e-7, (tag, lname, (a1 - a0).abs().max().item(), s)


@pytest.mark.parametrize("viewdirs", [True, False])
def test_tc_forward_hidden_256_matches_oracle_and_fp32_kernel(viewdirs):
    """pretrained/*/config.yml as written (8 x 256, skip 4): the tcgen05 inference forward (one tile in flight, N = 256 MMAs)
    against the fp64 oracle MLP and the fp32 CUDA-core kernel, ragged tile tail included."""
    from nerf_pytorch_b200 import ops
    from oracle import nerf_oracle as O

    arch = ops.ArchSpec(num_layers=8, hidden=256, skip_every=4, n_freq_xyz=10, n_freq_dir=4, use_viewdirs=viewdirs)
    g = torch.Generator().manual_seed(3)
    sd = O.init_flexible_nerf(8, 256, 4, 10, 4, use_viewdirs=viewdirs, generator=g)
    n, s_ = 37, 52   # 1924 points: 15 full tiles + a ragged one, tiles straddle rays
    d = torch.randn(n, 3, generator=g)
    rays = torch.cat([torch.randn(n, 3, generator=g) * 0.3, d, torch.full((n, 1), 2.0), torch.full((n, 1), 6.0),
                      d / d.norm(dim=-1, keepdim=True)], -1)
    rays = (rays if viewdirs else rays[:, :8]).contiguous()
    z = torch.sort(torch.rand(n, s_, generator=g) * 4 + 2, -1).values.contiguous()
    blob = ops.pack_weights(arch, ops.flatten_state_dict(arch, sd, "cuda"))
    raw0 = ops.mlp_fwd(arch, blob, rays.cuda(), z.cuda(), impl=ops.IMPL_SIMT)
    raw1 = ops.mlp_fwd(arch, blob, rays.cuda(), z.cuda(), impl=ops.IMPL_TC)
    torch.cuda.synchronize()
    sd64 = {k: v.double() for k, v in sd.items()}
    pts = (rays[:, None, :3] + rays[:, None, 3:6] * z[..., None]).double()
    want = O.run_network(sd64, pts, rays.double(), 1 << 20, (10, True, True), (4, True, True) if viewdirs else None)
    scale = want.abs().max().item()
    e_tc = (raw1.cpu().double() - want).abs().max().item()
    e_simt = (raw0.cpu().double() - want).abs().max().item()
    assert e_tc <= 3e-5 * scale + 1e-6, (e_tc, e_simt, scale)
    assert torch.isfinite(raw1).all()


def test_hidden_256_renders_on_tcgen05_in_inference_and_trains_on_fp32_kernels():
    """The automatic implementation choice for the shipped-checkpoint architecture (8 x 256): torch.no_grad() renders go
    through the tcgen05 forward and agree with the fp32 CUDA-core path; a training step silently uses the fp32 kernels."""
    import nerf_pytorch_b200 as nb
    from nerf_pytorch_b200 import ops, train_utils
    from oracle import nerf_oracle as O

    torch.manual_seed(0)
    kw = dict(num_layers=8, hidden_size=256, skip_connect_every=4, num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    mc, mf = nb.FlexibleNeRFModel(**kw).cuda(), nb.FlexibleNeRFModel(**kw).cuda()
    epf, edf = nb.get_embedding_function(10), nb.get_embedding_function(4)
    arch = train_utils._arch_of(mc, (10, True, True), (4, True, True))
    assert train_utils._auto_impl(arch, arch, 64, 64, training=False) == ops.IMPL_TC
    assert train_utils._auto_impl(arch, arch, 64, 64, training=True) == ops.IMPL_SIMT
    n = 300
    d = torch.randn(n, 3, device="cuda")
    ro, rd = torch.randn(n, 3, device="cuda") * 0.2 + torch.tensor([0.0, 0.0, 4.0], device="cuda"), -d.abs() * 0.3 - torch.tensor([0, 0, 1.0], device="cuda")
    det = O.make_options(num_coarse=64, num_fine=64, perturb=False, radiance_field_noise_std=0.0)
    with torch.no_grad():
        out_tc = nb.run_one_iter_of_nerf(100, 100, 120.0, mc, mf, ro, rd, det, mode="validation", encode_position_fn=epf,
                                         encode_direction_fn=edf)
        out_32 = nb.run_one_iter_of_nerf(100, 100, 120.0, mc, mf, ro, rd, det, mode="validation", encode_position_fn=epf,
                                         encode_direction_fn=edf, impl=ops.IMPL_SIMT)
    for i, (a, b) in enumerate(zip(out_tc, out_32)):   # (rgb, disp, acc) of the coarse and of the fine pass
        if i % 3 != 1:   # rgb and acc: the same render to fp32 round-off of the MLP outputs
            assert torch.isfinite(a).all()
            # coarse pass: fp32 round-off; fine pass: the resampler sits on the coarse weights, a 1e-7 difference there can move
            # a sample across a bin edge (the same sensitivity the reference has between its own fp32 and fp64 runs)
            tol = 2e-5 if i < 3 else 5e-4
            assert (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item()), (i, (a - b).abs().max().item())
            assert (a - b).abs().mean().item() <= 2e-6, (i, (a - b).abs().mean().item())
        else:            # disp = 1 / (depth / acc) is NaN on an empty ray (reference quirk) and ill-conditioned on a nearly
            acc = out_32[i + 1]   # empty one: compared where the ray has hit something
            hit = acc > 1e-2
            assert (torch.isnan(a) == torch.isnan(b)).all()
            assert (a[hit] - b[hit]).abs().max().item() <= 1e-3 * b[hit].abs().max().item(), (i, (a[hit] - b[hit]).abs().max().item())
    # training: no tcgen05 backward for hidden 256 -> the fp32 kernels, chosen by the library, gradients arrive
    tr = O.make_options(num_coarse=32, num_fine=32, perturb=True, radiance_field_noise_std=0.2)
    out = nb.run_one_iter_of_nerf(100, 100, 120.0, mc, mf, ro[:64], rd[:64], tr, mode="train", encode_position_fn=epf,
                                  encode_direction_fn=edf)
    (out[0].sum() + out[3].sum()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in list(mc.parameters()) + list(mf.parameters()))


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
        assert (r1 - r0).abs().max().item() <= 3e-5 * scale + 1e-6, (n, s, (r1 - r0).abs().max().item(), scale)


def _check_grads(arch, got_flat, want_sd, tol, ctx):
    for lname, w_off, b_off, fin, fout in arch.flat_layout():
        gw = got_flat[w_off:w_off + fin * fout].view(fout, fin).double().cpu()
        gb = got_flat[b_off:b_off + fout].double().cpu()
        for got, want, what in ((gw, want_sd[lname + ".weight"], "weight"), (gb, want_sd[lname + ".bias"], "bias")):
            want = want.double().cpu()
            scale = want.abs().max().item() + 1e-30
            e = (got - want).abs().max().item()
            assert e <= tol * scale, ctx + (lname, what, e, scale)


@pytest.mark.parametrize("name", TC_TRAIN_CASES)
def test_tc_backward_matches_fp64_oracle(name):
    """The fused tcgen05 backward (data-gradient chain + every weight gradient in one kernel) against fp64 autograd of
    the oracle MLP, directly: every weight / bias gradient tensor within 1e-4 of its own scale (worst element)."""
    from nerf_pytorch_b200 import ops

    c = Case(name)
    rays, _, aux = c.aux()
    arch = _arch(c)
    for tag, sd in (("coarse", c.sd_c), ("fine", c.sd_f)):
        if sd is None:
            continue
        z = aux["z_" + tag]
        gen = torch.Generator().manual_seed(11)
        G = torch.randn(z.shape[0], z.shape[1], 4, generator=gen)
        # realistic upstream gradients span decades across the samples of a ray (compositing weights): scale the rows
        G = G * torch.exp(torch.randn(z.shape[0], z.shape[1], 1, generator=gen) * 3.0)
        want = oracle_mlp_grads(c, sd, rays, z, G)
        blob = ops.pack_weights(arch, ops.flatten_state_dict(arch, sd, "cuda"))
        zc, Gc = z.cuda().contiguous(), G.cuda().contiguous()
        _, stash = ops.mlp_fwd(arch, blob, rays.cuda(), zc, impl=ops.IMPL_TC, want_stash=True)
        # ReLU ties: a unit whose pre-activation is within rounding of zero gets relu'() = 0 or 1 depending on the
        # arithmetic (fp64 oracle / fp32 reference / split-precision forward) -- a genuine discontinuity of the
        # function, not a property of the backward.  Count the forward's disagreements with the fp64 oracle, check
        # each is such a tie, then give the backward the oracle's masks so the comparison below measures the
        # backward's arithmetic alone.
        layers, cum_total = gemm_layers(arch)
        acts = oracle_mlp_acts(c, sd, rays, z)
        P = z.numel()
        n_ties = 0
        for li, (lname, n, _) in enumerate(layers):
            if li == 0:
                continue  # layer1 has no ReLU (models.py:238)
            mv = tc_mask_view(arch, stash, P, li, layers, cum_total)
            wantm = pack_mask((acts[li] > 0).cuda())
            diff = (mv ^ wantm).ne(0)
            if diff.any():
                a = acts[li].cuda()
                bits = ((mv ^ wantm).unsqueeze(-1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1
                flipped = bits.view(P, n).bool()
                n_ties += int(flipped.sum())
                # every disagreement sits within 2e-5 of the layer's scale of zero
                assert a[flipped].abs().max().item() <= 2e-5 * a.abs().max().item(), (name, tag, lname)
            mv.copy_(wantm)
        assert n_ties <= 1e-5 * P * cum_total + 2, (name, tag, n_ties)
        g1, _ = ops.mlp_bwd(arch, blob, rays.cuda(), zc, Gc, stash, impl=ops.IMPL_TC)
        torch.cuda.synchronize()
        _check_grads(arch, g1, want, 2e-5, (name, tag, f"{n_ties} relu ties patched"))


def _as_sd(arch, flat):
    out = {}
    for lname, w_off, b_off, fin, fout in arch.flat_layout():
        out[lname + ".weight"] = flat[w_off:w_off + fin * fout].view(fout, fin)
        out[lname + ".bias"] = flat[b_off:b_off + fout]
    return out


def test_tc_backward_ragged_sizes():
    """Point counts that are not multiples of the 128-point tile, tiles that straddle many rays, a single ray: the fused
    tcgen05 backward vs the fp32 CUDA-core kernels."""
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
            _, st0 = ops.mlp_fwd(arch, blob, rays, z, impl=ops.IMPL_SIMT, want_stash=True)
            _, st1 = ops.mlp_fwd(arch, blob, rays, z, impl=ops.IMPL_TC, want_stash=True)
            # same ReLU decisions for both (see test_tc_backward_matches_fp64_oracle): copy the fp32 kernel's masks
            layers, cum_total = gemm_layers(arch)
            P = n * s
            for li in range(1, len(layers)):
                n_l = layers[li][1]
                mask_cum = sum(m // 32 for _, m, _ in layers[:li])
                mask_base = cum_total + ((arch.dim_xyz + 15) & ~15) + ((arch.dim_dir + 7) & ~7)
                simt_words = st0.view(torch.int32)[P * (mask_base + mask_cum):P * (mask_base + mask_cum) + P * (n_l // 32)]
                tc_mask_view(arch, st1, P, li, layers, cum_total).copy_(simt_words.view(P, n_l // 32))
            g0, _ = ops.mlp_bwd(arch, blob, rays, z, G, st0, impl=ops.IMPL_SIMT)
            g1, _ = ops.mlp_bwd(arch, blob, rays, z, G, st1, impl=ops.IMPL_TC)
            _check_grads(arch, g1, _as_sd(arch, g0), 3e-5, (case, n, s))


def test_tc_backward_accumulates_and_repeats():
    """flat_grad is accumulated into (+=), and two calls give bit-identical... no: equal-to-rounding results (the
    reduction order of the bulk reduce-adds is not fixed)."""
    from nerf_pytorch_b200 import ops

    c = Case("a1_skip_lindisp")
    rays, _, aux = c.aux()
    arch = _arch(c)
    blob = ops.pack_weights(arch, ops.flatten_state_dict(arch, c.sd_c, "cuda"))
    z = aux["z_coarse"].cuda().contiguous()
    G = torch.randn(z.shape[0], z.shape[1], 4, generator=torch.Generator().manual_seed(3)).cuda()
    _, stash = ops.mlp_fwd(arch, blob, rays.cuda(), z, impl=ops.IMPL_TC, want_stash=True)
    g1, _ = ops.mlp_bwd(arch, blob, rays.cuda(), z, G, stash, impl=ops.IMPL_TC)
    g2, _ = ops.mlp_bwd(arch, blob, rays.cuda(), z, G, stash, impl=ops.IMPL_TC)
    scale = g1.abs().max().item()
    assert (g1 - g2).abs().max().item() <= 2e-6 * scale


def test_tc_unsupported_configs_are_refused():
    from nerf_pytorch_b200 import ops

    arch = ops.ArchSpec(num_layers=8, hidden=256, skip_every=4, n_freq_xyz=10)
    blob = ops.pack_weights(arch, torch.zeros(arch.flat_param_count(), device="cuda"))
    rays = torch.zeros(4, 11, device="cuda")
    z = torch.ones(4, 64, device="cuda")
    with pytest.raises(NotImplementedError):   # hidden 256: no activation stash / fused backward on tcgen05 ...
        ops.mlp_fwd(arch, blob, rays, z, impl=ops.IMPL_TC, want_stash=True)
    ops.mlp_fwd(arch, blob, rays, z, impl=ops.IMPL_TC)   # ... but the inference forward runs
    assert not ops.impl_supported(arch, 64, ops.IMPL_TC)
    assert ops.impl_supported(arch, 64, ops.IMPL_TC_FWD)
    assert ops.impl_supported(arch, 64, ops.IMPL_SIMT)
    # the fused backward rides the heads' weight gradients on layers_dir[0]: no view directions -> CUDA cores
    assert not ops.impl_supported(ops.ArchSpec(use_viewdirs=False), 64, ops.IMPL_TC)
    assert ops.impl_supported(ops.ArchSpec(num_layers=8, hidden=128, skip_every=3, n_freq_xyz=10), 64, ops.IMPL_TC)
    # the stage-level halves do not exist on tcgen05
    a0 = ops.ArchSpec()
    with pytest.raises(NotImplementedError):
        ops.mlp_dgrad(a0, blob, torch.zeros(4, 64, 4, device="cuda"), torch.zeros(8, device="cuda"), impl=ops.IMPL_TC)


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
