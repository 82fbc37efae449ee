"""Torch-tensor wrappers over the C ABI (include/nerfb200.h): device memory, streams, nothing else.

Every function checks that its tensors are CUDA / fp32 / contiguous, allocates the outputs with
``torch.empty`` and launches on ``torch.cuda.current_stream()``.  No math happens in Python."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from functools import lru_cache
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import Arch, RenderOpts

IMPL_SIMT = 0  # fp32 CUDA cores
IMPL_TC = 1    # tcgen05 tensor cores (fp16x2 three-term splits, fp32 accumulation)
IMPL_TC_FWD = 2  # query only (impl_supported): the tcgen05 forward alone (inference), which also runs hidden 256


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("nerfb200: tensor is not on a CUDA device (there is no CPU path)")
    if t.dtype not in (torch.float32, torch.int32, torch.int64, torch.uint8):
        raise RuntimeError(f"nerfb200: unsupported dtype {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError("nerfb200: tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    """Kernels launched by libnerfb200.so in this process so far."""
    return int(_lib.load().nerfb200_launch_count())


def frequency_bands(n: int, log_sampling: bool) -> torch.Tensor:
    """The bands exactly as the reference builds them (nerf/nerf_helpers.py:131-147), fp32, on CPU."""
    if n == 0:
        return torch.zeros(0)
    if log_sampling:
        return 2.0 ** torch.linspace(0.0, n - 1, n, dtype=torch.float32)
    return torch.linspace(2.0 ** 0.0, 2.0 ** (n - 1), n, dtype=torch.float32)


@dataclass(frozen=True)
class ArchSpec:
    """Hashable description of one FlexibleNeRFModel + its encoders (-> nerfb200_arch_t)."""

    num_layers: int = 4
    hidden: int = 128
    skip_every: int = 4
    use_viewdirs: bool = True
    n_freq_xyz: int = 6
    n_freq_dir: int = 4
    include_input_xyz: bool = True
    include_input_dir: bool = True
    log_sampling_xyz: bool = True
    log_sampling_dir: bool = True

    @property
    def dim_xyz(self):
        return (3 if self.include_input_xyz else 0) + 6 * self.n_freq_xyz

    @property
    def dim_dir(self):
        return ((3 if self.include_input_dir else 0) + 6 * self.n_freq_dir) if self.use_viewdirs else 0

    @lru_cache(maxsize=None)
    def c_struct(self) -> Arch:
        a = Arch()
        a.num_layers, a.hidden, a.skip_every = self.num_layers, self.hidden, self.skip_every
        a.use_viewdirs = int(self.use_viewdirs)
        a.n_freq_xyz, a.n_freq_dir = self.n_freq_xyz, self.n_freq_dir
        a.include_input_xyz, a.include_input_dir = int(self.include_input_xyz), int(self.include_input_dir)
        if self.n_freq_xyz > _lib.MAX_FREQS or self.n_freq_dir > _lib.MAX_FREQS:
            raise NotImplementedError("nerfb200: more than 16 encoding frequencies")
        for i, f in enumerate(frequency_bands(self.n_freq_xyz, self.log_sampling_xyz).tolist()):
            a.freq_xyz[i] = f
        for i, f in enumerate(frequency_bands(self.n_freq_dir, self.log_sampling_dir).tolist()):
            a.freq_dir[i] = f
        return a

    # ---- flat parameter vector layout -------------------------------------------------------
    def slot_names(self) -> Sequence[str]:
        """Reference state_dict prefixes in the canonical slot order of include/nerfb200.h."""
        names = ["layer1"] + [f"layers_xyz.{i}" for i in range(self.num_layers - 1)]
        if self.use_viewdirs:
            names += ["fc_feat", "fc_alpha", "layers_dir.0", "fc_rgb"]
        else:
            names += ["fc_out"]
        return names

    @lru_cache(maxsize=None)
    def flat_layout(self):
        """[(name, w_off, b_off, in, out)] from the library (single source of truth)."""
        lib = _lib.load()
        a = self.c_struct()
        out = []
        for slot, name in enumerate(self.slot_names()):
            w, b, i, o = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
            _lib.check(lib.nerfb200_flat_layout(C.byref(a), slot, C.byref(w), C.byref(b), C.byref(i), C.byref(o)),
                       "flat_layout")
            out.append((name, w.value, b.value, i.value, o.value))
        return out

    @lru_cache(maxsize=None)
    def flat_param_count(self) -> int:
        n = _lib.load().nerfb200_flat_param_count(C.byref(self.c_struct()))
        if n < 0:
            _lib.check(_lib.ERR_UNSUPPORTED, "flat_param_count")
        return n

    @lru_cache(maxsize=None)
    def blob_floats(self) -> int:
        n = _lib.load().nerfb200_blob_floats(C.byref(self.c_struct()))
        if n < 0:
            _lib.check(_lib.ERR_UNSUPPORTED, "blob_floats")
        return n


def make_opts(n_coarse, n_fine, perturb, lindisp, white_bkgd, noise_std) -> RenderOpts:
    o = RenderOpts()
    o.n_coarse, o.n_fine = int(n_coarse), int(n_fine)
    o.perturb, o.lindisp, o.white_bkgd = int(bool(perturb)), int(bool(lindisp)), int(bool(white_bkgd))
    o.noise_std = float(noise_std)
    return o


# ------------------------------------------------------------------------------------------------
# parameters
# ------------------------------------------------------------------------------------------------
def flatten_state_dict(arch: ArchSpec, sd, device) -> torch.Tensor:
    """state_dict (reference key names) -> flat fp32 vector in canonical order."""
    parts = []
    for name, _, _, fin, fout in arch.flat_layout():
        w, b = sd[name + ".weight"], sd[name + ".bias"]
        if tuple(w.shape) != (fout, fin) or tuple(b.shape) != (fout,):
            raise RuntimeError(f"nerfb200: {name} has shape {tuple(w.shape)}, expected {(fout, fin)}")
        parts += [w.reshape(-1), b.reshape(-1)]
    return torch.cat([p.to(device=device, dtype=torch.float32) for p in parts]).contiguous()


def pack_weights(arch: ArchSpec, flat: torch.Tensor, blob: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    if flat.numel() != arch.flat_param_count():
        raise RuntimeError("nerfb200: flat parameter vector has the wrong length")
    if blob is None:
        blob = torch.empty(arch.blob_floats(), dtype=torch.float32, device=flat.device)
    _lib.check(lib.nerfb200_pack_weights(C.byref(arch.c_struct()), _ptr(flat), _ptr(blob), _stream()), "pack_weights")
    return blob


# ------------------------------------------------------------------------------------------------
# stage-level ops (test hooks + building blocks)
# ------------------------------------------------------------------------------------------------
def sample_coarse(rays, t_vals, t_rand, n_coarse, perturb, lindisp):
    lib = _lib.load()
    n = rays.shape[0]
    z = torch.empty(n, n_coarse, dtype=torch.float32, device=rays.device)
    _lib.check(lib.nerfb200_sample_coarse(_ptr(rays), rays.shape[1], n, _ptr(t_vals), _ptr(t_rand) if perturb else None,
                                          n_coarse, int(bool(perturb)), int(bool(lindisp)), _ptr(z), _stream()),
               "sample_coarse")
    return z


def gen_rays(c2w, height, width, focal, pixel_ids, device, ndc=False, near=0.0, far=1.0, use_viewdirs=True, stride=None):
    """Rays of the given pixels (int64 ids j * W + i on `device`, or None for the whole image) of a pinhole camera with
    camera-to-world matrix ``c2w`` (anything convertible to 12+ host floats: the top 3 x 4 block is used)."""
    lib = _lib.load()
    m = torch.as_tensor(c2w, dtype=torch.float32, device="cpu").reshape(-1, 4)[:3].contiguous().reshape(-1)
    c12 = (C.c_float * 12)(*m.tolist())
    n = int(pixel_ids.numel()) if pixel_ids is not None else int(height) * int(width)
    if stride is None:
        stride = 11 if use_viewdirs else 8
    out = torch.empty(n, stride, dtype=torch.float32, device=device)
    _lib.check(lib.nerfb200_gen_rays(c12, int(height), int(width), float(focal), _ptr(pixel_ids), n, int(bool(ndc)),
                                     float(near), float(far), int(bool(use_viewdirs) and stride == 11), stride, _ptr(out),
                                     _stream()), "gen_rays")
    return out


def pack_rays(ro, rd, height, width, focal, ndc, near, far, use_viewdirs, stride=None):
    """[o d near far (viewdir)] rows from caller-supplied origins / directions (train_utils.py:143-168);
    stride 6 returns [o d] only (ndc_rays)."""
    lib = _lib.load()
    n = ro.shape[0]
    if stride is None:
        stride = 11 if use_viewdirs else 8
    out = torch.empty(n, stride, dtype=torch.float32, device=ro.device)
    _lib.check(lib.nerfb200_pack_rays(_ptr(ro), _ptr(rd), n, int(height), int(width), float(focal), int(bool(ndc)),
                                      float(near), float(far), int(bool(use_viewdirs)), stride, _ptr(out), _stream()),
               "pack_rays")
    return out


def encode(arch: ArchSpec, which: int, x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    n = x.shape[0]
    dim = arch.dim_dir if which else arch.dim_xyz
    out = torch.empty(n, dim, dtype=torch.float32, device=x.device)
    _lib.check(lib.nerfb200_encode(C.byref(arch.c_struct()), which, _ptr(x), n, _ptr(out), _stream()), "encode")
    return out


def stash_floats(arch: ArchSpec, n_points: int) -> int:
    return _lib.load().nerfb200_stash_floats(C.byref(arch.c_struct()), n_points)


def mlp_fwd(arch: ArchSpec, blob, rays, z, impl=IMPL_SIMT, want_stash=False):
    lib = _lib.load()
    n, s = z.shape
    raw = torch.empty(n, s, 4, dtype=torch.float32, device=z.device)
    stash = torch.empty(stash_floats(arch, n * s), dtype=torch.float32, device=z.device) if want_stash else None
    _lib.check(lib.nerfb200_mlp_fwd(C.byref(arch.c_struct()), _ptr(blob), _ptr(rays), rays.shape[1], _ptr(z), n, s,
                                    _ptr(raw), _ptr(stash), impl, _stream()), "mlp_fwd")
    return (raw, stash) if want_stash else raw


def impl_supported(arch: ArchSpec, n_samples: int, impl: int) -> bool:
    """Can `impl` run this architecture forward and backward (the library decides, not Python)?"""
    return _lib.load().nerfb200_impl_supported(C.byref(arch.c_struct()), int(n_samples), int(impl)) == _lib.OK


def mlp_bwd(arch: ArchSpec, blob, rays, z, d_raw, stash, impl=IMPL_SIMT):
    """Returns (flat_grad, scratch): scratch is the gradient stash for impl 0, the gradient blob for impl 1."""
    lib = _lib.load()
    n, s = z.shape
    gstash = torch.empty(lib.nerfb200_bwd_scratch_floats(C.byref(arch.c_struct()), n * s, impl), dtype=torch.float32,
                         device=z.device)
    flat_grad = torch.zeros(arch.flat_param_count(), dtype=torch.float32, device=z.device)
    _lib.check(lib.nerfb200_mlp_bwd(C.byref(arch.c_struct()), _ptr(blob), _ptr(rays), rays.shape[1], _ptr(z), n, s,
                                    _ptr(d_raw), _ptr(stash), _ptr(gstash), _ptr(flat_grad), impl, _stream()),
               "mlp_bwd")
    return flat_grad, gstash


def mlp_dgrad(arch: ArchSpec, blob, d_raw, stash, impl=IMPL_SIMT, gstash=None):
    """First half of mlp_bwd: the pre-activation gradients of every layer (the gradient stash)."""
    lib = _lib.load()
    if gstash is None:
        gstash = torch.empty_like(stash)
    _lib.check(lib.nerfb200_mlp_dgrad(C.byref(arch.c_struct()), _ptr(blob), _ptr(d_raw), _ptr(stash), _ptr(gstash),
                                      d_raw.numel() // 4, impl, _stream()), "mlp_dgrad")
    return gstash


def mlp_wgrad(arch: ArchSpec, rays, z, d_raw, stash, gstash, impl=IMPL_SIMT, flat_grad=None):
    """Second half of mlp_bwd: dW = dY^T X over all points, accumulated into ``flat_grad``."""
    lib = _lib.load()
    n, s = z.shape
    if flat_grad is None:
        flat_grad = torch.zeros(arch.flat_param_count(), dtype=torch.float32, device=z.device)
    _lib.check(lib.nerfb200_mlp_wgrad(C.byref(arch.c_struct()), _ptr(rays), rays.shape[1], _ptr(z), n, s, _ptr(d_raw),
                                      _ptr(stash), _ptr(gstash), _ptr(flat_grad), impl, _stream()), "mlp_wgrad")
    return flat_grad


def bwd_bytes_per_point(arch: ArchSpec) -> int:
    """HBM bytes the tcgen05 backward reads per point (activation tiles of its weight-gradient jobs, masks, d_raw)."""
    return int(_lib.load().nerfb200_bwd_bytes_per_point(C.byref(arch.c_struct())))


def composite_fwd(raw, z, rays, noise, noise_std, white_bkgd, want_weights=True):
    lib = _lib.load()
    n, s = z.shape
    out = torch.empty(n, 8, dtype=torch.float32, device=z.device)
    w = torch.empty(n, s, dtype=torch.float32, device=z.device) if want_weights else None
    _lib.check(lib.nerfb200_composite_fwd(_ptr(raw), _ptr(z), _ptr(rays), rays.shape[1], _ptr(noise), n, s,
                                          float(noise_std), int(bool(white_bkgd)), _ptr(out), _ptr(w), _stream()),
               "composite_fwd")
    return out, w


def composite_bwd(raw, z, rays, noise, g_out, noise_std, white_bkgd):
    lib = _lib.load()
    n, s = z.shape
    d_raw = torch.empty(n, s, 4, dtype=torch.float32, device=z.device)
    _lib.check(lib.nerfb200_composite_bwd(_ptr(raw), _ptr(z), _ptr(rays), rays.shape[1], _ptr(noise), _ptr(g_out), n,
                                          s, float(noise_std), int(bool(white_bkgd)), _ptr(d_raw), _stream()),
               "composite_bwd")
    return d_raw


def sample_pdf_merge(z_coarse, weights_coarse, u, n_fine, cdf_in=None, want_aux=False):
    """u: (n_rays, n_fine) or (n_fine,) shared by all rays (the det=True linspace)."""
    lib = _lib.load()
    n, nc = z_coarse.shape
    dev = z_coarse.device
    z_fine = torch.empty(n, nc + n_fine, dtype=torch.float32, device=dev)
    zs = torch.empty(n, n_fine, dtype=torch.float32, device=dev) if want_aux else None
    inds = torch.empty(n, n_fine, dtype=torch.int32, device=dev) if want_aux else None
    cdf = torch.empty(n, nc - 1, dtype=torch.float32, device=dev) if want_aux else None
    u_stride = 0 if u.dim() == 1 else n_fine
    _lib.check(lib.nerfb200_sample_pdf_merge(_ptr(z_coarse), _ptr(weights_coarse), _ptr(u), u_stride, _ptr(cdf_in), n,
                                             nc, n_fine, _ptr(z_fine), _ptr(zs), _ptr(inds), _ptr(cdf), _stream()),
               "sample_pdf_merge")
    return (z_fine, zs, inds, cdf) if want_aux else z_fine


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    lib = _lib.load()
    _lib.check(lib.nerfb200_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), int(step), float(lr), float(beta1),
                                      float(beta2), float(eps), float(grad_scale), _stream()), "adam_step")


# ------------------------------------------------------------------------------------------------
# whole path
# ------------------------------------------------------------------------------------------------
WS_SECTIONS = ("z_coarse", "raw_coarse", "weights_coarse", "z_fine", "raw_fine", "stash_coarse", "stash_fine",
               "gstash", "d_raw")


def render_workspace_bytes(arch_c, arch_f, opts: RenderOpts, n_rays: int, training: bool) -> int:
    lib = _lib.load()
    ac = arch_c.c_struct()
    af = arch_f.c_struct() if arch_f is not None else None
    n = lib.nerfb200_render_workspace_bytes(C.byref(ac), C.byref(af) if af is not None else None, C.byref(opts),
                                            n_rays, int(training))
    if n < 0:
        _lib.check(_lib.ERR_UNSUPPORTED, "render_workspace_bytes")
    return n


def render_workspace_layout(arch_c, arch_f, opts: RenderOpts, n_rays: int, training: bool) -> dict:
    lib = _lib.load()
    ac = arch_c.c_struct()
    af = arch_f.c_struct() if arch_f is not None else None
    offs = (C.c_int64 * 9)()
    _lib.check(lib.nerfb200_render_workspace_layout(C.byref(ac), C.byref(af) if af is not None else None,
                                                    C.byref(opts), n_rays, int(training), offs), "workspace_layout")
    return dict(zip(WS_SECTIONS, list(offs)))


def render_fwd(arch_c, arch_f, opts: RenderOpts, blob_c, blob_f, rays, t_vals, t_rand, noise_c, u, noise_f,
               training: bool, impl: int = IMPL_SIMT, workspace: Optional[torch.Tensor] = None):
    """One ray chunk through the whole path.  Returns (out_coarse[N,8], out_fine[N,8] | None, workspace)."""
    lib = _lib.load()
    n = rays.shape[0]
    dev = rays.device
    fine = opts.n_fine > 0
    need = render_workspace_bytes(arch_c, arch_f if fine else None, opts, n, training)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    out_c = torch.empty(n, 8, dtype=torch.float32, device=dev)
    out_f = torch.empty(n, 8, dtype=torch.float32, device=dev) if fine else None
    ac = arch_c.c_struct()
    af = arch_f.c_struct() if fine else None
    u_stride = 0 if (u is not None and u.dim() == 1) else opts.n_fine
    _lib.check(lib.nerfb200_render_fwd(C.byref(ac), C.byref(af) if af is not None else None, C.byref(opts),
                                       _ptr(blob_c), _ptr(blob_f) if fine else None, _ptr(rays), rays.shape[1], n,
                                       _ptr(t_vals), _ptr(t_rand), _ptr(noise_c), _ptr(u) if fine else None, u_stride,
                                       _ptr(noise_f) if fine else None, _ptr(out_c), _ptr(out_f), _ptr(workspace),
                                       int(training), impl, _stream()), "render_fwd")
    return out_c, out_f, workspace


def render_bwd(arch_c, arch_f, opts: RenderOpts, blob_c, blob_f, rays, noise_c, noise_f, g_c, g_f, workspace,
               flat_grad_c, flat_grad_f, impl: int = IMPL_SIMT, parts: int = 3):
    lib = _lib.load()
    fine = opts.n_fine > 0
    ac = arch_c.c_struct()
    af = arch_f.c_struct() if fine else None
    _lib.check(lib.nerfb200_render_bwd(C.byref(ac), C.byref(af) if af is not None else None, C.byref(opts),
                                       _ptr(blob_c), _ptr(blob_f) if fine else None, _ptr(rays), rays.shape[1],
                                       rays.shape[0], _ptr(noise_c), _ptr(noise_f) if fine else None, _ptr(g_c),
                                       _ptr(g_f) if fine else None, _ptr(workspace), _ptr(flat_grad_c),
                                       _ptr(flat_grad_f) if fine else None, impl, int(parts), _stream()), "render_bwd")
