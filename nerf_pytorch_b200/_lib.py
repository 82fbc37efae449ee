"""ctypes binding of libnerfb200.so (C ABI declared in include/nerfb200.h).

The product path has NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised (never a silent PyTorch / CPU path)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnerfb200.so")

MAX_FREQS = 16

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA = 0, -1, -2, -3


class Arch(C.Structure):
    """nerfb200_arch_t"""

    _fields_ = [
        ("num_layers", C.c_int32),
        ("hidden", C.c_int32),
        ("skip_every", C.c_int32),
        ("use_viewdirs", C.c_int32),
        ("n_freq_xyz", C.c_int32),
        ("n_freq_dir", C.c_int32),
        ("include_input_xyz", C.c_int32),
        ("include_input_dir", C.c_int32),
        ("freq_xyz", C.c_float * MAX_FREQS),
        ("freq_dir", C.c_float * MAX_FREQS),
    ]


class RenderOpts(C.Structure):
    """nerfb200_render_opts_t"""

    _fields_ = [
        ("n_coarse", C.c_int32),
        ("n_fine", C.c_int32),
        ("perturb", C.c_int32),
        ("lindisp", C.c_int32),
        ("white_bkgd", C.c_int32),
        ("noise_std", C.c_float),
    ]


P = C.c_void_p
I32, I64, F32 = C.c_int32, C.c_int64, C.c_float
AP, OP = C.POINTER(Arch), C.POINTER(RenderOpts)

# name -> (restype, argtypes); mirrors include/nerfb200.h one-to-one
SIGNATURES = {
    "nerfb200_version": (I32, []),
    "nerfb200_last_error": (C.c_char_p, []),
    "nerfb200_launch_count": (I64, []),
    "nerfb200_bwd_bytes_per_point": (I64, [AP]),
    "nerfb200_impl_supported": (I32, [AP, I32, I32]),
    "nerfb200_num_linear": (I64, [AP]),
    "nerfb200_flat_param_count": (I64, [AP]),
    "nerfb200_blob_floats": (I64, [AP]),
    "nerfb200_flat_layout": (I32, [AP, I32, C.POINTER(I64), C.POINTER(I64), C.POINTER(I32), C.POINTER(I32)]),
    "nerfb200_pack_weights": (I32, [AP, P, P, P]),
    "nerfb200_sample_coarse": (I32, [P, I32, I64, P, P, I32, I32, I32, P, P]),
    "nerfb200_gen_rays": (I32, [C.POINTER(F32), I32, I32, F32, P, I64, I32, F32, F32, I32, I32, P, P]),
    "nerfb200_pack_rays": (I32, [P, P, I64, I32, I32, F32, I32, F32, F32, I32, I32, P, P]),
    "nerfb200_encode": (I32, [AP, I32, P, I64, P, P]),
    "nerfb200_stash_floats": (I64, [AP, I64]),
    "nerfb200_mlp_fwd": (I32, [AP, P, P, I32, P, I64, I32, P, P, I32, P]),
    "nerfb200_composite_fwd": (I32, [P, P, P, I32, P, I64, I32, F32, I32, P, P, P]),
    "nerfb200_composite_bwd": (I32, [P, P, P, I32, P, P, I64, I32, F32, I32, P, P]),
    "nerfb200_sample_pdf_merge": (I32, [P, P, P, I32, P, I64, I32, I32, P, P, P, P, P]),
    "nerfb200_bwd_scratch_floats": (I64, [AP, I64, I32]),
    "nerfb200_mlp_bwd": (I32, [AP, P, P, I32, P, I64, I32, P, P, P, P, I32, P]),
    "nerfb200_mlp_dgrad": (I32, [AP, P, P, P, P, I64, I32, P]),
    "nerfb200_mlp_wgrad": (I32, [AP, P, I32, P, I64, I32, P, P, P, P, I32, P]),
    "nerfb200_render_workspace_bytes": (I64, [AP, AP, OP, I64, I32]),
    "nerfb200_render_workspace_layout": (I32, [AP, AP, OP, I64, I32, C.POINTER(I64)]),
    "nerfb200_render_fwd": (I32, [AP, AP, OP, P, P, P, I32, I64, P, P, P, P, I32, P, P, P, P, I32, I32, P]),
    "nerfb200_render_bwd": (I32, [AP, AP, OP, P, P, P, I32, I64, P, P, P, P, P, P, P, I32, I32, P]),
    "nerfb200_adam_step": (I32, [P, P, P, P, I64, I32, F32, F32, F32, F32, F32, P]),
}

_lib = None


def build_library(verbose: bool = False) -> str:
    """Compile csrc/*.cu into libnerfb200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libnerfb200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C nerf_pytorch_b200/csrc`). There is no fallback path."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != OK:
        msg = load().nerfb200_last_error().decode(errors="replace")
        kind = {ERR_INVALID: "invalid argument", ERR_UNSUPPORTED: "unsupported configuration", ERR_CUDA: "CUDA error"}.get(rc, "error")
        if rc == ERR_UNSUPPORTED:
            raise NotImplementedError(f"nerfb200 {what}: {kind}: {msg}")
        raise RuntimeError(f"nerfb200 {what}: {kind}: {msg}")
