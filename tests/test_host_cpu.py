"""CPU (no GPU needed): the C-ABI library loads and exports every symbol include/nerfb200.h declares,
its host-side planning functions answer correctly, and the Python host logic (encoder probing,
model introspection, sharding) behaves.  No compute kernels are launched here."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from nerf_pytorch_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "nerfb200.h")).read()
    declared = sorted(set(re.findall(r"\b(nerfb200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == declared
    assert _lib.load().nerfb200_version() == 100


def test_flat_layout_matches_reference_state_dict_shapes():
    from nerf_pytorch_b200 import ops
    from oracle import nerf_oracle as O

    for kw, n_params in ((dict(num_layers=4, hidden=128, skip_every=4), 84548),
                         (dict(num_layers=8, hidden=128, skip_every=3), 166724),
                         (dict(num_layers=8, hidden=256, skip_every=4), 595844)):
        arch = ops.ArchSpec(n_freq_xyz=10, n_freq_dir=4, **kw)
        assert arch.flat_param_count() == n_params
        sd = O.init_flexible_nerf(kw["num_layers"], kw["hidden"], kw["skip_every"], 10, 4)
        off = 0
        for name, w_off, b_off, fin, fout in arch.flat_layout():
            assert tuple(sd[name + ".weight"].shape) == (fout, fin), name
            assert w_off == off and b_off == off + fin * fout
            off = b_off + fout
        assert off == n_params
        assert set(n + s for n in arch.slot_names() for s in (".weight", ".bias")) == set(sd)


def test_unsupported_architectures_are_rejected_not_approximated():
    from nerf_pytorch_b200 import ops

    with pytest.raises(NotImplementedError):
        ops.ArchSpec(hidden=96).flat_param_count()
    with pytest.raises(NotImplementedError):
        ops.ArchSpec(num_layers=40).blob_floats()
    with pytest.raises(NotImplementedError):
        ops.ArchSpec(n_freq_xyz=17).c_struct()


def test_frequency_bands_follow_reference():
    from nerf_pytorch_b200 import ops

    assert ops.frequency_bands(10, True).tolist() == [2.0 ** i for i in range(10)]
    assert torch.equal(ops.frequency_bands(5, False), torch.linspace(1.0, 16.0, 5))


def test_encoder_probe_identifies_reference_lambdas():
    from nerf_pytorch_b200 import train_utils
    from nerf_pytorch_b200.nerf_helpers import get_embedding_function
    from oracle import nerf_oracle as O

    for L, inc, log in ((10, True, True), (4, True, True), (6, False, True), (5, True, False), (0, True, True)):
        fn = lambda x, L=L, inc=inc, log=log: O.positional_encoding(x, L, inc, log)  # noqa: E731
        got = train_utils._probe_encoder(fn, None)
        if L == 0:
            assert got[0] == 0 and got[1] is True
        else:
            assert got == (L, inc, log), (got, (L, inc, log))
    assert train_utils._probe_encoder(get_embedding_function(7, False, True), None) == (7, False, True)
    with pytest.raises(NotImplementedError):
        train_utils._probe_encoder(lambda x: torch.tanh(x), None)


def test_model_introspection_and_state_dict_compat():
    import nerf_pytorch_b200 as nb
    from nerf_pytorch_b200 import train_utils
    from oracle import nerf_oracle as O

    m = nb.FlexibleNeRFModel(num_layers=8, hidden_size=128, skip_connect_every=3, num_encoding_fn_xyz=10)
    sd = O.init_flexible_nerf(8, 128, 3, 10, 4)
    m.load_state_dict(sd)  # same keys/shapes as the reference class
    arch = train_utils._arch_of(m, (10, True, True), (4, True, True))
    assert (arch.num_layers, arch.hidden, arch.skip_every) == (8, 128, 3)
    x = torch.randn(5, 63 + 27)
    assert torch.allclose(m(x), O.flexible_nerf_forward(sd, x, 63), atol=1e-6)
    with pytest.raises(NotImplementedError):
        train_utils._arch_of(torch.nn.Linear(3, 3), (10, True, True), (4, True, True))
    with pytest.raises(RuntimeError):
        train_utils._arch_of(m, (6, True, True), (4, True, True))  # encoder / layer1 width mismatch


def test_product_path_refuses_cpu_tensors():
    import nerf_pytorch_b200 as nb
    from oracle import nerf_oracle as O

    m = nb.FlexibleNeRFModel(num_encoding_fn_xyz=10)
    rays = torch.zeros(4, 11)
    with pytest.raises(NotImplementedError):
        nb.predict_and_render_radiance(rays, m, m, O.make_options(), encode_position_fn=nb.get_embedding_function(10),
                                       encode_direction_fn=nb.get_embedding_function(4))


def test_shard_bounds_partition():
    from nerf_pytorch_b200.parallel import shard_bounds

    for n, w in ((32768, 8), (4096, 1), (10, 3), (7, 8)):
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
