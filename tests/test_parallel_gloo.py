"""CPU, world_size 2, gloo: the data-parallel host logic of SURVEY.md section 8(e).

Each rank renders its contiguous shard of the global batch and the flat gradient is sum-all-reduced and
averaged once; with equal shards that equals the gradient of the global-batch loss.  The per-shard
gradient here comes from the oracle (CPU stand-in for the CUDA backward, which the gloo test cannot run)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from helpers import Case
    from nerf_pytorch_b200 import parallel
    from oracle import nerf_oracle as O

    r, l, w = parallel.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    parallel.enable_gradient_sync()
    from nerf_pytorch_b200 import train_utils
    assert train_utils._GRAD_SYNC is not None and train_utils._GRAD_SYNC[1] == world

    c = Case("a0_noview_coarse_only")   # coarse only: deterministic given injected randoms, cheap on CPU
    n = c.ro.shape[0]
    lo, hi = parallel.shard_bounds(n, rank, world)
    sl = slice(lo, hi)

    def grads(rows):
        sd = {k: v.clone().requires_grad_(True) for k, v in c.sd_c.items()}
        rnd = {k: v[rows] for k, v in c.randoms.items()}
        out = O.run_one_iter_of_nerf(c.H, c.W, c.focal, sd, None, c.ro[rows], c.rd[rows], c.options,
                                     enc_xyz=c.enc_xyz, enc_dir=c.enc_dir, randoms=rnd)
        O.nerf_loss(out, c.target[rows]).backward()
        return torch.cat([sd[k].grad.reshape(-1) for k in sorted(sd)])

    flat = grads(sl)
    parallel.allreduce_flat_(flat, world)            # the one collective of a step
    full = grads(slice(0, n))
    ok = torch.allclose(flat, full, rtol=1e-4, atol=1e-7)
    ro_s, rd_s = parallel.shard_rays(rank, world, c.ro, c.rd)
    ok = ok and ro_s.shape[0] == hi - lo and torch.equal(rd_s, c.rd[sl])
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_global_batch():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _worker_sync(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from nerf_pytorch_b200 import parallel, train_utils

    parallel.init_distributed(backend="gloo")
    parallel.enable_gradient_sync()
    # a multi-chunk backward leaves the summed per-chunk gradients unsynchronised and raises the flag; ONE all-reduce
    # over everything follows (never one per chunk: ranks may differ in their chunk counts)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5))]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1 + 10 * i))
    ok = parallel.sync_gradients(params) is False          # nothing pending: no collective, gradients untouched
    ok = ok and float(params[0].grad[0, 0]) == rank + 1
    train_utils._PENDING_SYNC = True
    ok = ok and parallel.sync_gradients(params) is True
    want0, want1 = (1 + 2) / 2.0, (11 + 12) / 2.0
    ok = ok and torch.allclose(params[0].grad, torch.full((3, 4), want0)) and torch.allclose(params[1].grad, torch.full((5,), want1))
    ok = ok and train_utils._PENDING_SYNC is False
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_multi_chunk_gradients_are_all_reduced_once():
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_sync, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
