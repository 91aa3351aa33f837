"""Multi-process (gloo, world_size 2, CPU) test of the data-parallel gradient exchange."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as td
    from glue_b200 import dist as gdist
    gdist.init_from_env(backend="gloo")
    assert gdist.is_distributed() and gdist.rank() == rank and gdist.world_size() == world
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    extra = torch.nn.Parameter(torch.ones(4))  # no gradient on rank 1
    bucket = gdist.GradBucket(list(lin.parameters()) + [extra])
    x = torch.full((2, 5), float(rank + 1))
    lin(x).sum().backward()
    if rank == 0:
        extra.grad = torch.full((4,), 2.0)
    bucket.all_reduce_mean()
    # a second step must reuse the views without re-copying stale data
    w1 = lin.weight.grad.clone()
    lin.zero_grad(set_to_none=True)
    extra.grad = None
    lin(x).sum().backward()
    bucket.all_reduce_mean()
    out[rank] = (w1, lin.weight.grad.clone(), lin.bias.grad.clone(), extra.grad.clone())
    td.barrier()
    td.destroy_process_group()


def test_grad_bucket_all_reduce_mean_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    w0, w0b, b0, e0 = out[0]
    w1, w1b, b1, e1 = out[1]
    # d/dW of sum(lin(x)) = sum over batch of x -> rank r contributes 2*(r+1) per entry
    assert torch.allclose(w0, torch.full((3, 5), 3.0)) and torch.equal(w0, w1)
    assert torch.allclose(w0b, torch.full((3, 5), 3.0)) and torch.equal(w0b, w1b)
    assert torch.allclose(b0, torch.full((3,), 2.0)) and torch.equal(b0, b1)
    assert torch.allclose(e0, torch.zeros(4)) and torch.equal(e0, e1)  # second step: no grads
