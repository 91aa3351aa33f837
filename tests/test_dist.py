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


def _cpu_spmm(csr_part, inp, n_rows=None, out=None):
    """CPU stand-in for the CSR GraphConv kernel (tests of the host-side partition logic)."""
    indptr, col, val = csr_part
    n = indptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n), (indptr[1:] - indptr[:-1]).long())
    res = torch.zeros(n, inp.shape[1], dtype=inp.dtype)
    res.index_add_(0, rows, inp[col.long()] * val.unsqueeze(1))
    if out is None:
        return res
    out[:n] = res
    return out


def _partition_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import numpy as np
    import torch.distributed as td
    from glue_b200 import dist as gdist
    from glue_b200 import ops
    from oracle import sampling
    from tests.util import random_graph
    gdist.init_from_env(backend="gloo")
    V, E, K = 101, 700, 6   # V not divisible by the world size: the last block is short
    rs = np.random.RandomState(5)
    eidx, ewt, esgn = random_graph(rs, V, E)
    enorm = sampling.normalize_edges(eidx, ewt)
    csr = ops.GraphCSR.__new__(ops.GraphCSR)  # host-built CSR (the device builder needs a GPU)
    csr.vnum, csr.n_edges, csr.device = V, eidx.shape[1], torch.device("cpu")
    csr.fwd = tuple(torch.as_tensor(a) for a in sampling.csr_forward(eidx, enorm, esgn, V))
    csr.bwd = tuple(torch.as_tensor(a) for a in sampling.csr_transposed(eidx, enorm, esgn, V))
    ops._spmm = lambda part, inp: _cpu_spmm(part, inp)
    ops._spmm_rows = lambda part, inp, n_rows, o: _cpu_spmm(part, inp, n_rows, o)
    part = ops.PartitionedGraphCSR(csr, rank, world)
    torch.manual_seed(0)
    inp = torch.randn(V, K, requires_grad=True)
    res = ops.graph_conv_partitioned(inp, part)
    gout = torch.randn(V, K, generator=torch.Generator().manual_seed(10 + rank))  # rank-specific
    (gin,) = torch.autograd.grad(res, inp, gout)
    want = _cpu_spmm(csr.fwd, inp.detach())
    want_gin = _cpu_spmm(csr.bwd, gout)
    out[rank] = (torch.equal(res.detach(), want), torch.allclose(gin, want_gin, atol=1e-6),
                 part.n_rows, part.row0)
    td.barrier()
    td.destroy_process_group()


def test_partitioned_graphconv_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_partition_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] and out[1][0], "all-gathered row blocks must equal the unpartitioned result"
    assert out[0][1] and out[1][1], "backward must apply the whole transposed graph locally"
    assert (out[0][2], out[0][3]) == (51, 0) and (out[1][2], out[1][3]) == (50, 51)


def _fit_worker(rank, world, port, out, tmpdir):
    """Two ranks drive the real fit plumbing (loaders, engine, LR scheduler, early stopping) with a
    stub step whose validation loss differs per rank: rank 0 alone would see steady improvement,
    rank 1 alone a plateau.  With rank-averaged epoch metrics both take the same decisions."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as td
    from glue_b200 import dist as gdist
    from glue_b200 import models
    from glue_b200.models import scglue as gs
    from glue_b200.models.base import LossVector
    from glue_b200.utils import config
    from tests.synth import make_synthetic
    gdist.init_from_env(backend="gloo")
    config.CPU_ONLY = True
    adatas, graph = make_synthetic(n_cells=(96, 96), n_feat=(12, 30), n_pairs=40, seed=1, rep_dim=6, paired=True)
    for a in adatas.values():
        models.configure_dataset(a, "NB", use_highly_variable=False, use_rep="X_rep", use_obs_names=True)
    steps = []

    class StubTrainer(gs.PairedSCGLUETrainer):
        def _vec(self, engine, val):
            e = engine.state.epoch
            # rank 0: 1/e (always improving); rank 1: constant after epoch 2
            score = 1.0 / e if rank == 0 else (1.0 if e <= 2 else 0.9)
            vals = torch.full((len(self.required_losses),), score if val else 1.0)
            return LossVector(self.required_losses, vals)

        def train_step(self, engine, data):
            steps.append(engine.state.epoch)
            return self._vec(engine, False)

        def val_step(self, engine, data):
            return self._vec(engine, True)

    gs.PairedSCGLUEModel.TRAINER_TYPE = StubTrainer
    model = gs.PairedSCGLUEModel(adatas, sorted(graph.nodes), latent_dim=4, h_dim=8, random_seed=0)
    model.compile()
    model.fit(adatas, graph, data_batch_size=16, max_epochs=12, patience=3, reduce_lr_patience=1,
              align_burnin=1, directory=None)
    lr = model.trainer.vae_optim.param_groups[0]["lr"]
    out[rank] = (max(steps), len(steps), lr)
    td.barrier()
    td.destroy_process_group()


def test_data_parallel_fit_takes_the_same_decisions_on_every_rank(tmp_path):
    """ADVICE (round 1): LR scheduling / early stopping acted on rank-local validation metrics, so one
    rank could reduce its LR or leave the loop alone.  Epoch metrics are now averaged over the ranks."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_fit_worker, args=(world, port, out, str(tmp_path)), nprocs=world, join=True)
    assert out[0] == out[1], f"ranks diverged: {dict(out)}"
    assert out[0][0] >= 3
