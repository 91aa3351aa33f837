#!/usr/bin/env python
"""Benchmark of the SCGLUE training step (BASELINE.json metric: train cells/s, with the fused
decoder's fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfgN]

A *step* is one ``train_step`` (discriminator update + generator update, RMSprop) on one
minibatch per rank.  "A cell" = one view row of the dataset (one row per modality).

Workloads (BASELINE.json ``configs``; ``--config``, default ``cfg2`` = the configuration the
metric is quoted on; the others are extra lines, run through ``gpurun`` and kept under
``profiles/``):

  cfg1  SCGLUE RNA + ATAC, unpaired, PBMC shape (10 k + 10 k cells, 2 000 genes, 50 000 peaks,
        100 k-edge guidance graph + self-loops), NB decoders, batch 128
  cfg2  PairedSCGLUE, same shape, all cells paired, joint-cross / real-cross / cosine terms
  cfg3  triple omics: RNA NB 2 000 + ATAC NB 150 000 + methylation ZILN 48 000 (sign -1 edges),
        200 k vertices
  cfg4  ZINB decoders with 20 batches (``use_batch``), PBMC feature shape
  cfg5  atlas shape: 4 000 HVGs + 196 000 peaks, h_dim 512, batch 512 per GPU, 1.2 M edges

The step cost does not depend on how many cells the dataset holds, so cfg3-cfg5 materialise a
bounded number of synthetic cells (``config.cells_materialised``); features, graph, batch and
model sizes are the configuration's.

``value``   : steps over minibatches already resident in HBM (device-resident pipeline).
``e2e``     : the same steps through the reference-facing call ``trainer.train_step(engine,
              host_batch)`` with pinned host tensors: H2D of the minibatch and a D2H read of the
              loss are inside the timed region.
``fit_e2e`` : wall-clock cells/s of a real ``model.fit`` (negative sampler, loaders, validation,
              plugins) on the materialised dataset.
``roofline``: the dominant kernel sequence -- the fused decoder call of the largest modality as
              the step issues it (all its loss terms in one launch sequence) -- timed with CUDA
              events on the launching stream (the step itself is one CUDA-graph launch, so the
              sequence is timed in its own region of the same run); algorithmic bytes / time vs
              the measured HBM copy bandwidth.
``cpu_baseline`` / ``--impl reference``: the CPU port of the reference's train_step
              (``oracle/scglue_cpu.py``, pinned against the unmodified reference) on the host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # (see glue_b200/__init__.py)

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# modality: (name, feature prefix, prob_model, n_features, density of the count matrix)
_PBMC = [("rna", "g", "NB", 2000, 0.26), ("atac", "p", "NB", 50000, 0.05)]
WORKLOADS = {
    "cfg1": dict(name="cfg1_scglue_pbmc", paired=False, modalities=_PBMC, n_cells=10000, cells_total=20000,
                 n_pairs=50000, h_dim=256, batch=128),
    "cfg2": dict(name="cfg2_paired_scglue_pbmc", paired=True, modalities=_PBMC, n_cells=10000, cells_total=10000,
                 n_pairs=50000, h_dim=256, batch=128),
    "cfg3": dict(name="cfg3_triple_omics", paired=False,
                 modalities=[("rna", "g", "NB", 2000, 0.26), ("atac", "p", "NB", 150000, 0.02),
                             ("met", "m", "ZILN", 48000, 0.5)],
                 n_cells=4000, cells_total=30000, n_pairs=150000, h_dim=256, batch=128),
    "cfg4": dict(name="cfg4_zinb_batches", paired=False,
                 modalities=[("rna", "g", "ZINB", 2000, 0.26), ("atac", "p", "ZINB", 50000, 0.05)],
                 n_cells=10000, cells_total=100000, n_pairs=50000, h_dim=256, batch=128, n_batches=20),
    "cfg5": dict(name="cfg5_atlas", paired=False,
                 modalities=[("rna", "g", "NB", 4000, 0.15), ("atac", "p", "NB", 196000, 0.01)],
                 n_cells=4000, cells_total=4700000, n_pairs=500000, h_dim=512, batch=512, align_burnin=0),
}
for _w in WORKLOADS.values():
    _w.update(latent_dim=50, rep_dim=100, neg_samples=10, seed=0)


# ----------------------------------------------------------------------- workload
def make_workload(cfg):
    """Seeded synthetic data: CSR fp32 counts (log-normal values with exact zeros for the ZILN
    modality), 100-d encoder inputs, guidance graph: gene--peak pairs (sign +1), one gene per
    methylation feature (sign -1), self-loops."""
    import networkx as nx
    import pandas as pd
    import scipy.sparse

    from glue_b200.anndata_lite import AnnData

    rs = np.random.RandomState(cfg["seed"])
    n = cfg["n_cells"]
    adatas, names = {}, {}
    for key, prefix, prob, f, rate in cfg["modalities"]:
        cells = [f"cell{i}" for i in range(n)] if cfg["paired"] else [f"{key}_cell{i}" for i in range(n)]
        nnz = int(n * f * rate)
        rows, cols = rs.randint(0, n, nnz), rs.randint(0, f, nnz)
        if prob == "ZILN":
            vals = np.exp(rs.randn(nnz)).astype(np.float32)
        else:
            vals = (1 + rs.poisson(0.3, nnz)).astype(np.float32)
        x = scipy.sparse.coo_matrix((vals, (rows, cols)), shape=(n, f)).tocsr()
        x.sum_duplicates()
        x = (x + scipy.sparse.csr_matrix((np.ones(n, np.float32), (np.arange(n), np.zeros(n, int))),
                                         shape=(n, f))).tocsr().astype(np.float32)  # no empty library
        names[key] = [f"{prefix}{i}" for i in range(f)]
        obs = pd.DataFrame(index=cells)
        if cfg.get("n_batches"):
            obs["batch"] = rs.randint(0, cfg["n_batches"], n).astype(str)
        adatas[key] = AnnData(X=x, obs=obs, var=pd.DataFrame(index=names[key]),
                              obsm={"X_rep": rs.randn(n, cfg["rep_dim"]).astype(np.float32)})
    g = nx.Graph()
    for key in names:
        g.add_nodes_from(names[key])
    genes = names["rna"]
    peaks = names["atac"]
    src, dst = rs.randint(0, len(genes), cfg["n_pairs"]), rs.randint(0, len(peaks), cfg["n_pairs"])
    wts = rs.uniform(0.05, 1.0, cfg["n_pairs"])
    g.add_edges_from((genes[s], peaks[t], {"weight": float(w), "sign": 1}) for s, t, w in zip(src, dst, wts))
    if "met" in names:  # gene-body methylation represses: weight 1, sign -1
        met = names["met"]
        owner = rs.randint(0, len(genes), len(met))
        g.add_edges_from((genes[s], m, {"weight": 1.0, "sign": -1}) for s, m in zip(owner, met))
    g.add_edges_from((v, v, {"weight": 1.0, "sign": 1}) for v in list(g.nodes))
    return adatas, g


def build_model(cfg, adatas, graph):
    from glue_b200 import models

    for (key, _, prob, _, _) in cfg["modalities"]:
        models.configure_dataset(adatas[key], prob, use_highly_variable=False, use_rep="X_rep",
                                 use_obs_names=cfg["paired"], use_batch="batch" if cfg.get("n_batches") else None)
    cls = models.PairedSCGLUEModel if cfg["paired"] else models.SCGLUEModel
    # (vertices sorted as fit_SCGLUE does, scglue/models/__init__.py:217-273)
    model = cls(adatas, sorted(graph.nodes), latent_dim=cfg["latent_dim"], h_dim=cfg["h_dim"], random_seed=cfg["seed"])
    model.compile()
    return model


def make_loaders(model, adatas, graph, cfg, device_resident):
    """Datasets + loaders exactly as ``SCGLUEModel.fit`` / ``GLUETrainer.fit`` build them."""
    from glue_b200.models.data import AnnDataset, GraphDataset
    from glue_b200.utils import config

    net, trainer = model.net, model.trainer
    data = AnnDataset([adatas[k] for k in net.keys], [model.modalities[k] for k in net.keys], mode="train")
    gds = GraphDataset(graph, model.vertices, neg_samples=cfg["neg_samples"],
                       weighted_sampling=True, deemphasize_loops=True)
    if device_resident:
        data.to_device(net.device)
        gds.to_device(net.device)
    graph_batch = math.ceil(gds.size / model.GRAPH_BATCHES)
    fetches = config.DATALOADER_FETCHES_PER_BATCH
    data.getitem_size = max(1, round(cfg["batch"] / fetches))
    gds.getitem_size = max(1, round(graph_batch / fetches))
    data.prepare_shuffle(num_workers=0, random_seed=cfg["seed"])
    gds.prepare_shuffle(num_workers=0, random_seed=cfg["seed"])
    loader = trainer._loaders(data, gds, fetches, fetches, cfg["seed"], True, True, shard=True)
    return gds, loader, graph_batch


def config_block(cfg, world, gds_edges=None, e_mb=None, extra=None):
    """The ``config`` object of the JSON line: identical keys for both arms."""
    feats = {key: f for key, _, _, f, _ in cfg["modalities"]}
    out = {"workload": cfg["name"], "trainer": "PairedSCGLUE" if cfg["paired"] else "SCGLUE",
           "prob_models": {key: prob for key, _, prob, _, _ in cfg["modalities"]},
           "features": feats, "vertices": int(sum(feats.values())), "cells": cfg["cells_total"],
           "cells_materialised": cfg["n_cells"] * (1 if cfg["paired"] else len(feats)),
           "n_batches": cfg.get("n_batches", 0), "h_dim": cfg["h_dim"], "latent_dim": cfg["latent_dim"],
           "graph_edges": gds_edges, "edge_minibatch": e_mb, "batch_per_gpu": cfg["batch"],
           "global_batch": cfg["batch"] * world, "parallelism": f"dp{world}"}
    out.update(extra or {})
    return out


class _Engine:
    class state:
        epoch = 1


# ------------------------------------------------------------------- clock sampling
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, r[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------- CPU oracle arm
def cpu_step_runner(cfg, model, host_batches, gds, threads):
    """Returns ``run(n) -> seconds`` executing n train steps of the CPU port of the reference."""
    from glue_b200.num import normalize_edges
    from oracle import scglue_cpu as oc

    torch.set_num_threads(threads)
    net = model.net
    state = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    probs = {key: prob for key, _, prob, _, _ in cfg["modalities"]}
    paired = cfg["paired"]
    burnin = cfg.get("align_burnin", 57)
    step_cfg = oc.StepConfig(net.keys, probs, {k: probs[k] in ("NB", "ZINB") for k in net.keys},
                             {k: max(cfg.get("n_batches", 0), 1) for k in net.keys}, paired=paired,
                             lam_joint_cross=0.02 if paired else 0.0, lam_real_cross=0.02 if paired else 0.0,
                             lam_cos=0.02 if paired else 0.0, align_burnin=burnin or None)
    full = (torch.as_tensor(gds.eidx), torch.as_tensor(normalize_edges(gds.eidx, gds.ewt)),
            torch.as_tensor(gds.esgn))
    trainer = oc.CpuTrainer(step_cfg, state, full, lr=2e-3)
    K = len(net.keys)

    def to_batch(data):
        from glue_b200.models.data import is_packed_sparse, unpack_sparse_rows
        data = [t.cpu() for t in data]
        for i, k in enumerate(net.keys):  # (host batches carry sparse count rows packed)
            if is_packed_sparse(data[i]):
                data[i] = unpack_sparse_rows(data[i], data[2 * K + i].shape[0], getattr(net, f"{k}_idx").shape[0])
        groups = [dict(zip(net.keys, data[i * K:(i + 1) * K])) for i in range(5)]
        rest = data[5 * K:]
        pmsk = rest[0] if len(rest) == 4 else None
        return dict(x=groups[0], xrep=groups[1], xbch=groups[2], xdwt=groups[4], pmsk=pmsk,
                    eidx=rest[-3], ewt=rest[-2], esgn=rest[-1])

    batches = [to_batch(b) for b in host_batches]

    def run(n, offset=0):
        tic = time.perf_counter()
        for i in range(n):
            trainer.train_step(batches[(offset + i) % len(batches)], epoch=1)
        return time.perf_counter() - tic

    return run


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profile(name):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(name)
    return None


def decoder_bytes(B, F, K, n_param_tables, n_batches, n_terms):
    """x rows count once per term, v / dv and the parameter tables once per fused call."""
    return (n_terms * B * F * 4 + 2 * F * K * 4 + 2 * n_param_tables * n_batches * F * 4
            + n_terms * (2 * B * K * 4 + 16 * B))


# ------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--no-fit", action="store_true", help="skip the fit_e2e measurement")
    ap.add_argument("--partition-graph", action="store_true",
                    help="N > 1: run the steps with the edge-partitioned GraphConv (row-block aggregate + all-gather)")
    ap.add_argument("--profile", action="store_true",
                    help="only the device-resident steps (short run for ncu launch lists)")
    args = ap.parse_args()
    cfg = dict(WORKLOADS[args.config])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        return run_reference(args, cfg, world)

    import __graft_entry__ as entry
    if local_rank == 0:
        entry.build()
    from glue_b200 import dist as gdist
    from glue_b200 import ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    if args.partition_graph:
        from glue_b200.utils import config as _cfg
        _cfg.PARTITION_GRAPH = True
    torch.cuda.set_device(local_rank)
    gdist.init_from_env()
    if world > 1:
        torch.distributed.barrier()
    device = torch.device("cuda", local_rank)
    warmup = max(args.warmup, 3)

    adatas, graph = make_workload(cfg)
    model = build_model(cfg, adatas, graph)
    trainer, net = model.trainer, model.net
    gds, dev_loader, graph_batch = make_loaders(model, adatas, graph, cfg, device_resident=True)
    trainer.set_full_graph(gds)
    view_cells = cfg["n_cells"] * (1 if cfg["paired"] else len(cfg["modalities"]))
    auto_burnin = max(math.ceil(8.0 / 2e-3 / (view_cells * 0.9 / cfg["batch"])), 8)  # AUTO rule
    trainer.align_burnin = cfg.get("align_burnin", auto_burnin)

    # pools of distinct minibatches: larger than L2 in total
    pool_n = 8
    it = iter(dev_loader)
    dev_pool = [next(it) for _ in range(pool_n)]
    _, host_loader, _ = make_loaders(model, adatas, graph, cfg, device_resident=False)
    it = iter(host_loader)
    host_pool = [[t.pin_memory() for t in next(it)] for _ in range(pool_n)]
    if os.environ.get("GLUE_BENCH_SHORT_RANK") == str(rank):
        # test knob: give this rank alone a short minibatch (as the block-shuffled loader can),
        # to exercise the rank-uniform warm-up below
        n_short = cfg["batch"] - 16
        cut = lambda batch, pin: [(t[:n_short].contiguous().pin_memory() if pin else t[:n_short].contiguous())
                                  if t.dim() and t.shape[0] == cfg["batch"] else t for t in batch]
        dev_pool[3], host_pool[5] = cut(dev_pool[3], False), cut(host_pool[5], True)
    h2d = sum(t.numel() * t.element_size() for t in host_pool[0])
    pool_mb = pool_n * h2d / 1e6
    e_mb = int(dev_pool[0][-1].numel())

    def sync():
        torch.cuda.synchronize(device)
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize(device)

    def timed(fn, steps):
        sync()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for i in range(steps):
            fn(i)
        stop.record()
        sync()
        ms = start.elapsed_time(stop)
        if world > 1:
            t = torch.tensor([ms], device=device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # --- device-resident steps -------------------------------------------------------
    # (after two eager steps the trainer replays the step from a captured CUDA graph)
    step_dev = lambda i: trainer.train_step(_Engine, dev_pool[i % pool_n])
    with ClockSampler(local_rank) as clocks:  # sampled from warm-up on (the timed region is short)
        # every distinct minibatch signature of the pool (the block-shuffled loader can yield a
        # short minibatch) is stepped until its graph exists; at least `warmup` steps in any case
        n_sig = len({tuple(tuple(t.shape) for t in b) for b in dev_pool + host_pool})
        if world > 1:  # every rank must take the same number of steps (gradient all-reduce)
            t_sig = torch.tensor([n_sig], device=device)
            torch.distributed.all_reduce(t_sig, op=torch.distributed.ReduceOp.MAX)
            n_sig = int(t_sig.item())
        warmup = max(warmup, 4, 3 * pool_n if n_sig > 1 else 0)
        for i in range(warmup):
            step_dev(i)
        launches0, replays0 = ops.kernel_launches(), trainer.graphs.replays
        if args.profile:  # (`ncu --profile-from-start off`: only the replayed steps are profiled)
            torch.cuda.synchronize(device)
            torch.cuda.cudart().cudaProfilerStart()
        ms = timed(step_dev, args.steps)
        if args.profile:
            torch.cuda.cudart().cudaProfilerStop()
        ginfo = trainer.graphs.info()
        launches = (ops.kernel_launches() - launches0
                    + (trainer.graphs.replays - replays0) * int(ginfo["kernels_per_step"] or 0))
    cells = cfg["batch"] * args.steps * world
    value = cells / (ms * 1e-3)
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms / args.steps, "value": value,
                              "cuda_graph": dict(ginfo)}))
        return

    # --- end to end: host batches through the reference-facing calls ----------------------
    # as GLUETrainer.fit drives them: pinned host minibatches -> DevicePrefetcher (H2D of batch
    # i + 1 on a side stream while step i runs) -> trainer.train_step -> D2H read of the loss
    from glue_b200.models.data import DevicePrefetcher

    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event() for _ in range(2)]

    def run_e2e(n, prefetch):
        """Every step's loss is read back to the host (pinned, asynchronous D2H): the value of
        step i is consumed while step i + 1 is already queued, as a training loop that logs
        metrics does."""
        feed = (host_pool[i % pool_n] for i in range(n))
        if prefetch:  # (the loader object lives as long as the training loop: built once, like its stream)
            prefetcher.loader = feed
            feed = prefetcher
        seen = []
        marks = e2e_marks[prefetch][:n]  # where the GPU time of the region goes: inside the steps / between them
        for i, batch in enumerate(feed):
            marks[i][0].record()
            losses = trainer.train_step(_Engine, batch)
            marks[i][1].record()
            loss_host[i & 1].copy_(losses["vae_loss"].reshape(1), non_blocking=True)
            loss_done[i & 1].record()
            if i:
                loss_done[(i - 1) & 1].synchronize()
                seen.append(float(loss_host[(i - 1) & 1]))
        loss_done[(n - 1) & 1].synchronize()
        seen.append(float(loss_host[(n - 1) & 1]))
        assert len(seen) == n and all(v == v for v in seen), "every step's loss must reach the host"
        return seen[-1]

    prefetcher = DevicePrefetcher(None, device)
    n_marks = max(args.steps, 2 * pool_n)
    e2e_runs = {}
    e2e_marks = {p: [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                     for _ in range(n_marks)] for p in (False, True)}
    for mode, prefetch in (("direct", False), ("prefetch", True)):
        run_e2e(2 * pool_n if n_sig > 1 else 4, prefetch)  # (every signature of the host pool has its graph)
        e2e_runs[mode] = timed(lambda i, p=prefetch: run_e2e(args.steps, p) if i == 0 else None, 1)
    e2e_feed = min(e2e_runs, key=e2e_runs.get)
    ms_e2e = e2e_runs[e2e_feed]
    marks = e2e_marks[e2e_feed == "prefetch"][:args.steps]
    e2e_step_gpu_ms = sum(a.elapsed_time(b) for a, b in marks) / len(marks)
    e2e_gap_ms = sum(marks[i][1].elapsed_time(marks[i + 1][0]) for i in range(len(marks) - 1)) / max(1, len(marks) - 1)
    e2e_value = cells / (ms_e2e * 1e-3)

    # --- N > 1: edge-partitioned GraphConv against the replicated one, atlas graph shape ------
    partitioned = measure_partitioned(device, world, rank) if world > 1 else None

    if rank != 0:
        return

    # --- the native kernels, timed alone on this stream -----------------------------------
    # CUDA events around back-to-back calls queued behind a blocker (so the host launch path is
    # off the clock); x cycles through the pooled minibatches (more than L2 in total).
    B, K = cfg["batch"], cfg["latent_dim"]
    keys = net.keys
    x_of = {k: [batch[i] for batch in dev_pool] for i, k in enumerate(keys)}
    xbch_of = {k: [batch[2 * len(keys) + i] for batch in dev_pool] for i, k in enumerate(keys)}

    def time_calls(fn, n=24):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for i in range(n):
            fn(i)
        host_s = time.perf_counter() - t0  # what the Python launch path needs to enqueue the n calls
        torch.cuda.synchronize(device)
        # a spinning blocker the n launches queue behind, long enough to cover their enqueue time (a GEMM
        # blocker would pull the SM clock down under the power cap and inflate what is timed behind it)
        torch.cuda._sleep(max(12_000_000, int(host_s * 1.5 * 2.0e9)))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize(device)
        return a.elapsed_time(b) * 1e3 / n

    ktimes, kbytes = {}, {}
    vtab = (torch.randn(len(model.vertices), K, device=device) * 0.3).requires_grad_(True)
    n_terms = 3 if (cfg["paired"] and len(keys) == 2) else 1
    for k in keys:
        dec, idx = net.u2x[k], getattr(net, f"{k}_idx")
        fam = dec.FAMILY
        Fk = x_of[k][0].shape[1]
        us = [(torch.randn(B, K, device=device) * 0.5).requires_grad_(True) for _ in range(n_terms)]
        rws = [None] + [torch.rand(B, device=device) / (B * Fk) for _ in range(n_terms - 1)]
        need_l = fam in ("NB", "ZINB")
        ls = [x.sum(1, keepdim=True) if need_l else None for x in x_of[k]]
        disp = getattr(dec, dec.DISPERSION)
        zi = dec.zi_logits if dec.ZERO_INFLATED else None
        nb = dec.scale_lin.shape[0]
        npar = 4 if zi is not None else 3

        def dec_single(i, dec=dec, idx=idx, us=us, ls=ls, k=k, fam=fam, disp=disp, zi=zi):
            xb = x_of[k][i % pool_n]
            n = xb.shape[0]  # the block-shuffled loader can yield a short minibatch
            return ops.decoder_nll(fam, us[0][:n], vtab, dec.scale_lin, dec.bias, disp, zi,
                                   b=xbch_of[k][i % pool_n], l=ls[i % pool_n], x=xb, vidx=idx, reduce="mean")

        def dec_fused(i, dec=dec, idx=idx, us=us, ls=ls, k=k, fam=fam, disp=disp, zi=zi, rws=rws):
            xb = x_of[k][i % pool_n]
            n = xb.shape[0]
            return ops.decoder_nll_multi(fam, [u[:n] for u in us], vtab, dec.scale_lin, dec.bias, disp, zi,
                                         b=xbch_of[k][i % pool_n], l=ls[i % pool_n], x=xb, vidx=idx, reduce="mean",
                                         row_weights=[None if w is None else w[:n] for w in rws],
                                         upstreams=[1.0, 0.02, 0.02][:n_terms])

        name1 = f"decoder_fwd_bwd_{fam}_F{Fk}_B{B}"
        ktimes[name1] = (time_calls(dec_single), 0 if n_terms > 1 else 1)
        kbytes[name1] = decoder_bytes(B, Fk, K, npar, nb, 1)
        if n_terms > 1:
            name3 = f"decoder_fused{n_terms}_fwd_bwd_{fam}_F{Fk}_B{B}"
            ktimes[name3] = (time_calls(dec_fused), 1)
            kbytes[name3] = decoder_bytes(B, Fk, K, npar, nb, n_terms)
    V, E = len(model.vertices), int(gds.eidx.shape[1])
    csr = ops.csr_for(trainer.eidx, trainer.enorm, trainer.esgn, V)
    ktimes["graphconv_fwd"] = (time_calls(lambda i: ops._spmm(csr.fwd, vtab.detach())), 2)
    ktimes["graphconv_bwd"] = (time_calls(lambda i: ops._spmm(csr.bwd, vtab.detach())), 1)
    kbytes["graphconv_fwd"] = kbytes["graphconv_bwd"] = 2 * V * K * 4 + 8 * E + 4 * (V + 1)
    eb = dev_pool[0]
    s_eidx, s_ewt, s_esgn = eb[-3], eb[-2], eb[-1]

    def gnll(i):
        loss = ops.graph_nll(vtab, s_eidx, s_esgn, s_ewt)
        (g,) = torch.autograd.grad(loss, vtab)
        return g

    ktimes["graph_nll_fwd_bwd"] = (time_calls(gnll), 1)
    kbytes["graph_nll_fwd_bwd"] = e_mb * 16 + min(2 * e_mb, V) * K * 4 + V * K * 4

    # --- roofline of the dominant kernel sequence ------------------------------------------
    peak, peak_src = peaks()
    step_us = ms * 1e3 / args.steps
    dom = max((n for n in ktimes if n.startswith("decoder") and ktimes[n][1]), key=lambda n: ktimes[n][0])
    us_dom, per_step = ktimes[dom]
    roof = {"bound": "hbm", "kernel": f"{dom}: fused decoder call of the largest modality as the step issues it "
                                       f"({n_terms} loss term(s) in one launch sequence), tcgen05 + TMA path",
            "achieved": kbytes[dom] / (us_dom * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": kbytes[dom] / (us_dom * 1e-6) / 1e9 / peak, "traffic": traffic_from_profile(dom),
            "peak_source": peak_src, "algorithmic_bytes": kbytes[dom], "launch_us": us_dom,
            "calls_per_step": per_step, "share_of_step": per_step * us_dom / step_us,
            "how": "CUDA events around 24 back-to-back calls on the launching stream behind a spinning blocker "
                   "that outlasts their enqueue time, "
                   f"x from {pool_n} pooled minibatches (HBM resident, {pool_mb:.0f} MB > L2)"}
    kernels = {name: {"us_per_call": t, "calls_per_step": n, "share_of_step": n * t / step_us,
                      "algorithmic_bytes": kbytes.get(name),
                      "frac_of_hbm_peak": kbytes[name] / (t * 1e-6) / 1e9 / peak if name in kbytes else None}
               for name, (t, n) in sorted(ktimes.items())}

    # --- a real fit: sampler + loaders + validation + plugins, wall clock -------------------
    fit_e2e = None
    if not args.no_fit and world == 1:
        fit_e2e = measure_fit(cfg, adatas, graph, device)

    # --- CPU baseline (bounded sample of the same workload) --------------------------------
    cpu = None
    if not args.skip_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        run = cpu_step_runner(cfg, model, host_pool, gds, cores)
        run(1)
        n_cpu = 5 if cfg["batch"] * max(f for *_, f, _ in cfg["modalities"]) < 2e7 else 2
        secs = run(n_cpu, offset=1)
        cpu = {"value": cfg["batch"] * n_cpu / secs, "unit": "cells/s", "cores": cores,
               "kind": "port", "ms_per_step": secs * 1e3 / n_cpu,
               "sample": f"{n_cpu} train steps (1 warm-up) of the same workload, "
                         f"torch CPU, {cores} threads"}

    out = {
        "metric": "train_cells_per_s", "value": value, "unit": "cells/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_block(cfg, world, E, e_mb, {
            "parallelism": f"dp{world}" + ("+graph-partition" if args.partition_graph else ""),
            "l2": f"inputs cycle through {pool_n} minibatches ({pool_n} x {h2d / 1e6:.0f} MB > 126 MB L2)"}),
        "e2e": {"value": e2e_value, "unit": "cells/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "feed": e2e_feed,
                "ms_per_step_by_feed": {k: v / args.steps for k, v in e2e_runs.items()},
                "gpu_ms_inside_steps": e2e_step_gpu_ms, "gpu_ms_between_steps": e2e_gap_ms},
        "fit_e2e": fit_e2e,
        "gpu_launches": int(launches),
        "cuda_graph": {k: ginfo[k] for k in ("graphs", "kernels_per_step", "disabled")},
        "clocks": clocks.summary(),
        "roofline": roof,
        "kernels": kernels,
        "partitioned_graphconv": partitioned,
        "cpu_baseline": cpu,
    }
    print(json.dumps(out))


def measure_fit(cfg, adatas, graph, device):
    """Wall-clock throughput of ``model.fit`` itself: negative sampler, block-shuffled loaders,
    device-resident row gathers, per-epoch validation, metric read-back, plugins.  A fresh model is
    fitted once for one epoch (graph capture, allocator warm-up) and once more for nine epochs;
    ``value`` is the steady state (the last four epochs of the second fit, wall clock between the ends of
    epochs); ``epoch_s`` lists every epoch of that fit after the first, so the epochs that still capture a
    step graph (the one short minibatch of the training and of the validation split is a signature of its
    own and is captured at its third occurrence) are on record; ``setup_s`` is what precedes the first step:
    dataset and graph construction, upload."""
    from glue_b200.models.base import EPOCH_COMPLETED, EPOCH_STARTED, TrainingPlugin
    from glue_b200.utils import config

    class EpochClock(TrainingPlugin):
        def __init__(self):
            self.started, self.ended = [], []

        def attach(self, net, trainer, train_engine, val_engine, train_loader, val_loader, directory):
            def end(engine):
                torch.cuda.synchronize(device)
                self.ended.append(time.perf_counter())
            train_engine.add_event_handler(EPOCH_STARTED, lambda engine: self.started.append(time.perf_counter()))
            train_engine.add_event_handler(EPOCH_COMPLETED, end)

    model = build_model(cfg, adatas, graph)
    view_cells = cfg["n_cells"] * (1 if cfg["paired"] else len(cfg["modalities"]))
    kw = dict(data_batch_size=cfg["batch"], patience=None, reduce_lr_patience=None)
    if "align_burnin" in cfg:
        kw["align_burnin"] = cfg["align_burnin"]
    old = config.PRINT_LOSS_INTERVAL
    config.PRINT_LOSS_INTERVAL = 1000
    try:
        tic = time.perf_counter()
        first_clock = EpochClock()
        run_fit(model, adatas, graph, cfg, kw, 1, first_clock)
        torch.cuda.synchronize(device)
        first = time.perf_counter() - tic
        epochs, timed = 9, 4
        clock = EpochClock()
        tic = time.perf_counter()
        run_fit(model, adatas, graph, cfg, kw, epochs, clock)
        torch.cuda.synchronize(device)
        wall = time.perf_counter() - tic
    finally:
        config.PRINT_LOSS_INTERVAL = old
    fetches = config.DATALOADER_FETCHES_PER_BATCH
    block = max(1, round(cfg["batch"] / fetches))
    n_train = round(view_cells * 0.9)
    steps = (math.ceil(n_train / block)) // fetches  # drop_last over blocks
    steady_s = clock.ended[-1] - clock.ended[-1 - timed]
    trained = steps * cfg["batch"] * timed
    return {"value": trained / steady_s, "unit": "cells/s", "epochs_timed": timed,
            "train_steps_per_epoch": steps, "s_per_epoch": steady_s / timed,
            "epoch_s": [round(b - a, 4) for a, b in zip(clock.ended, clock.ended[1:])],
            "setup_s": clock.started[0] - tic, "fit_wall_s": wall, "first_fit_s_1_epoch": first,
            "includes": "per epoch: negative sampling of the guidance graph, block-shuffled loaders, device row "
                        "gathers, the training steps, validation on the held-out 10 %, metric read-back; "
                        "setup_s: dataset + graph construction and upload before the first step"}


def run_fit(model, adatas, graph, cfg, kw, epochs, clock):
    """``model.fit`` with an extra plugin (the public ``fit`` takes none): same code path, the clock rides
    along through ``GLUETrainer.fit(plugins=...)``."""
    from glue_b200.models import glue as glue_mod
    orig = glue_mod.GLUETrainer.fit

    def fit_with_clock(self, *a, **k):
        k["plugins"] = list(k.get("plugins") or []) + [clock]
        return orig(self, *a, **k)

    glue_mod.GLUETrainer.fit = fit_with_clock
    try:
        model.fit(adatas, graph, max_epochs=epochs, **kw)
    finally:
        glue_mod.GLUETrainer.fit = orig


def measure_partitioned(device, world, rank, V=200000, E=1000000, K=50, n=20):
    """GraphConv forward at the atlas graph shape: replicated (every rank aggregates all rows) vs
    edge-partitioned by target-row blocks + all-gather (``ops.PartitionedGraphCSR``); device time, max
    over ranks."""
    from glue_b200 import ops
    rs = np.random.RandomState(1)
    src = np.concatenate([rs.randint(0, V, E), np.arange(V)])
    dst = np.concatenate([rs.randint(0, V, E), np.arange(V)])
    eidx = torch.as_tensor(np.stack([src, dst]).astype(np.int64), device=device)
    enorm = torch.rand(eidx.shape[1], device=device)
    esgn = torch.ones(eidx.shape[1], device=device)
    csr = ops.GraphCSR(eidx, enorm, esgn, V)
    part = ops.PartitionedGraphCSR(csr, rank, world)
    xs = [torch.randn(V, K, device=device) for _ in range(8)]

    def run(fn):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize(device)
        torch.distributed.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize(device)
        t = torch.tensor([a.elapsed_time(b) * 1e3 / n], device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        rep = run(lambda i: ops.graph_conv(xs[i % 8], csr))
        par = run(lambda i: ops.graph_conv_partitioned(xs[i % 8], part))
        same = bool(torch.equal(ops.graph_conv(xs[0], csr), ops.graph_conv_partitioned(xs[0], part)))
    bytes_ = 2 * V * K * 4 + 8 * int(eidx.shape[1]) + 4 * (V + 1)
    return {"vertices": V, "edges": int(eidx.shape[1]), "replicated_us": rep, "partitioned_us": par,
            "bit_identical": same, "algorithmic_bytes": bytes_,
            "note": "forward aggregation only; partitioned = local row block + all-gather of [V/N, K] blocks"}


def run_reference(args, cfg, world):
    """The reference's CPU path (port pinned against the unmodified reference) on host cores."""
    from glue_b200.utils import config
    config.CPU_ONLY = True
    adatas, graph = make_workload(cfg)
    model = build_model(cfg, adatas, graph)
    gds, host_loader, _ = make_loaders(model, adatas, graph, cfg, device_resident=False)
    it = iter(host_loader)
    host_pool = [next(it) for _ in range(8)]
    e_mb = int(host_pool[0][-1].numel())
    cores = os.cpu_count() or 1
    run = cpu_step_runner(cfg, model, host_pool, gds, cores)
    warmup = max(args.warmup, 3)           # (the same rule as the GPU arm)
    steps = max(1, min(args.steps, 40))    # bounded: ~0.3 s per step on 16 cores
    run(warmup)
    secs = run(steps, offset=warmup)
    value = cfg["batch"] * steps / secs
    sample = (f"{steps} train steps ({warmup} warm-up) of the same workload; CPU port of the "
              f"reference train_step, torch CPU, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": "train_cells_per_s", "value": value, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": secs * 1e3 / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_block(cfg, max(world, args.gpus), int(gds.eidx.shape[1]), e_mb, {
            "parallelism": f"dp{max(world, args.gpus)}",
            "l2": "host arm: CPU caches, no flush"}),
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }))


def _only_json_on_stdout(fn):
    """The contract is ONE JSON line on stdout: libraries (NCCL prints its version banner there)
    get stderr as their fd 1 for the whole run, the result line goes to the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real, "w")
    captured = []
    import builtins
    orig_print = builtins.print

    def tee(*a, **k):
        if k.get("file") in (None, sys.stdout):
            captured.append(" ".join(str(x) for x in a))
        else:
            orig_print(*a, **k)

    builtins.print = tee
    try:
        fn()
    finally:
        builtins.print = orig_print
        sys.stdout.flush()
        for line in captured:
            out.write(line + "\n")
        out.flush()


if __name__ == "__main__":
    _only_json_on_stdout(main)
