"""Debug driver: two fits in a row on a bench workload (python tools/repro_fit.py cfg3 600 [probe])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
cfgname = sys.argv[1]
cfg = dict(bench.WORKLOADS[cfgname]); cfg["n_cells"] = int(sys.argv[2])
probe = len(sys.argv) > 3 and sys.argv[3] == "probe"
from glue_b200.utils import config as _config
from glue_b200 import ops as _ops
for item in filter(None, os.environ.get("GLUE_CFG", "").split(",")):
    name, val = item.split("=")
    setattr(_config, name, bool(int(val)))
if os.environ.get("NOEMPTY"):
    torch.cuda.empty_cache = lambda: None
_ops.DECODER_IMPL = int(os.environ.get("IMPL", "0"))
_orig_GraphCSR_init = _ops.GraphCSR.__init__


def _logged_init(self, eidx, enorm, esgn, vnum):
    print("GraphCSR build: capturing =", torch.cuda.is_current_stream_capturing(), "eidx ptr", hex(eidx.data_ptr()),
          "enorm ptr", hex(enorm.data_ptr()), "cache size", len(_ops._CSR_CACHE), flush=True)
    _orig_GraphCSR_init(self, eidx, enorm, esgn, vnum)


_ops.GraphCSR.__init__ = _logged_init
adatas, graph = bench.make_workload(cfg)
model = bench.build_model(cfg, adatas, graph)
kw = dict(data_batch_size=cfg["batch"], patience=None, reduce_lr_patience=None)
model.fit(adatas, graph, max_epochs=1, **kw)
torch.cuda.synchronize()
print("fit ok 1", flush=True)
if probe:
    from glue_b200 import ops
    from glue_b200.models.data import GraphDataset
    tr = model.trainer
    gds = GraphDataset(graph, model.vertices, neg_samples=10)
    tr.set_full_graph(gds)
    torch.cuda.synchronize()
    V = len(model.vertices)
    print("V", V, "E", tr.eidx.shape, "eidx max", int(tr.eidx.max()), "min", int(tr.eidx.min()), flush=True)
    csr = ops.csr_for(tr.eidx, tr.enorm, tr.esgn, V)
    torch.cuda.synchronize()
    for name, (indptr, col, val) in (("fwd", csr.fwd), ("bwd", csr.bwd)):
        ip = indptr.cpu()
        print(name, "indptr[0], [-1]", int(ip[0]), int(ip[-1]), "monotone", bool((ip[1:] >= ip[:-1]).all()),
              "col range", int(col.min()), int(col.max()), "finite", bool(torch.isfinite(val).all()), flush=True)
    out = ops.graph_conv(model.net.g2v.vrepr.detach(), csr)
    torch.cuda.synchronize()
    print("graph_conv ok", float(out.abs().sum()), flush=True)
    with torch.no_grad():
        model.net.eval()
        v = model.net.g2v(tr.eidx, tr.enorm, tr.esgn)
    torch.cuda.synchronize()
    print("g2v ok", flush=True)
    tr.clear_full_graph()
    losses = model.get_losses(adatas, graph)
    torch.cuda.synchronize()
    print("get_losses ok", {k: round(v, 4) for k, v in list(losses.items())[:4]}, flush=True)
import numpy as np
from glue_b200.models.data import GraphDataset as _GD
from glue_b200.num import normalize_edges as _ne
ref = _GD(graph, model.vertices, neg_samples=10)
ref_eidx, ref_enorm, ref_esgn = ref.eidx, _ne(ref.eidx, ref.ewt), ref.esgn
tr = model.trainer
orig_val = tr.val_step
orig_train = tr.train_step
state = {"n": 0}


def check(tag):
    torch.cuda.synchronize()
    ok = (np.array_equal(tr.eidx.cpu().numpy(), ref_eidx), np.array_equal(tr.enorm.cpu().numpy(), ref_enorm),
          np.array_equal(tr.esgn.cpu().numpy(), ref_esgn))
    vr = model.net.g2v.vrepr
    print(tag, "full graph intact:", ok, "vrepr finite", bool(torch.isfinite(vr).all()),
          "eidx ptr", hex(tr.eidx.data_ptr()), flush=True)


def build_check(tag, stream=None):
    V = len(model.vertices)
    ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())
    with ctx:
        csr = _ops.GraphCSR(tr.eidx, tr.enorm, tr.esgn, V)
    torch.cuda.synchronize()
    for name, (indptr, col, val) in (("fwd", csr.fwd), ("bwd", csr.bwd)):
        ip = indptr.cpu()
        print(tag, name, "indptr", int(ip[0]), int(ip[-1]), "monotone", bool((ip[1:] >= ip[:-1]).all()),
              "col", int(col.min()), int(col.max()), "ptrs", hex(indptr.data_ptr()), hex(col.data_ptr()), flush=True)


def val_step(engine, data):
    if state["n"] == 0:
        check("before first validation step")
        build_check("main-stream build")
        build_check("side-stream build", tr._side_streams[-1] if tr._side_streams else torch.cuda.Stream())
        build_check("fresh-stream build", torch.cuda.Stream())
    state["n"] += 1
    return orig_val(engine, data)


tstate = {"n": 0}


def train_step(engine, data):
    if tstate["n"] < 3:
        check(f"before train step {tstate['n']}")
    tstate["n"] += 1
    return orig_train(engine, data)


_orig_spmm = _ops._spmm
_keys = []
_orig_key = tr.graphs._key


def logged_key(data, anneal):
    k = _orig_key(data, anneal)
    if k not in _keys:
        if _keys:
            for i, (a, b) in enumerate(zip(_keys[-1], k)):
                if a != b:
                    print("NEW KEY differs at field", i, ":", str(a)[:300], "->", str(b)[:300], flush=True)
        else:
            print("first key seen in fit 2:", str(k)[:200], flush=True)
        _keys.append(k)
    return k


tr.graphs._key = logged_key
print("entries before fit 2:", len(tr.graphs.entries), [str(k)[:120] for k in tr.graphs.entries], flush=True)
for k in tr.graphs.entries:
    _keys.append(k)


def checked_spmm(csr_part, inp):
    indptr, col, val = csr_part
    if torch.cuda.is_current_stream_capturing():
        return _orig_spmm(csr_part, inp)
    torch.cuda.synchronize()
    ip = indptr.cpu()
    print("spmm check: indptr", int(ip[0]), int(ip[-1]), "monotone", bool((ip[1:] >= ip[:-1]).all()), "n", ip.numel(),
          "col", int(col.min()), int(col.max()), col.numel(), "inp", tuple(inp.shape), inp.is_contiguous(),
          hex(inp.data_ptr()), "stream", torch.cuda.current_stream().cuda_stream, flush=True)
    out = _orig_spmm(csr_part, inp)
    torch.cuda.synchronize()
    print("spmm ok", flush=True)
    return out


_ops._spmm = checked_spmm
tr.val_step = val_step
tr.train_step = train_step
model.fit(adatas, graph, max_epochs=2, **kw)
torch.cuda.synchronize()
print("fit ok 2", flush=True)
