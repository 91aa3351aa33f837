"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the *unmodified* reference hot-path modules from ``/root/reference``
under the alias package name ``refglue`` so that golden vectors can be
generated from the reference itself.  Six third-party
packages the reference imports at module scope are absent from this image
(anndata, pytorch-ignite, pyro-ppl, pybedtools, h5py, parse); none of them is
touched by the functions we call, so they are replaced by empty stand-in
modules.  ``/root/reference`` exists only in the build container: callers must
check :func:`available` first.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GLUE_REFERENCE_ROOT", "/root/reference")
ALIAS = "refglue"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scglue", "models"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__dict__["__stub__"] = True
        sys.modules[name] = mod
    for key, val in attrs.items():
        setattr(mod, key, val)
    return mod


def _install_third_party_stubs() -> None:
    class _Anything:  # attribute sink usable as a base class / decorator target
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, item):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

    if "anndata" not in sys.modules:
        core = _stub("anndata._core")
        sparse_dataset = _stub(
            "anndata._core.sparse_dataset", SparseDataset=type("SparseDataset", (), {})
        )
        core.sparse_dataset = sparse_dataset
        ad = _stub("anndata", AnnData=type("AnnData", (), {}), _core=core)
        ad.__version__ = "0.0-stub"
    if "h5py" not in sys.modules:
        _stub("h5py", Dataset=type("Dataset", (), {}), File=type("File", (), {}))
    if "pybedtools" not in sys.modules:
        helpers = _stub("pybedtools.helpers", set_bedtools_path=lambda *a, **k: None)
        _stub("pybedtools", helpers=helpers, BedTool=type("BedTool", (), {}))
    if "pyro" not in sys.modules:
        dist = _stub("pyro.distributions", BetaBinomial=type("BetaBinomial", (), {}))
        _stub("pyro", distributions=dist)
    if "parse" not in sys.modules:
        _stub("parse", parse=lambda *a, **k: None)
    if "ignite" not in sys.modules:
        events = type("Events", (), {k: k for k in (
            "STARTED", "EPOCH_STARTED", "EPOCH_COMPLETED", "ITERATION_COMPLETED",
            "COMPLETED", "TERMINATE", "EXCEPTION_RAISED")})
        engine = _stub("ignite.engine", Engine=_Anything, Events=events, State=_Anything)
        handlers = _stub("ignite.handlers", TerminateOnNan=_Anything, Checkpoint=_Anything,
                         DiskSaver=_Anything, EarlyStopping=_Anything)
        metrics = _stub("ignite.metrics", Average=_Anything)
        tb = _stub("ignite.contrib.handlers.tensorboard_logger",
                   TensorboardLogger=_Anything, OutputHandler=_Anything)
        ch = _stub("ignite.contrib.handlers", tensorboard_logger=tb)
        contrib = _stub("ignite.contrib", handlers=ch)
        _stub("ignite", engine=engine, handlers=handlers, metrics=metrics, contrib=contrib)


_loaded = None


def load():
    """Return the alias package with ``num``, ``utils`` and ``models.*`` imported."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_third_party_stubs()
    pkg = types.ModuleType(ALIAS)
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "scglue")]
    pkg.__package__ = ALIAS
    sys.modules[ALIAS] = pkg
    models = types.ModuleType(ALIAS + ".models")
    models.__path__ = [os.path.join(REFERENCE_ROOT, "scglue", "models")]
    models.__package__ = ALIAS + ".models"
    sys.modules[ALIAS + ".models"] = models
    pkg.models = models
    for name in ("num", "utils"):
        setattr(pkg, name, importlib.import_module(f"{ALIAS}.{name}"))
    pkg.utils.config.CPU_ONLY = True
    for name in ("nn", "prob", "glue", "sc", "data", "base", "plugins", "scglue"):
        setattr(models, name, importlib.import_module(f"{ALIAS}.models.{name}"))
    _loaded = pkg
    return pkg
