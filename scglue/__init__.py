r"""
Drop-in module paths of the reference package (gao-lab/GLUE, ``scglue`` 0.4.1) for the B200-native
implementation in :mod:`glue_b200`.

``import scglue`` makes ``scglue.models``, ``scglue.models.{scglue,sc,nn,prob,glue,data,base,plugins}``,
``scglue.num``, ``scglue.utils``, ``scglue.graph`` and ``scglue.data`` resolve to the modules of
``glue_b200`` with the same names.  Besides letting user code keep its imports, this is what makes
model files written by the reference loadable: ``Model.save`` dill-pickles the whole model object
(``scglue/models/base.py:364-383``), so class identities such as ``scglue.models.scglue.SCGLUEModel`` or
``scglue.models.sc.NBDataDecoder`` are baked into the file; they resolve to the classes here, whose
attribute layout (parameter / buffer / sub-module names) is the reference's.

Only the training hot path and what it needs is implemented ("Out of scope"); the
genomics / plotting / metrics parts of the reference package are not.
"""
import importlib
import sys

import glue_b200 as _impl

__version__ = _impl.__version__

_ALIASES = {
    "scglue.models": "glue_b200.models",
    "scglue.models.scglue": "glue_b200.models.scglue",
    "scglue.models.sc": "glue_b200.models.sc",
    "scglue.models.nn": "glue_b200.models.nn",
    "scglue.models.prob": "glue_b200.models.prob",
    "scglue.models.glue": "glue_b200.models.glue",
    "scglue.models.data": "glue_b200.models.data",
    "scglue.models.base": "glue_b200.models.base",
    "scglue.models.plugins": "glue_b200.models.plugins",
    "scglue.num": "glue_b200.num",
    "scglue.utils": "glue_b200.utils",
    "scglue.graph": "glue_b200.graph",
    "scglue.data": "glue_b200.data",
}
for _alias, _target in _ALIASES.items():
    sys.modules[_alias] = importlib.import_module(_target)

models = sys.modules["scglue.models"]
num = sys.modules["scglue.num"]
utils = sys.modules["scglue.utils"]
graph = sys.modules["scglue.graph"]
data = sys.modules["scglue.data"]
config = utils.config
log = utils.log
