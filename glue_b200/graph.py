r"""
Guidance-graph validation (the part of ``scglue/graph.py`` that ``SCGLUEModel.fit`` calls).
"""

from itertools import chain
from typing import Iterable

import networkx as nx

from .utils import logged


@logged
def check_graph(graph: nx.Graph, adatas: Iterable, cov: str = "error", attr: str = "error",
                loop: str = "error", sym: str = "error", verbose: bool = True) -> None:
    r"""
    Check a guidance graph against the datasets it should connect
    (``scglue/graph.py:112-195``): every configured feature is a vertex (``cov``), every edge
    carries ``weight`` in (0, 1] and ``sign`` in {-1, 1} (``attr``), every vertex has a
    self-loop (``loop``), and the graph is symmetric (``sym``).  Each check's action is one
    of ``"ignore"``, ``"warn"``, ``"error"``.
    """
    for name, action in dict(cov=cov, attr=attr, loop=loop, sym=sym).items():
        if action not in ("ignore", "warn", "error"):
            raise ValueError(f"Invalid action for `{name}`: {action!r}")
    passed = True

    def report(action: str, ok: bool, msg: str) -> None:
        nonlocal passed
        if ok or action == "ignore":
            return
        passed = False
        if action == "error":
            raise ValueError(msg)
        check_graph.logger.warning(msg)

    if cov != "ignore":
        nodes = set(graph.nodes)
        missing = [v for v in chain.from_iterable(a.var_names for a in adatas) if v not in nodes]
        report(cov, not missing, "Some variables are not covered by the graph!")
    if attr != "ignore":
        ok = all("weight" in d and "sign" in d and 0 < d["weight"] <= 1 and d["sign"] in (-1, 1)
                 for _, _, d in graph.edges(data=True))
        report(attr, ok, "Missing or invalid edge attributes (weight in (0, 1], sign in {-1, 1})!")
    if loop != "ignore":
        report(loop, all(graph.has_edge(v, v) for v in graph.nodes),
               "Missing self-loop!")
    if sym != "ignore" and graph.is_directed():
        report(sym, all(graph.has_edge(t, s) for s, t in graph.edges), "Graph is not symmetric!")
    if verbose and passed:
        check_graph.logger.info("All checks passed!")
