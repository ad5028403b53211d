"""
TEST / BENCH INFRASTRUCTURE -- imports the byte-compiled reference from ``oracle/_ref`` (``make_ref.py``).

``load()`` returns the reference's own modules (``graphkir.typing_mulit_allele`` ...), unmodified; the
packages it imports for plots, FASTA lengths and MSA files are replaced by empty stubs (they take no part
in the typing arithmetic: ``plot()`` typing_mulit_allele.py:600-619, ``readAlleleLength`` typing_em.py:32-34,
the MSA writers of msa2hisat.py).  Only ``tests/``, ``bench.py --impl reference`` / ``cpu_baseline`` and
``__graft_entry__`` may import this module.
"""
from __future__ import annotations

import importlib.abc
import importlib.util
import marshal
import os
import sys
import types
from dataclasses import asdict

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


EXT = ".gkref"


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "graphkir", "typing_mulit_allele" + EXT))


class _RefFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Serves the package ``graphkir`` from the byte-compiled modules under oracle/_ref/graphkir."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "graphkir" and not fullname.startswith("graphkir."):
            return None
        name = "__init__" if fullname == "graphkir" else fullname.split(".", 1)[1]
        origin = os.path.join(REF_DIR, "graphkir", name + EXT)
        if not os.path.exists(origin):
            return None
        return importlib.util.spec_from_loader(fullname, self, origin=origin, is_package=fullname == "graphkir")

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        with open(module.__spec__.origin, "rb") as f:
            code = marshal.loads(f.read()[16:])           # skip the .pyc header (magic, flags, hash)
        module.__file__ = module.__spec__.origin
        exec(code, module.__dict__)


_finder = _RefFinder()


def _stubs() -> None:
    for name in ("plotly", "plotly.express", "plotly.graph_objects", "plotly.subplots", "Bio", "Bio.SeqIO", "pyhlamsa"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["plotly"].express = sys.modules["plotly.express"]
    sys.modules["plotly"].graph_objects = sys.modules["plotly.graph_objects"]
    if not hasattr(sys.modules["plotly.graph_objects"], "Figure"):
        sys.modules["plotly.graph_objects"].Figure = object
    if not hasattr(sys.modules["plotly.subplots"], "make_subplots"):
        sys.modules["plotly.subplots"].make_subplots = lambda *a, **k: None
    sys.modules["Bio"].SeqIO = sys.modules["Bio.SeqIO"]
    if not hasattr(sys.modules["pyhlamsa"], "Genemsa"):
        sys.modules["pyhlamsa"].Genemsa = object


def load():
    """(typing_mulit_allele, typing_em, kir_typing, hisat2, msa2hisat) of the reference."""
    if not available():
        raise RuntimeError("oracle/_ref is not built: run `python oracle/make_ref.py` where /root/reference exists")
    _stubs()
    if _finder not in sys.meta_path:
        sys.meta_path.insert(0, _finder)
    import graphkir.typing_mulit_allele as tma
    import graphkir.typing_em as tem
    import graphkir.kir_typing as kt
    import graphkir.hisat2 as h2
    import graphkir.msa2hisat as m2h
    if not os.path.abspath(tma.__file__).startswith(REF_DIR):
        raise RuntimeError(f"graphkir was imported from {tma.__file__}, not from oracle/_ref")
    return tma, tem, kt, h2, m2h


def load_cn():
    """(cn_model, kir_cn) of the reference (they import scipy, scikit-learn and pandas, which the image has)."""
    load()
    import graphkir.cn_model as cm
    import graphkir.kir_cn as kc
    return cm, kc


def to_ref_objects(reads, variants):
    """Reads / variants of this repository's dataclasses -> the reference's own (same fields)."""
    _, _, _, h2, m2h = load()
    return [h2.PairRead(**asdict(r)) for r in reads], [m2h.Variant(**asdict(v)) for v in variants]
