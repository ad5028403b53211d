"""The C++ work-item planners (csrc/gk_plan.cu) against the NumPy statements in engine.py they restate: the
tables must be equal row for row, for random allele counts, kept-set counts and read counts, and for every
kept-set count around the tile boundaries.  CPU only (host routines of the library)."""
import numpy as np
import pytest

from kir_graph_b200 import engine, packing, synthetic
from tests.fake_backend import FakeBackend


@pytest.fixture(scope="module")
def group():
    packs = [packing.pack_synthetic(synthetic.make_gene([31, i], f"KIRPL{i}*BACKBONE", 6 + i, 64, 2, 40))
             for i in range(8)]
    batch = engine.MatrixBatch(packs, backend=FakeBackend(), run=False)
    return engine.SearchGroup(batch, list(range(8)), 300)


def _both(monkeypatch, fn):
    monkeypatch.setattr(engine, "PLAN_NATIVE", True)
    native = fn()
    monkeypatch.setattr(engine, "PLAN_NATIVE", False)
    return native, fn()


def _randomise(group, rng, kept=None):
    n = group.n_search
    group.A = rng.integers(1, 700, n).astype(group.A.dtype)
    group.n_cand = group.A.copy()
    group.R = rng.integers(1, 70000, n).astype(group.R.dtype)
    group.r_pad = np.maximum(128, -(-group.R // 128) * 128).astype(group.r_pad.dtype)
    group.kept = (rng.integers(0, 420, n) if kept is None else np.full(n, kept)).astype(np.int32)
    group.restricted = {}


@pytest.mark.parametrize("seed", range(12))
def test_score_items_equal_the_numpy_statement(group, monkeypatch, seed):
    rng = np.random.default_rng(seed)
    _randomise(group, rng)
    active = np.flatnonzero(rng.random(group.n_search) < 0.8)
    native, ref = _both(monkeypatch, lambda: group._score_items(active))
    assert native.dtype == ref.dtype and len(native) == len(ref) > 0
    assert np.array_equal(native, ref)


def test_score_items_for_every_kept_count(group, monkeypatch):
    rng = np.random.default_rng(99)
    for kept in list(range(1, 140)) + list(range(248, 262)) + [300, 383, 384, 385, 511, 512]:
        _randomise(group, rng, kept=kept)
        group.A[:4] = (35, 128, 130, 257)
        group.n_cand = group.A.copy()
        native, ref = _both(monkeypatch, lambda: group._score_items(np.arange(group.n_search)))
        assert np.array_equal(native, ref), kept


@pytest.mark.parametrize("seed", range(6))
def test_count_and_p_items_equal_the_numpy_statements(group, monkeypatch, seed):
    rng = np.random.default_rng(100 + seed)
    _randomise(group, rng)
    active = np.flatnonzero(rng.random(group.n_search) < 0.8)
    n_alive = rng.integers(0, 600, group.n_search).astype(np.int64)
    native, ref = _both(monkeypatch, lambda: group._count_items(active, n_alive))
    assert native.dtype == ref.dtype and np.array_equal(native, ref)
    native, ref = _both(monkeypatch, lambda: group._p_items(active))
    assert native.dtype == ref.dtype and np.array_equal(native, ref)


def test_planners_report_a_short_buffer():
    from kir_graph_b200 import _cabi
    lib = _cabi.load()
    ids = np.arange(2, dtype=np.int32)
    count = np.array([3, 2], dtype=np.int64)
    extent = np.array([5000, 100], dtype=np.int64)
    out = np.zeros((4, 4), dtype=np.int32)
    need = lib.gk_plan_grid_items(2, ids.ctypes.data, count.ctypes.data, 8, extent.ctypes.data, 2048, out.ctypes.data, 4)
    assert need == -(3 * 3 + 2 * 1)
