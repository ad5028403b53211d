"""CN model: the oracle's restatement of CNgroup / depthToCN against goldens
written by the unmodified reference (tests/golden/make_golden_cn.py), the mirror classes on the NumPy test
double of gk_cn_fit, and (`-m gpu`) on the CUDA kernel.  Tolerances: the likelihood curve and the CN-group
probabilities are float64 sums / exponentials (1e-10 / 1e-12 relative); the fitted base - a grid point - and
every CN call must be identical."""
import json
import os

import numpy as np
import pytest

from kir_graph_b200 import cn_model, kir_cn
from oracle import cn_oracle
from tests.fake_backend import FakeBackend
from tests.helpers import load_golden

CASES = load_golden("cn_model")["cases"]
NAMES = [c["name"] for c in CASES]


def _case(name):
    return CASES[NAMES.index(name)]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_equals_the_reference(name):
    c = _case(name)
    kw = c["kwargs"]
    with np.errstate(all="ignore"):
        cns, p, curve = cn_oracle.depth_to_cn(c["depths"], diploid=c["diploid"], kwargs=kw.get("cluster_method_kwargs"),
                                              assume_3dl3_diploid=kw.get("assume_3DL3_diploid", False))
        prob = cn_oracle.group_prob(p, p.base)
    assert cns == c["cns"] and p.base == c["base"] and p.bin_num == c["bin_num"] and p.x_max == c["x_max"]
    assert np.array_equal(curve, np.array(c["likelihood"]), equal_nan=True)          # bit-identical
    np.testing.assert_allclose(prob, np.array(c["group_prob"]), rtol=1e-13, atol=0, equal_nan=True)


def _run_mirror(c, backend, tmp_path):
    kw = dict(c["kwargs"])
    path = ""
    if c["diploid"] is not None:
        path = str(tmp_path / "dp")
        json.dump({"mean": c["diploid"][0], "std": c["diploid"][1]}, open(path + ".json", "w"))
    cns, dist = kir_cn.depthToCN(c["depths"], diploid_depth=path, _backend=backend, **kw)
    assert [{k: int(v) for k, v in x.items()} for x in cns] == c["cns"]
    assert dist.base == c["base"] and dist.bin_num == c["bin_num"] and dist.x_max == c["x_max"]
    assert dist.base_dev == c["base_dev"]
    ref = np.array(c["likelihood"])
    assert np.array_equal(dist.likelihood[:, 0], ref[:, 0])                          # the grid of candidate bases
    np.testing.assert_allclose(dist.likelihood[:, 1], ref[:, 1], rtol=1e-10, atol=0, equal_nan=True)
    np.testing.assert_allclose(dist.calcCNGroupProb(dist.base), np.array(c["group_prob"]), rtol=1e-12, atol=0,
                               equal_nan=True)
    return dist


@pytest.mark.parametrize("name", NAMES)
def test_mirror_on_the_numpy_statement(name, tmp_path):
    _run_mirror(_case(name), FakeBackend(), tmp_path)


def test_parameters_round_trip_and_errors(tmp_path):
    dist = _run_mirror(_case("cohort8_diploid_bounds"), FakeBackend(), tmp_path)
    file = str(tmp_path / "model.json")
    dist.save(file)
    again = cn_model.CNgroup.load(file)
    again._backend = FakeBackend()
    assert again.getParams().keys() == dist.getParams().keys() and again.base == dist.base
    depths = list(_case("cohort8_diploid_bounds")["depths"][0].values())
    assert again.assignCN(depths) == dist.assignCN(depths)
    with pytest.raises(NotImplementedError):
        kir_cn.depthToCN([{"KIR3DL3*BACKBONE": 10.0}], cluster_method="kde", _backend=FakeBackend())
    with pytest.raises(NotImplementedError):
        kir_cn.depthToCN([{"KIR3DL3*BACKBONE": 10.0}], cluster_method="other", _backend=FakeBackend())
    odd = cn_model.CNgroup(_backend=FakeBackend())
    odd.start_base = 3
    with pytest.raises(NotImplementedError):                                         # as the reference (:200-201)
        odd.fit([1.0, 2.0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_mirror_on_cuda(name, tmp_path):
    from kir_graph_b200 import engine
    _run_mirror(_case(name), engine.default_backend(), tmp_path)
