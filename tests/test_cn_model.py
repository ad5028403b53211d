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


# --- the callers: samtools depth tables -> .cn.tsv files (kir_cn.py:126-231, cn_model.py:382-390) -----------
PREDICT = load_golden("cn_predict")
PREDICT_NAMES = [c["name"] for c in PREDICT["cases"]]


def _slim(params):
    if isinstance(params, list):
        return [_slim(p) for p in params]
    return {k: (len(v) if k == "likelihood" else v) for k, v in params.items()}


def _run_predict(c, backend, d):
    d = str(d)
    depth_files, cn_files = [], []
    for i, text in enumerate(c["tables"]):
        depth_files.append(os.path.join(d, f"s{i}.depth.tsv"))
        cn_files.append(os.path.join(d, f"s{i}.cn.tsv"))
        open(depth_files[-1], "w").write(text)
    path = ""
    if c["diploid"] is not None:
        path = os.path.join(d, "dp")
        json.dump({"mean": c["diploid"][0], "std": c["diploid"][1]}, open(path + ".json", "w"))
    model_path = os.path.join(d, "model.json")
    kir_cn.predictSamplesCN(depth_files, cn_files, diploid_depth=path, save_cn_model_path=model_path,
                            _backend=backend, **c["kwargs"])
    assert [open(f).read() for f in cn_files] == c["cn_tsv"]                # byte for byte what the reference wrote
    models = {f[len("model.json"):]: _slim(json.loads(open(os.path.join(d, f)).read().replace(d, "@DIR@")))
              for f in sorted(os.listdir(d)) if f.startswith("model.json")}
    assert list(models) == list(c["models"])
    for suffix, want in c["models"].items():
        got = models[suffix]
        pairs = zip(got, want) if isinstance(want, list) else [(got, want)]
        for g, w in pairs:
            assert list(g) == list(w)                                       # same keys in the same order
            assert g == w                                                   # base, x_max, data, raw_df, gene ...
    if c["loaded"] is not None:
        dist = cn_model.loadCNModel(model_path)
        dist._backend = backend
        data = json.loads(open(model_path).read())["data"][:8]
        assert dist.base == c["loaded"]["base"]
        assert [int(x) for x in dist.assignCN([float(v) for v in data])] == c["loaded"]["cn_of_first"]
    merged = os.path.join(d, "cohort.cn.tsv")
    main_mod.mergeCN(cn_files, merged)
    assert open(merged).read().replace(d, "@DIR@") == c["merged_cn"]
    assert main_mod.loadCN(cn_files[0]) == {k: int(v) for k, v in (line.split("\t")[:2] for line in
                                                                   c["cn_tsv"][0].splitlines()[1:])}


from kir_graph_b200 import main as main_mod  # noqa: E402


@pytest.mark.parametrize("name", PREDICT_NAMES)
def test_predict_samples_cn_writes_the_reference_files(name, tmp_path):
    _run_predict(PREDICT["cases"][PREDICT_NAMES.index(name)], FakeBackend(), tmp_path)


def test_filter_depth_and_aggregation_modes(tmp_path):
    f = PREDICT["filter"]
    src, dst = str(tmp_path / "a.tsv"), str(tmp_path / "b.tsv")
    open(src, "w").write(f["table"])
    kir_cn.filterDepth(src, dst, {k: [tuple(r) for r in v] for k, v in f["regions"].items()})
    assert open(dst).read() == f["filtered"]
    table = kir_cn.readSamtoolsDepth(src)
    assert list(table.columns) == ["gene", "pos", "depth"]
    with pytest.raises(NotImplementedError):
        kir_cn.aggrDepths(table, "max")
    with pytest.raises(ValueError):                                         # no region at all (pd.concat of nothing)
        kir_cn.selectSamtoolsDepth(table, {})
    kde = str(tmp_path / "kde.json")
    json.dump({"method": "KDEcut"}, open(kde, "w"))
    with pytest.raises(NotImplementedError):
        cn_model.loadCNModel(kde)


def test_per_gene_prediction_with_a_dash_in_the_file_names(tmp_path):
    """The reference splits its ``{gene}-{file}`` keys at every ``-`` and fails on such paths; the mirror splits
    at the first one only and gives what it gives for the same tables under other names."""
    c = PREDICT["cases"][PREDICT_NAMES.index("cohort6_per_gene")]
    d = tmp_path / "run-2"
    d.mkdir()
    depth_files, cn_files = [], []
    for i, text in enumerate(c["tables"]):
        depth_files.append(str(d / f"s-{i}.depth.tsv"))
        cn_files.append(str(d / f"s-{i}.cn.tsv"))
        open(depth_files[-1], "w").write(text)
    kir_cn.predictSamplesCN(depth_files, cn_files, per_gene=True, _backend=FakeBackend())
    assert [open(f).read() for f in cn_files] == c["cn_tsv"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", PREDICT_NAMES)
def test_predict_samples_cn_on_cuda(name, tmp_path):
    from kir_graph_b200 import engine
    _run_predict(PREDICT["cases"][PREDICT_NAMES.index(name)], engine.default_backend(), tmp_path)


def test_allele_name_helpers_against_the_reference():
    from kir_graph_b200 import utils
    for name, resolution, field, limited, gene in PREDICT["allele_fields"]:
        assert utils.getAlleleField(name, resolution) == field
        assert utils.limitAlleleField(name, resolution) == limited
        assert utils.getGeneName(name) == gene
