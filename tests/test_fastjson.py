"""C++ scanner of .variant.json + array-level packing against the object path
(json.load -> PairRead objects -> pack_gene), which follows the reference reader
(graphkir/hisat2.py:859-866, :943-948; kir_typing.py:15-28, :92-97)."""
import json

import numpy as np
import pytest

from kir_graph_b200 import fastjson, packing, synthetic
from kir_graph_b200.hisat2 import (PairRead, loadReadsAndVariantsData, removeMultipleMapped,
                                   writeReadsAndVariantsData)
from kir_graph_b200.kir_typing import groupReads, groupVariants
from tests.helpers import load_golden


def _sample(tmp_path, seed=31, total_reads=2500, n_genes=5):
    genes = synthetic.make_wgs30x_sample(seed=seed, total_reads=total_reads)[:n_genes]
    reads, variants = [], []
    for g in genes:
        r, v = g.to_objects()
        reads += r
        variants += v
    rng = np.random.default_rng(seed)
    order = rng.permutation(len(reads))                  # genes interleaved, as in a real name-sorted file
    reads = [reads[i] for i in order]
    for i in rng.choice(len(reads), size=40, replace=False):
        reads[i].multiple = int(rng.integers(2, 5))      # multi-mapped pairs are dropped (hisat2.py:943-948)
    reads[7].l_sam = 'name\\tq"uo\\\\te\tFLAGé中\U0001F600 {"lpv": ["hv0"]}'   # nasty SAM text is skipped
    reads[11].lpv, reads[11].rpv, reads[11].lnv, reads[11].rnv = [], [], [], []
    reads.append(PairRead(l_sam="x", r_sam="y", backbone="KIRELSE*BACKBONE", lpv=[], lnv=[], rpv=[], rnv=[]))
    path = str(tmp_path / "sample.variant.json")
    writeReadsAndVariantsData({"variants": variants, "reads": reads}, path)
    return path, genes


def _assert_same_pack(a: packing.GenePack, b: packing.GenePack):
    assert a.gene == b.gene and a.allele_names == b.allele_names and a.variant_ids == b.variant_ids
    for name in ("mem_words", "ent_off", "ent_word", "ent_pos", "ent_neg", "k_obs", "kept_reads", "var_pos",
                 "var_is_del", "obs_pos", "obs_neg"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert a.var_val == b.var_val
    for name in packing.LIST_NAMES:
        assert np.array_equal(a.csr.offsets[name], b.csr.offsets[name]), name
        assert np.array_equal(a.csr.indices[name], b.csr.indices[name]), name


def test_scan_equals_json_load(tmp_path):
    path, _ = _sample(tmp_path)
    sc = fastjson.scan(path)
    ref = json.load(open(path))
    assert sc.n_reads == len(ref["reads"])
    assert [sc.genes[i] for i in sc.backbone] == [r["backbone"] for r in ref["reads"]]
    assert sc.multiple.tolist() == [r["multiple"] for r in ref["reads"]]
    for name in fastjson.SCAN_LISTS:
        off, idx = sc.offsets[name], sc.indices[name]
        got = [[sc.ids[j] for j in idx[off[i]:off[i + 1]]] for i in range(sc.n_reads)]
        assert got == [r[name] for r in ref["reads"]], name
    assert [v.id for v in sc.variants] == [v["id"] for v in ref["variants"]]


@pytest.mark.parametrize("variant_correction,no_empty", [(True, True), (False, True), (True, False)])
def test_packs_equal_object_path(tmp_path, variant_correction, no_empty):
    path, genes = _sample(tmp_path)
    data = removeMultipleMapped(loadReadsAndVariantsData(path))
    reads_by_gene, variants_by_gene = groupReads(data["reads"]), groupVariants(data["variants"])
    fast = fastjson.load_packs(path, variant_correction=variant_correction, no_empty=no_empty)
    assert list(fast) == list(variants_by_gene)
    for gene, variants in variants_by_gene.items():
        want = packing.pack_gene(reads_by_gene.get(gene, []), variants, variant_correction=variant_correction,
                                 no_empty=no_empty, mutate_reads=False, gene=gene)
        _assert_same_pack(fast[gene], want)
    only = fastjson.load_packs(path, genes=[genes[1].gene])
    assert list(only) == [genes[1].gene]


def test_reference_sample_fixture(tmp_path):
    """The input of the golden sample (written by the reference's own writer semantics)."""
    case = load_golden("sample_small")
    path = str(tmp_path / "golden.json")
    json.dump(case["input"], open(path, "w"))
    data = removeMultipleMapped(loadReadsAndVariantsData(path))
    reads_by_gene, variants_by_gene = groupReads(data["reads"]), groupVariants(data["variants"])
    fast = fastjson.load_packs(path)
    for gene, variants in variants_by_gene.items():
        _assert_same_pack(fast[gene], packing.pack_gene(reads_by_gene.get(gene, []), variants, mutate_reads=False,
                                                        gene=gene))


def test_layouts_and_errors(tmp_path):
    doc = {"reads": [{"lpv": ["a"], "backbone": "G*BACKBONE"}, {}], "extra": {"x": [1, 2.5e3, None, True, {"y": "z"}]},
           "variants": []}
    sc = fastjson.scan_bytes(json.dumps(doc, indent=2).encode())       # keys in any order, defaults, pretty-printed
    assert sc.n_reads == 2 and sc.multiple.tolist() == [1, 1] and [sc.genes[i] for i in sc.backbone] == ["G*BACKBONE", ""]
    assert sc.offsets["lpv"].tolist() == [0, 1, 1] and sc.ids == ["a"] and sc.variants == []
    assert fastjson.scan_bytes(b'{"reads": []}').n_reads == 0
    assert fastjson.scan_bytes(b"{}").n_reads == 0
    for bad in (b'{"reads": [{"lpv": [1]}]}', b'{"reads": [{"lpv": ["a"}]}', b'{"reads": [', b'[1, 2]', b'{"reads": [{"multiple": "x"}]}'):
        with pytest.raises(ValueError):
            fastjson.scan_bytes(bad)
    path, genes = _sample(tmp_path)
    data = json.load(open(path))
    data["reads"][0]["lpv"].append("hv_unknown")
    data["reads"][0]["multiple"] = 1
    json.dump(data, open(path, "w"))
    with pytest.raises(KeyError):
        fastjson.load_packs(path)


def test_entry_packing_equals_array_statement():
    """gk_pack_entries (host C++) against the NumPy statement, with duplicated observations."""
    rng = np.random.default_rng(3)
    for trial in range(6):
        n, n_var = int(rng.integers(1, 60)), int(rng.integers(1, 200))
        offsets, indices = {}, {}
        for name in packing.LIST_NAMES:
            lens = rng.integers(0, 9, size=n)
            if trial == 0:
                lens[:] = 0
            off = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(lens, out=off[1:])
            idx = rng.integers(0, n_var, size=int(off[-1])).astype(np.int32)
            if len(idx) > 4:
                idx[1:4] = idx[0]                                   # repeated observation
            offsets[name], indices[name] = off, idx
        csr = packing.ReadCSR(n, offsets, indices)
        for a, b in zip(packing.pack_entries(csr), packing.pack_entries_numpy(csr)):
            assert a.dtype == b.dtype and np.array_equal(a, b)
    off = {name: np.array([0, 0], dtype=np.int64) for name in packing.LIST_NAMES}
    idx = {name: np.zeros(0, dtype=np.int32) for name in packing.LIST_NAMES}
    off["lpv"] = np.array([0, 300], dtype=np.int64)
    idx["lpv"] = np.full(300, 5, dtype=np.int32)
    with pytest.raises(ValueError, match="repeated more than 255 times"):
        packing.pack_entries(packing.ReadCSR(1, off, idx))


@pytest.mark.parametrize("method", ["full"])
def test_fast_driver_equals_object_driver(tmp_path, method):
    """selectKirTypingModel(..., _fast=True) calls the same alleles as the object path."""
    from kir_graph_b200.kir_typing import selectKirTypingModel
    from tests.fake_backend import FakeBackend
    path, genes = _sample(tmp_path, seed=37, total_reads=4000, n_genes=6)
    gene_cn = {g.gene: g.cn for g in genes}
    slow = selectKirTypingModel(method, path, top_n=60, variant_correction=True, _backend=FakeBackend())
    fast = selectKirTypingModel(method, path, top_n=60, variant_correction=True, _backend=FakeBackend(), _fast=True)
    a_slow, w_slow = slow.typing(gene_cn)
    a_fast, w_fast = fast.typing(gene_cn)
    assert a_fast == a_slow and w_fast == w_slow
    assert fast.getAllPossibleTyping() == slow.getAllPossibleTyping()


def test_scanner_fuzz_against_json_loads():
    """tools/fuzz_json_scan.py for a few seconds: shuffled / missing keys, unknown nested values, escapes,
    non-ASCII text, every json.dumps layout - the scanner must agree with json.loads on every case."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_json_scan.py"), "11", "4"],
                          capture_output=True, text=True, timeout=120)
    assert proc.returncode == 0 and " bad 0" in proc.stdout, proc.stdout[-1500:] + proc.stderr[-1500:]
