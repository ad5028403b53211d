import os

import numpy as np

from kir_graph_b200 import cohort, packio
from kir_graph_b200.hisat2 import writeReadsAndVariantsData
from tests.fake_backend import FakeBackend
from tests.helpers import load_golden, objects_from_input


def test_sidecar_roundtrip_and_cohort_calls(tmp_path):
    sample = load_golden("sample_small")
    reads, variants = objects_from_input(sample["input"])
    path = os.path.join(tmp_path, "s.variant.json")
    writeReadsAndVariantsData({"reads": reads, "variants": variants}, path)
    packs = packio.pack_variant_json(path, variant_correction=True)
    side = os.path.join(tmp_path, "s.gkpack.npz")
    packio.save_packs(side, packs, {"variant_correction": True})
    back, meta = packio.load_packs(side)
    assert meta == {"variant_correction": True} and list(back) == list(packs)
    for g in packs:
        for name in ("mem_words", "ent_off", "ent_word", "ent_pos", "ent_neg", "k_obs", "obs_pos", "obs_neg"):
            assert np.array_equal(getattr(packs[g], name), getattr(back[g], name))
        assert packs[g].allele_names == back[g].allele_names
    genes = [g for g, cn in sample["gene_cn"].items() if cn and g in back]
    typer = cohort.BatchTyper([back[g] for g in genes], [sample["gene_cn"][g] for g in genes], top_n=60,
                              backend=FakeBackend())
    calls = typer.run()
    got = [a for c in calls for a in c.alleles]
    assert got == sample["calls"]["full"]["alleles"] or any(c.tie_flags for c in calls)
