"""Host-side numbers of the typing path, for bench.py's ``host_prep`` and ``api_e2e`` blocks (SURVEY.md
section 8d: "JSON parsing excluded ... and reported separately").

``host_prep(scale)``   one synthetic cfg3 sample written as the reference's ``.variant.json`` (with SAM text
                       of realistic size): seconds per sample for .json -> packed genes on the object path
                       (json.load + PairRead / Variant objects + pack_gene, what the reference reader does)
                       and on the fast path (C++ scanner + array packing), the sidecar load, the wire
                       encoding, and SAM text -> packed genes through the native extraction loop.
``api_e2e(scale)``     the drop-in call ``selectKirTypingModel("full", json, top_n=600).typing(cn)``
                       (graphkir/kir_typing.py:207-228, main.py:192-199) on that file, object path and
                       ``_fast``, beside the unmodified reference (oracle/_ref, in a process of its own that
                       never loads the CUDA library) on the same file.

Both are bounded samples (``scale`` of the 200k read pairs); the times are linear in the reads."""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SAM_LINE = ("r%d\t99\tKIR*BACKBONE\t1234\t60\t150M\t=\t1500\t416\t" + "ACGT" * 37 + "AC\t" + "F" * 150 +
            "\tNM:i:1\tMD:Z:75A74\tZs:Z:75|S|hv12\tNH:i:1")


def write_sample(folder: str, scale: float, seed: int = 3) -> tuple[str, dict, int]:
    """A cfg3 sample as ``{folder}/s.variant.json`` (+ its CN table); returns (prefix, gene -> cn, pairs)."""
    from kir_graph_b200 import synthetic
    from kir_graph_b200.hisat2 import writeReadsAndVariantsData
    genes = synthetic.make_wgs30x_sample(seed=seed, scale=scale)
    reads, variants = [], []
    for g in genes:
        r, v = g.to_objects()
        reads += r
        variants += v
    for i, r in enumerate(reads):                   # realistic record size: ~0.7 KB of SAM text per pair
        r.l_sam = SAM_LINE % i
        r.r_sam = SAM_LINE % i
    prefix = os.path.join(folder, "s.variant")
    writeReadsAndVariantsData({"variants": variants, "reads": reads}, prefix + ".json")
    return prefix, {g.gene: g.cn for g in genes}, len(reads)


def host_prep(scale: float = 0.25, sam_pairs: int = 8000) -> dict:
    import numpy as np
    from kir_graph_b200 import fastjson, fastsam, packing, packio
    from kir_graph_b200.hisat2 import loadReadsAndVariantsData, removeMultipleMapped
    from kir_graph_b200.kir_typing import groupReads, groupVariants
    from kir_graph_b200.msa2hisat import Variant
    out: dict = {}
    with tempfile.TemporaryDirectory() as d:
        prefix, _, n = write_sample(d, scale)
        path = prefix + ".json"
        t0 = time.perf_counter()
        data = removeMultipleMapped(loadReadsAndVariantsData(path))
        rg, vg = groupReads(data["reads"]), groupVariants(data["variants"])
        slow = {g: packing.pack_gene(rg.get(g, []), v, mutate_reads=False, gene=g) for g, v in vg.items()}
        t1 = time.perf_counter()
        fast = fastjson.load_packs(path)
        t2 = time.perf_counter()
        for p in fast.values():
            packing.wire_encode(p)
        t3 = time.perf_counter()
        packio.save_packs(packio.sidecar_path(path), fast, packio.sidecar_meta(path))
        t4 = time.perf_counter()
        side = packio.load_sample_packs(path)
        t5 = time.perf_counter()
        assert all(np.array_equal(slow[g].ent_pos, side[g].ent_pos) for g in slow)
        us = lambda t: 1e6 * t / n
        out.update({
            "read_pairs": n, "json_mb": os.path.getsize(path) / 1e6,
            "sidecar_mb": os.path.getsize(packio.sidecar_path(path)) / 1e6,
            "json_to_packs_object_path": {"s": t1 - t0, "us_per_pair": us(t1 - t0)},
            "json_to_packs_fast_path": {"s": t2 - t1, "us_per_pair": us(t2 - t1)},
            "wire_encode": {"s": t3 - t2, "us_per_pair": us(t3 - t2)},
            "sidecar_load": {"s": t5 - t4, "us_per_pair": us(t5 - t4)},
        })
    # SAM text -> packs (native extraction loop + array packing) on simulated 2 x 150 bp records
    from tests import sam_sim
    rng = np.random.default_rng(1)
    table, pairs = [], []
    genes = [f"KIR{g}*BACKBONE" for g in range(8)]
    for g, name in enumerate(genes):
        seq, variants = sam_sim.make_table(rng, name, length=6000, n_single=120, n_del=12)
        for i, v in enumerate(variants):
            v.id = f"hv{10000 * g + i}"
            v.allele = [f"{name.split('*')[0]}*{k:03d}" for k in range(6) if (i + k) % 3]
        table += variants
        for i in range(sam_pairs // len(genes)):
            s1 = int(rng.integers(0, len(seq) - 400))
            s2 = s1 + int(rng.integers(100, 200))
            pairs.append((sam_sim.simulate_record(rng, f"g{g}r{i}", 99, name, seq, variants, s1, n_ref=150, novel=0.001, soft=0.02),
                          sam_sim.simulate_record(rng, f"g{g}r{i}", 147, name, seq, variants, s2, n_ref=150, novel=0.001, soft=0.02)))
    table.sort()
    raw = sam_sim.sam_text(pairs).encode()
    Variant.novel_id = 0
    t0 = time.perf_counter()
    ext = fastsam.extract(raw, table)
    packs = fastjson.packs_from_scan(ext.scan())
    t1 = time.perf_counter()
    out["sam_to_packs"] = {"s": t1 - t0, "us_per_pair": 1e6 * (t1 - t0) / len(pairs), "pairs": len(pairs),
                           "sam_mb": len(raw) / 1e6, "genes": len(packs)}
    out["note"] = ("one synthetic cfg3 sample (17 genes, 900 alleles) at a fraction of its 200k read pairs, one host "
                   "core; each time has a part that is fixed per sample (the variant table: ~7k Variant objects and "
                   "the membership bitsets) and a part linear in the read pairs")
    return out


_REF_CODE = """
import json, sys, time
sys.path.insert(0, %(root)r)
from oracle import ref_loader
kt = ref_loader.load()[2]
cn = json.load(open(%(cn)r))
t0 = time.perf_counter()
t = kt.selectKirTypingModel("full", %(json)r, top_n=600, variant_correction=True)
t1 = time.perf_counter()
alleles, warnings = t.typing(cn)
t2 = time.perf_counter()
maps = open('/proc/self/maps').read()
print(json.dumps({"load_s": t1 - t0, "typing_s": t2 - t1, "alleles": alleles,
                  "cuda_library_mapped": 'libgk_typing' in maps}))
"""


def api_e2e(scale: float = 0.1, backend=None) -> dict:
    """The unchanged API on one cfg3 sample file: ours (object path and ``_fast``) beside the reference."""
    from kir_graph_b200.kir_typing import selectKirTypingModel
    from oracle import ref_loader
    out: dict = {}
    with tempfile.TemporaryDirectory() as d:
        prefix, cn, n = write_sample(d, scale, seed=4)
        path = prefix + ".json"
        json.dump(cn, open(os.path.join(d, "cn.json"), "w"))
        res = {}
        for label, kw in (("object_path", {}), ("fast_path", {"_fast": True})):
            for attempt in range(2):                 # the second attempt is warm (CUDA context, allocator)
                t0 = time.perf_counter()
                t = selectKirTypingModel("full", path, top_n=600, variant_correction=True, _backend=backend, **kw)
                t1 = time.perf_counter()
                alleles, _ = t.typing(cn)
                t2 = time.perf_counter()
            res[label] = alleles
            out[label] = {"load_s": t1 - t0, "typing_s": t2 - t1, "total_s": t2 - t0}
        out["read_pairs"] = n
        out["calls_equal_between_paths"] = res["object_path"] == res["fast_path"]
        if ref_loader.available():
            code = _REF_CODE % {"root": ROOT, "cn": os.path.join(d, "cn.json"), "json": path}
            proc = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1200)
            if proc.returncode == 0:
                ref = json.loads(proc.stdout.strip().splitlines()[-1])
                out["reference"] = {"load_s": ref["load_s"], "typing_s": ref["typing_s"],
                                    "total_s": ref["load_s"] + ref["typing_s"],
                                    "cuda_library_mapped": ref["cuda_library_mapped"]}
                same = sum(sorted(a) == sorted(b) for a, b in zip(_per_gene(res["object_path"], cn), _per_gene(ref["alleles"], cn)))
                out["genes_called_as_the_reference"] = same
                out["genes"] = len(cn)
                out["speedup_total"] = out["reference"]["total_s"] / out["fast_path"]["total_s"]
                out["speedup_typing"] = out["reference"]["typing_s"] / out["fast_path"]["typing_s"]
            else:
                out["reference"] = {"error": proc.stderr[-500:]}
    out["call"] = 'selectKirTypingModel("full", "<sample>.variant.json", top_n=600, variant_correction=True).typing(cn)'
    out["note"] = (f"one synthetic cfg3 sample at {scale:g} of the reads ({out['read_pairs']} pairs); load = reading the "
                   ".json and building the typing model, typing = .typing(cn) over the 17 genes; second (warm) call")
    return out


def cn_model(n_samples: int = 96, backend=None) -> dict:
    """``depthToCN`` (graphkir/kir_cn.py:41-123: CNgroup.fit over 500 candidate bases x 500 depth bins x 7 copy
    numbers, then assignCN per gene) on the gene depths of a synthetic cohort: the mirror on the GPU beside the
    unmodified reference on one host core, same CN calls and fitted base required."""
    import numpy as np
    from kir_graph_b200 import kir_cn
    from oracle import ref_loader
    rng = np.random.default_rng(11)
    genes = [f"KIR{g}*BACKBONE" for g in ("2DL1", "2DL2", "2DL3", "2DL4", "2DL5", "2DP1", "2DS1", "2DS2", "2DS3", "2DS4",
                                           "2DS5", "3DL1", "3DL2", "3DL3", "3DP1", "3DS1")]
    depths = []
    for _ in range(n_samples):
        per_copy = 15.0 * rng.uniform(0.9, 1.1)
        depths.append({g: float(max(0.0, (2 if "3DL3" in g else int(rng.choice([0, 1, 1, 2, 2, 2, 3]))) * per_copy
                                    * (1 + 0.08 * rng.standard_normal()))) for g in genes})
    out: dict = {"samples": n_samples, "genes": len(genes)}
    for attempt in range(2):                         # second call warm
        t0 = time.perf_counter()
        cns, dist = kir_cn.depthToCN(depths, assume_3DL3_diploid=True, _backend=backend)
        t1 = time.perf_counter()
    out["b200_s"] = t1 - t0
    out["base"] = float(dist.base)
    if ref_loader.available():
        _, kc = ref_loader.load_cn()
        t0 = time.perf_counter()
        ref_cns, ref_dist = kc.depthToCN(depths, assume_3DL3_diploid=True)
        t1 = time.perf_counter()
        out["reference_s"] = t1 - t0
        out["same_cn_calls"] = [{k: int(v) for k, v in c.items()} for c in cns] == [{k: int(v) for k, v in c.items()} for c in ref_cns]
        out["same_base"] = float(ref_dist.base) == float(dist.base)
        out["likelihood_max_rel_diff"] = float(np.max(np.abs(dist.likelihood[:, 1] - ref_dist.likelihood[:, 1])
                                                      / np.abs(ref_dist.likelihood[:, 1])))
        out["speedup"] = out["reference_s"] / out["b200_s"]
    out["call"] = "depthToCN(sample_gene_depths, assume_3DL3_diploid=True)  (kir_cn.py:41-123)"
    return out


def _per_gene(alleles: list[str], cn: dict) -> list[list[str]]:
    out, at = [], 0
    for gene, c in cn.items():
        if c:
            out.append(alleles[at: at + c])
            at += c
    return out


if __name__ == "__main__":
    print(json.dumps(host_prep(float(sys.argv[1]) if len(sys.argv) > 1 else 0.05), indent=1))
