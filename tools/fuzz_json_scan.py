"""Randomised check of the C++ .variant.json scanner (kir_graph_b200.fastjson) against json.loads:
shuffled and missing keys, unknown nested values, escapes and non-ASCII text in the strings, every
json.dumps layout.  CPU only.

    python tools/fuzz_json_scan.py [seed=0] [seconds=60]
"""
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kir_graph_b200 import fastjson  # noqa: E402

rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ALPHABET = ["a", "b", '"', "\\", "\t", "\n", "/", "{", "}", "[", "]", ",", ":", " ", "é", "\U0001F600", "x", "0"]


def rstr():
    return "".join(rnd.choice(ALPHABET) for _ in range(rnd.randint(0, 12)))


def rid():
    return rnd.choice(["hv", "nv", ""]) + str(rnd.randint(0, 30)) + (rstr() if rnd.random() < 0.05 else "")


t0 = time.time()
n = bad = 0
while time.time() - t0 < float(sys.argv[2] if len(sys.argv) > 2 else 60):
    reads = []
    for _ in range(rnd.randint(0, 8)):
        r = {"l_sam": rstr(), "r_sam": rstr(), "multiple": rnd.randint(0, 5),
             "backbone": rnd.choice(["G1*BACKBONE", "G2", rstr()]),
             "lpv": [rid() for _ in range(rnd.randint(0, 4))], "lnv": [rid() for _ in range(rnd.randint(0, 4))],
             "rpv": [rid() for _ in range(rnd.randint(0, 4))], "rnv": [rid() for _ in range(rnd.randint(0, 4))]}
        if rnd.random() < 0.3:
            r["extra"] = rnd.choice([None, 1.5e3, {"a": [1, {"b": "}"}]}, [[], [{}]], True, "lpv"])
        if rnd.random() < 0.3:
            for k in rnd.sample(list(r), rnd.randint(0, 3)):
                if k in ("l_sam", "r_sam", "extra"):
                    r.pop(k, None)
        items = list(r.items())
        rnd.shuffle(items)
        reads.append(dict(items))
    variants = [{"pos": i, "typ": "single", "ref": "G1*BACKBONE", "val": "A", "id": f"hv{i}", "length": 1,
                 "allele": [rstr()], "freq": None, "ignore": False, "in_exon": False}
                for i in range(rnd.randint(0, 3))]
    top = [("variants", variants), ("reads", reads)]
    if rnd.random() < 0.3:
        top.append(("other", {"reads": [1, 2], "x": rstr()}))
    rnd.shuffle(top)
    kw = rnd.choice([{}, {"indent": 2}, {"separators": (",", ":")}, {"ensure_ascii": False}])
    text = json.dumps(dict(top), **kw)
    n += 1
    try:
        sc = fastjson.scan_bytes(text.encode("utf-8"))
    except Exception as exc:
        print("EXC", repr(exc), text[:300])
        bad += 1
        continue
    ref = json.loads(text)
    ok = sc.n_reads == len(ref["reads"])
    if ok:
        for i, r in enumerate(ref["reads"]):
            ok &= sc.genes[sc.backbone[i]] == r.get("backbone", "")
            ok &= int(sc.multiple[i]) == r.get("multiple", 1)
            for name in fastjson.SCAN_LISTS:
                got = [sc.ids[j] for j in sc.indices[name][sc.offsets[name][i]:sc.offsets[name][i + 1]]]
                ok &= got == r.get(name, [])
        ok &= [v.id for v in sc.variants] == [v["id"] for v in ref["variants"]]
    if not ok:
        print("DIFF", text[:400])
        bad += 1
        if bad > 5:
            break
print("cases", n, "bad", bad)
sys.exit(1 if bad else 0)
