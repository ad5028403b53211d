"""Small simulator of HISAT2-style SAM records (CIGAR / MD / Zs) over a known variant table.
Used to generate inputs for the golden fixtures of the SAM -> variant walk (the expected outputs
come from the reference itself, tests/golden/make_golden.py)."""
from __future__ import annotations

import numpy as np

from kir_graph_b200.msa2hisat import Variant

BASES = "ACGT"


def make_table(rng: np.random.Generator, ref_name: str, length: int = 600, n_single: int = 40,
               n_del: int = 8) -> tuple[str, list[Variant]]:
    seq = "".join(rng.choice(list(BASES), size=length))
    variants = []
    taken = set()
    alleles = [f"{ref_name.split('*')[0]}*{i:03d}" for i in range(6)]
    for _ in range(n_single):
        pos = int(rng.integers(5, length - 5))
        alt = str(rng.choice([b for b in BASES if b != seq[pos]]))
        if (pos, alt) in taken:
            continue
        taken.add((pos, alt))
        variants.append(Variant(pos=pos, typ="single", ref=ref_name, val=alt))
    for _ in range(n_del):
        pos = int(rng.integers(5, length - 10))
        n = int(rng.integers(1, 4))
        if (pos, n) in taken:
            continue
        taken.add((pos, n))
        variants.append(Variant(pos=pos, typ="deletion", ref=ref_name, val=n))
    variants.sort()
    for i, v in enumerate(variants):
        v.id = f"hv{i}"
        v.allele = [a for a in alleles if rng.random() < 0.4]
        v.in_exon = bool(rng.random() < 0.3)
    return seq, variants


def simulate_record(rng: np.random.Generator, name: str, flag: int, ref_name: str, seq: str,
                    variants: list[Variant], start: int, n_ref: int = 70, novel: float = 0.01,
                    soft: float = 0.05, nh: int = 1) -> str:
    singles = {}
    dels = {}
    for v in variants:
        (singles if v.typ == "single" else dels).setdefault(v.pos, []).append(v)
    ops: list[tuple[str, int]] = []        # CIGAR
    md: list = []                          # ints (match runs), str ref bases, "^XYZ"
    zs: list[str] = []
    read = []
    run = 0                                # current MD match run
    zs_pos = 0                             # aligned read offset consumed by Zs

    def push(op, n):
        if ops and ops[-1][0] == op:
            ops[-1] = (op, ops[-1][1] + n)
        else:
            ops.append((op, n))

    p = start
    end = min(start + n_ref, len(seq) - 5)
    last_event = "M"
    while p < end:
        r = rng.random()
        if p in dels and rng.random() < 0.5 and p > start and last_event == "M":
            v = dels[p][int(rng.integers(len(dels[p])))]
            md.append(run); run = 0
            md.append("^" + seq[p:p + v.val])
            push("D", v.val)
            zs.append(f"{len(read) - zs_pos}|D|{v.id}")
            zs_pos = len(read)
            p += v.val
            last_event = "D"
            continue
        if r < novel / 2 and p > start + 2 and last_event == "M":          # novel deletion
            n = int(rng.integers(1, 3))
            md.append(run); run = 0
            md.append("^" + seq[p:p + n])
            push("D", n)
            p += n
            last_event = "D"
            continue
        if r < novel and p > start + 2 and last_event == "M":              # novel insertion
            n = int(rng.integers(1, 3))
            read.extend(rng.choice(list(BASES), size=n))
            push("I", n)
            last_event = "I"
            continue
        base = seq[p]
        if p in singles and rng.random() < 0.5:
            v = singles[p][int(rng.integers(len(singles[p])))]
            md.append(run); run = 0
            md.append(seq[p])
            zs.append(f"{len(read) - zs_pos}|S|{v.id}")
            zs_pos = len(read) + 1
            base = v.val
        elif rng.random() < novel:                                          # novel mismatch
            md.append(run); run = 0
            md.append(seq[p])
            base = str(rng.choice([b for b in BASES if b != seq[p]]))
        else:
            run += 1
        read.append(base)
        push("M", 1)
        p += 1
        last_event = "M"
    md.append(run)
    head = int(rng.integers(1, 6)) if rng.random() < soft else 0
    tail = int(rng.integers(1, 6)) if rng.random() < soft else 0
    cigar = (f"{head}S" if head else "") + "".join(f"{n}{op}" for op, n in ops) + (f"{tail}S" if tail else "")
    read_seq = "".join(rng.choice(list(BASES), size=head)) + "".join(read) + "".join(rng.choice(list(BASES), size=tail))
    md_str = "".join(str(x) for x in md)
    nm = sum(1 for x in md if isinstance(x, str) and not x.startswith("^")) + \
        sum(n for op, n in ops if op in "ID")
    fields = [name, str(flag), ref_name, str(start + 1), "60", cigar, "=", "300", "350", read_seq,
              "I" * len(read_seq), f"NM:i:{nm}", f"MD:Z:{md_str}"]
    if zs:
        fields.append("Zs:Z:" + ",".join(zs))
    fields.append(f"NH:i:{nh}")
    return "\t".join(fields)


def simulate_pairs(seed: int, n_pairs: int = 60) -> tuple[str, list[Variant], list[tuple[str, str]]]:
    rng = np.random.default_rng(seed)
    ref_name = "KIRSAM*BACKBONE"
    seq, variants = make_table(rng, ref_name)
    pairs = []
    for i in range(n_pairs):
        s1 = int(rng.integers(0, len(seq) - 200))
        s2 = s1 + int(rng.integers(40, 110))
        nh = 2 if rng.random() < 0.05 else 1
        left = simulate_record(rng, f"read{i}", 99, ref_name, seq, variants, s1, nh=nh)
        right = simulate_record(rng, f"read{i}", 147, ref_name, seq, variants, s2, nh=nh)
        pairs.append((left, right))
    return seq, variants, pairs


def sam_text(pairs, header=True):
    """Name-sorted SAM of (left, right) pairs with PNEXT set to the mate position; the right record
    comes first so that readPair yields (left, right) (it yields the later record first, :266)."""
    lines = ["@HD\tVN:1.0\tSO:queryname", "[bam_sort_core] merging from 0 files"] if header else []
    for left, right in pairs:
        lf, rf = left.split("\t"), right.split("\t")
        lf[7], rf[7] = rf[3], lf[3]
        lines += ["\t".join(rf), "\t".join(lf)]
    return "\n".join(lines) + "\n"


def multi_gene(seed, n_pairs=80, novel=0.03):
    """Two backbones in one sorted table, pairs of both interleaved, plus records that readPair skips."""
    rng = np.random.default_rng(seed)
    table, pairs = [], []
    for g, name in enumerate(("KIRA*BACKBONE", "KIRB*BACKBONE")):
        seq, variants = make_table(rng, name, n_single=50, n_del=10)
        for i, v in enumerate(variants):
            v.id = f"hv{1000 * g + i}"
        # an insertion in the table exercises the type order ins < single < del at one position
        table += variants + [Variant(pos=variants[3].pos, typ="insertion", ref=name, val="AC",
                                     id=f"hv{1000 * g + 900}", length=2)]
        for i in range(n_pairs):
            s1 = int(rng.integers(0, len(seq) - 200))
            s2 = s1 + int(rng.integers(40, 110))
            nh = 3 if rng.random() < 0.1 else 1
            pairs.append((simulate_record(rng, f"g{g}r{i}", 99, name, seq, variants, s1, nh=nh, novel=novel),
                          simulate_record(rng, f"g{g}r{i}", 147, name, seq, variants, s2, nh=nh, novel=novel)))
    table.sort()
    order = rng.permutation(len(pairs))
    return table, [pairs[i] for i in order]
