#!/usr/bin/env python
"""
TEST / BENCH INFRASTRUCTURE -- builds ``oracle/_ref``: the reference's own typing path, byte-compiled.

The reference (linnil1/KIR_graph) is pure Python, so "building" it means byte-compiling the twelve modules
of ``graphkir`` that the typing path and the CN model import, from the sources where they lie under ``/root/reference``,
into ``oracle/_ref/graphkir/*.gkref`` (byte code only: no reference source text is copied into this
repository; ``oracle/_ref/`` is git-ignored and travels to the GPU box with the snapshot like the built
``.so`` - the files do not carry the ``.pyc`` suffix because snapshot tools commonly drop ``*.pyc``).  ``oracle/ref_loader.py`` imports them there, with the three plotting / FASTA / MSA packages the
reference imports at module top (plotly, Bio, pyhlamsa: none takes part in the typing arithmetic) stubbed.

Used by: ``bench.py --impl reference`` and the ``cpu_baseline`` leg (the unmodified reference timed on the
box's host cores), and ``tests/`` (pinning the oracle).  Never imported by ``kir_graph_b200``.

    python oracle/make_ref.py [--src /root/reference] [--force]
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "graphkir")
# typing_mulit_allele / typing_em / kir_typing and what they import
EXT = ".gkref"            # a .pyc by content (16-byte header + marshalled code object)
MODULES = ("__init__", "typing_mulit_allele", "typing_em", "kir_typing", "hisat2", "msa2hisat", "utils",
           "external_tools", "pileup",
           "cn_model", "kir_cn", "samtools_utils")       # the CN model and its caller


def build(src: str = "/root/reference", force: bool = False) -> str | None:
    """Returns the output directory, or None when the reference tree is absent (GPU box: prebuilt files)."""
    pkg = os.path.join(src, "graphkir")
    if not os.path.isdir(pkg):
        return OUT if os.path.exists(os.path.join(OUT, "typing_mulit_allele" + EXT)) else None
    os.makedirs(OUT, exist_ok=True)
    for name in MODULES:
        source = os.path.join(pkg, name + ".py")
        target = os.path.join(OUT, name + EXT)
        if force or not os.path.exists(target) or os.path.getmtime(target) < os.path.getmtime(source):
            py_compile.compile(source, cfile=target, dfile=f"graphkir/{name}.py", doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(HERE, "_ref", "BUILT_FROM"), "w") as f:
        f.write(f"{pkg}\npython {sys.version.split()[0]}\nmodules {' '.join(MODULES)}\n")
    return OUT


if __name__ == "__main__":
    args = sys.argv[1:]
    src = args[args.index("--src") + 1] if "--src" in args else "/root/reference"
    out = build(src, force="--force" in args)
    print(out if out else "reference tree not found and no prebuilt oracle/_ref")
    sys.exit(0 if out else 1)
