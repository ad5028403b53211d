#!/usr/bin/env python
"""
Golden fixture of the caller side, from the UNMODIFIED reference: graphkir.main.alleleTyping
(main.py:171-220) and utils.mergeAllele (:161-165) over a three-sample synthetic cohort, run in a
scratch directory with relative file names (the "name" column holds the path).  Stubs only for
modules that take no part in it (plotly, Bio, pyhlamsa, pysam).  Build container only:

    python tests/golden/make_golden_main.py
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from make_golden import dump, import_reference  # noqa: E402

from tests.cohort_sim import write_cohort  # noqa: E402


def main() -> None:
    import_reference()
    for extra in ("Bio.Seq", "Bio.SeqRecord", "pysam"):
        sys.modules.setdefault(extra, types.ModuleType(extra))
    sys.modules["pyhlamsa"].KIRmsa = object
    sys.modules["Bio.Seq"].Seq = object
    sys.modules["Bio.SeqRecord"].SeqRecord = object
    sys.modules["pysam"].AlignmentFile = object
    import re
    for _ in range(20):                      # names the unused modules import from the stubs
        try:
            import graphkir.main as gm
            break
        except ImportError as exc:
            found = re.search(r"cannot import name '(\w+)' from '([\w.]+)'", str(exc))
            if found:
                setattr(sys.modules[found.group(2)], found.group(1), object)
            elif exc.name:
                sys.modules[exc.name] = types.ModuleType(exc.name)
            else:
                raise
    from graphkir.utils import mergeAllele
    out = {"kind": "main", "methods": {}}
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        names, cn_files, inputs = write_cohort(".")
        out["inputs"] = inputs
        for method in ("full", "exonfirst"):
            files = gm.alleleTyping(names, cn_files, method)
            mergeAllele(files, f"cohort.{method}.allele.tsv")
            out["methods"][method] = {
                "files": files,
                "tsv": [open(f).read() for f in files],
                "possible": [open(f[:-4] + ".possible.tsv").read() for f in files],
                "merged": open(f"cohort.{method}.allele.tsv").read(),
            }
            print(method, [t.split("\n")[1][:120] for t in out["methods"][method]["tsv"]])
        os.chdir(HERE)
    dump("main_tsv", out)


if __name__ == "__main__":
    main()
