"""
Read records of the typing path: ``PairRead`` and the ``.variant.json`` contract.

Field names and JSON layout follow the reference (graphkir/hisat2.py:24-52 for
``PairRead``, :847-866 for the JSON reader/writer, :943-948 for the
multi-mapping filter) so files are interchangeable.  Running HISAT2/samtools is
out of scope and stays in the reference.
"""
from __future__ import annotations

import json
from dataclasses import asdict, dataclass, field
from typing import TypedDict

from .msa2hisat import Variant


@dataclass
class PairRead:
    """A read pair with the variant ids it supports (p) or contradicts (n)."""

    l_sam: str = ""
    r_sam: str = ""
    multiple: int = 1
    backbone: str = ""
    lpv: list[str] = field(default_factory=list)
    lnv: list[str] = field(default_factory=list)
    rpv: list[str] = field(default_factory=list)
    rnv: list[str] = field(default_factory=list)


class ReadsAndVariantsData(TypedDict):
    variants: list[Variant]
    reads: list[PairRead]


def writeReadsAndVariantsData(reads_data: ReadsAndVariantsData, filename: str) -> None:
    """Serialise to the reference's ``{prefix}.json`` layout."""
    payload = {
        "variants": [asdict(v) for v in reads_data["variants"]],
        "reads": [asdict(r) for r in reads_data["reads"]],
    }
    with open(filename, "w") as handle:
        json.dump(payload, handle)


def loadReadsAndVariantsData(filename: str) -> ReadsAndVariantsData:
    """Inverse of :func:`writeReadsAndVariantsData`."""
    with open(filename) as handle:
        raw = json.load(handle)
    return {
        "variants": [Variant(**item) for item in raw["variants"]],
        "reads": [PairRead(**item) for item in raw["reads"]],
    }


def removeMultipleMapped(reads_data: ReadsAndVariantsData) -> ReadsAndVariantsData:
    """Keep pairs whose NH tag is exactly one."""
    return {
        "variants": reads_data["variants"],
        "reads": [r for r in reads_data["reads"] if r.multiple == 1],
    }
