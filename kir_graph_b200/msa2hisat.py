"""
Variant record: the data contract shared by read extraction and allele typing.

Mirrors the public fields, ordering, equality and hashing of the reference's
``Variant`` dataclass (reference: graphkir/msa2hisat.py:15-63) so that
``{prefix}.variant.json`` files written by either implementation load in both.
Only the record type lives here; building a HISAT2 index from an MSA is out of
scope (stays as in the reference).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import ClassVar

_TYPE_RANK = {"insertion": 0, "single": 1, "deletion": 2, "match": 3}


@dataclass
class Variant:
    """One graph variant (or a match segment produced while walking a read)."""

    pos: int
    typ: str
    ref: str
    val: None | int | str = None
    id: None | str = None
    length: int = 0
    allele: list[str] = field(default_factory=list)
    freq: None | float = None
    ignore: None | bool = False
    in_exon: bool = False

    # class-level state; ``novel_id`` numbers the nv* ids handed out during
    # read extraction (reference: graphkir/hisat2.py:599-600)
    min_freq_threshold: ClassVar[float] = 0.1
    count: ClassVar[int] = 0
    haplo_id: ClassVar[int] = 0
    novel_id: ClassVar[int] = 0
    order_type: ClassVar[dict[str, int]] = _TYPE_RANK
    order_nuc: ClassVar[dict[str, int]] = {"A": 0, "C": 1, "G": 2, "T": 3}

    def sort_key(self) -> tuple:
        """(ref, pos, ins<single<del<match, val): the index order of variants."""
        return (self.ref, self.pos, _TYPE_RANK[self.typ], self.val)

    def identity(self) -> tuple:
        """Fields that decide whether two records are the same variant."""
        return (self.pos, self.ref, self.typ, self.val)

    def __lt__(self, other: object) -> bool:
        if not isinstance(other, Variant):
            return NotImplemented
        return self.sort_key() < other.sort_key()

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, Variant):
            return NotImplemented
        return self.identity() == other.identity()

    def __hash__(self) -> int:
        return hash(self.identity())
