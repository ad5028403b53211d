"""Logger shared with the reference (same name, so ``--log-level`` keeps working) and the allele-name helpers
of ``graphkir/utils.py:130-158``."""
import logging
import re

logger = logging.getLogger("graphkir")

_FIELD = re.compile(r"^\w+\*(\d+\w*)")


def getGeneName(allele: str) -> str:
    """``KIR3DP1*BACKBONE`` -> ``KIR3DP1`` (utils.py:130-132)."""
    return allele.split("*")[0]


def getAlleleField(allele: str, resolution: int = 7) -> str:
    """Digits of an allele name cut to ``resolution`` characters; 7 means everything, suffix letters included
    (``KIR2DL1*0320102N`` -> ``0320102N``); a name whose field does not start with a digit gives ``new``, a
    name without ``*`` the empty string (utils.py:140-158)."""
    if "*" not in allele:
        return ""
    found = _FIELD.findall(allele)
    field = str(found[0]) if found else "new"
    return field if resolution == 7 else field[:resolution]


def limitAlleleField(allele: str, resolution: int = 7) -> str:
    """``KIR3DP1*0010101`` at resolution 5 -> ``KIR3DP1*00101`` (utils.py:135-137)."""
    return getGeneName(allele) + "*" + getAlleleField(allele, resolution)
