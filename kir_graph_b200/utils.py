"""Logger shared with the reference (same name, so ``--log-level`` keeps working)."""
import logging

logger = logging.getLogger("graphkir")
