"""
Mirror of ``cutadapt.align`` (src/cutadapt/align.py:1-34): the EndSkip flags and the names
``cutadapt.adapters`` imports from it.
"""
from enum import IntFlag

from ._align import (  # noqa: F401
    Aligner,
    PrefixComparer,
    SuffixComparer,
    hamming_sphere,
    hamming_environment,
    edit_environment,
)

__all__ = [
    "EndSkip",
    "Aligner",
    "PrefixComparer",
    "SuffixComparer",
    "hamming_sphere",
    "hamming_environment",
    "edit_environment",
]


class EndSkip(IntFlag):
    """
    Which ends of reference (adapter) or query (read) may be skipped at no cost
    (align.py:24-34).  All four together give a semiglobal alignment.
    """

    REFERENCE_START = 1
    QUERY_START = 2
    REFERENCE_END = 4
    QUERY_STOP = 8
    SEMIGLOBAL = 15
