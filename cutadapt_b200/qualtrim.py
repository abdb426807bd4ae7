"""
Mirror of the hot-path part of ``cutadapt.qualtrim`` (src/cutadapt/qualtrim.pyx:18-73):
``quality_trim_index`` with the reference's signature, executed on the GPU, plus a batched
variant.  (In the fused trimming pass the same device function runs as the first stage of the
kernel; see cutadapt_b200.pipeline.)
"""
import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib


class HasNoQualities(Exception):
    pass


def quality_trim_index_batch(
    qualities: Sequence[str], cutoff_front: int, cutoff_back: int, base: int = 33
) -> np.ndarray:
    """(n, 2) int32 array of (start, stop) for every quality string."""
    if any(q is None for q in qualities):
        raise HasNoQualities("Cannot do quality trimming when no qualities are available")
    try:
        joined = "".join(qualities).encode("latin-1")
    except UnicodeEncodeError:
        raise ValueError("Quality data is not ASCII") from None
    n = len(qualities)
    out = np.zeros((n, 2), dtype=np.int32)
    if n == 0:
        return out
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(q) for q in qualities], out=offsets[1:])
    data = np.frombuffer(joined, dtype=np.uint8) if joined else np.zeros(1, dtype=np.uint8)
    ctx = _lib.default_context()
    _lib.check(
        _lib.lib().cg_quality_trim_batch(
            ctx.handle, data.ctypes.data, offsets.ctypes.data, n, int(cutoff_front), int(cutoff_back),
            int(base), out.ctypes.data,
        )
    )
    return out


def quality_trim_index(
    qualities: Optional[str], cutoff_front: int, cutoff_back: int, base: int = 33
) -> Tuple[int, int]:
    """
    Positions (start, stop) of the good-quality segment, BWA style (qualtrim.pyx:22-73).
    """
    if qualities is None:
        raise HasNoQualities("Cannot do quality trimming when no qualities are available")
    start, stop = quality_trim_index_batch([qualities], cutoff_front, cutoff_back, base)[0]
    return int(start), int(stop)


def nextseq_trim_index_batch(sequences: Sequence[str], qualities: Sequence[str], cutoff: int, base: int = 33) -> np.ndarray:
    """int32 array: for every read the index at which its 3' end is cut (NextSeq two-colour chemistry)."""
    if any(q is None for q in qualities):
        raise HasNoQualities("Cannot do quality trimming when no qualities are available")
    n = len(qualities)
    out = np.zeros(n, dtype=np.int32)
    if n == 0:
        return out
    for s, q in zip(sequences, qualities):
        if len(s) != len(q):
            raise ValueError("sequence and qualities differ in length")
    try:
        qjoined = "".join(qualities).encode("latin-1")
    except UnicodeEncodeError:
        raise ValueError("Quality data is not ASCII") from None
    sjoined = "".join(sequences).encode("latin-1", errors="replace")   # only 'G' is ever compared
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(q) for q in qualities], out=offsets[1:])
    qdata = np.frombuffer(qjoined, dtype=np.uint8) if qjoined else np.zeros(1, dtype=np.uint8)
    sdata = np.frombuffer(sjoined, dtype=np.uint8) if sjoined else np.zeros(1, dtype=np.uint8)
    ctx = _lib.default_context()
    _lib.check(_lib.lib().cg_nextseq_trim_batch(ctx.handle, sdata.ctypes.data, qdata.ctypes.data, offsets.ctypes.data, n,
                                                int(cutoff), int(base), out.ctypes.data))
    return out


def nextseq_trim_index(sequence, cutoff: int, base: int = 33) -> int:
    """
    Variant of quality trimming for NextSeq data: qualities of 'G' bases count as cutoff - 1
    (qualtrim.pyx:76-117).  ``sequence`` is a record with ``.sequence`` and ``.qualities``.
    """
    if sequence.qualities is None:
        raise HasNoQualities("Cannot do quality trimming when no qualities are available")
    return int(nextseq_trim_index_batch([sequence.sequence], [sequence.qualities], cutoff, base)[0])


def poly_a_trim_index_batch(sequences: Sequence[str], revcomp: bool = False) -> np.ndarray:
    """int32 array: start index of every read's poly-A tail (revcomp: end of its poly-T head)."""
    n = len(sequences)
    out = np.zeros(n, dtype=np.int32)
    if n == 0:
        return out
    try:
        joined = "".join(sequences).encode("latin-1")
    except UnicodeEncodeError:
        raise ValueError("Sequence is not ASCII") from None
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(s) for s in sequences], out=offsets[1:])
    data = np.frombuffer(joined, dtype=np.uint8) if joined else np.zeros(1, dtype=np.uint8)
    ctx = _lib.default_context()
    _lib.check(_lib.lib().cg_poly_a_trim_batch(ctx.handle, data.ctypes.data, offsets.ctypes.data, n, int(bool(revcomp)),
                                               out.ctypes.data))
    return out


def poly_a_trim_index(s: str, revcomp: bool = False) -> int:
    """Start index of the poly-A tail; with revcomp the end of the poly-T head (qualtrim.pyx:120-169)."""
    return int(poly_a_trim_index_batch([s], revcomp)[0])


def expected_errors_batch(qualities: Sequence[str], base: int = 33) -> np.ndarray:
    """float64 array: expected number of errors of every quality string (Edgar & Flyvbjerg 2015)."""
    n = len(qualities)
    out = np.zeros(n, dtype=np.float64)
    if n == 0:
        return out
    for q in qualities:
        if not q.isascii():
            raise ValueError(f"Quality string contains non-ASCII values: {q}")
    joined = "".join(qualities).encode("ascii")
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(q) for q in qualities], out=offsets[1:])
    data = np.frombuffer(joined, dtype=np.uint8) if joined else np.zeros(1, dtype=np.uint8)
    ctx = _lib.default_context()
    _lib.check(_lib.lib().cg_expected_errors_batch(ctx.handle, data.ctypes.data, offsets.ctypes.data, n, int(base),
                                                   out.ctypes.data))
    for i in np.nonzero(out < 0.0)[0]:
        for ch in qualities[int(i)]:
            if ord(ch) < base or ord(ch) > 126:
                raise ValueError(f"Not a valid phred value {ord(ch)} for character {ch}")
    return out


def expected_errors(qualities: str, base: int = 33) -> float:
    """Number of expected errors of a read from its qualities (qualtrim.pyx:172-197), bit-identical double."""
    return float(expected_errors_batch([qualities], base)[0])
