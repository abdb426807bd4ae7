"""
Mirror of ``cutadapt._align`` (src/cutadapt/_align.pyx; stub _align.pyi): ``Aligner``,
``PrefixComparer``, ``SuffixComparer``, ``hamming_sphere``, ``edit_environment`` with the
reference's signatures, attributes, pickling and error behaviour.

``locate(query)`` is a batch of one on the GPU; ``locate_batch(queries)`` is what the batched
pipeline uses.  Adapter tables are compiled and uploaded lazily on first use, so the objects
stay cheap to construct and picklable (the reference ships the whole Pipeline to worker
processes, runners.py:345-356).
"""
import ctypes as C
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

Alignment = Tuple[int, int, int, int, int, int]


class _Located:
    """Shared batch plumbing: one adapter, no prefilter, results as locate() tuples."""

    _set = None

    def _descriptor(self) -> dict:  # pragma: no cover - overridden
        raise NotImplementedError

    def _adapterset(self) -> "_lib.AdapterSet":
        current = self._set
        ctx = _lib.default_context()
        if current is None or current.ctx is not ctx:
            current = _lib.AdapterSet(_lib.AdapterSetSpec([self._descriptor()]), ctx)
            self._set = current
        return current

    def locate_batch(self, queries: Sequence[str]) -> List[Optional[Alignment]]:
        data, offsets = _lib.pack_strings(queries)
        matches, _ = self._adapterset().process(data, offsets)
        m = matches[:, 0, 0]
        out: List[Optional[Alignment]] = []
        for rec in m:
            if rec["adapter"] < 0:
                out.append(None)
            else:
                out.append(
                    (int(rec["astart"]), int(rec["astop"]), int(rec["rstart"]), int(rec["rstop"]),
                     int(rec["score"]), int(rec["errors"]))
                )
        return out

    def locate(self, query: str) -> Optional[Alignment]:
        if not isinstance(query, str):
            raise TypeError("query must be str")
        return self.locate_batch([query])[0]

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_set", None)
        return state


class Aligner(_Located):
    """
    Find a full or partial occurrence of a query string in a reference string allowing errors
    (see the reference docstring, _align.pyx:93-169, and doc/algorithms.rst).
    """

    def __init__(
        self,
        reference: str,
        max_error_rate: float,
        flags: int = 15,
        wildcard_ref: bool = False,
        wildcard_query: bool = False,
        indel_cost: int = 1,
        min_overlap: int = 1,
    ):
        if not isinstance(reference, str):
            raise TypeError("reference must be str")
        self.max_error_rate = float(max_error_rate)
        self.flags = int(flags) & 15
        self.wildcard_ref = bool(wildcard_ref)
        self.wildcard_query = bool(wildcard_query)
        self._min_overlap = int(min_overlap)
        if indel_cost < 1:
            raise ValueError("indel_cost must be at least 1")
        self._indel_cost = int(indel_cost)
        if not reference.isascii():
            raise ValueError("String must contain only ASCII characters")
        self.reference = reference
        n_count = reference.count("N") + reference.count("n")
        self._effective_length = len(reference)
        if self.wildcard_ref:
            self._effective_length = len(reference) - n_count
            if self._effective_length == 0:
                raise ValueError("Cannot have only N wildcards in the sequence")
        self._debug = False

    @property
    def effective_length(self) -> int:
        return self._effective_length

    def _descriptor(self) -> dict:
        return dict(
            sequence=self.reference,
            max_error_rate=self.max_error_rate,
            flags=self.flags,
            wildcard_ref=self.wildcard_ref,
            wildcard_query=self.wildcard_query,
            indel_cost=self._indel_cost,
            min_overlap=self._min_overlap,
            kind=_lib.CG_KIND_ALIGNER,
        )

    def __reduce__(self):
        return (
            Aligner,
            (self.reference, self.max_error_rate, self.flags, self.wildcard_ref, self.wildcard_query,
             self._indel_cost, self._min_overlap),
        )

    def __repr__(self):
        return (
            f"Aligner(reference='{self.reference}', max_error_rate={self.max_error_rate}, "
            f"flags={self.flags}, wildcard_ref={self.wildcard_ref}, "
            f"wildcard_query={self.wildcard_query}, indel_cost={self._indel_cost}, "
            f"min_overlap={self._min_overlap})"
        )

    # -- debugging aid: the DP matrices of one read (enable_debug / dpmatrix / scorematrix, _align.pyx:279-296) -------
    dpmatrix = None
    scorematrix = None

    def enable_debug(self):
        """Keep the dynamic-programming matrices of every following locate() call in .dpmatrix / .scorematrix."""
        self._debug = True

    def locate(self, query: str) -> Optional[Alignment]:
        if not getattr(self, "_debug", False):
            return super().locate(query)
        if not isinstance(query, str):
            raise TypeError("query must be str")
        import ctypes as C

        spec = _lib.AdapterSetSpec([self._descriptor()])
        arr, _, _, _ = spec.to_ctypes()
        q = query.encode("ascii", "replace") if query.isascii() else query.encode("latin-1", "replace")
        m, n = len(self.reference), len(q)
        cost = np.empty((m + 1, n + 1), dtype=np.int32)
        score = np.empty((m + 1, n + 1), dtype=np.int32)
        res = np.zeros(8, dtype=np.int32)
        _lib.check(_lib.lib().cg_locate_debug(_lib.default_context().handle, arr, q, n, cost.ctypes.data,
                                              score.ctypes.data, res.ctypes.data))
        self.dpmatrix = DPMatrix(self.reference, query, cost)
        self.scorematrix = DPMatrix(self.reference, query, score)
        return tuple(int(x) for x in res[1:7]) if res[0] else None


class DPMatrix:
    """The cells the search computed, None where it never went; prints like the reference's (_align.pyx:58-90)."""

    NONE = -(1 << 31)

    def __init__(self, reference: str, query: str, values):
        self.reference, self.query = reference, query
        self._rows = [[None if v == self.NONE else int(v) for v in row] for row in values]

    def __str__(self):
        lines = ["     " + " ".join(c.rjust(2) for c in self.query)]
        for label, row in zip(" " + self.reference, self._rows):
            lines.append(label + " " + " ".join("  " if v is None else f"{v:2d}" for v in row))
        return "\n".join(lines)


class PrefixComparer(_Located):
    """Anchored 5' comparison without indels (_align.pyx:594-693)."""

    _kind = _lib.CG_KIND_PREFIX_COMPARER

    def __init__(
        self,
        reference: str,
        max_error_rate: float,
        wildcard_ref: bool = False,
        wildcard_query: bool = False,
        min_overlap: int = 1,
    ):
        if not reference.isascii():
            raise ValueError("String must contain only ASCII characters")
        self.wildcard_ref = bool(wildcard_ref)
        self.wildcard_query = bool(wildcard_query)
        self._reference = reference
        self.m = len(reference)
        self._effective_length = self.m
        if self.wildcard_ref:
            # sic: count('N') - count('n'), as in the reference (_align.pyx:628)
            self._effective_length -= reference.count("N") - reference.count("n")
            if self._effective_length == 0:
                raise ValueError("Cannot have only N wildcards in the sequence")
        if not (0 <= max_error_rate <= 1.0):
            raise ValueError("max_error_rate must be between 0 and 1")
        self.max_error_rate = float(max_error_rate)
        self.max_k = int(max_error_rate * self._effective_length)
        if min_overlap < 1:
            raise ValueError("min_overlap must be at least 1")
        self.min_overlap = int(min_overlap)

    @property
    def effective_length(self) -> int:
        return self._effective_length

    def _descriptor(self) -> dict:
        return dict(
            sequence=self._reference,
            max_error_rate=self.max_error_rate,
            wildcard_ref=self.wildcard_ref,
            wildcard_query=self.wildcard_query,
            min_overlap=self.min_overlap,
            kind=self._kind,
        )

    def __reduce__(self):
        return (
            type(self),
            (self._reference, self.max_error_rate, self.wildcard_ref, self.wildcard_query, self.min_overlap),
        )

    def __repr__(self):
        return "{}(reference={!r}, max_k={}, wildcard_ref={}, wildcard_query={})".format(
            self.__class__.__name__, self._reference, self.max_k, self.wildcard_ref, self.wildcard_query
        )


class SuffixComparer(PrefixComparer):
    """Anchored 3' comparison without indels (_align.pyx:696-714)."""

    _kind = _lib.CG_KIND_SUFFIX_COMPARER


# ---- neighbourhood generators (host side; used to build the anchored-adapter index) --------


def _environment(fn_name: str, s: str, k: int, with_lengths: bool):
    if not s.isascii():
        raise ValueError("String must contain only ASCII characters")
    lib = _lib.lib()
    data = s.encode("ascii")
    n = len(data)
    stride = n + k + 1
    fn = getattr(lib, fn_name)
    buf = C.create_string_buffer(data, max(n, 1))
    if with_lengths:
        count = fn(buf, n, k, stride, None, None, None, None, 0)
    else:
        count = fn(buf, n, k, stride, None, None, None, 0)
    if count < 0:
        _lib.check(int(count))
    strings = np.zeros((max(count, 1), stride), dtype=np.uint8)
    lengths = np.zeros(max(count, 1), dtype=np.int32)
    errors = np.zeros(max(count, 1), dtype=np.int32)
    matches = np.zeros(max(count, 1), dtype=np.int32)
    if with_lengths:
        fn(buf, n, k, stride, strings.ctypes.data, lengths.ctypes.data, errors.ctypes.data,
           matches.ctypes.data, count)
    else:
        fn(buf, n, k, stride, strings.ctypes.data, errors.ctypes.data, matches.ctypes.data, count)
        lengths[:] = n
    raw = strings.tobytes()
    for i in range(count):
        yield raw[i * stride : i * stride + lengths[i]].decode("ascii"), int(errors[i]), int(matches[i])


def edit_environment(t: str, k: int) -> Iterator[Tuple[str, int, int]]:
    """
    All strings s over ACGT with edit distance(s, t) <= k as (s, errors, matches)
    (_align.pyx:785-882).
    """
    return _environment("cg_edit_environment", t, k, True)


def hamming_environment(s: str, k: int) -> Iterator[Tuple[str, int, int]]:
    """All strings t over ACGT with Hamming distance(s, t) <= k as (t, errors, matches) (align.py:63-75)."""
    return _environment("cg_hamming_environment", s, k, False)


def hamming_sphere(s: str, k: int) -> Iterator[str]:
    """All strings t over ACGT with Hamming distance(s, t) == k (_align.pyx:717-782)."""
    for t, errors, _ in hamming_environment(s, k):
        if errors == k:
            yield t
