"""
Adapter objects with cutadapt's interface (``cutadapt.adapters``, src/cutadapt/adapters.py) in front of the GPU
library: the same class names, constructor arguments, attributes and ``match_to(sequence) -> Match | None``
contract, so that code written against the reference keeps working -- but organised for batches.

* One ``SingleAdapter`` implementation; what distinguishes ``FrontAdapter`` ... ``SuffixAdapter``
  (adapters.py:684-1089) is a row of the ``_KINDS`` table (aligner flags, which side a match removes, the
  arguments of the k-mer heuristic, anchoring, the strings the report uses).  The subclasses only name a row
  and keep the reference's inheritance (``isinstance(a, FrontAdapter)`` holds for the same objects).
* Every adapter can describe itself to the library (``descriptor()``, ``_flatten()``) and match a whole chunk
  with one call (``match_to_batch``); ``match_to`` is a batch of one.  ``matches_from_records`` turns the
  device's 32-byte records back into ``Match`` objects where a caller wants them.
* ``Match`` objects carry the coordinates and know how to apply them (adapters.py:292-493, 1092-1178): one
  ``SingleMatch`` implementation with the removal side as a class attribute.
* Statistics (adapters.py:71-289) are counters keyed by (removed length, errors) plus the five adjacent-base
  counts; they can be fed match by match like the reference's, or in bulk from the dense histograms the
  device reduces (``EndStatistics.add_counts``), and merge with ``+=``.
"""
from collections import Counter
from enum import IntFlag
from typing import Dict, Iterable, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._kmer_finder import KmerFinder
from .align import Aligner, EndSkip, PrefixComparer, SuffixComparer, edit_environment, hamming_sphere
from .kmer_heuristic import create_positions_and_kmers


class InvalidCharacter(Exception):
    pass


class MockKmerFinder:
    """Used where no prefilter applies: every sequence "contains" the k-mers (adapters.py:29-31)."""

    def kmers_present(self, sequence: str):
        return True


class Where(IntFlag):
    """Which ends of adapter and read may be skipped for free, per adapter type (adapters.py:39-53)."""

    BACK = EndSkip.QUERY_START | EndSkip.QUERY_STOP | EndSkip.REFERENCE_END
    FRONT = EndSkip.QUERY_START | EndSkip.QUERY_STOP | EndSkip.REFERENCE_START
    PREFIX = EndSkip.QUERY_STOP
    SUFFIX = EndSkip.QUERY_START
    FRONT_NOT_INTERNAL = EndSkip.REFERENCE_START | EndSkip.QUERY_STOP
    BACK_NOT_INTERNAL = EndSkip.QUERY_START | EndSkip.REFERENCE_END
    ANYWHERE = EndSkip.SEMIGLOBAL


# =============================================================================================================
# Statistics
# =============================================================================================================

ADJACENT_KEYS = ("A", "C", "G", "T", "")


class EndStatistics:
    """
    What was removed at one end by one adapter: how often a piece of length l was removed with e errors, and
    (3' ends) which base preceded it.  Same attributes as the reference's class (adapters.py:71-111):
    ``errors[l][e]``, ``adjacent_bases``, ``lengths``, the adapter's parameters.
    """

    # the adapter parameters the report prints next to the counts (and __iadd__ compares)
    _COPIED = ("max_error_rate", "sequence", "effective_length", "indels", "allows_partial_matches")

    def __init__(self, adapter: "SingleAdapter"):
        for field in self._COPIED:
            setattr(self, field, getattr(adapter, field))
        self.has_wildcards: bool = adapter.adapter_wildcards
        self.adapter_type: str = adapter.descriptive_identifier()
        self._remove_prefix = isinstance(adapter, FrontAdapter)   # (which way random_match_probabilities reads)
        self._counts: Counter = Counter()                    # (length, errors) -> n
        self.adjacent_bases: Dict[str, int] = {k: 0 for k in ADJACENT_KEYS}

    # -- feeding ------------------------------------------------------------------------------------------
    def count(self, length: int, errors: int, adjacent: Optional[str] = None, n: int = 1) -> None:
        self._counts[(length, errors)] += n
        if adjacent is not None:
            self.adjacent_bases[adjacent if adjacent in self.adjacent_bases else ""] += n

    def add_counts(self, histogram, adjacent=None) -> None:
        """Bulk update from dense arrays: histogram[length, errors] and adjacent[5] in ADJACENT_KEYS order."""
        h = np.asarray(histogram)
        for length, errors in zip(*np.nonzero(h)):
            self._counts[(int(length), int(errors))] += int(h[length, errors])
        if adjacent is not None:
            for key, n in zip(ADJACENT_KEYS, adjacent):
                self.adjacent_bases[key] += int(n)

    # -- reading (the reference's attribute names) --------------------------------------------------------------
    @property
    def errors(self) -> Dict[int, Dict[int, int]]:
        out: Dict[int, Dict[int, int]] = {}
        for (length, e), n in self._counts.items():
            if n:
                out.setdefault(length, {})[e] = n
        return out

    @property
    def lengths(self) -> Dict[int, int]:
        out: Dict[int, int] = {}
        for (length, _), n in self._counts.items():
            if n:
                out[length] = out.get(length, 0) + n
        return out

    def __repr__(self):
        return f"EndStatistics(max_error_rate={self.max_error_rate}, errors={self.errors}, adjacent_bases={self.adjacent_bases})"

    def __iadd__(self, other):
        if not isinstance(other, EndStatistics):
            raise ValueError("Cannot compare")
        mine = (self.max_error_rate, self.sequence, self.effective_length, self.indels)
        theirs = (other.max_error_rate, other.sequence, other.effective_length, other.indels)
        if mine != theirs:
            raise RuntimeError("Incompatible EndStatistics, cannot be added")
        self._counts.update(other._counts)
        for key in ADJACENT_KEYS:
            self.adjacent_bases[key] += other.adjacent_bases[key]
        return self

    def random_match_probabilities(self, gc_content: float) -> List[float]:
        """p[i] = probability that i bases of this adapter end match a random sequence with the given GC content
        (no indels), as the report uses it (adapters.py:113-139)."""
        if not 0.0 <= gc_content <= 1.0:
            raise ValueError("gc_content out of range")
        letters = self.sequence[::-1] if self._remove_prefix else self.sequence
        gc_like = "CGRYSKMBDHVN" if self.has_wildcards else "GC"
        out, p = [1.0], 1.0
        for c in letters:
            p *= (gc_content if c in gc_like else 1.0 - gc_content) / 2.0
            out.append(p)
        return out


class AdapterStatistics:
    """Counters of one adapter: a 5' end, a 3' end, or both (adapters.py:142-289)."""

    reverse_complemented: int = 0

    def __init__(self, adapter: "Adapter", front: Optional["SingleAdapter"], back: Optional["SingleAdapter"]):
        self.name = adapter.name
        self.adapter = adapter
        self.front = EndStatistics(front) if front is not None else None
        self.back = EndStatistics(back) if back is not None else None
        self.reverse_complemented = 0

    def end_statistics(self) -> Tuple[Optional[EndStatistics], Optional[EndStatistics]]:
        return self.front, self.back

    def _add_single(self, match: "SingleMatch") -> None:
        if match.remove_before:
            self.front.count(match.removed_sequence_length(), match.errors)
        else:
            self.back.count(match.removed_sequence_length(), match.errors, match.adjacent_base())

    def add_match(self, match) -> None:
        self._add_single(match)

    def __iadd__(self, other):
        if type(other) is not type(self):
            raise ValueError("Cannot iadd")
        for mine, theirs in ((self.front, other.front), (self.back, other.back)):
            if mine is not None:
                mine += theirs
        self.reverse_complemented += other.reverse_complemented
        return self

    def __repr__(self):
        return f"{type(self).__name__}(name={self.name}, front={self.front}, back={self.back})"


class SingleAdapterStatistics(AdapterStatistics):
    """One end only; ``end`` is that end."""

    @property
    def end(self) -> EndStatistics:
        return self.front if self.front is not None else self.back


class FrontAdapterStatistics(SingleAdapterStatistics):
    def __init__(self, adapter: "SingleAdapter"):
        super().__init__(adapter, adapter, None)


class BackAdapterStatistics(SingleAdapterStatistics):
    def __init__(self, adapter: "SingleAdapter"):
        super().__init__(adapter, None, adapter)


class AnywhereAdapterStatistics(AdapterStatistics):
    def __init__(self, adapter: "AnywhereAdapter"):
        super().__init__(adapter, adapter, adapter)


class LinkedAdapterStatistics(AdapterStatistics):
    def __init__(self, adapter: "LinkedAdapter", front: "SingleAdapter", back: "SingleAdapter"):
        super().__init__(adapter, front, back)

    def add_match(self, match: "LinkedMatch") -> None:
        for part in (match.front_match, match.back_match):
            if part:
                self._add_single(part)


# =============================================================================================================
# Matches
# =============================================================================================================


class Match:
    """A place where an adapter was found, and what removing it means."""

    adapter: "Adapter"


class SingleMatch(Match):
    """
    One adapter against one string: adapter[astart:astop] aligned to sequence[rstart:rstop] with `errors` errors
    (adapters.py:321-425).  ``remove_before`` (class attribute) says which side of the read goes with it.
    """

    __slots__ = ("astart", "astop", "rstart", "rstop", "score", "errors", "adapter", "sequence", "length")
    remove_before = False

    def __init__(self, astart: int, astop: int, rstart: int, rstop: int, score: int, errors: int,
                 adapter: "SingleAdapter", sequence: str):
        self.astart, self.astop, self.rstart, self.rstop = astart, astop, rstart, rstop
        self.score, self.errors = score, errors
        self.adapter = adapter
        self.sequence = sequence
        self.length = astop - astart                     # aligned adapter characters

    def _key(self):
        return (self.astart, self.astop, self.rstart, self.rstop, self.score, self.errors)

    def __repr__(self):
        a0, a1, r0, r1, s, e = self._key()
        return f"{type(self).__name__}(astart={a0}, astop={a1}, rstart={r0}, rstop={r1}, score={s}, errors={e})"

    def __eq__(self, other) -> bool:
        return (type(other) is type(self) and self._key() == other._key() and self.adapter is other.adapter
                and self.sequence == other.sequence)

    __hash__ = None

    def match_sequence(self) -> str:
        return self.sequence[self.rstart:self.rstop]

    def wildcards(self, wildcard_char: str = "N") -> str:
        """The read characters under the adapter's wildcard positions (not reliable with indels: the alignment
        itself is not kept)."""
        ref, seq = self.adapter.sequence, self.sequence
        return "".join(seq[self.rstart + i] for i in range(self.length)
                       if ref[self.astart + i] == wildcard_char and self.rstart + i < len(seq))

    def get_info_records(self, read) -> List[List]:
        """One row of the --info-file: errors, coordinates, the three pieces of the read and of its qualities."""
        cuts = ((0, self.rstart), (self.rstart, self.rstop), (self.rstop, None))
        seq, qual = read.sequence, read.qualities
        row = ["", self.errors, self.rstart, self.rstop] + [seq[a:b] for a, b in cuts] + [self.adapter.name]
        row += [qual[a:b] for a, b in cuts] if qual else ["", "", ""]
        return [row]

    # -- what removing the adapter leaves ---------------------------------------------------------------------
    def rest(self) -> str:
        """The part of the read that goes away with the adapter (before a 5' adapter, after a 3' adapter)."""
        return self.sequence[:self.rstart] if self.remove_before else self.sequence[self.rstop:]

    def trim_slice(self) -> slice:
        return slice(self.rstop, None) if self.remove_before else slice(None, self.rstart)

    def remainder_interval(self) -> Tuple[int, int]:
        return (self.rstop, len(self.sequence)) if self.remove_before else (0, self.rstart)

    def retained_adapter_interval(self) -> Tuple[int, int]:
        return (self.rstart, len(self.sequence)) if self.remove_before else (0, self.rstop)

    def trimmed(self, read):
        return read[self.trim_slice()]

    def removed_sequence_length(self) -> int:
        return self.rstop if self.remove_before else len(self.sequence) - self.rstart


class RemoveBeforeMatch(SingleMatch):
    """A match that takes everything in front of it along (5' adapters)."""

    __slots__ = ()
    remove_before = True


class RemoveAfterMatch(SingleMatch):
    """A match that takes everything behind it along (3' adapters)."""

    __slots__ = ()
    remove_before = False

    def adjacent_base(self) -> str:
        return self.sequence[self.rstart - 1:self.rstart]


def remainder(matches: Sequence[Match]) -> Tuple[int, int]:
    """(start, stop) of what is left of the original read after applying the matches in turn, each found in what
    the previous ones left (adapters.py:1598-1613)."""
    if not matches:
        raise ValueError("matches must not be empty")
    offset = 0
    lo = hi = 0
    for match in matches:
        lo, hi = match.remainder_interval()
        offset += lo
    return offset, offset + (hi - lo)


class LinkedMatch(Match):
    """A 5' and/or a 3' match of a LinkedAdapter; the 3' one refers to the read without the 5' part."""

    def __init__(self, front_match: Optional[RemoveBeforeMatch], back_match: Optional[RemoveAfterMatch],
                 adapter: "LinkedAdapter"):
        if front_match is None and back_match is None:
            raise AssertionError("a LinkedMatch needs at least one part")
        self.front_match = front_match
        self.back_match = back_match
        self.adapter = adapter

    def _parts(self) -> List[SingleMatch]:
        return [m for m in (self.front_match, self.back_match) if m is not None]

    def __repr__(self):
        return f"<LinkedMatch(front_match={self.front_match!r}, back_match={self.back_match}, adapter={self.adapter})>"

    @property
    def score(self) -> int:
        return sum(m.score for m in self._parts())

    @property
    def errors(self) -> int:
        return sum(m.errors for m in self._parts())

    def trimmed(self, read):
        for m in self._parts():
            read = m.trimmed(read)
        return read

    def remainder_interval(self) -> Tuple[int, int]:
        return remainder(self._parts())

    def retained_adapter_interval(self) -> Tuple[int, int]:
        start = offset = 0
        if self.front_match:
            start, offset = self.front_match.rstart, self.front_match.rstop
        end = self.back_match.rstop + offset if self.back_match else len(self.front_match.sequence)
        return start, end

    def get_info_records(self, read) -> List[List]:
        label = "none" if self.adapter.name is None else self.adapter.name
        rows = []
        for m, suffix in ((self.front_match, ";1"), (self.back_match, ";2")):
            if m is None:
                continue
            row = m.get_info_records(read)[0]
            row[7] = label + suffix
            rows.append(row)
            read = m.trimmed(read)
        return rows

    def match_sequence(self) -> str:
        front = self.front_match.match_sequence() if self.front_match else ""
        back = self.back_match.match_sequence() if self.back_match else ""
        return front + "," + back


# =============================================================================================================
# Adapters
# =============================================================================================================

_name_counter = [0]


def _generate_adapter_name() -> str:
    _name_counter[0] += 1
    return str(_name_counter[0])


class Matchable:
    """Anything with match_to() -- and, here, match_to_batch()."""

    _device_set = None                                   # (AdapterSet, single adapters, owner of every group)

    def __init__(self, name: Optional[str], *args, **kwargs):
        self.name = name

    def match_to(self, sequence: str):
        return self.match_to_batch([sequence])[0]

    # -- description for the library ---------------------------------------------------------------------------
    def _flatten(self) -> Tuple[List["SingleAdapter"], List[tuple], List["Matchable"]]:
        """(single adapters in device order, group tuples (type, a0, a1, front_required, back_required), the
        Matchable every group belongs to)"""
        raise NotImplementedError

    def _flatten_indexes(self) -> List[dict]:
        """AdapterSetSpec index dicts for the CG_GROUP_INDEXED groups of _flatten(), in a0 order."""
        return []

    def adapter_set(self) -> "_lib.AdapterSet":
        """The compiled tables of this object on the device (built once per context)."""
        ctx = _lib.default_context()
        cached = self._device_set
        if cached is None or cached[0] is None or cached[0].ctx is not ctx:
            singles, groups, owners = self._flatten()
            spec = _lib.AdapterSetSpec([s.descriptor() for s in singles], groups, self._flatten_indexes())
            cached = (_lib.AdapterSet(spec, ctx), singles, owners)
            self._device_set = cached
        return cached[0]

    def matches_from_records(self, records: np.ndarray, sequence: str):
        """The device records of ONE read and ONE round (shape (slots,)) as a Match, or None."""
        _, singles, owners = self._device_set
        first = records[0]
        second = records[1] if len(records) > 1 else None
        shown = first if first["adapter"] >= 0 else second
        if shown is None or shown["adapter"] < 0:
            return None
        owner = owners[int(shown["info"]) & 255]

        def one(rec, seq):
            cls = RemoveAfterMatch if (int(rec["info"]) >> 8) & 1 else RemoveBeforeMatch
            return cls(*(int(rec[f]) for f in ("astart", "astop", "rstart", "rstop", "score", "errors")),
                       adapter=singles[int(rec["adapter"])], sequence=seq)

        if not isinstance(owner, LinkedAdapter):
            return one(first, sequence)
        front = one(first, sequence) if first["adapter"] >= 0 else None
        inner = sequence[front.trim_slice()] if front is not None else sequence
        back = one(second, inner) if second is not None and second["adapter"] >= 0 else None
        return LinkedMatch(front, back, owner)

    def match_to_batch(self, sequences: Sequence[str]) -> List[Optional[Match]]:
        """match_to() for every sequence with one pass of the trimming kernels."""
        data, offsets = _lib.pack_strings(sequences)
        records, _ = self.adapter_set().process(data, offsets)
        return [self.matches_from_records(records[i, 0], s) for i, s in enumerate(sequences)]

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_device_set", None)                   # device handles do not travel between processes
        return state


class Adapter(Matchable):
    description = "adapter with one component"

    def spec(self) -> str:
        raise NotImplementedError

    def create_statistics(self) -> AdapterStatistics:
        raise NotImplementedError

    def descriptive_identifier(self) -> str:
        raise NotImplementedError

    def enable_debug(self) -> None:
        raise NotImplementedError


class _Kind(NamedTuple):
    """What distinguishes the adapter types (one row per class of adapters.py:684-1089)."""
    where: int              # aligner flags
    remove: int             # CG_REMOVE_*: which side a match takes along
    heuristic: tuple        # create_positions_and_kmers(back_adapter, front_adapter, internal)
    reverse: bool           # search the reversed read with the reversed adapter (Rightmost*)
    anchored: Optional[str] # "prefix" / "suffix": min_overlap = whole adapter; comparer when indels are off
    identifier: str         # descriptive_identifier()
    description: str
    spec: str               # spec() with {s} = the sequence
    statistics: type


_BEFORE, _AFTER, _AUTO = _lib.CG_REMOVE_BEFORE, _lib.CG_REMOVE_AFTER, _lib.CG_REMOVE_AUTO
_KINDS = {
    "front": _Kind(Where.FRONT, _BEFORE, (False, True, True), False, None, "regular_five_prime", "regular 5'",
                   "{s}...", FrontAdapterStatistics),
    "rightmost_front": _Kind(Where.BACK, _BEFORE, (True, False, True), True, None, "rightmost_five_prime",
                             "rightmost 5'", "{s}...;rightmost", FrontAdapterStatistics),
    "back": _Kind(Where.BACK, _AFTER, (True, False, True), False, None, "regular_three_prime", "regular 3'", "{s}",
                  BackAdapterStatistics),
    "rightmost_back": _Kind(Where.FRONT, _AFTER, (False, True, True), True, None, "rightmost_three_prime",
                            "rightmost 3'", "{s};rightmost", BackAdapterStatistics),
    "anywhere": _Kind(Where.ANYWHERE, _AUTO, (True, True, True), False, None, "anywhere", "variable 5'/3'",
                      "...{s}...", AnywhereAdapterStatistics),
    "noninternal_front": _Kind(Where.FRONT_NOT_INTERNAL, _BEFORE, (False, True, False), False, None,
                               "noninternal_five_prime", "non-internal 5'", "X{s}...", FrontAdapterStatistics),
    "noninternal_back": _Kind(Where.BACK_NOT_INTERNAL, _AFTER, (True, False, False), False, None,
                              "noninternal_three_prime", "non-internal 3'", "{s}X", BackAdapterStatistics),
    "prefix": _Kind(Where.PREFIX, _BEFORE, (False, True, False), False, "prefix", "anchored_five_prime",
                    "anchored 5'", "^{s}...", FrontAdapterStatistics),
    "suffix": _Kind(Where.SUFFIX, _AFTER, (True, False, False), False, "suffix", "anchored_three_prime",
                    "anchored 3'", "{s}$", BackAdapterStatistics),
}

_IUPAC_LETTERS = frozenset("ABCDGHKMNRSTUVWXY")


class SingleAdapter(Adapter):
    """
    One adapter: sequence, error rate, minimum overlap, wildcard and indel switches (the reference's constructor
    arguments, adapters.py:533-599) plus its type (``_kind``).  Construction normalises the sequence, builds the
    aligner (or comparer) and the k-mer prefilter tables; matching is done by the library.
    """

    _kind: _Kind = None

    def __init__(self, sequence: str, max_errors: float = 0.1, min_overlap: int = 3, read_wildcards: bool = False,
                 adapter_wildcards: bool = True, name: Optional[str] = None, indels: bool = True,
                 force_anywhere: bool = False):
        kind = self._kind
        if kind is None:
            raise TypeError("SingleAdapter is abstract: use one of the adapter type classes")
        self.name: str = name if name is not None else _generate_adapter_name()
        super().__init__(self.name)
        self._debug = False
        # "-a ADAPTER;anywhere" (parser.py:540): a 5' or 3' adapter searched like -b but removed on its own side
        self._force_anywhere = bool(force_anywhere)
        text = sequence.upper().replace("U", "T").replace("I", "N")
        if not text:
            raise ValueError("Adapter sequence is empty")
        if kind.anchored:
            min_overlap = len(sequence)                  # anchored adapters match as a whole (adapters.py:1025, 1060)
        n_wild = text.count("N")
        if max_errors >= 1 and n_wild != len(text):      # an absolute number of errors becomes a rate
            max_errors /= len(text) - n_wild
        self.sequence: str = text
        self.max_error_rate: float = max_errors
        self.min_overlap: int = min(min_overlap, len(text))
        letters = set(text)
        if adapter_wildcards:
            unknown = [c for c in text if c not in _IUPAC_LETTERS]
            if unknown:
                raise InvalidCharacter(
                    f"Character '{unknown[0]}' in adapter sequence '{text}' is "
                    f"not a valid IUPAC code. Use only characters 'ABCDGHIKMNRSTUVWXY'.")
        # plain A/C/G/T adapters are matched without the wildcard tables
        self.adapter_wildcards: bool = bool(adapter_wildcards) and not letters <= set("ACGT")
        self.read_wildcards: bool = read_wildcards
        self.indels: bool = indels
        self.aligner = self._aligner()
        self.kmer_finder = self._kmer_finder()

    # -- the reference's class attributes, from the table ------------------------------------------------------
    @property
    def description(self) -> str:
        return self._kind.description

    @property
    def allows_partial_matches(self) -> bool:
        return self._kind.anchored is None

    @property
    def remove_before(self) -> bool:
        return self._kind.remove == _BEFORE

    def descriptive_identifier(self) -> str:
        return self._kind.identifier

    def spec(self) -> str:
        return self._kind.spec.format(s=self.sequence)

    def create_statistics(self) -> AdapterStatistics:
        return self._kind.statistics(self)

    @property
    def effective_length(self) -> int:
        return self.aligner.effective_length

    def enable_debug(self) -> None:
        self._debug = True
        if hasattr(self.aligner, "enable_debug"):
            self.aligner.enable_debug()

    def __len__(self) -> int:
        return len(self.sequence)

    def __repr__(self):
        fields = ("name", "sequence", "max_error_rate", "min_overlap", "read_wildcards", "adapter_wildcards", "indels")
        inner = ", ".join(f"{f}={getattr(self, f)!r}" for f in fields)
        return f"<{type(self).__name__}({inner})>"

    # -- native objects ------------------------------------------------------------------------------------------
    def _searched(self) -> str:
        return self.sequence[::-1] if self._kind.reverse else self.sequence

    def _make_aligner(self, sequence: str, flags: int) -> Aligner:
        # without indels the aligner still runs, with indels priced out of reach (adapters.py:601-616)
        return Aligner(sequence, self.max_error_rate, flags=flags, wildcard_ref=self.adapter_wildcards,
                       wildcard_query=self.read_wildcards, indel_cost=1 if self.indels else 100000,
                       min_overlap=self.min_overlap)

    def _aligner(self):
        kind = self._kind
        if kind.anchored and not self.indels:
            comparer = PrefixComparer if kind.anchored == "prefix" else SuffixComparer
            return comparer(self.sequence, self.max_error_rate, wildcard_ref=self.adapter_wildcards,
                            wildcard_query=self.read_wildcards, min_overlap=self.min_overlap)
        where = kind.where
        if self._force_anywhere and where in (Where.FRONT, Where.BACK):
            where = Where.ANYWHERE
        return self._make_aligner(self._searched(), int(where))

    def _make_kmer_finder(self, sequence: str, back_adapter: bool, front_adapter: bool, internal: bool = True):
        tables = create_positions_and_kmers(sequence, self.min_overlap, self.max_error_rate, back_adapter,
                                            front_adapter, internal)
        try:
            return KmerFinder(tables, self.adapter_wildcards, self.read_wildcards)
        except ValueError:                               # a k-mer longer than one machine word
            return MockKmerFinder()

    def _kmer_finder(self):
        if not isinstance(self.aligner, Aligner):        # comparers look at a fixed place: nothing to prefilter
            return MockKmerFinder()
        back, front, internal = self._kind.heuristic
        if self._force_anywhere:
            back = front = True
        return self._make_kmer_finder(self._searched(), back_adapter=back, front_adapter=front, internal=internal)

    # -- description for the library ---------------------------------------------------------------------------
    def descriptor(self) -> dict:
        """The cg_adapter_desc of this adapter: its aligner's arguments plus the prefilter tables."""
        d = self.aligner._descriptor()
        d["reverse_read"] = self._kind.reverse
        d["remove"] = self._kind.remove
        if isinstance(self.kmer_finder, KmerFinder):
            d["kmer_entries"], d["kmer_masks"] = self.kmer_finder.tables
        return d

    def _flatten(self):
        return [self], [(_lib.CG_GROUP_SINGLE, 0, -1, 0, 0)], [self]


# The type classes: a table row each, with the reference's inheritance (adapters.py:684, 733, 792, 841, 896, 944,
# 981, 1018, 1052).


class FrontAdapter(SingleAdapter):
    """A 5' adapter: everything up to and including the match goes."""
    _kind = _KINDS["front"]


class RightmostFrontAdapter(FrontAdapter):
    """A 5' adapter that prefers the rightmost occurrence: searched on the reversed read."""
    _kind = _KINDS["rightmost_front"]


class BackAdapter(SingleAdapter):
    """A 3' adapter: the match and everything behind it go."""
    _kind = _KINDS["back"]


class RightmostBackAdapter(BackAdapter):
    """A 3' adapter that prefers the rightmost occurrence: searched on the reversed read."""
    _kind = _KINDS["rightmost_back"]


class AnywhereAdapter(SingleAdapter):
    """5' or 3': a match that starts at the first base of the read counts as 5', any other as 3'."""
    _kind = _KINDS["anywhere"]


class NonInternalFrontAdapter(FrontAdapter):
    """A 5' adapter that may hang over the start of the read but not lie inside it (-g XADAPTER)."""
    _kind = _KINDS["noninternal_front"]


class NonInternalBackAdapter(BackAdapter):
    """A 3' adapter that may hang over the end of the read but not lie inside it (-a ADAPTERX)."""
    _kind = _KINDS["noninternal_back"]


class PrefixAdapter(NonInternalFrontAdapter):
    """An anchored 5' adapter (-g ^ADAPTER)."""
    _kind = _KINDS["prefix"]


class SuffixAdapter(NonInternalBackAdapter):
    """An anchored 3' adapter (-a ADAPTER$)."""
    _kind = _KINDS["suffix"]


class LinkedAdapter(Adapter):
    """A 5' adapter followed by a 3' adapter, the latter searched in what the former leaves (adapters.py:1181-1243)."""

    description = "linked"

    def __init__(self, front_adapter: SingleAdapter, back_adapter: SingleAdapter, front_required: bool,
                 back_required: bool, name: Optional[str]):
        super().__init__(name)
        self.name: str = name if name is not None else _generate_adapter_name()
        self.where = "linked"
        self.front_required, self.back_required = front_required, back_required
        self.front_adapter, self.back_adapter = front_adapter, back_adapter
        front_adapter.name = self.name                   # the report shows the pair under one name

    def __repr__(self):
        return f"{type(self).__name__}(front_adapter={self.front_adapter}, back_adapter={self.back_adapter})"

    def descriptive_identifier(self) -> str:
        return "linked"

    def enable_debug(self) -> None:
        self.front_adapter.enable_debug()
        self.back_adapter.enable_debug()

    def create_statistics(self) -> LinkedAdapterStatistics:
        return LinkedAdapterStatistics(self, front=self.front_adapter, back=self.back_adapter)

    @property
    def sequence(self) -> str:
        return f"{self.front_adapter.sequence}...{self.back_adapter.sequence}"

    @property
    def remove(self):
        return None

    def spec(self) -> str:
        return f"{self.front_adapter.spec()}...{self.back_adapter.spec()}"

    def _flatten(self):
        group = (_lib.CG_GROUP_LINKED, 0, 1, int(bool(self.front_required)), int(bool(self.back_required)))
        return [self.front_adapter, self.back_adapter], [group], [self]


class MultipleAdapters(Matchable):
    """
    Several adapters at once; the best match wins: highest score, then fewest errors, then the adapter listed
    first (adapters.py:1246-1286).  The library evaluates all of them on the device and applies that rule there.
    """

    def __init__(self, adapters: Sequence[Matchable]):
        super().__init__(name="multiple_adapters")
        self._adapters = adapters

    def __getitem__(self, item):
        return self._adapters[item]

    def __len__(self):
        return len(self._adapters)

    def enable_debug(self) -> None:
        for adapter in self._adapters:
            adapter.enable_debug()

    def _flatten(self):
        singles: List[SingleAdapter] = []
        groups: List[tuple] = []
        owners: List[Matchable] = []
        n_indexes = 0
        for adapter in self._adapters:
            sub_singles, sub_groups, sub_owners = adapter._flatten()
            shift = len(singles)
            for typ, a0, a1, front_required, back_required in sub_groups:
                if typ == _lib.CG_GROUP_INDEXED:
                    groups.append((typ, a0 + n_indexes, -1, front_required, back_required))
                else:
                    groups.append((typ, a0 + shift, a1 + shift if a1 >= 0 else -1, front_required, back_required))
            n_indexes += sum(g[0] == _lib.CG_GROUP_INDEXED for g in sub_groups)
            singles += sub_singles
            owners += sub_owners
        return singles, groups, owners

    def _flatten_indexes(self):
        out: List[dict] = []
        shift = 0
        for adapter in self._adapters:
            for index in adapter._flatten_indexes():
                out.append(dict(index, adapter=[a + shift for a in index["adapter"]]))
            shift += len(adapter._flatten()[0])
        return out


class AdapterIndex:
    """
    Many anchored adapters of one type behind a dictionary (adapters.py:1289-1551): every string within the allowed
    edit (indels) or Hamming (no indels) distance of an adapter is a key; a read is matched by looking up its prefix
    (suffix) once per distinct key length.  Keys that two adapters reach with the same number of matches are
    ambiguous and dropped (such reads stay untrimmed).  The dictionary is built here and handed to the library,
    which keeps it as a hash table in HBM and does the lookups for a whole chunk inside the trimming kernel.

    Acceptable are the reference's adapters (no wildcards, at most 3 errors, adapters.py:1366-1378) that also fit
    the device table: A/C/G/T only, keys of at most 32 characters, fewer errors than characters; callers fall back
    to MultipleAdapters for the rest exactly as they do for the reference's own restrictions.
    """

    def __init__(self, adapters, prefix: bool):
        if not adapters:
            raise ValueError("Adapter list is empty")
        for adapter in adapters:
            self._accept(adapter, prefix)
        self._adapters = adapters
        self._prefix = prefix
        self._lengths, self._index, self._ambiguous = self._make_index()
        if len(self._lengths) == 1:
            self._length = self._lengths[0]

    def __repr__(self):
        return f"{type(self).__name__}(adapters={self._adapters!r})"

    @classmethod
    def _accept(cls, adapter: SingleAdapter, prefix: bool) -> None:
        wanted = PrefixAdapter if prefix else SuffixAdapter
        if not isinstance(adapter, wanted):
            raise ValueError("Only 5' anchored adapters are allowed" if prefix else "Only 3' anchored adapters are allowed")
        if adapter.read_wildcards:
            raise ValueError("Wildcards in the read not supported")
        if adapter.adapter_wildcards:
            raise ValueError("Wildcards in the adapter not supported")
        k = int(len(adapter) * adapter.max_error_rate)
        if k > 3:
            raise ValueError("Error rate too high")
        if k >= len(adapter):
            raise ValueError("As many errors as characters allowed: not indexable on the device")
        if len(adapter) + (k if adapter.indels else 0) > 32:
            raise ValueError("Adapter too long for the device index")
        if set(adapter.sequence) - set("ACGT"):
            raise ValueError("Only A, C, G, T adapters can be indexed on the device")

    @classmethod
    def is_acceptable(cls, adapter: SingleAdapter, prefix: bool) -> bool:
        try:
            cls._accept(adapter, prefix)
        except ValueError:
            return False
        return True

    @staticmethod
    def _neighbourhood(adapter: SingleAdapter) -> Iterable[Tuple[str, int, int]]:
        """(string, errors, matches) for everything the adapter may look like in a read."""
        text = adapter.sequence
        k = int(adapter.max_error_rate * len(text))
        if adapter.indels:
            yield from edit_environment(text, k)
        else:
            for errors in range(k + 1):
                for s in hamming_sphere(text, errors):
                    yield s, errors, len(text) - errors

    def _make_index(self):
        table: Dict[str, Tuple[SingleAdapter, int, int]] = {}
        lengths = set()
        contested = set()
        for adapter in self._adapters:
            for s, errors, matches in self._neighbourhood(adapter):
                holder = table.get(s)
                if holder is not None:
                    if matches < holder[2]:
                        continue                         # the earlier adapter explains this string better
                    if matches == holder[2]:
                        contested.add(s)
                table[s] = (adapter, errors, matches)
                lengths.add(len(s))
        for s in contested:
            del table[s]
        return sorted(lengths, reverse=True), table, len(contested)

    def descriptor(self) -> dict:
        """The cg_index_desc content: the keys with (adapter number, errors, matches) each."""
        number = {id(a): i for i, a in enumerate(self._adapters)}
        keys = list(self._index)
        values = [self._index[k] for k in keys]
        return {"prefix": self._prefix, "lengths": list(self._lengths), "keys": keys,
                "adapter": [number[id(v[0])] for v in values], "errors": [v[1] for v in values],
                "matches": [v[2] for v in values]}


class _IndexedAdapters(Matchable):
    _is_prefix = True

    def __init__(self, adapters, name):
        super().__init__(name=name)
        self._index = AdapterIndex(adapters, prefix=self._is_prefix)

    def _flatten(self):
        return list(self._index._adapters), [(_lib.CG_GROUP_INDEXED, 0, -1, 0, 0)], [self]

    def _flatten_indexes(self):
        return [self._index.descriptor()]


class IndexedPrefixAdapters(_IndexedAdapters):
    """Anchored 5' adapters behind one index (adapters.py:1554-1561)."""

    _is_prefix = True

    def __init__(self, adapters):
        super().__init__(adapters, name="indexed_prefix_adapters")


class IndexedSuffixAdapters(_IndexedAdapters):
    """Anchored 3' adapters behind one index (adapters.py:1564-1571)."""

    _is_prefix = False

    def __init__(self, adapters):
        super().__init__(adapters, name="indexed_suffix_adapters")


def warn_duplicate_adapters(adapters) -> List[str]:
    """Adapters given more than once (same type and sequence), as messages; the reference logs them
    (adapters.py:1574-1595)."""
    seen = set()
    messages = []
    for adapter in adapters:
        key = (type(adapter), adapter.sequence)
        if key in seen:
            messages.append(f"Adapter {adapter.sequence!r} ({adapter.description}) was specified multiple times! "
                            "Please make sure that this is what you want.")
        seen.add(key)
    return messages
