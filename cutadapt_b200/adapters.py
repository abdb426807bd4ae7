"""
Adapter finding classes with cutadapt's interface, backed by the GPU kernels.

Mirror of the matching part of ``cutadapt.adapters`` (src/cutadapt/adapters.py): the adapter
type zoo (``FrontAdapter`` ... ``SuffixAdapter``, lines 684-1089), ``LinkedAdapter`` (1181-1243),
``MultipleAdapters`` (1246-1286) and the ``Match`` classes that describe / apply a hit
(292-493, 1092-1178).  Every class keeps the reference's constructor arguments, attributes and
``match_to(sequence) -> Match | None`` contract.

What is different is *where the work happens*: ``match_to`` is a batch of one on the device,
and every Matchable additionally has ``match_to_batch(sequences)`` which runs prefilter,
alignment and best-adapter selection for a whole chunk of reads in one fused kernel launch
(see ``cutadapt_b200.pipeline`` for the per-chunk driver).
"""
from abc import ABC, abstractmethod
from collections import defaultdict
from enum import IntFlag
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from ._kmer_finder import KmerFinder
from .align import Aligner, EndSkip, PrefixComparer, SuffixComparer, edit_environment, hamming_sphere
from .kmer_heuristic import create_positions_and_kmers


class MockKmerFinder:
    """Stand-in used when no prefilter applies (adapters.py:29-31)."""

    def kmers_present(self, sequence: str):
        return True


class InvalidCharacter(Exception):
    pass


class Where(IntFlag):
    """Aligner flag combinations for all adapter types (adapters.py:39-53)."""

    BACK = EndSkip.QUERY_START | EndSkip.QUERY_STOP | EndSkip.REFERENCE_END
    FRONT = EndSkip.QUERY_START | EndSkip.QUERY_STOP | EndSkip.REFERENCE_START
    PREFIX = EndSkip.QUERY_STOP
    SUFFIX = EndSkip.QUERY_START
    FRONT_NOT_INTERNAL = EndSkip.REFERENCE_START | EndSkip.QUERY_STOP
    BACK_NOT_INTERNAL = EndSkip.QUERY_START | EndSkip.REFERENCE_END
    ANYWHERE = EndSkip.SEMIGLOBAL


# ---------------------------------------------------------------------------------------------
# Statistics (adapters.py:71-289) -- the per-adapter counters the hot path feeds
# ---------------------------------------------------------------------------------------------


def returns_defaultdict_int():
    return defaultdict(int)


class EndStatistics:
    """Statistics about the 5' or 3' end (adapters.py:71-175)."""

    def __init__(self, adapter: "SingleAdapter"):
        self.max_error_rate: float = adapter.max_error_rate
        self.sequence: str = adapter.sequence
        self.effective_length: int = adapter.effective_length
        self.has_wildcards: bool = adapter.adapter_wildcards
        self.indels: bool = adapter.indels
        self.adapter_type: str = adapter.descriptive_identifier()
        self.allows_partial_matches: bool = adapter.allows_partial_matches
        # errors[removed_length][n_errors] = count
        self.errors: Dict[int, Dict[int, int]] = defaultdict(returns_defaultdict_int)
        self.adjacent_bases = {"A": 0, "C": 0, "G": 0, "T": 0, "": 0}
        self._remove_prefix = adapter.remove_before

    def __repr__(self):
        errors = {k: dict(v) for k, v in self.errors.items()}
        return f"EndStatistics(sequence={self.sequence!r}, errors={errors}, adjacent_bases={self.adjacent_bases})"

    def __iadd__(self, other: Any):
        if not isinstance(other, self.__class__):
            raise ValueError("Cannot compare")
        if (
            self.max_error_rate != other.max_error_rate
            or self.sequence != other.sequence
            or self.effective_length != other.effective_length
            or self.indels != other.indels
        ):
            raise RuntimeError("Incompatible EndStatistics, cannot be added")
        for base in ("A", "C", "G", "T", ""):
            self.adjacent_bases[base] += other.adjacent_bases[base]
        for length, error_dict in other.errors.items():
            for errors in error_dict:
                self.errors[length][errors] += other.errors[length][errors]
        return self

    @property
    def lengths(self):
        return {length: sum(errors.values()) for length, errors in self.errors.items()}


def _count_adjacent(end: EndStatistics, base: str) -> None:
    """adjacent_bases[base] += 1, anything that is not A/C/G/T goes to "" (adapters.py:195-199)"""
    end.adjacent_bases[base if base in ("A", "C", "G", "T") else ""] += 1


class AdapterStatistics(ABC):
    reverse_complemented: int = 0
    name: str
    adapter: "Adapter"

    @abstractmethod
    def __iadd__(self, other):
        pass

    @abstractmethod
    def end_statistics(self) -> Tuple[Optional[EndStatistics], Optional[EndStatistics]]:
        pass

    @abstractmethod
    def add_match(self, match) -> None:
        pass


class SingleAdapterStatistics(AdapterStatistics, ABC):
    def __init__(self, adapter: "SingleAdapter"):
        self.name = adapter.name
        self.adapter = adapter
        self.end = EndStatistics(adapter)

    def add_match(self, match: "SingleMatch"):
        self.end.errors[match.removed_sequence_length()][match.errors] += 1

    def __repr__(self):
        return f"SingleAdapterStatistics(name={self.name}, end={self.end})"

    def __iadd__(self, other: "SingleAdapterStatistics"):
        if not isinstance(other, self.__class__):
            raise ValueError("Cannot iadd")
        self.end += other.end
        self.reverse_complemented += other.reverse_complemented
        return self


class FrontAdapterStatistics(SingleAdapterStatistics):
    def end_statistics(self):
        return self.end, None


class BackAdapterStatistics(SingleAdapterStatistics):
    def add_match(self, match: "SingleMatch"):
        super().add_match(match)
        _count_adjacent(self.end, match.adjacent_base())

    def end_statistics(self):
        return None, self.end


class LinkedAdapterStatistics(AdapterStatistics):
    def __init__(self, adapter: "LinkedAdapter", front: "SingleAdapter", back: "SingleAdapter"):
        self.name = adapter.name
        self.adapter = adapter
        self.front = EndStatistics(front)
        self.back = EndStatistics(back)
        self.reverse_complemented = 0

    def __repr__(self):
        return f"LinkedAdapterStatistics(name={self.name}, front={self.front}, back={self.back})"

    def __iadd__(self, other: "LinkedAdapterStatistics"):
        if not isinstance(other, self.__class__):
            raise ValueError("Cannot iadd")
        self.front += other.front
        self.back += other.back
        self.reverse_complemented += other.reverse_complemented
        return self

    def add_match(self, match: "LinkedMatch"):
        if match.front_match:
            self.front.errors[match.front_match.removed_sequence_length()][match.front_match.errors] += 1
        if match.back_match:
            self.back.errors[match.back_match.removed_sequence_length()][match.back_match.errors] += 1
            _count_adjacent(self.back, match.back_match.adjacent_base())

    def end_statistics(self):
        return self.front, self.back


class AnywhereAdapterStatistics(AdapterStatistics):
    def __init__(self, adapter: "AnywhereAdapter"):
        self.name = adapter.name
        self.adapter = adapter
        self.front = EndStatistics(adapter)
        self.back = EndStatistics(adapter)
        self.reverse_complemented = 0

    def __repr__(self):
        return f"AnywhereAdapterStatistics(name={self.name}, front={self.front}, back={self.back})"

    def __iadd__(self, other: "AnywhereAdapterStatistics"):
        if not isinstance(other, AnywhereAdapterStatistics):
            raise ValueError("Cannot add")
        self.front += other.front
        self.back += other.back
        self.reverse_complemented += other.reverse_complemented
        return self

    def add_match(self, match: "SingleMatch") -> None:
        if isinstance(match, RemoveBeforeMatch):
            self.front.errors[match.removed_sequence_length()][match.errors] += 1
        else:
            self.back.errors[match.removed_sequence_length()][match.errors] += 1
            _count_adjacent(self.back, match.adjacent_base())

    def end_statistics(self):
        return self.front, self.back


# ---------------------------------------------------------------------------------------------
# Matches (adapters.py:292-493, 1092-1178)
# ---------------------------------------------------------------------------------------------


class Match(ABC):
    adapter: "Adapter"

    @abstractmethod
    def remainder_interval(self) -> Tuple[int, int]:
        pass

    @abstractmethod
    def retained_adapter_interval(self) -> Tuple[int, int]:
        pass

    @abstractmethod
    def get_info_records(self, read) -> List[List]:
        pass

    @abstractmethod
    def trimmed(self, read):
        pass

    @abstractmethod
    def match_sequence(self):
        pass


class SingleMatch(Match, ABC):
    """One adapter matched to one string; coordinates as in adapters.py:334-356."""

    __slots__ = ["astart", "astop", "rstart", "rstop", "score", "errors", "adapter", "sequence", "length"]

    def __init__(self, astart, astop, rstart, rstop, score, errors, adapter, sequence):
        self.astart: int = astart
        self.astop: int = astop
        self.rstart: int = rstart
        self.rstop: int = rstop
        self.score: int = score
        self.errors: int = errors
        self.adapter = adapter
        self.sequence = sequence
        self.length: int = astop - astart

    def __repr__(self):
        return (
            f"{self.__class__.__name__}(astart={self.astart}, astop={self.astop}, "
            f"rstart={self.rstart}, rstop={self.rstop}, score={self.score}, errors={self.errors})"
        )

    def __eq__(self, other) -> bool:
        return (
            other.__class__ is self.__class__
            and self.astart == other.astart
            and self.astop == other.astop
            and self.rstart == other.rstart
            and self.rstop == other.rstop
            and self.score == other.score
            and self.errors == other.errors
            and self.adapter is other.adapter
            and self.sequence == other.sequence
        )

    def wildcards(self, wildcard_char: str = "N") -> str:
        """Characters of the read that were matched by wildcards in the adapter (adapters.py:378-393)."""
        return "".join(
            self.sequence[self.rstart + i]
            for i in range(self.length)
            if self.adapter.sequence[self.astart + i] == wildcard_char and self.rstart + i < len(self.sequence)
        )

    def get_info_records(self, read) -> List[List]:
        seq, qualities = read.sequence, read.qualities
        info = [
            "", self.errors, self.rstart, self.rstop,
            seq[0 : self.rstart], seq[self.rstart : self.rstop], seq[self.rstop :], self.adapter.name,
        ]
        if qualities:
            info += [qualities[0 : self.rstart], qualities[self.rstart : self.rstop], qualities[self.rstop :]]
        else:
            info += ["", "", ""]
        return [info]

    def match_sequence(self):
        return self.sequence[self.rstart : self.rstop]

    @abstractmethod
    def removed_sequence_length(self) -> int:
        pass


class RemoveBeforeMatch(SingleMatch):
    """A match that removes sequence before the match (5' adapters)."""

    def rest(self) -> str:
        return self.sequence[: self.rstart]

    def remainder_interval(self) -> Tuple[int, int]:
        return self.rstop, len(self.sequence)

    def retained_adapter_interval(self) -> Tuple[int, int]:
        return self.rstart, len(self.sequence)

    def trim_slice(self):
        return slice(self.rstop, None)

    def trimmed(self, read):
        return read[self.rstop :]

    def removed_sequence_length(self) -> int:
        return self.rstop


class RemoveAfterMatch(SingleMatch):
    """A match that removes sequence after the match (3' adapters)."""

    def rest(self) -> str:
        return self.sequence[self.rstop :]

    def remainder_interval(self) -> Tuple[int, int]:
        return 0, self.rstart

    def retained_adapter_interval(self) -> Tuple[int, int]:
        return 0, self.rstop

    def trim_slice(self):
        return slice(None, self.rstart)

    def trimmed(self, read):
        return read[: self.rstart]

    def adjacent_base(self) -> str:
        return self.sequence[self.rstart - 1 : self.rstart]

    def removed_sequence_length(self) -> int:
        return len(self.sequence) - self.rstart


def remainder(matches: Sequence[Match]) -> Tuple[int, int]:
    """Interval of the read that remains after applying all matches in order (adapters.py:remainder)."""
    if not matches:
        raise ValueError("matches must not be empty")
    start = 0
    for match in matches:
        match_start, match_stop = match.remainder_interval()
        start += match_start
    length = match_stop - match_start
    return (start, start + length)


class LinkedMatch(Match):
    """Match of a LinkedAdapter (adapters.py:1092-1178)."""

    def __init__(self, front_match, back_match, adapter: "LinkedAdapter"):
        assert front_match is not None or back_match is not None
        self.front_match = front_match
        self.back_match = back_match
        self.adapter = adapter

    def __repr__(self):
        return "<LinkedMatch(front_match={!r}, back_match={}, adapter={})>".format(
            self.front_match, self.back_match, self.adapter
        )

    @property
    def score(self):
        return (self.front_match.score if self.front_match is not None else 0) + (
            self.back_match.score if self.back_match is not None else 0
        )

    @property
    def errors(self):
        return (self.front_match.errors if self.front_match is not None else 0) + (
            self.back_match.errors if self.back_match is not None else 0
        )

    def trimmed(self, read):
        if self.front_match:
            read = self.front_match.trimmed(read)
        if self.back_match:
            read = self.back_match.trimmed(read)
        return read

    def remainder_interval(self) -> Tuple[int, int]:
        return remainder([m for m in (self.front_match, self.back_match) if m is not None])

    def retained_adapter_interval(self) -> Tuple[int, int]:
        if self.front_match:
            start = self.front_match.rstart
            offset = self.front_match.rstop
        else:
            start = offset = 0
        if self.back_match:
            end = self.back_match.rstop + offset
        else:
            end = len(self.front_match.sequence)
        return start, end

    def get_info_records(self, read) -> List[List]:
        records = []
        for match, namesuffix in ((self.front_match, ";1"), (self.back_match, ";2")):
            if match is None:
                continue
            record = match.get_info_records(read)[0]
            record[7] = ("none" if self.adapter.name is None else self.adapter.name) + namesuffix
            records.append(record)
            read = match.trimmed(read)
        return records

    def match_sequence(self):
        return (
            (self.front_match.match_sequence() if self.front_match else "")
            + ","
            + (self.back_match.match_sequence() if self.back_match else "")
        )


# ---------------------------------------------------------------------------------------------
# Adapters
# ---------------------------------------------------------------------------------------------


def _generate_adapter_name(_start=[1]) -> str:
    name = str(_start[0])
    _start[0] += 1
    return name


class Matchable(ABC):
    """Something that has a match_to() method -- and, here, a match_to_batch()."""

    def __init__(self, name: Optional[str], *args, **kwargs):
        self.name = name

    @abstractmethod
    def match_to(self, sequence: str):
        pass

    # -- batched dispatch ---------------------------------------------------------------------
    _device_set = None

    def _flatten(self) -> Tuple[List["SingleAdapter"], List[tuple], List["Matchable"]]:
        """(single adapters in device order, group tuples, the Matchable of every group)"""
        raise NotImplementedError

    def _flatten_indexes(self) -> List[dict]:
        """AdapterSetSpec index dicts for the CG_GROUP_INDEXED groups of _flatten() (in a0 order)."""
        return []

    def adapter_set(self) -> "_lib.AdapterSet":
        """Compile + upload this Matchable's tables (once per process/context)."""
        ctx = _lib.default_context()
        cached = self._device_set
        if cached is None or cached[0].ctx is not ctx:
            singles, groups, owners = self._flatten()
            spec = _lib.AdapterSetSpec([s.descriptor() for s in singles], groups, self._flatten_indexes())
            cached = (_lib.AdapterSet(spec, ctx), singles, owners)
            self._device_set = cached
        return cached[0]

    def matches_from_records(self, records: np.ndarray, sequence: str):
        """Turn the device records of ONE read and ONE round (shape (slots,)) into a Match or None."""
        _, singles, owners = self._device_set
        first, second = records[0], (records[1] if len(records) > 1 else None)
        present = first if first["adapter"] >= 0 else second
        if present is None or present["adapter"] < 0:
            return None
        owner = owners[int(present["info"]) & 255]

        def single(rec, seq):
            adapter = singles[int(rec["adapter"])]
            cls = RemoveAfterMatch if (int(rec["info"]) >> 8) & 1 else RemoveBeforeMatch
            return cls(int(rec["astart"]), int(rec["astop"]), int(rec["rstart"]), int(rec["rstop"]),
                       int(rec["score"]), int(rec["errors"]), adapter=adapter, sequence=seq)

        if isinstance(owner, LinkedAdapter):
            front = single(first, sequence) if first["adapter"] >= 0 else None
            rest = sequence[front.trim_slice()] if front is not None else sequence
            back = single(second, rest) if second is not None and second["adapter"] >= 0 else None
            return LinkedMatch(front, back, owner)
        return single(first, sequence)

    def match_to_batch(self, sequences: Sequence[str]) -> List[Optional[Match]]:
        """match_to() for every sequence with one fused kernel launch."""
        data, offsets = _lib.pack_strings(sequences)
        aset = self.adapter_set()
        records, _ = aset.process(data, offsets)
        return [self.matches_from_records(records[i, 0], seq) for i, seq in enumerate(sequences)]

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_device_set", None)
        return state


class Adapter(Matchable, ABC):
    description = "adapter with one component"

    @abstractmethod
    def spec(self) -> str:
        """Return string representation of this adapter"""

    @abstractmethod
    def create_statistics(self) -> AdapterStatistics:
        pass

    @abstractmethod
    def descriptive_identifier(self) -> str:
        pass

    @abstractmethod
    def enable_debug(self) -> None:
        pass


class SingleAdapter(Adapter, ABC):
    """
    One adapter characterised by sequence, error rate, type ... (adapters.py:533-681; the
    constructor arguments have the reference's meaning).
    """

    allows_partial_matches: bool = True
    remove_before: bool = False          # which Match class wraps a hit
    _remove_mode = _lib.CG_REMOVE_AFTER  # CG_REMOVE_* handed to the device
    _reverse_read = False

    def __init__(
        self,
        sequence: str,
        max_errors: float = 0.1,
        min_overlap: int = 3,
        read_wildcards: bool = False,
        adapter_wildcards: bool = True,
        name: Optional[str] = None,
        indels: bool = True,
    ):
        self.name: str = _generate_adapter_name() if name is None else name
        super().__init__(self.name)
        self._debug: bool = False
        self.sequence: str = sequence.upper().replace("U", "T").replace("I", "N")
        if not self.sequence:
            raise ValueError("Adapter sequence is empty")
        if max_errors >= 1 and self.sequence.count("N") != len(self.sequence):
            max_errors /= len(self.sequence) - self.sequence.count("N")
        self.max_error_rate: float = max_errors
        self.min_overlap: int = min(min_overlap, len(self.sequence))
        iupac = frozenset("ABCDGHKMNRSTUVWXY")
        if adapter_wildcards and not set(self.sequence) <= iupac:
            for c in self.sequence:
                if c not in iupac:
                    raise InvalidCharacter(
                        f"Character '{c}' in adapter sequence '{self.sequence}' is "
                        f"not a valid IUPAC code. Use only characters 'ABCDGHIKMNRSTUVWXY'."
                    )
        # non-wildcard matching is used when only ACGT occurs (adapters.py:592-595)
        self.adapter_wildcards: bool = adapter_wildcards and not set(self.sequence) <= set("ACGT")
        self.read_wildcards: bool = read_wildcards
        self.indels: bool = indels
        self.aligner = self._aligner()
        self.kmer_finder = self._kmer_finder()

    def _make_aligner(self, sequence: str, flags: int) -> Aligner:
        indel_cost = 1 if self.indels else 100000      # adapters.py:605
        return Aligner(
            sequence,
            self.max_error_rate,
            flags=flags,
            wildcard_ref=self.adapter_wildcards,
            wildcard_query=self.read_wildcards,
            indel_cost=indel_cost,
            min_overlap=self.min_overlap,
        )

    def _make_kmer_finder(self, sequence, back_adapter, front_adapter, internal=True):
        positions_and_kmers = create_positions_and_kmers(
            sequence, self.min_overlap, self.max_error_rate, back_adapter, front_adapter, internal
        )
        try:
            return KmerFinder(positions_and_kmers, self.adapter_wildcards, self.read_wildcards)
        except ValueError:
            return MockKmerFinder()                      # k-mers too long (adapters.py:637-639)

    def __repr__(self):
        return (
            "<{cls}(name={name!r}, sequence={sequence!r}, max_error_rate={max_error_rate}, "
            "min_overlap={min_overlap}, read_wildcards={read_wildcards}, "
            "adapter_wildcards={adapter_wildcards}, indels={indels})>".format(
                cls=self.__class__.__name__, name=self.name, sequence=self.sequence,
                max_error_rate=self.max_error_rate, min_overlap=self.min_overlap,
                read_wildcards=self.read_wildcards, adapter_wildcards=self.adapter_wildcards,
                indels=self.indels,
            )
        )

    @property
    def effective_length(self) -> int:
        return self.aligner.effective_length

    def enable_debug(self) -> None:
        self._debug = True

    @abstractmethod
    def _aligner(self):
        pass

    @abstractmethod
    def _kmer_finder(self):
        pass

    def __len__(self) -> int:
        return len(self.sequence)

    # -- device description -------------------------------------------------------------------
    def descriptor(self) -> dict:
        """The cg_adapter_desc of this adapter: its aligner's arguments + prefilter tables."""
        d = self.aligner._descriptor()
        d["reverse_read"] = self._reverse_read
        d["remove"] = self._remove_mode
        if isinstance(self.kmer_finder, KmerFinder):
            entries, masks = self.kmer_finder.tables
            d["kmer_entries"] = entries
            d["kmer_masks"] = masks
        return d

    def _flatten(self):
        return [self], [(_lib.CG_GROUP_SINGLE, 0, -1, 0, 0)], [self]

    def match_to(self, sequence: str):
        """
        Attempt to match this adapter to the given read; a Match, or None if nothing satisfies
        minimum overlap and maximum error rate.
        """
        return self.match_to_batch([sequence])[0]


class FrontAdapter(SingleAdapter):
    """A 5' adapter (adapters.py:684-730)"""

    description = "regular 5'"
    remove_before = True
    _remove_mode = _lib.CG_REMOVE_BEFORE

    def __init__(self, *args, **kwargs):
        self._force_anywhere = kwargs.pop("force_anywhere", False)
        super().__init__(*args, **kwargs)

    def descriptive_identifier(self) -> str:
        return "regular_five_prime"

    def _aligner(self) -> Aligner:
        return self._make_aligner(
            self.sequence, Where.ANYWHERE.value if self._force_anywhere else Where.FRONT.value
        )

    def _kmer_finder(self):
        return self._make_kmer_finder(self.sequence, back_adapter=self._force_anywhere, front_adapter=True)

    def spec(self) -> str:
        return f"{self.sequence}..."

    def create_statistics(self) -> FrontAdapterStatistics:
        return FrontAdapterStatistics(self)


class RightmostFrontAdapter(FrontAdapter):
    """A 5' adapter that prefers rightmost matches (adapters.py:733-789)"""

    description = "rightmost 5'"
    _reverse_read = True

    def descriptive_identifier(self) -> str:
        return "rightmost_five_prime"

    def _aligner(self) -> Aligner:
        return self._make_aligner(
            self.sequence[::-1], Where.ANYWHERE.value if self._force_anywhere else Where.BACK.value
        )

    def _kmer_finder(self):
        return self._make_kmer_finder(
            self.sequence[::-1], back_adapter=True, front_adapter=self._force_anywhere
        )

    def spec(self) -> str:
        return f"{self.sequence}...;rightmost"


class BackAdapter(SingleAdapter):
    """A 3' adapter (adapters.py:792-838)"""

    description = "regular 3'"

    def __init__(self, *args, **kwargs):
        self._force_anywhere = kwargs.pop("force_anywhere", False)
        super().__init__(*args, **kwargs)

    def descriptive_identifier(self) -> str:
        return "regular_three_prime"

    def _aligner(self):
        return self._make_aligner(
            self.sequence, Where.ANYWHERE.value if self._force_anywhere else Where.BACK.value
        )

    def _kmer_finder(self):
        return self._make_kmer_finder(self.sequence, back_adapter=True, front_adapter=self._force_anywhere)

    def spec(self) -> str:
        return f"{self.sequence}"

    def create_statistics(self) -> BackAdapterStatistics:
        return BackAdapterStatistics(self)


class RightmostBackAdapter(BackAdapter):
    """A 3' adapter that prefers rightmost matches (adapters.py:841-893)"""

    description = "rightmost 3'"
    _reverse_read = True

    def descriptive_identifier(self) -> str:
        return "rightmost_three_prime"

    def _aligner(self) -> Aligner:
        return self._make_aligner(
            self.sequence[::-1], Where.ANYWHERE.value if self._force_anywhere else Where.FRONT.value
        )

    def _kmer_finder(self):
        return self._make_kmer_finder(
            self.sequence[::-1], back_adapter=self._force_anywhere, front_adapter=True
        )

    def spec(self) -> str:
        return f"{self.sequence};rightmost"


class AnywhereAdapter(SingleAdapter):
    """
    An adapter that can be 5' or 3': a match that involves the first base of the read is taken
    to be 5', otherwise 3' (adapters.py:896-941).
    """

    description = "variable 5'/3'"
    _remove_mode = _lib.CG_REMOVE_AUTO

    def descriptive_identifier(self) -> str:
        return "anywhere"

    def _aligner(self):
        return self._make_aligner(self.sequence, Where.ANYWHERE.value)

    def _kmer_finder(self):
        return self._make_kmer_finder(self.sequence, back_adapter=True, front_adapter=True)

    def spec(self) -> str:
        return f"...{self.sequence}..."

    def create_statistics(self) -> AnywhereAdapterStatistics:
        return AnywhereAdapterStatistics(self)


class NonInternalFrontAdapter(FrontAdapter):
    """A non-internal 5' adapter (adapters.py:944-978)"""

    description = "non-internal 5'"

    def descriptive_identifier(self) -> str:
        return "noninternal_five_prime"

    def _aligner(self):
        return self._make_aligner(self.sequence, Where.FRONT_NOT_INTERNAL.value)

    def _kmer_finder(self):
        return self._make_kmer_finder(
            self.sequence, front_adapter=True, back_adapter=self._force_anywhere, internal=False
        )

    def spec(self) -> str:
        return f"X{self.sequence}..."


class NonInternalBackAdapter(BackAdapter):
    """A non-internal 3' adapter (adapters.py:981-1015)"""

    description = "non-internal 3'"

    def descriptive_identifier(self) -> str:
        return "noninternal_three_prime"

    def _aligner(self):
        return self._make_aligner(self.sequence, Where.BACK_NOT_INTERNAL.value)

    def _kmer_finder(self):
        return self._make_kmer_finder(
            self.sequence, back_adapter=True, front_adapter=self._force_anywhere, internal=False
        )

    def spec(self) -> str:
        return f"{self.sequence}X"


class PrefixAdapter(NonInternalFrontAdapter):
    """An anchored 5' adapter (adapters.py:1018-1052)"""

    description = "anchored 5'"
    allows_partial_matches = False

    def __init__(self, sequence: str, *args, **kwargs):
        kwargs["min_overlap"] = len(sequence)
        super().__init__(sequence, *args, **kwargs)

    def descriptive_identifier(self) -> str:
        return "anchored_five_prime"

    def _aligner(self):
        if not self.indels:
            return PrefixComparer(
                self.sequence, self.max_error_rate, wildcard_ref=self.adapter_wildcards,
                wildcard_query=self.read_wildcards, min_overlap=self.min_overlap,
            )
        return self._make_aligner(self.sequence, Where.PREFIX.value)

    def _kmer_finder(self):
        if isinstance(self.aligner, PrefixComparer):
            return MockKmerFinder()
        return super()._kmer_finder()

    def spec(self) -> str:
        return f"^{self.sequence}..."


class SuffixAdapter(NonInternalBackAdapter):
    """An anchored 3' adapter (adapters.py:1055-1089)"""

    description = "anchored 3'"
    allows_partial_matches = False

    def __init__(self, sequence: str, *args, **kwargs):
        kwargs["min_overlap"] = len(sequence)
        super().__init__(sequence, *args, **kwargs)

    def descriptive_identifier(self) -> str:
        return "anchored_three_prime"

    def _aligner(self):
        if not self.indels:
            return SuffixComparer(
                self.sequence, self.max_error_rate, wildcard_ref=self.adapter_wildcards,
                wildcard_query=self.read_wildcards, min_overlap=self.min_overlap,
            )
        return self._make_aligner(self.sequence, Where.SUFFIX.value)

    def _kmer_finder(self):
        if isinstance(self.aligner, SuffixComparer):
            return MockKmerFinder()
        return super()._kmer_finder()

    def spec(self) -> str:
        return f"{self.sequence}$"


class LinkedAdapter(Adapter):
    """A 5' adapter combined with a 3' adapter (adapters.py:1181-1243)"""

    description = "linked"

    def __init__(self, front_adapter, back_adapter, front_required, back_required, name):
        super().__init__(name)
        self.front_required = front_required
        self.back_required = back_required
        self.where = "linked"
        self.name: str = _generate_adapter_name() if name is None else name
        self.front_adapter = front_adapter
        self.front_adapter.name = self.name
        self.back_adapter = back_adapter

    def __repr__(self):
        return f"{self.__class__.__name__}(front_adapter={self.front_adapter}, back_adapter={self.back_adapter})"

    def descriptive_identifier(self) -> str:
        return "linked"

    def enable_debug(self):
        self.front_adapter.enable_debug()
        self.back_adapter.enable_debug()

    def _flatten(self):
        return (
            [self.front_adapter, self.back_adapter],
            [(_lib.CG_GROUP_LINKED, 0, 1, int(bool(self.front_required)), int(bool(self.back_required)))],
            [self],
        )

    def match_to(self, sequence: str) -> Optional[LinkedMatch]:
        """Match the two linked adapters against a string (adapters.py:1215-1227)"""
        return self.match_to_batch([sequence])[0]

    def create_statistics(self) -> LinkedAdapterStatistics:
        return LinkedAdapterStatistics(self, front=self.front_adapter, back=self.back_adapter)

    @property
    def sequence(self):
        return self.front_adapter.sequence + "..." + self.back_adapter.sequence

    @property
    def remove(self):
        return None

    def spec(self) -> str:
        return f"{self.front_adapter.spec()}...{self.back_adapter.spec()}"


class MultipleAdapters(Matchable):
    """
    Several adapters at once: the best match wins -- highest score, then fewest errors, then the
    adapter listed first (adapters.py:1246-1286).  On the device all adapters are evaluated for
    a read inside one kernel, sharing the staged read bytes.
    """

    def __init__(self, adapters: Sequence[Matchable]):
        super().__init__(name="multiple_adapters")
        self._adapters = adapters

    def enable_debug(self):
        for a in self._adapters:
            a.enable_debug()

    def __getitem__(self, item):
        return self._adapters[item]

    def __len__(self):
        return len(self._adapters)

    def _flatten(self):
        singles: List[SingleAdapter] = []
        groups: List[tuple] = []
        owners: List[Matchable] = []
        n_indexes = 0
        for adapter in self._adapters:
            sub_singles, sub_groups, sub_owners = adapter._flatten()
            base = len(singles)
            singles.extend(sub_singles)
            for typ, a0, a1, freq, breq in sub_groups:
                if typ == _lib.CG_GROUP_INDEXED:
                    groups.append((typ, a0 + n_indexes, -1, freq, breq))
                else:
                    groups.append((typ, a0 + base, a1 + base if a1 >= 0 else -1, freq, breq))
            n_indexes += sum(1 for g in sub_groups if g[0] == _lib.CG_GROUP_INDEXED)
            owners.extend(sub_owners)
        return singles, groups, owners

    def _flatten_indexes(self):
        out: List[dict] = []
        base = 0
        for adapter in self._adapters:
            for ix in adapter._flatten_indexes():
                ix = dict(ix)
                ix["adapter"] = [a + base for a in ix["adapter"]]
                out.append(ix)
            base += len(adapter._flatten()[0])
        return out

    def match_to(self, sequence: str) -> Optional[Match]:
        """Find the adapter that best matches the sequence; a Match or None."""
        return self.match_to_batch([sequence])[0]


class AdapterIndex:
    """
    Index of multiple anchored adapters of the same type (adapters.py:1289-1551): every string
    within the allowed edit/Hamming distance of every adapter is a dictionary key, so matching is
    one lookup of the read's prefix (suffix) per distinct key length.

    The dictionary is built here exactly as the reference builds it (including the removal of
    ambiguous keys) and handed to the library, which keeps it as a hash table in HBM; the lookups
    for a chunk of reads happen inside the fused kernel (CG_GROUP_INDEXED).

    Restrictions of the reference (adapters.py:1366-1378): no wildcards, at most 3 errors.
    Additional restrictions of this implementation (``is_acceptable`` is False, so callers fall
    back to ``MultipleAdapters`` as they do for the reference's own restrictions): adapter
    alphabet A/C/G/T, length <= 32, fewer errors than characters.
    """

    AdapterIndexDict = Dict[str, Tuple["SingleAdapter", int, int]]

    def __init__(self, adapters, prefix: bool):
        """All given adapters must be of the same type"""
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
        return f"{self.__class__.__name__}(adapters={self._adapters!r})"

    @classmethod
    def _accept(cls, adapter: "SingleAdapter", prefix: bool):
        """Raise a ValueError if the adapter is not acceptable (adapters.py:1366-1378)"""
        if prefix and not isinstance(adapter, PrefixAdapter):
            raise ValueError("Only 5' anchored adapters are allowed")
        elif not prefix and not isinstance(adapter, SuffixAdapter):
            raise ValueError("Only 3' anchored adapters are allowed")
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
    def is_acceptable(cls, adapter: "SingleAdapter", prefix: bool):
        """Whether this adapter can be used in an index (adapters.py:1380-1392)"""
        try:
            cls._accept(adapter, prefix)
        except ValueError:
            return False
        return True

    def _make_index(self) -> Tuple[List[int], "AdapterIndexDict", int]:
        """adapters.py:1394-1472"""
        index: Dict[str, Tuple[SingleAdapter, int, int]] = dict()
        lengths = set()
        ambiguous = {}
        for adapter in self._adapters:
            sequence = adapter.sequence
            k = int(adapter.max_error_rate * len(sequence))
            if adapter.indels:
                for s, errors, matches in edit_environment(sequence, k):
                    if s in index:
                        other_adapter, other_errors, other_matches = index[s]
                        if matches < other_matches:
                            continue
                        if other_matches == matches and s not in ambiguous:
                            ambiguous[s] = (adapter, other_adapter, k, matches)
                    index[s] = (adapter, errors, matches)
                    lengths.add(len(s))
            else:
                n = len(sequence)
                for errors in range(k + 1):
                    matches = n - errors
                    for s in hamming_sphere(sequence, errors):
                        if s in index:
                            other_adapter, other_errors, other_matches = index[s]
                            if matches < other_matches:
                                continue
                            if other_matches == matches and s not in ambiguous:
                                ambiguous[s] = (adapter, other_adapter, k, matches)
                        index[s] = (adapter, errors, matches)
                lengths.add(n)
        for s in ambiguous:
            del index[s]
        return sorted(lengths, reverse=True), index, len(ambiguous)

    def descriptor(self) -> dict:
        """The cg_index_desc content: keys and (adapter number, errors, matches) per key."""
        number = {id(a): i for i, a in enumerate(self._adapters)}
        keys = list(self._index)
        return {
            "prefix": self._prefix,
            "lengths": list(self._lengths),
            "keys": keys,
            "adapter": [number[id(self._index[k][0])] for k in keys],
            "errors": [self._index[k][1] for k in keys],
            "matches": [self._index[k][2] for k in keys],
        }


class _IndexedAdapters(Matchable):
    _is_prefix = True

    def __init__(self, adapters, name):
        super().__init__(name=name)
        self._index = AdapterIndex(adapters, prefix=self._is_prefix)

    def _flatten(self):
        return list(self._index._adapters), [(_lib.CG_GROUP_INDEXED, 0, -1, 0, 0)], [self]

    def _flatten_indexes(self):
        return [self._index.descriptor()]

    def match_to(self, sequence: str):
        """AdapterIndex.match_to (adapters.py:1474-1551) as a batch of one on the device."""
        return self.match_to_batch([sequence])[0]


class IndexedPrefixAdapters(_IndexedAdapters):
    """adapters.py:1554-1561"""

    _is_prefix = True

    def __init__(self, adapters):
        super().__init__(adapters, name="indexed_prefix_adapters")


class IndexedSuffixAdapters(_IndexedAdapters):
    """adapters.py:1564-1571"""

    _is_prefix = False

    def __init__(self, adapters):
        super().__init__(adapters, name="indexed_suffix_adapters")
