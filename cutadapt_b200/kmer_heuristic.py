"""
Pigeonhole decomposition of an adapter into k-mer search sets (host side, once per adapter).

Same interface and results as the reference's ``cutadapt.kmer_heuristic``
(src/cutadapt/kmer_heuristic.py:6-164): ``create_positions_and_kmers`` returns a list of
``(start, stop, [kmers])`` such that if none of the k-mers occurs in its window, no alignment
within the error rate exists.  The lists feed ``KmerFinder``; on the device the windows are
evaluated by the fused prefilter stage of the trimming kernel.

The order of k-mers *within* one search set is unspecified in the reference too (it builds
Python sets, kmer_heuristic.py:21-26); only the boolean outcome matters.  Here the order is
made deterministic (first occurrence in the adapter).
"""
from typing import Dict, List, Optional, Set, Tuple

SearchSet = Tuple[int, Optional[int], List[str]]


def kmer_chunks(sequence: str, chunks: int) -> Set[str]:
    """The pieces as a set, the reference's return type (kmer_heuristic.py:6-22); the tables are built from the
    ordered form, ``_ordered_chunks``."""
    return set(_ordered_chunks(sequence, chunks))


def create_back_overlap_searchsets(adapter: str, min_overlap: int, error_rate: float):
    """The reference's name and return type for ``_back_overlap_search_sets`` (kmer_heuristic.py:87-117)."""
    return [(start, stop, set(kmers)) for start, stop, kmers in _back_overlap_search_sets(adapter, min_overlap, error_rate)]


def minimize_kmer_search_list(kmer_search_list):
    """(kmer, start, stop) triples with every k-mer kept only in its widest window(s) (kmer_heuristic.py:29-84)."""
    merged = _merge_windows([(start, stop, [kmer]) for kmer, start, stop in kmer_search_list])
    return [(kmer, start, stop) for start, stop, kmers in merged for kmer in kmers]


def _ordered_chunks(sequence: str, chunks: int) -> List[str]:
    """
    Split ``sequence`` into ``chunks`` nearly equal pieces, the longer ones first, and
    return the distinct pieces in order of appearance (kmer_heuristic.py:6-22).
    """
    base, extra = divmod(len(sequence), chunks)
    pieces: List[str] = []
    pos = 0
    for index in range(chunks):
        size = base + 1 if index < extra else base
        piece = sequence[pos : pos + size]
        pos += size
        if piece not in pieces:
            pieces.append(piece)
    return pieces


def _back_overlap_search_sets(adapter: str, min_overlap: int, error_rate: float) -> List[SearchSet]:
    """kmer_heuristic.py:87-117"""
    # (allowed errors, longest prefix length with that many allowed errors)
    brackets: List[Tuple[int, int]] = []
    allowed = 0
    for length in range(len(adapter) + 1):
        if int(length * error_rate) > allowed:
            brackets.append((allowed, length - 1))
            allowed += 1
    brackets.append((allowed, len(adapter)))

    search_sets: List[SearchSet] = []
    shortest = min_overlap
    for errors, longest in brackets:
        if shortest > longest:
            continue
        if errors == 0 and shortest < 5:
            # very short exact overlaps are only checked at their exact offset
            for size in range(shortest, 5):
                search_sets.append((-size, None, [adapter[:size]]))
            shortest = 5
        search_sets.append((-longest, None, _ordered_chunks(adapter[:shortest], errors + 1)))
        shortest = longest + 1
    return search_sets


def _merge_windows(search_sets: List[SearchSet]) -> List[SearchSet]:
    """
    A k-mer that is searched in several windows is only searched in the widest one
    (kmer_heuristic.py:29-84).
    """
    windows: Dict[str, List[Tuple[int, Optional[int]]]] = {}
    for start, stop, kmers in search_sets:
        for kmer in kmers:
            windows.setdefault(kmer, []).append((start, stop))
    merged: Dict[Tuple[int, Optional[int]], List[str]] = {}

    def emit(kmer, start, stop):
        merged.setdefault((start, stop), []).append(kmer)

    for kmer, positions in windows.items():
        if len(positions) == 1:
            emit(kmer, *positions[0])
            continue
        if (0, None) in positions:
            emit(kmer, 0, None)
            continue
        anchored_front = [stop for start, stop in positions if start == 0]
        anchored_back = [start for start, stop in positions if stop is None]
        if any(start != 0 and stop is not None for start, stop in positions):
            raise NotImplementedError(
                "Situations with searches starting in the middle have not been considered."
            )
        if anchored_front:
            emit(kmer, 0, max(anchored_front))
        if anchored_back:
            emit(kmer, min(anchored_back), None)
    return [(start, stop, kmers) for (start, stop), kmers in merged.items()]


def create_positions_and_kmers(
    adapter: str,
    min_overlap: int,
    error_rate: float,
    back_adapter: bool,
    front_adapter: bool,
    internal: bool = True,
) -> List[SearchSet]:
    """kmer_heuristic.py:120-164"""
    search_sets: List[SearchSet] = []
    if back_adapter:
        search_sets.extend(_back_overlap_search_sets(adapter, min_overlap, error_rate))
    if front_adapter:
        # mirror image of the back-adapter construction
        for start, _stop, kmers in _back_overlap_search_sets(adapter[::-1], min_overlap, error_rate):
            search_sets.append((0, -start, [kmer[::-1] for kmer in kmers]))
    if internal:
        max_errors = int(len(adapter) * error_rate)
        search_sets.append((0, None, _ordered_chunks(adapter, max_errors + 1)))
    return _merge_windows(search_sets)
