"""
Bit-sliced formulation of KmerFinder.kmers_present -- a host prototype of the scan planned for the next round
(DESIGN.md section 7).  Not on the product path.

Instead of one shift-and step per character (scan_core, cg_core.cuh), the read becomes a few position bitmaps
("planes", one per pattern character class: bit p set iff read[p] matches that pattern character), and a k-mer
occurs at p iff  AND_i (plane[kmer[i]] >> i)  has bit p set.  A KmerFinder entry (start, stop, kmers) then passes
iff some k-mer starts at a position p with start <= p and p + len(kmer) <= stop (window normalised as in
_kmer_finder.pyx:188-204).  Python integers stand in for the multi-word bitmaps of a 150-base read (5 x 32 bits
per plane on the device; shifts become funnel shifts, ANDs stay ANDs).

``kmers_present(positions_and_kmers, ref_wildcards, query_wildcards, sequence)`` must equal the reference's verdict
(tests/test_host_logic.py fuzzes it against the oracle); ``chunk_end_positions`` gives the exact end positions of
the locator chunks, which the shift-and scan only delivers per 16-character group (the plan kernel re-scans the
hit groups today; with bit-slicing the positions fall out of the same ANDs).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cutadapt_b200._match_tables import matches_lookup  # noqa: E402


def planes(sequence: str, pattern_chars, ref_wildcards: bool, query_wildcards: bool) -> dict:
    """{pattern character: bitmap of the positions of `sequence` it matches}"""
    lookup = matches_lookup(ref_wildcards, query_wildcards)
    data = sequence.encode("ascii")
    out = {}
    for ch in pattern_chars:
        accepted = lookup[ord(ch)]
        bits = 0
        for p, c in enumerate(data):
            if c in accepted:
                bits |= 1 << p
        out[ch] = bits
    return out


def occurrences(kmer: str, pl: dict, n: int) -> int:
    """bitmap of the start positions of `kmer` (the AND of its shifted planes)"""
    occ = (1 << max(n - len(kmer) + 1, 0)) - 1        # only starts that leave room for the whole k-mer
    for i, ch in enumerate(kmer):
        occ &= pl[ch] >> i
        if not occ:
            break
    return occ


def window(start, stop, n):
    """(first, last) positions of the searched window, or None if the entry is skipped (_kmer_finder.pyx:188-204)"""
    stop = 0 if stop is None else stop
    if start < 0:
        start = max(n + start, 0)
    elif start > n:
        return None
    if stop < 0:
        stop = n + stop
        if stop <= 0:
            return None
    elif stop == 0 or stop > n:
        stop = n
    if stop <= start:
        return None
    return start, stop


def kmers_present(positions_and_kmers, ref_wildcards, query_wildcards, sequence: str) -> bool:
    n = len(sequence)
    chars = sorted({c for _, _, ks in positions_and_kmers for k in ks for c in k})
    pl = planes(sequence, chars, ref_wildcards, query_wildcards)
    for start, stop, kmers in positions_and_kmers:
        w = window(start, stop, n)
        if w is None:
            continue
        lo, hi = w
        for kmer in kmers:
            if not kmer or len(kmer) > hi - lo:
                continue
            in_window = ((1 << (hi - len(kmer) + 1)) - 1) & ~((1 << lo) - 1)     # lo <= p <= hi - len
            if occurrences(kmer, pl, n) & in_window:
                return True
    return False


def chunk_end_positions(chunks, ref_wildcards, query_wildcards, sequence: str) -> dict:
    """{chunk: bitmap of the positions where an occurrence of the chunk ends}"""
    n = len(sequence)
    pl = planes(sequence, sorted({c for k in chunks for c in k}), ref_wildcards, query_wildcards)
    return {k: occurrences(k, pl, n) << (len(k) - 1) for k in chunks if k}


def word_operations(positions_and_kmers, read_len=150, word=32) -> int:
    """funnel-shift + AND pairs per read for the matching step (planes not included)"""
    words = (read_len + word - 1) // word
    return sum(len(k) for _, _, ks in positions_and_kmers for k in ks) * words


if __name__ == "__main__":
    from cutadapt_b200.kmer_heuristic import create_positions_and_kmers

    pk = create_positions_and_kmers("AGATCGGAAGAGC", 3, 0.1, True, False, True)
    print("BASELINE configs[1]: search sets", pk)
    print("matching step: %d shift/AND pairs per 150-base read = %.2f per character (+ the planes)"
          % (word_operations(pk), word_operations(pk) / 150))
