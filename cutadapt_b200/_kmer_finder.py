"""
Mirror of ``cutadapt._kmer_finder`` (src/cutadapt/_kmer_finder.pyx): ``KmerFinder`` with the
reference's constructor, attributes, pickling and error behaviour.  The needle-mask tables are
built here on the host exactly as KmerFinder.__cinit__ does (lines 106-165); the scan itself
(``kmers_present``, lines 170-213, 241-257) runs on the GPU.
"""
import ctypes as C
from typing import List, Sequence

import numpy as np

from . import _lib
from ._match_tables import matches_lookup

MAXIMUM_WORD_SIZE = 64


def build_kmer_tables(positions_and_kmers, ref_wildcards: bool, query_wildcards: bool):
    """
    Pack every (start, stop, kmers) search set into 64-bit shift-and words.

    Returns (entries, masks): entries is a list of (search_start, search_stop, init_mask,
    found_mask) -- ``stop is None`` is encoded as 0 (_kmer_finder.pyx:156-157) -- and masks a
    uint64 array of 128 words per entry, indexed by ASCII code (_kmer_finder.pyx:226-238).
    """
    lookup = matches_lookup(ref_wildcards, query_wildcards)
    entries: List[tuple] = []
    masks: List[np.ndarray] = []
    for start, stop, kmers in positions_and_kmers:
        index = 0
        while index < len(kmers):
            word = bytearray()
            init_mask = 0
            found_mask = 0
            while index < len(kmers):
                kmer = kmers[index]
                if type(kmer) is not str:
                    raise TypeError(f"Kmer should be a string not {type(kmer)}")
                if not kmer.isascii():
                    raise ValueError("Only ASCII strings are supported")
                if len(kmer) > MAXIMUM_WORD_SIZE:
                    raise ValueError(
                        f"{kmer} of length {len(kmer)} is longer than the maximum of {MAXIMUM_WORD_SIZE}."
                    )
                if len(word) + len(kmer) > MAXIMUM_WORD_SIZE:
                    break
                # Shift counts are taken modulo 64 like the x86-64 build of the reference does
                # for the degenerate empty k-mer that kmer_chunks() emits when there are more
                # allowed errors than adapter characters (1 << (0 + 0 - 1) in
                # _kmer_finder.pyx:143-147): the result is a harmless duplicate bit.
                init_mask |= 1 << (len(word) & 63)
                word += kmer.encode("ascii")
                found_mask |= 1 << ((len(word) - 1) & 63)
                index += 1
            mask = [0] * 128
            for bit, char in enumerate(word):
                if char == 0:
                    continue
                for code in lookup[char]:
                    mask[code] |= 1 << bit
            entries.append((int(start), 0 if stop is None else int(stop), init_mask, found_mask))
            masks.append(np.array(mask, dtype=np.uint64))
    mask_array = np.concatenate(masks) if masks else np.zeros(0, dtype=np.uint64)
    return entries, mask_array


class KmerFinder:
    """
    Find k-mers in strings, case-independent and IUPAC-aware (see the reference docstring,
    _kmer_finder.pyx:66-97).
    """

    def __init__(self, positions_and_kmers, ref_wildcards: bool = False, query_wildcards: bool = False):
        self.ref_wildcards = bool(ref_wildcards)
        self.query_wildcards = bool(query_wildcards)
        self._entries, self._masks = build_kmer_tables(
            positions_and_kmers, self.ref_wildcards, self.query_wildcards
        )
        self.positions_and_kmers = positions_and_kmers

    def __reduce__(self):
        return KmerFinder, (self.positions_and_kmers, self.ref_wildcards, self.query_wildcards)

    @property
    def tables(self):
        """(entries, masks) in the C-ABI form (cg_kmer_entry / 128 x uint64 per entry)."""
        return self._entries, self._masks

    def kmers_present_batch(self, sequences: Sequence[str]) -> np.ndarray:
        try:
            data, offsets = _lib.pack_strings(sequences)
        except ValueError:
            raise ValueError("Only ASCII strings are supported") from None
        n = len(sequences)
        out = np.zeros(n, dtype=np.uint8)
        if n == 0:
            return out.astype(bool)
        ents = (_lib.cg_kmer_entry * max(len(self._entries), 1))()
        for j, (start, stop, init, found) in enumerate(self._entries):
            ents[j].search_start, ents[j].search_stop = start, stop
            ents[j].init_mask, ents[j].found_mask = init, found
        masks = self._masks if self._masks.size else np.zeros(128, dtype=np.uint64)
        ctx = _lib.default_context()
        _lib.check(
            _lib.lib().cg_kmers_present_batch(
                ctx.handle, ents, masks.ctypes.data, len(self._entries), data.ctypes.data,
                offsets.ctypes.data, n, out.ctypes.data,
            )
        )
        return out.astype(bool)

    def kmers_present(self, sequence: str) -> bool:
        if not isinstance(sequence, str):
            raise TypeError("sequence must be str")
        return bool(self.kmers_present_batch([sequence])[0])
