"""
Character-class tables of the hot path (host side, built once).

Mirrors the interface of the reference's ``cutadapt._match_tables``
(src/cutadapt/_match_tables.py:4-98): ``_acgt_table``, ``_iupac_table``, ``_upper_table``
and ``matches_lookup``.  The same three 256-byte tables are rebuilt in C++ for the device
(csrc/cg_setbuild.cpp: cg_build_enc_tables); tests/test_host_tables.py checks that both agree
with the golden tables generated from the reference.
"""

_IUPAC_BITS = {
    "X": 0b0000,
    "A": 0b0001,
    "C": 0b0010,
    "G": 0b0100,
    "T": 0b1000,
    "U": 0b1000,
    "R": 0b0101,
    "Y": 0b1010,
    "S": 0b0110,
    "W": 0b1001,
    "K": 0b1100,
    "M": 0b0011,
    "B": 0b1110,
    "D": 0b1101,
    "H": 0b1011,
    "V": 0b0111,
    # N matches everything, including characters that are not A/C/G/T: it carries the
    # 0x80 bit that _acgt_table assigns to "anything else" (_match_tables.py:20-61)
    "N": 0b1111 | 0x80,
}


def _fill(table: bytearray, mapping) -> None:
    for char, value in mapping.items():
        table[ord(char)] = value
        table[ord(char.lower())] = value


def _acgt_table() -> bytes:
    """A=1, C=2, G=4, T/U=8 (either case); every other byte is 0x80."""
    table = bytearray(b"\x80" * 256)
    _fill(table, {"A": 1, "C": 2, "G": 4, "T": 8, "U": 8})
    return bytes(table)


def _iupac_table() -> bytes:
    """IUPAC codes as 4-bit sets (either case); non-IUPAC bytes are 0."""
    table = bytearray(256)
    _fill(table, _IUPAC_BITS)
    return bytes(table)


def _upper_table() -> bytes:
    return bytes(range(256)).upper()


def matches_lookup(ref_wildcards: bool, query_wildcards: bool):
    """
    For every byte value r (a pattern character), the ASCII codes 1..127 that it matches
    under the given wildcard mode, as a ``bytes`` object -- the table KmerFinder builds its
    needle masks from (_match_tables.py:81-98).
    """
    if not ref_wildcards and not query_wildcards:
        ref_t, query_t = _upper_table(), _upper_table()
        same = True
    else:
        ref_t = _iupac_table() if ref_wildcards else _acgt_table()
        query_t = _iupac_table() if query_wildcards else _acgt_table()
        same = False
    result = []
    for r in range(256):
        rv = ref_t[r]
        if same:
            codes = [q for q in range(1, 128) if query_t[q] == rv]
        else:
            codes = [q for q in range(1, 128) if query_t[q] & rv]
        result.append(bytes(codes))
    return result
