"""
oracle/oracle.py -- Python face of the CPU oracle.

TEST INFRASTRUCTURE (see the header of cutadapt_oracle.c).  Loads ``liboracle.so`` (the plain-C
restatement of Aligner.locate / comparers / KmerFinder / quality_trim_index), compiling it with
gcc on first use, and restates in Python the *composition* the reference does above those
native functions:

    <Adapter>.match_to          src/cutadapt/adapters.py:707-724, 758-786, 815-832, 862-890,
                                915-935, 963-975, 1000-1012
    LinkedAdapter.match_to      adapters.py:1215-1227   (score/errors: 1113-1130)
    MultipleAdapters.match_to   adapters.py:1265-1286
    QualityTrimmer + AdapterCutter rounds   modifiers.py:853-858, 225-231

``oracle_process`` consumes the same plain-dict adapter descriptions as the product's
``cutadapt_b200._lib.AdapterSetSpec`` and produces records in the cg_match layout, so the
parity tests compare arrays element by element.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cutadapt_oracle.c")
SO = os.path.join(HERE, "liboracle.so")

KIND_ALIGNER, KIND_PREFIX, KIND_SUFFIX = 0, 1, 2
REMOVE_BEFORE, REMOVE_AFTER, REMOVE_AUTO = 0, 1, 2
GROUP_SINGLE, GROUP_LINKED = 0, 1

MATCH_DTYPE = np.dtype(
    [("adapter", "<i4"), ("astart", "<i4"), ("astop", "<i4"), ("rstart", "<i4"), ("rstop", "<i4"),
     ("score", "<i4"), ("errors", "<i4"), ("info", "<i4")]
)


class KmerEntry(C.Structure):
    _fields_ = [("search_start", C.c_int64), ("search_stop", C.c_int64),
                ("init_mask", C.c_uint64), ("found_mask", C.c_uint64)]


class OracleAdapter(C.Structure):
    _fields_ = [("seq", C.c_char_p), ("m", C.c_int32), ("flags", C.c_int32), ("wildcard_ref", C.c_int32),
                ("wildcard_query", C.c_int32), ("indel_cost", C.c_int32), ("min_overlap", C.c_int32),
                ("kind", C.c_int32), ("reverse_read", C.c_int32), ("remove", C.c_int32), ("n_entries", C.c_int32),
                ("max_error_rate", C.c_double), ("entries", C.POINTER(KmerEntry)), ("masks", C.c_void_p)]


class OracleGroup(C.Structure):
    _fields_ = [("type", C.c_int32), ("a0", C.c_int32), ("a1", C.c_int32), ("front_required", C.c_int32),
                ("back_required", C.c_int32)]


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", SO, SRC, "-lm"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        handle = C.CDLL(build())
        u8p, ip = C.c_char_p, C.POINTER(C.c_int)
        handle.oracle_locate.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_double, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, ip]
        handle.oracle_prefix_compare.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_double, C.c_int,
                                                 C.c_int, C.c_int, ip]
        handle.oracle_suffix_compare.argtypes = handle.oracle_prefix_compare.argtypes
        handle.oracle_effective_length.argtypes = [u8p, C.c_int, C.c_int]
        handle.oracle_kmer_pack.argtypes = [C.c_int64, C.c_int64, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                            C.POINTER(KmerEntry), C.c_void_p]
        handle.oracle_kmers_present.argtypes = [C.POINTER(KmerEntry), C.c_void_p, C.c_int, u8p, C.c_int64]
        handle.oracle_quality_trim_index.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip]
        handle.oracle_nextseq_trim_index.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int]
        handle.oracle_poly_a_trim_index.argtypes = [u8p, C.c_int, C.c_int]
        handle.oracle_expected_errors.argtypes = [u8p, C.c_size_t, C.c_ubyte]
        handle.oracle_expected_errors.restype = C.c_double
        handle.oracle_locate_batch.argtypes = [u8p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                               C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.POINTER(KmerEntry), C.c_void_p, C.c_int, C.c_void_p]
        handle.oracle_process_batch.argtypes = [
            C.POINTER(OracleAdapter), C.c_int, C.POINTER(OracleGroup), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_void_p, C.c_void_p]
        _lib = handle
    return _lib


def _b(s):
    return s if isinstance(s, bytes) else s.encode("latin-1")


def _raise(rc):
    if rc == -2:
        raise ValueError("String must contain only ASCII characters")
    if rc == -4:
        raise MemoryError()
    raise ValueError(f"oracle: invalid argument (rc={rc})")


def locate(reference, query, max_error_rate, flags=15, wildcard_ref=False, wildcard_query=False,
           indel_cost=1, min_overlap=1):
    """Aligner(reference, ...).locate(query)   (_align.pyx:298-587)"""
    out = (C.c_int * 6)()
    r, q = _b(reference), _b(query)
    rc = lib().oracle_locate(r, len(r), q, len(q), max_error_rate, flags, int(wildcard_ref),
                             int(wildcard_query), indel_cost, min_overlap, out)
    if rc < 0:
        _raise(rc)
    return tuple(out) if rc == 1 else None


def prefix_compare(reference, query, max_error_rate, wildcard_ref=False, wildcard_query=False, min_overlap=1):
    out = (C.c_int * 6)()
    r, q = _b(reference), _b(query)
    rc = lib().oracle_prefix_compare(r, len(r), q, len(q), max_error_rate, int(wildcard_ref),
                                     int(wildcard_query), min_overlap, out)
    if rc < 0:
        _raise(rc)
    return tuple(out) if rc == 1 else None


def suffix_compare(reference, query, max_error_rate, wildcard_ref=False, wildcard_query=False, min_overlap=1):
    out = (C.c_int * 6)()
    r, q = _b(reference), _b(query)
    rc = lib().oracle_suffix_compare(r, len(r), q, len(q), max_error_rate, int(wildcard_ref),
                                     int(wildcard_query), min_overlap, out)
    if rc < 0:
        _raise(rc)
    return tuple(out) if rc == 1 else None


def quality_trim_index(qualities, cutoff_front, cutoff_back, base=33):
    q = _b(qualities)
    a, b = C.c_int(), C.c_int()
    lib().oracle_quality_trim_index(q, len(q), cutoff_front, cutoff_back, base, C.byref(a), C.byref(b))
    return a.value, b.value


def nextseq_trim_index(sequence, qualities, cutoff, base=33):
    """qualtrim.pyx:76-117"""
    b, q = _b(sequence), _b(qualities)
    return int(lib().oracle_nextseq_trim_index(b, q, len(q), cutoff, base))


def poly_a_trim_index(sequence, revcomp=False):
    """qualtrim.pyx:120-169"""
    b = _b(sequence)
    return int(lib().oracle_poly_a_trim_index(b, len(b), int(bool(revcomp))))


def expected_errors(qualities, base=33):
    """qualtrim.pyx:172-197 (a negative result = invalid quality character)"""
    q = _b(qualities)
    return float(lib().oracle_expected_errors(q, len(q), base))


class KmerTables:
    """KmerFinder(positions_and_kmers, ref_wildcards, query_wildcards) tables (_kmer_finder.pyx:106-165)"""

    def __init__(self, positions_and_kmers, ref_wildcards=False, query_wildcards=False):
        total = sum(len(k) for _, _, k in positions_and_kmers)
        self.entries = (KmerEntry * max(total, 1))()
        self.masks = np.zeros(128 * max(total, 1), dtype=np.uint64)
        n = 0
        for start, stop, kmers in positions_and_kmers:
            blob = b"".join(_b(k) + b"\0" for k in kmers)
            got = lib().oracle_kmer_pack(
                start, 0 if stop is None else stop, blob, len(kmers), int(ref_wildcards), int(query_wildcards),
                C.cast(C.byref(self.entries, n * C.sizeof(KmerEntry)), C.POINTER(KmerEntry)),
                self.masks.ctypes.data + 128 * 8 * n,
            )
            if got < 0:
                raise ValueError("k-mer longer than 64 or not ASCII")
            n += got
        self.n = n

    def present(self, sequence):
        s = _b(sequence)
        rc = lib().oracle_kmers_present(self.entries, self.masks.ctypes.data, self.n, s, len(s))
        if rc < 0:
            _raise(rc)
        return bool(rc)

    def as_lists(self):
        """(entries as [(start, stop, init, found)], masks uint64[128*n]) -- the C-ABI form."""
        ents = [(e.search_start, e.search_stop, e.init_mask, e.found_mask) for e in self.entries[: self.n]]
        return ents, self.masks[: 128 * self.n].copy()


# ---- composition ----------------------------------------------------------------------------


def _kmer_tables_of(adapter):
    """Accept either reference-form lists (kmer_entries/kmer_masks) or positions_and_kmers."""
    if adapter.get("_oracle_kt") is not None:
        return adapter["_oracle_kt"]
    kt = None
    entries = adapter.get("kmer_entries")
    if entries is not None and len(entries):
        kt = KmerTables([], False, False)
        kt.entries = (KmerEntry * len(entries))()
        for j, (start, stop, init, found) in enumerate(entries):
            kt.entries[j].search_start, kt.entries[j].search_stop = start, stop
            kt.entries[j].init_mask, kt.entries[j].found_mask = init, found
        kt.masks = np.ascontiguousarray(adapter["kmer_masks"], dtype=np.uint64).reshape(-1)
        kt.n = len(entries)
    adapter["_oracle_kt"] = kt
    return kt


def match_single(adapter, sequence):
    """One SingleAdapter.match_to: returns dict(astart, astop, rstart, rstop, score, errors, remove) or None."""
    seq = sequence[::-1] if adapter.get("reverse_read") else sequence
    kt = _kmer_tables_of(adapter)
    if kt is not None and not kt.present(seq):
        return None
    kind = adapter.get("kind", KIND_ALIGNER)
    args = dict(wildcard_ref=adapter.get("wildcard_ref", False), wildcard_query=adapter.get("wildcard_query", False),
                min_overlap=adapter.get("min_overlap", 1))
    if kind == KIND_ALIGNER:
        res = locate(adapter["sequence"], seq, adapter["max_error_rate"], adapter.get("flags", 15),
                     indel_cost=adapter.get("indel_cost", 1), **args)
    elif kind == KIND_PREFIX:
        res = prefix_compare(adapter["sequence"], seq, adapter["max_error_rate"], **args)
    else:
        res = suffix_compare(adapter["sequence"], seq, adapter["max_error_rate"], **args)
    if res is None:
        return None
    astart, astop, rstart, rstop, score, errors = res
    if adapter.get("reverse_read"):                      # adapters.py:777-785, 881-889
        m, n = len(adapter["sequence"]), len(sequence)
        astart, astop, rstart, rstop = m - astop, m - astart, n - rstop, n - rstart
    remove = adapter.get("remove", REMOVE_AFTER)
    if remove == REMOVE_AUTO:                            # adapters.py:930-935
        remove = REMOVE_BEFORE if rstart == 0 else REMOVE_AFTER
    return dict(astart=astart, astop=astop, rstart=rstart, rstop=rstop, score=score, errors=errors, remove=remove)


def _trim(hit, s, e):
    if hit["remove"] == REMOVE_BEFORE:
        return s + hit["rstop"], e
    return s, s + hit["rstart"]


def match_group(adapters, group, sequence):
    """Returns (score, errors, front_hit_or_None, back_hit_or_None) or None."""
    typ, a0, a1, front_required, back_required = group
    if typ == GROUP_SINGLE:
        h = match_single(adapters[a0], sequence)
        if h is None:
            return None
        h["adapter"] = a0
        return h["score"], h["errors"], h, None
    front = match_single(adapters[a0], sequence)         # adapters.py:1219-1227
    if front_required and front is None:
        return None
    rest = sequence
    if front is not None:
        front["adapter"] = a0
        s, e = _trim(front, 0, len(sequence))
        rest = sequence[s:e]
    back = match_single(adapters[a1], rest)
    if back is None and (back_required or front is None):
        return None
    if back is not None:
        back["adapter"] = a1
    score = (front["score"] if front else 0) + (back["score"] if back else 0)
    errors = (front["errors"] if front else 0) + (back["errors"] if back else 0)
    return score, errors, front, back


def oracle_process(adapters, groups, sequences, qualities=None, quality_trim=False, cutoff_front=0,
                   cutoff_back=0, quality_base=33, times=1, nextseq_cutoff=None):
    """Whole per-read pass; returns (matches[n, times, slots], qtrim[n, 2]).
    Modifier order as cutadapt builds it (cli.py:940-953): NextseqQualityTrimmer, QualityTrimmer, AdapterCutter."""
    if groups is None:
        groups = [(GROUP_SINGLE, i, -1, 0, 0) for i in range(len(adapters))]
    slots = 2 if any(g[0] == GROUP_LINKED for g in groups) else 1
    n = len(sequences)
    out = np.zeros((n, times, slots), dtype=MATCH_DTYPE)
    out["adapter"] = -1
    qtrim = np.zeros((n, 2), dtype=np.int32)

    def put(rec, hit, gi, searched):
        rec["adapter"] = hit["adapter"]
        for f in ("astart", "astop", "rstart", "rstop", "score", "errors"):
            rec[f] = hit[f]
        info = (gi & 255) | (256 if hit["remove"] == REMOVE_AFTER else 0) | ((searched & 0xFFFF) << 16)
        rec["info"] = info - (1 << 32) if info >= (1 << 31) else info   # stored as int32

    for i, seq in enumerate(sequences):
        s, e = 0, len(seq)
        if nextseq_cutoff is not None:                       # modifiers.py:834-837: read[:stop]
            e = nextseq_trim_index(seq, qualities[i], nextseq_cutoff, quality_base)
        if quality_trim:                                     # modifiers.py:854-858 on what is left
            s, e = quality_trim_index(qualities[i][:e], cutoff_front, cutoff_back, quality_base)
        qtrim[i] = (s, e)
        for r in range(times):
            cur = seq[s:e]
            best = None
            for gi, g in enumerate(groups):                  # adapters.py:1271-1286
                m = match_group(adapters, g, cur)
                if m is None:
                    continue
                if best is None or m[0] > best[1][0] or (m[0] == best[1][0] and m[1] < best[1][1]):
                    best = (gi, m)
            if best is None:
                break                                        # modifiers.py:227-229
            gi, (_, _, h0, h1) = best
            searched = e - s
            if h0 is not None:
                put(out[i, r, 0], h0, gi, searched)
            if h1 is not None:
                s2, e2 = (0, searched) if h0 is None else _trim(h0, 0, searched)
                put(out[i, r, 1], h1, gi, e2 - s2)
            if h0 is not None:
                s, e = _trim(h0, s, e)
            if h1 is not None:
                s, e = _trim(h1, s, e)
    return out, qtrim


# ---- AdapterIndex (adapters.py:1289-1551): anchored adapters looked up in a dictionary of their neighbourhoods ----


def hamming_sphere(s, k):
    """All strings at Hamming distance exactly k from s over ACGT (_align.pyx:717-782)."""
    import itertools

    for positions in itertools.combinations(range(len(s)), k):
        choices = [[c for c in "ACGT" if c != s[p]] for p in positions]
        for repl in itertools.product(*choices):
            t = list(s)
            for p, c in zip(positions, repl):
                t[p] = c
            yield "".join(t)


def edit_environment(t, k):
    """
    (s, errors, matches) for every string s over ACGT within edit distance k of t (_align.pyx:785-882): a depth-first
    walk over the strings with one DP row per character; `matches` is carried along the same tie-broken optimal
    path as there (diagonal if <= both, else left if <= up, else up).
    """
    n = len(t)
    big = (k + 1) * 0x01010101
    costs = [[big] * (n + 1) for _ in range(n + k + 1)]
    matches = [[0] * (n + 1) for _ in range(n + k + 1)]
    for i in range(n + k + 1):
        costs[i][0] = i
    for j in range(n + 1):
        costs[0][j] = j
    s = [0] * (n + k + 1)
    tt = ["ACGT".index(c) for c in t.upper()]
    i = 0
    while True:
        if i > 0:
            ch = s[i - 1]
            min_cost = 999999999
            for j in range(max(1, i - k), min(n + 1, i + k + 1)):
                match = 0 if tt[j - 1] == ch else 1
                diag = costs[i - 1][j - 1] + match
                left = costs[i][j - 1] + 1
                up = costs[i - 1][j] + 1
                if diag <= left and diag <= up:
                    c, m = diag, matches[i - 1][j - 1] + (1 - match)
                elif left <= up:
                    c, m = left, matches[i][j - 1]
                else:
                    c, m = up, matches[i - 1][j]
                costs[i][j] = c
                matches[i][j] = m
                min_cost = min(min_cost, c)
        else:
            min_cost = 0
        if costs[i][n] <= k:
            yield "".join("ACGT"[c] for c in s[:i]), costs[i][n], matches[i][n]
        if min_cost <= k and i < n + k:
            s[i] = 0
            i += 1
        else:
            while True:
                if i == 0:
                    return
                i -= 1
                # (the rows above i keep stale cells of the previous branch outside their band, exactly like the
                # reference's single matrix: they are re-initialised only where the band writes)
                if s[i] < 3:
                    break
            s[i] += 1
            i += 1


def index_build(sequences, max_error_rate, indels):
    """AdapterIndex._make_index (adapters.py:1398-1472): (lengths descending, {string: (adapter no, errors, matches)})."""
    index, lengths, ambiguous = {}, set(), {}
    for a, sequence in enumerate(sequences):
        k = int(max_error_rate * len(sequence))
        if indels:
            env = edit_environment(sequence, k)
        else:
            env = ((s, e, len(sequence) - e) for e in range(k + 1) for s in hamming_sphere(sequence, e))
        for s, errors, matches in env:
            if s in index:
                _, _, other_matches = index[s]
                if matches < other_matches:
                    continue
                if other_matches == matches and s not in ambiguous:
                    ambiguous[s] = True
            index[s] = (a, errors, matches)
            lengths.add(len(s))
    for s in ambiguous:
        del index[s]
    return sorted(lengths, reverse=True), index


def oracle_index_process(sequences, max_error_rate, indels, prefix, data, offsets, descriptors=None):
    """
    IndexedPrefixAdapters / IndexedSuffixAdapters.match_to over a packed batch (adapters.py:1474-1551): records with
    the fields adapter, astart, astop, rstart, rstop, score (= matches), errors.  `descriptors` (one adapter
    description per sequence, as for oracle_process) are needed only for reads with an N in the looked-up affix
    (_lookup_with_n re-aligns with the adapter itself).
    """
    lengths, index = index_build(sequences, max_error_rate, indels)
    bindex = {k.encode(): v for k, v in index.items()}
    n = offsets.size - 1
    out = np.zeros(n, dtype=MATCH_DTYPE)
    out["adapter"] = -1
    raw = np.ascontiguousarray(data, dtype=np.uint8).tobytes()
    for i in range(n):
        seq = raw[offsets[i]:offsets[i + 1]].upper()
        best = None
        best_m, best_e, best_len = -1, 1000, 0
        for length in lengths:
            if length < best_m:
                break
            affix = seq[:length] if prefix else seq[-length:]
            if b"N" in affix:
                hit = bindex.get(affix.replace(b"N", b"A"))
                if hit is None:
                    continue
                m = match_single(descriptors[hit[0]], affix.decode())
                if m is None:
                    continue
                a, e, mm = hit[0], m["errors"], m["score"]
            else:
                hit = bindex.get(affix)
                if hit is None:
                    continue
                a, e, mm = hit
            if mm > best_m or (mm == best_m and e < best_e):
                best, best_e, best_m, best_len = a, e, mm, length
        if best_m == -1:
            continue
        r = out[i]
        r["adapter"] = best
        r["astart"], r["astop"] = 0, len(sequences[best])
        if prefix:
            r["rstart"], r["rstop"] = 0, best_len
        else:
            r["rstart"], r["rstop"] = len(seq) - best_len, len(seq)
        r["score"], r["errors"] = best_m, best_e
    return out


def oracle_process_packed(adapters, groups, data, offsets, qdata=None, quality_trim=False, cutoff_front=0,
                          cutoff_back=0, quality_base=33, times=1, nextseq_cutoff=None, threads=None):
    """
    oracle_process on a packed batch (uint8 bytes + int64 offsets), run by the C loop oracle_process_batch from
    `threads` Python threads over disjoint read ranges (ctypes releases the GIL): 10^6 reads in seconds.
    Same results as oracle_process (tests/test_oracle.py checks that).
    """
    from concurrent.futures import ThreadPoolExecutor

    if groups is None:
        groups = [(GROUP_SINGLE, i, -1, 0, 0) for i in range(len(adapters))]
    slots = 2 if any(g[0] == GROUP_LINKED for g in groups) else 1
    data = np.ascontiguousarray(data, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    keep = []
    ads = (OracleAdapter * max(len(adapters), 1))()
    for i, a in enumerate(adapters):
        seq = _b(a["sequence"])
        keep.append(seq)
        ads[i].seq = seq; ads[i].m = len(seq)
        ads[i].flags = a.get("flags", 15)
        ads[i].wildcard_ref = int(a.get("wildcard_ref", False)); ads[i].wildcard_query = int(a.get("wildcard_query", False))
        ads[i].indel_cost = a.get("indel_cost", 1); ads[i].min_overlap = a.get("min_overlap", 1)
        ads[i].kind = a.get("kind", KIND_ALIGNER); ads[i].reverse_read = int(bool(a.get("reverse_read")))
        ads[i].remove = a.get("remove", REMOVE_AFTER)
        ads[i].max_error_rate = a["max_error_rate"]
        kt = _kmer_tables_of(a)
        if kt is not None:
            keep.append(kt)
            ads[i].n_entries = kt.n
            ads[i].entries = C.cast(kt.entries, C.POINTER(KmerEntry))
            ads[i].masks = kt.masks.ctypes.data
    grs = (OracleGroup * max(len(groups), 1))()
    for i, g in enumerate(groups):
        grs[i].type, grs[i].a0, grs[i].a1, grs[i].front_required, grs[i].back_required = [int(x) for x in g]
    out = np.zeros((n, times, slots), dtype=MATCH_DTYPE)
    qtrim = np.zeros((n, 2), dtype=np.int32)
    if n == 0:
        return out, qtrim
    qd = np.ascontiguousarray(qdata, dtype=np.uint8) if qdata is not None else None
    if (quality_trim or nextseq_cutoff is not None) and qd is None:
        raise ValueError("qualities needed")
    L = lib()
    L.oracle_quality_trim_index(b"I", 1, 0, 0, 33, C.byref(C.c_int()), C.byref(C.c_int()))   # (init tables once)
    L.oracle_locate(b"A", 1, b"A", 1, 0.0, 15, 1, 1, 1, 1, (C.c_int * 6)())
    threads = threads or min(32, len(os.sched_getaffinity(0)))
    step = max(1, (n + threads - 1) // threads)
    rec_stride = times * slots * 32

    def work(r0):
        r1 = min(n, r0 + step)
        return L.oracle_process_batch(
            ads, len(adapters), grs, len(groups), data.ctypes.data, qd.ctypes.data if qd is not None else None,
            offsets.ctypes.data, r0, r1, int(bool(quality_trim)), cutoff_front, cutoff_back, quality_base, times,
            int(nextseq_cutoff is not None), nextseq_cutoff or 0, slots,
            out.ctypes.data + r0 * rec_stride, qtrim.ctypes.data + r0 * 8)

    with ThreadPoolExecutor(threads) as ex:
        for rc in ex.map(work, range(0, n, step)):
            if rc < 0:
                _raise(rc)
    return out, qtrim


# ---------------------------------------------------------------------------------------------
# FASTQ chunk in -> trimmed FASTQ out: restatement of what one worker of the reference does with a
# chunk (runners.py:174-214 -> pipeline.py:47-73 -> steps.py:299-319), for the product's
# cg_fastq_trim_chunk.  dnaio (absent here) only parses and formats; its on-disk conventions are the
# FASTQ format itself: 4 lines per record, "\r\n" tolerated on input, "@name\nseq\n+\nqual\n" on output.
# ---------------------------------------------------------------------------------------------
class FastqFormatError(ValueError):
    pass


def parse_fastq(data: bytes):
    """[(name, sequence, qualities)] as str; raises FastqFormatError like dnaio's parser does."""
    if not data:
        return []
    lines = data.split(b"\n")
    if lines[-1] == b"":
        lines.pop()
    if len(lines) % 4:
        raise FastqFormatError(f"FASTQ chunk does not consist of complete 4-line records ({len(lines)} lines)")
    lines = [ln[:-1] if ln.endswith(b"\r") else ln for ln in lines]
    records = []
    for r in range(len(lines) // 4):
        h, s, p, q = lines[4 * r:4 * r + 4]
        if not h.startswith(b"@"):
            raise FastqFormatError(f"record {r}: a record does not start with '@'")
        if not p.startswith(b"+"):
            raise FastqFormatError(f"record {r}: the third line of a record does not start with '+'")
        if len(p) > 1 and p[1:] != h[1:]:
            raise FastqFormatError(f"record {r}: sequence descriptions don't match")
        if len(s) != len(q):
            raise FastqFormatError(f"record {r}: sequence and qualities differ in length")
        records.append((h[1:].decode("latin-1"), s.decode("latin-1"), q.decode("latin-1")))
    return records


def _apply_cuts(records, cut):
    """UnconditionalCutter, first in the chain (modifiers.py:66-95)"""
    for c_len in cut:
        records = [(nm, sq[c_len:], q[c_len:]) if c_len > 0 else (nm, sq[:c_len], q[:c_len]) for nm, sq, q in records]
    return records


def _quality_trimmed(records, quality_trim, cutoff_front, cutoff_back, quality_base, nextseq_cutoff):
    """NextseqQualityTrimmer + QualityTrimmer as modifiers of their own: (records, bases removed)"""
    trimmed, removed = [], 0
    for name, seq, q in records:
        s, e = 0, len(seq)
        if nextseq_cutoff is not None:
            e = nextseq_trim_index(seq, q, nextseq_cutoff, quality_base)
        if quality_trim:
            s, e = quality_trim_index(q[:e], cutoff_front, cutoff_back, quality_base)
        removed += len(seq) - (e - s)
        trimmed.append((name, seq[s:e], q[s:e]))
    return trimmed, removed


FILTER_CHAIN = ("too_short", "too_long", "too_many_n", "too_many_expected_errors", "casava_filtered",
                "discard_trimmed", "discard_untrimmed")     # the order cli.py:700-830, 870-910 appends them
_FILTER_COUNTER = {"discard_trimmed": "discarded", "discard_untrimmed": "discarded"}


def _fastq_evaluate(data, adapters, groups, quality_trim=False, cutoff_front=0, cutoff_back=0, quality_base=33, times=1,
                    nextseq_cutoff=None, minimum_length=0, maximum_length=-1, discard_trimmed=False,
                    discard_untrimmed=False, max_n=-1.0, max_expected_errors=-1.0, cut=(), poly_a=False, length=None,
                    trim_n=False, discard_casava=False, second_mate=False, want_last_adapter=False, action="trim",
                    revcomp=False, rc_suffix=True, match_override=None, info_names=None, info_rows=None,
                    rest_rows=None, wildcard_rows=None, adapter_sequences=None):
    """Modifier chain on every record of a chunk + the verdict of every enabled filter.
    Returns ([(name, sequence, qualities, {filter: bool})], enabled filters, per-read counters).
    match_override: match records (n, 1, slots) found by the caller on the quality-trimmed reads (--pair-adapters).
    info_names / info_rows: name per adapter + a list that receives the --info-file rows (InfoFileWriter.__call__,
    steps.py:222-253: coordinates applied to info.original_read, the read as it came)."""
    records = parse_fastq(data)
    originals = list(records)
    is_rc = [False] * len(records)
    bp_in = sum(len(r[1]) for r in records)             # before any modifier (pipeline.py:58-64)
    records = _apply_cuts(records, cut)
    seqs = [r[1] for r in records]
    quals = [r[2] for r in records]
    n = len(records)
    pre_trimmed_bp = 0
    reverse_complemented = 0
    if (adapters and revcomp) or match_override is not None:
        # ReverseComplementer (modifiers.py:264-308) wraps the AdapterCutter only, PairedAdapterCutter replaces it:
        # the quality trimmers come first, the cutter sees the trimmed read
        records, pre_trimmed_bp = _quality_trimmed(records, quality_trim, cutoff_front, cutoff_back, quality_base,
                                                   nextseq_cutoff)
        seqs = [r[1] for r in records]
        if match_override is not None:
            matches = match_override
        else:
            comp = bytes.maketrans(b"ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", b"TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn")
            rc_seqs = [sq.encode("latin-1").translate(comp)[::-1].decode("latin-1") for sq in seqs]
            fwd, _ = oracle_process(adapters, groups, seqs, None, False, 0, 0, quality_base, times, None)
            rev, _ = oracle_process(adapters, groups, rc_seqs, None, False, 0, 0, quality_base, times, None)
            matches = fwd.copy()
            for i in range(n):
                score_f = int(fwd[i]["score"][fwd[i]["adapter"] >= 0].sum())
                score_r = int(rev[i]["score"][rev[i]["adapter"] >= 0].sum())
                if score_r > score_f:
                    reverse_complemented += 1
                    is_rc[i] = True
                    matches[i] = rev[i]
                    name, _, q = records[i]
                    records[i] = (name + (" rc" if rc_suffix else ""), rc_seqs[i], q[::-1])
        quals = [r[2] for r in records]
        qtrim = np.array([(0, len(r[1])) for r in records], dtype=np.int32).reshape(n, 2)
    elif adapters:
        matches, qtrim = oracle_process(adapters, groups, seqs, quals, quality_trim, cutoff_front, cutoff_back,
                                        quality_base, times, nextseq_cutoff)
    else:
        matches = None
        qtrim = np.zeros((n, 2), dtype=np.int32)
        for i, (seq, q) in enumerate(zip(seqs, quals)):
            s, e = 0, len(seq)
            if nextseq_cutoff is not None:
                e = nextseq_trim_index(seq, q, nextseq_cutoff, quality_base)
            if quality_trim:
                s, e = quality_trim_index(q[:e], cutoff_front, cutoff_back, quality_base)
            qtrim[i] = (s, e)
    enabled = [name for name, on in zip(FILTER_CHAIN, (minimum_length > 0, maximum_length >= 0, max_n >= 0,
                                                       max_expected_errors >= 0, discard_casava, discard_trimmed,
                                                       discard_untrimmed)) if on]
    c = dict(n_records=n, bp_in=bp_in, with_adapters=0, quality_trimmed_bp=pre_trimmed_bp,
             reverse_complemented=reverse_complemented)
    out = []
    for i, (name, seq, q) in enumerate(records):
        s, e = int(qtrim[i, 0]), int(qtrim[i, 1])
        c["quality_trimmed_bp"] += len(seq) - (e - s)
        matched = False
        last_adapter = -1
        if matches is not None:
            for r in range(matches.shape[1]):
                for slot in range(matches.shape[2]):
                    m = matches[i, r, slot]
                    if m["adapter"] < 0:
                        continue
                    matched = True
                    last_adapter = int(m["adapter"])
                    cur = range(s, e)                   # Python slicing of the current read, clamping included
                    if (int(m["info"]) >> 8) & 1:      # RemoveAfterMatch: read[:rstart]      adapters.py:486-487
                        cur = cur[:int(m["rstart"])]
                    else:                               # RemoveBeforeMatch: read[rstop:]      adapters.py:453-454
                        cur = cur[int(m["rstop"]):]
                    s, e = (cur.start, cur.stop) if len(cur) else (s, s)
        c["with_adapters"] += matched
        ts, tq = seq[s:e], q[s:e]
        if action != "trim":                           # AdapterCutter.match_and_trim, modifiers.py:214-251
            b0, b1 = int(qtrim[i, 0]), int(qtrim[i, 1])
            rseq, rq = seq[b0:b1], q[b0:b1]             # the read as the cutter saw it
            k0, k1 = s - b0, e - b0                     # remainder(matches), adapters.py:1588-1602
            if action == "lowercase":
                rseq = rseq.upper()                     # modifiers.py:222-223, also for reads without a match
            if not matched or action in (None, "none"):
                ts, tq = rseq, rq
            elif action == "mask":                      # masked_read, modifiers.py:175-182
                ts, tq = "N" * k0 + rseq[k0:k1] + "N" * (len(rseq) - k1), rq
            elif action == "lowercase":                 # lowercased_read, modifiers.py:184-193
                ts, tq = rseq[:k0].lower() + rseq[k0:k1].upper() + rseq[k1:].lower(), rq
            else:
                m0 = matches[i, 0, 0]
                m1 = matches[i, 0, 1] if matches.shape[2] > 1 else None
                has0, has1 = m0["adapter"] >= 0, m1 is not None and m1["adapter"] >= 0
                if action == "crop":                    # cropped_read, modifiers.py:195-198
                    m = m0 if has0 else m1
                    a, b = int(m["rstart"]), int(m["rstop"])
                elif has0 and (int(m0["info"]) >> 8) & 1:          # RemoveAfterMatch.retained_adapter_interval
                    a, b = 0, int(m0["rstop"])                      # adapters.py:479-480
                else:                                   # RemoveBeforeMatch (446-447) / LinkedMatch (1145-1155)
                    a = int(m0["rstart"]) if has0 else 0
                    offset = int(m0["rstop"]) if has0 else 0
                    b = int(m1["rstop"]) + offset if has1 else len(rseq)
                ts, tq = rseq[a:b], rq[a:b]
        if poly_a:                                      # PolyATrimmer (modifiers.py:861-879); revcomp form for R2
            if second_mate:
                idx = poly_a_trim_index(ts, revcomp=True)
                ts, tq = ts[idx:], tq[idx:]
            else:
                idx = poly_a_trim_index(ts)
                ts, tq = ts[:idx], tq[:idx]
        if length is not None:                          # Shortener (modifiers.py:882-899)
            ts, tq = (ts[:length], tq[:length]) if length >= 0 else (ts[length:], tq[length:])
        if trim_n:                                      # NEndTrimmer (modifiers.py:902-918): upper-case N only
            a = len(ts) - len(ts.lstrip("N"))
            b = len(ts.rstrip("N"))
            ts, tq = ts[a:b], tq[a:b]
        if info_rows is not None:
            _, cur_s, cur_q = originals[i]
            if is_rc[i]:
                comp = bytes.maketrans(b"ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", b"TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn")
                cur_s, cur_q = cur_s.encode("latin-1").translate(comp)[::-1].decode("latin-1"), cur_q[::-1]
            flag = ("1" if is_rc[i] else "0") if revcomp else ""
            if matched:
                for r in range(matches.shape[1]):
                    if not (matches[i, r]["adapter"] >= 0).any():
                        break
                    for slot in range(matches.shape[2]):
                        m = matches[i, r, slot]
                        if m["adapter"] < 0:
                            continue
                        rs, re_ = int(m["rstart"]), int(m["rstop"])
                        info_rows.append("\t".join([name, str(int(m["errors"])), str(rs), str(re_), cur_s[0:rs], cur_s[rs:re_],
                                                    cur_s[re_:], info_names[int(m["adapter"])], cur_q[0:rs], cur_q[rs:re_],
                                                    cur_q[re_:], flag]))
                        if (int(m["info"]) >> 8) & 1:
                            cur_s, cur_q = cur_s[:rs], cur_q[:rs]
                        else:
                            cur_s, cur_q = cur_s[re_:], cur_q[re_:]
            else:
                info_rows.append("\t".join([name, "-1", ts, tq]))
        if (rest_rows is not None or wildcard_rows is not None) and matched:
            # RestFileWriter / WildcardFileWriter (steps.py:193-220): the LAST match and the sequence its round searched
            cur = seq[int(qtrim[i, 0]):int(qtrim[i, 1])]
            last = None
            for r in range(matches.shape[1]):
                present = [m for m in matches[i, r] if m["adapter"] >= 0]
                if not present:
                    break
                for m in present:
                    last = (m, cur)
                    cur = cur[:int(m["rstart"])] if (int(m["info"]) >> 8) & 1 else cur[int(m["rstop"]):]
            m, cur = last
            if rest_rows is not None:
                rest = cur[int(m["rstop"]):] if (int(m["info"]) >> 8) & 1 else cur[:int(m["rstart"])]
                if rest:
                    rest_rows.append(f"{rest} {name}")
            if wildcard_rows is not None:
                aseq = adapter_sequences[int(m["adapter"])]
                astart, rstart = int(m["astart"]), int(m["rstart"])
                chars = [cur[rstart + k] for k in range(int(m["astop"]) - astart)
                         if aseq[astart + k] == "N" and rstart + k < len(cur)]
                wildcard_rows.append(f"{''.join(chars)} {name}")
        n_count = ts.lower().count("n")
        fails = {
            "too_short": len(ts) < minimum_length,                                        # predicates.py:29-40
            "too_long": len(ts) > maximum_length,                                         # predicates.py:43-53
            "too_many_n": (n_count / len(ts) > max_n if len(ts) else False) if max_n < 1.0 else n_count > max_n,
            "too_many_expected_errors": max_expected_errors >= 0 and expected_errors(tq) > max_expected_errors,
            "casava_filtered": name.partition(" ")[2][1:4] == ":Y:",                      # predicates.py:125-139
            "discard_trimmed": matched, "discard_untrimmed": not matched,                 # predicates.py:142-175
        }
        out.append((name, ts, tq, fails, last_adapter) if want_last_adapter else (name, ts, tq, fails))
    return out, enabled, c


def _fastq_record(name, ts, tq):
    return f"@{name}\n{ts}\n+\n{tq}\n".encode("latin-1")


def oracle_fastq_demux(data: bytes, adapters, groups, adapter_names, unknown="unknown", **options):
    """{name: bytes}: Demultiplexer.__call__ (steps.py:397-409) after the filters -- a surviving read goes to the
    output named after the adapter of its most recent match (adapter_names[i] = output of adapter i), else to
    `unknown`."""
    evaluated, enabled, _ = _fastq_evaluate(data, adapters, groups, want_last_adapter=True, **options)
    outputs = {name: [] for name in list(dict.fromkeys(adapter_names)) + [unknown]}
    for name, ts, tq, fails, last in evaluated:
        if any(fails[f] for f in enabled):
            continue
        outputs[adapter_names[last] if last >= 0 else unknown].append(_fastq_record(name, ts, tq))
    return {k: b"".join(v) for k, v in outputs.items()}


def oracle_fastq_trim(data: bytes, adapters=None, groups=None, **options):
    """(output bytes, counters dict) of one single-end chunk.  Options: see _fastq_evaluate.  Filters in the order
    cli.py appends them; the first that matches counts the read (SingleEndFilter, steps.py:70-102)."""
    evaluated, enabled, c = _fastq_evaluate(data, adapters, groups, **options)
    c.update(n_written=0, bp_out=0, too_short=0, too_long=0, too_many_n=0, too_many_expected_errors=0, discarded=0,
             casava_filtered=0)
    out = []
    for name, ts, tq, fails in evaluated:
        fired = next((f for f in enabled if fails[f]), None)
        if fired is not None:
            c[_FILTER_COUNTER.get(fired, fired)] += 1
            continue
        c["n_written"] += 1
        c["bp_out"] += len(ts)
        out.append(_fastq_record(name, ts, tq))
    return b"".join(out), c


def oracle_fastq_trim_paired(data1: bytes, data2: bytes, adapters1=None, groups1=None, adapters2=None, groups2=None,
                             options1=None, options2=None, pair_filter="any", pair_specs=None, route=None):
    """(out1, out2, counters1, counters2) of one paired-end chunk (PairedEndPipeline.process_reads, pipeline.py:125-153).
    Each filter works on the pair like PairedEndFilter (steps.py:105-180); with adapters on one mate only,
    --discard-untrimmed tests "both" (cli.py:859-893).
    pair_specs: --pair-adapters, [((adapters, groups) of adapter i alone for R1, the same for R2)].
    route: demultiplexing, a function (last adapter of R1 or -1, of R2 or -1) -> key or None (no writer: dropped, not
    counted); out1 / out2 are then dicts key -> bytes."""
    options1, options2 = dict(options1 or {}), dict(options2 or {})
    if pair_specs:
        adapters1, groups1 = pair_specs[0][0]
        adapters2, groups2 = pair_specs[0][1]
        options1["match_override"], options2["match_override"] = _best_adapter_pairs(data1, data2, pair_specs, options1,
                                                                                     options2)
    ev1, en1, c1 = _fastq_evaluate(data1, adapters1, groups1, want_last_adapter=True, **options1)
    ev2, en2, c2 = _fastq_evaluate(data2, adapters2, groups2, want_last_adapter=True, second_mate=True, **options2)
    if len(ev1) != len(ev2):
        raise FastqFormatError("paired FASTQ chunks differ in their number of records")
    for c in (c1, c2):
        c.update(n_written=0, bp_out=0, too_short=0, too_long=0, too_many_n=0, too_many_expected_errors=0, discarded=0,
                 casava_filtered=0)
    out1, out2 = ({}, {}) if route else ([], [])
    for (n1, s1, q1, f1, last1), (n2, s2, q2, f2, last2) in zip(ev1, ev2):
        fired = None
        for flt in FILTER_CHAIN:
            e1, e2 = flt in en1, flt in en2
            if not e1 and not e2:
                continue
            mode = pair_filter
            if flt == "discard_untrimmed" and (not adapters1 or not adapters2):
                mode = "both"
            if not e2:
                hit = f1[flt]
            elif not e1:
                hit = f2[flt]
            elif mode == "any":
                hit = f1[flt] or f2[flt]
            elif mode == "both":
                hit = f1[flt] and f2[flt]
            else:
                hit = f1[flt]
            if hit:
                fired = flt
                break
        if fired is not None:
            for c in (c1, c2):
                c[_FILTER_COUNTER.get(fired, fired)] += 1
            continue
        key = None
        if route:
            key = route(last1, last2)
            if key is None:
                continue
        for c in (c1, c2):
            c["n_written"] += 1
        c1["bp_out"] += len(s1)
        c2["bp_out"] += len(s2)
        if route:
            out1[key] = out1.get(key, b"") + _fastq_record(n1, s1, q1)
            out2[key] = out2.get(key, b"") + _fastq_record(n2, s2, q2)
        else:
            out1.append(_fastq_record(n1, s1, q1))
            out2.append(_fastq_record(n2, s2, q2))
    if route:
        return out1, out2, c1, c2
    return b"".join(out1), b"".join(out2), c1, c2


def _best_adapter_pairs(data1, data2, pair_specs, options1, options2):
    """PairedAdapterCutter._find_best_match_pair (modifiers.py:480-503) for every pair of reads: match records
    (n, 1, slots) for R1 and R2, `adapter` = number of the chosen adapter pair, -1 where no pair matches both mates."""
    seqs = []
    for data, o in ((data1, options1), (data2, options2)):
        records = _apply_cuts(parse_fastq(data), o.get("cut", ()))
        records, _ = _quality_trimmed(records, o.get("quality_trim", False), o.get("cutoff_front", 0), o.get("cutoff_back", 0),
                                      o.get("quality_base", 33), o.get("nextseq_cutoff"))
        seqs.append([r[1] for r in records])
    if len(seqs[0]) != len(seqs[1]):
        raise FastqFormatError("paired FASTQ chunks differ in their number of records")
    n = len(seqs[0])
    per_pair = []
    for (a1, g1), (a2, g2) in pair_specs:
        m1, _ = oracle_process(a1, g1, seqs[0], None, False, 0, 0, 33, 1, None)
        m2, _ = oracle_process(a2, g2, seqs[1], None, False, 0, 0, 33, 1, None)
        per_pair.append((m1, m2))
    slots = max(max(m1.shape[2], m2.shape[2]) for m1, m2 in per_pair)
    best1 = np.zeros((n, 1, slots), dtype=per_pair[0][0].dtype)
    best2 = np.zeros((n, 1, slots), dtype=per_pair[0][0].dtype)
    best1["adapter"] = -1
    best2["adapter"] = -1
    for i in range(n):
        best = None
        for k, (m1, m2) in enumerate(per_pair):
            h1 = m1[i, 0][m1[i, 0]["adapter"] >= 0]
            h2 = m2[i, 0][m2[i, 0]["adapter"] >= 0]
            if not len(h1) or not len(h2):
                continue
            key = (int(h1["score"].sum() + h2["score"].sum()), int(h1["errors"].sum() + h2["errors"].sum()))
            if best is None or key[0] > best[0] or (key[0] == best[0] and key[1] < best[1]):
                best = (key[0], key[1], k)
        if best is not None:
            for src, dst in ((per_pair[best[2]][0], best1), (per_pair[best[2]][1], best2)):
                row = src[i, 0].copy()
                row["adapter"] = np.where(row["adapter"] >= 0, best[2], -1)
                dst[i, 0, :len(row)] = row
    return best1, best2
