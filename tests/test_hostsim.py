"""
The per-read device functions (cutadapt_b200/csrc/cg_core.cuh) compiled for the host and checked
against the oracle and the golden vectors.  This exercises exactly the code the kernels run per
lane -- packed-cell DP, wide-cell DP, prefilter, comparers, quality trimming, linked / multiple
composition, rounds -- without a GPU.  The host build lives under tests/ and is never loaded by
the product.
"""
import random

import numpy as np
import pytest

from cutadapt_b200 import _lib as L
from oracle import oracle
from util import golden, hostsim_process, build_adapters, match_desc, spec_of, random_reads


def _single(ref, rate, flags, wr, wq, ic, mo, kind=0):
    return L.AdapterSetSpec([dict(sequence=ref, max_error_rate=rate, flags=flags, wildcard_ref=wr,
                                  wildcard_query=wq, indel_cost=ic, min_overlap=mo, kind=kind)])


@pytest.mark.parametrize("force_wide", [0, 1, 2, 10, 18, 34, 64])   # packed, wide, two-phase: smem column / registers+refine / registers+inline runs / split main+end passes / planned runs + exact shortcut
def test_locate_golden(force_wide):
    cases = golden("locate_kat.json.gz")
    for ref, q, rate, flags, wr, wq, ic, mo, expected in cases:
        rec, _ = hostsim_process(_single(ref, rate, flags, wr, wq, ic, mo), [q], force_wide=force_wide)
        r = rec[0, 0, 0]
        got = None if r["adapter"] < 0 else [int(r[f]) for f in ("astart", "astop", "rstart", "rstop", "score", "errors")]
        assert got == expected, (ref, q, rate, flags, wr, wq, ic, mo)


def test_comparers_golden():
    for ref, q, rate, wr, wq, mo, p, s in golden("comparer_kat.json.gz"):
        for kind, expected in ((1, p), (2, s)):
            rec, _ = hostsim_process(_single(ref, rate, 0, wr, wq, 1, mo, kind), [q])
            r = rec[0, 0, 0]
            got = None if r["adapter"] < 0 else [int(r[f]) for f in ("astart", "astop", "rstart", "rstop", "score", "errors")]
            assert got == expected, (kind, ref, q)


def test_adapters_golden():
    import cutadapt_b200.adapters as PA

    for case in golden("adapters_kat.json.gz"):
        multi = build_adapters(PA, case["adapters"])
        spec = spec_of(multi)
        reads = [r for r, _ in case["reads"]]
        recs, _ = hostsim_process(spec, reads)
        for i, (read, expected) in enumerate(case["reads"]):
            assert match_desc(multi.matches_from_records(recs[i, 0], read)) == expected, (case["adapters"], read)


def test_indexed_adapters_golden():
    """IndexedPrefixAdapters / IndexedSuffixAdapters (adapters.py:1289-1571): index build + lookups."""
    import cutadapt_b200.adapters as PA

    hits = 0
    for case in golden("index_kat.json.gz"):
        multi = build_adapters(PA, case["adapters"])
        indexed = [a for a in multi if isinstance(a, PA._IndexedAdapters)][0]
        assert len(indexed._index._index) == case["n_keys"]
        assert list(indexed._index._lengths) == case["lengths"]
        assert indexed._index._ambiguous == case["ambiguous"]
        spec = spec_of(multi)
        reads = [r for r, _ in case["reads"]]
        recs, _ = hostsim_process(spec, reads)
        for i, (read, expected) in enumerate(case["reads"]):
            assert match_desc(multi.matches_from_records(recs[i, 0], read)) == expected, (case["adapters"], read)
            hits += expected is not None
    assert hits > 1000


def test_rounds_and_quality_trim_against_oracle():
    import cutadapt_b200.adapters as PA

    rng = random.Random(5)
    for _ in range(60):
        ads = ["".join(rng.choice("ACGT") for _ in range(rng.randint(5, 20))) for _ in range(rng.randint(1, 3))]
        objs = [rng.choice([PA.BackAdapter, PA.FrontAdapter, PA.AnywhereAdapter])(a, max_errors=0.15, name="a") for a in ads]
        multi = PA.MultipleAdapters(objs)
        spec = spec_of(multi)
        reads = random_reads(rng, ads, 40)
        quals = ["".join(chr(33 + rng.choice([2, 2, 15, 30, 38])) for _ in r) for r in reads]
        times = rng.choice([1, 2, 3])
        params = L.make_params(quality_trim=True, cutoff_front=rng.choice([0, 10]), cutoff_back=20, times=times)
        got, qt = hostsim_process(spec, reads, quals, params)
        exp, eqt = oracle.oracle_process(spec.adapters, spec.groups, reads, quals, True, params.cutoff_front, 20, 33, times)
        assert (qt == eqt).all()
        assert (got == exp).all()


def test_empty_and_ragged_reads():
    import cutadapt_b200.adapters as PA

    multi = PA.MultipleAdapters([PA.BackAdapter("AGATCGGAAGAGC", name="a")])
    spec = spec_of(multi)
    reads = ["", "A", "AGA", "AGATCGGAAGAGC", "T" * 300 + "AGATCGGAAGAGC", "", "ACGT" * 50]
    got, _ = hostsim_process(spec, reads)
    exp, _ = oracle.oracle_process(spec.adapters, spec.groups, reads)
    assert (got == exp).all()
    assert got["adapter"][3, 0, 0] == 0 and got["rstart"][3, 0, 0] == 0


def test_two_phase_path_equals_general_path():
    """
    The scan + windowed-DP schedule of the two-phase kernel (process_read_simple: fused 32-bit scan
    words, locator hits, restarted DP windows) must reproduce the one-phase path bit for bit --
    every adapter type, wildcards, short and long reads (both group sizes), several adapter copies
    per read, partial occurrences at both ends, with and without quality trimming.
    """
    import cutadapt_b200.adapters as PA

    rng = random.Random(2024)
    types = ["FrontAdapter", "RightmostFrontAdapter", "BackAdapter", "RightmostBackAdapter", "AnywhereAdapter",
             "NonInternalFrontAdapter", "NonInternalBackAdapter", "PrefixAdapter", "SuffixAdapter", "BackAdapter"]
    n_windowed = 0
    for trial in range(250):
        alpha = rng.choice(["ACGT", "ACGT", "ACGTN", "AC", "ACGTacgtn"])
        m = rng.choice([rng.randint(3, 12), rng.randint(10, 34), 13, rng.randint(30, 70)])
        seq = "".join(rng.choice(rng.choice(["ACGT", "ACGT", "ACGTN", "ACGTRYN"])) for _ in range(m))
        if set(seq) <= {"N"}:
            seq = "A" + seq
        kw = dict(max_errors=rng.choice([0, 0.05, 0.1, 0.1, 0.15, 0.2, 0.3]), min_overlap=rng.randint(1, 6),
                  read_wildcards=rng.random() < 0.15, adapter_wildcards=rng.random() < 0.7)
        ad = getattr(PA, rng.choice(types))(seq, name="x", **kw)
        spec = spec_of(ad)
        reads = random_reads(rng, [seq], 50, alpha, rng.choice([60, 150, 150, 300, 700]))
        quals = ["".join(chr(33 + rng.choice([2, 2, 20, 30, 38])) for _ in r) for r in reads]
        qt = rng.random() < 0.3
        params = L.make_params(quality_trim=qt, cutoff_front=5, cutoff_back=20)
        a, qa = hostsim_process(spec, reads, quals if qt else None, params, 0)
        for mode in (2, 10, 18, 34, 64):
            b, qb = hostsim_process(spec, reads, quals if qt else None, params, mode)
            assert (a == b).all() and (qa == qb).all(), (mode, repr(ad))
        n_windowed += 1
    assert n_windowed == 250


def test_two_phase_on_golden_single_adapters():
    import cutadapt_b200.adapters as PA

    n = 0
    for case in golden("adapters_kat.json.gz"):
        if len(case["adapters"]) != 1 or case["adapters"][0][0] == "Linked":
            continue
        multi = build_adapters(PA, case["adapters"])
        spec = spec_of(multi)
        reads = [r for r, _ in case["reads"]]
        for mode in (2, 10, 18, 34, 64):
            recs, _ = hostsim_process(spec, reads, force_wide=mode)
            for i, (read, expected) in enumerate(case["reads"]):
                assert match_desc(multi.matches_from_records(recs[i, 0], read)) == expected, (mode, case["adapters"], read)
        n += 1
    assert n > 40


def _multipass(spec, reads, quals=None, params=None):
    """Mode 128 = the multi-pass schedule (per-component passes + select_best); None if not planned."""
    try:
        return hostsim_process(spec, reads, quals, params, force_wide=128)
    except RuntimeError as e:
        if e.args[0][0] == 100:
            return None
        raise


def test_multipass_schedule_golden():
    """Per-adapter passes + selection == MultipleAdapters / LinkedAdapter / AdapterIndex of the reference."""
    import cutadapt_b200.adapters as PA

    planned = 0
    for name in ("adapters_kat.json.gz", "index_kat.json.gz"):
        for case in golden(name):
            multi = build_adapters(PA, case["adapters"])
            spec = spec_of(multi)
            reads = [r for r, _ in case["reads"]]
            res = _multipass(spec, reads)
            if res is None:
                continue
            planned += 1
            for i, (read, expected) in enumerate(case["reads"]):
                assert match_desc(multi.matches_from_records(res[0][i, 0], read)) == expected, (case["adapters"], read)
    assert planned > 100


def test_multipass_schedule_with_quality_trim_against_oracle():
    import cutadapt_b200.adapters as PA

    rng = random.Random(77)
    planned = 0
    for _ in range(60):
        ads = ["".join(rng.choice("ACGT") for _ in range(rng.randint(5, 45))) for _ in range(rng.randint(2, 4))]
        objs = [rng.choice([PA.BackAdapter, PA.FrontAdapter, PA.AnywhereAdapter, PA.SuffixAdapter])(a, max_errors=0.15, name="a")
                for a in ads]
        if rng.random() < 0.5:
            objs.append(PA.LinkedAdapter(PA.PrefixAdapter(ads[0][:12], max_errors=0.2), PA.BackAdapter(ads[1], max_errors=0.1),
                                         rng.random() < 0.5, rng.random() < 0.5, "l"))
        multi = PA.MultipleAdapters(objs)
        spec = spec_of(multi)
        reads = random_reads(rng, ads, 40, max_len=120)
        quals = ["".join(chr(33 + rng.choice([2, 2, 15, 30, 38])) for _ in r) for r in reads]
        params = L.make_params(quality_trim=True, cutoff_front=rng.choice([0, 10]), cutoff_back=20)
        res = _multipass(spec, reads, quals, params)
        assert res is not None
        planned += 1
        exp, eqt = oracle.oracle_process(spec.adapters, spec.groups, reads, quals, True, params.cutoff_front, 20, 33, 1)
        assert (res[1] == eqt).all()
        assert (res[0] == exp).all()
    assert planned == 60


def test_trim_scan_device_functions_golden():
    """The device functions behind cg_nextseq_trim_batch / cg_poly_a_trim_batch, compiled for the host."""
    import ctypes as C
    from util import hostsim_lib

    lib = hostsim_lib()
    g = golden("trim_scans_kat.json.gz")
    for seq, qual, cutoff, base, expected in g["nextseq"]:
        assert lib.hs_nextseq_trim(seq.encode(), qual.encode(), len(qual), cutoff, base) == expected
    for seq, revcomp, expected in g["polya"]:
        assert lib.hs_poly_a_trim(seq.encode(), len(seq), int(revcomp)) == expected
    lib.hs_expected_errors.restype = C.c_double
    for qual, base, expected in g["expected_errors"]:
        assert float(lib.hs_expected_errors(qual.encode(), len(qual), base)).hex() == expected


def test_fused_nextseq_and_quality_trim_against_oracle():
    """NextseqQualityTrimmer -> QualityTrimmer -> AdapterCutter fused in one pass, all schedules."""
    import cutadapt_b200.adapters as PA

    rng = random.Random(91)
    for trial in range(40):
        ads = ["".join(rng.choice("ACGT") for _ in range(rng.randint(6, 30))) for _ in range(rng.randint(1, 3))]
        objs = [rng.choice([PA.BackAdapter, PA.FrontAdapter])(a, max_errors=0.1, name="a") for a in ads]
        multi = PA.MultipleAdapters(objs)
        spec = spec_of(multi)
        reads = [r + "G" * rng.choice([0, 0, 3, 10]) for r in random_reads(rng, ads, 40, max_len=100)]
        quals = ["".join(chr(33 + rng.choice([2, 2, 15, 30, 38])) for _ in r) for r in reads]
        qt = rng.random() < 0.6
        params = L.make_params(quality_trim=qt, cutoff_front=rng.choice([0, 10]), cutoff_back=20,
                               nextseq_cutoff=rng.choice([10, 20, 30]))
        exp, eqt = oracle.oracle_process(spec.adapters, spec.groups, reads, quals, qt, params.cutoff_front, 20, 33, 1,
                                         nextseq_cutoff=params.nextseq_cutoff)
        modes = [0, 128] if len(ads) > 1 else [0, 2, 10, 34, 64]
        for mode in modes:
            try:
                got, gqt = hostsim_process(spec, reads, quals, params, force_wide=mode)
            except RuntimeError as e:
                if e.args[0][0] == 100:
                    continue
                raise
            assert (gqt == eqt).all(), mode
            assert (got == exp).all(), mode


def test_reference_fasta_goldens_through_the_device_functions():
    """
    The command-line known answers of the reference with anchored / non-internal / linked adapters, --no-indels, -N
    and --match-read-wildcards (tests/golden/fastq_kat.json.gz, "fasta_cases"): match records from the host build of
    the device functions + pipeline.kept_intervals must give the expected reads.
    """
    import numpy as np
    from oracle import oracle
    from util import golden, fastq_file, adapter_from_spec, hostsim_process, spec_of
    import cutadapt_b200.adapters as PA
    from cutadapt_b200.pipeline import kept_intervals

    done = 0
    for c in golden("fastq_kat.json.gz")["fasta_cases"]:
        o = c["options"]
        if not o["specs"] or any(k in o for k in ("trim_n", "poly_a", "max_n")):
            continue
        params = dict(max_errors=o.get("error_rate", 0.1), min_overlap=o.get("min_overlap", 3),
                      adapter_wildcards=not o.get("no_wildcards", False), read_wildcards=o.get("read_wildcards", False),
                      indels=not o.get("no_indels", False))
        ads = [adapter_from_spec(spec, kind, name=f"a{i}", **params) for i, (kind, spec) in enumerate(o["specs"])]
        spec = spec_of(PA.MultipleAdapters(ads))
        records = oracle.parse_fastq(fastq_file(f"fa_{c['name']}.in.fastq"))
        seqs = [r[1] for r in records]
        matches, _ = hostsim_process(spec, seqs)
        iv = kept_intervals(matches, None, np.array([len(x) for x in seqs]))
        got = []
        for (name, seq, q), (a, b) in zip(records, iv):
            a, b = int(a), int(b)
            if o.get("minimum_length") and b - a < o["minimum_length"]:
                continue
            if "maximum_length" in o and b - a > o["maximum_length"]:
                continue
            got.append(f"@{name}\n{seq[a:b]}\n+\n{q[a:b]}\n")
        assert "".join(got).encode() == fastq_file(f"fa_{c['name']}.out.fastq"), (c["name"], c["command"])
        done += 1
    assert done >= 22
