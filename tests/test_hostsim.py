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


@pytest.mark.parametrize("force_wide", [0, 1])
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
