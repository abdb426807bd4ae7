"""
GPU tests (-m gpu) of the FASTQ entry point: FASTQ chunk in, trimmed FASTQ
out, against the expected output files of the reference's command-line tests and against the oracle.
Byte work: the bar is byte-identical output.
"""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cutadapt_b200.pipeline import FastqTrimmer, PairedFastqTrimmer  # noqa: E402
from oracle import oracle  # noqa: E402
from util import (fastq_cases, fastq_case_adapters, fastq_case_kwargs, spec_of, fastq_paired_cases,  # noqa: E402
                  oracle_paired)


def trimmer_for(options, **extra):
    kw = fastq_case_kwargs(options)
    kw.update(extra)
    if kw.pop("quality_trim", False):
        kw["quality_cutoff"] = (kw.pop("cutoff_front"), kw.pop("cutoff_back"))
    return FastqTrimmer(fastq_case_adapters(options), **kw)


def oracle_for(options, data, **extra):
    import cutadapt_b200.adapters as PA

    ads = fastq_case_adapters(options)
    descs = groups = None
    if ads:
        spec = spec_of(PA.MultipleAdapters(ads))
        descs, groups = spec.adapters, spec.groups
    kw = fastq_case_kwargs(options)
    kw.update(extra)
    return oracle.oracle_fastq_trim(data, descs, groups, **kw)


def test_reference_command_line_goldens():
    """Every FASTQ known-answer case of the reference's test_commandline.py, byte for byte."""
    for c in fastq_cases():
        t = trimmer_for(c["options"])
        got = t.process_chunk(c["input_bytes"])
        assert got == c["expected_bytes"], c["name"]
        _, counters = oracle_for(c["options"], c["input_bytes"])
        for k, v in counters.items():
            assert t.statistics[k] == v, (c["name"], k)
        assert t.statistics["out_bytes"] == len(got)


def synthetic_fastq(n, seed, crlf=False, final_newline=True):
    rng = random.Random(seed)
    adapter = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA"
    out = []
    for i in range(n):
        ln = rng.choice((0, 1, 5, 30, 75, 100, 150, 150, 150, 151, 250))
        seq = "".join(rng.choice("ACGT") for _ in range(ln))
        if ln > 30 and rng.random() < 0.6:
            cut = rng.randrange(5, ln)
            seq = (seq[:cut] + adapter + seq)[:ln]
        if rng.random() < 0.05 and ln:
            seq = "".join(c if rng.random() > 0.2 else "N" for c in seq)
        if rng.random() < 0.1 and ln > 20:
            tail = rng.randrange(3, 20)
            seq = seq[:ln - tail] + "".join("A" if rng.random() < 0.93 else "C" for _ in range(tail))
        if rng.random() < 0.1 and ln > 10:
            a, b = rng.randrange(0, 4), rng.randrange(0, 4)
            seq = ("N" * a + seq[a:ln - b] + "N" * b)[:ln]
        if rng.random() < 0.02:
            seq = seq.lower()
        qual = "".join(chr(33 + min(41, max(2, int(rng.gauss(32 - 25 * (j / max(ln, 1)) ** 2, 6))))) for j in range(ln))
        name = f"read{i}" + ("" if rng.random() < 0.5 else f" 1:{rng.choice('NY')}:0:{'ACGT' * rng.randrange(0, 9)}")
        plus = "+" if rng.random() < 0.9 else "+" + name
        out.append(f"@{name}\n{seq}\n{plus}\n{qual}\n")
    data = "".join(out)
    if crlf:
        data = data.replace("\n", "\r\n")
    if not final_newline:
        data = data[:-2] if crlf else data[:-1]
    return data.encode()


@pytest.mark.parametrize("variant", ["plain", "crlf", "no_final_newline", "filters", "quality_only", "times2",
                                     "modifiers", "modifiers2", "mask", "lowercase", "none", "retain", "crop",
                                     "revcomp", "revcomp_quality", "revcomp_linked"])
def test_random_chunks_against_oracle(variant):
    options = dict(adapters=[["back", "AGATCGGAAGAGC"], ["front", "TTGACNNACG"]], quality_cutoff=[5, 20])
    extra = {}
    seed = {"plain": 1, "crlf": 2, "no_final_newline": 3, "filters": 4, "quality_only": 5, "times2": 6,
            "modifiers": 7, "modifiers2": 8, "mask": 9, "lowercase": 10, "none": 11, "retain": 12, "crop": 13,
            "revcomp": 14, "revcomp_quality": 15, "revcomp_linked": 16}[variant]
    data = synthetic_fastq(6000, seed=seed, crlf=variant == "crlf")
    if variant.startswith("revcomp"):       # every other record arrives as its reverse complement
        data = flip_records(data, seed)
    if variant == "no_final_newline":       # a last record whose quality line is not terminated
        data += b"@last\nACGTACGTAGATCGGAAGAGCAAA\n+\nIIIIIIIIIIIIIIIIIIIIIIII"
    if variant == "filters":
        extra = dict(minimum_length=20, maximum_length=140, max_n=0.1, max_expected_errors=2.5, discard_untrimmed=True)
    elif variant == "quality_only":
        options = dict(adapters=[], quality_cutoff=[0, 25], nextseq_cutoff=20)
        extra = dict(minimum_length=1, max_n=3)
    elif variant == "times2":
        extra = dict(times=2, discard_trimmed=False)
    elif variant == "modifiers":
        extra = dict(cut=[3, -2], poly_a=True, length=-90, trim_n=True, discard_casava=True, minimum_length=1)
    elif variant == "modifiers2":
        options = dict(adapters=[["anywhere", "AGATCGGAAGAGC"]])
        extra = dict(cut=[-4, -3, 2], poly_a=True, length=60, trim_n=True, max_n=0, discard_trimmed=True)
    elif variant == "mask":
        extra = dict(action="mask", times=2, trim_n=True, max_n=0.3, minimum_length=5)
    elif variant == "lowercase":
        extra = dict(action="lowercase", times=2, poly_a=True)
    elif variant == "none":
        extra = dict(action="none", discard_untrimmed=True, length=100)
    elif variant == "retain":
        options = dict(adapters=[["linked", "TTGACNNACG", "AGATCGGAAGAGC"], ["back", "CACGTCTGAACTC"],
                                 ["front", "ACGTACGTAC"]], quality_cutoff=[0, 15])
        extra = dict(action="retain", minimum_length=1)
    elif variant == "crop":
        options = dict(adapters=[["back", "AGATCGGAAGAGC"], ["front", "TTGACNNACG"], ["anywhere", "CACGTCTGAA"]])
        extra = dict(action="crop", discard_untrimmed=True, trim_n=True)
    elif variant == "revcomp":
        options = dict(adapters=[["back", "AGATCGGAAGAGC"], ["front", "TTGACNNACG"]])
        extra = dict(revcomp=True, minimum_length=10, trim_n=True)
    elif variant == "revcomp_quality":      # the quality trimmers run before the reverse complementer
        extra = dict(revcomp=True, rc_suffix=False, nextseq_cutoff=15, cut=[2, -3], times=2, discard_untrimmed=True)
    elif variant == "revcomp_linked":
        options = dict(adapters=[["linked", "TTGACNNACG", "AGATCGGAAGAGC"], ["back", "CACGTCTGAACTC"]], quality_cutoff=[0, 15])
        extra = dict(revcomp=True, action="mask", poly_a=True)
    t = trimmer_for(options, **extra)
    got = t.process_chunk(data)
    exp, counters = oracle_for(options, data, **extra)
    assert got == exp
    for k, v in counters.items():
        assert t.statistics[k] == v, k
    if variant.startswith("revcomp"):
        assert 0 < t.statistics["reverse_complemented"] < 6000


def flip_records(data: bytes, seed: int) -> bytes:
    from cutadapt_b200.pipeline import reverse_complement

    rng = random.Random(seed)
    out = []
    for name, seq, qual in oracle.parse_fastq(data):
        if rng.random() < 0.5:
            seq, qual = reverse_complement(seq), qual[::-1]
        out.append(f"@{name}\n{seq}\n+\n{qual}\n")
    return "".join(out).encode("latin-1")


def test_revcomp_reference_golden_on_the_device():
    """--revcomp --no-index -g ^TTATTTGTCT -g ^TCCGCACTGG on revcomp.1.fastq (reference test_commandline.py:827-835):
    both orientations, the choice, the in-place replacement and the " rc" suffix inside the FASTQ kernels."""
    import cutadapt_b200.adapters as PA
    from util import fastq_file

    t = FastqTrimmer([PA.PrefixAdapter("TTATTTGTCT", name="a"), PA.PrefixAdapter("TCCGCACTGG", name="b")], revcomp=True)
    assert t.process_chunk(fastq_file("revcomp.in.fastq")) == fastq_file("revcomp.out.fastq")
    assert t.statistics["reverse_complemented"] == 2


def test_many_chunks_in_flight_and_format_errors():
    """process_chunks (two slots, one chunk in flight) gives the concatenation of the per-chunk results."""
    options = dict(adapters=[["back", "AGATCGGAAGAGC"]])
    chunks = [synthetic_fastq(n, seed=50 + i) for i, n in enumerate((1, 3000, 0, 17, 40000, 5, 2500))]
    t = trimmer_for(options, minimum_length=10)
    got = list(t.process_chunks(chunks))
    exp = [oracle_for(options, c, minimum_length=10)[0] for c in chunks]
    assert got == exp
    assert t.statistics["n_records"] == sum((1, 3000, 0, 17, 40000, 5, 2500))
    for bad in (b"@r\nACGT\n+\nIII\n", b"@r\nACGT\n+\n", b"r\nACGT\n+\nIIII\n", b"@r\nACGT\n-\nIIII\n"):
        with pytest.raises(ValueError):
            t.process_chunk(bad)
    # the context stays usable after an error
    assert t.process_chunk(chunks[3]) == exp[3]
    # nothing to trim and no final newline: the output is one byte longer than the input
    plain = b"@r1\nACGTTGCA\n+\nIIIIIIII\n@r2\nTTGACCAT\n+\nIIIIIIII"
    assert FastqTrimmer(None, minimum_length=1).process_chunk(plain) == plain + b"\n"


def test_large_chunk_properties():
    """150 MB of FASTQ (1 M records x 150 bp): size-independent properties + the oracle on a sample of records."""
    n, L = 400_000, 150
    rng = np.random.default_rng(3)
    seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), (n, L))
    ad = np.frombuffer(b"AGATCGGAAGAGCACACGTCTGAACTCCAGTCAC", dtype=np.uint8)
    starts = rng.integers(20, L, n)
    has = rng.random(n) < 0.5
    for j in range(ad.size):
        pos = starts + j
        ok = has & (pos < L)
        seq[np.nonzero(ok)[0], pos[ok]] = ad[j]
    qual = (33 + np.clip(rng.normal(34, 5, (n, L)), 2, 41)).astype(np.uint8)
    recs = []
    for i in range(n):
        recs.append(b"@r%d\n%s\n+\n%s\n" % (i, seq[i].tobytes(), qual[i].tobytes()))
    data = b"".join(recs)
    options = dict(adapters=[["back", "AGATCGGAAGAGC"]], quality_cutoff=[0, 20])
    t = trimmer_for(options, minimum_length=25)
    got = t.process_chunk(data)
    st = t.statistics
    assert st["n_records"] == n and st["bp_in"] == n * L
    assert st["n_written"] + st["too_short"] == n
    lines = got.split(b"\n")
    assert lines[-1] == b"" and len(lines) - 1 == 4 * st["n_written"]
    assert sum(len(x) for x in lines[1::4]) == st["bp_out"]
    assert all(len(a) == len(b) >= 25 for a, b in zip(lines[1::4][:5000], lines[3::4][:5000]))
    # output order = input order; a strided sample goes through the oracle record by record
    idx = list(range(0, n, 1009))
    sample = b"".join(recs[i] for i in idx)
    exp, _ = oracle_for(options, sample, minimum_length=25)
    by_name = {lines[k]: (lines[k + 1], lines[k + 3]) for k in range(0, len(lines) - 1, 4)}
    exp_lines = exp.split(b"\n")
    for k in range(0, len(exp_lines) - 1, 4):
        assert by_name[exp_lines[k]] == (exp_lines[k + 1], exp_lines[k + 3])
    kept_names = {exp_lines[k] for k in range(0, len(exp_lines) - 1, 4)}
    for i in idx:
        assert ((b"@r%d" % i) in by_name) == ((b"@r%d" % i) in kept_names)


# ---- paired-end ----------------------------------------------------------------------------------------

def trimmer_kwargs(options):
    kw = fastq_case_kwargs(options)
    if kw.pop("quality_trim", False):
        kw["quality_cutoff"] = (kw.pop("cutoff_front"), kw.pop("cutoff_back"))
    return kw


def paired_trimmer_for(options):
    return PairedFastqTrimmer(fastq_case_adapters(options, "adapters1"), fastq_case_adapters(options, "adapters2"),
                              trimmer_kwargs(options["options1"]), trimmer_kwargs(options["options2"]),
                              options.get("pair_filter", "any"))


def test_paired_reference_goldens():
    """The paired-end FASTQ known-answer cases of the reference's tests/test_paired.py, byte for byte."""
    for c in fastq_paired_cases():
        t = paired_trimmer_for(c["options"])
        got = t.process_chunk(*c["input_bytes"])
        assert list(got) == c["expected_bytes"], c["name"]
        _, _, c1, c2 = oracle_paired(oracle, c["options"], *c["input_bytes"])
        for st, cc in zip(t.statistics, (c1, c2)):
            for k, v in cc.items():
                assert st[k] == v, (c["name"], k)


@pytest.mark.parametrize("mode", ["any", "both", "first"])
def test_paired_random_chunks_against_oracle(mode):
    data1 = synthetic_fastq(5000, seed=71)
    data2 = synthetic_fastq(5000, seed=72)
    options = dict(
        adapters1=[["back", "AGATCGGAAGAGC"]], adapters2=[["back", "AGATCGGAAGAGC"], ["front", "TTGACNNACG"]],
        pair_filter=mode,
        options1=dict(quality_cutoff=[0, 20], minimum_length=30, max_n=2, cut=[2], poly_a=True, discard_casava=True,
                      max_expected_errors=3.0),
        options2=dict(quality_cutoff=[3, 15], minimum_length=25, maximum_length=200, max_n=2, cut=[-3], poly_a=True,
                      trim_n=True, discard_casava=True, max_expected_errors=3.0))
    t = paired_trimmer_for(options)
    got = t.process_chunk(data1, data2)
    e1, e2, c1, c2 = oracle_paired(oracle, options, data1, data2)
    assert got == (e1, e2)
    for st, cc in zip(t.statistics, (c1, c2)):
        for k, v in cc.items():
            assert st[k] == v, k
    # one-sided adapters with --discard-untrimmed: the pair mode of that filter becomes "both"
    options = dict(adapters1=[["back", "AGATCGGAAGAGC"]], adapters2=[], pair_filter=mode,
                   options1=dict(discard_untrimmed=True, minimum_length=20), options2=dict(discard_untrimmed=True))
    t = paired_trimmer_for(options)
    e1, e2, _, _ = oracle_paired(oracle, options, data1, data2)
    assert t.process_chunk(data1, data2) == (e1, e2)
    with pytest.raises(ValueError):
        t.process_chunk(data1, synthetic_fastq(4999, seed=72))


# ---- demultiplexing ----------------------------------------------------------------------------------

def test_demultiplex_reference_golden_and_barcodes():
    import cutadapt_b200.adapters as PA
    from util import fastq_demux_case

    c = fastq_demux_case()
    ads = [PA.BackAdapter(seq, max_errors=0.1, min_overlap=3, name=name) for name, seq in c["adapters"]]
    t = FastqTrimmer(ads)
    assert t.process_chunk_demux(c["input_bytes"]) == c["expected"]

    # BASELINE config 5 in small: 24 anchored 5' barcodes (device index), several tiles of records, filters on
    rng = random.Random(5)
    barcodes = []
    while len(barcodes) < 24:
        b = "".join(rng.choice("ACGT") for _ in range(10))
        if all(sum(x != y for x, y in zip(b, o)) >= 3 for o in barcodes):
            barcodes.append(b)
    recs = []
    for i in range(20000):
        r = rng.random()
        bc = rng.choice(barcodes) if r < 0.95 else "".join(rng.choice("ACGT") for _ in range(10))
        if rng.random() < 0.05:
            p = rng.randrange(10)
            bc = bc[:p] + rng.choice("ACGT") + bc[p + 1:]
        ins = "".join(rng.choice("ACGT") for _ in range(rng.choice((0, 5, 40, 140))))
        seq = bc + ins
        recs.append(f"@r{i}\n{seq}\n+\n{'F' * len(seq)}\n")
    data = "".join(recs).encode()
    members = [PA.PrefixAdapter(b, max_errors=0.1, name=f"bc{j}") for j, b in enumerate(barcodes)]
    multi = PA.MultipleAdapters(members)                                   # 24 adapters compared one by one
    indexed = PA.MultipleAdapters([PA.IndexedPrefixAdapters([PA.PrefixAdapter(b, max_errors=0.1, name=f"bc{j}")
                                                             for j, b in enumerate(barcodes)])])   # the device index
    for extra in ({}, {"minimum_length": 6, "discard_untrimmed": True}):
        t = FastqTrimmer(multi, **extra)
        got = t.process_chunk_demux(data)
        spec = spec_of(multi)
        exp = oracle.oracle_fastq_demux(data, spec.adapters, spec.groups, [m.name for m in members], **extra)
        assert set(got) == set(exp)
        for k in exp:
            assert got[k] == exp[k], k
        assert sum(len(v) for v in got.values()) == t.statistics["out_bytes"]
        # the same through the device index (its matches may differ from one-by-one alignment, as in the
        # reference): routing must agree with the index's own match records, order inside an output = input order
        ti = FastqTrimmer(indexed, **extra)
        gi = ti.process_chunk_demux(data)
        flat = ti.process_chunk(data)

        def records(b):
            lines = b.split(b"\n")
            return [tuple(lines[k:k + 4]) for k in range(0, len(lines) - 1, 4)]

        assert sorted(r for v in gi.values() for r in records(v)) == sorted(records(flat))
        seqs = [r.split("\n")[1] for r in recs]
        from cutadapt_b200 import _lib as L
        m, _ = indexed.adapter_set().process(*L.pack_strings(seqs))
        where = {}
        for name, v in gi.items():
            ids = [int(r[0][2:]) for r in records(v)]
            assert ids == sorted(ids), name
            where.update({i: name for i in ids})
        for i, name in where.items():
            a = int(m["adapter"][i, 0, 0])
            assert name == (f"bc{a}" if a >= 0 else "unknown"), i

