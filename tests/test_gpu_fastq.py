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
    for bad in (b"@r\nACGT\n+\nIII\n", b"@r\nACGT\n+\n", b"r\nACGT\n+\nIIII\n", b"@r\nACGT\n-\nIIII\n",
                b"@r 1\nACGT\n+r 2\nIIII\n"):
        with pytest.raises(ValueError):
            t.process_chunk(bad)
        with pytest.raises(oracle.FastqFormatError):
            oracle.parse_fastq(bad)
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



# ---- --pair-adapters and paired demultiplexing on the device --------------------------------------------

def _single_spec(adapter):
    import cutadapt_b200.adapters as PA

    spec = spec_of(PA.MultipleAdapters([adapter]))
    return spec.adapters, spec.groups


def test_pair_adapters_reference_golden_on_the_device():
    """--pair-adapters -a GTCTCCAGCT -A GACAAATAAC (reference test_paired.py:668-676) through
    cg_fastq_collect_pair_adapters."""
    import cutadapt_b200.adapters as PA
    from util import fastq_file

    t = PairedFastqTrimmer([PA.BackAdapter("GTCTCCAGCT", name="a")], [PA.BackAdapter("GACAAATAAC", name="b")],
                           pair_adapters=True)
    got = t.process_chunk(fastq_file("pair_adapters.in1.fastq"), fastq_file("pair_adapters.in2.fastq"))
    assert got == (fastq_file("pair_adapters.out1.fastq"), fastq_file("pair_adapters.out2.fastq"))
    with pytest.raises(ValueError):
        PairedFastqTrimmer([PA.BackAdapter("GTCTCCAGCT")], [], pair_adapters=True)


@pytest.mark.parametrize("variant", ["trim", "mask_quality", "retain_linked"])
def test_pair_adapters_random_chunks_against_oracle(variant):
    import cutadapt_b200.adapters as PA

    rng = random.Random(91)
    firsts = ["AGATCGGAAGAGC", "CTGTCTCTTATACAC", "TGGAATTCTCGGGTGCC"]
    seconds = ["AGATCGGAAGAGC", "GACAAATAACGGT", "CACGTCTGAACTC"]
    n = 4000
    recs1, recs2 = [], []
    for i in range(n):
        def read(adapter):
            ln = rng.choice((40, 90, 150))
            seq = "".join(rng.choice("ACGT") for _ in range(ln))
            if adapter is not None:
                cut = rng.randrange(5, ln - 3)
                ad = "".join(c if rng.random() > 0.04 else rng.choice("ACGT") for c in adapter)
                seq = (seq[:cut] + ad + seq)[:ln]
            qual = "".join(chr(33 + min(41, max(2, int(rng.gauss(34 - 25 * (j / ln) ** 2, 6))))) for j in range(ln))
            return seq, qual
        k = rng.randrange(len(firsts))
        r = rng.random()
        # both mates carry pair k / different pairs / only one mate / none
        a1 = firsts[k] if r < 0.8 else None
        a2 = seconds[k] if r < 0.5 else (seconds[(k + 1) % 3] if r < 0.65 else (None if r < 0.9 else seconds[k]))
        (s1, q1), (s2, q2) = read(a1), read(a2)
        recs1.append(f"@p{i}/1\n{s1}\n+\n{q1}\n")
        recs2.append(f"@p{i}/2\n{s2}\n+\n{q2}\n")
    data1, data2 = "".join(recs1).encode(), "".join(recs2).encode()
    if variant == "retain_linked":
        ads1 = [PA.LinkedAdapter(PA.PrefixAdapter("ACGTAC", name="f"), PA.BackAdapter(firsts[0], name="b"), False, True, "l0"),
                PA.BackAdapter(firsts[1], name="a1"), PA.BackAdapter(firsts[2], name="a2")]
        o1 = dict(action="retain", minimum_length=5)
        o2 = dict(action="retain", minimum_length=5)
    else:
        ads1 = [PA.BackAdapter(s, name=f"a{i}") for i, s in enumerate(firsts)]
        o1 = dict(minimum_length=20, discard_untrimmed=True) if variant == "trim" else \
            dict(action="mask", quality_cutoff=(0, 20), cut=[1])
        o2 = dict(minimum_length=20, discard_untrimmed=True) if variant == "trim" else \
            dict(action="mask", quality_cutoff=(5, 15), nextseq_cutoff=None)
    ads2 = [PA.BackAdapter(s, name=f"b{i}") for i, s in enumerate(seconds)]
    t = PairedFastqTrimmer(ads1, ads2, o1, o2, "any", pair_adapters=True)
    got = t.process_chunk(data1, data2)

    def okw(o):
        o = dict(o)
        if "quality_cutoff" in o:
            c = o.pop("quality_cutoff")
            o.update(quality_trim=True, cutoff_front=c[0], cutoff_back=c[1])
        return o
    e1, e2, c1, c2 = oracle.oracle_fastq_trim_paired(
        data1, data2, options1=okw(o1), options2=okw(o2), pair_filter="any",
        pair_specs=[(_single_spec(a), _single_spec(b)) for a, b in zip(ads1, ads2)])
    assert got == (e1, e2)
    for st, cc in zip(t.statistics, (c1, c2)):
        for k, v in cc.items():
            assert st[k] == v, k
    assert 0 < t.statistics[0]["with_adapters"] < n


@pytest.mark.parametrize("combinatorial", [False, True])
@pytest.mark.parametrize("discard_untrimmed", [False, True])
def test_paired_demultiplexing_against_oracle(combinatorial, discard_untrimmed):
    """PairedDemultiplexer / CombinatorialDemultiplexer (steps.py:422-581): both mates partitioned by the adapter of the
    most recent match of R1 (or of both mates) on the device, every output in input order."""
    import cutadapt_b200.adapters as PA

    rng = random.Random(17 + combinatorial)
    bc1 = ["ACGTACGTAC", "TTGCAAGGTC", "GGATCCTTAA", "CATGCATGGA"]
    bc2 = ["TCAGTCAGTC", "AACCGGTTAC", "GTGTGACACA"]
    n = 7000
    recs1, recs2 = [], []
    for i in range(n):
        def read(barcodes):
            seq = "".join(rng.choice("ACGT") for _ in range(rng.choice((30, 60, 100))))
            if rng.random() < 0.85:
                b = rng.choice(barcodes)
                if rng.random() < 0.2:
                    j = rng.randrange(len(b))
                    b = b[:j] + rng.choice("ACGT") + b[j + 1:]
                seq = b + seq
            return seq
        s1, s2 = read(bc1), read(bc2)
        recs1.append(f"@p{i}/1\n{s1}\n+\n{'I' * len(s1)}\n")
        recs2.append(f"@p{i}/2\n{s2}\n+\n{'F' * len(s2)}\n")
    data1, data2 = "".join(recs1).encode(), "".join(recs2).encode()
    ads1 = [PA.PrefixAdapter(b, max_errors=1, name=f"x{i}", indels=False) for i, b in enumerate(bc1)]
    ads2 = [PA.PrefixAdapter(b, max_errors=1, name=f"y{i}", indels=False) for i, b in enumerate(bc2)]
    o = dict(minimum_length=35)
    t = PairedFastqTrimmer(ads1, ads2, o, o, "any")
    got = t.process_chunk_demux(data1, data2, combinatorial=combinatorial, discard_untrimmed=discard_untrimmed)
    names1, names2 = [a.name for a in ads1], [a.name for a in ads2]

    def route(last1, last2):
        k1 = names1[last1] if last1 >= 0 else None
        k2 = names2[last2] if last2 >= 0 else None
        if combinatorial:
            return None if discard_untrimmed and (k1 is None or k2 is None) else (k1, k2)
        if k1 is None:
            return None if discard_untrimmed else "unknown"
        return k1
    s1 = spec_of(PA.MultipleAdapters(ads1))
    s2 = spec_of(PA.MultipleAdapters(ads2))
    e1, e2, c1, c2 = oracle.oracle_fastq_trim_paired(data1, data2, s1.adapters, s1.groups, s2.adapters, s2.groups, o, o, "any",
                                                     route=route)
    assert set(e1) <= set(got)
    for key, (g1, g2) in got.items():
        assert g1 == e1.get(key, b"") and g2 == e2.get(key, b""), key
    for st, cc in zip(t.statistics, (c1, c2)):
        for k, v in cc.items():
            assert st[k] == v, k


# ---- --info-file rows on the device ------------------------------------------------------------------------

def test_info_file_reference_goldens_on_the_device():
    """tests/cut/illumina.info.txt and illumina5.info.txt (--times 2) of the reference (tests/test_info_file.py:14-55):
    the rows come formatted from the device (cg_fastq_collect_info)."""
    import cutadapt_b200.adapters as PA
    from util import fastq_file

    cases = [("iupac.in.fastq", "info_illumina.txt", [("adapt", "GCCGAACTTCTTAGACTGCCTTAAGGACGT")], 1),
             ("info_illumina5.in.fastq", "info_illumina5.txt", [("adapt", "GCCGAACTTCTTA"), ("adapt2", "GACTGCCTTAAGGACGT")], 2)]
    for fastq, expected, ads, times in cases:
        t = FastqTrimmer([PA.BackAdapter(s, max_errors=0.1, min_overlap=3, name=n) for n, s in ads], times=times)
        _, info = t.process_chunk_info(fastq_file(fastq))
        want = [w.rstrip() for w in fastq_file(expected).decode().split("\n")]
        got = [g.rstrip() for g in info.decode("latin-1").split("\n")]
        assert got[-1] == "" and got == want                        # assert_files_equal(ignore_trailing_space)


@pytest.mark.parametrize("variant", ["plain", "times_linked", "revcomp_cuts", "lowercase_filters"])
def test_info_rows_random_chunks_against_oracle(variant):
    import cutadapt_b200.adapters as PA

    seed = {"plain": 31, "times_linked": 32, "revcomp_cuts": 33, "lowercase_filters": 34}[variant]
    data = synthetic_fastq(5000, seed=seed)
    options = dict(adapters=[["back", "AGATCGGAAGAGC"], ["front", "TTGACNNACG"]], quality_cutoff=[5, 20])
    extra = {}
    if variant == "times_linked":
        options = dict(adapters=[["linked", "TTGACNNACG", "AGATCGGAAGAGC"], ["back", "CACGTCTGAACTC"],
                                 ["anywhere", "ACGTACGTAC"]], quality_cutoff=[0, 15])
        extra = dict(times=3, poly_a=True)
    elif variant == "revcomp_cuts":
        data = flip_records(data, seed)
        extra = dict(revcomp=True, cut=[3, -2], nextseq_cutoff=12, times=2, trim_n=True)
    elif variant == "lowercase_filters":
        extra = dict(action="lowercase", minimum_length=40, discard_trimmed=True, length=70)
    t = trimmer_for(options, **extra)
    out, info = t.process_chunk_info(data)
    ads = fastq_case_adapters(options)
    singles, groups, owners = PA.MultipleAdapters(ads)._flatten()
    names = [s.name for s in singles]
    for (typ, a0, a1, _, _), owner in zip(groups, owners):
        if typ == 1:
            names[a0], names[a1] = owner.name + ";1", owner.name + ";2"
    rows = []
    exp, counters = oracle_for(options, data, info_names=names, info_rows=rows, **extra)
    assert out == exp
    assert info.decode("latin-1") == "".join(r + "\n" for r in rows)
    assert len(rows) >= 5000


def test_rest_and_wildcard_rows_on_the_device():
    """--rest-file / --wildcard-file rows formatted on the device (cg_fastq_collect_rows): the reference's known answers
    (tests/data/rest.txt, restfront.txt, test_adapter_wildcard; test_commandline.py:110-122, 345-367) and randomized
    chunks against the oracle (several rounds, quality trimming in front, --revcomp)."""
    import cutadapt_b200.adapters as PA
    from util import fastq_file

    def fasta_as_fastq(name):
        lines = [l for l in fastq_file(name).decode().split("\n") if l]
        return "".join(f"@{h[1:]}\n{s}\n+\n{'I' * len(s)}\n" for h, s in zip(lines[0::2], lines[1::2])).encode()

    data = fasta_as_fastq("rest.in.fasta")
    for cls, expected in ((PA.AnywhereAdapter, "rest.txt"), (PA.FrontAdapter, "restfront.txt")):
        t = FastqTrimmer([cls("ADAPTER", max_errors=0.1, min_overlap=3, adapter_wildcards=False, name="a")])
        _, rows = t.process_chunk_rest(data)
        assert rows == fastq_file(expected), expected
    data = fasta_as_fastq("wildcard_adapter.in.fasta")
    for cls in (PA.BackAdapter, PA.AnywhereAdapter):
        t = FastqTrimmer([cls("ACGTNNNACGT", max_errors=0.1, min_overlap=3, name="a")])
        _, rows = t.process_chunk_wildcards(data)
        assert rows == b"AAA 1\nGGG 2\nCCC 3b\nTTT 4b\n"
    # randomized
    for seed, extra in ((41, dict(times=2)), (42, dict(revcomp=True, cut=[2]))):
        chunk = synthetic_fastq(4000, seed=seed)
        if extra.get("revcomp"):
            chunk = flip_records(chunk, seed)
        options = dict(adapters=[["back", "AGATCGGAAGAGC"], ["front", "TTGACNNACG"], ["anywhere", "CACGTNTGAAC"]],
                       quality_cutoff=[5, 20])
        ads = fastq_case_adapters(options)
        seqs = [s.sequence for s in PA.MultipleAdapters(ads)._flatten()[0]]
        rest, wild = [], []
        exp, _ = oracle_for(options, chunk, rest_rows=rest, wildcard_rows=wild, adapter_sequences=seqs, **extra)
        t = trimmer_for(options, **extra)
        out, rows = t.process_chunk_rest(chunk)
        assert out == exp and rows.decode("latin-1") == "".join(r + "\n" for r in rest)
        out, rows = trimmer_for(options, **extra).process_chunk_wildcards(chunk)
        assert out == exp and rows.decode("latin-1") == "".join(r + "\n" for r in wild)
        assert len(rest) > 500 and len(wild) > 1000
