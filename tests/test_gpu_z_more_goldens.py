"""
GPU tests (-m gpu) that run last: further known answers of the reference's command-line tests through the device
path (FASTA-vector goldens of test_commandline.py, --revcomp, --pair-adapters).  Byte-identical output is the bar.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cutadapt_b200.pipeline import FastqTrimmer  # noqa: E402
from oracle import oracle  # noqa: E402
from test_gpu_fastq import trimmer_kwargs  # noqa: E402


def test_reference_fasta_goldens():
    """32 further command-line known answers of the reference whose vectors are FASTA (anchored / non-internal / linked
    adapters, --no-indels, -N, --match-read-wildcards, --trim-n, --poly-a, --max-n; stored as FASTQ with constant
    qualities, tests/golden/make_fastq_golden.py): byte-identical output of the device path."""
    from util import golden, fastq_file, adapter_from_spec

    cases = golden("fastq_kat.json.gz")["fasta_cases"]
    assert len(cases) >= 30
    for c in cases:
        o = c["options"]
        params = dict(max_errors=o.get("error_rate", 0.1), min_overlap=o.get("min_overlap", 3),
                      adapter_wildcards=not o.get("no_wildcards", False), read_wildcards=o.get("read_wildcards", False),
                      indels=not o.get("no_indels", False))
        ads = [adapter_from_spec(spec, kind, name=f"a{i}", **params) for i, (kind, spec) in enumerate(o["specs"])]
        t = FastqTrimmer(ads, **trimmer_kwargs(o))
        assert t.process_chunk(fastq_file(f"fa_{c['name']}.in.fastq")) == fastq_file(f"fa_{c['name']}.out.fastq"), \
            (c["name"], c["command"])


def test_revcomp_and_pair_adapters_compositions():
    """--revcomp and --pair-adapters as compositions of device passes (BatchTrimmer.process_revcomp,
    PairedAdapterBatch) on the reference's known answers (test_commandline.py:827-835, test_paired.py:668-676)."""
    import cutadapt_b200.adapters as PA
    from cutadapt_b200 import pipeline
    from util import fastq_file

    records = oracle.parse_fastq(fastq_file("revcomp.in.fastq"))
    names, seqs, quals = zip(*records)
    bt = pipeline.BatchTrimmer([PA.PrefixAdapter("TTATTTGTCT", name="a"), PA.PrefixAdapter("TCCGCACTGG", name="b")])
    res, is_rc = bt.process_revcomp(list(seqs))
    assert int(is_rc.sum()) == 2
    out = []
    for i, name in enumerate(names):
        s, q = (pipeline.reverse_complement(seqs[i]), quals[i][::-1]) if is_rc[i] else (seqs[i], quals[i])
        a, b = (int(x) for x in res.intervals[i])
        out.append(f"@{name}{' rc' if is_rc[i] else ''}\n{s[a:b]}\n+\n{q[a:b]}\n")
    assert "".join(out).encode() == fastq_file("revcomp.out.fastq")

    recs = [oracle.parse_fastq(fastq_file(f"pair_adapters.in{k}.fastq")) for k in (1, 2)]
    pb = pipeline.PairedAdapterBatch([PA.BackAdapter("GTCTCCAGCT", name="a")], [PA.BackAdapter("GACAAATAAC", name="b")])
    best, t1, t2 = pb.process([r[1] for r in recs[0]], [r[1] for r in recs[1]])
    for rec, t, k in ((recs[0], t1, 1), (recs[1], t2, 2)):
        text = "".join(f"@{n}\n{s[a:b]}\n+\n{q[a:b]}\n" for (n, s, q), (a, b) in zip(rec, t.intervals.tolist()))
        assert text.encode() == fastq_file(f"pair_adapters.out{k}.fastq")


def test_paired_revcomp_composition():
    """--revcomp on pairs (PairedReverseComplementer, modifiers.py:311-400) as four device passes + paired_revcomp_select:
    the reference's known answers with adapters on R1 only, on R2 only and on both (test_paired.py:786-833)."""
    import cutadapt_b200.adapters as PA
    from cutadapt_b200 import pipeline
    from util import fastq_file

    rec1, rec2 = (oracle.parse_fastq(fastq_file(f)) for f in ("revcomp.in.fastq", "revcomp.in2.fastq"))

    def written(swapped, t1, t2, first, second):
        out = []
        for (a, b), res in (((first, second), t1), ((second, first), t2)):
            lines = []
            for i in range(len(first)):
                name, s, q = (b if swapped[i] else a)[i]
                lo, hi = (int(x) for x in res.intervals[i]) if res is not None else (0, len(s))
                lines.append(f"@{name}{' rc' if swapped[i] else ''}\n{s[lo:hi]}\n+\n{q[lo:hi]}\n")
            out.append("".join(lines).encode())
        return out

    both = [PA.PrefixAdapter("TTATTTGTCT", name="a"), PA.PrefixAdapter("TCCGCACTGGC", name="b")]
    one = [fastq_file("revcomp_one_mate.out1.fastq"), fastq_file("revcomp_one_mate.out2.fastq")]
    swapped, t1, t2 = pipeline.PairedRevcompBatch(both, None).process([r[1] for r in rec1], [r[1] for r in rec2])
    assert written(swapped, t1, t2, rec1, rec2) == one
    swapped, t1, t2 = pipeline.PairedRevcompBatch(None, both).process([r[1] for r in rec2], [r[1] for r in rec1])
    assert written(swapped, t1, t2, rec2, rec1) == one[::-1]
    swapped, t1, t2 = pipeline.PairedRevcompBatch(both[:1], both[1:]).process([r[1] for r in rec1], [r[1] for r in rec2])
    assert int(swapped.sum()) == 2
    assert written(swapped, t1, t2, rec1, rec2) == [fastq_file("revcomp_r1r2.out1.fastq"), fastq_file("revcomp_r1r2.out2.fastq")]
