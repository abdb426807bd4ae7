#!/usr/bin/env python3
"""
Copies the FASTQ known-answer cases of the reference's own command-line tests
(/root/reference/tests/test_commandline.py: run(params, expected, input) compares cutadapt's output
with tests/cut/<expected>) into tests/golden/fastq/ as (<case>.in.fastq, <case>.out.fastq) pairs plus
cases.json, which restates each command line in terms of cutadapt_b200's FASTQ entry point.
These are test vectors (inputs and expected outputs), not source code.

    python tests/golden/make_fastq_golden.py      (needs /root/reference; run once, results committed)
"""
import gzip
import json
import os
import shutil

REF = "/root/reference/tests"
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fastq")

# name, reference test (test_commandline.py line), command line, input, expected, options for the product
CASES = [
    ("small", 79, "-a TTAGACATATCTCCGTCG", "small.fastq", "small.fastq",
     dict(adapters=[["back", "TTAGACATATCTCCGTCG"]])),
    ("empty", 91, "-a TTAGACATATCTCCGTCG", "empty.fastq", "empty.fastq",
     dict(adapters=[["back", "TTAGACATATCTCCGTCG"]])),
    ("dos", 104, "-e 0.12 -a TTAGACATATCTCCGTCG", "dos.fastq", "dos.fastq",
     dict(adapters=[["back", "TTAGACATATCTCCGTCG"]], error_rate=0.12)),
    ("lowercase", 109, "-a ttagacatatctccgtcg", "small.fastq", "lowercase.fastq",
     dict(adapters=[["back", "ttagacatatctccgtcg"]])),
    ("discard", 129, "-b TTAGACATATCTCCGTCG --discard", "small.fastq", "discard.fastq",
     dict(adapters=[["anywhere", "TTAGACATATCTCCGTCG"]], discard_trimmed=True)),
    ("discard_untrimmed", 134, "-b CAAGAT --discard-untrimmed", "small.fastq", "discard-untrimmed.fastq",
     dict(adapters=[["anywhere", "CAAGAT"]], discard_untrimmed=True)),
    ("lowqual", 248, "-q 10 -a XXXXXX", "lowqual.fastq", "lowqual.fastq",
     dict(adapters=[["back", "XXXXXX"]], quality_cutoff=[0, 10])),
    ("illumina64", 253, "-q 10 --quality-base 64 -a XXXXXX", "illumina64.fastq", "illumina64.fastq",
     dict(adapters=[["back", "XXXXXX"]], quality_cutoff=[0, 10], quality_base=64)),
    ("illumina64_no_adapter", 258, "-q 10 --quality-base 64", "illumina64.fastq", "illumina64.fastq",
     dict(adapters=[], quality_cutoff=[0, 10], quality_base=64)),
    ("iupac", 376, "-a VCCGAMCYUCKHRKDCUBBCNUWNSGHCGU", "illumina.fastq.gz", "illumina.fastq",
     dict(adapters=[["back", "VCCGAMCYUCKHRKDCUBBCNUWNSGHCGU"]])),
    ("rna", 456, "-a GCCGAACUUCUUAGACUGCCUUAAGGACGU", "illumina.fastq.gz", "illumina.fastq",
     dict(adapters=[["back", "GCCGAACUUCUUAGACUGCCUUAAGGACGU"]])),
    ("nextseq", 667, "--nextseq-trim 22", "nextseq.fastq", "nextseq.fastq",
     dict(adapters=[], nextseq_cutoff=22)),
    ("anywhere_small", 780, "-b TTAGACATATCTCCGTCG", "small.fastq", "small.fastq",
     dict(adapters=[["anywhere", "TTAGACATATCTCCGTCG"]])),
    ("paired_separate_1", 785, "-a TTAGACATAT", "paired.1.fastq", "paired-separate.1.fastq",
     dict(adapters=[["back", "TTAGACATAT"]])),
    ("paired_separate_2", 786, "-a CAGTGGAGTA", "paired.2.fastq", "paired-separate.2.fastq",
     dict(adapters=[["back", "CAGTGGAGTA"]])),
    ("front_empty", 790, "-g CWC", "empty.fastq", "empty.fastq",
     dict(adapters=[["front", "CWC"]])),
    ("unconditional_front", 502, "-u 5", "small.fastq", "unconditional-front.fastq", dict(adapters=[], cut=[5])),
    ("unconditional_back", 506, "-u -5", "small.fastq", "unconditional-back.fastq", dict(adapters=[], cut=[-5])),
    ("unconditional_both", 510, "-u -5 -u 5", "small.fastq", "unconditional-both.fastq",
     dict(adapters=[], cut=[-5, 5])),
    ("shortened", 738, "--length 5", "small.fastq", "shortened.fastq", dict(adapters=[], length=5)),
    ("shortened_negative", 742, "--length -5", "small.fastq", "shortened-negative.fastq",
     dict(adapters=[], length=-5)),
    ("casava", 770, "--discard-casava", "casava.fastq", "casava.fastq", dict(adapters=[], discard_casava=True)),
    ("maxee", 838, "--max-ee=0.9", "maxee.fastq", "maxee.fastq",
     dict(adapters=[], max_expected_errors=0.9)),
]


def main():
    os.makedirs(HERE, exist_ok=True)
    index = []
    for name, line, cmd, inp, exp, opts in CASES:
        src = os.path.join(REF, "data", inp)
        dst_in = os.path.join(HERE, f"{name}.in.fastq")
        if inp.endswith(".gz"):
            with gzip.open(src, "rb") as f, open(dst_in, "wb") as g:
                g.write(f.read())
        else:
            shutil.copyfile(src, dst_in)
        shutil.copyfile(os.path.join(REF, "cut", exp), os.path.join(HERE, f"{name}.out.fastq"))
        index.append(dict(name=name, reference_test=f"tests/test_commandline.py:{line}", command=cmd,
                          input=f"tests/data/{inp}", expected=f"tests/cut/{exp}", options=opts))
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(index, f, indent=1)
    print(len(index), "cases")


if __name__ == "__main__":
    main()
