#!/usr/bin/env python3
"""
Copies the FASTQ known-answer cases of the reference's own command-line tests
(/root/reference/tests/test_commandline.py: run(params, expected, input) compares cutadapt's output
with tests/cut/<expected>) into tests/golden/fastq_kat.json.gz: the (<case>.in.fastq, <case>.out.fastq) pairs as
text plus the case lists, which restate each command line in terms of cutadapt_b200's FASTQ entry point.
These are test vectors (inputs and expected outputs), not source code.

    python tests/golden/make_fastq_golden.py      (needs /root/reference; run once, results committed)
"""
import gzip
import json
import os
import shutil

REF = "/root/reference/tests"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fastq_kat.json.gz")
FILES = {}     # fixture name -> content (latin-1 text)


def put(name, data):
    FILES[name] = (data if isinstance(data, bytes) else data.encode()).decode("latin-1")


def copy(src, name):
    with open(src, "rb") as f:
        put(name, f.read())

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
    ("action_none", 291, "--action=none --discard-untrimmed -a CCCTAGTTAAAC", "small.fastq", "no-trim.fastq",
     dict(adapters=[["back", "CCCTAGTTAAAC"]], action="none", discard_untrimmed=True)),
    ("action_mask", 305, "-b CAAG -n 3 --action=mask", "anywhere_repeat.fastq", "anywhere_repeat.fastq",
     dict(adapters=[["anywhere", "CAAG"]], times=3, action="mask")),
    # FASTA vectors of the reference, stored as FASTQ with constant qualities (single-line records: 1:1)
    ("action_lowercase", 309, "-b CAAG -n 3 --action=lowercase", "action_lowercase.fasta", "action_lowercase.fasta",
     dict(adapters=[["anywhere", "CAAG"]], times=3, action="lowercase")),
    ("action_retain", 317, "-g GGTTAACC -a CAAG --action=retain", "action_retain.fasta", "action_retain.fasta",
     dict(adapters=[["front", "GGTTAACC"], ["back", "CAAG"]], action="retain")),
    ("action_crop", 330, "-g GGTTAA -a CAAG --action=crop --discard-untrimmed", "action_retain.fasta",
     "action_crop.fasta", dict(adapters=[["front", "GGTTAA"], ["back", "CAAG"]], action="crop", discard_untrimmed=True)),
    ("maxee", 838, "--max-ee=0.9", "maxee.fastq", "maxee.fastq",
     dict(adapters=[], max_expected_errors=0.9)),
]


# paired-end: name, reference test (test_paired.py line), command line, in1, in2, expected1, expected2, options
Q10 = dict(quality_cutoff=[0, 10])
Q20 = dict(quality_cutoff=[0, 20])
PAIRED = [
    ("m14", 48, "-a TTAGACATAT -m 14 -q 10", "paired.1.fastq", "paired.2.fastq", "paired.m14.1.fastq", "paired.m14.2.fastq",
     dict(adapters1=[["back", "TTAGACATAT"]], adapters2=[], options1=dict(minimum_length=14, **Q10),
          options2=dict(minimum_length=14, **Q10))),
    ("m27", 260, "-a XXX -m 27", "paired.1.fastq", "paired.2.fastq", "paired-m27.1.fastq", "paired-m27.2.fastq",
     dict(adapters1=[["back", "XXX"]], adapters2=[], options1=dict(minimum_length=27), options2=dict(minimum_length=27))),
    ("both_adapters", 272, "-a TTAGACATAT -A CAGTGGAGTA -m 14", "paired.1.fastq", "paired.2.fastq", "paired.1.fastq",
     "paired.2.fastq", dict(adapters1=[["back", "TTAGACATAT"]], adapters2=[["back", "CAGTGGAGTA"]],
                            options1=dict(minimum_length=14), options2=dict(minimum_length=14))),
    ("qualtrim", 295, "-q 20 -a TTAGACATAT -A CAGTGGAGTA -m 14 -M 90", "paired.1.fastq", "paired.2.fastq",
     "pairedq.1.fastq", "pairedq.2.fastq",
     dict(adapters1=[["back", "TTAGACATAT"]], adapters2=[["back", "CAGTGGAGTA"]],
          options1=dict(minimum_length=14, maximum_length=90, **Q20),
          options2=dict(minimum_length=14, maximum_length=90, **Q20))),
    ("qualtrim_swapped", 307, "-q 20 -a CAGTGGAGTA -A TTAGACATAT -m 14", "paired.2.fastq", "paired.1.fastq",
     "pairedq.2.fastq", "pairedq.1.fastq",
     dict(adapters1=[["back", "CAGTGGAGTA"]], adapters2=[["back", "TTAGACATAT"]],
          options1=dict(minimum_length=14, **Q20), options2=dict(minimum_length=14, **Q20))),
    ("q10_Q0", 325, "-q 10 -Q 0", "lowqual.fastq", "lowqual.fastq", "lowqual.fastq", "lowqual.unchanged.fastq",
     dict(adapters1=[], adapters2=[], options1=dict(**Q10), options2=dict())),
    ("Q10_only", 323, "-Q 10", "lowqual.fastq", "lowqual.fastq", "lowqual.unchanged.fastq", "lowqual.fastq",
     dict(adapters1=[], adapters2=[], options1=dict(), options2=dict(**Q10))),
    ("cut", 340, "-u 3 -u -1 -U 4 -U -2", "paired.1.fastq", "paired.2.fastq", "pairedu.1.fastq", "pairedu.2.fastq",
     dict(adapters1=[], adapters2=[], options1=dict(cut=[3, -1]), options2=dict(cut=[4, -2]))),
    ("length5", 351, "--length 5", "paired.1.fastq", "paired.2.fastq", "length5.1.fastq", "length5.2.fastq",
     dict(adapters1=[], adapters2=[], options1=dict(length=5), options2=dict(length=5))),
    ("length_neg5", 362, "--length -5", "paired.1.fastq", "paired.2.fastq", "length-5.1.fastq", "length-5.2.fastq",
     dict(adapters1=[], adapters2=[], options1=dict(length=-5), options2=dict(length=-5))),
    ("L5_only", 384, "-L 5", "paired.1.fastq", "paired.2.fastq", "paired-unchanged.1.fastq", "length5.2.fastq",
     dict(adapters1=[], adapters2=[], options1=dict(), options2=dict(length=5))),
    ("only_A", 396, "-A CAGTGGAGTA", "paired.1.fastq", "paired.2.fastq", "paired-onlyA.1.fastq", "paired-onlyA.2.fastq",
     dict(adapters1=[], adapters2=[["back", "CAGTGGAGTA"]], options1=dict(), options2=dict())),
    ("discard_untrimmed", 408, "-a CTCCAGCTTAGACATATC -A XXXXXXXX --discard-untrimmed", "paired.1.fastq",
     "paired.2.fastq", "empty.fastq", "empty.fastq",
     dict(adapters1=[["back", "CTCCAGCTTAGACATATC"]], adapters2=[["back", "XXXXXXXX"]],
          options1=dict(discard_untrimmed=True), options2=dict(discard_untrimmed=True))),
    ("discard_trimmed", 419, "-A C -O 1 --discard-trimmed", "paired.1.fastq", "paired.2.fastq", "empty.fastq",
     "empty.fastq", dict(adapters1=[], adapters2=[["back", "C"]], min_overlap=1,
                         options1=dict(discard_trimmed=True), options2=dict(discard_trimmed=True))),
    ("filter_both", 493, "--pair-filter=both -a TTAGACATAT -A GGAGTA -m 14", "paired.1.fastq", "paired.2.fastq",
     "paired-filterboth.1.fastq", "paired-filterboth.2.fastq",
     dict(adapters1=[["back", "TTAGACATAT"]], adapters2=[["back", "GGAGTA"]], pair_filter="both",
          options1=dict(minimum_length=14), options2=dict(minimum_length=14))),
    ("filter_first", 504, "--pair-filter=first -a TTAGACATAT -A GGAGTA -m 14", "paired.1.fastq", "paired.2.fastq",
     "paired-filterfirst.1.fastq", "paired-filterfirst.2.fastq",
     dict(adapters1=[["back", "TTAGACATAT"]], adapters2=[["back", "GGAGTA"]], pair_filter="first",
          options1=dict(minimum_length=14), options2=dict(minimum_length=14))),
]


def fasta_to_fastq(path):
    lines = open(path).read().split("\n")
    recs = []
    for i in range(0, len(lines) - 1, 2):
        assert lines[i].startswith(">") and not lines[i + 1].startswith(">")
        recs.append("@%s\n%s\n+\n%s\n" % (lines[i][1:], lines[i + 1], "I" * len(lines[i + 1])))
    return "".join(recs).encode()


def main():
    index = []
    for name, line, cmd, inp, exp, opts in CASES:
        src = os.path.join(REF, "data", inp)
        if inp.endswith(".gz"):
            with gzip.open(src, "rb") as f:
                put(f"{name}.in.fastq", f.read())
        elif inp.endswith(".fasta"):
            put(f"{name}.in.fastq", fasta_to_fastq(src))
        else:
            copy(src, f"{name}.in.fastq")
        if exp.endswith(".fasta"):
            put(f"{name}.out.fastq", fasta_to_fastq(os.path.join(REF, "cut", exp)))
        else:
            copy(os.path.join(REF, "cut", exp), f"{name}.out.fastq")
        index.append(dict(name=name, reference_test=f"tests/test_commandline.py:{line}", command=cmd,
                          input=f"tests/data/{inp}", expected=f"tests/cut/{exp}", options=opts))
    print(len(index), "cases")
    pindex = []
    for name, line, cmd, in1, in2, exp1, exp2, opts in PAIRED:
        for k, (inp, exp) in enumerate(((in1, exp1), (in2, exp2)), 1):
            copy(os.path.join(REF, "data", inp), f"paired_{name}.in{k}.fastq")
            copy(os.path.join(REF, "cut", exp), f"paired_{name}.out{k}.fastq")
        pindex.append(dict(name=name, reference_test=f"tests/test_paired.py:{line}", command=cmd,
                           inputs=[f"tests/data/{in1}", f"tests/data/{in2}"],
                           expected=[f"tests/cut/{exp1}", f"tests/cut/{exp2}"], options=opts))
    print(len(pindex), "paired cases")
    # demultiplexing (test_commandline.py:581-601: -a first=AATTTCAGGAATT -a second=GTTCTCTAGTTCT -o {name}.fasta
    # twoadapters.fasta): the reference's vectors are FASTA; they are stored as FASTQ with constant qualities 'I'
    # (single-line records, so the conversion is 1:1 and the expected sequences are untouched)
    put("demux_twoadapters.in.fastq", fasta_to_fastq(os.path.join(REF, "data", "twoadapters.fasta")))
    for name in ("first", "second", "unknown"):
        put(f"demux_twoadapters.{name}.out.fastq", fasta_to_fastq(os.path.join(REF, "cut", f"twoadapters.{name}.fasta")))
    # --info-file known answers (tests/test_info_file.py:14-55): the text files
    for src, dst in (("cut/illumina.info.txt", "info_illumina.txt"), ("cut/illumina5.info.txt", "info_illumina5.txt"),
                     ("data/illumina5.fastq", "info_illumina5.in.fastq")):
        copy(os.path.join(REF, src), dst)
    # --rest-file / --wildcard-file known answers (test_commandline.py:110-122, 345-367)
    for src, dst in (("data/rest.fa", "rest.in.fasta"), ("data/rest.txt", "rest.txt"),
                     ("data/restfront.txt", "restfront.txt"), ("data/wildcard_adapter.fa", "wildcard_adapter.in.fasta")):
        copy(os.path.join(REF, src), dst)
    with gzip.open(OUT, "wt", compresslevel=9) as f:
        json.dump(dict(cases=index, paired_cases=pindex, files=FILES), f)
    print(len(FILES), "fixture files ->", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
