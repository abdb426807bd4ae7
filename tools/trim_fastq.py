#!/usr/bin/env python3
"""
Trim FASTQ files on the GPU: a minimal driver around cutadapt_b200.pipeline.FastqTrimmer / PairedFastqTrimmer
that understands the subset of cutadapt's options the device path implements.  Not a replacement for cutadapt's
command line (no reports, no compressed files): it shows the per-chunk worker of INTEGRATION.md section 3 running
on real files.

  python tools/trim_fastq.py -a AGATCGGAAGAGC -q 20 -m 20 -o out.fastq in.fastq
  python tools/trim_fastq.py -a ADAPT1 -A ADAPT2 -q 20 -m 20 -o out.1.fastq -p out.2.fastq in.1.fastq in.2.fastq
  python tools/trim_fastq.py -g bc1=^ACGTACGTAC -g bc2=^TTGCATTGCA -o 'demux-{name}.fastq' in.fastq     (demultiplex)
"""
import argparse
import json
import sys

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import cutadapt_b200.adapters as PA  # noqa: E402
from cutadapt_b200.pipeline import (FastqTrimmer, PairedFastqTrimmer, read_fastq_chunks,  # noqa: E402
                                    read_paired_fastq_chunks)


def make_adapters(specs, kind, error_rate, overlap):
    """-a / -g / -b values: [name=]SEQUENCE with the anchoring characters ^ (5') and $ (3')."""
    out = []
    for i, spec in enumerate(specs or []):
        name, _, seq = spec.rpartition("=")
        name = name or f"{kind}{i + 1}"
        kw = dict(max_errors=error_rate, min_overlap=overlap, name=name)
        if kind == "front" and seq.startswith("^"):
            out.append(PA.PrefixAdapter(seq[1:], **kw))
        elif kind == "back" and seq.endswith("$"):
            out.append(PA.SuffixAdapter(seq[:-1], **kw))
        else:
            out.append({"back": PA.BackAdapter, "front": PA.FrontAdapter, "anywhere": PA.AnywhereAdapter}[kind](seq, **kw))
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    for flag, dest in (("-a", "back"), ("-g", "front"), ("-b", "anywhere"), ("-A", "back2"), ("-G", "front2"),
                       ("-B", "anywhere2")):
        ap.add_argument(flag, dest=dest, action="append")
    ap.add_argument("-e", "--error-rate", type=float, default=0.1)
    ap.add_argument("-O", "--overlap", type=int, default=3)
    ap.add_argument("-n", "--times", type=int, default=1)
    ap.add_argument("-q", "--quality-cutoff", default=None, help="[5'CUTOFF,]3'CUTOFF")
    ap.add_argument("--quality-base", type=int, default=33)
    ap.add_argument("--nextseq-trim", type=int, default=None)
    ap.add_argument("-u", "--cut", type=int, action="append", default=[])
    ap.add_argument("-m", "--minimum-length", type=int, default=0)
    ap.add_argument("-M", "--maximum-length", type=int, default=None)
    ap.add_argument("--max-n", type=float, default=None)
    ap.add_argument("--max-ee", type=float, default=None)
    ap.add_argument("--length", "-l", type=int, default=None)
    ap.add_argument("--poly-a", action="store_true")
    ap.add_argument("--trim-n", action="store_true")
    ap.add_argument("--discard-casava", action="store_true")
    ap.add_argument("--discard-trimmed", action="store_true")
    ap.add_argument("--discard-untrimmed", action="store_true")
    ap.add_argument("--action", default="trim", choices=["trim", "none", "mask", "lowercase", "retain", "crop"])
    ap.add_argument("--pair-filter", default="any", choices=["any", "both", "first"])
    ap.add_argument("--buffer-size", type=int, default=64 << 20)
    ap.add_argument("-o", "--output", required=True, help="output FASTQ; with {name}: one file per adapter name")
    ap.add_argument("-p", "--paired-output")
    ap.add_argument("inputs", nargs="+")
    args = ap.parse_args()

    qc = None
    if args.quality_cutoff is not None:
        parts = [int(x) for x in args.quality_cutoff.split(",")]
        qc = (0, parts[0]) if len(parts) == 1 else (parts[0], parts[1])
    common = dict(times=args.times, quality_cutoff=qc, quality_base=args.quality_base, nextseq_cutoff=args.nextseq_trim,
                  minimum_length=args.minimum_length, maximum_length=args.maximum_length, max_n=args.max_n,
                  max_expected_errors=args.max_ee, discard_trimmed=args.discard_trimmed,
                  discard_untrimmed=args.discard_untrimmed, cut=args.cut, poly_a=args.poly_a, length=args.length,
                  trim_n=args.trim_n, discard_casava=args.discard_casava, action=args.action)
    ads1 = (make_adapters(args.back, "back", args.error_rate, args.overlap)
            + make_adapters(args.front, "front", args.error_rate, args.overlap)
            + make_adapters(args.anywhere, "anywhere", args.error_rate, args.overlap))
    ads2 = (make_adapters(args.back2, "back", args.error_rate, args.overlap)
            + make_adapters(args.front2, "front", args.error_rate, args.overlap)
            + make_adapters(args.anywhere2, "anywhere", args.error_rate, args.overlap))

    if len(args.inputs) == 2:
        if not args.paired_output:
            ap.error("paired-end input needs -p")
        t = PairedFastqTrimmer(ads1, ads2, common, common, args.pair_filter)
        with open(args.inputs[0], "rb") as f1, open(args.inputs[1], "rb") as f2, \
                open(args.output, "wb") as o1, open(args.paired_output, "wb") as o2:
            for c1, c2 in read_paired_fastq_chunks(f1, f2, args.buffer_size):
                r1, r2 = t.process_chunk(c1, c2)
                o1.write(r1)
                o2.write(r2)
        stats = {"read1": t.statistics[0], "read2": t.statistics[1]}
    elif "{name}" in args.output:
        t = FastqTrimmer(ads1, **common)
        files = {}
        with open(args.inputs[0], "rb") as f:
            for chunk in read_fastq_chunks(f, args.buffer_size):
                for name, data in t.process_chunk_demux(chunk).items():
                    if name == "unknown" and args.discard_untrimmed:
                        continue
                    if name not in files:
                        files[name] = open(args.output.replace("{name}", name), "wb")
                    files[name].write(data)
        for fh in files.values():
            fh.close()
        stats = t.statistics
    else:
        t = FastqTrimmer(ads1, **common)
        with open(args.inputs[0], "rb") as f, open(args.output, "wb") as o:
            for out in t.process_chunks(read_fastq_chunks(f, args.buffer_size), copy=False):
                o.write(out.tobytes() if hasattr(out, "tobytes") else out)
        stats = t.statistics
    print(json.dumps(stats), file=sys.stderr)


if __name__ == "__main__":
    main()
