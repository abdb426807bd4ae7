"""Resident-throughput of the other BASELINE configurations (not bench lines).

  python tools/measure_configs.py [n_reads] [case-name prefix]

Prints one JSON line per configuration: reads/s of DeviceBatch.run on reads resident in HBM,
CUDA-event timed, plus the fraction of reads with a match.
"""
import json
import random
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import cutadapt_b200.adapters as PA  # noqa: E402
from cutadapt_b200.pipeline import DeviceBatch  # noqa: E402
from cutadapt_b200.synth import make_read_tensor  # noqa: E402


def timed(batch, seq, offsets, qual, reps=3):
    n = offsets.numel() - 1
    out = torch.empty((n * batch.times * batch.adapter_set.slots, 8), dtype=torch.int32, device=seq.device)
    qt = torch.empty((n, 2), dtype=torch.int32, device=seq.device)
    for _ in range(1 if reps == 1 else 2):
        batch.run(seq, offsets, qual, max_read_len=150, out=out, qtrim_out=qt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        res = batch.run(seq, offsets, qual, max_read_len=150, out=out, qtrim_out=qt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    hit = (out.view(n, -1, 8)[:, :, 0] >= 0).any(dim=1).float().mean().item()
    return n / ms * 1e3, hit


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    dev = "cuda"
    offsets = torch.arange(0, (n + 1) * 150, 150, dtype=torch.int64, device=dev)
    results = []

    seq, qual = make_read_tensor(n, config=2, device=dev, with_qualities=True)
    pad = torch.zeros(64, dtype=torch.uint8, device=dev)
    flat = torch.cat([seq.view(-1), pad])
    qflat = torch.cat([qual.view(-1), pad])

    cases = {}
    cases["config2: one 3' adapter AGATCGGAAGAGC e=0.1"] = (PA.BackAdapter("AGATCGGAAGAGC", max_errors=0.1), None, None)
    five = [PA.BackAdapter(s, max_errors=0.15, name=f"a{i}") for i, s in enumerate(
        ["AGATCGGAAGAGC", "CTGTCTCTTATACACATCT", "VCCGAMCYUCKHRKDCUBBCNUWNSGHCGU", "AGATCGGAAGAGCNNNNNNNNATCTCGTATGCC"])]
    five.append(PA.LinkedAdapter(PA.PrefixAdapter("GTTCAGAGTTCTACAGTCCGACGATC", max_errors=0.15, name="f"),
                                 PA.BackAdapter("TGGAATTCTCGGGTGCCAAGG", max_errors=0.15, name="b"), True, False, "l"))
    cases["config3: 5 adapters incl. IUPAC + linked e=0.15"] = (PA.MultipleAdapters(five), None, None)
    cases["config4 (per mate): -q 20 + 33-mer 3' adapter e=0.1"] = (
        PA.BackAdapter("AGATCGGAAGAGCACACGTCTGAACTCCAGTCA", max_errors=0.1), (0, 20), qflat)

    rng = random.Random(96)
    barcodes = set()
    while len(barcodes) < 96:
        barcodes.add("".join(rng.choice("ACGT") for _ in range(10)))
    barcodes = sorted(barcodes)
    pre = [PA.PrefixAdapter(b, max_errors=1, indels=False, name=f"bc{i}") for i, b in enumerate(barcodes)]
    bc = torch.tensor([list(b.encode()) for b in barcodes], dtype=torch.uint8, device=dev)
    demux = seq.clone()
    demux[:, :10] = bc[torch.randint(0, 96, (n,), device=dev)]
    mism = torch.rand(n, device=dev) < 0.2
    col = torch.randint(0, 10, (n,), device=dev)
    rows = torch.nonzero(mism).squeeze(1)
    demux[rows, col[rows]] = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (rows.numel(),), device=dev)]
    dflat = torch.cat([demux.view(-1), pad])
    cases["config5: 96 anchored 5' barcodes, device index"] = (PA.IndexedPrefixAdapters(pre), None, None, dflat)
    cases["config5 without index: 96 PrefixAdapters one by one"] = (PA.MultipleAdapters(pre), None, None, dflat)

    only = sys.argv[2] if len(sys.argv) > 2 else None
    for name, c in cases.items():
        if only and not name.startswith(only):
            continue
        ad, qc, q = c[0], c[1], c[2]
        data = c[3] if len(c) > 3 else flat
        batch = DeviceBatch(ad, quality_cutoff=qc)
        rps, hit = timed(batch, data, offsets, q, reps=1 if only else 3)
        line = {"case": name, "reads": n, "reads_per_s": rps, "matched_frac": round(hit, 4)}
        print(json.dumps(line), flush=True)
        results.append(line)


if __name__ == "__main__":
    main()
