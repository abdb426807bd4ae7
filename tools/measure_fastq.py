"""End-to-end throughput of the FASTQ entry point (not a bench line).

  python tools/measure_fastq.py [n_reads] [chunk_megabytes]

Builds n_reads synthetic FASTQ records of BASELINE configs[1]'s shape (150 bp, Phred+33 qualities, names
"@SIM2:000000123") in pinned host memory, cuts the buffer into chunks of whole records and streams them
through FastqTrimmer.process_chunks (-a AGATCGGAAGAGC -q 20 -m 20): raw FASTQ bytes in, trimmed FASTQ bytes
out, host -> device -> host inside the timed region.  Prints one JSON line.
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import cutadapt_b200.adapters as PA  # noqa: E402
from cutadapt_b200.pipeline import FastqTrimmer  # noqa: E402
from cutadapt_b200.synth import make_read_tensor  # noqa: E402


def build_fastq(n, pinned=True):
    seq, qual = make_read_tensor(n, config=2, device="cuda", with_qualities=True)
    name_len = 6 + 9
    rec_len = 1 + name_len + 1 + 150 + 3 + 150 + 1
    rec = torch.empty((n, rec_len), dtype=torch.uint8, device="cuda")
    rec[:, 0] = ord("@")
    rec[:, 1:6] = torch.tensor(list(b"SIM2:"), dtype=torch.uint8, device="cuda")
    idx = torch.arange(n, device="cuda")
    for d in range(10):
        rec[:, 6 + 9 - d] = (48 + (idx // 10 ** d) % 10).to(torch.uint8)
    o = 1 + name_len
    rec[:, o] = 10
    rec[:, o + 1:o + 151] = seq
    rec[:, o + 151] = 10
    rec[:, o + 152] = ord("+")
    rec[:, o + 153] = 10
    rec[:, o + 154:o + 304] = qual
    rec[:, o + 304] = 10
    host = torch.empty(n * rec_len, dtype=torch.uint8, pin_memory=pinned)
    host.copy_(rec.view(-1))
    torch.cuda.synchronize()
    return host.numpy(), rec_len


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    chunk_mb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    data, rec_len = build_fastq(n)
    per_chunk = max(1, (chunk_mb << 20) // rec_len)
    chunks = [data[i * rec_len:min(n, i + per_chunk) * rec_len] for i in range(0, n, per_chunk)]
    t = FastqTrimmer([PA.BackAdapter("AGATCGGAAGAGC", max_errors=0.1)], quality_cutoff=(0, 20), minimum_length=20)
    out_bytes = sum(len(o) for o in t.process_chunks(chunks[:9], copy=False))   # warm-up: every slot's buffers, pool
    t.statistics.clear()
    t0 = time.perf_counter()
    out_bytes = 0
    for o in t.process_chunks(chunks, copy=False):
        out_bytes += len(o)
    wall = time.perf_counter() - t0
    st = t.statistics
    print(json.dumps({
        "what": "FASTQ bytes in -> trimmed FASTQ bytes out (-a AGATCGGAAGAGC -q 20 -m 20), host to host",
        "reads": n, "chunk_mb": chunk_mb, "chunks": len(chunks), "reads_per_s": n / wall,
        "in_GB_per_s": data.size / wall / 1e9, "out_GB_per_s": out_bytes / wall / 1e9,
        "in_bytes": int(data.size), "out_bytes": out_bytes, "wall_s": wall,
        "statistics": {k: int(v) for k, v in st.items()},
    }))


if __name__ == "__main__":
    main()
