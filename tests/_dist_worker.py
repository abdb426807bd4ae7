"""Worker of tests/test_distributed_cpu.py (gloo, world_size 2; CPU only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def host_statistics(matches, qtrim, lengths, n_adapters, max_len, kmax):
    """numpy restatement of cg_stats_kernel's vector layout (include/cutadapt_b200.h)."""
    from cutadapt_b200.pipeline import stats_layout

    lay = stats_layout(n_adapters, max_len, kmax)
    v = np.zeros(lay["size"], dtype=np.int64)
    v[0] = len(lengths)
    v[1] = int(lengths.sum())
    v[2] = int((matches["adapter"] >= 0).any(axis=(1, 2)).sum())
    if qtrim is not None:
        v[3] = int((lengths - (qtrim[:, 1] - qtrim[:, 0])).sum())
    hist = v[lay["hist"]:].reshape(lay["shape"])
    for rec in matches.reshape(-1):
        if rec["adapter"] < 0:
            continue
        searched = (int(rec["info"]) >> 16) & 0xFFFF
        removed = searched - rec["rstart"] if (int(rec["info"]) >> 8) & 1 else rec["rstop"]
        v[4] += removed
        hist[rec["adapter"], min(max(removed, 0), max_len), min(max(rec["errors"], 0), kmax)] += 1
    return v


def run(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    import cutadapt_b200.adapters as PA
    from cutadapt_b200.pipeline import shard_range, allreduce_statistics
    from cutadapt_b200.synth import make_reads
    from util import hostsim_process, spec_of

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    reads, _ = make_reads(3001, config=2, seed=99)          # same data on every rank ...
    lo, hi = shard_range(len(reads), rank, world)           # ... each trims its own shard
    spec = spec_of(PA.MultipleAdapters([PA.BackAdapter("AGATCGGAAGAGC", max_errors=0.1, name="a")]))
    matches, _ = hostsim_process(spec, reads[lo:hi])
    lengths = np.array([len(r) for r in reads[lo:hi]], dtype=np.int64)
    local = host_statistics(matches, None, lengths, 1, 150, 3)
    t = torch.from_numpy(local.copy())
    allreduce_statistics(t)
    # FASTQ-level counters: every rank "trims" its shard of a FASTQ file with the oracle; the totals are all-reduced
    import json
    from cutadapt_b200.pipeline import allreduce_fastq_statistics
    from oracle import oracle

    fq = "".join(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads[lo:hi])).encode()
    _, counters = oracle.oracle_fastq_trim(fq, spec.adapters, spec.groups, minimum_length=100)
    counters["out_bytes"] = 0
    with open(os.path.join(out_dir, f"fq{rank}.json"), "w") as f:
        json.dump({"local": counters, "total": allreduce_fastq_statistics(counters)}, f)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), t.numpy())
    np.save(os.path.join(out_dir, f"local{rank}.npy"), local)
    dist.barrier()
    dist.destroy_process_group()
