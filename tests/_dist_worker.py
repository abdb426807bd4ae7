"""Worker of tests/test_distributed_cpu.py (gloo, world_size 2; CPU only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def host_statistics(reads, matches, qtrim, n_adapters, max_len, kmax):
    """The statistics vector of a shard: the host build of the function cg_stats_kernel runs per read."""
    from util import hostsim_statistics

    return hostsim_statistics(reads, matches, qtrim, n_adapters, max_len, kmax)


def workload():
    """3001 config-2 reads plus reads for a 5' adapter, an anywhere adapter and a linked adapter (lower case and N too)."""
    import random
    import cutadapt_b200.adapters as PA
    from cutadapt_b200.synth import make_reads
    from util import random_reads

    reads, _ = make_reads(3001, config=2, seed=99)
    rng = random.Random(7)
    extra = ["GTTCAGAGTTCTACAGTCCGACGATC", "TGGAATTCTCGGGTGCCAAGG", "CTGTCTCTTATACACATCT", "ACGTTGCATTGAC"]
    reads += random_reads(rng, extra, 1500, "ACGTNacgt", 120)
    multi = PA.MultipleAdapters([
        PA.BackAdapter("AGATCGGAAGAGC", max_errors=0.1, name="a"),
        PA.FrontAdapter("ACGTTGCATTGAC", max_errors=0.1, name="f"),
        PA.AnywhereAdapter("CTGTCTCTTATACACATCT", max_errors=0.1, name="w"),
        PA.LinkedAdapter(PA.PrefixAdapter("GTTCAGAGTTCTACAGTCCGACGATC", max_errors=0.1, name="lf"),
                         PA.BackAdapter("TGGAATTCTCGGGTGCCAAGG", max_errors=0.1, name="lb"), False, False, "linked"),
    ])
    return reads, multi


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
    reads, multi = workload()                               # same data on every rank ...
    lo, hi = shard_range(len(reads), rank, world)           # ... each trims its own shard
    spec = spec_of(multi)
    matches, _ = hostsim_process(spec, reads[lo:hi])
    local = host_statistics(reads[lo:hi], matches, None, len(spec.adapters), 200, 3)
    t = torch.from_numpy(local.copy())
    allreduce_statistics(t)
    # FASTQ-level counters: every rank "trims" its shard of a FASTQ file with the oracle; the totals are all-reduced
    import json
    from cutadapt_b200.pipeline import allreduce_fastq_statistics
    from oracle import oracle

    reads2, _ = make_reads(3001, config=2, seed=99)
    lo2, hi2 = shard_range(len(reads2), rank, world)
    spec2 = spec_of(PA.MultipleAdapters([PA.BackAdapter("AGATCGGAAGAGC", max_errors=0.1, name="a")]))
    fq = "".join(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads2[lo2:hi2])).encode()
    _, counters = oracle.oracle_fastq_trim(fq, spec2.adapters, spec2.groups, minimum_length=100)
    counters["out_bytes"] = 0
    with open(os.path.join(out_dir, f"fq{rank}.json"), "w") as f:
        json.dump({"local": counters, "total": allreduce_fastq_statistics(counters)}, f)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), t.numpy())
    np.save(os.path.join(out_dir, f"local{rank}.npy"), local)
    dist.barrier()
    dist.destroy_process_group()


def chunk_workload():
    """A FASTQ file of 6000 config-2 reads and the per-chunk function of the CPU tests (the FASTQ oracle)."""
    import cutadapt_b200.adapters as PA
    from cutadapt_b200.synth import make_reads
    from oracle import oracle
    from util import spec_of

    reads, quals = make_reads(6000, config=2, seed=5, with_qualities=True)
    data = "".join(f"@r{i} x\n{r}\n+\n{q}\n" for i, (r, q) in enumerate(zip(reads, quals))).encode()
    spec = spec_of(PA.MultipleAdapters([PA.BackAdapter("AGATCGGAAGAGC", max_errors=0.1, name="a")]))

    def process(chunk):
        return oracle.oracle_fastq_trim(bytes(chunk), spec.adapters, spec.groups, quality_trim=True, cutoff_back=20,
                                        minimum_length=30)[0]

    return data, process


def run_chunks(rank, world, port, out_dir):
    """RoundRobinRunner with the oracle as the per-chunk worker: the protocol (round-robin, ordered merge) on CPU."""
    import io
    import torch.distributed as dist
    from cutadapt_b200.runners import RoundRobinRunner

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data, process = chunk_workload()
    runner = RoundRobinRunner(None, buffer_size=24 * 1024, process_chunk=process)
    sink = io.BytesIO() if rank == 0 else None
    total = runner.run(data, sink)
    if rank == 0:
        with open(os.path.join(out_dir, "merged.fastq"), "wb") as f:
            f.write(sink.getvalue())
        with open(os.path.join(out_dir, "chunks.txt"), "w") as f:
            f.write(str(total["chunks"]))
    dist.barrier()
    dist.destroy_process_group()


def gpu_chunk_workload(n=120_000):
    from cutadapt_b200.synth import make_reads

    reads, quals = make_reads(n, config=2, seed=21, with_qualities=True)
    return "".join(f"@r{i} x\n{r}\n+\n{q}\n" for i, (r, q) in enumerate(zip(reads, quals))).encode()


def run_gpu_chunks(rank, world, port, out_dir):
    """RoundRobinRunner with one GPU per rank (NCCL for the statistics, gloo for the ordered output merge)."""
    import io
    import json
    import torch
    import torch.distributed as dist
    import cutadapt_b200.adapters as PA
    from cutadapt_b200 import _lib
    from cutadapt_b200.pipeline import FastqTrimmer
    from cutadapt_b200.runners import RoundRobinRunner

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    data = gpu_chunk_workload()
    trimmer = FastqTrimmer([PA.BackAdapter("AGATCGGAAGAGC", max_errors=0.1, name="a")], quality_cutoff=(0, 20),
                           minimum_length=30, ctx=_lib.Context(rank))
    runner = RoundRobinRunner(trimmer, buffer_size=256 * 1024)
    sink = io.BytesIO() if rank == 0 else None
    total = runner.run(data, sink)
    if rank == 0:
        with open(os.path.join(out_dir, "merged.fastq"), "wb") as f:
            f.write(sink.getvalue())
        with open(os.path.join(out_dir, "total.json"), "w") as f:
            json.dump(total, f)
    dist.barrier()
    dist.destroy_process_group()
