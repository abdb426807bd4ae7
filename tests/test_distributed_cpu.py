"""
The N>1 path on CPU: two gloo ranks shard the reads, trim their shard (device functions compiled for
the host -- test infrastructure), reduce the statistics vector with the product's
``allreduce_statistics`` and must arrive at the single-process totals.
"""
import socket

import numpy as np


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_range_partitions_exactly():
    from cutadapt_b200.pipeline import shard_range

    for n in (0, 1, 7, 128, 1000003):
        for world in (1, 2, 3, 8):
            bounds = [shard_range(n, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in bounds]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_statistics_allreduce(tmp_path):
    import torch.multiprocessing as mp
    import _dist_worker
    import cutadapt_b200.adapters as PA
    from cutadapt_b200.synth import make_reads
    from util import hostsim_process, spec_of

    mp.spawn(_dist_worker.run, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert (r0 == r1).all()
    assert (np.load(tmp_path / "local0.npy") + np.load(tmp_path / "local1.npy") == r0).all()
    # single-process reference of the same totals
    reads, _ = make_reads(3001, config=2, seed=99)
    spec = spec_of(PA.MultipleAdapters([PA.BackAdapter("AGATCGGAAGAGC", max_errors=0.1, name="a")]))
    matches, _ = hostsim_process(spec, reads)
    total = _dist_worker.host_statistics(matches, None, np.array([len(r) for r in reads]), 1, 150, 3)
    assert (total == r0).all()
    assert r0[0] == 3001 and 1300 < r0[2] < 1700
    # the FASTQ counters of both ranks add up to the single-process totals
    import json
    from oracle import oracle

    f0, f1 = (json.load(open(tmp_path / f"fq{r}.json")) for r in (0, 1))
    assert f0["total"] == f1["total"]
    fq = "".join(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads)).encode()
    _, whole = oracle.oracle_fastq_trim(fq, spec.adapters, spec.groups, minimum_length=100)
    for k, v in whole.items():
        assert f0["total"][k] == v == f0["local"][k] + f1["local"][k], k
    assert whole["n_written"] + whole["too_short"] == 3001 and whole["too_short"] > 100


def test_kept_intervals_compose_like_the_reference():
    """kept_intervals == applying QualityTrimmer then every Match.trimmed() (adapters.py:453-487)."""
    import random
    import cutadapt_b200.adapters as PA
    from cutadapt_b200 import _lib as L
    from cutadapt_b200.pipeline import kept_intervals
    from util import hostsim_process, spec_of, random_reads

    rng = random.Random(4)
    ads = ["ACGTTGCAAC", "TTGACCAGTA"]
    multi = PA.MultipleAdapters([PA.BackAdapter(ads[0], name="a"), PA.FrontAdapter(ads[1], name="b")])
    spec = spec_of(multi)
    reads = random_reads(rng, ads, 300, "ACGT", 90)
    quals = ["".join(chr(33 + rng.choice([2, 20, 35])) for _ in r) for r in reads]
    params = L.make_params(quality_trim=True, cutoff_front=0, cutoff_back=15, times=2)
    matches, qtrim = hostsim_process(spec, reads, quals, params)
    iv = kept_intervals(matches, qtrim, np.array([len(r) for r in reads]))
    for i, read in enumerate(reads):
        cur = read[qtrim[i, 0]:qtrim[i, 1]]
        for r in range(2):
            m = multi.matches_from_records(matches[i, r], cur)
            if m is None:
                break
            cur = m.trimmed(cur)
        assert read[iv[i, 0]:iv[i, 1]] == cur
