"""
Multi-GPU chunk driver (cutadapt_b200.runners.RoundRobinRunner): chunk c is trimmed by GPU c mod G, rank 0 writes the
outputs in chunk order.  Needs at least two GPUs on the box (`gpurun --gpus 2`); skipped otherwise.
"""
import io
import json
import socket

import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_gpus_write_what_one_gpu_writes(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import _dist_worker
    import cutadapt_b200.adapters as PA
    from cutadapt_b200.pipeline import FastqTrimmer, read_fastq_chunks
    from cutadapt_b200.runners import SerialRunner
    from oracle import oracle
    from util import spec_of

    world = min(4, torch.cuda.device_count())
    mp.spawn(_dist_worker.run_gpu_chunks, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    merged = open(tmp_path / "merged.fastq", "rb").read()
    total = json.load(open(tmp_path / "total.json"))
    data = _dist_worker.gpu_chunk_workload()
    n_chunks = len(list(read_fastq_chunks(io.BytesIO(data), 256 * 1024)))
    assert n_chunks >= 64 and total["chunks"] == n_chunks
    # one GPU, same chunks
    trimmer = FastqTrimmer([PA.BackAdapter("AGATCGGAAGAGC", max_errors=0.1, name="a")], quality_cutoff=(0, 20), minimum_length=30)
    sink = io.BytesIO()
    single = SerialRunner(trimmer, buffer_size=256 * 1024).run(data, sink)
    assert merged == sink.getvalue()
    for key in ("n_records", "n_written", "bp_in", "bp_out", "with_adapters", "quality_trimmed_bp", "too_short"):
        assert total[key] == single[key], key
    # and the oracle on the whole file
    spec = spec_of(PA.MultipleAdapters([PA.BackAdapter("AGATCGGAAGAGC", max_errors=0.1, name="a")]))
    assert merged == oracle.oracle_fastq_trim(data, spec.adapters, spec.groups, quality_trim=True, cutoff_back=20, minimum_length=30)[0]
