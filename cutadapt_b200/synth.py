"""
Synthetic read generator for the benchmark configurations.

150 bp reads, bases uniform over ACGT with 0.1 % N; 50 % of the reads carry the adapter after an
insert of 20..149 bases followed by the TruSeq continuation and poly-A (so that partial 3'
adapters of every length occur); 0.5 %/base substitutions; one random single-base indel inside
the adapter copy in 2 % of the adapter-bearing reads.  Qualities follow
clip(round(N(37 - 12 (i/149)^3, 3)), 2, 41) with a Q2 tail of 5..60 bases in 10 % of the reads.

Written with torch ops so that the same code produces a few thousand reads on the CPU for the
tests and 10^8 reads directly in HBM for bench.py (one generator per shard: seed*1000 + shard).
"""
from typing import List, Optional, Tuple

import numpy as np
import torch

READ_LEN = 150
ADAPTER = "AGATCGGAAGAGC"
CONTINUATION = "ACACGTCTGAACTCCAGTCAC"
BASE_SEED = 20240923


def _codes(s: str) -> torch.Tensor:
    return torch.tensor(list(s.encode("ascii")), dtype=torch.uint8)


@torch.no_grad()
def make_read_tensor(
    n_reads: int,
    config: int = 2,
    shard: int = 0,
    seed: Optional[int] = None,
    device: str = "cpu",
    adapter: str = ADAPTER,
    read_len: int = READ_LEN,
    with_qualities: bool = False,
    chunk: int = 4_000_000,
    out_seq: Optional[torch.Tensor] = None,
    out_qual: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """uint8 tensors [n_reads, read_len] of bases (and Phred+33 qualities)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(((BASE_SEED + config) if seed is None else seed) * 1000 + shard)
    seq = out_seq if out_seq is not None else torch.empty((n_reads, read_len), dtype=torch.uint8, device=dev)
    qual = out_qual
    if with_qualities and qual is None:
        qual = torch.empty((n_reads, read_len), dtype=torch.uint8, device=dev)
    lut = _codes("ACGT").to(dev)
    tail_str = (adapter + CONTINUATION + "A" * read_len)[:read_len]
    tail = _codes(tail_str).to(dev)
    pos = torch.arange(read_len, device=dev)
    for c0 in range(0, n_reads, chunk):
        c1 = min(n_reads, c0 + chunk)
        m = c1 - c0
        bases = lut[torch.randint(0, 4, (m, read_len), generator=gen, device=dev)]
        has_adapter = torch.rand(m, generator=gen, device=dev) < 0.5
        insert = torch.randint(20, read_len, (m,), generator=gen, device=dev)
        # one indel inside the adapter copy for 2 % of the adapter-bearing reads: shift the tail
        # by one from a random adapter position on (deletion) or repeat a base (insertion)
        indel = (torch.rand(m, generator=gen, device=dev) < 0.02) & has_adapter
        indel_pos = torch.randint(1, len(adapter), (m,), generator=gen, device=dev)
        indel_is_del = torch.rand(m, generator=gen, device=dev) < 0.5
        rel = pos[None, :] - insert[:, None]                       # index into the tail
        shift = torch.where(indel_is_del[:, None], 1, -1) * (rel >= indel_pos[:, None]) * indel[:, None]
        src = (rel + shift).clamp_(0, read_len - 1)
        from_tail = (rel >= 0) & has_adapter[:, None]
        bases = torch.where(from_tail, tail[src], bases)
        subst = torch.rand((m, read_len), generator=gen, device=dev) < 0.005
        bases = torch.where(subst, lut[torch.randint(0, 4, (m, read_len), generator=gen, device=dev)], bases)
        n_mask = torch.rand((m, read_len), generator=gen, device=dev) < 0.001
        bases = torch.where(n_mask, torch.full_like(bases, ord("N")), bases)
        seq[c0:c1] = bases
        if with_qualities:
            mean = 37.0 - 12.0 * (pos.float() / (read_len - 1)) ** 3
            q = torch.randn((m, read_len), generator=gen, device=dev) * 3.0 + mean[None, :]
            q = q.round_().clamp_(2, 41)
            low = torch.rand(m, generator=gen, device=dev) < 0.10
            tail_len = torch.randint(5, 61, (m,), generator=gen, device=dev)
            in_tail = (pos[None, :] >= (read_len - tail_len)[:, None]) & low[:, None]
            q = torch.where(in_tail, torch.full_like(q, 2.0), q)
            qual[c0:c1] = (q + 33.0).to(torch.uint8)
    return seq, qual


def make_reads(n_reads: int, config: int = 2, shard: int = 0, seed: Optional[int] = None,
               with_qualities: bool = False, adapter: str = ADAPTER):
    """Small batches as Python strings (tests, smoke, CPU baselines)."""
    seq, qual = make_read_tensor(n_reads, config, shard, seed, "cpu", adapter, with_qualities=with_qualities)
    raw = seq.numpy().tobytes()
    L = seq.shape[1]
    seqs = [raw[i * L : (i + 1) * L].decode("ascii") for i in range(n_reads)]
    quals = None
    if with_qualities:
        qraw = qual.numpy().tobytes()
        quals = [qraw[i * L : (i + 1) * L].decode("ascii") for i in range(n_reads)]
    return seqs, quals
