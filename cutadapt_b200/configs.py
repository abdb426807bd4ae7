"""
The benchmark configurations of BASELINE.json (configs[1..4] = "config 2..5" of SURVEY.md section 8(d)): adapter
sets, parameters and the synthetic read generators, shared by bench.py, the parity tests and the tools.

    2  100 M x 150 bp SE, one 3' adapter AGATCGGAAGAGC, e = 0.1
    3  100 M x 150 bp SE, five adapters (two plain, one IUPAC, one with an N run, one linked), e = 0.15;
       every read is generated from the construct of one of the five adapters (20 % each)
    4  50 M pairs 2 x 150 bp, 3' adapters on R1 / R2, quality trimming -q 20; R2 = reverse complement of the
       insert + the R2 adapter, independent quality strings
    5  demultiplexing: 96 anchored 5' barcodes of length 10 (pairwise Hamming distance >= 3), e = 0.1 with indels
       (edit-environment index); read = barcode (1 % with one substitution) + random insert, 2 % without barcode

Everything is generated with torch ops (reproducible per shard: seed * 1000 + shard) so that the same code gives a
few thousand reads on the CPU for the tests and 10^8 reads directly in HBM for bench.py.
"""
from typing import Dict, List, Optional

import torch

from .synth import BASE_SEED, CONTINUATION, READ_LEN, make_read_tensor

CONFIG2_ADAPTER = "AGATCGGAAGAGC"
CONFIG3_BACK = ["AGATCGGAAGAGC", "CTGTCTCTTATACACATCT", "VCCGAMCYUCKHRKDCUBBCNUWNSGHCGU",
                "AGATCGGAAGAGCNNNNNNNNATCTCGTATGCC"]
CONFIG3_LINKED = ("GTTCAGAGTTCTACAGTCCGACGATC", "TGGAATTCTCGGGTGCCAAGG")
CONFIG4_R1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA"
CONFIG4_R2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT"
CONFIG_ERROR_RATE = {2: 0.1, 3: 0.15, 4: 0.1, 5: 0.1}

_IUPAC = {"A": "A", "C": "C", "G": "G", "T": "T", "U": "T", "R": "AG", "Y": "CT", "S": "GC", "W": "AT", "K": "GT",
          "M": "AC", "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT"}


def config5_barcodes(n: int = 96, length: int = 10, seed: int = BASE_SEED + 5) -> List[str]:
    """`n` barcodes with pairwise Hamming distance >= 3, drawn with a seeded generator (rejection sampling)."""
    import random

    rng = random.Random(seed)
    out: List[str] = []
    while len(out) < n:
        cand = "".join(rng.choice("ACGT") for _ in range(length))
        if all(sum(a != b for a, b in zip(cand, o)) >= 3 for o in out):
            out.append(cand)
    return out


def config_adapters(config: int):
    """The adapters of a configuration as cutadapt_b200 objects: (adapters for R1, adapters for R2 or None)."""
    from . import adapters as PA

    e = CONFIG_ERROR_RATE[config]
    if config == 2:
        return PA.MultipleAdapters([PA.BackAdapter(CONFIG2_ADAPTER, max_errors=e, min_overlap=3, name="adapter")]), None
    if config == 3:
        objs = [PA.BackAdapter(s, max_errors=e, min_overlap=3, name=f"a{i}") for i, s in enumerate(CONFIG3_BACK)]
        objs.append(PA.LinkedAdapter(PA.PrefixAdapter(CONFIG3_LINKED[0], max_errors=e, min_overlap=3, name="lf"),
                                     PA.BackAdapter(CONFIG3_LINKED[1], max_errors=e, min_overlap=3, name="lb"),
                                     True, False, "linked"))
        return PA.MultipleAdapters(objs), None
    if config == 4:
        return (PA.MultipleAdapters([PA.BackAdapter(CONFIG4_R1, max_errors=e, min_overlap=3, name="r1")]),
                PA.MultipleAdapters([PA.BackAdapter(CONFIG4_R2, max_errors=e, min_overlap=3, name="r2")]))
    if config == 5:
        pre = [PA.PrefixAdapter(b, max_errors=e, min_overlap=3, indels=True, name=f"bc{i}")
               for i, b in enumerate(config5_barcodes())]
        return PA.MultipleAdapters([PA.IndexedPrefixAdapters(pre)]), None
    raise ValueError(config)


def _codes(s: str, dev) -> torch.Tensor:
    return torch.tensor(list(s.encode("ascii")), dtype=torch.uint8, device=dev)


def _iupac_tables(dev):
    """pick[c, r]: the r-th (cyclically) base the IUPAC letter with ASCII code c stands for."""
    pick = torch.zeros((128, 4), dtype=torch.uint8, device=dev)
    for c, alts in _IUPAC.items():
        for r in range(4):
            pick[ord(c), r] = ord(alts[r % len(alts)])
    return pick


@torch.no_grad()
def make_config_batch(config: int, n: int, shard: int = 0, device: str = "cpu", seed: Optional[int] = None,
                      read_len: int = READ_LEN, chunk: int = 4_000_000) -> Dict[str, Optional[torch.Tensor]]:
    """
    {"seq", "qual", "seq2", "qual2"}: uint8 tensors [n, read_len] (None where the configuration has none).
    For config 4, n is the number of PAIRS.
    """
    dev = torch.device(device)
    if config == 2:
        seq, _ = make_read_tensor(n, 2, shard, seed, device, CONFIG2_ADAPTER, read_len)
        return {"seq": seq, "qual": None, "seq2": None, "qual2": None}
    gen = torch.Generator(device=dev)
    gen.manual_seed(((BASE_SEED + config) if seed is None else seed) * 1000 + shard)
    lut = _codes("ACGT", dev)
    pos = torch.arange(read_len, device=dev)

    def rnd_bases(m, width=read_len):
        return lut[torch.randint(0, 4, (m, width), generator=gen, device=dev)]

    def noise(bases):
        m = bases.shape[0]
        subst = torch.rand((m, read_len), generator=gen, device=dev) < 0.005
        bases = torch.where(subst, rnd_bases(m), bases)
        n_mask = torch.rand((m, read_len), generator=gen, device=dev) < 0.001
        return torch.where(n_mask, torch.full_like(bases, ord("N")), bases)

    def qualities(m):
        mean = 37.0 - 12.0 * (pos.float() / (read_len - 1)) ** 3
        q = torch.randn((m, read_len), generator=gen, device=dev) * 3.0 + mean[None, :]
        q = q.round_().clamp_(2, 41)
        low = torch.rand(m, generator=gen, device=dev) < 0.10
        tail_len = torch.randint(5, 61, (m,), generator=gen, device=dev)
        in_tail = (pos[None, :] >= (read_len - tail_len)[:, None]) & low[:, None]
        return (torch.where(in_tail, torch.full_like(q, 2.0), q) + 33.0).to(torch.uint8)

    out = {"seq": torch.empty((n, read_len), dtype=torch.uint8, device=dev), "qual": None, "seq2": None, "qual2": None}
    if config == 4:
        out["qual"] = torch.empty((n, read_len), dtype=torch.uint8, device=dev)
        out["seq2"] = torch.empty((n, read_len), dtype=torch.uint8, device=dev)
        out["qual2"] = torch.empty((n, read_len), dtype=torch.uint8, device=dev)

    if config == 3:
        pick = _iupac_tables(dev)
        # tails of the five constructs: adapter + continuation + poly-A, IUPAC letters kept (resolved per read)
        tails = torch.stack([_codes((a + CONTINUATION + "A" * read_len)[:read_len], dev)
                             for a in CONFIG3_BACK + [CONFIG3_LINKED[1]]])
        front = _codes(CONFIG3_LINKED[0], dev)
    if config == 5:
        bcs = config5_barcodes()
        bc = torch.stack([_codes(b, dev) for b in bcs])

    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        m = c1 - c0
        if config == 3:
            bases = rnd_bases(m)
            construct = torch.randint(0, 5, (m,), generator=gen, device=dev)
            has_adapter = torch.rand(m, generator=gen, device=dev) < 0.5
            insert = torch.randint(20, read_len, (m,), generator=gen, device=dev)
            rel = pos[None, :] - insert[:, None]
            src = rel.clamp(0, read_len - 1)
            tail = tails[construct[:, None], src]                          # letters incl. IUPAC
            r = torch.randint(0, 4, (m, read_len), generator=gen, device=dev)
            tail = pick[tail.long(), r]                                    # one concrete instance per read
            bases = torch.where((rel >= 0) & has_adapter[:, None], tail, bases)
            # the linked construct starts with the anchored 5' adapter (90 % of its reads)
            linked = (construct == 4) & (torch.rand(m, generator=gen, device=dev) < 0.9)
            fpos = pos[None, : front.numel()].expand(m, -1)
            head = torch.where(linked[:, None], front[fpos], bases[:, : front.numel()])
            bases = torch.cat([head, bases[:, front.numel():]], dim=1)
            out["seq"][c0:c1] = noise(bases)
        elif config == 4:
            has_adapter = torch.rand(m, generator=gen, device=dev) < 0.5
            insert = torch.randint(20, read_len, (m,), generator=gen, device=dev)
            frag = rnd_bases(m)                                            # the insert (its first `insert` bases)
            rel = pos[None, :] - insert[:, None]
            src = rel.clamp(0, read_len - 1)
            for key, adapter, rc in (("seq", CONFIG4_R1, False), ("seq2", CONFIG4_R2, True)):
                tail = _codes((adapter + "A" * read_len)[:read_len], dev)
                if rc:
                    # reverse complement of the insert: base j of R2 = complement(frag[insert - 1 - j])
                    j = (insert[:, None] - 1 - pos[None, :]).clamp(0, read_len - 1)
                    comp = torch.zeros(128, dtype=torch.uint8, device=dev)
                    for a, b in zip("ACGT", "TGCA"):
                        comp[ord(a)] = ord(b)
                    mate = comp[torch.gather(frag, 1, j).long()]
                    mate = torch.where(has_adapter[:, None], mate, rnd_bases(m))   # long fragments: independent mate
                else:
                    mate = frag
                mate = torch.where((rel >= 0) & has_adapter[:, None], tail[src], mate)
                out[key][c0:c1] = noise(mate)
            out["qual"][c0:c1] = qualities(m)
            out["qual2"][c0:c1] = qualities(m)
        elif config == 5:
            bases = rnd_bases(m)
            which = torch.randint(0, bc.shape[0], (m,), generator=gen, device=dev)
            code = bc[which]                                               # [m, 10]
            err = torch.rand(m, generator=gen, device=dev) < 0.01
            epos = torch.randint(0, code.shape[1], (m,), generator=gen, device=dev)
            sub = lut[torch.randint(0, 4, (m,), generator=gen, device=dev)]
            code = torch.where(err[:, None] & (pos[None, : code.shape[1]] == epos[:, None]), sub[:, None], code)
            none = torch.rand(m, generator=gen, device=dev) < 0.02
            head = torch.where(none[:, None], bases[:, : code.shape[1]], code)
            out["seq"][c0:c1] = torch.cat([head, bases[:, code.shape[1]:]], dim=1)
        else:
            raise ValueError(config)
    return out


def to_strings(t: torch.Tensor) -> List[str]:
    raw = t.cpu().numpy().tobytes()
    L = t.shape[1]
    return [raw[i * L:(i + 1) * L].decode("ascii") for i in range(t.shape[0])]
