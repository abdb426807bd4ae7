"""Stress of the banded DP runs (run_band_d, cg_core.cuh) on the host build of the device functions: reads with
several mutated adapter copies at small distances from each other (overlapping, adjacent, a few bases apart), indels,
partial copies at both ends, repetitive adapters -- the bit-plane path (hostsim mode 256) against the oracle.

  python tools/fuzz_band.py [seed0] [n_trials]
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle  # noqa: E402
from util import hostsim_process, spec_of  # noqa: E402
from cutadapt_b200 import _lib as L  # noqa: E402
import cutadapt_b200.adapters as PA  # noqa: E402


def mutate(rng, s, n_edits):
    s = list(s)
    for _ in range(n_edits):
        if not s:
            break
        op = rng.random()
        i = rng.randrange(len(s))
        if op < 0.5:
            s[i] = rng.choice("ACGT")
        elif op < 0.75:
            del s[i]
        else:
            s.insert(i, rng.choice("ACGT"))
    return "".join(s)


def make_read(rng, adapter, length):
    k = max(1, int(len(adapter) * 0.3))
    parts = []
    total = 0
    while total < length:
        r = rng.random()
        if r < 0.45:
            piece = mutate(rng, adapter, rng.choice([0, 0, 1, 1, 2, k]))
        elif r < 0.6:
            a = rng.randrange(len(adapter))
            piece = adapter[a:a + rng.randint(1, len(adapter))]
        elif r < 0.7:
            piece = adapter[: rng.randint(1, len(adapter))]
        else:
            piece = "".join(rng.choice("ACGT") for _ in range(rng.choice([0, 1, 2, 3, 5, 8, 20, 60])))
        parts.append(piece)
        total += len(piece)
    return "".join(parts)[:length]


def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    trials = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    n_reads = 0
    for trial in range(trials):
        rng = random.Random(seed0 * 100003 + trial)
        style = rng.random()
        m = rng.choice([rng.randint(6, 12), 13, 13, rng.randint(14, 33), 33, rng.randint(34, 60)])
        if style < 0.25:        # repetitive
            unit = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 4)))
            seq = (unit * m)[:m]
            seq = mutate(rng, seq, rng.randint(0, 2))[:m] or "ACGTAC"
        else:
            seq = "".join(rng.choice("ACGT") for _ in range(m))
        if len(seq) < 5:
            seq = seq + "ACGTA"
        kw = dict(max_errors=rng.choice([0.05, 0.1, 0.1, 0.15, 0.2, 0.25, 0.3]), min_overlap=rng.randint(1, 6))
        cls = rng.choice([PA.BackAdapter, PA.BackAdapter, PA.BackAdapter, PA.FrontAdapter, PA.AnywhereAdapter])
        ad = cls(seq, name="x", **kw)
        spec = spec_of(ad)
        reads = [make_read(rng, seq, rng.choice([40, 100, 150, 150, 160, 200, 256])) for _ in range(400)]
        params = L.make_params(quality_trim=False)
        exp, _ = oracle.oracle_process(spec.adapters, spec.groups, reads, None, False, 0, 0, 33, 1)
        got, _ = hostsim_process(spec, reads, None, params, 256)
        bad = np.nonzero((got != exp).reshape(len(reads), -1).any(axis=1))[0]
        if len(bad):
            i = int(bad[0])
            print("MISMATCH", repr(ad), kw, reads[i], "\n got", got[i], "\n exp", exp[i])
            sys.exit(1)
        n_reads += len(reads)
    print(f"seed {seed0}: {trials} adapters, {n_reads} reads, no mismatch")


if __name__ == "__main__":
    main()
