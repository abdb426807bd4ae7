"""Index lookups (IndexedPrefixAdapters / IndexedSuffixAdapters) of the host build against the oracle's index: random barcode
sets (equal and mixed lengths, with and without indels), mutated barcodes with N, lower case and other letters.

  python tools/fuzz_index.py seed trials
"""
import sys, random
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import oracle
from util import hostsim_process, spec_of
import cutadapt_b200.adapters as PA
rng=random.Random(int(sys.argv[1])); tot=0
for trial in range(int(sys.argv[2])):
    n_bc=rng.randint(2,12); Ls=[rng.randint(4,14) for _ in range(n_bc)]
    if rng.random()<0.5: Ls=[Ls[0]]*n_bc
    bcs=[]
    while len(bcs)<n_bc:
        b="".join(rng.choice("ACGT") for _ in range(Ls[len(bcs)]))
        if b not in bcs: bcs.append(b)
    prefix=rng.random()<0.5; indels=rng.random()<0.5; rate=rng.choice([0.0,0.1,0.1,0.2])
    cls=PA.PrefixAdapter if prefix else PA.SuffixAdapter
    ads=[cls(b,max_errors=rate,indels=indels,name=f"b{i}") for i,b in enumerate(bcs)]
    try:
        multi=(PA.IndexedPrefixAdapters if prefix else PA.IndexedSuffixAdapters)(ads)
    except Exception as e:
        continue
    spec=spec_of(PA.MultipleAdapters([multi]) if not isinstance(multi, PA.Matchable) else multi)
    reads=[]
    for _ in range(300):
        b=list(rng.choice(bcs))
        for _ in range(rng.choice([0,0,1,1,2])):
            op=rng.random(); i=rng.randrange(len(b)) if b else 0
            if not b: break
            if op<0.6: b[i]=rng.choice("ACGTNacgtnX")
            elif op<0.8: del b[i]
            else: b.insert(i,rng.choice("ACGT"))
        body="".join(rng.choice("ACGTN") for _ in range(rng.choice([0,1,3,10,40])))
        reads.append(("".join(b)+body) if prefix else (body+"".join(b)))
    data=np.frombuffer("".join(reads).encode(),dtype=np.uint8)
    offsets=np.zeros(len(reads)+1,dtype=np.int64); offsets[1:]=np.cumsum([len(r) for r in reads])
    exp=oracle.oracle_index_process(bcs,rate,indels,prefix,data,offsets,[a.descriptor() for a in ads])
    got,_=hostsim_process(spec,reads)
    g=got[:,0,0]
    for f in ("adapter","astart","astop","rstart","rstop","score","errors"):
        if not (g[f]==exp[f]).all():
            i=int(np.nonzero(g[f]!=exp[f])[0][0]); print("MISMATCH",f,reads[i],g[i],exp[i],bcs,prefix,indels,rate); sys.exit(1)
    tot+=int((exp["adapter"]>=0).sum())
print("ok, hits",tot)
