"""Adapters with N runs (UMIs) and IUPAC codes, reads crowded with their instances: every host-sim schedule against the oracle.

  python tools/fuzz_wildcard_adapters.py seed trials
"""
import sys, random
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import oracle
from util import hostsim_process, spec_of
from cutadapt_b200 import _lib as L
import cutadapt_b200.adapters as PA
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import fuzz_band as fb
seed=int(sys.argv[1])
tot=0
for trial in range(int(sys.argv[2])):
    rng=random.Random(seed*7919+trial)
    a=rng.randint(5,16); b=rng.randint(3,12); c=rng.randint(5,16)
    seq="".join(rng.choice("ACGT") for _ in range(a))+"N"*b+"".join(rng.choice("ACGT") for _ in range(c))
    if rng.random()<0.3: seq="".join(ch if rng.random()>0.15 else rng.choice("RYKMSWN") for ch in seq)
    cls=rng.choice([PA.BackAdapter,PA.BackAdapter,PA.AnywhereAdapter,PA.FrontAdapter])
    ad=cls(seq,name="x",max_errors=rng.choice([0.1,0.15,0.15,0.2,0.3]),min_overlap=rng.randint(1,6))
    spec=spec_of(ad)
    def inst(s): 
        m={'N':'ACGT','R':'AG','Y':'CT','K':'GT','M':'AC','S':'GC','W':'AT'}
        return "".join(rng.choice(m.get(ch,ch)) for ch in s)
    reads=[]
    for _ in range(300):
        reads.append(fb.make_read(rng, inst(seq), rng.choice([60,150,150,200,256])))
    exp,_=oracle.oracle_process(spec.adapters,spec.groups,reads,None,False,0,0,33,1)
    for mode in (0,2,10,64):
        got,_=hostsim_process(spec,reads,None,L.make_params(quality_trim=False),mode)
        bad=np.nonzero((got!=exp).reshape(len(reads),-1).any(axis=1))[0]
        if len(bad):
            i=int(bad[0]); print("MISMATCH",mode,repr(ad),reads[i],got[i],exp[i]); sys.exit(1)
    tot+=int((exp["adapter"]>=0).sum())
print("seed",seed,"ok, matches",tot)
