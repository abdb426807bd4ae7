import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
from util import golden
from cutadapt_b200._align import Aligner
cases = golden("locate_kat.json.gz")
by = {}
for ref, q, rate, flags, wr, wq, ic, mo, expected in cases:
    by.setdefault((ref, rate, flags, wr, wq, ic, mo), []).append((q, expected))
for i, (k, items) in enumerate(by.items()):
    print(i, k, [q for q,_ in items][:2], flush=True)
    al = Aligner(*k)
    got = al.locate_batch([q for q, _ in items])
    pass
print("ok")
