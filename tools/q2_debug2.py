"""Debug aid: is the sequence-pair kernel's disagreement tied to the long-tile path? (GPU box only)"""
import sys
import numpy as np
sys.path.insert(0, ".")
from swimm_b200 import gpu as g, host, synth

rng = np.random.default_rng(11)
q = synth.make_queries(rng, [144, 189, 850, 1000, 2005, 2504])
db = synth.make_db(10, 40_000, queries=q)
_, dl, dc = synth.length_sorted(db)
_, ql, qc = synth.length_sorted(q)
qo = np.zeros(q.n + 1, np.uint32)
np.cumsum(ql.astype(np.uint32), out=qo[1:])
b62 = host.submat("blosum62")
s = g.GpuSearch(0)
s.load_db(dl, dc)
s.set_option("query_pairing", 2)
truth, _ = s.search(qc, ql, qo[:-1], b62, 10, 2, 0, want_scores=True)
s.set_option("query_pairing", 0)
print("max db len", dl.max())
for name, thr, fg, fr in [("auto", 0, 0, 0), ("auto", 0, 0, 0), ("no-long", 65000, 0, 0), ("no-long", 65000, 0, 0), ("long>3072", 3072, 0, 0),
                          ("long>512", 512, 0, 0), ("no-long K=32", 65000, 32, 32), ("no-long K=20", 65000, 32, 20)]:
    s.set_option("long_threshold", thr)
    s.set_option("force_group", fg)
    s.set_option("force_rows", fr)
    got, _ = s.search(qc, ql, qo[:-1], b62, 10, 2, 0, want_scores=True)
    print(name, "mismatches per query", [int((got[i] != truth[i]).sum()) for i in range(q.n)], "launches", s.stats()["launches"])
