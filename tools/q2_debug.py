"""Debug aid: where do the query-pair kernel and the sequence-pair kernel disagree? (GPU box only)"""
import sys
import numpy as np
sys.path.insert(0, ".")
from swimm_b200 import gpu as g, host, synth
from tests.helpers import load_oracle

rng = np.random.default_rng(11)
q = synth.make_queries(rng, [144, 189, 850, 1000, 2005, 2504])
db = synth.make_db(10, int(sys.argv[1]) if len(sys.argv) > 1 else 40_000, queries=q)
_, dl, dc = synth.length_sorted(db)
_, ql, qc = synth.length_sorted(q)
qo = np.zeros(q.n + 1, np.uint32)
np.cumsum(ql.astype(np.uint32), out=qo[1:])
b62 = host.submat("blosum62")
s = g.GpuSearch(0)
s.load_db(dl, dc)
s.set_option("query_pairing", 0)
single, _ = s.search(qc, ql, qo[:-1], b62, 10, 2, 0, want_scores=True)
do = np.zeros(db.n + 1, np.uint64)
np.cumsum(dl.astype(np.uint64), out=do[1:])
orc = load_oracle()
for G, K in [(0, 0), (32, 8)]:
    s.set_option("query_pairing", 2)
    s.set_option("q2_group", G)
    s.set_option("q2_rows", K)
    paired, _ = s.search(qc, ql, qo[:-1], b62, 10, 2, 0, want_scores=True)
    bad = np.argwhere(paired != single)
    print("shape", G, K, "mismatches", len(bad), "per query", [int((paired[i] != single[i]).sum()) for i in range(q.n)])
    for qi, si in bad[:12]:
        print("   q", qi, "len", ql[qi], "seq", si, "dblen", dl[si], "tile", si // 16, "in-tile", si % 16, "paired", paired[qi, si], "single", single[qi, si],
              "oracle", orc.score(qc[qo[qi]:qo[qi + 1]], dc[int(do[si]):int(do[si + 1])], b62, 10, 2))
    if len(bad):
        b = bad[:, 1]
        print("   seq idx range", b.min(), b.max(), "dblen range", dl[b].min(), dl[b].max(), "in-tile hist", np.bincount(b % 16, minlength=16).tolist())
