"""Per-query GCUPS on cfg2 (scale arg) for a few option settings."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swimm_b200 import gpu, host, synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
q = synth.make_queries(np.random.default_rng(7), synth.QUERY_LENGTHS)
db = synth.make_db(1000, int(570_000 * scale) // 16 * 16, mu=5.675, queries=q)
_, dl, dc = synth.length_sorted(db)
_, ql, qc = synth.length_sorted(q)
qo = np.zeros(q.n + 1, np.uint32); np.cumsum(ql.astype(np.uint32), out=qo[1:])
b62 = host.submat("blosum62")
s = gpu.GpuSearch(0)
s.load_db(dl, dc)
print("max len", dl.max(), "n", len(dl))
for name, opts in [("default", {}), ("no_long", {"long_threshold": 65535}), ("fixed3072", {"long_threshold": 3072})] + [(a, json.loads(b)) for a, b in zip(sys.argv[2::2], sys.argv[3::2])]:
    for k, v in opts.items():
        s.set_option(k, v)
    s.set_queries(qc, ql, qo[:-1], b62, 10, 2)
    best = np.full(q.n, 1e9)
    tot = 1e9
    for rep in range(3):
        s.run(10); s.sync()
        best = np.minimum(best, s.query_seconds())
        tot = min(tot, s.stats()["search_seconds"])
    g = ql.astype(np.float64) * len(dc) / best / 1e9
    print(name, "total %.0f GCUPS" % (ql.sum() * len(dc) / tot / 1e9), " ".join("%d:%.0f" % (a, b) for a, b in zip(ql, g)))
    s.set_option("long_threshold", 0)
