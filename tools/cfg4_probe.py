"""cfg4 (long-sequence stress) timing probe on the GPU box: per-query GCUPS with the long-tile path automatic, off and forced."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from swimm_b200 import gpu, host, synth

rng = np.random.default_rng(44)
q = synth.make_queries(rng, [144, 1000, 3100, 5478])
n = 8000
lens = np.concatenate([rng.integers(3001, 20000, n - 8), [30000, 40000, 50000, 60000, 65535, 65535, 3001, 3002]])
db = synth.make_seqset(rng, lens)
synth.plant(rng, db, q, fraction=0.02, frag_range=(50, 3000), rate=0.1)
_, dl, dc = synth.length_sorted(db)
_, ql, qc = synth.length_sorted(q)
qo = np.zeros(q.n + 1, np.uint32)
np.cumsum(ql.astype(np.uint32), out=qo[1:])
b62 = host.submat("blosum62")
s = gpu.GpuSearch(0)
s.load_db(dl, dc)
ref = None
for name, opts in [("auto", {}), ("auto verbose", {"verbose": 1}), ("old long kernel", {"long_kernel": 0}),
                   ("no long path", {"long_threshold": 65535}),
                   ("all long (xw)", {"long_threshold": 8}), ("all long W16", {"long_threshold": 8, "xw_warps": 16}),
                   ("all long W8", {"long_threshold": 8, "xw_warps": 8}), ("all long W4", {"long_threshold": 8, "xw_warps": 4}),
                   ("long > 20000", {"long_threshold": 20000}), ("long > 30000", {"long_threshold": 30000}),
                   ("no pairs", {"query_pairing": 0}), ("no pairs, old long", {"query_pairing": 0, "long_kernel": 0})]:
    for k, v in {"long_threshold": 0, "query_pairing": 1, "long_kernel": 1, "xw_warps": 0, "verbose": 0}.items():
        s.set_option(k, v)
    for k, v in opts.items():
        s.set_option(k, v)
    s.search(qc, ql, qo[:-1], b62, 10, 2, 10)
    sc, _ = s.search(qc, ql, qo[:-1], b62, 10, 2, 10, want_scores=True)
    st = s.stats()
    if ref is None:
        ref = sc
    qs = s.query_seconds()
    print("%-18s search %.3f s = %6.0f GCUPS, launches %d, per query %s, same scores: %s" % (
        name, st["search_seconds"], st["cells"] / st["search_seconds"] / 1e9, st["launches"],
        [int(float(ql[i]) * len(dc) / qs[i] / 1e9) for i in range(q.n)], np.array_equal(sc, ref)), flush=True)
