"""Single queries and a batch with default / BLAST-style / other gap penalties on the cfg2-sized database: GCUPS per
instantiation family (both penalties immediates, gap-extend immediate, generic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
sys.argv = sys.argv[:1]
import bench
from swimm_b200 import gpu, host
w = bench.Workload("cfg2", 0.5)
s = gpu.GpuSearch(0)
s.load_db(w.dl, w.dc)
for label, sel in (("single query 375", [3]), ("single query 2005", [11]), ("batch of 20", list(range(20)))):
    ql = w.ql[sel]
    qc = np.concatenate([w.qc[w.qo[i]:w.qo[i + 1]] for i in sel])
    qo = np.concatenate([[0], np.cumsum(ql.astype(np.uint32))]).astype(np.uint32)
    for matrix, go, ge in (("blosum62", 10, 2), ("blosum62", 11, 1), ("blosum62", 9, 2), ("blosum62", 10, 3)):
        s.set_queries(qc, ql, qo[:-1], host.submat(matrix), go, ge)
        best = 1e9
        for _ in range(4):
            s.run(10); s.sync()
            best = min(best, s.stats()["search_seconds"])
        print("%-18s %s %2d/%d: %6.0f GCUPS" % (label, matrix, go, ge, float(ql.astype(np.int64).sum()) * len(w.dc) / best / 1e9), flush=True)
