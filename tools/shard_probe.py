"""Launch-bound floor: cfg1 split 8 ways, one shard searched on one GPU (what every rank of the strong-scaling leg does)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
sys.argv = sys.argv[:1]
import bench
from swimm_b200 import gpu, host
b62 = host.submat("blosum62")
w = bench.Workload("cfg1")
s = gpu.GpuSearch(0)
for shards in (1, 2, 4, 8):
    worst = 0.0
    for sh in sorted({0, 1, shards - 1}):
        if sh >= shards:
            continue
        s.load_db(w.dl, w.dc, shard=sh, num_shards=shards)
        s.set_queries(w.qc, w.ql, w.qo[:-1], b62, 10, 2)
        if shards == 8 and sh == 1:
            s.set_option("verbose", 1); s.run(10); s.sync(); s.set_option("verbose", 0)
        per, _ = bench.timed_steps(s, lambda: (s.run(10), s.sync()), 20, 5)
        worst = max(worst, float(np.median(per)))
        print("cfg1 shard %d of %d: device ms median %.3f min %.3f (search %.3f)" % (sh, shards, float(np.median(per)) * 1e3, min(per) * 1e3, s.stats()["search_seconds"] * 1e3), flush=True)
    print("  -> %d GPUs: %.0f GCUPS (slowest sampled shard)" % (shards, w.cells / worst / 1e9), flush=True)
