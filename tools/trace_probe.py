"""Timeline (option trace) of one search on the GPU box for small workloads: where a 1 ms search spends its time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from swimm_b200 import gpu, host
sys.argv = sys.argv[:1]
import bench

b62 = host.submat("blosum62")
s = gpu.GpuSearch(0)
for name in ("cfg1", "cfg4"):
    w = bench.Workload(name)
    s.load_db(w.dl, w.dc)
    variants = [("auto", {})]
    if name == "cfg1":
        variants += [("G=16 K=9", {"force_group": 16, "force_rows": 9}), ("G=32 K=5, no split", {"force_group": 32, "force_rows": 5, "long_threshold": 65535}),
                     ("G=8 K=18, long > 3000", {"long_threshold": 3000}), ("grid 136", {"grid_blocks": 136})]
    for tag, opts in variants:
        for k, v in {"force_group": 0, "force_rows": 0, "long_threshold": 0, "grid_blocks": 0}.items():
            s.set_option(k, v)
        for k, v in opts.items():
            s.set_option(k, v)
        s.set_queries(w.qc, w.ql, w.qo[:-1], b62, 10, 2)
        for _ in range(3):
            s.run(10); s.sync()
        print("==== %s %s" % (name, tag), flush=True)
        s.set_option("trace", 1)
        s.run(10); s.sync()
        s.set_option("trace", 0)
        best = 1e9
        for _ in range(5):
            s.run(10); s.sync()
            best = min(best, s.stats()["search_seconds"])
        print("search %.3f ms = %.0f GCUPS" % (best * 1e3, s.stats()["cells"] / best / 1e9), flush=True)
# fixed cost of a search: a database of 64 short sequences
rng = np.random.default_rng(0)
from swimm_b200 import synth
db = synth.make_seqset(rng, rng.integers(20, 60, 64))
_, dl, dc = synth.length_sorted(db)
s.load_db(dl, dc)
w = bench.Workload("cfg1")
s.set_queries(w.qc, w.ql, w.qo[:-1], b62, 10, 2)
for _ in range(3):
    s.run(10); s.sync()
s.set_option("trace", 1)
s.run(10); s.sync()
print("==== 64 short sequences: search %.3f ms (fixed cost)" % (s.stats()["search_seconds"] * 1e3))
