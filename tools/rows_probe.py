"""The benchmark batch with the planner's rows per thread and with forced ones (option q2_rows) for the query-pair
kernel: does the rate table still pick the fastest K after a kernel change?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
sys.argv = sys.argv[:1]
import bench
from swimm_b200 import gpu, host
w = bench.Workload("cfg2", 1.0)
s = gpu.GpuSearch(0)
s.load_db(w.dl, w.dc)
s.set_queries(w.qc, w.ql, w.qo[:-1], host.submat("blosum62"), 10, 2)
for _ in range(2):
    s.run(10); s.sync()
for rows in (0, 24, 26, 28, 30, 32, 0):
    s.set_option("q2_rows", rows)
    best = 1e9
    for _ in range(3):
        s.run(10); s.sync()
        best = min(best, s.stats()["search_seconds"])
    st = s.stats()
    print("q2_rows %2d: %7.0f GCUPS, %d launches (%d of the query-pair kernel)" % (rows, w.cells / best / 1e9, st["launches"], st["pair_launches"]), flush=True)
