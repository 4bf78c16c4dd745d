"""ncu target: one cfg4 search (long-sequence stress), so that the long-sequence kernel's launches can be captured."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = sys.argv[:1]
import bench
from swimm_b200 import gpu, host
w = bench.Workload("cfg4")
s = gpu.GpuSearch(0)
s.load_db(w.dl, w.dc)
s.set_option("verbose", 1)
for _ in range(2):
    s.search(w.qc, w.ql, w.qo[:-1], host.submat("blosum62"), 10, 2, 10)
    s.set_option("verbose", 0)
st = s.stats()
print("cfg4 search %.3f ms = %.0f GCUPS, %d launches" % (st["search_seconds"] * 1e3, st["cells"] / st["search_seconds"] / 1e9, st["launches"]))
