import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
sys.argv = sys.argv[:1]
import bench
from swimm_b200 import gpu, host
b62 = host.submat("blosum62")
w = bench.Workload("cfg1")
s = gpu.GpuSearch(0)
s.load_db(w.dl, w.dc)
s.set_queries(w.qc, w.ql, w.qo[:-1], b62, 10, 2)
for _ in range(300):
    s.run(10); s.sync()
s.set_option("trace", 1)
for i in range(8):
    s.run(10); s.sync()
    st = s.stats()
    print("step %d: device %.3f ms search %.3f topr %.3f" % (i, st["device_seconds"] * 1e3, st["search_seconds"] * 1e3, st["topr_seconds"] * 1e3), file=sys.stderr, flush=True)
