"""Why is a millisecond-sized search slower inside bench.py's rank 0 than in a fresh process?  cfg1 timed after each
suspect in turn."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
sys.argv = sys.argv[:1]
import bench
from swimm_b200 import gpu, host
import torch
torch.cuda.set_device(0)
b62 = host.submat("blosum62")
w = bench.Workload("cfg1")
s = gpu.GpuSearch(0)
s.load_db(w.dl, w.dc)
s.set_queries(w.qc, w.ql, w.qo[:-1], b62, 10, 2)

def t(tag):
    def one():
        s.run(10); s.sync()
    per, n = bench.timed_steps(s, one, 20, 5)
    w0 = time.time()
    for _ in range(20):
        one()
    wall = (time.time() - w0) / 20
    print("%-40s device ms: min %.3f median %.3f max %.3f; wall per step %.3f ms; SM %s MHz" % (
        tag, min(per) * 1e3, float(np.median(per)) * 1e3, max(per) * 1e3, wall * 1e3, bench.sm_clock_mhz(0)), flush=True)

t("fresh")
chk = gpu.GpuSearch(0); chk.load_db(w.dl[:1600], w.dc[:int(w.dl[:1600].astype(np.int64).sum())]); chk.close()
t("after a second context came and went")
s.pipebench()
t("after pipebench")
smp = bench.ClockSampler(0); smp.start(); time.sleep(1.0); smp.stop_flag.set(); smp.join(timeout=3)
t("after the clock sampler thread")
x = torch.zeros(1 << 20, device="cuda"); torch.cuda.synchronize()
t("after a torch allocation")
pin = torch.zeros(1 << 20).pin_memory()
t("after pinning host memory with torch")
