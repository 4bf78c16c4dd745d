"""Long-tile decisions and timings on the GPU box: cfg1 (two generators), cfg4, cfg5 with the planner's choice, the split
forced off, and the previous long-tile kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from swimm_b200 import gpu, host, synth
sys.argv = sys.argv[:1]
import bench

b62 = host.submat("blosum62")
s = gpu.GpuSearch(0)


def run(tag, w, opts, reps=5, verbose=False):
    for k, v in {"long_threshold": 0, "query_pairing": 1, "long_kernel": 1, "xw_warps": 0, "xw_rows": 0, "verbose": 0, "chunk_columns": 0}.items():
        s.set_option(k, v)
    for k, v in opts.items():
        s.set_option(k, v)
    s.set_queries(w.qc, w.ql, w.qo[:-1], b62, 10, 2)
    if verbose:
        s.set_option("verbose", 1)
        s.run(10); s.sync()
        s.set_option("verbose", 0)
    best = 1e9
    for _ in range(reps):
        s.run(10); s.sync()
        best = min(best, s.stats()["search_seconds"])
    st = s.stats()
    print("%-28s search %8.3f ms = %6.0f GCUPS, launches %d" % (tag, best * 1e3, st["cells"] / best / 1e9, st["launches"]), flush=True)


class W1:      # cfg1 with the FASTA-order generator of round 1 (longest sequence 7367)
    def __init__(self):
        q = synth.make_queries(np.random.default_rng(42), [144])
        db = synth.make_db(42, 100_000, queries=q)
        _, self.dl, self.dc = synth.length_sorted(db)
        _, self.ql, self.qc = synth.length_sorted(q)
        self.qo = np.zeros(2, np.uint32); self.qo[1] = 144


for name, w in [("cfg1", bench.Workload("cfg1")), ("cfg5", bench.Workload("cfg5")), ("cfg4", bench.Workload("cfg4"))]:
    s.load_db(w.dl, w.dc)
    print(name, "longest", w.dl[-3:], "residues", len(w.dc), flush=True)
    run(name + " auto", w, {}, verbose=True)
    run(name + " no split", w, {"long_threshold": 65535})
    run(name + " old long kernel", w, {"long_kernel": 0})
    run(name + " no column chunks", w, {"chunk_columns": 1})
    for c in (1024, 1536, 3072, 4096):
        if name.startswith("cfg1"):
            run(name + " chunks of %d" % c, w, {"chunk_columns": c})
    for thr in (1024, 2048):
        if name.startswith("cfg1"):
            run(name + " long > %d" % thr, w, {"long_threshold": thr})
