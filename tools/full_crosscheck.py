"""Full-size cross-check on the GPU box: ALL scores of a workload computed three ways -- the sequence-pair kernel
only (query_pairing = 0), the query-pair kernel for every pair (2) and the planner's mix (1) -- must be identical
(two independent implementations of the recurrence, 11.4 M scores on cfg2).  Usage: python tools/full_crosscheck.py [scale] [cfg2|cfg3]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from swimm_b200 import gpu, host, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
workload = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
sys.argv = sys.argv[:1]
import bench                      # the bench's workloads (databases generated straight in preprocessed form)
w = bench.Workload(workload, scale)
if workload == "cfg2" and os.environ.get("SWG_TITIN"):
    # a Swiss-Prot-like tail: a few sequences of up to 35 213 residues (titin) appended to the sorted database
    rng = np.random.default_rng(99)
    extra = np.sort(rng.integers(12000, 35214, 15)).astype(np.uint16)
    extra[-1] = 35213
    w.dl = np.concatenate([w.dl, extra])
    w.dc = np.concatenate([w.dc, synth.encode(synth.random_residues(rng, int(extra.astype(np.int64).sum())))])
q, dl, dc, ql, qc, qo = w.q, w.dl, w.dc, w.ql, w.qc, w.qo
b62 = host.submat("blosum62")
s = gpu.GpuSearch(0)
s.load_db(dl, dc)
res = {}
print("%s: %d sequences, %d residues" % (workload, len(dl), len(dc)), flush=True)
modes = (0, 1) if workload in ("cfg3", "cfg4") else (0, 2, 1)
for mode in modes:
    s.set_option("query_pairing", mode)
    # mode 0 is also the run WITHOUT long-tile split and column chunks: the plainest path is the yardstick
    s.set_option("long_threshold", 65535 if mode == 0 else 0)
    s.set_option("chunk_columns", 1 if mode == 0 else 0)
    s.set_option("verbose", 1 if mode == 1 and os.environ.get("SWG_VERBOSE") else 0)
    t0 = time.time()
    sc, keys = s.search(qc, ql, qo[:-1], b62, 10, 2, 10, want_scores=True)
    st = s.stats()
    res[mode] = (sc, keys)
    print("query_pairing=%d: %d queries on the query-pair kernel, %d launches, search %.3f s = %.0f GCUPS, max score %d, checksum %d"
          % (mode, int(s.query_kernels().sum()), st["launches"], st["search_seconds"], st["cells"] / st["search_seconds"] / 1e9,
             int(sc.max()), int(sc.astype(np.int64).sum())))
ok = True
for mode in modes[1:]:
    same_s = np.array_equal(res[0][0], res[mode][0])
    same_k = np.array_equal(res[0][1], res[mode][1])
    print("query_pairing=%d vs 0: %d x %d scores identical: %s, top-10 keys identical: %s (%d differing scores)"
          % (mode, res[0][0].shape[0], res[0][0].shape[1], same_s, same_k, int((res[0][0] != res[mode][0]).sum())))
    ok = ok and same_s and same_k
sys.exit(0 if ok else 1)
