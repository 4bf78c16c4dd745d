"""Measure the long-sequence kernel's (wavefront_xw.cuh) throughput for (warps W, rows K) shapes on the GPU: a database of
long sequences, every tile forced onto the kernel, one query of exactly W*32*K rows.  Prints JSON rates[W][K] = GCUPS."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swimm_b200 import gpu, host, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
Ws = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1,2,4,8,16".split(","))]
Ks = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "1,2,3,4,6,8,10,12,14,16,18,20,22,24,28,32".split(","))]
rng = np.random.default_rng(5)
db = synth.make_seqset(rng, rng.integers(3001, 20000, n))
_, dl, dc = synth.length_sorted(db)
b62 = host.submat("blosum62")
s = gpu.GpuSearch(0)
s.load_db(dl, dc)
s.set_option("long_threshold", 8)
s.set_option("query_pairing", 0)
res = {}
for W in Ws:
    res[W] = {}
    for K in Ks:
        if W * ((K + 15) // 16) > 16:
            continue
        m = W * 32 * K
        qc = synth.encode(synth.random_residues(rng, m))
        s.set_option("xw_warps", W)
        s.set_option("xw_rows", K)
        s.set_queries(qc, np.array([m], np.uint16), np.array([0], np.uint32), b62, 10, 2)
        best = 0.0
        for rep in range(2):
            s.run(10)
            s.sync()
            t = s.query_seconds()[0]
            best = max(best, m * len(dc) / t / 1e9)
        res[W][K] = round(best, 1)
        print(W, K, res[W][K], file=sys.stderr, flush=True)
print(json.dumps(res))
