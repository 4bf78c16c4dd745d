"""Measure the search kernel's throughput for every (G, K) shape on the GPU (one synthetic database, one query
whose length fits the shape exactly, single pass and 3 passes).  Prints JSON: rates[G][K] = [GCUPS 1 pass, GCUPS 3 passes]."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swimm_b200 import gpu, host, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.35
rng = np.random.default_rng(5)
db = synth.make_db(77, int(570_000 * scale) // 16 * 16, mu=5.675)
_, dl, dc = synth.length_sorted(db)
b62 = host.submat("blosum62")
s = gpu.GpuSearch(0)
s.load_db(dl, dc)
res = {}
Gs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "4,8,16,32".split(","))]
for G in Gs:
    res[G] = {}
    for K in range(4, 33):
        out = []
        for passes in ([1, 3] if G == 32 else [1]):
            m = G * K * passes
            q = synth.random_residues(rng, m)
            qc = synth.encode(q)
            s.set_option("force_group", G)
            s.set_option("force_rows", K)
            s.set_queries(qc, np.array([m], np.uint16), np.array([0], np.uint32), b62, 10, 2)
            best = 0.0
            for rep in range(3):
                s.run(10)
                s.sync()
                t = s.query_seconds()[0]
                best = max(best, m * len(dc) / t / 1e9)
            out.append(round(best, 1))
        res[G][K] = out
        print(G, K, out, file=sys.stderr, flush=True)
print(json.dumps(res))
