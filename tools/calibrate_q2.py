"""Measure the query-pair kernel's throughput for every (G, K) shape on the GPU: two queries whose length fits the
shape exactly; single pass for G = 8, 16, 32 and 3 passes for G = 32.  GCUPS count the cells of both queries.
Prints JSON: rates[G][K] = [GCUPS 1 pass, GCUPS 3 passes (G = 32 only)]."""
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
s.set_option("query_pairing", 2)
res = {}
Gs = [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else '8,16,32'.split(','))]
Ks = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else list(range(8, 33, 2))
for G in Gs:
    res[G] = {}
    for K in Ks:
        out = []
        for passes in ([1, 3] if G == 32 else [1]):
            m = G * K * passes
            qc = np.concatenate([synth.encode(synth.random_residues(rng, m)) for _ in range(2)])
            s.set_option("q2_group", G)
            s.set_option("q2_rows", K)
            s.set_queries(qc, np.array([m, m], np.uint16), np.array([0, m], np.uint32), b62, 10, 2)
            best = 0.0
            for rep in range(3):
                s.run(10)
                s.sync()
                t = float(s.query_seconds().sum())
                best = max(best, 2 * m * len(dc) / t / 1e9)
            out.append(round(best, 1))
        res[G][K] = out
        print(G, K, out, file=sys.stderr, flush=True)
print(json.dumps(res))
