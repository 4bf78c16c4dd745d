"""cfg1 (one 144-residue query x 100 k sequences) timing probe on the GPU box: shapes and long-tile settings."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from swimm_b200 import gpu, host, synth

q = synth.make_queries(np.random.default_rng(42), [144])
db = synth.make_db(42, 100_000, queries=q)
_, dl, dc = synth.length_sorted(db)
_, ql, qc = synth.length_sorted(q)
b62 = host.submat("blosum62")
s = gpu.GpuSearch(0)
s.load_db(dl, dc)
print("longest sequences:", dl[-5:], "residues", len(dc))
s.set_option("verbose", 1)
s.search(qc, ql, np.zeros(1, np.uint32), b62, 10, 2, 10)
s.set_option("verbose", 0)
for name, opts in [("auto", {}), ("no long", {"long_threshold": 65535}), ("G8 K18 no long", {"force_group": 8, "force_rows": 18, "long_threshold": 65535}),
                   ("G16 K9 no long", {"force_group": 16, "force_rows": 9, "long_threshold": 65535}),
                   ("G32 K5 no long", {"force_group": 32, "force_rows": 5, "long_threshold": 65535}),
                   ("G16 K9 auto", {"force_group": 16, "force_rows": 9}), ("G8 K18 auto", {"force_group": 8, "force_rows": 18}),
                   ("G8 K18 long>1024", {"force_group": 8, "force_rows": 18, "long_threshold": 1024}),
                   ("G8 K18 long>2048", {"force_group": 8, "force_rows": 18, "long_threshold": 2048})]:
    for k in ("force_group", "force_rows", "long_threshold"):
        s.set_option(k, 0)
    for k, v in opts.items():
        s.set_option(k, v)
    best = 1e9
    for rep in range(5):
        s.run(10)
        s.sync()
        st = s.stats()
        best = min(best, st["search_seconds"])
    print("%-20s search %.3f ms = %5.0f GCUPS, launches %d, top-r %.3f ms" % (name, best * 1e3, st["cells"] / best / 1e9, st["launches"], st["topr_seconds"] * 1e3))
