"""A small run of every kernel path checked against the oracle: sequence-pair kernel single- and multi-pass, long-tile
path, 32-bit recomputation, query-pair kernel single launch and lane streams (one CTA and full grid), top-r."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from swimm_b200 import gpu, host, synth
from tests.helpers import load_oracle

rng = np.random.default_rng(5)
q = synth.make_queries(rng, [0, 33, 144, 150, 700, 1100, 1300, 2500])
db = synth.make_seqset(rng, np.concatenate([synth.lognormal_lengths(rng, 330, 4.5, 0.8, 1, 900), [3, 1, 2500, 6000]]))
synth.plant(rng, db, q, fraction=0.1, frag_range=(10, 400), rate=0.1)
w = np.full(3300, ord("W"), np.uint8)
db.residues[db.offsets[333]:db.offsets[333] + 3300] = w            # W x 3300 inside the 6000-residue sequence
q = synth.SeqSet(np.concatenate([q.residues, w[:3200]]), np.concatenate([q.offsets, [q.offsets[-1] + 3200]]), q.titles + [">w"])
_, dl, dc = synth.length_sorted(db)
_, ql, qc = synth.length_sorted(q)
qo = np.zeros(q.n + 1, np.uint32); np.cumsum(ql.astype(np.uint32), out=qo[1:])
do = np.zeros(db.n + 1, np.uint64); np.cumsum(dl.astype(np.uint64), out=do[1:])
b62 = host.submat("blosum62")
want = load_oracle().search(qc, qo, dc, do, b62, 10, 2)
s = gpu.GpuSearch(0)
s.load_db(dl, dc)
for name, opts in [("planner", {}), ("pairs forced", {"query_pairing": 2}), ("pairs forced, 1 CTA", {"query_pairing": 2, "grid_blocks": 1}),
                   ("no pairs, long tiles > 64 columns", {"query_pairing": 0, "long_threshold": 64}),
                   ("no pairs, 1 CTA", {"query_pairing": 0, "grid_blocks": 1})]:
    for k in ("query_pairing", "grid_blocks", "long_threshold"):
        s.set_option(k, 1 if k == "query_pairing" else 0)
    for k, v in opts.items():
        s.set_option(k, v)
    got, keys = s.search(qc, ql, qo[:-1], b62, 10, 2, 7, want_scores=True)
    st = s.stats()
    print("%-36s ok=%s launches=%d pair_launches=%d rescored=%d" % (name, np.array_equal(got, want), st["launches"], st["pair_launches"], st["rescored"]), flush=True)
    assert np.array_equal(got, want)
s.close()
print("done, max score", int(want.max()))
