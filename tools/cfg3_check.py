"""cfg3 (Environmental-NR-sized synthetic database, ~6 M sequences / ~1.3 G residues, 20 queries, top 10) on the GPU box:
the whole database on one B200, then the same database as 8 shards (tile round-robin, searched one after the other
on the same GPU) with the per-shard hit lists merged -- the merged lists must equal the unsharded ones, and every
shard must hold 1/8 of the residues.  Usage: python tools/cfg3_check.py [scale] [shards]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from swimm_b200 import gpu, host, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
shards = int(sys.argv[2]) if len(sys.argv) > 2 else 8
t0 = time.time()
q = synth.make_queries(np.random.default_rng(7), synth.QUERY_LENGTHS)
db = synth.make_db(3, int(6_000_000 * scale) // 16 * 16, mu=5.2, sigma=0.6, queries=q)
_, dl, dc = synth.length_sorted(db)
_, ql, qc = synth.length_sorted(q)
qo = np.zeros(q.n + 1, np.uint32)
np.cumsum(ql.astype(np.uint32), out=qo[1:])
b62 = host.submat("blosum62")
print("cfg3: %d sequences, %d residues, longest %d (%.0f s to generate)" % (len(dl), len(dc), int(dl.max()), time.time() - t0), flush=True)
s = gpu.GpuSearch(0)
t0 = time.time()
s.load_db(dl, dc)
t_load = time.time() - t0
_, keys = s.search(qc, ql, qo[:-1], b62, 10, 2, 10)
_, keys = s.search(qc, ql, qo[:-1], b62, 10, 2, 10)
st = s.stats()
print("1 GPU, unsharded: load %.2f s, search %.3f s = %.0f GCUPS, %d launches (%d of the query-pair kernel), %d lanes recomputed in 32 bits"
      % (t_load, st["search_seconds"], st["cells"] / st["search_seconds"] / 1e9, st["launches"], st["pair_launches"], st["rescored"]), flush=True)
parts, residues, secs = [], [], []
for sh in range(shards):
    s.load_db(dl, dc, shard=sh, num_shards=shards)
    residues.append(s.local_residues)
    _, k = s.search(qc, ql, qo[:-1], b62, 10, 2, 10)
    secs.append(s.stats()["search_seconds"])
    parts.append(k)
merged = gpu.merge_top_keys(parts, 10)
ok = np.array_equal(merged, keys)
print("%d shards: residues per shard %d..%d (max/mean %.4f), search seconds per shard %.3f..%.3f -> %.0f GCUPS if run on %d GPUs side by side"
      % (shards, min(residues), max(residues), max(residues) / (sum(residues) / shards), min(secs), max(secs),
         float(ql.astype(np.int64).sum()) * len(dc) / max(secs) / 1e9, shards))
print("merged top-10 of the %d shards == unsharded top-10 for all %d queries: %s; best scores %s" % (shards, q.n, ok, [int(k >> np.uint64(32)) for k in keys[:, 0]][:6]))
sys.exit(0 if ok else 1)
