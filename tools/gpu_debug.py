"""GPU-side diagnostics: tiled layout and profile against numpy, then tiny searches against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swimm_b200 import gpu, host, synth
from tests.helpers import load_oracle

oracle = load_oracle()
rng = np.random.default_rng(3)
b62 = host.submat("blosum62")
s = gpu.GpuSearch(0)

def case(n, qlens, hi=80):
    q = synth.make_queries(rng, qlens)
    db = synth.make_seqset(rng, rng.integers(1, hi, n))
    _, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    do = np.zeros(db.n + 1, np.uint64); np.cumsum(dl.astype(np.uint64), out=do[1:])
    qo = np.zeros(q.n + 1, np.uint32); np.cumsum(ql.astype(np.uint32), out=qo[1:])
    return qc, ql, qo, dc, dl, do

qc, ql, qo, dc, dl, do = case(40, [20])
s.load_db(dl, dc)
tile_off = s.debug_read("tile_off", np.uint64)
tile_cols = s.debug_read("tile_cols", np.uint32)
db = s.debug_read("db", np.uint8)
print("ntiles", len(tile_cols), "tile_cols", tile_cols, "tile_off", tile_off)
# expected layout
bad = 0
for t in range(len(tile_cols)):
    for chunk in range(tile_cols[t] // 8):
        for pair in range(8):
            unit = int(tile_off[t]) + chunk * 8 + pair
            for c in range(8):
                for h in range(2):
                    sidx = t * 16 + 2 * pair + h
                    col = chunk * 8 + c
                    exp = 24
                    if sidx < len(dl) and col < dl[sidx]:
                        exp = dc[int(do[sidx]) + col]
                    got = db[unit * 16 + 2 * c + h]
                    if got != exp * 4:
                        bad += 1
                        if bad < 5: print("layout mismatch t", t, "chunk", chunk, "pair", pair, "c", c, "h", h, got, exp * 4)
print("layout mismatches:", bad)

want = oracle.search(qc, qo, dc, do, b62, 10, 2)
for G in [4, 8, 16, 32]:
    for K in [1, 2, 5, 8]:
        if G * K < 20: continue
        s.set_option("force_group", G); s.set_option("force_rows", K)
        got, _ = s.search(qc, ql, qo[:-1], b62, 10, 2, 0, want_scores=True)
        ok = np.array_equal(got, want)
        print("G", G, "K", K, "ok", ok)
        if not ok:
            print(" got ", got[0][:24]); print(" want", want[0][:24])
            prof = s.debug_read("profile", np.int8)
            prof32 = s.debug_read("profile32", np.int8)
            # expected profile for (G,K)
            exp = np.zeros(25 * 1024, np.int8)
            for row in range(20):
                t, x = row // K, row % K
                for letter in range(25):
                    exp[letter * 1024 + (x >> 4) * (G * 16) + t * 16 + (x & 15)] = b62[qc[row], letter]
            print(" profile ok:", np.array_equal(prof[:25 * 1024], exp), "profile32 first bytes", prof32[:8])
            print(" counters", s.debug_read("counters", np.uint32))
            break
s.set_option("force_group", 0); s.set_option("force_rows", 0)
