#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the UNMODIFIED reference binary.

Dev-container only (needs oracle/_ref/swimm, built by `make -C oracle ref` from /root/reference).
For every case it writes the FASTA inputs it fed to the reference (<case>.db.fasta, <case>.q.fasta)
and <case>.json with, per run (matrix, gap penalties) and per query, the complete printed hit list
`[(score, original FASTA index of the database sequence), ...]` in the order the reference printed it
(`-r <n>` makes it print all n scores, which pins both the scores and the tie order).

    python tests/golden/make_golden.py
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from swimm_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "swimm")
MATS = ["blosum45", "blosum50", "blosum62", "blosum80", "blosum90", "pam30", "pam70", "pam250"]


def run_reference(db_fasta, q_fasta, n, matrix, go, ge, vec=32, threads=4):
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, "db")
        subprocess.run([REF, "-S", "preprocess", "-i", db_fasta, "-o", prefix, "-c", str(threads)],
                       check=True, stdout=subprocess.DEVNULL)
        out = subprocess.run([REF, "-S", "search", "-q", q_fasta, "-d", prefix, "-m", "0", "-v", str(vec),
                              "-c", str(threads), "-r", str(n), "-s", matrix, "-g", str(go), "-e", str(ge)],
                             check=True, stdout=subprocess.PIPE).stdout.decode("latin-1")
    queries, cur = [], None
    for line in out.split("\n"):
        if line.startswith("Query no."):
            cur = []
            queries.append(cur)
        m = re.match(r"^(-?\d+)\t.*syn\|(\d+)\|", line)
        if m and cur is not None:
            cur.append([int(m.group(1)), int(m.group(2))])
    return queries


def w_run(k):
    return np.full(k, ord("W"), dtype=np.uint8)


def build_case(name):
    rng = np.random.default_rng({"basic": 11, "overflow": 12, "matrices": 13, "edge": 14}[name])
    runs = [("blosum62", 10, 2)]
    if name == "basic":
        q = synth.make_queries(rng, [37, 144, 200])
        db = synth.make_seqset(rng, synth.lognormal_lengths(rng, 300, 4.6, 0.6, 5, 900))
        synth.plant(rng, db, q, fraction=0.08, frag_range=(10, 120), rate=0.2)
        # exact duplicates -> guaranteed score ties between different database indices
        for a, b in [(3, 250), (17, 18), (100, 299)]:
            la = db.offsets[a + 1] - db.offsets[a]
            lb = db.offsets[b + 1] - db.offsets[b]
            k = min(la, lb)
            db.residues[db.offsets[b]:db.offsets[b] + k] = db.residues[db.offsets[a]:db.offsets[a] + k]
    elif name == "overflow":
        qa = synth.random_residues(rng, 300)
        qb = synth.random_residues(rng, 1201)
        qs = [qa, qb, w_run(3100)]
        off = np.cumsum([0] + [len(x) for x in qs])
        q = synth.SeqSet(np.concatenate(qs), off,
                         [">query|%04d| overflow query %d" % (i, len(x)) for i, x in enumerate(qs)])
        lens = np.concatenate([synth.lognormal_lengths(rng, 40, 5.5, 0.7, 30, 2500),
                               rng.integers(3001, 3400, 8), [3200, 1400, 700, 400, 64]])
        db = synth.make_seqset(rng, lens)
        n0 = 48
        db.residues[db.offsets[n0]:db.offsets[n0 + 1]] = ord("W")                 # 3200 W  -> 34100 vs W*3100
        s = db.offsets[n0 + 1]
        db.residues[s + 100:s + 100 + 1201] = synth.mutate(rng, qb, 0.03)          # near copy -> thousands
        s = db.offsets[n0 + 2]
        db.residues[s + 200:s + 500] = qa                                          # exact copy -> > 127
        s = db.offsets[n0 + 3]
        db.residues[s + 10:s + 30] = qa[50:70]                                     # 20-residue overlap
        s = db.offsets[n0 + 4]
        db.residues[s:s + 64] = ord("W")
    elif name == "matrices":
        q = synth.make_queries(rng, [25, 60, 101])
        db = synth.make_seqset(rng, synth.lognormal_lengths(rng, 120, 4.4, 0.5, 8, 400))
        synth.plant(rng, db, q, fraction=0.15, frag_range=(8, 90), rate=0.25)
        runs = [(m, go, ge) for m in MATS for go, ge in [(10, 2), (5, 1), (12, 3), (8, 2)]]
    elif name == "edge":
        q = synth.make_queries(rng, [1, 2, 5, 33])
        db = synth.make_seqset(rng, np.arange(1, 34))       # 33 sequences: one full 32-lane group + 1
        synth.plant(rng, db, q, fraction=0.3, frag_range=(2, 30), rate=0.0)
    else:
        raise KeyError(name)
    return db, q, runs


def main():
    if not os.path.exists(REF):
        sys.exit("build the reference first: make -C oracle ref")
    for name in ["basic", "overflow", "matrices", "edge"]:
        db, q, runs = build_case(name)
        dbf = os.path.join(HERE, name + ".db.fasta")
        qf = os.path.join(HERE, name + ".q.fasta")
        synth.write_fasta(dbf, db)
        synth.write_fasta(qf, q)
        out = {"case": name, "n": db.n, "q_lengths": [int(x) for x in q.lengths], "runs": []}
        for (mat, go, ge) in runs:
            hits = run_reference(dbf, qf, db.n, mat, go, ge)
            assert len(hits) == q.n and all(len(h) == db.n for h in hits), (name, mat, [len(h) for h in hits])
            if name in ("basic", "edge"):
                sse = run_reference(dbf, qf, db.n, mat, go, ge, vec=16)
                assert sse == hits, "reference SSE and AVX2 paths disagree"
            out["runs"].append({"matrix": mat, "go": go, "ge": ge, "hits": hits})
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f, separators=(",", ":"))
        print(name, "n=%d" % db.n, "runs=%d" % len(runs), "max score", max(h[0][0] for r in out["runs"] for h in r["hits"]))


if __name__ == "__main__":
    main()
