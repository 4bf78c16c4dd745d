"""Shared test plumbing: FASTA reader, golden-case loader, ctypes view of the oracle (the checker)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

from swimm_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["basic", "overflow", "matrices", "edge"]


def read_fasta(path) -> synth.SeqSet:
    titles, seqs, cur = [], [], None
    with open(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\n")
            if line.startswith(b">"):
                titles.append(line.decode())
                cur = []
                seqs.append(cur)
            elif cur is not None:
                cur.append(line)
    arrs = [np.frombuffer(b"".join(s), dtype=np.uint8) for s in seqs]
    off = np.zeros(len(arrs) + 1, dtype=np.int64)
    np.cumsum([len(a) for a in arrs], out=off[1:])
    res = np.concatenate(arrs) if arrs else np.zeros(0, np.uint8)
    return synth.SeqSet(res.copy(), off, titles)


class GoldenCase:
    """Inputs exactly as the reference saw them + the hit lists it printed."""

    def __init__(self, name):
        self.name = name
        self.db = read_fasta(os.path.join(GOLDEN, name + ".db.fasta"))
        self.q = read_fasta(os.path.join(GOLDEN, name + ".q.fasta"))
        self.meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
        self.perm, self.db_len, self.db_codes = synth.length_sorted(self.db)      # sorted pos -> FASTA index
        self.rank_of = np.empty(self.db.n, dtype=np.int64)                          # FASTA index -> sorted pos
        self.rank_of[self.perm] = np.arange(self.db.n)
        self.db_off = np.zeros(self.db.n + 1, dtype=np.uint64)
        np.cumsum(self.db_len.astype(np.uint64), out=self.db_off[1:])
        qperm, self.q_len, self.q_codes = synth.length_sorted(self.q)
        assert (qperm == np.arange(self.q.n)).all(), "golden query files are written in ascending length order"
        self.q_off = np.zeros(self.q.n + 1, dtype=np.uint32)
        np.cumsum(self.q_len.astype(np.uint32), out=self.q_off[1:])

    def runs(self):
        for r in self.meta["runs"]:
            # per query: scores indexed by sorted position, and the printed order as sorted positions
            exp_scores = np.zeros((self.q.n, self.db.n), dtype=np.int32)
            order = np.zeros((self.q.n, self.db.n), dtype=np.int64)
            for qi, hits in enumerate(r["hits"]):
                h = np.array(hits, dtype=np.int64)
                pos = self.rank_of[h[:, 1]]
                exp_scores[qi, pos] = h[:, 0]
                order[qi] = pos
            yield r["matrix"], r["go"], r["ge"], exp_scores, order


def ensure_oracle_built():
    lib = os.path.join(ROOT, "oracle", "liboracle.so")
    src = os.path.join(ROOT, "oracle", "sw_oracle.c")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return lib


class Oracle:
    """ctypes view of oracle/liboracle.so -- the CPU checker (never on the product path)."""

    def __init__(self):
        L = C.CDLL(ensure_oracle_built())
        self.L = L
        L.swo_encode_residue.restype = C.c_int
        L.swo_score.restype = C.c_int32
        L.swo_score.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int]
        L.swo_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                 C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.swo_sort_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.swo_top.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.swo_length_order.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.swo_align_ends.restype = C.c_int32
        L.swo_align_ends.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_void_p]

    def encode_residue(self, c):
        return self.L.swo_encode_residue(int(c))

    def score(self, q, d, submat, go, ge):
        q = np.ascontiguousarray(q, np.int8)
        d = np.ascontiguousarray(d, np.int8)
        submat = np.ascontiguousarray(submat, np.int8)
        return self.L.swo_score(q.ctypes.data, len(q), d.ctypes.data, len(d), submat.ctypes.data, go, ge)

    def align_ends(self, q, d, submat, go, ge):
        """EXTENSION (no reference counterpart): (score, [q_start, q_end, d_start, d_end]) 0-based inclusive."""
        q = np.ascontiguousarray(q, np.int8)
        d = np.ascontiguousarray(d, np.int8)
        submat = np.ascontiguousarray(submat, np.int8)
        c = np.zeros(4, np.int32)
        s = self.L.swo_align_ends(q.ctypes.data, len(q), d.ctypes.data, len(d), submat.ctypes.data, go, ge, c.ctypes.data)
        return s, c

    def search(self, q_codes, q_off, db_codes, db_off, submat, go, ge, threads=0):
        q_codes = np.ascontiguousarray(q_codes, np.int8)
        q_off = np.ascontiguousarray(q_off, np.uint32)
        db_codes = np.ascontiguousarray(db_codes, np.int8)
        db_off = np.ascontiguousarray(db_off, np.uint64)
        submat = np.ascontiguousarray(submat, np.int8)
        nq, n = len(q_off) - 1, len(db_off) - 1
        out = np.zeros((nq, n), dtype=np.int32)
        self.L.swo_search(q_codes.ctypes.data, q_off.ctypes.data, nq, db_codes.ctypes.data, db_off.ctypes.data, n,
                          submat.ctypes.data, go, ge, threads, out.ctypes.data)
        return out

    def sort_scores(self, scores):
        s = np.array(scores, dtype=np.int32)
        idx = np.arange(len(s), dtype=np.uint64)
        self.L.swo_sort_scores(s.ctypes.data, idx.ctypes.data, len(s))
        return s, idx

    def top(self, scores, r):
        s = np.ascontiguousarray(scores, np.int32)
        r = min(r, len(s))
        ts = np.zeros(r, np.int32)
        ti = np.zeros(r, np.uint64)
        self.L.swo_top(s.ctypes.data, len(s), r, ts.ctypes.data, ti.ctypes.data)
        return ts, ti

    def length_order(self, lengths):
        l = np.ascontiguousarray(lengths, np.uint16)
        p = np.zeros(len(l), np.uint64)
        self.L.swo_length_order(l.ctypes.data, len(l), p.ctypes.data)
        return p


def load_oracle():
    return Oracle()
