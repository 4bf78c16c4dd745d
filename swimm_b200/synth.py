"""Seeded synthetic protein data for parity tests and benchmarks.

Everything here is host-side input plumbing: residues are drawn i.i.d. from Swiss-Prot-like
background frequencies, lengths from a clipped log-normal, and a small fraction of database
sequences receive mutated copies of query fragments so the hit lists are not pure noise.
The writers respect every input constraint of the reference parser (uppercase A-Z only, short
lines, '\\n' endings, length <= 65535, multi-query files in ascending length order;
reference sequences.c:28-50, :276/:344).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

# alphabet order of the re-encoded residues (reference sequences.c:165-175)
ALPHABET = "ABCDEFGHIKLMNPQRSTVWXYZ"
_STD = "ARNDCQEGHILKMFPSTWYV"
_FREQ = np.array([8.25, 5.53, 4.06, 5.45, 1.37, 3.93, 6.75, 7.07, 2.27, 5.96, 9.66, 5.84, 2.42, 3.86,
                  4.70, 6.56, 5.34, 1.08, 2.92, 6.87])
_RARE = "BZX"      # ~0.1 % together
_DUMMY = "JOU"     # a few, to exercise code 23

# classic Swiss-Prot query lengths used by the SWIPE / CUDASW++ / SWIMM papers
QUERY_LENGTHS = [144, 189, 222, 375, 464, 567, 657, 729, 850, 1000, 1500, 2005, 2504, 3005, 3564, 4061,
                 4548, 4743, 5147, 5478]

_ENC = np.full(256, 23, dtype=np.int8)
for _c in range(ord("A"), ord("Z") + 1):
    _cc = ord("Z") + 1 if chr(_c) in "JOU" else _c
    _ENC[_c] = _cc - (ord("A") + (_cc > ord("J")) + (_cc > ord("O")) + (_cc > ord("U")))


def encode(ascii_bytes: np.ndarray) -> np.ndarray:
    """ASCII 'A'..'Z' (uint8) -> residue codes 0..23 (int8); J/O/U -> 23."""
    return _ENC[np.asarray(ascii_bytes, dtype=np.uint8)]


def _letter_table() -> tuple[np.ndarray, np.ndarray]:
    letters = np.frombuffer((_STD + _RARE + _DUMMY).encode(), dtype=np.uint8)
    p = np.concatenate([_FREQ / _FREQ.sum() * 0.9988, np.full(3, 0.001 / 3), np.full(3, 0.0002 / 3)])
    return letters, p / p.sum()


def random_residues(rng: np.random.Generator, n: int) -> np.ndarray:
    """n ASCII residues (uint8) with Swiss-Prot-like background frequencies."""
    letters, p = _letter_table()
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    return letters[np.searchsorted(cdf, rng.random(n), side="right")]


def lognormal_lengths(rng: np.random.Generator, n: int, mu: float, sigma: float, lo: int, hi: int) -> np.ndarray:
    return np.clip(np.rint(rng.lognormal(mu, sigma, n)), lo, hi).astype(np.int64)


@dataclass
class SeqSet:
    """A set of sequences in FASTA (input) order."""
    residues: np.ndarray          # uint8 ASCII, concatenated
    offsets: np.ndarray           # int64[n+1]
    titles: list[str] | None = None

    @property
    def n(self) -> int:
        return len(self.offsets) - 1

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets)

    def seq(self, i: int) -> np.ndarray:
        return self.residues[self.offsets[i]:self.offsets[i + 1]]

    def title(self, i: int) -> str:
        return self.titles[i] if self.titles is not None else ">syn|%09d| synthetic protein %d" % (i, i)


def make_seqset(rng: np.random.Generator, lengths: np.ndarray) -> SeqSet:
    lengths = np.asarray(lengths, dtype=np.int64)
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=off[1:])
    return SeqSet(random_residues(rng, int(off[-1])), off)


def mutate(rng: np.random.Generator, frag: np.ndarray, rate: float) -> np.ndarray:
    out = frag.copy()
    k = rng.random(len(out)) < rate
    out[k] = random_residues(rng, int(k.sum()))
    return out


def plant(rng: np.random.Generator, db: SeqSet, queries: SeqSet, fraction: float = 0.001,
          frag_range: tuple[int, int] = (20, 300), rate: float = 0.15) -> None:
    """Overwrite windows of ~fraction of the database sequences with mutated query fragments."""
    n_plant = max(1, int(db.n * fraction))
    targets = rng.choice(db.n, size=min(n_plant, db.n), replace=False)
    for t in targets:
        q = int(rng.integers(queries.n))
        qs = queries.seq(q)
        tl = int(db.offsets[t + 1] - db.offsets[t])
        fl = int(min(rng.integers(frag_range[0], frag_range[1] + 1), len(qs), tl))
        if fl <= 0:
            continue
        qa = int(rng.integers(0, len(qs) - fl + 1))
        ta = int(rng.integers(0, tl - fl + 1))
        db.residues[db.offsets[t] + ta: db.offsets[t] + ta + fl] = mutate(rng, qs[qa:qa + fl], rate)


def make_queries(rng: np.random.Generator, lengths) -> SeqSet:
    lengths = sorted(int(x) for x in lengths)     # ascending: reference sequences.c:276 vs :344
    qs = make_seqset(rng, np.array(lengths))
    qs.titles = [">query|%04d| synthetic query of length %d" % (i, l) for i, l in enumerate(lengths)]
    return qs


def make_db(seed: int, n: int, mu: float = 5.65, sigma: float = 0.65, lo: int = 20, hi: int = 35000,
            queries: SeqSet | None = None, plant_fraction: float = 0.001) -> SeqSet:
    rng = np.random.default_rng(seed)
    db = make_seqset(rng, lognormal_lengths(rng, n, mu, sigma, lo, hi))
    if queries is not None and plant_fraction > 0:
        plant(rng, db, queries, plant_fraction)
    return db


def write_fasta(path: str, s: SeqSet, width: int = 60) -> None:
    """FASTA the reference parser accepts: '\\n' endings, lines < 1000 chars, every line terminated."""
    with open(path, "wb") as f:
        for i in range(s.n):
            f.write(s.title(i).encode() + b"\n")
            seq = s.seq(i).tobytes()
            for a in range(0, len(seq), width):
                f.write(seq[a:a + width] + b"\n")


def length_sorted(s: SeqSet) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(perm, sorted_lengths u16, sorted residue CODES int8) in the reference's preprocessed order:
    stable ascending by length (reference sequences.c:125, :770-865)."""
    lens = s.lengths
    if lens.max(initial=0) > 65535:
        raise ValueError("sequence longer than 65535 residues (reference stores lengths as unsigned short)")
    perm = np.argsort(lens, kind="stable")
    sl = lens[perm]
    off = np.zeros(s.n + 1, dtype=np.int64)
    np.cumsum(sl, out=off[1:])
    # gather: index vector built from per-sequence starts
    starts = s.offsets[:-1][perm]
    idx = np.repeat(starts - off[:-1], sl) + np.arange(off[-1], dtype=np.int64)
    codes = encode(s.residues[idx])
    return perm, sl.astype(np.uint16), codes


def write_preprocessed(prefix: str, s: SeqSet) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Write <prefix>.info/.seq/.desc byte-compatible with `swimm -S preprocess`
    (reference sequences.c:128-205).  Returns length_sorted(s)."""
    perm, sl, codes = length_sorted(s)
    titles = [s.title(int(p)) for p in perm]
    max_title = max((len(t) + 2 for t in titles), default=0)   # strlen(line incl. '\n') + 1 (sequences.c:41)
    with open(prefix + ".info", "w") as f:
        f.write("%d %d %d" % (s.n, int(sl.astype(np.int64).sum()), max_title))
    with open(prefix + ".seq", "wb") as f:
        f.write(sl.astype("<u2").tobytes())
        f.write(codes.tobytes())
    with open(prefix + ".desc", "w") as f:
        for t in titles:
            f.write(t + "\n")
    return perm, sl, codes


def read_preprocessed(prefix: str) -> tuple[np.ndarray, np.ndarray, int]:
    """(lengths u16, residue codes int8, max_title_len) from <prefix>.info/.seq."""
    n, d, mt = (int(x) for x in open(prefix + ".info").read().split())
    raw = np.fromfile(prefix + ".seq", dtype=np.uint8)
    lengths = raw[:2 * n].view("<u2").astype(np.uint16)
    codes = raw[2 * n:2 * n + d].view(np.int8)
    if len(codes) != d:
        raise ValueError("%s.seq is truncated" % prefix)
    return lengths, codes, mt


def interleave_reference(lengths: np.ndarray, codes: np.ndarray, vector_length: int = 32, block: int = 60):
    """The lane-interleaved arrays the reference driver hands to its kernels (what
    assemble_single_chunk_db builds, reference sequences.c:618-734): groups of `vector_length`
    consecutive sorted sequences, group length = longest member rounded up to a multiple of 5,
    residue j of lane k at disp[g] + j*vector_length + k, padding code 24.
    Returns (vect_db int8, vect_lengths u16, vect_blocks u16, vect_disp u64[groups+1])."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    groups = (n + vector_length - 1) // vector_length
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lengths, out=off[1:])
    vlen = np.zeros(groups, dtype=np.int64)
    for g in range(groups):
        last = min(n, (g + 1) * vector_length) - 1
        vlen[g] = (lengths[last] + 4) // 5 * 5
    disp = np.zeros(groups + 1, dtype=np.uint64)
    np.cumsum((vlen * vector_length).astype(np.uint64), out=disp[1:])
    vdb = np.full(int(disp[-1]), 24, dtype=np.int8)
    for g in range(groups):
        view = vdb[int(disp[g]):int(disp[g + 1])].reshape(int(vlen[g]), vector_length)
        for k in range(vector_length):
            s = g * vector_length + k
            if s < n:
                view[:lengths[s], k] = codes[off[s]:off[s + 1]]
    blocks = ((vlen + block - 1) // block).astype(np.uint16)
    return vdb, vlen.astype(np.uint16), blocks, disp


# ---- large benchmark databases, generated directly in the preprocessed (length-sorted, encoded) form -----------------
def sorted_db(seed: int, n: int, mu: float = 5.65, sigma: float = 0.65, lo: int = 20, hi: int = 35000,
              queries: SeqSet | None = None, plant_fraction: float = 0.001) -> tuple[np.ndarray, np.ndarray]:
    """(lengths u16 ascending, residue codes int8 concatenated in that order) of a synthetic database of n sequences:
    what `swimm -S preprocess` would write to <db>.seq.  Same distributions as make_db, but generated straight in
    sorted order with a 16-bit lookup table (about 1 s per 100 M residues), so that every rank of a multi-GPU job can
    build the SAME billion-residue database and keep only its shard."""
    rng = np.random.default_rng(seed)
    lengths = np.sort(lognormal_lengths(rng, n, mu, sigma, lo, hi)).astype(np.int64)
    letters, p = _letter_table()
    codes_of = encode(letters)
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    lut = codes_of[np.searchsorted(cdf, (np.arange(65536) + 0.5) / 65536.0, side="right")].astype(np.int8)
    total = int(lengths.sum())
    codes = np.empty(total, dtype=np.int8)
    step = 1 << 26
    for a in range(0, total, step):
        b = min(total, a + step)
        codes[a:b] = lut[rng.integers(0, 65536, b - a, dtype=np.uint16)]
    if queries is not None and plant_fraction > 0:
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lengths, out=off[1:])
        n_plant = max(1, int(n * plant_fraction))
        for t in rng.choice(n, size=min(n_plant, n), replace=False):
            qs = encode(queries.seq(int(rng.integers(queries.n))))
            tl = int(lengths[t])
            fl = int(min(rng.integers(20, 301), len(qs), tl))
            if fl <= 0:
                continue
            qa = int(rng.integers(0, len(qs) - fl + 1))
            ta = int(rng.integers(0, tl - fl + 1))
            frag = qs[qa:qa + fl].copy()
            k = rng.random(fl) < 0.15
            frag[k] = lut[rng.integers(0, 65536, int(k.sum()), dtype=np.uint16)]
            codes[off[t] + ta: off[t] + ta + fl] = frag
    return lengths.astype(np.uint16), codes


def shard_of(lengths: np.ndarray, codes: np.ndarray, shard: int, num_shards: int, tile: int = 16):
    """The sequences of shard `shard`: tiles (16 consecutive sorted sequences) t with t % num_shards == shard, back to
    back -- what one process per GPU hands to swg_gpu_load_db_shard.  Returns (local lengths, local codes, global
    index of every local sequence)."""
    n = len(lengths)
    if num_shards == 1:
        return lengths, codes, np.arange(n, dtype=np.int64)
    idx = np.arange(n, dtype=np.int64)
    mine = idx[(idx // tile) % num_shards == shard]
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lengths.astype(np.int64), out=off[1:])
    sl = lengths[mine].astype(np.int64)
    loff = np.zeros(len(mine) + 1, dtype=np.int64)
    np.cumsum(sl, out=loff[1:])
    gather = np.repeat(off[mine] - loff[:-1], sl) + np.arange(loff[-1], dtype=np.int64)
    return lengths[mine], codes[gather], mine


def workload(name: str, scale: float = 1.0):
    """Named BASELINE.json configurations -> (db SeqSet, queries SeqSet).  `scale` shrinks the
    sequence count (tests use small scales; bench uses 1.0)."""
    if name == "cfg1":       # q 144 vs 100k sequences
        q = make_queries(np.random.default_rng(42), [144])
        return make_db(42, max(32, int(100_000 * scale)), queries=q), q
    if name == "cfg2":       # Swiss-Prot-sized, 20 queries
        q = make_queries(np.random.default_rng(7), QUERY_LENGTHS)
        return make_db(2, max(32, int(570_000 * scale)), mu=5.675, queries=q), q
    if name == "cfg3":       # Environmental-NR-sized
        q = make_queries(np.random.default_rng(7), QUERY_LENGTHS)
        return make_db(3, max(32, int(6_000_000 * scale)), mu=5.2, sigma=0.6, queries=q), q
    raise KeyError(name)


def cache_dir() -> str:
    d = os.environ.get("SWIMM_B200_CACHE", "/tmp/swimm_b200_cache")
    os.makedirs(d, exist_ok=True)
    return d
