"""Parity of the long-sequence kernel (K3, csrc/wavefront_xw.cuh: the passes of one sequence run concurrently on
the warps of a CTA) with the oracle, through the C ABI.  Bar: bit-exact scores, identical hit order."""
import numpy as np
import pytest

from swimm_b200 import host, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "gpu-marked test collected without a CUDA device"
    from swimm_b200 import gpu as g
    s = g.GpuSearch(0)
    yield s
    s.close()


def _case(seed, lens, qlens, plant=0.2, copies=()):
    """Database with the given lengths, queries, planted fragments and (sequence, query, mutation rate) near-copies."""
    rng = np.random.default_rng(seed)
    q = synth.make_queries(rng, qlens)
    db = synth.make_seqset(rng, np.asarray(lens))
    synth.plant(rng, db, q, fraction=plant, frag_range=(5, 400), rate=0.1)
    for t, qi, rate in copies:
        s0 = db.offsets[t]
        L = min(len(q.seq(qi)), db.offsets[t + 1] - s0)
        db.residues[s0:s0 + L] = synth.mutate(rng, q.seq(qi)[:L], rate)
    _, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    do = np.zeros(db.n + 1, np.uint64)
    np.cumsum(dl.astype(np.uint64), out=do[1:])
    qo = np.zeros(q.n + 1, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    return qc, ql, qo, dc, dl, do


def _with(gpu, opts, fn):
    try:
        for k, v in opts.items():
            gpu.set_option(k, v)
        return fn()
    finally:
        for k in opts:
            gpu.set_option(k, 1 if k in ("query_pairing", "long_kernel") else 0)


@pytest.mark.parametrize("W,K", [(1, 5), (2, 1), (2, 3), (2, 16), (4, 2), (4, 7), (4, 17), (8, 4), (8, 16), (8, 32),
                                 (16, 1), (16, 9), (16, 16)])
@pytest.mark.parametrize("blocks", [0, 1])
def test_forced_shapes_every_tile_long(gpu, oracle, W, K, blocks):
    """Every (warps, rows) family of the long-sequence kernel, all tiles forced onto it; with one CTA every group runs
    dozens of sequence pairs back to back (task ring, step counters and the last-row rings wrap many times)."""
    m = W * 32 * K
    rng = np.random.default_rng(1000 + 37 * W + K)
    lens = np.concatenate([rng.integers(1, 90, 60), rng.integers(100, 2500, 180), [1, 4097, 5000]])
    qc, ql, qo, dc, dl, do = _case(50 + W + K, lens, [max(1, m // 3), max(1, m - 1), m])
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    got, keys = _with(gpu, {"long_threshold": 8, "xw_warps": W, "xw_rows": K, "grid_blocks": blocks, "query_pairing": 0},
                      lambda: gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 10, want_scores=True))
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]
    from swimm_b200.gpu import split_key
    for qi in range(len(ql)):
        ts, ti = oracle.top(want[qi], 10)
        ks, ki = split_key(keys[qi])
        assert np.array_equal(ki, ti.astype(np.int64)) and np.array_equal(ks, ts)


@pytest.mark.parametrize("pairing", [0, 2])
@pytest.mark.parametrize("threshold", [0, 8, 1000])
def test_split_between_main_and_long_kernels(gpu, oracle, pairing, threshold):
    """Tiles below the threshold on the sequence-pair / query-pair kernels, the others on the long-sequence kernel, side
    by side on two streams; non-default penalties (generic instantiations)."""
    rng = np.random.default_rng(5)
    lens = np.concatenate([rng.integers(20, 600, 900), rng.integers(1000, 6000, 60), [12000, 30000]])
    qc, ql, qo, dc, dl, do = _case(6, lens, [47, 144, 600, 1500, 2100], plant=0.05)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum50"), 11, 1)
    gpu.load_db(dl, dc)
    got, _ = _with(gpu, {"long_threshold": threshold, "query_pairing": pairing},
                   lambda: gpu.search(qc, ql, qo[:-1], host.submat("blosum50"), 11, 1, 0, want_scores=True))
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]


@pytest.mark.parametrize("W", [4, 8, 16])
def test_longest_sequence_32bit_across_pass_boundaries(gpu, oracle, W):
    """A 65535-column sequence (the format's maximum) through W >= 4 concurrent passes, with an alignment whose score
    passes 32767 while it crosses the pass boundaries: W x 3400 against W x 65535 scores 3400 * 11 = 37400 along a
    diagonal that visits every warp of the group.  The 16-bit kernel must list it, the 32-bit instantiation of the
    long-sequence kernel must recompute it exactly."""
    rng = np.random.default_rng(9)
    q = synth.make_queries(rng, [200, 3400])
    q.residues[q.offsets[1]:q.offsets[2]] = ord("W")
    lens = np.concatenate([rng.integers(3000, 9000, 29), [65535, 65535, 40000]])
    db = synth.make_seqset(rng, lens)
    db.residues[db.offsets[29]:db.offsets[30]] = ord("W")                    # all W: score 37400
    db.residues[db.offsets[30] + 20000:db.offsets[30] + 23000] = ord("W")   # 3000 x 11 = 33000 somewhere in the middle
    db.residues[db.offsets[31] + 100:db.offsets[31] + 2900] = ord("W")      # 2800 x 11 = 30800: stays in 16 bits
    _, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    do = np.zeros(db.n + 1, np.uint64)
    np.cumsum(dl.astype(np.uint64), out=do[1:])
    qo = np.zeros(q.n + 1, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    assert sorted(want[1])[-3:] == [30800, 33000, 37400]
    gpu.load_db(dl, dc)
    K = -(-3400 // (32 * W))
    got, _ = _with(gpu, {"long_threshold": 2000, "xw_warps": W, "xw_rows": K, "query_pairing": 0},
                   lambda: gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 0, want_scores=True))
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]
    assert gpu.stats()["rescored"] == 2


def test_long_kernel_off_matches(gpu, oracle):
    """Option long_kernel = 0 (the 32-thread shape of the sequence-pair kernel for long tiles) gives the same scores."""
    rng = np.random.default_rng(12)
    lens = np.concatenate([rng.integers(20, 500, 500), [5000, 9000, 20000]])
    qc, ql, qo, dc, dl, do = _case(13, lens, [144, 1200])
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    for lk in (0, 1):
        got, _ = _with(gpu, {"long_kernel": lk, "long_threshold": 1000, "query_pairing": 0},
                       lambda: gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 0, want_scores=True))
        assert np.array_equal(got, want), (lk, np.argwhere(got != want)[:8])


def _gappy_copy(rng, frag, p_ins=0.04, p_del=0.03, p_sub=0.08):
    """A homolog of `frag` with insertions, deletions and substitutions (alignments with gaps span more columns)."""
    out = []
    for c in frag:
        if rng.random() < p_del:
            continue
        out.append(c if rng.random() >= p_sub else synth.random_residues(rng, 1)[0])
        while rng.random() < p_ins:
            out.extend(synth.random_residues(rng, int(rng.integers(1, 12))))
    return np.array(out, dtype=np.uint8)


@pytest.mark.parametrize("chunk", [0, 1, 1200, 4096])
@pytest.mark.parametrize("matrix,go,ge", [("blosum62", 10, 2), ("blosum62", 11, 1), ("pam250", 5, 1)])
def test_short_queries_column_chunks_of_long_sequences(gpu, oracle, chunk, matrix, go, ge):
    """One-pass queries against long sequences: the long tiles are cut into overlapping column chunks (overlap = the
    longest span an alignment of the query can have), searched as independent tasks, and merged with max -- exact.
    Gappy homologs planted across chunk boundaries; chunk length by the planner (0), off (1) and forced."""
    rng = np.random.default_rng(31 + chunk + ge)
    q = synth.make_queries(rng, [25, 60, 144])
    lens = np.concatenate([rng.integers(30, 400, 600), rng.integers(6000, 30000, 24)])
    db = synth.make_seqset(rng, lens)
    for t in range(600, 624):                      # homologs at many offsets of every long sequence
        s0, L = int(db.offsets[t]), int(lens[t])
        for pos in range(150, L - 700, 997):
            qi = int(rng.integers(q.n))
            frag = _gappy_copy(rng, q.seq(qi))
            db.residues[s0 + pos:s0 + pos + len(frag)] = frag[:max(0, L - pos)][:len(frag)]
    _, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    do = np.zeros(db.n + 1, np.uint64)
    np.cumsum(dl.astype(np.uint64), out=do[1:])
    qo = np.zeros(q.n + 1, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    want = oracle.search(qc, qo, dc, do, host.submat(matrix), go, ge)
    gpu.load_db(dl, dc)
    for pairing in (0, 1):
        got, _ = _with(gpu, {"long_threshold": 2000, "chunk_columns": chunk, "query_pairing": pairing},
                       lambda: gpu.search(qc, ql, qo[:-1], host.submat(matrix), go, ge, 0, want_scores=True))
        assert np.array_equal(got, want), (pairing, np.argwhere(got != want)[:8])
