"""Parity of the query-pair kernel (wavefront_q2.cuh: two queries in the 16-bit halves of a register, one launch per
pass, per-sequence pass lines in HBM) with the oracle and the golden vectors of the unmodified reference.  The
planner only pairs queries when it expects a gain; these tests force pairing (`query_pairing` = 2) and every
shape, so that the path is covered whatever the planner would choose.  Bar: bit-exact scores, identical hit order."""
import numpy as np
import pytest

from swimm_b200 import host, synth
from tests.helpers import GOLDEN_CASES, GoldenCase
from tests.test_gpu_parity import _keys_to_order, _random_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "gpu-marked test collected without a CUDA device"
    from swimm_b200 import gpu as g
    s = g.GpuSearch(0)
    s.set_option("query_pairing", 2)
    yield s
    s.close()


@pytest.fixture
def forced(gpu):
    """Set the forced shape of the query-pair kernel for one test and always reset it."""
    def _set(G, K):
        gpu.set_option("q2_group", G)
        gpu.set_option("q2_rows", K)
    yield _set
    gpu.set_option("q2_group", 0)
    gpu.set_option("q2_rows", 0)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_scores_and_order_with_pairs(gpu, name):
    case = GoldenCase(name)
    gpu.load_db(case.db_len, case.db_codes)
    for matrix, go, ge, exp_scores, order in case.runs():
        scores, keys = gpu.search(case.q_codes, case.q_len, case.q_off[:-1], host.submat(matrix), go, ge, case.db.n,
                                  want_scores=True)
        assert np.array_equal(scores, exp_scores), (name, matrix, go, ge, np.argwhere(scores != exp_scores)[:5])
        ks, ki = _keys_to_order(keys)
        assert np.array_equal(ki, order), (name, matrix, "hit order")
    if name == "overflow":
        assert gpu.stats()["rescored"] > 0


@pytest.mark.parametrize("G", [8, 16, 32])
@pytest.mark.parametrize("K", [8, 10, 14, 16, 22, 24, 30, 32])
def test_every_shape_single_and_multi_pass(gpu, oracle, forced, G, K):
    """Queries of exactly one pass, one row more, and three passes and a bit; an odd query count leaves one query to
    the sequence-pair kernel; the two queries of a pair differ in length by up to 2x.  (A pair that needs several
    passes always runs 32-thread groups, whatever group size is forced; K = 4n + 2 uses half of its last chunk.)"""
    R = G * K
    qlens = [max(1, R // 2), R, R - 1, R + 1, 2 * R + 3, 3 * R + 5, 7]
    qc, ql, qo, dc, dl, do = _random_case(1000 + 7 * G + K, 260, qlens, hi=500)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    forced(G, K)
    got, keys = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 15, want_scores=True)
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]
    for qi in range(len(ql)):
        ts, ti = oracle.top(want[qi], 15)
        ks, ki = _keys_to_order(keys[qi])
        assert np.array_equal(ki, ti.astype(np.int64)) and np.array_equal(ks, ts)


@pytest.mark.parametrize("blocks", [1, 2])
@pytest.mark.parametrize("shape", [(32, 32), (32, 12), (16, 24), (8, 8)])
def test_pairs_many_tasks_per_warp(gpu, oracle, forced, blocks, shape):
    """With one or two CTAs every thread group runs dozens of sequences back to back (real -> real segment
    transitions, pass lines of neighbouring sequences, the carried last column), single- and multi-pass."""
    qc, ql, qo, dc, dl, do = _random_case(950 + blocks + shape[1], 1200, [100, 250, 900, 1100, 2100, 2600], mu=4.2, sigma=0.9,
                                          hi=1200, plant=0.2)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    forced(*shape)
    gpu.set_option("grid_blocks", blocks)
    try:
        got, _ = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 0, want_scores=True)
    finally:
        gpu.set_option("grid_blocks", 0)
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_streams_random_batches(gpu, oracle, seed):
    """The two lanes as streams: batches of 5 to 11 queries of random lengths (1 ... 3000 rows, duplicates and
    zero-length included), so that queries start and end in either lane at unrelated launches -- fresh lane next to a
    continuing one (cleared line halves), idle lane tails, scores merged over several launches.  Planner's own choice
    of launch heights and of the stream threshold; one CTA for two of the seeds (dozens of sequences per thread group)."""
    rng = np.random.default_rng(700 + seed)
    n = int(rng.integers(5, 12))
    qlens = sorted(int(x) for x in np.concatenate([rng.integers(1, 3000, n - 2), [1100, 1100]]))
    qc, ql, qo, dc, dl, do = _random_case(800 + seed, 320, qlens, mu=4.6, sigma=0.8, hi=1500, plant=0.15)
    if seed == 4:                                   # a zero-length query in the batch
        ql = np.concatenate([[0], ql]).astype(np.uint16)
        qo = np.concatenate([[0], qo]).astype(np.uint32)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    gpu.set_option("grid_blocks", 1 if seed % 2 else 0)
    try:
        got, keys = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 5, want_scores=True)
        kinds = gpu.query_kernels()
    finally:
        gpu.set_option("grid_blocks", 0)
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]
    assert kinds.sum() >= len(ql) - 1               # everything but at most one query ran the query-pair kernel
    for qi in range(len(ql)):
        ts, ti = oracle.top(want[qi], 5)
        ks, ki = _keys_to_order(keys[qi])
        assert np.array_equal(ki, ti.astype(np.int64)) and np.array_equal(ks, ts)


def test_planner_shapes_vs_oracle(gpu, oracle):
    """The shapes the planner itself picks for the benchmark's query lengths (scaled-down database)."""
    qlens = [144, 189, 222, 375, 464, 567, 657, 729, 850, 1000, 1500, 2005]
    qc, ql, qo, dc, dl, do = _random_case(4242, 400, qlens, hi=900)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    got, _ = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 0, want_scores=True)
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]


def test_matrix_and_penalty_sweep_with_pairs(gpu, oracle):
    """Generic-penalty instantiations (everything but 10/2) and every matrix family, multi-query batch."""
    qc, ql, qo, dc, dl, do = _random_case(31, 300, [30, 90, 150, 222, 700, 710])
    gpu.load_db(dl, dc)
    for mat in ["blosum45", "blosum80", "pam30", "pam250"]:
        for go, ge in [(5, 1), (8, 3), (10, 2), (12, 1), (0, 0), (127, 127)]:
            want = oracle.search(qc, qo, dc, do, host.submat(mat), go, ge)
            got, _ = gpu.search(qc, ql, qo[:-1], host.submat(mat), go, ge, 0, want_scores=True)
            assert np.array_equal(got, want), (mat, go, ge)


@pytest.mark.parametrize("shape", [(32, 32), (32, 8), (8, 16)])
def test_pairs_long_sequences_and_32bit_rescore(gpu, oracle, forced, shape):
    """Scores above 32767 in one lane of a pair, in both lanes, and crossing pass boundaries; database sequences of
    1 to 40000 residues (tiles shorter than the minimum segment and far longer than a pass line prefetch)."""
    rng = np.random.default_rng(78)
    q = synth.make_queries(rng, [300, 6800, 7000, 7001])
    lens = np.concatenate([rng.integers(3001, 9000, 24), rng.integers(1, 60, 40), rng.integers(50, 400, 40),
                           [7500, 8000, 20000, 40000]])
    db = synth.make_seqset(rng, lens)
    for qi, targets in [(2, [(104, 0.0), (105, 0.02), (3, 0.3)]), (3, [(106, 0.01)]), (1, [(107, 0.0), (5, 0.25)])]:
        big = q.seq(qi)
        for t, rate in targets:
            s = db.offsets[t]
            L = min(len(big), db.offsets[t + 1] - s)
            db.residues[s:s + L] = synth.mutate(rng, big[:L], rate)
    _, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    do = np.zeros(db.n + 1, np.uint64)
    np.cumsum(dl.astype(np.uint64), out=do[1:])
    qo = np.zeros(q.n + 1, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    assert (want.max(axis=1)[1:] > 32767).all()
    gpu.load_db(dl, dc)
    forced(*shape)
    got, keys = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 10, want_scores=True)
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]
    assert gpu.stats()["rescored"] >= 4
    for qi in range(q.n):
        ts, ti = oracle.top(want[qi], 10)
        ks, ki = _keys_to_order(keys[qi])
        assert np.array_equal(ki, ti.astype(np.int64)) and np.array_equal(ks, ts)


@pytest.mark.parametrize("shards", [2, 5])
def test_pairs_sharded_search_merges_to_unsharded(gpu, oracle, shards):
    from swimm_b200.gpu import merge_top_keys
    qc, ql, qo, dc, dl, do = _random_case(62, 900, [50, 144, 150, 1100])
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    top = 20
    parts, all_scores = [], np.zeros_like(want)
    for s in range(shards):
        gpu.load_db(dl, dc, shard=s, num_shards=shards)
        sc, keys = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, top, want_scores=True)
        parts.append(keys)
        all_scores |= sc
    assert np.array_equal(all_scores, want)
    merged = merge_top_keys(parts, top)
    for qi in range(len(ql)):
        ts, ti = oracle.top(want[qi], top)
        ks, ki = _keys_to_order(merged[qi])
        assert np.array_equal(ki, ti.astype(np.int64)) and np.array_equal(ks, ts)


def test_pairs_equal_single_query_path_at_scale(gpu):
    """Size-independent property on a database too large for the oracle: the query-pair kernel and the
    sequence-pair kernel (two independent implementations of the recurrence) give identical score vectors."""
    rng = np.random.default_rng(11)
    q = synth.make_queries(rng, [144, 189, 850, 1000, 2005, 2504])
    db = synth.make_db(10, 40_000, queries=q)
    _, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    qo = np.zeros(q.n + 1, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    b62 = host.submat("blosum62")
    gpu.load_db(dl, dc)
    paired, keys_p = gpu.search(qc, ql, qo[:-1], b62, 10, 2, 10, want_scores=True)
    gpu.set_option("query_pairing", 0)
    try:
        single, keys_s = gpu.search(qc, ql, qo[:-1], b62, 10, 2, 10, want_scores=True)
    finally:
        gpu.set_option("query_pairing", 2)
    assert np.array_equal(paired, single)
    assert np.array_equal(keys_p, keys_s)
    assert paired.max() > 100          # the planted homologs are there


def test_without_pass_lines_only_single_launch_pairs(gpu, oracle):
    """`pass_lines` = 0 (or not enough device memory for 8 bytes per database column): queries longer than one pass
    fall back to the sequence-pair kernel, the short ones are still paired; same scores."""
    qc, ql, qo, dc, dl, do = _random_case(77, 300, [100, 130, 600, 640, 1500, 2600])
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    gpu.set_option("pass_lines", 0)
    try:
        got, _ = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 0, want_scores=True)
        kinds = gpu.query_kernels()
        st = gpu.stats()
    finally:
        gpu.set_option("pass_lines", 1)
    assert np.array_equal(got, want)
    assert list(kinds) == [1, 1, 1, 1, 0, 0] and st["pair_launches"] == 2


def test_pairs_degenerate_inputs(gpu):
    b62 = host.submat("blosum62")
    # two queries of one residue against one sequence of one residue
    gpu.load_db(np.array([1], np.uint16), np.array([5], np.int8))
    sc, keys = gpu.search(np.array([5, 7], np.int8), np.array([1, 1], np.uint16), np.array([0, 1], np.uint32), b62, 10, 2, 3,
                          want_scores=True)
    assert sc[0, 0] == b62[5, 5] and sc[1, 0] == max(0, int(b62[7, 5]))
    # a zero-length query paired with a real one
    gpu.load_db(np.array([3, 3, 4], np.uint16), np.arange(10, dtype=np.int8))
    sc, _ = gpu.search(np.array([1, 2, 3], np.int8), np.array([0, 3], np.uint16), np.array([0, 0], np.uint32), b62, 10, 2, 0,
                       want_scores=True)
    assert (sc[0] == 0).all() and sc[1].max() > 0
    # an empty database
    gpu.load_db(np.zeros(0, np.uint16), np.zeros(0, np.int8))
    sc, keys = gpu.search(np.array([5, 6], np.int8), np.array([1, 1], np.uint16), np.array([0, 1], np.uint32), b62, 10, 2, 3,
                          want_scores=True)
    assert sc.shape == (2, 0) and (keys == 0).all()


@pytest.mark.parametrize("shape", [(32, 8), (32, 30), (16, 10), (8, 8)])
def test_full_length_sequences_last_residue_across_segment_boundaries(gpu, oracle, forced, shape):
    """Every sequence fills its tile's columns to the last one (lengths are multiples of 8, equal inside a tile, so no
    padding separates consecutive sequences) and ends with a copy of a query's tail: the last residue of a sequence
    reaches thread 0 in the first step of the NEXT sequence (the other one of the pair, or another pair) and must
    still be read as this sequence's residue.  One CTA, so that every group runs many sequences back to back."""
    rng = np.random.default_rng(77 + shape[1])
    q = synth.make_queries(rng, [90, 150, 700, 1300])
    lens = np.repeat(np.array([88, 96, 96, 160, 168, 400, 408, 1024]), 16)
    db = synth.make_seqset(rng, lens)
    for t in range(db.n):
        qi = int(rng.integers(q.n))
        tail = q.seq(qi)[-min(60, int(lens[t])):]
        db.residues[db.offsets[t + 1] - len(tail):db.offsets[t + 1]] = tail
    _, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    do = np.zeros(db.n + 1, np.uint64)
    np.cumsum(dl.astype(np.uint64), out=do[1:])
    qo = np.zeros(q.n + 1, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    forced(*shape)
    gpu.set_option("grid_blocks", 1)
    try:
        got, _ = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 0, want_scores=True)
    finally:
        gpu.set_option("grid_blocks", 0)
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]
