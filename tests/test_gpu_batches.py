"""Batch handling of the CUDA library through the C ABI: score-row window (large batches searched in chunks of queries),
streaming submit / poll with two batches in flight, sharded full-sort top-r, multi-GPU reference-signature entry."""
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


def _db(seed, n, hi=900):
    rng = np.random.default_rng(seed)
    db = synth.make_seqset(rng, synth.lognormal_lengths(rng, n, 4.6, 0.7, 1, hi))
    _, dl, dc = synth.length_sorted(db)
    do = np.zeros(db.n + 1, np.uint64)
    np.cumsum(dl.astype(np.uint64), out=do[1:])
    return db, dl, dc, do


def _queries(seed, lens, db=None):
    rng = np.random.default_rng(seed)
    q = synth.make_queries(rng, lens)
    if db is not None:
        synth.plant(rng, db, q, fraction=0.05, frag_range=(5, 200), rate=0.1)
    _, ql, qc = synth.length_sorted(q)
    qo = np.zeros(q.n + 1, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    return ql, qc, qo


def _check_keys(oracle, want, keys, top):
    from swimm_b200.gpu import split_key
    for qi in range(want.shape[0]):
        ts, ti = oracle.top(want[qi], top)
        ks, ki = split_key(keys[qi])
        assert np.array_equal(ki, ti.astype(np.int64)) and np.array_equal(ks, ts), qi


def test_score_window_chunks_match_one_piece(gpu, oracle):
    """25 queries with a 1 MB score budget on a 40000-sequence shard: 160 KB per row -> the batch is searched in five
    chunks of queries (each planned, paired and top-r-selected on its own); hit lists equal the oracle's, and a fetch of
    score rows after such a run is refused."""
    from swimm_b200.gpu import SwgError
    db, dl, dc, do = _db(3, 40000, hi=200)
    lens = list(np.random.default_rng(4).integers(20, 700, 25))
    ql, qc, qo = _queries(5, lens, db)
    _, dl, dc = synth.length_sorted(db)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    gpu.set_option("score_budget_mb", 1)
    try:
        _, keys = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 15)
        _check_keys(oracle, want, keys, 15)
        launches_chunked = gpu.stats()["launches"]
        with pytest.raises(SwgError):
            gpu.fetch(True, False)
        # the full score vectors are still available on request (keep_scores): one piece, whatever the budget
        got, keys2 = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 15, want_scores=True)
        assert np.array_equal(got, want) and np.array_equal(keys2, keys)
        assert launches_chunked > gpu.stats()["launches"]          # more chunks -> more (smaller) groups and selections
    finally:
        gpu.set_option("score_budget_mb", 4096)


def test_submit_poll_two_batches_in_flight(gpu, oracle):
    """Streaming: three batches through one resident database, the next one submitted before the previous is polled;
    a third submit without a poll is refused; results equal the blocking call's and the oracle's."""
    from swimm_b200.gpu import SwgError
    db, dl, dc, do = _db(11, 2500)
    batches = [_queries(20 + i, lens, db) for i, lens in enumerate([[40, 144, 333], [1200, 2100], [5, 77, 600, 601, 999]])]
    _, dl, dc = synth.length_sorted(db)
    gpu.load_db(dl, dc)
    b62 = host.submat("blosum62")
    t0 = gpu.submit(batches[0][1], batches[0][0], batches[0][2][:-1], b62, 10, 2, 12)
    t1 = gpu.submit(batches[1][1], batches[1][0], batches[1][2][:-1], b62, 10, 2, 12)
    with pytest.raises(SwgError):
        gpu.submit(batches[2][1], batches[2][0], batches[2][2][:-1], b62, 10, 2, 12)
    k0, s0 = gpu.poll(t0)
    t2 = gpu.submit(batches[2][1], batches[2][0], batches[2][2][:-1], b62, 10, 2, 12)
    k1, s1 = gpu.poll(t1)
    k2, s2 = gpu.poll(t2)
    assert s0 > 0 and s1 > 0 and s2 > 0
    for (ql, qc, qo), keys in zip(batches, [k0, k1, k2]):
        want = oracle.search(qc, qo, dc, do, b62, 10, 2)
        _check_keys(oracle, want, keys, 12)
        _, blocking = gpu.search(qc, ql, qo[:-1], b62, 10, 2, 12)
        assert np.array_equal(blocking, keys)


@pytest.mark.parametrize("shards", [2, 5])
def test_sharded_full_sort_top_beyond_local_count(gpu, oracle, shards):
    """`-r <n>` on a sharded context: top is clamped to the whole database, which exceeds what one shard holds (the
    full-sort path then has fewer keys than `top`: the rest of the row must read as "no hit", not as scratch memory)."""
    from swimm_b200.gpu import merge_top_keys
    db, dl, dc, do = _db(31, 3100, hi=300)
    ql, qc, qo = _queries(32, [60, 200], db)
    _, dl, dc = synth.length_sorted(db)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    top = db.n                           # > 2048: full sort; > the shard's padded count
    parts = []
    for s in range(shards):
        gpu.load_db(dl, dc, shard=s, num_shards=shards)
        _, keys = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, top)
        local = gpu.local_sequences
        assert (keys[:, local:] == 0).all(), "keys beyond the shard's sequences must be empty"
        parts.append(keys)
    merged = merge_top_keys(parts, top)
    _check_keys(oracle, want, merged, top)


def test_reference_signature_entry_on_all_gpus(oracle):
    """swimm_gpu_search_avx2_compat with n_threads = number of GPUs (0 = all visible): the database is sharded over them
    and every GPU fills its own entries of the reference-layout score array."""
    from swimm_b200 import gpu as g
    db, dl, dc, do = _db(41, 1000, hi=500)
    ql, qc, qo = _queries(42, [30, 144, 1100], db)
    _, dl, dc = synth.length_sorted(db)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    vdb, vlen, vblocks, vdisp = synth.interleave_reference(dl, dc)
    for n_threads in sorted({0, 1, g.device_count()}):
        scores, _ = g.compat_search_avx2(qc, ql, qo, vdb, vlen, vblocks, vdisp, host.submat("blosum62"), 10, 2,
                                         n_threads=n_threads)
        assert np.array_equal(scores[:, :db.n], want), n_threads
        assert (scores[:, db.n:] == 0).all()


def test_alignment_coordinates_of_the_hits(gpu, oracle):
    """Opt-in coordinate pass (no counterpart in the reference, which is score-only): for every hit of every query the
    start and end of an optimal alignment as oracle.align_ends defines them (ties: smallest database position, then
    smallest query position; start from the reversed prefixes).  Planted fragments, repeats with many tied cells,
    a query longer than a strip of 32 rows, hits in long sequences, two shards."""
    from swimm_b200.gpu import split_key
    rng = np.random.default_rng(17)
    q = synth.make_queries(rng, [9, 31, 33, 150, 700])
    q.residues[q.offsets[0]:q.offsets[1]] = ord("W")                      # WWWWWWWWW: every cell of a W run ties
    lens = np.concatenate([synth.lognormal_lengths(rng, 400, 4.8, 0.7, 1, 900), [3000, 5000]])
    db = synth.make_seqset(rng, lens)
    synth.plant(rng, db, q, fraction=0.2, frag_range=(5, 600), rate=0.1)
    db.residues[db.offsets[400] + 100:db.offsets[400] + 140] = ord("W")   # a W run inside a long sequence
    db.residues[db.offsets[7]:db.offsets[8]] = ord("W")
    _, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    do = np.zeros(db.n + 1, np.uint64)
    np.cumsum(dl.astype(np.uint64), out=do[1:])
    qo = np.zeros(q.n + 1, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    b62 = host.submat("blosum62")
    for shard, shards in [(0, 1), (1, 2)]:
        gpu.load_db(dl, dc, shard=shard, num_shards=shards)
        _, keys = gpu.search(qc, ql, qo[:-1], b62, 10, 2, 8)
        coords = gpu.align_ends()
        assert coords.shape == (q.n, 8, 4)
        checked = 0
        for qi in range(q.n):
            ks, ki = split_key(keys[qi])
            for h in range(8):
                if ks[h] == 0:
                    assert (coords[qi, h] == -1).all()
                    continue
                d = dc[int(do[ki[h]]):int(do[ki[h] + 1])]
                s, c = oracle.align_ends(qc[qo[qi]:qo[qi + 1]], d, b62, 10, 2)
                assert s == ks[h] and np.array_equal(coords[qi, h], c), (shard, qi, h, ks[h], coords[qi, h], c)
                assert 0 <= c[0] <= c[1] < ql[qi] and 0 <= c[2] <= c[3] < dl[ki[h]]
                checked += 1
        assert checked >= 30


def test_256_queries_one_million_sequences_under_a_score_budget(gpu, oracle):
    """256 queries x 1 000 000 sequences would keep 1 GB of score rows; under a 256 MB budget the batch is searched in
    four chunks of 64 queries.  Size-independent checks: the hit lists equal those of the one-piece run (4 GB budget),
    and -- for a few queries -- the oracle's on a sample of the database that contains all their hits."""
    from swimm_b200.gpu import split_key
    rng = np.random.default_rng(77)
    q = synth.make_queries(rng, list(rng.integers(25, 90, 256)))
    dl, dc = synth.sorted_db(78, 1_000_000, mu=3.6, sigma=0.4, lo=8, hi=400, queries=q, plant_fraction=0.002)
    _, ql, qc = synth.length_sorted(q)
    qo = np.zeros(q.n + 1, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    b62 = host.submat("blosum62")
    gpu.load_db(dl, dc)
    try:
        gpu.set_option("score_budget_mb", 256)
        _, chunked = gpu.search(qc, ql, qo[:-1], b62, 10, 2, 10)
        launches_chunked = gpu.stats()["launches"]
        gpu.set_option("score_budget_mb", 4096)
        _, whole = gpu.search(qc, ql, qo[:-1], b62, 10, 2, 10)
        assert gpu.stats()["launches"] != launches_chunked
    finally:
        gpu.set_option("score_budget_mb", 4096)
    assert np.array_equal(chunked, whole)
    # oracle on a sample: the hit sequences of five queries + every 500th sequence
    off = np.zeros(len(dl) + 1, np.int64)
    np.cumsum(dl.astype(np.int64), out=off[1:])
    for qi in (0, 63, 64, 200, 255):
        ks, ki = split_key(chunked[qi])
        sub = np.unique(np.concatenate([ki, np.arange(0, len(dl), 500)]))
        sl = dl[sub]
        so = np.zeros(len(sub) + 1, np.uint64)
        np.cumsum(sl.astype(np.uint64), out=so[1:])
        idx = np.repeat(off[sub] - so[:-1].astype(np.int64), sl.astype(np.int64)) + np.arange(int(so[-1]))
        want = oracle.search(qc[qo[qi]:qo[qi + 1]], np.array([0, ql[qi]], np.uint32), dc[idx], so, b62, 10, 2)[0]
        pos = {int(g): i for i, g in enumerate(sub)}
        assert np.array_equal(want[[pos[int(g)] for g in ki]], ks)
        keys_sub = (want.astype(np.uint64) << np.uint64(32)) | sub.astype(np.uint64)
        assert np.isin(sub[keys_sub > np.uint64(chunked[qi, -1])], ki).all()
