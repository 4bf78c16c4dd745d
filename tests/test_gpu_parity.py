"""Parity of the CUDA path (through the C ABI, include/swimm_gpu.h) with the oracle and with the golden
vectors of the unmodified reference.  Integer work: the bar is bit-exact scores and identical hit order."""
import numpy as np
import pytest

from swimm_b200 import host, synth
from tests.helpers import GOLDEN_CASES, GoldenCase

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "gpu-marked test collected without a CUDA device"
    from swimm_b200 import gpu as g
    s = g.GpuSearch(0)
    yield s
    s.close()


def _keys_to_order(keys):
    from swimm_b200.gpu import split_key
    return split_key(keys)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_scores_and_order(gpu, name):
    case = GoldenCase(name)
    gpu.load_db(case.db_len, case.db_codes)
    for matrix, go, ge, exp_scores, order in case.runs():
        scores, keys = gpu.search(case.q_codes, case.q_len, case.q_off[:-1], host.submat(matrix), go, ge, case.db.n,
                                  want_scores=True)
        assert np.array_equal(scores, exp_scores), (name, matrix, go, ge, np.argwhere(scores != exp_scores)[:5])
        ks, ki = _keys_to_order(keys)
        assert np.array_equal(ki, order), (name, matrix, "hit order")
        assert np.array_equal(ks, np.take_along_axis(exp_scores.astype(np.int64), order, axis=1))
    st = gpu.stats()
    if name == "overflow":
        assert st["rescored"] > 0          # the 32-bit kernel must have run


@pytest.mark.parametrize("top", [1, 10, 100])
def test_golden_top_r_small(gpu, top):
    case = GoldenCase("basic")
    gpu.load_db(case.db_len, case.db_codes)
    matrix, go, ge, exp_scores, order = next(case.runs())
    _, keys = gpu.search(case.q_codes, case.q_len, case.q_off[:-1], host.submat(matrix), go, ge, top)
    ks, ki = _keys_to_order(keys)
    assert np.array_equal(ki, order[:, :top])


def _random_case(seed, n, qlens, mu=4.5, sigma=0.7, hi=1500, plant=0.05):
    rng = np.random.default_rng(seed)
    q = synth.make_queries(rng, qlens)
    db = synth.make_seqset(rng, synth.lognormal_lengths(rng, n, mu, sigma, 1, hi))
    synth.plant(rng, db, q, fraction=plant, frag_range=(5, 200), rate=0.1)
    perm, dlen, dcodes = synth.length_sorted(db)
    _, qlen, qcodes = synth.length_sorted(q)
    doff = np.zeros(db.n + 1, np.uint64)
    np.cumsum(dlen.astype(np.uint64), out=doff[1:])
    qoff = np.zeros(q.n + 1, np.uint32)
    np.cumsum(qlen.astype(np.uint32), out=qoff[1:])
    return qcodes, qlen, qoff, dcodes, dlen, doff


@pytest.mark.parametrize("qlens", [[1, 2, 3], [4, 17, 31, 32, 33], [63, 64, 65, 127, 128, 129], [144, 255, 257],
                                   [511, 513, 700], [1023, 1025], [1500, 2100]])
def test_random_vs_oracle_query_shapes(gpu, oracle, qlens):
    """Every (group size, rows per thread, passes) shape the planner can choose."""
    qc, ql, qo, dc, dl, do = _random_case(100 + len(qlens) + qlens[0], 700, qlens)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    got, keys = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 25, want_scores=True)
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]
    for qi in range(len(ql)):
        ts, ti = oracle.top(want[qi], 25)
        ks, ki = _keys_to_order(keys[qi])
        assert np.array_equal(ki, ti.astype(np.int64)) and np.array_equal(ks, ts)


@pytest.mark.parametrize("G", [4, 8, 16, 32])
@pytest.mark.parametrize("K", [1, 2, 7, 16, 17, 32])
def test_forced_shapes_vs_oracle(gpu, oracle, G, K):
    m = G * K - (1 if K > 1 else 0)
    qc, ql, qo, dc, dl, do = _random_case(7 * G + K, 300, [max(1, m // 2), m], hi=600)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum50"), 11, 1)
    gpu.load_db(dl, dc)
    gpu.set_option("force_group", G)
    gpu.set_option("force_rows", K)
    try:
        got, _ = gpu.search(qc, ql, qo[:-1], host.submat("blosum50"), 11, 1, 0, want_scores=True)
    finally:
        gpu.set_option("force_group", 0)
        gpu.set_option("force_rows", 0)
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]


@pytest.mark.parametrize("blocks", [1, 2])
@pytest.mark.parametrize("qlens,force", [([144, 375, 1000], (0, 0)), ([1500, 2005, 2600], (0, 0)), ([700, 1300], (32, 20)),
                                         ([9000], (0, 0))])
def test_many_tasks_per_warp(gpu, oracle, blocks, qlens, force):
    """Regression: with one or two CTAs every warp runs dozens of tasks back to back, so whatever one task leaves
    behind (scratch-line columns, the prefetch ring, the carried last column, the running best) meets the next
    one.  Single-pass, multi-pass and global-profile shapes.  (A ring slot that kept the previous task's last column
    through pass 0 of the next task once slipped through the few-tasks-per-warp cases.)"""
    qc, ql, qo, dc, dl, do = _random_case(900 + len(qlens) + blocks, 1200, qlens, mu=4.2, sigma=0.9, hi=1200, plant=0.2)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    gpu.set_option("grid_blocks", blocks)
    gpu.set_option("query_pairing", 0)
    gpu.set_option("force_group", force[0])
    gpu.set_option("force_rows", force[1])
    try:
        got, _ = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 0, want_scores=True)
    finally:
        gpu.set_option("grid_blocks", 0)
        gpu.set_option("query_pairing", 1)
        gpu.set_option("force_group", 0)
        gpu.set_option("force_rows", 0)
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]


def test_matrix_and_penalty_sweep(gpu, oracle):
    qc, ql, qo, dc, dl, do = _random_case(31, 400, [30, 90, 150, 222])
    gpu.load_db(dl, dc)
    for mat in ["blosum45", "blosum80", "pam30", "pam250"]:
        for go in [5, 8, 10, 12]:
            for ge in [1, 2, 3]:
                want = oracle.search(qc, qo, dc, do, host.submat(mat), go, ge)
                got, _ = gpu.search(qc, ql, qo[:-1], host.submat(mat), go, ge, 0, want_scores=True)
                assert np.array_equal(got, want), (mat, go, ge)


@pytest.mark.parametrize("long_threshold", [0, 3072, 64])
def test_long_sequences_and_32bit_rescore(gpu, oracle, long_threshold):
    """cfg4 in miniature: database sequences above the long-sequence threshold, planted near-copies of a long
    query (score > 32767 -> 32-bit kernel) and partial homologs."""
    rng = np.random.default_rng(77)
    q = synth.make_queries(rng, [300, 7000])
    lens = np.concatenate([rng.integers(3001, 9000, 40), rng.integers(50, 400, 60), [7500, 8000, 20000, 40000]])
    db = synth.make_seqset(rng, lens)
    big = q.seq(1)
    for t, rate in [(100, 0.0), (101, 0.02), (3, 0.3), (41, 0.3)]:
        s = db.offsets[t]
        L = min(len(big), db.offsets[t + 1] - s)
        db.residues[s:s + L] = synth.mutate(rng, big[:L], rate)
    perm, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    do = np.zeros(db.n + 1, np.uint64)
    np.cumsum(dl.astype(np.uint64), out=do[1:])
    qo = np.zeros(q.n + 1, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    assert want.max() > 32767
    gpu.load_db(dl, dc)
    gpu.set_option("long_threshold", long_threshold)     # 0 = per-query estimate; > 0 = fixed column count
    try:
        got, keys = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 10, want_scores=True)
    finally:
        gpu.set_option("long_threshold", 0)
    assert np.array_equal(got, want), np.argwhere(got != want)[:8]
    assert gpu.stats()["rescored"] > 0
    for qi in range(2):
        ts, ti = oracle.top(want[qi], 10)
        ks, ki = _keys_to_order(keys[qi])
        assert np.array_equal(ki, ti.astype(np.int64)) and np.array_equal(ks, ts)


def test_very_long_query_global_profile(gpu, oracle):
    """More passes than fit in shared memory: the kernel variant that reads the profile through L1."""
    qc, ql, qo, dc, dl, do = _random_case(5, 64, [9000], hi=300)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    got, _ = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, 0, want_scores=True)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_search_merges_to_unsharded(gpu, oracle, shards):
    """Multi-GPU path on one device: every shard searched in turn, hit lists merged on the host."""
    from swimm_b200.gpu import merge_top_keys
    qc, ql, qo, dc, dl, do = _random_case(61, 1000, [50, 144])
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    top = 20
    parts, all_scores = [], np.zeros_like(want)
    residues = 0
    for s in range(shards):
        gpu.load_db(dl, dc, shard=s, num_shards=shards)
        residues += gpu.local_residues
        sc, keys = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, top, want_scores=True)
        parts.append(keys)
        all_scores |= sc            # each shard writes only its own sequences, the rest stay 0
    assert residues == int(dl.astype(np.int64).sum())
    assert np.array_equal(all_scores, want)
    merged = merge_top_keys(parts, top)
    for qi in range(len(ql)):
        ts, ti = oracle.top(want[qi], top)
        ks, ki = _keys_to_order(merged[qi])
        assert np.array_equal(ki, ti.astype(np.int64)) and np.array_equal(ks, ts)
        one = host.merge_top_keys(np.stack([p[qi] for p in parts]), top)
        assert np.array_equal(one, merged[qi])


def test_reference_signature_entry_point(oracle):
    """swimm_gpu_search_avx2_compat takes exactly what swimm.c:74-76 hands to cpu_search_avx2_sp."""
    from swimm_b200 import gpu as g
    qc, ql, qo, dc, dl, do = _random_case(88, 150, [21, 144])
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    # the reference's query padding (odd -> even with the dummy residue, sequences.c:367-387)
    m = (ql + (ql & 1)).astype(np.uint16)
    qdisp = np.zeros(len(ql) + 1, np.uint32)
    np.cumsum(m, out=qdisp[1:])
    qpad = np.full(int(qdisp[-1]), 23, np.int8)
    for i in range(len(ql)):
        qpad[qdisp[i]:qdisp[i] + ql[i]] = qc[qo[i]:qo[i + 1]]
    vdb, vlen, vblocks, vdisp = synth.interleave_reference(dl, dc, 32, 60)
    scores, work = g.compat_search_avx2(qpad, m, qdisp, vdb, vlen, vblocks, vdisp, host.submat("blosum62"), 10, 2)
    n = len(dl)
    assert np.array_equal(scores[:, :n], want)
    assert (scores[:, n:] == 0).all()
    assert work > 0


def test_round_trip_properties_at_scale(gpu):
    """Size-independent checks on a database too large for the oracle to finish in seconds:
    self-hit = sum of diagonal scores, top-r idempotent under re-sharding, scores symmetric in (q, d) roles."""
    rng = np.random.default_rng(3)
    q = synth.make_queries(rng, [200])
    db = synth.make_db(9, 60_000, queries=q)
    # plant the query itself: its score against the copy is the sum of the matrix diagonal over its residues
    t = int(np.argmax(db.lengths >= 400))
    db.residues[db.offsets[t] + 7: db.offsets[t] + 207] = q.seq(0)
    perm, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    b62 = host.submat("blosum62")
    gpu.load_db(dl, dc)
    scores, keys = gpu.search(qc, ql, np.zeros(1, np.uint32), b62, 10, 2, 10, want_scores=True)
    self_score = int(sum(int(b62[c, c]) for c in qc))
    pos = int(np.where(perm == t)[0][0])
    assert scores[0, pos] == self_score
    from swimm_b200.gpu import split_key, merge_top_keys
    ks, ki = split_key(keys[0])
    assert ki[0] == pos and ks[0] == self_score
    # order property of the keys: strictly descending
    assert (np.diff(keys[0].astype(np.int64)) < 0).all()
    # the same top-10 from 4 shards
    parts = []
    for s in range(4):
        gpu.load_db(dl, dc, shard=s, num_shards=4)
        _, k = gpu.search(qc, ql, np.zeros(1, np.uint32), b62, 10, 2, 10)
        parts.append(k)
    assert np.array_equal(merge_top_keys(parts, 10), keys)
    # numpy check of the top-10 against the full score vector
    order = np.lexsort((-np.arange(len(dl)), -scores[0].astype(np.int64)))[:10]
    assert np.array_equal(ki, order)


@pytest.mark.parametrize("name", ["basic", "edge"])
def test_cli_search_prints_the_reference_hit_lists(tmp_path, name):
    """`swimm -S preprocess` + `swimm -S search -m 3` (the C host driving the C ABI): the printed hit lists
    equal the ones the unmodified reference printed for the same FASTA files (tests/golden/*.json)."""
    import json
    import os
    import re
    import subprocess
    from tests.helpers import GOLDEN, ROOT
    exe = os.path.join(ROOT, "swimm_b200", "swimm")
    prefix = str(tmp_path / "db")
    subprocess.run([exe, "-S", "preprocess", "-i", os.path.join(GOLDEN, name + ".db.fasta"), "-o", prefix], check=True,
                   stdout=subprocess.DEVNULL)
    meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
    run = meta["runs"][0]
    out = subprocess.run([exe, "-S", "search", "-q", os.path.join(GOLDEN, name + ".q.fasta"), "-d", prefix, "-m", "3",
                          "-r", str(meta["n"]), "-s", run["matrix"], "-g", str(run["go"]), "-e", str(run["ge"])],
                         check=True, capture_output=True, text=True).stdout
    queries, cur = [], None
    for line in out.split("\n"):
        if line.startswith("Query no."):
            cur = []
            queries.append(cur)
        m = re.match(r"^(-?\d+)\t.*syn\|(\d+)\|", line)
        if m and cur is not None:
            cur.append([int(m.group(1)), int(m.group(2))])
    assert queries == run["hits"]
    assert "Search speed:" in out and "Execution mode:\t\t\tB200 GPU only (1 GPU" in out


def test_cli_several_query_files_against_the_resident_database(tmp_path):
    """`-q a.fasta,b.fasta`: the database is uploaded once, every file gets its own report, and each report's hit
    lists equal the ones of a separate run (and therefore the reference's)."""
    import json
    import os
    import re
    import subprocess
    from tests.helpers import GOLDEN, ROOT
    exe = os.path.join(ROOT, "swimm_b200", "swimm")
    prefix = str(tmp_path / "db")
    subprocess.run([exe, "-S", "preprocess", "-i", os.path.join(GOLDEN, "basic.db.fasta"), "-o", prefix], check=True,
                   stdout=subprocess.DEVNULL)
    meta = json.load(open(os.path.join(GOLDEN, "basic.json")))
    run = meta["runs"][0]
    qa, qb = os.path.join(GOLDEN, "basic.q.fasta"), os.path.join(GOLDEN, "edge.q.fasta")
    common = ["-d", prefix, "-m", "3", "-r", str(meta["n"]), "-s", run["matrix"], "-g", str(run["go"]), "-e", str(run["ge"])]

    def reports(files):
        out = subprocess.run([exe, "-S", "search", "-q", files] + common, check=True, capture_output=True, text=True).stdout
        parts = out.split("Query filename:")[1:]
        res = []
        for part in parts:
            hits = [[int(m.group(1)), int(m.group(2))] for m in re.finditer(r"^(-?\d+)\t.*syn\|(\d+)\|", part, re.M)]
            res.append((part.split("\n")[0].strip(), part.count("Query no."), hits))
        return res

    both = reports(qa + "," + qb)
    assert [r[0] for r in both] == [qa, qb]
    assert both[0] == reports(qa)[0] and both[1] == reports(qb)[0]
    flat = [h for q in run["hits"] for h in q]
    assert both[0][2] == flat                                   # the first file: the reference's own printed lists


def test_full_sort_path_top_larger_than_select_limit(gpu, oracle):
    """`-r <n>` with n > 2048: the bitonic full sort instead of the selection kernels."""
    qc, ql, qo, dc, dl, do = _random_case(17, 5000, [64], mu=3.8, sigma=0.5, hi=200, plant=0.02)
    want = oracle.search(qc, qo, dc, do, host.submat("blosum62"), 10, 2)
    gpu.load_db(dl, dc)
    for top in (2049, 5000, 7000):
        _, keys = gpu.search(qc, ql, qo[:-1], host.submat("blosum62"), 10, 2, top)
        n = min(top, 5000)
        ts, ti = oracle.top(want[0], n)
        ks, ki = _keys_to_order(keys[0][:n])
        assert np.array_equal(ki, ti.astype(np.int64)) and np.array_equal(ks, ts)
        assert (keys[0][n:] == 0).all()


def test_degenerate_inputs(gpu, oracle):
    b62 = host.submat("blosum62")
    # a database of one sequence of one residue, a query of one residue
    gpu.load_db(np.array([1], np.uint16), np.array([5], np.int8))
    sc, keys = gpu.search(np.array([5], np.int8), np.array([1], np.uint16), np.array([0], np.uint32), b62, 10, 2, 3,
                          want_scores=True)
    assert sc.shape == (1, 1) and sc[0, 0] == b62[5, 5]
    assert keys[0, 0] == (np.uint64(int(b62[5, 5])) << np.uint64(32)) and (keys[0, 1:] == 0).all()
    # no queries at all
    sc, keys = gpu.search(np.zeros(0, np.int8), np.zeros(0, np.uint16), np.zeros(0, np.uint32), b62, 10, 2, 3,
                          want_scores=True)
    assert sc.shape == (0, 1) and keys.shape == (0, 3)
    # an empty database
    gpu.load_db(np.zeros(0, np.uint16), np.zeros(0, np.int8))
    sc, keys = gpu.search(np.array([5, 6], np.int8), np.array([2], np.uint16), np.array([0], np.uint32), b62, 10, 2, 3,
                          want_scores=True)
    assert sc.shape == (1, 0) and (keys == 0).all()
    # only dummy residues (J/O/U = 23): every score is 0 and the order is index descending
    gpu.load_db(np.array([3, 3, 4], np.uint16), np.full(10, 23, np.int8))
    sc, keys = gpu.search(np.full(7, 23, np.int8), np.array([7], np.uint16), np.array([0], np.uint32), b62, 10, 2, 3,
                          want_scores=True)
    assert (sc == 0).all() and list(keys[0]) == [2, 1, 0]


def test_extreme_penalties_and_zero_gap(gpu, oracle):
    qc, ql, qo, dc, dl, do = _random_case(23, 200, [40, 120], hi=300)
    gpu.load_db(dl, dc)
    for mat, go, ge in [("blosum62", 0, 0), ("blosum62", 0, 1), ("pam30", 127, 127), ("blosum90", 1, 0)]:
        want = oracle.search(qc, qo, dc, do, host.submat(mat), go, ge)
        got, _ = gpu.search(qc, ql, qo[:-1], host.submat(mat), go, ge, 0, want_scores=True)
        assert np.array_equal(got, want), (mat, go, ge)
