"""The host-side planner (swimm_b200/csrc/swg_plan.cu) through swg_plan_describe: pure host code, no GPU needed.
Invariants of every schedule: each query is searched exactly once, a streamed query's launches cover its rows in
order without gaps, a lane serves one query at a time, launch heights are legal kernel shapes."""
import re

import numpy as np
import pytest

from swimm_b200 import gpu, synth

CFG2 = dict(n_sequences=570_000, n_residues=205_136_075, longest_sequence=7779)
LAUNCH = re.compile(r"G=(\d+) K=(\d+) \| lane 0: q (-?\d+) \((\d+) rows\) from row (\d+)( first)?( last)? \| "
                    r"lane 1: q (-?\d+) \((\d+) rows\) from row (\d+)( first)?( last)?")


def parse(text):
    """-> (single: {q: (G, K, passes)}, groups: [[(G, K, [(q, rows, row0, first, last)] * 2)]])"""
    single, groups = {}, []
    for line in text.splitlines():
        m = re.match(r"\[swg\] query (\d+) \((\d+) rows\): sequence-pair kernel G=(\d+) K=(\d+) passes=(\d+)", line)
        if m:
            single[int(m.group(1))] = (int(m.group(3)), int(m.group(4)), int(m.group(5)))
            continue
        if "query-pair kernel," in line:
            groups.append([])
            continue
        m = LAUNCH.search(line)
        assert m, line
        g = m.groups()
        lanes = [(int(g[2]), int(g[3]), int(g[4]), g[5] is not None, g[6] is not None),
                 (int(g[7]), int(g[8]), int(g[9]), g[10] is not None, g[11] is not None)]
        groups[-1].append((int(g[0]), int(g[1]), lanes))
    return single, groups


def check_schedule(q_lengths, text):
    single, groups = parse(text)
    seen = set(single)
    for q, (G, K, passes) in single.items():
        assert G in (4, 8, 16, 32) and 1 <= K <= 32 and passes * G * K >= q_lengths[q]
        assert passes == 1 or G == 32
    for launches in groups:
        progress = {}                      # query -> rows done
        current = [None, None]             # the query each lane is working on
        for G, K, lanes in launches:
            assert G in (8, 16, 32) and K % 2 == 0 and 8 <= K <= 32
            assert len(launches) == 1 or G == 32
            for l, (q, rows, row0, first, last) in enumerate(lanes):
                if q < 0:
                    assert current[l] is None
                    continue
                assert rows == q_lengths[q]
                if first:
                    assert q not in progress and q not in seen and current[l] is None and row0 == 0
                    current[l] = q
                    progress[q] = 0
                assert current[l] == q and row0 == progress[q]          # in order, no gap, no overlap
                progress[q] += G * K
                assert last == (progress[q] >= max(rows, 1))
                if last:
                    current[l] = None
                    seen.add(q)
        assert current == [None, None]
    assert seen == set(range(len(q_lengths)))
    return single, groups


def test_cfg2_schedule():
    ql = synth.QUERY_LENGTHS
    text = gpu.plan_describe(ql, **CFG2)
    single, groups = check_schedule(ql, text)
    streamed = sum(len({q for _, _, lanes in g for q, *_ in lanes if q >= 0}) for g in groups)
    assert streamed >= 16 and len(single) <= 4
    lane_rows = sum(32 * K for g in groups for G, K, _ in g if G == 32 and len(g) > 1)
    query_rows = sum(ql[q] for g in groups if len(g) > 1 for q in {q for _, _, lanes in g for q, *_ in lanes if q >= 0})
    assert query_rows / (2 * lane_rows) > 0.96           # the lanes are almost free of padding rows


def test_pairing_off_and_single_query():
    ql = synth.QUERY_LENGTHS
    single, groups = check_schedule(ql, gpu.plan_describe(ql, query_pairing=0, **CFG2))
    assert not groups and len(single) == len(ql)
    single, groups = check_schedule([2005], gpu.plan_describe([2005], **CFG2))
    assert not groups and single[0][2] > 1               # one query: the sequence-pair kernel, several passes
    single, groups = check_schedule([], gpu.plan_describe([], **CFG2))
    assert not groups and not single


def test_forced_pairing_covers_everything_but_an_odd_one():
    ql = [1, 7, 144, 189, 1000, 1024, 1025, 3000, 65535]
    single, groups = check_schedule(ql, gpu.plan_describe(ql, query_pairing=2, **CFG2))
    assert len(single) <= 1


def test_chain_bound_shard_prefers_the_pair_kernel_and_wider_groups():
    """cfg4-like shard (few, very long sequences): the longest sequence's serial chain decides."""
    ql = [144, 1000, 3100, 5478]
    single, groups = check_schedule(ql, gpu.plan_describe(ql, n_sequences=8000, n_residues=91_932_856,
                                                          longest_sequence=65535))
    assert len(groups) == 1 and set(single) == set()         # the short query fills the idle lane of the stream
    # a small shard with one long sequence: a shape with more threads per sequence than the saturated optimum (G=8)
    # only when the long-tile path cannot cap the chain -- here it can, so the throughput shape stays
    single, _ = check_schedule([144], gpu.plan_describe([144], n_sequences=100_000, n_residues=35_092_341,
                                                        longest_sequence=7367))
    assert single[0][0] in (8, 16, 32)


@pytest.mark.parametrize("seed", range(8))
def test_random_batches_keep_the_invariants(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 40))
    ql = [int(x) for x in rng.integers(0, 6000, n)]
    shard = dict(n_sequences=int(rng.integers(16, 2_000_000)), n_residues=int(rng.integers(10_000, 400_000_000)),
                 longest_sequence=int(rng.integers(50, 40_000)))
    for pairing in (0, 1, 2):
        check_schedule(ql, gpu.plan_describe(ql, query_pairing=pairing, **shard))


def test_column_chunks_cover_every_possible_alignment():
    """The exactness of the column chunks rests on two host-side facts: the span bound B = m + m*Smax/ge + 1, and
    that every window of at most B columns of a long tile lies entirely inside one chunk.  Checked on random tiles,
    queries, matrices' maxima, penalties and chunk options; chunk starts are multiples of 8, chunks stay inside
    their tile, together they cover it, and they come longest first."""
    rng = np.random.default_rng(3)
    for trial in range(300):
        m = int(rng.integers(1, 1025))
        smax = int(rng.integers(1, 18))
        ge = int(rng.integers(1, 4))
        option = int(rng.choice([0, 0, 0, 8, 1000, 2048, 5000, 20000]))
        tiles = (rng.integers(1, 8192, int(rng.integers(1, 12))) * 8).astype(np.uint32)
        chunks, B = gpu.plan_column_chunks(m, smax, ge, tiles, option)
        assert B == m + m * smax // ge + 1
        if len(chunks) == 0:
            continue
        assert (np.diff(chunks[:, 2].astype(np.int64)) <= 0).all()
        for t, cols in enumerate(tiles):
            mine = chunks[chunks[:, 0] == t]
            mine = mine[np.argsort(mine[:, 1])]
            assert len(mine) >= 1 and (mine[:, 1] % 8 == 0).all() and (mine[:, 1] + mine[:, 2] <= cols).all()
            assert mine[0, 1] == 0 and mine[-1, 1] + mine[-1, 2] == cols
            # every window [a, a + w) with w <= B inside the tile is inside some chunk: it suffices to check, for every
            # chunk start s_k, that the window starting just before the NEXT chunk's start still fits into chunk k
            starts, ends = mine[:, 1].astype(np.int64), (mine[:, 1] + mine[:, 2]).astype(np.int64)
            for k in range(len(mine) - 1):
                a = starts[k + 1] - 1                     # the last start that chunk k + 1 does not cover
                assert min(a + B, cols) <= ends[k], (m, smax, ge, option, cols, mine)
    # option 1 switches chunking off; a matrix or penalty without a bound does too
    assert len(gpu.plan_column_chunks(144, 11, 2, [40000], 1)[0]) == 0
    none, unbounded = gpu.plan_column_chunks(144, 11, 0, [40000], 0)
    assert len(none) == 0 and unbounded == 0
    chunks, B = gpu.plan_column_chunks(144, 11, 2, [40000], 0)
    assert B == 937 and chunks[0, 2] == 2048 and len(chunks) > 30
