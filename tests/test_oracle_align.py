"""The oracle's coordinate extension (oracle/sw_oracle.c swo_align_ends; no reference counterpart, parity unpinned):
internal consistency on CPU -- same score as swo_score, the reported region alone reproduces the score, the region is
tight, and the documented tie rule holds on a repeat."""
import numpy as np

from swimm_b200 import host, synth


def test_align_ends_is_consistent_with_the_score(oracle):
    b62 = host.submat("blosum62")
    rng = np.random.default_rng(0)
    for t in range(300):
        m, n = int(rng.integers(1, 50)), int(rng.integers(1, 80))
        q = synth.encode(synth.random_residues(rng, m))
        d = synth.encode(synth.random_residues(rng, n))
        if t % 3 == 0:
            k = min(m, n) // 2 + 1
            d[:k] = q[:k]
        go, ge = [(10, 2), (5, 1), (12, 3)][t % 3]
        s, c = oracle.align_ends(q, d, b62, go, ge)
        assert s == oracle.score(q, d, b62, go, ge)
        if s == 0:
            assert (c == -1).all()
            continue
        assert 0 <= c[0] <= c[1] < m and 0 <= c[2] <= c[3] < n
        assert oracle.score(q[c[0]:c[1] + 1], d[c[2]:c[3] + 1], b62, go, ge) == s
        assert oracle.score(q[c[0]:c[1]], d[c[2]:c[3] + 1], b62, go, ge) < s
        assert oracle.score(q[c[0]:c[1] + 1], d[c[2]:c[3]], b62, go, ge) < s


def test_align_ends_tie_rule(oracle):
    """WWW against WWWWWW: the score 33 is reached at database positions 2..5; the smallest one wins."""
    b62 = host.submat("blosum62")
    w = synth.encode(np.full(6, ord("W"), np.uint8))
    s, c = oracle.align_ends(w[:3], w, b62, 10, 2)
    assert s == 33 and list(c) == [0, 2, 0, 2]
