"""The data flow of the query-pair kernel's two-column step (csrc/wavefront_q2.cuh), restated in plain Python and
checked against the oracle: thread t works two columns behind thread t-1, a step walks columns a and b with the
(H, F) of the row above handed down from the previous step's outputs, DS is formed one column ahead from the two
letters of the step's word, a thread restarts its rows in head step t of a segment, segments follow each other without
draining the pipeline, the last letter of a segment enters with the first step of the next one, and a sequence's score
is read G steps into the following segment.  This pins the schedule (skew, restart step, where the score is
complete), not the CUDA code: that is what the -m gpu parity tests do."""
import numpy as np
import pytest

from swimm_b200 import host, synth

PAD = -1


def model_scores(q, seqs, S, goe, ge, G, K):
    m = len(q)
    assert m <= G * K
    sc = lambda row, letter: int(S[int(q[row]), int(letter)]) if row < m and letter != PAD else 0
    minseg = 2 * G + 16
    DS = [[0] * K for _ in range(G)]
    E = [[0] * K for _ in range(G)]
    best, bsave = [0] * G, [0] * G
    pk = [(PAD, PAD)] * G
    outs = [(0, 0, 0, 0)] * G
    result, pending, last_letter = [], False, PAD
    for seg in list(seqs) + [None]:                       # None: the flush segment that pushes the last sequence through
        L = [] if seg is None else [int(c) for c in seg]
        L += [PAD] * (-len(L) % 4)
        L += [PAD] * max(0, minseg - len(L))
        for s in range(len(L) // 2):
            feed = (last_letter if s == 0 else L[2 * s - 1], L[2 * s])
            old_pk, old_outs = list(pk), list(outs)
            for t in range(G):
                pkn = feed if t == 0 else old_pk[t - 1]
                r_ha, r_fa, r_hb, r_fb = (0, 0, 0, 0) if t == 0 else old_outs[t - 1]
                pk[t] = pkn
                res = []
                for hp, f, letter in ((r_ha, r_fa, pkn[0]), (r_hb, r_fb, pkn[1])):    # column a, then column b
                    for x in range(K):
                        ds = DS[t][x]
                        h = max(ds, E[t][x], f, 0)
                        E[t][x] = max(E[t][x] - ge, h - goe)
                        f = max(f - ge, h - goe)
                        best[t] = max(best[t], ds)
                        DS[t][x] = hp + sc(t * K + x, letter)
                        hp = h
                    res += [hp, f]
                outs[t] = tuple(res)
                if s == t:                                  # head step t: the next column starts a new sequence
                    for x in range(K):
                        DS[t][x] = sc(t * K + x, pkn[1])
                        E[t][x] = 0
                    bsave[t], best[t] = best[t], 0
            if s == G - 1 and pending:
                result.append(max(bsave))
        last_letter, pending = L[-1], seg is not None
    return result


@pytest.mark.parametrize("G,K,qlen", [(4, 3, 12), (4, 3, 7), (8, 2, 16), (2, 6, 11)])
def test_two_column_schedule_matches_oracle(oracle, G, K, qlen):
    rng = np.random.default_rng(100 * G + K)
    S = host.submat("blosum62")
    q = synth.make_queries(rng, [qlen])
    # lengths around the padding rules: shorter than a minimal segment, multiples of four (no padding between two
    # sequences), one and two columns, long
    lens = np.array([1, 2, 3, 4, 8, 2 * G + 16, 2 * G + 17, 40, 40, 41, 64, 97, 150])
    db = synth.make_seqset(rng, lens)
    synth.plant(rng, db, q, fraction=0.6, frag_range=(3, qlen), rate=0.1)
    for t in (4, 8, 10):                                    # a match that ends in a sequence's last column
        tail = q.seq(0)[-min(qlen, int(lens[t])):]
        db.residues[db.offsets[t + 1] - len(tail):db.offsets[t + 1]] = tail
    dc = synth.encode(db.residues)
    qc = synth.encode(q.residues)
    do = db.offsets.astype(np.uint64)
    qo = q.offsets.astype(np.uint32)
    want = oracle.search(qc, qo, dc, do, S, 10, 2)[0]
    seqs = [dc[int(do[i]):int(do[i + 1])] for i in range(db.n)]
    got = model_scores(qc, seqs, S, 12, 2, G, K)
    assert got == [int(v) for v in want]
