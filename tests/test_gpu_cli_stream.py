"""The C host's streaming mode and opt-in extensions on the GPU box: query file names on stdin through ONE process with
the database uploaded once, --keep-input-order, --coordinates, -x N on all visible GPUs."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from swimm_b200 import host, synth
from tests.helpers import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "swimm_b200", "swimm")


def _reports(out):
    res = []
    for part in out.split("Query filename:")[1:]:
        hits = [[int(m.group(1)), int(m.group(2))] for m in re.finditer(r"^(-?\d+)\t.*syn\|(\d+)\|", part, re.M)]
        res.append((part.split("\n")[0].strip(), part.count("Query no."), hits))
    return res


def test_stdin_stream_of_three_files_one_process(tmp_path):
    """`-q -`: three query files named on stdin go through one process; the database is loaded once (one header block),
    every file gets its own report with the hit lists of a separate run, an unreadable name in between is reported and
    skipped."""
    prefix = str(tmp_path / "db")
    subprocess.run([EXE, "-S", "preprocess", "-i", os.path.join(GOLDEN, "basic.db.fasta"), "-o", prefix], check=True,
                   stdout=subprocess.DEVNULL)
    meta = json.load(open(os.path.join(GOLDEN, "basic.json")))
    run = meta["runs"][0]
    common = ["-d", prefix, "-m", "3", "-r", str(meta["n"]), "-s", run["matrix"], "-g", str(run["go"]), "-e", str(run["ge"])]
    files = [os.path.join(GOLDEN, "basic.q.fasta"), os.path.join(GOLDEN, "edge.q.fasta"), os.path.join(GOLDEN, "basic.q.fasta")]
    feed = "\n".join([files[0], files[1], str(tmp_path / "missing.fasta"), "", files[2]]) + "\n"
    out = subprocess.run([EXE, "-S", "search", "-q", "-"] + common, input=feed, check=True, capture_output=True, text=True).stdout
    assert out.count("Database size:") == 1 and out.count("An error occurred while opening input sequence file") == 1
    streamed = _reports(out)
    assert [r[0] for r in streamed] == files
    for f, rep in zip(files, streamed):
        single = subprocess.run([EXE, "-S", "search", "-q", f] + common, check=True, capture_output=True, text=True).stdout
        assert _reports(single)[0] == rep
    assert streamed[0][2] == [h for q in run["hits"] for h in q]          # the reference's own printed lists


def test_keep_input_order_and_unsorted_query_file(tmp_path):
    """A multi-query file in DESCENDING length order (the reference mis-pairs lengths and sequences on such a file,
    sequences.c:276 vs :344): by default the reports come in ascending length, with --keep-input-order in file order;
    the hit lists themselves are the same."""
    rng = np.random.default_rng(3)
    db = synth.make_db(5, 400, mu=4.5, sigma=0.6, lo=5, hi=500)
    synth.write_fasta(str(tmp_path / "db.fasta"), db)
    q = synth.make_queries(rng, [200, 90, 41])
    order = [2, 0, 1]                                                          # lengths 200, 41, 90 in the file
    qs = synth.SeqSet(np.concatenate([q.seq(i) for i in order]),
                      np.concatenate([[0], np.cumsum([len(q.seq(i)) for i in order])]).astype(np.int64),
                      [">q%d" % i for i in order])
    synth.write_fasta(str(tmp_path / "q.fasta"), qs)
    prefix = str(tmp_path / "db")
    subprocess.run([EXE, "-S", "preprocess", "-i", str(tmp_path / "db.fasta"), "-o", prefix], check=True, stdout=subprocess.DEVNULL)

    def run(extra):
        out = subprocess.run([EXE, "-S", "search", "-q", str(tmp_path / "q.fasta"), "-d", prefix, "-m", "3", "-r", "5"] + extra,
                             check=True, capture_output=True, text=True).stdout
        blocks = out.split("Query no.")[1:]
        return [(re.search(r"Query description: \t\t(\S+)", b).group(1), int(re.search(r"Query length:\t\t\t(\d+)", b).group(1)),
                 re.findall(r"^(-?\d+)\t(.*)$", b, re.M)) for b in blocks]

    sorted_rep, file_rep = run([]), run(["--keep-input-order"])
    assert [r[1] for r in sorted_rep] == [41, 90, 200]
    assert [r[0] for r in file_rep] == ["q2", "q0", "q1"] and [r[1] for r in file_rep] == [200, 41, 90]
    assert sorted(sorted_rep) == sorted(file_rep)


def test_coordinates_flag_and_all_gpus(tmp_path, oracle):
    """--coordinates adds the query / sequence range of every hit (1-based; oracle.align_ends is the definition), on one
    GPU and with -x 0 (all visible GPUs, shards merged) alike; without the flag the line layout is the reference's."""
    rng = np.random.default_rng(8)
    q = synth.make_queries(rng, [60, 333])
    db = synth.make_db(9, 600, mu=4.8, sigma=0.6, lo=5, hi=700, queries=q, plant_fraction=0.1)
    synth.write_fasta(str(tmp_path / "db.fasta"), db)
    synth.write_fasta(str(tmp_path / "q.fasta"), q)
    prefix = str(tmp_path / "db")
    subprocess.run([EXE, "-S", "preprocess", "-i", str(tmp_path / "db.fasta"), "-o", prefix], check=True, stdout=subprocess.DEVNULL)
    perm, dl, dc = synth.length_sorted(db)
    do = np.zeros(db.n + 1, np.int64)
    np.cumsum(dl.astype(np.int64), out=do[1:])
    rank_of = np.empty(db.n, np.int64)
    rank_of[perm] = np.arange(db.n)
    _, ql, qc = synth.length_sorted(q)
    qo = np.concatenate([[0], np.cumsum(ql.astype(np.int64))])
    b62 = host.submat("blosum62")
    base = [EXE, "-S", "search", "-q", str(tmp_path / "q.fasta"), "-d", prefix, "-m", "3", "-r", "6"]
    plain = subprocess.run(base, check=True, capture_output=True, text=True).stdout
    assert "Query range" not in plain
    outs = [subprocess.run(base + ["--coordinates", "-x", x], check=True, capture_output=True, text=True).stdout for x in ("1", "0")]
    for out in outs:
        blocks = out.split("Query no.")[1:]
        assert len(blocks) == 2
        for qi, b in enumerate(blocks):
            rows = re.findall(r"^(\d+)\t(\d+)-(\d+)\t(\d+)-(\d+)\t.*syn\|(\d+)\|", b, re.M)
            assert len(rows) == 6
            for sc, qs_, qe, ds, de, fasta_idx in rows:
                pos = rank_of[int(fasta_idx)]
                s, c = oracle.align_ends(qc[qo[qi]:qo[qi + 1]], dc[do[pos]:do[pos + 1]], b62, 10, 2)
                assert int(sc) == s and [int(qs_) - 1, int(qe) - 1, int(ds) - 1, int(de) - 1] == list(c)
    strip = lambda o: re.sub(r"\t\d+-\d+\t\d+-\d+", "", o.split("Search date:")[0]).replace("\tQuery range\tSequence range", "")
    assert strip(outs[0]) == plain.split("Search date:")[0] == strip(outs[1])
