"""Host-side logic and the C-ABI surface.  No GPU: nothing here computes a score on a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from swimm_b200 import gpu, host, sharding, synth
from tests.helpers import GOLDEN, ROOT


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "swimm_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sw(?:g|imm)_(?:gpu|plan)_\w+)\s*\(", hdr))
    assert declared == set(gpu.ABI_SYMBOLS), declared ^ set(gpu.ABI_SYMBOLS)
    L = C.CDLL(gpu.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    with pytest.raises(gpu.SwgError):
        gpu.GpuSearch(0)
    n = C.c_int(-1)
    assert gpu.load_library().swg_gpu_device_count(C.byref(n)) != 0 and n.value == 0


def test_cli_refuses_cpu_modes_and_fails_loudly_without_gpu(tmp_path):
    exe = os.path.join(ROOT, "swimm_b200", "swimm")
    prefix = str(tmp_path / "db")
    r = subprocess.run([exe, "-S", "preprocess", "-i", os.path.join(GOLDEN, "edge.db.fasta"), "-o", prefix],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "Database size:\t\t\t33 sequences" in r.stdout
    q = os.path.join(GOLDEN, "edge.q.fasta")
    for mode in ["0", "1", "2"]:
        r = subprocess.run([exe, "-S", "search", "-q", q, "-d", prefix, "-m", mode], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "-S", "search", "-q", q, "-d", prefix, "-m", "3"], capture_output=True, text=True)
        assert r.returncode == 4 and "GPU discovery failed" in r.stdout
    r = subprocess.run([exe, "-S", "bogus"], capture_output=True, text=True)
    assert r.returncode == 1


@pytest.mark.parametrize("name", ["basic", "overflow", "matrices", "edge"])
def test_preprocess_format_matches_reference_layout(tmp_path, name):
    """<prefix>.info/.seq/.desc as the reference writes them (sequences.c:128-205): the same bytes
    tests/golden/make_golden.py fed to the reference binary (synth.write_preprocessed is that layout)."""
    fasta = os.path.join(GOLDEN, name + ".db.fasta")
    mine, ref = str(tmp_path / "mine"), str(tmp_path / "ref")
    host.preprocess_db(fasta, mine)
    from tests.helpers import read_fasta
    synth.write_preprocessed(ref, read_fasta(fasta))
    for ext in ("info", "seq", "desc"):
        assert open(mine + "." + ext, "rb").read() == open(ref + "." + ext, "rb").read(), ext
    db = host.load_db(mine, headers=True)
    lengths, codes, mt = synth.read_preprocessed(mine)
    assert np.array_equal(db.lengths, lengths) and np.array_equal(db.codes, codes) and db.max_title == mt
    assert (np.diff(db.lengths.astype(np.int64)) >= 0).all()
    assert db.titles[0].startswith(">")


def test_fasta_reader_sorts_stably_and_encodes(tmp_path):
    p = tmp_path / "x.fasta"
    p.write_text(">b two\nAC\nDE\n>a one\nJOU\n>c three\nWXYZ\n>d four\nBBB\n")
    s = host.read_fasta(str(p))
    assert list(s.lengths) == [3, 3, 4, 4]
    assert s.titles == [">a one", ">d four", ">b two", ">c three"]          # ties keep the input order
    assert list(s.codes[:3]) == [23, 23, 23] and list(s.codes[3:6]) == [1, 1, 1]
    assert list(s.codes[6:10]) == [synth.ALPHABET.index(c) for c in "ACDE"]


def test_submat_tables_have_the_reference_shape():
    for name in host.MATRICES:
        m = host.submat(name)
        assert m.shape == (24, 32) and (m[23] == 0).all() and (m[:, 23:] == 0).all()
        assert (m[:23, :23] == m[:23, :23].T).all()
    assert host.submat("blosum62")[synth.ALPHABET.index("W"), synth.ALPHABET.index("W")] == 11
    assert host.submat("pam30").min() == -17 and host.submat("pam250").max() == 17


def test_merge_top_keys_matches_sort():
    rng = np.random.default_rng(2)
    parts = np.sort(rng.integers(1, 1 << 40, (5, 12)).astype(np.uint64), axis=1)[:, ::-1].copy()
    want = np.sort(parts.reshape(-1))[::-1][:12]
    assert np.array_equal(host.merge_top_keys(parts, 12), want)
    assert np.array_equal(gpu.merge_top_keys([parts[i:i + 1] for i in range(5)], 12)[0], want)


@pytest.mark.parametrize("n,shards", [(1, 1), (33, 2), (100, 3), (1000, 8)])
def test_shard_index_arithmetic(n, shards):
    seen = np.concatenate([sharding.shard_sequences(n, s, shards) for s in range(shards)])
    assert np.array_equal(np.sort(seen), np.arange(n))
    for s in range(shards):
        g = sharding.shard_sequences(n, s, shards)
        assert np.array_equal(sharding.local_to_global(np.arange(len(g)), s, shards), g)


def test_sharded_oracle_search_merges_to_unsharded(oracle):
    """The multi-GPU data flow on the CPU: shard, score each shard (the oracle stands in for a device),
    build per-shard hit lists with global indices, merge -> the unsharded reference order."""
    rng = np.random.default_rng(4)
    q = synth.make_queries(rng, [40, 90])
    db = synth.make_seqset(rng, synth.lognormal_lengths(rng, 333, 4.3, 0.6, 1, 500))
    _, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    qo = np.zeros(3, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    do = np.zeros(db.n + 1, np.uint64)
    np.cumsum(dl.astype(np.uint64), out=do[1:])
    b62 = host.submat("blosum62")
    want = oracle.search(qc, qo, dc, do, b62, 10, 2)
    top = 15
    for shards in (2, 5):
        parts = []
        for s in range(shards):
            ll, lc, gidx = sharding.extract_shard(dl, dc, s, shards)
            lo = np.zeros(len(ll) + 1, np.uint64)
            np.cumsum(ll.astype(np.uint64), out=lo[1:])
            sc = oracle.search(qc, qo, lc, lo, b62, 10, 2)
            assert np.array_equal(sc, want[:, gidx])
            parts.append(np.stack([sharding.top_keys(sharding.make_keys(sc[i], gidx), top) for i in range(2)]))
        merged = gpu.merge_top_keys(parts, top)
        for i in range(2):
            ts, ti = oracle.top(want[i], top)
            ks, ki = gpu.split_key(merged[i])
            assert np.array_equal(ki, ti.astype(np.int64)) and np.array_equal(ks, ts)
