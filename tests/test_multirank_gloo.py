"""The N>1 path of bench.py / a multi-process search on CPU: world_size-2 gloo, one shard per rank, the
per-rank hit lists all-gathered and merged on every rank.  The oracle stands in for the device (checker only)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from swimm_b200 import gpu, host, sharding, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.helpers import load_oracle
    oracle = load_oracle()
    rng = np.random.default_rng(21)                       # every rank builds the same whole database
    q = synth.make_queries(rng, [30, 77, 144])
    db = synth.make_seqset(rng, synth.lognormal_lengths(rng, 257, 4.4, 0.6, 1, 600))
    synth.plant(rng, db, q, fraction=0.1, frag_range=(10, 70), rate=0.1)
    _, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    qo = np.zeros(q.n + 1, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    b62 = host.submat("blosum62")
    top = 10
    ll, lc, gidx = sharding.extract_shard(dl, dc, rank, world)
    lo = np.zeros(len(ll) + 1, np.uint64)
    np.cumsum(ll.astype(np.uint64), out=lo[1:])
    sc = oracle.search(qc, qo, lc, lo, b62, 10, 2)
    mine = np.stack([sharding.top_keys(sharding.make_keys(sc[i], gidx), top) for i in range(q.n)])
    gathered = [torch.zeros((q.n, top), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine.view(np.int64)))
    merged = gpu.merge_top_keys([g.numpy().view(np.uint64) for g in gathered], top)
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    if rank == 0:
        do = np.zeros(db.n + 1, np.uint64)
        np.cumsum(dl.astype(np.uint64), out=do[1:])
        want = oracle.search(qc, qo, dc, do, b62, 10, 2)
        ok = True
        for i in range(q.n):
            ts, ti = oracle.top(want[i], top)
            ks, ki = gpu.split_key(merged[i])
            ok &= bool(np.array_equal(ki, ti.astype(np.int64)) and np.array_equal(ks, ts))
        open(os.path.join(out_dir, "ok"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_merge(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"
