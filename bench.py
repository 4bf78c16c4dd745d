#!/usr/bin/env python3
"""bench.py -- GCUPS of the SWIMM `search` hot path on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2] [--scale S]

A step = one pass of the hot path over one batch: all queries of the workload against the resident
database shard (query profiles, 16-bit kernels, 32-bit recomputation, top-r).  Headline workload: BASELINE.json
configs[1] ("cfg2": 20 queries of 144..5478 residues vs a synthetic Swiss-Prot-sized database, ~570k
sequences / ~205M residues per GPU).  N>1: one process per GPU (torchrun); ALL ranks build the same seeded
database of N x 570k sequences and keep their own shard (tile round-robin), so per-GPU work is fixed (weak
scaling) and the merged hit list can be verified; no data-path collective, the per-rank hit lists are gathered
and merged once per step.

value  = cells of all ranks / max-over-ranks device time (CUDA events on the library's stream),
         inputs resident in HBM.
e2e    = the same through swg_gpu_search() with host buffers: query H2D + kernels + top-r + hit-list D2H
         (+ the cross-rank gather/merge) inside the timed region, wall clock, max over ranks.
strong = measured in the same run after the headline: fixed databases (cfg3 = 6M sequences / 1.3G residues, cfg1 =
         100k sequences) split over the N ranks; GCUPS of the whole job.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  Libraries (NCCL's version banner, for one) print to file descriptor 1
# behind Python's back, so fd 1 is pointed at stderr for the whole run and the line is written to the saved fd.
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


from swimm_b200 import host, synth  # noqa: E402

METRIC = "GCUPS (whole box) Swiss-Prot-scale synthetic search"
GO, GE, MATRIX, TOP = 10, 2, "blosum62", 10
# one sampling rule for both CPU legs (cpu_baseline of our arm, and --impl reference): every k-th sequence of the
# shard, k chosen so that the sample is about this many cells per step (all queries)
CPU_SAMPLE_CELLS = 6e11
NCU_JSON = "profiles/r02_ncu_full_cfg2_kernels.json"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def sm_clock_mhz(index):
    """Current SM clock of one GPU (NVML); None when NVML is not usable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        return float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    except Exception:
        return None


def timed_steps(s, step, steps, warm, min_warm_seconds=0.4):
    """`warm` untimed steps -- and as many more as it takes to keep the GPU busy for min_warm_seconds, so that a
    millisecond-sized workload is not timed while the SM clock is still ramping up from idle -- then `steps` timed steps.
    Returns the per-step device seconds (CUDA events of the library)."""
    t0 = time.time()
    n = 0
    while n < warm or time.time() - t0 < min_warm_seconds:
        step()
        n += 1
        if n >= 5000:
            break
    per = []
    for _ in range(steps):
        step()
        per.append(s.stats()["device_seconds"])
    return per, n


def physical_device_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        parts = [p for p in vis.split(",") if p.strip() != ""]
        if local_rank < len(parts) and parts[local_rank].strip().isdigit():
            return int(parts[local_rank])
    return local_rank


# ---------------------------------------------------------------------------------------------------
class Workload:
    """Queries + the WHOLE length-sorted database of a BASELINE.json configuration (every rank builds the same one)."""

    def __init__(self, name, scale=1.0, world=1, weak=False):
        t0 = time.time()
        self.name = name
        mult = world if weak else 1
        if name == "cfg2":      # Swiss-Prot-sized, 20 queries; weak scaling: `world` times as many sequences
            self.q = synth.make_queries(np.random.default_rng(7), synth.QUERY_LENGTHS)
            n = max(64, int(570_000 * scale)) // 16 * 16 * mult
            self.dl, self.dc = synth.sorted_db(1000, n, mu=5.675, queries=self.q)
        elif name == "cfg1":    # one 144-residue query vs 100k sequences
            self.q = synth.make_queries(np.random.default_rng(42), [144])
            n = max(64, int(100_000 * scale)) // 16 * 16 * mult
            self.dl, self.dc = synth.sorted_db(42, n, queries=self.q)
        elif name == "cfg3":    # Environmental-NR-sized
            self.q = synth.make_queries(np.random.default_rng(7), synth.QUERY_LENGTHS)
            n = max(64, int(6_000_000 * scale)) // 16 * 16 * mult
            self.dl, self.dc = synth.sorted_db(3000, n, mu=5.2, sigma=0.6, queries=self.q)
        elif name == "cfg4":    # long-sequence stress: every sequence longer than 3000 residues, up to the format's 65535
            rng = np.random.default_rng(44)
            self.q = synth.make_queries(rng, [144, 1000, 3100, 5478])
            n = max(64, int(8000 * scale)) // 16 * 16
            lens = np.concatenate([rng.integers(3001, 20000, n - 8), [30000, 40000, 50000, 60000, 65535, 65535, 3001, 3002]])
            db = synth.make_seqset(rng, lens)
            synth.plant(rng, db, self.q, fraction=0.02, frag_range=(50, 3000), rate=0.1)
            _, self.dl, self.dc = synth.length_sorted(db)
        elif name == "cfg5":    # multi-query batch for the matrix / penalty sweep, on the cfg1 database
            self.q = synth.make_queries(np.random.default_rng(55), [61, 144, 222, 375, 567, 850, 1321, 2005])
            n = max(64, int(100_000 * scale)) // 16 * 16
            self.dl, self.dc = synth.sorted_db(42, n, queries=self.q, plant_fraction=0.005)
        else:
            raise SystemExit("unknown workload %s" % name)
        _, self.ql, self.qc = synth.length_sorted(self.q)
        self.qo = np.zeros(self.q.n + 1, np.uint32)
        np.cumsum(self.ql.astype(np.uint32), out=self.qo[1:])
        self.q_res = int(self.ql.astype(np.int64).sum())
        self.n_total = len(self.dl)
        self.residues = len(self.dc)
        self.cells = self.q_res * self.residues
        self.gen_s = time.time() - t0

    def shard(self, rank, world):
        return synth.shard_of(self.dl, self.dc, rank, world)

    def describe(self, world=1):
        return "%s: %d queries (%d..%d residues, %d total) x %d sequences / %d residues%s" % (
            self.name, self.q.n, int(self.ql.min()), int(self.ql.max()), self.q_res, self.n_total // world,
            self.residues // world, " per GPU" if world > 1 else "")


def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "swimm")
    return p if os.path.exists(p) and os.access(p, os.X_OK) else None


def run_reference_once(q, dl, dc, threads, tmp):
    """Time the unmodified reference (oracle/_ref/swimm -S search -m 0 -v 32) on preprocessed files written
    in its own format; returns (search_seconds_printed, cells)."""
    prefix = os.path.join(tmp, "db")
    if not os.path.exists(prefix + ".seq"):
        n = len(dl)
        with open(prefix + ".info", "w") as f:
            f.write("%d %d %d" % (n, len(dc), 40))
        with open(prefix + ".seq", "wb") as f:
            f.write(np.asarray(dl, "<u2").tobytes())
            f.write(np.asarray(dc, np.int8).tobytes())
        with open(prefix + ".desc", "w") as f:
            for i in range(n):
                f.write(">syn|%09d| s\n" % i)
    qf = os.path.join(tmp, "q.fasta")
    if not os.path.exists(qf):
        synth.write_fasta(qf, q)
    out = subprocess.run([reference_binary(), "-S", "search", "-q", qf, "-d", prefix, "-m", "0", "-v", "32", "-c",
                          str(threads), "-r", str(TOP)], capture_output=True, text=True, check=True).stdout
    secs = None
    for line in out.splitlines():
        if line.startswith("Search time:"):
            secs = float(line.split()[2])
    cells = int(sum(len(q.seq(i)) for i in range(q.n))) * int(len(dc))
    return secs, cells


def cpu_sample(q, dl, dc):
    """The bounded sample of the workload both CPU legs time: every k-th sequence of the (length-sorted) shard, which
    keeps the length mix, and all queries; k from CPU_SAMPLE_CELLS."""
    n = len(dl)
    want_res = CPU_SAMPLE_CELLS / float(sum(len(q.seq(i)) for i in range(q.n)))
    stride = max(1, int(round(max(1, len(dc)) / want_res)))
    off = np.zeros(n + 1, np.int64)
    np.cumsum(dl.astype(np.int64), out=off[1:])
    keep = np.arange(0, n, stride)
    sl = dl[keep]
    idx = np.repeat(off[keep] - np.concatenate([[0], np.cumsum(sl.astype(np.int64))[:-1]]), sl.astype(np.int64)) + \
        np.arange(int(sl.astype(np.int64).sum()))
    return sl, dc[idx], "every %d-th sequence of the shard (%d sequences, %d residues) x all %d queries" % (
        stride, len(sl), len(idx), q.n)


def full_size_crosscheck():
    """The reference CLI timed on the FULL cfg2 workload on a GPU box (tools/full_parity.py, committed transcript)."""
    for name in ("r02_cfg2_full_size_parity_vs_reference_cli.txt", "r01_cfg2_full_size_parity_vs_reference_cli.txt"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            for line in open(p):
                if line.startswith("reference:"):
                    return {"source": "profiles/" + name, "line": line.strip()}
    return None


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipebench", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the strong-scaling legs and the other BASELINE configs")
    ap.add_argument("--strong-scale", type=float, default=1.0, help="size of the fixed cfg3 database of the strong-scaling leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        return bench_reference(args, cores)

    import torch
    import torch.distributed as dist
    from swimm_b200 import gpu
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    b62 = host.submat(MATRIX)
    wl = Workload(args.workload, args.scale, world, weak=True)
    dl, dc, gidx = wl.shard(rank, world)
    log("[rank %d] %s; this shard: %d sequences, %d residues (%.1f s to generate)"
        % (rank, wl.describe(world), len(dl), len(dc), wl.gen_s))
    q, ql, qc, qo = wl.q, wl.ql, wl.qc, wl.qo
    cells_local = wl.q_res * int(len(dc))
    cells_all = float(wl.cells)

    s = gpu.GpuSearch(local_rank)
    t0 = time.time()
    s.load_db_shard(dl, dc, rank, world, wl.n_total)
    db_load_s = time.time() - t0
    st0 = s.stats()

    # pinned host buffers for the end-to-end leg
    pin_q = torch.from_numpy(qc.copy()).pin_memory()
    pin_keys = torch.zeros((q.n, TOP), dtype=torch.int64).pin_memory()
    keys_np = pin_keys.numpy().view(np.uint64)
    gather = [torch.zeros((q.n, TOP), dtype=torch.int64, device="cuda") for _ in range(world)] if world > 1 else None

    def merge(keys):
        if world == 1:
            return keys
        # the one exchange step of the path: r keys per query per GPU, merged on every rank
        dist.all_gather(gather, torch.from_numpy(keys.view(np.int64)).cuda(non_blocking=True))
        return gpu.merge_top_keys([g.cpu().numpy().view(np.uint64) for g in gather], TOP)

    def step_resident():
        s.run(TOP)
        s.sync()

    def step_e2e():
        s.set_queries(pin_q.numpy(), ql, qo[:-1], b62, GO, GE)
        s.run(TOP)
        s.fetch(False, True, keys_out=keys_np)
        return merge(keys_np)

    # ---- value: inputs resident ----
    s.set_queries(qc, ql, qo[:-1], b62, GO, GE)
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(physical_device_index(local_rank))
    barrier()
    s.sync()
    sampler.start()
    dev_s, search_s, topr_s, launches = 0.0, 0.0, 0.0, 0
    q_secs = np.zeros(q.n)
    w0 = time.time()
    for _ in range(args.steps):
        step_resident()
        st = s.stats()
        dev_s += st["device_seconds"]
        search_s += st["search_seconds"]
        topr_s += st["topr_seconds"]
        launches += st["launches"]
        q_secs += s.query_seconds()
    barrier()
    wall_s = time.time() - w0
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    dev_max = max_over_ranks(dev_s)
    wall_max = max_over_ranks(wall_s)
    value = cells_all * args.steps / dev_max / 1e9

    # ---- e2e: host buffers in, hit lists out ----
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    e0 = time.time()
    for _ in range(args.steps):
        merged = step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.time() - e0)
    merged = np.array(merged, copy=True)          # (one rank: the pinned buffer itself, reused below)
    st = s.stats()
    rescored = st["rescored"]             # filled by fetch (the resident loop does not download the counters)
    e2e = {"value": cells_all * args.steps / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(st["h2d_bytes"]),
           "d2h_bytes_per_step": int(st["d2h_bytes"]), "ms_per_step": e2e_s / args.steps * 1e3}

    # ---- e2e, streaming: back-to-back batches through submit/poll (batch k+1 is uploaded and queued while k computes) ----
    nb = max(4, args.steps + 1)
    barrier()
    s0_ = time.time()
    tickets = [s.submit(pin_q.numpy(), ql, qo[:-1], b62, GO, GE, TOP)]
    for b in range(1, nb):
        tickets.append(s.submit(pin_q.numpy(), ql, qo[:-1], b62, GO, GE, TOP))
        k_, _ = s.poll(tickets[b - 1], True, keys_out=keys_np)
        merged_stream = merge(keys_np)
    s.poll(tickets[-1], True, keys_out=keys_np)
    merged_stream = merge(keys_np)
    barrier()
    stream_s = max_over_ranks(time.time() - s0_)
    e2e_stream = {"value": cells_all * nb / stream_s / 1e9, "unit": "GCUPS", "batches": nb,
                  "ms_per_batch": stream_s / nb * 1e3, "same_hits_as_blocking_call": bool(np.array_equal(merged_stream, merged))}

    # ---- merged hit list: verified on rank 0 against a single-GPU search of a sampled subset of the ONE database ----
    merged_ok, merged_note = None, None
    if rank == 0:
        ks, ki = gpu.split_key(merged)
        hits = np.unique(ki.ravel())
        stride = max(1, wl.n_total // 20000)
        sub = np.unique(np.concatenate([hits, np.arange(0, wl.n_total, stride)]))
        off = np.zeros(wl.n_total + 1, np.int64)
        np.cumsum(wl.dl.astype(np.int64), out=off[1:])
        sl = wl.dl[sub]
        idx = np.repeat(off[sub] - np.concatenate([[0], np.cumsum(sl.astype(np.int64))[:-1]]), sl.astype(np.int64)) + \
            np.arange(int(sl.astype(np.int64).sum()))
        chk = gpu.GpuSearch(local_rank)
        chk.load_db(sl, wl.dc[idx])
        sc, _ = chk.search(qc, ql, qo[:-1], b62, GO, GE, 0, want_scores=True)
        chk.close()
        # (1) every merged hit has exactly the score the single-GPU search gives that sequence; (2) no sampled sequence
        # that is missing from a query's list beats the list's last hit (key order: score, then index)
        pos = {int(g): i for i, g in enumerate(sub)}
        ok = True
        for qi in range(q.n):
            cols = np.array([pos[int(g)] for g in ki[qi]])
            ok = ok and bool(np.array_equal(sc[qi, cols], ks[qi]))
            keys_sub = (sc[qi].astype(np.uint64) << np.uint64(32)) | sub.astype(np.uint64)
            last = np.uint64(merged[qi, -1])
            better = sub[keys_sub > last]
            ok = ok and bool(np.isin(better, ki[qi]).all())
            ok = ok and bool((np.diff(merged[qi].astype(np.int64)) < 0).all())
        merged_ok = ok
        digest = hashlib.sha256(np.ascontiguousarray(merged).tobytes()).hexdigest()[:16]
        merged_note = ("merged top-%d of all %d ranks (ONE database of %d sequences, tile round-robin shards) checked on rank 0 "
                       "against an unsharded single-GPU search of %d sampled sequences + all hit sequences: hit scores equal, "
                       "no sampled non-hit beats a list's last key, keys strictly descending; sha256[:16] of the merged keys %s"
                       % (TOP, world, wl.n_total, len(sub) - len(hits), digest))

    # ---- roofline: the integer/SIMD pipe for each kernel's own instruction mix, measured live ----
    roof, peaks = None, None
    if rank == 0 and not args.no_pipebench:
        roof = roofline(s, gpu, q_secs, ql, len(dc), args.steps, st, search_s, cells_local, dev_s)
        peaks = (roof.pop("_peaks"), roof.pop("_pb"))

    # ---- strong scaling + the other BASELINE configurations, measured in the same run ----
    strong, configs = None, None
    if not args.no_extra:
        strong = {}
        for name, sc_, steps_, warm_ in (("cfg3", args.strong_scale, 2, 1), ("cfg1", 1.0, 20, 5)):
            w2 = Workload(name, sc_, world, weak=False)
            l2, c2, _ = w2.shard(rank, world)
            s.load_db_shard(l2, c2, rank, world, w2.n_total)
            s.set_queries(w2.qc, w2.ql, w2.qo[:-1], b62, GO, GE)
            barrier()
            per_step, warmed = timed_steps(s, step_resident, steps_, warm_)
            mhz = sm_clock_mhz(physical_device_index(local_rank))
            barrier()
            d2 = max_over_ranks(sum(per_step))
            log("[rank %d] strong %s per-step ms: %s" % (rank, name, " ".join("%.3f" % (x * 1e3) for x in per_step)))
            strong[name] = {"workload": w2.describe(), "n_gpus": world, "steps": steps_, "warmup": warmed,
                            "ms_per_step": d2 / steps_ * 1e3, "gcups": w2.cells * steps_ / d2 / 1e9,
                            "ms_per_step_min_this_rank": min(per_step) * 1e3, "sm_mhz_after": mhz,
                            "note": "fixed total database split over the ranks by tiles; efficiency = gcups / (N x the N=1 line's gcups)"}
            log("[rank %d] strong %s: %.3f ms/step" % (rank, name, d2 / steps_ * 1e3))
            del w2, l2, c2
        if world == 1:
            configs = other_configs(s, gpu, peaks)
    s.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu_base = None
    if not args.no_cpu_baseline and reference_binary():
        sl, sc, what = cpu_sample(q, dl, dc)
        with tempfile.TemporaryDirectory() as tmp:
            run_reference_once(q, sl, sc, cores, tmp)                 # warm-up (page cache, OpenMP pool)
            t_cpu, runs = [], 0
            t_begin = time.time()
            while runs < 3 or (time.time() - t_begin < 12 and runs < 12):
                secs, cells = run_reference_once(q, sl, sc, cores, tmp)
                t_cpu.append(secs)
                runs += 1
        med = float(np.median(t_cpu))
        cpu_base = {"value": cells / med / 1e9, "unit": "GCUPS", "cores": cores, "kind": "reference",
                    "sample": what + "; swimm -S search -m 0 -v 32 -c %d, median printed Search time of %d runs %.3f s "
                                     "(the same sample --impl reference times)" % (cores, runs, med),
                    "full_size_crosscheck": full_size_crosscheck()}

    line = {"metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_max / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "s16x2 (exact int32 result)", "data": "synthetic",
            "config": {"workload": wl.describe(world) + ", BLOSUM62, gap 10/2, top %d" % TOP,
                       "sharding": "ONE seeded database of %d sequences; tile round-robin, one shard per GPU, hit lists "
                                   "merged per step" % wl.n_total,
                       "l2": "tiled database per GPU (%d MB) exceeds the 126 MB L2; every search launch streams it once"
                             % (st0["db_bytes"] >> 20)},
            "e2e": e2e, "e2e_stream": e2e_stream, "gpu_launches": int(launches), "clocks": sampler.summary(),
            "merged_ok": merged_ok, "merged_note": merged_note,
            "roofline": roof, "cpu_baseline": cpu_base, "strong": strong,
            "detail": {"search_ms_per_step": search_s / args.steps * 1e3, "topr_ms_per_step": topr_s / args.steps * 1e3,
                       "wall_ms_per_step": wall_max / args.steps * 1e3, "db_load_seconds": db_load_s,
                       "rescored_lanes": int(rescored), "db_bytes": int(st0["db_bytes"]),
                       "per_query_gcups": {str(int(ql[i])): round(float(ql[i]) * len(dc) * args.steps / q_secs[i] / 1e9, 1)
                                           for i in range(q.n) if q_secs[i] > 0},
                       "configs": configs}}
    if world > 1:
        dist.destroy_process_group()
    emit(line)
    return 0


KERNELS = [
    # kind, name, pipebench probes (default penalties as immediates / gap-extend as an immediate / generic), instr per
    # 2 cells, ALU-pipe instr per 2 cells
    (0, "wavefront_kernel<Lane16,G,K> (one query x two database sequences per register)",
     ("mix_v2_immediate_penalties", "mix_v2_extend_immediate_open_register", "mix_v2_4p5_alu_2_viadd"), 6.5, 4.5,
     "6.5 integer instr per 2 cells: 4.5 on the ALU pipe (PRMT, VIMNMX3.RELU, VIADDMNMX x2, VIMNMX3/2) + 2 VIADD.16x2"),
    (1, "wavefront_q2_kernel<G,K> (two queries x one database sequence per register)",
     ("mix_q2_3p5alu_2viadd_immediate", "mix_q2_extend_immediate_open_register", "mix_c_hef_best_3p5alu_2viadd"), 5.5, 3.5,
     "5.5 integer instr per 2 cells: 3.5 on the ALU pipe (VIMNMX3.RELU, VIADDMNMX x2, VIMNMX3/2) + 2 VIADD.16x2"),
]


def kernel_peaks(pb, go, ge):
    """Per kernel: peak GCUPS of its own measured dependency-free mix -- the instantiation family the penalties select:
    SWIMM's defaults (10/2) as immediates, gap-extend 1 or 2 as an immediate (query-pair kernel), or penalties in
    registers -- and the strict ALU-pipe bound (64 lanes/clk/SM)."""
    out = {}
    for kind, name, probes, ipc2, alu2, mix_text in KERNELS:
        probe = probes[0] if (go + ge, ge) == (12, 2) else probes[1] if (ge in (1, 2) and probes[1]) else probes[2]
        mix = pb["probes"][probe]
        strict = pb["sms"] * mix["sm_mhz"] * 1e6 * 64.0 * 2.0 / alu2 / 1e9
        out[kind] = {"name": name, "peak": mix["ginstr_per_s"] * 2.0 / ipc2, "peak_strict": strict,
                     "mix": mix_text + "; probe %s: measured %.1f thread-instr/clk/SM at %.0f MHz"
                            % (probe, mix["thread_instr_per_clk_per_sm"], mix["sm_mhz"])}
    return out


def roofline(s, gpu, q_secs, ql, residues, steps, st, search_s, cells_local, dev_s):
    pb = s.pipebench()
    peaks = kernel_peaks(pb, GO, GE)
    kinds = s.query_kernels()
    kernels = {}
    for kind, pk in peaks.items():
        sel = kinds == kind
        secs = float(q_secs[sel].sum())
        if secs <= 0:
            continue
        cells_k = float(ql[sel].astype(np.int64).sum()) * residues * steps
        kernels[kind] = {"kernel": pk["name"], "queries": int(sel.sum()), "share_of_search_time": secs / float(q_secs.sum()),
                         "achieved": cells_k / secs / 1e9, "peak": pk["peak"], "frac": cells_k / secs / 1e9 / pk["peak"],
                         "peak_strict_alu_pipe": pk["peak_strict"], "frac_strict_alu_pipe": cells_k / secs / 1e9 / pk["peak_strict"],
                         "mix": pk["mix"]}
    dom = max(kernels.values(), key=lambda k: k["share_of_search_time"])
    others = [k for k in kernels.values() if k is not dom]
    per_gpu = cells_local * steps / dev_s / 1e9
    have_peaks = os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if have_peaks else 6650.0
    # algorithmic HBM bytes: the tiled database once per search launch + the pass lines of the query-pair kernel
    algo_gbs = st["stream_bytes"] * steps / search_s / 1e9
    # DRAM traffic of the dominant kernel's launches from the committed ncu --set full captures (read + written)
    traffic, traffic_note = None, None
    try:
        cap = json.load(open(os.path.join(ROOT, NCU_JSON)))
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        key = "q2_middle_pass" if dom is kernels.get(1) else "seqpair_q144"
        traffic = sum(float(cap[key][k].split()[0]) * unit[cap[key][k].split()[1]]
                      for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic_note = ("ncu dram bytes read+written by one full-size launch of the dominant kernel (%s, %s); algorithmic bytes of "
                        "such a launch: the tiled database (%d) + -- query-pair kernel, launches that continue a query -- the "
                        "pass lines, 8 B per database column read and 8 B written" % (NCU_JSON, key, st["db_bytes"]))
    except Exception:
        pass
    return {"bound": "int_alu", "achieved": dom["achieved"], "peak": dom["peak"], "unit": "GCUPS",
            "frac": dom["frac"], "peak_strict_alu_pipe": dom["peak_strict_alu_pipe"],
            "frac_strict_alu_pipe": dom["frac_strict_alu_pipe"], "traffic": traffic, "traffic_note": traffic_note,
            "kernel": dom["kernel"], "mix": dom["mix"], "share_of_search_time": dom["share_of_search_time"],
            "other_kernels": others, "whole_step_gcups_per_gpu": per_gpu,
            "hbm": {"bound": "hbm", "achieved": algo_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": algo_gbs / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if have_peaks else "fallback"},
            "pipebench": {k: round(v["thread_instr_per_clk_per_sm"], 2) for k, v in pb["probes"].items()},
            "_peaks": peaks, "_pb": pb}


def other_configs(s, gpu, peaks):
    """BASELINE.json configs[0], [3], [4] on this GPU, after the headline: GCUPS of the whole batch, which kernel took most
    of the time and the fraction of that kernel's roofline."""
    out = {}
    runs = [("cfg1", "blosum62", 10, 2, 20, 5), ("cfg4", "blosum62", 10, 2, 3, 2),
            ("cfg5", "blosum45", 8, 2, 10, 3), ("cfg5", "pam250", 12, 1, 10, 3)]
    loaded = None
    for name, matrix, go, ge, steps, warm in runs:
        if loaded is None or loaded.name != name:
            loaded = Workload(name)
            s.load_db(loaded.dl, loaded.dc)
        w = loaded
        s.set_queries(w.qc, w.ql, w.qo[:-1], host.submat(matrix), go, ge)

        def one():
            s.run(TOP)
            s.sync()
        timed_steps(s, one, 0, warm)
        dev, srch, qs = 0.0, 0.0, np.zeros(w.q.n)
        for _ in range(steps):
            one()
            dev += s.stats()["device_seconds"]
            srch += s.stats()["search_seconds"]
            qs += s.query_seconds()
        kinds = s.query_kernels()
        share1 = float(qs[kinds == 1].sum()) / max(float(qs.sum()), 1e-12)
        kind = 1 if share1 >= 0.5 else 0
        gc = w.cells * steps / dev / 1e9           # the whole step: search kernels + top-r
        kgc = w.cells * steps / srch / 1e9        # the search kernels alone (what the roofline fraction is about)
        rec = {"workload": w.describe() + ", %s, gap %d/%d" % (matrix, go, ge), "gcups": gc, "ms_per_step": dev / steps * 1e3,
               "search_kernels_gcups": kgc, "search_ms_per_step": srch / steps * 1e3,
               "sm_mhz_after": sm_clock_mhz(0),
               "dominant_kernel": KERNELS[kind][1], "share_of_search_time": share1 if kind else 1.0 - share1}
        if peaks:
            pk = kernel_peaks(peaks[1], go, ge)[kind]
            rec.update({"peak": pk["peak"], "frac": kgc / pk["peak"], "frac_strict_alu_pipe": kgc / pk["peak_strict"],
                        "frac_whole_step": gc / pk["peak"], "mix": pk["mix"]})
        out["%s %s %d/%d" % (name, matrix, go, ge)] = rec
        log("[config] %s %s %d/%d: %.0f GCUPS" % (name, matrix, go, ge, gc))
    return out


def bench_reference(args, cores):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    wl = Workload(args.workload, args.scale, 1)
    q, ql, qc, qo, dl, dc = wl.q, wl.ql, wl.qc, wl.qo, wl.dl, wl.dc
    kind = "reference" if reference_binary() else "port"
    sl, sc, what = cpu_sample(q, dl, dc)
    times = []
    with tempfile.TemporaryDirectory() as tmp:
        if kind == "reference":
            for i in range(args.warmup + args.steps):
                secs, cells = run_reference_once(q, sl, sc, cores, tmp)
                if i >= args.warmup:
                    times.append(secs)
        else:
            from tests.helpers import load_oracle
            oracle = load_oracle()
            so = np.zeros(len(sl) + 1, np.uint64)
            np.cumsum(sl.astype(np.uint64), out=so[1:])
            cells = wl.q_res * len(sc)
            for i in range(args.warmup + args.steps):
                t0 = time.time()
                oracle.search(qc, qo, sc, so, host.submat(MATRIX), GO, GE, threads=cores)
                if i >= args.warmup:
                    times.append(time.time() - t0)
    total = sum(times)
    value = cells * len(times) / total / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s8/s16/s32 AVX2",
            "data": "synthetic",
            "config": {"workload": wl.describe() + ", BLOSUM62, gap 10/2, top %d" % TOP},
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": cores, "kind": kind,
                             "sample": what + ("; swimm -S search -m 0 -v 32 -c %d (printed Search time); GCUPS is a rate: "
                                               "the full-size run agrees, see full_size_crosscheck" % cores
                                               if kind == "reference" else "; scalar oracle port, OpenMP"),
                             "full_size_crosscheck": full_size_crosscheck()},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
