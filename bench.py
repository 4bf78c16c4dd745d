#!/usr/bin/env python3
"""bench.py -- GCUPS of the SWIMM `search` hot path on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2] [--scale S]

A step = one pass of the hot path over one batch: all queries of the workload against the resident
database shard (query profiles, 16-bit kernels, 32-bit recomputation, top-r).  N=1 workload: BASELINE.json
configs[1] ("cfg2": 20 queries of 144..5478 residues vs a synthetic Swiss-Prot-sized database, ~570k
sequences / ~205M residues).  N>1: one process per GPU (torchrun), every rank holds one shard of an
N-times larger database (weak scaling), no data-path collective; the per-rank hit lists are gathered
and merged once per step.

value  = cells of all ranks / max-over-ranks device time (CUDA events on the library's stream),
         inputs resident in HBM.
e2e    = the same through swg_gpu_search() with host buffers: query H2D + kernels + top-r + hit-list D2H
         (+ the cross-rank gather/merge) inside the timed region, wall clock, max over ranks.
"""
import argparse
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


def physical_device_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        parts = [p for p in vis.split(",") if p.strip() != ""]
        if local_rank < len(parts) and parts[local_rank].strip().isdigit():
            return int(parts[local_rank])
    return local_rank


# ---------------------------------------------------------------------------------------------------
def make_shard(workload, scale, rank, world):
    """This rank's shard of the (world x larger) database + the queries, in the library's input form."""
    t0 = time.time()
    if workload == "cfg2":
        q = synth.make_queries(np.random.default_rng(7), synth.QUERY_LENGTHS)
        n_local = max(64, int(570_000 * scale)) // 16 * 16
        db = synth.make_db(1000 + rank, n_local, mu=5.675, queries=q)
    elif workload == "cfg1":
        q = synth.make_queries(np.random.default_rng(42), [144])
        n_local = max(64, int(100_000 * scale)) // 16 * 16
        db = synth.make_db(42 + rank, n_local, queries=q)
    elif workload == "cfg3":      # Environmental-NR-sized TOTAL, split over the ranks
        q = synth.make_queries(np.random.default_rng(7), synth.QUERY_LENGTHS)
        n_local = max(64, int(6_000_000 * scale / world)) // 16 * 16
        db = synth.make_db(3000 + rank, n_local, mu=5.2, sigma=0.6, queries=q)
    else:
        raise SystemExit("unknown workload %s" % workload)
    _, dl, dc = synth.length_sorted(db)
    _, ql, qc = synth.length_sorted(q)
    qo = np.zeros(q.n + 1, np.uint32)
    np.cumsum(ql.astype(np.uint32), out=qo[1:])
    log("[rank %d] shard: %d sequences, %d residues, %d queries (%.1f s to generate)"
        % (rank, len(dl), len(dc), q.n, time.time() - t0))
    return q, ql, qc, qo, dl, dc


def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "swimm")
    return p if os.path.exists(p) and os.access(p, os.X_OK) else None


def run_reference_once(q, dl, dc, threads, tmp, queries_subset=None):
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
    idx = list(range(q.n)) if queries_subset is None else queries_subset
    qf = os.path.join(tmp, "q_%s.fasta" % "_".join(map(str, idx)))
    if not os.path.exists(qf):
        sub = synth.SeqSet(np.concatenate([q.seq(i) for i in idx]),
                           np.concatenate([[0], np.cumsum([len(q.seq(i)) for i in idx])]).astype(np.int64),
                           [q.title(i) for i in idx])
        synth.write_fasta(qf, sub)
    out = subprocess.run([reference_binary(), "-S", "search", "-q", qf, "-d", prefix, "-m", "0", "-v", "32", "-c",
                          str(threads), "-r", str(TOP)], capture_output=True, text=True, check=True).stdout
    secs = None
    for line in out.splitlines():
        if line.startswith("Search time:"):
            secs = float(line.split()[2])
    cells = int(sum(len(q.seq(i)) for i in idx)) * int(len(dc))
    return secs, cells


def cpu_baseline_sample(q, dl, dc, target_cells=1.5e12):
    """A bounded sample of the workload for the CPU legs: the first sequences of the shard (every 8th tile keeps
    the length mix) and all queries, sized to ~10-30 s of reference time."""
    n = len(dl)
    want_res = target_cells / float(sum(len(q.seq(i)) for i in range(q.n)))
    frac = min(1.0, want_res / max(1, len(dc)))
    stride = max(1, int(round(1.0 / frac)))
    off = np.zeros(n + 1, np.int64)
    np.cumsum(dl.astype(np.int64), out=off[1:])
    keep = np.arange(0, n, stride)
    sl = dl[keep]
    idx = np.repeat(off[keep] - np.concatenate([[0], np.cumsum(sl.astype(np.int64))[:-1]]), sl.astype(np.int64)) + \
        np.arange(int(sl.astype(np.int64).sum()))
    return sl, dc[idx], "every %d-th sequence of the shard (%d sequences, %d residues) x all %d queries" % (
        stride, len(sl), len(idx), q.n)


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

    q, ql, qc, qo, dl, dc = make_shard(args.workload, args.scale, rank, world)
    n_local = len(dl)
    n_total = n_local * world
    b62 = host.submat(MATRIX)
    q_res = int(ql.astype(np.int64).sum())
    cells_local = q_res * int(len(dc))

    s = gpu.GpuSearch(local_rank)
    t0 = time.time()
    s.load_db_shard(dl, dc, rank, world, n_total)
    db_load_s = time.time() - t0
    st0 = s.stats()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s.sync()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # pinned host buffers for the end-to-end leg
    pin_q = torch.from_numpy(qc.copy()).pin_memory()
    pin_keys = torch.zeros((q.n, TOP), dtype=torch.int64).pin_memory()
    keys_np = pin_keys.numpy().view(np.uint64)
    gather = [torch.zeros((q.n, TOP), dtype=torch.int64, device="cuda") for _ in range(world)] if world > 1 else None

    def step_resident():
        s.run(TOP)
        s.sync()

    def step_e2e():
        s.set_queries(pin_q.numpy(), ql, qo[:-1], b62, GO, GE)
        s.run(TOP)
        s.fetch(False, True, keys_out=keys_np)
        if world > 1:      # the one exchange step of the path: r keys per query per GPU, merged on every rank
            dist.all_gather(gather, pin_keys.cuda(non_blocking=True))
            merged = gpu.merge_top_keys([g.cpu().numpy().view(np.uint64) for g in gather], TOP)
        else:
            merged = keys_np
        return merged

    # ---- value: inputs resident ----
    s.set_queries(qc, ql, qo[:-1], b62, GO, GE)
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(physical_device_index(local_rank))
    barrier()
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
    cells_all = sum_over_ranks(cells_local)
    value = cells_all * args.steps / dev_max / 1e9
    rescored = s.stats()["rescored"]

    # ---- e2e: host buffers in, hit lists out ----
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    e0 = time.time()
    for _ in range(args.steps):
        merged = step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.time() - e0)
    st = s.stats()
    e2e = {"value": cells_all * args.steps / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(st["h2d_bytes"]),
           "d2h_bytes_per_step": int(st["d2h_bytes"]), "ms_per_step": e2e_s / args.steps * 1e3}

    if rank != 0:
        s.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline: the integer/SIMD pipe for each kernel's own instruction mix, measured live ----
    roof = None
    if not args.no_pipebench:
        pb = s.pipebench()
        fast = (GO + GE, GE) == (12, 2)         # default penalties run the kernels whose penalties are immediates
        kinds = s.query_kernels()
        kernels = {}
        # sequence-pair kernel: 6.5 integer instructions per 2 cells; query-pair kernel: 5.5 (no score-pack PRMT).
        # peak cells/s = dependency-free issue rate of that mix * 2 / instructions per cell pair
        for kind, name, probe, ipc2, mix_text in [
                (0, "wavefront_kernel<Lane16,G,K> (one query x two database sequences per register)",
                 "mix_v2_immediate_penalties" if fast else "mix_v2_4p5_alu_2_viadd", 6.5,
                 "6.5 integer instr per 2 cells: 4.5 on the ALU pipe (PRMT, VIMNMX3.RELU, VIADDMNMX x2, VIMNMX3/2) + 2 VIADD.16x2"),
                (1, "wavefront_q2_kernel<G,K> (two queries x one database sequence per register)",
                 "mix_q2_3p5alu_2viadd_immediate" if fast else "mix_c_hef_best_3p5alu_2viadd", 5.5,
                 "5.5 integer instr per 2 cells: 3.5 on the ALU pipe (VIMNMX3.RELU, VIADDMNMX x2, VIMNMX3/2) + 2 VIADD.16x2")]:
            sel = kinds == kind
            secs = float(q_secs[sel].sum())
            if secs <= 0:
                continue
            mix = pb["probes"][probe]
            cells_k = float(ql[sel].astype(np.int64).sum()) * len(dc) * args.steps
            peak = mix["ginstr_per_s"] * 2.0 / ipc2
            kernels[kind] = {"kernel": name, "queries": int(sel.sum()), "share_of_search_time": secs / float(q_secs.sum()),
                             "achieved": cells_k / secs / 1e9, "peak": peak, "frac": cells_k / secs / 1e9 / peak,
                             "mix": mix_text + "; measured %.1f thread-instr/clk/SM at %.0f MHz"
                                    % (mix["thread_instr_per_clk_per_sm"], mix["sm_mhz"])}
        dom = max(kernels.values(), key=lambda k: k["share_of_search_time"])
        others = [k for k in kernels.values() if k is not dom]
        per_gpu = cells_local * args.steps / dev_s / 1e9
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        # algorithmic HBM bytes: the tiled database once per search launch + the pass lines of the query-pair kernel
        algo_gbs = st["stream_bytes"] * args.steps / search_s / 1e9
        # DRAM traffic of the dominant kernel's launches from the committed ncu --set full captures (read + written)
        traffic, traffic_note = None, None
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_cfg2_kernels.json")))
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            key = "q2_middle_pass" if dom is kernels.get(1) else "seqpair_q144"
            traffic = sum(float(cap[key][k].split()[0]) * unit[cap[key][k].split()[1]]
                          for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            algo_launch = st["stream_bytes"] / max(1, st["pair_launches"]) if dom is kernels.get(1) else st["db_bytes"]
            traffic_note = ("ncu dram bytes read+written by one full-size launch (profiles/r01_ncu_full_cfg2_kernels.json, %s); "
                            "algorithmic bytes: the tiled database (%d) once per launch, plus -- query-pair kernel, launches that "
                            "continue a query -- the pass lines, 8 B per database column read and 8 B written; average over "
                            "this run's launches %.0f" % (key, st["db_bytes"], algo_launch))
        except Exception:
            pass
        roof = {"bound": "int_alu", "achieved": dom["achieved"], "peak": dom["peak"], "unit": "GCUPS",
                "frac": dom["frac"], "traffic": traffic, "traffic_note": traffic_note,
                "kernel": dom["kernel"], "mix": dom["mix"], "share_of_search_time": dom["share_of_search_time"],
                "other_kernels": others, "whole_step_gcups_per_gpu": per_gpu,
                "hbm": {"bound": "hbm", "achieved": algo_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": algo_gbs / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json"))
                        else "fallback"},
                "pipebench": {k: round(v["thread_instr_per_clk_per_sm"], 2) for k, v in pb["probes"].items()}}
    s.close()

    cpu_base = None
    if not args.no_cpu_baseline and reference_binary():
        sl, sc, what = cpu_baseline_sample(q, dl, dc)
        with tempfile.TemporaryDirectory() as tmp:
            run_reference_once(q, sl, sc, cores, tmp)                 # warm-up (page cache, OpenMP pool)
            secs, cells = run_reference_once(q, sl, sc, cores, tmp)
        cpu_base = {"value": cells / secs / 1e9, "unit": "GCUPS", "cores": cores, "kind": "reference",
                    "sample": what + "; swimm -S search -m 0 -v 32 -c %d, printed Search time %.3f s" % (cores, secs)}

    line = {"metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_max / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "s16x2 (exact int32 result)", "data": "synthetic",
            "config": {"workload": "%s: %d queries (%d..%d residues, %d total) x %d sequences / %d residues per GPU, "
                                   "BLOSUM62, gap 10/2, top %d" % (args.workload, q.n, int(ql.min()), int(ql.max()), q_res,
                                                                    n_local, len(dc), TOP),
                       "sharding": "tile round-robin, one shard per GPU, hit lists merged per step",
                       "l2": "tiled database per GPU (%d MB) exceeds the 126 MB L2; every search launch streams it once"
                             % (st["db_bytes"] >> 20)},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(),
            "roofline": roof, "cpu_baseline": cpu_base,
            "detail": {"search_ms_per_step": search_s / args.steps * 1e3, "topr_ms_per_step": topr_s / args.steps * 1e3,
                       "wall_ms_per_step": wall_max / args.steps * 1e3, "db_load_seconds": db_load_s,
                       "rescored_lanes": int(rescored), "db_bytes": int(st0["db_bytes"]),
                       "per_query_gcups": {str(int(ql[i])): round(float(ql[i]) * len(dc) * args.steps / q_secs[i] / 1e9, 1)
                                           for i in range(q.n) if q_secs[i] > 0}}}
    if world > 1:
        dist.destroy_process_group()
    emit(line)
    return 0


def bench_reference(args, cores):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    q, ql, qc, qo, dl, dc = make_shard(args.workload, args.scale, 0, 1)
    q_res = int(ql.astype(np.int64).sum())
    kind = "reference" if reference_binary() else "port"
    sl, sc, what = cpu_baseline_sample(q, dl, dc, target_cells=6e11)
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
            cells = q_res * len(sc)
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
            "config": {"workload": "%s: %d queries (%d..%d residues, %d total), BLOSUM62, gap 10/2, top %d"
                                   % (args.workload, q.n, int(ql.min()), int(ql.max()), q_res, TOP)},
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": cores, "kind": kind,
                             "sample": what + ("; swimm -S search -m 0 -v 32 -c %d (printed Search time)" % cores
                                               if kind == "reference" else "; scalar oracle port, OpenMP")},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
