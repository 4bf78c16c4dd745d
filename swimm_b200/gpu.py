"""ctypes binding of libswimm_cuda.so -- the C ABI declared in include/swimm_gpu.h.

This is host-side plumbing only: every score is computed by the CUDA library.  There is no CPU
path here; if the library is missing or no B200 is visible the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (SWIMM_B200_LIB: an experimental build of the same library, e.g. another CTA size: tools only)
LIB_PATH = os.environ.get("SWIMM_B200_LIB") or os.path.join(_HERE, "libswimm_cuda.so")

# every symbol include/swimm_gpu.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "swg_gpu_device_count", "swg_gpu_create", "swg_gpu_destroy", "swg_gpu_last_error", "swg_gpu_load_db",
    "swg_gpu_load_db_shard", "swg_gpu_load_db_interleaved", "swg_gpu_db_local_sequences", "swg_gpu_db_local_residues", "swg_gpu_search",
    "swg_gpu_set_queries", "swg_gpu_run", "swg_gpu_fetch", "swg_gpu_sync", "swg_gpu_get_stats", "swg_gpu_get_query_seconds",
    "swg_gpu_get_query_kernels", "swg_plan_describe", "swg_gpu_pipebench",
    "swg_gpu_set_option", "swg_gpu_debug_read", "swimm_gpu_search_avx2_compat", "swg_gpu_submit", "swg_gpu_poll",
    "swg_gpu_load_db_offsets", "swg_gpu_align_ends", "swg_plan_column_chunks",
]


class SwgError(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [("device_seconds", C.c_double), ("search_seconds", C.c_double), ("topr_seconds", C.c_double),
                ("cells", C.c_uint64), ("padded_cells", C.c_uint64), ("launches", C.c_uint64),
                ("rescored", C.c_uint64), ("db_bytes", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("pair_launches", C.c_uint64), ("stream_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load_library() -> C.CDLL:
    """dlopen the product library; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SwgError("libswimm_cuda.so is not built: run `make -C swimm_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.swg_gpu_device_count.argtypes = [C.POINTER(i32)]
    L.swg_gpu_create.argtypes = [i32, C.POINTER(vp)]
    L.swg_gpu_destroy.argtypes = [vp]
    L.swg_gpu_destroy.restype = None
    L.swg_gpu_last_error.argtypes = [vp]
    L.swg_gpu_last_error.restype = C.c_char_p
    L.swg_gpu_load_db.argtypes = [vp, vp, vp, u64, u64, i32, i32]
    L.swg_gpu_load_db_offsets.argtypes = [vp, vp, vp, vp, u64, u64, i32, i32]
    L.swg_gpu_load_db_shard.argtypes = [vp, vp, vp, u64, u64, i32, i32, u64]
    L.swg_gpu_load_db_interleaved.argtypes = [vp, vp, vp, u64, vp, i32, u64, i32, i32]
    L.swg_gpu_db_local_sequences.argtypes = [vp]
    L.swg_gpu_db_local_sequences.restype = u64
    L.swg_gpu_db_local_residues.argtypes = [vp]
    L.swg_gpu_db_local_residues.restype = u64
    L.swg_gpu_search.argtypes = [vp, vp, vp, vp, u64, vp, i32, i32, u64, vp, vp, C.POINTER(C.c_double)]
    L.swg_gpu_set_queries.argtypes = [vp, vp, vp, vp, u64, vp, i32, i32]
    L.swg_gpu_run.argtypes = [vp, u64, i32]
    L.swg_gpu_fetch.argtypes = [vp, vp, vp]
    L.swg_gpu_sync.argtypes = [vp]
    L.swg_gpu_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.swg_gpu_get_query_seconds.argtypes = [vp, vp, u64]
    L.swg_gpu_get_query_kernels.argtypes = [vp, vp, u64]
    L.swg_plan_describe.argtypes = [vp, u64, u64, u64, C.c_uint32, i32, C.c_char_p, u64]
    L.swg_gpu_pipebench.argtypes = [vp, i32, vp, vp, vp, C.POINTER(i32), C.POINTER(i32)]
    L.swg_gpu_debug_read.argtypes = [vp, C.c_char_p, vp, u64, C.POINTER(u64)]
    L.swg_gpu_set_option.argtypes = [vp, C.c_char_p, C.c_long]
    L.swg_gpu_align_ends.argtypes = [vp, vp]
    L.swg_plan_column_chunks.argtypes = [C.c_uint32, i32, i32, C.c_long, vp, C.c_uint32, vp, u64, C.POINTER(u64), C.POINTER(u64)]
    L.swg_gpu_submit.argtypes = [vp, vp, vp, vp, u64, vp, i32, i32, u64, C.POINTER(i32)]
    L.swg_gpu_poll.argtypes = [vp, i32, i32, vp, C.POINTER(C.c_double), C.POINTER(i32)]
    L.swimm_gpu_search_avx2_compat.argtypes = [vp, vp, C.c_ulong, vp, vp, vp, vp, C.c_ulong, vp, vp, i32, i32, i32, i32,
                                               vp, C.POINTER(C.c_double)]
    for f in ABI_SYMBOLS:
        if f not in ("swg_gpu_destroy", "swg_gpu_last_error", "swg_gpu_db_local_sequences", "swg_gpu_db_local_residues"):
            getattr(L, f).restype = i32
    _lib = L
    return L


def device_count() -> int:
    L = load_library()
    n = C.c_int(0)
    L.swg_gpu_device_count(C.byref(n))
    return n.value


def split_key(keys: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """SWG_KEY -> (score, database index)."""
    keys = np.asarray(keys, dtype=np.uint64)
    return (keys >> np.uint64(32)).astype(np.int64), (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)


def merge_top_keys(parts: list[np.ndarray], top: int) -> np.ndarray:
    """Top-r merge of per-shard hit lists ([q][r] each): the r largest keys per query, descending."""
    allk = np.concatenate(parts, axis=1)
    allk = -np.sort(-allk.astype(np.int64), axis=1)      # keys are < 2^63 (scores are non-negative int32)
    return allk[:, :top].astype(np.uint64)


def plan_describe(q_lengths, n_sequences: int, n_residues: int, longest_sequence: int, query_pairing: int = 1) -> str:
    """The schedule the library would use for a batch (host-only: works without a GPU)."""
    L = load_library()
    ql = np.ascontiguousarray(q_lengths, dtype=np.uint16)
    buf = C.create_string_buffer(1 << 20)
    st = L.swg_plan_describe(ql.ctypes.data, len(ql), n_sequences, n_residues, longest_sequence, query_pairing, buf, len(buf))
    if st != 0:
        raise SwgError("swg_plan_describe -> %d: %s" % (st, L.swg_gpu_last_error(None).decode()))
    return buf.value.decode()


def plan_column_chunks(m: int, smax: int, ge: int, tile_cols, option: int = 0):
    """(chunks [n][3] = tile, first column, columns; span bound B) the library would use (host-only)."""
    L = load_library()
    tc = np.ascontiguousarray(tile_cols, dtype=np.uint32)
    cap = int(tc.astype(np.int64).sum() // 8 + len(tc) + 8)
    out = np.zeros((cap, 3), dtype=np.uint32)
    n, b = C.c_uint64(0), C.c_uint64(0)
    st = L.swg_plan_column_chunks(m, smax, ge, option, tc.ctypes.data, len(tc), out.ctypes.data, cap, C.byref(n), C.byref(b))
    if st != 0:
        raise SwgError("swg_plan_column_chunks -> %d" % st)
    return out[:n.value], b.value


class GpuSearch:
    """One context = one GPU.  Mirrors the order of calls in the reference driver (swimm.c:38-76)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        self.ctx = C.c_void_p()
        st = self.L.swg_gpu_create(device, C.byref(self.ctx))
        if st != 0:
            raise SwgError("swg_gpu_create(%d) -> %d: %s" % (device, st, self.L.swg_gpu_last_error(None).decode()))
        self.device = device
        self.n_total = 0
        self.q_count = 0
        self._keep = []

    def close(self):
        if self.ctx:
            self.L.swg_gpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st, what):
        if st != 0:
            raise SwgError("%s -> %d: %s" % (what, st, self.L.swg_gpu_last_error(self.ctx).decode()))

    def set_option(self, name: str, value: int):
        self._check(self.L.swg_gpu_set_option(self.ctx, name.encode(), int(value)), "set_option(%s)" % name)

    def load_db(self, lengths, codes, shard: int = 0, num_shards: int = 1, offsets=None):
        lengths = np.ascontiguousarray(lengths, dtype=np.uint16)
        codes = np.ascontiguousarray(codes, dtype=np.int8)
        self.n_total = len(lengths)
        if offsets is not None:      # prefix sums of the lengths: the shard visits only its own tiles
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
            self._check(self.L.swg_gpu_load_db_offsets(self.ctx, lengths.ctypes.data, offsets.ctypes.data, codes.ctypes.data,
                                                       len(lengths), len(codes), shard, num_shards), "load_db_offsets")
            return
        self._check(self.L.swg_gpu_load_db(self.ctx, lengths.ctypes.data, codes.ctypes.data, len(lengths), len(codes),
                                           shard, num_shards), "load_db")

    def load_db_shard(self, local_lengths, local_codes, shard: int, num_shards: int, n_total: int):
        local_lengths = np.ascontiguousarray(local_lengths, dtype=np.uint16)
        local_codes = np.ascontiguousarray(local_codes, dtype=np.int8)
        self.n_total = n_total
        self._check(self.L.swg_gpu_load_db_shard(self.ctx, local_lengths.ctypes.data, local_codes.ctypes.data,
                                                 len(local_lengths), len(local_codes), shard, num_shards, n_total),
                    "load_db_shard")

    def load_db_interleaved(self, vect_db, vect_lengths, vect_disp, vector_length, n_sequences, shard=0, num_shards=1):
        vect_db = np.ascontiguousarray(vect_db, dtype=np.int8)
        vect_lengths = np.ascontiguousarray(vect_lengths, dtype=np.uint16)
        vect_disp = np.ascontiguousarray(vect_disp, dtype=np.uint64)
        self.n_total = n_sequences
        self._check(self.L.swg_gpu_load_db_interleaved(self.ctx, vect_db.ctypes.data, vect_lengths.ctypes.data,
                                                       len(vect_lengths), vect_disp.ctypes.data, vector_length,
                                                       n_sequences, shard, num_shards), "load_db_interleaved")

    @property
    def local_sequences(self) -> int:
        return self.L.swg_gpu_db_local_sequences(self.ctx)

    @property
    def local_residues(self) -> int:
        return self.L.swg_gpu_db_local_residues(self.ctx)

    def set_queries(self, q_codes, q_lengths, q_disp, submat, go: int, ge: int):
        q_codes = np.ascontiguousarray(q_codes, dtype=np.int8)
        q_lengths = np.ascontiguousarray(q_lengths, dtype=np.uint16)
        q_disp = np.ascontiguousarray(q_disp, dtype=np.uint32)
        submat = np.ascontiguousarray(submat, dtype=np.int8)
        assert submat.size == 768
        self.q_count = len(q_lengths)
        self._keep = [q_codes, q_lengths, q_disp, submat]
        self._check(self.L.swg_gpu_set_queries(self.ctx, q_codes.ctypes.data, q_lengths.ctypes.data, q_disp.ctypes.data,
                                               self.q_count, submat.ctypes.data, go, ge), "set_queries")

    def run(self, top: int, keep_scores: bool = False):
        self.top = int(top)
        self._check(self.L.swg_gpu_run(self.ctx, self.top, int(keep_scores)), "run")

    def sync(self):
        self._check(self.L.swg_gpu_sync(self.ctx), "sync")

    def fetch(self, want_scores: bool = False, want_keys: bool = True, keys_out: np.ndarray | None = None):
        scores = np.zeros((self.q_count, self.n_total), dtype=np.int32) if want_scores else None
        keys = None
        if want_keys:
            keys = keys_out if keys_out is not None else np.zeros((self.q_count, self.top), dtype=np.uint64)
        self._check(self.L.swg_gpu_fetch(self.ctx, scores.ctypes.data if scores is not None else None,
                                         keys.ctypes.data if keys is not None else None), "fetch")
        return scores, keys

    def search(self, q_codes, q_lengths, q_disp, submat, go, ge, top, want_scores=False):
        """The whole call with host buffers, as the reference driver would make it."""
        self.set_queries(q_codes, q_lengths, q_disp, submat, go, ge)
        self.run(top, want_scores)
        return self.fetch(want_scores, top > 0)

    def align_ends(self) -> np.ndarray:
        """Opt-in coordinate pass after a search with top > 0: [q][top][4] = q_start, q_end, d_start, d_end (0-based)."""
        out = np.full((self.q_count, self.top, 4), -1, dtype=np.int32)
        self._check(self.L.swg_gpu_align_ends(self.ctx, out.ctypes.data), "align_ends")
        return out

    def submit(self, q_codes, q_lengths, q_disp, submat, go, ge, top) -> int:
        """Streaming: enqueue a whole batch (upload, kernels, hit-list download) and return a ticket at once; two
        batches may be in flight."""
        q_codes = np.ascontiguousarray(q_codes, dtype=np.int8)
        q_lengths = np.ascontiguousarray(q_lengths, dtype=np.uint16)
        q_disp = np.ascontiguousarray(q_disp, dtype=np.uint32)
        submat = np.ascontiguousarray(submat, dtype=np.int8)
        t = C.c_int(-1)
        self._check(self.L.swg_gpu_submit(self.ctx, q_codes.ctypes.data, q_lengths.ctypes.data, q_disp.ctypes.data,
                                          len(q_lengths), submat.ctypes.data, go, ge, int(top), C.byref(t)), "submit")
        self._tickets = getattr(self, "_tickets", {})
        self._tickets[t.value] = (len(q_lengths), int(top))
        return t.value

    def poll(self, ticket: int, wait: bool = True, keys_out: np.ndarray | None = None):
        """-> (hit keys [q][top] or None when not finished, device seconds)."""
        nq, top = self._tickets[ticket]
        keys = keys_out if keys_out is not None else np.zeros((nq, top), dtype=np.uint64)
        done, secs = C.c_int(0), C.c_double(0)
        self._check(self.L.swg_gpu_poll(self.ctx, ticket, int(wait), keys.ctypes.data, C.byref(secs), C.byref(done)), "poll")
        if not done.value:
            return None, 0.0
        del self._tickets[ticket]
        return keys, secs.value

    def stats(self) -> dict:
        s = Stats()
        self._check(self.L.swg_gpu_get_stats(self.ctx, C.byref(s)), "get_stats")
        return s.as_dict()

    def query_seconds(self) -> np.ndarray:
        out = np.zeros(self.q_count, dtype=np.float64)
        self._check(self.L.swg_gpu_get_query_seconds(self.ctx, out.ctypes.data, self.q_count), "get_query_seconds")
        return out

    def query_kernels(self) -> np.ndarray:
        """0 = sequence-pair kernel, 1 = query-pair kernel, per query of the last run."""
        out = np.zeros(self.q_count, dtype=np.int32)
        self._check(self.L.swg_gpu_get_query_kernels(self.ctx, out.ctypes.data, self.q_count), "get_query_kernels")
        return out

    def debug_read(self, name: str, dtype=np.uint8) -> np.ndarray:
        n = C.c_uint64(0)
        self._check(self.L.swg_gpu_debug_read(self.ctx, name.encode(), None, 0, C.byref(n)), "debug_read")
        buf = np.zeros(n.value, dtype=np.uint8)
        if n.value:
            self._check(self.L.swg_gpu_debug_read(self.ctx, name.encode(), buf.ctypes.data, n.value, C.byref(n)), "debug_read")
        return buf.view(dtype)

    def pipebench(self) -> dict:
        n, sms = C.c_int(0), C.c_int(0)
        rates = np.zeros(64, dtype=np.float64)
        mhz = np.zeros(64, dtype=np.float64)
        names = (C.c_char_p * 64)()
        self._check(self.L.swg_gpu_pipebench(self.ctx, 64, rates.ctypes.data, mhz.ctypes.data, names, C.byref(n),
                                             C.byref(sms)), "pipebench")
        out = {"sms": sms.value, "probes": {}}
        for p in range(n.value):
            out["probes"][names[p].decode()] = {
                "ginstr_per_s": float(rates[p]), "sm_mhz": float(mhz[p]),
                "thread_instr_per_clk_per_sm": float(rates[p] * 1e9 / (mhz[p] * 1e6) / sms.value) if mhz[p] else 0.0}
        return out


def compat_search_avx2(query_sequences, m, query_disp, vect_db, vect_lengths, vect_blocks, vect_disp, submat, go, ge,
                       n_threads=0, block=0):
    """Call the reference-signature entry point (CPUsearch.h:37-39) with numpy buffers."""
    L = load_library()
    q = np.ascontiguousarray(query_sequences, dtype=np.int8)
    m = np.ascontiguousarray(m, dtype=np.uint16)
    qd = np.ascontiguousarray(query_disp, dtype=np.uint32)
    vdb = np.ascontiguousarray(vect_db, dtype=np.int8)
    vl = np.ascontiguousarray(vect_lengths, dtype=np.uint16)
    vb = np.ascontiguousarray(vect_blocks, dtype=np.uint16)
    vd = np.ascontiguousarray(vect_disp, dtype=np.uint64)
    sm = np.ascontiguousarray(submat, dtype=np.int8)
    scores = np.full((len(m), len(vl) * 32), -1, dtype=np.int32)
    wt = C.c_double(0)
    st = L.swimm_gpu_search_avx2_compat(q.ctypes.data, m.ctypes.data, len(m), qd.ctypes.data, vdb.ctypes.data,
                                        vl.ctypes.data, vb.ctypes.data, len(vl), vd.ctypes.data, sm.ctypes.data, go, ge,
                                        n_threads, block, scores.ctypes.data, C.byref(wt))
    if st != 0:
        raise SwgError("swimm_gpu_search_avx2_compat -> %d: %s" % (st, L.swg_gpu_last_error(None).decode()))
    return scores, wt.value
