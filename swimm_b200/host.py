"""ctypes binding of libswimm_host.so -- the plain-C host data layer (swimm_b200/host/*.c):
substitution matrices, `-S preprocess`, FASTA and preprocessed-database loading, top-r merge."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libswimm_host.so")
MATRICES = ["blosum45", "blosum50", "blosum62", "blosum80", "blosum90", "pam30", "pam70", "pam250"]


class SeqSetC(C.Structure):
    _fields_ = [("count", C.c_uint64), ("residues", C.c_uint64), ("lengths", C.POINTER(C.c_uint16)),
                ("offsets", C.POINTER(C.c_uint64)), ("codes", C.POINTER(C.c_int8)),
                ("titles", C.POINTER(C.c_char_p)), ("max_title", C.c_int), ("input_pos", C.POINTER(C.c_uint64))]


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libswimm_host.so is not built: run `make -C swimm_b200/host`")
        L = C.CDLL(LIB_PATH)
        L.swg_submat_index.argtypes = [C.c_char_p]
        L.swg_submat_table.argtypes = [C.c_int]
        L.swg_submat_table.restype = C.POINTER(C.c_int8)
        L.swg_encode_residue.argtypes = [C.c_int]
        L.swg_read_fasta.argtypes = [C.c_char_p, C.POINTER(SeqSetC)]
        L.swg_seqset_free.argtypes = [C.POINTER(SeqSetC)]
        L.swg_seqset_free.restype = None
        L.swg_preprocess_db.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        L.swg_load_db.argtypes = [C.c_char_p, C.POINTER(SeqSetC)]
        L.swg_load_db_headers.argtypes = [C.c_char_p, C.POINTER(SeqSetC)]
        L.swg_merge_top_keys.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]
        L.swg_merge_top_keys.restype = None
        _lib = L
    return _lib


def submat(name: str) -> np.ndarray:
    """The 24x32 signed-byte table of one scoring matrix (row = query code, column = database code)."""
    L = load_library()
    k = L.swg_submat_index(name.encode())
    if k < 0:
        raise KeyError(name)
    return np.ctypeslib.as_array(L.swg_submat_table(k), shape=(768,)).copy().reshape(24, 32)


class HostSeqSet:
    """Length-sorted, encoded sequences as the C host holds them (copied into numpy)."""

    def __init__(self, c: SeqSetC, with_titles: bool):
        n, d = c.count, c.residues
        self.count, self.residues, self.max_title = n, d, c.max_title
        self.lengths = np.ctypeslib.as_array(c.lengths, shape=(max(n, 1),))[:n].copy()
        self.offsets = np.ctypeslib.as_array(c.offsets, shape=(n + 1,)).copy()
        self.codes = np.ctypeslib.as_array(c.codes, shape=(max(d, 1),))[:d].copy()
        self.titles = [c.titles[i].decode("latin-1") for i in range(n)] if with_titles and c.titles else None


def read_fasta(path: str) -> HostSeqSet:
    L = load_library()
    c = SeqSetC()
    rc = L.swg_read_fasta(path.encode(), C.byref(c))
    if rc != 0:
        raise RuntimeError("swg_read_fasta(%s) -> %d" % (path, rc))
    out = HostSeqSet(c, True)
    L.swg_seqset_free(C.byref(c))
    return out


def preprocess_db(fasta: str, prefix: str, threads: int = 1) -> None:
    rc = load_library().swg_preprocess_db(fasta.encode(), prefix.encode(), threads, 0)
    if rc != 0:
        raise RuntimeError("swg_preprocess_db -> %d" % rc)


def load_db(prefix: str, headers: bool = False) -> HostSeqSet:
    L = load_library()
    c = SeqSetC()
    rc = L.swg_load_db(prefix.encode(), C.byref(c))
    if rc != 0:
        raise RuntimeError("swg_load_db(%s) -> %d" % (prefix, rc))
    if headers:
        rc = L.swg_load_db_headers(prefix.encode(), C.byref(c))
        if rc != 0:
            raise RuntimeError("swg_load_db_headers(%s) -> %d" % (prefix, rc))
    out = HostSeqSet(c, headers)
    L.swg_seqset_free(C.byref(c))
    return out


def merge_top_keys(parts: np.ndarray, top: int) -> np.ndarray:
    """parts: [n_parts][top] descending key lists of ONE query -> the `top` largest, descending."""
    parts = np.ascontiguousarray(parts, dtype=np.uint64)
    out = np.zeros(top, dtype=np.uint64)
    load_library().swg_merge_top_keys(parts.ctypes.data, parts.shape[0], top, out.ctypes.data)
    return out
