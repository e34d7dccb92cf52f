"""ctypes access to the parity oracle (oracle/liboracle.so) and, when it was built, to the
unmodified reference (oracle/_ref/libquack_ref.so, oracle/_ref/quack).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (quack_b200/) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROW = 97
TABLE_SIZE = 1 << 20

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


class _Stats(C.Structure):
    _fields_ = [
        ("rows", _u64p),
        ("cap", C.c_uint64),
        ("max_length", C.c_uint64),
        ("n_reads", C.c_uint64),
        ("n_invalid_qual", C.c_uint64),
    ]


def build(ref: bool = True) -> None:
    """Compile the checker (and the reference shim when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)
    if ref and os.path.isdir(os.environ.get("QB_REF", "/root/reference")):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.qo_base_code.argtypes = [C.c_int]
        L.qo_table_new.restype = C.c_void_p
        L.qo_table_free.argtypes = [C.c_void_p]
        L.qo_table_add_record.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.qo_read_adapters.argtypes = [C.c_char_p, C.c_void_p]
        L.qo_read_adapters.restype = C.c_long
        L.qo_table_keys.argtypes = [C.c_void_p, _u32p]
        L.qo_table_keys.restype = C.c_uint32
        L.qo_stats_init.argtypes = [C.POINTER(_Stats)]
        L.qo_stats_free.argtypes = [C.POINTER(_Stats)]
        L.qo_accumulate_batch.argtypes = [C.POINTER(_Stats), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_uint64, C.c_void_p]
        L.qo_read_fastq.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(_Stats)]
        L.qo_transform.argtypes = [C.POINTER(_Stats), _u64p]
        L.qo_extras.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, _u64p, _u64p, C.c_uint64, _u64p]
        L.qo_reader_open.argtypes = [C.c_char_p]
        L.qo_reader_open.restype = C.c_void_p
        L.qo_reader_next.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_size_t)]
        L.qo_reader_next.restype = C.c_long
        L.qo_reader_close.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def base_code(c: int) -> int:
    return lib().qo_base_code(c)


class AdapterTable:
    """The reference's kmers[] table (1<<20 entries) restated; keys are first-base-most-significant."""

    def __init__(self):
        self._p = lib().qo_table_new()

    def add_record(self, seq: bytes) -> None:
        lib().qo_table_add_record(self._p, seq, len(seq))

    @classmethod
    def from_file(cls, path: str) -> "AdapterTable":
        t = cls()
        n = lib().qo_read_adapters(path.encode(), t._p)
        if n < 0:
            raise OSError(f"cannot open {path}")
        return t

    @classmethod
    def from_records(cls, records) -> "AdapterTable":
        t = cls()
        for r in records:
            t.add_record(r)
        return t

    def keys(self) -> np.ndarray:
        n = lib().qo_table_keys(self._p, None)
        out = np.zeros(n, dtype=np.uint32)
        if n:
            lib().qo_table_keys(self._p, out.ctypes.data_as(_u32p))
        return out

    def as_bytes(self) -> np.ndarray:
        return np.ctypeslib.as_array(C.cast(self._p, _u8p), shape=(TABLE_SIZE,)).copy()

    def __del__(self):
        if getattr(self, "_p", None):
            lib().qo_table_free(self._p)
            self._p = None


class Result:
    """rows: uint64 [max_length][97] raw (pre-transform) counts."""

    def __init__(self, rows: np.ndarray, max_length: int, n_reads: int, n_invalid_qual: int = 0):
        self.rows, self.max_length, self.n_reads, self.n_invalid_qual = rows, max_length, n_reads, n_invalid_qual


def _take(st: _Stats) -> Result:
    ml = int(st.max_length)
    if ml:
        rows = np.ctypeslib.as_array(st.rows, shape=(ml, ROW)).copy()
    else:
        rows = np.zeros((0, ROW), dtype=np.uint64)
    res = Result(rows, ml, int(st.n_reads), int(st.n_invalid_qual))
    lib().qo_stats_free(C.byref(st))
    return res


def accumulate_batch(seq, qual, offset, length, table: AdapterTable | None) -> Result:
    """Oracle over one packed batch (the layout the C-ABI takes)."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    qual = np.ascontiguousarray(qual, dtype=np.uint8)
    offset = np.ascontiguousarray(offset, dtype=np.uint32)
    length = np.ascontiguousarray(length, dtype=np.uint32)
    st = _Stats()
    lib().qo_stats_init(C.byref(st))
    rc = lib().qo_accumulate_batch(C.byref(st), seq.ctypes.data, qual.ctypes.data, offset.ctypes.data,
                                   length.ctypes.data, len(offset), table._p if table else None)
    if rc:
        raise MemoryError
    return _take(st)


def extras(seq, qual, offset, length, rows: int):
    """PARITY UNPINNED side outputs (qo_extras): (n_count[rows], qual_sum[rows], mean_hist[94])."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    qual = np.ascontiguousarray(qual, dtype=np.uint8)
    offset = np.ascontiguousarray(offset, dtype=np.uint32)
    length = np.ascontiguousarray(length, dtype=np.uint32)
    n_count, qual_sum, mean = (np.zeros(rows, dtype=np.uint64), np.zeros(rows, dtype=np.uint64),
                               np.zeros(94, dtype=np.uint64))
    lib().qo_extras(seq.ctypes.data, qual.ctypes.data, offset.ctypes.data, length.ctypes.data, len(offset),
                    n_count.ctypes.data_as(_u64p), qual_sum.ctypes.data_as(_u64p), rows, mean.ctypes.data_as(_u64p))
    return n_count, qual_sum, mean


def read_fastq(path: str, table: AdapterTable | None) -> Result:
    st = _Stats()
    lib().qo_stats_init(C.byref(st))
    if lib().qo_read_fastq(path.encode(), table._p if table else None, C.byref(st)):
        raise OSError(f"cannot open {path}")
    return _take(st)


def parse_records(path: str):
    """[(seq, qual or None)] and the final return code, through the oracle's kseq restatement."""
    r = lib().qo_reader_open(path.encode())
    if not r:
        raise OSError(path)
    out = []
    s, q, ql = C.c_void_p(), C.c_void_p(), C.c_size_t()
    while True:
        l = lib().qo_reader_next(r, C.byref(s), C.byref(q), C.byref(ql))
        if l < 0:
            break
        out.append((C.string_at(s, l), C.string_at(q, ql.value) if q.value else None))
    lib().qo_reader_close(r)
    return out, int(l)


def transform(rows: np.ndarray, max_length: int, n_reads: int):
    """qo_transform on a copy; returns (rows, new_max_length, original_max_length)."""
    rows = np.ascontiguousarray(rows, dtype=np.uint64).copy()
    st = _Stats(rows.ctypes.data_as(_u64p), rows.shape[0], max_length, n_reads, 0)
    orig = C.c_uint64()
    lib().qo_transform(C.byref(st), C.byref(orig))
    return rows, int(st.max_length), int(orig.value)


# ---------------------------------------------------------------- unmodified reference

REF_SO = os.path.join(HERE, "_ref", "libquack_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "quack")
_ref = None


def have_ref() -> bool:
    return os.path.exists(REF_SO) and os.path.exists(REF_BIN)


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.qref_read_fastq.argtypes = [C.c_char_p, C.c_char_p, _u64p, C.c_uint64, _u64p, _u64p]
        L.qref_read_adapters.argtypes = [C.c_char_p, _u8p]
        L.qref_transform.argtypes = [_u64p, _u64p, C.c_uint64, _u64p]
        L.qref_base_code.argtypes = [C.c_int]
        L.qref_parse.argtypes = [C.c_char_p, C.c_char_p, C.c_long, C.POINTER(C.c_int)]
        L.qref_parse.restype = C.c_long
        _ref = L
    return _ref


def ref_read_fastq(path: str, adapters: str | None, cap_rows: int = 4096) -> Result:
    rows = np.zeros((cap_rows, ROW), dtype=np.uint64)
    ml, n = C.c_uint64(), C.c_uint64()
    rc = ref().qref_read_fastq(path.encode(), adapters.encode() if adapters else None,
                               rows.ctypes.data_as(_u64p), cap_rows, C.byref(ml), C.byref(n))
    if rc:
        raise ValueError("cap_rows too small")
    return Result(rows[: ml.value].copy(), int(ml.value), int(n.value))


def ref_adapter_table(path: str) -> np.ndarray:
    t = np.zeros(TABLE_SIZE, dtype=np.uint8)
    ref().qref_read_adapters(path.encode(), t.ctypes.data_as(_u8p))
    return t


def ref_transform(rows: np.ndarray, max_length: int, n_reads: int):
    rows = np.ascontiguousarray(rows, dtype=np.uint64).copy()
    ml, orig = C.c_uint64(max_length), C.c_uint64()
    ref().qref_transform(rows.ctypes.data_as(_u64p), C.byref(ml), n_reads, C.byref(orig))
    return rows, int(ml.value), int(orig.value)


def ref_parse_records(path: str, cap: int = 1 << 24):
    """[(seq, qual)] and the final kseq_read return code from the reference reader itself."""
    buf = C.create_string_buffer(cap)
    rc = C.c_int()
    n = ref().qref_parse(path.encode(), buf, cap, C.byref(rc))
    if n < 0:
        raise ValueError("cap too small")
    recs = []
    for line in buf.raw[:n].split(b"\n")[:-1]:
        s_, q_ = line.split(b"\t")
        recs.append((s_, q_ if q_ else None))
    return recs, int(rc.value)


def ref_svg(args: list[str]) -> bytes:
    """stdout of the unmodified reference binary for a CLI argument list."""
    return subprocess.run([REF_BIN, *args], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
