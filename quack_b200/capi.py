"""ctypes mirror of include/quack_b200.h (same names, same argument meaning, same error codes).

Raises at import-of-use time when the CUDA library has not been built: there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import lib_path

ROW = 97
COL_CONTENT, COL_LENGTH, COL_KMER = 91, 95, 96
KERNEL_AUTO, KERNEL_SIMPLE, KERNEL_FUSED, KERNEL_WTILE, KERNEL_PERIOD, KERNEL_FLAT = 0, 1, 2, 3, 4, 5
NCCL_ID_BYTES = 128

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


class QbConfig(C.Structure):
    _fields_ = [
        ("n_devices", C.c_int), ("device_ids", C.POINTER(C.c_int)), ("len_cap", C.c_uint32),
        ("n_mates", C.c_int), ("adapters_enabled", C.c_int), ("adapter_keys", _u32p),
        ("n_adapter_keys", C.c_uint32), ("batch_bytes", C.c_uint64), ("batch_reads", C.c_uint32),
        ("ring_depth", C.c_int), ("kernel", C.c_int),
    ]


class QbBatch(C.Structure):
    _fields_ = [
        ("seq", _u8p), ("qual", _u8p), ("offset", _u32p), ("length", _u32p),
        ("cap_bytes", C.c_uint64), ("cap_reads", C.c_uint32), ("device_index", C.c_int), ("slot", C.c_int),
    ]


class QbText(C.Structure):
    _fields_ = [("text", C.POINTER(C.c_uint8)), ("cap_bytes", C.c_uint64), ("device_index", C.c_int), ("slot", C.c_int)]


class QbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"quack_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(quack_b200 has no CPU fallback)")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.qb_create.argtypes = [C.POINTER(QbConfig), C.POINTER(vp)]
    L.qb_destroy.argtypes = [vp]
    L.qb_destroy.restype = None
    L.qb_last_error.argtypes = [vp]
    L.qb_last_error.restype = C.c_char_p
    L.qb_acquire.argtypes = [vp, C.POINTER(QbBatch)]
    L.qb_submit.argtypes = [vp, C.POINTER(QbBatch), C.c_int, C.c_uint32, C.c_uint64, C.c_uint32]
    L.qb_submit_from.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_uint32, C.c_uint64, C.c_uint32]
    L.qb_accumulate_host.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_uint64]
    L.qb_sync.argtypes = [vp]
    L.qb_text_acquire.argtypes = [vp, C.POINTER(QbText)]
    L.qb_text_submit.argtypes = [vp, C.POINTER(QbText), C.c_int, C.c_uint64, C.c_int]
    L.qb_text_status.argtypes = [vp, C.c_int, _u64p, _u64p]
    L.qb_bgzf_submit.argtypes = [vp, C.POINTER(QbText), C.c_int, C.c_uint64, C.c_int]
    L.qb_bgzf_submit_from.argtypes = [vp, C.c_int, vp, C.c_uint64, C.c_int]
    L.qb_text_submit_from.argtypes = [vp, C.c_int, vp, C.c_uint64, C.c_int]
    L.qb_text_capacity.argtypes = [vp, _u64p]
    L.qb_bgzf_fit.argtypes = [vp, C.c_uint64, C.c_uint64, _u64p, _u64p]
    L.qb_bgzf_inflate.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, _u64p]
    L.qb_extras_enable.argtypes = [vp]
    L.qb_extras_finish.argtypes = [vp, C.c_int, _u64p, _u64p, C.c_uint64, _u64p]
    L.qb_bgzf_inflate_bench.argtypes = [vp, vp, C.c_uint64, C.c_int, C.POINTER(C.c_float), _u64p, _u32p]
    L.fqr_read_raw.argtypes = [vp, vp, C.c_size_t]
    L.fqr_read_raw.restype = C.c_long
    L.qb_finish.argtypes = [vp, C.c_int, vp, C.c_uint64, _u64p, _u64p]
    L.qb_finish_transformed.argtypes = [vp, C.c_int, vp, C.c_uint64, _u64p, _u64p, _u64p]
    L.qb_reset.argtypes = [vp, C.c_int]
    L.qb_invalid_quality_count.argtypes = [vp, C.c_int, _u64p]
    L.qb_nccl_unique_id.argtypes = [vp]
    L.qb_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, vp]
    L.qb_dbatch_upload.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_uint32, C.c_uint64, C.c_uint32, C.POINTER(vp)]
    L.qb_dbatch_generate.argtypes = [vp, C.c_int, C.c_uint64, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32,
                                     C.c_uint32, C.c_double, C.POINTER(vp)]
    L.qb_dbatch_run.argtypes = [vp, vp, C.c_int]
    L.qb_dbatch_time.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                 C.POINTER(C.c_float)]
    L.qb_dbatch_info.argtypes = [vp, _u32p, _u64p]
    L.qb_dbatch_free.argtypes = [vp, vp]
    L.qb_dbatch_free.restype = None
    L.qb_launch_count.argtypes = [vp]
    L.qb_launch_count.restype = C.c_uint64
    L.qb_kernel_counts.argtypes = [vp, _u64p, _u64p]
    L.qb_period_launch_count.argtypes = [vp]
    L.qb_period_launch_count.restype = C.c_uint64
    L.qb_flat_launch_count.argtypes = [vp]
    L.qb_flat_launch_count.restype = C.c_uint64
    L.qb_h2d_bytes.argtypes = [vp]
    L.qb_h2d_bytes.restype = C.c_uint64
    L.qb_profile_enable.argtypes = [vp, C.c_int]
    L.qb_profile_collect.argtypes = [vp, C.POINTER(C.c_float), _u64p, C.c_int]
    L.qb_timer_start.argtypes = [vp, C.c_int]
    L.qb_timer_stop.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    L.qb_measure_h2d.argtypes = [vp, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_double)]
    L.qb_measure_h2d_list.argtypes = [vp, C.c_int, C.POINTER(vp), _u64p, C.c_uint32, C.POINTER(C.c_double)]
    L.qb_base_code.argtypes = [C.c_int]
    L.qb_adapter_record_keys.argtypes = [C.c_char_p, C.c_size_t, _u32p, C.c_size_t]
    L.qb_gen_reads.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                               vp, vp, vp, vp, _u64p]
    L.qb_host_alloc.argtypes = [C.c_size_t]
    L.qb_host_alloc.restype = vp
    L.qb_host_free.argtypes = [vp]
    L.qb_host_free.restype = None
    # host program pieces linked into the same library (quack_b200/host/*.c)
    L.fqr_open.argtypes = [C.c_char_p]
    L.fqr_open.restype = vp
    L.fqr_open_mt.argtypes = [C.c_char_p, C.c_int]
    L.fqr_open_mt.restype = vp
    L.fqr_decode_threads.argtypes = [vp]
    L.fqr_close.argtypes = [vp]
    L.fqr_close.restype = None
    L.fqr_fill.argtypes = [vp, vp, vp, vp, vp, C.c_uint64, C.c_uint32, _u32p, _u64p, _u32p]
    L.fqr_status.argtypes = [vp]
    L.fqr_next.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.fqr_next.restype = C.c_long
    L.fqr_bytes_in.argtypes = [vp]
    L.fqr_bytes_in.restype = C.c_uint64
    L.fqr_inflate_seconds.argtypes = [vp]
    L.fqr_inflate_seconds.restype = C.c_double
    L.fqr_read_adapter_keys.argtypes = [C.c_char_p, C.POINTER(_u32p)]
    L.fqr_read_adapter_keys.restype = C.c_long
    L.qr_render_to_path.argtypes = [vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_int, C.c_char_p,
                                    C.c_char_p]
    L.qb_microbench.argtypes = [C.c_char_p, C.c_size_t]
    L.qb_adapter_filter_info.argtypes = [vp, _u32p, C.POINTER(C.c_double)]
    _lib = L
    return L


def base_code(c: int) -> int:
    return lib().qb_base_code(c)


def adapter_record_keys(seq: bytes) -> np.ndarray:
    """Keys read_adapters() inserts for one FASTA record (reference quack.c:165-172)."""
    out = np.zeros(max(len(seq), 1), dtype=np.uint32)
    n = lib().qb_adapter_record_keys(seq, len(seq), out.ctypes.data_as(_u32p), len(out))
    if n < 0:
        raise QbError(n, "qb_adapter_record_keys")
    return out[:n]


def gen_reads(seed: int, mate: int, first_read: int, n_reads: int, len_min: int, len_max: int,
              adapter_rate: float):
    """Deterministic synthetic reads.  Returns (seq, qual, offset, length) numpy arrays."""
    seq = np.zeros(n_reads * len_max + 64, dtype=np.uint8)
    qual = np.zeros(n_reads * len_max + 64, dtype=np.uint8)
    off = np.zeros(n_reads, dtype=np.uint32)
    ln = np.zeros(n_reads, dtype=np.uint32)
    nb = C.c_uint64()
    rc = lib().qb_gen_reads(seed, mate, first_read, n_reads, len_min, len_max, adapter_rate, seq.ctypes.data,
                            qual.ctypes.data, off.ctypes.data, ln.ctypes.data, C.byref(nb))
    if rc:
        raise QbError(rc, "qb_gen_reads")
    return seq[: nb.value], qual[: nb.value], off, ln


class Result:
    def __init__(self, rows, max_length, n_reads):
        self.rows, self.max_length, self.n_reads = rows, max_length, n_reads


class DeviceBatch:
    def __init__(self, ctx: "Context", handle):
        self.ctx, self.h = ctx, handle

    @property
    def info(self):
        n, b = C.c_uint32(), C.c_uint64()
        lib().qb_dbatch_info(self.h, C.byref(n), C.byref(b))
        return n.value, b.value

    def run(self, mate: int = 0):
        self.ctx._chk(lib().qb_dbatch_run(self.ctx.h, self.h, mate))

    def time(self, mate: int = 0, warmup: int = 3, iters: int = 10, flush_l2: bool = False):
        avg, mn = C.c_float(), C.c_float()
        self.ctx._chk(lib().qb_dbatch_time(self.ctx.h, self.h, mate, warmup, iters, int(flush_l2), C.byref(avg),
                                           C.byref(mn)))
        return avg.value, mn.value

    def free(self):
        if self.h:
            lib().qb_dbatch_free(self.ctx.h, self.h)
            self.h = None


class Context:
    """qb_ctx: what the host program creates once per run in place of calling read_fastq() per mate."""

    def __init__(self, len_cap: int, n_mates: int = 1, adapter_keys=None, n_devices: int = 1, device_ids=None,
                 batch_bytes: int = 0, batch_reads: int = 0, ring_depth: int = 0, kernel: int = KERNEL_AUTO):
        cfg = QbConfig()
        cfg.n_devices = n_devices
        self._ids = (C.c_int * n_devices)(*device_ids) if device_ids is not None else None
        cfg.device_ids = self._ids
        cfg.len_cap = len_cap
        cfg.n_mates = n_mates
        cfg.adapters_enabled = 0 if adapter_keys is None else 1
        self._keys = None
        if adapter_keys is not None:
            self._keys = np.ascontiguousarray(adapter_keys, dtype=np.uint32)
            cfg.adapter_keys = self._keys.ctypes.data_as(_u32p)
            cfg.n_adapter_keys = len(self._keys)
        cfg.batch_bytes, cfg.batch_reads, cfg.ring_depth, cfg.kernel = batch_bytes, batch_reads, ring_depth, kernel
        self.len_cap = len_cap
        h = C.c_void_p()
        rc = lib().qb_create(C.byref(cfg), C.byref(h))
        if rc:
            raise QbError(rc, lib().qb_last_error(None).decode())
        self.h = h

    def _chk(self, rc: int):
        if rc:
            raise QbError(rc, lib().qb_last_error(self.h).decode())

    def close(self):
        if self.h:
            lib().qb_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def acquire(self) -> QbBatch:
        b = QbBatch()
        self._chk(lib().qb_acquire(self.h, C.byref(b)))
        return b

    def submit(self, b: QbBatch, mate: int, n_reads: int, n_bytes: int, max_len: int = 0):
        self._chk(lib().qb_submit(self.h, C.byref(b), mate, n_reads, n_bytes, max_len))

    def submit_from(self, mate, seq_ptr, qual_ptr, off_ptr, len_ptr, n_reads, n_bytes, max_len=0):
        self._chk(lib().qb_submit_from(self.h, mate, seq_ptr, qual_ptr, off_ptr, len_ptr, n_reads, n_bytes, max_len))

    def accumulate_host(self, mate: int, seq, qual, offset, length):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        qual = np.ascontiguousarray(qual, dtype=np.uint8)
        offset = np.ascontiguousarray(offset, dtype=np.uint32)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        self._chk(lib().qb_accumulate_host(self.h, mate, seq.ctypes.data, qual.ctypes.data, offset.ctypes.data,
                                           length.ctypes.data, len(offset)))

    def text_accumulate(self, mate: int, text: bytes, cuts=None):
        """FASTQ text framed on the device (qb_text_submit).  cuts: ascending byte positions where the text is cut
        into chunks (default: the slot capacity); the last chunk ends the stream."""
        L = lib()
        t = QbText()
        self._chk(L.qb_text_acquire(self.h, C.byref(t)))
        cap = int(t.cap_bytes)
        edges = [0]
        for b in sorted(set(int(c) for c in (cuts or []) if 0 < c < len(text))) + [len(text)]:
            while b - edges[-1] > cap:
                edges.append(edges[-1] + cap)
            if b > edges[-1]:
                edges.append(b)
        if len(edges) == 1:
            edges.append(len(text))
        for i in range(len(edges) - 1):
            a, b = edges[i], edges[i + 1]
            if i:
                self._chk(L.qb_text_acquire(self.h, C.byref(t)))
            C.memmove(t.text, text[a:b], b - a)
            self._chk(L.qb_text_submit(self.h, C.byref(t), mate, b - a, 1 if i == len(edges) - 2 else 0))

    def bgzf_accumulate(self, mate: int, comp: bytes, max_chunk: int = 0):
        """A BGZF file's bytes, inflated and framed on the device (qb_bgzf_submit): chunks of whole blocks."""
        L = lib()
        t = QbText()
        pos, first = 0, True
        while first or pos < len(comp):
            first = False
            self._chk(L.qb_text_acquire(self.h, C.byref(t)))
            cap = int(t.cap_bytes)
            take = min(len(comp) - pos, max_chunk or cap, cap)
            buf = comp[pos:pos + take]
            whole, text = C.c_uint64(0), C.c_uint64(0)
            rc = L.qb_bgzf_fit(buf, len(buf), cap, C.byref(whole), C.byref(text))
            n = int(whole.value) if rc == 0 else len(buf)   # (a bad header: let the submit refuse it)
            if rc == 0 and n == 0 and take < len(comp) - pos and take < cap:
                n, buf = 0, b""                              # chunk smaller than a block: grow it
                max_chunk = take * 2
            C.memmove(t.text, buf, n)
            last = pos + n >= len(comp) or (rc == 0 and n == 0 and take == len(comp) - pos)
            self._chk(L.qb_bgzf_submit(self.h, C.byref(t), mate, n, 1 if last else 0))
            pos += n
            if last:
                break

    def text_cap(self) -> int:
        """Capacity of a text / BGZF chunk (bytes of text)."""
        c = C.c_uint64(0)
        self._chk(lib().qb_text_capacity(self.h, C.byref(c)))
        return int(c.value)

    def bgzf_submit_from(self, mate: int, ptr: int, n_bytes: int, last: bool):
        self._chk(lib().qb_bgzf_submit_from(self.h, mate, ptr, n_bytes, 1 if last else 0))

    def bgzf_inflate(self, comp: bytes, text_cap: int) -> bytes:
        out = (C.c_uint8 * max(text_cap, 1))()
        n = C.c_uint64(0)
        self._chk(lib().qb_bgzf_inflate(self.h, comp, len(comp), out, text_cap, C.byref(n)))
        return C.string_at(out, int(n.value))

    def bgzf_inflate_bench(self, comp: bytes, iters: int = 5):
        """(ms per launch, text bytes, blocks) of the inflate kernel alone."""
        ms, n, nb = C.c_float(0), C.c_uint64(0), C.c_uint32(0)
        self._chk(lib().qb_bgzf_inflate_bench(self.h, comp, len(comp), iters, C.byref(ms), C.byref(n), C.byref(nb)))
        return float(ms.value), int(n.value), int(nb.value)

    def extras_enable(self):
        self._chk(lib().qb_extras_enable(self.h))

    def extras_finish(self, mate: int = 0):
        """(n_count[max_length], qual_sum[max_length], mean_hist[94]): side outputs without a reference oracle."""
        ml = self.finish(mate).max_length
        n_count, qual_sum, mean = (np.zeros(max(ml, 1), dtype=np.uint64), np.zeros(max(ml, 1), dtype=np.uint64),
                                   np.zeros(94, dtype=np.uint64))
        self._chk(lib().qb_extras_finish(self.h, mate, n_count.ctypes.data_as(_u64p), qual_sum.ctypes.data_as(_u64p),
                                         max(ml, 1), mean.ctypes.data_as(_u64p)))
        return n_count[:ml], qual_sum[:ml], mean

    def text_status(self, mate: int = 0):
        """(records framed, bytes left behind the last complete record); raises QbError(QB_ERR_TEXT) for text the
        device does not frame."""
        n, tail = C.c_uint64(0), C.c_uint64(0)
        self._chk(lib().qb_text_status(self.h, mate, C.byref(n), C.byref(tail)))
        return int(n.value), int(tail.value)

    def sync(self):
        self._chk(lib().qb_sync(self.h))

    def reset(self, mate: int = 0):
        self._chk(lib().qb_reset(self.h, mate))

    def finish(self, mate: int = 0) -> Result:
        ml, n = C.c_uint64(), C.c_uint64()
        self._chk(lib().qb_finish(self.h, mate, None, 0, C.byref(ml), C.byref(n)))   # the reduce; longest read
        rows = np.zeros((max(int(ml.value), 1), ROW), dtype=np.uint64)
        self._chk(lib().qb_finish(self.h, mate, rows.ctypes.data, rows.shape[0], C.byref(ml), C.byref(n)))
        return Result(rows[: ml.value].copy(), int(ml.value), int(n.value))

    def finish_transformed(self, mate: int = 0):
        """(rows[max_length][97] transformed on the device, max_length, n_reads, original_max_length)."""
        ml, nr, orig = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._chk(lib().qb_finish_transformed(self.h, mate, None, 0, C.byref(ml), C.byref(nr), C.byref(orig)))
        rows = np.zeros((max(int(ml.value), 1), ROW), dtype=np.uint64)
        self._chk(lib().qb_finish_transformed(self.h, mate, rows.ctypes.data, rows.shape[0], C.byref(ml), C.byref(nr),
                                              C.byref(orig)))
        return rows[: int(ml.value)], int(ml.value), int(nr.value), int(orig.value)

    def invalid_quality_count(self, mate: int = 0) -> int:
        v = C.c_uint64()
        self._chk(lib().qb_invalid_quality_count(self.h, mate, C.byref(v)))
        return int(v.value)

    def comm_init_rank(self, n_ranks: int, rank: int, nccl_id: bytes):
        buf = (C.c_uint8 * NCCL_ID_BYTES).from_buffer_copy(nccl_id)
        self._chk(lib().qb_comm_init_rank(self.h, n_ranks, rank, buf))

    def upload(self, seq, qual, offset, length, max_len: int = 0, device_index: int = 0) -> DeviceBatch:
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        qual = np.ascontiguousarray(qual, dtype=np.uint8)
        offset = np.ascontiguousarray(offset, dtype=np.uint32)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        h = C.c_void_p()
        self._chk(lib().qb_dbatch_upload(self.h, device_index, seq.ctypes.data, qual.ctypes.data, offset.ctypes.data,
                                         length.ctypes.data, len(offset), len(seq), max_len, C.byref(h)))
        return DeviceBatch(self, h)

    def generate(self, seed, mate, first_read, n_reads, len_min, len_max, adapter_rate, device_index=0) -> DeviceBatch:
        h = C.c_void_p()
        self._chk(lib().qb_dbatch_generate(self.h, device_index, seed, mate, first_read, n_reads, len_min, len_max,
                                           adapter_rate, C.byref(h)))
        return DeviceBatch(self, h)

    def profile_enable(self, max_launches: int):
        self._chk(lib().qb_profile_enable(self.h, max_launches))

    def profile_collect(self, cap: int = 4096):
        """[(ms, algorithmic_bytes)] of the launches recorded since profile_enable()."""
        ms = (C.c_float * cap)()
        by = (C.c_uint64 * cap)()
        n = lib().qb_profile_collect(self.h, ms, by, cap)
        if n < 0:
            self._chk(n)
        return [(ms[i], by[i]) for i in range(n)]

    def timer_start(self, device_index: int = 0):
        self._chk(lib().qb_timer_start(self.h, device_index))

    def timer_stop(self, device_index: int = 0) -> float:
        ms = C.c_float()
        self._chk(lib().qb_timer_stop(self.h, device_index, C.byref(ms)))
        return ms.value

    @property
    def launch_count(self) -> int:
        return int(lib().qb_launch_count(self.h))

    @property
    def h2d_bytes(self) -> int:
        """Bytes queued for host-to-device copy by the submit calls so far."""
        return int(lib().qb_h2d_bytes(self.h))

    @property
    def flat_launch_count(self) -> int:
        """Launches that took the flat kernel (ragged batches of back-to-back reads)."""
        return int(lib().qb_flat_launch_count(self.h))

    @property
    def period_launch_count(self) -> int:
        """Launches that took the period kernel (v5)."""
        return int(lib().qb_period_launch_count(self.h))

    @property
    def kernel_counts(self):
        """(launches of the simple kernel, launches of the fused kernel)."""
        a, b = C.c_uint64(), C.c_uint64()
        self._chk(lib().qb_kernel_counts(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def measure_h2d_list(self, ptrs, sizes, device_index: int = 0) -> float:
        """GB/s of one pass over pinned host buffers (addresses, byte counts)."""
        n = len(ptrs)
        pa = (C.c_void_p * n)(*ptrs)
        sa = (C.c_uint64 * n)(*sizes)
        g = C.c_double(0)
        self._chk(lib().qb_measure_h2d_list(self.h, device_index, pa, sa, n, C.byref(g)))
        return float(g.value)

    def measure_h2d(self, nbytes: int = 256 << 20, iters: int = 5, device_index: int = 0) -> float:
        g = C.c_double()
        self._chk(lib().qb_measure_h2d(self.h, device_index, nbytes, iters, C.byref(g)))
        return g.value

    def adapter_filter_info(self):
        m, f = C.c_uint32(), C.c_double()
        lib().qb_adapter_filter_info(self.h, C.byref(m), C.byref(f))
        return m.value, f.value


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * NCCL_ID_BYTES)()
    rc = lib().qb_nccl_unique_id(buf)
    if rc:
        raise QbError(rc, "qb_nccl_unique_id")
    return bytes(buf)


def microbench() -> str:
    buf = C.create_string_buffer(1 << 16)
    rc = lib().qb_microbench(buf, len(buf))
    if rc:
        raise QbError(rc, "qb_microbench")
    return buf.value.decode()


# ------------------------------------------------------------------ host program pieces (C, in the same .so)

def read_adapters(path: str) -> np.ndarray:
    """read_adapters(), reference quack.c:154-178, through the host reader: keys in reference order."""
    p = _u32p()
    n = lib().fqr_read_adapter_keys(path.encode(), C.byref(p))
    if n < 0:
        raise OSError(f"cannot open {path}")
    out = np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].copy()
    C.CDLL(None).free(p)
    return out


def parse_records(path: str, threads: int = 0):
    """[(seq, qual or None)] and the final status, record by record through the host reader
    (threads: BGZF pool threads, 0 = default, 1 = gzread path)."""
    L = lib()
    r = L.fqr_open_mt(path.encode(), threads)
    if not r:
        raise OSError(path)
    out = []
    s, q, ql = C.c_void_p(), C.c_void_p(), C.c_size_t()
    while True:
        n = L.fqr_next(r, C.byref(s), C.byref(q), C.byref(ql))
        if n < 0:
            break
        out.append((C.string_at(s, n), C.string_at(q, ql.value) if q.value else None))
    L.fqr_close(r)
    return out, int(n)


def decode_throughput(path: str, threads: int = 0, cap_bytes: int = 32 << 20):
    """Host decode + framing + packing of a whole file into scratch batches (no GPU): dict with reads,
    decompressed bytes, wall seconds, seconds waiting for inflated data, pool threads."""
    import time
    L = lib()
    r = L.fqr_open_mt(path.encode(), threads)
    if not r:
        raise OSError(path)
    cap_reads = cap_bytes // 32 + 1
    seq = np.zeros(cap_bytes + 64, dtype=np.uint8)
    qual = np.zeros(cap_bytes + 64, dtype=np.uint8)
    off = np.zeros(cap_reads, dtype=np.uint32)
    ln = np.zeros(cap_reads, dtype=np.uint32)
    n, nb, ml = C.c_uint32(), C.c_uint64(), C.c_uint32()
    reads = bases = 0
    t0 = time.perf_counter()
    more = 1
    while more > 0:
        more = L.fqr_fill(r, seq.ctypes.data, qual.ctypes.data, off.ctypes.data, ln.ctypes.data, cap_bytes, cap_reads,
                          C.byref(n), C.byref(nb), C.byref(ml))
        reads += n.value
        bases += nb.value
    dt = time.perf_counter() - t0
    out = {"reads": reads, "bases": bases, "text_bytes": int(L.fqr_bytes_in(r)), "seconds": dt,
           "wait_inflate_s": float(L.fqr_inflate_seconds(r)), "threads": int(L.fqr_decode_threads(r)),
           "status": int(L.fqr_status(r))}
    L.fqr_close(r)
    return out


def read_batches(path: str, cap_bytes: int, cap_reads: int, threads: int = 0):
    """All batches the host reader packs from a file: [(seq, qual, offset, length, max_len)], status."""
    L = lib()
    r = L.fqr_open_mt(path.encode(), threads)
    if not r:
        raise OSError(path)
    batches = []
    more = 1
    while more > 0:
        seq = np.zeros(cap_bytes + 64, dtype=np.uint8)
        qual = np.zeros(cap_bytes + 64, dtype=np.uint8)
        off = np.zeros(cap_reads, dtype=np.uint32)
        ln = np.zeros(cap_reads, dtype=np.uint32)
        n, nb, ml = C.c_uint32(), C.c_uint64(), C.c_uint32()
        more = L.fqr_fill(r, seq.ctypes.data, qual.ctypes.data, off.ctypes.data, ln.ctypes.data, cap_bytes, cap_reads,
                          C.byref(n), C.byref(nb), C.byref(ml))
        batches.append((seq[: nb.value], qual[: nb.value], off[: n.value], ln[: n.value], ml.value))
    status = L.fqr_status(r)
    L.fqr_close(r)
    return batches, status


def render_svg(first: "Result", second: "Result | None", adapters_used: bool, name: str | None, path: str):
    """transform() + draw() of the reference (quack.c:230-856) restated: SVG for raw accumulators."""
    r1 = np.ascontiguousarray(first.rows, dtype=np.uint64)
    r2 = np.ascontiguousarray(second.rows, dtype=np.uint64) if second is not None else None
    rc = lib().qr_render_to_path(r1.ctypes.data, first.max_length, first.n_reads,
                                 r2.ctypes.data if r2 is not None else None,
                                 second.max_length if second is not None else 0,
                                 second.n_reads if second is not None else 0, int(adapters_used),
                                 name.encode() if name is not None else None, path.encode())
    if rc:
        raise OSError(path)
