"""Synthetic FASTQ text from the library's deterministic read generator.

Used by bench.py (CPU-baseline sample files) and the tests of the host program; the generator itself is
qb_gen_reads() in the C library, so files and device-resident batches hold exactly the same reads.
"""
from __future__ import annotations

import gzip
import os

import numpy as np

from . import capi

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
ADAPTER_FA = os.path.join(GOLDEN, "adapters_all.fa")


def adapter_records(path: str = ADAPTER_FA) -> list[bytes]:
    recs, cur = [], None
    with open(path, "rb") as f:
        for line in f.read().split(b"\n"):
            if line.startswith(b">"):
                if cur is not None:
                    recs.append(cur)
                cur = b""
            elif cur is not None:
                cur += line.strip()
    if cur:
        recs.append(cur)
    return recs


def adapter_keys(path: str = ADAPTER_FA) -> np.ndarray:
    """The keys read_adapters() would set for a FASTA file (one record per '>' header, single-line
    or multi-line sequences), through the library's qb_adapter_record_keys()."""
    return np.concatenate([capi.adapter_record_keys(r) for r in adapter_records(path)])


def fastq_text(seed: int, mate: int, first_read: int, n_reads: int, length: int, adapter_rate: float) -> bytes:
    """4-line FASTQ records '@r<9-digit index>/<mate>', fixed read length (vectorised assembly)."""
    seq, qual, _, _ = capi.gen_reads(seed, mate, first_read, n_reads, length, length, adapter_rate)
    name_len = 1 + 1 + 9 + 2  # '@' 'r' digits '/m'
    rec = name_len + 1 + length + 3 + length + 1
    out = np.empty((n_reads, rec), dtype=np.uint8)
    out[:, 0] = ord("@")
    out[:, 1] = ord("r")
    idx = np.arange(first_read, first_read + n_reads, dtype=np.int64)
    for d in range(9):
        out[:, 2 + 8 - d] = (idx % 10) + ord("0")
        idx //= 10
    out[:, 11] = ord("/")
    out[:, 12] = ord("0") + mate
    out[:, 13] = ord("\n")
    out[:, 14: 14 + length] = seq.reshape(n_reads, length)
    p = 14 + length
    out[:, p: p + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    out[:, p + 3: p + 3 + length] = qual.reshape(n_reads, length)
    out[:, p + 3 + length] = ord("\n")
    return out.tobytes()


def write_fastq(path: str, seed: int, mate: int, n_reads: int, length: int = 150, adapter_rate: float = 0.1,
                first_read: int = 0, gz_level: int | None = None, chunk: int = 250_000, bgzf: bool = False) -> int:
    """Writes n_reads records; gz_level=None -> plain text, else one gzip member per chunk (multi-member
    files are what the reference's gzread handles too) or, with bgzf=True, BGZF blocks (also a multi-member
    gzip file to every gzip reader; the host reader inflates them in parallel).  Returns bytes of text written."""
    total = 0
    with open(path, "wb") as f:
        for r0 in range(0, n_reads, chunk):
            n = min(chunk, n_reads - r0)
            txt = fastq_text(seed, mate, first_read + r0, n, length, adapter_rate)
            total += len(txt)
            if bgzf:
                f.write(bgzf_bytes(txt, gz_level if gz_level is not None else 1)[: -len(BGZF_EOF)])
            else:
                f.write(txt if gz_level is None else gzip.compress(txt, compresslevel=gz_level, mtime=0))
        if bgzf:
            f.write(BGZF_EOF)
    return total


BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_bytes(data: bytes, level: int = 1, block: int = 65280) -> bytes:
    """`data` as a BGZF stream (SAM spec 4.1; reference klib/bgzf.c): independent gzip members of at most
    `block` uncompressed bytes, each with its size in a 'BC' extra subfield, then the empty end marker.  Any
    gzip reader (the reference's gzread included) sees one multi-member gzip file."""
    import struct
    import zlib
    out = []
    for i in range(0, len(data), block):
        chunk = data[i: i + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        payload = co.compress(chunk) + co.flush()
        if len(payload) + 26 > 65536:   # incompressible: store
            co = zlib.compressobj(0, zlib.DEFLATED, -15)
            payload = co.compress(chunk) + co.flush()
        bsize = len(payload) + 25
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + payload +
                   struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    out.append(BGZF_EOF)
    return b"".join(out)
