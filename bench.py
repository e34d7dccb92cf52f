#!/usr/bin/env python
"""bench.py -- reads/s of quack's per-read statistics accumulation on B200 (BASELINE.json config 2:
paired-end 2 x 150 bp, 10 M pairs per GPU, adapter set all.fa.gz).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (the reference's own CPU path, host cores)

One step = one pass of the hot path over the rank's 10 M synthetic pairs (2 kernel launches, one per
mate, then qb_finish of both mates = NCCL reduce to rank 0 + D2H of the count arrays).
  value    : whole-job reads/s with the batches resident in HBM (CUDA events on the launching stream,
             max over ranks); inputs (3.1 GB per launch) are far larger than the 126 MB L2.
  e2e      : the same workload through the streaming C-ABI from pinned HOST buffers: every step copies
             all inputs host->device (ring of 3 slots per GPU, copies overlapped with kernels) and
             reads the count arrays back.
  roofline : algorithmic bytes (2*l + 8 per read) / mean launch time of the statistics
             kernel, measured live with a CUDA event pair around every launch of the timed region,
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline : the unmodified reference read_fastq() (oracle/_ref) on a bounded uncompressed sample,
             1 core (it is single-threaded), timed on this box.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

# stdout carries exactly one JSON line, and NCCL writes its version banner there (NCCL_DEBUG=VERSION, which
# the GPU boxes set; NCCL honours NCCL_DEBUG_FILE only above that level): WARN + log file = stderr
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 150
SEED = 2
ADAPTER_RATE = 0.1
METRIC = "reads/s, quack per-read statistics, paired-end 2x150 bp, 10M pairs per GPU, -a all.fa.gz"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled while the timed region runs.  In-process NVML (nvidia-ml-py) every few
    milliseconds: spawning nvidia-smi next to a 10 ms timed region perturbs it (its NVML start-up takes driver locks;
    measured: +0.1-0.2 ms per 2 ms step).  nvidia-smi -lms stays as the fallback when NVML cannot be loaded."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.stop_flag = threading.Event()
        self.nvml, self.handle, self.samples, self.mask, self.max_mhz = None, None, [], 0, None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                pr = torch.cuda.get_device_properties(gpu_index)
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.nvml, self.handle = pynvml, h
        except Exception:
            self.nvml = None

    def _reasons(self):
        n = self.nvml
        for f in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
            if hasattr(n, f):
                return int(getattr(n, f)(self.handle))
        return 0

    def run(self):
        if self.nvml is not None:
            while not self.stop_flag.is_set():
                try:
                    self.samples.append(float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
                    self.mask |= self._reasons()
                except Exception:
                    pass
                self.stop_flag.wait(0.004)
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        if self.nvml is not None:
            return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "reasons": [name for name, bit in self.REASONS if self.mask & bit], "samples": len(self.samples),
                    "source": "nvml"}
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6),
                          ("sw_power_cap", 7)):
            if any(len(r) >= 8 and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi"}


# --------------------------------------------------------------------------------- inputs on disk

ADAPTER_FA = os.path.join(ROOT, "tests", "golden", "adapters_all.fa")
GEN_BIN = os.path.join(ROOT, "quack_b200", "bin", "qb_gen_fastq")
QUACK_BIN = os.path.join(ROOT, "quack_b200", "bin", "quack")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "quack")
WORKLOAD = ("configs[1]: paired-end 2x150 bp, adapters all.fa.gz (333 10-mers), read-through adapters in 10% of pairs")


def bench_config(pairs: int, world: int) -> dict:
    """The `config` object: the same in both arms (the reference arm says in cpu_baseline.sample what it timed)."""
    return {"workload": WORKLOAD, "pairs_per_gpu": pairs, "reads_per_step": 2 * pairs * world, "generator_seed": SEED}


def scratch_dir(need_bytes: int):
    """/dev/shm when it has the room (the files then never touch a disk), else the default temp directory."""
    try:
        st = os.statvfs("/dev/shm")
        if st.f_bavail * st.f_frsize > need_bytes * 1.25:
            return "/dev/shm"
    except OSError:
        pass
    return None


def gen_fastq(path: str, mate: int, first: int, n: int, mode: str = "plain", threads: int | None = None):
    """Config-2 reads [first, first + n) of one mate through the standalone generator tool (tools/gen_fastq.cpp):
    the same reads qb_gen_reads() produces, written without loading libquack_b200.so."""
    cmd = [GEN_BIN, path, str(SEED), str(mate), str(first), str(n), str(READ_LEN), str(READ_LEN), str(ADAPTER_RATE),
           mode, "1"]
    if threads:
        cmd.append(str(threads))
    subprocess.run(cmd, check=True)


# --------------------------------------------------------------------------------- CPU baseline

def cpu_baseline(pairs: int = 500_000, gz_pairs: int = 100_000):
    """Reference read_fastq() (or the oracle port when oracle/_ref is absent) on host cores, 1 thread."""
    from oracle import pyoracle as po
    from quack_b200 import capi
    use_ref = po.have_ref()
    out = {"unit": "reads/s", "cores": 1, "kind": "reference" if use_ref else "port"}
    with tempfile.TemporaryDirectory(dir=scratch_dir(pairs * 700)) as tmp:
        paths = {}
        for mate in (1, 2):
            paths[mate] = os.path.join(tmp, f"sample_{mate}.fq")
            gen_fastq(paths[mate], mate, 0, pairs)
            if gz_pairs:
                paths[(mate, "gz")] = os.path.join(tmp, f"sample_{mate}.fq.gz")
                gen_fastq(paths[(mate, "gz")], mate, 0, gz_pairs, "gz")
        table = po.AdapterTable.from_file(ADAPTER_FA)
        run = (lambda p: po.ref_read_fastq(p, ADAPTER_FA)) if use_ref else (lambda p: po.read_fastq(p, table))
        t0 = time.perf_counter()
        n = sum(run(paths[m]).n_reads for m in (1, 2))
        dt = time.perf_counter() - t0
        out["value"] = n / dt
        out["sample"] = (f"{pairs} pairs 2x{READ_LEN} bp (both mates), uncompressed FASTQ, -a adapters, "
                         f"{'unmodified reference read_fastq()' if use_ref else 'oracle C port'}, {dt:.2f} s")
        if gz_pairs:
            t0 = time.perf_counter()
            n = sum(run(paths[(m, "gz")]).n_reads for m in (1, 2))
            out["gz_value"] = n / (time.perf_counter() - t0)  # same path incl. zlib inflate (gzip -1 members)
        # statistics only (no inflate, no parsing): the oracle port on packed batches
        seq, qual, off, ln = capi.gen_reads(SEED, 1, 0, 200_000, READ_LEN, READ_LEN, ADAPTER_RATE)
        t0 = time.perf_counter()
        po.accumulate_batch(seq, qual, off, ln, table)
        out["stats_only_port_value"] = 200_000 / (time.perf_counter() - t0)
    return out


def e2e_file(pairs: int = 2_000_000):
    """File -> SVG, wall clock: the `quack` program of this repo and the unmodified reference binary on the SAME
    config-2 shaped .fq.gz files (multi-member gzip as SURVEY 8d specifies, and the same reads as BGZF), SVGs compared
    byte for byte; host gzip decode and CUDA start-up of the process broken out (QB_STATS_JSON)."""
    out = {"pairs": pairs, "unit": "reads/s", "reads": 2 * pairs}
    with tempfile.TemporaryDirectory(dir=scratch_dir(pairs * 700)) as tmp:
        files = {}
        for mode in ("gz", "bgzf"):
            for mate in (1, 2):
                files[(mode, mate)] = os.path.join(tmp, f"{mode}_{mate}.fq.gz")
                gen_fastq(files[(mode, mate)], mate, 0, pairs, mode)
        out["gz_bytes"] = sum(os.path.getsize(files[("gz", m)]) for m in (1, 2))
        common = ["-a", ADAPTER_FA, "-n", "cfg2"]
        svg_ref = None
        if os.path.exists(REF_BIN):
            t0 = time.perf_counter()
            r = subprocess.run([REF_BIN, "-1", files[("gz", 1)], "-2", files[("gz", 2)], *common], stdout=subprocess.PIPE,
                               stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            svg_ref = r.stdout
            out["reference"] = {"seconds": dt, "value": 2 * pairs / dt, "cores": 1, "input": "multi-member gzip",
                                "program": "oracle/_ref/quack (unmodified reference, single-threaded)"}
        # third arm: the same BGZF files with the DEVICE inflating and framing (qb_bgzf_submit): the host only reads the file
        for mode, fmode, extra_env in (("gz", "gz", {}), ("bgzf", "bgzf", {}), ("bgzf_device_inflate", "bgzf", {"QB_DEVICE_INFLATE": "1"})):
            js = os.path.join(tmp, "stats.json")
            best = None
            for _ in range(2):
                env = dict(os.environ, QB_STATS_JSON=js, **extra_env)
                t0 = time.perf_counter()
                r = subprocess.run([QUACK_BIN, "-1", files[(fmode, 1)], "-2", files[(fmode, 2)], *common],
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
                dt = time.perf_counter() - t0
                if r.returncode != 0:
                    raise RuntimeError(r.stderr.decode()[-400:])
                if best is None or dt < best[0]:
                    best = (dt, r.stdout, json.load(open(js)))
            dt, svg, st = best
            rec = {"seconds": dt, "value": 2 * pairs / dt,
                   "input": {"gz": "multi-member gzip", "bgzf": "BGZF", "bgzf_device_inflate": "BGZF, inflated and framed on the device"}[mode],
                   "create_s": st["create_s"], "stream_s": st["stream_s"], "finish_s": st["finish_s"],
                   "render_s": st["render_s"],
                   # host gzip decode, broken out: inflate runs in a pool per file (QUACK_DECODE_THREADS), ahead of the
                   # framing code; MBps_text = decompressed bytes of both mates / the time the stream took
                   "host_decode": {"MBps_text": st["text_bytes"] / max(st["stream_s"] - st["create_s"], 1e-9) / 1e6,
                                   "threads_per_file": st["decode_threads"],
                                   "seconds_framing_waited_for_inflate": st["host_gzip_decode_s_max_over_mates"],
                                   "cores": os.cpu_count()},
                   "svg_identical_to_reference": (svg == svg_ref) if svg_ref is not None else None}
            if svg_ref is not None and svg != svg_ref:
                raise AssertionError(f"SVG of the {mode} run differs from the reference's")
            if svg_ref is not None:
                rec["speedup_vs_reference"] = out["reference"]["seconds"] / dt
            out["ours_" + mode] = rec
    return out


_REF_ADAPTERS = None


def _ref_worker(job):
    path1, path2 = job
    from oracle import pyoracle as po
    if po.have_ref():
        return po.ref_read_fastq(path1, ADAPTER_FA).n_reads + po.ref_read_fastq(path2, ADAPTER_FA).n_reads
    t = po.AdapterTable.from_file(ADAPTER_FA)
    return po.read_fastq(path1, t).n_reads + po.read_fastq(path2, t).n_reads


def bench_reference(args):
    """--impl reference: the reference's own read_fastq() on this box's host cores, on the GPU arm's workload.  The
    reference is single-threaded, so 'all the host threads it can use' = one independent reference instance per
    core, each over its shard of the step's pairs (uncompressed FASTQ text: no inflate in the timed region, which
    favours the reference).  libquack_b200.so is never loaded here; the inputs come from the standalone generator."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import pyoracle as po
    kind = "reference" if po.have_ref() else "port"
    if po.have_ref():
        po.ref()          # the parent maps oracle/_ref/libquack_ref.so too (the workers are forks of it)
    else:
        po.lib()
    cores = min(os.cpu_count() or 1, 64)
    pairs = int(os.environ.get("QB_BENCH_PAIRS", str(args.pairs)))
    # the step's workload is `pairs` per GPU; with N > 1 the CPU arm times a bounded sample of it: one GPU's share
    sample_pairs = int(os.environ.get("QB_REF_SAMPLE_PAIRS", str(pairs)))
    shard = (sample_pairs + cores - 1) // cores
    with tempfile.TemporaryDirectory(dir=scratch_dir(sample_pairs * 700)) as tmp:
        jobs = []
        for i in range(cores):
            first, n = i * shard, min(shard, sample_pairs - i * shard)
            if n <= 0:
                break
            p = [os.path.join(tmp, f"shard{i}_{m}.fq") for m in (1, 2)]
            for m in (1, 2):
                gen_fastq(p[m - 1], m, first, n)
            jobs.append(tuple(p))
        times = []
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            for i in range(args.warmup + args.steps):
                t0 = time.perf_counter()
                n = sum(pool.map(_ref_worker, jobs, chunksize=1))
                dt = time.perf_counter() - t0
                assert n == 2 * sample_pairs, (n, sample_pairs)
                if i >= args.warmup:
                    times.append(dt)
        sec = sum(times) / len(times)
        value = n / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "bases_per_s": value * READ_LEN,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(pairs, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": len(jobs), "kind": kind,
                         "sample": (f"{sample_pairs} pairs per step ({'all of' if args.gpus <= 1 else 'one GPU share of'} the "
                                    f"step's workload), {len(jobs)} independent single-threaded reference instances, one "
                                    f"shard of {shard} pairs each, uncompressed FASTQ text (kseq parsing inside, no inflate)")},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------- GPU arm

class Arena:
    """The step's inputs in pinned host memory, laid out as slot-sized batches (what the host reader
    would have produced), generated by the library's deterministic generator."""

    def __init__(self, capi, first_read: int, pairs: int, reads_per_batch: int):
        self.capi, self.batches, self.bytes = capi, {0: [], 1: []}, 0
        L = capi.lib()
        for mate in (0, 1):
            for r0 in range(0, pairs, reads_per_batch):
                n = min(reads_per_batch, pairs - r0)
                seq = L.qb_host_alloc(n * READ_LEN + 64)
                qual = L.qb_host_alloc(n * READ_LEN + 64)
                off = L.qb_host_alloc(n * 4 + 64)
                ln = L.qb_host_alloc(n * 4 + 64)
                if not (seq and qual and off and ln):
                    raise MemoryError("pinned host allocation failed")
                nb = ctypes.c_uint64()
                rc = L.qb_gen_reads(SEED, mate + 1, first_read + r0, n, READ_LEN, READ_LEN, ADAPTER_RATE, seq, qual,
                                    off, ln, ctypes.byref(nb))
                if rc:
                    raise RuntimeError(f"qb_gen_reads failed: {rc}")
                self.batches[mate].append((seq, qual, off, ln, n, nb.value))
                self.bytes += 2 * nb.value + 8 * n

    def free(self):
        L = self.capi.lib()
        for m in self.batches.values():
            for seq, qual, off, ln, _, _ in m:
                for p in (seq, qual, off, ln):
                    L.qb_host_free(p)
        self.batches = {0: [], 1: []}


def bench_e2e_bgzf(capi, keys, local_rank, rank, world, args, barrier, allmax, allsum):
    """Own context (8 slots of 128 MiB, no NCCL: every rank checks its own counts), one host thread per mate like the
    `quack` program, so that up to eight chunks (~16 k blocks) are being inflated at a time."""
    import torch
    L = capi.lib()
    cpairs = int(os.environ.get("QB_BENCH_BGZF_PAIRS", "4000000"))
    mb = int(os.environ.get("QB_BENCH_BGZF_BATCH_MB", "128"))
    ring = int(os.environ.get("QB_BENCH_BGZF_RING", "8"))
    ctx = capi.Context(READ_LEN, n_mates=2, adapter_keys=keys, device_ids=[local_rank], batch_bytes=mb << 20, ring_depth=ring)
    cap = ctx.text_cap()
    bufs, chunks = [], {0: [], 1: []}
    comp_bytes = text_bytes = 0
    with tempfile.TemporaryDirectory(dir=scratch_dir(cpairs * 400 * max(world, 1))) as tmp:
        for mate in (0, 1):
            path = os.path.join(tmp, f"r{rank}_{mate}.fq.gz")
            gen_fastq(path, mate + 1, rank * cpairs, cpairs, "bgzf", threads=max(2, (os.cpu_count() or 8) // max(world, 1)))
            n = os.path.getsize(path)
            p = L.qb_host_alloc(n + 64)
            if not p:
                raise MemoryError("pinned host allocation failed")
            bufs.append(p)
            with open(path, "rb") as f:
                got = f.readinto((ctypes.c_char * n).from_address(p))
            assert got == n
            os.unlink(path)
            off = 0
            while off < n:  # chunks of whole blocks that hold at most `cap` bytes of text
                whole, text = ctypes.c_uint64(0), ctypes.c_uint64(0)
                rc = L.qb_bgzf_fit(p + off, n - off, cap, ctypes.byref(whole), ctypes.byref(text))
                if rc or whole.value == 0:
                    raise RuntimeError(f"qb_bgzf_fit failed at {off} of {n}: {rc}")
                chunks[mate].append((p + off, whole.value))
                off += whole.value
                text_bytes += text.value
            comp_bytes += n

    def feed(mate, errors):
        try:
            for i, (ptr, nb) in enumerate(chunks[mate]):
                ctx.bgzf_submit_from(mate, ptr, nb, i == len(chunks[mate]) - 1)
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)

    def one_pass():
        errors = []
        th = [threading.Thread(target=feed, args=(m, errors)) for m in (0, 1)]
        [t.start() for t in th]
        [t.join() for t in th]
        if errors:
            raise errors[0]
        return ctx.finish(0), ctx.finish(1)

    steps = max(1, min(args.steps, 5))
    ctx.reset(0)
    ctx.reset(1)
    one_pass()
    ctx.reset(0)
    ctx.reset(1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = one_pass()
    torch.cuda.synchronize()
    dt = allmax(time.perf_counter() - t0) / steps
    for mate in (0, 1):
        n, tail = ctx.text_status(mate)
        assert (n, tail) == (cpairs * steps, 0), (n, tail)
    barrier()
    for r in res:
        assert r.n_reads == cpairs * steps and r.max_length == READ_LEN, (r.n_reads, r.max_length)
        assert int(r.rows[READ_LEN - 1, capi.COL_LENGTH]) == r.n_reads
    ctx.close()
    for p in bufs:
        L.qb_host_free(p)
    return {"value": 2 * cpairs * world / dt, "unit": "reads/s", "pairs_per_gpu": cpairs, "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": comp_bytes, "text_bytes_per_step": text_bytes,
            "h2d_gbs_achieved": comp_bytes / dt / 1e9, "text_gbs_per_gpu": text_bytes / dt / 1e9,
            "chunks_per_step": len(chunks[0]) + len(chunks[1]),
            "input": "BGZF (zlib level 1) in pinned host memory, qb_bgzf_submit_from: inflate + framing + statistics on the device"}


def bench_ours(args):
    import torch
    import torch.distributed as dist
    import quack_b200
    from quack_b200 import capi, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: quack_b200 has no CPU fallback")
    if not os.path.exists(quack_b200.lib_path()):
        raise SystemExit("libquack_b200.so missing: run python -c 'import __graft_entry__ as g; g.build()'")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # every rank streams its batches from its own pinned arena: keep the ranks' host threads on disjoint cores
        try:
            cpus = sorted(os.sched_getaffinity(0))
            per = max(1, len(cpus) // world)
            os.sched_setaffinity(0, set(cpus[local_rank * per:(local_rank + 1) * per]) or set(cpus))
        except (AttributeError, OSError):
            pass

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    pairs = int(os.environ.get("QB_BENCH_PAIRS", str(args.pairs)))
    keys = synth.adapter_keys()
    batch_bytes = 64 << 20
    reads_per_batch = batch_bytes // READ_LEN
    ctx = capi.Context(READ_LEN, n_mates=2, adapter_keys=keys, device_ids=[local_rank], batch_bytes=batch_bytes,
                       batch_reads=reads_per_batch, ring_depth=3)
    if world > 1:
        idt = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.comm_init_rank(world, rank, bytes(idt.cpu().numpy().tobytes()))

    first = rank * pairs  # reads shard naturally: rank r owns pairs [r*pairs, (r+1)*pairs)
    db = [ctx.generate(SEED, 1, first, pairs, READ_LEN, READ_LEN, ADAPTER_RATE, 0),
          ctx.generate(SEED, 2, first, pairs, READ_LEN, READ_LEN, ADAPTER_RATE, 0)]
    alg_bytes_launch = 2 * db[0].info[1] + 8 * db[0].info[0]

    def step_resident():
        db[0].run(0)
        db[1].run(1)
        return ctx.finish(0), ctx.finish(1)

    # ---- value: batches resident in HBM ----
    for _ in range(args.warmup):
        step_resident()
    barrier()
    launches0 = ctx.launch_count
    ctx.profile_enable(2 * args.steps)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    w0 = time.perf_counter()
    ctx.timer_start(0)
    for _ in range(args.steps):
        res = step_resident()
    ms_dev = ctx.timer_stop(0)
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    prof = ctx.profile_collect()
    ms_step = allmax(ms_dev / args.steps)
    wall_ms_step = allmax((w1 - w0) * 1e3 / args.steps)
    total_reads_step = 2 * pairs * world
    value = total_reads_step / (ms_step * 1e-3)
    gpu_launches = int(allsum(launches))

    # sanity on the result the timed loop produced (linearity: every step adds the same counts)
    done = args.warmup + args.steps
    checked = False
    if rank == 0:
        for r in res:
            assert r.n_reads == pairs * world * done and r.max_length == READ_LEN, (r.n_reads, r.max_length)
            assert int(r.rows[READ_LEN - 1, capi.COL_LENGTH]) == r.n_reads
            assert np.all(r.rows[:, capi.COL_CONTENT:capi.COL_CONTENT + 4].sum(axis=1) == r.n_reads)
            assert np.all(r.rows[:, :91].sum(axis=1) == r.n_reads)
            assert np.all(r.rows % done == 0)
        checked = True

    ker_ms = [m for m, _ in prof if m > 0]
    ker_mean = sum(ker_ms) / len(ker_ms) if ker_ms else float("nan")
    peak, peak_src = load_peaks()
    achieved = alg_bytes_launch / (ker_mean * 1e-3) / 1e9
    # which kernel the timed launches took, and its measured DRAM traffic (one `ncu --set full` capture of the same
    # kernel on the same read shape, dram__bytes_read.sum + dram__bytes_write.sum, scaled per read)
    kname = "qb::period_kernel<adapters,5 steps,20 warps>" if ctx.period_launch_count else "qb::fused_kernel<true,96>"
    traffic, traffic_src = None, None
    try:  # one `ncu --set full` capture of this kernel at this launch size (profiles/bench_traffic.json says which)
        with open(os.path.join(ROOT, "profiles", "bench_traffic.json")) as f:
            tj = json.load(f)
            if tj.get("kernel_family") == ("period" if ctx.period_launch_count else "fused"):
                traffic = tj["dram_bytes_per_read"] * db[0].info[0]
                traffic_src = tj.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": kname,
                "algorithmic_bytes_per_launch": alg_bytes_launch, "kernel_ms_mean": ker_mean,
                "kernel_ms_min": min(ker_ms) if ker_ms else None, "launches_timed": len(ker_ms),
                "kernel_share_of_step": 2 * ker_mean / (ms_dev / args.steps)}

    # ---- e2e: the same workload streamed from pinned host memory through qb_submit_from ----
    e2e_pairs = int(os.environ.get("QB_BENCH_E2E_PAIRS", str(pairs)))
    h2d_gbs = ctx.measure_h2d(256 << 20, 3, 0)
    # the N-GPU ingest ceiling: every rank copies at the same time (shared PCIe switches / host memory show up here)
    barrier()
    h2d_conc = ctx.measure_h2d(256 << 20, 6, 0)
    barrier()
    h2d_conc_sum = allsum(h2d_conc)
    h2d_conc_min = -allmax(-h2d_conc)
    arena = Arena(capi, first, e2e_pairs, reads_per_batch)
    # the streaming ceiling: every rank copies its own arena once, all at the same time, no kernels.  (The 256 MiB
    # repeated copy above can be served from the CPUs' last-level cache; one pass over 6 GB per rank cannot -- on a
    # node whose host memory feeds 8 GPUs this is the ceiling that counts.)
    ptrs, sizes = [], []
    for m in (0, 1):
        for seq_p, qual_p, _, _, _, nbytes in arena.batches[m]:
            ptrs += [seq_p, qual_p]
            sizes += [nbytes, nbytes]
    ctx.measure_h2d_list(ptrs[:4], sizes[:4], 0)
    barrier()
    h2d_stream = ctx.measure_h2d_list(ptrs, sizes, 0)
    barrier()
    h2d_stream_sum = allsum(h2d_stream)
    h2d_stream_min = -allmax(-h2d_stream)
    ctx.reset(0)
    ctx.reset(1)

    def step_e2e():
        for mate in (0, 1):
            for seq, qual, off, ln, n, nb in arena.batches[mate]:
                ctx.submit_from(mate, seq, qual, off, ln, n, nb, READ_LEN)
        return ctx.finish(0), ctx.finish(1)

    e2e_steps = max(1, min(args.steps, int(os.environ.get("QB_BENCH_E2E_STEPS", str(args.steps)))))
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    ctx.reset(0)
    ctx.reset(1)
    barrier()
    h2d0 = ctx.h2d_bytes
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res2 = step_e2e()
    torch.cuda.synchronize()
    e2e_s = allmax(time.perf_counter() - t0) / e2e_steps
    h2d_step = (ctx.h2d_bytes - h2d0) // e2e_steps  # bytes the library really queued for H2D copy per step
    barrier()
    if rank == 0 and e2e_pairs == pairs:
        for a, b in zip(res, res2):  # streamed path produced the same counts as the resident path
            assert np.array_equal(a.rows // done, b.rows // e2e_steps)
    d2h = 2 * (READ_LEN * capi.ROW + 4) * 8
    e2e = {"value": 2 * e2e_pairs * world / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": h2d_step,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3, "pairs_per_gpu": e2e_pairs,
           "host_batch_bytes_per_step": arena.bytes,
           "h2d_gbs_achieved": h2d_step / e2e_s / 1e9, "h2d_gbs_link_measured": h2d_gbs,
           "frac_of_h2d_roofline": h2d_step / e2e_s / 1e9 / h2d_gbs,
           "h2d_gbs_all_ranks_copying": {"sum": h2d_conc_sum, "slowest_rank": h2d_conc_min,
                                         "note": "256 MiB pinned copies repeated by every rank at once (source may stay in the CPUs' caches)"},
           "h2d_gbs_all_ranks_streaming": {"sum": h2d_stream_sum, "slowest_rank": h2d_stream_min,
                                           "note": "every rank copies its whole pinned arena once, all ranks at once, no kernels: "
                                                   "the node's ingest ceiling for input that is read from host memory once"},
           "frac_of_concurrent_h2d_roofline": h2d_step * world / e2e_s / 1e9 / h2d_stream_sum,
           # the step ends when the slowest rank has streamed its share (weak scaling, no work stealing across processes)
           "frac_of_slowest_rank_streaming": h2d_step / e2e_s / 1e9 / h2d_stream_min,
           "note": "offsets/lengths (8 B/read) stay on the host for batches the period kernel takes: the host verified their shape"}
    arena.free()
    for b in db:
        b.free()

    # ---- e2e from COMPRESSED host buffers (secondary): BGZF blocks in pinned memory -> qb_bgzf_submit_from: the device
    # inflates, frames and counts; the host-to-device copy carries ~half the bytes per read.  Slower than the raw path
    # while one link feeds one GPU (the inflate kernel, not the link, is the limit); it matters where the node's host
    # fabric is the wall (N = 8). ----
    e2e_bgzf = None
    if not args.no_e2e_bgzf:
        try:
            ctx.close()
            ctx = None
            e2e_bgzf = bench_e2e_bgzf(capi, keys, local_rank, rank, world, args, barrier, allmax, allsum)
        except Exception as ex:
            e2e_bgzf = {"error": repr(ex)}
    if ctx is not None:
        ctx.close()

    # ---- the other kernels of the path, kernel only, CUDA events (rank 0; secondary to the headline) ----
    other = {}
    if rank == 0 and not args.no_other_kernels:
        def kernel_only(len_min, len_max, cap, ad, n):
            with capi.Context(cap, adapter_keys=keys if ad else None, device_ids=[local_rank]) as c2:
                b = c2.generate(SEED if len_min == len_max else 4, 1, 0, n, len_min, len_max, ADAPTER_RATE)
                nr, nb = b.info
                avg, mn = b.time(0, warmup=3, iters=10, flush_l2=False)
                b.free()
            alg = 2 * nb + 8 * nr
            return {"reads_per_launch": nr, "kernel_ms_mean": avg, "achieved": alg / avg / 1e6, "unit": "GB/s",
                    "frac": alg / avg / 1e6 / peak, "reads_per_s": nr / avg * 1e3}
        other["config1_5_150bp_no_adapters"] = kernel_only(READ_LEN, READ_LEN, READ_LEN, False, 10_000_000)
        other["config4_ragged_35_300_no_adapters"] = kernel_only(35, 300, 304, False, 5_000_000)
        other["config4_ragged_35_300_adapters"] = kernel_only(35, 300, 304, True, 5_000_000)
        try:    # the inflate kernel of the compressed-input path (qb_bgzf_submit): one warp per BGZF block; not an HBM-bound
            # kernel (a serial bit stream per block: instruction-issue bound), so throughput only, no roofline fraction
            with tempfile.TemporaryDirectory(dir=scratch_dir(2_000_000 * 400)) as tmp:
                pth = os.path.join(tmp, "inflate.fq.gz")
                gen_fastq(pth, 1, 0, 2_000_000, "bgzf")
                comp = open(pth, "rb").read()
            with capi.Context(READ_LEN, device_ids=[local_rank]) as c2:
                ms, n_text, n_blocks = c2.bgzf_inflate_bench(comp, 5)
            other["bgzf_inflate_150bp_level1"] = {"blocks_per_launch": n_blocks, "kernel_ms_mean": ms, "text_GBps": n_text / ms / 1e6,
                                                  "compressed_GBps": len(comp) / ms / 1e6, "reads_per_s": 2_000_000 / ms * 1e3,
                                                  "bound": "issue"}
            del comp
        except Exception as ex:
            other["bgzf_inflate_150bp_level1"] = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "bases_per_s": value * READ_LEN,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "wall_ms_per_step": wall_ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": bench_config(pairs, world),
            "notes": {"parallelism": f"shard{world}: reads shard over the ranks, one ncclReduce of the count arrays",
                      "l2": "inputs (3.1 GB per launch) larger than L2"},
            "roofline": roofline, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks, "checked": checked,
        }
        if e2e_bgzf is not None:
            line["e2e_compressed"] = e2e_bgzf
        if other:
            line["roofline_other_kernels"] = other
        if world == 1 and not args.no_e2e_file:
            try:
                line["e2e_file"] = e2e_file(int(os.environ.get("QB_BENCH_FILE_PAIRS", "2000000")))
            except Exception as ex:
                line["e2e_file"] = {"error": repr(ex)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline()
            except Exception as ex:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"error": repr(ex)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=10_000_000, help="pairs per GPU (configs[1]: 10M)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-bgzf", action="store_true", help="skip the e2e line from compressed (BGZF) host buffers")
    ap.add_argument("--no-e2e-file", action="store_true", help="skip the file -> SVG comparison with the reference binary")
    ap.add_argument("--no-other-kernels", action="store_true", help="skip the secondary kernel-only numbers")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
