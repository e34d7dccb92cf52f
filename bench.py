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
    """nvidia-smi clocks and throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6),
                          ("sw_power_cap", 7)):
            if any(len(r) >= 8 and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------- CPU baseline

def make_sample_files(tmp: str, pairs: int, gz_pairs: int):
    from quack_b200 import synth
    paths = {}
    for mate in (1, 2):
        p = os.path.join(tmp, f"sample_{mate}.fq")
        synth.write_fastq(p, SEED, mate, pairs, READ_LEN, ADAPTER_RATE)
        paths[mate] = p
        if gz_pairs:
            g = os.path.join(tmp, f"sample_{mate}.fq.gz")
            synth.write_fastq(g, SEED, mate, gz_pairs, READ_LEN, ADAPTER_RATE, gz_level=1)
            paths[(mate, "gz")] = g
    return paths


def cpu_baseline(pairs: int = 500_000, gz_pairs: int = 100_000):
    """Reference read_fastq() (or the oracle port when oracle/_ref is absent) on host cores, 1 thread."""
    from oracle import pyoracle as po
    from quack_b200 import capi, synth
    use_ref = po.have_ref()
    out = {"unit": "reads/s", "cores": 1, "kind": "reference" if use_ref else "port"}
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        paths = make_sample_files(tmp, pairs, gz_pairs)
        table = po.AdapterTable.from_file(synth.ADAPTER_FA)
        run = (lambda p: po.ref_read_fastq(p, synth.ADAPTER_FA)) if use_ref else (lambda p: po.read_fastq(p, table))
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


def _ref_worker(args):
    path1, path2, adapters = args
    from oracle import pyoracle as po
    if po.have_ref():
        return po.ref_read_fastq(path1, adapters).n_reads + po.ref_read_fastq(path2, adapters).n_reads
    t = po.AdapterTable.from_file(adapters)
    return po.read_fastq(path1, t).n_reads + po.read_fastq(path2, t).n_reads


def bench_reference(args):
    """--impl reference: the reference's own read_fastq() on this box's host cores.  The reference is
    single-threaded, so the box's capacity is shown as one independent reference instance per core,
    each over the same bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import pyoracle as po
    from quack_b200 import synth
    cores = min(os.cpu_count() or 1, 64)
    pairs = int(os.environ.get("QB_REF_SAMPLE_PAIRS", "100000"))
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        paths = make_sample_files(tmp, pairs, 0)
        job = (paths[1], paths[2], synth.ADAPTER_FA)
        times = []
        with mp.get_context("fork").Pool(cores) as pool:
            for i in range(args.warmup + args.steps):
                t0 = time.perf_counter()
                n = sum(pool.map(_ref_worker, [job] * cores))
                dt = time.perf_counter() - t0
                if i >= args.warmup:
                    times.append(dt)
        sec = sum(times) / len(times)
        value = n / sec
    kind = "reference" if po.have_ref() else "port"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "bases_per_s": value * READ_LEN,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "configs[1]: paired-end 2x150 bp, adapters all.fa.gz", "sample_pairs_per_instance": pairs,
                   "instances": cores},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind,
                         "sample": f"{cores} independent single-threaded instances x {pairs} pairs, uncompressed FASTQ"},
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
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "bench_traffic.json")) as f:
            tj = json.load(f)
            if tj.get("kernel_family") == ("period" if ctx.period_launch_count else "fused"):
                traffic = tj["dram_bytes_per_read"] * db[0].info[0]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": kname,
                "algorithmic_bytes_per_launch": alg_bytes_launch, "kernel_ms_mean": ker_mean,
                "kernel_ms_min": min(ker_ms) if ker_ms else None, "launches_timed": len(ker_ms),
                "kernel_share_of_step": 2 * ker_mean / (ms_dev / args.steps)}

    # ---- e2e: the same workload streamed from pinned host memory through qb_submit_from ----
    e2e_pairs = int(os.environ.get("QB_BENCH_E2E_PAIRS", str(pairs)))
    h2d_gbs = ctx.measure_h2d(256 << 20, 3, 0)
    arena = Arena(capi, first, e2e_pairs, reads_per_batch)
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
           "note": "offsets/lengths (8 B/read) stay on the host for batches the period kernel takes: the host verified their shape"}
    arena.free()
    for b in db:
        b.free()
    ctx.close()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "bases_per_s": value * READ_LEN,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "wall_ms_per_step": wall_ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: paired-end 2x150 bp, adapters all.fa.gz (333 10-mers), "
                                   "read-through adapters in 10% of pairs",
                       "pairs_per_gpu": pairs, "reads_per_step": total_reads_step, "parallelism": f"shard{world}",
                       "l2": "inputs (3.1 GB per launch) larger than L2", "generator_seed": SEED},
            "roofline": roofline, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks, "checked": checked,
        }
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
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
