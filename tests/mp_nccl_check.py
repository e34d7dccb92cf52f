"""Launched by tests/test_gpu_multi.py under torch.distributed.run: one rank per GPU, each counts its shard
through the C-ABI, qb_finish() sums over NCCL; rank 0 compares with the oracle over all reads."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qb_testutil as util  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from quack_b200 import capi, shard, synth  # noqa: E402

N = 200_001
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
keys = synth.adapter_keys()
ctx = capi.Context(304, n_mates=1, adapter_keys=keys, device_ids=[local])
idt = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
ctx.comm_init_rank(world, rank, idt.cpu().numpy().tobytes())
b, e = shard.shard_range(rank, world, N)
seq, qual, off, ln = capi.gen_reads(4, 1, b, e - b, 35, 300, 0.1)
ctx.accumulate_host(0, seq, qual, off, ln)
res = ctx.finish(0)
if rank == 0:
    s = capi.gen_reads(4, 1, 0, N, 35, 300, 0.1)
    want = po.accumulate_batch(*s, util.oracle_table())
    util.assert_same(res, want, f"{world}-rank NCCL reduce")
    print(f"MP_NCCL_OK world={world} reads={res.n_reads}")
else:
    assert res.n_reads == e - b   # non-root ranks keep their own partial result
ctx.close()
dist.destroy_process_group()
