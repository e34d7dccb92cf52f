"""Launched by tests/test_gpu_configs.py under torch.distributed.run: config 3 -- the config-2
batch of each mate counted 20 times, the repetitions dealt round-robin to the ranks (one per GPU), one NCCL reduce in
qb_finish(); rank 0 checks that the result is exactly 20 x the oracle's counts of the config-2 batch."""
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
from quack_b200 import capi, synth  # noqa: E402

PAIRS = int(os.environ.get("QB_CFG3_PAIRS", "1000000"))
REPS = 20
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
keys = synth.adapter_keys()
ctx = capi.Context(150, n_mates=2, adapter_keys=keys, device_ids=[local])
idt = torch.zeros(capi.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
ctx.comm_init_rank(world, rank, idt.cpu().numpy().tobytes())
db = [ctx.generate(2, m + 1, 0, PAIRS, 150, 150, 0.1) for m in (0, 1)]
mine = [r for r in range(REPS) if r % world == rank]
for _ in mine:
    db[0].run(0)
    db[1].run(1)
res = [ctx.finish(0), ctx.finish(1)]
if rank == 0:
    table = util.oracle_table()
    for m in (0, 1):
        w = po.accumulate_batch(*capi.gen_reads(2, m + 1, 0, PAIRS, 150, 150, 0.1), table)
        util.assert_same(res[m], capi.Result(w.rows * np.uint64(REPS), w.max_length, w.n_reads * REPS),
                         f"config 3, mate {m + 1}, {world} ranks")
    print(f"MP_CONFIG3_OK world={world} pairs={PAIRS * REPS}")
else:
    assert res[0].n_reads == PAIRS * len(mine)
ctx.close()
dist.destroy_process_group()
