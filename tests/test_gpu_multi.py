"""Multi-GPU paths on hardware (skipped on boxes with one GPU): one process driving several devices with a
single grouped ncclReduce, and one process per GPU under torch.distributed.run."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import qb_testutil as util
from oracle import pyoracle as po
from quack_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_in_process_multi_device_reduce():
    n = min(_ngpu(), 8)
    batch = util.random_batch(8, 120_000, 35, 150, plant=0.1)
    keys = util.oracle_table().keys()
    with capi.Context(150, adapter_keys=keys, n_devices=n, batch_bytes=1 << 20, batch_reads=9000) as ctx:
        ctx.accumulate_host(0, *batch)       # slots rotate over the devices round-robin
        got = ctx.finish(0)
        assert ctx.launch_count >= 2 * n
    util.assert_same(got, po.accumulate_batch(*batch, util.oracle_table()), f"{n} devices, one process")


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_one_process_per_gpu_nccl_reduce():
    n = min(_ngpu(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(ROOT, "tests", "mp_nccl_check.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and f"MP_NCCL_OK world={n}" in r.stdout, r.stdout[-3000:]
