#!/usr/bin/env python
"""Development probe (not the judged bench): kernel-only throughput of the simple and fused kernels on
generated 150-bp reads, with and without the adapter set, plus pinned H2D bandwidth."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quack_b200 import capi

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import qb_testutil as util

HBM = 6545.3
KN = {capi.KERNEL_SIMPLE: "simple", capi.KERNEL_FUSED: "fused", capi.KERNEL_PERIOD: "period",
      capi.KERNEL_AUTO: "auto", capi.KERNEL_FLAT: "flat"}
# QB_QUICK_KERNELS=3,2 restricts the sweep (default: wtile, fused, simple)
KSEL = [int(k) for k in os.environ.get("QB_QUICK_KERNELS", "0,2,1").split(",")]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    keys = np.concatenate([capi.adapter_record_keys(r) for r in util.adapter_records()])
    out = []
    shapes = ((150, 150, 150),) if os.environ.get("QB_QUICK_ONLY150") else ((150, 150, 150), (35, 300, 304))
    if os.environ.get("QB_QUICK_LENS"):  # fixed-length shapes, e.g. 50,76,100
        shapes = tuple((int(x), int(x), int(x)) for x in os.environ["QB_QUICK_LENS"].split(","))
    for lmin, lmax, cap in shapes:
        for ad in (None, keys):
            for kernel in KSEL:
                if kernel == capi.KERNEL_PERIOD and lmin != lmax:
                    continue
                nn = n if kernel != capi.KERNEL_SIMPLE else n // 8
                with capi.Context(cap, adapter_keys=ad, kernel=kernel) as ctx:
                    b = ctx.generate(2, 1, 0, nn, lmin, lmax, 0.1)
                    nr, nb = b.info
                    avg, mn = b.time(0, warmup=2, iters=5, flush_l2=False)
                    alg = 2 * nb + 8 * nr
                    rec = {"len": [lmin, lmax], "adapters": ad is not None,
                           "kernel": KN[kernel], "reads": nr,
                           "ms_avg": round(avg, 4), "ms_min": round(mn, 4), "Greads_s": round(nr / avg / 1e6, 3),
                           "Gbases_s": round(nb / avg / 1e6, 2), "GBps": round(alg / avg / 1e6, 1),
                           "frac_hbm": round(alg / avg / 1e6 / HBM, 4)}
                    for k in ("QB_LIB", "QB_WT_READS", "QB_PT_STAGES", "QB_PT_BYTES", "QB_PT_WARPS", "QB_PT_NATURAL"):
                        if os.environ.get(k):
                            rec[k] = os.path.basename(os.environ[k])
                    print(json.dumps(rec), flush=True)
                    out.append(rec)
                    b.free()
    if os.environ.get("QB_QUICK_ONLY150") or os.environ.get("QB_QUICK_LENS"):
        return
    with capi.Context(150) as ctx:
        print(json.dumps({"h2d_GBps": round(ctx.measure_h2d(256 << 20, 5), 2)}), flush=True)


if __name__ == "__main__":
    main()
