"""Where a resident bench step spends its time outside the two kernels (development probe)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quack_b200 import capi, synth  # noqa: E402

pairs = 10_000_000
ctx = capi.Context(150, n_mates=2, adapter_keys=synth.adapter_keys(), batch_bytes=64 << 20, batch_reads=(64 << 20) // 150, ring_depth=3)
db = [ctx.generate(2, 1, 0, pairs, 150, 150, 0.1, 0), ctx.generate(2, 2, 0, pairs, 150, 150, 0.1, 0)]
for _ in range(3):
    db[0].run(0), db[1].run(1), ctx.finish(0), ctx.finish(1)


def timeit(f, n=20):
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    ctx.sync()
    return (time.perf_counter() - t0) / n * 1e3


print("2 launches + sync          %.3f ms" % timeit(lambda: (db[0].run(0), db[1].run(1), ctx.sync())))
print("2 launches + finish x2     %.3f ms" % timeit(lambda: (db[0].run(0), db[1].run(1), ctx.finish(0), ctx.finish(1))))
print("launch 0 + sync            %.3f ms" % timeit(lambda: (db[0].run(0), ctx.sync())))
print("sync only                  %.3f ms" % timeit(lambda: ctx.sync()))
print("finish x2 only (cached)    %.3f ms" % timeit(lambda: (ctx.finish(0), ctx.finish(1))))
avg, mn = db[0].time(0, warmup=3, iters=10, flush_l2=False)
print("kernel alone (events)      %.3f ms" % avg)
