"""The inflate kernel alone (qb_bgzf_inflate_bench): GB/s of text over a generated BGZF file, vs zlib on one core."""
import json
import os
import subprocess
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quack_b200 import capi  # noqa: E402
from quack_b200.build import gen_bin  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    level = sys.argv[2] if len(sys.argv) > 2 else "1"
    path = "/dev/shm/qb_inflate_bench.fq.gz"
    subprocess.run([gen_bin(), path, "7", "1", "0", str(n), "150", "150", "0.1",
                    "bgzf", level], check=True)
    comp = open(path, "rb").read()
    os.unlink(path)
    with capi.Context(160, batch_bytes=1 << 20, ring_depth=2) as ctx:
        ms, text, blocks = ctx.bgzf_inflate_bench(comp, 5)
    t0 = time.perf_counter()
    sample = comp[: 64 << 20]
    d, out = zlib.decompressobj(31), 0
    rest = sample
    while rest:
        try:
            out += len(d.decompress(rest))
        except zlib.error:
            break
        rest = d.unused_data
        d = zlib.decompressobj(31)
    cpu_s = time.perf_counter() - t0
    print(json.dumps({"reads": n, "level": int(level), "comp_MB": round(len(comp) / 1e6, 1), "text_MB": round(text / 1e6, 1),
                      "blocks": blocks, "ms_per_launch": round(ms, 3), "text_GBps": round(text / ms / 1e6, 2),
                      "comp_GBps": round(len(comp) / ms / 1e6, 2), "Mreads_s": round(n / ms / 1e3, 1),
                      "zlib_one_core_text_GBps": round(out / cpu_s / 1e9, 3)}))


if __name__ == "__main__":
    main()
