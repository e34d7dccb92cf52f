"""In-tree build of the CUDA library (and the host program) with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIBDIR = os.path.join(HERE, "lib")
BINDIR = os.path.join(HERE, "bin")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]


def lib_path() -> str:
    # QB_LIB: an alternative build of the library (kernel experiments, e.g. a different warp count)
    return os.environ.get("QB_LIB") or os.path.join(LIBDIR, "libquack_b200.so")


def quack_bin() -> str:
    return os.path.join(BINDIR, "quack")


def gen_bin() -> str:
    return os.path.join(BINDIR, "qb_gen_fastq")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _nvcc() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libquack_b200.so (kernels + C-ABI) and, when present, the host program."""
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in ("qb_kernels.cu", "qb_period.cu", "qb_flat.cu", "qb_text.cu", "qb_inflate.cu", "qb_extras.cu", "qb_transform.cu", "qb_api.cu", "qb_host.cpp", "qb_gen.cpp")]
    host_lib_srcs = [os.path.join(HOST, f) for f in ("fq_reader.c", "render.c")
                     if os.path.exists(os.path.join(HOST, f))]
    deps = srcs + host_lib_srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps += [os.path.join(HOST, f) for f in (os.listdir(HOST) if os.path.isdir(HOST) else []) if f.endswith(".h")]
    deps.append(os.path.join(HERE, "..", "include", "quack_b200.h"))
    out = lib_path()
    if force or not _newer(out, deps):
        hdrs = [d for d in deps if d.endswith((".h", ".cuh"))]
        extra = os.environ.get("QB_NVCC_EXTRA", "").split()
        tag = ("." + "".join(c if c.isalnum() else "_" for c in " ".join(extra))) if extra else ""
        jobs = []
        objs = []
        for c in host_lib_srcs:  # plain C translation units, compiled as C
            o = os.path.join(LIBDIR, os.path.basename(c) + ".o")
            objs.append(o)
            if force or not _newer(o, [c] + hdrs):
                jobs.append(["gcc", "-O3", "-std=c11", "-D_DEFAULT_SOURCE", "-fPIC", "-Wall", "-c", c, "-o", o,
                             "-I" + os.path.join(HERE, "..", "include")])
        for c in srcs:  # one object per translation unit: only what changed is recompiled, all of them in parallel
            o = os.path.join(LIBDIR, os.path.basename(c) + tag + ".o")
            objs.append(o)
            if force or not _newer(o, [c] + hdrs):
                cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", c, "-o", o]
                if verbose:
                    cmd.insert(1, "-Xptxas=-v")
                jobs.append(cmd)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
            for r in ex.map(lambda cmd: subprocess.run(cmd, check=True), jobs):
                pass
        subprocess.run([_nvcc(), *NVCC_FLAGS, "-shared", "-o", out, *objs, "-ldl", "-lz", "-lpthread"], check=True)
    main_c = os.path.join(HOST, "quack_main.c")
    if os.path.exists(main_c):
        os.makedirs(BINDIR, exist_ok=True)
        qdeps = [main_c, out] + deps
        if force or not _newer(quack_bin(), qdeps):
            subprocess.run(["gcc", "-O3", "-std=c11", "-D_DEFAULT_SOURCE", "-Wall", "-o", quack_bin(), main_c,
                            "-I" + os.path.join(HERE, "..", "include"), "-L" + LIBDIR, "-lquack_b200",
                            "-Wl,-rpath,$ORIGIN/../lib", "-lz", "-lm", "-lpthread"], check=True)
    # the standalone writer of the synthetic benchmark inputs: the generator's translation unit + a main, no library
    gen_cpp = os.path.join(HERE, "..", "tools", "gen_fastq.cpp")
    gen_dep = [gen_cpp, os.path.join(CSRC, "qb_gen.cpp"), os.path.join(CSRC, "qb_host.h")]
    if os.path.exists(gen_cpp):
        os.makedirs(BINDIR, exist_ok=True)
        if force or not _newer(gen_bin(), gen_dep):
            subprocess.run(["g++", "-O3", "-std=c++17", "-Wall", "-o", gen_bin(), gen_cpp, os.path.join(CSRC, "qb_gen.cpp"),
                            "-lz", "-lpthread"], check=True)
    return out
