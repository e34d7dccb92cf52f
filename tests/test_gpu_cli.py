"""The `quack` host program end to end on a GPU: same command lines as the reference, SVG on stdout
byte-identical to what the unmodified reference printed (committed goldens, and the reference binary
itself when oracle/_ref travelled with the snapshot)."""
import gzip
import json
import os
import subprocess

import pytest

import qb_testutil as util
from oracle import pyoracle as po
from quack_b200 import build, synth
from quack_b200.build import quack_bin

pytestmark = pytest.mark.gpu

COMBOS = {
    "u_kat_t": ["-u", "kat_t.fq"],
    "u_kat_t_ad": ["-u", "kat_t.fq", "-a", "adapters_all.fa"],
    "u_kat_t_ad_name": ["-u", "kat_t.fq", "-a", "adapters_all.fa", "-n", "tiny"],
    "u_kat_k_ad": ["-u", "kat_k.fq", "-a", "adapters_all.fa"],
    "pe_kat_t_kat_k_ad_name": ["-1", "kat_t.fq", "-2", "kat_k.fq", "-a", "adapters_all.fa", "-n", "pair"],
    "u_rand": ["-u", "rand_small.fq.gz"],
    "u_rand_ad": ["-u", "rand_small.fq.gz", "-a", "adapters_all.fa"],
    "pe_rand_kat_k": ["-1", "rand_small.fq.gz", "-2", "kat_k.fq"],
    "pe_rand_rand_ad_name": ["--forward", "rand_small.fq.gz", "--reverse", "rand_small.fq.gz", "--adapters",
                             "adapters_all.fa", "--name", "rand x2"],
}


@pytest.fixture(scope="module", autouse=True)
def built():
    build()


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([quack_bin(), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)


def _golden(golden_dir, name):
    p = os.path.join(golden_dir, "svg", name + ".svg")
    return open(p, "rb").read() if os.path.exists(p) else gzip.open(p + ".gz", "rb").read()


@pytest.mark.parametrize("name", sorted(COMBOS))
@pytest.mark.parametrize("kernel", ["1", "2", "0"], ids=["simple", "fused", "auto"])
def test_cli_svg_matches_reference_golden(name, kernel, golden_dir):
    args = [os.path.join(golden_dir, a) if os.path.exists(os.path.join(golden_dir, a)) else a for a in COMBOS[name]]
    r = _run(args, {"QB_KERNEL": kernel, "QB_LEN_CAP": "320" if kernel == "2" else "65536"})
    assert r.returncode == 0, r.stderr
    assert r.stdout == _golden(golden_dir, name)


def test_cli_paired_200k_vs_reference_binary(tmp_path, golden_dir):
    """Config-2 shaped files (gzip members, adapters) through both programs; stage times to a JSON file."""
    f1, f2 = str(tmp_path / "s_1.fq.gz"), str(tmp_path / "s_2.fq.gz")
    synth.write_fastq(f1, 2, 1, 200_000, 150, 0.1, gz_level=1)
    synth.write_fastq(f2, 2, 2, 200_000, 150, 0.1, gz_level=1)
    js = str(tmp_path / "stats.json")
    args = ["-1", f1, "-2", f2, "-a", util.ADAPTER_FA, "-n", "cfg2 sample"]
    r = _run(args, {"QB_STATS_JSON": js, "QB_BATCH_MB": "8"})
    assert r.returncode == 0, r.stderr
    st = json.load(open(js))
    assert st["reads"] == 400_000 and st["bases"] == 400_000 * 150 and st["launches"] >= 8
    if po.have_ref():
        assert r.stdout == po.ref_svg(args)
    else:  # same arrays as the oracle, rendered by the same host code
        from quack_b200 import capi
        t = util.oracle_table()
        a, b = po.read_fastq(f1, t), po.read_fastq(f2, t)
        out = str(tmp_path / "o.svg")
        capi.render_svg(capi.Result(a.rows, a.max_length, a.n_reads), capi.Result(b.rows, b.max_length, b.n_reads),
                        True, "cfg2 sample", out)
        assert r.stdout == open(out, "rb").read()


def test_cli_bgzf_input_parallel_decode(tmp_path):
    """The same reads as BGZF files (decoded by the reader's thread pool) and as gzip files (gzread, like the
    reference): identical SVG; the reference binary reads the BGZF files too (multi-member gzip)."""
    g1, g2 = str(tmp_path / "g_1.fq.gz"), str(tmp_path / "g_2.fq.gz")
    b1, b2 = str(tmp_path / "b_1.fq.gz"), str(tmp_path / "b_2.fq.gz")
    for mate, g, b in ((1, g1, b1), (2, g2, b2)):
        synth.write_fastq(g, 2, mate, 60_000, 150, 0.1, gz_level=1)
        synth.write_fastq(b, 2, mate, 60_000, 150, 0.1, gz_level=1, bgzf=True)
    outs = {}
    for tag, f1, f2, thr in (("gz", g1, g2, "1"), ("bgzf1", b1, b2, "1"), ("bgzf4", b1, b2, "4")):
        js = str(tmp_path / (tag + ".json"))
        r = _run(["-1", f1, "-2", f2, "-a", util.ADAPTER_FA, "-n", "x"], {"QUACK_DECODE_THREADS": thr, "QB_STATS_JSON": js})
        assert r.returncode == 0, r.stderr
        outs[tag] = r.stdout
        assert json.load(open(js))["decode_threads"] == (4 if tag == "bgzf4" else 1)
    assert outs["gz"] == outs["bgzf1"] == outs["bgzf4"]
    if po.have_ref():
        assert outs["gz"] == po.ref_svg(["-1", b1, "-2", b2, "-a", util.ADAPTER_FA, "-n", "x"])


def test_cli_errors(tmp_path):
    r = _run(["-u", str(tmp_path / "missing.fq")])
    assert r.returncode == 2 and b"cannot open" in r.stderr and r.stdout == b""
    empty = tmp_path / "empty.fq"
    empty.write_bytes(b"")
    r = _run(["-u", str(empty)])
    assert r.returncode == 2 and b"no reads" in r.stderr
