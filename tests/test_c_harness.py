"""The C ABI driven from plain C (tests/c/abi_harness.c): compiled with gcc against libbdx.so.  Without a GPU it
covers config validation, error codes, the FASTQ scanner / packer, the loader and the "no CPU fallback" rule;
on the GPU box (marker gpu) it also classifies and the results are compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import bdx_b200 as bdx
from bdx_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "abi_harness")
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_harness.c"), "-o", exe, "-L", libdir, "-lbdx",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def _run(exe):
    return subprocess.run([exe], capture_output=True, text=True, timeout=300)


def test_c_harness_without_gpu(tmp_path):
    capi.load_library()
    r = _run(_build(tmp_path))
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    if "devices 0" not in r.stdout:
        pytest.skip("a CUDA device is visible: covered by the gpu test")


@pytest.mark.gpu
def test_c_harness_on_gpu(tmp_path):
    import orc
    r = _run(_build(tmp_path))
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    got = [tuple(int(x) for x in ln.split()[1:]) for ln in r.stdout.splitlines() if ln.startswith("result ")]
    bcs = ["ACGTACGT", "AAGGCCTT", "AACCTTGG", "AACCGGTT", "AGGATTCC", "AATTGGCC", "GCGCATAT", "GCATGCAT", "TAGCTAGC",
           "TAGGATCC", "GGATCCTT"]
    cfg = bdx.DemuxConfig(bc_seqs=bcs, bc_lengths_no_N=[8] * 11, ids=[str(i) for i in range(11)], trim_side=5)
    ref = orc.Oracle(cfg).classify_reads([b"GGACGTACGTCC", b"TTAAGGCCTTAA", b"ACG"])
    want = [tuple(int(ref[f][i]) for f in ("status", "bc1", "bc2", "keep_start", "keep_end")) for i in range(3)]
    assert got == want
