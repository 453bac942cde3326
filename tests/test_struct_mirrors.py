"""The hand-kept mirrors of include/bdx.h -- the Julia structs of julia/BioDemuXB200.jl (which cannot be executed
in this image) and the ctypes Structures of capi.py -- against the header itself: field order, names and widths.
A C struct and its mirror are both parsed from their source text; drift in either fails the test."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "bdx.h")).read()
JULIA = open(os.path.join(ROOT, "biodemux.jl_b200", "julia", "BioDemuXB200.jl")).read()

C_TYPES = {"int32_t": ("Int32", 4), "uint32_t": ("UInt32", 4), "int64_t": ("Int64", 8), "uint64_t": ("UInt64", 8),
           "double": ("Float64", 8), "float": ("Float32", 4), "int": ("Int32", 4)}
# C struct -> Julia mirror
PAIRS = {"bdx_range": "BdxRange", "bdx_barcode_set": "BdxBarcodeSet", "bdx_params": "BdxParams",
         "bdx_result": "BdxResult", "bdx_stats_overflow": "BdxStatsOverflow", "bdx_stats_layout": "BdxStatsLayout",
         "bdx_demux_bucket": "BdxDemuxBucket", "bdx_demux_out": "BdxDemuxOut"}


def c_struct(name):
    """[(field, julia type)] of `typedef struct name { ... } name;` in bdx.h"""
    m = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", HEADER, re.S)
    assert m, name
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        mm = re.match(r"(const )?(\w+) (\*?)(.*)$", decl)
        ctype, ptr, names = mm.group(2), mm.group(3), mm.group(4)
        for nm in names.split(","):
            nm = nm.strip()
            is_ptr = bool(ptr) or nm.startswith("*")
            nm = nm.lstrip("*")
            arr = re.match(r"(\w+)\[(\d+)\]", nm)
            if is_ptr:
                jt = "Ptr"
            elif ctype in C_TYPES:
                jt = C_TYPES[ctype][0]
            else:
                jt = PAIRS.get(ctype, ctype)
            if arr:
                out.append((arr.group(1), f"NTuple{{{arr.group(2)},{jt}}}"))
            else:
                out.append((nm, jt))
    return out


def julia_struct(name):
    m = re.search(r"^struct " + name + r"\b[^\n]*\n(.*?)^end", JULIA, re.S | re.M)
    assert m, name
    out = []
    for line in m.group(1).splitlines():
        line = line.split("#")[0].strip()
        if "::" in line:
            f, t = line.split("::")
            out.append((f.strip(), "Ptr" if t.strip().startswith("Ptr") else t.strip()))
    return out


@pytest.mark.parametrize("cname", sorted(PAIRS))
def test_julia_mirror_matches_header(cname):
    assert julia_struct(PAIRS[cname]) == c_struct(cname)


def test_ctypes_mirrors_match_header():
    from bdx_b200 import capi
    ct = {"bdx_range": capi.Range, "bdx_barcode_set": capi.BarcodeSet, "bdx_params": capi.Params,
          "bdx_stats_layout": capi.StatsLayout, "bdx_synth_spec": capi.SynthSpec, "bdx_demux_out": capi.DemuxOut}
    widths = {"Int32": 4, "UInt32": 4, "Int64": 8, "UInt64": 8, "Float64": 8, "Ptr": 8}
    for cname, cls in ct.items():
        want = c_struct(cname)
        got = [(n, t) for n, t in cls._fields_]
        assert [n for n, _ in got] == [n for n, _ in want], cname
        for (n, t), (_, jt) in zip(got, want):
            if jt in widths:
                assert C.sizeof(t) == widths[jt], (cname, n)


def test_abi_version_and_in_flight_constants():
    assert re.search(r"#define BDX_ABI_VERSION (\d+)", HEADER).group(1) == re.search(r"BDX_ABI_VERSION\s*=\s*(?:UInt32\()?(\d+)", JULIA).group(1)
