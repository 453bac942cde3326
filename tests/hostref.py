"""Test helpers: run the host pipeline with the ORACLE as the classifier (to pin
the oracle against the reference's golden files) and compare output trees."""
from __future__ import annotations

import gzip
import os
import re

import numpy as np

import bdx_b200 as bdx
import orc


def oracle_classifier(cfg, want_stats=None):
    o = orc.Oracle(cfg, want_stats=want_stats)

    def classify(seqs):
        res = o.classify_reads(seqs)
        out = np.zeros(len(seqs), dtype=bdx.RESULT_DTYPE)
        for f in ("status", "bc1", "bc2", "keep_start", "keep_end"):
            out[f] = res[f]
        return out

    return classify


def strip_ext(p):
    return re.sub(r"\.fastq(\.gz)?$", "", os.path.basename(p))


def check_output_files(output_dir, ideal_dir):
    """test/common.jl:4-21 -- every golden file must exist and be byte-identical."""
    n = 0
    for name in sorted(os.listdir(ideal_dir)):
        out_p, ideal_p = os.path.join(output_dir, name), os.path.join(ideal_dir, name)
        assert os.path.isfile(out_p), f"missing output {name}"
        if name.lower().endswith(".gz"):
            a, b = gzip.open(out_p).read(), gzip.open(ideal_p).read()
        else:
            a, b = open(out_p, "rb").read(), open(ideal_p, "rb").read()
        assert a == b, f"content differs: {name}"
        n += 1
    return n


def demo_cases(refdata):
    """The three golden integration cases of test/integration/single_barcode.jl:2-45
    as (name, [(fastq1, fastq2, prefix1, prefix2)], barcode_file, kwargs, ideal_dir)."""
    fq = os.path.join(refdata, "FASTQ_files")
    rf = os.path.join(refdata, "reference_files")
    res = os.path.join(refdata, "results")
    r1 = sorted(os.listdir(os.path.join(fq, "demo1_R1")))
    r2 = sorted(os.listdir(os.path.join(fq, "demo1_R2")))
    g1 = sorted(os.listdir(os.path.join(fq, "demo2_R1")))
    g2 = sorted(os.listdir(os.path.join(fq, "demo2_R2")))
    cases = []
    cases.append(("demo1_R1", [(os.path.join(fq, "demo1_R1", f), None, strip_ext(f), "") for f in r1],
                  os.path.join(rf, "demo1.tsv"), {}, os.path.join(res, "demo1_R1")))
    cases.append(("demo1_R2", [(os.path.join(fq, "demo1_R1", a), os.path.join(fq, "demo1_R2", b),
                                strip_ext(a), strip_ext(b)) for a, b in zip(r1, r2)],
                  os.path.join(rf, "demo1.tsv"), {}, os.path.join(res, "demo1_R2")))
    kw = dict(max_error_rate=0.25, min_delta=0.15, mismatch=1, indel=2, classify_both=True,
              bc_complement=True, bc_rev=True, gzip_output=False)
    cases.append(("demo2", [(os.path.join(fq, "demo2_R1", a), os.path.join(fq, "demo2_R2", b),
                             "test_prefix1." + strip_ext(a), "test_prefix2." + strip_ext(b))
                            for a, b in zip(g1, g2)],
                  os.path.join(rf, "demo2.csv"), kw, os.path.join(res, "demo2")))
    return cases


def build_cfg(barcode_file, fastqs, **kw):
    d = dict(barcode_file2=None, gzip_output=None, bc_complement=False, bc_rev=False, classify_both=False,
             max_error_rate=0.2, min_delta=0.0, match=0, mismatch=1, indel=1, nindel=None,
             ref_search_range="1:end", barcode_start_range="1:end", barcode_end_range="1:end",
             ref_search_range2="1:end", barcode_start_range2="1:end", barcode_end_range2="1:end",
             trim_side=None, trim_side2=None, summary=False, summary_format="html",
             matching_algorithm="semiglobal")
    d.update(kw)
    return bdx.build_config(barcode_file, d["barcode_file2"], fastqs, d["gzip_output"], d["bc_complement"],
                            d["bc_rev"], d["classify_both"], d["max_error_rate"], d["min_delta"], d["match"],
                            d["mismatch"], d["indel"], d["nindel"], d["ref_search_range"],
                            d["barcode_start_range"], d["barcode_end_range"], d["ref_search_range2"],
                            d["barcode_start_range2"], d["barcode_end_range2"], d["trim_side"],
                            d["trim_side2"], d["summary"], d["summary_format"], d["matching_algorithm"])


def run_case(case, out_dir, make_classifier):
    name, files, bc_file, kw, ideal = case
    for f1, f2, p1, p2 in files:
        fastqs = [f1] if f2 is None else [f1, f2]
        kw2 = dict(kw)
        if f2 is None:
            kw2["classify_both"] = False
        cfg = build_cfg(bc_file, fastqs, **kw2)
        bdx.run_pipeline(cfg, make_classifier(cfg), f1, f2, out_dir, p1, p2)
    return check_output_files(out_dir, ideal)
